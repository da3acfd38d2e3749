// ktn_handle.h -- the handle behind the C ABI (shared by ktn_api.cu and ktn_comm.cu).
#ifndef KTN_HANDLE_H
#define KTN_HANDLE_H
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/ktn.h"
#include "ktn_compile.h"
#include "ktn_kernels.cuh"

#define KTN_PX_MAX_RANKS 16
#define KTN_PX_SLOTS 3
#define KTN_PX_CTRL 4096u   // control page at the head of a receive arena: ack[rank] words
#define KTN_LANE_LIMIT 1536u   // max per-lane shared-memory bytes of a regular (shared-memory staged) shape

struct DevBuf {
    void* p = nullptr; size_t bytes = 0;
    cudaError_t alloc(size_t n) { release(); bytes = n; if (n == 0) return cudaSuccess; return cudaMalloc(&p, n); }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct ktn_handle {
    ktn_options opt;
    int device = 0, num_sms = 0, max_smem = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    // per-round kernel timing: ring of (start, after K1, after K3, after K2) events, drained by ktn_timings_get
    static const int RING = 128;
    cudaEvent_t ring[RING][4];
    int mid_every = 8;             // the timing events of a round are recorded in one round of mid_every (enqueue_round)
    int ring_head = 0, ring_tail = 0;      // [tail, head) not yet drained
    double eval_ms_sum = 0, compact_ms_sum = 0, cut_ms_sum = 0; int64_t rounds_timed = 0;
    double exchange_ms_sum = 0; int64_t exchanges_timed = 0;
    KtnProblem prob;
    bool loading = false, loaded = false, round_pending = false, have_round = false;
    bool forced_last = false;              // the last round was unconditional (ktn_gencut_rows): rows outside nlconstr_ixs may carry selection flags
    DevBuf chunks, shapes, prog, blob, chunk_rows, chunk_lb, chunk_ub, jac_ptr, jac_col, row_lb, row_ub, row_slot, row_nl, ladder, rec, worklist, park, cut_off, errpos, blk_off, chunk_jp, dump;
    DevBuf topk_key, topk_state, topk_eqcnt;      // top-k selection (allocated when ktn_options.topk > 0)
    DevBuf x, force, g_row, b_row, sel, stage_val, big_scratch, ticket, blk_cnt, counts, table;
    // the round's cuts: one blob written by K2 (ktn_pack_layout).  Sharded handles rotate three blobs, so that the exchange of
    // round i reads its blob while rounds i+1, i+2 write theirs; blob_ev = the exchange that last read a blob has finished
    DevBuf out_blob[3]; int out_cur = 0; size_t out_cap = 0;
    cudaEvent_t blob_ev[3] = {nullptr, nullptr, nullptr}; bool blob_busy[3] = {false, false, false};
    int64_t lay_cuts = 0, lay_nnz = 0;      // untruncated totals of the last synced round: the arguments of the blob's layout
    double* h_x = nullptr;                 // pinned
    unsigned long long* h_counts = nullptr;  // pinned [8]
    int64_t n_cuts = 0, nnz_cuts = 0, err_row = -1;
    unsigned char* h_view[2] = {nullptr, nullptr}; size_t h_view_cap[2] = {0, 0}; int view_cur = 0;   // pinned buffers behind ktn_fetch_cuts_view
    // KTN_FLAG_DIRECT_VIEW: the cut blob itself lives in mapped pinned HOST memory (two alternate), K2 / K3 store the batch over PCIe
    unsigned char* h_direct[2] = {nullptr, nullptr}; unsigned char* d_direct[2] = {nullptr, nullptr}; int direct_cur = 0; bool direct = false, round_lean = false;
    uint32_t warp_bytes = 0, blob_cap = 0, table_bytes = 0, table_prog_off = 0, epoch = 0, blk_stride = 0;
    ktn_timings tm;
    std::string err;
    // sharding (ktn_comm.cu): NCCL communicator on its own stream; three exchange slots rotate so that the payload of round i
    // travels while later rounds compute and the host never waits for sizes that are not there yet
    void* comm = nullptr; int nranks = 1, rank = 0;
    cudaStream_t comm_stream = nullptr;
    struct Exchange {
        DevBuf gathered, all_counts, stage;             // stage: send buffer of the common slot size when this rank's own blob is smaller
        unsigned long long* h_all_counts = nullptr;     // pinned [8 * nranks]: the blob headers of all ranks
        std::vector<int64_t> g_cuts, g_nnz, g_off;      // per rank, once the headers are on the host (ranks behind the first non-finite cut: 0)
        int64_t g_err_row = -1;                         // global index of the first non-finite row of the gathered batch (-1: none)
        cudaEvent_t t0 = nullptr, t1 = nullptr;         // around the exchange's transfer on the exchange stream (exchange_ms)
        std::vector<int64_t> g_lay_cuts, g_lay_nnz, g_bytes;   // layout arguments and size of every rank's blob
        int src_idx = 0;                                // which of the handle's cut blobs this exchange ships
        cudaEvent_t packed = nullptr, sizes = nullptr;    // the round's K2 has finished; the headers of all ranks are on the host
        int state = 0;                                  // 0 idle, 1 sizes in flight (payload not launched), 2 payload in flight / complete
        int64_t gathered_bytes = 0;
    } xch[3];                                           // the payload of exchange k is launched when exchange k+2 is enqueued
    int xch_cur = 0;                                    // slot of the last ktn_allgather_cuts_async
    // peer-push exchange (default when every rank can map every other rank's receive arena): the pack kernel's blob is stored
    // straight into the peers' HBM over NVLink by ktn_push_kernel; NCCL only bootstraps it (and stays the fallback transport)
    struct PeerExchange {
        bool tried = false, on = false;
        size_t slot_cap = 0;                            // bytes per (exchange slot, source rank)
        DevBuf arena;                                   // control page + [KTN_PX_SLOTS][nranks][slot_cap], exported over CUDA IPC
        DevBuf boot, hdr, ctr;                          // bootstrap buffer, gathered headers (+ error word), block counter + error word
        unsigned char* peer[KTN_PX_MAX_RANKS] = {};     // every rank's arena as mapped here (own arena at [rank])
        unsigned long long* h_boot = nullptr;           // pinned
        unsigned long long seq = 0;                     // exchanges enqueued so far (1-based sequence number of the last one)
        int blocks = 16;                                // grid of the push kernel: KTN_PUSH_BLOCKS, else sized per round (ktn_comm_plan_blocks)
        bool blocks_fixed = false;
        bool reserve = true;                            // K1 leaves that many SMs free (KTN_PUSH_RESERVE=0 turns it off)
    } px;
    int64_t row_offset = 0;
    // single-process sharded operation (ktn_options.ngpus > 1): this handle is only the FRONT of one handle per device
    std::vector<ktn_handle*> shards;
    std::vector<int64_t> shard_begin;               // rows [shard_begin[s], shard_begin[s + 1]) live on shards[s]
    std::vector<int64_t> sh_cuts, sh_nnz;           // last round: cuts / nnz every shard contributes to the combined batch
    int64_t g_num_var = 0, g_num_constr = 0;
    double* g_hx = nullptr;                         // pinned staging of x* (all devices upload from it)
    // eager download (KTN_FLAG_EAGER_VIEW): two pinned buffers whose sections are laid out for the WORST case (every nonlinear row cut),
    // so a shard's cuts can be copied to their final place as soon as that shard has finished, while later shards still compute
    unsigned char* g_eager[2] = {nullptr, nullptr}; int eager_cur = 0; bool eager_valid = false;
    int64_t eager_cap_cuts = 0, eager_cap_nnz = 0;
    // every shard on ONE device (KatanaGPUSeparator(pipeline = S)): the shards share a stream, x* is uploaded once, and a small kernel per
    // shard stores that shard's cuts into the combined batch in mapped pinned memory while the next shards compute (group_round)
    bool push_view = false; int push_blocks = 12; float push_pace = 0.f;      // KTN_HOSTPUSH_GBS: bytes per nanosecond the push may store, 0 = unpaced (measured best: the SM store path, 43-46 GB/s, is the bottleneck either way)
    unsigned char* g_eager_d[2] = {nullptr, nullptr};
    unsigned long long* g_hdr = nullptr; unsigned long long* g_hdr_d = nullptr;
    cudaStream_t g_copy_stream = nullptr; std::vector<cudaEvent_t> g_done;
    int reserve_sms = 0;                            // (a shard) K1 leaves that many SMs to the host push of the shard before it
    DevBuf rp_shift;                                // shard: row_ptr of the last round shifted to the combined batch's entry offsets
};

static inline int fail(ktn_handle* h, int code, const char* fmt, ...) {
    char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (h) h->err = buf;
    return code;
}
#define CK(h, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail(h, KTN_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); } while (0)


#endif
