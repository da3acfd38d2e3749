// ktn_kernels.cuh -- launch interface of the sm_100a separation kernels (ktn_kernels.cu).
#ifndef KTN_KERNELS_CUH
#define KTN_KERNELS_CUH
#include <cuda_runtime.h>
#include <stdint.h>
#include "ktn_program.h"

// device state of the top-k radix select (ktn_kernels.cu, "Top-k selection")
struct KtnTopkState { unsigned long long prefix, mask, remaining, eq_total; unsigned int hist[256]; unsigned int done, all; };

enum { KTN_MODE_SEPARATE = 0, KTN_MODE_FORCE = 1, KTN_MODE_EVAL = 2 };   // EVAL: g only (set by ktn_launch_eval)

struct KtnRoundParams {
    // compiled problem (device)
    const KtnChunkDesc* chunks;
    const KtnShapeDesc* shapes;
    const KtnIns* prog;
    const uint8_t* blob;
    const int32_t* chunk_rows;
    const double* chunk_lb;
    const double* chunk_ub;
    const int64_t* jac_ptr;    // row -> first entry of the row in the static Jacobian CSR (= staging layout)
    const int32_t* jac_col;
    const double* row_lb;
    const double* row_ub;
    const int32_t* row_slot;   // row -> chunk * 32 + lane
    const uint32_t* chunk_jp;  // chunk * 32 + lane -> first Jacobian entry of the row (jac_ptr)
    double* dump; uint64_t dump_nnz;
    // round inputs
    const double* x;
    const uint8_t* force;      // KTN_MODE_FORCE: per-row mask
    double f_tol, rng;
    int64_t topk;              // > 0: keep only the k most violated rows (build extension)
    unsigned long long* topk_key; KtnTopkState* topk_state; unsigned int* topk_eqcnt;
    int32_t mode, do_round;
    int32_t clear_unselected;  // separation round after an unconditional one: also visit the chunks of rows outside nlconstr_ixs (their flags are reset)
    int64_t num_var, num_rows, row_offset;
    uint32_t chunk_begin, chunk_end;   // chunk range this launch covers
    uint32_t warp_bytes;               // shared-memory bytes per warp (regular kernel)
    uint32_t blob_cap;                 // bytes reserved for the blob inside a warp's region
    // round outputs
    double* g_row;             // g_i(x*) for every evaluated row (sep.g)
    double* b_row;             // cut constant of selected rows
    uint32_t* sel;             // 0 = not selected, else nnz | KTN_SEL_ERRBIT | KTN_SEL_DEFER
    double4* rec;              // family rows K1 selected: {g, aux, lb, ub}, one 32-byte sector per row, read by K2's cut
    double* stage_val;         // cut coefficients in the static CSR layout
    double* big_scratch;
    unsigned int* ticket;      // KTN_TICKETS dynamic work tickets (interpreter K1: one; family K1: one per class); re-armed by K2
    uint32_t ticket_idx;       // first ticket this launch draws from
    uint32_t fam_begin[KTN_FAM__COUNT + 1];                 // chunk range of every family,
    uint32_t cls_begin[KTN_FAM__COUNT][KTN_FAM_NCLS + 1];   // of every class of a family,
    uint64_t cls_blob_off[KTN_FAM__COUNT][KTN_FAM_NCLS];    // blob offset of the class's first chunk (blobs of a class >= 1 are KTN_FAM_BLOB_BYTES apart)
    // block-shared table of the regular kernel: shape descriptors, then the programs of the regular shapes
    const unsigned char* table; uint32_t table_bytes, table_prog_off;
    uint32_t epoch;                    // round counter (>= 1): look-back flags and error slots are epoch-stamped
    // compaction: per block of KTN_CROWS rows, cuts << KTN_BLK_SHIFT | nnz of the selected rows, added by K1 and summed
    // by K2.  Two copies indexed by epoch parity; K2 zeroes the copy the next round adds into.
    unsigned long long* blk_cnt; uint32_t blk_stride;
    unsigned long long* blk_off;       // large problems: exclusive scan of blk_cnt (ktn_blkscan_kernel), 2 words per block + totals
    double4* park;             // per cut of the compacted list: the record of a family row, parked by K2 where K3 reads it coalesced
    unsigned long long* cut_off;       // per cut of the compacted list (+ 1): first entry of the cut in the round's CSR (device copy of row_ptr: the cut
                                       // blob may live in HOST memory, KTN_FLAG_DIRECT_VIEW, and is then never read back by a kernel)
    int lean_out;                      // 1: g | viol | b sections of the cut blob are not written (KTN_FLAG_DIRECT_VIEW with KTN_FLAG_LEAN_VIEW)
    unsigned long long* worklist;      // per cut of the compacted list: 0 = cut built by K1, else (nnz | family << 8 | 1 << 16) << 32 | chunk slot (cut kernel)
    unsigned long long* errpos;        // per compaction block: (cut index, entry offset) of the block's first non-finite row (K2)
    // [0] n_cuts [1] nnz (both truncated at the first non-finite row) [2],[3] first-error row + 1 of even / odd epochs
    // (~0 = none) [4] n_cuts_total [5] nnz_total [6] first-error row + 1 of the last round [7] K2 blocks finished (the last one
    // to finish writes the totals and the blob header and re-arms the per-round state)
    unsigned long long* counts;
    // K2 writes the round's cuts as ONE blob (ktn_pack_layout of the round's total counts): 64-byte header
    // {n_cuts, nnz (both truncated at the first non-finite row), first-error row + 1 (~0 = none), blob bytes, row offset,
    //  n_cuts_total, nnz_total (the layout's arguments), reserved}, then row_id (GLOBAL ids) | row_ptr | lo | hi | g | viol | b |
    // col | val.  The host downloads it, and the sharded exchange ships it, as it lies.
    unsigned char* out_blob;
};

// Chunk ranges of one problem: regular chunks sorted by (family, class), then the BIG chunks.
struct KtnLaunchPlan {
    uint32_t fam_begin[KTN_FAM__COUNT + 1]; uint32_t cls_begin[KTN_FAM__COUNT][KTN_FAM_NCLS + 1];
    uint64_t cls_blob_off[KTN_FAM__COUNT][KTN_FAM_NCLS];
    uint32_t n_regular, n_total;
};
enum { KTN_TICKET_GENERIC = 0, KTN_TICKET_LSE = 8, KTN_TICKET_QUAD = 32, KTN_TICKET_SOC = 56, KTN_TICKETS = 80 };

// Launches the kernels of one round on `stream`; returns the number of kernels launched.
int ktn_launch_round(const KtnRoundParams& p, const KtnLaunchPlan& plan, int num_sms,
                     int max_smem_optin, uint32_t epoch, cudaStream_t stream, cudaEvent_t after_eval, cudaEvent_t after_compact, cudaError_t* err);
void ktn_plan_occupancy(uint32_t table_bytes, uint32_t warp_bytes, int max_smem_optin, int* warps_per_block, int* blocks_per_sm);
// forward evaluation only (ktn_eval_g): writes g_row for every row
int ktn_launch_eval(const KtnRoundParams& p, const KtnLaunchPlan& plan, int num_sms,
                    int max_smem_optin, cudaStream_t stream, cudaError_t* err);
cudaError_t ktn_kernels_configure(int max_smem_optin);

// the cut blob K2 writes (byte offsets of the sections)
struct KtnPackLayout { unsigned long long row_id, row_ptr, lo, hi, g, viol, b, col, val, total; };
#if defined(__CUDACC__)
__host__ __device__
#endif
inline KtnPackLayout ktn_pack_layout(unsigned long long n, unsigned long long nz) {
    KtnPackLayout L; unsigned long long o = 64;
    L.row_id = o; o = (o + 8 * n + 15) & ~15ull;
    L.row_ptr = o; o = (o + 8 * (n + 1) + 15) & ~15ull;
    L.lo = o; o = (o + 8 * n + 15) & ~15ull;
    L.hi = o; o = (o + 8 * n + 15) & ~15ull;
    L.g = o; o = (o + 8 * n + 15) & ~15ull;
    L.viol = o; o = (o + 8 * n + 15) & ~15ull;
    L.b = o; o = (o + 8 * n + 15) & ~15ull;
    L.col = o; o = (o + 4 * nz + 15) & ~15ull;
    L.val = o; o = (o + 8 * nz + 15) & ~15ull;
    L.total = o;
    return L;
}
// Pipelined single-device handles (KTN_FLAG_EAGER_VIEW, every shard on one device): a shard's cuts are stored by a small kernel
// straight into the combined batch in mapped pinned HOST memory, at the place the cuts of the shards before it end (ktn_api.cu).
#define KTN_HP_MAX_PREV 16
struct KtnHostPushParams {
    const unsigned char* src;                          // the shard's cut blob (device)
    const unsigned long long* counts;                  // the shard's round counters (device): [0] cuts, [1] entries, [4] / [5] layout totals, [6] first non-finite row + 1
    const unsigned long long* prev[KTN_HP_MAX_PREV];   // the counters of the shards before it (same device; their rounds precede this kernel in stream order)
    int nprev, lean;
    unsigned char* dst;                                // the combined batch (device alias of the pinned buffer)
    KtnPackLayout EL;                                  // its sections: laid out for the worst case, filled compactly
    float pace;                                        // bytes per nanosecond the whole grid may store (0: unpaced), see ktn_hostpush_kernel
    unsigned long long* hdr;                           // host-mapped: {cuts, entries, first non-finite row + 1 (or ~0), 0} of this shard as it enters the batch
};

#endif
