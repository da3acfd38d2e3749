// ktn_comm.cu -- sharded operation: every GPU separates its contiguous slice of the constraint rows, then the
// compacted cuts of all ranks are combined on every GPU over NCCL / NVLink.
// Rank-major concatenation is ascending row order, the reference's emission order (src/model.jl:272).
// NCCL is loaded lazily with dlopen so libktn.so has no link-time dependency on it.
//
// Two transports, chosen collectively at the first exchange:
//  * peer push (default): every rank exports a receive arena over CUDA IPC; ktn_push_kernel streams the rank's packed cut blob
//    straight into the arenas of ALL ranks with 16-byte stores over NVLink / NVSwitch, then publishes a header carrying the
//    exchange's sequence number behind a system-scope fence.  No sizes travel ahead of the data (slots are sized for the
//    worst case), no collective kernel has to be co-scheduled on all GPUs, and nothing waits on the host: the push of round i
//    runs on a few SMs beside the kernels of round i+1.  Flow control is an ack word per rank (the highest sequence number a
//    rank has begun to push, i.e. it no longer reads older slots): a slot is overwritten only when its owner has moved on.
//    NCCL is used for the bootstrap only (slot size, IPC handles, agreement) and as a barrier at tear-down.
//  * NCCL (KTN_EXCHANGE=nccl, or when a peer arena cannot be mapped): sizes all-gather, then one ncclAllGather of the blobs
//    in slots of the largest blob, launched two calls later.
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include "ktn_handle.h"

KtnRoundParams ktn_make_params(ktn_handle* h, const double* d_x, int mode, int do_round);

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint8 = 1, ncclUint64 = 5 };
struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
} N;

bool load_nccl(std::string* why) {
    if (N.ok) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { N.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (N.lib) break; }
    if (!N.lib) { *why = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
#define SYM(f) *(void**)(&N.f) = dlsym(N.lib, "nccl" #f); if (!N.f) { *why = "missing symbol nccl" #f; return false; }
    SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(AllGather) SYM(Broadcast) SYM(GroupStart) SYM(GroupEnd) SYM(GetErrorString)
#undef SYM
    N.ok = true;
    return true;
}
}  // namespace

// ---- peer-push kernels -------------------------------------------------------------------------------------------------------
struct KtnPushParams {
    unsigned char* dst[KTN_PX_MAX_RANKS];             // this exchange's (slot, source = me) region in every rank's arena
    unsigned long long* ack_dst[KTN_PX_MAX_RANKS];    // my ack word in every rank's control page
    const unsigned long long* ack_local;              // the ack words of all ranks in MY control page
    const unsigned char* src;                         // packed blob (ktn_pack_kernel)
    unsigned long long seq;
    unsigned int* ctr;                                // [0] finished blocks, [1] error word
    int nranks, rank;
};

__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long* p) {
    unsigned long long v; asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long now_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define KTN_PX_TIMEOUT_NS 20000000000ull     // a peer that does not show up within 20 s is reported, not waited for forever

__global__ void __launch_bounds__(512) ktn_push_kernel(const KtnPushParams q) {
    __shared__ int s_flag;
    if (threadIdx.x == 0) s_flag = 0;
    __syncthreads();
    // flow control: announce that this rank has moved on to exchange `seq` (it no longer reads the slots of older exchanges),
    // then wait until every destination has moved past the exchange whose slot is about to be overwritten
    if (threadIdx.x < (unsigned)q.nranks) {
        const int r = threadIdx.x;
        if (blockIdx.x == 0) st_sys(q.ack_dst[r], q.seq);
        if (q.seq > KTN_PX_SLOTS) {
            const unsigned long long need = q.seq - KTN_PX_SLOTS + 1, t0 = now_ns();
            while (ld_sys(q.ack_local + r) < need) {
                if (now_ns() - t0 > KTN_PX_TIMEOUT_NS) { s_flag = 1; break; }
                __nanosleep(200);
            }
        }
    }
    __syncthreads();
    if (s_flag) { if (threadIdx.x == 0) atomicOr(q.ctr + 1, 1u); return; }          // peers will time out waiting for the header
    const unsigned long long* hd = reinterpret_cast<const unsigned long long*>(q.src);
    const unsigned long long n16 = hd[3] >> 4;                                        // blob bytes / 16 (sections are 16-byte aligned)
    const uint4* s = reinterpret_cast<const uint4*>(q.src);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long i = 4 + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;   // the 64-byte header goes last
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 v0 = s[i], v1 = s[i + stride], v2 = s[i + 2 * stride], v3 = s[i + 3 * stride];
#pragma unroll 1
        for (int k = 0; k < q.nranks; ++k) {
            int r = q.rank + 1 + k; if (r >= q.nranks) r -= q.nranks;              // start at the next rank: spreads the switch ports
            uint4* d = reinterpret_cast<uint4*>(q.dst[r]);
            d[i] = v0; d[i + stride] = v1; d[i + 2 * stride] = v2; d[i + 3 * stride] = v3;
        }
    }
    for (; i < n16; i += stride) {
        uint4 v = s[i];
        for (int k = 0; k < q.nranks; ++k) { int r = q.rank + 1 + k; if (r >= q.nranks) r -= q.nranks; reinterpret_cast<uint4*>(q.dst[r])[i] = v; }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_flag = (atomicAdd(q.ctr, 1u) == gridDim.x - 1) ? 2 : 0;
    __syncthreads();
    if (s_flag != 2) return;
    // last block: every block's stores are ordered before its counter increment; publish the header, sequence number last
    __threadfence_system();
    if (threadIdx.x < (unsigned)q.nranks) {
        unsigned long long* d = reinterpret_cast<unsigned long long*>(q.dst[threadIdx.x]);
        for (int k = 0; k < 7; ++k) st_sys(d + k, hd[k]);
        __threadfence_system();
        st_sys(d + 7, q.seq);
    }
    if (threadIdx.x == 0) *q.ctr = 0;
}



// Waits (on the exchange stream) until the headers of exchange `seq` from all ranks have landed in this rank's arena, and copies
// them out: hdr[8 * r + k], then the error word.
__global__ void ktn_wait_kernel(const unsigned char* slot_base, unsigned long long slot_cap, int nranks, unsigned long long seq,
                                unsigned long long* hdr, unsigned int* ctr) {
    const int r = threadIdx.x;
    if (r < nranks) {
        const unsigned long long* hd = reinterpret_cast<const unsigned long long*>(slot_base + slot_cap * (unsigned long long)r);
        const unsigned long long t0 = now_ns();
        bool ok = true;
        while (ld_sys(hd + 7) != seq) {
            if (now_ns() - t0 > KTN_PX_TIMEOUT_NS) { ok = false; atomicOr(ctr + 1, 2u); break; }
            __nanosleep(500);
        }
        __threadfence_system();
        for (int k = 0; k < 8; ++k) hdr[8 * r + k] = ok ? ld_sys(hd + k) : 0ull;
    }
    __syncthreads();
    if (r == 0) hdr[8 * nranks] = ctr[1];
}

#define NK(h, call) do { ncclResult_t r__ = (call); if (r__ != 0) return fail(h, KTN_ERR_NCCL, "%s failed: %s", #call, N.GetErrorString(r__)); } while (0)

// collective tear-down of the peer-push state: nobody frees an arena that a peer still has mapped or is still writing to
static void peer_release(ktn_handle* h) {
    ktn_handle::PeerExchange& px = h->px;
    if (px.arena.p) {
        cudaDeviceSynchronize();                                         // my pushes have left
        for (int r = 0; r < h->nranks && r < KTN_PX_MAX_RANKS; ++r)
            if (r != h->rank && px.peer[r]) cudaIpcCloseMemHandle(px.peer[r]);
        if (h->comm && N.ok && px.boot.p) {                              // barrier: every rank has drained its pushes and unmapped
            N.AllGather(px.boot.p, (char*)px.boot.p + 128, 8, ncclUint8, (ncclComm_t)h->comm, h->comm_stream);
            cudaStreamSynchronize(h->comm_stream);
        }
    }
    for (auto& q : px.peer) q = nullptr;
    px.arena.release(); px.boot.release(); px.hdr.release(); px.ctr.release();
    if (px.h_boot) { cudaFreeHost(px.h_boot); px.h_boot = nullptr; }
    px.on = px.tried = false; px.seq = 0; px.slot_cap = 0;
}

void ktn_comm_release(ktn_handle* h) {
    if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
    peer_release(h);
    if (h->comm && N.ok) N.CommDestroy((ncclComm_t)h->comm);
    h->comm = nullptr;
    for (auto& x : h->xch) {
        x.gathered.release(); x.all_counts.release(); x.stage.release();
        if (x.h_all_counts) { cudaFreeHost(x.h_all_counts); x.h_all_counts = nullptr; }
        if (x.packed) { cudaEventDestroy(x.packed); cudaEventDestroy(x.sizes); cudaEventDestroy(x.t0); cudaEventDestroy(x.t1); x.packed = x.sizes = x.t0 = x.t1 = nullptr; }
        x.state = 0;
    }
    for (int k = 0; k < 3; ++k) { if (h->blob_ev[k]) { cudaEventDestroy(h->blob_ev[k]); h->blob_ev[k] = nullptr; } h->blob_busy[k] = false; }
    if (h->comm_stream) { cudaStreamDestroy(h->comm_stream); h->comm_stream = nullptr; }
}

extern "C" int ktn_comm_unique_id(void* id128) {
    std::string why;
    if (!id128 || !load_nccl(&why)) { fprintf(stderr, "libktn: %s\n", why.c_str()); return KTN_ERR_NCCL; }
    ncclUniqueId id; if (N.GetUniqueId(&id) != 0) return KTN_ERR_NCCL;
    memcpy(id128, &id, 128);
    return KTN_OK;
}

extern "C" int ktn_comm_init(ktn_handle* h, int32_t nranks, int32_t rank, const void* id128) {
    if (!h || nranks < 1 || rank < 0 || rank >= nranks || !id128) return fail(h, KTN_ERR_USAGE, "bad communicator arguments");
    std::string why;
    if (!load_nccl(&why)) return fail(h, KTN_ERR_NCCL, "%s", why.c_str());
    cudaSetDevice(h->device);
    ktn_comm_release(h);
    ncclUniqueId id; memcpy(&id, id128, 128);
    ncclComm_t c = nullptr;
    NK(h, N.CommInitRank(&c, nranks, id, rank));
    h->comm = c; h->nranks = nranks; h->rank = rank;
    { int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi);       // the exchange gets its few SMs ahead of the next round's persistent kernel
      CK(h, cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, hi)); }
    for (auto& x : h->xch) {
        CK(h, x.all_counts.alloc(64 * (size_t)nranks + 64));
        CK(h, cudaMallocHost(&x.h_all_counts, 64 * (size_t)nranks + 64));
        x.g_cuts.assign(nranks, 0); x.g_nnz.assign(nranks, 0); x.g_off.assign(nranks + 1, 0);
        x.g_lay_cuts.assign(nranks, 0); x.g_lay_nnz.assign(nranks, 0); x.g_bytes.assign(nranks, 0);
        CK(h, cudaEventCreateWithFlags(&x.packed, cudaEventDisableTiming)); CK(h, cudaEventCreateWithFlags(&x.sizes, cudaEventDisableTiming));
        CK(h, cudaEventCreate(&x.t0)); CK(h, cudaEventCreate(&x.t1));
        x.state = 0;
    }
    for (int k = 0; k < 3; ++k) { CK(h, cudaEventCreateWithFlags(&h->blob_ev[k], cudaEventDisableTiming)); h->blob_busy[k] = false; }
    return KTN_OK;
}

extern "C" int ktn_set_row_offset(ktn_handle* h, int64_t first_global_row) {
    if (!h) return KTN_ERR_USAGE;
    h->row_offset = first_global_row;
    return KTN_OK;
}

// One small all-gather through NCCL, result on the host (bootstrap only).  `bytes` per rank, at most 128.
static int boot_allgather(ktn_handle* h, const void* mine, size_t bytes, unsigned char* all) {
    ktn_handle::PeerExchange& px = h->px;
    unsigned char* d = px.boot.as<unsigned char>();
    CK(h, cudaMemcpyAsync(d, mine, bytes, cudaMemcpyHostToDevice, h->comm_stream));
    NK(h, N.AllGather(d, d + 128, bytes, ncclUint8, (ncclComm_t)h->comm, h->comm_stream));
    CK(h, cudaMemcpyAsync(px.h_boot, d + 128, bytes * (size_t)h->nranks, cudaMemcpyDeviceToHost, h->comm_stream));
    CK(h, cudaStreamSynchronize(h->comm_stream));
    memcpy(all, px.h_boot, bytes * (size_t)h->nranks);
    return KTN_OK;
}

// Collective (first exchange): agree on the slot size, export / map the receive arenas, agree on the transport.
static int peer_setup(ktn_handle* h, size_t my_cap) {
    ktn_handle::PeerExchange& px = h->px;
    px.tried = true; px.on = false;
    const int R = h->nranks;
    const char* ex = getenv("KTN_EXCHANGE");
    const char* pb = getenv("KTN_PUSH_BLOCKS");
    if (pb && atoi(pb) > 0) { px.blocks = atoi(pb); px.blocks_fixed = true; }
    { const char* rs = getenv("KTN_PUSH_RESERVE"); if (rs) px.reserve = atoi(rs) != 0; }
    CK(h, px.boot.alloc(128 + 128 * (size_t)R)); CK(h, cudaMallocHost(&px.h_boot, 128 * (size_t)R));
    std::vector<unsigned char> all(128 * (size_t)R);
    // 1. slot size = the largest worst-case blob; every rank must want the peer transport
    unsigned long long a[2] = {(unsigned long long)my_cap, (unsigned long long)((!ex || strcmp(ex, "nccl") != 0) && R <= KTN_PX_MAX_RANKS)};
    int rc = boot_allgather(h, a, 16, all.data()); if (rc) return rc;
    unsigned long long cap = 0, want = 1;
    for (int r = 0; r < R; ++r) { unsigned long long v[2]; memcpy(v, all.data() + 16 * r, 16); if (v[0] > cap) cap = v[0]; want &= v[1]; }
    if (!want) return KTN_OK;
    // 2. allocate, clear the control page and the header of every sub-slot, export
    struct { cudaIpcMemHandle_t mh; unsigned long long ok; } mine, theirs;
    memset(&mine, 0, sizeof mine);
    const size_t bytes = KTN_PX_CTRL + (size_t)KTN_PX_SLOTS * R * cap;
    size_t free_b = 0, total_b = 0; cudaMemGetInfo(&free_b, &total_b);
    bool ok = bytes <= free_b / 4 && px.arena.alloc(bytes) == cudaSuccess && px.hdr.alloc(8 * (8 * (size_t)R + 8)) == cudaSuccess &&
              px.ctr.alloc(64) == cudaSuccess;
    if (ok) {
        ok = cudaMemset(px.arena.p, 0, KTN_PX_CTRL) == cudaSuccess && cudaMemset(px.ctr.p, 0, 64) == cudaSuccess;
        for (int k = 0; ok && k < KTN_PX_SLOTS * R; ++k) ok = cudaMemset(px.arena.as<unsigned char>() + KTN_PX_CTRL + (size_t)k * cap, 0, 64) == cudaSuccess;
        ok = ok && cudaDeviceSynchronize() == cudaSuccess && cudaIpcGetMemHandle(&mine.mh, px.arena.p) == cudaSuccess;
    }
    cudaGetLastError();
    mine.ok = ok;
    rc = boot_allgather(h, &mine, sizeof mine, all.data()); if (rc) return rc;
    bool all_ok = true;
    for (int r = 0; r < R; ++r) { memcpy(&theirs, all.data() + sizeof mine * r, sizeof mine); all_ok = all_ok && theirs.ok; }
    // 3. map the peers' arenas (enables peer access over NVLink), agree
    unsigned long long opened = all_ok;
    if (all_ok) {
        for (int r = 0; r < R && opened; ++r) {
            if (r == h->rank) { px.peer[r] = px.arena.as<unsigned char>(); continue; }
            memcpy(&theirs, all.data() + sizeof mine * r, sizeof mine);
            void* q = nullptr;
            if (cudaIpcOpenMemHandle(&q, theirs.mh, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { opened = 0; cudaGetLastError(); }
            else px.peer[r] = (unsigned char*)q;
        }
    }
    rc = boot_allgather(h, &opened, 8, all.data()); if (rc) return rc;
    for (int r = 0; r < R; ++r) { unsigned long long v; memcpy(&v, all.data() + 8 * r, 8); opened &= v; }
    if (!opened) {                                                       // some rank cannot reach some arena: everybody uses NCCL
        for (int r = 0; r < R; ++r) { if (r != h->rank && px.peer[r]) cudaIpcCloseMemHandle(px.peer[r]); px.peer[r] = nullptr; }
        rc = boot_allgather(h, &opened, 8, all.data()); if (rc) return rc;     // barrier before the arenas go away
        px.arena.release();
        return KTN_OK;
    }
    px.slot_cap = cap; px.on = true; px.seq = 0;
    return KTN_OK;
}

// Peer-push exchange of the last round: K2 left the cuts as one blob; the push runs on the exchange stream.
static int peer_exchange(ktn_handle* h, size_t my_cap) {
    ktn_handle::PeerExchange& px = h->px;
    if (my_cap > px.slot_cap) return fail(h, KTN_ERR_USAGE, "a larger problem was loaded after the first exchange: call ktn_comm_init again on every rank");
    const unsigned long long seq = ++px.seq;
    const int s = (int)(seq % KTN_PX_SLOTS);
    h->xch_cur = s;
    ktn_handle::Exchange& x = h->xch[s];
    x.src_idx = h->out_cur;
    CK(h, cudaEventRecord(x.packed, h->stream));
    CK(h, cudaStreamWaitEvent(h->comm_stream, x.packed, 0));
    KtnPushParams q; memset(&q, 0, sizeof q);
    for (int r = 0; r < h->nranks; ++r) {
        q.dst[r] = px.peer[r] + KTN_PX_CTRL + ((size_t)s * h->nranks + h->rank) * px.slot_cap;
        q.ack_dst[r] = reinterpret_cast<unsigned long long*>(px.peer[r]) + h->rank;
    }
    q.ack_local = px.arena.as<unsigned long long>(); q.src = h->out_blob[x.src_idx].as<unsigned char>(); q.seq = seq;
    q.ctr = px.ctr.as<unsigned int>(); q.nranks = h->nranks; q.rank = h->rank;
    CK(h, cudaEventRecord(x.t0, h->comm_stream));
    ktn_push_kernel<<<px.blocks, 512, 0, h->comm_stream>>>(q);
    CK(h, cudaGetLastError());
    CK(h, cudaEventRecord(x.t1, h->comm_stream));
    h->tm.launches += 1;
    CK(h, cudaEventRecord(h->blob_ev[x.src_idx], h->comm_stream)); h->blob_busy[x.src_idx] = true;
    x.state = 2;
    return KTN_OK;
}

// Peer-push: wait until the last exchange has arrived from every rank; sizes and offsets of the blobs.
static int peer_sync(ktn_handle* h) {
    ktn_handle::PeerExchange& px = h->px;
    if (px.seq == 0) return fail(h, KTN_ERR_USAGE, "no exchange has been enqueued");
    const int s = (int)(px.seq % KTN_PX_SLOTS), R = h->nranks;
    ktn_handle::Exchange& x = h->xch[s];
    const unsigned char* base = px.arena.as<unsigned char>() + KTN_PX_CTRL + (size_t)s * R * px.slot_cap;
    ktn_wait_kernel<<<1, 32, 0, h->comm_stream>>>(base, px.slot_cap, R, px.seq, px.hdr.as<unsigned long long>(), px.ctr.as<unsigned int>());
    CK(h, cudaGetLastError());
    CK(h, cudaMemcpyAsync(px.h_boot, px.hdr.p, 8 * (8 * (size_t)R + 1), cudaMemcpyDeviceToHost, h->comm_stream));
    CK(h, cudaStreamSynchronize(h->comm_stream));
    if (px.h_boot[8 * R]) return fail(h, KTN_ERR_NCCL, "peer exchange timed out (error word %llu): a rank did not enqueue the same exchanges", px.h_boot[8 * R]);
    for (int r = 0; r < R; ++r) {
        const unsigned long long* hd = px.h_boot + 8 * r;
        x.g_cuts[r] = (int64_t)hd[0]; x.g_nnz[r] = (int64_t)hd[1]; x.g_bytes[r] = (int64_t)hd[3];
        x.g_lay_cuts[r] = (int64_t)hd[5]; x.g_lay_nnz[r] = (int64_t)hd[6];
        x.g_off[r] = (int64_t)(px.slot_cap * (size_t)r);
    }
    x.g_off[R] = (int64_t)(px.slot_cap * (size_t)R);
    return KTN_OK;
}

// Second half of an exchange: the sizes of every rank are on the host (their copy was enqueued a round ago, so the wait is
// short); size the common slot and move the payloads with one all-gather over NVLink / NVSwitch.
static int launch_payload(ktn_handle* h, ktn_handle::Exchange& x) {
    if (x.state != 1) return KTN_OK;
    ncclComm_t comm = (ncclComm_t)h->comm;
    CK(h, cudaEventSynchronize(x.sizes));
    // every rank's blob travels in a slot of the size of the largest one: ONE ncclAllGather (NVSwitch-friendly) instead of a
    // broadcast per rank; with balanced shards the padding is negligible
    size_t slot = 0;
    for (int r = 0; r < h->nranks; ++r) {
        const unsigned long long* hd = x.h_all_counts + 8 * r;
        x.g_cuts[r] = (int64_t)hd[0]; x.g_nnz[r] = (int64_t)hd[1]; x.g_bytes[r] = (int64_t)hd[3];
        x.g_lay_cuts[r] = (int64_t)hd[5]; x.g_lay_nnz[r] = (int64_t)hd[6];
        if ((size_t)hd[3] > slot) slot = (size_t)hd[3];
    }
    slot = (slot + 127) & ~(size_t)127;
    for (int r = 0; r <= h->nranks; ++r) x.g_off[r] = (int64_t)(slot * (size_t)r);
    const size_t off = slot * (size_t)h->nranks;
    x.gathered_bytes = (int64_t)off;
    if (x.gathered.bytes < off) { CK(h, cudaStreamSynchronize(h->comm_stream)); CK(h, x.gathered.alloc(off + off / 4)); }
    // very unbalanced shards: the common slot may be larger than this rank's blob buffer.  No rank may bail out between collective
    // calls (the others would wait in the all-gather forever): the blob is staged into a buffer of the slot size instead
    const void* send = h->out_blob[x.src_idx].p;
    if (h->out_cap < slot) {
        if (x.stage.bytes < slot) { CK(h, cudaStreamSynchronize(h->comm_stream)); CK(h, x.stage.alloc(slot + slot / 4)); }
        const size_t mine = (size_t)x.g_bytes[h->rank] < h->out_cap ? (size_t)x.g_bytes[h->rank] : h->out_cap;
        CK(h, cudaMemcpyAsync(x.stage.p, h->out_blob[x.src_idx].p, mine, cudaMemcpyDeviceToDevice, h->comm_stream));
        send = x.stage.p;
    }
    CK(h, cudaEventRecord(x.t0, h->comm_stream));
    NK(h, N.AllGather(send, x.gathered.p, slot, ncclUint8, comm, h->comm_stream));
    CK(h, cudaEventRecord(x.t1, h->comm_stream));
    CK(h, cudaEventRecord(h->blob_ev[x.src_idx], h->comm_stream)); h->blob_busy[x.src_idx] = true;
    x.state = 2;
    return KTN_OK;
}

// Enqueue the exchange of the last round (K2 left its cuts as one blob).  Peer push: one kernel on the exchange stream, see the
// head of this file.  NCCL: the 64-byte headers are all-gathered now, nothing is waited for; the payloads go out two calls
// later (ktn_comm_launch_pending, from the round launcher), when their sizes have long reached the host, and overlap the
// kernels of later rounds; ktn_sync_gathered launches what is still outstanding.  Every rank makes the same sequence of calls.
extern "C" int ktn_allgather_cuts_async(ktn_handle* h) {
    if (!h || !h->loaded) return fail(h, KTN_ERR_USAGE, "no problem loaded");
    if (!h->comm) return fail(h, KTN_ERR_USAGE, "ktn_comm_init has not been called");
    cudaSetDevice(h->device);
    ncclComm_t comm = (ncclComm_t)h->comm;
    {
        const size_t m0 = (size_t)h->prob.num_constr, NZ0 = (size_t)h->prob.jac_ptr[m0];
        const size_t my_cap = ((ktn_pack_layout(m0, NZ0).total + 127) & ~(size_t)127) + 128;
        if (!h->round_pending && !h->have_round) return fail(h, KTN_ERR_USAGE, "no round has been enqueued");
        if (!h->px.tried) { int rc = peer_setup(h, my_cap); if (rc) return rc; }
        if (h->px.on) return peer_exchange(h, my_cap);
    }
    h->xch_cur = (h->xch_cur + 1) % 3;
    ktn_handle::Exchange& x = h->xch[h->xch_cur];
    // the slot reused here was last used three exchanges ago; its payload (and the one after it) goes out before anything of
    // this exchange is enqueued on the exchange stream, in the same order on every rank
    int rc = launch_payload(h, x); if (rc) return rc;
    rc = launch_payload(h, h->xch[(h->xch_cur + 1) % 3]); if (rc) return rc;
    x.src_idx = h->out_cur;
    CK(h, cudaEventRecord(x.packed, h->stream));
    CK(h, cudaStreamWaitEvent(h->comm_stream, x.packed, 0));
    // K2 wrote the 64-byte header at the head of the round's blob: all-gather the headers
    NK(h, N.AllGather(h->out_blob[x.src_idx].p, x.all_counts.p, 8, ncclUint64, comm, h->comm_stream));
    CK(h, cudaMemcpyAsync(x.h_all_counts, x.all_counts.p, 64 * (size_t)h->nranks, cudaMemcpyDeviceToHost, h->comm_stream));
    CK(h, cudaEventRecord(x.sizes, h->comm_stream));
    x.state = 1;
    return KTN_OK;
}

// Called by the round launcher before it enqueues the kernels of a new round: the payload of the exchange enqueued TWO calls ago
// goes out now (its sizes reached the host long ago, so the host does not stall), and runs beside the new round.
int ktn_comm_launch_pending(ktn_handle* h) {
    if (!h->comm || h->px.on) return KTN_OK;
    return launch_payload(h, h->xch[(h->xch_cur + 2) % 3]);       // the slot used before the previous one
}

// Called by the round launcher (peer-push transport) before it sizes K1's grid: how many SMs the push kernel gets.
// MEASURED defaults (10^6 log-sum-exp rows per GPU, v = 0.1; profiles/scale_r02_{4,8}gpu.log): 2 GPUs: 16 blocks (151 us per
// round); 4 GPUs: 32 blocks 180 us (16: 187, 24: 188, 8: 246); 8 GPUs: 32 blocks 323 us = 508 GB/s inbound per GPU (16: 391,
// 48: 350); final build (faster kernels, scripts/scale_final8.sh, one box): 24: 350 us, 32: 335, 40: 301 -> 40 blocks from 8 GPUs on.
// More blocks push faster but are taken from K1.  A volume-driven rule shipped unmeasured in round 1 and cost the
// 4- and 8-GPU runs a factor 1.8; it is available as KTN_PUSH_PLAN=volume for experiments, KTN_PUSH_BLOCKS fixes the grid.
void ktn_comm_plan_blocks(ktn_handle* h) {
    ktn_handle::PeerExchange& px = h->px;
    if (!h->comm || !px.on || px.blocks_fixed || !h->have_round) return;
    static const bool by_volume = getenv("KTN_PUSH_PLAN") && !strcmp(getenv("KTN_PUSH_PLAN"), "volume");
    if (!by_volume) { px.blocks = h->nranks <= 2 ? 16 : h->nranks < 8 ? 32 : 40; return; }
    const double out = (double)ktn_pack_layout((unsigned long long)h->lay_cuts, (unsigned long long)h->lay_nnz).total * (double)h->nranks;
    int b = (int)(out / 3.0e6) + 1;
    const int hi = h->num_sms / 3 < 48 ? h->num_sms / 3 : 48;
    if (b < 8) b = 8;
    if (b > hi) b = hi;
    if (b < 1) b = 1;
    px.blocks = b;
}

// Called by the round launcher before K2 may overwrite cut blob `idx`: whatever exchange still has to read it goes first.
int ktn_comm_release_blob(ktn_handle* h, int idx) {
    if (!h->comm) return KTN_OK;
    if (!h->px.on) for (auto& x : h->xch) if (x.state == 1 && x.src_idx == idx) { int rc = launch_payload(h, x); if (rc) return rc; }
    if (h->blob_busy[idx]) { CK(h, cudaStreamWaitEvent(h->stream, h->blob_ev[idx], 0)); h->blob_busy[idx] = false; }
    return KTN_OK;
}

extern "C" int ktn_exchange_transport(ktn_handle* h) {
    if (!h || !h->comm || !h->px.tried) return 0;
    return h->px.on ? 2 : 1;
}

extern "C" int ktn_sync_gathered(ktn_handle* h, int64_t* total_cuts, int64_t* total_nnz) {
    if (!h || !h->comm) return fail(h, KTN_ERR_USAGE, "no communicator");
    cudaSetDevice(h->device);
    ktn_handle::Exchange& x = h->xch[h->xch_cur];
    if (h->px.on) { int rc = peer_sync(h); if (rc) return rc; }
    else {
        int rc = launch_payload(h, h->xch[(h->xch_cur + 1) % 3]); if (rc) return rc;      // oldest first: the same order on every rank
        rc = launch_payload(h, h->xch[(h->xch_cur + 2) % 3]); if (rc) return rc;
        rc = launch_payload(h, x); if (rc) return rc;
        if (x.state != 2) return fail(h, KTN_ERR_USAGE, "no exchange has been enqueued");
        CK(h, cudaStreamSynchronize(h->comm_stream));
    }
    { float ms = 0.f; if (cudaEventElapsedTime(&ms, x.t0, x.t1) == cudaSuccess) { h->tm.exchange_ms = ms; h->exchange_ms_sum += ms; h->exchanges_timed++; } }
    // the reference stops at the first non-finite cut (src/model.jl:69-73, :278): the batch ends inside the first rank that saw
    // one (its own cuts are already truncated there); the ranks behind it contribute nothing.  Every rank reads the same headers,
    // so every rank returns the same status.
    const unsigned long long* hd = h->px.on ? h->px.h_boot : x.h_all_counts;
    int64_t c = 0, z = 0; x.g_err_row = -1;
    for (int r = 0; r < h->nranks; ++r) {
        if (x.g_err_row >= 0) { x.g_cuts[r] = 0; x.g_nnz[r] = 0; continue; }
        c += x.g_cuts[r]; z += x.g_nnz[r];
        if (hd[8 * r + 2] != ~0ull) x.g_err_row = (int64_t)(hd[8 * r + 2] - 1ull) + (int64_t)hd[8 * r + 4];
    }
    if (total_cuts) *total_cuts = c;
    if (total_nnz) *total_nnz = z;
    return x.g_err_row >= 0 ? KTN_NUMERIC_NONFINITE : KTN_OK;
}

extern "C" int ktn_gathered_error_row(ktn_handle* h, int64_t* err_row) {
    if (!h || !h->comm || !err_row) return fail(h, KTN_ERR_USAGE, "no communicator");
    *err_row = h->xch[h->xch_cur].g_err_row;
    return KTN_OK;
}

// Unpacks the gathered blobs of the LAST exchange into one CSR; row ids are global (K2 applied each rank's row offset).
extern "C" int ktn_fetch_gathered(ktn_handle* h, int64_t* row_id, int64_t* row_ptr, int32_t* col, double* val,
                                  double* lo, double* hi, double* g, double* viol, double* bconst) {
    const int status = ktn_sync_gathered(h, nullptr, nullptr); if (status < 0) return status;
    ktn_handle::Exchange& x = h->xch[h->xch_cur];
    // every rank's blob at its real size (the slots are larger: worst case, or the largest blob)
    const unsigned char* base = h->px.on ? h->px.arena.as<unsigned char>() + KTN_PX_CTRL + (size_t)h->xch_cur * h->nranks * h->px.slot_cap
                                         : x.gathered.as<unsigned char>();
    std::vector<size_t> hoff((size_t)h->nranks + 1, 0);
    for (int r = 0; r < h->nranks; ++r) hoff[r + 1] = hoff[r] + (size_t)x.g_bytes[r];
    std::vector<unsigned char> host(hoff[h->nranks] + 16);
    for (int r = 0; r < h->nranks; ++r)
        CK(h, cudaMemcpyAsync(host.data() + hoff[r], base + x.g_off[r], hoff[r + 1] - hoff[r], cudaMemcpyDeviceToHost, h->comm_stream));
    CK(h, cudaStreamSynchronize(h->comm_stream));
    int64_t co = 0, zo = 0;
    if (row_ptr) row_ptr[0] = 0;
    for (int r = 0; r < h->nranks; ++r) {
        const unsigned char* b = host.data() + hoff[r];
        const int64_t n = x.g_cuts[r], nz = x.g_nnz[r];                                      // truncated at a non-finite cut
        const KtnPackLayout L = ktn_pack_layout((unsigned long long)x.g_lay_cuts[r], (unsigned long long)x.g_lay_nnz[r]);
        if (row_id) memcpy(row_id + co, b + L.row_id, 8 * (size_t)n);     // global ids: K2 applied the rank's row offset
        if (row_ptr) { const int64_t* s = reinterpret_cast<const int64_t*>(b + L.row_ptr); for (int64_t i = 0; i < n; ++i) row_ptr[co + i + 1] = s[i + 1] + zo; }
        if (lo) memcpy(lo + co, b + L.lo, 8 * (size_t)n);
        if (hi) memcpy(hi + co, b + L.hi, 8 * (size_t)n);
        if (g) memcpy(g + co, b + L.g, 8 * (size_t)n);
        if (viol) memcpy(viol + co, b + L.viol, 8 * (size_t)n);
        if (bconst) memcpy(bconst + co, b + L.b, 8 * (size_t)n);
        if (col) memcpy(col + zo, b + L.col, 4 * (size_t)nz);
        if (val) memcpy(val + zo, b + L.val, 8 * (size_t)nz);
        co += n; zo += nz;
    }
    return status;
}
