// ktn_comm.cu -- sharded operation: every GPU separates its contiguous slice of the constraint rows, then the
// compacted cuts of all ranks are combined on every GPU over NCCL / NVLink.
// Rank-major concatenation is ascending row order, the reference's emission order (src/model.jl:272).
// NCCL is loaded lazily with dlopen so libktn.so has no link-time dependency on it.
#include <dlfcn.h>
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

#define NK(h, call) do { ncclResult_t r__ = (call); if (r__ != 0) return fail(h, KTN_ERR_NCCL, "%s failed: %s", #call, N.GetErrorString(r__)); } while (0)

void ktn_comm_release(ktn_handle* h) {
    if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
    if (h->comm && N.ok) N.CommDestroy((ncclComm_t)h->comm);
    h->comm = nullptr;
    for (auto& x : h->xch) {
        x.sendbuf.release(); x.gathered.release(); x.all_counts.release();
        if (x.h_all_counts) { cudaFreeHost(x.h_all_counts); x.h_all_counts = nullptr; }
        if (x.packed) { cudaEventDestroy(x.packed); cudaEventDestroy(x.sizes); cudaEventDestroy(x.done); x.packed = x.sizes = x.done = nullptr; }
        x.state = 0;
    }
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
    if (h->opt.topk > 0) return fail(h, KTN_ERR_UNSUPPORTED, "topk > 0 on a sharded handle: a global top-k needs a cross-rank selection, which is not built");
    std::string why;
    if (!load_nccl(&why)) return fail(h, KTN_ERR_NCCL, "%s", why.c_str());
    cudaSetDevice(h->device);
    ktn_comm_release(h);
    ncclUniqueId id; memcpy(&id, id128, 128);
    ncclComm_t c = nullptr;
    NK(h, N.CommInitRank(&c, nranks, id, rank));
    h->comm = c; h->nranks = nranks; h->rank = rank;
    CK(h, cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
    for (auto& x : h->xch) {
        CK(h, x.all_counts.alloc(16 * (size_t)nranks + 64));
        CK(h, cudaMallocHost(&x.h_all_counts, 16 * (size_t)nranks + 64));
        x.g_cuts.assign(nranks, 0); x.g_nnz.assign(nranks, 0); x.g_off.assign(nranks + 1, 0);
        CK(h, cudaEventCreateWithFlags(&x.packed, cudaEventDisableTiming)); CK(h, cudaEventCreateWithFlags(&x.sizes, cudaEventDisableTiming));
        CK(h, cudaEventCreateWithFlags(&x.done, cudaEventDisableTiming));
        x.state = 0;
    }
    return KTN_OK;
}

extern "C" int ktn_set_row_offset(ktn_handle* h, int64_t first_global_row) {
    if (!h) return KTN_ERR_USAGE;
    h->row_offset = first_global_row;
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
        x.g_cuts[r] = (int64_t)x.h_all_counts[2 * r]; x.g_nnz[r] = (int64_t)x.h_all_counts[2 * r + 1];
        const size_t t = ktn_pack_layout(x.g_cuts[r], x.g_nnz[r]).total;
        if (t > slot) slot = t;
    }
    slot = (slot + 127) & ~(size_t)127;
    for (int r = 0; r <= h->nranks; ++r) x.g_off[r] = (int64_t)(slot * (size_t)r);
    const size_t off = slot * (size_t)h->nranks;
    x.gathered_bytes = (int64_t)off;
    if (x.gathered.bytes < off) { CK(h, cudaStreamSynchronize(h->comm_stream)); CK(h, x.gathered.alloc(off + off / 4)); }
    if (x.sendbuf.bytes < slot) return fail(h, KTN_ERR_NCCL, "exchange slot larger than the send buffer");   // cannot happen: the buffer holds every row
    NK(h, N.AllGather(x.sendbuf.p, x.gathered.p, slot, ncclUint8, comm, h->comm_stream));
    CK(h, cudaEventRecord(x.done, h->comm_stream));
    x.state = 2;
    return KTN_OK;
}

// Enqueue the exchange of the last round.  First half now: pack this rank's cuts into one blob (on the round's stream) and
// all-gather the sizes (16 bytes per rank, on the exchange stream), nothing is waited for.  Payloads go out two calls later
// (ktn_comm_launch_pending, from the round launcher), when their sizes have long reached the host, and overlap the kernels
// of later rounds; ktn_sync_gathered launches what is still outstanding.  Every rank makes the same sequence of NCCL calls.
extern "C" int ktn_allgather_cuts_async(ktn_handle* h) {
    if (!h || !h->loaded) return fail(h, KTN_ERR_USAGE, "no problem loaded");
    if (!h->comm) return fail(h, KTN_ERR_USAGE, "ktn_comm_init has not been called");
    cudaSetDevice(h->device);
    ncclComm_t comm = (ncclComm_t)h->comm;
    h->xch_cur = (h->xch_cur + 1) % 3;
    ktn_handle::Exchange& x = h->xch[h->xch_cur];
    // the slot reused here was last used three exchanges ago; its payload (and the one after it) goes out before anything of
    // this exchange is enqueued on the exchange stream, in the same order on every rank
    int rc = launch_payload(h, x); if (rc) return rc;
    rc = launch_payload(h, h->xch[(h->xch_cur + 1) % 3]); if (rc) return rc;
    const size_t m = (size_t)h->prob.num_constr, NZ = (size_t)h->prob.jac_ptr[m];
    const size_t cap = ((ktn_pack_layout(m, NZ).total + 127) & ~(size_t)127) + 128;
    if (x.sendbuf.bytes < cap) { CK(h, cudaStreamSynchronize(h->comm_stream)); CK(h, x.sendbuf.alloc(cap)); }
    if (x.state == 2) CK(h, cudaStreamWaitEvent(h->stream, x.done, 0));
    CK(h, cudaEventRecord(h->evx0, h->stream));
    KtnRoundParams p = ktn_make_params(h, nullptr, 0, 0);
    ktn_launch_pack(p, x.sendbuf.as<unsigned char>(), h->num_sms, h->stream);
    h->tm.launches += 1;
    CK(h, cudaEventRecord(x.packed, h->stream));
    CK(h, cudaStreamWaitEvent(h->comm_stream, x.packed, 0));
    // the pack kernel wrote {n_cuts, nnz} at the head of the blob: all-gather those 16 bytes
    NK(h, N.AllGather(x.sendbuf.p, x.all_counts.p, 2, ncclUint64, comm, h->comm_stream));
    CK(h, cudaMemcpyAsync(x.h_all_counts, x.all_counts.p, 16 * (size_t)h->nranks, cudaMemcpyDeviceToHost, h->comm_stream));
    CK(h, cudaEventRecord(x.sizes, h->comm_stream));
    x.state = 1;
    // the next round must not overwrite the compacted outputs before the pack has read them: same stream, nothing to do.
    return KTN_OK;
}

// Called by the round launcher before it enqueues the kernels of a new round: the payload of the exchange enqueued TWO calls ago
// goes out now (its sizes reached the host long ago, so the host does not stall), and runs beside the new round.
int ktn_comm_launch_pending(ktn_handle* h) {
    if (!h->comm) return KTN_OK;
    return launch_payload(h, h->xch[(h->xch_cur + 2) % 3]);       // the slot used before the previous one
}

extern "C" int ktn_sync_gathered(ktn_handle* h, int64_t* total_cuts, int64_t* total_nnz) {
    if (!h || !h->comm) return fail(h, KTN_ERR_USAGE, "no communicator");
    cudaSetDevice(h->device);
    ktn_handle::Exchange& x = h->xch[h->xch_cur];
    int rc = launch_payload(h, h->xch[(h->xch_cur + 1) % 3]); if (rc) return rc;      // oldest first: the same order on every rank
    rc = launch_payload(h, h->xch[(h->xch_cur + 2) % 3]); if (rc) return rc;
    rc = launch_payload(h, x); if (rc) return rc;
    if (x.state != 2) return fail(h, KTN_ERR_USAGE, "no exchange has been enqueued");
    CK(h, cudaStreamSynchronize(h->comm_stream));
    int64_t c = 0, z = 0;
    for (int r = 0; r < h->nranks; ++r) { c += x.g_cuts[r]; z += x.g_nnz[r]; }
    if (total_cuts) *total_cuts = c;
    if (total_nnz) *total_nnz = z;
    return KTN_OK;
}

// Unpacks the gathered blobs of the LAST exchange into one CSR; row ids are global (K2 applied each rank's row offset).
extern "C" int ktn_fetch_gathered(ktn_handle* h, int64_t* row_id, int64_t* row_ptr, int32_t* col, double* val,
                                  double* lo, double* hi, double* g, double* viol, double* bconst) {
    int rc = ktn_sync_gathered(h, nullptr, nullptr); if (rc) return rc;
    ktn_handle::Exchange& x = h->xch[h->xch_cur];
    std::vector<unsigned char> host((size_t)x.gathered_bytes + 16);
    CK(h, cudaMemcpy(host.data(), x.gathered.p, (size_t)x.gathered_bytes, cudaMemcpyDeviceToHost));
    int64_t co = 0, zo = 0;
    if (row_ptr) row_ptr[0] = 0;
    for (int r = 0; r < h->nranks; ++r) {
        const unsigned char* b = host.data() + x.g_off[r];
        const unsigned long long* hd = reinterpret_cast<const unsigned long long*>(b);
        const int64_t n = (int64_t)hd[0], nz = (int64_t)hd[1];
        const KtnPackLayout L = ktn_pack_layout(n, nz);
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
    return KTN_OK;
}
