// ktn_api.cu -- the C ABI of include/ktn.h on top of the tape compiler and the sm_100a kernels.
// Host-side counterpart of KatanaFirstOrderSeparator's state (reference src/separators.jl:58-77):
// the handle owns the compiled problem, device buffers, a stream and pinned staging.
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include "ktn_handle.h"

void ktn_comm_release(ktn_handle* h);
int ktn_comm_launch_pending(ktn_handle* h);
int ktn_comm_release_blob(ktn_handle* h, int idx);
void ktn_comm_plan_blocks(ktn_handle* h);

// single-process sharded operation (ktn_options.ngpus > 1): defined at the end of this file
static int group_create(ktn_handle* front, ktn_handle** out);
static void group_destroy(ktn_handle* f);
static int group_fail(ktn_handle* f, ktn_handle* s, int rc);
static int group_load_begin(ktn_handle* f, int64_t num_var, int64_t num_constr);
static int group_add_rows(ktn_handle* f, int64_t first_row, int64_t nrows, const int64_t* eptr, const int32_t* op, const int32_t* arg, const double* val,
                          const double* lb, const double* ub, const uint8_t* flags);
static int group_load_end(ktn_handle* f);
static int group_round(ktn_handle* f, const double* x, const int64_t* rows, int64_t nrows, int do_round, int64_t* n_cuts, int64_t* nnz, int64_t* err_row);
static int group_fetch(ktn_handle* f, ktn_cut_view* view, int64_t* row_id, int64_t* row_ptr, int32_t* col, double* val, double* lo, double* hi, double* g, double* viol, double* bconst);

extern "C" const char* ktn_backend(void) { return "cuda"; }
extern "C" const char* ktn_last_error(ktn_handle* h) { return h ? h->err.c_str() : "null handle"; }

static void free_problem(ktn_handle* h) {
    DevBuf* all[] = {&h->chunks, &h->shapes, &h->prog, &h->blob, &h->chunk_rows, &h->chunk_lb, &h->chunk_ub, &h->jac_ptr, &h->jac_col,
                     &h->row_lb, &h->row_ub, &h->row_slot, &h->row_nl, &h->ladder, &h->rec, &h->worklist, &h->errpos, &h->blk_off, &h->chunk_jp, &h->dump, &h->x, &h->force, &h->g_row, &h->b_row, &h->sel, &h->stage_val, &h->big_scratch,
                     &h->blk_cnt, &h->topk_key, &h->topk_state, &h->topk_eqcnt, &h->table, &h->out_blob[0], &h->out_blob[1], &h->out_blob[2]};
    for (DevBuf* b : all) b->release();
    if (h->h_x) { cudaFreeHost(h->h_x); h->h_x = nullptr; }
    for (int i = 0; i < 2; ++i) { if (h->h_direct[i]) cudaFreeHost(h->h_direct[i]); h->h_direct[i] = h->d_direct[i] = nullptr; }
    h->direct = false;
    h->prob = KtnProblem();
    h->loaded = h->loading = h->round_pending = h->have_round = h->forced_last = false;
    h->n_cuts = h->nnz_cuts = 0; h->err_row = -1;
}

extern "C" int ktn_create(const ktn_options* o, ktn_handle** out) {
    if (!out) return KTN_ERR_USAGE;
    *out = nullptr;
    ktn_handle* h = new (std::nothrow) ktn_handle();
    if (!h) return KTN_ERR_NOMEM;
    memset(&h->opt, 0, sizeof h->opt); memset(&h->tm, 0, sizeof h->tm);
    h->opt.f_tol = 1e-6; h->opt.cut_coef_rng = 1e9; h->opt.device = -1;
    if (o) memcpy(&h->opt, o, (size_t)o->struct_size < sizeof(ktn_options) ? (size_t)o->struct_size : sizeof(ktn_options));
    if (h->opt.ngpus > 1) return group_create(h, out);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {   // no CPU fallback: the product path fails loudly without a GPU
        fprintf(stderr, "libktn: no CUDA device available (%s); this library has no CPU path\n", cudaGetErrorString(e));
        delete h; return KTN_ERR_CUDA;
    }
    if (h->opt.device >= 0) { if (cudaSetDevice(h->opt.device) != cudaSuccess) { delete h; return KTN_ERR_CUDA; } }
    cudaGetDevice(&h->device);
    cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device);
    cudaDeviceGetAttribute(&h->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return KTN_ERR_CUDA; }
    h->stream = h->own_stream;
    if (getenv("KTN_K1_EVENT_EVERY")) { const int v = atoi(getenv("KTN_K1_EVENT_EVERY")); if (v >= 1) h->mid_every = v; }
    cudaEventCreate(&h->ev0); cudaEventCreate(&h->ev1); cudaEventCreate(&h->ev2); cudaEventCreate(&h->ev3);
    for (int i = 0; i < ktn_handle::RING; ++i) for (int j = 0; j < 4; ++j) cudaEventCreate(&h->ring[i][j]);
    if (cudaMallocHost(&h->h_counts, 8 * sizeof(unsigned long long)) != cudaSuccess) { delete h; return KTN_ERR_CUDA; }
    if (h->ticket.alloc(4 * KTN_TICKETS) != cudaSuccess || h->counts.alloc(8 * sizeof(unsigned long long)) != cudaSuccess) { delete h; return KTN_ERR_CUDA; }
    cudaMemset(h->ticket.p, 0, 4 * KTN_TICKETS);
    { unsigned long long init[8] = {0, 0, ~0ull, ~0ull, 0, 0, ~0ull, 0}; cudaMemcpy(h->counts.p, init, 64, cudaMemcpyHostToDevice); }
    if (ktn_kernels_configure(h->max_smem) != cudaSuccess) { fprintf(stderr, "libktn: kernel configuration failed (is this an sm_100a device?)\n"); delete h; return KTN_ERR_CUDA; }
    *out = h;
    return KTN_OK;
}

extern "C" void ktn_destroy(ktn_handle* h) {
    if (!h) return;
    if (!h->shards.empty()) { group_destroy(h); return; }
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    free_problem(h);
    h->ticket.release(); h->counts.release();
    if (h->h_counts) cudaFreeHost(h->h_counts);
    for (int i = 0; i < 2; ++i) if (h->h_view[i]) cudaFreeHost(h->h_view[i]);
    ktn_comm_release(h);
    if (h->ev0) {
        cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1); cudaEventDestroy(h->ev2); cudaEventDestroy(h->ev3);
        for (int i = 0; i < ktn_handle::RING; ++i) for (int j = 0; j < 4; ++j) cudaEventDestroy(h->ring[i][j]);
    }
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

// device buffers of the top-k stage: one key per row, the select state, one tie count per compaction block
static int ensure_topk(ktn_handle* h) {
    if (h->opt.topk <= 0 || !h->loaded || h->topk_key.p) return KTN_OK;
    const size_t m = (size_t)h->prob.num_constr;
    CK(h, h->topk_key.alloc(8 * (m + 1))); CK(h, h->topk_state.alloc(sizeof(KtnTopkState))); CK(h, h->topk_eqcnt.alloc(4 * (size_t)h->blk_stride + 16));
    return KTN_OK;
}

extern "C" int ktn_set_params(ktn_handle* h, double f_tol, double rng, int64_t topk) {
    if (!h || topk < 0) return fail(h, KTN_ERR_USAGE, "bad parameters");
    if (!h->shards.empty()) {
        if (topk > 0) return fail(h, KTN_ERR_UNSUPPORTED, "topk is not available on a multi-device handle");
        for (ktn_handle* s : h->shards) { int rc = ktn_set_params(s, f_tol, rng, 0); if (rc) return group_fail(h, s, rc); }
        h->opt.f_tol = f_tol; h->opt.cut_coef_rng = rng;
        return KTN_OK;
    }
    h->opt.f_tol = f_tol; h->opt.cut_coef_rng = rng; h->opt.topk = topk;
    cudaSetDevice(h->device);
    return ensure_topk(h);
}

extern "C" int ktn_load_begin(ktn_handle* h, int64_t num_var, int64_t num_constr) {
    if (!h || num_var < 0 || num_constr < 0 || num_constr > 0x7fffff00ll || num_var > 0x7fffff00ll) return fail(h, KTN_ERR_USAGE, "bad problem sizes");
    if (!h->shards.empty()) return group_load_begin(h, num_var, num_constr);
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    free_problem(h);
    h->prob.reset(num_var, num_constr);
    h->loading = true;
    return KTN_OK;
}

extern "C" int ktn_add_rows(ktn_handle* h, int64_t first_row, int64_t nrows, const int64_t* eptr, const int32_t* op,
                            const int32_t* arg, const double* val, const double* lb, const double* ub, const uint8_t* flags) {
    if (h && !h->shards.empty()) return group_add_rows(h, first_row, nrows, eptr, op, arg, val, lb, ub, flags);
    if (!h || !h->loading) return fail(h, KTN_ERR_USAGE, "ktn_add_rows before ktn_load_begin");
    int rc = h->prob.add_rows(first_row, nrows, eptr, op, arg, val, lb, ub, flags);
    if (rc != KTN_OK) h->err = h->prob.err;
    return rc;
}

template <class T> static cudaError_t upload(DevBuf& b, const std::vector<T>& v, size_t extra = 0) {
    cudaError_t e = b.alloc((v.size() + extra) * sizeof(T) + 16);
    if (e != cudaSuccess) return e;
    if (!v.empty()) e = cudaMemcpy(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}

extern "C" int ktn_load_end(ktn_handle* h) {
    if (h && !h->shards.empty()) return group_load_end(h);
    if (!h || !h->loading) return fail(h, KTN_ERR_USAGE, "ktn_load_end before ktn_load_begin");
    cudaSetDevice(h->device);
    KtnProblem& P = h->prob;
    int64_t sigma = 8192;
    if (const char* s = getenv("KTN_SIGMA")) sigma = atoll(s);
    int rc = P.finalize(sigma, KTN_LANE_LIMIT);
    if (rc != KTN_OK) { h->err = P.err; return rc; }
    // block-shared table (shape descriptors + programs of the regular shapes) and the shared-memory plan.
    // If the table would not fit beside the warp regions every shape falls back to the global-memory kernel.
    std::vector<unsigned char> table;
    {
        std::vector<KtnShapeDesc> sh = P.shapes;
        std::vector<KtnIns> pr;
        for (KtnShapeDesc& s : sh) if (!(s.flags & KTN_SH_BIG)) { const uint32_t off = (uint32_t)pr.size(); pr.insert(pr.end(), P.prog.begin() + s.prog_off, P.prog.begin() + s.prog_off + s.n_ins); s.prog_off = off; }
        const size_t sb = (sh.size() * sizeof(KtnShapeDesc) + 15) & ~(size_t)15;
        table.assign(sb + pr.size() * sizeof(KtnIns) + 16, 0);
        memcpy(table.data(), sh.data(), sh.size() * sizeof(KtnShapeDesc));
        if (!pr.empty()) memcpy(table.data() + sb, pr.data(), pr.size() * sizeof(KtnIns));
        h->table_prog_off = (uint32_t)sb; h->table_bytes = (uint32_t)((sb + pr.size() * sizeof(KtnIns) + 15) & ~(size_t)15);
    }
    uint32_t blob_cap = 0, scratch_cap = 0;
    for (const KtnShapeDesc& s : P.shapes) if (!(s.flags & KTN_SH_BIG)) {
        uint32_t sec_col = (8u * s.n_const * 32u + 15u) & ~15u, sec_ord = (sec_col + 4u * s.n_uniq * 32u + 15u) & ~15u;
        uint32_t bytes = (sec_ord + s.order_bytes * s.n_uniq * 32u + 15u) & ~15u;
        if (bytes > blob_cap) blob_cap = bytes;
        if (s.n_scratch * 256u > scratch_cap) scratch_cap = s.n_scratch * 256u;
    }
    blob_cap = (blob_cap + 127u) & ~127u;
    h->blob_cap = blob_cap; h->warp_bytes = 128u + blob_cap + ((scratch_cap + 127u) & ~127u);
    const size_t m = (size_t)P.num_constr, N = (size_t)P.jac_ptr[m];
    if (N > 0xfffffff0ull) return fail(h, KTN_ERR_UNSUPPORTED, "more than 2^32 Jacobian entries on one device: shard the rows over more GPUs");
    CK(h, upload(h->chunks, P.chunks)); CK(h, upload(h->shapes, P.shapes)); CK(h, upload(h->prog, P.prog));
    CK(h, upload(h->blob, P.blob)); CK(h, upload(h->chunk_rows, P.chunk_rows));
    CK(h, upload(h->chunk_lb, P.chunk_lb)); CK(h, upload(h->chunk_ub, P.chunk_ub));
    CK(h, upload(h->jac_ptr, P.jac_ptr)); CK(h, upload(h->jac_col, P.jac_col));
    CK(h, upload(h->row_lb, P.lb)); CK(h, upload(h->row_ub, P.ub)); CK(h, upload(h->row_slot, P.row_slot));
    { std::vector<uint8_t> nl(P.flags.size()); for (size_t i = 0; i < nl.size(); ++i) nl[i] = (P.flags[i] & KTN_ROW_NL) ? 1 : 0; CK(h, upload(h->row_nl, nl)); }
    CK(h, h->rec.alloc(32 * (m + 1))); CK(h, h->worklist.alloc(8 * (m + 1))); CK(h, h->park.alloc(32 * (m + 1))); CK(h, h->cut_off.alloc(8 * (m + 2)));
    CK(h, upload(h->chunk_jp, P.chunk_jp));
    CK(h, h->x.alloc(8 * ((size_t)P.num_var + 1))); CK(h, h->force.alloc(m + 16));
    CK(h, h->g_row.alloc(8 * (m + 1))); CK(h, h->b_row.alloc(8 * (m + 1))); CK(h, h->sel.alloc(4 * (m + 1)));
    CK(h, cudaMemset(h->sel.p, 0, 4 * (m + 1))); CK(h, cudaMemset(h->g_row.p, 0, 8 * (m + 1)));
    CK(h, h->stage_val.alloc(8 * (N + 1))); CK(h, h->big_scratch.alloc(8 * (P.big_scratch_doubles + 1)));
    h->blk_stride = (uint32_t)((m + KTN_CROWS - 1) / KTN_CROWS + 1);
    CK(h, h->blk_cnt.alloc(16 * (size_t)h->blk_stride));
    CK(h, cudaMemset(h->blk_cnt.p, 0, 16 * (size_t)h->blk_stride));
    CK(h, h->errpos.alloc(16 * (size_t)h->blk_stride)); CK(h, h->blk_off.alloc(16 * ((size_t)h->blk_stride + 1)));
    CK(h, upload(h->table, table));
    h->out_cap = ((size_t)ktn_pack_layout(m, N).total + 127) / 128 * 128 + 128;      // every row selected
    h->out_cur = 0; for (bool& b : h->blob_busy) b = false;
    CK(h, h->out_blob[0].alloc(h->out_cap)); h->out_blob[1].release(); h->out_blob[2].release();     // [1], [2]: sharded handles, on first use
    CK(h, cudaMemset(h->out_blob[0].p, 0, 128));       // a shard without rows never runs K2: its blob header must read "nothing"
    CK(h, cudaMallocHost(&h->h_x, 8 * ((size_t)P.num_var + 1)));
    if ((h->opt.flags & KTN_FLAG_DIRECT_VIEW) && 2 * h->out_cap <= ((size_t)2 << 30)) {      // the cut blob in mapped pinned host memory
        for (int i = 0; i < 2; ++i) {
            void* q = nullptr; void* d = nullptr;
            CK(h, cudaHostAlloc(&q, h->out_cap, cudaHostAllocMapped));
            h->h_direct[i] = static_cast<unsigned char*>(q); memset(q, 0, 128);
            CK(h, cudaHostGetDevicePointer(&d, q, 0));
            h->d_direct[i] = static_cast<unsigned char*>(d);
        }
        h->direct = true; h->direct_cur = 0;
    }
    // the packed blob lives on the device now
    std::vector<uint8_t>().swap(P.blob);
    h->loading = false; h->loaded = true;
    return ensure_topk(h);
}

extern "C" int ktn_set_bounds(ktn_handle* h, const double* lb, const double* ub) {
    if (h && !h->shards.empty()) {
        for (size_t s = 0; s < h->shards.size(); ++s) { int rc = ktn_set_bounds(h->shards[s], lb + h->shard_begin[s], ub + h->shard_begin[s]); if (rc) return group_fail(h, h->shards[s], rc); }
        return KTN_OK;
    }
    if (!h || !h->loaded) return fail(h, KTN_ERR_USAGE, "no problem loaded");
    cudaSetDevice(h->device);
    KtnProblem& P = h->prob;
    P.lb.assign(lb, lb + P.num_constr); P.ub.assign(ub, ub + P.num_constr);
    P.repack_bounds();
    CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaMemcpy(h->row_lb.p, P.lb.data(), 8 * P.lb.size(), cudaMemcpyHostToDevice));
    CK(h, cudaMemcpy(h->row_ub.p, P.ub.data(), 8 * P.ub.size(), cudaMemcpyHostToDevice));
    CK(h, cudaMemcpy(h->chunk_lb.p, P.chunk_lb.data(), 8 * P.chunk_lb.size(), cudaMemcpyHostToDevice));
    CK(h, cudaMemcpy(h->chunk_ub.p, P.chunk_ub.data(), 8 * P.chunk_ub.size(), cudaMemcpyHostToDevice));
    return KTN_OK;
}

extern "C" int64_t ktn_num_rows(ktn_handle* h) {
    if (h && !h->shards.empty()) { int64_t n = 0; for (ktn_handle* s : h->shards) n += s->prob.rows_loaded; return n; }
    return h ? h->prob.rows_loaded : 0;
}
extern "C" int64_t ktn_jac_nnz(ktn_handle* h) {
    if (h && !h->shards.empty()) { int64_t n = 0; for (ktn_handle* s : h->shards) n += ktn_jac_nnz(s); return n; }
    return (h && !h->prob.jac_ptr.empty()) ? h->prob.jac_ptr.back() : 0;
}
extern "C" int ktn_jac_structure(ktn_handle* h, int64_t* row_ptr, int32_t* cols) {
    if (h && !h->shards.empty()) {      // concatenation of the shards' structures
        int64_t rows = 0, nz = 0;
        for (ktn_handle* s : h->shards) {
            const KtnProblem& P = s->prob;
            if (P.jac_ptr.empty()) return fail(h, KTN_ERR_USAGE, "no problem");
            if (row_ptr) for (int64_t i = 0; i < P.num_constr; ++i) row_ptr[rows + i] = nz + P.jac_ptr[i];
            if (cols && !P.jac_col.empty()) memcpy(cols + nz, P.jac_col.data(), 4 * P.jac_col.size());
            rows += P.num_constr; nz += P.jac_ptr[P.num_constr];
        }
        if (row_ptr) row_ptr[rows] = nz;
        return KTN_OK;
    }
    if (!h || h->prob.jac_ptr.empty()) return fail(h, KTN_ERR_USAGE, "no problem");
    if (row_ptr) memcpy(row_ptr, h->prob.jac_ptr.data(), 8 * h->prob.jac_ptr.size());
    if (cols) memcpy(cols, h->prob.jac_col.data(), 4 * h->prob.jac_col.size());
    return KTN_OK;
}

static KtnLaunchPlan make_plan(const ktn_handle* h) {
    KtnLaunchPlan pl;
    for (int f = 0; f <= KTN_FAM__COUNT; ++f) pl.fam_begin[f] = h->prob.fam_begin[f];
    memcpy(pl.cls_begin, h->prob.cls_begin, sizeof pl.cls_begin);
    memcpy(pl.cls_blob_off, h->prob.cls_blob_off, sizeof pl.cls_blob_off);
    pl.n_regular = h->prob.n_regular_chunks; pl.n_total = (uint32_t)h->prob.chunks.size();
    return pl;
}

KtnRoundParams ktn_make_params(ktn_handle* h, const double* d_x, int mode, int do_round) {
    KtnRoundParams p; memset(&p, 0, sizeof p);
    p.chunks = h->chunks.as<KtnChunkDesc>(); p.shapes = h->shapes.as<KtnShapeDesc>(); p.prog = h->prog.as<KtnIns>();
    p.blob = h->blob.as<uint8_t>(); p.chunk_rows = h->chunk_rows.as<int32_t>();
    p.chunk_lb = h->chunk_lb.as<double>(); p.chunk_ub = h->chunk_ub.as<double>();
    p.jac_ptr = h->jac_ptr.as<int64_t>(); p.jac_col = h->jac_col.as<int32_t>();
    p.row_lb = h->row_lb.as<double>(); p.row_ub = h->row_ub.as<double>(); p.row_slot = h->row_slot.as<int32_t>();
    p.chunk_jp = h->chunk_jp.as<uint32_t>(); p.dump = nullptr; p.dump_nnz = 0;      // (register-dump experiment of round 2, DESIGN.md section 5: measured, dropped; the buffer is no longer allocated)
    p.rec = h->rec.as<double4>(); p.park = h->park.as<double4>(); p.cut_off = h->cut_off.as<unsigned long long>(); p.worklist = h->worklist.as<unsigned long long>(); p.errpos = h->errpos.as<unsigned long long>(); p.blk_off = h->blk_off.as<unsigned long long>();
    for (int f = 0; f <= KTN_FAM__COUNT; ++f) p.fam_begin[f] = h->prob.fam_begin[f];
    memcpy(p.cls_begin, h->prob.cls_begin, sizeof p.cls_begin); memcpy(p.cls_blob_off, h->prob.cls_blob_off, sizeof p.cls_blob_off);
    p.x = d_x; p.force = h->force.as<uint8_t>();
    p.f_tol = h->opt.f_tol; p.rng = h->opt.cut_coef_rng; p.mode = mode; p.do_round = do_round;
    p.topk = h->opt.topk;      // sharded handles: a LOCAL top-k of this rank's rows; the global one is the merge of the union (sharding.py merge_topk)
    p.topk_key = h->topk_key.as<unsigned long long>(); p.topk_state = h->topk_state.as<KtnTopkState>(); p.topk_eqcnt = h->topk_eqcnt.as<unsigned int>();
    p.num_var = h->prob.num_var; p.num_rows = h->prob.num_constr;
    p.warp_bytes = h->warp_bytes; p.blob_cap = h->blob_cap;
    p.g_row = h->g_row.as<double>(); p.b_row = h->b_row.as<double>(); p.sel = h->sel.as<uint32_t>();
    p.stage_val = h->stage_val.as<double>(); p.big_scratch = h->big_scratch.as<double>(); p.ticket = h->ticket.as<unsigned int>();
    p.blk_cnt = h->blk_cnt.as<unsigned long long>(); p.blk_stride = h->blk_stride;
    p.counts = h->counts.as<unsigned long long>();
    p.row_offset = h->row_offset;
    p.table = h->table.as<unsigned char>(); p.table_bytes = h->table_bytes; p.table_prog_off = h->table_prog_off; p.epoch = h->epoch;
    p.out_blob = h->out_blob[h->out_cur].as<unsigned char>();
    if (h->direct && !h->comm) {       // forced rounds (gencut) keep every section: their callers fetch copies with g / viol / b
        p.out_blob = h->d_direct[h->direct_cur]; p.lean_out = ((h->opt.flags & KTN_FLAG_LEAN_VIEW) && mode == KTN_MODE_SEPARATE) ? 1 : 0;
        h->round_lean = p.lean_out != 0;
    }
    return p;
}

static void drain_ring(ktn_handle* h, bool all) {
    while (h->ring_tail != h->ring_head) {
        cudaEvent_t* e = h->ring[h->ring_tail % ktn_handle::RING];
        if (!all && cudaEventQuery(e[2]) != cudaSuccess) break;
        float a = 0.f, b = 0.f, c = 0.f;      // K1 | K2 + K3 | K2 alone (KTN_FLAG_TIME_KERNELS: an event is recorded between K2 and K3, which costs the round a few microseconds)
        const bool detail = (h->opt.flags & KTN_FLAG_TIME_KERNELS) != 0;
        if (cudaEventElapsedTime(&a, e[0], e[1]) == cudaSuccess && cudaEventElapsedTime(&b, e[1], e[2]) == cudaSuccess && (!detail || cudaEventElapsedTime(&c, e[1], e[3]) == cudaSuccess)) {
            if (!detail) c = b;
            h->eval_ms_sum += a; h->compact_ms_sum += c; h->cut_ms_sum += b - c; h->rounds_timed++;
            h->tm.kernel_ms = a + b; h->tm.eval_ms = a; h->tm.compact_ms = c; h->tm.cut_ms = b - c;
        }
        h->ring_tail++;
    }
}

static int enqueue_round(ktn_handle* h, const double* d_x, int mode, int do_round) {
    { int rc = ktn_comm_launch_pending(h); if (rc) return rc; }       // sharded runs: the previous round's cut payload travels beside this round
    if (h->comm) {      // the next of three cut blobs; the exchange that last read it (three rounds ago) must have finished
        h->out_cur = (h->out_cur + 1) % 3;
        if (!h->out_blob[h->out_cur].p) { CK(h, h->out_blob[h->out_cur].alloc(h->out_cap)); CK(h, cudaMemsetAsync(h->out_blob[h->out_cur].p, 0, 128, h->stream)); }
        { int rc = ktn_comm_release_blob(h, h->out_cur); if (rc) return rc; }
    }
    h->epoch = (h->epoch % 0x3ffffff0u) + 1u;
    if (h->direct && !h->comm) h->direct_cur ^= 1;      // a view stays valid until the round after the next one
    KtnRoundParams p = ktn_make_params(h, d_x, mode, do_round);
    p.clear_unselected = h->forced_last ? 1 : 0; h->forced_last = mode == KTN_MODE_FORCE;
    cudaError_t e = cudaSuccess;
    (void)cudaGetLastError();   // a stale non-sticky error must not be blamed on this round's launches
    // the kernels of a round are timed with CUDA events around and between them in one round of mid_every (KTN_K1_EVENT_EVERY; every
    // round with KTN_FLAG_TIME_KERNELS): the records cost 2-4 us per round and keep the next kernel from being placed early
    const bool detail = (h->opt.flags & KTN_FLAG_TIME_KERNELS) != 0;
    const bool mid = detail || h->mid_every <= 1 || (h->tm.rounds % h->mid_every) == 0;
    if (mid && h->ring_head - h->ring_tail >= ktn_handle::RING) drain_ring(h, false);
    if (mid && h->ring_head - h->ring_tail >= ktn_handle::RING) { CK(h, cudaEventSynchronize(h->ring[h->ring_tail % ktn_handle::RING][2])); drain_ring(h, false); }
    cudaEvent_t* ev = h->ring[h->ring_head % ktn_handle::RING];
    if (mid) CK(h, cudaEventRecord(ev[0], h->stream));
    // peer-push exchange: the persistent K1 leaves the push kernel's SMs free, so that the push of the previous round starts at once
    // beside this round instead of queueing behind K1's resident blocks (K1 owns every register of the SMs it runs on)
    ktn_comm_plan_blocks(h);
    int sms = h->num_sms;
    if (h->comm && h->px.on && h->px.reserve && sms > 2 * h->px.blocks) sms -= h->px.blocks;
    if (h->reserve_sms > 0 && sms > 2 * h->reserve_sms) sms -= h->reserve_sms;
    int n = ktn_launch_round(p, make_plan(h), sms, h->max_smem, h->epoch, h->stream, mid ? ev[1] : nullptr, detail ? ev[3] : nullptr, &e);
    h->tm.launches += n;
    if (e != cudaSuccess) return fail(h, KTN_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    if (mid) { CK(h, cudaEventRecord(ev[2], h->stream)); h->ring_head++; }
    h->round_pending = true; h->tm.rounds++;      // the counts are read back when somebody asks for them (finish_round): no copy between back-to-back rounds
    return KTN_OK;
}

static int finish_round(ktn_handle* h, int64_t* n_cuts, int64_t* nnz, int64_t* err_row) {
    if (!h->round_pending) return fail(h, KTN_ERR_USAGE, "no round is pending");
    CK(h, cudaMemcpyAsync(h->h_counts, h->counts.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    h->round_pending = false; h->have_round = true;
    drain_ring(h, true);
    h->n_cuts = (int64_t)h->h_counts[0]; h->nnz_cuts = (int64_t)h->h_counts[1];
    h->lay_cuts = (int64_t)h->h_counts[4]; h->lay_nnz = (int64_t)h->h_counts[5];
    h->err_row = h->h_counts[6] == ~0ull ? -1 : (int64_t)h->h_counts[6] - 1;
    if (n_cuts) *n_cuts = h->n_cuts;
    if (nnz) *nnz = h->nnz_cuts;
    if (err_row) *err_row = h->err_row;
    return h->err_row >= 0 ? KTN_NUMERIC_NONFINITE : KTN_OK;
}

static int upload_x(ktn_handle* h, const double* x) {
    memcpy(h->h_x, x, 8 * (size_t)h->prob.num_var);
    if (h->opt.flags & KTN_FLAG_TIME_KERNELS) CK(h, cudaEventRecord(h->ev0, h->stream));
    CK(h, cudaMemcpyAsync(h->x.p, h->h_x, 8 * (size_t)h->prob.num_var, cudaMemcpyHostToDevice, h->stream));
    if (h->opt.flags & KTN_FLAG_TIME_KERNELS) CK(h, cudaEventRecord(h->ev1, h->stream));
    return KTN_OK;
}

extern "C" int ktn_separate(ktn_handle* h, const double* xstar, int64_t* n_cuts, int64_t* nnz, int64_t* err_row) {
    if (h && !h->shards.empty()) return group_round(h, xstar, nullptr, 0, 1, n_cuts, nnz, err_row);
    if (!h || !h->loaded) return fail(h, KTN_ERR_USAGE, "no problem loaded");
    cudaSetDevice(h->device);
    int rc = upload_x(h, xstar); if (rc) return rc;
    rc = enqueue_round(h, h->x.as<double>(), KTN_MODE_SEPARATE, 1); if (rc) return rc;
    rc = finish_round(h, n_cuts, nnz, err_row);
    float ms = 0.f; if ((h->opt.flags & KTN_FLAG_TIME_KERNELS) && cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->tm.h2d_ms = ms;
    return rc;
}

extern "C" int ktn_gencut_rows(ktn_handle* h, const double* x, const int64_t* rows, int64_t nrows, int do_round,
                               int64_t* n_cuts, int64_t* nnz, int64_t* err_row) {
    if (h && !h->shards.empty()) {
        if (nrows < 0 || (nrows > 0 && !rows)) return fail(h, KTN_ERR_USAGE, "bad row list");
        for (int64_t j = 0; j < nrows; ++j) if (rows[j] < 0 || rows[j] >= h->g_num_constr || (j && rows[j] <= rows[j - 1])) return fail(h, KTN_ERR_USAGE, "rows must be ascending and in range");
        return group_round(h, x, rows ? rows : (const int64_t*)&nrows, nrows, do_round ? 1 : 0, n_cuts, nnz, err_row);
    }
    if (!h || !h->loaded) return fail(h, KTN_ERR_USAGE, "no problem loaded");
    cudaSetDevice(h->device);
    const int64_t m = h->prob.num_constr;
    std::vector<uint8_t> mask((size_t)m + 1, 0);
    for (int64_t j = 0; j < nrows; ++j) {
        if (rows[j] < 0 || rows[j] >= m || (j && rows[j] <= rows[j - 1])) return fail(h, KTN_ERR_USAGE, "rows must be ascending and in range");
        mask[rows[j]] = 1;
    }
    CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaMemcpy(h->force.p, mask.data(), (size_t)m, cudaMemcpyHostToDevice));
    int rc = upload_x(h, x); if (rc) return rc;
    rc = enqueue_round(h, h->x.as<double>(), KTN_MODE_FORCE, do_round ? 1 : 0); if (rc) return rc;
    return finish_round(h, n_cuts, nnz, err_row);
}

// boundroutine's ladder, src/model.jl:175-197: see include/ktn.h
void ktn_launch_scale(const double* ray, double* x, int64_t n, double scale, cudaStream_t stream);
void ktn_launch_anyviol(const KtnRoundParams& p, const uint8_t* row_nl, unsigned int* flag, cudaStream_t stream);
extern "C" int ktn_separate_ladder(ktn_handle* h, const double* ray, int32_t n_first, int32_t n_last, int32_t* n_hit,
                                   int64_t* n_cuts, int64_t* nnz, int64_t* err_row) {
    if (h && !h->shards.empty()) return fail(h, KTN_ERR_UNSUPPORTED, "not available on a multi-device handle");
    if (!h || !h->loaded || !ray || n_first < 0 || n_last > 1023 || n_last < n_first) return fail(h, KTN_ERR_USAGE, "bad ladder");
    cudaSetDevice(h->device);
    const int64_t nv = h->prob.num_var;
    const int B = 16;
    // device copy of the ray behind x (x itself is rewritten per point), one violation flag per point of a batch
    if (h->ladder.bytes < 8 * ((size_t)nv + 1) + 4 * B) CK(h, h->ladder.alloc(8 * ((size_t)nv + 1) + 4 * B));
    double* d_ray = h->ladder.as<double>();
    unsigned int* d_flag = reinterpret_cast<unsigned int*>(h->ladder.as<unsigned char>() + 8 * ((size_t)nv + 1));
    memcpy(h->h_x, ray, 8 * (size_t)nv);
    CK(h, cudaMemcpyAsync(d_ray, h->h_x, 8 * (size_t)nv, cudaMemcpyHostToDevice, h->stream));
    int hit = -1;
    unsigned int flags[B];
    for (int n0 = n_first; n0 <= n_last && hit < 0; n0 += B) {
        const int cnt = n_last - n0 + 1 < B ? n_last - n0 + 1 : B;
        CK(h, cudaMemsetAsync(d_flag, 0, 4 * B, h->stream));
        for (int k = 0; k < cnt; ++k) {       // no host round trip between the points of a batch
            ktn_launch_scale(d_ray, h->x.as<double>(), nv, ldexp(1.0, n0 + k), h->stream);
            KtnRoundParams p = ktn_make_params(h, h->x.as<double>(), KTN_MODE_SEPARATE, 0);
            cudaError_t e = cudaSuccess;
            h->tm.launches += 2 + ktn_launch_eval(p, make_plan(h), h->num_sms, h->max_smem, h->stream, &e);
            if (e != cudaSuccess) return fail(h, KTN_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
            ktn_launch_anyviol(p, h->row_nl.as<uint8_t>(), d_flag + k, h->stream);
        }
        CK(h, cudaMemcpyAsync(flags, d_flag, 4 * (size_t)cnt, cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
        for (int k = 0; k < cnt; ++k) if (flags[k]) { hit = n0 + k; break; }
    }
    if (n_hit) *n_hit = hit;
    if (hit < 0) {       // every row satisfied along the whole ladder: an empty batch (a round at the last point, which selects nothing)
        hit = n_last;
    }
    std::vector<double> x((size_t)nv);
    const double sc = ldexp(1.0, hit);
    for (int64_t j = 0; j < nv; ++j) x[j] = sc * ray[j];      // (2.0^n) * ray, src/model.jl:183
    return ktn_separate(h, x.data(), n_cuts, nnz, err_row);
}

extern "C" int ktn_fetch_cuts(ktn_handle* h, int64_t* row_id, int64_t* row_ptr, int32_t* col, double* val,
                              double* lo, double* hi, double* g, double* viol, double* bconst) {
    if (h && !h->shards.empty()) return group_fetch(h, nullptr, row_id, row_ptr, col, val, lo, hi, g, viol, bconst);
    if (!h || !h->loaded) return fail(h, KTN_ERR_USAGE, "no problem loaded");
    if (h->round_pending) { int rc = finish_round(h, nullptr, nullptr, nullptr); if (rc < 0) return rc; }
    cudaSetDevice(h->device);
    const size_t nc = (size_t)h->n_cuts, nz = (size_t)h->nnz_cuts;
    const KtnPackLayout S = ktn_pack_layout((unsigned long long)h->lay_cuts, (unsigned long long)h->lay_nnz);      // where K2 put the sections
    const unsigned char* src = h->out_blob[h->out_cur].as<unsigned char>();
    if (h->direct && !h->comm) {       // the batch is in host memory already (the round has been waited for)
        const unsigned char* b = h->h_direct[h->direct_cur];
        const bool lean = h->round_lean;
        if ((g || viol || bconst) && lean) return fail(h, KTN_ERR_USAGE, "g, viol and bconst are not produced with KTN_FLAG_DIRECT_VIEW | KTN_FLAG_LEAN_VIEW");
        if (row_id) memcpy(row_id, b + S.row_id, 8 * nc);
        if (row_ptr) { memcpy(row_ptr, b + S.row_ptr, 8 * nc); row_ptr[nc] = (int64_t)nz; }
        if (col) memcpy(col, b + S.col, 4 * nz);
        if (val) memcpy(val, b + S.val, 8 * nz);
        if (lo) memcpy(lo, b + S.lo, 8 * nc);
        if (hi) memcpy(hi, b + S.hi, 8 * nc);
        if (g) memcpy(g, b + S.g, 8 * nc);
        if (viol) memcpy(viol, b + S.viol, 8 * nc);
        if (bconst) memcpy(bconst, b + S.b, 8 * nc);
        return KTN_OK;
    }
    if (h->opt.flags & KTN_FLAG_TIME_KERNELS) CK(h, cudaEventRecord(h->ev2, h->stream));
    if (row_id && nc) CK(h, cudaMemcpyAsync(row_id, src + S.row_id, 8 * nc, cudaMemcpyDeviceToHost, h->stream));
    if (row_ptr) { if (nc) CK(h, cudaMemcpyAsync(row_ptr, src + S.row_ptr, 8 * nc, cudaMemcpyDeviceToHost, h->stream)); }
    if (col && nz) CK(h, cudaMemcpyAsync(col, src + S.col, 4 * nz, cudaMemcpyDeviceToHost, h->stream));
    if (val && nz) CK(h, cudaMemcpyAsync(val, src + S.val, 8 * nz, cudaMemcpyDeviceToHost, h->stream));
    if (lo && nc) CK(h, cudaMemcpyAsync(lo, src + S.lo, 8 * nc, cudaMemcpyDeviceToHost, h->stream));
    if (hi && nc) CK(h, cudaMemcpyAsync(hi, src + S.hi, 8 * nc, cudaMemcpyDeviceToHost, h->stream));
    if (g && nc) CK(h, cudaMemcpyAsync(g, src + S.g, 8 * nc, cudaMemcpyDeviceToHost, h->stream));
    if (viol && nc) CK(h, cudaMemcpyAsync(viol, src + S.viol, 8 * nc, cudaMemcpyDeviceToHost, h->stream));
    if (bconst && nc) CK(h, cudaMemcpyAsync(bconst, src + S.b, 8 * nc, cudaMemcpyDeviceToHost, h->stream));
    if (h->opt.flags & KTN_FLAG_TIME_KERNELS) CK(h, cudaEventRecord(h->ev3, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    if (row_ptr) row_ptr[nc] = (int64_t)nz;   // the device array ends at the untruncated total
    float ms = 0.f; if ((h->opt.flags & KTN_FLAG_TIME_KERNELS) && cudaEventElapsedTime(&ms, h->ev2, h->ev3) == cudaSuccess) h->tm.d2h_ms = ms;
    return KTN_OK;
}

// Zero-copy download into ONE library-owned pinned buffer (two buffers alternate, so a view stays valid until the round
// after the next one): the device blob already has the view's layout, so it comes down in one copy (lean view: two, around
// the g | viol | b sections); one synchronisation, no host-side copy.
extern "C" int ktn_fetch_cuts_view(ktn_handle* h, ktn_cut_view* out) {
    if (h && out && !h->shards.empty()) return group_fetch(h, out, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (!h || !h->loaded || !out) return fail(h, KTN_ERR_USAGE, "no problem loaded");
    if (h->round_pending) { int rc = finish_round(h, nullptr, nullptr, nullptr); if (rc < 0) return rc; }
    cudaSetDevice(h->device);
    const size_t nc = (size_t)h->n_cuts, nz = (size_t)h->nnz_cuts;
    const KtnPackLayout L = ktn_pack_layout(nc, nz);
    const bool lean = (h->opt.flags & KTN_FLAG_LEAN_VIEW) != 0;
    if (h->direct && !h->comm) {       // the kernels stored the batch in host memory: nothing to copy, the sections are where K2 laid them out
        const KtnPackLayout S = ktn_pack_layout((unsigned long long)h->lay_cuts, (unsigned long long)h->lay_nnz);
        unsigned char* b = h->h_direct[h->direct_cur];
        h->tm.d2h_ms = 0.0;
        out->n_cuts = (int64_t)nc; out->nnz = (int64_t)nz;     // a truncated round: row_ptr[nc] is the first entry of the cut the batch ends before = nz
        out->row_id = reinterpret_cast<const int64_t*>(b + S.row_id); out->row_ptr = reinterpret_cast<const int64_t*>(b + S.row_ptr);
        out->col = reinterpret_cast<const int32_t*>(b + S.col); out->val = reinterpret_cast<const double*>(b + S.val);
        out->lo = reinterpret_cast<const double*>(b + S.lo); out->hi = reinterpret_cast<const double*>(b + S.hi);
        out->g = lean ? nullptr : reinterpret_cast<const double*>(b + S.g); out->viol = lean ? nullptr : reinterpret_cast<const double*>(b + S.viol);
        out->bconst = lean ? nullptr : reinterpret_cast<const double*>(b + S.b);
        return KTN_OK;
    }
    h->view_cur ^= 1;
    unsigned char*& buf = h->h_view[h->view_cur]; size_t& cap = h->h_view_cap[h->view_cur];
    if (cap < L.total) {
        if (buf) cudaFreeHost(buf);
        buf = nullptr; cap = 0;
        const size_t want = L.total + L.total / 4 + 4096;
        CK(h, cudaMallocHost(&buf, want));
        cap = want;
    }
    const KtnPackLayout S = ktn_pack_layout((unsigned long long)h->lay_cuts, (unsigned long long)h->lay_nnz);      // where K2 put the sections
    const unsigned char* src = h->out_blob[h->out_cur].as<unsigned char>();
    if (h->opt.flags & KTN_FLAG_TIME_KERNELS) CK(h, cudaEventRecord(h->ev2, h->stream));
    if ((size_t)h->lay_cuts == nc && (size_t)h->lay_nnz == nz) {      // no truncation: source and view layouts coincide
        if (!lean) { if (nc) CK(h, cudaMemcpyAsync(buf + L.row_id, src + S.row_id, L.total - L.row_id, cudaMemcpyDeviceToHost, h->stream)); }
        else if (nc) {
            CK(h, cudaMemcpyAsync(buf + L.row_id, src + S.row_id, L.g - L.row_id, cudaMemcpyDeviceToHost, h->stream));
            if (nz) CK(h, cudaMemcpyAsync(buf + L.col, src + S.col, L.total - L.col, cudaMemcpyDeviceToHost, h->stream));
        }
    } else {                                                          // a non-finite cut truncated the round: section by section
        if (nc) {
            CK(h, cudaMemcpyAsync(buf + L.row_id, src + S.row_id, 8 * nc, cudaMemcpyDeviceToHost, h->stream));
            CK(h, cudaMemcpyAsync(buf + L.row_ptr, src + S.row_ptr, 8 * nc, cudaMemcpyDeviceToHost, h->stream));
            CK(h, cudaMemcpyAsync(buf + L.lo, src + S.lo, 8 * nc, cudaMemcpyDeviceToHost, h->stream));
            CK(h, cudaMemcpyAsync(buf + L.hi, src + S.hi, 8 * nc, cudaMemcpyDeviceToHost, h->stream));
            if (!lean) {
                CK(h, cudaMemcpyAsync(buf + L.g, src + S.g, 8 * nc, cudaMemcpyDeviceToHost, h->stream));
                CK(h, cudaMemcpyAsync(buf + L.viol, src + S.viol, 8 * nc, cudaMemcpyDeviceToHost, h->stream));
                CK(h, cudaMemcpyAsync(buf + L.b, src + S.b, 8 * nc, cudaMemcpyDeviceToHost, h->stream));
            }
        }
        if (nz) {
            CK(h, cudaMemcpyAsync(buf + L.col, src + S.col, 4 * nz, cudaMemcpyDeviceToHost, h->stream));
            CK(h, cudaMemcpyAsync(buf + L.val, src + S.val, 8 * nz, cudaMemcpyDeviceToHost, h->stream));
        }
    }
    if (h->opt.flags & KTN_FLAG_TIME_KERNELS) CK(h, cudaEventRecord(h->ev3, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    reinterpret_cast<int64_t*>(buf + L.row_ptr)[nc] = (int64_t)nz;   // the device array ends at the untruncated total
    float ms = 0.f; if ((h->opt.flags & KTN_FLAG_TIME_KERNELS) && cudaEventElapsedTime(&ms, h->ev2, h->ev3) == cudaSuccess) h->tm.d2h_ms = ms;
    out->n_cuts = (int64_t)nc; out->nnz = (int64_t)nz;
    out->row_id = reinterpret_cast<const int64_t*>(buf + L.row_id); out->row_ptr = reinterpret_cast<const int64_t*>(buf + L.row_ptr);
    out->col = reinterpret_cast<const int32_t*>(buf + L.col); out->val = reinterpret_cast<const double*>(buf + L.val);
    out->lo = reinterpret_cast<const double*>(buf + L.lo); out->hi = reinterpret_cast<const double*>(buf + L.hi);
    out->g = lean ? nullptr : reinterpret_cast<const double*>(buf + L.g); out->viol = lean ? nullptr : reinterpret_cast<const double*>(buf + L.viol);
    out->bconst = lean ? nullptr : reinterpret_cast<const double*>(buf + L.b);
    return KTN_OK;
}

extern "C" int ktn_get_g(ktn_handle* h, double* g_out) {
    if (h && !h->shards.empty()) {
        for (size_t s = 0; s < h->shards.size(); ++s) { int rc = ktn_get_g(h->shards[s], g_out + h->shard_begin[s]); if (rc) return group_fail(h, h->shards[s], rc); }
        return KTN_OK;
    }
    if (!h || !h->have_round) return fail(h, KTN_ERR_USAGE, "no round has run");
    cudaSetDevice(h->device);
    CK(h, cudaMemcpyAsync(g_out, h->g_row.p, 8 * (size_t)h->prob.num_constr, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return KTN_OK;
}

extern "C" int ktn_eval_g(ktn_handle* h, const double* x, double* g_out) {
    if (h && !h->shards.empty()) {
        for (size_t s = 0; s < h->shards.size(); ++s) { int rc = ktn_eval_g(h->shards[s], x, g_out + h->shard_begin[s]); if (rc) return group_fail(h, h->shards[s], rc); }
        return KTN_OK;
    }
    if (!h || !h->loaded) return fail(h, KTN_ERR_USAGE, "no problem loaded");
    cudaSetDevice(h->device);
    int rc = upload_x(h, x); if (rc) return rc;
    KtnRoundParams p = ktn_make_params(h, h->x.as<double>(), KTN_MODE_SEPARATE, 0);
    cudaError_t e = cudaSuccess;
    CK(h, cudaEventRecord(h->ev2, h->stream));
    h->tm.launches += ktn_launch_eval(p, make_plan(h), h->num_sms, h->max_smem, h->stream, &e);
    if (e != cudaSuccess) return fail(h, KTN_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    CK(h, cudaEventRecord(h->ev3, h->stream));
    CK(h, cudaMemcpyAsync(g_out, h->g_row.p, 8 * (size_t)h->prob.num_constr, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f; if (cudaEventElapsedTime(&ms, h->ev2, h->ev3) == cudaSuccess) h->tm.eval_ms = ms;      // the evaluation kernels alone
    return KTN_OK;
}

extern "C" int ktn_timings_get(ktn_handle* h, ktn_timings* out) {
    if (!h || !out) return KTN_ERR_USAGE;
    if (!h->shards.empty()) {      // the devices work side by side: times are the slowest device's, counts are sums
        ktn_timings t; memset(&t, 0, sizeof t);
        for (ktn_handle* s : h->shards) {
            ktn_timings q; int rc = ktn_timings_get(s, &q); if (rc) return rc;
            t.h2d_ms = q.h2d_ms > t.h2d_ms ? q.h2d_ms : t.h2d_ms; t.kernel_ms = q.kernel_ms > t.kernel_ms ? q.kernel_ms : t.kernel_ms;
            t.d2h_ms = q.d2h_ms > t.d2h_ms ? q.d2h_ms : t.d2h_ms; t.eval_ms = q.eval_ms > t.eval_ms ? q.eval_ms : t.eval_ms;
            t.compact_ms = q.compact_ms > t.compact_ms ? q.compact_ms : t.compact_ms; t.cut_ms = q.cut_ms > t.cut_ms ? q.cut_ms : t.cut_ms;
            t.eval_ms_sum = q.eval_ms_sum > t.eval_ms_sum ? q.eval_ms_sum : t.eval_ms_sum; t.compact_ms_sum = q.compact_ms_sum > t.compact_ms_sum ? q.compact_ms_sum : t.compact_ms_sum;
            t.cut_ms_sum = q.cut_ms_sum > t.cut_ms_sum ? q.cut_ms_sum : t.cut_ms_sum; t.rounds_timed = q.rounds_timed; t.rounds = q.rounds; t.launches += q.launches;
        }
        *out = t;
        return KTN_OK;
    }
    cudaSetDevice(h->device);
    drain_ring(h, false);
    h->tm.eval_ms_sum = h->eval_ms_sum; h->tm.compact_ms_sum = h->compact_ms_sum; h->tm.cut_ms_sum = h->cut_ms_sum;
    h->tm.exchange_ms_sum = h->exchange_ms_sum; h->tm.exchanges_timed = h->exchanges_timed; h->tm.rounds_timed = h->rounds_timed;
    *out = h->tm;
    return KTN_OK;
}

extern "C" int ktn_set_stream(ktn_handle* h, void* s) {
    if (!h) return KTN_ERR_USAGE;
    if (!h->shards.empty()) return fail(h, KTN_ERR_UNSUPPORTED, "not available on a multi-device handle");
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    h->stream = s ? (cudaStream_t)s : h->own_stream;
    return KTN_OK;
}

extern "C" int ktn_separate_device_async(ktn_handle* h, const double* d_x) {
    if (h && !h->shards.empty()) return fail(h, KTN_ERR_UNSUPPORTED, "not available on a multi-device handle");
    if (!h || !h->loaded) return fail(h, KTN_ERR_USAGE, "no problem loaded");
    cudaSetDevice(h->device);
    return enqueue_round(h, d_x, KTN_MODE_SEPARATE, 1);
}

extern "C" int ktn_sync_counts(ktn_handle* h, int64_t* n_cuts, int64_t* nnz, int64_t* err_row) {
    if (h && !h->shards.empty()) return fail(h, KTN_ERR_UNSUPPORTED, "not available on a multi-device handle");
    if (!h || !h->loaded) return fail(h, KTN_ERR_USAGE, "no problem loaded");
    cudaSetDevice(h->device);
    return finish_round(h, n_cuts, nnz, err_row);
}

extern "C" int64_t ktn_algorithmic_bytes(ktn_handle* h) {
    if (h && !h->shards.empty()) { int64_t b = 0; for (size_t s = 0; s < h->shards.size(); ++s) b += h->shards[s]->prob.alg_bytes_static + 12 * h->sh_nnz[s] + 28 * h->sh_cuts[s] - (s ? 8 * h->g_num_var : 0); return b; }      // SURVEY 8d counts x* once
    if (!h || !h->loaded) return 0;
    return h->prob.alg_bytes_static + 12 * h->nnz_cuts + 28 * h->n_cuts;
}

// ---------------------------------------------------------------------------------------------
// Single-process sharded operation (ktn_options.ngpus > 1).  The reference owns ONE separator in ONE
// process (src/Katana.jl:18, the loop src/model.jl:257-283): here one FRONT handle drives one handle per device.  Rows are split
// into contiguous ranges; a round is enqueued on every device before any is waited for; the combined batch is the rank-major
// concatenation (= ascending row order), cut at the first non-finite row as the reference does (src/model.jl:69-73, :278); every
// device downloads its share over its own PCIe link straight into its place in ONE pinned buffer.
// ---------------------------------------------------------------------------------------------
#include <thread>
void ktn_launch_shift(const int64_t* in, int64_t* out, int64_t n, int64_t add, cudaStream_t stream);

static int group_fail(ktn_handle* f, ktn_handle* s, int rc) { if (f && s) f->err = s->err; return rc; }

static int group_create(ktn_handle* f, ktn_handle** out) {
    const int n = f->opt.ngpus;
    if (n > KTN_MAX_GPUS) { delete f; return KTN_ERR_USAGE; }
    if (f->opt.topk > 0) { fprintf(stderr, "libktn: topk is not available on a multi-device handle\n"); delete f; return KTN_ERR_UNSUPPORTED; }
    for (int s = 0; s < n; ++s) {
        ktn_options o = f->opt;
        o.struct_size = (int32_t)sizeof(ktn_options); o.ngpus = 0; o.flags &= ~KTN_FLAG_DIRECT_VIEW; o.device = f->opt.devices[s] >= 0 ? f->opt.devices[s] : s;
        ktn_handle* sh = nullptr;
        const int rc = ktn_create(&o, &sh);
        if (rc != KTN_OK) { for (ktn_handle* q : f->shards) ktn_destroy(q); f->shards.clear(); delete f; return rc; }
        f->shards.push_back(sh);
    }
    f->shard_begin.assign((size_t)n + 1, 0); f->sh_cuts.assign((size_t)n, 0); f->sh_nnz.assign((size_t)n, 0);
    f->device = f->shards[0]->device;
    bool one_device = true;
    for (ktn_handle* q : f->shards) one_device = one_device && q->device == f->device;
    if (one_device && (f->opt.flags & KTN_FLAG_EAGER_VIEW) && n <= KTN_HP_MAX_PREV && !(getenv("KTN_HOSTPUSH") && !strcmp(getenv("KTN_HOSTPUSH"), "0"))) {
        // one device, several shards: one stream for all of them (they run back to back anyway), a second one for the host pushes
        cudaSetDevice(f->device);
        for (ktn_handle* q : f->shards) q->stream = f->shards[0]->own_stream;
        bool ok = cudaStreamCreateWithFlags(&f->g_copy_stream, cudaStreamNonBlocking) == cudaSuccess;
        f->g_done.assign((size_t)n, nullptr);
        for (int s = 0; ok && s < n; ++s) ok = cudaEventCreateWithFlags(&f->g_done[s], cudaEventDisableTiming) == cudaSuccess;
        void* q = nullptr; void* d = nullptr;
        ok = ok && cudaHostAlloc(&q, 32 * (size_t)n, cudaHostAllocMapped) == cudaSuccess && cudaHostGetDevicePointer(&d, q, 0) == cudaSuccess;
        if (!ok) { group_destroy(f); return KTN_ERR_CUDA; }
        f->g_hdr = static_cast<unsigned long long*>(q); f->g_hdr_d = static_cast<unsigned long long*>(d);
        if (getenv("KTN_HOSTPUSH_GBS")) { const double r = atof(getenv("KTN_HOSTPUSH_GBS")); if (r >= 0.0 && r < 1000.0) f->push_pace = (float)r; }
        if (getenv("KTN_HOSTPUSH_BLOCKS")) { const int b = atoi(getenv("KTN_HOSTPUSH_BLOCKS")); if (b >= 1 && b <= 64) f->push_blocks = b; }
        f->push_view = true;
    }
    *out = f;
    return KTN_OK;
}

static void group_destroy(ktn_handle* f) {
    if (f->push_view) { cudaSetDevice(f->device); cudaDeviceSynchronize(); for (ktn_handle* s : f->shards) s->stream = s->own_stream; }
    for (ktn_handle* s : f->shards) ktn_destroy(s);
    f->shards.clear();
    cudaSetDevice(f->device);
    if (f->g_hx) cudaFreeHost(f->g_hx);
    for (int i = 0; i < 2; ++i) { if (f->h_view[i]) cudaFreeHost(f->h_view[i]); if (f->g_eager[i]) cudaFreeHost(f->g_eager[i]); }
    if (f->g_hdr) cudaFreeHost(f->g_hdr);
    for (cudaEvent_t e : f->g_done) if (e) cudaEventDestroy(e);
    if (f->g_copy_stream) cudaStreamDestroy(f->g_copy_stream);
    delete f;
}

static int group_load_begin(ktn_handle* f, int64_t num_var, int64_t num_constr) {
    const int64_t n = (int64_t)f->shards.size();
    f->g_num_var = num_var; f->g_num_constr = num_constr; f->loaded = false; f->have_round = false;
    for (int64_t s = 0; s <= n; ++s) f->shard_begin[s] = num_constr * s / n;      // contiguous, equal row counts
    for (int64_t s = 0; s < n; ++s) {
        int rc = ktn_load_begin(f->shards[s], num_var, f->shard_begin[s + 1] - f->shard_begin[s]); if (rc) return group_fail(f, f->shards[s], rc);
        f->shards[s]->row_offset = f->shard_begin[s];
    }
    cudaSetDevice(f->device);
    if (f->g_hx) { cudaFreeHost(f->g_hx); f->g_hx = nullptr; }
    for (int i = 0; i < 2; ++i) if (f->g_eager[i]) { cudaFreeHost(f->g_eager[i]); f->g_eager[i] = nullptr; }
    f->eager_valid = false; f->eager_cap_cuts = f->eager_cap_nnz = 0;
    CK(f, cudaHostAlloc(&f->g_hx, 8 * ((size_t)num_var + 1), cudaHostAllocPortable));
    f->loading = true;
    return KTN_OK;
}

static int group_add_rows(ktn_handle* f, int64_t first_row, int64_t nrows, const int64_t* eptr, const int32_t* op, const int32_t* arg, const double* val,
                          const double* lb, const double* ub, const uint8_t* flags) {
    if (!f->loading) return fail(f, KTN_ERR_USAGE, "ktn_add_rows before ktn_load_begin");
    std::vector<int64_t> ep;
    for (size_t s = 0; s < f->shards.size(); ++s) {
        const int64_t r0 = first_row > f->shard_begin[s] ? first_row : f->shard_begin[s];
        const int64_t r1 = first_row + nrows < f->shard_begin[s + 1] ? first_row + nrows : f->shard_begin[s + 1];
        if (r0 >= r1) continue;
        const int64_t a = r0 - first_row, cnt = r1 - r0, base = eptr[a];
        ep.resize((size_t)cnt + 1);
        for (int64_t i = 0; i <= cnt; ++i) ep[i] = eptr[a + i] - base;
        int rc = ktn_add_rows(f->shards[s], r0 - f->shard_begin[s], cnt, ep.data(), op + base, arg + base, val + base, lb + a, ub + a, flags + a);
        if (rc) return group_fail(f, f->shards[s], rc);
    }
    return KTN_OK;
}

static int group_load_end(ktn_handle* f) {
    if (!f->loading) return fail(f, KTN_ERR_USAGE, "ktn_load_end before ktn_load_begin");
    const size_t n = f->shards.size();
    std::vector<int> rc(n, KTN_OK);
    std::vector<std::thread> th;
    for (size_t s = 0; s < n; ++s) th.emplace_back([&, s] { rc[s] = ktn_load_end(f->shards[s]); });      // the tape compiler is host work: one thread per shard
    for (std::thread& t : th) t.join();
    for (size_t s = 0; s < n; ++s) if (rc[s]) return group_fail(f, f->shards[s], rc[s]);
    f->loading = false; f->loaded = true;
    if (f->opt.flags & KTN_FLAG_EAGER_VIEW) {       // worst-case layout of the combined batch: every nonlinear row cut
        int64_t cc = 0, zz = 0;
        for (ktn_handle* s : f->shards) for (int64_t i = 0; i < s->prob.num_constr; ++i) if (s->prob.flags[i] & KTN_ROW_NL) { ++cc; zz += s->prob.jac_ptr[i + 1] - s->prob.jac_ptr[i]; }
        const size_t total = (size_t)ktn_pack_layout((unsigned long long)cc, (unsigned long long)zz).total;
        if (total <= ((size_t)1 << 30)) {
            cudaSetDevice(f->device);
            for (int i = 0; i < 2; ++i) {
                CK(f, cudaHostAlloc(&f->g_eager[i], total + 64, cudaHostAllocPortable | (f->push_view ? cudaHostAllocMapped : 0)));
                if (f->push_view) { void* d = nullptr; CK(f, cudaHostGetDevicePointer(&d, f->g_eager[i], 0)); f->g_eager_d[i] = static_cast<unsigned char*>(d); }
            }
            f->eager_cap_cuts = cc; f->eager_cap_nnz = zz;
        }
    }
    return KTN_OK;
}

// Separation round of a pipelined single-device handle: x* goes up once; the shards' rounds are enqueued back to back on one stream;
// behind every shard a small kernel on a second stream stores that shard's cuts into the combined batch in mapped pinned memory (the
// PCIe transfer of shard s runs beside the kernels of shard s + 1, whose K1 leaves the push kernel's SMs free).  The host enqueues
// everything without waiting and synchronises ONCE.
void ktn_launch_hostpush(const KtnHostPushParams& q, int blocks, cudaStream_t stream);
static int group_round_pushed(ktn_handle* f, int do_round, int64_t* n_cuts, int64_t* nnz, int64_t* err_row) {
    const size_t n = f->shards.size();
    ktn_handle* h0 = f->shards[0];
    cudaSetDevice(f->device);
    f->eager_valid = false;
    f->eager_cur ^= 1;
    const bool lean = (f->opt.flags & KTN_FLAG_LEAN_VIEW) != 0;
    const KtnPackLayout EL = ktn_pack_layout((unsigned long long)f->eager_cap_cuts, (unsigned long long)f->eager_cap_nnz);
    static const bool trace = getenv("KTN_HOSTPUSH_TRACE") != nullptr;       // stderr: when every shard's kernels and push ended (microseconds after the upload began)
    static std::vector<cudaEvent_t> tev;
    if (trace && tev.size() < 3 * n + 1) { tev.resize(3 * n + 1); for (cudaEvent_t& e : tev) cudaEventCreate(&e); }
    if (trace) cudaEventRecord(tev[3 * n], h0->stream);
    CK(f, cudaMemcpyAsync(h0->x.p, f->g_hx, 8 * (size_t)f->g_num_var, cudaMemcpyHostToDevice, h0->stream));
    for (size_t s = 0; s < n; ++s) {
        ktn_handle* h = f->shards[s];
        h->reserve_sms = s ? f->push_blocks : 0;
        int rc = enqueue_round(h, h0->x.as<double>(), KTN_MODE_SEPARATE, do_round); if (rc) return group_fail(f, h, rc);
        if (trace) cudaEventRecord(tev[3 * s], h->stream);
        CK(f, cudaEventRecord(f->g_done[s], h->stream));
        CK(f, cudaStreamWaitEvent(f->g_copy_stream, f->g_done[s], 0));
        if (trace) cudaEventRecord(tev[3 * s + 1], f->g_copy_stream);
        KtnHostPushParams q; memset(&q, 0, sizeof q);
        q.src = h->out_blob[h->out_cur].as<unsigned char>(); q.counts = h->counts.as<unsigned long long>();
        for (size_t k = 0; k < s; ++k) q.prev[k] = f->shards[k]->counts.as<unsigned long long>();
        q.pace = f->push_pace; q.nprev = (int)s; q.lean = lean ? 1 : 0; q.dst = f->g_eager_d[f->eager_cur]; q.EL = EL; q.hdr = f->g_hdr_d + 4 * s;
        ktn_launch_hostpush(q, f->push_blocks, f->g_copy_stream);
        CK(f, cudaGetLastError());
        h->tm.launches += 1;
        if (trace) cudaEventRecord(tev[3 * s + 2], f->g_copy_stream);
    }
    CK(f, cudaStreamSynchronize(f->g_copy_stream));
    if (trace) {
        for (size_t s = 0; s < n; ++s) {
            float a = 0, b = 0, c = 0;
            cudaEventElapsedTime(&a, tev[3 * n], tev[3 * s]); cudaEventElapsedTime(&b, tev[3 * n], tev[3 * s + 1]); cudaEventElapsedTime(&c, tev[3 * n], tev[3 * s + 2]);
            fprintf(stderr, "libktn hostpush: shard %zu kernels done %.0f us, push %.0f .. %.0f us (%llu cuts)\n", s, 1e3 * a, 1e3 * b, 1e3 * c, (unsigned long long)f->g_hdr[4 * s]);
        }
    }
    int64_t tc = 0, tz = 0, er = -1;
    for (size_t s = 0; s < n; ++s) {
        const unsigned long long* hd = f->g_hdr + 4 * s;
        f->sh_cuts[s] = (int64_t)hd[0]; f->sh_nnz[s] = (int64_t)hd[1]; tc += f->sh_cuts[s]; tz += f->sh_nnz[s];
        if (er < 0 && hd[2] != ~0ull) er = (int64_t)hd[2] - 1 + f->shard_begin[s];
    }
    f->eager_valid = true;
    f->n_cuts = tc; f->nnz_cuts = tz; f->err_row = er; f->have_round = true;
    if (n_cuts) *n_cuts = tc;
    if (nnz) *nnz = tz;
    if (err_row) *err_row = er;
    return er >= 0 ? KTN_NUMERIC_NONFINITE : KTN_OK;
}

// One round on every device: rows == nullptr: separation (violated rows); else unconditional cuts of the listed rows.
static int group_round(ktn_handle* f, const double* x, const int64_t* rows, int64_t nrows, int do_round, int64_t* n_cuts, int64_t* nnz, int64_t* err_row) {
    if (!f->loaded) return fail(f, KTN_ERR_USAGE, "no problem loaded");
    const size_t n = f->shards.size();
    memcpy(f->g_hx, x, 8 * (size_t)f->g_num_var);
    if (f->push_view && f->g_eager[0] && !rows) return group_round_pushed(f, do_round, n_cuts, nnz, err_row);
    for (ktn_handle* h : f->shards) h->reserve_sms = 0;
    std::vector<uint8_t> mask;
    for (size_t s = 0; s < n; ++s) {       // enqueue everywhere first: the devices run side by side
        ktn_handle* h = f->shards[s];
        cudaSetDevice(h->device);
        if (rows) {
            const int64_t m = h->prob.num_constr;
            mask.assign((size_t)m + 1, 0);
            for (int64_t j = 0; j < nrows; ++j) if (rows[j] >= f->shard_begin[s] && rows[j] < f->shard_begin[s + 1]) mask[rows[j] - f->shard_begin[s]] = 1;
            CK(f, cudaStreamSynchronize(h->stream));
            CK(f, cudaMemcpy(h->force.p, mask.data(), (size_t)m, cudaMemcpyHostToDevice));
        }
        if (f->opt.flags & KTN_FLAG_TIME_KERNELS) CK(f, cudaEventRecord(h->ev0, h->stream));
        CK(f, cudaMemcpyAsync(h->x.p, f->g_hx, 8 * (size_t)f->g_num_var, cudaMemcpyHostToDevice, h->stream));
        if (f->opt.flags & KTN_FLAG_TIME_KERNELS) CK(f, cudaEventRecord(h->ev1, h->stream));
        int rc = enqueue_round(h, h->x.as<double>(), rows ? KTN_MODE_FORCE : KTN_MODE_SEPARATE, do_round); if (rc) return group_fail(f, h, rc);
    }
    int64_t tc = 0, tz = 0, er = -1; bool stopped = false;
    const bool eager = f->g_eager[0] != nullptr && !rows;      // separation rounds only: the unconditional rounds may select rows outside nlconstr_ixs
    const bool lean = (f->opt.flags & KTN_FLAG_LEAN_VIEW) != 0;
    unsigned char* ebuf = nullptr; KtnPackLayout EL;
    if (eager) { f->eager_cur ^= 1; ebuf = f->g_eager[f->eager_cur]; EL = ktn_pack_layout((unsigned long long)f->eager_cap_cuts, (unsigned long long)f->eager_cap_nnz); }
    f->eager_valid = false;
    for (size_t s = 0; s < n; ++s) {
        ktn_handle* h = f->shards[s];
        cudaSetDevice(h->device);
        int64_t c = 0, z = 0, e = -1;
        int rc = finish_round(h, &c, &z, &e); if (rc < 0) return group_fail(f, h, rc);
        if (eager && !stopped && c > 0) {      // this shard's cuts go to their final place now; the later shards are still computing
            const KtnPackLayout S = ktn_pack_layout((unsigned long long)h->lay_cuts, (unsigned long long)h->lay_nnz);
            const unsigned char* src = h->out_blob[h->out_cur].as<unsigned char>();
            const size_t cb = (size_t)tc, zb = (size_t)tz, cs = (size_t)c, zs = (size_t)z;
            if (h->rp_shift.bytes < 8 * (cs + 1)) CK(f, h->rp_shift.alloc(8 * (cs + 1) + 8 * (cs + 1) / 4));
            ktn_launch_shift(reinterpret_cast<const int64_t*>(src + S.row_ptr), h->rp_shift.as<int64_t>(), (int64_t)cs, (int64_t)zb, h->stream);
            CK(f, cudaMemcpyAsync(ebuf + EL.row_id + 8 * cb, src + S.row_id, 8 * cs, cudaMemcpyDeviceToHost, h->stream));
            CK(f, cudaMemcpyAsync(ebuf + EL.row_ptr + 8 * cb, h->rp_shift.p, 8 * cs, cudaMemcpyDeviceToHost, h->stream));
            CK(f, cudaMemcpyAsync(ebuf + EL.lo + 8 * cb, src + S.lo, 8 * cs, cudaMemcpyDeviceToHost, h->stream));
            CK(f, cudaMemcpyAsync(ebuf + EL.hi + 8 * cb, src + S.hi, 8 * cs, cudaMemcpyDeviceToHost, h->stream));
            if (!lean) {
                CK(f, cudaMemcpyAsync(ebuf + EL.g + 8 * cb, src + S.g, 8 * cs, cudaMemcpyDeviceToHost, h->stream));
                CK(f, cudaMemcpyAsync(ebuf + EL.viol + 8 * cb, src + S.viol, 8 * cs, cudaMemcpyDeviceToHost, h->stream));
                CK(f, cudaMemcpyAsync(ebuf + EL.b + 8 * cb, src + S.b, 8 * cs, cudaMemcpyDeviceToHost, h->stream));
            }
            if (zs) {
                CK(f, cudaMemcpyAsync(ebuf + EL.col + 4 * zb, src + S.col, 4 * zs, cudaMemcpyDeviceToHost, h->stream));
                CK(f, cudaMemcpyAsync(ebuf + EL.val + 8 * zb, src + S.val, 8 * zs, cudaMemcpyDeviceToHost, h->stream));
            }
        }
        float ms = 0.f; if ((h->opt.flags & KTN_FLAG_TIME_KERNELS) && cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->tm.h2d_ms = ms;
        if (stopped) { c = 0; z = 0; }                         // the reference never reaches the rows behind the first non-finite cut
        f->sh_cuts[s] = c; f->sh_nnz[s] = z; tc += c; tz += z;
        if (!stopped && e >= 0) { er = e + f->shard_begin[s]; stopped = true; }
    }
    f->eager_valid = eager;
    f->n_cuts = tc; f->nnz_cuts = tz; f->err_row = er; f->have_round = true;
    if (n_cuts) *n_cuts = tc;
    if (nnz) *nnz = tz;
    if (err_row) *err_row = er;
    return er >= 0 ? KTN_NUMERIC_NONFINITE : KTN_OK;
}

// The combined batch of the last round, into a view (library-owned pinned memory) or into caller buffers.
static int group_fetch(ktn_handle* f, ktn_cut_view* view, int64_t* row_id, int64_t* row_ptr, int32_t* col, double* val, double* lo, double* hi, double* g, double* viol, double* bconst) {
    if (!f->have_round) return fail(f, KTN_ERR_USAGE, "no round has run");
    const size_t nc = (size_t)f->n_cuts, nz = (size_t)f->nnz_cuts;
    const bool lean = view && (f->opt.flags & KTN_FLAG_LEAN_VIEW) != 0;
    if (f->eager_valid) {       // the copies were started by ktn_separate, shard by shard: wait for them
        if (!f->push_view) for (ktn_handle* h : f->shards) { cudaSetDevice(h->device); CK(f, cudaStreamSynchronize(h->stream)); }
        unsigned char* buf = f->g_eager[f->eager_cur];
        const KtnPackLayout EL = ktn_pack_layout((unsigned long long)f->eager_cap_cuts, (unsigned long long)f->eager_cap_nnz);
        const bool lv = (f->opt.flags & KTN_FLAG_LEAN_VIEW) != 0;
        reinterpret_cast<int64_t*>(buf + EL.row_ptr)[nc] = (int64_t)nz;
        if (view) {
            view->n_cuts = (int64_t)nc; view->nnz = (int64_t)nz;
            view->row_id = reinterpret_cast<const int64_t*>(buf + EL.row_id); view->row_ptr = reinterpret_cast<const int64_t*>(buf + EL.row_ptr);
            view->col = reinterpret_cast<const int32_t*>(buf + EL.col); view->val = reinterpret_cast<const double*>(buf + EL.val);
            view->lo = reinterpret_cast<const double*>(buf + EL.lo); view->hi = reinterpret_cast<const double*>(buf + EL.hi);
            view->g = lv ? nullptr : reinterpret_cast<const double*>(buf + EL.g); view->viol = lv ? nullptr : reinterpret_cast<const double*>(buf + EL.viol);
            view->bconst = lv ? nullptr : reinterpret_cast<const double*>(buf + EL.b);
            return KTN_OK;
        }
        if (!lv) {      // caller buffers: from the pinned copy (a lean handle has no g / viol / bconst there: fall through to the device copies)
            if (row_id) memcpy(row_id, buf + EL.row_id, 8 * nc);
            if (row_ptr) memcpy(row_ptr, buf + EL.row_ptr, 8 * (nc + 1));
            if (col) memcpy(col, buf + EL.col, 4 * nz);
            if (val) memcpy(val, buf + EL.val, 8 * nz);
            if (lo) memcpy(lo, buf + EL.lo, 8 * nc);
            if (hi) memcpy(hi, buf + EL.hi, 8 * nc);
            if (g) memcpy(g, buf + EL.g, 8 * nc);
            if (viol) memcpy(viol, buf + EL.viol, 8 * nc);
            if (bconst) memcpy(bconst, buf + EL.b, 8 * nc);
            return KTN_OK;
        }
    }
    for (ktn_handle* h : f->shards) if (h->round_pending) { cudaSetDevice(h->device); int rc = finish_round(h, nullptr, nullptr, nullptr); if (rc < 0) return group_fail(f, h, rc); }
    if (view) {
        const KtnPackLayout L = ktn_pack_layout(nc, nz);
        f->view_cur ^= 1;
        unsigned char*& buf = f->h_view[f->view_cur]; size_t& cap = f->h_view_cap[f->view_cur];
        if (cap < L.total) {
            cudaSetDevice(f->device);
            if (buf) cudaFreeHost(buf);
            buf = nullptr; cap = 0;
            const size_t want = L.total + L.total / 4 + 4096;
            CK(f, cudaHostAlloc(&buf, want, cudaHostAllocPortable));
            cap = want;
        }
        row_id = reinterpret_cast<int64_t*>(buf + L.row_id); row_ptr = reinterpret_cast<int64_t*>(buf + L.row_ptr);
        col = reinterpret_cast<int32_t*>(buf + L.col); val = reinterpret_cast<double*>(buf + L.val);
        lo = reinterpret_cast<double*>(buf + L.lo); hi = reinterpret_cast<double*>(buf + L.hi);
        g = lean ? nullptr : reinterpret_cast<double*>(buf + L.g); viol = lean ? nullptr : reinterpret_cast<double*>(buf + L.viol);
        bconst = lean ? nullptr : reinterpret_cast<double*>(buf + L.b);
    }
    size_t cb = 0, zb = 0;
    for (size_t s = 0; s < f->shards.size(); ++s) {       // every device copies its share to its place; all copies are in flight together
        ktn_handle* h = f->shards[s];
        const size_t c = (size_t)f->sh_cuts[s], z = (size_t)f->sh_nnz[s];
        if (c == 0) continue;
        cudaSetDevice(h->device);
        const KtnPackLayout S = ktn_pack_layout((unsigned long long)h->lay_cuts, (unsigned long long)h->lay_nnz);
        const unsigned char* src = h->out_blob[h->out_cur].as<unsigned char>();
        if (f->opt.flags & KTN_FLAG_TIME_KERNELS) CK(f, cudaEventRecord(h->ev2, h->stream));
        if (row_id) CK(f, cudaMemcpyAsync(row_id + cb, src + S.row_id, 8 * c, cudaMemcpyDeviceToHost, h->stream));
        if (row_ptr) {     // the shard's entry offsets start at 0: shifted on the device to the combined batch's
            if (h->rp_shift.bytes < 8 * (c + 1)) CK(f, h->rp_shift.alloc(8 * (c + 1) + 8 * (c + 1) / 4));
            ktn_launch_shift(reinterpret_cast<const int64_t*>(src + S.row_ptr), h->rp_shift.as<int64_t>(), (int64_t)c, (int64_t)zb, h->stream);
            CK(f, cudaMemcpyAsync(row_ptr + cb, h->rp_shift.p, 8 * c, cudaMemcpyDeviceToHost, h->stream));
        }
        if (lo) CK(f, cudaMemcpyAsync(lo + cb, src + S.lo, 8 * c, cudaMemcpyDeviceToHost, h->stream));
        if (hi) CK(f, cudaMemcpyAsync(hi + cb, src + S.hi, 8 * c, cudaMemcpyDeviceToHost, h->stream));
        if (g) CK(f, cudaMemcpyAsync(g + cb, src + S.g, 8 * c, cudaMemcpyDeviceToHost, h->stream));
        if (viol) CK(f, cudaMemcpyAsync(viol + cb, src + S.viol, 8 * c, cudaMemcpyDeviceToHost, h->stream));
        if (bconst) CK(f, cudaMemcpyAsync(bconst + cb, src + S.b, 8 * c, cudaMemcpyDeviceToHost, h->stream));
        if (col && z) CK(f, cudaMemcpyAsync(col + zb, src + S.col, 4 * z, cudaMemcpyDeviceToHost, h->stream));
        if (val && z) CK(f, cudaMemcpyAsync(val + zb, src + S.val, 8 * z, cudaMemcpyDeviceToHost, h->stream));
        if (f->opt.flags & KTN_FLAG_TIME_KERNELS) CK(f, cudaEventRecord(h->ev3, h->stream));
        cb += c; zb += z;
    }
    for (size_t s = 0; s < f->shards.size(); ++s) {
        ktn_handle* h = f->shards[s];
        if (f->sh_cuts[s] == 0) continue;
        cudaSetDevice(h->device);
        CK(f, cudaStreamSynchronize(h->stream));
        float ms = 0.f; if ((h->opt.flags & KTN_FLAG_TIME_KERNELS) && cudaEventElapsedTime(&ms, h->ev2, h->ev3) == cudaSuccess) h->tm.d2h_ms = ms;
    }
    if (row_ptr) row_ptr[nc] = (int64_t)nz;
    if (view) {
        view->n_cuts = (int64_t)nc; view->nnz = (int64_t)nz; view->row_id = row_id; view->row_ptr = row_ptr; view->col = col; view->val = val;
        view->lo = lo; view->hi = hi; view->g = g; view->viol = viol; view->bconst = bconst;
    }
    return KTN_OK;
}
