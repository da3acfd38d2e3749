// ktn_emu.cpp -- TEST-ONLY host emulator of the CUDA round kernels.
// Runs the SAME compiled artefacts (shape programs, packed chunk blobs, sort orders) that the
// sm_100a kernels consume, lane by lane on the CPU, through the same interpreter core
// (csrc/ktn_interp.h).  It lets the CPU test-suite check the tape compiler and the chunk packing
// against the oracle without a GPU.  It is not part of the product and is never shipped in libktn.so.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/ktn.h"
#include "../../katana.jl_b200/csrc/ktn_compile.h"
#include "../../katana.jl_b200/csrc/ktn_interp.h"
#include "../../katana.jl_b200/csrc/ktn_family.h"

struct ktn_handle {
    ktn_options opt; KtnProblem prob; bool loaded = false, have_round = false;
    std::vector<double> g_row, b_row, aux_row, stage_val, x_last; std::vector<uint32_t> sel; int do_round = 1;
    std::vector<int64_t> c_row, c_ptr; std::vector<int32_t> c_col; std::vector<double> c_val, c_lo, c_hi, c_g, c_viol, c_b;
    int64_t err_row = -1; std::string err;
};
extern "C" {
const char* ktn_backend(void) { return "emu"; }
const char* ktn_last_error(ktn_handle* h) { return h ? h->err.c_str() : "null"; }
int ktn_create(const ktn_options* o, ktn_handle** out) { ktn_handle* h = new ktn_handle(); h->opt = *o; *out = h; return 0; }
void ktn_destroy(ktn_handle* h) { delete h; }
int ktn_set_params(ktn_handle* h, double f, double r, int64_t k) { h->opt.f_tol = f; h->opt.cut_coef_rng = r; h->opt.topk = k; return 0; }
int ktn_load_begin(ktn_handle* h, int64_t n, int64_t m) { h->prob.reset(n, m); h->loaded = false; return 0; }
int ktn_add_rows(ktn_handle* h, int64_t first, int64_t nrows, const int64_t* ep, const int32_t* op, const int32_t* arg, const double* val,
                 const double* lb, const double* ub, const uint8_t* fl) {
    int rc = h->prob.add_rows(first, nrows, ep, op, arg, val, lb, ub, fl); if (rc) h->err = h->prob.err; return rc; }
int ktn_load_end(ktn_handle* h) {
    int64_t sigma = 8192; if (const char* s = getenv("KTN_SIGMA")) sigma = atoll(s);
    uint32_t limit = 1536; if (const char* s = getenv("KTN_EMU_LANE_LIMIT")) limit = (uint32_t)atoi(s);
    int rc = h->prob.finalize(sigma, limit); if (rc) { h->err = h->prob.err; return rc; }
    size_t m = (size_t)h->prob.num_constr;
    h->g_row.assign(m, 0.0); h->b_row.assign(m, 0.0); h->aux_row.assign(m, 0.0); h->sel.assign(m, 0u); h->stage_val.assign((size_t)h->prob.jac_ptr[m], 0.0);
    h->loaded = true; return 0; }
int ktn_set_bounds(ktn_handle* h, const double* lb, const double* ub) { h->prob.lb.assign(lb, lb + h->prob.num_constr); h->prob.ub.assign(ub, ub + h->prob.num_constr); h->prob.repack_bounds(); return 0; }
int64_t ktn_num_rows(ktn_handle* h) { return h->prob.rows_loaded; }
int64_t ktn_jac_nnz(ktn_handle* h) { return h->prob.jac_ptr.back(); }
int ktn_jac_structure(ktn_handle* h, int64_t* rp, int32_t* cols) {
    if (rp) memcpy(rp, h->prob.jac_ptr.data(), 8 * h->prob.jac_ptr.size());
    if (cols) memcpy(cols, h->prob.jac_col.data(), 4 * h->prob.jac_col.size()); return 0; }
}

// family shapes (ktn_family.h): the same row functions the sm_100a kernels run, on the same packed chunk
struct EmuFamRow {
    const double* C; const int32_t* cols; const uint8_t* rk; const double* X; uint32_t nu, L, lane;
    double cst(uint32_t i) const { return C[(size_t)i * L + lane]; }
    int32_t col(uint32_t u) const { return cols[(size_t)u * L + lane]; }
    double xat(int32_t c) const { return X[c]; }
    double x(uint32_t u) const { return X[col(u)]; }
    uint32_t rank(uint32_t u) const { return rk[(size_t)u * L + lane]; }
    // class >= 1 blobs (grouped layout, ktn_program.h): C = the blob
    uint64_t rankword() const { return ((const uint64_t*)((const uint8_t*)C + KTN_FAM_ORD_OFF(nu)))[lane]; }
    void pairs2(uint32_t g, double& a0, double& a1, double& b0, double& b1) const { const double* q = (const double*)((const uint8_t*)C + g * 1024u + lane * 32u); a0 = q[0]; a1 = q[1]; b0 = q[2]; b1 = q[3]; }
    void cols8(uint32_t g, int32_t (&c)[8]) const { memcpy(c, (const uint8_t*)C + KTN_FAM_COL_OFF(nu) + g * 1024u + lane * 32u, 32); }
    void pair(uint32_t u, double& p0, double& p1) const { const double* q = (const double*)((const uint8_t*)C + KTN_FAM_PAIR_AT(u, lane)); p0 = q[0]; p1 = q[1]; }
    int32_t gcol(uint32_t u) const { return *(const int32_t*)((const uint8_t*)C + KTN_FAM_COL_AT(nu, u, lane)); }
};
struct EmuFamSink {
    double* out; const int32_t* scol; const double* X;
    void put_j(uint32_t q, double v) { out[q] = v; } double get_j(uint32_t q) const { return out[q]; }
    double xsorted(uint32_t q) const { return X[scol[q]]; }
};
struct EmuCutSink {     // the cut kernel's sink: coefficients and columns of the row (entry order) and the products -x* J
    double* val; const int32_t* scol; bool* colmismatch; double t[KTN_FAM_REGS];
    void put(uint32_t q, double v, int32_t c) { val[q] = v; if (scol[q] != c) *colmismatch = true; } double get(uint32_t q) const { return val[q]; } void set(uint32_t q, double v) { val[q] = v; }
    void put_t(uint32_t q, double v) { t[q] = v; } double get_t(uint32_t q) const { return t[q]; }
};
static EmuFamRow emu_row(const KtnProblem& P, const KtnChunkDesc& cd, uint32_t lane, const double* x) {
    const uint32_t nu = (uint32_t)cd.aux;
    const uint8_t* blob = P.blob.data() + cd.blob_off;
    const size_t sec_col = (size_t)256 * KTN_FAM_NCONST(P.shapes[cd.shape].family, nu), sec_rk = sec_col + (size_t)128 * nu;   // long rows: constants | columns | rank bytes
    return EmuFamRow{(const double*)blob, (const int32_t*)(blob + sec_col), blob + sec_rk, x, nu, cd.stride, lane};
}
// K1: forward, test, record {g, aux}; rows of more than KTN_FAM_REGS variables build their cut here (streaming fallback)
template <int FAM, int N>
static void run_family_row(ktn_handle* h, const KtnChunkDesc& cd, uint32_t lane, int32_t row, const double* x, int mode, bool forced, double lb, double ub, int do_round) {
    KtnProblem& P = h->prob;
    const EmuFamRow r = emu_row(P, cd, lane, x);
    const uint32_t nu = r.nu;
    constexpr int NR = N > 0 ? N : 1;
    double aux, g;
    if constexpr (N > 0) {
        typedef KtnFamily<FAM> F;
        KtnFamRegs<NR> v;
        for (int u = 0; u < N; ++u) { r.pair(u, v.p0[u], v.p1[u]); v.x[u] = x[r.gcol(u)]; }
        g = F::template forward<NR>(v, aux);
    } else g = KtnFamily<FAM>::forward_stream(r, aux);
    h->g_row[row] = g;
    if (mode == 2) return;
    const bool selected = mode == 1 ? forced : !((g >= lb - h->opt.f_tol) && (g <= ub + h->opt.f_tol));
    if (!selected) { h->sel[row] = 0; return; }
    if constexpr (N > 0) { h->aux_row[row] = aux; h->sel[row] = nu | KTN_SEL_DEFER; }
    else {
        const int64_t base = P.jac_ptr[row];
        EmuFamSink s{h->stage_val.data() + base, P.jac_col.data() + base, x};
        double b;
        const bool bad = ktn_family_cut_stream<FAM>(r, s, g, aux, do_round != 0, h->opt.cut_coef_rng, b);
        h->b_row[row] = b; h->sel[row] = nu | (bad ? KTN_SEL_ERRBIT : 0u);
    }
}
template <int FAM, class... A> static void run_family_dispatch(uint32_t nu, A... a) {
    switch (ktn_family_class(nu)) {
#define KTN_CASE(n) case n: run_family_row<FAM, n>(a...); break;
        KTN_CASE(1) KTN_CASE(2) KTN_CASE(3) KTN_CASE(4) KTN_CASE(5) KTN_CASE(6) KTN_CASE(7) KTN_CASE(8)
        KTN_CASE(9) KTN_CASE(10) KTN_CASE(11) KTN_CASE(12) KTN_CASE(13) KTN_CASE(14) KTN_CASE(15) KTN_CASE(16)
#undef KTN_CASE
        default: run_family_row<FAM, 0>(a...); break;
    }
}
// K2: the cut of a deferred family row, into the staging row (the kernel writes the round's CSR directly)
static bool run_family_cut(ktn_handle* h, int64_t row, double& b) {
    KtnProblem& P = h->prob;
    const uint32_t slot = (uint32_t)P.row_slot[row], c = slot >> 5, lane = slot & 31u;
    const KtnChunkDesc& cd = P.chunks[c];
    const int fam = (int)P.shapes[cd.shape].family;
    if (c < P.fam_begin[fam] || c >= P.fam_begin[fam + 1] || cd.blob_off != P.cls_blob_off[fam][cd.aux] + (uint64_t)(c - P.cls_begin[fam][cd.aux]) * KTN_FAM_BLOB_BYTES(cd.aux)) {
        fprintf(stderr, "emu: family blob addressing violated\n"); abort(); }
    const EmuFamRow r = emu_row(P, cd, lane, h->x_last.data());
    const int64_t base = P.jac_ptr[row];
    const uint64_t rw = r.rankword();
    for (uint32_t u = 0; u < r.nu; ++u) if (P.jac_col[base + ((rw >> (4 * u)) & 15u)] != r.gcol(u)) { fprintf(stderr, "emu: rank word does not reproduce the Jacobian structure\n"); abort(); }
    bool mismatch = false;
    EmuCutSink s{h->stage_val.data() + base, P.jac_col.data() + base, &mismatch, {0}};
    const bool bad = fam == KTN_FAM_LSE ? ktn_family_cut_terms<KTN_FAM_LSE>(r, r.nu, rw, s, h->g_row[row], h->aux_row[row], h->do_round != 0, h->opt.cut_coef_rng, b)
                   : fam == KTN_FAM_QUAD ? ktn_family_cut_terms<KTN_FAM_QUAD>(r, r.nu, rw, s, h->g_row[row], h->aux_row[row], h->do_round != 0, h->opt.cut_coef_rng, b)
                                         : ktn_family_cut_terms<KTN_FAM_SOC>(r, r.nu, rw, s, h->g_row[row], h->aux_row[row], h->do_round != 0, h->opt.cut_coef_rng, b);
    if (mismatch) { fprintf(stderr, "emu: the cut's columns differ from the Jacobian structure\n"); abort(); }
    return bad;
}

static uint32_t ord_at(const uint8_t* ord, uint32_t ob, size_t e) { return ob == 1 ? ord[e] : ob == 2 ? ((const uint16_t*)ord)[e] : ((const uint32_t*)ord)[e]; }

// mode 0 = separate, 1 = force(mask), 2 = eval only
static void run_chunks(ktn_handle* h, const double* x, int mode, const std::vector<uint8_t>& force, int do_round) {
    KtnProblem& P = h->prob;
    std::vector<double> S;
    h->x_last.assign(x, x + P.num_var); h->do_round = do_round;
    for (size_t c = 0; c < P.chunks.size(); ++c) {
        const KtnChunkDesc& cd = P.chunks[c]; const KtnShapeDesc& sd = P.shapes[cd.shape];
        const uint32_t L = cd.stride, nu = sd.n_uniq;
        const uint8_t* blob = P.blob.data() + cd.blob_off;
        const size_t sec_col = ((size_t)8 * sd.n_const * L + 15) & ~(size_t)15, sec_ord = (sec_col + (size_t)4 * nu * L + 15) & ~(size_t)15;
        const int32_t* cols = (const int32_t*)(blob + sec_col); const uint8_t* ord = blob + sec_ord;
        S.assign((size_t)sd.n_scratch * L, 0.0);
        for (uint32_t lane = 0; lane < cd.nrows; ++lane) {
            const int32_t row = P.chunk_rows[cd.row_slot + lane];
            if (mode == 0 && !(sd.flags & KTN_SH_NL)) { h->sel[row] = 0; continue; }
            if (sd.family != KTN_FAM_GENERIC) {   // family chunks never take the interpreter (their blob carries rank, not order)
                const double flb = P.chunk_lb[cd.row_slot + lane], fub = P.chunk_ub[cd.row_slot + lane];
                if (cd.row_slot != c * 32 || L != 32 || cd.aux != nu) { fprintf(stderr, "emu: family chunk layout violated\n"); abort(); }
                const bool forced = mode == 1 && force[row] != 0;
                if (sd.family == KTN_FAM_LSE) run_family_dispatch<KTN_FAM_LSE>(nu, h, cd, lane, row, x, mode, forced, flb, fub, do_round);
                else if (sd.family == KTN_FAM_QUAD) run_family_dispatch<KTN_FAM_QUAD>(nu, h, cd, lane, row, x, mode, forced, flb, fub, do_round);
                else run_family_dispatch<KTN_FAM_SOC>(nu, h, cd, lane, row, x, mode, forced, flb, fub, do_round);
                continue;
            }
            for (uint32_t u = 0; u < nu; ++u) S[(size_t)u * L + lane] = x[cols[(size_t)u * L + lane]];
            // aliased shapes write into their constants: work on a private copy of the chunk blob, as the kernel's
            // shared-memory staging does
            std::vector<uint8_t> blobcopy(blob, blob + cd.blob_bytes);
            double* Cc = (double*)blobcopy.data();
            double* Jb = sd.j_in_blob ? Cc : S.data();
            struct EmuMem { double* C; double* S; double* Jp; size_t jmul; uint32_t lane, L;
                double c(uint32_t i) const { return C[(size_t)i * L + lane]; } void cst(uint32_t i, double v) { C[(size_t)i * L + lane] = v; }
                double lds(uint32_t i) const { return S[(size_t)i * L + lane]; } void sts(uint32_t i, double v) { S[(size_t)i * L + lane] = v; }
                double jld(uint32_t u) const { return Jp[(size_t)u * jmul]; } void jst(uint32_t u, double v) { Jp[(size_t)u * jmul] = v; }
                size_t stride() const { return L; } size_t jstride() const { return jmul; }
                double* caddr(uint32_t i) const { return C + (size_t)i * L + lane; } double* saddr(uint32_t i) const { return S + (size_t)i * L + lane; }
                double* jaddr(uint32_t u) const { return Jp + (size_t)u * jmul; } };
            EmuMem m{Cc, S.data(), Jb + (size_t)sd.j_base * L + lane, (size_t)L * sd.j_stride, lane, L};
            const KtnIns* prog = P.prog.data() + sd.prog_off;
            const double g = run_program(prog, 0, sd.n_fwd, m, 0u);
            h->g_row[row] = g;
            if (mode == 2) continue;
            const double lb = P.chunk_lb[cd.row_slot + lane], ub = P.chunk_ub[cd.row_slot + lane];
            bool selected = mode == 1 ? force[row] != 0 : !((g >= lb - h->opt.f_tol) && (g <= ub + h->opt.f_tol));
            if (!selected) { h->sel[row] = 0; continue; }
            run_program(prog, sd.n_fwd, sd.n_ins, m, 0u);
            const int64_t base = P.jac_ptr[row];
            if (!(sd.flags & KTN_SH_DENSE)) {
                double b = g, mx = 0.0;
                for (uint32_t q = 0; q < nu; ++q) { uint32_t u = ord_at(ord, sd.order_bytes, (size_t)q * L + lane); double jv = m.jld(u), xv = S[(size_t)u * L + lane]; b = b + (-xv) * jv; mx = q == 0 ? jv : ktn_jlmax(mx, jv); }
                bool bad = false;
                for (uint32_t q = 0; q < nu; ++q) { uint32_t u = ord_at(ord, sd.order_bytes, (size_t)q * L + lane); double jv = m.jld(u); if (do_round && jv + h->opt.cut_coef_rng < mx) jv = 0.0; bad = bad || !ktn_isfinite(jv); h->stage_val[base + q] = jv; }
                h->b_row[row] = b; h->sel[row] = nu | (bad ? KTN_SEL_ERRBIT : 0u);
            } else {
                const int64_t n = P.num_var; double* out = h->stage_val.data() + base;
                for (int64_t j = 0; j < n; ++j) out[j] = 0.0;
                for (uint32_t u = 0; u < nu; ++u) out[cols[(size_t)u * L + lane]] = m.jld(u);
                double b = g, mx = -ktn_inf();
                for (int64_t j = 0; j < n; ++j) { b = b + (-x[j]) * out[j]; mx = ktn_jlmax(mx, out[j]); }
                bool bad = false;
                for (int64_t j = 0; j < n; ++j) { double jv = out[j]; if (do_round && jv + h->opt.cut_coef_rng < mx) jv = 0.0; bad = bad || !ktn_isfinite(jv); out[j] = jv; }
                h->b_row[row] = b; h->sel[row] = (uint32_t)n | (bad ? KTN_SEL_ERRBIT : 0u);
            }
        }
    }
}

static int compact(ktn_handle* h, int64_t* n_cuts, int64_t* nnz, int64_t* err_row) {
    KtnProblem& P = h->prob;
    h->c_row.clear(); h->c_ptr.assign(1, 0); h->c_col.clear(); h->c_val.clear(); h->c_lo.clear(); h->c_hi.clear(); h->c_g.clear(); h->c_viol.clear(); h->c_b.clear();
    h->err_row = -1;
    for (int64_t i = 0; i < P.num_constr; ++i) {
        uint32_t s = h->sel[i]; if (!s) continue;
        if (s & KTN_SEL_DEFER) { double b; const bool bad = run_family_cut(h, i, b); h->b_row[i] = b; s = KTN_SEL_NNZ(s) | (bad ? KTN_SEL_ERRBIT : 0u); }
        if (s & KTN_SEL_ERRBIT) { h->err_row = i; break; }
        const int64_t base = P.jac_ptr[i];
        for (uint32_t q = 0; q < s; ++q) { h->c_col.push_back(P.jac_col[base + q]); h->c_val.push_back(h->stage_val[base + q]); }
        h->c_row.push_back(i); h->c_ptr.push_back((int64_t)h->c_col.size());
        const double g = h->g_row[i], b = h->b_row[i];
        h->c_lo.push_back(P.lb[i] - b); h->c_hi.push_back(P.ub[i] - b); h->c_g.push_back(g); h->c_b.push_back(b);
        const double v1 = P.lb[i] - g, v2 = g - P.ub[i]; h->c_viol.push_back(g == g ? (v1 > v2 ? v1 : v2) : g);
    }
    if (n_cuts) *n_cuts = (int64_t)h->c_row.size();
    if (nnz) *nnz = (int64_t)h->c_col.size();
    if (err_row) *err_row = h->err_row;
    h->have_round = true;
    return h->err_row >= 0 ? KTN_NUMERIC_NONFINITE : KTN_OK;
}

extern "C" {
int ktn_separate(ktn_handle* h, const double* x, int64_t* nc, int64_t* nz, int64_t* er) { run_chunks(h, x, 0, {}, 1); return compact(h, nc, nz, er); }
int ktn_gencut_rows(ktn_handle* h, const double* x, const int64_t* rows, int64_t nrows, int do_round, int64_t* nc, int64_t* nz, int64_t* er) {
    std::vector<uint8_t> mask((size_t)h->prob.num_constr, 0); for (int64_t j = 0; j < nrows; ++j) mask[rows[j]] = 1;
    run_chunks(h, x, 1, mask, do_round); return compact(h, nc, nz, er); }
int ktn_fetch_cuts(ktn_handle* h, int64_t* row_id, int64_t* row_ptr, int32_t* col, double* val, double* lo, double* hi, double* g, double* viol, double* bconst) {
    size_t nc = h->c_row.size(), nz = h->c_col.size();
    if (row_id && nc) memcpy(row_id, h->c_row.data(), 8 * nc);
    if (row_ptr) memcpy(row_ptr, h->c_ptr.data(), 8 * (nc + 1));
    if (col && nz) memcpy(col, h->c_col.data(), 4 * nz);
    if (val && nz) memcpy(val, h->c_val.data(), 8 * nz);
    if (lo && nc) memcpy(lo, h->c_lo.data(), 8 * nc);
    if (hi && nc) memcpy(hi, h->c_hi.data(), 8 * nc);
    if (g && nc) memcpy(g, h->c_g.data(), 8 * nc);
    if (viol && nc) memcpy(viol, h->c_viol.data(), 8 * nc);
    if (bconst && nc) memcpy(bconst, h->c_b.data(), 8 * nc);
    return 0; }
int ktn_fetch_cuts_view(ktn_handle* h, ktn_cut_view* v) {
    v->n_cuts = (int64_t)h->c_row.size(); v->nnz = (int64_t)h->c_col.size(); v->row_id = h->c_row.data(); v->row_ptr = h->c_ptr.data(); v->col = h->c_col.data(); v->val = h->c_val.data();
    v->lo = h->c_lo.data(); v->hi = h->c_hi.data(); v->g = h->c_g.data(); v->viol = h->c_viol.data(); v->bconst = h->c_b.data(); return 0; }
int ktn_get_g(ktn_handle* h, double* g) { memcpy(g, h->g_row.data(), 8 * h->g_row.size()); return 0; }
int ktn_eval_g(ktn_handle* h, const double* x, double* g) { run_chunks(h, x, 2, {}, 0); memcpy(g, h->g_row.data(), 8 * h->g_row.size()); return 0; }
int ktn_timings_get(ktn_handle*, ktn_timings* t) { memset(t, 0, sizeof *t); return 0; }
int64_t ktn_algorithmic_bytes(ktn_handle* h) { return h->prob.alg_bytes_static + 12 * (int64_t)h->c_col.size() + 28 * (int64_t)h->c_row.size(); }
// number of shapes / chunks, for tests of the packing
int64_t ktn_emu_num_shapes(ktn_handle* h) { return (int64_t)h->prob.shapes.size(); }
int64_t ktn_emu_num_chunks(ktn_handle* h) { return (int64_t)h->prob.chunks.size(); }
int64_t ktn_emu_num_family_chunks(ktn_handle* h, int32_t fam) { return fam < 0 || fam >= KTN_FAM__COUNT ? -1 : (int64_t)h->prob.fam_begin[fam + 1] - (int64_t)h->prob.fam_begin[fam]; }
int64_t ktn_emu_num_big_chunks(ktn_handle* h) { return (int64_t)h->prob.chunks.size() - h->prob.n_regular_chunks; }
int ktn_separate_ladder(ktn_handle*, const double*, int32_t, int32_t, int32_t*, int64_t*, int64_t*, int64_t*) { return KTN_ERR_UNSUPPORTED; }
int ktn_set_stream(ktn_handle*, void*) { return KTN_ERR_UNSUPPORTED; }
int ktn_separate_device_async(ktn_handle*, const double*) { return KTN_ERR_UNSUPPORTED; }
int ktn_sync_counts(ktn_handle*, int64_t*, int64_t*, int64_t*) { return KTN_ERR_UNSUPPORTED; }
int ktn_comm_unique_id(void*) { return KTN_ERR_UNSUPPORTED; }
int ktn_comm_init(ktn_handle*, int32_t, int32_t, const void*) { return KTN_ERR_UNSUPPORTED; }
int ktn_set_row_offset(ktn_handle*, int64_t) { return KTN_ERR_UNSUPPORTED; }
int ktn_allgather_cuts_async(ktn_handle*) { return KTN_ERR_UNSUPPORTED; }
int ktn_exchange_transport(ktn_handle*) { return 0; }
int ktn_sync_gathered(ktn_handle*, int64_t*, int64_t*) { return KTN_ERR_UNSUPPORTED; }
int ktn_gathered_error_row(ktn_handle*, int64_t*) { return KTN_ERR_UNSUPPORTED; }
int ktn_fetch_gathered(ktn_handle*, int64_t*, int64_t*, int32_t*, double*, double*, double*, double*, double*, double*) { return KTN_ERR_UNSUPPORTED; }
}
extern "C" int ktn_emu_shape_info(ktn_handle* h, int64_t sid, uint32_t* out /* n_fwd, n_ins, n_uniq, n_const, n_scratch, flags */) {
    if (sid < 0 || sid >= (int64_t)h->prob.shapes.size()) return -1;
    const KtnShapeDesc& s = h->prob.shapes[sid];
    out[0] = s.n_fwd; out[1] = s.n_ins; out[2] = s.n_uniq; out[3] = s.n_const; out[4] = s.n_scratch; out[5] = s.flags; out[6] = s.j_in_blob; out[7] = s.j_base; out[8] = s.j_stride; return 0; }
extern "C" int ktn_emu_shape_prog(ktn_handle* h, int64_t sid, uint32_t* out /* 4 words per instruction */) {
    const KtnShapeDesc& s = h->prob.shapes[sid];
    memcpy(out, h->prob.prog.data() + s.prog_off, 16 * (size_t)s.n_ins); return 0; }
// vectorised access to the shared math header for tests/test_math.py
extern "C" void ktn_test_exp(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = ktn_exp(x[i]); }
extern "C" void ktn_test_log(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = ktn_log(x[i]); }
extern "C" void ktn_test_sin(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = ktn_sin(x[i]); }
extern "C" void ktn_test_cos(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = ktn_cos(x[i]); }
extern "C" void ktn_test_pow(const double* x, const double* p, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = ktn_pow(x[i], p[i]); }
