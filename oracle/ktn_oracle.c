/* ktn_oracle.c -- CPU ORACLE for the ECP separation round.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a plain-C restatement of the reference algorithm (lanl-ansi/Katana.jl).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it; the product (katana.jl_b200/) never does.
 *
 * PARITY STATUS: "parity unpinned" at the separator boundary.  The reference cannot run
 * here (no julia binary), and its own tests pin only end-to-end optima (test/runtests.jl:16-20),
 * never g, Jacobian entries or cut rows.  The arithmetic of eval_g / eval_jac_g lives in
 * un-vendored dependencies (REQUIRE:3-4: MathProgBase 0.6-0.7, JuMP 0.17-0.18 with
 * ReverseDiffSparse); this file restates their published tape algorithm:
 *   forward_eval   children before parents, storage[] + partials_storage[]
 *   reverse_eval   reverse[k] = reverse[parent] * partials[k], 0 if parent adjoint is 0 and the partial is not finite
 *   reverse_extract  grad[var] += reverse[k] in node order; structure = sorted unique columns
 * The oracle is pinned against (a) hand-derived per-round KATs and mpmath-evaluated
 * analytic gradients of the reference's test expressions (tests/golden/), (b) the
 * reference's end-to-end optima (test/2d.jl, 3d.jl, misc.jl, basic.jl, lpqp.jl) through
 * the ECP driver.
 *
 * It implements the SAME C ABI as the CUDA library (include/ktn.h).
 * Build: gcc -O2 -ffp-contract=off -mfma -fopenmp (see oracle/Makefile).  The OpenMP
 * variant parallelises over rows only; per-row arithmetic is unchanged.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include "../include/ktn.h"
#include "../katana.jl_b200/csrc/ktn_math.h"
#ifdef _OPENMP
#include <omp.h>
#endif

struct ktn_handle {
    double f_tol, cut_coef_rng;
    int64_t topk;
    int64_t num_var, num_constr, rows_loaded;
    /* expression store (all rows concatenated) */
    int64_t n_nodes, cap_nodes;
    int64_t* expr_ptr;     /* num_constr+1 */
    int32_t* op; int32_t* arg; double* val;
    int32_t* parent;       /* parent node (row-relative), -1 for root */
    int32_t* send;         /* row-relative index one past the node's subtree (= next sibling) */
    double* lb; double* ub; uint8_t* flags;
    /* jac structure: reference sep.sp_cols / sep.sp_col_inds (src/separators.jl:93-100) */
    int64_t* jac_ptr; int32_t* jac_col; int64_t jac_cap;
    /* reference sep.g / sep.jac / sep.xstar (src/separators.jl:69-71) */
    double* g; double* jac; double* xstar;
    int have_round;
    int64_t max_nodes;
    /* last cut batch */
    int64_t n_cuts, nnz_cuts, err_row;
    int64_t* c_row; int64_t* c_ptr; int32_t* c_col; double* c_val;
    double* c_lo; double* c_hi; double* c_g; double* c_viol; double* c_b;
    int64_t cap_cuts, cap_nnz;
    int threads;
    ktn_timings tm;
    char err[512];
};

static int fail(ktn_handle* h, int code, const char* fmt, ...) {
    if (h) { va_list ap; va_start(ap, fmt); vsnprintf(h->err, sizeof h->err, fmt, ap); va_end(ap); }
    return code;
}

const char* ktn_backend(void) { return "oracle"; }
const char* ktn_last_error(ktn_handle* h) { return h ? h->err : "null handle"; }

static void free_problem(ktn_handle* h) {
    free(h->expr_ptr); free(h->op); free(h->arg); free(h->val); free(h->parent); free(h->send);
    free(h->lb); free(h->ub); free(h->flags); free(h->jac_ptr); free(h->jac_col);
    free(h->g); free(h->jac); free(h->xstar);
    free(h->c_row); free(h->c_ptr); free(h->c_col); free(h->c_val);
    free(h->c_lo); free(h->c_hi); free(h->c_g); free(h->c_viol); free(h->c_b);
    h->expr_ptr = NULL; h->op = NULL; h->arg = NULL; h->val = NULL; h->parent = NULL; h->send = NULL;
    h->lb = h->ub = NULL; h->flags = NULL; h->jac_ptr = NULL; h->jac_col = NULL;
    h->g = h->jac = h->xstar = NULL;
    h->c_row = h->c_ptr = NULL; h->c_col = NULL; h->c_val = h->c_lo = h->c_hi = h->c_g = h->c_viol = h->c_b = NULL;
    h->cap_cuts = h->cap_nnz = 0; h->n_nodes = h->cap_nodes = 0; h->rows_loaded = 0; h->jac_cap = 0;
    h->have_round = 0; h->n_cuts = h->nnz_cuts = 0; h->err_row = -1; h->max_nodes = 0;
}

int ktn_create(const ktn_options* o, ktn_handle** out) {
    if (!out) return KTN_ERR_USAGE;
    ktn_handle* h = (ktn_handle*)calloc(1, sizeof *h);
    if (!h) return KTN_ERR_NOMEM;
    h->f_tol = o ? o->f_tol : 1e-6; h->cut_coef_rng = o ? o->cut_coef_rng : 1e9; h->topk = o ? o->topk : 0;
    h->err_row = -1; h->threads = 1;
    const char* t = getenv("KTN_ORACLE_THREADS");
    if (t) h->threads = atoi(t) > 0 ? atoi(t) : 1;
    *out = h; return KTN_OK;
}
void ktn_destroy(ktn_handle* h) { if (h) { free_problem(h); free(h); } }
int ktn_set_params(ktn_handle* h, double f_tol, double rng, int64_t topk) {
    if (!h) return KTN_ERR_USAGE;
    h->f_tol = f_tol; h->cut_coef_rng = rng; h->topk = topk; return KTN_OK;
}

int ktn_load_begin(ktn_handle* h, int64_t num_var, int64_t num_constr) {
    if (!h || num_var < 0 || num_constr < 0) return fail(h, KTN_ERR_USAGE, "bad sizes");
    free_problem(h);
    h->num_var = num_var; h->num_constr = num_constr;
    h->expr_ptr = (int64_t*)calloc((size_t)num_constr + 1, sizeof(int64_t));
    h->lb = (double*)malloc(sizeof(double) * (size_t)(num_constr + 1));
    h->ub = (double*)malloc(sizeof(double) * (size_t)(num_constr + 1));
    h->flags = (uint8_t*)malloc((size_t)num_constr + 1);
    h->jac_ptr = (int64_t*)calloc((size_t)num_constr + 1, sizeof(int64_t));
    h->g = (double*)calloc((size_t)num_constr + 1, sizeof(double));
    h->xstar = (double*)calloc((size_t)num_var + 1, sizeof(double));
    return KTN_OK;
}

static int cmp_i32(const void* a, const void* b) { int32_t x = *(const int32_t*)a, y = *(const int32_t*)b; return x < y ? -1 : x > y; }

/* number of children each op requires; -1 = n-ary (>= 1) */
static int arity(int op) {
    switch (op) {
        case KTN_OP_CONST: case KTN_OP_VAR: return 0;
        case KTN_OP_ADD: case KTN_OP_MUL: return -1;
        case KTN_OP_SUB: case KTN_OP_DIV: case KTN_OP_POW: case KTN_OP_LE: case KTN_OP_LT: case KTN_OP_GE: case KTN_OP_GT: case KTN_OP_EQ: return 2;
        case KTN_OP_NEG: case KTN_OP_EXP: case KTN_OP_LOG: case KTN_OP_SQRT: case KTN_OP_ABS: case KTN_OP_SIN: case KTN_OP_COS: return 1;
        case KTN_OP_IFELSE: return 3;
        default: return -2;
    }
}

int ktn_add_rows(ktn_handle* h, int64_t first_row, int64_t nrows, const int64_t* eptr, const int32_t* op,
                 const int32_t* arg, const double* val, const double* lb, const double* ub, const uint8_t* flags) {
    if (!h || !h->expr_ptr) return fail(h, KTN_ERR_USAGE, "ktn_add_rows before ktn_load_begin");
    if (first_row != h->rows_loaded || first_row + nrows > h->num_constr) return fail(h, KTN_ERR_USAGE, "rows must be added in ascending order");
    int64_t nn = eptr[nrows] - eptr[0];
    if (h->n_nodes + nn > h->cap_nodes) {
        int64_t cap = (h->n_nodes + nn) * 3 / 2 + 64;
        h->op = (int32_t*)realloc(h->op, sizeof(int32_t) * (size_t)cap);
        h->arg = (int32_t*)realloc(h->arg, sizeof(int32_t) * (size_t)cap);
        h->val = (double*)realloc(h->val, sizeof(double) * (size_t)cap);
        h->parent = (int32_t*)realloc(h->parent, sizeof(int32_t) * (size_t)cap);
        h->send = (int32_t*)realloc(h->send, sizeof(int32_t) * (size_t)cap);
        h->cap_nodes = cap;
    }
    memcpy(h->op + h->n_nodes, op + eptr[0], sizeof(int32_t) * (size_t)nn);
    memcpy(h->arg + h->n_nodes, arg + eptr[0], sizeof(int32_t) * (size_t)nn);
    memcpy(h->val + h->n_nodes, val + eptr[0], sizeof(double) * (size_t)nn);
    int32_t* tmpc = NULL; int64_t tmpcap = 0;
    for (int64_t r = 0; r < nrows; ++r) {
        int64_t row = first_row + r, b = h->n_nodes + (eptr[r] - eptr[0]), e = h->n_nodes + (eptr[r + 1] - eptr[0]);
        int64_t n = e - b;
        if (n <= 0) return fail(h, KTN_ERR_USAGE, "row %lld has an empty expression", (long long)row);
        if (n > h->max_nodes) h->max_nodes = n;
        h->expr_ptr[row] = b; h->expr_ptr[row + 1] = e;
        h->lb[row] = lb[r]; h->ub[row] = ub[r]; h->flags[row] = flags[r];
        /* parents from prefix order + child counts (explicit stack of pending child slots) */
        int32_t* par = h->parent + b;
        {
            int64_t sp = 0; /* stack of (node, remaining) kept in two scratch arrays */
            int32_t* stn = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + 1));
            int32_t* str = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + 1));
            for (int64_t k = 0; k < n; ++k) {
                int o = h->op[b + k], a = arity(o);
                if (a == -2) { free(stn); free(str); return fail(h, KTN_ERR_USAGE, "row %lld: unknown op %d", (long long)row, o); }
                if (sp == 0) { if (k != 0) { free(stn); free(str); return fail(h, KTN_ERR_USAGE, "row %lld: more than one root", (long long)row); } par[k] = -1; }
                else { par[k] = stn[sp - 1]; if (--str[sp - 1] == 0) --sp; }
                int nc = (a == 0) ? 0 : h->arg[b + k];
                if (a > 0 && nc != a) { free(stn); free(str); return fail(h, KTN_ERR_USAGE, "row %lld: op %d needs %d children", (long long)row, o, a); }
                if (a == -1 && nc < 1) { free(stn); free(str); return fail(h, KTN_ERR_USAGE, "row %lld: n-ary op without children", (long long)row); }
                if (o == KTN_OP_VAR && (h->arg[b + k] < 0 || h->arg[b + k] >= h->num_var)) { free(stn); free(str); return fail(h, KTN_ERR_USAGE, "row %lld: variable index out of range", (long long)row); }
                if (nc > 0) { stn[sp] = (int32_t)k; str[sp] = nc; ++sp; }
            }
            free(stn); free(str);
            if (sp != 0) return fail(h, KTN_ERR_USAGE, "row %lld: truncated expression", (long long)row);
            int32_t* se = h->send + b;
            for (int64_t k = n - 1; k >= 0; --k) {
                int64_t c = k + 1; int nc = arity(h->op[b + k]) == 0 ? 0 : h->arg[b + k];
                for (int i = 0; i < nc; ++i) c = se[c];
                se[k] = (int32_t)c;
            }
        }
        /* jac structure: grad_sparsity = sorted unique variable indices; dense row = every column */
        int64_t cnt = 0;
        if (flags[r] & KTN_ROW_DENSE) cnt = h->num_var;
        else {
            if (n > tmpcap) { tmpcap = n * 2; tmpc = (int32_t*)realloc(tmpc, sizeof(int32_t) * (size_t)tmpcap); }
            int64_t nv = 0;
            for (int64_t k = 0; k < n; ++k) if (h->op[b + k] == KTN_OP_VAR) tmpc[nv++] = h->arg[b + k];
            qsort(tmpc, (size_t)nv, sizeof(int32_t), cmp_i32);
            for (int64_t k = 0; k < nv; ++k) if (k == 0 || tmpc[k] != tmpc[k - 1]) tmpc[cnt++] = tmpc[k];
        }
        int64_t jp = h->jac_ptr[row];
        if (jp + cnt > h->jac_cap) { h->jac_cap = (jp + cnt) * 3 / 2 + 64; h->jac_col = (int32_t*)realloc(h->jac_col, sizeof(int32_t) * (size_t)h->jac_cap); }
        if (flags[r] & KTN_ROW_DENSE) for (int64_t k = 0; k < cnt; ++k) h->jac_col[jp + k] = (int32_t)k;
        else memcpy(h->jac_col + jp, tmpc, sizeof(int32_t) * (size_t)cnt);
        h->jac_ptr[row + 1] = jp + cnt;
    }
    free(tmpc);
    h->n_nodes += nn; h->rows_loaded += nrows;
    return KTN_OK;
}

int ktn_load_end(ktn_handle* h) {
    if (!h || !h->expr_ptr) return fail(h, KTN_ERR_USAGE, "no problem");
    if (h->rows_loaded != h->num_constr) return fail(h, KTN_ERR_USAGE, "loaded %lld of %lld rows", (long long)h->rows_loaded, (long long)h->num_constr);
    free(h->jac); h->jac = (double*)calloc((size_t)h->jac_ptr[h->num_constr] + 1, sizeof(double)); /* sep.jac = zeros(N), src/separators.jl:103 */
    return KTN_OK;
}

int64_t ktn_num_rows(ktn_handle* h) { return h ? h->rows_loaded : 0; }
int64_t ktn_jac_nnz(ktn_handle* h) { return (h && h->jac_ptr) ? h->jac_ptr[h->rows_loaded] : 0; }
int ktn_jac_structure(ktn_handle* h, int64_t* row_ptr, int32_t* cols) {
    if (!h || !h->jac_ptr) return fail(h, KTN_ERR_USAGE, "no problem");
    if (row_ptr) memcpy(row_ptr, h->jac_ptr, sizeof(int64_t) * (size_t)(h->rows_loaded + 1));
    if (cols) memcpy(cols, h->jac_col, sizeof(int32_t) * (size_t)h->jac_ptr[h->rows_loaded]);
    return KTN_OK;
}

/* ---- tape interpreter: forward_eval (children before parents), reverse_eval, reverse_extract ---- */
typedef struct { double* st; double* pa; double* rv; double* gw; } scratch_t;

/* forward sweep over one row; returns g_i(x).  Fills st[] (values) and pa[] (d parent / d node). */
static double forward_row(const ktn_handle* h, int64_t row, const double* x, double* st, double* pa) {
    int64_t b = h->expr_ptr[row], n = h->expr_ptr[row + 1] - b;
    const int32_t* op = h->op + b; const int32_t* arg = h->arg + b; const double* val = h->val + b; const int32_t* se = h->send + b;
    for (int64_t k = n - 1; k >= 0; --k) {
        switch (op[k]) {
            case KTN_OP_CONST: st[k] = val[k]; break;
            case KTN_OP_VAR: st[k] = x[arg[k]]; break;
            case KTN_OP_ADD: { /* tmp_sum = 0; tmp_sum += child; partial = 1 */
                double s = 0.0; int nc = arg[k]; int64_t c = k + 1;
                for (int i = 0; i < nc; ++i) { s = s + st[c]; pa[c] = 1.0; c = se[c]; }
                st[k] = s; break; }
            case KTN_OP_SUB: { int64_t c1 = k + 1, c2 = se[c1];
                pa[c1] = 1.0; pa[c2] = -1.0; st[k] = st[c1] - st[c2]; break; }
            case KTN_OP_MUL: { /* tmp_prod = 1; tmp_prod *= child */
                int nc = arg[k]; double p = 1.0; int64_t c = k + 1;
                for (int i = 0; i < nc; ++i) { p = p * st[c]; c = se[c]; }
                if (p == 0.0 || nc <= 2) { /* product of the others, left to right from 1 */
                    int64_t ci = k + 1;
                    for (int i = 0; i < nc; ++i) {
                        double po = 1.0; int64_t cj = k + 1;
                        for (int j = 0; j < nc; ++j) { if (j != i) po = po * st[cj]; cj = se[cj]; }
                        pa[ci] = po; ci = se[ci];
                    }
                } else {
                    int64_t ci = k + 1;
                    for (int i = 0; i < nc; ++i) { pa[ci] = p / st[ci]; ci = se[ci]; }
                }
                st[k] = p; break; }
            case KTN_OP_DIV: { int64_t c1 = k + 1, c2 = se[c1];
                double num = st[c1], den = st[c2], rec = 1.0 / den;
                pa[c1] = rec; pa[c2] = (-num * rec) * rec; st[k] = num * rec; break; }
            case KTN_OP_POW: { int64_t c1 = k + 1, c2 = se[c1];
                double base = st[c1], ex = st[c2];
                if (ex == 2.0) { st[k] = base * base; pa[c1] = 2.0 * base; }
                else if (ex == 1.0) { st[k] = base; pa[c1] = 1.0; }
                else { st[k] = ktn_pow(base, ex); pa[c1] = ex * ktn_pow(base, ex - 1.0); }
                /* d/d exponent = value * log(base); only observable when the exponent holds a variable */
                pa[c2] = (op[c2] == KTN_OP_CONST) ? 0.0 : st[k] * ktn_log(base);
                break; }
            case KTN_OP_NEG: st[k] = -st[k + 1]; pa[k + 1] = -1.0; break;
            case KTN_OP_EXP: st[k] = ktn_exp(st[k + 1]); pa[k + 1] = st[k]; break;
            case KTN_OP_LOG: st[k] = ktn_log(st[k + 1]); pa[k + 1] = 1.0 / st[k + 1]; break;
            case KTN_OP_SQRT: st[k] = ktn_sqrt(st[k + 1]); pa[k + 1] = 0.5 / st[k]; break; /* Calculus.jl: 1 / 2 / sqrt(x) */
            case KTN_OP_ABS: st[k] = ktn_fabs(st[k + 1]); pa[k + 1] = (st[k + 1] >= 0.0) ? 1.0 : -1.0; break;
            case KTN_OP_SIN: st[k] = ktn_sin(st[k + 1]); pa[k + 1] = ktn_cos(st[k + 1]); break;      /* Calculus.jl: cos(x) */
            case KTN_OP_COS: st[k] = ktn_cos(st[k + 1]); pa[k + 1] = -ktn_sin(st[k + 1]); break;     /* Calculus.jl: -sin(x) */
            case KTN_OP_IFELSE: { int64_t c1 = k + 1, c2 = se[c1], c3 = se[c2];   /* [recalled] ReverseDiffSparse forward_eval: condition == 1 */
                int sel = st[c1] == 1.0;
                st[k] = sel ? st[c2] : st[c3]; pa[c1] = 0.0; pa[c2] = sel ? 1.0 : 0.0; pa[c3] = sel ? 0.0 : 1.0; break; }
            case KTN_OP_LE: case KTN_OP_LT: case KTN_OP_GE: case KTN_OP_GT: case KTN_OP_EQ: { int64_t c1 = k + 1, c2 = se[c1];
                double a = st[c1], b = st[c2];
                int r = op[k] == KTN_OP_LE ? a <= b : op[k] == KTN_OP_LT ? a < b : op[k] == KTN_OP_GE ? a >= b : op[k] == KTN_OP_GT ? a > b : a == b;
                st[k] = r ? 1.0 : 0.0; pa[c1] = 0.0; pa[c2] = 0.0; break; }
        }
    }
    return st[0];
}

/* reverse sweep + extraction into the row's jac slots (eval_jac_g for one row). */
static void reverse_row(const ktn_handle* h, int64_t row, const double* pa, double* rv, double* gw, double* jac_out) {
    int64_t b = h->expr_ptr[row], n = h->expr_ptr[row + 1] - b;
    const int32_t* op = h->op + b; const int32_t* arg = h->arg + b; const int32_t* par = h->parent + b;
    int64_t jp = h->jac_ptr[row], nz = h->jac_ptr[row + 1] - jp;
    const int32_t* cols = h->jac_col + jp;
    for (int64_t p = 0; p < nz; ++p) gw[cols[p]] = 0.0;
    rv[0] = 1.0;
    if (op[0] == KTN_OP_VAR) gw[arg[0]] += rv[0];
    for (int64_t k = 1; k < n; ++k) {
        if (op[k] == KTN_OP_CONST) continue;
        double rp = rv[par[k]], p = pa[k];
        rv[k] = (rp == 0.0 && !ktn_isfinite(p)) ? rp : rp * p;
        if (op[k] == KTN_OP_VAR) gw[arg[k]] += rv[k];
    }
    for (int64_t p = 0; p < nz; ++p) jac_out[p] = gw[cols[p]];
}

static scratch_t make_scratch(const ktn_handle* h) {
    scratch_t s; size_t n = (size_t)h->max_nodes + 1;
    s.st = (double*)malloc(n * 8); s.pa = (double*)malloc(n * 8); s.rv = (double*)malloc(n * 8);
    s.gw = (double*)calloc((size_t)h->num_var + 1, 8);
    return s;
}
static void free_scratch(scratch_t* s) { free(s->st); free(s->pa); free(s->rv); free(s->gw); }

/* precompute!(sep, xstar): eval_jac_g then eval_g over ALL rows (src/separators.jl:111-116). */
static void precompute(ktn_handle* h, const double* x) {
    memcpy(h->xstar, x, sizeof(double) * (size_t)h->num_var);
    int64_t m = h->num_constr;
#ifdef _OPENMP
#pragma omp parallel num_threads(h->threads)
#endif
    {
        scratch_t s = make_scratch(h);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int64_t i = 0; i < m; ++i) {
            h->g[i] = forward_row(h, i, h->xstar, s.st, s.pa);
            reverse_row(h, i, s.pa, s.rv, s.gw, h->jac + h->jac_ptr[i]);
        }
        free_scratch(&s);
    }
    h->have_round = 1;
}

/* isconstrsat (src/separators.jl:120) */
static int isconstrsat(const ktn_handle* h, int64_t i, double lb, double ub, double f_tol) {
    return (h->g[i] >= lb - f_tol) && (h->g[i] <= ub + f_tol);
}

static void reserve_cuts(ktn_handle* h, int64_t nc, int64_t nz) {
    if (nc + 1 > h->cap_cuts) {
        int64_t c = (nc + 1) * 3 / 2 + 16;
        h->c_row = (int64_t*)realloc(h->c_row, 8 * (size_t)c); h->c_ptr = (int64_t*)realloc(h->c_ptr, 8 * (size_t)(c + 1));
        h->c_lo = (double*)realloc(h->c_lo, 8 * (size_t)c); h->c_hi = (double*)realloc(h->c_hi, 8 * (size_t)c);
        h->c_g = (double*)realloc(h->c_g, 8 * (size_t)c); h->c_viol = (double*)realloc(h->c_viol, 8 * (size_t)c); h->c_b = (double*)realloc(h->c_b, 8 * (size_t)c);
        h->cap_cuts = c;
    }
    if (nz > h->cap_nnz) {
        int64_t c = nz * 3 / 2 + 64;
        h->c_col = (int32_t*)realloc(h->c_col, 4 * (size_t)c); h->c_val = (double*)realloc(h->c_val, 8 * (size_t)c);
        h->cap_nnz = c;
    }
}

/* gencut -> linear_oa_cut (src/algorithms.jl:3-18), round_coefs (src/model.jl:200-207),
 * _addcut (src/model.jl:68-79).  Returns 0 if the cut was appended, 1 if it had a non-finite
 * coefficient (reference: warn, m.status = :Error, no cut). */
static int emit_cut(ktn_handle* h, int64_t i, int do_round) {
    int64_t jp = h->jac_ptr[i], nz = h->jac_ptr[i + 1] - jp;
    reserve_cuts(h, h->n_cuts + 1, h->nnz_cuts + nz);
    int64_t o = h->nnz_cuts;
    double b = h->g[i];                                   /* algorithms.jl:8 */
    for (int64_t p = 0; p < nz; ++p) {                    /* algorithms.jl:9-16 */
        int32_t col = h->jac_col[jp + p];
        double partial = h->jac[jp + p];
        h->c_col[o + p] = col; h->c_val[o + p] = partial;
        double t = -h->xstar[col] * partial;
        b = b + t;                                        /* b += -sep.xstar[col]*partial */
    }
    if (do_round && nz > 0) {                             /* model.jl:201-205; max() is NaN-propagating */
        double mx = h->c_val[o];
        for (int64_t p = 1; p < nz; ++p) mx = ktn_jlmax(mx, h->c_val[o + p]);
        for (int64_t p = 0; p < nz; ++p) if (h->c_val[o + p] + h->cut_coef_rng < mx) h->c_val[o + p] = 0.0;
    }
    for (int64_t p = 0; p < nz; ++p) if (!ktn_isfinite(h->c_val[o + p])) return 1;   /* model.jl:69-73 */
    int64_t c = h->n_cuts;
    h->c_row[c] = i; h->c_ptr[c] = o; h->c_ptr[c + 1] = o + nz;
    h->c_lo[c] = h->lb[i] - b; h->c_hi[c] = h->ub[i] - b;  /* model.jl:74-75 */
    h->c_g[c] = h->g[i]; h->c_b[c] = b;
    double v1 = h->lb[i] - h->g[i], v2 = h->g[i] - h->ub[i];
    h->c_viol[c] = (h->g[i] == h->g[i]) ? (v1 > v2 ? v1 : v2) : h->g[i];
    h->n_cuts = c + 1; h->nnz_cuts = o + nz;
    return 0;
}

typedef struct { double v; int64_t i; } vi_t;
static int cmp_vi(const void* a, const void* b) { /* NaN first, then violation descending, then index ascending */
    const vi_t* x = (const vi_t*)a; const vi_t* y = (const vi_t*)b;
    int xn = !(x->v == x->v), yn = !(y->v == y->v);
    if (xn != yn) return xn ? -1 : 1;
    if (!xn && x->v != y->v) return x->v > y->v ? -1 : 1;
    return x->i < y->i ? -1 : x->i > y->i;
}
static int cmp_i64(const void* a, const void* b) { int64_t x = *(const int64_t*)a, y = *(const int64_t*)b; return x < y ? -1 : x > y; }

/* the loop body src/model.jl:265-283 */
int ktn_separate(ktn_handle* h, const double* xstar, int64_t* n_cuts, int64_t* nnz_cuts, int64_t* err_row) {
    if (!h || !h->jac) return fail(h, KTN_ERR_USAGE, "no problem loaded");
    precompute(h, xstar);                                 /* model.jl:268 */
    h->n_cuts = 0; h->nnz_cuts = 0; h->err_row = -1;
    reserve_cuts(h, 0, 0); h->c_ptr[0] = 0;
    int status = KTN_OK;
    if (h->topk <= 0) {
        for (int64_t i = 0; i < h->num_constr; ++i) {     /* for i in m.nlconstr_ixs (ascending, epigraph row last) */
            if (!(h->flags[i] & KTN_ROW_NL)) continue;
            if (isconstrsat(h, i, h->lb[i], h->ub[i], h->f_tol)) continue;
            if (emit_cut(h, i, 1)) { status = KTN_NUMERIC_NONFINITE; h->err_row = i; break; }   /* model.jl:278 */
        }
    } else { /* build extension: rank violated rows, keep the top k, emit ascending */
        int64_t nv = 0; vi_t* v = (vi_t*)malloc(sizeof(vi_t) * (size_t)(h->num_constr + 1));
        for (int64_t i = 0; i < h->num_constr; ++i) {
            if (!(h->flags[i] & KTN_ROW_NL) || isconstrsat(h, i, h->lb[i], h->ub[i], h->f_tol)) continue;
            double a = h->lb[i] - h->g[i], b = h->g[i] - h->ub[i];
            v[nv].v = (h->g[i] == h->g[i]) ? (a > b ? a : b) : h->g[i]; v[nv].i = i; ++nv;
        }
        qsort(v, (size_t)nv, sizeof(vi_t), cmp_vi);
        int64_t k = nv < h->topk ? nv : h->topk;
        int64_t* sel = (int64_t*)malloc(8 * (size_t)(k + 1));
        for (int64_t j = 0; j < k; ++j) sel[j] = v[j].i;
        qsort(sel, (size_t)k, 8, cmp_i64);
        for (int64_t j = 0; j < k; ++j) if (emit_cut(h, sel[j], 1)) { status = KTN_NUMERIC_NONFINITE; h->err_row = sel[j]; break; }
        free(sel); free(v);
    }
    if (n_cuts) *n_cuts = h->n_cuts;
    if (nnz_cuts) *nnz_cuts = h->nnz_cuts;
    if (err_row) *err_row = h->err_row;
    h->tm.rounds++;
    return status;
}

int ktn_gencut_rows(ktn_handle* h, const double* x, const int64_t* rows, int64_t nrows, int do_round,
                    int64_t* n_cuts, int64_t* nnz_cuts, int64_t* err_row) {
    if (!h || !h->jac) return fail(h, KTN_ERR_USAGE, "no problem loaded");
    for (int64_t j = 0; j < nrows; ++j) if (rows[j] < 0 || rows[j] >= h->num_constr || (j && rows[j] <= rows[j - 1])) return fail(h, KTN_ERR_USAGE, "rows must be ascending and in range");
    precompute(h, x);
    h->n_cuts = 0; h->nnz_cuts = 0; h->err_row = -1;
    reserve_cuts(h, 0, 0); h->c_ptr[0] = 0;
    int status = KTN_OK;
    for (int64_t j = 0; j < nrows; ++j) if (emit_cut(h, rows[j], do_round)) { status = KTN_NUMERIC_NONFINITE; h->err_row = rows[j]; break; }
    if (n_cuts) *n_cuts = h->n_cuts;
    if (nnz_cuts) *nnz_cuts = h->nnz_cuts;
    if (err_row) *err_row = h->err_row;
    return status;
}

int ktn_fetch_cuts_view(ktn_handle* h, ktn_cut_view* out) {
    if (!h || !out) return KTN_ERR_USAGE;
    out->n_cuts = h->n_cuts; out->nnz = h->nnz_cuts;
    out->row_id = h->c_row; out->row_ptr = h->c_ptr; out->col = h->c_col; out->val = h->c_val;
    out->lo = h->c_lo; out->hi = h->c_hi; out->g = h->c_g; out->viol = h->c_viol; out->bconst = h->c_b;
    return KTN_OK;
}
int ktn_fetch_cuts(ktn_handle* h, int64_t* row_id, int64_t* row_ptr, int32_t* col, double* val,
                   double* lo, double* hi, double* g, double* viol, double* bconst) {
    if (!h) return KTN_ERR_USAGE;
    size_t nc = (size_t)h->n_cuts, nz = (size_t)h->nnz_cuts;
    if (row_ptr) { if (h->c_ptr) memcpy(row_ptr, h->c_ptr, 8 * (nc + 1)); else row_ptr[0] = 0; }
    if (row_id && nc) memcpy(row_id, h->c_row, 8 * nc);
    if (col && nz) memcpy(col, h->c_col, 4 * nz);
    if (val && nz) memcpy(val, h->c_val, 8 * nz);
    if (lo && nc) memcpy(lo, h->c_lo, 8 * nc);
    if (hi && nc) memcpy(hi, h->c_hi, 8 * nc);
    if (g && nc) memcpy(g, h->c_g, 8 * nc);
    if (viol && nc) memcpy(viol, h->c_viol, 8 * nc);
    if (bconst && nc) memcpy(bconst, h->c_b, 8 * nc);
    return KTN_OK;
}

int ktn_get_g(ktn_handle* h, double* g_out) {
    if (!h || !h->have_round) return fail(h, KTN_ERR_USAGE, "no round has run");
    memcpy(g_out, h->g, 8 * (size_t)h->num_constr); return KTN_OK;
}

int ktn_eval_g(ktn_handle* h, const double* x, double* g_out) {
    if (!h || !h->jac) return fail(h, KTN_ERR_USAGE, "no problem loaded");
    int64_t m = h->num_constr;
#ifdef _OPENMP
#pragma omp parallel num_threads(h->threads)
#endif
    {
        scratch_t s = make_scratch(h);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int64_t i = 0; i < m; ++i) g_out[i] = forward_row(h, i, x, s.st, s.pa);
        free_scratch(&s);
    }
    return KTN_OK;
}

int ktn_set_bounds(ktn_handle* h, const double* lb, const double* ub) {
    if (!h || !h->jac) return fail(h, KTN_ERR_USAGE, "no problem loaded");
    memcpy(h->lb, lb, 8 * (size_t)h->num_constr); memcpy(h->ub, ub, 8 * (size_t)h->num_constr); return KTN_OK;
}

int ktn_timings_get(ktn_handle* h, ktn_timings* out) { if (!h || !out) return KTN_ERR_USAGE; *out = h->tm; return KTN_OK; }

/* sum_NL (4 nnz + 8 C + 16) + 8 n + sum_selected (12 nnz + 28) */
int64_t ktn_algorithmic_bytes(ktn_handle* h) {
    if (!h || !h->jac_ptr) return 0;
    int64_t by = 8 * h->num_var;
    for (int64_t i = 0; i < h->num_constr; ++i) {
        if (!(h->flags[i] & KTN_ROW_NL)) continue;
        int64_t C = 0, b0 = h->expr_ptr[i];
        for (int64_t k = b0; k < h->expr_ptr[i + 1]; ++k) {
            if (h->op[k] != KTN_OP_CONST) continue;
            /* the literal exponent of x^2 / x^1 is structure (ReverseDiffSparse special-cases it), not per-row data */
            int32_t pk = h->parent[k];
            if (pk >= 0 && h->op[b0 + pk] == KTN_OP_POW && b0 + h->send[pk + 1] == k && (h->val[k] == 2.0 || h->val[k] == 1.0)) continue;
            ++C;
        }
        by += 4 * (h->jac_ptr[i + 1] - h->jac_ptr[i]) + 8 * C + 16;
    }
    by += 12 * h->nnz_cuts + 28 * h->n_cuts;
    return by;
}

/* boundroutine, src/model.jl:175-197: the sequential ladder as the reference runs it: one full round per point, stop at the first
 * point where a row was violated (cuts were made) or a cut was not finite */
int ktn_separate_ladder(ktn_handle* h, const double* ray, int32_t n_first, int32_t n_last, int32_t* n_hit,
                        int64_t* n_cuts, int64_t* nnz_cuts, int64_t* err_row) {
    if (!h || !h->jac || !ray || n_first < 0 || n_last > 1023 || n_last < n_first) return fail(h, KTN_ERR_USAGE, "bad ladder");
    double* x = (double*)malloc(sizeof(double) * (size_t)(h->num_var + 1));
    int status = KTN_OK; int64_t nc = 0, nz = 0, er = -1;
    if (n_hit) *n_hit = -1;
    for (int32_t n = n_first; n <= n_last; ++n) {
        const double sc = ldexp(1.0, n);
        for (int64_t j = 0; j < h->num_var; ++j) x[j] = sc * ray[j];
        status = ktn_separate(h, x, &nc, &nz, &er);
        if (status < 0) break;
        if (nc > 0 || status == KTN_NUMERIC_NONFINITE) { if (n_hit) *n_hit = n; break; }      /* !allsat: stop searching in this direction */
    }
    free(x);
    if (n_cuts) *n_cuts = nc;
    if (nnz_cuts) *nnz_cuts = nz;
    if (err_row) *err_row = er;
    return status;
}

/* device-resident and sharded entry points have no CPU meaning */
int ktn_set_stream(ktn_handle* h, void* s) { (void)s; return fail(h, KTN_ERR_UNSUPPORTED, "oracle has no stream"); }
int ktn_separate_device_async(ktn_handle* h, const double* d) { (void)d; return fail(h, KTN_ERR_UNSUPPORTED, "oracle has no device path"); }
int ktn_sync_counts(ktn_handle* h, int64_t* a, int64_t* b, int64_t* c) { (void)a; (void)b; (void)c; return fail(h, KTN_ERR_UNSUPPORTED, "oracle has no device path"); }
int ktn_comm_unique_id(void* id) { (void)id; return KTN_ERR_UNSUPPORTED; }
int ktn_comm_init(ktn_handle* h, int32_t n, int32_t r, const void* id) { (void)n; (void)r; (void)id; return fail(h, KTN_ERR_UNSUPPORTED, "oracle is single process"); }
int ktn_set_row_offset(ktn_handle* h, int64_t r) { (void)r; return fail(h, KTN_ERR_UNSUPPORTED, "oracle is single process"); }
int ktn_allgather_cuts_async(ktn_handle* h) { return fail(h, KTN_ERR_UNSUPPORTED, "oracle is single process"); }
int ktn_exchange_transport(ktn_handle* h) { (void)h; return 0; }
int ktn_sync_gathered(ktn_handle* h, int64_t* a, int64_t* b) { (void)a; (void)b; return fail(h, KTN_ERR_UNSUPPORTED, "oracle is single process"); }
int ktn_gathered_error_row(ktn_handle* h, int64_t* e) { (void)e; return fail(h, KTN_ERR_UNSUPPORTED, "oracle is single process"); }
int ktn_fetch_gathered(ktn_handle* h, int64_t* a, int64_t* b, int32_t* c, double* d, double* e, double* f, double* g, double* v, double* w) {
    (void)w; (void)a; (void)b; (void)c; (void)d; (void)e; (void)f; (void)g; (void)v; return fail(h, KTN_ERR_UNSUPPORTED, "oracle is single process"); }
