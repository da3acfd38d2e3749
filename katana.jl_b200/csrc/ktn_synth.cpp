// ktn_synth.cpp -- deterministic synthetic instances.
// Test / bench support object of the library: C, Python and Julia callers see the same
// instance.  A row depends only on (kind, seed, num_var, global row index).
//   kind 0  sparse convex QCQP   g_i = sum_k a_ik x_jk^2 + sum_k b_ik x_jk          8 distinct columns
//   kind 1  log-sum-exp          g_i = log sum_{k<K_i} exp(a_ik x_jk + b_ik)        K_i in {4..16}
//   kind 2  SOC-like risk row    g_i = sqrt(sum_{k<8} (s_ik x_jk)^2) - x_t           as test/3d.jl:161
//   kind 3  portfolio            nine sparse LINEAR rows sum_{k<16} a_ik x_jk (not in nlconstr_ixs: never separated) for every
//                                SOC-like row of kind 2 (rows with index = 9 mod 10)                       BASELINE.json configs[3]
// Expressions are emitted the way JuMP's @NLconstraint parser builds them (n-ary +, binary *, ^).
#include <cstdint>
#include <cmath>
#include <vector>
#include "../../include/ktn.h"

namespace {
struct Rng {
    uint64_t s;
    uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    double uni(double a, double b) { return a + (b - a) * (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};
static Rng row_rng(uint64_t seed, int64_t row) { Rng r{seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(row + 1))}; r.next(); return r; }
static void distinct_cols(Rng& r, int64_t n, int k, int32_t* out) {
    for (int i = 0; i < k; ++i) {
        for (;;) {
            int32_t c = (int32_t)(r.next() % (uint64_t)n);
            bool dup = false;
            for (int j = 0; j < i; ++j) dup = dup || out[j] == c;
            if (!dup) { out[i] = c; break; }
        }
    }
}
struct Emit {
    int32_t* op; int32_t* arg; double* val; int64_t n = 0;
    void call(int o, int nc) { if (op) { op[n] = o; arg[n] = nc; val[n] = 0.0; } ++n; }
    void var(int32_t c) { if (op) { op[n] = KTN_OP_VAR; arg[n] = c; val[n] = 0.0; } ++n; }
    void cst(double v) { if (op) { op[n] = KTN_OP_CONST; arg[n] = 0; val[n] = v; } ++n; }
};
static int lse_terms(uint64_t seed, int64_t row) { Rng r = row_rng(seed ^ 0x5151, row); return 4 + (int)(r.next() % 13u); }
}  // namespace

extern "C" int ktn_synth_rows(int32_t kind, uint64_t seed, int64_t num_var, int64_t row_begin, int64_t nrows,
                              int64_t* n_nodes, int64_t* expr_ptr, int32_t* op, int32_t* arg, double* val,
                              double* lb, double* ub, uint8_t* flags) {
    if (kind < 0 || kind > 3 || num_var < 17 || nrows < 0 || !n_nodes) return KTN_ERR_USAGE;
    Emit e{op, arg, val};
    int32_t cols[17];
    for (int64_t r = 0; r < nrows; ++r) {
        const int64_t row = row_begin + r;
        if (op) { expr_ptr[r] = e.n; lb[r] = -INFINITY; ub[r] = 0.0; flags[r] = KTN_ROW_NL; }
        Rng g = row_rng(seed, row);
        if (kind == 0) {
            distinct_cols(g, num_var, 8, cols);
            double a[8], b[8];
            for (int k = 0; k < 8; ++k) { a[k] = g.uni(0.5, 1.5); b[k] = g.uni(-1.0, 1.0); }
            e.call(KTN_OP_ADD, 16);
            for (int k = 0; k < 8; ++k) { e.call(KTN_OP_MUL, 2); e.cst(a[k]); e.call(KTN_OP_POW, 2); e.var(cols[k]); e.cst(2.0); }
            for (int k = 0; k < 8; ++k) { e.call(KTN_OP_MUL, 2); e.cst(b[k]); e.var(cols[k]); }
        } else if (kind == 1) {
            const int K = lse_terms(seed, row);
            distinct_cols(g, num_var, K, cols);
            e.call(KTN_OP_LOG, 1);
            e.call(KTN_OP_ADD, K);
            for (int k = 0; k < K; ++k) {
                const double a = g.uni(-1.0, 1.0), b = g.uni(-1.0, 1.0);
                e.call(KTN_OP_EXP, 1); e.call(KTN_OP_ADD, 2); e.call(KTN_OP_MUL, 2); e.cst(a); e.var(cols[k]); e.cst(b);
            }
        } else if (kind == 3 && row % 10 != 9) {
            if (op) flags[r] = 0;                       // a linear row: copied into the LP at loadproblem! (src/model.jl:115-118), not tested per round
            distinct_cols(g, num_var, 16, cols);
            e.call(KTN_OP_ADD, 16);
            for (int k = 0; k < 16; ++k) { e.call(KTN_OP_MUL, 2); e.cst(g.uni(-1.0, 1.0)); e.var(cols[k]); }
        } else {
            distinct_cols(g, num_var, 9, cols);
            e.call(KTN_OP_SUB, 2);
            e.call(KTN_OP_SQRT, 1);
            e.call(KTN_OP_ADD, 8);
            for (int k = 0; k < 8; ++k) { e.call(KTN_OP_POW, 2); e.call(KTN_OP_MUL, 2); e.cst(g.uni(0.1, 0.5)); e.var(cols[k]); e.cst(2.0); }
            e.var(cols[8]);
        }
    }
    if (op) expr_ptr[nrows] = e.n;
    *n_nodes = e.n;
    return KTN_OK;
}

extern "C" int ktn_synth_point(int32_t kind, uint64_t seed, int64_t num_var, double* x) {
    if (kind < 0 || kind > 3 || !x) return KTN_ERR_USAGE;
    const double lo = kind == 0 ? -1.0 : kind == 1 ? -2.0 : 0.05, hi = kind == 0 ? 1.0 : kind == 1 ? 2.0 : 1.0;
    for (int64_t j = 0; j < num_var; ++j) { Rng r = row_rng(seed ^ 0xA5A5A5A5ull, j); x[j] = r.uni(lo, hi); }
    return KTN_OK;
}
