// ktn_family.h -- row evaluators of the shape FAMILIES (KTN_FAM_* in ktn_program.h).
//
// A family is a program pattern the tape compiler recognises exactly (ktn_compile.cpp, "family detection").
// For those shapes the round kernels do not interpret the program: they run the functions below, which perform
// the SAME fp64 operations in the SAME order as the shape's program (the program stays the definition; the
// CPU suite runs both through tests/emu and checks them against the oracle bit for bit).
// Reference: these replace forward_eval / reverse_eval of the per-constraint tape behind eval_g / eval_jac_g
// (called at src/separators.jl:112-113), plus linear_oa_cut (src/algorithms.jl:3-18), round_coefs
// (src/model.jl:200-207) and _addcut's finiteness test (src/model.jl:69) for the selected rows.
//
// Data of one family row (chunk blob, lane stride L):
//   constants   two per unique variable u (LSE: c_u = slot 2u, d_u = slot 2u+1;  QUAD: a_u = slot u, b_u = slot nu+u;  SOC: s_u = slot u,
//               none for the linear variable); rows of
//               <= 16 unique variables store them as PAIRS (p0_u, p1_u), 16 bytes per row and variable (ktn_program.h)
//   cols        column of unique variable u (first-occurrence order = the order of the terms)
//   rank        position of unique variable u among the row's ascending columns = its Jacobian entry index;
//               nu <= 16: ONE 64-bit word per row, 4 bits per u;  nu > 16: one byte per u
//
// The work is split between the two kernels of a round (ktn_kernels.cu):
//   K1  forward<N>: the whole row in registers, every constant and column requested at once; g, the violation test.  Nothing
//       of the row is kept for the cut: a selected row leaves one 32-byte record {g, aux, lb, ub}.
//   K2  ktn_family_cut_terms (the cut kernel): one thread per SELECTED row recomputes the row's terms from the blob and x*,
//       scatters coefficient and product to their Jacobian entry (rank word), and accumulates the constant b = g + sum -x_q J_q
//       in entry order, as the reference does.
// Rows with more than 16 unique variables take the streaming fallbacks (cut built in K1).
#ifndef KTN_FAMILY_H
#define KTN_FAMILY_H
#include "ktn_interp.h"

#define KTN_FAM_DMAX 1.7976931348623157e308

template <int N> struct KtnFamRegs { double p0[N], p1[N], x[N]; };   // the row's constants and gathered x*; forward may overwrite p1

template <int FAM> struct KtnFamily;

// log(sum_u exp(c_u * x_u + d_u))
// Program: KF_TERMS(EXP_AFF, FIRST); STORE S; LOG | KR_ONE; MULRCP S; STORE R1; KR_TERMS(EXP_AFF); END
template <> struct KtnFamily<KTN_FAM_LSE> {
    static KTN_HDM double arg(double c, double d, double x) { return (0.0 + c * x) + d; }      // LOAD c; MUL x; ADDZ; ADD d
    // forward over the N register-resident terms; aux = the sum (the cut's adjoint is its reciprocal)
    template <int N> static KTN_HDM double forward(KtnFamRegs<N>& r, double& aux) {
        double a[N];
        bool slow = false;
#pragma unroll
        for (int u = 0; u < N; ++u) a[u] = arg(r.p0[u], r.p1[u], r.x[u]);
#pragma unroll
        for (int u = 0; u < N; ++u) r.p1[u] = ktn_exp_fast(a[u]);      // branch-free: N independent chains
#pragma unroll
        for (int u = 0; u < N; ++u) slow = slow || !ktn_exp_is_fast(a[u]);
        if (slow) {
#pragma unroll
            for (int u = 0; u < N; ++u) if (!ktn_exp_is_fast(a[u])) r.p1[u] = ktn_exp_slow(a[u]);
        }
        double acc = 0.0;        // n-ary sum starts from zero(T): first step is 0.0 + e_0
#pragma unroll
        for (int u = 0; u < N; ++u) acc = acc + r.p1[u];
        aux = acc;
        return ktn_log(acc);
    }
    static KTN_HDM double adjoint(double aux) { return revmul(1.0, 1.0 / aux); }                  // KR_ONE; KR_MULRCP S
    static KTN_HDM double jac(double adj, double c, double e, double) { return 0.0 + revmul(revmul(adj, e), c); }
    static KTN_HDM double jac_plain(double adj, double c, double e, double) { return 0.0 + (adj * e) * c; }
    // the cut evaluates the exponentials again, with the forward pass's operations: eight at a time, branch-free (p1: d -> exp)
    static KTN_HDM void pre8(const double (&p0)[8], double (&p1)[8], const double (&x)[8]) {
        double a[8];
        bool slow = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = arg(p0[k], p1[k], x[k]);
#pragma unroll
        for (int k = 0; k < 8; ++k) p1[k] = ktn_exp_fast(a[k]);
#pragma unroll
        for (int k = 0; k < 8; ++k) slow = slow || !ktn_exp_is_fast(a[k]);
        if (slow) {
#pragma unroll
            for (int k = 0; k < 8; ++k) if (!ktn_exp_is_fast(a[k])) p1[k] = ktn_exp_slow(a[k]);
        }
    }
    static KTN_HDM double entry(double adj, double c, double e, double x, bool exact, uint32_t, uint32_t) { return exact ? jac(adj, c, e, x) : jac_plain(adj, c, e, x); }
    // streaming fallback (any nu)
    template <class R> static KTN_HDM double forward_stream(const R& r, double& aux) {
        double acc = 0.0;
        for (uint32_t u = 0; u < r.nu; ++u) acc = acc + ktn_exp(arg(r.cst(2 * u), r.cst(2 * u + 1), r.x(u)));
        aux = acc;
        return ktn_log(acc);
    }
    template <class R> static KTN_HDM double jac_stream(const R& r, uint32_t u, double adj) {
        const double c = r.cst(2 * u);
        return jac(adj, c, ktn_exp(arg(c, r.cst(2 * u + 1), r.x(u))), 0.0);
    }
};

// sum_u a_u * x_u^2 + sum_u b_u * x_u
// Program: KF_TERMS(MULC_SQ, FIRST); KF_TERMS(MULC_X) | KR_ONE; STORE R1; KR_TERMS(MULC_SQ); KR_TERMS(MULC_X, JACC); END
template <> struct KtnFamily<KTN_FAM_QUAD> {
    template <int N> static KTN_HDM double forward(KtnFamRegs<N>& r, double& aux) {
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < N; ++u) acc = acc + (r.x[u] * r.x[u]) * r.p0[u];
#pragma unroll
        for (int u = 0; u < N; ++u) acc = acc + r.p1[u] * r.x[u];
        aux = 0.0;
        return acc;
    }
    static KTN_HDM double adjoint(double) { return 1.0; }                                         // KR_ONE
    static KTN_HDM double jac(double adj, double a, double b, double x) { return (0.0 + revmul(revmul(adj, a), 2.0 * x)) + revmul(adj, b); }
    static KTN_HDM double jac_plain(double adj, double a, double b, double x) { return (0.0 + (adj * a) * (2.0 * x)) + adj * b; }
    static KTN_HDM void pre8(const double (&)[8], double (&)[8], const double (&)[8]) {}
    static KTN_HDM double entry(double adj, double a, double b, double x, bool exact, uint32_t, uint32_t) { return exact ? jac(adj, a, b, x) : jac_plain(adj, a, b, x); }
    template <class R> static KTN_HDM double forward_stream(const R& r, double& aux) {
        double acc = 0.0;
        for (uint32_t u = 0; u < r.nu; ++u) { const double x = r.x(u); acc = acc + (x * x) * r.cst(u); }
        for (uint32_t u = 0; u < r.nu; ++u) acc = acc + r.cst(r.nu + u) * r.x(u);
        aux = 0.0;
        return acc;
    }
    template <class R> static KTN_HDM double jac_stream(const R& r, uint32_t u, double adj) { return jac(adj, r.cst(u), r.cst(r.nu + u), r.x(u)); }
};

// sqrt(sum_{u < nu-1} (s_u * x_u)^2) - x_{nu-1}
// Program: KF_TERMS(SQ_MULC, FIRST); SQRT; STORE S; SUB x_t | KR_ONE; STORE R1; MULHRCP S; STORE R2; KR_TERMS(SQ_MULC); LOAD R1; NEG; JSET t; END
// At a point where every squared term vanishes the root is 0, its partial 0.5 / 0 is infinite and the coefficients are NaN: the
// row ends the batch, as in the reference (src/model.jl:69-73; test/3d.jl:153-171 starts away from the apex for that reason).
template <> struct KtnFamily<KTN_FAM_SOC> {
    template <int N> static KTN_HDM double forward(KtnFamRegs<N>& r, double& aux) {
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < N - 1; ++u) { const double q = r.p0[u] * r.x[u]; acc = acc + q * q; }
        const double s = ktn_sqrt(acc);
        aux = s;
        return s - r.x[N - 1];
    }
    static KTN_HDM double adjoint(double aux) { return revmul(1.0, 0.5 / aux); }                  // KR_ONE; KR_MULHRCP S
    static KTN_HDM void pre8(const double (&)[8], double (&)[8], const double (&)[8]) {}
    static KTN_HDM double entry(double adj, double s, double, double x, bool exact, uint32_t u, uint32_t nu) {
        if (u + 1 == nu) return 0.0 + (-1.0);                                                     // LOAD R1 (= 1); NEG; JSET
        return exact ? 0.0 + revmul(revmul(adj, 2.0 * (s * x)), s) : 0.0 + (adj * (2.0 * (s * x))) * s;
    }
    template <class R> static KTN_HDM double forward_stream(const R& r, double& aux) {
        double acc = 0.0;
        for (uint32_t u = 0; u + 1 < r.nu; ++u) { const double q = r.cst(u) * r.x(u); acc = acc + q * q; }
        const double s = ktn_sqrt(acc);
        aux = s;
        return s - r.x(r.nu - 1);
    }
    template <class R> static KTN_HDM double jac_stream(const R& r, uint32_t u, double adj) {
        if (u + 1 == r.nu) return 0.0 + (-1.0);
        const double s = r.cst(u);
        return 0.0 + revmul(revmul(adj, 2.0 * (s * r.x(u))), s);
    }
};

// NaN-skipping max / min (one body for host and device: the rounding decision below must not depend on the compiler's fmax)
KTN_HDM double ktn_dmax(double a, double b) { return a > b ? a : (b != b ? a : b); }
KTN_HDM double ktn_dmin(double a, double b) { return a < b ? a : (b != b ? a : b); }

// ---- the cut of a selected row (the cut kernel; tests/emu) --------------------------------------------------------------
// Row context R: pairs2(g, a0, a1, b0, b1) = the constants of unique variables 2g (a) and 2g + 1 (b); cols8(g, c[8]) = the
// columns of unique variables 8g .. 8g + 7; xat(col).  `rw`: 4 bits per unique variable u = its Jacobian entry index (rank).
// Sink S: put(q, J, col): coefficient and column of entry q, get(q) / set(q, J);  put_t(q, t) / get_t(q): the product -x* J of entry q.
// The terms are walked in TERM order, eight at a time (their loads are in flight together: one 256-bit load per two terms'
// constants, one per eight columns, then the eight x* gathers); coefficients and products are scattered to their entry index,
// and the constant b = g; b += -x*_q J_q then accumulates in ENTRY order, as the reference does (src/algorithms.jl:8-16).
// round_coefs (src/model.jl:200-207); returns true when a coefficient is not finite (src/model.jl:69).
//
// reverse_eval's product rule revmul(a, p) equals a * p whenever a * p is not NaN, and NaN operands stay NaN through the
// later products, so the coefficients are first formed with plain multiplications; only a row in which one of them came out
// NaN repeats the sweep with the exact rule.  round_coefs zeroes J when J + rng < maximum(J): J + rng is monotone in J, so when
// the smallest coefficient passes (and everything is finite) all pass, and the second sweep only runs for rows that need it.
template <int FAM, class R, class S>
KTN_HDM bool ktn_family_cut_terms(const R& r, uint32_t nu, uint64_t rw, S& s, double g, double aux, bool do_round, double rng, double& b_out) {
    typedef KtnFamily<FAM> F;
    const double adj = F::adjoint(aux);
    double mx, mn;
    bool anynan, exact = false;
    for (;;) {
        mx = -ktn_inf(); mn = ktn_inf(); anynan = false;
        for (uint32_t u0 = 0; u0 < nu; u0 += 8) {
            int32_t c[8]; double p0[8], p1[8], x[8];
            r.cols8(u0 >> 3, c);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (u0 + 2 * k < nu) r.pairs2((u0 >> 1) + k, p0[2 * k], p1[2 * k], p0[2 * k + 1], p1[2 * k + 1]);
                else { p0[2 * k] = p1[2 * k] = p0[2 * k + 1] = p1[2 * k + 1] = 0.0; }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = r.xat(c[k]);          // columns past the row's end are padded with 0: a valid index
            F::pre8(p0, p1, x);                                      // (padding: exp(0 * x + 0) = 1, never used)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (u0 + k < nu) {
                    const double jv = F::entry(adj, p0[k], p1[k], x[k], exact, u0 + k, nu);
                    const uint32_t q = (uint32_t)(rw >> (4 * (u0 + k))) & 15u;
                    s.put(q, jv, c[k]); s.put_t(q, (-x[k]) * jv);
                    mx = ktn_dmax(mx, jv); mn = ktn_dmin(mn, jv); anynan = anynan || (jv != jv);
                }
            }
        }
        if (!anynan || exact) break;
        exact = true;                                       // rare: repeat the sweep with reverse_eval's exact product rule
    }
    double b = g;
    for (uint32_t q = 0; q < nu; ++q) b = b + s.get_t(q);
    b_out = b;
    bool bad = false;
    if (anynan || !(ktn_fabs(mn) <= KTN_FAM_DMAX) || !(ktn_fabs(mx) <= KTN_FAM_DMAX) || (do_round && (mn + rng < mx))) {
        if (anynan) mx = ktn_nan();     // Julia's maximum() propagates NaN
        for (uint32_t q = 0; q < nu; ++q) {
            double c = s.get(q);
            if (do_round && (c + rng < mx)) c = 0.0;
            bad = bad || !(ktn_fabs(c) <= KTN_FAM_DMAX);
            s.set(q, c);
        }
    }
    return bad;
}

// ---- streaming fallback for rows with more than KTN_FAM_REGS unique variables (cut built where the row is evaluated) ------
// Row context R adds x(u), rank(u), nu.  Sink S: put_j / get_j on the coefficient row, and xsorted(q) = x* of the q-th ascending column.
template <int FAM, class R, class S>
KTN_HDM bool ktn_family_cut_stream(const R& r, S& s, double g, double aux, bool do_round, double rng, double& b_out) {
    typedef KtnFamily<FAM> F;
    const uint32_t nu = r.nu;
    const double adj = F::adjoint(aux);
    double mx = -ktn_inf(), mn = ktn_inf();
    bool anynan = false;
    for (uint32_t u = 0; u < nu; ++u) {
        const double jv = F::jac_stream(r, u, adj);
        s.put_j(r.rank(u), jv);
        mx = jv > mx ? jv : mx; mn = jv < mn ? jv : mn; anynan = anynan || (jv != jv);
    }
    double b = g;
    for (uint32_t q = 0; q < nu; ++q) b = b + (-s.xsorted(q)) * s.get_j(q);
    b_out = b;
    bool bad = false;
    if (anynan || !(ktn_fabs(mn) <= KTN_FAM_DMAX) || !(ktn_fabs(mx) <= KTN_FAM_DMAX) || (do_round && (mn + rng < mx))) {
        if (anynan) mx = ktn_nan();
        for (uint32_t q = 0; q < nu; ++q) {
            double jv = s.get_j(q);
            if (do_round && (jv + rng < mx)) jv = 0.0;
            bad = bad || !(ktn_fabs(jv) <= KTN_FAM_DMAX);
            s.put_j(q, jv);
        }
    }
    return bad;
}

#endif
