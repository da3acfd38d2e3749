// ktn_family.h -- row evaluators of the shape FAMILIES (KTN_FAM_* in ktn_program.h).
//
// A family is a program pattern the tape compiler recognises exactly (ktn_compile.cpp, "family detection").
// For those shapes the round kernel does not interpret the program: it runs the functions below, which perform
// the SAME fp64 operations in the SAME order as the shape's program (the program stays the definition; the
// CPU suite runs both through tests/emu and checks them against the oracle bit for bit).
// Reference: these replace forward_eval / reverse_eval of the per-constraint tape behind eval_g / eval_jac_g
// (called at src/separators.jl:112-113), plus linear_oa_cut (src/algorithms.jl:3-18), round_coefs
// (src/model.jl:200-207) and _addcut's finiteness test (src/model.jl:69) for the selected rows.
//
// Data of one family row (chunk blob, lane stride L):
//   constants   two per unique variable u (LSE: c_u = slot 2u, d_u = slot 2u+1;  QUAD: a_u = slot u, b_u = slot nu+u)
//   cols        column of unique variable u (first-occurrence order = the order of the terms)
//   rank        position of unique variable u among the row's ascending columns = its Jacobian entry index;
//               nu <= 16: ONE 64-bit word per row, 4 bits per u;  nu > 16: one byte per u
//
// Fast path (nu <= KTN_FAM_REGS, one instantiation per nu): the whole row lives in registers.  ktn_family_forward loads every constant
// and column at once (one memory round trip), gathers x*, evaluates g.  ktn_family_cut builds the cut from the
// same registers: no value is read twice.  Rows with more unique variables take the streaming fallbacks.
#ifndef KTN_FAMILY_H
#define KTN_FAMILY_H
#include "ktn_interp.h"

#define KTN_FAM_DMAX 1.7976931348623157e308
#ifndef KTN_FWD_GROUP
#define KTN_FWD_GROUP 8            // terms in flight per row in the evaluation-only forward pass
#endif

template <int N> struct KtnFamRegs { double p0[N], p1[N], x[N]; };   // LSE: c, exp value;  QUAD: a, b

template <int FAM> struct KtnFamily;

// log(sum_u exp(c_u * x_u + d_u))
// Program: KF_TERMS(EXP_AFF, FIRST); STORE S; LOG | KR_ONE; MULRCP S; STORE R1; KR_TERMS(EXP_AFF); END
template <> struct KtnFamily<KTN_FAM_LSE> {
    static const bool KEEP_P1 = true;     // p1 holds exp(c x + d) after the forward pass: kept for the cut
    static KTN_HDM uint32_t slot0(uint32_t u, uint32_t) { return 2 * u; }
    static KTN_HDM uint32_t slot1(uint32_t u, uint32_t) { return 2 * u + 1; }
    static KTN_HDM double arg(double c, double d, double x) { return (0.0 + c * x) + d; }      // LOAD c; MUL x; ADDZ; ADD d
    // forward over the N register-resident terms
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
    // evaluation only (ktn_eval_g): the same operations in the same order as forward<N>, but the terms are loaded and
    // consumed in groups of KTN_FWD_GROUP, so a row never holds more than one group in registers.  `hook` runs once, behind
    // the first group's loads (the kernel draws its next work ticket there).
    template <int N, class R, class H> static KTN_HDM double forward_only(const R& r, H hook) {
        double acc = 0.0;
#pragma unroll
        for (int u0 = 0; u0 < N; u0 += KTN_FWD_GROUP) {
            double a[KTN_FWD_GROUP], e[KTN_FWD_GROUP];
            {
                double c[KTN_FWD_GROUP], d[KTN_FWD_GROUP]; int32_t col[KTN_FWD_GROUP];
#pragma unroll
                for (int k = 0; k < KTN_FWD_GROUP; ++k) if (u0 + k < N) { c[k] = r.cst(2 * (u0 + k)); d[k] = r.cst(2 * (u0 + k) + 1); col[k] = r.col(u0 + k); }
                if (u0 == 0) hook();
#pragma unroll
                for (int k = 0; k < KTN_FWD_GROUP; ++k) if (u0 + k < N) a[k] = arg(c[k], d[k], r.xat(col[k]));
            }
            bool slow = false;
#pragma unroll
            for (int k = 0; k < KTN_FWD_GROUP; ++k) if (u0 + k < N) e[k] = ktn_exp_fast(a[k]);
#pragma unroll
            for (int k = 0; k < KTN_FWD_GROUP; ++k) if (u0 + k < N) slow = slow || !ktn_exp_is_fast(a[k]);
            if (slow) {
#pragma unroll
                for (int k = 0; k < KTN_FWD_GROUP; ++k) if (u0 + k < N) if (!ktn_exp_is_fast(a[k])) e[k] = ktn_exp_slow(a[k]);
            }
#pragma unroll
            for (int k = 0; k < KTN_FWD_GROUP; ++k) if (u0 + k < N) acc = acc + e[k];
        }
        return ktn_log(acc);
    }
    static KTN_HDM double adjoint(double aux) { return revmul(1.0, 1.0 / aux); }                  // KR_ONE; KR_MULRCP S
    static KTN_HDM double jac(double adj, double c, double e, double) { return 0.0 + revmul(revmul(adj, e), c); }
    static KTN_HDM double jac_plain(double adj, double c, double e, double) { return 0.0 + (adj * e) * c; }
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
    static const bool KEEP_P1 = false;
    static KTN_HDM uint32_t slot0(uint32_t u, uint32_t) { return u; }
    static KTN_HDM uint32_t slot1(uint32_t u, uint32_t nu) { return nu + u; }
    template <int N> static KTN_HDM double forward(KtnFamRegs<N>& r, double& aux) {
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < N; ++u) acc = acc + (r.x[u] * r.x[u]) * r.p0[u];
#pragma unroll
        for (int u = 0; u < N; ++u) acc = acc + r.p1[u] * r.x[u];
        aux = 0.0;
        return acc;
    }
    template <int N, class R, class H> static KTN_HDM double forward_only(const R& r, H hook) {      // see KtnFamily<KTN_FAM_LSE>
        double x[N], acc = 0.0;
#pragma unroll
        for (int u0 = 0; u0 < N; u0 += KTN_FWD_GROUP) {
            double a[KTN_FWD_GROUP]; int32_t col[KTN_FWD_GROUP];
#pragma unroll
            for (int k = 0; k < KTN_FWD_GROUP; ++k) if (u0 + k < N) { a[k] = r.cst(u0 + k); col[k] = r.col(u0 + k); }
            if (u0 == 0) hook();
#pragma unroll
            for (int k = 0; k < KTN_FWD_GROUP; ++k) if (u0 + k < N) x[u0 + k] = r.xat(col[k]);
#pragma unroll
            for (int k = 0; k < KTN_FWD_GROUP; ++k) if (u0 + k < N) acc = acc + (x[u0 + k] * x[u0 + k]) * a[k];
        }
#pragma unroll
        for (int u0 = 0; u0 < N; u0 += KTN_FWD_GROUP) {
            double b[KTN_FWD_GROUP];
#pragma unroll
            for (int k = 0; k < KTN_FWD_GROUP; ++k) if (u0 + k < N) b[k] = r.cst(N + u0 + k);
#pragma unroll
            for (int k = 0; k < KTN_FWD_GROUP; ++k) if (u0 + k < N) acc = acc + b[k] * x[u0 + k];
        }
        return acc;
    }
    static KTN_HDM double adjoint(double) { return 1.0; }                                         // KR_ONE
    static KTN_HDM double jac(double adj, double a, double b, double x) { return (0.0 + revmul(revmul(adj, a), 2.0 * x)) + revmul(adj, b); }
    static KTN_HDM double jac_plain(double adj, double a, double b, double x) { return (0.0 + (adj * a) * (2.0 * x)) + adj * b; }
    template <class R> static KTN_HDM double forward_stream(const R& r, double& aux) {
        double acc = 0.0;
        for (uint32_t u = 0; u < r.nu; ++u) { const double x = r.x(u); acc = acc + (x * x) * r.cst(u); }
        for (uint32_t u = 0; u < r.nu; ++u) acc = acc + r.cst(r.nu + u) * r.x(u);
        aux = 0.0;
        return acc;
    }
    template <class R> static KTN_HDM double jac_stream(const R& r, uint32_t u, double adj) { return jac(adj, r.cst(u), r.cst(r.nu + u), r.x(u)); }
};

// ---- register-resident rows -------------------------------------------------------------------------------
// Row context R: cst(i), col(u), xat(col), rankword() (register-resident rows) / rank(u) (streaming rows).
// ktn_family_load requests every constant and column id of the row at once (the kernel reads them from the shared-memory
// slot a bulk copy filled, and gives the slot back right after); ktn_family_eval gathers x* and evaluates g.
template <int FAM, int N, class R>
KTN_HDM void ktn_family_load(const R& r, KtnFamRegs<N>& v, int32_t (&col)[N]) {
    typedef KtnFamily<FAM> F;
#pragma unroll
    for (int u = 0; u < N; ++u) { v.p0[u] = r.cst(F::slot0(u, N)); v.p1[u] = r.cst(F::slot1(u, N)); col[u] = r.col(u); }
}
template <int FAM, int N, class R>
KTN_HDM double ktn_family_eval(const R& r, KtnFamRegs<N>& v, const int32_t (&col)[N], double& aux) {
#pragma unroll
    for (int u = 0; u < N; ++u) v.x[u] = r.xat(col[u]);
    return KtnFamily<FAM>::template forward<N>(v, aux);
}

KTN_HDM double ktn_dmax(double a, double b) {
#if defined(__CUDA_ARCH__)
    return fmax(a, b);
#else
    return a > b ? a : (b != b ? a : b);
#endif
}
KTN_HDM double ktn_dmin(double a, double b) {
#if defined(__CUDA_ARCH__)
    return fmin(a, b);
#else
    return a < b ? a : (b != b ? a : b);
#endif
}

// Cut row from the registers ktn_family_eval left behind.  Sink S: t(q) scratch cells (one per Jacobian entry) and the
// coefficient row out[0..nu).  Coefficients and the products -x_u * J_u are computed in term order and scattered to
// their Jacobian entry index; the constant is then accumulated sequentially in entry order, as the reference does
// (b = g; b += -xstar[col] * partial).  Returns true when a coefficient is not finite.
//
// reverse_eval's product rule revmul(a, p) equals a * p whenever a * p is not NaN, and NaN operands stay NaN through the
// later products, so the coefficients are first formed with plain multiplications (`plain`); only a row in which one of
// them came out NaN repeats the sweep with the exact rule.  min / max skip NaN (tracked separately).
template <int FAM, int N, class R, class S>
KTN_HDM bool ktn_family_cut(const R& r, const KtnFamRegs<N>& v, S& s, double g, double aux, bool do_round, double rng, double& b_out) {
    typedef KtnFamily<FAM> F;
    const double adj = F::adjoint(aux);
    const uint64_t rw = r.rankword();                       // 4 bits per unique variable: its Jacobian entry index
    double p0[N], p1[N];
#pragma unroll
    for (int u = 0; u < N; ++u) { p0[u] = r.cst(F::slot0(u, N)); p1[u] = F::KEEP_P1 ? v.p1[u] : r.cst(F::slot1(u, N)); }   // constants are re-read, not kept
    double mx = -ktn_inf(), mn = ktn_inf();
    bool anynan = false;
#pragma unroll
    for (int u = 0; u < N; ++u) {
        const double jv = F::jac_plain(adj, p0[u], p1[u], v.x[u]);
        const uint32_t q = (uint32_t)(rw >> (4 * u)) & 15u;
        s.put_t(q, (-v.x[u]) * jv);
        s.put_j(q, jv);
        mx = ktn_dmax(mx, jv); mn = ktn_dmin(mn, jv); anynan = anynan || (jv != jv);
    }
    if (anynan) {                                           // rare: repeat the sweep with reverse_eval's exact product rule
        anynan = false; mx = -ktn_inf(); mn = ktn_inf();
#pragma unroll
        for (int u = 0; u < N; ++u) {
            const double jv = F::jac(adj, p0[u], p1[u], v.x[u]);
            const uint32_t q = (uint32_t)(rw >> (4 * u)) & 15u;
            s.put_t(q, (-v.x[u]) * jv);
            s.put_j(q, jv);
            mx = ktn_dmax(mx, jv); mn = ktn_dmin(mn, jv); anynan = anynan || (jv != jv);
        }
    }
    double b = g;
#pragma unroll
    for (int k = 0; k < N; ++k) b = b + s.get_t(k);
    b_out = b;
    // round_coefs zeroes jv when jv + rng < maximum(coefs).  jv + rng is monotone in jv, so when the smallest coefficient
    // passes (and everything is finite) all pass: the exact second sweep only runs for rows that need it.
    bool bad = false;
    if (anynan || !(ktn_fabs(mn) <= KTN_FAM_DMAX) || !(ktn_fabs(mx) <= KTN_FAM_DMAX) || (do_round && (mn + rng < mx))) {
        if (anynan) mx = ktn_nan();     // Julia's maximum() propagates NaN
        for (uint32_t k = 0; k < (uint32_t)N; ++k) {
            double c = s.get_j(k);
            if (do_round && (c + rng < mx)) c = 0.0;
            bad = bad || !(ktn_fabs(c) <= KTN_FAM_DMAX);
            s.put_j(k, c);
        }
    }
    return bad;
}

// ---- streaming fallback for rows with more than KTN_FAM_REGS unique variables -----------------------------
// Sink S here: put_j / get_j on the coefficient row, and xsorted(q) = x* of the q-th ascending column.
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
