// ktn_interp.h -- the tape interpreter core (ISA in ktn_program.h), shared by the sm_100a kernels
// and by the test-only host emulator (tests/emu/) that checks compiler output without a GPU.
// Restates forward_eval / reverse_eval of the reference's tape evaluator (called at
// src/separators.jl:112-113) as an accumulator machine; arithmetic order matches oracle/ktn_oracle.c.
#ifndef KTN_INTERP_H
#define KTN_INTERP_H
#include "ktn_math.h"
#include "ktn_program.h"

struct KtnInsWord { uint32_t x, y, z, w; };
#if defined(__CUDACC__)
#define KTN_HDM __host__ __device__ __forceinline__
#define KTN_HD_NOINLINE __host__ __device__ __noinline__
#else
#define KTN_HDM inline
#define KTN_HD_NOINLINE static inline
#endif
#if defined(__CUDA_ARCH__)
#define KTN_ANY_LANE(mask, pred) __any_sync(mask, pred)
#else
#define KTN_ANY_LANE(mask, pred) (true)   /* host emulation runs lane by lane: never take the warp-uniform skip */
#endif
KTN_HD KtnInsWord ktn_fetch_ins(const KtnIns* p) {
#if defined(__CUDA_ARCH__)
    const uint4 w = *reinterpret_cast<const uint4*>(p); return KtnInsWord{w.x, w.y, w.z, w.w};   // shared (K1) or global (BIG) program
#else
    KtnInsWord w; memcpy(&w, p, 16); return w;
#endif
}

// reverse_eval's product rule: a * p, except a zero adjoint stays zero under a non-finite partial.
// a * p is NaN in exactly those cases (and when a NaN is genuinely propagated), so only NaN results re-check; the re-check is
// an out-of-line call so that the common path is one multiply, one compare and a branch that is never taken.
KTN_HD_NOINLINE double revmul_nan(double a, double p, double r) { return (a == 0.0 && !ktn_isfinite(p)) ? a : r; }
#ifndef KTN_OPT_REVMUL_CALL
#define KTN_OPT_REVMUL_CALL 1
#endif
KTN_HD double revmul(double a, double p) {
    double r = a * p;
#if KTN_OPT_REVMUL_CALL
    if (r != r) r = revmul_nan(a, p, r);
#else
    if (r != r) r = (a == 0.0 && !ktn_isfinite(p)) ? a : r;
#endif
    return r;
}

// reference forward rule for `^` (ReverseDiffSparse forward_eval): exponent 2 and 1 are special-cased at run time
KTN_HD_NOINLINE double pow_value(double base, double ex) { return ex == 2.0 ? base * base : ex == 1.0 ? base : ktn_pow(base, ex); }
KTN_HD_NOINLINE double pow_dbase(double base, double ex) { return ex == 2.0 ? 2.0 * base : ex == 1.0 ? 1.0 : ex * ktn_pow(base, ex - 1.0); }

// regular chunks: SoA sections with lane stride 32 in shared memory
struct SmemMem {
    double* C; double* S; double* Jp; uint32_t jmul, lane;   // C, S, Jp already include the lane offset
    KTN_HDM double c(uint32_t i) const { return C[i * 32]; }
    KTN_HDM void cst(uint32_t i, double v) { C[i * 32] = v; }
    KTN_HDM double lds(uint32_t i) const { return S[i * 32]; }
    KTN_HDM void sts(uint32_t i, double v) { S[i * 32] = v; }
    KTN_HDM double jld(uint32_t u) const { return Jp[u * jmul]; }
    KTN_HDM void jst(uint32_t u, double v) { Jp[u * jmul] = v; }
    KTN_HDM size_t stride() const { return 32; }
    KTN_HDM size_t jstride() const { return jmul; }
    KTN_HDM double* caddr(uint32_t i) const { return C + i * 32; }
    KTN_HDM double* saddr(uint32_t i) const { return S + i * 32; }
    KTN_HDM double* jaddr(uint32_t u) const { return Jp + u * jmul; }
};
// BIG chunks: constants straight from the blob in global memory, scratch in a global arena
struct GlobalMem {
    const double* C; double* S; double* Jp; size_t jmul; uint32_t lane, L;
    KTN_HDM double c(uint32_t i) const { return C[(size_t)i * L + lane]; }
    KTN_HDM void cst(uint32_t, double) {}   // BIG shapes never alias into the (persistent) global blob
    KTN_HDM double lds(uint32_t i) const { return S[(size_t)i * L + lane]; }
    KTN_HDM void sts(uint32_t i, double v) { S[(size_t)i * L + lane] = v; }
    KTN_HDM double jld(uint32_t u) const { return Jp[(size_t)u * jmul]; }
    KTN_HDM void jst(uint32_t u, double v) { Jp[(size_t)u * jmul] = v; }
    KTN_HDM size_t stride() const { return L; }
    KTN_HDM size_t jstride() const { return jmul; }
    KTN_HDM double* caddr(uint32_t i) const { return const_cast<double*>(C) + (size_t)i * L + lane; }   // never written for BIG shapes
    KTN_HDM double* saddr(uint32_t i) const { return S + (size_t)i * L + lane; }
    KTN_HDM double* jaddr(uint32_t u) const { return Jp + (size_t)u * jmul; }
};

// ---- fused term runs (KF_TERMS / KR_TERMS) ----
// Each case restates, op for op, what the primitive program for the term would do:
//   c*x        LOAD c; MUL x                      x^2       LOAD x; POW2
//   c*x^2      LOAD x; POW2; MUL c                (c*x)^2   LOAD c; MUL x; POW2
//   exp(c*x+d) LOAD c; MUL x; ADDZ; ADD d; EXP
// and the n-ary sum around them:  acc = 0.0 + t0 (first child) ; acc = acc + t_k.
// Memory is addressed through strided lane pointers so the loops compile to pointer bumps.
#define KTN_SUM_STEP(v) { acc = first ? 0.0 + (v) : acc + (v); first = false; }
template <class M>
KTN_HD double terms_fwd(M& m, uint32_t tk, uint32_t n, uint32_t c0, uint32_t u0, uint32_t s0, bool first, bool saveblob, double acc) {
    const size_t st = m.stride();
    const double* xp = m.saddr(u0);
    switch (tk) {
        case KTN_T_X:
            for (uint32_t t = 0; t < n; ++t, xp += st) KTN_SUM_STEP(*xp)
            break;
        case KTN_T_MULC_X: {
            const double* cp = m.caddr(c0);
            for (uint32_t t = 0; t < n; ++t, xp += st, cp += st) KTN_SUM_STEP(*cp * *xp)
            break; }
        case KTN_T_SQ:
            for (uint32_t t = 0; t < n; ++t, xp += st) { const double x = *xp; KTN_SUM_STEP(x * x) }
            break;
        case KTN_T_MULC_SQ: {
            const double* cp = m.caddr(c0);
            for (uint32_t t = 0; t < n; ++t, xp += st, cp += st) { const double x = *xp; KTN_SUM_STEP((x * x) * *cp) }
            break; }
        case KTN_T_SQ_MULC: {
            const double* cp = m.caddr(c0);
            for (uint32_t t = 0; t < n; ++t, xp += st, cp += st) { const double q = *cp * *xp; KTN_SUM_STEP(q * q) }
            break; }
        case KTN_T_EXP_AFF: {
            double* cp = m.caddr(c0);
            double* sp = saveblob ? cp + st : m.saddr(s0);       // exp values: the term's dead `d` slot, or scratch
            const size_t ss = saveblob ? 2 * st : st;
            uint32_t t = 0;
            for (; t + 2 <= n; t += 2, xp += 2 * st, cp += 4 * st, sp += 2 * ss) {   // two independent exp chains in flight
                const double a0 = (0.0 + cp[0] * xp[0]) + cp[st];
                const double a1 = (0.0 + cp[2 * st] * xp[st]) + cp[3 * st];
                const double v0 = ktn_exp(a0), v1 = ktn_exp(a1);
                sp[0] = v0; sp[ss] = v1;
                KTN_SUM_STEP(v0)
                acc = acc + v1;
            }
            if (t < n) {
                const double v = ktn_exp((0.0 + cp[0] * xp[0]) + cp[st]);
                sp[0] = v;
                KTN_SUM_STEP(v)
            }
            break; }
        default: break;
    }
    return acc;
}

template <class M>
KTN_HD void terms_rev(M& m, uint32_t tk, uint32_t n, uint32_t c0, uint32_t u0, uint32_t s0, bool jacc, bool saveblob, double adj) {
    const size_t st = m.stride(), js = m.jstride();
    const double* xp = m.saddr(u0);
    double* jp = m.jaddr(u0);
#define KTN_J_STEP(a) { const double a_ = (a); *jp = jacc ? *jp + a_ : 0.0 + a_; }
    switch (tk) {
        case KTN_T_X:
            for (uint32_t t = 0; t < n; ++t, jp += js) KTN_J_STEP(adj)
            break;
        case KTN_T_MULC_X: {
            const double* cp = m.caddr(c0);
            for (uint32_t t = 0; t < n; ++t, jp += js, cp += st) KTN_J_STEP(revmul(adj, *cp))
            break; }
        case KTN_T_SQ:
            for (uint32_t t = 0; t < n; ++t, jp += js, xp += st) KTN_J_STEP(revmul(adj, 2.0 * *xp))
            break;
        case KTN_T_MULC_SQ: {
            const double* cp = m.caddr(c0);
            for (uint32_t t = 0; t < n; ++t, jp += js, xp += st, cp += st) KTN_J_STEP(revmul(revmul(adj, *cp), 2.0 * *xp))
            break; }
        case KTN_T_SQ_MULC: {
            const double* cp = m.caddr(c0);
            for (uint32_t t = 0; t < n; ++t, jp += js, xp += st, cp += st) { const double c = *cp; KTN_J_STEP(revmul(revmul(adj, 2.0 * (c * *xp)), c)) }
            break; }
        case KTN_T_EXP_AFF: {
            const double* cp = m.caddr(c0);
            const double* sp = saveblob ? cp + st : m.saddr(s0);
            const size_t ss = saveblob ? 2 * st : st;
            for (uint32_t t = 0; t < n; ++t, jp += js, cp += 2 * st, sp += ss) KTN_J_STEP(revmul(revmul(adj, *sp), *cp))
            break; }
        default: break;
    }
#undef KTN_J_STEP
}
#undef KTN_SUM_STEP

// The tape interpreter.  Warp-uniform control flow: every active lane executes the same op.
template <class M>
KTN_HD double run_program(const KtnIns* prog, uint32_t pc, uint32_t end, M& m, unsigned mask) {
    double acc = 0.0, aux = 0.0, r1 = 0.0, r2 = 0.0;
    for (; pc < end; ++pc) {
        const KtnInsWord w = ktn_fetch_ins(prog + pc);
        const uint32_t op = w.x & 0xffu, kind = (w.x >> 8) & 0xffu, idx = w.y;
        double src = 0.0;
        switch ((op == KF_TERMS || op == KR_TERMS) ? (uint32_t)KTN_K_NONE : kind) {
            case KTN_K_C: src = m.c(idx); break;
            case KTN_K_S: if (op != KF_STORE) src = m.lds(idx); break;
            case KTN_K_R1: src = r1; break;
            case KTN_K_R2: src = r2; break;
            default: break;
        }
        switch (op) {
            case KF_LOAD: acc = src; break;
            case KF_ADD: acc = acc + src; break;
            case KF_ADDZ: acc = 0.0 + acc; break;
            case KF_SUB: acc = acc - src; break;
            case KF_RSUB: acc = src - acc; break;
            case KF_MUL: acc = acc * src; break;
            case KF_DIV: aux = 1.0 / src; acc = acc * aux; break;
            case KF_RDIV: aux = 1.0 / acc; acc = src * aux; break;
            case KF_FDIV: acc = acc / src; break;
            case KF_POW2: acc = acc * acc; break;
            case KF_NEG: acc = -acc; break;
            case KF_EXP: acc = ktn_exp(acc); break;
            case KF_LOG: acc = ktn_log(acc); break;
            case KF_SQRT: acc = ktn_sqrt(acc); break;
            case KF_ABS: acc = ktn_fabs(acc); break;
            case KF_SIN: acc = ktn_sin(acc); break;
            case KF_COS: acc = ktn_cos(acc); break;
            case KF_STORE:
                if (kind == KTN_K_S) m.sts(idx, acc); else if (kind == KTN_K_R1) r1 = acc; else r2 = acc;
                break;
            case KF_LDAUX: aux = src; break;
            case KF_STAUX: m.sts(idx, aux); break;
            case KF_DENP: m.sts(idx, (-acc) * aux); break;
            case KF_POWG: acc = pow_value(acc, aux); break;
            case KF_POWPB: m.sts(idx, pow_dbase(acc, aux)); break;
            case KF_POWPE: m.sts(idx, pow_value(acc, aux) * ktn_log(acc)); break;
            case KF_SELZ: acc = (src == 0.0) ? aux : acc; break;
            case KF_SEL1: acc = (src == 1.0) ? acc : aux; break;
            case KF_CMP: { const uint32_t cmp = w.x >> 16;
                const bool r = cmp == 0u ? src <= acc : cmp == 1u ? src < acc : cmp == 2u ? src >= acc : cmp == 3u ? src > acc : src == acc;
                acc = r ? 1.0 : 0.0; break; }
            case KF_SKIPNZ: if (!KTN_ANY_LANE(mask, src == 0.0)) pc += (w.x >> 16); break;
            case KR_ONE: acc = 1.0; break;
            case KR_MUL: acc = revmul(acc, src); break;
            case KR_NEG: acc = -acc; break;
            case KR_MUL2: acc = revmul(acc, 2.0 * src); break;
            case KR_MULRCP: acc = revmul(acc, 1.0 / src); break;
            case KR_MULHRCP: acc = revmul(acc, 0.5 / src); break;
            case KR_MULSGN: acc = revmul(acc, src >= 0.0 ? 1.0 : -1.0); break;
            case KR_MULCOS: acc = revmul(acc, ktn_cos(src)); break;
            case KR_MULNSIN: acc = revmul(acc, -ktn_sin(src)); break;
            case KR_MULZERO: acc = revmul(acc, 0.0); break;
            case KR_MULEQ1: acc = revmul(acc, src == 1.0 ? 1.0 : 0.0); break;
            case KR_MULNE1: acc = revmul(acc, src == 1.0 ? 0.0 : 1.0); break;
            case KR_JSET: m.jst(idx, 0.0 + acc); break;
            case KR_JACC: m.jst(idx, m.jld(idx) + acc); break;
            case KF_TERMS: acc = terms_fwd(m, kind & 0xfu, w.x >> 16, idx, w.z, w.w, (kind & KTN_TF_FIRST) != 0, (kind & KTN_TF_SAVEBLOB) != 0, acc); break;
            case KR_TERMS: terms_rev(m, kind & 0xfu, w.x >> 16, idx, w.z, w.w, (kind & KTN_TF_JACC) != 0, (kind & KTN_TF_SAVEBLOB) != 0, acc); break;
            default: break;
        }
    }
    return acc;
}


#endif
