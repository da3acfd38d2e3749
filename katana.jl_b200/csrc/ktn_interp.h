// ktn_interp.h -- the tape interpreter core (ISA in ktn_program.h), shared by the sm_100a kernels
// and by the test-only host emulator (tests/emu/) that checks compiler output without a GPU.
// Restates forward_eval / reverse_eval of the reference's tape evaluator (called at
// src/separators.jl:112-113) as an accumulator machine; arithmetic order matches oracle/ktn_oracle.c.
#ifndef KTN_INTERP_H
#define KTN_INTERP_H
#include "ktn_math.h"
#include "ktn_program.h"

struct KtnInsWord { uint32_t x, y; };
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
    const uint2 w = __ldg(reinterpret_cast<const uint2*>(p)); return KtnInsWord{w.x, w.y};
#else
    KtnInsWord w; memcpy(&w, p, 8); return w;
#endif
}

KTN_HD double revmul(double a, double p) { return (a == 0.0 && !ktn_isfinite(p)) ? a : a * p; }

// reference forward rule for `^` (ReverseDiffSparse forward_eval): exponent 2 and 1 are special-cased at run time
KTN_HD_NOINLINE double pow_value(double base, double ex) { return ex == 2.0 ? base * base : ex == 1.0 ? base : ktn_pow(base, ex); }
KTN_HD_NOINLINE double pow_dbase(double base, double ex) { return ex == 2.0 ? 2.0 * base : ex == 1.0 ? 1.0 : ex * ktn_pow(base, ex - 1.0); }

// regular chunks: SoA sections with lane stride 32 in shared memory
struct SmemMem {
    const double* C; double* S; uint32_t lane;
    KTN_HDM double c(uint32_t i) const { return C[i * 32 + lane]; }
    KTN_HDM double lds(uint32_t i) const { return S[i * 32 + lane]; }
    KTN_HDM void sts(uint32_t i, double v) { S[i * 32 + lane] = v; }
};
// BIG chunks: constants straight from the blob in global memory, scratch in a global arena
struct GlobalMem {
    const double* C; double* S; uint32_t lane, L;
    KTN_HDM double c(uint32_t i) const { return C[(size_t)i * L + lane]; }
    KTN_HDM double lds(uint32_t i) const { return S[(size_t)i * L + lane]; }
    KTN_HDM void sts(uint32_t i, double v) { S[(size_t)i * L + lane] = v; }
};

// The tape interpreter.  Warp-uniform control flow: every active lane executes the same op.
template <class M>
KTN_HD double run_program(const KtnIns* prog, uint32_t pc, uint32_t end, M& m, uint32_t nu, unsigned mask) {
    double acc = 0.0, aux = 0.0, r1 = 0.0, r2 = 0.0;
    for (; pc < end; ++pc) {
        const KtnInsWord w = ktn_fetch_ins(prog + pc);
        const uint32_t op = w.x & 0xffu, kind = (w.x >> 8) & 0xffu, idx = w.y;
        double src = 0.0;
        switch (kind) {
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
            case KF_SKIPNZ: if (!KTN_ANY_LANE(mask, src == 0.0)) pc += (w.x >> 16); break;
            case KR_ONE: acc = 1.0; break;
            case KR_MUL: acc = revmul(acc, src); break;
            case KR_NEG: acc = -acc; break;
            case KR_MUL2: acc = revmul(acc, 2.0 * src); break;
            case KR_MULRCP: acc = revmul(acc, 1.0 / src); break;
            case KR_MULHRCP: acc = revmul(acc, 0.5 / src); break;
            case KR_MULSGN: acc = revmul(acc, src >= 0.0 ? 1.0 : -1.0); break;
            case KR_JSET: m.sts(nu + idx, 0.0 + acc); break;
            case KR_JACC: m.sts(nu + idx, m.lds(nu + idx) + acc); break;
            default: break;
        }
    }
    return acc;
}


#endif
