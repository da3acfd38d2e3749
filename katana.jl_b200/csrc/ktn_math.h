// ktn_math.h -- deterministic fp64 elementary functions shared by host and device.
//
// Why this exists: the ECP separation test `g <= ub + f_tol` (reference
// src/separators.jl:120) is a hard threshold, and the selected cut set must be
// bit-identical between the CUDA path and the CPU oracle.  libm's and CUDA's
// exp/log/pow differ in the last bit, so both sides evaluate transcendentals
// with THIS header: only IEEE-754 correctly rounded +,-,*,/,sqrt and explicit
// fma() are used, which round identically on x86-64 and sm_100a.  Build rules:
// host `-ffp-contract=off`, device `--fmad=false` (no implicit contraction).
//
// Accuracy (checked in tests/test_math.py against mpmath): exp, log < 1 ulp;
// sin, cos < 1 ulp for |x| <= 1.6e6; pow <= 2 ulp for |y*log(x)| <= 64.  That is far inside the 1e-12 relative
// agreement the north star asks for against Julia's libm.
#ifndef KTN_MATH_H
#define KTN_MATH_H

#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define KTN_HD __host__ __device__ __forceinline__
#else
#define KTN_HD static inline
#endif

KTN_HD double ktn_bits2d(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d; memcpy(&d, &b, 8); return d;
#endif
}
KTN_HD uint64_t ktn_d2bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b; memcpy(&b, &d, 8); return b;
#endif
}
KTN_HD double ktn_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
KTN_HD double ktn_inf(void) { return ktn_bits2d(0x7FF0000000000000ull); }
KTN_HD double ktn_nan(void) { return ktn_bits2d(0x7FF8000000000000ull); }
KTN_HD int ktn_isfinite(double x) { return ((ktn_d2bits(x) >> 52) & 0x7FF) != 0x7FF; }
KTN_HD double ktn_fabs(double x) { return ktn_bits2d(ktn_d2bits(x) & 0x7FFFFFFFFFFFFFFFull); }
KTN_HD double ktn_sqrt(double x) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(x);
#else
    return __builtin_sqrt(x);
#endif
}
// 2^k for -1022 <= k <= 1023
KTN_HD double ktn_pow2i(int k) { return ktn_bits2d((uint64_t)(k + 1023) << 52); }

#define KTN_LN2_HI 6.93147180369123816490e-01 /* 0x3FE62E42FEE00000: 32 significant bits */
#define KTN_LN2_LO 1.90821492927058770002e-10 /* 0x3DEA39EF35793C76 */
#define KTN_INV_LN2 1.44269504088896338700e+00 /* 0x3FF71547652B82FE */

// exp(x).  k = round(x/ln2); r = x - k ln2 (two fma steps); exp(r) = 1 + r + r^2 q(r),
// q = sum_{j=0..11} r^j/(j+2)!  split into even/odd halves for ILP; result scaled by 2^k.
// Full-range path: NaN, +-inf, overflow, gradual underflow.
// Coefficient tables.  On the device they live in constant memory so every DFMA reads its coefficient as a
// constant-bank operand instead of materialising a 64-bit immediate with two moves.
#define KTN_EXP_COEFFS { \
    2.08767569878680989792e-09 /* 1/12! */, 2.75573192239858906526e-07 /* 1/10! */, 2.48015873015873015873e-05 /* 1/8! */, \
    1.38888888888888888889e-03 /* 1/6!  */, 4.16666666666666666667e-02 /* 1/4!  */, 5.00000000000000000000e-01 /* 1/2! */, \
    1.60590438368216145994e-10 /* 1/13! */, 2.50521083854417187751e-08 /* 1/11! */, 2.75573192239858906526e-06 /* 1/9! */, \
    1.98412698412698412698e-04 /* 1/7!  */, 8.33333333333333333333e-03 /* 1/5!  */, 1.66666666666666666667e-01 /* 1/3! */ }
#define KTN_LOG_COEFFS { \
    8.69565217391304347826e-02 /* 2/23 */, 1.05263157894736842105e-01 /* 2/19 */, 1.33333333333333333333e-01 /* 2/15 */, \
    1.81818181818181818182e-01 /* 2/11 */, 2.85714285714285714286e-01 /* 2/7  */, 6.66666666666666666667e-01 /* 2/3  */, \
    9.52380952380952380952e-02 /* 2/21 */, 1.17647058823529411765e-01 /* 2/17 */, 1.53846153846153846154e-01 /* 2/13 */, \
    2.22222222222222222222e-01 /* 2/9  */, 4.00000000000000000000e-01 /* 2/5  */, 0.0 }
static const double KTN_EXPC_HOST[12] = KTN_EXP_COEFFS;
static const double KTN_LOGC_HOST[12] = KTN_LOG_COEFFS;
#if defined(__CUDACC__)
static __constant__ double KTN_EXPC_DEV[12] = KTN_EXP_COEFFS;
static __constant__ double KTN_LOGC_DEV[12] = KTN_LOG_COEFFS;
#endif
#if defined(__CUDA_ARCH__)
#define KTN_EXPC(i) KTN_EXPC_DEV[i]
#define KTN_LOGC(i) KTN_LOGC_DEV[i]
#else
#define KTN_EXPC(i) KTN_EXPC_HOST[i]
#define KTN_LOGC(i) KTN_LOGC_HOST[i]
#endif

KTN_HD double ktn_exp_poly(double r) {
    double z = r * r;
    // even part A(z): 1/12!, 1/10!, 1/8!, 1/6!, 1/4!, 1/2!
    double a = KTN_EXPC(0);
    a = ktn_fma(a, z, KTN_EXPC(1));
    a = ktn_fma(a, z, KTN_EXPC(2));
    a = ktn_fma(a, z, KTN_EXPC(3));
    a = ktn_fma(a, z, KTN_EXPC(4));
    a = ktn_fma(a, z, KTN_EXPC(5));
    // odd part B(z): 1/13!, 1/11!, 1/9!, 1/7!, 1/5!, 1/3!
    double b = KTN_EXPC(6);
    b = ktn_fma(b, z, KTN_EXPC(7));
    b = ktn_fma(b, z, KTN_EXPC(8));
    b = ktn_fma(b, z, KTN_EXPC(9));
    b = ktn_fma(b, z, KTN_EXPC(10));
    b = ktn_fma(b, z, KTN_EXPC(11));
    double q = ktn_fma(b, r, a);
    double s = ktn_fma(z, q, r);
    return 1.0 + s;
}
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
static
#endif
double ktn_exp_slow(double x) {
    if (!(x == x)) return x + x;
    if (x > 709.782712893384) return ktn_inf();
    if (x < -745.1332191019412) return 0.0;
    const double SHIFT = 6755399441055744.0; /* 1.5 * 2^52 */
    double t = x * KTN_INV_LN2;
    double kd = (t + SHIFT) - SHIFT;
    double r = ktn_fma(-kd, KTN_LN2_HI, x);
    r = ktn_fma(-kd, KTN_LN2_LO, r);
    double y = ktn_exp_poly(r);
    int k = (int)kd;
    if (k > 1023) return (y * ktn_pow2i(k - 1)) * 2.0;
    if (k < -1021) return (y * ktn_pow2i(k + 54)) * 5.5511151231257827e-17; /* 2^-54: one rounding into the subnormals */
    return y * ktn_pow2i(k);
}
// Fast path for |x| <= 708 (result is a normal number): identical arithmetic, k is read from the low
// word of (t + SHIFT) and added to the exponent field, which equals the multiplication by 2^k bit for bit.
// ktn_exp_fast is branch-free and may be evaluated on ANY input (the result is only meaningful when
// ktn_exp_is_fast(x)); callers that evaluate several exponentials at once run it unconditionally and
// patch the rare out-of-range lanes with ktn_exp_slow afterwards: the same value ktn_exp returns.
KTN_HD int ktn_exp_is_fast(double x) { return ktn_fabs(x) <= 708.0; }
KTN_HD double ktn_exp_fast(double x) {
    const double SHIFT = 6755399441055744.0;
    double ts = x * KTN_INV_LN2 + SHIFT;
    double kd = ts - SHIFT;
    double r = ktn_fma(-kd, KTN_LN2_HI, x);
    r = ktn_fma(-kd, KTN_LN2_LO, r);
    double y = ktn_exp_poly(r);
    int64_t k = (int64_t)(int32_t)(uint32_t)ktn_d2bits(ts);
    return ktn_bits2d(ktn_d2bits(y) + ((uint64_t)k << 52));
}
KTN_HD double ktn_exp(double x) {
    if (!ktn_exp_is_fast(x)) return ktn_exp_slow(x);
    return ktn_exp_fast(x);
}

// Taylor tail of 2*atanh(s) = 2s + s*z*P(z), z = s^2, P(z) = sum_{n>=1} 2/(2n+1) z^(n-1), n = 1..11
KTN_HD double ktn_log_tail_poly(double z) {
    double w = z * z;
    // odd n = 11,9,..,1: 2/23, 2/19, 2/15, 2/11, 2/7, 2/3
    double a = KTN_LOGC(0);
    a = ktn_fma(a, w, KTN_LOGC(1));
    a = ktn_fma(a, w, KTN_LOGC(2));
    a = ktn_fma(a, w, KTN_LOGC(3));
    a = ktn_fma(a, w, KTN_LOGC(4));
    a = ktn_fma(a, w, KTN_LOGC(5));
    // even n = 10,8,..,2: 2/21, 2/17, 2/13, 2/9, 2/5
    double b = KTN_LOGC(6);
    b = ktn_fma(b, w, KTN_LOGC(7));
    b = ktn_fma(b, w, KTN_LOGC(8));
    b = ktn_fma(b, w, KTN_LOGC(9));
    b = ktn_fma(b, w, KTN_LOGC(10));
    return ktn_fma(b, z, a);
}

// Splits a positive finite x into 2^k * m with m in [sqrt(1/2), sqrt(2)).
KTN_HD double ktn_frexp_sqrt2(double x, int* kout) {
    int k = 0;
    if (x < 2.2250738585072014e-308) { x *= 18014398509481984.0; k = -54; } /* 2^54 */
    uint64_t bits = ktn_d2bits(x);
    k += (int)((bits >> 52) & 0x7FF) - 1023;
    uint64_t mant = bits & 0x000FFFFFFFFFFFFFull;
    uint64_t ebits = 0x3FF0000000000000ull;
    if (mant >= 0x6A09E667F3BCDull) { ebits = 0x3FE0000000000000ull; k += 1; } /* m >= sqrt(2): halve */
    *kout = k;
    return ktn_bits2d(mant | ebits);
}

// log(x), natural.  Published fdlibm-style assembly f - (hfsq - s*(hfsq+R)) with
// a plain Taylor R (coefficients 2/(2n+1)); log(<0) = NaN as NaNMath.log does.
KTN_HD double ktn_log(double x) {
    if (!(x == x)) return x + x;
    if (x < 0.0) return ktn_nan();
    if (x == 0.0) return -ktn_inf();
    if (!ktn_isfinite(x)) return x;
    int k;
    double m = ktn_frexp_sqrt2(x, &k);
    double f = m - 1.0;
    double s = f / (2.0 + f);
    double z = s * s;
    double R = z * ktn_log_tail_poly(z);
    double hfsq = 0.5 * f * f;
    double dk = (double)k;
    return dk * KTN_LN2_HI - ((hfsq - (s * (hfsq + R) + dk * KTN_LN2_LO)) - f);
}

// pow(x, y) with C99 / NaNMath.pow special cases.  General case:
// exp(y * log|x|) with log|x| carried as a double-double (~2^-58 relative).
KTN_HD double ktn_pow(double x, double y) {
    if (y == 0.0) return 1.0;
    if (x == 1.0) return 1.0;
    if (!(x == x) || !(y == y)) return x + y;
    double ax = ktn_fabs(x), ay = ktn_fabs(y);
    int yint = 0; /* 0 non-integer, 1 odd integer, 2 even integer */
    if (ay >= 9007199254740992.0) yint = 2;
    else if (ay >= 1.0) { int64_t iy = (int64_t)ay; if ((double)iy == ay) yint = 2 - (int)(iy & 1); }
    int xneg = (int)(ktn_d2bits(x) >> 63);
    if (!ktn_isfinite(y)) {
        if (ax == 1.0) return 1.0;
        return ((ax > 1.0) == (y > 0.0)) ? ktn_inf() : 0.0;
    }
    if (ax == 0.0 || !ktn_isfinite(ax)) {
        double r = ((ax == 0.0) == (y > 0.0)) ? 0.0 : ktn_inf();
        return (xneg && yint == 1) ? -r : r;
    }
    if (xneg && yint == 0) return ktn_nan();
    double sign = (xneg && yint == 1) ? -1.0 : 1.0;
    int k;
    double m = ktn_frexp_sqrt2(ax, &k);
    double f = m - 1.0;
    // d = m + 1 as an exact double-double (TwoSum)
    double dh = m + 1.0;
    double bb = dh - m;
    double dl = (m - (dh - bb)) + (1.0 - bb);
    // s = f / d as hi + lo
    double sh = f / dh;
    double rem = ktn_fma(-sh, dh, f);
    rem = rem - sh * dl;
    double sl = rem / dh;
    double z = sh * sh;
    double T = sh * (z * ktn_log_tail_poly(z));
    double lh = 2.0 * sh;
    double ll = 2.0 * sl + T;
    // (h, l) = Fast2Sum(lh, ll)
    double h = lh + ll;
    double l = ll - (h - lh);
    double dk = (double)k;
    double kh = dk * KTN_LN2_HI, kl = dk * KTN_LN2_LO;
    // (H, e1) = TwoSum(kh, h)
    double H = kh + h;
    double b2 = H - kh;
    double e1 = (kh - (H - b2)) + (h - b2);
    double L = e1 + (kl + l);
    double ph = y * H;
    double pe = ktn_fma(y, H, -ph);
    double pl = pe + y * L;
    if (ph > 709.79) return sign * ktn_inf();
    if (ph < -745.2) return sign * 0.0;
    double e = ktn_exp(ph);
    return sign * ktn_fma(e, pl, e);
}

// sin(x), cos(x).  n = round(x * 2/pi); r = x - n * pi/2 with pi/2 = P1 + P2 + P3 (33 + 33 + 53 bits: the two fma steps are exact for
// |n| < 2^20, the differences are compensated: r = rh + rl); on |r| <= pi/4 Taylor polynomials through r^17 (sin)
// and r^16 (cos), with the first-order correction for rl; the quadrant n mod 4 picks function and sign.  Measured against mpmath
// (tests/test_math.py): < 1 ulp for |x| <= 1.6e6 (2^20 * pi/2).  Beyond that the reduction loses accuracy gradually (absolute
// error about |x| * 2^-86); |x| >= 2^45, +-inf and NaN give NaN -- a row evaluated there is reported as not finite.
#define KTN_2_OVER_PI 6.36619772367581382433e-01 /* 0x3FE45F306DC9C883 */
#define KTN_PIO2_1 1.57079632673412561417e+00    /* 0x3FF921FB54400000: first 33 bits of pi/2 */
#define KTN_PIO2_2 6.07710050630396597660e-11    /* 0x3DD0B4611A600000: next 33 bits */
#define KTN_PIO2_3 2.02226624879595063154e-21    /* 0x3BA3198A2E037073: pi/2 - P1 - P2 */
KTN_HD double ktn_sin_kernel(double rh, double rl) {
    const double z = rh * rh;
    double p = 2.81145725434552059811e-15;                 /*  1/17! */
    p = ktn_fma(p, z, -7.64716373181981640551e-13);        /* -1/15! */
    p = ktn_fma(p, z, 1.60590438368216133409e-10);         /*  1/13! */
    p = ktn_fma(p, z, -2.50521083854417202239e-08);        /* -1/11! */
    p = ktn_fma(p, z, 2.75573192239858925110e-06);         /*  1/9!  */
    p = ktn_fma(p, z, -1.98412698412698412526e-04);        /* -1/7!  */
    p = ktn_fma(p, z, 8.33333333333333321769e-03);         /*  1/5!  */
    p = ktn_fma(p, z, -1.66666666666666657415e-01);        /* -1/3!  */
    const double corr = ktn_fma(-0.5 * z, rl, rl);         /* rl * cos(rh), first order */
    return rh + ktn_fma(z * rh, p, corr);
}
KTN_HD double ktn_cos_kernel(double rh, double rl) {
    const double z = rh * rh;
    double p = 4.77947733238738525345e-14;                 /*  1/16! */
    p = ktn_fma(p, z, -1.14707455977297245073e-11);        /* -1/14! */
    p = ktn_fma(p, z, 2.08767569878681001866e-09);         /*  1/12! */
    p = ktn_fma(p, z, -2.75573192239858882758e-07);        /* -1/10! */
    p = ktn_fma(p, z, 2.48015873015873015658e-05);         /*  1/8!  */
    p = ktn_fma(p, z, -1.38888888888888894189e-03);        /* -1/6!  */
    p = ktn_fma(p, z, 4.16666666666666643537e-02);         /*  1/4!  */
    const double hz = 0.5 * z, w = 1.0 - hz;
    const double tail = ktn_fma(z * z, p, -(rh * rl));     /* z^2 P(z) - rl * sin(rh), first order */
    return w + (((1.0 - w) - hz) + tail);                  /* 1 - z/2 without losing the bits of z/2 */
}
// quadrant (0..3) and remainder; returns -1 when x is outside the supported range
KTN_HD int ktn_rem_pio2(double x, double* rh, double* rl) {
    if (!(ktn_fabs(x) < 35184372088832.0)) return -1;      /* 2^45; also NaN and +-inf */
    const double SHIFT = 6755399441055744.0;               /* 1.5 * 2^52 */
    const double ts = x * KTN_2_OVER_PI + SHIFT;
    const double nd = ts - SHIFT;
    const double r1 = ktn_fma(-nd, KTN_PIO2_1, x);         /* exact for |n| < 2^20 */
    const double w = nd * KTN_PIO2_2;                      /* exact for |n| < 2^20 */
    const double h = r1 - w, b1 = h - r1;                  /* two-sum: r1 - w = h + l */
    const double l = (r1 - (h - b1)) - (w + b1);
    const double t = l - nd * KTN_PIO2_3;
    const double s = h + t, b2 = s - h;                    /* two-sum: h + t = s + e */
    *rh = s; *rl = (h - (s - b2)) + (t - b2);
    return (int)(ktn_d2bits(ts) & 3ull);
}
KTN_HD double ktn_sin(double x) {
    if (x == 0.0) return x;                                /* keeps -0.0 */
    double rh, rl;
    const int q = ktn_rem_pio2(x, &rh, &rl);
    if (q < 0) return ktn_nan();
    const double v = (q & 1) ? ktn_cos_kernel(rh, rl) : ktn_sin_kernel(rh, rl);
    return (q & 2) ? -v : v;
}
KTN_HD double ktn_cos(double x) {
    double rh, rl;
    const int q = ktn_rem_pio2(x, &rh, &rl);
    if (q < 0) return ktn_nan();
    const double v = (q & 1) ? ktn_sin_kernel(rh, rl) : ktn_cos_kernel(rh, rl);
    return ((q + 1) & 2) ? -v : v;
}

// Julia's max(): NaN-propagating (used by round_coefs, reference src/model.jl:201).
KTN_HD double ktn_jlmax(double a, double b) {
    if (!(a == a)) return a;
    if (!(b == b)) return b;
    return a > b ? a : b;
}

#endif /* KTN_MATH_H */
