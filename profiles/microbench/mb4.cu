// mb4.cu -- round 2 microbenchmark: a fully asynchronous family-row pipeline against the register-resident row of round 1.
//
// Workload = the headline one: 10^6 log-sum-exp rows in chunks of 32 rows (one warp), K terms per row, per term c, d (fp64) and
// a column id (int32), x* of 10^5 doubles gathered at random.  Two data sets: uniform K = 10 and the mixed classes K = 4..16.
//
// P  pipeline per warp (no block-wide synchronisation at all):
//      A(j)  TMA bulk copy of chunk j's column ids                 -> shared-memory ring of NCOL slots      (DRAM latency)
//      B(j)  cp.async (LDGSTS) 8-byte gathers x*[col] of chunk j    -> x slot j&1; TMA of chunk j's constants (L2 / DRAM latency)
//      C(j)  compute chunk j from shared memory only: exp, sum, log, optional sparse cut
//    iteration i runs B(i+1), A(i+NCOL), C(i): no instruction of C waits on global memory.
// R  round 1's scheme: every constant and column of the row loaded to registers at once, then the gathers, then exp.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -lineinfo -o mb4 mb4.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../../katana.jl_b200/csrc/ktn_math.h"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct Params {
    const unsigned char* blob; const double* x; double* g; double* jout; unsigned* ticket; unsigned nchunks;
    unsigned cls_begin[18]; unsigned long long cls_off[17];
};
#define NSEG 16

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mb_bulk(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cpa8(void* dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s32(dst)), "l"(src) : "memory"); }
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void chunk_lookup(const Params& p, unsigned c, unsigned& N, const unsigned char*& b) {
    unsigned k = 1;
    while (k < 16 && c >= p.cls_begin[k + 1]) ++k;
    N = k; b = p.blob + p.cls_off[k] + (size_t)(c - p.cls_begin[k]) * (640u * k + 256u);
}

// chunk order: static = round robin over all warps; dynamic = NSEG segment counters (a warp starts in the segment of its SM
// and moves on cyclically when a segment runs dry)
struct Sched {
    unsigned next_static, stride, seg, tried;
    __device__ void init(const Params& p, bool dyn) {
        const unsigned gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        next_static = gw; stride = gridDim.x * (blockDim.x >> 5); seg = blockIdx.x % NSEG; tried = 0;
    }
    __device__ unsigned next(const Params& p, bool dyn, unsigned lane) {
        if (!dyn) { const unsigned c = next_static; next_static += stride; return c < p.nchunks ? c : 0xffffffffu; }
        unsigned c = 0xffffffffu;
        if (lane == 0) {
            const unsigned per = (p.nchunks + NSEG - 1) / NSEG;
            while (tried < NSEG) {
                const unsigned t = atomicAdd(&p.ticket[seg * 32], 1u);
                const unsigned lo = seg * per, hi = lo + per < p.nchunks ? lo + per : p.nchunks;
                if (lo + t < hi) { c = lo + t; break; }
                seg = (seg + 1) % NSEG; ++tried;
            }
        }
        return __shfl_sync(0xffffffffu, c, 0);
    }
};

template <int NMAX, int NCOL, bool EXP, bool CUT, bool GATH, bool DYN, int G>
__global__ void __launch_bounds__(512, 1) k_pipe(const Params p) {
    extern __shared__ __align__(128) unsigned char sm[];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr unsigned CB = 512 * NMAX + 256, RB = 128 * NMAX, XB = 256 * NMAX, WB = 2 * CB + NCOL * RB + 2 * XB + 64;
    unsigned char* wb = sm + (size_t)warp * WB;
    unsigned char* cbuf = wb; unsigned char* rbuf = wb + 2 * CB; unsigned char* xbuf = rbuf + NCOL * RB;
    uint64_t* cbar = reinterpret_cast<uint64_t*>(xbuf + 2 * XB); uint64_t* rbar = cbar + 2;
    if (lane == 0) { for (int s = 0; s < 2; ++s) mb_init(&cbar[s], 1); for (int s = 0; s < NCOL; ++s) mb_init(&rbar[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    Sched sc; sc.init(p, DYN);
    unsigned id[NCOL + 1];
    auto issueA = [&](unsigned j, unsigned c) {      // column ids of chunk c (the j-th chunk of this warp)
        if (lane == 0) { unsigned N; const unsigned char* b; chunk_lookup(p, c, N, b); const unsigned s = j % NCOL; mb_expect(&rbar[s], 128u * N); mb_bulk(rbuf + s * RB, b, 128u * N, &rbar[s]); }
    };
    auto stageB = [&](unsigned j, unsigned c) {      // gathers + constants of chunk c
        unsigned N; const unsigned char* b; chunk_lookup(p, c, N, b);
        const unsigned s = j % NCOL;
        mb_wait(&rbar[s], (j / NCOL) & 1u);
        if (GATH) {
            const int32_t* col = reinterpret_cast<const int32_t*>(rbuf + s * RB) + lane;
            double* xs = reinterpret_cast<double*>(xbuf + (j & 1u) * XB) + lane;
#pragma unroll 4
            for (unsigned u = 0; u < N; ++u) cpa8(xs + u * 32, p.x + col[u * 32]);
        }
        cpa_commit();
        if (lane == 0) { mb_expect(&cbar[j & 1u], 512u * N + 256u); mb_bulk(cbuf + (j & 1u) * CB, b + 128u * N, 512u * N + 256u, &cbar[j & 1u]); }
    };
#pragma unroll
    for (int j = 0; j < NCOL; ++j) { id[j] = sc.next(p, DYN, lane); if (id[j] != 0xffffffffu) issueA(j, id[j]); }
    if (id[0] != 0xffffffffu) stageB(0, id[0]);
    for (unsigned i = 0; id[0] != 0xffffffffu; ++i) {
        if (id[1] != 0xffffffffu) stageB(i + 1, id[1]); else cpa_commit();
        id[NCOL] = sc.next(p, DYN, lane);
        if (id[NCOL] != 0xffffffffu) issueA(i + NCOL, id[NCOL]);
        // ---- C(i): shared memory only ----
        unsigned N; const unsigned char* bdummy; chunk_lookup(p, id[0], N, bdummy);
        cpa_wait<1>();
        mb_wait(&cbar[i & 1u], (i >> 1) & 1u);
        double* C = reinterpret_cast<double*>(cbuf + (i & 1u) * CB) + lane;
        const double* X = reinterpret_cast<const double*>(xbuf + (i & 1u) * XB) + lane;
        double acc = 0.0;
        for (unsigned u0 = 0; u0 < N; u0 += G) {
            double a[G], e[G];
#pragma unroll
            for (int k = 0; k < G; ++k) { const unsigned u = u0 + k < N ? u0 + k : N - 1; const double xx = GATH ? X[u * 32] : 1.0; a[k] = (0.0 + C[(2 * u) * 32] * xx) + C[(2 * u + 1) * 32]; }
            if (EXP) {
                bool slow = false;
#pragma unroll
                for (int k = 0; k < G; ++k) { e[k] = ktn_exp_fast(a[k]); slow = slow || !ktn_exp_is_fast(a[k]); }
                if (slow) {
#pragma unroll
                    for (int k = 0; k < G; ++k) if (!ktn_exp_is_fast(a[k])) e[k] = ktn_exp_slow(a[k]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < G; ++k) e[k] = a[k];
            }
#pragma unroll
            for (int k = 0; k < G; ++k) if (u0 + k < N) { if (CUT) C[(2 * (u0 + k) + 1) * 32] = e[k]; acc = acc + e[k]; }
        }
        const double g = EXP ? ktn_log(acc) : acc;
        const size_t slot = (size_t)id[0] * 32 + lane;
        p.g[slot] = g;
        if (CUT) {
            const bool sel = ((slot * 2654435761u) >> 7) % 10u == 0u;       // ~10 % of the rows, sparse lanes
            if (sel) {
                const uint64_t ow = reinterpret_cast<const uint64_t*>(cbuf + (i & 1u) * CB + 512u * N)[lane];
                const double adj = 1.0 / acc;
                double b = g;
                double* out = p.jout + slot * 16;
#pragma unroll 4
                for (unsigned q = 0; q < N; ++q) {
                    const unsigned u = (unsigned)(ow >> (4 * q)) & 15u;
                    const double jv = 0.0 + (adj * C[(2 * u + 1) * 32]) * C[(2 * u) * 32];
                    b = b + (-(GATH ? X[u * 32] : 1.0)) * jv;
                    out[q] = jv;
                }
                p.g[slot] = b;
            }
        }
        fence_async();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < NCOL; ++j) id[j] = id[j + 1];
    }
}

// R: round 1's scheme for K compile-time (uniform data set): all loads of the row at once, gathers, exp, log
template <int K, int WPB, int BPS, bool EXP>
__global__ void __launch_bounds__(WPB * 32, BPS) k_reg(const Params p) {
    const unsigned lane = threadIdx.x & 31;
    Sched sc; sc.init(p, true);
    for (;;) {
        const unsigned c = sc.next(p, true, lane);
        if (c == 0xffffffffu) break;
        const unsigned char* b = p.blob + (size_t)c * (640u * K + 256u);
        const int32_t* col = reinterpret_cast<const int32_t*>(b) + lane;
        const double* C = reinterpret_cast<const double*>(b + 128u * K) + lane;
        double cc[K], dd[K]; int cl[K];
#pragma unroll
        for (int u = 0; u < K; ++u) { cc[u] = __ldg(C + (2 * u) * 32); dd[u] = __ldg(C + (2 * u + 1) * 32); cl[u] = __ldg(col + u * 32); }
        double a[K];
#pragma unroll
        for (int u = 0; u < K; ++u) a[u] = (0.0 + cc[u] * __ldg(p.x + cl[u])) + dd[u];
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < K; ++u) acc = acc + (EXP ? ktn_exp_fast(a[u]) : a[u]);
        p.g[(size_t)c * 32 + lane] = EXP ? ktn_log(acc) : acc;
    }
}

static inline uint64_t splitmix(uint64_t& s) { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static inline double u01(uint64_t& s) { return (double)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0); }

struct Data { Params p; size_t bytes; std::vector<double> ref; };

static Data make(bool mixed, int n) {
    Data D; memset(&D.p, 0, sizeof D.p);
    const unsigned nchunks = 31250;
    std::vector<unsigned> cnt(18, 0);
    if (mixed) for (unsigned c = 0; c < nchunks; ++c) cnt[4 + c % 13]++; else cnt[10] = nchunks;
    unsigned at = 0; unsigned long long off = 0;
    for (int k = 0; k <= 16; ++k) { D.p.cls_begin[k] = at; D.p.cls_off[k] = off; at += cnt[k]; off += (unsigned long long)cnt[k] * (640ull * k + 256ull); }
    D.p.cls_begin[17] = at; D.bytes = off; D.p.nchunks = nchunks;
    std::vector<unsigned char> h(off + 256);
    std::vector<double> x(n);
    uint64_t s = 7;
    for (int i = 0; i < n; ++i) x[i] = 4.0 * u01(s) - 2.0;
    D.ref.assign((size_t)nchunks * 32, 0.0);
    for (int k = 1; k <= 16; ++k) for (unsigned c = D.p.cls_begin[k]; c < D.p.cls_begin[k + 1]; ++c) {
        unsigned char* b = h.data() + D.p.cls_off[k] + (size_t)(c - D.p.cls_begin[k]) * (640u * k + 256u);
        int32_t* col = reinterpret_cast<int32_t*>(b); double* C = reinterpret_cast<double*>(b + 128u * k); uint64_t* ow = reinterpret_cast<uint64_t*>(b + 640u * k);
        for (int lane = 0; lane < 32; ++lane) {
            double acc = 0.0; uint64_t w = 0;
            for (int u = 0; u < k; ++u) {
                const int cl = (int)(splitmix(s) % (uint64_t)n); const double cc = 2.0 * u01(s) - 1.0, dd = 2.0 * u01(s) - 1.0;
                col[u * 32 + lane] = cl; C[(2 * u) * 32 + lane] = cc; C[(2 * u + 1) * 32 + lane] = dd;
                acc = acc + ktn_exp((0.0 + cc * x[cl]) + dd); w |= (uint64_t)((u * 7 + 3) % k) << (4 * u);
            }
            ow[lane] = w; D.ref[(size_t)c * 32 + lane] = ktn_log(acc);
        }
    }
    unsigned char* d; double *dx, *g, *jout; unsigned* ticket;
    CK(cudaMalloc(&d, h.size())); CK(cudaMalloc(&dx, 8 * (size_t)n)); CK(cudaMalloc(&g, 8 * (size_t)nchunks * 32)); CK(cudaMalloc(&jout, 8 * (size_t)nchunks * 32 * 16)); CK(cudaMalloc(&ticket, 4 * 32 * NSEG));
    CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dx, x.data(), 8 * (size_t)n, cudaMemcpyHostToDevice));
    D.p.blob = d; D.p.x = dx; D.p.g = g; D.p.jout = jout; D.p.ticket = ticket;
    return D;
}

template <class F> static float timeit(F f, const Data& D) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 6; ++it) {
        CK(cudaMemset(D.p.ticket, 0, 4 * 32 * NSEG));
        cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (it >= 2 && ms < best) best = ms;
    }
    return best;
}
static void check(const Data& D, const char* what) {
    std::vector<double> g(D.ref.size());
    CK(cudaMemcpy(g.data(), D.p.g, 8 * g.size(), cudaMemcpyDeviceToHost));
    size_t bad = 0; for (size_t i = 0; i < g.size(); ++i) if (memcmp(&g[i], &D.ref[i], 8)) ++bad;
    printf("   check %s: %zu of %zu rows differ from the host evaluation\n", what, bad, g.size());
}

static int sms;
template <int NMAX, int NCOL, bool EXP, bool CUT, bool GATH, bool DYN, int G = 4> static void run_pipe(const Data& D, int W, const char* tag, bool chk = false) {
    constexpr unsigned WB = 2 * (512 * NMAX + 256) + NCOL * 128 * NMAX + 2 * 256 * NMAX + 64;
    const int smem = W * WB;
    if (smem > 232448) { printf("pipe %s NMAX=%d NCOL=%d W=%d: %d B of shared memory do not fit\n", tag, NMAX, NCOL, W, smem); return; }
    CK(cudaFuncSetAttribute(k_pipe<NMAX, NCOL, EXP, CUT, GATH, DYN, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const float ms = timeit([&] { k_pipe<NMAX, NCOL, EXP, CUT, GATH, DYN, G><<<sms, W * 32, smem>>>(D.p); }, D);
    printf("pipe %-7s NMAX=%2d NCOL=%d exp=%d cut=%d gather=%d dyn=%d G=%d W=%2d smem=%6d: %6.1f us  %5.0f GB/s\n", tag, NMAX, NCOL, EXP, CUT, GATH, DYN, G, W, smem, ms * 1e3, D.bytes / ms / 1e6);
    if (chk && EXP && GATH && !CUT) check(D, "pipe");
}
template <int K, int WPB, int BPS, bool EXP> static void run_reg(const Data& D, bool chk = false, int smem = 0) {
    CK(cudaFuncSetAttribute(k_reg<K, WPB, BPS, EXP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem > 0 ? smem : 1024));
    const float ms = timeit([&] { k_reg<K, WPB, BPS, EXP><<<sms * BPS, WPB * 32, smem>>>(D.p); }, D);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_reg<K, WPB, BPS, EXP>);
    printf("reg  uniform K=%d warps/SM=%2d exp=%d regs=%d spill=%zu idle smem/block=%d: %6.1f us  %5.0f GB/s\n", K, WPB * BPS, EXP, fa.numRegs, (size_t)fa.localSizeBytes, smem, ms * 1e3, D.bytes / ms / 1e6);
    if (chk && EXP) check(D, "reg");
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int n = 100000;
    Data U = make(false, n), M = make(true, n);
    printf("uniform K=10: %.1f MB; mixed K=4..16: %.1f MB; %d SMs\n", U.bytes / 1e6, M.bytes / 1e6, sms);
    // round 1's scheme, with and without the exponentials
    run_reg<10, 16, 1, true>(U, true); run_reg<10, 8, 4, true>(U); run_reg<10, 16, 1, false>(U); run_reg<10, 8, 4, false>(U);
    // does an idle shared-memory carve-out (a smaller L1) slow the register scheme down?
    run_reg<10, 16, 1, true>(U, false, 100 * 1024); run_reg<10, 16, 1, true>(U, false, 200 * 1024); run_reg<10, 16, 1, false>(U, false, 200 * 1024);
    // pipeline, uniform data: warps, ring depth, pieces
    run_pipe<10, 3, true, false, true, false>(U, 4, "uniform", true);
    run_pipe<10, 3, true, false, true, false>(U, 6, "uniform");
    run_pipe<10, 3, true, false, true, false>(U, 8, "uniform");
    run_pipe<10, 3, true, false, true, false>(U, 11, "uniform");
    run_pipe<10, 2, true, false, true, false>(U, 8, "uniform");
    run_pipe<10, 2, true, false, true, false>(U, 12, "uniform");
    run_pipe<10, 3, true, false, true, true>(U, 8, "uniform");
    run_pipe<10, 3, true, false, true, true>(U, 11, "uniform");
    run_pipe<10, 3, false, false, true, false>(U, 8, "uniform");      // no exp: the memory pattern alone
    run_pipe<10, 3, false, false, true, false>(U, 11, "uniform");
    run_pipe<10, 3, false, false, false, false>(U, 8, "uniform");     // no exp, no gathers: the streams alone
    run_pipe<10, 3, true, false, false, false>(U, 8, "uniform");      // exp, no gathers
    run_pipe<10, 3, true, true, true, false>(U, 8, "uniform");        // with the sparse cut
    run_pipe<10, 3, true, true, true, false>(U, 11, "uniform");
    run_pipe<10, 3, true, false, true, false, 8>(U, 8, "uniform");    // 8 exponentials in flight per lane
    run_pipe<10, 3, true, false, true, false, 2>(U, 8, "uniform");
    run_pipe<10, 3, true, true, true, false, 8>(U, 8, "uniform");
    run_pipe<16, 3, true, false, true, false>(U, 7, "uniform");       // slots sized for 16 terms
    run_pipe<16, 2, true, false, true, false>(U, 7, "uniform");
    // mixed classes (the headline instance): slots sized for 16 terms
    run_pipe<16, 3, true, false, true, false>(M, 7, "mixed", true);
    run_pipe<16, 3, true, false, true, true>(M, 7, "mixed");
    run_pipe<16, 2, true, false, true, true>(M, 7, "mixed");
    run_pipe<16, 3, true, true, true, true>(M, 7, "mixed");
    run_pipe<16, 3, true, true, true, true>(M, 5, "mixed");
    run_pipe<16, 3, true, true, true, true, 8>(M, 7, "mixed");
    run_pipe<16, 3, false, false, true, true>(M, 7, "mixed");
    printf("done\n");
    return 0;
}
