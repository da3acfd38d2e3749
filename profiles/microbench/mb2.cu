// mb2.cu -- access-pattern microbenchmark for the family kernel (sm_100a).
// One warp per 32-row chunk of K terms: per lane K x (c, d) fp64 + K x int32 column, lane stride 32 (coalesced 256 B / 128 B
// warp loads from a contiguous 672*K byte blob), then K dependent 8-byte gathers from an n-double table, a token of compute.
// Variants: warps per SM, loads-at-once vs per-term, gathers on/off, L1 no_allocate on/off.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb2 mb2.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <bool NA> __device__ __forceinline__ double ldd(const double* p) {
    double v;
    if (NA) asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); else v = __ldg(p);
    return v;
}
template <bool NA> __device__ __forceinline__ int ldi(const int* p) {
    int v;
    if (NA) asm("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p)); else v = __ldg(p);
    return v;
}

// K compile-time; all loads of the row issued at once
template <int K, bool GATHER, bool NA>
__global__ void __launch_bounds__(128) k_once(const unsigned char* blob, const double* x, double* out, unsigned* ticket, unsigned nchunks) {
    const unsigned lane = threadIdx.x & 31;
    double acc = 0;
    for (;;) {
        unsigned c = 0;
        if (lane == 0) c = atomicAdd(ticket, 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= nchunks) break;
        const unsigned char* b = blob + (size_t)c * (672 * K);
        const double* C = reinterpret_cast<const double*>(b) + lane;
        const int* col = reinterpret_cast<const int*>(b + 512 * K) + lane;
        double cc[K], dd[K]; int cl[K];
#pragma unroll
        for (int u = 0; u < K; ++u) { cc[u] = ldd<NA>(C + (2 * u) * 32); dd[u] = ldd<NA>(C + (2 * u + 1) * 32); cl[u] = ldi<NA>(col + u * 32); }
        double xs[K];
#pragma unroll
        for (int u = 0; u < K; ++u) xs[u] = GATHER ? __ldg(x + cl[u]) : (double)cl[u];
#pragma unroll
        for (int u = 0; u < K; ++u) acc += cc[u] * xs[u] + dd[u];
    }
    if (acc == 1.2345) out[0] = acc;
}

// static round-robin chunk assignment (no tickets); BATCH>1: dynamic tickets of BATCH chunks
template <int K, bool GATHER, int BATCH>
__global__ void __launch_bounds__(128) k_static(const unsigned char* blob, const double* x, double* out, unsigned* ticket, unsigned nchunks) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    double acc = 0;
    unsigned c = gw, cend = 0;
    for (;;) {
        if (BATCH == 0) { if (c >= nchunks) break; }
        else {
            if (c >= cend) { unsigned t = 0; if (lane == 0) t = atomicAdd(ticket, (unsigned)BATCH); c = __shfl_sync(0xffffffffu, t, 0); cend = c + BATCH; if (c >= nchunks) break; if (cend > nchunks) cend = nchunks; }
        }
        const unsigned char* b = blob + (size_t)c * (672 * K);
        const double* C = reinterpret_cast<const double*>(b) + lane;
        const int* col = reinterpret_cast<const int*>(b + 512 * K) + lane;
        double cc[K], dd[K]; int cl[K];
#pragma unroll
        for (int u = 0; u < K; ++u) { cc[u] = ldd<true>(C + (2 * u) * 32); dd[u] = ldd<true>(C + (2 * u + 1) * 32); cl[u] = ldi<true>(col + u * 32); }
        double xs[K];
#pragma unroll
        for (int u = 0; u < K; ++u) xs[u] = GATHER ? __ldg(x + cl[u]) : (double)cl[u];
#pragma unroll
        for (int u = 0; u < K; ++u) acc += cc[u] * xs[u] + dd[u];
        c += BATCH == 0 ? nw : 1;
    }
    if (acc == 1.2345) out[0] = acc;
}

// warp-specialised pipeline: warp 0 lane 0 = producer (tickets in batches + cp.async.bulk into a smem ring), other warps consume
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_bulk(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t parity, int who = 0, unsigned info = 0) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(s32(b)), "r"(parity) : "memory");
        if (spin > (1u << 22)) { if ((threadIdx.x & 31) == 0) printf("HANG who=%d info=%u block=%d warp=%d parity=%u\n", who, info, blockIdx.x, threadIdx.x >> 5, parity); __trap(); }
    }
}
template <int K, bool GATHER, int NS, int BATCH>
__global__ void __launch_bounds__(512, 1) k_pipe(const unsigned char* blob, const double* x, double* out, unsigned* ticket, unsigned nchunks, int nwarps) {
    extern __shared__ __align__(128) unsigned char sm[];
    constexpr int SB = 640 * K;
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + NS * SB); uint64_t* empty = full + NS;
    unsigned* meta = reinterpret_cast<unsigned*>(empty + NS); unsigned* head = meta + NS;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int s = 0; s < NS; ++s) { mb_init(&full[s], 1); mb_init(&empty[s], 1); } *head = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (warp == 0) {
        if (lane != 0) return;
        unsigned i = 0, base = 0, end = 0;
        unsigned pend = atomicAdd(ticket, (unsigned)BATCH);
        for (;;) {
            if (base >= end) { if (pend >= nchunks) break; base = pend; end = min(pend + BATCH, nchunks); pend = atomicAdd(ticket, (unsigned)BATCH); }
            const unsigned s = i % NS;
            mb_wait(&empty[s], ((i / NS) & 1) ^ 1, 1, i);
            meta[s] = base;
            mb_expect(&full[s], SB);
            mb_bulk(sm + s * SB, blob + (size_t)base * (672 * K), SB, &full[s]);
            ++i; ++base;
        }
        for (int j = 0; j < nwarps - 1; ++j, ++i) { const unsigned s = i % NS; mb_wait(&empty[s], ((i / NS) & 1) ^ 1); meta[s] = 0xffffffffu; mb_arrive(&full[s]); }
        return;
    }
    double acc = 0;
    for (;;) {
        unsigned idx = 0;
        if (lane == 0) idx = atomicAdd(head, 1u);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        const unsigned s = idx % NS;
        mb_wait(&full[s], (idx / NS) & 1, 2, idx);
        if (meta[s] == 0xffffffffu) break;
        const double* C = reinterpret_cast<const double*>(sm + s * SB) + lane;
        const int* col = reinterpret_cast<const int*>(sm + s * SB + 512 * K) + lane;
        double cc[K], dd[K]; int cl[K];
#pragma unroll
        for (int u = 0; u < K; ++u) { cc[u] = C[(2 * u) * 32]; dd[u] = C[(2 * u + 1) * 32]; cl[u] = col[u * 32]; }
        __syncwarp();
        if (lane == 0) mb_arrive(&empty[s]);
        double xs[K];
#pragma unroll
        for (int u = 0; u < K; ++u) xs[u] = GATHER ? __ldg(x + cl[u]) : (double)cl[u];
#pragma unroll
        for (int u = 0; u < K; ++u) acc += cc[u] * xs[u] + dd[u];
    }
    if (acc == 1.2345) out[0] = acc;
}

// self-prefetch: every warp owns one smem slot + mbarrier; after moving chunk i to registers it issues the bulk copy of its
// chunk i+1 itself, then gathers / computes chunk i.  WORK = dependent DFMA steps per term (stand-in for exp).
template <int K, bool GATHER, int BATCH, int WORK>
__global__ void __launch_bounds__(1024, 1) k_self(const unsigned char* blob, const double* x, double* out, unsigned* ticket, unsigned nchunks) {
    extern __shared__ __align__(128) unsigned char sm[];
    constexpr int SB = 640 * K;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + nw * SB) + warp;
    unsigned char* slot = sm + warp * SB;
    if (lane == 0) { mb_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    unsigned base = 0, end = 0, parity = 0;
    auto next_ticket = [&]() -> unsigned {     // returns chunk or 0xffffffff
        if (base >= end) { unsigned t = 0; if (lane == 0) t = atomicAdd(ticket, (unsigned)BATCH); t = __shfl_sync(0xffffffffu, t, 0); base = t; end = t + BATCH; }
        unsigned c = base++; return c < nchunks ? c : 0xffffffffu;
    };
    auto issue = [&](unsigned c) { if (lane == 0) { mb_expect(bar, SB); mb_bulk(slot, blob + (size_t)c * (672 * K), SB, bar); } };
    unsigned cur = next_ticket();
    if (cur != 0xffffffffu) issue(cur);
    double acc = 0;
    while (cur != 0xffffffffu) {
        const unsigned nxt = next_ticket();
        mb_wait(bar, parity, 3, cur); parity ^= 1;
        const double* C = reinterpret_cast<const double*>(slot) + lane;
        const int* col = reinterpret_cast<const int*>(slot + 512 * K) + lane;
        double cc[K], dd[K]; int cl[K];
#pragma unroll
        for (int u = 0; u < K; ++u) { cc[u] = C[(2 * u) * 32]; dd[u] = C[(2 * u + 1) * 32]; cl[u] = col[u * 32]; }
        __syncwarp();
        if (nxt != 0xffffffffu) issue(nxt);
        double xs[K];
#pragma unroll
        for (int u = 0; u < K; ++u) xs[u] = GATHER ? __ldg(x + cl[u]) : (double)cl[u];
#pragma unroll
        for (int u = 0; u < K; ++u) {
            double a = cc[u] * xs[u] + dd[u];
#pragma unroll
            for (int w = 0; w < WORK; ++w) a = fma(a, 1.0000001, 1e-9);
            acc += a;
        }
        cur = nxt;
    }
    if (acc == 1.2345) out[0] = acc;
}

// per-term loop (low registers, high occupancy), unroll U
template <int U, bool GATHER, bool NA>
__global__ void __launch_bounds__(256) k_loop(const unsigned char* blob, const double* x, double* out, unsigned* ticket, unsigned nchunks, int K) {
    const unsigned lane = threadIdx.x & 31;
    double acc = 0;
    for (;;) {
        unsigned c = 0;
        if (lane == 0) c = atomicAdd(ticket, 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= nchunks) break;
        const unsigned char* b = blob + (size_t)c * (672 * K);
        const double* C = reinterpret_cast<const double*>(b) + lane;
        const int* col = reinterpret_cast<const int*>(b + 512 * K) + lane;
        for (int u0 = 0; u0 < K; u0 += U) {
            double cc[U], dd[U]; int cl[U]; double xs[U];
#pragma unroll
            for (int j = 0; j < U; ++j) { int u = u0 + j < K ? u0 + j : K - 1; cc[j] = ldd<NA>(C + (2 * u) * 32); dd[j] = ldd<NA>(C + (2 * u + 1) * 32); cl[j] = ldi<NA>(col + u * 32); }
#pragma unroll
            for (int j = 0; j < U; ++j) xs[j] = GATHER ? __ldg(x + cl[j]) : (double)cl[j];
#pragma unroll
            for (int j = 0; j < U; ++j) acc += cc[j] * xs[j] + dd[j];
        }
    }
    if (acc == 1.2345) out[0] = acc;
}

static inline uint64_t splitmix(uint64_t& s) { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

template <class F> float timeit(F f, unsigned* ticket) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 5; ++it) {
        CK(cudaMemset(ticket, 0, 4));
        cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (it && ms < best) best = ms;
    }
    return best;
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    const int K = 10, n = 100000; const unsigned nchunks = 31250;
    const size_t bytes = (size_t)nchunks * 672 * K;
    std::vector<unsigned char> h(bytes);
    uint64_t s = 1;
    for (unsigned c = 0; c < nchunks; ++c) {
        double* C = reinterpret_cast<double*>(h.data() + (size_t)c * 672 * K);
        int* col = reinterpret_cast<int*>(h.data() + (size_t)c * 672 * K + 512 * K);
        for (int i = 0; i < 2 * K * 32; ++i) C[i] = 1.0;
        for (int i = 0; i < K * 32; ++i) col[i] = (int)(splitmix(s) % n);
    }
    unsigned char* d; double *x, *out; unsigned* ticket;
    CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&x, 8 * n)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&ticket, 64));
    CK(cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice)); CK(cudaMemset(x, 0, 8 * n));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("chunks %u K %d bytes %.1f MB, %d SMs\n", nchunks, K, bytes / 1e6, sms);
#define RUN_ONCE(G, NA, bps) { float ms = timeit([&] { k_once<K, G, NA><<<sms * bps, 128>>>(d, x, out, ticket, nchunks); }, ticket); \
        printf("once  gather=%d na=%d warps/SM=%2d: %.1f us  %.0f GB/s\n", G, NA, bps * 4, ms * 1e3, bytes / ms / 1e6); }
    RUN_ONCE(true, true, 4) RUN_ONCE(true, true, 6) RUN_ONCE(true, true, 8) RUN_ONCE(true, true, 12) RUN_ONCE(true, true, 16)
    RUN_ONCE(false, true, 4) RUN_ONCE(false, true, 8) RUN_ONCE(true, false, 4) RUN_ONCE(true, false, 8)
#define RUN_ST(G, B, bps) { float ms = timeit([&] { k_static<K, G, B><<<sms * bps, 128>>>(d, x, out, ticket, nchunks); }, ticket); \
        printf("static/batch=%d gather=%d warps/SM=%2d: %.1f us  %.0f GB/s\n", B, G, bps * 4, ms * 1e3, bytes / ms / 1e6); }
    RUN_ST(false, 0, 4) RUN_ST(false, 0, 8) RUN_ST(false, 0, 16) RUN_ST(true, 0, 4) RUN_ST(true, 0, 8) RUN_ST(true, 0, 16)
    RUN_ST(false, 4, 4) RUN_ST(false, 4, 8) RUN_ST(true, 4, 4) RUN_ST(true, 4, 8) RUN_ST(true, 2, 8) RUN_ST(true, 8, 8)
#define RUN_PIPE(G, NS, B, NW) { const int smem = NS * 640 * K + NS * 24 + 64; CK(cudaFuncSetAttribute(k_pipe<K, G, NS, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        float ms = timeit([&] { k_pipe<K, G, NS, B><<<sms, NW * 32, smem>>>(d, x, out, ticket, nchunks, NW); }, ticket); \
        printf("pipe gather=%d slots=%d batch=%d warps=%2d: %.1f us  %.0f GB/s\n", G, NS, B, NW, ms * 1e3, bytes / ms / 1e6); }
    // RUN_PIPE variants hang: bulk copies complete out of order, so a consumer can run a full phase ahead of a slot (parity aliasing)
#define RUN_SELF(G, B, NW, WORK) { const int smem = NW * 640 * K + NW * 8 + 64; CK(cudaFuncSetAttribute(k_self<K, G, B, WORK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        float ms = timeit([&] { k_self<K, G, B, WORK><<<sms, NW * 32, smem>>>(d, x, out, ticket, nchunks); }, ticket); \
        printf("self gather=%d batch=%d warps=%2d work=%d: %.1f us  %.0f GB/s\n", G, B, NW, WORK, ms * 1e3, bytes / ms / 1e6); }
    RUN_SELF(false, 4, 16, 0) RUN_SELF(true, 4, 16, 0) RUN_SELF(true, 4, 24, 0) RUN_SELF(true, 4, 32, 0) RUN_SELF(true, 2, 16, 0) RUN_SELF(true, 8, 16, 0)
    RUN_SELF(true, 4, 16, 20) RUN_SELF(true, 4, 16, 40) RUN_SELF(true, 4, 24, 20) RUN_SELF(true, 4, 24, 40) RUN_SELF(true, 4, 32, 20)
#define RUN_LOOP(U, G, NA, bps) { float ms = timeit([&] { k_loop<U, G, NA><<<sms * bps, 256>>>(d, x, out, ticket, nchunks, K); }, ticket); \
        printf("loop U=%d gather=%d na=%d warps/SM=%2d: %.1f us  %.0f GB/s\n", U, G, NA, bps * 8, ms * 1e3, bytes / ms / 1e6); }
    RUN_LOOP(2, true, true, 4) RUN_LOOP(2, true, true, 8) RUN_LOOP(4, true, true, 4) RUN_LOOP(4, true, true, 8) RUN_LOOP(1, true, true, 8)
    RUN_LOOP(4, false, true, 8) RUN_LOOP(2, true, false, 8)
    return 0;
}
