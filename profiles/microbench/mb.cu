// Design microbenchmarks for the separation kernels (sm_100a).
// Measures the four rates DESIGN.md's per-row cycle budget is built on:
//   1. scattered fp64 gathers of x* from L2 (n = 1e4, 1e5, 1e6 doubles)
//   2. the same gathers from shared memory (n = 1e4) and from cluster DSMEM (n = 1e5)
//   3. dependent DFMA chains at several warps/SM and ILP levels
//   4. streaming reads: LDG.128 vs per-warp cp.async.bulk (TMA 1D) into shared memory
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o mb mb.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static inline uint64_t splitmix(uint64_t& s) { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

// ---- 1. L2 gather: each thread reads G indices (coalesced) and gathers x[idx]
template <int G>
__global__ void gather_l2(const double* __restrict__ x, const int* __restrict__ idx, double* out, int nthreads_total) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads_total) return;
    double acc = 0;
#pragma unroll
    for (int k = 0; k < G; ++k) acc += __ldg(&x[idx[(size_t)k * nthreads_total + t]]);
    if (acc == 12345.678) out[0] = acc;
}

// ---- 2. smem gather: x (n doubles) staged into smem by each CTA, rows grid-strided
template <int G>
__global__ void gather_smem(const double* __restrict__ x, int n, const int* __restrict__ idx, double* out, int nrows) {
    extern __shared__ double sx[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sx[i] = x[i];
    __syncthreads();
    double acc = 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nrows; t += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < G; ++k) acc += sx[idx[(size_t)k * nrows + t]];
    }
    if (acc == 12345.678) out[0] = acc;
}

// ---- 3. DSMEM gather: cluster of CS CTAs, x split evenly across the cluster's smem
template <int G, int CS>
__global__ void gather_dsmem(const double* __restrict__ x, int n, const int* __restrict__ idx, double* out, int nrows) {
    extern __shared__ double sx[];
    cg::cluster_group cl = cg::this_cluster();
    int per = (n + CS - 1) / CS;
    int r = cl.block_rank();
    for (int i = threadIdx.x; i < per; i += blockDim.x) { int gi = r * per + i; sx[i] = gi < n ? x[gi] : 0.0; }
    cl.sync();
    const double* peers[CS];
#pragma unroll
    for (int c = 0; c < CS; ++c) peers[c] = cl.map_shared_rank(sx, c);
    double acc = 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nrows; t += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < G; ++k) {
            int j = idx[(size_t)k * nrows + t];
            int c = j / per, o = j - c * per;
            const double* p = peers[0];
#pragma unroll
            for (int q = 1; q < CS; ++q) if (c == q) p = peers[q];
            acc += p[o];
        }
    }
    if (acc == 12345.678) out[0] = acc;
    cl.sync();
}

// ---- 4. DFMA chains
template <int ILP>
__global__ void dfma(double* out, int iters, double a, double b) {
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 12345.678) out[0] = s;
}

// ---- 5a. streaming LDG.128
__global__ void stream_ldg(const double2* __restrict__ p, size_t n2, double* out) {
    double acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        double2 v = __ldg(&p[i]); acc += v.x + v.y;
    }
    if (acc == 12345.678) out[0] = acc;
}

// ---- 5b. per-warp TMA 1D bulk copies into smem with an mbarrier, CHB bytes per chunk, NST stages per warp
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
template <int NST>
__global__ void stream_tma(const char* __restrict__ p, size_t nchunks, int chb, double* out) {
    extern __shared__ __align__(128) char sm[];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    uint64_t* bars = (uint64_t*)sm;                       // nw*NST barriers
    char* buf = sm + 1024 + (size_t)warp * NST * chb;
    if (lane == 0) for (int s = 0; s < NST; ++s) mbar_init(&bars[warp * NST + s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    size_t gw = (size_t)blockIdx.x * nw + warp, tw = (size_t)gridDim.x * nw;
    double acc = 0;
    // prologue
    size_t c = gw; int issued = 0;
    size_t cs[NST];
    for (int s = 0; s < NST; ++s) { cs[s] = c + (size_t)s * tw; }
    if (lane == 0) for (int s = 0; s < NST; ++s) if (cs[s] < nchunks) { mbar_expect(&bars[warp * NST + s], chb); bulk_g2s(buf + (size_t)s * chb, p + cs[s] * chb, chb, &bars[warp * NST + s]); }
    uint32_t phase = 0; int s = 0;
    for (; c < nchunks; c += tw) {
        mbar_wait(&bars[warp * NST + s], phase);
        const double* d = (const double*)(buf + (size_t)s * chb);
        for (int i = lane; i < chb / 8; i += 32) acc += d[i];
        __syncwarp();
        size_t nx = c + (size_t)NST * tw;
        if (lane == 0 && nx < nchunks) { mbar_expect(&bars[warp * NST + s], chb); bulk_g2s(buf + (size_t)s * chb, p + nx * chb, chb, &bars[warp * NST + s]); }
        if (++s == NST) { s = 0; phase ^= 1; }
    }
    (void)issued;
    if (acc == 12345.678) out[0] = acc;
}

template <class F> float timeit(F f, int reps = 5) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms; }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    int nsm = pr.multiProcessorCount;
    printf("device %s sms %d smem/blk optin %zu clock %d kHz\n", pr.name, nsm, pr.sharedMemPerBlockOptin, pr.clockRate);
    double* out; CK(cudaMalloc(&out, 64));
    const int G = 10;
    // ---- gathers
    for (int n : {10000, 100000, 1000000}) {
        int nrows = 1 << 20;
        std::vector<double> hx(n); for (int i = 0; i < n; ++i) hx[i] = i * 1e-6;
        std::vector<int> hi((size_t)G * nrows); uint64_t s = 42 + n; for (auto& v : hi) v = (int)(splitmix(s) % (uint64_t)n);
        double* dx; int* di; CK(cudaMalloc(&dx, n * 8)); CK(cudaMalloc(&di, hi.size() * 4));
        CK(cudaMemcpy(dx, hx.data(), n * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(di, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice));
        for (int bs : {128, 256, 512}) {
            float ms = timeit([&] { gather_l2<G><<<(nrows + bs - 1) / bs, bs>>>(dx, di, out, nrows); });
            double gps = (double)G * nrows / (ms * 1e-3);
            printf("gather_l2 n=%d bs=%d: %.3f ms  %.2f Ggather/s  %.3f gathers/clk/SM @1.9GHz\n", n, bs, ms, gps * 1e-9, gps / nsm / 1.9e9);
        }
        if ((size_t)n * 8 <= 200 * 1024) {
            CK(cudaFuncSetAttribute(gather_smem<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 8));
            for (int bs : {256, 512, 1024}) {
                float ms = timeit([&] { gather_smem<G><<<nsm, bs, n * 8>>>(dx, n, di, out, nrows); });
                double gps = (double)G * nrows / (ms * 1e-3);
                printf("gather_smem n=%d bs=%d: %.3f ms  %.2f Ggather/s  %.3f gathers/clk/SM\n", n, bs, ms, gps * 1e-9, gps / nsm / 1.9e9);
            }
        }
        if (n == 100000) {
            constexpr int CS = 8;
            int per = (n + CS - 1) / CS; size_t sm = (size_t)per * 8;
            CK(cudaFuncSetAttribute(gather_dsmem<G, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            for (int bs : {256, 512, 1024}) {
                cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3((nsm / CS) * CS); cfg.blockDim = dim3(bs); cfg.dynamicSmemBytes = sm;
                cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                float ms = timeit([&] { CK(cudaLaunchKernelEx(&cfg, gather_dsmem<G, CS>, (const double*)dx, n, (const int*)di, out, nrows)); });
                double gps = (double)G * nrows / (ms * 1e-3);
                printf("gather_dsmem n=%d cs=%d bs=%d: %.3f ms  %.2f Ggather/s  %.3f gathers/clk/SM\n", n, CS, bs, ms, gps * 1e-9, gps / nsm / 1.9e9);
            }
        }
        CK(cudaFree(dx)); CK(cudaFree(di));
    }
    // ---- DFMA
    {
        int iters = 4096;
        for (int wps : {4, 8, 16, 32}) {
            int bs = wps * 32;
            float m1 = timeit([&] { dfma<1><<<nsm, bs>>>(out, iters, 1.0000001, 1e-9); });
            float m2 = timeit([&] { dfma<2><<<nsm, bs>>>(out, iters, 1.0000001, 1e-9); });
            float m4 = timeit([&] { dfma<4><<<nsm, bs>>>(out, iters, 1.0000001, 1e-9); });
            auto rate = [&](float ms, int ilp) { return (double)nsm * bs * iters * ilp / (ms * 1e-3) * 1e-12; };
            printf("dfma warps/SM=%d: ILP1 %.2f  ILP2 %.2f  ILP4 %.2f TDFMA/s (x2 = TFLOP/s)\n", wps, rate(m1, 1), rate(m2, 2), rate(m4, 4));
        }
    }
    // ---- streaming
    {
        size_t bytes = (size_t)1 << 30; char* p; CK(cudaMalloc(&p, bytes)); CK(cudaMemset(p, 0, bytes));
        for (int bs : {256, 512}) for (int mult : {2, 4, 8}) {
            float ms = timeit([&] { stream_ldg<<<nsm * mult, bs>>>((const double2*)p, bytes / 16, out); });
            printf("stream_ldg128 grid=%dx bs=%d: %.3f ms %.1f GB/s\n", mult, bs, ms, bytes / (ms * 1e-3) * 1e-9);
        }
        for (int chb : {4096, 8192, 16384}) for (int nw : {4, 8, 16}) {
            {
                constexpr int NST = 1; size_t sm = 1024 + (size_t)nw * NST * chb; if (sm > 220 * 1024) continue;
                CK(cudaFuncSetAttribute(stream_tma<NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                float ms = timeit([&] { stream_tma<NST><<<nsm, nw * 32, sm>>>(p, bytes / chb, chb, out); });
                printf("stream_tma chunk=%d warps=%d stages=%d: %.3f ms %.1f GB/s\n", chb, nw, NST, ms, bytes / (ms * 1e-3) * 1e-9);
            }
            {
                constexpr int NST = 2; size_t sm = 1024 + (size_t)nw * NST * chb; if (sm > 220 * 1024) continue;
                CK(cudaFuncSetAttribute(stream_tma<NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                float ms = timeit([&] { stream_tma<NST><<<nsm, nw * 32, sm>>>(p, bytes / chb, chb, out); });
                printf("stream_tma chunk=%d warps=%d stages=%d: %.3f ms %.1f GB/s\n", chb, nw, NST, ms, bytes / (ms * 1e-3) * 1e-9);
            }
        }
        CK(cudaFree(p));
    }
    printf("done\n");
    return 0;
}
