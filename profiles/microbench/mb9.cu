// mb9: does the cache operator of the x* gathers change the gather rate?  (K1's floor is ~0.9 gathers/clk/SM with __ldg: every miss
// holds an L1 line until the sector returns from the L2.)  Random 8-byte gathers from a 0.8 MB table (L2 resident), indices streamed.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mb9 mb9.cu && ./mb9
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE> __device__ __forceinline__ double gat(const double* p) {
    double v;
    if (MODE == 0) v = __ldg(p);
    else if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else if (MODE == 2) v = __ldcg(p);
    else if (MODE == 3) asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else if (MODE == 4) asm volatile("ld.global.nc.L1::evict_first.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
// each thread: G index vectors of 8 (one 256-bit load), 8 gathers in flight per vector
template <int MODE, int IDXNA> __global__ void __launch_bounds__(512, 1) gather(const int* __restrict__ idx, const double* __restrict__ x, double* out, size_t nvec) {
    double acc = 0;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
        int c[8];
        if (IDXNA) asm volatile("ld.global.nc.L1::no_allocate.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7]) : "l"(idx + 8 * v));
        else asm volatile("ld.global.nc.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7]) : "l"(idx + 8 * v));
        double g[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = gat<MODE>(x + c[k]);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += g[k];
    }
    if (acc == 1.2345) out[0] = acc;
}

int main() {
    const size_t nvar = 100000, ngat = 10240000, nvec = ngat / 8;
    std::vector<int> h(ngat); uint64_t s = 88172645463325252ull;
    for (size_t i = 0; i < ngat; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % nvar); }
    int* idx; double *x, *out;
    CK(cudaMalloc(&idx, 4 * ngat)); CK(cudaMalloc(&x, 8 * nvar)); CK(cudaMalloc(&out, 64));
    CK(cudaMemcpy(idx, h.data(), 4 * ngat, cudaMemcpyHostToDevice)); CK(cudaMemset(x, 0, 8 * nvar));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    int sms = 0, khz = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0)); CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    auto run = [&](const char* name, auto kern) {
        float best = 1e9f;
        for (int it = 0; it < 8; ++it) {
            CK(cudaEventRecord(a)); kern<<<sms, 512>>>(idx, x, out, nvec); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
            float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (it >= 2 && ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("%-58s %7.1f us  %.2f gathers/clk/SM\n", name, 1e3 * best, ngat / (best * 1e-3) / sms / (khz * 1e3));
    };
    run("__ldg (ld.global.nc)", gather<0, 0>);
    run("ld.global.nc.L1::no_allocate", gather<1, 0>);
    run("__ldcg (ld.global.cg: L2 only)", gather<2, 0>);
    run("ld.global.nc.L1::evict_last", gather<3, 0>);
    run("ld.global.nc.L1::evict_first", gather<4, 0>);
    run("ld.global.L1::no_allocate", gather<5, 0>);
    run("__ldg, index stream L1::no_allocate", gather<0, 1>);
    run("gathers L1::no_allocate, index stream L1::no_allocate", gather<1, 1>);
    run("gathers evict_last, index stream L1::no_allocate", gather<3, 1>);
    run("gathers __ldcg, index stream L1::no_allocate", gather<2, 1>);
    return 0;
}
