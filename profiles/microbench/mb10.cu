// mb10: x* staged in (distributed) shared memory instead of gathered from the L2.  Random 8-byte gathers, indices streamed as in mb9.
//   (a) table in the L2 (__ldg)                       -- mb9's baseline
//   (b) 80 KB table in every CTA's shared memory       -- the QCQP configuration (10^4 variables)
//   (c) 800 KB table spread over a cluster of 8 / 16 CTAs (100 / 50 KB each), gathers through DSMEM (ld.shared::cluster)
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mb10 mb10.cu && ./mb10
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void ldidx(const int* p, int (&c)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7]) : "l"(p));
}
__global__ void __launch_bounds__(512, 1) gather_l2(const int* __restrict__ idx, const double* __restrict__ x, double* out, size_t nvec) {
    double acc = 0;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
        int c[8]; ldidx(idx + 8 * v, c); double g[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = __ldg(x + c[k]);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += g[k];
    }
    if (acc == 1.2345) out[0] = acc;
}
__global__ void __launch_bounds__(512, 1) gather_smem(const int* __restrict__ idx, const double* __restrict__ x, double* out, size_t nvec, int nvar) {
    extern __shared__ double sx[];
    for (int i = threadIdx.x; i < nvar; i += blockDim.x) sx[i] = x[i];
    __syncthreads();
    double acc = 0;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
        int c[8]; ldidx(idx + 8 * v, c); double g[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = sx[c[k] % nvar];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += g[k];
    }
    if (acc == 1.2345) out[0] = acc;
}
// table of nvar doubles split over the cluster: CTA r holds [r * slice, (r + 1) * slice)
__global__ void __launch_bounds__(512, 1) gather_dsmem(const int* __restrict__ idx, const double* __restrict__ x, double* out, size_t nvec, int nvar, int slice) {
    extern __shared__ double sx[];
    cg::cluster_group cl = cg::this_cluster();
    const unsigned r = cl.block_rank(), C = cl.num_blocks();
    for (int i = threadIdx.x; i < slice; i += blockDim.x) { const int j = (int)r * slice + i; sx[i] = j < nvar ? x[j] : 0.0; }
    cl.sync();
    unsigned base; asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(base) : "l"(sx));
    double acc = 0;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
        int c[8]; ldidx(idx + 8 * v, c); double g[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const unsigned owner = (unsigned)c[k] / (unsigned)slice, off = (unsigned)c[k] - owner * (unsigned)slice;
            unsigned ra; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(base + 8u * off), "r"(owner));
            asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(g[k]) : "r"(ra));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += g[k];
    }
    if (acc == 1.2345) out[0] = acc;
    cl.sync();          // nobody leaves while a peer may still read its slice
    (void)C;
}

int main() {
    const size_t ngat = 10240000, nvec = ngat / 8;
    int sms = 0, khz = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0)); CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    double *x, *out; int* idx;
    CK(cudaMalloc(&x, 8 * 100000)); CK(cudaMemset(x, 0, 8 * 100000)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&idx, 4 * ngat));
    auto fill = [&](size_t nvar) { std::vector<int> h(ngat); uint64_t s = 88172645463325252ull; for (size_t i = 0; i < ngat; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % nvar); } CK(cudaMemcpy(idx, h.data(), 4 * ngat, cudaMemcpyHostToDevice)); };
    auto timeit = [&](const char* name, int blocks, auto launch) {
        float best = 1e9f;
        for (int it = 0; it < 8; ++it) { CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (it >= 2 && ms < best) best = ms; }
        CK(cudaGetLastError());
        printf("%-64s %4d CTAs %7.1f us  %.2f gathers/clk/SM(of %d)\n", name, blocks, 1e3 * best, ngat / (best * 1e-3) / sms / (khz * 1e3), sms);
    };
    fill(100000);
    timeit("L2 gathers, 10^5 variables (__ldg)", sms, [&] { gather_l2<<<sms, 512>>>(idx, x, out, nvec); });
    fill(10000);
    timeit("L2 gathers, 10^4 variables (__ldg)", sms, [&] { gather_l2<<<sms, 512>>>(idx, x, out, nvec); });
    CK(cudaFuncSetAttribute(gather_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 80000));
    timeit("shared-memory table, 10^4 variables (80 KB per CTA)", sms, [&] { gather_smem<<<sms, 512, 80000>>>(idx, x, out, nvec, 10000); });
    fill(100000);
    for (int C : {8, 16, 4}) {
        const int slice = (100000 + C - 1) / C; const size_t smem = 8 * (size_t)slice;
        CK(cudaFuncSetAttribute(gather_dsmem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (C > 8) CK(cudaFuncSetAttribute(gather_dsmem, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg = {}; cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int ncl = 0; cfg.gridDim = dim3(C);
        if (cudaOccupancyMaxActiveClusters(&ncl, gather_dsmem, &cfg) != cudaSuccess || ncl == 0) { printf("cluster of %d: not launchable (%s)\n", C, cudaGetErrorString(cudaGetLastError())); continue; }
        cfg.gridDim = dim3(ncl * C);
        char nm[128]; snprintf(nm, sizeof nm, "DSMEM table, 10^5 variables, cluster of %d (%d KB per CTA)", C, (int)(smem / 1000));
        timeit(nm, ncl * C, [&] { CK(cudaLaunchKernelEx(&cfg, gather_dsmem, (const int*)idx, (const double*)x, out, nvec, 100000, slice)); });
    }
    return 0;
}
