// mb8: can a kernel deliver its output to HOST memory (mapped pinned buffer) as fast as a copy engine does?
// Decides whether K3 should write the cut batch straight into the caller's pinned view (no separate download).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mb8 mb8.cu && ./mb8
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <typename T> __global__ void put(const T* __restrict__ s, T* __restrict__ d, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
}
// tile pattern of the cut kernel: a block owns a contiguous 16 KB piece and writes it after some compute
__global__ void put_tiles(const double* __restrict__ s, double* __restrict__ d, size_t n, int spin) {
    const size_t tile = 2048;
    for (size_t t = blockIdx.x; t * tile < n; t += gridDim.x) {
        double acc = 0;
        for (int k = 0; k < spin; ++k) acc = acc * 1.0000001 + 1e-9;
        for (size_t i = t * tile + threadIdx.x; i < (t + 1) * tile && i < n; i += blockDim.x) d[i] = s[i] + (acc > 1e30 ? 1.0 : 0.0);
    }
}

int main() {
    const size_t bytes = 20u << 20;
    void *dsrc, *hbuf, *hdev;
    CK(cudaMalloc(&dsrc, bytes)); CK(cudaMemset(dsrc, 1, bytes));
    CK(cudaHostAlloc(&hbuf, bytes, cudaHostAllocMapped)); CK(cudaHostGetDevicePointer(&hdev, hbuf, 0));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    auto time = [&](const char* name, auto fn) {
        float best = 1e9f, sum = 0;
        for (int it = 0; it < 12; ++it) {
            CK(cudaEventRecord(a, st)); fn(); CK(cudaEventRecord(b, st)); CK(cudaStreamSynchronize(st));
            float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (it >= 2) { sum += ms; if (ms < best) best = ms; }
        }
        printf("%-44s mean %.3f ms (%.1f GB/s)  best %.3f ms (%.1f GB/s)\n", name, sum / 10, bytes / (sum / 10) * 1e-6, best, bytes / best * 1e-6);
    };
    time("cudaMemcpyAsync D2H 20 MB", [&] { CK(cudaMemcpyAsync(hbuf, dsrc, bytes, cudaMemcpyDeviceToHost, st)); });
    for (int grid : {16, 37, 74, 148, 296, 592}) for (int blk : {128, 512}) {
        char nm[96];
        snprintf(nm, sizeof nm, "kernel uint4 stores, grid %d x %d", grid, blk);
        time(nm, [&] { put<uint4><<<grid, blk, 0, st>>>((const uint4*)dsrc, (uint4*)hdev, bytes / 16); });
        snprintf(nm, sizeof nm, "kernel double stores, grid %d x %d", grid, blk);
        time(nm, [&] { put<double><<<grid, blk, 0, st>>>((const double*)dsrc, (double*)hdev, bytes / 8); });
    }
    for (int grid : {148, 444}) for (int spin : {0, 2000}) {
        char nm[96]; snprintf(nm, sizeof nm, "tiles of 16 KB, grid %d x 128, spin %d", grid, spin);
        time(nm, [&] { put_tiles<<<grid, 128, 0, st>>>((const double*)dsrc, (double*)hdev, bytes / 8, spin); });
    }
    // alignment: the same contiguous stream, shifted by 8 / 16 / 24 / 64 bytes against the 32-byte sectors and 128-byte lines
    for (int sh : {1, 2, 3, 8}) {
        char nm[96]; snprintf(nm, sizeof nm, "double stores shifted by %d bytes, 444 x 128", 8 * sh);
        time(nm, [&] { put<double><<<444, 128, 0, st>>>((const double*)dsrc, (double*)hdev + sh, bytes / 8 - 16); });
    }
    for (int sh : {1, 3}) {
        char nm[96]; snprintf(nm, sizeof nm, "int stores shifted by %d bytes, 444 x 128", 4 * sh);
        time(nm, [&] { put<int><<<444, 128, 0, st>>>((const int*)dsrc, (int*)hdev + sh, bytes / 4 - 16); });
    }
    time("kernel int stores, grid 148 x 512", [&] { put<int><<<148, 512, 0, st>>>((const int*)dsrc, (int*)hdev, bytes / 4); });
    // end to end: kernel into a DEVICE buffer + copy, against kernel straight into host memory
    void* ddst; CK(cudaMalloc(&ddst, bytes));
    time("tiles -> device, then D2H copy", [&] { put_tiles<<<444, 128, 0, st>>>((const double*)dsrc, (double*)ddst, bytes / 8, 2000); CK(cudaMemcpyAsync(hbuf, ddst, bytes, cudaMemcpyDeviceToHost, st)); });
    time("tiles -> host directly", [&] { put_tiles<<<444, 128, 0, st>>>((const double*)dsrc, (double*)hdev, bytes / 8, 2000); });
    printf("check %d\n", ((unsigned char*)hbuf)[12345]);
    return 0;
}
