// mb5.cu -- round 2 microbenchmark: what separates the bare register-resident row (mb4 `reg`: 59.5 us) from the product's
// family kernel (eval only 86 us, with cuts 112 us)?  The product's features are added one by one:
//   mixed classes K = 4..16 with one specialised code path per class, rows scattered over a sigma-window (chunk_rows indirection,
//   scattered g / sel stores), bounds + violation test, and three cut schemes for the ~10 % selected rows:
//     CUT 1  stream the row again (constants, columns, x* come back from L1), recompute the exponentials, entries in Jacobian
//            order through the row's ORDER word (sorted position -> term): no scratch, no kept registers
//     CUT 2  keep the exponentials in registers, re-read c and x* from L1, scatter by RANK word through shared memory
//     CUT 3  round 1's scheme: keep exp and x*, re-read c
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -lineinfo -o mb5 mb5.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "../../katana.jl_b200/csrc/ktn_math.h"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

#define NCLS 17
#define PASS 16
struct Params {
    const unsigned char* blob; const double* x; const int32_t* chunk_rows; const double* chunk_lb; const double* chunk_ub; const long long* jac_ptr;
    double* g_row; double* b_row; uint32_t* sel; double* stage; unsigned long long* blk_cnt; unsigned* ticket; unsigned nchunks;
    unsigned cls_begin[NCLS + 1]; unsigned long long cls_off[NCLS];
    double ftol;
};

template <bool NA> __device__ __forceinline__ double ldd(const double* p) {
    double v; if (NA) asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); else v = __ldg(p); return v;
}
template <bool NA> __device__ __forceinline__ int ldi(const int* p) {
    int v; if (NA) asm("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p)); else v = __ldg(p); return v;
}

__device__ __forceinline__ void count_selected(const Params& p, unsigned act, int32_t row, uint32_t nnz) {
    const uint32_t blk = (uint32_t)row >> 12;
    const unsigned grp = __match_any_sync(act, blk);
    const uint32_t nz = __reduce_add_sync(grp, nnz);
    if ((threadIdx.x & 31u) == (uint32_t)(__ffs(grp) - 1)) atomicAdd(p.blk_cnt + blk, (1ull << 48) * __popc(grp) + nz);
}

// FEAT bits: 1 = rows scattered (chunk_rows, g_row[row]); 2 = bounds + test + sel store; CUT as above (needs FEAT 3)
template <int N, int FEAT, int CUT, bool NA>
__device__ __forceinline__ void chunk(const Params& p, unsigned c, unsigned lane, double* scratch) {
    const unsigned char* b = p.blob + p.cls_off[N] + (size_t)(c - p.cls_begin[N]) * (640u * N + 256u);
    const double* C = reinterpret_cast<const double*>(b) + lane;
    const int32_t* col = reinterpret_cast<const int32_t*>(b + 512u * N) + lane;
    const unsigned slot = c * 32u + lane;
    double cc[N], dd[N]; int cl[N];
#pragma unroll
    for (int u = 0; u < N; ++u) { cc[u] = ldd<NA>(C + (2 * u) * 32); dd[u] = ldd<NA>(C + (2 * u + 1) * 32); cl[u] = ldi<NA>(col + u * 32); }
    int32_t row = (int32_t)slot; double lb = 0.0, ub = 0.0;
    if (FEAT & 1) row = ldi<NA>(p.chunk_rows + slot);
    if (FEAT & 2) { lb = ldd<NA>(p.chunk_lb + slot); ub = ldd<NA>(p.chunk_ub + slot); }
    double xs[N], e[N];
#pragma unroll
    for (int u = 0; u < N; ++u) xs[u] = __ldg(p.x + cl[u]);
    bool slow = false;
#pragma unroll
    for (int u = 0; u < N; ++u) { const double a = (0.0 + cc[u] * xs[u]) + dd[u]; e[u] = ktn_exp_fast(a); slow = slow || !ktn_exp_is_fast(a); }
    if (slow) {
#pragma unroll
        for (int u = 0; u < N; ++u) { const double a = (0.0 + cc[u] * xs[u]) + dd[u]; if (!ktn_exp_is_fast(a)) e[u] = ktn_exp_slow(a); }
    }
    double acc = 0.0;
#pragma unroll
    for (int u = 0; u < N; ++u) acc = acc + e[u];
    const double g = ktn_log(acc);
    if (row >= 0) p.g_row[row] = g;
    if (!(FEAT & 2)) return;
    const bool selected = row >= 0 && !((g >= lb - p.ftol) && (g <= ub + p.ftol));
    if (row >= 0 && !selected) p.sel[row] = 0u;
    unsigned selm = __ballot_sync(0xffffffffu, selected);
    if (CUT == 0) { if (selected) p.sel[row] = N; return; }
    if (CUT == 1) {
        if (selected) {
            const uint64_t ow = __ldg(reinterpret_cast<const unsigned long long*>(b + 640u * N) + lane);
            const double adj = 1.0 / acc;
            double bb = g, mx = -ktn_inf(), mn = ktn_inf();
            double* out = p.stage + p.jac_ptr[row];
#pragma unroll
            for (int q0 = 0; q0 < N; q0 += 4) {
                double jv[4], xv[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) if (q0 + k < N) {
                    const unsigned u = (unsigned)(ow >> (4 * (q0 + k))) & 15u;
                    const double c1 = __ldg(C + (2 * u) * 32), d1 = __ldg(C + (2 * u + 1) * 32);
                    xv[k] = __ldg(p.x + __ldg(col + u * 32));
                    jv[k] = 0.0 + (adj * ktn_exp((0.0 + c1 * xv[k]) + d1)) * c1;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) if (q0 + k < N) { bb = bb + (-xv[k]) * jv[k]; out[q0 + k] = jv[k]; mx = fmax(mx, jv[k]); mn = fmin(mn, jv[k]); }
            }
            p.b_row[row] = bb; p.sel[row] = N | ((mn + 1e9 < mx) ? 0x40000000u : 0u);
            count_selected(p, selm, row, N);
        }
        return;
    }
    // CUT 2 / 3: PASS selected lanes at a time share the warp's scratch
    while (selm) {
        const uint32_t cut = __fns(selm, 0, PASS + 1);
        const unsigned grp = cut == 0xffffffffu ? selm : (selm & ((1u << cut) - 1u));
        if ((grp >> lane) & 1u) {
            double* t = scratch + __popc(grp & ((1u << lane) - 1u));
            double* out = p.stage + p.jac_ptr[row];
            const uint64_t rw = __ldg(reinterpret_cast<const unsigned long long*>(b + 640u * N) + lane);
            const double adj = 1.0 / acc;
            double mx = -ktn_inf(), mn = ktn_inf();
#pragma unroll
            for (int u = 0; u < N; ++u) {
                const double c1 = __ldg(C + (2 * u) * 32);
                const double xv = CUT == 2 ? __ldg(p.x + __ldg(col + u * 32)) : xs[u];
                const double jv = 0.0 + (adj * e[u]) * c1;
                const unsigned q = (unsigned)(rw >> (4 * u)) & 15u;
                t[q * PASS] = (-xv) * jv; out[q] = jv; mx = fmax(mx, jv); mn = fmin(mn, jv);
            }
            double bb = g;
#pragma unroll
            for (int q = 0; q < N; ++q) bb = bb + t[q * PASS];
            p.b_row[row] = bb; p.sel[row] = N | ((mn + 1e9 < mx) ? 0x40000000u : 0u);
            count_selected(p, grp, row, N);
        }
        __syncwarp();
        selm &= ~grp;
    }
}

template <int FEAT, int CUT, bool NA>
__device__ __forceinline__ void dispatch(const Params& p, unsigned cls, unsigned c, unsigned lane, double* scratch) {
    switch (cls) {
#define CASE(n) case n: chunk<n, FEAT, CUT, NA>(p, c, lane, scratch); break;
        CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16)
#undef CASE
        default: break;
    }
}

// the product's scheduling: per-class ticket counters, the warps of an SM start in the same class and move on together
template <int FEAT, int CUT, bool NA, int WARPS, int BPS>
__global__ void __launch_bounds__(WARPS * 32, BPS) k_fam(const Params p) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* scratch = reinterpret_cast<double*>(smem) + (size_t)(threadIdx.x >> 5) * (16 * PASS);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned my_n = lane < NCLS ? p.cls_begin[lane + 1] - p.cls_begin[lane] : 0u;
    unsigned cls = 0;
    {
        unsigned smid, nsm;
        asm("mov.u32 %0, %%smid;" : "=r"(smid)); asm("mov.u32 %0, %%nsmid;" : "=r"(nsm));
        unsigned long long total = 0, acc = 0;
        for (unsigned k = 0; k < NCLS; ++k) total += (unsigned long long)(p.cls_begin[k + 1] - p.cls_begin[k]) * (k + 3u);
        const unsigned long long target = (total * (2ull * smid + 1ull)) / (2ull * nsm);
        for (unsigned k = 0; k < NCLS; ++k) { acc += (unsigned long long)(p.cls_begin[k + 1] - p.cls_begin[k]) * (k + 3u); if (acc > target) { cls = k; break; } }
    }
    unsigned n_cls = p.cls_begin[cls + 1] - p.cls_begin[cls];
    for (;;) {
        unsigned cur = 0;
        if (lane == 0) cur = atomicAdd(&p.ticket[cls], 1u);
        cur = __shfl_sync(0xffffffffu, cur, 0);
        if (cur >= n_cls) {
            const bool live = lane < NCLS && my_n > 0 && __ldcg(&p.ticket[lane]) < my_n;
            const unsigned livem = __ballot_sync(0xffffffffu, live);
            if (!livem) return;
            const unsigned ahead = livem & ~((2u << cls) - 1u);
            cls = (unsigned)__ffs(ahead ? ahead : livem) - 1u;
            n_cls = p.cls_begin[cls + 1] - p.cls_begin[cls];
            continue;
        }
        dispatch<FEAT, CUT, NA>(p, cls, p.cls_begin[cls] + cur, lane, scratch);
    }
}

static inline uint64_t splitmix(uint64_t& s) { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static inline double u01(uint64_t& s) { return (double)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0); }

struct Data { Params p; size_t bytes; size_t m; std::vector<double> gref; };

static Data make(int n, double vfrac) {
    Data D; memset(&D.p, 0, sizeof D.p);
    const size_t m = 1000000, sigma = 8192; D.m = m;
    // chunks: per window, per class, 32 rows each (like ktn_compile.cpp), then sorted by class
    struct Ch { int k; int rows[32]; };
    std::vector<Ch> chs;
    for (size_t w0 = 0; w0 < m; w0 += sigma) {
        const size_t w1 = std::min(m, w0 + sigma);
        for (int k = 4; k <= 16; ++k) {
            Ch c; c.k = k; int nr = 0;
            for (size_t r = w0; r < w1; ++r) if ((int)(4 + r % 13) == k) { c.rows[nr++] = (int)r; if (nr == 32) { chs.push_back(c); nr = 0; } }
            if (nr) { for (int q = nr; q < 32; ++q) c.rows[q] = -1; chs.push_back(c); }
        }
    }
    std::stable_sort(chs.begin(), chs.end(), [](const Ch& a, const Ch& b) { return a.k < b.k; });
    const unsigned nchunks = (unsigned)chs.size();
    std::vector<unsigned> cnt(NCLS + 1, 0);
    for (auto& c : chs) cnt[c.k]++;
    unsigned at = 0; unsigned long long off = 0;
    for (int k = 0; k < NCLS; ++k) { D.p.cls_begin[k] = at; D.p.cls_off[k] = off; at += cnt[k]; off += (unsigned long long)cnt[k] * (640ull * k + 256ull); }
    D.p.cls_begin[NCLS] = at; D.p.nchunks = nchunks;
    std::vector<unsigned char> h(off + 256);
    std::vector<double> x(n), lb((size_t)nchunks * 32, -1e300), ub((size_t)nchunks * 32, 0.0), g(m, 0.0);
    std::vector<int32_t> rows((size_t)nchunks * 32);
    std::vector<long long> jp(m + 1, 0);
    for (size_t r = 0; r < m; ++r) jp[r + 1] = jp[r] + (long long)(4 + r % 13);
    uint64_t s = 7;
    for (int i = 0; i < n; ++i) x[i] = 4.0 * u01(s) - 2.0;
    for (unsigned c = 0; c < nchunks; ++c) {
        const int k = chs[c].k;
        unsigned char* b = h.data() + D.p.cls_off[k] + (size_t)(c - D.p.cls_begin[k]) * (640u * k + 256u);
        double* C = reinterpret_cast<double*>(b); int32_t* col = reinterpret_cast<int32_t*>(b + 512u * k); uint64_t* ow = reinterpret_cast<uint64_t*>(b + 640u * k);
        for (int lane = 0; lane < 32; ++lane) {
            double acc = 0.0; uint64_t w = 0;
            for (int u = 0; u < k; ++u) {
                const int cl = (int)(splitmix(s) % (uint64_t)n); const double cc = 2.0 * u01(s) - 1.0, dd = 2.0 * u01(s) - 1.0;
                col[u * 32 + lane] = cl; C[(2 * u) * 32 + lane] = cc; C[(2 * u + 1) * 32 + lane] = dd;
                acc = acc + ktn_exp((0.0 + cc * x[cl]) + dd); w |= (uint64_t)((u * 7 + 3) % k) << (4 * u);
            }
            ow[lane] = w; rows[(size_t)c * 32 + lane] = chs[c].rows[lane];
            if (chs[c].rows[lane] >= 0) g[chs[c].rows[lane]] = ktn_log(acc);
        }
    }
    // ub at the (1 - v) quantile of g
    { std::vector<double> sg(g); std::sort(sg.begin(), sg.end()); const double thr = sg[(size_t)((1.0 - vfrac) * (m - 1))];
      for (size_t i = 0; i < ub.size(); ++i) ub[i] = thr; }
    D.gref = g;
    D.bytes = off + (size_t)nchunks * 32 * 20 + 8 * (size_t)n;
    unsigned char* d; double *dx, *dlb, *dub, *dg, *dbr, *dst; int32_t* drows; long long* djp; uint32_t* dsel; unsigned long long* dblk; unsigned* ticket;
    CK(cudaMalloc(&d, h.size())); CK(cudaMalloc(&dx, 8 * (size_t)n)); CK(cudaMalloc(&dlb, 8 * lb.size())); CK(cudaMalloc(&dub, 8 * ub.size()));
    CK(cudaMalloc(&drows, 4 * rows.size())); CK(cudaMalloc(&djp, 8 * jp.size())); CK(cudaMalloc(&dg, 8 * m)); CK(cudaMalloc(&dbr, 8 * m)); CK(cudaMalloc(&dsel, 4 * m));
    CK(cudaMalloc(&dst, 8 * (size_t)jp[m])); CK(cudaMalloc(&dblk, 8 * 1024)); CK(cudaMalloc(&ticket, 4 * 64));
    CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dx, x.data(), 8 * (size_t)n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dlb, lb.data(), 8 * lb.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dub, ub.data(), 8 * ub.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(drows, rows.data(), 4 * rows.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(djp, jp.data(), 8 * jp.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dblk, 0, 8 * 1024));
    D.p.blob = d; D.p.x = dx; D.p.chunk_rows = drows; D.p.chunk_lb = dlb; D.p.chunk_ub = dub; D.p.jac_ptr = djp; D.p.g_row = dg; D.p.b_row = dbr; D.p.sel = dsel;
    D.p.stage = dst; D.p.blk_cnt = dblk; D.p.ticket = ticket; D.p.ftol = 1e-6;
    return D;
}

template <class F> static float timeit(F f, const Data& D) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 6; ++it) {
        CK(cudaMemset(D.p.ticket, 0, 4 * 64));
        cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (it >= 2 && ms < best) best = ms;
    }
    return best;
}
static int sms;
template <int FEAT, int CUT, bool NA, int WARPS, int BPS> static void run(const Data& D, const char* note = "") {
    const int smem = (CUT >= 2) ? WARPS * 16 * PASS * 8 : 0;
    CK(cudaFuncSetAttribute(k_fam<FEAT, CUT, NA, WARPS, BPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem > 1024 ? smem : 1024));
    CK(cudaMemset(D.p.g_row, 0, 8 * D.m));
    const float ms = timeit([&] { k_fam<FEAT, CUT, NA, WARPS, BPS><<<sms * BPS, WARPS * 32, smem>>>(D.p); }, D);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_fam<FEAT, CUT, NA, WARPS, BPS>);
    size_t bad = 0;
    if (FEAT & 1) { std::vector<double> g(D.m); CK(cudaMemcpy(g.data(), D.p.g_row, 8 * D.m, cudaMemcpyDeviceToHost)); for (size_t i = 0; i < D.m; ++i) if (memcmp(&g[i], &D.gref[i], 8)) ++bad; }
    printf("fam feat=%d cut=%d na=%d warps/SM=%2d (%dx%d) regs=%3d spill=%4zu smem=%5d: %6.1f us  %5.0f GB/s  bad_g=%zu %s\n", FEAT, CUT, NA, WARPS * BPS, WARPS, BPS, fa.numRegs,
           (size_t)fa.localSizeBytes, smem, ms * 1e3, D.bytes / ms / 1e6, bad, note);
}

int main(int argc, char** argv) {
    setvbuf(stdout, nullptr, _IONBF, 0);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double v = argc > 1 ? atof(argv[1]) : 0.1;
    Data D = make(100000, v);
    printf("mixed K=4..16, 10^6 rows, %u chunks, v=%.2f, %.1f MB read per round (blobs + rows/bounds + x*)\n", D.p.nchunks, v, D.bytes / 1e6);
    // features one by one (16 warps, plain __ldg)
    run<0, 0, false, 16, 1>(D, "classes only");
    run<1, 0, false, 16, 1>(D, "+ scattered rows");
    run<3, 0, false, 16, 1>(D, "+ bounds, test, sel");
    run<3, 0, true, 16, 1>(D, "same, L1::no_allocate streams");
    run<3, 0, false, 8, 4>(D, "32 warps");
    run<3, 0, false, 12, 2>(D, "24 warps");
    run<3, 0, false, 8, 3>(D, "24 warps (8x3)");
    // cut schemes
    run<3, 1, false, 16, 1>(D, "cut 1: recompute from L1");
    run<3, 1, true, 16, 1>(D, "cut 1, no_allocate streams");
    run<3, 1, false, 12, 2>(D, "cut 1, 24 warps");
    run<3, 1, false, 8, 3>(D, "cut 1, 24 warps (8x3)");
    run<3, 1, false, 8, 4>(D, "cut 1, 32 warps");
    run<3, 2, false, 16, 1>(D, "cut 2: keep exp");
    run<3, 2, false, 12, 2>(D, "cut 2, 24 warps");
    run<3, 3, false, 16, 1>(D, "cut 3: keep exp and x (round 1)");
    run<3, 3, true, 16, 1>(D, "cut 3, no_allocate (round 1)");
    printf("done\n");
    return 0;
}
