// ktn_kernels.cu -- hand-written sm_100a kernels of the ECP separation round.
//
// What the reference does per round (src/model.jl:265-283): precompute! evaluates every
// constraint and the full sparse Jacobian on one CPU thread by interpreting one tape per
// constraint (src/separators.jl:111-116), then tests each NL row (src/separators.jl:120) and
// builds a cut per violated row (src/algorithms.jl:3-18, src/model.jl:200-207, :68-79).
//
// Here: one WARP runs 32 rows of the same shape in lock step (thread-per-constraint inside a
// chunk).  Per chunk:
//   1. one elected lane issues a TMA bulk copy (cp.async.bulk, mbarrier completion) of the
//      chunk's SoA blob (constants, column ids, sort order) into the warp's shared memory;
//   2. x* values are gathered once per unique column into shared-memory scratch;
//   3. the shape's forward program runs (acc machine, operands from shared memory) -> g;
//   4. violation test; if any lane is violated the reverse program runs -> Jacobian row;
//   5. violated lanes build the cut row (constant b, round_coefs, finiteness) and store the
//      coefficients at the row's slot of the static Jacobian CSR layout.
// A second phase (count / scan / scatter) compacts the selected rows, in ascending row order,
// into the CSR the host LP consumes.  All arithmetic is fp64, unfused, in the oracle's order.
#include "ktn_kernels.cuh"
#include "ktn_math.h"
#include "ktn_interp.h"

#define KTN_WARPS_PER_BLOCK 4
#define KTN_CBLOCK 1024   // rows per compaction block

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nKW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra KD_%=;\nbra KW_%=;\nKD_%=:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}

__device__ __forceinline__ uint32_t load_order(const uint8_t* ord, uint32_t order_bytes, size_t e) {
    return order_bytes == 1 ? (uint32_t)ord[e] : order_bytes == 2 ? (uint32_t)((const uint16_t*)ord)[e] : ((const uint32_t*)ord)[e];
}

__device__ __forceinline__ bool row_selected(const KtnRoundParams& p, double g, double lb, double ub, int32_t row) {
    if (p.mode == KTN_MODE_FORCE) return p.force[row] != 0;
    const bool sat = (g >= lb - p.f_tol) && (g <= ub + p.f_tol);   // src/separators.jl:120 (NaN -> not satisfied)
    return !sat;
}

// ---------------------------------------------------------------------------------------------
// regular chunks: shared-memory staged, one warp per chunk, dynamic chunk tickets
// ---------------------------------------------------------------------------------------------
template <bool EVAL_ONLY>
__global__ void __launch_bounds__(KTN_WARPS_PER_BLOCK * 32) ktn_round_kernel(const KtnRoundParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* wbase = smem + (size_t)warp * p.warp_bytes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(wbase);
    unsigned char* blobbuf = wbase + 128;
    double* S = reinterpret_cast<double*>(blobbuf + p.blob_cap);
    if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    uint32_t parity = 0;
    for (;;) {
        uint32_t c = 0;
        if (lane == 0) c = p.chunk_begin + atomicAdd(&p.ticket[0], 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= p.chunk_end) break;
        const KtnChunkDesc cd = p.chunks[c];
        const KtnShapeDesc sd = p.shapes[cd.shape];
        const int32_t row = p.chunk_rows[cd.row_slot + lane];
        if (!EVAL_ONLY && p.mode == KTN_MODE_SEPARATE && !(sd.flags & KTN_SH_NL)) {   // rows outside nlconstr_ixs are never tested
            if (row >= 0) p.sel[row] = 0u;
            continue;
        }
        if (lane == 0) { mbar_expect(bar, cd.blob_bytes); bulk_g2s(blobbuf, p.blob + cd.blob_off, cd.blob_bytes, bar); }
        const double lb = p.chunk_lb[cd.row_slot + lane], ub = p.chunk_ub[cd.row_slot + lane];
        const uint32_t nu = sd.n_uniq;
        const uint32_t sec_col = (8u * sd.n_const * 32u + 15u) & ~15u;
        const uint32_t sec_ord = (sec_col + 4u * nu * 32u + 15u) & ~15u;
        mbar_wait(bar, parity); parity ^= 1u;
        const int32_t* cols = reinterpret_cast<const int32_t*>(blobbuf + sec_col);
        const uint8_t* ord = blobbuf + sec_ord;
        // gather x* once per unique column (precompute! reads xstar through the evaluator)
        {
            uint32_t u = 0;
            for (; u + 4 <= nu; u += 4) {
                const double a0 = __ldg(p.x + cols[(u + 0) * 32 + lane]), a1 = __ldg(p.x + cols[(u + 1) * 32 + lane]);
                const double a2 = __ldg(p.x + cols[(u + 2) * 32 + lane]), a3 = __ldg(p.x + cols[(u + 3) * 32 + lane]);
                S[(u + 0) * 32 + lane] = a0; S[(u + 1) * 32 + lane] = a1; S[(u + 2) * 32 + lane] = a2; S[(u + 3) * 32 + lane] = a3;
            }
            for (; u < nu; ++u) S[u * 32 + lane] = __ldg(p.x + cols[u * 32 + lane]);
        }
        SmemMem m{reinterpret_cast<const double*>(blobbuf), S, lane};
        const KtnIns* prog = p.prog + sd.prog_off;
        const double g = run_program(prog, 0, sd.n_fwd, m, nu, 0xffffffffu);
        if (row >= 0) p.g_row[row] = g;
        if (!EVAL_ONLY) {
            const bool selected = row >= 0 && row_selected(p, g, lb, ub, row);
            if (__any_sync(0xffffffffu, selected)) {
                run_program(prog, sd.n_fwd, sd.n_ins, m, nu, 0xffffffffu);
                if (selected) {
                    // linear_oa_cut (src/algorithms.jl:8-16): b = g; b += -xstar[col]*partial, Jacobian-entry order
                    double b = g, mx = 0.0;
                    for (uint32_t q = 0; q < nu; ++q) {
                        const uint32_t u = load_order(ord, sd.order_bytes, (size_t)q * 32 + lane);
                        const double jv = S[(nu + u) * 32 + lane], xv = S[u * 32 + lane];
                        const double t = (-xv) * jv;
                        b = b + t;
                        mx = q == 0 ? jv : ktn_jlmax(mx, jv);
                    }
                    // round_coefs (src/model.jl:200-207) then _addcut's finiteness test (src/model.jl:69)
                    const int64_t base = p.jac_ptr[row];
                    bool bad = false;
                    for (uint32_t q = 0; q < nu; ++q) {
                        const uint32_t u = load_order(ord, sd.order_bytes, (size_t)q * 32 + lane);
                        double jv = S[(nu + u) * 32 + lane];
                        if (p.do_round && (jv + p.rng < mx)) jv = 0.0;
                        bad = bad || !ktn_isfinite(jv);
                        p.stage_val[base + q] = jv;
                    }
                    p.b_row[row] = b;
                    p.sel[row] = nu | (bad ? KTN_SEL_ERRBIT : 0u);
                } else if (row >= 0) p.sel[row] = 0u;
            } else if (row >= 0) p.sel[row] = 0u;
        }
        __syncwarp();   // every lane is done with the blob before the next bulk copy overwrites it
    }
}

// ---------------------------------------------------------------------------------------------
// BIG chunks (long tapes, the dense epigraph row src/nlpeval.jl:49-63): global scratch arena
// ---------------------------------------------------------------------------------------------
template <bool EVAL_ONLY>
__global__ void __launch_bounds__(128) ktn_big_kernel(const KtnRoundParams p) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t c = p.chunk_begin + gw; c < p.chunk_end; c += nw) {
        const KtnChunkDesc cd = p.chunks[c];
        const KtnShapeDesc sd = p.shapes[cd.shape];
        const uint32_t L = cd.stride, nu = sd.n_uniq;
        const int32_t row = lane < cd.nrows ? p.chunk_rows[cd.row_slot + lane] : -1;
        if (!EVAL_ONLY && p.mode == KTN_MODE_SEPARATE && !(sd.flags & KTN_SH_NL)) { if (row >= 0) p.sel[row] = 0u; continue; }
        const unsigned char* blob = p.blob + cd.blob_off;
        const size_t sec_col = ((size_t)8 * sd.n_const * L + 15) & ~(size_t)15;
        const size_t sec_ord = (sec_col + (size_t)4 * nu * L + 15) & ~(size_t)15;
        const int32_t* cols = reinterpret_cast<const int32_t*>(blob + sec_col);
        const uint8_t* ord = blob + sec_ord;
        double* S = p.big_scratch + cd.scratch_off;
        double g = 0.0; bool selected = false;
        const bool active = lane < cd.nrows;
        const unsigned amask = __ballot_sync(0xffffffffu, active);
        if (active) {
            for (uint32_t u = 0; u < nu; ++u) S[(size_t)u * L + lane] = __ldg(p.x + cols[(size_t)u * L + lane]);
            GlobalMem m{reinterpret_cast<const double*>(blob), S, lane, L};
            const KtnIns* prog = p.prog + sd.prog_off;
            g = run_program(prog, 0, sd.n_fwd, m, nu, amask);
            p.g_row[row] = g;
            if (!EVAL_ONLY) {
                const double lb = p.chunk_lb[cd.row_slot + lane], ub = p.chunk_ub[cd.row_slot + lane];
                selected = row_selected(p, g, lb, ub, row);
                if (__any_sync(amask, selected)) run_program(prog, sd.n_fwd, sd.n_ins, m, nu, amask);
            }
        }
        if (EVAL_ONLY) continue;
        if (!(sd.flags & KTN_SH_DENSE)) {
            if (selected) {
                double b = g, mx = 0.0;
                for (uint32_t q = 0; q < nu; ++q) {
                    const uint32_t u = load_order(ord, sd.order_bytes, (size_t)q * L + lane);
                    const double jv = S[(size_t)(nu + u) * L + lane], xv = S[(size_t)u * L + lane];
                    b = b + (-xv) * jv;
                    mx = q == 0 ? jv : ktn_jlmax(mx, jv);
                }
                const int64_t base = p.jac_ptr[row];
                bool bad = false;
                for (uint32_t q = 0; q < nu; ++q) {
                    const uint32_t u = load_order(ord, sd.order_bytes, (size_t)q * L + lane);
                    double jv = S[(size_t)(nu + u) * L + lane];
                    if (p.do_round && (jv + p.rng < mx)) jv = 0.0;
                    bad = bad || !ktn_isfinite(jv);
                    p.stage_val[base + q] = jv;
                }
                p.b_row[row] = b;
                p.sel[row] = nu | (bad ? KTN_SEL_ERRBIT : 0u);
            } else if (row >= 0) p.sel[row] = 0u;
        } else {
            // dense row: every column 0..num_var-1 is an entry (src/nlpeval.jl:49-54); columns the
            // expression does not touch carry an explicit 0.0.  The whole warp serves one row at a time.
            const unsigned selmask = __ballot_sync(0xffffffffu, selected);
            if (!selected && row >= 0) p.sel[row] = 0u;
            for (uint32_t r = 0; r < cd.nrows; ++r) {
                if (!((selmask >> r) & 1u)) continue;
                const int32_t rrow = __shfl_sync(0xffffffffu, row, r);
                const double rg = __shfl_sync(0xffffffffu, g, r);
                const int64_t base = p.jac_ptr[rrow];
                const int64_t n = p.num_var;
                double* out = p.stage_val + base;
                for (int64_t j = lane; j < n; j += 32) out[j] = 0.0;
                __syncwarp();
                for (uint32_t u = lane; u < nu; u += 32) out[cols[(size_t)u * L + r]] = S[(size_t)(nu + u) * L + r];
                __syncwarp();
                // b = g + sum_j -x_j * J_j in column order, exactly: blocks of 32 columns whose terms are all
                // +-0 leave a non-zero b unchanged and are skipped; any other block is added lane by lane.
                double b = rg;
                double mx = -ktn_inf();   // identity of the NaN-propagating max
                for (int64_t j0 = 0; j0 < n; j0 += 32) {
                    const int64_t j = j0 + lane;
                    const double jv = j < n ? out[j] : 0.0;
                    const double t = j < n ? (-__ldg(p.x + j)) * jv : 0.0;
                    const unsigned nzm = __ballot_sync(0xffffffffu, j < n && !(t == 0.0));
                    if (nzm != 0u || !(b != 0.0)) {
                        const int cnt = (int)((n - j0) < 32 ? (n - j0) : 32);
                        for (int l = 0; l < cnt; ++l) b = b + __shfl_sync(0xffffffffu, t, l);
                    }
                    if (j < n) mx = ktn_jlmax(mx, jv);
                }
                for (int o = 16; o > 0; o >>= 1) mx = ktn_jlmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                bool bad = false;
                for (int64_t j = lane; j < n; j += 32) {
                    double jv = out[j];
                    if (p.do_round && (jv + p.rng < mx)) jv = 0.0;
                    bad = bad || !ktn_isfinite(jv);
                    out[j] = jv;
                }
                bad = __any_sync(0xffffffffu, bad);
                if (lane == 0) { p.b_row[rrow] = b; p.sel[rrow] = (uint32_t)n | (bad ? KTN_SEL_ERRBIT : 0u); }
                __syncwarp();
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// ordered compaction: selected rows -> CSR in ascending row order (the loop order of
// src/model.jl:272).  count -> scan -> scatter over blocks of KTN_CBLOCK rows.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_scan2(uint32_t& a, unsigned long long& b, uint32_t& ta, unsigned long long& tb) {
    // exclusive scan of (a, b) over a 1024-thread block; totals in (ta, tb)
    __shared__ uint32_t wa[32];
    __shared__ unsigned long long wb[32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t ia = a; unsigned long long ib = b;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t na = __shfl_up_sync(0xffffffffu, ia, o);
        const unsigned long long nb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= (uint32_t)o) { ia += na; ib += nb; }
    }
    if (lane == 31) { wa[warp] = ia; wb[warp] = ib; }
    __syncthreads();
    if (warp == 0) {
        uint32_t va = wa[lane]; unsigned long long vb = wb[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t na = __shfl_up_sync(0xffffffffu, va, o);
            const unsigned long long nb = __shfl_up_sync(0xffffffffu, vb, o);
            if (lane >= (uint32_t)o) { va += na; vb += nb; }
        }
        wa[lane] = va; wb[lane] = vb;
    }
    __syncthreads();
    const uint32_t offa = warp ? wa[warp - 1] : 0u;
    const unsigned long long offb = warp ? wb[warp - 1] : 0ull;
    ta = wa[31]; tb = wb[31];
    a = offa + ia - a; b = offb + ib - b;
    __syncthreads();
}

__global__ void __launch_bounds__(KTN_CBLOCK) ktn_count_kernel(const KtnRoundParams p) {
    const int64_t i = (int64_t)blockIdx.x * KTN_CBLOCK + threadIdx.x;
    const uint32_t s = i < p.num_rows ? p.sel[i] : 0u;
    uint32_t a = s ? 1u : 0u; unsigned long long b = s & ~KTN_SEL_ERRBIT;
    if (s & KTN_SEL_ERRBIT) atomicMin(&p.counts[2], (unsigned long long)i + 1ull);
    uint32_t ta; unsigned long long tb;
    block_scan2(a, b, ta, tb);
    if (threadIdx.x == 0) { p.blk_cnt[blockIdx.x] = ta; p.blk_nnz[blockIdx.x] = tb; }
}

__global__ void __launch_bounds__(KTN_CBLOCK) ktn_scan_kernel(const KtnRoundParams p, uint32_t nblocks) {
    __shared__ uint32_t carry_a; __shared__ unsigned long long carry_b;
    if (threadIdx.x == 0) { carry_a = 0; carry_b = 0; }
    __syncthreads();
    for (uint32_t b0 = 0; b0 < nblocks; b0 += KTN_CBLOCK) {
        const uint32_t i = b0 + threadIdx.x;
        uint32_t a = i < nblocks ? p.blk_cnt[i] : 0u; unsigned long long b = i < nblocks ? p.blk_nnz[i] : 0ull;
        uint32_t ta; unsigned long long tb;
        block_scan2(a, b, ta, tb);
        if (i < nblocks) { p.blk_cnt[i] = carry_a + a; p.blk_nnz[i] = carry_b + b; }
        __syncthreads();
        if (threadIdx.x == 0) { carry_a += ta; carry_b += tb; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        p.counts[3] = carry_a; p.counts[4] = carry_b;
        if (p.counts[2] == ~0ull) { p.counts[0] = carry_a; p.counts[1] = carry_b; }   // no error row: totals stand
        p.out_ptr[carry_a] = (int64_t)carry_b;
        p.ticket[0] = 0u; p.ticket[1] = 0u;                                           // re-arm the chunk scheduler
    }
}

__global__ void __launch_bounds__(KTN_CBLOCK) ktn_scatter_kernel(const KtnRoundParams p) {
    __shared__ int32_t big_row[64]; __shared__ uint32_t big_cnt;
    if (threadIdx.x == 0) big_cnt = 0;
    const int64_t i = (int64_t)blockIdx.x * KTN_CBLOCK + threadIdx.x;
    const uint32_t s = i < p.num_rows ? p.sel[i] : 0u;
    const uint32_t nnz = s & ~KTN_SEL_ERRBIT;
    uint32_t a = s ? 1u : 0u; unsigned long long b = nnz;
    uint32_t ta; unsigned long long tb;
    block_scan2(a, b, ta, tb);
    int64_t cidx = -1, o = 0, base = 0;
    if (s) {
        cidx = (int64_t)p.blk_cnt[blockIdx.x] + a; o = (int64_t)(p.blk_nnz[blockIdx.x] + b); base = p.jac_ptr[i];
        const double g = p.g_row[i], bc = p.b_row[i], lb = p.row_lb[i], ub = p.row_ub[i];
        p.out_row[cidx] = (int32_t)i; p.out_ptr[cidx] = o;
        p.out_lo[cidx] = lb - bc; p.out_hi[cidx] = ub - bc;     // src/model.jl:74-75
        p.out_g[cidx] = g;
        const double v1 = lb - g, v2 = g - ub;
        p.out_viol[cidx] = (g == g) ? (v1 > v2 ? v1 : v2) : g;
        if ((unsigned long long)i + 1ull == p.counts[2]) { p.counts[0] = (unsigned long long)cidx; p.counts[1] = (unsigned long long)o; }
    }
    // rows with few entries are copied by their thread; long rows by the whole block
    bool deferred = false;
    if (s && nnz > 64u) { const uint32_t k = atomicAdd(&big_cnt, 1u); if (k < 64u) { big_row[k] = threadIdx.x; deferred = true; } }
    if (s && !deferred) for (uint32_t q = 0; q < nnz; ++q) { p.out_col[o + q] = p.jac_col[base + q]; p.out_val[o + q] = p.stage_val[base + q]; }
    __shared__ int64_t sh_o[KTN_CBLOCK / 16]; __shared__ int64_t sh_base[KTN_CBLOCK / 16]; __shared__ uint32_t sh_n[KTN_CBLOCK / 16];
    __syncthreads();
    const uint32_t nb = big_cnt < 64u ? big_cnt : 64u;
    for (uint32_t k = 0; k < nb; ++k) {
        if ((int32_t)threadIdx.x == big_row[k]) { sh_o[k] = o; sh_base[k] = base; sh_n[k] = nnz; }
    }
    __syncthreads();
    for (uint32_t k = 0; k < nb; ++k) {
        const int64_t oo = sh_o[k], bb = sh_base[k]; const uint32_t nn = sh_n[k];
        for (uint32_t q = threadIdx.x; q < nn; q += KTN_CBLOCK) { p.out_col[oo + q] = p.jac_col[bb + q]; p.out_val[oo + q] = p.stage_val[bb + q]; }
    }
}

__global__ void ktn_reset_kernel(const KtnRoundParams p) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { p.counts[0] = 0; p.counts[1] = 0; p.counts[2] = ~0ull; p.counts[3] = 0; p.counts[4] = 0; p.ticket[0] = 0u; p.ticket[1] = 0u; }
}

int regular_blocks_per_sm = 1;

}  // namespace

cudaError_t ktn_kernels_configure(int max_smem_optin) {
    cudaError_t e = cudaFuncSetAttribute(ktn_round_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(ktn_round_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
}

template <bool EVAL>
static int launch_eval_part(const KtnRoundParams& p0, uint32_t n_regular, uint32_t n_total, int num_sms, int max_smem_optin,
                            cudaStream_t stream, cudaError_t* err) {
    int launches = 0;
    KtnRoundParams p = p0;
    if (n_regular > 0) {
        const size_t smem = (size_t)KTN_WARPS_PER_BLOCK * p.warp_bytes;
        int per_sm = (int)((size_t)(max_smem_optin + 1024) / (smem + 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 16) per_sm = 16;
        uint32_t blocks = (uint32_t)(num_sms * per_sm);
        const uint32_t need = (n_regular + KTN_WARPS_PER_BLOCK - 1) / KTN_WARPS_PER_BLOCK;
        if (blocks > need) blocks = need;
        p.chunk_begin = 0; p.chunk_end = n_regular;
        ktn_round_kernel<EVAL><<<blocks, KTN_WARPS_PER_BLOCK * 32, smem, stream>>>(p);
        ++launches;
    }
    if (n_total > n_regular) {
        p.chunk_begin = n_regular; p.chunk_end = n_total;
        uint32_t blocks = (n_total - n_regular + 3) / 4;
        if (blocks > (uint32_t)num_sms * 8u) blocks = (uint32_t)num_sms * 8u;
        ktn_big_kernel<EVAL><<<blocks, 128, 0, stream>>>(p);
        ++launches;
    }
    *err = cudaGetLastError();
    return launches;
}

int ktn_launch_round(const KtnRoundParams& p, uint32_t n_regular, uint32_t n_total, int num_sms, int max_smem_optin,
                     cudaStream_t stream, cudaError_t* err) {
    int launches = 0;
    ktn_reset_kernel<<<1, 32, 0, stream>>>(p); ++launches;
    launches += launch_eval_part<false>(p, n_regular, n_total, num_sms, max_smem_optin, stream, err);
    if (*err != cudaSuccess) return launches;
    const uint32_t nblocks = (uint32_t)((p.num_rows + KTN_CBLOCK - 1) / KTN_CBLOCK);
    if (nblocks > 0) {
        ktn_count_kernel<<<nblocks, KTN_CBLOCK, 0, stream>>>(p);
        ktn_scan_kernel<<<1, KTN_CBLOCK, 0, stream>>>(p, nblocks);
        ktn_scatter_kernel<<<nblocks, KTN_CBLOCK, 0, stream>>>(p);
        launches += 3;
    }
    *err = cudaGetLastError();
    return launches;
}

int ktn_launch_eval(const KtnRoundParams& p, uint32_t n_regular, uint32_t n_total, int num_sms, int max_smem_optin,
                    cudaStream_t stream, cudaError_t* err) {
    int launches = 0;
    ktn_reset_kernel<<<1, 32, 0, stream>>>(p); ++launches;
    launches += launch_eval_part<true>(p, n_regular, n_total, num_sms, max_smem_optin, stream, err);
    return launches;
}
