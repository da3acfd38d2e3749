// ktn_kernels.cu -- hand-written sm_100a kernels of the ECP separation round.
//
// What the reference does per round (src/model.jl:265-283): precompute! evaluates every
// constraint and the full sparse Jacobian on one CPU thread by interpreting one tape per
// constraint (src/separators.jl:111-116), then tests each NL row (src/separators.jl:120) and
// builds a cut per violated row (src/algorithms.jl:3-18, src/model.jl:200-207, :68-79).
//
// Here a round is two launches (three when a problem mixes family and generic shapes):
//   K1f ktn_family_kernel  shapes of a recognised FAMILY (log-sum-exp, separable quadratic; ktn_family.h):
//        no interpreter.  Phase A: one thread per row streams the row's constants / column ids with coalesced
//        loads straight from the chunk blob (lane stride 32), gathers x* through L1/L2, evaluates g and tests it.
//        Selected rows are pushed on a block-wide shared-memory list; phase B drains the list 256 rows at a
//        time, one thread per SELECTED row (dense lanes whatever the violated fraction): Jacobian row in
//        Jacobian-entry order, cut constant, round_coefs, finiteness, coefficients to the staging CSR.
//   K1 ktn_round_kernel   every other shape: one WARP interprets 32 rows of the same shape in lock step.  Per chunk:
//        1. an elected lane issues a TMA bulk copy (cp.async.bulk + mbarrier) of the chunk's SoA
//           blob (constants, column ids, sort order) into the warp's shared memory; the next
//           chunk's ticket and descriptor are fetched while the current chunk computes;
//        2. x* is gathered once per unique column into shared-memory scratch;
//        3. the shape's forward program (block-shared copy in shared memory) runs -> g;
//        4. violation test; if any lane is violated the reverse program runs -> Jacobian row
//           (accumulators aliased into dead constant slots of the blob where the compiler could);
//        5. violated lanes build the cut row (constant b, round_coefs, finiteness) and store the
//           coefficients at the row's slot of the static Jacobian CSR layout.
//   K2 ktn_compact_kernel ordered stream compaction of the selected rows into the CSR the host LP consumes,
//        ascending row order, coalesced copies.  K1 leaves per-block cut counts (one 64-bit atomic per warp and
//        block of 4096 rows), so every K2 block finds its output offset with one short sum: no look-back chain.
// BIG shapes (long tapes, the dense epigraph row) take ktn_big_kernel with global scratch.
// All arithmetic is fp64, unfused, in the oracle's order.
#include "ktn_kernels.cuh"
#include "ktn_math.h"
#include "ktn_interp.h"
#include "ktn_family.h"

#ifndef KTN_CBLOCK
#define KTN_CBLOCK 256    // threads per compaction block
#endif
#ifndef KTN_CBPS
#define KTN_CBPS 3        // compaction blocks per SM the register budget is set for
#endif

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

// Adds the selected rows of the calling lanes to the per-block cut counts K2 turns into output offsets.
// Called by all lanes of `act` (the lanes that hold a selected row); lanes whose rows fall into the same
// compaction block share one 64-bit atomic.
__device__ __forceinline__ void count_selected(const KtnRoundParams& p, unsigned act, int32_t row, uint32_t nnz) {
    const uint32_t blk = (uint32_t)row >> KTN_CROWS_LOG2;
    const unsigned grp = __match_any_sync(act, blk);
    const uint32_t nz = __reduce_add_sync(grp, nnz);
    if ((threadIdx.x & 31u) == (uint32_t)(__ffs(grp) - 1))
        atomicAdd(p.blk_cnt + (size_t)(p.epoch & 1u) * p.blk_stride + blk, ((unsigned long long)__popc(grp) << KTN_BLK_SHIFT) + nz);
}

// ---------------------------------------------------------------------------------------------
// K1f: family shapes (ktn_family.h).  Persistent: ONE block of 16 warps per SM; one thread per row.
//   * a row is register resident: every constant and column id of the row is requested at once with coalesced loads straight
//     from the chunk blob (lane stride 32), x* is gathered through L1 / L2, g is evaluated and tested.  NOTHING of the row is
//     kept for the cut: a selected row leaves one 32-byte record {g, aux, lb, ub} and the compaction kernel builds its cut
//     (one thread per selected row, lane-dense whatever the violated fraction).  Building the cut here, in lanes that are 90 %
//     idle, costs 22-45 us of the round whichever way its operands are kept or re-read (profiles/microbench/mb5-mb7).
//   * every class (rows of exactly k unique variables) has its own ticket counter, its own contiguous range of equally sized
//     blobs (no descriptor to fetch) and its own fully unrolled code path; the warps of an SM start in the same class (classes
//     are spread over the SMs in proportion to their work) and move on together, so the instruction working set is one class.
//   * short rows run R = 16 / k chunks per warp iteration (R rows per thread): the loads in flight per warp, not the warps,
//     are what hides the two dependent round trips (constants, then the x* gather); measured 2.8 -> 3.3 TB/s for k = 4.
// Measured and dropped (profiles/microbench/mb2.cu, mb4.cu): TMA / cp.async pipelines through shared
// memory and a shared-memory cache of x* (every shared-memory byte is taken from L1, whose lines are the SM's outstanding-miss
// capacity: 200 KB of idle shared memory alone slow the kernel down 2.2x), L2 prefetch, more warps at fewer registers.
// ---------------------------------------------------------------------------------------------
#define KTN_FP_WARPS 16
#ifndef KTN_DUMP
#define KTN_DUMP 0
#endif
#define KTN_FP_SMEM 0

#ifdef KTN_OPT_TIMING
__device__ unsigned long long ktn_dbg_cycles[16];
#endif

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// 256-bit read-only loads (LDG.E.256): one per lane and group of the family blobs
__device__ __forceinline__ void ldg256(const void* p, double& a, double& b, double& c, double& d) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void ldg256(const void* p, int32_t (&v)[8]) {
    asm("ld.global.nc.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
}

struct FamRow {     // streaming row context of ktn_family.h (class 0): the chunk's SoA sections in global memory, lane offset applied
    const double* C; const int32_t* cols; const uint8_t* rk; const double* X; uint32_t nu;
    __device__ __forceinline__ double cst(uint32_t i) const { return __ldg(C + i * 32u); }
    __device__ __forceinline__ int32_t col(uint32_t u) const { return __ldg(cols + u * 32u); }
    __device__ __forceinline__ double xat(int32_t c) const { return __ldg(X + c); }
    __device__ __forceinline__ double x(uint32_t u) const { return xat(col(u)); }
    __device__ __forceinline__ uint32_t rank(uint32_t u) const { return __ldg(rk + u * 32u); }
};
struct FamStreamSink {
    double* out; const int32_t* scol; const double* X;
    __device__ __forceinline__ void put_j(uint32_t q, double v) { out[q] = v; }
    __device__ __forceinline__ double get_j(uint32_t q) const { return out[q]; }
    __device__ __forceinline__ double xsorted(uint32_t q) const { return __ldg(X + __ldg(scol + q)); }
};

// class 0: rows of more than KTN_FAM_REGS unique variables, one chunk per ticket, cut built here (streaming fallbacks)
template <int FAM>
__device__ __forceinline__ void family_stream_class(const KtnRoundParams& p, uint32_t lane, unsigned int* tk) {
    typedef KtnFamily<FAM> F;
    const uint32_t base = p.cls_begin[FAM][0], n = p.cls_begin[FAM][1] - base;
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(tk, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n) return;
        const uint32_t c = base + t, slotid = c * 32u + lane;
        const KtnChunkDesc cd = p.chunks[c];
        const uint32_t nu = (uint32_t)cd.aux;
        const unsigned char* blob = p.blob + cd.blob_off;
        const uint32_t sec_col = 256u * KTN_FAM_NCONST((uint32_t)FAM, nu), sec_rk = sec_col + 128u * nu;      // constants | columns | rank bytes
        const FamRow r{reinterpret_cast<const double*>(blob) + lane, reinterpret_cast<const int32_t*>(blob + sec_col) + lane, blob + sec_rk + lane, p.x, nu};
        double aux;
        const double g = F::forward_stream(r, aux);
        const int32_t row = __ldg(p.chunk_rows + slotid);
        if (row >= 0) p.g_row[row] = g;
        if (p.mode == KTN_MODE_EVAL) continue;
        const bool selected = row >= 0 && row_selected(p, g, __ldg(p.chunk_lb + slotid), __ldg(p.chunk_ub + slotid), row);
        if (row >= 0 && !selected) p.sel[row] = 0u;
        const unsigned selm = __ballot_sync(0xffffffffu, selected);
        if (selected) {
            const int64_t jb = p.jac_ptr[row];
            FamStreamSink s{p.stage_val + jb, p.jac_col + jb, p.x};
            double b;
            const bool bad = ktn_family_cut_stream<FAM>(r, s, g, aux, p.do_round != 0, p.rng, b);
            p.b_row[row] = b;
            p.sel[row] = nu | (bad ? KTN_SEL_ERRBIT : 0u);
            count_selected(p, selm, row, nu);
        }
    }
}

// class N = 1..16: R chunks per warp iteration, the R rows of a thread in registers
template <int FAM, int N, int R>
__device__ __forceinline__ void family_class(const KtnRoundParams& p, uint32_t lane, unsigned int* tk) {
    typedef KtnFamily<FAM> F;
    const uint32_t base = p.cls_begin[FAM][N], n = p.cls_begin[FAM][N + 1] - base;
    const unsigned char* const blob0 = p.blob + p.cls_blob_off[FAM][N];
    uint32_t t0 = 0;
    if (lane == 0) t0 = atomicAdd(tk, (unsigned)R);
    uint32_t cur = __shfl_sync(0xffffffffu, t0, 0);
    while (cur < n) {
        KtnFamRegs<N> v[R]; int32_t col[R][N]; int32_t row[R]; double lb[R], ub[R];
#if KTN_DUMP
        uint32_t jp[R];
#endif
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t ch = cur + r < n ? cur + r : n - 1;      // past the end: the last chunk again, results dropped
            const unsigned char* blob = blob0 + (size_t)ch * KTN_FAM_BLOB_BYTES(N);
            const uint32_t slotid = (base + ch) * 32u + lane;
#pragma unroll
            for (int gq = 0; gq < (N + 1) / 2; ++gq) {       // pair groups: unique variables 2 gq, 2 gq + 1
                double a0, a1, b0, b1;
                ldg256(blob + gq * 1024 + lane * 32u, a0, a1, b0, b1);
                v[r].p0[2 * gq] = a0; v[r].p1[2 * gq] = a1;
                if (2 * gq + 1 < N) { v[r].p0[2 * gq + 1] = b0; v[r].p1[2 * gq + 1] = b1; }
            }
#pragma unroll
            for (int gq = 0; gq < (N + 7) / 8; ++gq) {       // column groups: unique variables 8 gq .. 8 gq + 7
                int32_t c8[8];
                ldg256(blob + KTN_FAM_COL_OFF(N) + gq * 1024 + lane * 32u, c8);
#pragma unroll
                for (int k = 0; k < 8; ++k) if (8 * gq + k < N) col[r][8 * gq + k] = c8[k];
            }
            row[r] = __ldg(p.chunk_rows + slotid);
            if (cur + r >= n) row[r] = -1;
            lb[r] = __ldg(p.chunk_lb + slotid); ub[r] = __ldg(p.chunk_ub + slotid);
#if KTN_DUMP
            jp[r] = __ldg(p.chunk_jp + slotid);
#endif
        }
        // the warp's next ticket is drawn HERE, behind the rows' loads: its round trip overlaps theirs
        uint32_t tn = 0;
        if (lane == 0) tn = atomicAdd(tk, (unsigned)R);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int u = 0; u < N; ++u) v[r].x[u] = __ldg(p.x + col[r][u]);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double aux;
            const double g = F::template forward<N>(v[r], aux);
            if (row[r] >= 0) p.g_row[row[r]] = g;
            if (p.mode == KTN_MODE_EVAL) continue;
            const bool selected = row[r] >= 0 && row_selected(p, g, lb[r], ub[r], row[r]);
            if (row[r] >= 0 && !selected) p.sel[row[r]] = 0u;
            const unsigned selm = __ballot_sync(0xffffffffu, selected);
            if (selected) {
                p.rec[row[r]] = make_double4(g, aux, lb[r], ub[r]);
                p.sel[row[r]] = (uint32_t)N | KTN_SEL_DEFER;
                count_selected(p, selm, row[r], (uint32_t)N);
            }
        }
        cur = __shfl_sync(0xffffffffu, tn, 0);
    }
}

template <int FAM>
__global__ void __launch_bounds__(KTN_FP_WARPS * 32, 1) ktn_family_kernel(const KtnRoundParams p) {
    asm volatile("griddepcontrol.wait;" ::: "memory");      // launched programmatically behind the previous round's cut kernel

    const uint32_t lane = threadIdx.x & 31u;
    unsigned int* tickets = p.ticket + p.ticket_idx;
    const uint32_t my_n = lane < KTN_FAM_NCLS ? p.cls_begin[FAM][lane + 1] - p.cls_begin[FAM][lane] : 0u;     // lane k: chunks of class k
    uint32_t cls = 0;
    {
        uint32_t smid, nsm;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        asm("mov.u32 %0, %%nsmid;" : "=r"(nsm));
        unsigned long long total = 0, acc = 0;
        for (uint32_t k = 0; k < KTN_FAM_NCLS; ++k) total += (unsigned long long)(p.cls_begin[FAM][k + 1] - p.cls_begin[FAM][k]) * ((k ? k : 32u) + 3u);
        const unsigned long long target = (total * (2ull * smid + 1ull)) / (2ull * nsm);
        for (uint32_t k = 0; k < KTN_FAM_NCLS; ++k) {
            acc += (unsigned long long)(p.cls_begin[FAM][k + 1] - p.cls_begin[FAM][k]) * ((k ? k : 32u) + 3u);
            if (acc > target) { cls = k; break; }
        }
    }
    for (;;) {
        // rows per thread: 16 / k, at most 4
        switch (cls) {
#define KTN_CASE(n, r) case n: family_class<FAM, n, r>(p, lane, &tickets[n]); break;
            KTN_CASE(1, 4) KTN_CASE(2, 4) KTN_CASE(3, 4) KTN_CASE(4, 4) KTN_CASE(5, 3) KTN_CASE(6, 2) KTN_CASE(7, 2) KTN_CASE(8, 2)
            KTN_CASE(9, 1) KTN_CASE(10, 1) KTN_CASE(11, 1) KTN_CASE(12, 1) KTN_CASE(13, 1) KTN_CASE(14, 1) KTN_CASE(15, 1) KTN_CASE(16, 1)
#undef KTN_CASE
            default: family_stream_class<FAM>(p, lane, &tickets[0]); break;
        }
        // this class is dry: one look at every class counter picks the next live class
        const bool live = lane < KTN_FAM_NCLS && my_n > 0 && __ldcg(&tickets[lane]) < my_n;
        const unsigned livem = __ballot_sync(0xffffffffu, live);
        if (!livem) return;
        const unsigned ahead = livem & ~((2u << cls) - 1u);        // first live class after cls, cyclically
        cls = (uint32_t)__ffs(ahead ? ahead : livem) - 1u;
    }
}

// ---------------------------------------------------------------------------------------------
// K1: regular chunks, shared-memory staged, one warp per chunk, dynamic chunk tickets
// ---------------------------------------------------------------------------------------------
template <bool EVAL_ONLY>
__global__ void __launch_bounds__(512, 1) ktn_round_kernel(const KtnRoundParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    // block-shared shape descriptors + programs of the regular shapes
    for (uint32_t i = threadIdx.x; i < p.table_bytes / 16u; i += blockDim.x)
        reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(p.table) + i);
    const KtnShapeDesc* sh_shapes = reinterpret_cast<const KtnShapeDesc*>(smem);
    const KtnIns* sh_prog = reinterpret_cast<const KtnIns*>(smem + p.table_prog_off);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* wbase = smem + ((p.table_bytes + 127u) & ~127u) + (size_t)warp * p.warp_bytes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(wbase);
    unsigned char* blobbuf = wbase + 128;
    double* Sl = reinterpret_cast<double*>(blobbuf + p.blob_cap) + lane;    // scratch, lane offset applied
    double* Cl = reinterpret_cast<double*>(blobbuf) + lane;                 // constants, lane offset applied
    if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    uint32_t parity = 0;

    uint32_t c_cur = 0, c_nxt = 0;
    if (lane == 0) { c_cur = atomicAdd(&p.ticket[p.ticket_idx], 1u); c_nxt = atomicAdd(&p.ticket[p.ticket_idx], 1u); }
    c_cur = p.chunk_begin + __shfl_sync(0xffffffffu, c_cur, 0);
    c_nxt = p.chunk_begin + __shfl_sync(0xffffffffu, c_nxt, 0);
    KtnChunkDesc cd, cdn;
    bool have = c_cur < p.chunk_end, cand = false;
    if (have) {
        cd = p.chunks[c_cur];
        cand = EVAL_ONLY || p.mode != KTN_MODE_SEPARATE || (sh_shapes[cd.shape].flags & KTN_SH_NL);
        if (cand && lane == 0) { mbar_expect(bar, cd.blob_bytes); bulk_g2s(blobbuf, p.blob + cd.blob_off, cd.blob_bytes, bar); }
    }
    while (have) {
        // next-next ticket and next descriptor travel while this chunk computes
        uint32_t c_n2 = 0;
        if (lane == 0) c_n2 = atomicAdd(&p.ticket[p.ticket_idx], 1u);
        const bool have_nxt = c_nxt < p.chunk_end;
        if (have_nxt) cdn = p.chunks[c_nxt];
        const int32_t row = p.chunk_rows[cd.row_slot + lane];
        if (!cand) {   // rows outside nlconstr_ixs are never tested (src/model.jl:272)
            if (row >= 0) p.sel[row] = 0u;
        } else {
            const double lb = p.chunk_lb[cd.row_slot + lane], ub = p.chunk_ub[cd.row_slot + lane];
            const KtnShapeDesc& sd = sh_shapes[cd.shape];
            const uint32_t nu = sd.n_uniq;
            const uint32_t sec_col = (8u * sd.n_const * 32u + 15u) & ~15u;
            const uint32_t sec_ord = (sec_col + 4u * nu * 32u + 15u) & ~15u;
            const int32_t* cols = reinterpret_cast<const int32_t*>(blobbuf + sec_col) + lane;
            const uint8_t* ord = blobbuf + sec_ord;
            mbar_wait(bar, parity); parity ^= 1u;
            // gather x* once per unique column (precompute! reads xstar through the evaluator)
            {
                uint32_t u = 0;
                for (; u + 4 <= nu; u += 4) {
                    const double a0 = __ldg(p.x + cols[(u + 0) * 32]), a1 = __ldg(p.x + cols[(u + 1) * 32]);
                    const double a2 = __ldg(p.x + cols[(u + 2) * 32]), a3 = __ldg(p.x + cols[(u + 3) * 32]);
                    Sl[(u + 0) * 32] = a0; Sl[(u + 1) * 32] = a1; Sl[(u + 2) * 32] = a2; Sl[(u + 3) * 32] = a3;
                }
                for (; u < nu; ++u) Sl[u * 32] = __ldg(p.x + cols[u * 32]);
            }
            SmemMem m{Cl, Sl, (sd.j_in_blob ? Cl : Sl) + sd.j_base * 32u, sd.j_stride * 32u, lane};
            const KtnIns* prog = sh_prog + sd.prog_off;
            const double g = run_program(prog, 0, sd.n_fwd, m, 0xffffffffu);
            if (row >= 0) p.g_row[row] = g;
            if (!EVAL_ONLY) {
                const bool selected = row >= 0 && row_selected(p, g, lb, ub, row);
                uint32_t selv = 0u;
                const unsigned selm = __ballot_sync(0xffffffffu, selected);
                if (selm) {
                    run_program(prog, sd.n_fwd, sd.n_ins, m, 0xffffffffu);
                    if (selected) {
                        // linear_oa_cut (src/algorithms.jl:8-16): b = g; b += -xstar[col]*partial in Jacobian-entry order.
                        // max() of round_coefs (src/model.jl:201) is NaN-propagating: fmax ignores NaN, so track it.
                        double b = g, mx = -ktn_inf();
                        bool anynan = false, bad = false;
                        double* out = p.stage_val + p.jac_ptr[row];
                        if (sd.order_bytes == 1) {
                            const uint8_t* o8 = ord + lane;
                            for (uint32_t q = 0; q < nu; ++q) {
                                const uint32_t u = o8[q * 32];
                                const double jv = m.jld(u), xv = Sl[u * 32];
                                b = b + (-xv) * jv;
                                mx = fmax(mx, jv); anynan = anynan || (jv != jv);
                            }
                            if (anynan) mx = ktn_nan();
                            // round_coefs (src/model.jl:202-206) then _addcut's finiteness test (src/model.jl:69)
                            for (uint32_t q = 0; q < nu; ++q) {
                                double jv = m.jld(o8[q * 32]);
                                if (p.do_round && (jv + p.rng < mx)) jv = 0.0;
                                bad = bad || !(fabs(jv) <= 1.7976931348623157e308);
                                out[q] = jv;
                            }
                        } else {
                            for (uint32_t q = 0; q < nu; ++q) {
                                const uint32_t u = load_order(ord, sd.order_bytes, (size_t)q * 32 + lane);
                                const double jv = m.jld(u), xv = Sl[u * 32];
                                b = b + (-xv) * jv;
                                mx = fmax(mx, jv); anynan = anynan || (jv != jv);
                            }
                            if (anynan) mx = ktn_nan();
                            for (uint32_t q = 0; q < nu; ++q) {
                                double jv = m.jld(load_order(ord, sd.order_bytes, (size_t)q * 32 + lane));
                                if (p.do_round && (jv + p.rng < mx)) jv = 0.0;
                                bad = bad || !(fabs(jv) <= 1.7976931348623157e308);
                                out[q] = jv;
                            }
                        }
                        selv = nu | (bad ? KTN_SEL_ERRBIT : 0u);
                        p.b_row[row] = b;
                                    count_selected(p, selm, row, nu);
                    }
                }
                if (row >= 0) p.sel[row] = selv;
            }
            // the reverse sweep wrote Jacobian accumulators into the blob through the generic proxy; the next bulk copy writes the same
            // bytes through the async proxy: order the two (PTX: fence.proxy.async), then make sure every lane is done with the blob
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
        }
        c_cur = c_nxt; cd = cdn; have = have_nxt;
        c_nxt = p.chunk_begin + __shfl_sync(0xffffffffu, c_n2, 0);
        if (have) {
            cand = EVAL_ONLY || p.mode != KTN_MODE_SEPARATE || (sh_shapes[cd.shape].flags & KTN_SH_NL);
            if (cand && lane == 0) { mbar_expect(bar, cd.blob_bytes); bulk_g2s(blobbuf, p.blob + cd.blob_off, cd.blob_bytes, bar); }
        }
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
        double* S = p.big_scratch + cd.aux;
        double* J = S + (size_t)sd.j_base * L;     // BIG shapes keep their accumulators in scratch (j_in_blob == 0)
        const size_t jmul = (size_t)L * sd.j_stride;
        double g = 0.0; bool selected = false;
        const bool active = lane < cd.nrows;
        const unsigned amask = __ballot_sync(0xffffffffu, active);
        if (active) {
            for (uint32_t u = 0; u < nu; ++u) S[(size_t)u * L + lane] = __ldg(p.x + cols[(size_t)u * L + lane]);
            GlobalMem m{reinterpret_cast<const double*>(blob), S, J + lane, jmul, lane, L};
            const KtnIns* prog = p.prog + sd.prog_off;
            g = run_program(prog, 0, sd.n_fwd, m, amask);
            p.g_row[row] = g;
            if (!EVAL_ONLY) {
                const double lb = p.chunk_lb[cd.row_slot + lane], ub = p.chunk_ub[cd.row_slot + lane];
                selected = row_selected(p, g, lb, ub, row);
                if (__any_sync(amask, selected)) run_program(prog, sd.n_fwd, sd.n_ins, m, amask);
            }
        }
        if (EVAL_ONLY) continue;
        if (!(sd.flags & KTN_SH_DENSE)) {
            const unsigned selm = __ballot_sync(0xffffffffu, selected);
            if (selected) {
                double b = g, mx = 0.0;
                for (uint32_t q = 0; q < nu; ++q) {
                    const uint32_t u = load_order(ord, sd.order_bytes, (size_t)q * L + lane);
                    const double jv = J[(size_t)u * jmul + lane], xv = S[(size_t)u * L + lane];
                    b = b + (-xv) * jv;
                    mx = q == 0 ? jv : ktn_jlmax(mx, jv);
                }
                const int64_t base = p.jac_ptr[row];
                bool bad = false;
                for (uint32_t q = 0; q < nu; ++q) {
                    const uint32_t u = load_order(ord, sd.order_bytes, (size_t)q * L + lane);
                    double jv = J[(size_t)u * jmul + lane];
                    if (p.do_round && (jv + p.rng < mx)) jv = 0.0;
                    bad = bad || !ktn_isfinite(jv);
                    p.stage_val[base + q] = jv;
                }
                p.b_row[row] = b;
                p.sel[row] = nu | (bad ? KTN_SEL_ERRBIT : 0u);
                    count_selected(p, selm, row, nu);
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
                for (uint32_t u = lane; u < nu; u += 32) out[cols[(size_t)u * L + r]] = J[(size_t)u * jmul + r];
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
                if (lane == 0) {
                    p.b_row[rrow] = b; p.sel[rrow] = (uint32_t)n | (bad ? KTN_SEL_ERRBIT : 0u);
                    atomicAdd(p.blk_cnt + (size_t)(p.epoch & 1u) * p.blk_stride + ((uint32_t)rrow >> KTN_CROWS_LOG2), (1ull << KTN_BLK_SHIFT) + (unsigned long long)n);
                }
                __syncwarp();
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2: ordered stream compaction + the cuts of the family rows.  Selected rows -> CSR in ascending row order (the loop order of
// src/model.jl:272); block offsets come from the per-block cut counts K1 left in blk_cnt (no look-back chain).
//   A  flags of the block's KTN_CROWS rows, block scan -> compact list of the selected rows with their entry offsets
//   B  one thread per SELECTED row: bounds shift, violation, and -- for a family row (KTN_SEL_DEFER) -- the row's cut
//      (ktn_family_cut_entries: coefficients and columns written straight to their final place, constant b in entry order)
//   C  rows whose cut K1 built (interpreter shapes, long rows, the dense epigraph row): coalesced copy from the staging CSR
//   D  the LAST block to finish settles the first non-finite row (src/model.jl:69-73, :278), writes totals and blob header and
//      re-arms the per-round state
// ---------------------------------------------------------------------------------------------
#define KTN_CWARPS (KTN_CBLOCK / 32)
__device__ __forceinline__ void block_scan2(uint32_t& a, unsigned long long& b, uint32_t& ta, unsigned long long& tb) {
    // exclusive scan of (a, b) over the block; totals in (ta, tb)
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
        uint32_t va = lane < KTN_CWARPS ? wa[lane] : 0u; unsigned long long vb = lane < KTN_CWARPS ? wb[lane] : 0ull;
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

// Large problems: exclusive scan of K1's per-block counts by one block, so that the compaction blocks read their offsets instead
// of each summing all counts (quadratic in the number of blocks).  off[2 j] = cuts, off[2 j + 1] = nnz before block j; totals at j = nblocks.
__global__ void __launch_bounds__(1024) ktn_blkscan_kernel(const KtnRoundParams p, uint32_t nblocks, uint32_t epoch) {
    __shared__ unsigned long long s_c[1024], s_n[1024];
    const unsigned long long* bc = p.blk_cnt + (size_t)(epoch & 1u) * p.blk_stride;
    const uint32_t per = (nblocks + 1023u) / 1024u, j0 = threadIdx.x * per, j1 = j0 + per < nblocks ? j0 + per : nblocks;
    unsigned long long c = 0, n = 0;
    for (uint32_t j = j0; j < j1; ++j) { const unsigned long long v = bc[j]; c += v >> KTN_BLK_SHIFT; n += v & KTN_BLK_NNZ_MASK; }
    s_c[threadIdx.x] = c; s_n[threadIdx.x] = n;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned long long ac = threadIdx.x >= (unsigned)o ? s_c[threadIdx.x - o] : 0ull, an = threadIdx.x >= (unsigned)o ? s_n[threadIdx.x - o] : 0ull;
        __syncthreads();
        s_c[threadIdx.x] += ac; s_n[threadIdx.x] += an;
        __syncthreads();
    }
    c = s_c[threadIdx.x] - c; n = s_n[threadIdx.x] - n;      // exclusive
    for (uint32_t j = j0; j < j1; ++j) { const unsigned long long v = bc[j]; p.blk_off[2 * j] = c; p.blk_off[2 * j + 1] = n; c += v >> KTN_BLK_SHIFT; n += v & KTN_BLK_NNZ_MASK; }
    if (threadIdx.x == 1023) { p.blk_off[2 * (size_t)nblocks] = s_c[1023]; p.blk_off[2 * (size_t)nblocks + 1] = s_n[1023]; }
}

// One block = KTN_CBLOCK threads x KTN_CRPT consecutive rows per thread = KTN_CROWS rows.
#define KTN_CRPT (KTN_CROWS / KTN_CBLOCK)
__global__ void __launch_bounds__(KTN_CBLOCK, KTN_CBPS) ktn_compact_kernel(const KtnRoundParams p, uint32_t nblocks, uint32_t epoch, int scanned) {
    asm volatile("griddepcontrol.wait;" ::: "memory");      // programmatic dependent launch: the kernel may have been placed while K1 was draining

    __shared__ uint32_t s_cnt_base; __shared__ unsigned long long s_nnz_base, s_tot_n, s_tot_nz;
    __shared__ unsigned long long s_red[4][32];
    __shared__ uint32_t s_off[KTN_CROWS + 1];        // exclusive nnz offsets of the block's selected rows (compact list)
    __shared__ uint32_t s_src[KTN_CROWS];            // first entry of each selected row in the static Jacobian CSR (jac_ptr; the library caps nnz(J) at 2^32 - 1)
    __shared__ uint16_t s_rowl[KTN_CROWS];           // block-local row index of each selected row | 0x8000 non-finite (K1) | 0x4000 deferred
    const uint32_t bid = nblocks - 1u - blockIdx.x;      // the rows K1 read last first: what the L2 still holds of their chunk blobs need not come from DRAM
    const int64_t row0 = (int64_t)bid * KTN_CROWS, i0 = row0 + (int64_t)threadIdx.x * KTN_CRPT;
    // every independent load of the block is requested up front: the row flags, then K1's per-block counts
    uint32_t sv[KTN_CRPT];
    if (i0 + KTN_CRPT <= p.num_rows) {
#pragma unroll
        for (int r = 0; r < KTN_CRPT; r += 4) { const uint4 q = *reinterpret_cast<const uint4*>(p.sel + i0 + r); sv[r] = q.x; sv[r + 1] = q.y; sv[r + 2] = q.z; sv[r + 3] = q.w; }
    } else {
#pragma unroll
        for (int r = 0; r < KTN_CRPT; ++r) sv[r] = (i0 + r < p.num_rows) ? p.sel[i0 + r] : 0u;
    }
    // output offset of this block = cuts / nnz of all blocks before it; the totals of ALL blocks fix the blob's layout
    unsigned long long* bc = p.blk_cnt + (size_t)(epoch & 1u) * p.blk_stride;
    if (!scanned) {
        unsigned long long cb = 0, nb = 0, ca = 0, na = 0;
        for (uint32_t j = threadIdx.x; j < nblocks; j += KTN_CBLOCK) {
            const unsigned long long v = bc[j], c = v >> KTN_BLK_SHIFT, n = v & KTN_BLK_NNZ_MASK;
            ca += c; na += n;
            if (j < bid) { cb += c; nb += n; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            cb += __shfl_xor_sync(0xffffffffu, cb, o); nb += __shfl_xor_sync(0xffffffffu, nb, o);
            ca += __shfl_xor_sync(0xffffffffu, ca, o); na += __shfl_xor_sync(0xffffffffu, na, o);
        }
        if ((threadIdx.x & 31u) == 0) { const uint32_t w = threadIdx.x >> 5; s_red[0][w] = cb; s_red[1][w] = nb; s_red[2][w] = ca; s_red[3][w] = na; }
    }
    uint32_t a = 0; unsigned long long b = 0;
#pragma unroll
    for (int r = 0; r < KTN_CRPT; ++r) { a += sv[r] ? 1u : 0u; b += KTN_SEL_NNZ(sv[r]); }
    uint32_t ta; unsigned long long tb;
    block_scan2(a, b, ta, tb);      // contains the barriers that publish s_red
    if (threadIdx.x < 32) {
        unsigned long long cb, nb, ca, na;
        if (scanned) { cb = p.blk_off[2 * (size_t)bid]; nb = p.blk_off[2 * (size_t)bid + 1]; ca = p.blk_off[2 * (size_t)nblocks]; na = p.blk_off[2 * (size_t)nblocks + 1]; }
        else {
            const bool live = threadIdx.x < KTN_CWARPS;
            cb = live ? s_red[0][threadIdx.x] : 0ull; nb = live ? s_red[1][threadIdx.x] : 0ull;
            ca = live ? s_red[2][threadIdx.x] : 0ull; na = live ? s_red[3][threadIdx.x] : 0ull;
            for (int o = 16; o > 0; o >>= 1) {
                cb += __shfl_xor_sync(0xffffffffu, cb, o); nb += __shfl_xor_sync(0xffffffffu, nb, o);
                ca += __shfl_xor_sync(0xffffffffu, ca, o); na += __shfl_xor_sync(0xffffffffu, na, o);
            }
        }
        if (threadIdx.x == 0) {
            s_cnt_base = (uint32_t)cb; s_nnz_base = nb; s_tot_n = ca; s_tot_nz = na;
            p.blk_cnt[(size_t)((epoch & 1u) ^ 1u) * p.blk_stride + bid] = 0ull;     // re-arm the slot the NEXT round's K1 adds into
            if (bid == 0) { p.counts[4] = ca; p.counts[5] = na; }                   // the cut kernel's work list ends here
        }
    }
    __syncthreads();
    const uint32_t cbase = s_cnt_base; const unsigned long long nbase = s_nnz_base;
    const KtnPackLayout L = ktn_pack_layout(s_tot_n, s_tot_nz);
    int64_t* const out_row = reinterpret_cast<int64_t*>(p.out_blob + L.row_id); int64_t* const out_ptr = reinterpret_cast<int64_t*>(p.out_blob + L.row_ptr);
    double* const out_lo = reinterpret_cast<double*>(p.out_blob + L.lo); double* const out_hi = reinterpret_cast<double*>(p.out_blob + L.hi);
    double* const out_g = reinterpret_cast<double*>(p.out_blob + L.g); double* const out_viol = reinterpret_cast<double*>(p.out_blob + L.viol);
    double* const out_b = reinterpret_cast<double*>(p.out_blob + L.b);
    int32_t* const out_col = reinterpret_cast<int32_t*>(p.out_blob + L.col); double* const out_val = reinterpret_cast<double*>(p.out_blob + L.val);
    if (bid == nblocks - 1 && threadIdx.x == 0) { out_ptr[s_tot_n] = (int64_t)s_tot_nz; p.cut_off[s_tot_n] = s_tot_nz; }
    const bool full = !p.lean_out;
    // compact list of the block's selected rows: local row index (+ flags) and nnz offset
    bool copy = false;
#pragma unroll
    for (int r = 0; r < KTN_CRPT; ++r) {
        const uint32_t s = sv[r];
        if (!s) continue;
        s_off[a] = (uint32_t)b;
        s_rowl[a] = (uint16_t)((threadIdx.x * KTN_CRPT + r) | ((s & KTN_SEL_ERRBIT) ? 0x8000u : 0u) | ((s & KTN_SEL_DEFER) ? 0x4000u : 0u));
        a += 1u; b += KTN_SEL_NNZ(s);
        copy = copy || !(s & KTN_SEL_DEFER);
    }
    if (threadIdx.x == 0) s_off[ta] = (uint32_t)tb;
    const int any_copy = __syncthreads_or(copy ? 1 : 0);
    // one thread per SELECTED row: where it goes.  Rows whose cut K1 built get their scalars here; a family row gets an entry of
    // the cut kernel's work list (slot, nnz) with its record {g, aux, lb, ub} parked at the cut's index (so that the cut kernel
    // starts from coalesced loads), and the sectors of the chunk blob that hold the row are requested into L2.
    unsigned long long* const wl = p.worklist;
    for (uint32_t k = threadIdx.x; k < ta; k += KTN_CBLOCK) {
        const uint32_t rl = s_rowl[k];
        const int64_t i = row0 + (rl & 0x3fffu);
        const int64_t cidx = (int64_t)cbase + k, o = (int64_t)(nbase + s_off[k]);
        out_row[cidx] = i + p.row_offset; out_ptr[cidx] = o; p.cut_off[cidx] = (unsigned long long)o;
        if (!(rl & 0x4000u)) {
            s_src[k] = (uint32_t)p.jac_ptr[i];
            const double g = p.g_row[i], bcst = p.b_row[i], lb = p.row_lb[i], ub = p.row_ub[i];
            out_lo[cidx] = lb - bcst; out_hi[cidx] = ub - bcst;     // src/model.jl:74-75
            if (full) {
                out_g[cidx] = g; out_b[cidx] = bcst;
                const double v1 = lb - g, v2 = g - ub;
                out_viol[cidx] = (g == g) ? (v1 > v2 ? v1 : v2) : g;
            }
            wl[cidx] = 0ull;
            if (rl & 0x8000u) atomicMin(&p.counts[2 + (epoch & 1u)], (unsigned long long)cidx);      // first non-finite cut of the round
        } else {
            const double4 rc = p.rec[i];
            const uint32_t slot = (uint32_t)__ldg(p.row_slot + i), c = slot >> 5, ln = slot & 31u, nu = s_off[k + 1] - s_off[k];
            int fam = KTN_FAM_LSE;
            while (fam + 1 < KTN_FAM__COUNT && c >= p.fam_begin[fam + 1]) ++fam;      // chunks are sorted by family
            const unsigned char* blob = p.blob + p.cls_blob_off[fam][nu] + (size_t)(c - p.cls_begin[fam][nu]) * KTN_FAM_BLOB_BYTES(nu);
            for (uint32_t gq = 0; gq < KTN_FAM_PGROUPS(nu) + KTN_FAM_CGROUPS(nu); ++gq) prefetch_l2(blob + gq * 1024u + ln * 32u);
            prefetch_l2(blob + KTN_FAM_ORD_OFF(nu) + ln * 8u);
            p.park[cidx] = rc;
            wl[cidx] = ((unsigned long long)(nu | ((uint32_t)fam << 8) | 0x10000u) << 32) | slot;
        }
    }
    __syncthreads();
    // expand the rows whose cut K1 built: one thread per output entry, coalesced writes of columns and coefficients from the
    // static / staging CSR.  Every warp owns a contiguous range of the block's entries: one binary search (shared memory) finds
    // the row of the range's first entry, after that each lane walks forward through the row offsets (a few steps per 32
    // entries).  Four steps are unrolled: all loads are in flight before the first store.  (Family rows: the cut kernel.)
    if (!any_copy) return;
    const uint32_t nent = (uint32_t)tb, lane = threadIdx.x & 31u;
    const uint32_t per = ((nent + KTN_CBLOCK - 1) / KTN_CBLOCK) * 32u;         // entries per warp, a multiple of 32
    const uint32_t wbeg = (threadIdx.x >> 5) * per, wend = wbeg + per < nent ? wbeg + per : nent;
    if (wbeg < wend) {
        uint32_t lo = 0, hi = ta;   // largest k with s_off[k] <= wbeg
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (s_off[mid] <= wbeg) lo = mid; else hi = mid; }
        for (uint32_t e0 = wbeg + lane; e0 - lane < wend; e0 += 128u) {
            uint32_t src[4]; int32_t cv[4]; double vv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t e = e0 + 32u * (uint32_t)k;
                src[k] = 0xffffffffu;
                if (e < wend) { while (s_off[lo + 1] <= e) ++lo; if (!(s_rowl[lo] & 0x4000u)) src[k] = s_src[lo] + (e - s_off[lo]); }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) { cv[k] = src[k] != 0xffffffffu ? __ldg(p.jac_col + src[k]) : 0; vv[k] = src[k] != 0xffffffffu ? p.stage_val[src[k]] : 0.0; }
#pragma unroll
            for (int k = 0; k < 4; ++k) if (src[k] != 0xffffffffu) { const unsigned long long e = nbase + e0 + 32u * (uint32_t)k; out_col[e] = cv[k]; out_val[e] = vv[k]; }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K3: the cuts of the family rows (KTN_SEL_DEFER).  One thread per cut of the compacted list, KTN_XBLOCK consecutive cuts per
// block: consecutive cuts own consecutive slices of the round's CSR, so the block stages its coefficients in shared memory and
// writes them out coalesced.  Per row: record {g, aux, lb, ub}, slot -> the row's groups in the chunk blob (256-bit loads),
// x* gathers, ktn_family_cut_terms.  Every load of a row is independent of the row's arithmetic: eight terms are in flight at once.
// The LAST block to finish settles the first non-finite row (src/model.jl:69-73, :278), writes totals and blob header and
// re-arms the per-round state.
// ---------------------------------------------------------------------------------------------
#ifndef KTN_XBLOCK
#define KTN_XBLOCK 128
#endif
#ifndef KTN_XBPS
#define KTN_XBPS 3       // measured: 3 blocks of 128 threads per SM (166 registers) 30.8 us, 4 (128 registers) 37.4, 2: 34.8 (profiles/r02_ab_ab11.log)
#endif
struct CutRow {     // row context of ktn_family_cut_terms: the row's groups in the chunk blob (ktn_program.h)
    const unsigned char* blob; const double* X; uint32_t nu, lane;
    __device__ __forceinline__ void pairs2(uint32_t g, double& a0, double& a1, double& b0, double& b1) const { ldg256(blob + g * 1024u + lane * 32u, a0, a1, b0, b1); }
    __device__ __forceinline__ void cols8(uint32_t g, int32_t (&c)[8]) const { ldg256(blob + KTN_FAM_COL_OFF(nu) + g * 1024u + lane * 32u, c); }
    __device__ __forceinline__ double xat(int32_t c) const { return __ldg(X + c); }
};
struct CutSink {    // coefficients and columns: the row's slice of the block's staging (or of the round's CSR); products: a column of the block's scratch
    double* val; int32_t* col; double* t;
    __device__ __forceinline__ void put(uint32_t q, double v, int32_t c) { val[q] = v; col[q] = c; }
    __device__ __forceinline__ double get(uint32_t q) const { return val[q]; }
    __device__ __forceinline__ void set(uint32_t q, double v) { val[q] = v; }
    __device__ __forceinline__ void put_t(uint32_t q, double v) { t[q * KTN_XBLOCK] = v; }
    __device__ __forceinline__ double get_t(uint32_t q) const { return t[q * KTN_XBLOCK]; }
};

__global__ void __launch_bounds__(KTN_XBLOCK, KTN_XBPS) ktn_cut_kernel(const KtnRoundParams p, uint32_t epoch) {
    asm volatile("griddepcontrol.wait;" ::: "memory");

#ifndef KTN_X_NOSTAGE
    __shared__ double s_val[KTN_XBLOCK * KTN_FAM_REGS];      // coefficients of the block's cuts, in CSR order
#else
    double* const s_val = nullptr;
#endif
    __shared__ double s_t[KTN_FAM_REGS * KTN_XBLOCK];        // products -x* J: [entry][thread]
#ifndef KTN_X_NOSTAGE
    __shared__ int32_t s_col[KTN_XBLOCK * KTN_FAM_REGS];     // columns of the block's cuts, in CSR order
#else
    int32_t* const s_col = nullptr;
#endif
    __shared__ unsigned long long s_e0, s_e1; __shared__ uint32_t s_last;
    const unsigned long long ca = __ldcg(&p.counts[4]), na = __ldcg(&p.counts[5]);
    const KtnPackLayout L = ktn_pack_layout(ca, na);
    const int64_t* const out_row = reinterpret_cast<const int64_t*>(p.out_blob + L.row_id);
    double* const out_lo = reinterpret_cast<double*>(p.out_blob + L.lo); double* const out_hi = reinterpret_cast<double*>(p.out_blob + L.hi);
    double* const out_g = reinterpret_cast<double*>(p.out_blob + L.g); double* const out_viol = reinterpret_cast<double*>(p.out_blob + L.viol);
    double* const out_b = reinterpret_cast<double*>(p.out_blob + L.b); double* const out_val = reinterpret_cast<double*>(p.out_blob + L.val);
    int32_t* const out_col = reinterpret_cast<int32_t*>(p.out_blob + L.col);
    const bool full = !p.lean_out;
    for (unsigned long long c0 = (unsigned long long)blockIdx.x * KTN_XBLOCK; c0 < ca; c0 += (unsigned long long)gridDim.x * KTN_XBLOCK) {
        const unsigned long long cidx = c0 + threadIdx.x;
        const bool active = cidx < ca;
        int64_t o = 0; unsigned long long w = 0ull; double g = 0.0, aux = 0.0, lb = 0.0, ub = 0.0;
        if (active) {      // one round trip: the work-list entry, the cut's place in the CSR and the row's record, all parked by the compaction kernel
            w = __ldcg(p.worklist + cidx); o = (int64_t)__ldcg(p.cut_off + cidx);
            const double2* const pk = reinterpret_cast<const double2*>(p.park + cidx);
            const double2 r0 = __ldcg(pk), r1 = __ldcg(pk + 1);
            g = r0.x; aux = r0.y; lb = r1.x; ub = r1.y;
        }
        if (threadIdx.x == 0) { s_e0 = (unsigned long long)o; const unsigned long long cend = c0 + KTN_XBLOCK < ca ? c0 + KTN_XBLOCK : ca; s_e1 = __ldcg(p.cut_off + cend); }
        const bool mine = active && w != 0ull;
        // the block's slice of the CSR is staged when every cut of the block is a family cut (else: straight to the CSR)
#ifndef KTN_X_NOSTAGE
        const int staged = __syncthreads_and((mine || !active) ? 1 : 0);
#else
        const int staged = 0; __syncthreads();
#endif
        const unsigned long long e0 = s_e0, e1 = s_e1;
        if (mine) {
            const uint32_t slot = (uint32_t)w, c = slot >> 5, ln = slot & 31u, nu = (uint32_t)(w >> 32) & 0xffu;
            const int fam = (int)((w >> 40) & 0xffu);
            const unsigned char* blob = p.blob + p.cls_blob_off[fam][nu] + (size_t)(c - p.cls_begin[fam][nu]) * KTN_FAM_BLOB_BYTES(nu);
            const uint64_t rw = __ldg(reinterpret_cast<const unsigned long long*>(blob + KTN_FAM_ORD_OFF(nu)) + ln);
            const CutRow r{blob, p.x, nu, ln};
            CutSink s{staged ? s_val + ((unsigned long long)o - e0) : out_val + o, staged ? s_col + ((unsigned long long)o - e0) : out_col + o, s_t + threadIdx.x};
            double bcst; bool bad;
            if (fam == KTN_FAM_LSE) bad = ktn_family_cut_terms<KTN_FAM_LSE>(r, nu, rw, s, g, aux, p.do_round != 0, p.rng, bcst);
            else if (fam == KTN_FAM_QUAD) bad = ktn_family_cut_terms<KTN_FAM_QUAD>(r, nu, rw, s, g, aux, p.do_round != 0, p.rng, bcst);
            else bad = ktn_family_cut_terms<KTN_FAM_SOC>(r, nu, rw, s, g, aux, p.do_round != 0, p.rng, bcst);
            out_lo[cidx] = lb - bcst; out_hi[cidx] = ub - bcst;     // src/model.jl:74-75
            if (full) {
                out_g[cidx] = g; out_b[cidx] = bcst;
                const double v1 = lb - g, v2 = g - ub;
                out_viol[cidx] = (g == g) ? (v1 > v2 ? v1 : v2) : g;
            }
            if (bad) atomicMin(&p.counts[2 + (epoch & 1u)], cidx);      // first non-finite cut of the round
        }
        __syncthreads();
        if (staged) for (unsigned long long e = e0 + threadIdx.x; e < e1; e += KTN_XBLOCK) { out_val[e] = s_val[e - e0]; out_col[e] = s_col[e - e0]; }
        __syncthreads();
    }
    // the last block to finish settles the round
    if (threadIdx.x == 0) { __threadfence(); s_last = atomicAdd(&p.counts[7], 1ull) == (unsigned long long)gridDim.x - 1ull ? 1u : 0u; }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned long long* const hdr = reinterpret_cast<unsigned long long*>(p.out_blob);
        const unsigned long long errc = __ldcg(&p.counts[2 + (epoch & 1u)]);      // the reference stops at the first non-finite cut (src/model.jl:278): cuts before it stand
        unsigned long long n = ca, nz = na, err = ~0ull;
        if (errc != ~0ull) { n = errc; nz = __ldcg(p.cut_off + errc); err = (unsigned long long)(__ldcg(out_row + errc) - p.row_offset) + 1ull; }
        p.counts[0] = n; p.counts[1] = nz; p.counts[6] = err;
        p.counts[2 + ((epoch & 1u) ^ 1u)] = ~0ull;                    // re-arm the slot the NEXT round uses
        p.counts[7] = 0ull;
        hdr[0] = n; hdr[1] = nz; hdr[2] = err; hdr[3] = L.total; hdr[4] = (unsigned long long)p.row_offset; hdr[5] = ca; hdr[6] = na; hdr[7] = 0ull;
    }
    for (uint32_t i = threadIdx.x; i < KTN_TICKETS; i += KTN_XBLOCK) p.ticket[i] = 0u;     // K1 is over: re-arm its work tickets
}

// ---------------------------------------------------------------------------------------------
// Top-k selection (ktn_options.topk > 0; a build extension, the reference emits every violated row -- src/model.jl:272-283).
// Between K1 and K2: of the violated rows keep the k ranked first by (NaN first, violation max(lb - g, g - ub) descending,
// row index ascending); the survivors are emitted in ascending row order by the unchanged K2.
//   T1 keys      one 64-bit order-preserving key per row (0 = not violated)
//   T2 select    8 passes of an 8-bit radix select, most significant digit first: block-local shared-memory histograms of the
//                keys that still match the prefix, merged with global atomics; the LAST block to finish picks the digit that
//                holds the k-th key (no block ever waits for another)  ->  threshold key T and how many keys == T to keep
//   T3 ties      per 4096-row block: number of keys == T (ties at the threshold are taken in row order)
//   T4 demote    rows below the threshold (or beyond the tie quota) are deselected; per-block cut counts and the first
//                non-finite row are rewritten for the survivors; K2 then runs as always
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long topk_key(double g, double lb, double ub) {
    if (g != g) return ~0ull;                                       // NaN ranks first
    const double v1 = lb - g, v2 = g - ub, v = v1 > v2 ? v1 : v2;
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    u = (u >> 63) ? ~u : (u | 0x8000000000000000ull);               // order-preserving map of fp64 onto unsigned integers
    return u == ~0ull ? u - 1ull : (u == 0ull ? 1ull : u);          // keep clear of the two reserved values
}

__global__ void __launch_bounds__(256) ktn_topk_key_kernel(const KtnRoundParams p, unsigned long long* key, KtnTopkState* st, unsigned long long k) {
    if (blockIdx.x == 0 && threadIdx.x < 256) {
        st->hist[threadIdx.x] = 0u;
        if (threadIdx.x == 0) { st->prefix = 0ull; st->mask = 0ull; st->remaining = k; st->done = 0u; st->all = 0u; st->eq_total = 0ull; }
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.num_rows; i += (int64_t)gridDim.x * blockDim.x)
        key[i] = p.sel[i] ? topk_key(p.g_row[i], p.row_lb[i], p.row_ub[i]) : 0ull;
}

__global__ void __launch_bounds__(256) ktn_topk_select_kernel(const KtnRoundParams p, const unsigned long long* key, KtnTopkState* st, int pass) {
    __shared__ unsigned int sh[256];
    __shared__ unsigned int s_last;
    sh[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned long long prefix = st->prefix, mask = st->mask;      // written by the last block of the previous pass (a previous launch)
    const int shift = 56 - 8 * pass;
    if (!st->all) {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.num_rows; i += (int64_t)gridDim.x * blockDim.x) {
            const unsigned long long kk = key[i];
            if (kk != 0ull && (kk & mask) == prefix) atomicAdd(&sh[(unsigned)(kk >> shift) & 255u], 1u);
        }
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], sh[threadIdx.x]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&st->done, 1u) == gridDim.x - 1u) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last block: digit of the k-th key among the keys matching the prefix.  One bin per thread, suffix sums over the bins
    // (so that "above" = keys in higher bins), and exactly one thread finds above < k' <= above + hist[d].
    __shared__ unsigned long long suf[256];
    const unsigned int mybin = *reinterpret_cast<volatile unsigned int*>(&st->hist[threadIdx.x]);
    suf[threadIdx.x] = mybin;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        const unsigned long long add = threadIdx.x + o < 256 ? suf[threadIdx.x + o] : 0ull;
        __syncthreads();
        suf[threadIdx.x] += add;
        __syncthreads();
    }
    const unsigned long long rem = st->remaining, total = suf[0], above = suf[threadIdx.x] - mybin;
    const bool was_all = st->all != 0u;
    __syncthreads();
    if (pass == 0 && total <= rem) { if (threadIdx.x == 0) st->all = 1u; }              // fewer violated rows than k: everything survives
    else if (!was_all && mybin > 0u && above < rem && rem <= above + mybin) {
        st->prefix = prefix | ((unsigned long long)threadIdx.x << shift); st->mask = mask | (255ull << shift);
        st->remaining = rem - above;                                    // keys in higher bins all survive
        st->eq_total = mybin;                                           // after the last pass: number of keys equal to the threshold
    }
    st->hist[threadIdx.x] = 0u;
    if (threadIdx.x == 0) {
        st->done = 0u;
    }
}

__global__ void __launch_bounds__(KTN_CBLOCK) ktn_topk_ties_kernel(const KtnRoundParams p, const unsigned long long* key, const KtnTopkState* st, unsigned int* eqcnt) {
    __shared__ unsigned int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0u;
    __syncthreads();
    const unsigned long long T = st->prefix;
    const int64_t i0 = (int64_t)blockIdx.x * KTN_CROWS + (int64_t)threadIdx.x * KTN_CRPT;
    unsigned int c = 0;
    if (!st->all) for (int r = 0; r < KTN_CRPT; ++r) if (i0 + r < p.num_rows && key[i0 + r] == T) ++c;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31u) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) eqcnt[blockIdx.x] = s_cnt;
}

__global__ void __launch_bounds__(KTN_CBLOCK) ktn_topk_demote_kernel(const KtnRoundParams p, const unsigned long long* key, const KtnTopkState* st, const unsigned int* eqcnt) {
    __shared__ unsigned int s_w[32];
    __shared__ unsigned long long s_eq_before, s_cnt, s_nnz;
    const uint32_t bid = blockIdx.x, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_cnt = 0ull; s_nnz = 0ull; }
    const bool all = st->all != 0u;
    const unsigned long long T = st->prefix, quota = st->remaining;     // keys == T that survive, in row order
    // keys == T in the blocks before this one
    unsigned long long before = 0;
    if (!all) for (uint32_t j = threadIdx.x; j < bid; j += KTN_CBLOCK) before += eqcnt[j];
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    if (threadIdx.x == 0) s_eq_before = 0ull;
    __syncthreads();
    if (lane == 0 && before) atomicAdd(&s_eq_before, before);
    // ordered rank of this thread's keys == T inside the block
    const int64_t i0 = (int64_t)bid * KTN_CROWS + (int64_t)threadIdx.x * KTN_CRPT;
    unsigned long long kk[KTN_CRPT]; unsigned int mine = 0;
    for (int r = 0; r < KTN_CRPT; ++r) { kk[r] = i0 + r < p.num_rows ? key[i0 + r] : 0ull; if (!all && kk[r] == T) ++mine; }
    unsigned int incl = mine;
    for (int o = 1; o < 32; o <<= 1) { const unsigned int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += n; }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) { unsigned int v = lane < KTN_CWARPS ? s_w[lane] : 0u; for (int o = 1; o < 32; o <<= 1) { const unsigned int n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= (uint32_t)o) v += n; } s_w[lane] = v; }
    __syncthreads();
    unsigned long long rank = s_eq_before + (warp ? s_w[warp - 1] : 0u) + (incl - mine);
    unsigned long long cnt = 0, nnz = 0;
    for (int r = 0; r < KTN_CRPT; ++r) {
        if (kk[r] == 0ull) continue;
        const int64_t i = i0 + r;
        bool keep = all || kk[r] > T;
        if (!all && kk[r] == T) { keep = rank < quota; ++rank; }
        if (!keep) { p.sel[i] = 0u; continue; }
        const uint32_t s = p.sel[i];
        ++cnt; nnz += KTN_SEL_NNZ(s);
    }
    for (int o = 16; o > 0; o >>= 1) { cnt += __shfl_xor_sync(0xffffffffu, cnt, o); nnz += __shfl_xor_sync(0xffffffffu, nnz, o); }
    if (lane == 0 && cnt) { atomicAdd(&s_cnt, cnt); atomicAdd(&s_nnz, nnz); }
    __syncthreads();
    if (threadIdx.x == 0) p.blk_cnt[(size_t)(p.epoch & 1u) * p.blk_stride + bid] = (s_cnt << KTN_BLK_SHIFT) | s_nnz;     // replaces K1's count of ALL violated rows
}

}  // namespace

cudaError_t ktn_kernels_configure(int max_smem_optin) {
    cudaError_t e = cudaFuncSetAttribute(ktn_round_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(ktn_round_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
}

// warps per block x blocks per SM that maximise resident warps for the given per-warp shared-memory need
void ktn_plan_occupancy(uint32_t table_bytes, uint32_t warp_bytes, int max_smem_optin, int* wpb_out, int* bps_out) {
    const size_t sm_total = (size_t)max_smem_optin + 1024;   // per-SM carve-out incl. the 1 KB per-block reserve
    int best_w = 1, best_b = 1, best = 0;
    const size_t tb = (table_bytes + 127u) & ~127u;
    static int regs = 0;
    if (!regs) { cudaFuncAttributes fa; regs = (cudaFuncGetAttributes(&fa, ktn_round_kernel<false>) == cudaSuccess && fa.numRegs > 0) ? fa.numRegs : 128; }
    const int regs_alloc = (regs + 7) / 8 * 8;
    int max_warps = 65536 / (regs_alloc * 32);
    if (max_warps > 48) max_warps = 48;
    for (int w = 1; w <= 16; ++w) {
        const size_t blk = tb + (size_t)w * warp_bytes;
        if (blk > (size_t)max_smem_optin) break;
        int b = (int)(sm_total / (blk + 1024));
        if (b > 32) b = 32;
        while (b * w > max_warps) --b;              // register file
        if (b < 1) continue;
        if (b * w > best || (b * w == best && w > best_w)) { best = b * w; best_w = w; best_b = b; }
    }
    *wpb_out = best_w; *bps_out = best_b;
}

#ifdef KTN_OPT_TIMING
extern "C" int ktn_debug_cycles(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    if (out16 && cudaMemcpyFromSymbol(out16, ktn_dbg_cycles, sizeof ktn_dbg_cycles) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[16] = {0}; if (cudaMemcpyToSymbol(ktn_dbg_cycles, z, sizeof z) != cudaSuccess) return -1; }
    return 0;
}
#endif

template <int FAM>
static void launch_family(KtnRoundParams p, const KtnLaunchPlan& plan, uint32_t ticket_idx, int num_sms, cudaStream_t stream) {
    const uint32_t begin = plan.fam_begin[FAM], end = plan.fam_begin[FAM + 1];
    p.chunk_begin = begin; p.chunk_end = end; p.ticket_idx = ticket_idx;
    uint32_t blocks = (uint32_t)num_sms;      // persistent: one block per SM
    const uint32_t need = (end - begin + KTN_FP_WARPS - 1) / KTN_FP_WARPS;
    if (blocks > need) blocks = need;
    static const bool pdl = !(getenv("KTN_PDL") && atoi(getenv("KTN_PDL")) == 0);
    if (pdl) {
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        cudaLaunchConfig_t cfg = {}; cfg.stream = stream; cfg.attrs = at; cfg.numAttrs = 1;
        cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(KTN_FP_WARPS * 32); cfg.dynamicSmemBytes = KTN_FP_SMEM;
        cudaLaunchKernelEx(&cfg, ktn_family_kernel<FAM>, p);
    } else ktn_family_kernel<FAM><<<blocks, KTN_FP_WARPS * 32, KTN_FP_SMEM, stream>>>(p);
}

template <bool EVAL>
static int launch_eval_part(const KtnRoundParams& p0, const KtnLaunchPlan& plan, int num_sms, int max_smem_optin,
                            cudaStream_t stream, cudaError_t* err) {
    int launches = 0;
    KtnRoundParams p = p0;
    if (EVAL) p.mode = KTN_MODE_EVAL;
    // interpreted shapes: the rows of nlconstr_ixs first.  A separation round never looks at the others (src/model.jl:272): their
    // chunks are skipped unless an unconditional round (ktn_gencut_rows) selected some of them last time and their flags must be cleared
    const uint32_t g_begin = plan.fam_begin[KTN_FAM_GENERIC];
    const uint32_t g_end = (!EVAL && p.mode == KTN_MODE_SEPARATE && !p.clear_unselected) ? plan.cls_begin[KTN_FAM_GENERIC][1] : plan.fam_begin[KTN_FAM_GENERIC + 1];
    if (g_end > g_begin) {
        int wpb, bps; ktn_plan_occupancy(p.table_bytes, p.warp_bytes, max_smem_optin, &wpb, &bps);
        const size_t smem = ((p.table_bytes + 127u) & ~127u) + (size_t)wpb * p.warp_bytes;
        uint32_t blocks = (uint32_t)(num_sms * bps);
        const uint32_t need = (g_end - g_begin + wpb - 1) / wpb;
        if (blocks > need) blocks = need;
        p.chunk_begin = g_begin; p.chunk_end = g_end; p.ticket_idx = 0;
        ktn_round_kernel<EVAL><<<blocks, wpb * 32, smem, stream>>>(p);
        ++launches;
    }
    if (plan.fam_begin[KTN_FAM_LSE + 1] > plan.fam_begin[KTN_FAM_LSE]) { launch_family<KTN_FAM_LSE>(p, plan, KTN_TICKET_LSE, num_sms, stream); ++launches; }
    if (plan.fam_begin[KTN_FAM_QUAD + 1] > plan.fam_begin[KTN_FAM_QUAD]) { launch_family<KTN_FAM_QUAD>(p, plan, KTN_TICKET_QUAD, num_sms, stream); ++launches; }
    if (plan.fam_begin[KTN_FAM_SOC + 1] > plan.fam_begin[KTN_FAM_SOC]) { launch_family<KTN_FAM_SOC>(p, plan, KTN_TICKET_SOC, num_sms, stream); ++launches; }
    if (plan.n_total > plan.n_regular) {
        p.chunk_begin = plan.n_regular; p.chunk_end = plan.n_total;
        uint32_t blocks = (plan.n_total - plan.n_regular + 3) / 4;
        if (blocks > (uint32_t)num_sms * 8u) blocks = (uint32_t)num_sms * 8u;
        ktn_big_kernel<EVAL><<<blocks, 128, 0, stream>>>(p);
        ++launches;
    }
    *err = cudaGetLastError();
    return launches;
}

int ktn_launch_round(const KtnRoundParams& p, const KtnLaunchPlan& plan, int num_sms, int max_smem_optin,
                     uint32_t epoch, cudaStream_t stream, cudaEvent_t after_eval, cudaEvent_t after_compact, cudaError_t* err) {
    int launches = launch_eval_part<false>(p, plan, num_sms, max_smem_optin, stream, err);
    if (*err != cudaSuccess) return launches;
    if (after_eval) cudaEventRecord(after_eval, stream);
    const uint32_t nblocks = (uint32_t)((p.num_rows + KTN_CROWS - 1) / KTN_CROWS);
    if (nblocks > 0 && p.topk > 0 && p.mode == KTN_MODE_SEPARATE) {      // keep the k most violated rows
        const uint32_t grid = (uint32_t)num_sms * 4u;
        ktn_topk_key_kernel<<<grid, 256, 0, stream>>>(p, p.topk_key, p.topk_state, (unsigned long long)p.topk);
        for (int pass = 0; pass < 8; ++pass) ktn_topk_select_kernel<<<grid, 256, 0, stream>>>(p, p.topk_key, p.topk_state, pass);
        ktn_topk_ties_kernel<<<nblocks, KTN_CBLOCK, 0, stream>>>(p, p.topk_key, p.topk_state, p.topk_eqcnt);
        ktn_topk_demote_kernel<<<nblocks, KTN_CBLOCK, 0, stream>>>(p, p.topk_key, p.topk_state, p.topk_eqcnt);
        launches += 11;
    }
    if (nblocks > 0) {
        const int scanned = nblocks > 1024u ? 1 : 0;
        if (scanned) { ktn_blkscan_kernel<<<1, 1024, 0, stream>>>(p, nblocks, epoch); ++launches; }
        uint32_t xblocks = (uint32_t)((p.num_rows + KTN_XBLOCK - 1) / KTN_XBLOCK);
        if (xblocks > (uint32_t)num_sms * KTN_XBPS) xblocks = (uint32_t)num_sms * KTN_XBPS;      // resident blocks: every block loops over its share of the cuts
        // programmatic dependent launch (KTN_PDL=0 turns it off): K2 / K3 are placed while their predecessor drains and wait in
        // griddepcontrol.wait.  Measured (10^6 log-sum-exp rows): 125.7 -> 122.6 us per round; with the K1 | K2 event sampled: 118.0
        static const bool pdl = !(getenv("KTN_PDL") && atoi(getenv("KTN_PDL")) == 0);
        if (pdl) {
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
            cudaLaunchConfig_t cfg = {}; cfg.stream = stream; cfg.attrs = at; cfg.numAttrs = 1;
            cfg.gridDim = dim3(nblocks); cfg.blockDim = dim3(KTN_CBLOCK);
            cudaLaunchKernelEx(&cfg, ktn_compact_kernel, p, nblocks, epoch, scanned);
            if (after_compact) cudaEventRecord(after_compact, stream);
            cfg.gridDim = dim3(xblocks); cfg.blockDim = dim3(KTN_XBLOCK);
            cudaLaunchKernelEx(&cfg, ktn_cut_kernel, p, epoch);
        } else {
            ktn_compact_kernel<<<nblocks, KTN_CBLOCK, 0, stream>>>(p, nblocks, epoch, scanned);
            if (after_compact) cudaEventRecord(after_compact, stream);
            ktn_cut_kernel<<<xblocks, KTN_XBLOCK, 0, stream>>>(p, epoch);
        }
        launches += 2;
    }
    *err = cudaGetLastError();
    return launches;
}

// row_ptr of a shard's batch shifted to the combined batch's entry offsets (single-process sharded handles, ktn_api.cu)
namespace { __global__ void ktn_shift_kernel(const int64_t* in, int64_t* out, int64_t n, int64_t add) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i] + add;
} }
void ktn_launch_shift(const int64_t* in, int64_t* out, int64_t n, int64_t add, cudaStream_t stream) {
    if (n <= 0) return;
    const int64_t blocks = (n + 255) / 256;
    ktn_shift_kernel<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, stream>>>(in, out, n, add);
}

// Host push (KtnHostPushParams): the sections of one shard's batch go to their place in the combined batch in pinned host memory.
// Every warp stores whole 128-byte lines of the destination (the loops are aligned to the DESTINATION: misaligned stores cost
// 14 % of the PCIe rate, profiles/microbench/mb8.log); four loads are in flight per thread before the first store.  The fp64
// sections travel as 64-bit integers (x + 0.0 would turn a -0.0 coefficient into +0.0).
namespace {
struct HpPace {     // stores to host memory are posted: unpaced, they fill the queues between the L2 and the PCIe port and every other kernel's
                    // memory traffic waits behind them (measured: the next shard's kernels ran 2.3x slower).  The grid therefore never runs
                    // ahead of `rate` bytes per nanosecond.
    unsigned long long t0; float rate; unsigned long long sent;
    __device__ __forceinline__ void wait(unsigned long long grid_bytes) {
        sent += grid_bytes;
        if (rate <= 0.f) return;
        const unsigned long long due = (unsigned long long)((float)sent / rate);
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        while (t - t0 < due) { __nanosleep(200); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); }
    }
};
template <typename T> __device__ __forceinline__ void hp_section(T* dst, const T* __restrict__ src, unsigned long long n, T add, HpPace& pace) {
    const unsigned long long mis = (reinterpret_cast<unsigned long long>(dst) & 127ull) / sizeof(T);
    const unsigned long long total = n + mis, stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; j + 3 * stride < total; j += 4 * stride) {
        T v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const unsigned long long q = j + k * stride; v[k] = q >= mis ? src[q - mis] : T(0); }
        pace.wait(4ull * stride * sizeof(T));
#pragma unroll
        for (int k = 0; k < 4; ++k) { const unsigned long long q = j + k * stride; if (q >= mis) dst[q - mis] = v[k] + add; }
    }
    pace.wait((total - (j - ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x))) * sizeof(T));
    for (; j < total; j += stride) if (j >= mis) dst[j - mis] = src[j - mis] + add;
}
__global__ void __launch_bounds__(512) ktn_hostpush_kernel(const KtnHostPushParams q) {
    unsigned long long n = __ldcg(q.counts), nz = __ldcg(q.counts + 1);
    const unsigned long long err = __ldcg(q.counts + 6);
    const KtnPackLayout S = ktn_pack_layout(__ldcg(q.counts + 4), __ldcg(q.counts + 5));
    unsigned long long cb = 0, zb = 0; bool stopped = false;      // the reference never reaches the rows behind the first non-finite cut
    for (int k = 0; k < q.nprev; ++k) { cb += __ldcg(q.prev[k]); zb += __ldcg(q.prev[k] + 1); stopped = stopped || __ldcg(q.prev[k] + 6) != ~0ull; }
    if (stopped) { n = 0; nz = 0; }
    if (blockIdx.x == 0 && threadIdx.x == 0) { q.hdr[0] = n; q.hdr[1] = nz; q.hdr[2] = stopped ? ~0ull : err; q.hdr[3] = 0ull; }
    if (n == 0) return;
    const KtnPackLayout& E = q.EL;
    HpPace pace; pace.rate = q.pace; pace.sent = 0ull; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(pace.t0));
    hp_section<long long>(reinterpret_cast<long long*>(q.dst + E.val) + zb, reinterpret_cast<const long long*>(q.src + S.val), nz, 0ll, pace);
    hp_section<int32_t>(reinterpret_cast<int32_t*>(q.dst + E.col) + zb, reinterpret_cast<const int32_t*>(q.src + S.col), nz, 0, pace);
    hp_section<long long>(reinterpret_cast<long long*>(q.dst + E.row_id) + cb, reinterpret_cast<const long long*>(q.src + S.row_id), n, 0ll, pace);
    hp_section<long long>(reinterpret_cast<long long*>(q.dst + E.row_ptr) + cb, reinterpret_cast<const long long*>(q.src + S.row_ptr), n, (long long)zb, pace);
    hp_section<long long>(reinterpret_cast<long long*>(q.dst + E.lo) + cb, reinterpret_cast<const long long*>(q.src + S.lo), n, 0ll, pace);
    hp_section<long long>(reinterpret_cast<long long*>(q.dst + E.hi) + cb, reinterpret_cast<const long long*>(q.src + S.hi), n, 0ll, pace);
    if (!q.lean) {
        hp_section<long long>(reinterpret_cast<long long*>(q.dst + E.g) + cb, reinterpret_cast<const long long*>(q.src + S.g), n, 0ll, pace);
        hp_section<long long>(reinterpret_cast<long long*>(q.dst + E.viol) + cb, reinterpret_cast<const long long*>(q.src + S.viol), n, 0ll, pace);
        hp_section<long long>(reinterpret_cast<long long*>(q.dst + E.b) + cb, reinterpret_cast<const long long*>(q.src + S.b), n, 0ll, pace);
    }
}
}
void ktn_launch_hostpush(const KtnHostPushParams& q, int blocks, cudaStream_t stream) { ktn_hostpush_kernel<<<blocks, 512, 0, stream>>>(q); }

// boundroutine's ladder (ktn_separate_ladder): x = scale * ray, and "is any nonlinear row violated at the point just evaluated?"
namespace {
__global__ void ktn_scale_kernel(const double* ray, double* x, int64_t n, double scale) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = scale * ray[i];
}
__global__ void ktn_anyviol_kernel(const double* g, const double* lb, const double* ub, const uint8_t* nl, int64_t m, double f_tol, unsigned int* flag) {
    bool v = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
        if (nl[i]) { const double gi = g[i]; v = v || !((gi >= lb[i] - f_tol) && (gi <= ub[i] + f_tol)); }      // src/separators.jl:120 (NaN: not satisfied)
    if (__any_sync(0xffffffffu, v) && (threadIdx.x & 31u) == 0) *flag = 1u;
}
}
void ktn_launch_scale(const double* ray, double* x, int64_t n, double scale, cudaStream_t stream) {
    const int64_t b = (n + 255) / 256;
    ktn_scale_kernel<<<(unsigned)(b < 592 ? (b ? b : 1) : 592), 256, 0, stream>>>(ray, x, n, scale);
}
void ktn_launch_anyviol(const KtnRoundParams& p, const uint8_t* row_nl, unsigned int* flag, cudaStream_t stream) {
    const int64_t b = (p.num_rows + 255) / 256;
    ktn_anyviol_kernel<<<(unsigned)(b < 1184 ? (b ? b : 1) : 1184), 256, 0, stream>>>(p.g_row, p.row_lb, p.row_ub, row_nl, p.num_rows, p.f_tol, flag);
}

int ktn_launch_eval(const KtnRoundParams& p, const KtnLaunchPlan& plan, int num_sms, int max_smem_optin,
                    cudaStream_t stream, cudaError_t* err) {
    int launches = launch_eval_part<true>(p, plan, num_sms, max_smem_optin, stream, err);
    if (*err == cudaSuccess) *err = cudaMemsetAsync(p.ticket, 0, 4 * KTN_TICKETS, stream);   // the eval path has no K2 to re-arm the tickets
    return launches;
}
