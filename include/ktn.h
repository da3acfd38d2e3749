/* ktn.h -- C ABI of the B200 separation-round library (libktn.so).
 *
 * This is the drop-in boundary for ONE path of lanl-ansi/Katana.jl: the ECP
 * separation round (evaluate g(x*), reverse-mode sparse Jacobian rows, violation
 * test, first-order cut emission).  Every entry point names the reference
 * interface it replaces (paths relative to the reference repository root).
 * The Julia host reaches these by `ccall`, the Python mirror by `ctypes`;
 * INTEGRATION.md shows both bindings.  The CPU oracle (oracle/ktn_oracle.c)
 * exports the same symbols so tests can drive either library.
 *
 * Conventions: plain pointers and sizes only; caller-owned HOST buffers unless a
 * name says `_device`; 0-based indices on the wire (int32 columns, int64 rows /
 * row pointers); every function returns an int status: 0 = OK, < 0 = usage /
 * CUDA / NCCL error (message via ktn_last_error), > 0 = numeric condition.
 * No exception crosses the ABI.  A handle is not thread-safe.
 */
#ifndef KTN_H
#define KTN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- expression wire format (MathProgBase `constr_expr` / `obj_expr` flattened) ----
 * One expression = its nodes in PREFIX (pre-order) order.  For node i:
 *   op[i]  one of KTN_OP_*
 *   arg[i] KTN_OP_VAR: 0-based variable index; call nodes: number of children; CONST: 0
 *   val[i] KTN_OP_CONST: the value; otherwise ignored
 * expr_ptr[r] .. expr_ptr[r+1] delimits row r's nodes (offsets relative to the batch).
 * Semantics of evaluation and of the reverse sweep follow JuMP 0.18's
 * ReverseDiffSparse tape interpreter (forward_eval / reverse_eval / reverse_extract),
 * which is what `eval_g` / `eval_jac_g` run at reference src/separators.jl:112-113.
 */
enum {
    KTN_OP_CONST = 0,
    KTN_OP_VAR   = 1,
    KTN_OP_ADD   = 2,  /* n-ary, summed left to right starting from 0.0 */
    KTN_OP_SUB   = 3,  /* binary */
    KTN_OP_MUL   = 4,  /* n-ary */
    KTN_OP_DIV   = 5,  /* binary */
    KTN_OP_POW   = 6,  /* binary: base ^ exponent */
    KTN_OP_NEG   = 7,  /* unary minus */
    KTN_OP_EXP   = 8,
    KTN_OP_LOG   = 9,
    KTN_OP_SQRT  = 10,
    KTN_OP_ABS   = 11,
    KTN_OP_SIN   = 12,
    KTN_OP_COS   = 13,
    KTN_OP_IFELSE = 14, /* ifelse(cond, a, b): a where cond == 1.0, else b; BOTH branches are evaluated (JuMP's tape is linear); the
                           unselected branch and the condition receive a zero partial */
    KTN_OP_LE    = 15,  /* binary comparisons (JuMP allows them as ifelse conditions): 1.0 / 0.0, NaN compares false; zero partials */
    KTN_OP_LT    = 16,
    KTN_OP_GE    = 17,
    KTN_OP_GT    = 18,
    KTN_OP_EQ    = 19,
    KTN_OP__COUNT = 20
};

/* per-row flags */
enum {
    KTN_ROW_NL    = 1, /* row is in `nlconstr_ixs` (reference src/model.jl:120,148): tested every round */
    KTN_ROW_DENSE = 2  /* Jacobian row lists every column 0..num_var-1 (the epigraph row, src/nlpeval.jl:49-54) */
};

/* status codes */
enum {
    KTN_OK = 0,
    KTN_NUMERIC_NONFINITE = 1,  /* a selected cut had a non-finite coefficient: reference `m.status = :Error` (src/model.jl:69-73) */
    KTN_ERR_USAGE = -1,
    KTN_ERR_CUDA = -2,
    KTN_ERR_NCCL = -3,
    KTN_ERR_UNSUPPORTED = -4,
    KTN_ERR_NOMEM = -5
};

typedef struct ktn_handle ktn_handle;

/* Mirrors the fields of KatanaModelParams the round reads (reference src/Katana.jl:12-19,
 * src/solver.jl:34-43): f_tol (src/model.jl:273) and cut_coef_rng (src/model.jl:276).
 * topk is a build extension: 0 = every violated row becomes a cut (reference behaviour); k > 0 = of the violated rows only
 * the k ranked first by (NaN first, violation max(lb - g, g - ub) descending, row index ascending) become cuts, still emitted
 * in ascending row order; the first non-finite row among THOSE ends the batch.  On a sharded handle the selection is LOCAL to the rank's rows: the
 * union of the ranks' selections contains the global top-k, which is one deterministic merge of the gathered batch away
 * (katana.jl_b200/sharding.py merge_topk). */
typedef struct ktn_options {
    int32_t struct_size;   /* sizeof(ktn_options), for versioning */
    int32_t device;        /* CUDA device ordinal; -1 = current device */
    double  f_tol;         /* default 1e-6 */
    double  cut_coef_rng;  /* default 1e9 */
    int64_t topk;          /* 0 = all violated rows */
    int32_t flags;         /* KTN_FLAG_* */
    int32_t reserved;
    /* Single-process sharded operation: ngpus > 1 makes ONE handle drive `ngpus` devices.  The rows are
     * split into contiguous ranges, one per device (devices[s]; a negative entry means device s); every call of this header then
     * acts on the whole problem: ktn_separate runs the round on all devices at once, ktn_fetch_cuts / ktn_fetch_cuts_view deliver
     * ONE batch in ascending row order (rank-major concatenation; every device downloads its share over its own PCIe link), with
     * the reference's error semantics (the batch ends at the first non-finite row, src/model.jl:69-73, :278).  The same device
     * may be listed more than once (tests on one GPU).  Not available on such a handle: topk, ktn_set_stream and the
     * device-resident / ktn_comm_* entry points (those serve the one-process-per-GPU mode). */
    int32_t ngpus;         /* 0 or 1: one device (`device`) */
    int32_t devices[16];
} ktn_options;
#define KTN_MAX_GPUS 16

/* ktn_options.flags */
enum {
    KTN_FLAG_LEAN_VIEW = 1, /* ktn_fetch_cuts_view downloads only what the LP needs (row_id, row_ptr, col, val, lo, hi);
                               g, viol and bconst come back NULL: 11 % less PCIe traffic per round */
    KTN_FLAG_EAGER_VIEW = 4,  /* multi-device handles (ktn_options.ngpus > 1; the same device may be listed several times to pipeline ONE
                               device): ktn_separate starts every shard's cut download the moment that shard has finished, into a
                               pinned buffer laid out for the worst case (every nonlinear row cut: 32 bytes per nonlinear row + 12 per
                               Jacobian entry, twice); ktn_fetch_cuts_view then only waits.  The download of the first shards overlaps
                               the kernels of the later ones.  Ignored when that buffer would exceed 1 GiB.  When every shard is on ONE
                               device the shards share a stream, x* is uploaded once and a small kernel per shard stores its cuts into
                               the (mapped) pinned buffer: one host synchronisation per round. */
    KTN_FLAG_DIRECT_VIEW = 8, /* single-device handles without a communicator: the round's kernels store the cut batch straight into
                               mapped pinned host memory (two buffers of the worst-case size -- every row cut -- alternate) instead of
                               device memory, so the PCIe transfer runs WHILE the cuts are built and ktn_fetch_cuts_view copies
                               nothing.  With KTN_FLAG_LEAN_VIEW the g | viol | b sections are not produced at all.  Ignored when
                               the two buffers would exceed 2 GiB, on multi-device handles, and once ktn_comm_init has been called. */
    KTN_FLAG_TIME_KERNELS = 2 /* ktn_timings.compact_ms / cut_ms are timed separately (one more CUDA event per round, between the
                               compaction and the cut kernel); otherwise compact_ms covers both and cut_ms is 0 */
};

/* kernel_ms / eval_ms / compact_ms / cut_ms are SAMPLED: the CUDA events around and between the kernels of a round are recorded in
   one round of eight (KTN_K1_EVENT_EVERY; every round with KTN_FLAG_TIME_KERNELS). */
typedef struct ktn_timings {
    double h2d_ms;        /* x* upload (timed with KTN_FLAG_TIME_KERNELS only: two CUDA events per call otherwise saved) */
    double kernel_ms;     /* separation kernels, CUDA events on the library stream */
    double exchange_ms;   /* sharded handles: the last synced exchange's transfer on the exchange stream (the push kernel / the
                             ncclAllGather, CUDA events; it runs beside the following rounds, so it is NOT part of kernel_ms) */
    double d2h_ms;        /* cut download in ktn_fetch_cuts (KTN_FLAG_TIME_KERNELS only) */
    int64_t launches;     /* kernels launched by this library since creation */
    int64_t rounds;       /* separation rounds run since creation */
    double eval_ms;       /* last round: evaluation kernel(s) (K1) */
    double compact_ms;    /* last round: compaction kernel (K2); without KTN_FLAG_TIME_KERNELS: K2 + K3 */
    double eval_ms_sum;   /* sums over every round timed since creation (CUDA events on the round's stream) */
    double compact_ms_sum;
    int64_t rounds_timed;
    double cut_ms;        /* last round: cut kernel (K3: cuts of the family rows) */
    double cut_ms_sum;
    double exchange_ms_sum; /* sharded handles: sum of exchange_ms over the exchanges synced since creation */
    int64_t exchanges_timed;
} ktn_timings;

/* lifetime -- replaces constructing KatanaFirstOrderSeparator() (src/separators.jl:58-77). */
int  ktn_create(const ktn_options* opts, ktn_handle** out);
void ktn_destroy(ktn_handle* h);
const char* ktn_last_error(ktn_handle* h);
/* "cuda" for libktn.so, "oracle" for the CPU restatement. */
const char* ktn_backend(void);
int  ktn_set_params(ktn_handle* h, double f_tol, double cut_coef_rng, int64_t topk);

/* problem loading -- replaces initialize!(sep, linear_model, num_var, num_constr, oracle)
 * (src/separators.jl:81-107): MathProgBase.initialize + jac_structure + per-row bucketing.
 * May be called again on the same handle (test/runtests.jl:24 reuses one solver): a new
 * ktn_load_begin frees the previous problem.  Rows may be added in batches, in ascending
 * row order; `first_row` is the 0-based global index of the batch's first row. */
int ktn_load_begin(ktn_handle* h, int64_t num_var, int64_t num_constr);
int ktn_add_rows(ktn_handle* h, int64_t first_row, int64_t nrows,
                 const int64_t* expr_ptr, const int32_t* op, const int32_t* arg, const double* val,
                 const double* lb, const double* ub, const uint8_t* flags);
int ktn_load_end(ktn_handle* h);
/* Replace every loaded row's (lb, ub).  In the reference the bounds are model state passed to
 * isconstrsat / gencut per call (src/model.jl:273-277), not separator state. */
int ktn_set_bounds(ktn_handle* h, const double* lb, const double* ub);

/* Jacobian structure of the loaded rows -- replaces MathProgBase.jac_structure as consumed by
 * src/separators.jl:92-100: per row, the column list in evaluator entry order (ascending
 * unique columns for tape rows; 0..num_var-1 for KTN_ROW_DENSE rows).
 * row_ptr has (rows loaded)+1 entries; pass cols = NULL to query sizes only. */
int ktn_jac_structure(ktn_handle* h, int64_t* row_ptr, int32_t* cols);
int64_t ktn_num_rows(ktn_handle* h);
int64_t ktn_jac_nnz(ktn_handle* h);

/* One separation round at x* (HOST pointer, num_var doubles) -- replaces the loop body
 * src/model.jl:265-283: precompute! (src/separators.jl:111-116), isconstrsat (:120) for each
 * i in nlconstr_ixs ascending, gencut -> linear_oa_cut (src/algorithms.jl:3-18), round_coefs
 * (src/model.jl:200-207), _addcut's finiteness check and bound shift (src/model.jl:68-79).
 * Returns KTN_NUMERIC_NONFINITE when a selected row had a non-finite coefficient; cuts of
 * the selected rows BEFORE that row are still delivered (the reference adds them before
 * it returns :Error), and *err_row receives the offending row (else -1). */
int ktn_separate(ktn_handle* h, const double* xstar, int64_t* n_cuts, int64_t* nnz_cuts, int64_t* err_row);

/* Unconditional cuts for the listed rows (ascending 0-based indices) -- replaces the gencut
 * calls of loadproblem! on linear rows, the linear objective row and the initial vertex cut
 * (src/model.jl:115-118,129,160-163) and boundroutine's per-row gencut.  round_coefs != 0
 * applies src/model.jl:200-207 (the reference does so only for the vertex cut, :162). */
int ktn_gencut_rows(ktn_handle* h, const double* x, const int64_t* rows, int64_t nrows, int round_coefs,
                    int64_t* n_cuts, int64_t* nnz_cuts, int64_t* err_row);

/* boundroutine's ladder (src/model.jl:175-197): points x_n = 2^n * ray for n = n_first .. n_last (the reference: 2 .. 1023); every
 * nonlinear row is tested at each point in turn, and at the FIRST point where a row is violated the violated rows are cut (as
 * ktn_separate does) and the search stops.  The reference runs up to 1022 sequential precompute! rounds for this; here the points
 * are evaluated in batches of 16 without a host round trip in between (forward evaluation + one violation flag per point), then
 * ONE separation round runs at the point that was hit.  *n_hit receives that exponent (-1: no point violated any row: no cuts).
 * The cuts are fetched as after ktn_separate.  Not available on a multi-device handle. */
int ktn_separate_ladder(ktn_handle* h, const double* ray, int32_t n_first, int32_t n_last, int32_t* n_hit,
                        int64_t* n_cuts, int64_t* nnz_cuts, int64_t* err_row);

/* Download the cuts of the last ktn_separate / ktn_gencut_rows.  Cut c (ascending row order):
 *   row_id[c]                      0-based constraint index
 *   row_ptr[c] .. row_ptr[c+1]     its entries in col[] / val[]   (row_ptr has n_cuts+1 entries)
 *   lo[c], hi[c]                   LP row bounds lb - b, ub - b   (src/model.jl:74-75)
 *   g[c], viol[c]                  g_i(x*) and max(lb - g, g - ub)
 *   bconst[c]                      the AffExpr constant b of src/algorithms.jl:8-17 (lo = lb - b, hi = ub - b)
 * Any output pointer may be NULL to skip that array. */
int ktn_fetch_cuts(ktn_handle* h, int64_t* row_id, int64_t* row_ptr, int32_t* col, double* val,
                   double* lo, double* hi, double* g, double* viol, double* bconst);

/* Zero-copy variant of ktn_fetch_cuts: the same nine arrays, delivered as pointers into LIBRARY-OWNED pinned host memory
 * (one device->host transfer per array, no host-side copy).  Two buffers alternate: a view stays valid until the SECOND
 * later call of ktn_fetch_cuts_view on the handle, ktn_load_begin or ktn_destroy.  This is the call the batched
 * addconstr! hand-off of optimize! uses: the cut rows go to the LP solver straight from the view. */
typedef struct ktn_cut_view {
    int64_t n_cuts, nnz;
    const int64_t* row_id; const int64_t* row_ptr;   /* n_cuts, n_cuts + 1 */
    const int32_t* col; const double* val;           /* nnz */
    const double* lo; const double* hi; const double* g; const double* viol; const double* bconst;   /* n_cuts */
} ktn_cut_view;
int ktn_fetch_cuts_view(ktn_handle* h, ktn_cut_view* out);

/* All constraint values of the last round (sep.g, src/separators.jl:113) -- backs the
 * per-row isconstrsat(sep, i, lb, ub, f_tol) compatibility hook (src/separators.jl:120). */
int ktn_get_g(ktn_handle* h, double* g_out);
/* eval_g at an arbitrary point, all loaded rows (debug / parity). */
int ktn_eval_g(ktn_handle* h, const double* x, double* g_out);

int ktn_timings_get(ktn_handle* h, ktn_timings* out);

/* ---- device-resident round (benchmarks, GPU-side LP masters) ---- */
/* Use an external CUDA stream (cudaStream_t cast to void*); NULL restores the library stream. */
int ktn_set_stream(ktn_handle* h, void* cuda_stream);
/* Enqueue one round on the stream with x* already in device memory; does not synchronise. */
int ktn_separate_device_async(ktn_handle* h, const double* d_xstar);
/* Wait for the last enqueued round and read its counts. */
int ktn_sync_counts(ktn_handle* h, int64_t* n_cuts, int64_t* nnz_cuts, int64_t* err_row);
/* Algorithmic bytes of one round (SURVEY.md section 8d formula) for the last synced round. */
int64_t ktn_algorithmic_bytes(ktn_handle* h);

/* ---- sharded operation: one handle per GPU / process, rows [row_begin,row_end) each ---- */
/* 128-byte NCCL unique id, created on rank 0 and broadcast by the host's own plumbing. */
int ktn_comm_unique_id(void* id128);
int ktn_comm_init(ktn_handle* h, int32_t nranks, int32_t rank, const void* id128);
/* Global index of this handle's first row: row ids delivered by ktn_fetch_cuts / ktn_fetch_gathered are shifted by it. */
int ktn_set_row_offset(ktn_handle* h, int64_t first_global_row);
/* Combine every rank's compacted cuts (rank-major = ascending row order) on every GPU over
 * NCCL/NVLink; enqueued on the stream after the last round.  Totals are returned after a sync. */
int ktn_allgather_cuts_async(ktn_handle* h);
int ktn_sync_gathered(ktn_handle* h, int64_t* total_cuts, int64_t* total_nnz);
/* Transport the exchange runs on, decided collectively at the first ktn_allgather_cuts_async: 0 = not sharded / not decided yet,
 * 1 = NCCL all-gather, 2 = peer push (the round's cut blob is stored into every rank's receive arena over NVLink by a CUDA
 * kernel; arenas are shared through CUDA IPC).  KTN_EXCHANGE=nccl in the environment of any rank forces 1 on all ranks. */
int ktn_exchange_transport(ktn_handle* h);
int ktn_fetch_gathered(ktn_handle* h, int64_t* row_id, int64_t* row_ptr, int32_t* col, double* val,
                       double* lo, double* hi, double* g, double* viol, double* bconst);
/* ktn_sync_gathered / ktn_fetch_gathered return KTN_NUMERIC_NONFINITE on EVERY rank when a rank's round hit a non-finite cut: the
 * gathered batch then ends at that row, as the reference's loop does (src/model.jl:69-73, :278), and this returns its global index
 * (-1: none). */
int ktn_gathered_error_row(ktn_handle* h, int64_t* err_row);

/* ---- deterministic synthetic instances (test / bench support) ----
 * kind: 0 = sparse convex QCQP (config 2), 1 = log-sum-exp (config 3), 2 = SOC-like risk rows (config 4), 3 = portfolio (config 4 at
 * size: nine sparse linear rows, not flagged KTN_ROW_NL, for every SOC-like row).
 * Two-call protocol: call with op = NULL to get *n_nodes, then with caller buffers
 * (expr_ptr: nrows+1, op/arg/val: n_nodes, lb/ub: nrows, flags: nrows, xstar: num_var).
 * Rows [row_begin, row_begin+nrows) of the instance (num_var, seed) are produced; a row's
 * content depends only on (seed, global row index), so shards agree with the whole. */
int ktn_synth_rows(int32_t kind, uint64_t seed, int64_t num_var, int64_t row_begin, int64_t nrows,
                   int64_t* n_nodes, int64_t* expr_ptr, int32_t* op, int32_t* arg, double* val,
                   double* lb, double* ub, uint8_t* flags);
int ktn_synth_point(int32_t kind, uint64_t seed, int64_t num_var, double* xstar);

#ifdef __cplusplus
}
#endif
#endif /* KTN_H */
