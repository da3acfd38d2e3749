"""Separator plugin API and the B200 separator.

Mirrors reference src/separators.jl: `AbstractKatanaSeparator` with the four hooks
initialize! / precompute! / isconstrsat / gencut (src/separators.jl:23,34,43,53), and, in place of
`KatanaFirstOrderSeparator` (src/separators.jl:58-120), `KatanaGPUSeparator`, which keeps the same
state (g, Jacobian rows, xstar) on the device behind the C ABI of include/ktn.h.
"""
import numpy as np

from .binding import FLAG_DIRECT_VIEW, FLAG_EAGER_VIEW, FLAG_LEAN_VIEW, KTN_NUMERIC_NONFINITE, load_cuda_library
from .nlpeval import rows_to_wire


class AffExpr:
    """The JuMP.AffExpr the reference's gencut returns: sum coeffs[k] * x[vars[k]] + constant."""
    __slots__ = ("vars", "coeffs", "constant")

    def __init__(self, vars_, coeffs, constant):
        self.vars, self.coeffs, self.constant = np.asarray(vars_, np.int64), np.asarray(coeffs, np.float64), float(constant)


class AbstractKatanaSeparator:
    def initialize(self, linear_model, num_var, num_constr, oracle):          # src/separators.jl:23
        raise NotImplementedError("Not implemented: Katana.initialize!")

    def gencut(self, xstar, bounds, i):                                       # src/separators.jl:34
        raise NotImplementedError("Not implemented: Katana.gencut!")

    def isconstrsat(self, i, lb, ub, f_tol):                                  # src/separators.jl:43
        raise NotImplementedError("Not implemented: Katana.isconstrsat")

    def precompute(self, xstar):                                              # src/separators.jl:53
        return None

    def set_bounds(self, l_constr, u_constr):
        """Added hook (default no-op): lets a batched separator test all rows in one device pass."""
        return None


class KatanaGPUSeparator(AbstractKatanaSeparator):
    """First-order separator whose precompute! is one device round (include/ktn.h: ktn_separate).

    The batched `separate()` is what optimize! uses; `isconstrsat` / `gencut` stay available with
    the reference's per-row semantics and answer from the last round.
    `library` defaults to the CUDA library; there is no CPU path in the product (tests may pass
    another implementation of the same C ABI as the checker).
    """

    def __init__(self, library=None, topk=0, ngpus=1, devices=None, pipeline=None, direct=False):
        """ngpus > 1: the ONE separator of the model shards its rows over `ngpus` devices of this process (ktn_options.ngpus);
        `separate` still returns one combined batch in ascending row order to the one LP master.
        pipeline = S > 1: every device's rows are split into S consecutive shards whose cut downloads start as soon as each
        shard has finished (KTN_FLAG_EAGER_VIEW): the PCIe transfer of the first shards overlaps the kernels of the later ones.
        On ONE device the shards share a stream and a small kernel per shard stores its cuts into the pinned batch (ktn_api.cu
        group_round_pushed).  pipeline = None: the measured default (B200, 10^6 log-sum-exp rows, 10^5 cuts per round: one shard
        0.604 ms, two 0.575, three 0.587, four 0.632, eight 0.76): two shards between PIPELINE_MIN_ROWS and PIPELINE_MAX_ROWS rows, else one.
        direct (one device, one shard): the round's kernels store the batch straight into the pinned host buffer the views point
        into (KTN_FLAG_DIRECT_VIEW); measured no faster than the download (0.622 against 0.604 ms: stores from the SMs reach
        43-48 GB/s over PCIe, the copy engine 56), so it is off by default."""
        self._lib = library
        self.topk = topk
        self.ngpus, self.devices, self.pipeline, self.direct = ngpus, devices, pipeline, direct
        self.handle = None
        self.has_ladder = True
        self.last = None           # CutBatch of the last precompute!
        self.xstar = None
        self.g = None

    # initialize!(sep, linear_model, num_var, num_constr, oracle)  -- src/separators.jl:81-107
    def initialize(self, linear_model, num_var, num_constr, oracle, f_tol=1e-6, cut_coef_rng=1e9):
        lib = self._lib if self._lib is not None else load_cuda_library()
        self.linear_model, self.oracle = linear_model, oracle
        oracle.initialize(["ExprGraph"])                       # MathProgBase.initialize(oracle, ...) :88
        if self.handle is not None:                            # one separator is reused across models (test/runtests.jl:24)
            self.handle.close()
        # lean views: optimize! hands (row_ptr, col, val, lo, hi) to the LP; g / viol / bconst stay on the device
        lb = np.full(num_constr, -np.inf); ub = np.full(num_constr, np.inf)
        wire = rows_to_wire(oracle, num_constr, lb, ub)
        opts = self.handle_options(int((wire.flags & 1).sum()))      # the rows a round tests and cuts (nlconstr_ixs) decide the pipelining
        self.has_ladder = opts.get("ngpus", 0) <= 1        # ktn_separate_ladder runs on plain handles (one device, one shard)
        self.handle = lib.create(f_tol=f_tol, cut_coef_rng=cut_coef_rng, topk=self.topk, **opts)
        self.num_var, self.num_constr = num_var, num_constr
        self.handle.load(num_var, wire)
        self.l_constr, self.u_constr = lb, ub
        self.last = self.g = self.xstar = None

    PIPELINE_MIN_ROWS = 750_000        # measured at 10^6 rows only (29 us gained of 604); the gain shrinks with the kernel time it overlaps
    PIPELINE_MAX_ROWS = 4_000_000      # beyond: the worst-case pinned batch of a pipelined handle would pass the library's 1 GiB limit

    def shards_per_device(self, num_nl_rows):
        if self.pipeline is not None:
            return max(int(self.pipeline), 1)
        return 2 if self.PIPELINE_MIN_ROWS <= num_nl_rows <= self.PIPELINE_MAX_ROWS and not self.topk else 1

    def handle_options(self, num_nl_rows=0):
        """ktn_options of this separator: lean views (optimize! hands row_ptr, col, val, lo, hi to the LP), the device list, eager downloads."""
        devs = list(self.devices) if self.devices else list(range(max(self.ngpus, 1)))
        per = self.shards_per_device(num_nl_rows) if len(devs) == 1 else max(int(self.pipeline or 1), 1)      # the measured default is a one-device result
        shards = [d for d in devs for _ in range(per)]
        if len(shards) <= 1:
            return dict(flags=FLAG_LEAN_VIEW | (FLAG_DIRECT_VIEW if self.direct else 0), device=devs[0] if self.devices else -1)
        return dict(flags=FLAG_LEAN_VIEW | (FLAG_EAGER_VIEW if per > 1 else 0), ngpus=len(shards), devices=shards)

    def set_params(self, f_tol, cut_coef_rng):
        self.handle.set_params(f_tol, cut_coef_rng, self.topk)

    def set_bounds(self, l_constr, u_constr):
        self.l_constr = np.asarray(l_constr, np.float64).copy(); self.u_constr = np.asarray(u_constr, np.float64).copy()
        self.handle.set_bounds(self.l_constr, self.u_constr)

    # precompute!(sep, xstar) -- src/separators.jl:111-116
    def precompute(self, xstar):
        self.xstar = np.asarray(xstar, np.float64)
        # zero-copy views of the library's pinned cut buffer: optimize! hands the rows to the LP before the next round
        self.last = self.handle.separate(self.xstar, view=True)
        self.g = None

    def separate(self, xstar):
        """Batched round: every violated NL row as a CSR cut batch (ascending row order)."""
        self.precompute(xstar)
        return self.last

    def separate_ladder(self, ray, n_first=2, n_last=1023):
        """boundroutine's search along an unbounded ray (src/model.jl:175-197) as ONE call: (exponent hit or -1, CutBatch)."""
        n_hit, self.last = self.handle.separate_ladder(ray, n_first, n_last, view=True)
        self.xstar = (2.0 ** (n_hit if n_hit >= 0 else n_last)) * np.asarray(ray, np.float64)
        self.g = None
        return n_hit, self.last

    # isconstrsat(sep, i, lb, ub, f_tol) -- src/separators.jl:120
    def isconstrsat(self, i, lb, ub, f_tol):
        if self.g is None:
            self.g = self.handle.get_g()
        return bool((self.g[i] >= lb - f_tol) and (self.g[i] <= ub + f_tol))

    # gencut(sep, xstar, bounds, i) -> AffExpr -- src/separators.jl:118, src/algorithms.jl:3-18
    def gencut(self, xstar, bounds, i):
        b = self.handle.gencut_rows(self.xstar if self.xstar is not None else xstar, np.array([i], np.int64), round_coefs=False)
        if b.n_cuts != 1:
            cols = self.handle.jac_structure()
            s, e = cols[0][i], cols[0][i + 1]
            return AffExpr(cols[1][s:e], np.full(e - s, np.nan), np.nan)     # non-finite row: _addcut will flag :Error
        cols, vals = b.row(0)
        return AffExpr(cols, vals, b.bconst[0])


def linear_oa_cut(sep, a, b, i):
    """Name kept from src/algorithms.jl:3: the first-order cut of row i at the precomputed point."""
    return sep.gencut(a, b, i)
