"""The ECP driver: a line-by-line host mirror of reference src/model.jl around the B200 separator.

Everything numerical in the loop body (src/model.jl:265-283) runs behind the C ABI; this file
keeps the reference's orchestration: loadproblem! (LP set-up, linear rows copied as cuts at 0,
epigraph lifting of a nonlinear objective), the unbounded-ray presolve, the ECP loop, status and
getter semantics.  The LP master stays on the host (lp.HighsLP stands in for GLPK/Clp).
"""
import math
import time
import warnings

import numpy as np

from .binding import KTN_NUMERIC_NONFINITE
from .nlpeval import EpigraphNLPEvaluator
from .separators import AffExpr, KatanaGPUSeparator


class KatanaModelParams:                                       # src/Katana.jl:12-19
    def __init__(self, f_tol, iter_cap, log_level, cut_coef_rng, obj_eps, separator, cut_purge_age=0, cut_filter_duplicates=False):
        self.f_tol, self.iter_cap, self.log_level = f_tol, iter_cap, log_level
        self.cut_coef_rng, self.obj_eps, self.separator = cut_coef_rng, obj_eps, separator
        # cut management, SURVEY 8(f) item 2 -- extensions, OFF by default (the reference never removes or filters cuts, src/model.jl:215):
        # cut_purge_age = A > 0: a loop cut that has been slack at A consecutive LP optima is removed from the LP, but only after a
        # round in which the LP bound moved (a master whose bound is monotone and strictly moving cannot cycle);
        # cut_filter_duplicates: a cut whose (columns, coefficients, bounds) are bit-identical to one added earlier is not added again.
        self.cut_purge_age, self.cut_filter_duplicates = cut_purge_age, cut_filter_duplicates


def round_coefs(cut, cut_coef_rng):                            # src/model.jl:200-207
    mx = cut.coeffs[0]
    for c in cut.coeffs[1:]:
        mx = c if (mx != mx or c != c) and c != c else (mx if mx != mx else max(mx, c))
    for i in range(len(cut.coeffs)):
        if cut.coeffs[i] + cut_coef_rng < mx:
            cut.coeffs[i] = 0.0


class KatanaNonlinearModel:                                    # src/model.jl:9-61
    def __init__(self, lp_solver, features, params):
        self.lp_solver = lp_solver
        self.linear_model = None
        self.status = "None"
        self.objval = math.nan
        self.params = params
        self.features = {"VisData": False}
        for f in features:
            self.features[f] = True                            # src/model.jl:49-52
        self.nlconstr_ixs = []
        self.linear_cuts, self.lp_sols = [], []
        self.iter = 0
        self.in_loop = False
        self.numcuts = 0
        self.soltime = 0.0
        self.round_log = []                                    # per-round timers
        self.cut_keys = set()                                  # duplicate filter (params.cut_filter_duplicates)
        self.cuts_purged = self.cuts_filtered = 0

    # _addcut(m, cut, lb, ub) -- src/model.jl:68-79
    def _addcut(self, cut, lb, ub):
        if not np.all(np.isfinite(cut.coeffs)):
            warnings.warn("Nonlinear constraint or objective likely undefined within domain")
            self.status = "Error"
            return
        c = cut.constant
        self.linear_model.addconstr(cut.vars, cut.coeffs, lb - c, ub - c)
        self.numcuts += 1
        if self.features["VisData"]:
            self.linear_cuts.append((cut.vars.copy(), cut.coeffs.copy(), lb - c, ub - c))

    def _addbatch(self, batch):
        """Batched _addcut over a CutBatch: lo/hi already carry lb - c, ub - c (src/model.jl:74-75)."""
        skip = None
        if self.params.cut_filter_duplicates and batch.n_cuts:
            skip = np.zeros(batch.n_cuts, bool)
            for c in range(batch.n_cuts):
                s, e = batch.row_ptr[c], batch.row_ptr[c + 1]
                key = (batch.col[s:e].tobytes(), batch.val[s:e].tobytes(), float(batch.lo[c]), float(batch.hi[c]))
                if key in self.cut_keys: skip[c] = True
                else: self.cut_keys.add(key)
            self.cuts_filtered += int(skip.sum())
        if skip is not None or self.params.cut_purge_age > 0:
            self.linear_model.addconstrs_csr(batch.row_ptr, batch.col, batch.val, batch.lo, batch.hi, managed=self.in_loop and self.params.cut_purge_age > 0, skip=skip)
        else:
            self.linear_model.addconstrs_csr(batch.row_ptr, batch.col, batch.val, batch.lo, batch.hi)
        self.numcuts += batch.n_cuts - (int(skip.sum()) if skip is not None else 0)
        if self.features["VisData"]:
            for c in range(batch.n_cuts):
                cols, vals = batch.row(c)
                self.linear_cuts.append((cols.copy(), vals.copy(), batch.lo[c], batch.hi[c]))
        if batch.status == KTN_NUMERIC_NONFINITE:
            warnings.warn("Nonlinear constraint or objective likely undefined within domain")
            self.status = "Error"

    # MathProgBase.loadproblem!(m, num_var, num_constr, l_var, u_var, l_constr, u_constr, sense, d) -- src/model.jl:81-173
    def loadproblem(self, num_var, num_constr, l_var, u_var, l_constr, u_constr, sense, d):
        lm = self.linear_model = self.lp_solver()              # :89
        for i in range(num_var):
            lm.addvar(l_var[i], u_var[i])                      # :92
        vertex = np.full(num_var, np.nan)                      # :93
        if lm.solve() == "Optimal":                            # :94-97
            vertex = lm.getsolution()
        self.num_var, self.num_constr = num_var, num_constr
        self.l_constr, self.u_constr = list(map(float, l_constr)), list(map(float, u_constr))
        self.nlconstr_ixs = []

        # linear rows and the objective are recreated from first-order cuts at 0 (:105-133)
        sep_lib = getattr(self.params.separator, "_lib", None)
        fsep = KatanaGPUSeparator(library=sep_lib)             # :110 (default algo)
        epi_d = EpigraphNLPEvaluator(d, num_var + 1, num_constr + 1)   # :111
        fsep.initialize(lm, num_var + 1, num_constr + 1, epi_d, self.params.f_tol, self.params.cut_coef_rng)   # :112
        fsep.set_bounds(self.l_constr + [0.0], self.u_constr + [0.0])
        pt = np.zeros(num_var + 1)                             # :113
        fsep.xstar = pt
        lin_rows = [i for i in range(num_constr) if d.isconstrlinear(i)]
        self.nlconstr_ixs = [i for i in range(num_constr) if not d.isconstrlinear(i)]   # :115-122
        if lin_rows:
            b = fsep.handle.gencut_rows(pt, np.asarray(lin_rows, np.int64), round_coefs=False)   # gencut + _addcut, no rounding (:117-118)
            self._addbatch(b)
        self.objislinear = d.isobjlinear()                     # :125
        if self.objislinear:
            if self.params.log_level > 0: print("objective is linear")
            cut = fsep.gencut(pt, (0, 0), num_constr)          # :129
            assert cut.vars[-1] == num_var                     # :130
            lm.setobjective(sense, cut.vars[:-1], cut.coeffs[:-1], cut.constant)   # :131-133
        else:
            if self.params.log_level > 0: print("objective is nonlinear")
            y = lm.addvar()                                    # :137
            self.num_var += 1
            lm.setobjective(sense, [y], [1.0])                 # :139
            l_obj, u_obj = (0.0, math.inf) if sense == "Max" else (-math.inf, 0.0)   # :144
            self.l_constr.append(l_obj); self.u_constr.append(u_obj)
            self.num_constr += 1
            self.nlconstr_ixs.append(self.num_constr - 1)      # :148
            if np.any(np.isnan(vertex)):
                warnings.warn("Problem variables insufficiently bounded!")   # :156-157
            else:
                vertex = np.append(vertex, d.eval_f(vertex))   # :159
                fsep.set_bounds(self.l_constr, self.u_constr)
                b = fsep.handle.gencut_rows(vertex, np.array([self.num_constr - 1], np.int64), round_coefs=True)   # :160-163
                self._addbatch(b)
            d = EpigraphNLPEvaluator(d, num_var + 1, num_constr + 1)   # :166
        self.num_nlconstr = len(self.nlconstr_ixs)
        fsep.handle.close()
        sep = self.params.separator                            # :171-172
        if isinstance(sep, KatanaGPUSeparator):
            sep.initialize(lm, self.num_var, self.num_constr, d, self.params.f_tol, self.params.cut_coef_rng)
        else:
            sep.initialize(lm, self.num_var, self.num_constr, d)
        sep.set_bounds(self.l_constr, self.u_constr)
        self.oracle = d

    def _separate_round(self, x):
        """Loop body src/model.jl:268-283.  Returns (allsat, cuts_added)."""
        sep = self.params.separator
        if hasattr(sep, "separate"):                           # batched device round
            t0 = time.perf_counter()
            batch = sep.separate(x)
            t1 = time.perf_counter()
            self._addbatch(batch)
            t2 = time.perf_counter()
            self.round_log.append({"separate_s": t1 - t0, "addconstr_s": t2 - t1, "cuts": batch.n_cuts})
            return batch.n_cuts == 0 and self.status != "Error", batch.n_cuts
        sep.precompute(x)                                      # reference per-row path, any AbstractKatanaSeparator
        allsat, cuts = True, 0
        for i in self.nlconstr_ixs:
            sat = sep.isconstrsat(i, self.l_constr[i], self.u_constr[i], self.params.f_tol)
            if not sat:
                cut = sep.gencut(x, (self.l_constr[i], self.u_constr[i]), i)
                round_coefs(cut, self.params.cut_coef_rng)
                self._addcut(cut, self.l_constr[i], self.u_constr[i])
                if self.status == "Error":
                    return False, cuts
                cuts += 1
            allsat &= sat
        return allsat, cuts

    # boundroutine(m, ray) -- src/model.jl:175-197
    def boundroutine(self, ray):
        sep = self.params.separator
        if hasattr(sep, "separate_ladder") and getattr(sep, "has_ladder", False):
            # the whole search in one library call: the points are evaluated on the device in batches, the cuts are made at the
            # first point that violates a row -- what the loop below does with up to 1022 sequential rounds
            n_hit, batch = sep.separate_ladder(np.asarray(ray, np.float64))
            self._addbatch(batch)
            self.round_log.append({"ladder_hit": n_hit, "cuts": batch.n_cuts})
            return
        for n in range(2, 1024):
            x = (2.0 ** n) * ray
            allsat, _ = self._separate_round(x)
            if self.status == "Error":
                return
            if not allsat:
                break

    # MathProgBase.optimize!(m) -- src/model.jl:219-319
    def optimize(self):
        start = time.time()
        lm = self.linear_model
        status = lm.solve()                                    # :228
        if status == "Unbounded":
            warnings.warn("Automatically bounding unbounded LP")
        i = 0
        while status == "Unbounded" and i < self.num_var:      # :235-242
            ray = lm.getunboundedray()
            if self.params.log_level > 0: print(f"Unbounded ray along: {ray}")
            self.boundroutine(ray)
            if self.status == "Error": return self.status
            status = lm.solve()
            i += 1
        if status == "Unbounded":                              # :244-247
            warnings.warn("Katana could not resolve unbounded LP")
            self.status = status
            return self.status
        if self.params.log_level > 0: self.print_header()
        allsat = False
        self.in_loop = True                                    # cuts added from here on may be purged (cut_purge_age)
        self.purge_enabled, self.purge_bound = True, -math.inf
        cuts_lastprnt, max_viol, obj_prev = 0, 0, math.inf
        while not allsat and self.iter < self.params.iter_cap:     # :257
            self.iter += 1
            t0 = time.perf_counter()
            status = lm.solve()                                # :259
            if status == "Unbounded" and self.params.cut_purge_age > 0 and hasattr(lm, "restore_purged"):
                back = lm.restore_purged()                     # a purged cut was bounding the LP: all of them return for good,
                if back:                                       # and nothing is purged any more
                    self.cuts_purged -= back
                    self.purge_enabled = False
                    status = lm.solve()
            lp_s = time.perf_counter() - t0
            if status != "Optimal":                            # :261-263
                self.status = status
                return self.status
            xstar = lm.getsolution()                           # :265
            if self.params.cut_purge_age > 0 and self.purge_enabled and hasattr(lm, "purge_slack_rows"):
                # the LP bound only tightens while cuts are only added; a purge may loosen it.  The next purge waits until the bound
                # has passed the value it had at the last one: every purge is separated by strict progress, so the loop cannot cycle
                bound = lm.getobjval() if lm.sense == "Min" else -lm.getobjval()
                if bound > self.purge_bound + 1e-9 * max(1.0, abs(bound)):
                    n = lm.purge_slack_rows(self.params.cut_purge_age, 1e-7)
                    self.cuts_purged += n
                    if n: self.purge_bound = bound
            if self.features["VisData"]: self.lp_sols.append(xstar)
            allsat, cuts_viol = self._separate_round(xstar)    # :268-283
            if self.round_log: self.round_log[-1]["lp_s"] = lp_s
            if self.status == "Error": return self.status      # :278
            max_viol = max(max_viol, cuts_viol)                # :284
            cuts_lastprnt += cuts_viol
            obj = lm.getobjval()                               # :287
            # IEEE division as in Julia: 0/0 = NaN (NaN <= obj_eps is false: the loop goes on), x/0 = Inf
            with np.errstate(divide="ignore", invalid="ignore"):
                obj_delta = float(np.abs((np.float64(obj_prev) - np.float64(obj)) / np.float64(obj)))
            obj_prev = obj
            if self.params.log_level > 0:                      # :291-303
                r = self.iter % self.params.log_level
                if r == 0:
                    if self.iter % (self.params.log_level * 50) == 0: self.print_header()
                    self.print_stats(self.params.log_level, cuts_lastprnt, max_viol)
                    cuts_lastprnt = 0; max_viol = 0
                elif allsat:
                    self.print_stats(r, cuts_lastprnt, max_viol)
                elif obj_delta <= self.params.obj_eps:
                    self.print_stats(self.iter, cuts_lastprnt, max_viol)
            if obj_delta <= self.params.obj_eps:               # :306-308
                break
        self.soltime = time.time() - start                     # :311
        if self.iter >= self.params.iter_cap:                  # :313-315
            status = "UserLimit"
        self.status = status
        return self.status

    @staticmethod
    def print_header():                                        # src/model.jl:209-211
        print("%-10s %-15s %-15s %-20s %-20s %-15s" % ("Iteration", "Total cuts", "Cuts added", "Max constr. viol.", "Avg constr. viol.", "Current cuts"))

    def print_stats(self, iter_lastprnt, cuts_lastprnt, max_viol):   # src/model.jl:213-217
        avg = cuts_lastprnt / (iter_lastprnt * max(self.num_nlconstr, 1))
        print("%-10d %-15d %-15d %-20d %-20.2f %-15d" % (self.iter, self.numcuts, cuts_lastprnt, max_viol, avg, self.numcuts))

    # getters -- src/model.jl:326-343
    def numiters(self): return self.iter
    def getnumcuts(self): return self.numcuts
    def setwarmstart(self, x): return [0.0] * len(x)
    def getstatus(self): return self.status
    def getobjval(self): return self.linear_model.getobjval()
    def getsolution(self): return self.linear_model.getsolution()
    def getsolvetime(self): return self.soltime
