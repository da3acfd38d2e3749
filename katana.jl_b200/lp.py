"""Host LP master stand-in.

In the reference the LP master is a JuMP model on an external MathProgBase LP solver
(GLPK in test/runtests.jl:24); it stays on the host and is OUT OF SCOPE of the B200 path
(SURVEY.md section 2, row 9).  No GLPK/Clp exists in this image, so the Python mirror uses
scipy's bundled HiGHS behind the handful of calls src/model.jl makes on the LP:
solve, getsolution, getobjval, getunboundedray, addconstraint (src/model.jl:76,89-96,228-265).
"""
import numpy as np
from scipy.optimize import linprog
from scipy.sparse import csr_matrix


class HighsLP:
    def __init__(self):
        self.lb, self.ub = [], []
        self.c = np.zeros(0); self.c0 = 0.0; self.sense = "Min"
        self.rows = []          # (cols int array, vals float array, lo, hi)
        self.managed = []       # per row: -1 = permanent (model rows, vertex / bounding cuts); >= 0: a loop cut, solves in a row it has been slack
        self.x = None; self.objval = np.nan; self.status = "None"

    # --- model building -------------------------------------------------------------------
    def addvar(self, lb=-np.inf, ub=np.inf):
        self.lb.append(float(lb)); self.ub.append(float(ub))
        self.c = np.append(self.c, 0.0)
        return len(self.lb) - 1

    @property
    def numvar(self):
        return len(self.lb)

    def setobjective(self, sense, cols, coefs, const=0.0):
        self.sense = sense
        self.c = np.zeros(self.numvar)
        np.add.at(self.c, np.asarray(cols, dtype=np.int64), np.asarray(coefs, dtype=np.float64))
        self.c0 = float(const)

    def addconstr(self, cols, vals, lo, hi):
        """MathProgBase.addconstr!(m, varidx, coef, lb, ub): one LP row lo <= a.x <= hi."""
        self.rows.append((np.asarray(cols, np.int64), np.asarray(vals, np.float64), float(lo), float(hi)))
        self.managed.append(-1)

    def addconstrs_csr(self, row_ptr, col, val, lo, hi, managed=False, skip=None):
        """Batched hand-off of a CutBatch (the fast path that bypasses AffExpr objects).  managed: the rows may be purged later
        (purge_slack_rows); skip: boolean mask of cuts NOT to add (duplicate filter)."""
        for c in range(len(lo)):
            if skip is not None and skip[c]:
                continue
            s, e = row_ptr[c], row_ptr[c + 1]
            self.rows.append((col[s:e].astype(np.int64), val[s:e].copy(), float(lo[c]), float(hi[c])))
            self.managed.append(0 if managed else -1)

    def purge_slack_rows(self, age, tol):
        """Cut management (an extension: the reference keeps every cut, src/model.jl:215).  After a solve: a managed row whose
        slack at x* exceeds tol * max(1, |bound|) has been inactive one more solve; rows inactive `age` solves in a row are
        removed.  Returns the number of rows removed."""
        if self.x is None:
            return 0
        keep_rows, keep_tag, removed = [], [], 0
        if not hasattr(self, "purged"):
            self.purged = []    # every row removed so far: put back (for good) by restore_purged if the LP loses its bound
        for (cols, vals, lo, hi), tag in zip(self.rows, self.managed):
            if tag >= 0:
                a = float(vals @ self.x[cols])
                slack = min(hi - a if np.isfinite(hi) else np.inf, a - lo if np.isfinite(lo) else np.inf)
                ref = max(1.0, abs(hi) if np.isfinite(hi) else 0.0, abs(lo) if np.isfinite(lo) else 0.0)
                tag = tag + 1 if slack > tol * ref else 0
                if tag >= age:
                    removed += 1
                    self.purged.append((cols, vals, lo, hi))
                    continue
            keep_rows.append((cols, vals, lo, hi)); keep_tag.append(tag)
        self.rows, self.managed = keep_rows, keep_tag
        return removed

    def restore_purged(self):
        """Puts every purged row back as a permanent row; returns how many."""
        rows = getattr(self, "purged", [])
        for r in rows:
            self.rows.append(r); self.managed.append(-1)
        self.purged = []
        return len(rows)

    # --- solving ---------------------------------------------------------------------------
    def _matrices(self):
        n = self.numvar
        ub_rows, eq_rows = [], []
        for cols, vals, lo, hi in self.rows:
            if not np.all(np.isfinite(vals)):
                return None
            if np.isnan(lo) or np.isnan(hi):
                # a cut taken where g is undefined (NaN value, finite gradient: the reference adds it, src/model.jl:69-76)
                # has NaN bounds and constrains nothing; this stand-in drops it instead of guessing what GLPK does with NaN
                continue
            if lo == hi:
                eq_rows.append((cols, vals, lo))
                continue
            if np.isfinite(hi): ub_rows.append((cols, vals, hi))
            if np.isfinite(lo): ub_rows.append((cols, -vals, -lo))

        def build(rs):
            if not rs:
                return None, None
            indptr = np.zeros(len(rs) + 1, np.int64)
            for i, r in enumerate(rs): indptr[i + 1] = indptr[i] + len(r[0])
            A = csr_matrix((np.concatenate([r[1] for r in rs]), np.concatenate([r[0] for r in rs]), indptr), shape=(len(rs), n))
            return A, np.array([r[2] for r in rs])
        return build(ub_rows), build(eq_rows)

    def solve(self):
        n = self.numvar
        mats = self._matrices()
        if mats is None:
            self.status = "Error"; return self.status
        (A_ub, b_ub), (A_eq, b_eq) = mats
        c = self.c if self.sense == "Min" else -self.c
        bounds = [(l if np.isfinite(l) else None, u if np.isfinite(u) else None) for l, u in zip(self.lb, self.ub)]
        res = linprog(c, A_ub=A_ub, b_ub=b_ub, A_eq=A_eq, b_eq=b_eq, bounds=bounds, method="highs-ds")
        st = res.status
        if st in (2, 3, 4) or (st == 0 and res.x is None):
            # HiGHS may answer "infeasible or unbounded": settle it with a feasibility LP
            feas = linprog(np.zeros(n), A_ub=A_ub, b_ub=b_ub, A_eq=A_eq, b_eq=b_eq, bounds=bounds, method="highs-ds")
            self.status = "Unbounded" if feas.status == 0 else "Infeasible"
            self.x = None
            return self.status
        if st != 0:
            self.status = "Error"; return self.status
        self.x = np.asarray(res.x, np.float64)
        self.objval = float(self.c @ self.x + self.c0)
        self.status = "Optimal"
        return self.status

    def getsolution(self):
        return self.x.copy()

    def getobjval(self):
        return self.objval

    def getunboundedray(self):
        """An improving recession direction of the LP (MathProgBase.getunboundedray, src/model.jl:236)."""
        n = self.numvar
        mats = self._matrices()
        (A_ub, _), (A_eq, _) = mats
        c = self.c if self.sense == "Min" else -self.c
        bounds = [(0.0 if np.isfinite(l) else -1.0, 0.0 if np.isfinite(u) else 1.0) for l, u in zip(self.lb, self.ub)]
        res = linprog(c, A_ub=A_ub, b_ub=None if A_ub is None else np.zeros(A_ub.shape[0]),
                      A_eq=A_eq, b_eq=None if A_eq is None else np.zeros(A_eq.shape[0]), bounds=bounds, method="highs-ds")
        if res.status != 0 or res.x is None:
            return np.zeros(n)
        d = np.asarray(res.x, np.float64)
        # A simplex code reports an extreme ray with O(1) movement in the structural variables; normalise the same
        # way so boundroutine's probe points 2^n * ray (src/model.jl:181-182) leave the origin along x, not only along
        # the objective variable.
        mask = c == 0
        scale = np.max(np.abs(d[mask])) if mask.any() else 0.0
        return d / scale if scale > 1e-12 else d
