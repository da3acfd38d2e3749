"""Host LP master stand-in.

In the reference the LP master is a JuMP model on an external MathProgBase LP solver
(GLPK in test/runtests.jl:24); it stays on the host and is OUT OF SCOPE of the B200 path
(SURVEY.md section 2, row 9).  No GLPK/Clp exists in this image, so the Python mirror uses
scipy's bundled HiGHS behind the handful of calls src/model.jl makes on the LP:
solve, getsolution, getobjval, getunboundedray, addconstraint (src/model.jl:76,89-96,228-265).

Batched hand-off: a CutBatch arrives as CSR and is KEPT as one CSR block (`addconstrs_csr`: three
array copies, no per-cut Python work); the constraint matrix of a solve is assembled from the blocks with sparse row selections.
The row order of the assembled LP is the insertion order (a two-sided row contributes its upper row, then its negated lower row),
whether a row arrived alone or inside a block.
"""
import numpy as np
from scipy.optimize import linprog
from scipy.sparse import csr_matrix, diags, vstack


class _Block:
    """Consecutive LP rows lo <= A x <= hi as one CSR block.  age: None = permanent rows; else per row the number of consecutive
    solves it has been slack (cut management, purge_slack_rows)."""

    def __init__(self, indptr, col, val, lo, hi, managed):
        self.indptr = np.asarray(indptr, np.int64) - int(indptr[0])
        self.col, self.val = np.asarray(col, np.int64), np.asarray(val, np.float64)
        self.lo, self.hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
        self.age = np.zeros(len(self.lo), np.int32) if managed else None

    def __len__(self):
        return len(self.lo)

    def matrix(self, n):
        return csr_matrix((self.val, self.col, self.indptr), shape=(len(self.lo), n))

    def extend(self, o):
        """Appends the rows of block `o` (same kind: both permanent or both managed)."""
        self.indptr = np.concatenate([self.indptr, o.indptr[1:] + self.indptr[-1]])
        self.col, self.val = np.concatenate([self.col, o.col]), np.concatenate([self.val, o.val])
        self.lo, self.hi = np.concatenate([self.lo, o.lo]), np.concatenate([self.hi, o.hi])
        if self.age is not None:
            self.age = np.concatenate([self.age, o.age])

    def take(self, keep):
        """The block restricted to the rows of the boolean mask `keep`."""
        A = self.matrix(int(self.col.max()) + 1 if len(self.col) else 1)[np.flatnonzero(keep)]
        b = _Block(A.indptr, A.indices, A.data, self.lo[keep], self.hi[keep], False)
        b.age = None if self.age is None else self.age[keep]
        return b


class HighsLP:
    def __init__(self):
        self.lb, self.ub = [], []
        self.c = np.zeros(0); self.c0 = 0.0; self.sense = "Min"
        self.blocks = []        # _Block objects in insertion order (a single row is a block of one row)
        self.purged = []        # blocks of rows removed by purge_slack_rows: put back (for good) by restore_purged if the LP loses its bound
        self.x = None; self.objval = np.nan; self.status = "None"

    # --- model building -------------------------------------------------------------------
    def addvar(self, lb=-np.inf, ub=np.inf):
        self.lb.append(float(lb)); self.ub.append(float(ub))
        self.c = np.append(self.c, 0.0)
        return len(self.lb) - 1

    @property
    def numvar(self):
        return len(self.lb)

    @property
    def rows(self):
        """The LP rows as (cols, vals, lo, hi) tuples in LP order (inspection and tests; the solver path never builds this list)."""
        out = []
        for b in self.blocks:
            for i in range(len(b)):
                s, e = b.indptr[i], b.indptr[i + 1]
                out.append((b.col[s:e], b.val[s:e], float(b.lo[i]), float(b.hi[i])))
        return out

    @property
    def numrows(self):
        return sum(len(b) for b in self.blocks)

    def setobjective(self, sense, cols, coefs, const=0.0):
        self.sense = sense
        self.c = np.zeros(self.numvar)
        np.add.at(self.c, np.asarray(cols, dtype=np.int64), np.asarray(coefs, dtype=np.float64))
        self.c0 = float(const)

    def addconstr(self, cols, vals, lo, hi):
        """MathProgBase.addconstr!(m, varidx, coef, lb, ub): one LP row lo <= a.x <= hi."""
        cols = np.asarray(cols, np.int64)
        self._append(_Block([0, len(cols)], cols, np.array(vals, np.float64), [float(lo)], [float(hi)], False))

    def _append(self, b):
        """Rows in insertion order; a small block joins the block before it when both are of one kind (the ECP loop of a small model
        adds one or two cuts per round: hundreds of one-row blocks would make every solve assemble hundreds of matrices)."""
        last = self.blocks[-1] if self.blocks else None
        if last is not None and len(b) <= 4096 and len(last.val) <= 1 << 20 and (last.age is None) == (b.age is None):
            last.extend(b)
        else:
            self.blocks.append(b)

    def addconstrs_csr(self, row_ptr, col, val, lo, hi, managed=False, skip=None):
        """Batched hand-off of a CutBatch: the CSR arrays are copied once (they may be views of the library's pinned buffer) and kept
        as a block.  managed: the rows may be purged later (purge_slack_rows); skip: boolean mask of cuts NOT to add (duplicate filter)."""
        n = len(lo)
        if n == 0:
            return
        s, e = int(row_ptr[0]), int(row_ptr[n])
        b = _Block(np.array(row_ptr[:n + 1], np.int64), np.array(col[s:e], np.int64), np.array(val[s:e], np.float64),
                   np.array(lo, np.float64), np.array(hi, np.float64), managed)
        if skip is not None and np.any(skip):
            keep = ~np.asarray(skip, bool)
            if not keep.any():
                return
            b = b.take(keep)
        self._append(b)

    def purge_slack_rows(self, age, tol):
        """Cut management (an extension: the reference keeps every cut, src/model.jl:215).  After a solve: a managed row whose
        slack at x* exceeds tol * max(1, |bound|) has been inactive one more solve; rows inactive `age` solves in a row are
        removed.  Returns the number of rows removed."""
        if self.x is None:
            return 0
        removed, kept = 0, []
        n = self.numvar
        for b in self.blocks:
            if b.age is None or len(b) == 0:
                kept.append(b); continue
            a = b.matrix(n) @ self.x
            with np.errstate(invalid="ignore"):
                slack = np.minimum(np.where(np.isfinite(b.hi), b.hi - a, np.inf), np.where(np.isfinite(b.lo), a - b.lo, np.inf))
            ref = np.maximum(1.0, np.maximum(np.where(np.isfinite(b.hi), np.abs(b.hi), 0.0), np.where(np.isfinite(b.lo), np.abs(b.lo), 0.0)))
            b.age = np.where(slack > tol * ref, b.age + 1, 0).astype(np.int32)
            out = b.age >= age
            if out.any():
                removed += int(out.sum())
                gone = b.take(out); gone.age = None
                self.purged.append(gone)
                if (~out).any():
                    kept.append(b.take(~out))
            else:
                kept.append(b)
        self.blocks = kept
        return removed

    def restore_purged(self):
        """Puts every purged row back as a permanent row; returns how many."""
        k = sum(len(b) for b in self.purged)
        self.blocks.extend(self.purged)
        self.purged = []
        return k

    # --- solving ---------------------------------------------------------------------------
    def _matrices(self):
        """(A_ub, b_ub), (A_eq, b_eq) in LP order, or None when a coefficient is not finite."""
        n = self.numvar
        ub_parts, ub_rhs, eq_parts, eq_rhs = [], [], [], []
        for b in self.blocks:
            if len(b) == 0:
                continue
            if not np.all(np.isfinite(b.val)):
                return None
            # a cut taken where g is undefined (NaN value, finite gradient: the reference adds it, src/model.jl:69-76) has NaN bounds
            # and constrains nothing; this stand-in drops it instead of guessing what GLPK does with NaN
            ok = ~(np.isnan(b.lo) | np.isnan(b.hi))
            eq = ok & (b.lo == b.hi)
            up = ok & ~eq & np.isfinite(b.hi)
            dn = ok & ~eq & np.isfinite(b.lo)
            A = None
            if eq.any():
                A = b.matrix(n)
                idx = np.flatnonzero(eq)
                eq_parts.append(A[idx]); eq_rhs.append(b.lo[idx])
            if up.any() or dn.any():
                A = b.matrix(n) if A is None else A
                iu, il = np.flatnonzero(up), np.flatnonzero(dn)
                rows = np.concatenate([iu, il]); sign = np.concatenate([np.ones(len(iu)), -np.ones(len(il))])
                rhs = np.concatenate([b.hi[iu], -b.lo[il]])
                order = np.argsort(2 * rows + (sign < 0), kind="stable")       # a row's upper row first, then its negated lower row
                rows, sign, rhs = rows[order], sign[order], rhs[order]
                M = A[rows]
                if (sign < 0).any():
                    M = diags(sign) @ M
                ub_parts.append(M.tocsr()); ub_rhs.append(rhs)

        def build(parts, rhs):
            if not parts:
                return None, None
            return (parts[0] if len(parts) == 1 else vstack(parts, format="csr")), np.concatenate(rhs)
        return build(ub_parts, ub_rhs), build(eq_parts, eq_rhs)

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
