"""The LP / QP / QCQP route: reference src/solver.jl:46

    MathProgBase.LinearQuadraticModel(s::KatanaSolver) = MathProgBase.NonlinearToLPQPBridge(MathProgBase.NonlinearModel(s))

A JuMP model without @NL* macros reaches Katana through MathProgBase's LinearQuadratic interface (loadproblem!(m, A, l, u, c,
lb, ub, sense), setquadobj!, addquadconstr!); the bridge (third-party: MathProgBase 0.6-0.7, SolverInterface/nonlinear_to_lpqp.jl,
not vendored in the reference; conventions below are MathProgBase's documented ones) turns it into a nonlinear model whose
evaluator, LPQPEvaluator, offers [:Grad, :Jac, :Hess] and NO :ExprGraph.  The B200 path needs expressions, so this mirror -- and
the Julia shim, julia/gpu_separator.jl `lpqp_rows` -- CAPTURES A AND Q DIRECTLY and hands out the
rows as expression graphs of the same shape JuMP prints for quadratic expressions: +(q*x_i*x_j ..., a*x_k ...).

Difference from the reference on this route, stated: LPQPEvaluator's Jacobian has TWO entries per quadratic term (d/dx_i and
d/dx_j, also when i == j) and keeps the COO order of (A, Q); the reference's cut therefore carries duplicate columns which JuMP
merges when it hands the row to the LP solver.  Here the columns of a cut are sorted and unique (the merged row).  Same violated
rows (g is the same polynomial), coefficients equal after the merge up to the summation order (<= 1e-12 relative).
"""
import math

import numpy as np

from . import expr as E


class LPQPEvaluator:
    """Rows 0..numLin-1: A x in [lb, ub]; then one row per addquadconstr!.  Objective c'x + 0.5 x'Qx, Q given by one triangle."""

    def __init__(self, num_var, A_rows, c, qobj, qcons):
        self.num_var, self.A_rows, self.c, self.qobj, self.qcons = num_var, A_rows, np.asarray(c, np.float64), qobj, qcons
        self.num_lin = len(A_rows)

    def features_available(self): return ["ExprGraph"]           # the reference's offers [:Grad, :Jac, :Hess]; see the module docstring

    def initialize(self, requested_features):
        for f in requested_features:
            if f not in self.features_available():
                raise ValueError(f"Unsupported feature {f}")

    def isobjlinear(self): return len(self.qobj[2]) == 0
    def isconstrlinear(self, i): return i < self.num_lin
    def isconstrdense(self, i): return False

    def _obj_form(self):
        q = E.QuadForm(0.0, {j: v for j, v in enumerate(self.c) if v != 0.0})
        for i, j, v in zip(*self.qobj):                           # 0.5 x'Qx with the given triangle mirrored: v x_i x_j off the diagonal, v/2 x_i^2 on it
            key = (min(i, j), max(i, j))
            q.quad[key] = q.quad.get(key, 0.0) + (0.5 * v if i == j else v)
        return q

    def obj_expr(self): return self._obj_form().to_expr()

    def _row_form(self, i):
        if i < self.num_lin:
            cols, vals = self.A_rows[i]
            q = E.QuadForm()
            for j, v in zip(cols, vals): q.lin[int(j)] = q.lin.get(int(j), 0.0) + float(v)
            return q
        linidx, linval, qrow, qcol, qval = self.qcons[i - self.num_lin]
        q = E.QuadForm()
        for j, v in zip(linidx, linval): q.lin[int(j)] = q.lin.get(int(j), 0.0) + float(v)
        for a, b, v in zip(qrow, qcol, qval):                     # addquadconstr!: sum quadval * x_row * x_col, every entry as given
            key = (min(int(a), int(b)), max(int(a), int(b)))
            q.quad[key] = q.quad.get(key, 0.0) + float(v)
        return q

    def constr_expr(self, i): return self._row_form(i).to_expr()
    def eval_f(self, x): return E.evaluate(self.obj_expr(), x)


class NonlinearToLPQPBridge:
    """MathProgBase.NonlinearToLPQPBridge: collects the LinearQuadratic calls, loads the wrapped nonlinear model at optimize!."""

    def __init__(self, nlpmodel):
        self.nlpmodel = nlpmodel                                  # src/util.jl:4 reads this field
        self.qobj = ([], [], [])
        self.qcons, self.qbounds = [], []

    def loadproblem(self, A_rows, l, u, c, lb, ub, sense):
        self.A_rows, self.l, self.u, self.c = list(A_rows), np.asarray(l, float), np.asarray(u, float), np.asarray(c, float)
        self.lb, self.ub, self.sense = list(lb), list(ub), sense

    def setquadobj(self, rowidx, colidx, quadval):
        self.qobj = (list(rowidx), list(colidx), list(quadval))

    def addquadconstr(self, linearidx, linearval, quadrowidx, quadcolidx, quadval, sense, rhs):
        self.qcons.append((list(linearidx), list(linearval), list(quadrowidx), list(quadcolidx), list(quadval)))
        self.qbounds.append({"<": (-math.inf, rhs), ">": (rhs, math.inf), "=": (rhs, rhs)}[sense])

    def optimize(self):
        n = len(self.l)
        d = LPQPEvaluator(n, self.A_rows, self.c, self.qobj, self.qcons)
        lbs = np.array(self.lb + [b[0] for b in self.qbounds], dtype=float)
        ubs = np.array(self.ub + [b[1] for b in self.qbounds], dtype=float)
        self.nlpmodel.loadproblem(n, len(lbs), self.l, self.u, lbs, ubs, self.sense, d)
        return self.nlpmodel.optimize()

    def getobjval(self): return self.nlpmodel.getobjval()
    def getsolution(self): return self.nlpmodel.getsolution()
    def status(self): return self.nlpmodel.getstatus()
