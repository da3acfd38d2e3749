"""Solver facade and a JuMP-like modelling front end.

`KatanaSolver` mirrors reference src/solver.jl:6-46 (same keyword names and defaults).  `Model`
stands in for the JuMP layer above it (out of scope of the B200 path; it exists so the
reference's test problems, test/*.jl, can be written down and solved through the same calls).
"""
import math

import numpy as np

from . import expr as E
from .lp import HighsLP
from .lpqp import NonlinearToLPQPBridge
from .model import KatanaModelParams, KatanaNonlinearModel
from .nlpeval import ExprNLPEvaluator
from .separators import KatanaGPUSeparator


class KatanaSolver:                                            # src/solver.jl:6-10,34-43
    def __init__(self, lp_solver=HighsLP, separator=None, features=(), f_tol=1e-6, cut_coef_rng=1e9,
                 log_level=10, iter_cap=10000, obj_eps=-1.0, cut_purge_age=0, cut_filter_duplicates=False):
        self.lp_solver = lp_solver
        self.features = list(features)
        self.model_params = KatanaModelParams(f_tol, iter_cap, log_level, cut_coef_rng, obj_eps,
                                              separator if separator is not None else KatanaGPUSeparator(),
                                              cut_purge_age=cut_purge_age, cut_filter_duplicates=cut_filter_duplicates)


def NonlinearModel(s):                                         # src/model.jl:63-65
    return KatanaNonlinearModel(s.lp_solver, s.features, s.model_params)


def LinearQuadraticModel(s):                                   # src/solver.jl:46
    return NonlinearToLPQPBridge(NonlinearModel(s))


class Variable(E.Node):
    pass


class Model:
    """Subset of JuMP used by the reference's tests: @variable, @objective, @NLobjective,
    @constraint (affine / quadratic, normalised like JuMP: constants move to the bounds), @NLconstraint."""

    def __init__(self, solver=None):
        self.solver = solver
        self.lb, self.ub = [], []
        self.lin, self.quad, self.nl = [], [], []              # (expr Node, lb, ub)
        self.sense, self.obj, self.obj_lin = "Min", E.const(0.0), True
        self.obj_q = E.QuadForm()                              # polynomial objective (None after @NLobjective)
        self.internal = None

    def variable(self, lb=-math.inf, ub=math.inf, start=None):
        self.lb.append(float(lb)); self.ub.append(float(ub))
        return E.var(len(self.lb) - 1)

    def variables(self, n, lb=-math.inf, ub=math.inf):
        return [self.variable(lb, ub) for _ in range(n)]

    def objective(self, sense, expr):
        q = E.to_quadform(expr)
        self.sense, self.obj, self.obj_lin, self.obj_q = sense, q.to_expr(with_const=True), q.is_affine, q

    def nlobjective(self, sense, expr):
        self.sense, self.obj, self.obj_lin, self.obj_q = sense, E.wrap(expr), False, None

    @staticmethod
    def _bounds(sense, rhs):
        return {"<=": (-math.inf, rhs), ">=": (rhs, math.inf), "==": (rhs, rhs)}[sense]

    def constraint(self, lhs, sense, rhs):
        """@constraint(m, lhs sense rhs) with affine / quadratic sides."""
        q = E.to_quadform(E.wrap(lhs) - E.wrap(rhs))           # JuMP moves everything left, the constant right
        lo, hi = self._bounds(sense, -q.c)
        (self.lin if q.is_affine else self.quad).append((q.to_expr(), lo, hi, q))

    def nlconstraint(self, lhs, sense, rhs):
        """@NLconstraint(m, lhs sense rhs): JuMP stores lhs - rhs against 0 (rhs == 0 keeps lhs as is)."""
        rhs_n = E.wrap(rhs)
        body = E.wrap(lhs) if (rhs_n.op == E.OP_CONST and rhs_n.value == 0.0) else E.Node(E.OP_SUB, (E.wrap(lhs), rhs_n))
        lo, hi = self._bounds(sense, 0.0)
        self.nl.append((body, lo, hi))

    def _solve_lpqp(self):
        """No @NL* in the model: JuMP builds the solver's LinearQuadraticModel (src/solver.jl:46) and speaks MathProgBase's
        LinearQuadratic interface to it: loadproblem!(A, l, u, c, lb, ub, sense), setquadobj!, addquadconstr!."""
        b = self.bridge = LinearQuadraticModel(self.solver)
        self.internal = b.nlpmodel
        n, q0 = len(self.lb), self.obj_q
        c = np.zeros(n)
        for j, v in q0.lin.items(): c[j] = v
        A_rows = [(np.array(sorted(q.lin), np.int64), np.array([q.lin[j] for j in sorted(q.lin)])) for _, _, _, q in self.lin]
        b.loadproblem(A_rows, self.lb, self.ub, c, [r[1] for r in self.lin], [r[2] for r in self.lin], self.sense)
        if not q0.is_affine:                                   # 0.5 x'Qx, one triangle: the diagonal entry of v x_i^2 is 2 v
            keys = [k for k in sorted(q0.quad) if q0.quad[k] != 0.0]
            b.setquadobj([i for i, _ in keys], [j for _, j in keys], [q0.quad[k] * (2.0 if k[0] == k[1] else 1.0) for k in keys])
        for _, lo, hi, q in self.quad:                         # JuMP: one-sided or equality quadratic constraints only
            sense, rhs = ("<", hi) if lo == -math.inf else (">", lo) if hi == math.inf else ("=", hi)
            if sense == "=" and lo != hi: raise ValueError("two-sided quadratic constraints are not part of the LinearQuadratic interface")
            keys, lk = [k for k in sorted(q.quad) if q.quad[k] != 0.0], sorted(q.lin)
            b.addquadconstr(lk, [q.lin[j] for j in lk], [i for i, _ in keys], [j for _, j in keys], [q.quad[k] for k in keys], sense, rhs)
        self.obj_const = q0.c                                  # the LP form has no objective constant: JuMP adds it back
        return b.optimize()

    def solve(self):
        self.obj_const = 0.0
        if not self.nl and self.obj_q is not None:
            return self._solve_lpqp()
        rows = [(r[0], True) for r in self.lin] + [(r[0], False) for r in self.quad] + [(e, False) for e, _, _ in self.nl]
        allc = self.lin + self.quad + self.nl
        d = ExprNLPEvaluator(len(self.lb), rows, self.obj, self.obj_lin)
        m = self.internal = NonlinearModel(self.solver)
        m.loadproblem(len(self.lb), len(allc), np.array(self.lb), np.array(self.ub),
                      np.array([c[1] for c in allc], dtype=float), np.array([c[2] for c in allc], dtype=float), self.sense, d)
        # the reference proceeds into optimize! whatever loadproblem! left in m.status (src/solver.jl:34-43 -> src/model.jl:219):
        # optimize! ends by overwriting it
        return m.optimize()

    def getobjectivevalue(self):
        return self.internal.getobjval() + self.obj_const

    def getvalue(self, v):
        return float(self.internal.getsolution()[v.index])


def getKatanaModel(m):                                         # src/util.jl:3-5
    if isinstance(m, NonlinearToLPQPBridge): return m.nlpmodel
    return m.internal if isinstance(m, Model) else m
