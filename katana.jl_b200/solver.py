"""Solver facade and a JuMP-like modelling front end.

`KatanaSolver` mirrors reference src/solver.jl:6-46 (same keyword names and defaults).  `Model`
stands in for the JuMP layer above it (out of scope of the B200 path; it exists so the
reference's test problems, test/*.jl, can be written down and solved through the same calls).
"""
import math

import numpy as np

from . import expr as E
from .lp import HighsLP
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


def LinearQuadraticModel(s):                                   # src/solver.jl:46 (the LPQP bridge is the identity here:
    return NonlinearModel(s)                                   # quadratic rows already arrive as expression graphs)


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
        self.internal = None

    def variable(self, lb=-math.inf, ub=math.inf, start=None):
        self.lb.append(float(lb)); self.ub.append(float(ub))
        return E.var(len(self.lb) - 1)

    def variables(self, n, lb=-math.inf, ub=math.inf):
        return [self.variable(lb, ub) for _ in range(n)]

    def objective(self, sense, expr):
        q = E.to_quadform(expr)
        self.sense, self.obj, self.obj_lin = sense, q.to_expr(with_const=True), q.is_affine

    def nlobjective(self, sense, expr):
        self.sense, self.obj, self.obj_lin = sense, E.wrap(expr), False

    @staticmethod
    def _bounds(sense, rhs):
        return {"<=": (-math.inf, rhs), ">=": (rhs, math.inf), "==": (rhs, rhs)}[sense]

    def constraint(self, lhs, sense, rhs):
        """@constraint(m, lhs sense rhs) with affine / quadratic sides."""
        q = E.to_quadform(E.wrap(lhs) - E.wrap(rhs))           # JuMP moves everything left, the constant right
        lo, hi = self._bounds(sense, -q.c)
        (self.lin if q.is_affine else self.quad).append((q.to_expr(), lo, hi))

    def nlconstraint(self, lhs, sense, rhs):
        """@NLconstraint(m, lhs sense rhs): JuMP stores lhs - rhs against 0 (rhs == 0 keeps lhs as is)."""
        rhs_n = E.wrap(rhs)
        body = E.wrap(lhs) if (rhs_n.op == E.OP_CONST and rhs_n.value == 0.0) else E.Node(E.OP_SUB, (E.wrap(lhs), rhs_n))
        lo, hi = self._bounds(sense, 0.0)
        self.nl.append((body, lo, hi))

    def solve(self):
        rows = [(e, True) for e, _, _ in self.lin] + [(e, False) for e, _, _ in self.quad] + [(e, False) for e, _, _ in self.nl]
        allc = self.lin + self.quad + self.nl
        d = ExprNLPEvaluator(len(self.lb), rows, self.obj, self.obj_lin)
        m = self.internal = NonlinearModel(self.solver)
        m.loadproblem(len(self.lb), len(allc), np.array(self.lb), np.array(self.ub),
                      np.array([c[1] for c in allc], dtype=float), np.array([c[2] for c in allc], dtype=float), self.sense, d)
        # the reference proceeds into optimize! whatever loadproblem! left in m.status (src/solver.jl:34-43 -> src/model.jl:219):
        # optimize! ends by overwriting it
        return m.optimize()

    def getobjectivevalue(self):
        return self.internal.getobjval()

    def getvalue(self, v):
        return float(self.internal.getsolution()[v.index])


def getKatanaModel(m):                                         # src/util.jl:3-5
    return m.internal if isinstance(m, Model) else m
