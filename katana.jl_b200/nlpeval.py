"""NLP evaluators on the host side of the tape compiler.

Mirrors reference src/nlpeval.jl.  The north star replaces the JuMP / ReverseDiffSparse evaluator
behind `eval_g` / `eval_jac_g` by expression tapes: an evaluator here only has to hand out
MathProgBase-style expression graphs (`constr_expr`, `obj_expr`, feature :ExprGraph); flattening
them (expr.to_wire) and everything after happens behind the C ABI.
"""
import numpy as np

from . import expr as E


class ExprNLPEvaluator:
    """A MathProgBase.AbstractNLPEvaluator over expression trees (what JuMP hands to loadproblem!).

    rows: list of (Node, is_linear) in JuMP's row order: linear rows first, then quadratic, then
    @NLconstraint rows [recalled JuMP 0.18 ordering, SURVEY.md section 8a8].
    """

    def __init__(self, num_var, rows, obj, obj_is_linear):
        self.num_var, self.rows, self.obj, self.obj_is_linear = num_var, rows, obj, obj_is_linear
        self.initialized = False

    def features_available(self):
        return ["ExprGraph"]

    def initialize(self, requested_features):
        for f in requested_features:
            if f not in self.features_available():
                raise ValueError(f"Unsupported feature {f}")
        self.initialized = True

    def isobjlinear(self): return self.obj_is_linear
    def isconstrlinear(self, i): return self.rows[i][1]
    def obj_expr(self): return self.obj
    def constr_expr(self, i): return self.rows[i][0]
    def isconstrdense(self, i): return False
    def eval_f(self, x): return E.evaluate(self.obj, x)


class EpigraphNLPEvaluator:
    """Reference src/nlpeval.jl:6-63: the objective becomes the last constraint f(x[1:n]) - x[n+1].

    The reference appends a DENSE Jacobian row (every column, src/nlpeval.jl:49-54) whose last
    entry is -1 (src/nlpeval.jl:36-39); here the row is the expression `obj - x[num_var-1]`
    flagged dense, and the -1 falls out of the reverse sweep of the binary minus.
    """

    def __init__(self, d, num_var, num_constr):
        self.nlpeval, self.num_var, self.num_constr = d, num_var, num_constr

    def features_available(self): return ["ExprGraph"]          # src/nlpeval.jl:23 advertised [:Grad, :Jac]

    def initialize(self, requested_features):                    # src/nlpeval.jl:25-32
        for f in requested_features:
            if f not in self.features_available():
                raise ValueError(f"Unsupported feature {f}")
        self.nlpeval.initialize(requested_features)

    def isobjlinear(self): return self.nlpeval.isobjlinear()     # src/nlpeval.jl:17
    def obj_expr(self): return self.nlpeval.obj_expr()           # src/nlpeval.jl:20

    def isconstrlinear(self, i):
        return self.nlpeval.isconstrlinear(i) if i < self.num_constr - 1 else False

    def isconstrdense(self, i):
        return i == self.num_constr - 1

    def constr_expr(self, i):
        if i < self.num_constr - 1:
            return self.nlpeval.constr_expr(i)
        return E.Node(E.OP_SUB, (E.wrap(self.nlpeval.obj_expr()), E.var(self.num_var - 1)))   # src/nlpeval.jl:35,42-45

    def eval_f(self, x):                                         # src/nlpeval.jl:35
        return self.nlpeval.eval_f(x[:-1]) - x[-1]


def rows_to_wire(oracle, num_constr, l_constr, u_constr, nl_flags=None):
    """Flatten every row of an evaluator for ktn_add_rows.  NL rows are the ones src/model.jl:116-121
    keeps in nlconstr_ixs: those `isconstrlinear` rejects."""
    exprs, flags = [], []
    for i in range(num_constr):
        exprs.append(oracle.constr_expr(i))
        nl = (not oracle.isconstrlinear(i)) if nl_flags is None else bool(nl_flags[i])
        flags.append((E.ROW_NL if nl else 0) | (E.ROW_DENSE if oracle.isconstrdense(i) else 0))
    return E.to_wire(exprs, np.asarray(l_constr, np.float64), np.asarray(u_constr, np.float64), flags)
