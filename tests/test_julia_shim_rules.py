"""The flattening rules of julia/gpu_separator.jl, ported 1:1 to Python and run through the C ABI.

The shim itself cannot run here (no julia binary).  What CAN be checked is that its rules -- `flatten!` over MathProgBase
`Expr` trees, `constr_body` (one-sided `body <= rhs` calls and two-sided `lb <= body <= ub` comparisons), the row flags
(`KTN_ROW_NL` from `isconstrlinear`, as the reference keeps `nlconstr_ixs`, src/model.jl:116-121, :148; `KTN_ROW_DENSE` on
the epigraph row, src/nlpeval.jl:49-54) and the epigraph row `f(x[1:n]) - x[n+1]` -- produce exactly the wire rows of the
Python host mirror (nlpeval.rows_to_wire), on the reference's own test problems (test/2d.jl, 3d.jl, misc.jl, basic.jl), and
that the library separates them identically.  A port, not the shim: every function below names the shim line it restates.
"""
import math
import re

import numpy as np
import pytest

from conftest import ROOT
from helpers import assert_batches_identical
from katana_jl_b200 import expr as E
from katana_jl_b200.binding import ROW_DENSE, ROW_NL, WireRows
from katana_jl_b200.nlpeval import EpigraphNLPEvaluator, ExprNLPEvaluator, rows_to_wire
from reference_problems import PROBLEMS

KTN_OPS = {"+": 2, "-": 3, "*": 4, "/": 5, "^": 6, "exp": 8, "log": 9, "sqrt": 10, "abs": 11, "sin": 12, "cos": 13,
           "ifelse": 14, "<=": 15, "<": 16, ">=": 17, ">": 18, "(==)": 19}     # shim: const KTN_OPS
SYM = {E.OP_ADD: "+", E.OP_SUB: "-", E.OP_MUL: "*", E.OP_DIV: "/", E.OP_POW: "^", E.OP_EXP: "exp", E.OP_LOG: "log", E.OP_SQRT: "sqrt", E.OP_ABS: "abs", E.OP_SIN: "sin", E.OP_COS: "cos", E.OP_IFELSE: "ifelse"}
CMP = {E.OP_LE: "<=", E.OP_LT: "<", E.OP_GE: ">=", E.OP_GT: ">", E.OP_EQ: "(==)"}


def julia_expr(n):
    """A Node as the Julia Expr MathProgBase would hand out: numbers, ("ref", "x", i) with 1-based i, ("call", f, args...)."""
    if n.op == E.OP_CONST:
        return n.value
    if n.op == E.OP_VAR:
        return ("ref", "x", n.index + 1)
    if n.op == E.OP_NEG:
        return ("call", "-", julia_expr(n.children[0]))
    if n.op in CMP:                                              # Julia 0.5 / 0.6 parse a <= b as Expr(:comparison, a, :<=, b)
        return ("comparison", julia_expr(n.children[0]), CMP[n.op], julia_expr(n.children[1]))
    return ("call", SYM[n.op]) + tuple(julia_expr(c) for c in n.children)


def flatten(op, arg, val, ex):                                   # shim: flatten!
    if isinstance(ex, (int, float)):
        op.append(0); arg.append(0); val.append(float(ex))
    elif ex[0] == "ref":
        op.append(1); arg.append(ex[2] - 1); val.append(0.0)
    elif ex[0] == "comparison" and len(ex) == 4:
        op.append(KTN_OPS[ex[2]]); arg.append(2); val.append(0.0)
        flatten(op, arg, val, ex[1]); flatten(op, arg, val, ex[3])
    elif ex[0] == "call":
        f, a = ex[1], ex[2:]
        if f == "-" and len(a) == 1:
            op.append(7); arg.append(1); val.append(0.0)
        else:
            op.append(KTN_OPS[f]); arg.append(len(a)); val.append(0.0)
        for c in a:
            flatten(op, arg, val, c)
    else:
        raise ValueError(f"unsupported expression node {ex}")


def constr_body(c):                                              # shim: constr_body
    return c[3] if c[0] == "comparison" else c[2]


class JuliaOracle:
    """MathProgBase view of a Python evaluator: constr_expr returns comparison Exprs, as JuMP's evaluator does."""

    def __init__(self, d, lo, hi):
        self.d, self.lo, self.hi = d, lo, hi

    def constr_expr(self, i):                                    # 1-based
        body = julia_expr(E.wrap(self.d.constr_expr(i - 1)))
        lo, hi = self.lo[i - 1], self.hi[i - 1]
        if math.isfinite(lo) and math.isfinite(hi) and lo != hi:
            return ("comparison", lo, "<=", body, "<=", hi)
        if lo == hi:
            return ("call", "==", body, hi)
        return ("call", "<=", body, hi) if math.isfinite(hi) else ("call", ">=", body, lo)

    def isconstrlinear(self, i):
        return self.d.isconstrlinear(i - 1)


class JuliaEpigraph:
    """shim: the three MathProgBase methods it adds for EpigraphNLPEvaluator."""

    def __init__(self, inner, obj, num_var, num_constr):
        self.inner, self.obj, self.num_var, self.num_constr = inner, obj, num_var, num_constr

    def constr_expr(self, i):
        if i < self.num_constr:
            return self.inner.constr_expr(i)
        return ("call", "<=", ("call", "-", julia_expr(E.wrap(self.obj)), ("ref", "x", self.num_var)), 0.0)

    def isconstrlinear(self, i):
        return i < self.num_constr and self.inner.isconstrlinear(i)


def shim_wire(oracle, num_constr, is_epigraph):                  # shim: the loop of initialize!
    op, arg, val, eptr = [], [], [], [0]
    flags = np.zeros(num_constr, np.uint8)
    dense_row = num_constr if is_epigraph else 0
    for i in range(1, num_constr + 1):
        flatten(op, arg, val, constr_body(oracle.constr_expr(i))); eptr.append(len(op))
        flags[i - 1] = (0 if oracle.isconstrlinear(i) else 1) | (2 if i == dense_row else 0)
    return WireRows(np.asarray(eptr, np.int64), np.asarray(op, np.int32), np.asarray(arg, np.int32), np.asarray(val, np.float64),
                    np.full(num_constr, -np.inf), np.full(num_constr, np.inf), flags)


class Capture:
    """Records what Model.solve would hand to loadproblem! without solving."""

    def __init__(self):
        from katana_jl_b200.solver import Model
        self.m = Model(solver=None)

    def evaluator(self):
        m = self.m
        rows = [(r[0], True) for r in m.lin] + [(r[0], False) for r in m.quad] + [(e, False) for e, _, _ in m.nl]
        allc = m.lin + m.quad + m.nl
        return ExprNLPEvaluator(len(m.lb), rows, m.obj, m.obj_lin), [c[1] for c in allc], [c[2] for c in allc]


def problems():
    for name, cite, build, obj, sol in PROBLEMS:
        cap = Capture()
        build(cap.m)
        d, lo, hi = cap.evaluator()
        yield name, cite, d, lo, hi


def test_shim_rules_reproduce_the_host_mirror(oracle_lib):
    n_two_sided = n_lin = n_dense = 0
    for name, cite, d, lo, hi in problems():
        nv, nc = d.num_var, len(d.rows)
        # (1) the user separator's evaluator: plain when the objective is linear, epigraph-wrapped otherwise (src/model.jl:166-172)
        for epi in ((False, True) if not d.isobjlinear() else (False,)):
            if epi:
                twin_d = EpigraphNLPEvaluator(d, nv + 1, nc + 1)
                jl = JuliaEpigraph(JuliaOracle(d, lo, hi), d.obj_expr(), nv + 1, nc + 1)
                m_rows, m_var = nc + 1, nv + 1
            else:
                twin_d, jl, m_rows, m_var = d, JuliaOracle(d, lo, hi), nc, nv
            if m_rows == 0:
                continue
            twin = rows_to_wire(twin_d, m_rows, np.full(m_rows, -np.inf), np.full(m_rows, np.inf))
            shim = shim_wire(jl, m_rows, epi)
            for f in ("expr_ptr", "op", "arg", "val", "flags"):
                assert np.array_equal(getattr(twin, f), getattr(shim, f)), (name, cite, f, epi)
            n_two_sided += sum(1 for i in range(1, nc + 1) if JuliaOracle(d, lo, hi).constr_expr(i)[0] == "comparison")
            n_lin += int(np.sum((shim.flags & ROW_NL) == 0)); n_dense += int(np.sum((shim.flags & ROW_DENSE) != 0))
            # (2) through the C ABI: one round and the unconditional cuts of every row, at a point away from the optimum
            ht, hs = oracle_lib.create(), oracle_lib.create()
            ht.load(m_var, twin); hs.load(m_var, shim)
            bl = np.array(lo + ([-np.inf] if epi else [])); bu = np.array(hi + ([0.0] if epi else []))
            ht.set_bounds(bl, bu); hs.set_bounds(bl, bu)
            x = np.linspace(0.3, 1.7, m_var)
            assert_batches_identical(ht.separate(x), hs.separate(x), name)
            rows = np.arange(m_rows, dtype=np.int64)
            assert_batches_identical(ht.gencut_rows(x, rows, False), hs.gencut_rows(x, rows, False), name)
            ht.close(); hs.close()
    assert n_lin > 0 and n_dense > 0           # linear rows are NOT flagged NL, the epigraph row IS flagged dense


def test_shim_source_states_the_rules():
    """The shipped shim contains the rules this file ports (a rename there must be followed here)."""
    shim = open(f"{ROOT}/julia/gpu_separator.jl").read()
    assert "MathProgBase.isconstrlinear(oracle, i) ? 0x00 : 0x01" in shim
    assert "i == dense_row ? 0x02 : 0x00" in shim and "oracle isa EpigraphNLPEvaluator ? num_constr : 0" in shim
    assert "c.head == :comparison ? c.args[3] : c.args[2]" in shim
    assert re.search(r"set_bounds!\(m\.params\.separator, m\.l_constr, m\.u_constr\)", shim)
    assert "fill(NaN, length(vars))" in shim                    # gencut of a non-finite row: _addcut must see NaN (src/model.jl:69-73)
    ops = dict(re.findall(r":(\S+) => (\d+)", re.search(r"const KTN_OPS = Dict\((.*?=> 19)\)", shim, flags=re.S).group(1)))
    assert {k: int(v) for k, v in ops.items()} == KTN_OPS
