"""Generates tests/golden/kat_values.json: g and the gradient of every KAT expression
(tests/helpers.kat_problem -- the reference's own test expressions from test/2d.jl, test/3d.jl,
test/misc.jl plus operator coverage) at the KAT points, evaluated by sympy/mpmath at 60 digits from
the ANALYTIC derivative.  This is an independent pin for the oracle: the reference itself cannot
run here (no Julia), and its tests hold no per-round vectors.

Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import mpmath as mp
import sympy as sp

mp.mp.dps = 60
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, os.path.dirname(HERE))
import katana_jl_b200  # noqa: E402,F401
from katana_jl_b200 import expr as E  # noqa: E402
from helpers import kat_problem  # noqa: E402


def to_sympy(w, r, xs):
    """Rebuilds row r of a WireRows batch as a sympy expression (prefix decoding)."""
    pos = [int(w.expr_ptr[r])]

    def rec():
        k = pos[0]; pos[0] += 1
        op, arg, val = int(w.op[k]), int(w.arg[k]), float(w.val[k])
        if op == E.OP_CONST: return sp.Float(val, 60) if val != int(val) else sp.Integer(int(val))
        if op == E.OP_VAR: return xs[arg]
        ch = [rec() for _ in range(arg)]
        if op == E.OP_ADD: return sp.Add(*ch)
        if op == E.OP_SUB: return ch[0] - ch[1]
        if op == E.OP_MUL: return sp.Mul(*ch)
        if op == E.OP_DIV: return ch[0] / ch[1]
        if op == E.OP_POW: return ch[0] ** ch[1]
        if op == E.OP_NEG: return -ch[0]
        if op == E.OP_EXP: return sp.exp(ch[0])
        if op == E.OP_LOG: return sp.log(ch[0])
        if op == E.OP_SQRT: return sp.sqrt(ch[0])
        if op == E.OP_ABS: return sp.Abs(ch[0])
        if op == E.OP_SIN: return sp.sin(ch[0])
        if op == E.OP_COS: return sp.cos(ch[0])
        if op == E.OP_IFELSE: return sp.Piecewise((ch[1], ch[0]), (ch[2], True))
        if op == E.OP_LE: return sp.Le(ch[0], ch[1])
        if op == E.OP_LT: return sp.Lt(ch[0], ch[1])
        if op == E.OP_GE: return sp.Ge(ch[0], ch[1])
        if op == E.OP_GT: return sp.Gt(ch[0], ch[1])
        if op == E.OP_EQ: return sp.Eq(ch[0], ch[1])
        raise ValueError(op)
    return rec()


def main():
    nvar, w, pts = kat_problem()
    xs = sp.symbols(f"x0:{nvar}", real=True)
    out = {"num_var": nvar, "points": [list(map(float, p)) for p in pts], "rows": []}
    for r in range(w.nrows):
        e = to_sympy(w, r, xs)
        grads = [sp.diff(e, v) for v in xs]
        row = {"expr": str(e), "values": []}
        for p in pts:
            sub = {v: sp.Float(float(pv), 60) for v, pv in zip(xs, p)}
            def ev(t):
                try:
                    v = sp.N(t.subs(sub), 60)
                    c = complex(v)
                    if abs(c.imag) > 0 or c.real != c.real or abs(c.real) == float("inf"): return None
                    return mp.nstr(mp.mpf(str(v)), 25)
                except Exception:
                    return None
            # at a kink of |.| the reference's convention is d|u|/du = 1 (JuMP: ifelse(u >= 0, 1, -1)); sympy says 0: skip
            kink = any(sp.N(a.args[0].subs(sub), 60) == 0 for a in e.atoms(sp.Abs))
            row["values"].append({"g": ev(e), "grad": [None if kink else ev(g) for g in grads]})
        out["rows"].append(row)
    json.dump(out, open(os.path.join(HERE, "kat_values.json"), "w"), indent=1)
    print("wrote", os.path.join(HERE, "kat_values.json"))


if __name__ == "__main__":
    main()
