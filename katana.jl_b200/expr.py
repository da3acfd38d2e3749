"""Expression trees and their flattening to the wire format of include/ktn.h.

This is the host-side half of the tape compiler the north star puts in nlpeval.jl: the
MathProgBase `constr_expr` / `obj_expr` Expr trees are flattened to prefix
(op, arg, val) arrays; shape grouping and SoA packing happen inside the C-ABI library
(csrc/ktn_compile.cpp) so the Julia shim and this mirror share one compiler.
"""
import math

import numpy as np

from .binding import (OP_ABS, OP_ADD, OP_CONST, OP_COS, OP_SIN, OP_IFELSE, OP_LE, OP_LT, OP_GE, OP_GT, OP_EQ, OP_DIV, OP_EXP, OP_LOG, OP_MUL, OP_NEG, OP_POW, OP_SQRT, OP_SUB, OP_VAR,
                      ROW_DENSE, ROW_NL, WireRows)

_NARY = (OP_ADD, OP_MUL)


class Node:
    """One expression node.  `op` is a KTN_OP_* code; leaves carry `index` (VAR) or `value` (CONST)."""
    __slots__ = ("op", "children", "index", "value", "_chain")

    def __init__(self, op, children=(), index=0, value=0.0, chain=False):
        self.op, self.children, self.index, self.value = op, tuple(children), index, float(value)
        self._chain = chain  # built by an infix chain a+b+c / a*b*c, which Julia parses as ONE n-ary call

    # Julia parses a+b+c as +(a,b,c) and a*b*c as *(a,b,c); subtraction and division stay binary.
    def _nary(self, op, other, swap=False):
        o = wrap(other)
        if swap:
            return Node(op, (o, self), chain=True)
        if self.op == op and self._chain:
            return Node(op, self.children + (o,), chain=True)
        return Node(op, (self, o), chain=True)

    def __add__(self, o): return self._nary(OP_ADD, o)
    def __radd__(self, o): return self._nary(OP_ADD, o, swap=True)
    def __mul__(self, o): return self._nary(OP_MUL, o)
    def __rmul__(self, o): return self._nary(OP_MUL, o, swap=True)
    def __sub__(self, o): return Node(OP_SUB, (self, wrap(o)))
    def __rsub__(self, o): return Node(OP_SUB, (wrap(o), self))
    def __truediv__(self, o): return Node(OP_DIV, (self, wrap(o)))
    def __rtruediv__(self, o): return Node(OP_DIV, (wrap(o), self))
    def __pow__(self, o): return Node(OP_POW, (self, wrap(o)))
    def __rpow__(self, o): return Node(OP_POW, (wrap(o), self))
    def __neg__(self): return Node(OP_NEG, (self,))
    def __pos__(self): return self


def wrap(v):
    if isinstance(v, Node):
        if v._chain:  # a parenthesised chain used as an operand no longer extends
            return Node(v.op, v.children, v.index, v.value, chain=False)
        return v
    return Node(OP_CONST, value=float(v))


def var(index):
    return Node(OP_VAR, index=int(index))


def const(v):
    return Node(OP_CONST, value=v)


def call(op, *args):
    return Node(op, [wrap(a) for a in args])


def exp(a): return call(OP_EXP, a)
def log(a): return call(OP_LOG, a)
def sqrt(a): return call(OP_SQRT, a)
def abs_(a): return call(OP_ABS, a)
def sin(a): return call(OP_SIN, a)
def cos(a): return call(OP_COS, a)
def ifelse(cond, a, b): return call(OP_IFELSE, cond, a, b)      # JuMP: ifelse(x <= 1, x^2, 2x - 1); both branches are evaluated
def le(a, b): return call(OP_LE, a, b)
def lt(a, b): return call(OP_LT, a, b)
def ge(a, b): return call(OP_GE, a, b)
def gt(a, b): return call(OP_GT, a, b)
def eq(a, b): return call(OP_EQ, a, b)


def sum_(terms):
    """n-ary + over an iterable, like Julia's sum(... for ...) inside @NLconstraint."""
    terms = [wrap(t) for t in terms]
    return Node(OP_ADD, terms)


def prod_(terms):
    return Node(OP_MUL, [wrap(t) for t in terms])


def flatten_into(node, ops, args, vals):
    """Append `node` in prefix order (iterative: objective rows can be very deep / wide)."""
    stack = [node]
    while stack:
        n = stack.pop()
        ops.append(n.op)
        if n.op == OP_VAR:
            args.append(n.index); vals.append(0.0)
        elif n.op == OP_CONST:
            args.append(0); vals.append(n.value)
        else:
            args.append(len(n.children)); vals.append(0.0)
            stack.extend(reversed(n.children))


def to_wire(exprs, lb, ub, flags):
    """Flatten a list of expression trees into one WireRows batch."""
    if not (len(lb) == len(ub) == len(flags) == len(exprs)):
        raise ValueError(f"to_wire: {len(exprs)} rows but {len(lb)} lower bounds, {len(ub)} upper bounds, {len(flags)} flags")
    ops, args, vals, ptr = [], [], [], [0]
    for e in exprs:
        flatten_into(wrap(e), ops, args, vals)
        ptr.append(len(ops))
    return WireRows(np.asarray(ptr, np.int64), np.asarray(ops, np.int32), np.asarray(args, np.int32), np.asarray(vals, np.float64),
                    np.ascontiguousarray(lb, np.float64), np.ascontiguousarray(ub, np.float64), np.ascontiguousarray(flags, np.uint8))


def variables(node):
    out, stack = set(), [node]
    while stack:
        n = stack.pop()
        if n.op == OP_VAR:
            out.add(n.index)
        stack.extend(n.children)
    return out


def evaluate(node, x):
    """Host evaluation with libm (used for eval_f at load time only; never on the separation path)."""
    op = node.op
    if op == OP_CONST: return node.value
    if op == OP_VAR: return float(x[node.index])
    c = [evaluate(k, x) for k in node.children]
    if op == OP_ADD: return math.fsum(c) if False else sum(c, 0.0)
    if op == OP_SUB: return c[0] - c[1]
    if op == OP_MUL:
        p = 1.0
        for v in c: p *= v
        return p
    if op == OP_DIV: return c[0] / c[1] if c[1] != 0 else math.copysign(math.inf, c[0]) if c[0] != 0 else math.nan
    if op == OP_POW:
        try: return c[0] ** c[1]
        except (OverflowError, ZeroDivisionError): return math.inf
        except ValueError: return math.nan
    if op == OP_NEG: return -c[0]
    if op == OP_EXP:
        try: return math.exp(c[0])
        except OverflowError: return math.inf
    if op == OP_LOG: return math.log(c[0]) if c[0] > 0 else (-math.inf if c[0] == 0 else math.nan)
    if op == OP_SQRT: return math.sqrt(c[0]) if c[0] >= 0 else math.nan
    if op == OP_ABS: return abs(c[0])
    if op == OP_SIN: return math.sin(c[0]) if math.isfinite(c[0]) else math.nan
    if op == OP_COS: return math.cos(c[0]) if math.isfinite(c[0]) else math.nan
    if op == OP_IFELSE: return c[1] if c[0] == 1.0 else c[2]
    if op == OP_LE: return 1.0 if c[0] <= c[1] else 0.0
    if op == OP_LT: return 1.0 if c[0] < c[1] else 0.0
    if op == OP_GE: return 1.0 if c[0] >= c[1] else 0.0
    if op == OP_GT: return 1.0 if c[0] > c[1] else 0.0
    if op == OP_EQ: return 1.0 if c[0] == c[1] else 0.0
    raise ValueError(f"unknown op {op}")


# ---- polynomial normal form for @constraint / @objective (affine or quadratic) -------------------
class QuadForm:
    """const + sum lin[i] x_i + sum quad[(i,j)] x_i x_j with i <= j, like JuMP's AffExpr / QuadExpr."""

    def __init__(self, c=0.0, lin=None, quad=None):
        self.c, self.lin, self.quad = float(c), dict(lin or {}), dict(quad or {})

    def add(self, o, s=1.0):
        r = QuadForm(self.c + s * o.c, self.lin, self.quad)
        for k, v in o.lin.items(): r.lin[k] = r.lin.get(k, 0.0) + s * v
        for k, v in o.quad.items(): r.quad[k] = r.quad.get(k, 0.0) + s * v
        return r

    def mul(self, o):
        if self.quad and (o.lin or o.quad) or o.quad and (self.lin or self.quad):
            raise ValueError("expression is not quadratic; use the NL form")
        r = QuadForm(self.c * o.c)
        for k, v in self.lin.items(): r.lin[k] = r.lin.get(k, 0.0) + v * o.c
        for k, v in o.lin.items(): r.lin[k] = r.lin.get(k, 0.0) + v * self.c
        for k, v in self.quad.items(): r.quad[k] = r.quad.get(k, 0.0) + v * o.c
        for k, v in o.quad.items(): r.quad[k] = r.quad.get(k, 0.0) + v * self.c
        for i, a in self.lin.items():
            for j, b in o.lin.items():
                key = (min(i, j), max(i, j))
                r.quad[key] = r.quad.get(key, 0.0) + a * b
        return r

    @property
    def is_affine(self):
        return not any(v != 0.0 for v in self.quad.values())

    def to_expr(self, with_const=False):
        """JuMP-style Expr: +(c*x_i*x_j ..., c*x_k ..., [const])."""
        terms = [Node(OP_MUL, (const(v), var(i), var(j))) for (i, j), v in sorted(self.quad.items()) if v != 0.0]
        terms += [Node(OP_MUL, (const(v), var(i))) for i, v in sorted(self.lin.items()) if v != 0.0 or not self.quad]
        if with_const and self.c != 0.0:
            terms.append(const(self.c))
        if not terms:
            terms = [const(0.0)]
        return Node(OP_ADD, terms)


def to_quadform(node):
    node = wrap(node)
    op = node.op
    if op == OP_CONST: return QuadForm(node.value)
    if op == OP_VAR: return QuadForm(0.0, {node.index: 1.0})
    ch = [to_quadform(c) for c in node.children]
    if op == OP_ADD:
        r = ch[0]
        for c in ch[1:]: r = r.add(c)
        return r
    if op == OP_SUB: return ch[0].add(ch[1], -1.0)
    if op == OP_NEG: return QuadForm().add(ch[0], -1.0)
    if op == OP_MUL:
        r = ch[0]
        for c in ch[1:]: r = r.mul(c)
        return r
    if op == OP_DIV:
        if ch[1].lin or ch[1].quad: raise ValueError("division by a variable is not quadratic")
        return ch[0].mul(QuadForm(1.0 / ch[1].c))
    if op == OP_POW:
        e = ch[1]
        if e.lin or e.quad or e.c not in (0.0, 1.0, 2.0): raise ValueError("only powers 0, 1, 2 are polynomial here")
        return QuadForm(1.0) if e.c == 0.0 else ch[0] if e.c == 1.0 else ch[0].mul(ch[0])
    raise ValueError("expression is not polynomial; use the NL form")


__all__ = ["Node", "var", "const", "call", "exp", "log", "sqrt", "abs_", "sin", "cos", "ifelse", "le", "lt", "ge", "gt", "eq", "sum_", "prod_", "to_wire", "flatten_into", "evaluate",
           "variables", "QuadForm", "to_quadform", "wrap", "ROW_NL", "ROW_DENSE"]
