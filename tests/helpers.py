"""Shared test helpers: bit-level comparison of cut batches and the KAT expression set."""
import numpy as np

from katana_jl_b200 import expr as E
from katana_jl_b200.binding import ROW_DENSE, ROW_NL

BATCH_FIELDS = ("row_id", "row_ptr", "col", "val", "lo", "hi", "g", "viol", "bconst")


def bits_equal(a, b):
    """Bit-exact equality for integer / float arrays; NaNs must sit at the same places (payloads ignored)."""
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype.kind == "f":
        na, nb = np.isnan(a), np.isnan(b)
        return bool(np.array_equal(na, nb) and np.array_equal(a[~na].view(np.int64), b[~nb].view(np.int64)))
    return bool(np.array_equal(a, b))


def assert_batches_identical(ref, got, what=""):
    assert ref.status == got.status, f"{what}: status {ref.status} vs {got.status}"
    assert ref.err_row == got.err_row, f"{what}: err_row {ref.err_row} vs {got.err_row}"
    for f in BATCH_FIELDS:
        a, b = getattr(ref, f), getattr(got, f)
        assert bits_equal(a, b), f"{what}: field {f} differs (shapes {a.shape} {b.shape})"


def kat_problem():
    """The reference's test expressions (test/2d.jl, test/3d.jl, test/misc.jl) plus operator coverage.
    Returns (num_var, WireRows, points)."""
    x, y, z = E.var(0), E.var(1), E.var(2)
    exprs = [
        x**2 + y**2,                                  # disk, test/2d.jl:12
        E.const(np.e)**(x - 2.0) - 0.5 - y,           # test/2d.jl:282  e^(x-2.0) - 0.5 <= y
        y - (E.log(x) + 0.5),                         # test/2d.jl:283
        E.sqrt(x**2 + y**2) - (z - 0.25),             # test/3d.jl:161
        y * E.exp(x / y) - z,                         # test/3d.jl:229
        y * E.exp((-x) / y) - z,                      # test/3d.jl:230
        2 * x**2 - 4 * x * y - 4 * x + 4 - y,         # test/2d.jl:467
        E.Node(E.OP_MUL, (E.const(1.5), x, y, z)),    # n-ary product (all-but-one partials)
        E.Node(E.OP_MUL, (E.const(2.0), x, x)) + E.Node(E.OP_MUL, (E.const(-1.0), y)),   # JuMP quadratic form c*x*x
        x**3 + y**0.5 + 2.0**z,                       # general powers, variable exponent
        x**y,
        E.abs_(x - y) + z / (x * x + 1.0),
        (x - 1.0)**2 + (y - 2.0)**2 - z,              # test/basic.jl:86 lifted objective
        x,
        E.const(3.0) * x - y / 2.0 + 1.0,
        -(x + y + 2 * z),                             # test/3d.jl:95
        E.sum_([E.exp(E.var(i)) for i in range(3)]),  # test/2d.jl:92 style
        1.0 / x + x / (y + z),
        (x * y)**2 / z,
        E.sqrt(E.sum_([E.var(i)**2 for i in range(3)])),   # test/misc.jl:39
        E.sin(-x - 1.0) + x / 2.0 + 0.5 - y,               # test/2d.jl:367 (case 106, commented out upstream: not convex on the box)
        y - (E.cos(x - 0.5) + x / 4.0 - 0.5),              # test/2d.jl:368
        E.sin(x * y) * E.cos(z) + E.cos(E.exp(x)),          # sin / cos under products and of transcendental arguments
        E.ifelse(E.le(x, 1.0), x**2, 2.0 * x - 1.0) + E.ifelse(E.gt(y * z, x), E.exp(y), y + 1.0),   # ifelse with comparisons (JuMP's user-visible piecewise form)
        E.sum_([E.var(i)**2 for i in range(3)]) - z,       # dense epigraph-style row, last
    ]
    m = len(exprs)
    ub = np.array([1.0, 0, 0, 0, 0, 0, 0, 1, 1, 5, 2, 1, 0, 0.5, 1, 0, 4, 3, 1, 1, 0.5, -0.25, 0.75, 1.5, 0.0])
    assert len(ub) == m
    flags = [ROW_NL] * (m - 1) + [ROW_NL | ROW_DENSE]
    w = E.to_wire(exprs, np.full(m, -np.inf), ub, flags)
    pts = [np.array(p, float) for p in ([2, 2, 1], [0.5, 1.5, 0.25], [0.3, 0.7, 2.0], [-1, 3, 2], [1, 0.5, 2], [3, -2, 0.5], [0, 0, 0], [1, 0, 2])]
    return 3, w, pts


def random_tree(rng, nvar, depth):
    """A random expression over nvar variables using every operator of the wire format."""
    if depth == 0 or rng.random() < 0.25:
        return E.var(int(rng.integers(nvar))) if rng.random() < 0.7 else E.const(float(np.round(rng.uniform(-2, 2), 3)))
    k = rng.integers(15)
    sub = lambda: random_tree(rng, nvar, depth - 1)
    if k == 0: return E.sum_([sub() for _ in range(int(rng.integers(1, 5)))])
    if k == 1: return E.prod_([sub() for _ in range(int(rng.integers(1, 4)))])
    if k == 2: return E.Node(E.OP_SUB, (sub(), sub()))
    if k == 3: return E.Node(E.OP_DIV, (sub(), sub()))
    if k == 4: return E.Node(E.OP_POW, (sub(), E.const(2.0)))
    if k == 5: return E.Node(E.OP_POW, (sub(), E.const(float(rng.choice([1.0, 3.0, 0.5, -1.0, 2.5])))))
    if k == 6: return E.Node(E.OP_POW, (sub(), sub()))
    if k == 7: return E.exp(sub())
    if k == 8: return E.log(sub())
    if k == 9: return E.sqrt(sub())
    if k == 10: return E.abs_(sub())
    if k == 11: return E.sin(sub())
    if k == 12: return E.cos(sub())
    if k == 13:
        cmp = [E.le, E.lt, E.ge, E.gt, E.eq][int(rng.integers(5))]
        return E.ifelse(cmp(sub(), sub()), sub(), sub())
    return -sub()
