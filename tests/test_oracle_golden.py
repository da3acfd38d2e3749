"""Pins the CPU oracle: hand-derived per-round KATs and the committed
sympy/mpmath golden vectors (tests/golden/kat_values.json, generator beside it)."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT
from helpers import kat_problem
from katana_jl_b200 import expr as E
from katana_jl_b200.binding import KTN_NUMERIC_NONFINITE, ROW_DENSE, ROW_NL


def test_disk_kat_first_lp_vertex(oracle_lib):
    """x^2 + y^2 <= 1 at the first LP vertex (2,2) of test/2d.jl:5-20: g = 8, row (4,4), b = -8, cut 4x + 4y <= 9."""
    x, y = E.var(0), E.var(1)
    h = oracle_lib.create()
    h.load(2, E.to_wire([x**2 + y**2], [-np.inf], [1.0], [ROW_NL]))
    b = h.separate(np.array([2.0, 2.0]))
    assert b.status == 0 and b.n_cuts == 1
    assert list(b.col) == [0, 1] and list(b.val) == [4.0, 4.0]
    assert b.g[0] == 8.0 and b.bconst[0] == -8.0 and b.hi[0] == 9.0 and b.lo[0] == -np.inf and b.viol[0] == 7.0
    assert h.separate(np.array([0.6, 0.6])).n_cuts == 0          # g = 0.72 <= 1


def test_quadratic_evaluator_form_duplicates_merge(oracle_lib):
    """JuMP's quadratic form c*x*x (n-ary product): two occurrences of x accumulate into one sorted column."""
    x, y = E.var(0), E.var(1)
    e = E.sum_([E.Node(E.OP_MUL, (E.const(1.0), x, x)), E.Node(E.OP_MUL, (E.const(1.0), y, y))])
    h = oracle_lib.create(); h.load(2, E.to_wire([e], [-np.inf], [1.0], [ROW_NL]))
    b = h.separate(np.array([2.0, 2.0]))
    assert list(b.col) == [0, 1] and list(b.val) == [4.0, 4.0] and b.hi[0] == 9.0


def test_tolerance_is_a_hard_threshold(oracle_lib):
    """isconstrsat: g <= ub + f_tol (src/separators.jl:120), NaN counts as violated."""
    x = E.var(0)
    h = oracle_lib.create(f_tol=1e-6); h.load(1, E.to_wire([x * 1.0, E.log(x)], [-np.inf, -np.inf], [1.0, 5.0], [ROW_NL, ROW_NL]))
    assert h.separate(np.array([1.0 + 1e-6])).n_cuts == 0
    assert list(h.separate(np.array([np.nextafter(1.0 + 1e-6, 2.0)])).row_id) == [0]
    b = h.separate(np.array([-1.0]))                              # log(-1) = NaN: both comparisons false -> selected
    assert list(b.row_id) == [1] and np.isnan(b.g[0]) and list(b.val) == [-1.0]   # 1/x is finite, so _addcut accepts the row
    assert b.status == 0 and np.isnan(b.hi[0])


def test_sqrt_cone_origin_is_error(oracle_lib):
    """test/3d.jl:161 at the origin: gradient of sqrt(x^2+y^2) is NaN -> _addcut sets :Error (src/model.jl:69-73)."""
    x, y, z = E.var(0), E.var(1), E.var(2)
    h = oracle_lib.create(); h.load(3, E.to_wire([x**2 + y**2 - 1.0, E.sqrt(x**2 + y**2) - (z - 0.25)], [-np.inf] * 2, [-2.0, 0.0], [ROW_NL] * 2))
    b = h.separate(np.zeros(3))
    assert b.status == KTN_NUMERIC_NONFINITE and b.err_row == 1
    assert list(b.row_id) == [0]                                  # the cut before the failing row is still delivered


def test_round_coefs_signed_max(oracle_lib):
    """round_coefs (src/model.jl:200-207): coefficients more than cut_coef_rng below the SIGNED maximum are zeroed; b untouched."""
    x, y = E.var(0), E.var(1)
    h = oracle_lib.create(cut_coef_rng=10.0); h.load(2, E.to_wire([100.0 * x + 1.0 * y + (-50.0) * x * 0 + 0.0], [-np.inf], [0.0], [ROW_NL]))
    b = h.separate(np.array([1.0, 1.0]))
    assert list(b.val) == [100.0, 0.0]
    assert b.bconst[0] == (101.0 + -100.0) + -1.0                 # computed with the unrounded row


def test_epigraph_dense_row(oracle_lib):
    """src/nlpeval.jl:49-63: dense last row over all columns, -1 on the auxiliary variable."""
    x, y, t = E.var(0), E.var(1), E.var(3)
    h = oracle_lib.create(); h.load(4, E.to_wire([(x - 1.0)**2 + (y - 2.0)**2 - t], [-np.inf], [0.0], [ROW_NL | ROW_DENSE]))
    rp, cols = h.jac_structure()
    assert list(cols) == [0, 1, 2, 3]
    b = h.separate(np.array([0.0, 0.0, 7.0, 1.0]))
    assert list(b.col) == [0, 1, 2, 3] and list(b.val) == [-2.0, -4.0, 0.0, -1.0] and b.g[0] == 4.0


def test_against_mpmath_golden_vectors(oracle_lib):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "kat_values.json")))
    nvar, w, pts = kat_problem()
    assert gold["num_var"] == nvar and len(gold["rows"]) == w.nrows
    h = oracle_lib.create(); h.load(nvar, w)
    rp, cols = h.jac_structure()
    checked = 0
    for pi, p in enumerate(pts):
        assert list(map(float, p)) == gold["points"][pi]
        g = h.eval_g(p)
        cuts = h.gencut_rows(p, np.arange(w.nrows, dtype=np.int64), round_coefs=False)
        by_row = {int(r): c for c, r in enumerate(cuts.row_id)}
        for r in range(w.nrows):
            ref = gold["rows"][r]["values"][pi]
            if ref["g"] is None or any(v is None for v in ref["grad"]):
                continue
            if not np.isfinite(g[r]):   # singular point (0/0, 0*inf): IEEE yields NaN where sympy simplifies; pinned bit-level elsewhere
                continue
            assert g[r] == pytest.approx(float(ref["g"]), rel=2e-14, abs=1e-300), (r, pi)
            if r not in by_row:
                continue
            ccols, cvals = cuts.row(by_row[r])
            dense = np.zeros(nvar); dense[ccols] = cvals
            for j in range(nvar):
                assert dense[j] == pytest.approx(float(ref["grad"][j]), rel=1e-13, abs=1e-300), (r, pi, j)
                checked += 1
    assert checked > 200


def test_ladder_is_the_sequential_search(oracle_lib):
    """ktn_separate_ladder on the oracle IS boundroutine's loop (src/model.jl:175-197): rounds at 2^n * ray, n = 2, 3, ..., until a
    row is violated; the cuts are those of that round."""
    import numpy as np
    from katana_jl_b200 import expr as E
    from katana_jl_b200.binding import ROW_NL
    x, y = E.var(0), E.var(1)
    exprs = [x**2 + y**2, E.exp(x) - y, x + y]                       # disk of radius 100, exp(x) <= y + 5000, a linear row (never tested)
    w = E.to_wire(exprs, [-np.inf] * 3, [1e4, 5e3, 1.0], [ROW_NL, ROW_NL, 0])
    h = oracle_lib.create(); h.load(2, w)
    ray = np.array([1.0, 0.25])
    n_hit, b = h.separate_ladder(ray)
    want_n = next(n for n in range(2, 1024) if h.separate(2.0**n * ray).n_cuts > 0)
    assert n_hit == want_n == 4                                      # exp(2^4) - 4 > 5000 (2^3: exp(8) - 2 = 2979); the disk holds until 2^7
    ref = h.separate(2.0**want_n * ray)
    assert ref.n_cuts == b.n_cuts and np.array_equal(ref.val, b.val) and np.array_equal(ref.row_id, b.row_id)
    n_hit, b = h.separate_ladder(np.array([-1.0, 1.0]), 2, 5)        # a direction that stays feasible for the points asked: no cuts
    assert n_hit == -1 and b.n_cuts == 0
