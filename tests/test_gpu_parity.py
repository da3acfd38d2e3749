"""GPU parity: the CUDA library (through the C ABI) against the CPU oracle, bit for bit.
Small seeded cases the oracle finishes in seconds, the committed golden vectors, edge cases, and the
BASELINE.json sizes (10^6 rows) checked both against the oracle and through size-independent
properties.  Run on the B200 box:  python -m pytest tests -m gpu"""
import json
import os
import warnings

import numpy as np
import pytest

from conftest import ROOT
from helpers import BATCH_FIELDS, assert_batches_identical, bits_equal, kat_problem, random_tree
from katana_jl_b200 import expr as E
from katana_jl_b200.binding import KTN_NUMERIC_NONFINITE, ROW_DENSE, ROW_NL

pytestmark = pytest.mark.gpu


def both(oracle_lib, cuda_lib, nvar, w, **kw):
    ho, hc = oracle_lib.create(**kw), cuda_lib.create(**kw)
    ho.load(nvar, w); hc.load(nvar, w)
    return ho, hc


def test_backend_is_cuda(cuda_lib):
    assert cuda_lib.backend == "cuda"


def test_kat_rounds_identical(oracle_lib, cuda_lib):
    nvar, w, pts = kat_problem()
    ho, hc = both(oracle_lib, cuda_lib, nvar, w)
    assert all(np.array_equal(a, b) for a, b in zip(ho.jac_structure(), hc.jac_structure()))
    rows = np.arange(w.nrows, dtype=np.int64)
    for p in pts:
        assert_batches_identical(ho.separate(p), hc.separate(p), f"separate at {p}")
        assert bits_equal(ho.eval_g(p), hc.eval_g(p))
        assert_batches_identical(ho.gencut_rows(p, rows, False), hc.gencut_rows(p, rows, False), f"gencut at {p}")
        assert_batches_identical(ho.gencut_rows(p, rows[::3], True), hc.gencut_rows(p, rows[::3], True), f"gencut+round at {p}")


def test_golden_vectors_on_gpu(cuda_lib):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "kat_values.json")))
    nvar, w, pts = kat_problem()
    h = cuda_lib.create(); h.load(nvar, w)
    checked = 0
    for pi, p in enumerate(pts):
        g = h.eval_g(p)
        for r in range(w.nrows):
            ref = gold["rows"][r]["values"][pi]
            if ref["g"] is None or not np.isfinite(g[r]):
                continue
            assert g[r] == pytest.approx(float(ref["g"]), rel=2e-14, abs=1e-300)
            checked += 1
    assert checked > 100


@pytest.mark.parametrize("kind,nv,nr", [(0, 1000, 5000), (1, 5000, 20000), (2, 2000, 3000), (1, 100, 33), (0, 50, 1)])
def test_synthetic_families_identical(oracle_lib, cuda_lib, kind, nv, nr):
    w = cuda_lib.synth_rows(kind, 20260001 + kind, nv, 0, nr)
    x0 = cuda_lib.synth_point(kind, 20260001 + kind, nv)
    ho, hc = both(oracle_lib, cuda_lib, nv, w)
    g = ho.eval_g(x0)
    assert bits_equal(g, hc.eval_g(x0))
    for v in (0.0, 0.01, 0.1, 1.0):
        ub = np.full(nr, np.quantile(g, 1 - v) if v > 0 else g.max() + 1.0)
        ho.set_bounds(w.lb, ub); hc.set_bounds(w.lb, ub)
        for x in (x0, 0.5 * x0):
            assert_batches_identical(ho.separate(x), hc.separate(x), f"kind {kind} v {v}")
        assert ho.algorithmic_bytes() == hc.algorithmic_bytes()


@pytest.mark.parametrize("seed", range(4))
def test_random_expression_trees(oracle_lib, cuda_lib, seed):
    rng = np.random.default_rng(100 + seed)
    nvar = 6
    exprs = [random_tree(rng, nvar, int(rng.integers(1, 6))) for _ in range(400)]
    exprs = [e if E.variables(e) else e + E.var(0) for e in exprs]
    m = len(exprs)
    w = E.to_wire(exprs, np.full(m, -np.inf), rng.uniform(-1, 1, m), [ROW_NL] * m)
    ho, hc = both(oracle_lib, cuda_lib, nvar, w)
    rows = np.arange(m, dtype=np.int64)
    for _ in range(3):
        x = np.round(rng.uniform(-2, 2, nvar), 2)
        assert bits_equal(ho.eval_g(x), hc.eval_g(x))
        for r0 in range(0, m, 100):
            sub = rows[r0:r0 + 100]
            while len(sub):
                bo, bc = ho.gencut_rows(x, sub, True), hc.gencut_rows(x, sub, True)
                assert_batches_identical(bo, bc, f"seed {seed} rows {sub[0]}..")
                if bo.status == 0:
                    break
                sub = sub[sub > bo.err_row]
        assert_batches_identical(ho.separate(x), hc.separate(x), f"seed {seed} separate")


def test_error_row_truncation(oracle_lib, cuda_lib):
    """First non-finite row stops the round; cuts before it are delivered (src/model.jl:69-73,278)."""
    x, y, z = E.var(0), E.var(1), E.var(2)
    exprs = [x**2 + y**2 - 1.0] * 40 + [E.sqrt(x**2 + y**2) - (z - 0.25)] + [x**2 + y**2 - 1.0] * 40
    m = len(exprs)
    w = E.to_wire(exprs, np.full(m, -np.inf), np.full(m, -2.0), [ROW_NL] * m)
    ho, hc = both(oracle_lib, cuda_lib, 3, w)
    bo, bc = ho.separate(np.zeros(3)), hc.separate(np.zeros(3))
    assert bo.status == KTN_NUMERIC_NONFINITE and bo.err_row == 40 and bo.n_cuts == 40
    assert_batches_identical(bo, bc)
    # the same truncated round through the zero-copy view: K2 laid the blob out for ALL selected rows, the view holds the first 40
    assert_batches_identical(bo, hc.separate(np.zeros(3), view=True), "view of a truncated round")
    assert_batches_identical(ho.separate(np.ones(3)), hc.separate(np.ones(3)))    # and the state re-arms for the next round


def test_big_shapes_dense_rows_and_reload(oracle_lib, cuda_lib):
    rng = np.random.default_rng(5)
    nvar = 300
    # 200-term rows exceed the shared-memory lane budget -> global-scratch kernel; last row is a dense epigraph row
    exprs = [E.sum_([E.exp(E.const(float(rng.uniform(-1, 1))) * E.var(int(j)) + float(rng.uniform(-1, 1))) for j in rng.choice(nvar, 200, replace=False)]) for _ in range(70)]
    exprs += [x for x in [E.var(0)**2 + E.var(1)] * 5]
    exprs.append(E.sum_([(E.var(j) - 0.5)**2 for j in range(0, nvar - 1, 2)]) - E.var(nvar - 1))
    m = len(exprs)
    w = E.to_wire(exprs, np.full(m, -np.inf), np.full(m, 100.0), [ROW_NL] * (m - 1) + [ROW_NL | ROW_DENSE])
    ho, hc = both(oracle_lib, cuda_lib, nvar, w)
    for _ in range(3):
        xx = rng.uniform(-1, 2, nvar)
        bo = ho.separate(xx)
        assert bo.n_cuts > 0
        assert_batches_identical(bo, hc.separate(xx))
    nv2, w2, pts = kat_problem()                       # initialize! again on the same handle (test/runtests.jl:24)
    ho.load(nv2, w2); hc.load(nv2, w2)
    assert_batches_identical(ho.separate(pts[1]), hc.separate(pts[1]))


@pytest.mark.parametrize("kind,name", [(1, "lse"), (0, "qcqp"), (3, "portfolio")])
def test_baseline_size_round(oracle_lib, cuda_lib, kind, name):
    """BASELINE.json size: 10^6 rows, 10^5 variables (configs[1], [2]; configs[3]: 10^5 SOC-like NL rows among 9 * 10^5 linear rows
    that are never tested, src/model.jl:116-121).  Full comparison with the oracle plus properties."""
    nv, nr = 100000, 1000000
    w = cuda_lib.synth_rows(kind, 20260001 + kind, nv, 0, nr)
    x0 = cuda_lib.synth_point(kind, 20260001 + kind, nv)
    hc = cuda_lib.create(); hc.load(nv, w)
    g = hc.eval_g(x0)
    nl = (w.flags & ROW_NL) != 0
    ub = np.full(nr, np.quantile(g[nl], 0.9))
    hc.set_bounds(w.lb, ub)
    b1 = hc.separate(x0)
    b2 = hc.separate(x0)
    for f in BATCH_FIELDS:                              # idempotence: a round is a pure function of x*
        assert bits_equal(getattr(b1, f), getattr(b2, f)), f
    assert np.all(np.diff(b1.row_id) > 0)              # ascending row order = the reference's loop order
    assert b1.row_ptr[0] == 0 and np.all(np.diff(b1.row_ptr) > 0) and b1.row_ptr[-1] == len(b1.col) == len(b1.val)
    assert abs(b1.n_cuts - 0.1 * nl.sum()) < 0.001 * nr and np.all(nl[b1.row_id])
    assert np.all(b1.g > ub[0] + 1e-6) and np.all(b1.viol == b1.g - ub[0])
    sel = np.zeros(nr, bool); sel[b1.row_id] = True
    assert np.all(g[nl & ~sel] <= ub[0] + 1e-6)        # nothing violated was missed
    # first-order identity: sum_k J_k x*_k + b == g up to rounding of the accumulation
    lin = np.add.reduceat(b1.val * x0[b1.col], b1.row_ptr[:-1]) + b1.bconst
    assert np.allclose(lin, b1.g, rtol=1e-12, atol=1e-12)
    assert np.array_equal(b1.hi, ub[0] - b1.bconst)
    # and the whole round against the oracle
    ho = oracle_lib.create(); ho.load(nv, w); ho.set_bounds(w.lb, ub)
    assert_batches_identical(ho.separate(x0), b1, name)


def test_zero_copy_view_matches_copy(oracle_lib, cuda_lib):
    """ktn_fetch_cuts_view (pinned, library-owned, double-buffered) delivers the same arrays as ktn_fetch_cuts."""
    nv, nr = 3000, 40000
    w = cuda_lib.synth_rows(1, 9, nv, 0, nr); x0 = cuda_lib.synth_point(1, 9, nv)
    ho, hc = both(oracle_lib, cuda_lib, nv, w)
    g = ho.eval_g(x0)
    ub = np.full(nr, np.quantile(g, 0.8)); ho.set_bounds(w.lb, ub); hc.set_bounds(w.lb, ub)
    ref = ho.separate(x0)
    v1 = hc.separate(x0, view=True)
    assert_batches_identical(ref, v1, "view")
    v2 = hc.separate(0.5 * x0, view=True)              # the previous view survives one more fetch (two buffers alternate)
    assert_batches_identical(ref, v1, "view after the next round")
    assert_batches_identical(ho.separate(0.5 * x0), v2, "second view")
    ub0 = np.full(nr, g.max() + 1.0); hc.set_bounds(w.lb, ub0)
    assert hc.separate(x0, view=True).n_cuts == 0      # empty batch
    # lean views (what KatanaGPUSeparator asks for): the LP's arrays are there, the diagnostics are not downloaded
    from katana_jl_b200.binding import FLAG_LEAN_VIEW
    hl = cuda_lib.create(flags=FLAG_LEAN_VIEW); hl.load(nv, w); hl.set_bounds(w.lb, ub)
    vl = hl.separate(x0, view=True)
    for f in ("row_id", "row_ptr", "col", "val", "lo", "hi"):
        assert bits_equal(getattr(vl, f), getattr(ref, f)), f
    assert len(vl.g) == len(vl.viol) == len(vl.bconst) == 0


def test_many_rounds_keep_state_clean(oracle_lib, cuda_lib):
    """600 rounds with changing cut sets on one handle: per-round device state (selection flags, per-block counters by round
    parity, error slots, tickets) is re-armed by the kernels themselves; nothing from an earlier round may resurface."""
    nv, nr = 500, 3000
    w = cuda_lib.synth_rows(0, 5, nv, 0, nr); x0 = cuda_lib.synth_point(0, 5, nv)
    ho, hc = both(oracle_lib, cuda_lib, nv, w)
    g = ho.eval_g(x0)
    ub = np.full(nr, np.quantile(g, 0.7)); ho.set_bounds(w.lb, ub); hc.set_bounds(w.lb, ub)
    scales = [1.0, 0.2, 0.6, 1.3]
    refs = [ho.separate(s * x0) for s in scales]
    assert len({r.n_cuts for r in refs}) > 2
    for it in range(600):
        k = it % len(scales)
        if it % 97 == 0 or it in (254, 255, 256, 257, 509, 510, 511, 512):
            assert_batches_identical(refs[k], hc.separate(scales[k] * x0), f"round {it}")
            assert bits_equal(hc.get_g(), ho.eval_g(scales[k] * x0))
        else:
            hc.separate(scales[k] * x0, fetch=False)


@pytest.mark.parametrize("kind,nv,nr", [(1, 3000, 30000), (0, 400, 9000)])
def test_topk_matches_oracle(oracle_lib, cuda_lib, kind, nv, nr):
    """topk > 0: of the violated rows the k ranked first by (NaN first, violation descending, row ascending) become cuts,
    emitted in ascending row order -- radix select + tie handling on the device against the oracle's sort."""
    w = cuda_lib.synth_rows(kind, 31 + kind, nv, 0, nr); x0 = cuda_lib.synth_point(kind, 31 + kind, nv)
    ho, hc = both(oracle_lib, cuda_lib, nv, w)
    g = ho.eval_g(x0)
    ub = np.full(nr, np.quantile(g, 0.7)); ho.set_bounds(w.lb, ub); hc.set_bounds(w.lb, ub)
    nviol = ho.separate(x0).n_cuts
    for k in (1, 2, 17, 1000, nviol - 1, nviol, nviol + 5, 10 * nr):
        ho.set_params(1e-6, 1e9, k); hc.set_params(1e-6, 1e9, k)
        for x in (x0, 0.7 * x0):
            bo, bc = ho.separate(x), hc.separate(x)
            assert_batches_identical(bo, bc, f"kind {kind} topk {k}")
    ho.set_params(1e-6, 1e9, 0); hc.set_params(1e-6, 1e9, 0)
    assert_batches_identical(ho.separate(x0), hc.separate(x0), "back to all violated rows")


def test_topk_ties_nan_and_errors(oracle_lib, cuda_lib):
    """Ties at the threshold are broken by row index, NaN violations rank first, and the first non-finite row among the
    SELECTED rows ends the batch."""
    x, y, z = E.var(0), E.var(1), E.var(2)
    disk = x**2 + y**2 - 1.0
    exprs = []
    for i in range(6000):                               # many identical rows -> identical violations (ties), a few distinct
        exprs.append(disk if i % 7 else disk + float(i % 5))
    exprs[100] = E.log(z)                               # NaN at z < 0: ranks first, and its cut row is non-finite
    exprs[4000] = E.sqrt(x**2 + y**2) - z               # finite g, non-finite gradient at the origin
    m = len(exprs)
    w = E.to_wire(exprs, np.full(m, -np.inf), np.full(m, -5.0), [ROW_NL] * m)
    ho, hc = both(oracle_lib, cuda_lib, 3, w)
    for k in (1, 3, 50, 857, 858, 859, 5999, 6000, 7000):
        ho.set_params(1e-6, 1e9, k); hc.set_params(1e-6, 1e9, k)
        for pt in ([2.0, 2.0, 1.0], [2.0, 2.0, -1.0], [0.0, 0.0, 3.0], [0.0, 0.0, -1.0]):
            bo, bc = ho.separate(np.array(pt)), hc.separate(np.array(pt))
            assert_batches_identical(bo, bc, f"topk {k} at {pt}")


def test_empty_and_linear_only_problems(oracle_lib, cuda_lib):
    """Edge cases of initialize! / the loop body: no constraints at all, and a model whose rows are all linear
    (nlconstr_ixs empty, src/model.jl:115-122): a round is a no-op that reports zero cuts."""
    x0, x1 = E.var(0), E.var(1)
    empty = E.to_wire([], np.zeros(0), np.zeros(0), [])
    ho, hc = both(oracle_lib, cuda_lib, 2, empty)
    assert_batches_identical(ho.separate(np.array([1.0, 2.0])), hc.separate(np.array([1.0, 2.0])), "no rows")
    assert hc.separate(np.array([1.0, 2.0]), view=True).n_cuts == 0
    lin = E.to_wire([x0 + x1, E.const(2.0) * x0 - x1], [-np.inf, -np.inf], [-5.0, -5.0], [0, 0])
    ho, hc = both(oracle_lib, cuda_lib, 2, lin)
    pt = np.array([3.0, 4.0])
    assert_batches_identical(ho.separate(pt), hc.separate(pt), "linear rows only")
    assert hc.separate(pt).n_cuts == 0
    rows = np.array([0, 1], np.int64)                       # ... but their rows can still be asked for (loadproblem!, src/model.jl:115-118)
    assert_batches_identical(ho.gencut_rows(pt, rows, False), hc.gencut_rows(pt, rows, False), "gencut of linear rows")
    assert bits_equal(ho.eval_g(pt), hc.eval_g(pt))


def test_device_resident_round_and_counters(cuda_lib):
    import torch
    nv, nr = 2000, 50000
    w = cuda_lib.synth_rows(1, 3, nv, 0, nr); x0 = cuda_lib.synth_point(1, 3, nv)
    h = cuda_lib.create(); h.load(nv, w)
    g = h.eval_g(x0); h.set_bounds(w.lb, np.full(nr, np.quantile(g, 0.9)))
    ref = h.separate(x0)
    dx = torch.from_numpy(x0).cuda()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    h.set_stream(s.cuda_stream)
    l0 = h.timings()["launches"]
    h.separate_device_async(dx.data_ptr())
    got = h.fetch_last()
    assert h.timings()["launches"] - l0 == 3            # K1 (evaluate, test) + K2 (compact) + K3 (cuts of the family rows)
    assert_batches_identical(ref, got)
    h.set_stream(0)


def test_ecp_end_to_end_on_gpu(cuda_lib):
    import katana_jl_b200 as K
    from reference_problems import PROBLEMS
    for name, cite, build, obj, sol in PROBLEMS:
        if name.startswith("501") and not name.endswith(("n2", "n5")):
            continue
        m = K.Model(K.KatanaSolver(separator=K.KatanaGPUSeparator(), log_level=0))
        vars_ = build(m)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            status = m.solve()
        assert status == "Optimal", (name, cite)
        assert np.isclose(m.getobjectivevalue(), obj, rtol=1e-6, atol=1e-6), (name, cite)
        if sol is not None:
            assert np.allclose([m.getvalue(v) for v in vars_], sol, rtol=1e-3, atol=1e-3), (name, cite)


@pytest.mark.parametrize("family", ["lse", "quad", "soc"])
def test_family_rows_edge_values_on_gpu(oracle_lib, cuda_lib, family):
    """Device twin of test_compiler_emu.py::test_family_rows_edge_values: the family kernels (K1 forward, K3 cut incl. the exact
    revmul resweep, the rounding sweep, ktn_exp_slow, rows of more than 16 unique variables = the streaming class) on overflow,
    underflow, NaN / inf points, coefficient ranges that make round_coefs (src/model.jl:200-207) zero entries, ragged chunks
    and every unique-variable count 1..20, 40, 70 -- with cut_coef_rng 1e9 and 10."""
    rng = np.random.default_rng(11)
    nvar = 80
    exprs = []
    for nu in list(range(1, 21)) * 3 + [40, 70]:
        cols = rng.choice(nvar, nu, replace=False)
        scale = 10.0 ** rng.integers(-12, 12, nu)
        if family == "lse":
            exprs.append(E.log(E.sum_([E.exp(E.const(float(rng.uniform(-1, 1) * scale[k])) * E.var(int(cols[k])) + float(rng.uniform(-1, 1))) for k in range(nu)])))
        elif family == "quad":
            exprs.append(E.sum_([E.const(float(rng.uniform(0.5, 1.5) * scale[k])) * E.var(int(cols[k]))**2 for k in range(nu)] +
                                [E.const(float(rng.uniform(-1, 1))) * E.var(int(cols[k])) for k in range(nu)]))
        elif nu >= 2:                                                           # sqrt(sum (s x)^2) - t  (test/3d.jl:161)
            exprs.append(E.sqrt(E.sum_([(E.const(float(rng.uniform(0.1, 0.5) * scale[k])) * E.var(int(cols[k])))**2 for k in range(nu - 1)])) - E.var(int(cols[nu - 1])))
    m = len(exprs)
    w = E.to_wire(exprs, np.full(m, -np.inf), np.full(m, -1e300), [ROW_NL] * m)
    pts = [rng.uniform(-2, 2, nvar), np.zeros(nvar), np.full(nvar, 1e3), np.full(nvar, -1e3), np.full(nvar, 1e200), rng.uniform(-1e-9, 1e-9, nvar)]
    p = rng.uniform(-2, 2, nvar); p[3] = np.nan; pts.append(p)
    p = rng.uniform(-2, 2, nvar); p[5] = np.inf; p[7] = -np.inf; pts.append(p)
    ho, hc = both(oracle_lib, cuda_lib, nvar, w)
    ncuts = []
    for rng_coef in (1e9, 10.0):
        ho.set_params(1e-6, rng_coef, 0); hc.set_params(1e-6, rng_coef, 0)
        for x in pts:
            assert bits_equal(ho.eval_g(x), hc.eval_g(x))
            bo, bc = ho.separate(x), hc.separate(x)
            assert_batches_identical(bo, bc, f"{family} rng={rng_coef}")
            ncuts.append(bo.n_cuts)
            # row by row, so that the rows behind the first non-finite cut (which ends a batch, src/model.jl:278) are compared too
            for r0 in range(0, m, 7):
                sub = np.arange(r0, min(m, r0 + 7), dtype=np.int64)
                assert_batches_identical(ho.gencut_rows(x, sub, True), hc.gencut_rows(x, sub, True), f"{family} gencut+round rows {r0}")
    assert max(ncuts) > 0


def test_round_coefs_on_generic_and_big_kernels(oracle_lib, cuda_lib):
    """cut_coef_rng = 10 (round_coefs zeroes small coefficients, src/model.jl:200-207) on the interpreter kernel (generic
    shapes), the global-scratch kernel (200-term rows) and the dense epigraph row (src/nlpeval.jl:49-54)."""
    rng = np.random.default_rng(23)
    nvar = 240
    x = [E.var(j) for j in range(nvar)]
    exprs = []
    for _ in range(150):                                                         # generic shapes: products, divisions, sqrt, powers
        c = rng.choice(nvar, 4, replace=False); s = 10.0 ** rng.integers(-6, 6, 4)
        exprs.append(E.const(float(s[0])) * x[c[0]] * x[c[1]] + E.sqrt(E.const(float(s[1])) * x[c[2]]**2 + 1.0) + E.const(float(s[2])) * x[c[3]]**3 - x[c[0]] / (x[c[1]]**2 + 2.0))
    for _ in range(40):                                                          # 200-term rows: BIG kernel
        exprs.append(E.sum_([E.exp(E.const(float(rng.uniform(-1, 1) * 10.0 ** rng.integers(-4, 3))) * x[int(j)]) for j in rng.choice(nvar, 200, replace=False)]))
    exprs.append(E.sum_([E.const(float(10.0 ** rng.integers(-5, 5))) * (x[j] - 0.5)**2 for j in range(0, nvar - 1, 3)]) - x[nvar - 1])
    m = len(exprs)
    w = E.to_wire(exprs, np.full(m, -np.inf), np.full(m, -1e300), [ROW_NL] * (m - 1) + [ROW_NL | ROW_DENSE])
    ho, hc = both(oracle_lib, cuda_lib, nvar, w)
    zeroed = 0
    for rng_coef in (10.0, 1e3, 1e9):
        ho.set_params(1e-6, rng_coef, 0); hc.set_params(1e-6, rng_coef, 0)
        for _ in range(3):
            xx = rng.uniform(-1.5, 1.5, nvar)
            bo = ho.separate(xx)
            assert_batches_identical(bo, hc.separate(xx), f"rng={rng_coef}")
            zeroed += int(np.sum(bo.val == 0.0))
    assert zeroed > 0                                                            # the rounding sweep did run


def test_single_process_sharded_handle(oracle_lib, cuda_lib):
    """ktn_options.ngpus: ONE handle, rows sharded over several devices (here the same device three times, so that the path runs
    on a one-GPU box), ONE combined batch in ascending row order == the oracle over all rows, bit for bit: separate, views,
    gencut_rows, eval_g / get_g, jac_structure, reload, and the reference's stop at the first non-finite row when that row is
    on a shard that is NOT the last (src/model.jl:69-73, :278)."""
    for kind, nv, nr in ((1, 4000, 9001), (0, 1000, 5000), (2, 2000, 3000)):
        w = cuda_lib.synth_rows(kind, 77 + kind, nv, 0, nr); x0 = cuda_lib.synth_point(kind, 77 + kind, nv)
        ho = oracle_lib.create(); ho.load(nv, w)
        hs = cuda_lib.create(ngpus=3, devices=[0, 0, 0]); hs.load(nv, w)
        assert all(np.array_equal(a, b) for a, b in zip(ho.jac_structure(), hs.jac_structure()))
        g = ho.eval_g(x0)
        assert bits_equal(g, hs.eval_g(x0))
        for v in (0.0, 0.03, 0.5, 1.0):
            ub = np.full(nr, np.quantile(g, 1 - v) if v > 0 else g.max() + 1.0)
            ho.set_bounds(w.lb, ub); hs.set_bounds(w.lb, ub)
            bo = ho.separate(x0)
            assert_batches_identical(bo, hs.separate(x0), f"sharded kind {kind} v {v}")
            assert_batches_identical(bo, hs.separate(x0, view=True), f"sharded view kind {kind} v {v}")
            assert bits_equal(ho.get_g(), hs.get_g())
            assert ho.algorithmic_bytes() == hs.algorithmic_bytes()
        rows = np.unique(np.random.default_rng(kind).integers(0, nr, 300)).astype(np.int64)
        assert_batches_identical(ho.gencut_rows(x0, rows, True), hs.gencut_rows(x0, rows, True), "sharded gencut")
        hs.close(); ho.close()
    # first non-finite row on the middle shard: the cuts of the later shards are dropped, the earlier ones stand
    x, y, z = E.var(0), E.var(1), E.var(2)
    exprs = [x**2 + y**2 - 1.0] * 50 + [E.sqrt(x**2 + y**2) - (z - 0.25)] + [x**2 + y**2 - 1.0] * 39
    m = len(exprs)
    w = E.to_wire(exprs, np.full(m, -np.inf), np.full(m, -2.0), [ROW_NL] * m)
    ho = oracle_lib.create(); ho.load(3, w)
    hs = cuda_lib.create(ngpus=3, devices=[0, 0, 0]); hs.load(3, w)
    bo, bs = ho.separate(np.zeros(3)), hs.separate(np.zeros(3))
    assert bo.status == KTN_NUMERIC_NONFINITE and bo.err_row == 50 and bo.n_cuts == 50
    assert_batches_identical(bo, bs, "sharded, truncated on the middle shard")
    assert_batches_identical(bo, hs.separate(np.zeros(3), view=True), "sharded view, truncated")
    assert_batches_identical(ho.separate(np.ones(3)), hs.separate(np.ones(3)))
    nv2, w2, pts = kat_problem()                                                 # reload on the same (sharded) handle
    ho.load(nv2, w2); hs.load(nv2, w2)
    for p in pts:
        assert_batches_identical(ho.separate(p), hs.separate(p), "sharded KAT")


def test_long_rows_of_family_form_on_gpu(oracle_lib, cuda_lib):
    """Log-sum-exp / quadratic / SOC rows of 60 .. 300 terms: the streaming class of the family kernels (17 .. ~90 variables) and,
    beyond the shared-memory lane budget, the interpreted BIG shapes."""
    rng = np.random.default_rng(3)
    nvar = 400
    exprs = []
    for nu in (17, 40, 60, 91, 150, 256, 300):
        cols = rng.choice(nvar, nu, replace=False)
        exprs.append(E.log(E.sum_([E.exp(E.const(float(rng.uniform(-1, 1))) * E.var(int(c)) + float(rng.uniform(-1, 1))) for c in cols])))
        exprs.append(E.sum_([E.const(float(rng.uniform(0.5, 1.5))) * E.var(int(c))**2 for c in cols] + [E.const(float(rng.uniform(-1, 1))) * E.var(int(c)) for c in cols]))
        exprs.append(E.sqrt(E.sum_([(E.const(float(rng.uniform(0.1, 0.5))) * E.var(int(c)))**2 for c in cols[:-1]])) - E.var(int(cols[-1])))
    m = len(exprs)
    w = E.to_wire(exprs, np.full(m, -np.inf), np.full(m, -1e300), [ROW_NL] * m)
    ho, hc = both(oracle_lib, cuda_lib, nvar, w)
    for rng_coef in (1e9, 10.0):
        ho.set_params(1e-6, rng_coef, 0); hc.set_params(1e-6, rng_coef, 0)
        for _ in range(3):
            x = rng.uniform(-1, 1, nvar)
            assert bits_equal(ho.eval_g(x), hc.eval_g(x))
            assert_batches_identical(ho.separate(x), hc.separate(x), "long family-form rows")


def test_ladder_matches_the_sequential_search(oracle_lib, cuda_lib):
    """ktn_separate_ladder (boundroutine batching, src/model.jl:175-197): the device evaluates the points 2^n * ray in batches and
    cuts at the first violating one; the oracle runs the reference's sequential loop.  Same exponent, same cuts, bit for bit --
    on the KAT problem, on synthetic families (hit in the first batch, in a later batch, never), and with a non-finite cut."""
    nvar, w, pts = kat_problem()
    ho, hc = both(oracle_lib, cuda_lib, nvar, w)
    rng = np.random.default_rng(9)
    for _ in range(4):
        ray = rng.normal(size=nvar)
        (no, bo), (nc, bc) = ho.separate_ladder(ray), hc.separate_ladder(ray)
        assert no == nc
        assert_batches_identical(bo, bc, "ladder KAT")
    for kind, nv, nr, scale in ((0, 1000, 4000, 1.0), (1, 3000, 6000, 1e-7), (2, 2000, 3000, 1e-12)):
        ws = cuda_lib.synth_rows(kind, 5 + kind, nv, 0, nr); x0 = cuda_lib.synth_point(kind, 5 + kind, nv)
        ho, hc = both(oracle_lib, cuda_lib, nv, ws)
        ub = np.full(nr, 50.0); ho.set_bounds(ws.lb, ub); hc.set_bounds(ws.lb, ub)
        ray = x0 * scale                                      # small rays: the first violation comes many doublings out
        (no, bo), (nc, bc) = ho.separate_ladder(ray), hc.separate_ladder(ray)
        assert no == nc and (no >= 2)
        assert_batches_identical(bo, bc, f"ladder kind {kind}")
        (no, bo), (nc, bc) = ho.separate_ladder(ray, 2, 4), hc.separate_ladder(ray, 2, 4)      # a ladder that may end before anything is violated
        assert no == nc
        assert_batches_identical(bo, bc, f"short ladder kind {kind}")
        assert_batches_identical(ho.separate(x0), hc.separate(x0))                              # the handle is usable afterwards
    x, y, z = E.var(0), E.var(1), E.var(2)                    # sqrt at the apex: the first violated row is not finite -> :Error
    exprs = [E.sqrt(x**2 + y**2) - (z - 0.25), x**2 + y**2 - 1.0]
    w = E.to_wire(exprs, np.full(2, -np.inf), np.full(2, -2.0), [ROW_NL] * 2)
    ho, hc = both(oracle_lib, cuda_lib, 3, w)
    (no, bo), (nc, bc) = ho.separate_ladder(np.zeros(3)), hc.separate_ladder(np.zeros(3))
    assert no == nc == 2 and bo.status == KTN_NUMERIC_NONFINITE
    assert_batches_identical(bo, bc, "ladder, non-finite")


def test_ten_million_rows(oracle_lib, cuda_lib):
    """BASELINE.json configs[4], largest size: 10^7 log-sum-exp rows, 10^6 variables, on ONE handle (loaded in ten batches).  Rows are
    independent given x*, so the round's cuts restricted to rows [k 10^6, (k+1) 10^6) must equal, bit for bit, the round of an
    oracle that holds only those rows (same bounds); three such slices are compared in full (first, middle, last), the rest
    through size-independent properties."""
    nv, piece, pieces = 1000000, 1000000, 10
    nr = piece * pieces
    x0 = cuda_lib.synth_point(1, 20260002, nv)
    hc = cuda_lib.create()
    hc.load_begin(nv, nr)
    for k in range(pieces):
        hc.add_rows(k * piece, cuda_lib.synth_rows(1, 20260002, nv, k * piece, piece))
    hc.load_end()
    g = hc.eval_g(x0)
    ubv = float(np.quantile(g, 0.9))
    hc.set_bounds(np.full(nr, -np.inf), np.full(nr, ubv))
    b = hc.separate(x0)
    assert abs(b.n_cuts - 0.1 * nr) < 0.001 * nr and np.all(np.diff(b.row_id) > 0)
    assert b.row_ptr[0] == 0 and np.all(np.diff(b.row_ptr) > 0) and b.row_ptr[-1] == len(b.col) == len(b.val)
    assert np.all(b.g > ubv + 1e-6) and np.array_equal(b.g, g[b.row_id]) and np.array_equal(b.hi, ubv - b.bconst)
    sel = np.zeros(nr, bool); sel[b.row_id] = True
    assert np.all(g[~sel] <= ubv + 1e-6)
    lin = np.add.reduceat(b.val * x0[b.col], b.row_ptr[:-1]) + b.bconst
    assert np.allclose(lin, b.g, rtol=1e-12, atol=1e-12)
    os.environ["KTN_ORACLE_THREADS"] = str(os.cpu_count() or 1)
    try:
        for k in (0, 4, 9):
            w = cuda_lib.synth_rows(1, 20260002, nv, k * piece, piece)
            ho = oracle_lib.create(); ho.load(nv, w); ho.set_bounds(w.lb, np.full(piece, ubv))
            bo = ho.separate(x0)
            c0, c1 = np.searchsorted(b.row_id, [k * piece, (k + 1) * piece])
            assert bo.n_cuts == c1 - c0
            e0, e1 = b.row_ptr[c0], b.row_ptr[c1]
            assert np.array_equal(bo.row_id + k * piece, b.row_id[c0:c1]) and np.array_equal(bo.row_ptr, b.row_ptr[c0:c1 + 1] - e0)
            for f, sl in (("col", slice(e0, e1)), ("val", slice(e0, e1)), ("lo", slice(c0, c1)), ("hi", slice(c0, c1)), ("g", slice(c0, c1)),
                          ("viol", slice(c0, c1)), ("bconst", slice(c0, c1))):
                assert bits_equal(getattr(bo, f), getattr(b, f)[sl]), (k, f)
            assert bits_equal(ho.get_g(), g[k * piece:(k + 1) * piece])
            ho.close()
    finally:
        os.environ.pop("KTN_ORACLE_THREADS", None)


def test_direct_view_cuts_stored_in_host_memory(oracle_lib, cuda_lib):
    """KTN_FLAG_DIRECT_VIEW: K2 / K3 store the batch straight into mapped pinned host memory; views and copies must equal the oracle's
    batch bit for bit -- family rows (LSE, SOC), interpreter rows (KAT problem), full and lean, empty / partial / full rounds, views
    staying valid across the next round, forced rounds (gencut), a truncated round, reload, and the boundroutine ladder."""
    from katana_jl_b200.binding import FLAG_DIRECT_VIEW, FLAG_LEAN_VIEW
    lean_fields = ("row_id", "row_ptr", "col", "val", "lo", "hi")
    for kind, nv, nr in ((1, 4000, 9001), (2, 2000, 3000), (0, 3000, 5000)):
        w = cuda_lib.synth_rows(kind, 71 + kind, nv, 0, nr); x0 = cuda_lib.synth_point(kind, 71 + kind, nv)
        ho = oracle_lib.create(); ho.load(nv, w)
        g = ho.eval_g(x0)
        for flags in (FLAG_DIRECT_VIEW, FLAG_DIRECT_VIEW | FLAG_LEAN_VIEW):
            hd = cuda_lib.create(flags=flags); hd.load(nv, w)
            prev = None
            for v in (0.0, 0.07, 1.0, 0.3):
                ub = np.full(nr, np.quantile(g, 1 - v) if v > 0 else g.max() + 1.0)
                ho.set_bounds(w.lb, ub); hd.set_bounds(w.lb, ub)
                bo = ho.separate(x0)
                bv = hd.separate(x0, view=True)
                assert bo.status == bv.status and bo.err_row == bv.err_row and bo.n_cuts == bv.n_cuts
                for f in (lean_fields if flags & FLAG_LEAN_VIEW else BATCH_FIELDS):
                    assert bits_equal(getattr(bo, f), getattr(bv, f)), (kind, flags, v, f)
                if prev is not None:       # the view of the round before is still intact (two host buffers alternate)
                    for f in lean_fields:
                        assert bits_equal(getattr(prev[0], f), getattr(prev[1], f)), ("previous view", kind, flags, v, f)
                prev = (bo, bv)
            if not flags & FLAG_LEAN_VIEW:
                assert_batches_identical(bo, hd.separate(x0), f"direct, copy, kind {kind}")
            rows = np.arange(0, nr, 17, dtype=np.int64)
            assert_batches_identical(ho.gencut_rows(x0, rows, True), hd.gencut_rows(x0, rows, True), "direct handle, gencut")
            hd.close()
        ho.close()
    x, y, z = E.var(0), E.var(1), E.var(2)
    exprs = [x**2 + y**2 - 1.0] * 50 + [E.sqrt(x**2 + y**2) - (z - 0.25)] + [x**2 + y**2 - 1.0] * 39
    m = len(exprs)
    w = E.to_wire(exprs, np.full(m, -np.inf), np.full(m, -2.0), [ROW_NL] * m)
    ho = oracle_lib.create(); ho.load(3, w)
    hd = cuda_lib.create(flags=FLAG_DIRECT_VIEW); hd.load(3, w)
    bo = ho.separate(np.zeros(3))
    assert bo.status == KTN_NUMERIC_NONFINITE and bo.n_cuts == 50
    assert_batches_identical(bo, hd.separate(np.zeros(3), view=True), "direct view, truncated round")
    assert_batches_identical(bo, hd.separate(np.zeros(3)), "direct copy, truncated round")
    assert_batches_identical(ho.separate(np.ones(3)), hd.separate(np.ones(3), view=True))
    nv2, w2, pts = kat_problem()
    ho.load(nv2, w2); hd.load(nv2, w2)
    for p in pts:
        assert_batches_identical(ho.separate(p), hd.separate(p, view=True), "direct KAT")
    ray = np.array([0.25, -0.5, 0.125] + [0.0] * (nv2 - 3))[:nv2]
    no, bo = ho.separate_ladder(ray, 2, 40)
    nd, bd = hd.separate_ladder(ray, 2, 40, view=True)
    assert no == nd
    assert_batches_identical(bo, bd, "direct ladder")
    ho.close(); hd.close()


def test_sin_cos_edge_values_on_gpu(oracle_lib, cuda_lib):
    """sin / cos rows on the device at the edges of the shared implementation (csrc/ktn_math.h): zeros of either sign, multiples of
    pi/2, the end of the exact-reduction range (2^20 pi/2), arguments up to 2^44 (reduced with growing error, identically on both
    sides), 2^45 and beyond / inf / NaN (NaN: the row is reported as not finite), with and without coefficient rounding."""
    x, y, z = E.var(0), E.var(1), E.var(2)
    exprs = [E.sin(x), E.cos(x), E.sin(x) * E.cos(y) + z, E.cos(x * y) - E.sin(z / 3.0), E.sin(E.cos(x)) + y * y, E.exp(E.sin(x)) + E.cos(y)**2,
             x * E.sin(y) - z * E.cos(x), E.sin(x + y + z) + E.cos(x - y)] * 40          # enough rows of each shape to fill warps
    m = len(exprs)
    w = E.to_wire(exprs, np.full(m, -np.inf), np.full(m, -5.0), [ROW_NL] * m)              # every finite row is violated
    pts = [np.zeros(3), np.array([-0.0, 0.0, -0.0]), np.array([np.pi / 2, np.pi, -np.pi / 2]), np.array([1647099.0, -1647099.5, 3.0]),
           np.array([1e9, -1e12, 2.0**44]), np.array([1e-300, 5e-324, -1e-308]), np.array([0.5, 2.0**45, 1.0]), np.array([np.inf, 1.0, 1.0]),
           np.array([1.0, np.nan, 1.0]), np.array([1.0, 1.0, -np.inf]), np.array([355.0, 22.0 / 7.0, 1e5])]
    for rng_ in (1e9, 10.0):
        ho, hc = both(oracle_lib, cuda_lib, 3, w, cut_coef_rng=rng_)
        for p_ in pts:
            bo, bc = ho.separate(p_), hc.separate(p_)
            assert_batches_identical(bo, bc, f"sin/cos at {p_}, rng {rng_}")
            assert bits_equal(ho.eval_g(p_), hc.eval_g(p_))
        rows = np.arange(0, m, 3, dtype=np.int64)
        assert_batches_identical(ho.gencut_rows(pts[3], rows, True), hc.gencut_rows(pts[3], rows, True))
        ho.close(); hc.close()
    ho, hc = both(oracle_lib, cuda_lib, 3, w)
    assert ho.separate(pts[6]).status == KTN_NUMERIC_NONFINITE           # 2^45: outside the supported range on both sides


@pytest.mark.parametrize("hostpush", ["1", "0"])
def test_eager_view_pipelined_download(oracle_lib, cuda_lib, monkeypatch, hostpush):
    """KTN_FLAG_EAGER_VIEW: ktn_separate starts every shard's cut download when that shard has finished, into a pinned buffer laid out
    for the worst case; the view (and the copy) must equal the oracle's batch bit for bit -- full and lean views, empty and full
    rounds, several rounds in a row (the two pinned buffers alternate), a non-finite row on the middle shard, and reload.
    hostpush 1 (default when every shard is on one device): a kernel per shard stores the cuts into the mapped pinned batch, one
    host synchronisation per round; 0: copy-engine downloads started by the host shard by shard (the multi-device path)."""
    monkeypatch.setenv("KTN_HOSTPUSH", hostpush)
    from katana_jl_b200.binding import FLAG_EAGER_VIEW, FLAG_LEAN_VIEW
    lean_fields = ("row_id", "row_ptr", "col", "val", "lo", "hi")
    for kind, nv, nr in ((1, 4000, 9001), (2, 2000, 3000)):
        w = cuda_lib.synth_rows(kind, 91 + kind, nv, 0, nr); x0 = cuda_lib.synth_point(kind, 91 + kind, nv)
        ho = oracle_lib.create(); ho.load(nv, w)
        g = ho.eval_g(x0)
        for flags in (FLAG_EAGER_VIEW, FLAG_EAGER_VIEW | FLAG_LEAN_VIEW):
            hs = cuda_lib.create(ngpus=4, devices=[0, 0, 0, 0], flags=flags); hs.load(nv, w)
            for v in (0.0, 0.07, 1.0, 0.3):
                ub = np.full(nr, np.quantile(g, 1 - v) if v > 0 else g.max() + 1.0)
                ho.set_bounds(w.lb, ub); hs.set_bounds(w.lb, ub)
                bo = ho.separate(x0)
                bv = hs.separate(x0, view=True)
                assert bo.status == bv.status and bo.err_row == bv.err_row and bo.n_cuts == bv.n_cuts
                for f in (lean_fields if flags & FLAG_LEAN_VIEW else BATCH_FIELDS):
                    assert bits_equal(getattr(bo, f), getattr(bv, f)), (kind, flags, v, f)
                assert_batches_identical(bo, hs.separate(x0), f"eager, copy, kind {kind} v {v}")
            rows = np.arange(0, nr, 17, dtype=np.int64)
            assert_batches_identical(ho.gencut_rows(x0, rows, True), hs.gencut_rows(x0, rows, True), "eager handle, gencut")
            hs.close()
        ho.close()
    x, y, z = E.var(0), E.var(1), E.var(2)
    exprs = [x**2 + y**2 - 1.0] * 50 + [E.sqrt(x**2 + y**2) - (z - 0.25)] + [x**2 + y**2 - 1.0] * 39
    m = len(exprs)
    w = E.to_wire(exprs, np.full(m, -np.inf), np.full(m, -2.0), [ROW_NL] * m)
    ho = oracle_lib.create(); ho.load(3, w)
    hs = cuda_lib.create(ngpus=3, devices=[0, 0, 0], flags=FLAG_EAGER_VIEW); hs.load(3, w)
    bo = ho.separate(np.zeros(3))
    assert bo.status == KTN_NUMERIC_NONFINITE and bo.n_cuts == 50
    assert_batches_identical(bo, hs.separate(np.zeros(3), view=True), "eager view, truncated on the middle shard")
    assert_batches_identical(ho.separate(np.ones(3)), hs.separate(np.ones(3), view=True))
    nv2, w2, pts = kat_problem()
    ho.load(nv2, w2); hs.load(nv2, w2)
    for p in pts:
        assert_batches_identical(ho.separate(p), hs.separate(p, view=True), "eager KAT")
