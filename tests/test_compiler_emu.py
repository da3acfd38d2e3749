"""The tape compiler + chunk packing + interpreter core (the exact artefacts and code the sm_100a
kernels run) against the oracle, bit for bit, on the CPU through the test-only host emulator."""
import numpy as np
import pytest

from helpers import assert_batches_identical, bits_equal, kat_problem, random_tree
from katana_jl_b200 import expr as E
from katana_jl_b200.binding import ROW_DENSE, ROW_NL


def both(oracle_lib, emu_lib, nvar, w, **kw):
    ho, he = oracle_lib.create(**kw), emu_lib.create(**kw)
    ho.load(nvar, w); he.load(nvar, w)
    return ho, he


def test_jac_structure_sorted_unique(oracle_lib, emu_lib):
    nvar, w, _ = kat_problem()
    ho, he = both(oracle_lib, emu_lib, nvar, w)
    (rp0, c0), (rp1, c1) = ho.jac_structure(), he.jac_structure()
    assert np.array_equal(rp0, rp1) and np.array_equal(c0, c1)
    for r in range(w.nrows):
        cols = c0[rp0[r]:rp0[r + 1]]
        assert np.all(np.diff(cols) > 0)


def test_kat_rounds_identical(oracle_lib, emu_lib):
    nvar, w, pts = kat_problem()
    ho, he = both(oracle_lib, emu_lib, nvar, w)
    for p in pts:
        assert_batches_identical(ho.separate(p), he.separate(p), f"separate at {p}")
        assert bits_equal(ho.eval_g(p), he.eval_g(p))
        rows = np.arange(w.nrows, dtype=np.int64)
        assert_batches_identical(ho.gencut_rows(p, rows, False), he.gencut_rows(p, rows, False), f"gencut at {p}")
        assert_batches_identical(ho.gencut_rows(p, rows[::3], True), he.gencut_rows(p, rows[::3], True), f"gencut+round at {p}")


@pytest.mark.parametrize("kind,nv,nr", [(0, 1000, 3000), (1, 5000, 6000), (2, 2000, 2500)])
def test_synthetic_families_identical(oracle_lib, emu_lib, synth_lib, kind, nv, nr):
    w = synth_lib.synth_rows(kind, 20260001 + kind, nv, 0, nr)
    x0 = synth_lib.synth_point(kind, 20260001 + kind, nv)
    ho, he = both(oracle_lib, emu_lib, nv, w)
    g = ho.eval_g(x0)
    assert bits_equal(g, he.eval_g(x0))
    for v in (0.0, 0.01, 0.1, 1.0):
        ub = np.full(nr, np.quantile(g, 1 - v) if v > 0 else g.max() + 1.0)
        ho.set_bounds(w.lb, ub); he.set_bounds(w.lb, ub)
        bo, be = ho.separate(x0), he.separate(x0)
        assert_batches_identical(bo, be, f"kind {kind} v {v}")
        assert abs(bo.n_cuts - v * nr) <= max(2, 0.01 * nr)
        assert ho.algorithmic_bytes() == he.algorithmic_bytes()


def test_family_detection(emu_lib, synth_lib):
    """The three benchmark forms are recognised as families (interpreter-free kernels)."""
    import ctypes as C
    emu_lib.dll.ktn_emu_num_family_chunks.restype = C.c_int64
    for kind, fam in ((0, 2), (1, 1), (2, 3)):
        w = synth_lib.synth_rows(kind, 7, 1000, 0, 700)
        h = emu_lib.create(); h.load(1000, w)
        nch = emu_lib.dll.ktn_emu_num_chunks(h.h)
        assert emu_lib.dll.ktn_emu_num_family_chunks(h.h, C.c_int32(fam)) == nch
    # a row that is not in nlconstr_ixs never takes a family kernel; a 300-term row exceeds the family's 1-byte sort order
    x = [E.var(j) for j in range(4)]
    quad = E.sum_([E.const(1.5) * x[j]**2 for j in range(4)] + [E.const(0.5) * x[j] for j in range(4)])
    w = E.to_wire([quad, quad], [-np.inf] * 2, [1.0] * 2, [ROW_NL, 0])
    h = emu_lib.create(); h.load(4, w)
    assert emu_lib.dll.ktn_emu_num_family_chunks(h.h, C.c_int32(2)) == 1 and emu_lib.dll.ktn_emu_num_family_chunks(h.h, C.c_int32(0)) == 1


@pytest.mark.parametrize("family", ["lse", "quad", "soc"])
def test_family_rows_edge_values(oracle_lib, emu_lib, family, monkeypatch):
    """Family fast path == interpreter == oracle on overflow, underflow, NaN / inf points, repeated coefficients,
    coefficient ranges that make round_coefs zero entries, ragged chunks and every unique-variable count 1..20."""
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
    results = []
    for no_family in ("", "1"):
        if no_family:
            monkeypatch.setenv("KTN_NO_FAMILY", "1")
        ho, he = both(oracle_lib, emu_lib, nvar, w)
        for rng_coef in (1e9, 10.0):
            ho.set_params(1e-6, rng_coef, 0); he.set_params(1e-6, rng_coef, 0)
            for x in pts:
                assert bits_equal(ho.eval_g(x), he.eval_g(x))
                bo, be = ho.separate(x), he.separate(x)
                assert_batches_identical(bo, be, f"{family} no_family={no_family!r} rng={rng_coef}")
                results.append(bo.n_cuts)
                sub = np.arange(1, m, 3, dtype=np.int64)                      # forced cuts (loadproblem! / boundroutine path)
                assert_batches_identical(ho.gencut_rows(x, sub, False), he.gencut_rows(x, sub, False), f"{family} gencut")
    assert max(results) > 0


def test_fused_shapes_use_few_instructions(emu_lib, synth_lib):
    """The benchmark families compile to fused term runs (a handful of instructions per shape)."""
    import ctypes as C
    for kind in (0, 1, 2):
        w = synth_lib.synth_rows(kind, 7, 1000, 0, 256)
        h = emu_lib.create(); h.load(1000, w)
        for sid in range(emu_lib.dll.ktn_emu_num_shapes(h.h)):
            info = np.zeros(9, np.uint32)
            emu_lib.dll.ktn_emu_shape_info(h.h, C.c_int64(sid), info.ctypes.data_as(C.c_void_p))
            assert info[1] <= 16, (kind, sid, info)


@pytest.mark.parametrize("seed", range(6))
def test_random_expression_trees(oracle_lib, emu_lib, seed):
    """Every operator, arbitrary nesting, repeated variables, NaN / inf producing points."""
    rng = np.random.default_rng(seed)
    nvar = 6
    exprs = [random_tree(rng, nvar, int(rng.integers(1, 6))) for _ in range(300)]
    exprs = [e if E.variables(e) else e + E.var(0) for e in exprs]
    m = len(exprs)
    w = E.to_wire(exprs, np.full(m, -np.inf), rng.uniform(-1, 1, m), [ROW_NL] * m)
    ho, he = both(oracle_lib, emu_lib, nvar, w)
    for _ in range(4):
        x = np.round(rng.uniform(-2, 2, nvar), 2)
        rows = np.arange(m, dtype=np.int64)
        assert bits_equal(ho.eval_g(x), he.eval_g(x))
        # row by row, so non-finite rows (which stop a batch, src/model.jl:278) do not hide later rows
        for r0 in range(0, m, 50):
            sub = rows[r0:r0 + 50]
            bo, be = ho.gencut_rows(x, sub, True), he.gencut_rows(x, sub, True)
            while True:
                assert_batches_identical(bo, be, f"seed {seed} rows {sub[0]}..")
                if bo.status == 0 or bo.err_row == sub[-1]:
                    break
                sub = sub[sub > bo.err_row]
                bo, be = ho.gencut_rows(x, sub, True), he.gencut_rows(x, sub, True)
        assert_batches_identical(ho.separate(x), he.separate(x), f"seed {seed} separate")


def test_edge_cases(oracle_lib, emu_lib):
    x0, x1 = E.var(0), E.var(1)
    # single row; row touching one variable many times; constant-only subtree; ragged chunk (33 rows of one shape)
    exprs = [x0 * x0 * x0 + x0 / x0 - x0, (E.const(2.0) * E.const(3.0)) * x1] + [E.const(float(k)) * x0**2 + x1 for k in range(33)]
    m = len(exprs)
    w = E.to_wire(exprs, np.full(m, -np.inf), np.zeros(m), [ROW_NL] * m)
    ho, he = both(oracle_lib, emu_lib, 2, w)
    for x in ([1.0, 1.0], [0.0, -1.0], [-3.0, 2.0]):
        assert_batches_identical(ho.separate(np.array(x)), he.separate(np.array(x)), str(x))
    # rows that are not NL are never selected; two-sided bounds; equality rows
    w2 = E.to_wire([x0 + x1, x0**2, x1**2], [0.0, -np.inf, 1.0], [1.0, 4.0, 1.0], [0, ROW_NL, ROW_NL])
    ho, he = both(oracle_lib, emu_lib, 2, w2)
    for x in ([5.0, 5.0], [1.0, 1.0], [3.0, 0.5]):
        bo = ho.separate(np.array(x)); assert_batches_identical(bo, he.separate(np.array(x)), str(x))
        assert 0 not in bo.row_id


def test_big_shapes_and_dense_rows(oracle_lib, emu_lib, monkeypatch):
    """Shapes over the shared-memory budget and the dense epigraph row take the global-scratch path."""
    monkeypatch.setenv("KTN_EMU_LANE_LIMIT", "200")
    rng = np.random.default_rng(5)
    nvar = 40
    exprs = [E.sum_([E.exp(E.const(float(rng.uniform(-1, 1))) * E.var(int(j)) + float(rng.uniform(-1, 1))) for j in rng.choice(nvar, 30, replace=False)]) for _ in range(70)]
    exprs.append(E.sum_([(E.var(j) - 0.5)**2 for j in range(0, nvar - 1, 2)]) - E.var(nvar - 1))
    m = len(exprs)
    w = E.to_wire(exprs, np.full(m, -np.inf), np.full(m, 20.0), [ROW_NL] * (m - 1) + [ROW_NL | ROW_DENSE])
    ho, he = both(oracle_lib, emu_lib, nvar, w)
    assert emu_lib.dll.ktn_emu_num_big_chunks(he.h) >= 3
    for _ in range(3):
        x = rng.uniform(-1, 2, nvar)
        bo = ho.separate(x)
        assert_batches_identical(bo, he.separate(x))
        assert bo.n_cuts > 0 and bo.row_ptr[-1] >= nvar


def test_reload_on_same_handle(oracle_lib, emu_lib):
    """initialize! is called repeatedly on one separator (test/runtests.jl:24)."""
    nvar, w, pts = kat_problem()
    he, ho = emu_lib.create(), oracle_lib.create()
    for _ in range(2):
        he.load(nvar, w); ho.load(nvar, w)
        assert_batches_identical(ho.separate(pts[1]), he.separate(pts[1]))
    x0 = E.var(0)
    w2 = E.to_wire([x0**2], [-np.inf], [1.0], [ROW_NL])
    he.load(1, w2); ho.load(1, w2)
    assert_batches_identical(ho.separate(np.array([3.0])), he.separate(np.array([3.0])))


def test_long_rows_of_family_form(oracle_lib, emu_lib):
    """Log-sum-exp / quadratic / SOC rows too long for the shared-memory lane budget run as interpreted BIG shapes (their chunks
    carry the interpreter's sort order, not the families' rank bytes)."""
    rng = np.random.default_rng(3)
    nvar = 400
    exprs = []
    for nu in (60, 91, 150, 256, 300):
        cols = rng.choice(nvar, nu, replace=False)
        exprs.append(E.log(E.sum_([E.exp(E.const(float(rng.uniform(-1, 1))) * E.var(int(c)) + float(rng.uniform(-1, 1))) for c in cols])))
        exprs.append(E.sum_([E.const(float(rng.uniform(0.5, 1.5))) * E.var(int(c))**2 for c in cols] + [E.const(float(rng.uniform(-1, 1))) * E.var(int(c)) for c in cols]))
        exprs.append(E.sqrt(E.sum_([(E.const(float(rng.uniform(0.1, 0.5))) * E.var(int(c)))**2 for c in cols[:-1]])) - E.var(int(cols[-1])))
    m = len(exprs)
    w = E.to_wire(exprs, np.full(m, -np.inf), np.full(m, -1e300), [ROW_NL] * m)
    ho, he = both(oracle_lib, emu_lib, nvar, w)
    for _ in range(3):
        x = rng.uniform(-1, 1, nvar)
        assert bits_equal(ho.eval_g(x), he.eval_g(x))
        assert_batches_identical(ho.separate(x), he.separate(x), "long family-form rows")
