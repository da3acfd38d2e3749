"""End-to-end ECP solves of the reference's test problems through the Katana-mirroring host API
(KatanaSolver / loadproblem! / optimize! / getobjval / getsolution) with the reference's tolerances
(test/runtests.jl:16-20).  On the CPU the separator's backend is the oracle (as the checker of the host
logic); the GPU-marked twin in test_gpu_parity.py runs the same problems on the CUDA library."""
import warnings

import numpy as np
import pytest

import katana_jl_b200 as K
from reference_problems import PROBLEMS

OPT_TOL, SOL_TOL = 1e-6, 1e-3


def solve_problem(lib, build, **kw):
    m = K.Model(K.KatanaSolver(separator=K.KatanaGPUSeparator(library=lib), log_level=0, **kw))
    vars_ = build(m)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        status = m.solve()
    return m, vars_, status


@pytest.mark.parametrize("name,cite,build,obj,sol", PROBLEMS, ids=[p[0] for p in PROBLEMS])
def test_reference_optimum(oracle_lib, name, cite, build, obj, sol):
    m, vars_, status = solve_problem(oracle_lib, build)
    assert status == "Optimal", (name, cite, status)
    assert np.isclose(m.getobjectivevalue(), obj, rtol=OPT_TOL, atol=OPT_TOL), (name, cite, m.getobjectivevalue(), obj)
    if sol is not None:
        got = [m.getvalue(v) for v in vars_]
        assert np.allclose(got, sol, rtol=SOL_TOL, atol=SOL_TOL), (name, cite, got, sol)


def test_linear_problem_generates_no_nl_cuts(oracle_lib):
    """The TODO of test/lpqp.jl:26: a pure LP needs no separation round beyond the first."""
    m, _, status = solve_problem(oracle_lib, PROBLEMS[0][2])
    assert status == "Optimal" and m.internal.iter == 1
    assert m.internal.numcuts == 5          # the five linear rows copied as cuts at 0 (src/model.jl:115-118)


def test_per_row_hooks_match_batched_path(oracle_lib):
    """isconstrsat / gencut (reference per-row API) give the same cuts as the batched round."""
    from katana_jl_b200 import expr as E
    from katana_jl_b200.model import round_coefs
    m, vars_, _ = solve_problem(oracle_lib, PROBLEMS[4][2])
    sep = m.internal.params.separator
    x = np.array([1.5, -0.3])
    batch = sep.separate(x)
    assert batch.n_cuts == 1
    assert not sep.isconstrsat(0, -np.inf, 1.0, 1e-6)
    cut = sep.gencut(x, (-np.inf, 1.0), 0)
    round_coefs(cut, 1e9)
    assert list(cut.vars) == list(batch.col) and list(cut.coeffs) == list(batch.val)
    assert 1.0 - cut.constant == batch.hi[0]


def test_iteration_cap_and_status(oracle_lib):
    m, _, status = solve_problem(oracle_lib, PROBLEMS[4][2], iter_cap=3)
    assert status == "UserLimit" and m.internal.numiters() == 3


def test_error_status_on_undefined_gradient(oracle_lib):
    """sqrt-cone started at the origin: NaN gradient -> :Error (src/model.jl:69-73,278)."""
    from katana_jl_b200 import expr as E

    def build(m):
        x, y, z = m.variable(-1, 1), m.variable(-1, 1), m.variable(0, 0)
        m.objective("Min", x + y + z)
        m.nlconstraint(E.sqrt((x + 1)**2 + (y + 1)**2), "<=", z - 0.25)
        return [x, y, z]
    m, _, status = solve_problem(oracle_lib, build)
    assert status == "Error"


def test_cut_management_keeps_the_optimum_with_a_smaller_lp(oracle_lib):
    """Cut purging and duplicate filtering (extensions, off by default -- the reference never removes a cut,
    src/model.jl:215): same optima within the reference's tolerances, fewer rows left in the LP master, every removal counted."""
    by_name = {p[0]: p for p in PROBLEMS}
    shrunk = rows0 = rows1 = 0
    for name in ("101_01", "103_03", "105_01", "202_03", "210_02", "501_01_n8", "501_02_n13"):
        _, cite, build, obj, sol = by_name[name]
        m0, _, s0 = solve_problem(oracle_lib, build)
        m1, v1, s1 = solve_problem(oracle_lib, build, cut_purge_age=2, cut_filter_duplicates=True)
        assert s0 == s1 == "Optimal", (name, s0, s1)
        assert np.isclose(m1.getobjectivevalue(), obj, rtol=OPT_TOL, atol=OPT_TOL), (name, m1.getobjectivevalue(), obj)
        if sol is not None:
            assert np.allclose([m1.getvalue(v) for v in v1], sol, rtol=SOL_TOL, atol=SOL_TOL), name
        k0, k1 = m0.internal, m1.internal
        assert len(k1.linear_model.rows) == len(k0.linear_model.rows) - (k0.numcuts - k1.numcuts) - k1.cuts_purged
        rows0 += len(k0.linear_model.rows); rows1 += len(k1.linear_model.rows)       # (a single problem may take a few more rounds)
        shrunk += k1.cuts_purged
    assert shrunk > 0 and rows1 < rows0 / 1.5


def test_duplicate_cuts_are_filtered(oracle_lib):
    """The same constraint stated twice yields bit-identical cuts every round: the filter adds each of them once."""
    def build(m):
        x, y = m.variable(-2, 2), m.variable(-2, 2)
        m.objective("Min", -x - y)
        m.nlconstraint(x**2 + y**2, "<=", 1.0); m.nlconstraint(x**2 + y**2, "<=", 1.0)
        return [x, y]
    m0, _, s0 = solve_problem(oracle_lib, build)
    m1, _, s1 = solve_problem(oracle_lib, build, cut_filter_duplicates=True)
    assert s0 == s1 == "Optimal"
    assert np.isclose(m0.getobjectivevalue(), m1.getobjectivevalue(), rtol=1e-9)
    k0, k1 = m0.internal, m1.internal
    assert k1.cuts_filtered > 0 and k1.numcuts + k1.cuts_filtered == k0.numcuts
    assert k0.iter == k1.iter                 # the LP is the same polyhedron: same iterates


def test_models_without_nl_macros_take_the_lpqp_bridge(oracle_lib):
    """src/solver.jl:46: LP / QP / QCQP models reach Katana through NonlinearToLPQPBridge.  The mirror captures A and Q from the
    LinearQuadratic calls and hands them out as expression graphs; MathProgBase's conventions (0.5 x'Qx with one triangle given,
    addquadconstr! entries as given, no objective constant in the LP form) are checked on a worked example."""
    from katana_jl_b200 import expr as E
    from katana_jl_b200.lpqp import LPQPEvaluator, NonlinearToLPQPBridge
    from katana_jl_b200.solver import LinearQuadraticModel, getKatanaModel
    m = K.Model(K.KatanaSolver(separator=K.KatanaGPUSeparator(library=oracle_lib), log_level=0))
    x, y = m.variable(-2, 2), m.variable(-2, 2)
    m.objective("Min", 3 * x**2 + x * y + 2 * y**2 - x + 7.0)
    m.constraint(x + 2 * y, ">=", -1.0); m.constraint(x**2 + 2 * x * y + 4 * y**2 + y, "<=", 3.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert m.solve() == "Optimal"
    b = m.bridge
    assert isinstance(b, NonlinearToLPQPBridge) and getKatanaModel(b) is b.nlpmodel is m.internal
    assert isinstance(LinearQuadraticModel(m.solver), NonlinearToLPQPBridge)
    assert b.qobj == ([0, 0, 1], [0, 1, 1], [6.0, 1.0, 4.0])                     # 0.5 x'Qx: diagonal entries doubled
    assert b.qcons == [([0, 1], [0.0, 1.0], [0, 0, 1], [0, 1, 1], [1.0, 2.0, 4.0])] or b.qcons == [([1], [1.0], [0, 0, 1], [0, 1, 1], [1.0, 2.0, 4.0])]
    assert b.qbounds == [(-np.inf, 3.0)] and list(b.c) == [-1.0, 0.0]
    d = LPQPEvaluator(2, b.A_rows, b.c, b.qobj, b.qcons)
    assert d.features_available() == ["ExprGraph"] and d.isconstrlinear(0) and not d.isconstrlinear(1) and not d.isobjlinear()
    p = np.array([0.3, -0.7])
    assert np.isclose(d.eval_f(p), 3 * 0.09 + 0.3 * -0.7 + 2 * 0.49 - 0.3)       # no constant: JuMP adds the 7 back
    assert np.isclose(E.evaluate(d.constr_expr(1), p), 0.09 + 2 * 0.3 * -0.7 + 4 * 0.49 - 0.7)
    assert np.isclose(E.evaluate(d.constr_expr(0), p), 0.3 - 1.4)
    # unconstrained minimiser of the objective (13/23, -... ) is feasible here: check against the closed form
    H = np.array([[6.0, 1.0], [1.0, 4.0]]); xs = np.linalg.solve(H, [1.0, 0.0])
    assert np.isclose(m.getobjectivevalue(), 0.5 * xs @ H @ xs - xs[0] + 7.0, rtol=1e-5, atol=1e-5)
    # a model WITH an @NL macro stays on the NonlinearModel route
    m2 = K.Model(K.KatanaSolver(separator=K.KatanaGPUSeparator(library=oracle_lib), log_level=0))
    u = m2.variable(-1, 1); m2.objective("Min", -u); m2.nlconstraint(u**2, "<=", 0.25)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert m2.solve() == "Optimal" and not hasattr(m2, "bridge")


def test_pipelined_separator_bounds_an_unbounded_lp_without_the_ladder_call(oracle_lib):
    """A separator whose handle is a group of shards (pipeline > 1, several devices) has no ktn_separate_ladder: boundroutine
    (src/model.jl:175-197) falls back to the reference's sequential search and reaches the same optimum."""
    by_name = {p[0]: p for p in PROBLEMS}
    _, cite, build, obj, sol = by_name["501_01_n3"]
    sep = K.KatanaGPUSeparator(library=oracle_lib, pipeline=2)
    m = K.Model(K.KatanaSolver(separator=sep, log_level=0))
    vars_ = build(m)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert m.solve() == "Optimal"
    assert not sep.has_ladder and sep.handle_options(10)["ngpus"] == 2
    assert not any("ladder_hit" in r for r in m.internal.round_log)         # the sequential path was taken
    assert np.isclose(m.getobjectivevalue(), obj, rtol=OPT_TOL, atol=OPT_TOL)
    m2, _, s2 = solve_problem(oracle_lib, build)
    assert s2 == "Optimal" and any("ladder_hit" in r for r in m2.internal.round_log)      # a plain handle uses the one-call ladder
