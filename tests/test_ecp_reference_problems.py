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
