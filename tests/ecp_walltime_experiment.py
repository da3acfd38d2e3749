"""ECP wall time with and without cut management, host mirror + CPU restatement as the separator backend (test infrastructure: the
LP master dominates the wall time).  python tests/ecp_walltime_experiment.py NVAR NROWS   (e.g. 200 3000)"""
import sys, time, warnings
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import katana_jl_b200 as K
from katana_jl_b200 import expr as E
from katana_jl_b200.binding import KtnLibrary
lib = KtnLibrary(os.path.join(ROOT, 'oracle', 'libktn_oracle.so'))
def build(m, nvar, nrows, k, seed):
    rng = np.random.default_rng(seed)
    xs = m.variables(nvar, -10.0, 10.0)
    c = rng.normal(size=nvar)
    m.objective("Min", E.sum_([float(c[j]) * xs[j] for j in range(nvar)]))
    for i in range(nrows):      # convex quadratic rows: sum a_j (x_j - d_j)^2 <= r
        cols = rng.choice(nvar, k, replace=False)
        a = rng.uniform(0.5, 2.0, k); d = rng.uniform(-1, 1, k)
        m.nlconstraint(E.sum_([float(a[t]) * (xs[int(cols[t])] - float(d[t]))**2 for t in range(k)]), "<=", float(rng.uniform(20.0, 60.0)))
    return xs
nvar, nrows, k = int(sys.argv[1]), int(sys.argv[2]), 8
for kw in ({}, dict(cut_purge_age=3), dict(cut_purge_age=3, cut_filter_duplicates=True)):
    m = K.Model(K.KatanaSolver(separator=K.KatanaGPUSeparator(library=lib), log_level=0, f_tol=1e-6, iter_cap=400, **kw))
    build(m, nvar, nrows, k, 1)
    t0 = time.time()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore"); st = m.solve()
    dt = time.time() - t0
    km = m.internal
    lp_s = sum(r.get("lp_s", 0) for r in km.round_log); sep_s = sum(r.get("separate_s", 0) for r in km.round_log); add_s = sum(r.get("addconstr_s", 0) for r in km.round_log)
    print(kw or "plain", st, "obj %.8f" % m.getobjectivevalue(), "iters", km.iter, "cuts", km.numcuts, "LP rows", km.linear_model.numrows, "purged", km.cuts_purged,
          "wall %.1f s (LP %.1f, separation %.2f, hand-off %.2f)" % (dt, lp_s, sep_s, add_s), flush=True)
