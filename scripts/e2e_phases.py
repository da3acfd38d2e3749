"""Where the end-to-end round spends its time: ktn_separate (upload + kernels + wait) and ktn_fetch_cuts_view, per flag set.
   python scripts/e2e_phases.py          (B200; 10^6 log-sum-exp rows, v = 0.1)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from katana_jl_b200.binding import FLAG_DIRECT_VIEW, FLAG_EAGER_VIEW, FLAG_LEAN_VIEW, FLAG_TIME_KERNELS, load_cuda_library

lib = load_cuda_library()
kind, nv, rows, v = 1, 100_000, 1_000_000, 0.1
w = lib.synth_rows(kind, 1, nv, 0, rows); x0 = lib.synth_point(kind, 1, nv)
cases = [("copy", FLAG_LEAN_VIEW, 1), ("direct", FLAG_LEAN_VIEW | FLAG_DIRECT_VIEW, 1)]
cases += [(f"pipeline {S}", FLAG_LEAN_VIEW | FLAG_EAGER_VIEW, S) for S in (2, 3, 4, 6, 8)]
for name, flags, S in cases:
    h = lib.create(flags=flags, ngpus=S, devices=[0] * S) if S > 1 else lib.create(flags=flags)
    h.load(nv, w)
    g = h.eval_g(x0)
    ub = np.full(rows, np.quantile(g, 1 - v)); h.set_bounds(w.lb, ub)
    for _ in range(5):
        h.separate(x0, view=True)
    ts, tf = [], []
    for _ in range(30):
        t0 = time.perf_counter(); r = h.separate(x0, fetch=False); t1 = time.perf_counter(); b = h._fetch_view(r[0], r[3]); t2 = time.perf_counter()
        ts.append(t1 - t0); tf.append(t2 - t1)
    tm = h.timings() if S == 1 else dict(eval_ms=0, compact_ms=0, cut_ms=0, h2d_ms=0, d2h_ms=0)
    print(f"{name:24s} separate {1e3 * np.median(ts):.3f} ms  view {1e3 * np.median(tf):.3f} ms  total {1e3 * (np.median(ts) + np.median(tf)):.3f} ms | "
          f"K1 {1e3 * tm['eval_ms']:.1f} us  K2 {1e3 * tm['compact_ms']:.1f} us  K3 {1e3 * tm['cut_ms']:.1f} us  h2d {1e3 * tm['h2d_ms']:.1f} us  d2h {1e3 * tm['d2h_ms']:.1f} us  cuts {b.n_cuts}", flush=True)
    h.close()
