import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
from katana_jl_b200.binding import FLAG_EAGER_VIEW, FLAG_LEAN_VIEW, load_cuda_library
lib = load_cuda_library()
kind, nv, rows, v = 1, 100_000, 1_000_000, 0.1
w = lib.synth_rows(kind, 1, nv, 0, rows); x0 = lib.synth_point(kind, 1, nv)
S = int(sys.argv[1])
h = lib.create(flags=FLAG_LEAN_VIEW | FLAG_EAGER_VIEW, ngpus=S, devices=[0] * S); h.load(nv, w)
g = h.eval_g(x0); h.set_bounds(w.lb, np.full(rows, np.quantile(g, 1 - v)))
os.environ["X"] = "1"
for i in range(4):
    t0 = time.perf_counter(); h.separate(x0, view=True); print(f"round {i}: {1e3 * (time.perf_counter() - t0):.3f} ms", file=sys.stderr)
