"""Evaluation-only kernels (ktn_eval_g) timed on 1e6-row instances: the forward pass without the cut path.  Run under gpurun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import katana_jl_b200  # noqa: F401
from katana_jl_b200.binding import load_cuda_library
P = load_cuda_library()
for kind, nv, nr, name in ((1, 100000, 1000000, "lse1e6"), (0, 100000, 1000000, "qcqp1e6")):
    w = P.synth_rows(kind, 20260001 + kind, nv, 0, nr); x0 = P.synth_point(kind, 20260001 + kind, nv)
    h = P.create(); h.load(nv, w)
    ts = []
    for _ in range(8):
        h.eval_g(x0); ts.append(h.timings()["eval_ms"])
    print(f"{name}: evaluation-only kernels median {np.median(ts[2:]) * 1e3:.1f} us, min {min(ts) * 1e3:.1f} us", flush=True)
    h.close()
