"""BASELINE.json configs[4]: separation throughput sweep on one GPU (the 2/4/8-GPU points are bench.py --gpus N lines).
m in 1e4..1e7 rows (n = m / 10 variables), QCQP and log-sum-exp rows, violated fraction v in {0.01, 0.1, 1}.
Prints a markdown table: device time of K1 and of K2 + K3 (CUDA events inside the library), rows/s of the round, algorithmic GB/s
(SURVEY.md 8d formula) and its fraction of the measured HBM peak.  Run under gpurun: python scripts/sweep.py > profiles/sweep.md"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from katana_jl_b200.binding import load_cuda_library
peak = 6553.9
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
P = load_cuda_library()
sizes = [10**4, 10**5, 10**6, 10**7] if len(sys.argv) < 2 else [int(float(a)) for a in sys.argv[1:]]
print(f"| workload | rows m | vars n | v | cuts | K1 us | K2+K3 us | round rows/s | algorithmic MB | GB/s (round) | frac of {peak:.0f} GB/s (round) | frac (K1 alone) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for kind, name in ((0, "qcqp"), (1, "lse"), (2, "soc")):
    for m in sizes:
        nv = max(17, m // 10)
        w = P.synth_rows(kind, 20260001 + kind, nv, 0, m); x0 = P.synth_point(kind, 20260001 + kind, nv)
        h = P.create(); h.load(nv, w)
        g = h.eval_g(x0)
        for v in (0.01, 0.1, 1.0):
            h.set_bounds(w.lb, np.full(m, np.quantile(g, 1 - v)))
            k1, k2 = [], []
            for it in range(12):
                st, nc, nz, er = h.separate(x0, fetch=False)
                tm = h.timings(); k1.append(tm["eval_ms"]); k2.append(tm["compact_ms"])
            a, b = float(np.median(k1[2:])), float(np.median(k2[2:]))
            ab = h.algorithmic_bytes(); ab1 = ab - 12 * nz - 28 * nc
            print(f"| {name} | {m} | {nv} | {v} | {nc} | {1e3*a:.1f} | {1e3*b:.1f} | {m/((a+b)*1e-3):.3e} | {ab/1e6:.1f} | {ab/(a+b)/1e6:.0f} | {ab/(a+b)/1e6/peak:.3f} | {ab1/a/1e6/peak:.3f} |", flush=True)
        h.close(); del w
