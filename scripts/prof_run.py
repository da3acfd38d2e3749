"""Small driver for ncu captures: a few separation rounds of one synthetic config."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from katana_jl_b200.binding import load_cuda_library
kind = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nr = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1000000
v = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 3
nv = max(17, nr // 10)
P = load_cuda_library()
w = P.synth_rows(kind, 20260001 + kind, nv, 0, nr); x0 = P.synth_point(kind, 20260001 + kind, nv)
h = P.create(); h.load(nv, w)
g = h.eval_g(x0)
h.set_bounds(w.lb, np.full(nr, np.quantile(g, 1 - v)))
for it in range(rounds):
    st, nc, nz, er = h.separate(x0, fetch=False)
print('cuts', nc, 'nnz', nz, 'kernel_ms', h.timings()['kernel_ms'])
