"""GPU bring-up check: CUDA library vs CPU oracle, bit for bit, plus a first timing. Run under gpurun."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import katana_jl_b200 as K
from katana_jl_b200 import expr as E
from katana_jl_b200.binding import KtnLibrary, ROW_NL, ROW_DENSE, load_cuda_library
O = KtnLibrary('oracle/libktn_oracle.so'); P = load_cuda_library()

def same(a, b, name):
    a = np.asarray(a); b = np.asarray(b)
    ok = a.shape == b.shape
    if ok and a.dtype.kind == 'f':
        na, nb = np.isnan(a), np.isnan(b)
        ok = np.array_equal(na, nb) and np.array_equal(a[~na].view(np.int64), b[~nb].view(np.int64))
    elif ok:
        ok = np.array_equal(a, b)
    if not ok:
        print('MISMATCH', name, a.shape, b.shape, flush=True)
        if a.shape == b.shape:
            d = np.flatnonzero(~((a == b) | (np.isnan(a.astype(float)) & np.isnan(b.astype(float)))))
            print(' first diffs at', d[:5], a[d[:5]], b[d[:5]], 'ndiff', len(d))
    return ok

def compare(w, nvar, xs, label):
    ho, hp = O.create(), P.create()
    ho.load(nvar, w); hp.load(nvar, w)
    ok = True
    for x in xs:
        bo, bp = ho.separate(x), hp.separate(x)
        ok &= (bo.status == bp.status) & (bo.err_row == bp.err_row)
        if bo.status != bp.status: print('status', bo.status, bp.status, bo.err_row, bp.err_row)
        for f in ('row_id', 'row_ptr', 'col', 'val', 'lo', 'hi', 'g', 'viol'):
            ok &= same(getattr(bo, f), getattr(bp, f), label + '.' + f)
        ok &= same(ho.eval_g(x), hp.eval_g(x), label + '.eval_g')
        rows = np.arange(0, w.nrows, max(1, w.nrows // 7), dtype=np.int64)
        co, cp = ho.gencut_rows(x, rows, False), hp.gencut_rows(x, rows, False)
        for f in ('row_id', 'row_ptr', 'col', 'val', 'lo', 'hi', 'g'):
            ok &= same(getattr(co, f), getattr(cp, f), label + '.gencut.' + f)
    print(label, 'OK' if ok else 'FAIL', 'cuts', bo.n_cuts, 'status', bo.status, flush=True)
    return ok

x, y, z = E.var(0), E.var(1), E.var(2)
exprs = [x**2 + y**2, E.exp(x-2.0) - 0.5 - y, y - (E.log(x)+0.5), E.sqrt(x**2+y**2) - (z-0.25), y*E.exp(x/y) - z, y*E.exp((-x)/y) - z,
         2*x**2 - 4*x*y - 4*x + 4 - y, E.Node(4, (E.const(1.5), x, y, z)), E.Node(4, (E.const(2.0), x, x)) + E.Node(4, (E.const(-1.0), y)),
         x**3 + y**0.5 + 2.0**z, x**y, E.abs_(x-y) + z / (x*x + 1.0), (x-1.0)**2 + (y-2.0)**2 - z, x, E.const(3.0)*x - y/2.0 + 1.0, -(x+y+2*z),
         E.sum_([E.exp(E.var(i)) for i in range(3)]), 1.0/x + x/(y+z), (x*y)**2 / z]
m = len(exprs)
w = E.to_wire(exprs, [-np.inf]*m, [1.0, 0, 0, 0, 0, 0, 0, 1, 1, 5, 2, 1, 0, 0.5, 1, 0, 4, 3, 1], [ROW_NL]*(m-1) + [ROW_NL | ROW_DENSE])
xs = [np.array([2.0, 2.0, 1.0]), np.array([0.5, 1.5, 0.25]), np.array([0.0, 0.0, 0.0]), np.array([-1.0, 3.0, 2.0]), np.array([1.0, 0.0, 2.0]), np.array([3.0, -2.0, 0.0])]
allok = compare(w, 3, xs, 'kat')
for kind, nv, nr, name in ((0, 1000, 5000, 'qcqp'), (1, 5000, 20000, 'lse'), (2, 2000, 3000, 'soc')):
    w = P.synth_rows(kind, 20260001 + kind, nv, 0, nr)
    x0 = P.synth_point(kind, 20260001 + kind, nv)
    ho = O.create(); ho.load(nv, w); g = ho.eval_g(x0)
    for v in (0.1, 1.0, 0.0):
        w.ub[:] = np.quantile(g, 1 - v) if v > 0 else g.max() + 1
        allok &= compare(w, nv, [x0, x0 * 0.5], f'{name} v={v}')
print('PARITY', 'ALL OK' if allok else 'SOME FAILED', flush=True)

# ---- first timing: LSE 1e6 rows / QCQP 1e6 rows, device-resident rounds ----
import ctypes
for kind, nv, nr, name in ((1, 100000, 1000000, 'lse1e6'), (0, 100000, 1000000, 'qcqp1e6'), (0, 10000, 100000, 'qcqp1e5')):
    t0 = time.time(); w = P.synth_rows(kind, 20260001 + kind, nv, 0, nr); x0 = P.synth_point(kind, 20260001 + kind, nv); t1 = time.time()
    hp = P.create(); hp.load(nv, w); t2 = time.time()
    g = hp.eval_g(x0)
    for v in (0.1, 0.01, 1.0):
        ub = np.full(nr, np.quantile(g, 1 - v)); hp.set_bounds(w.lb, ub)
        st, nc, nz, er = hp.separate(x0, fetch=False)
        # timed: library events around the kernels of one round (x already on device from the last separate)
        ts = []
        for it in range(10):
            st, nc, nz, er = hp.separate(x0, fetch=False)
            tm = hp.timings(); ts.append(tm['kernel_ms']); k1s = k1s + [tm['eval_ms']] if it else [tm['eval_ms']]
        ab = hp.algorithmic_bytes()
        ms = float(np.median(ts[2:]))
        print(f'{name} v={v}: gen {t1-t0:.1f}s load {t2-t1:.1f}s cuts {nc} nnz {nz} kernel_ms median {ms:.4f} min {min(ts):.4f} (K1 {np.median(k1s[2:]):.4f}) -> {nr/ms/1e3:.1f} Mrows/s, alg bytes {ab/1e6:.1f} MB -> {ab/ms/1e6:.0f} GB/s ({ab/ms/1e6/6553.9:.3f} of measured copy peak)', flush=True)
    if name == 'lse1e6':
        # cross-check a full-size round against the oracle
        ho = O.create(); ho.load(nv, w); ho.set_bounds(w.lb, np.full(nr, np.quantile(g, 0.9))); hp.set_bounds(w.lb, np.full(nr, np.quantile(g, 0.9)))
        t = time.time(); bo = ho.separate(x0); to = time.time() - t
        bp = hp.separate(x0)
        ok = all(same(getattr(bo, f), getattr(bp, f), 'full.' + f) for f in ('row_id', 'row_ptr', 'col', 'val', 'lo', 'hi', 'g', 'viol'))
        print('full-size LSE parity', 'OK' if ok else 'FAIL', 'oracle 1-thread round %.2fs' % to, flush=True)
    hp.close()
