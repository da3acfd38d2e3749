"""A/B timing of library variants: K1 / K2 device times of one separation round per workload.  Run under gpurun:
python scripts/ab_time.py build/variants/libktn_a.so build/variants/libktn_b.so ..."""
import hashlib, os, sys
os.environ.setdefault("KTN_K1_EVENT_EVERY", "1")      # per-round K1 times (the library samples the K1 | K2 event by default)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from katana_jl_b200.binding import KtnLibrary
cases = [(1, 100000, 1000000, 0.1, "lse1e6"), (0, 100000, 1000000, 0.1, "qcqp1e6")]
if os.environ.get("AB_CASES") == "all":
    cases += [(1, 100000, 1000000, 1.0, "lse1e6"), (1, 100000, 1000000, 0.01, "lse1e6"), (0, 10000, 100000, 0.1, "qcqp1e5"), (2, 100000, 1000000, 0.1, "soc1e6")]
data = {}
if len(sys.argv) > 2:      # one process per variant: the libraries export the same symbols, and the first one loaded would serve the internal calls of all
    import subprocess
    for path in sys.argv[1:]:
        subprocess.run([sys.executable, os.path.abspath(__file__), path], check=False)
    sys.exit(0)
for path in sys.argv[1:]:
    P = KtnLibrary(path)
    out = [os.path.basename(path)]
    for kind, nv, nr, v, name in cases:
        if (kind, nv, nr) not in data:
            data[(kind, nv, nr)] = (P.synth_rows(kind, 20260001 + kind, nv, 0, nr), P.synth_point(kind, 20260001 + kind, nv))
        w, x0 = data[(kind, nv, nr)]
        h = P.create(flags=2 if os.environ.get('AB_DETAIL') else 0); h.load(nv, w)
        g = h.eval_g(x0)
        ev = []
        for _ in range(6):
            h.eval_g(x0); ev.append(h.timings()["eval_ms"])
        h.set_bounds(w.lb, np.full(nr, np.quantile(g, 1 - v)))
        b = h.separate(x0)        # results must not depend on the variant: compare the digests across the lines
        digest = hashlib.sha1(g.tobytes() + b.row_id.tobytes() + b.col.tobytes() + b.val.tobytes() + b.lo.tobytes() + b.hi.tobytes()).hexdigest()[:10]
        k1, k2, k3 = [], [], []
        dbg = getattr(P.dll, "ktn_debug_cycles", None) if hasattr(P.dll, "ktn_debug_cycles") else None
        import ctypes
        if dbg: dbg(None, 1)
        for it in range(12):
            st, nc, nz, er = h.separate(x0, fetch=False)
            tm = h.timings(); k1.append(tm["eval_ms"]); k2.append(tm["compact_ms"]); k3.append(tm["cut_ms"])
        out.append(f"{name} v={v}: K1 {1e3 * np.median(k1[2:]):.1f} us (min {1e3 * min(k1):.1f}) K2 {1e3 * np.median(k2[2:]):.1f} us K3 {1e3 * np.median(k3[2:]):.1f} us cuts {nc} eval-only {1e3 * np.median(ev[2:]):.1f} us sha {digest}")
        if dbg:
            arr = (ctypes.c_ulonglong * 16)(); dbg(arr, 1); a = [x / 12.0 for x in arr]
            nch = max(a[4], 1)
            out.append(f"[per chunk-warp cycles: fwd {a[0]/nch:.0f} sel {a[1]/nch:.0f} cut {a[2]/nch:.0f} ticket(per batch) {a[3]/nch:.0f}; chunks {nch:.0f} selected lanes/chunk {a[5]/nch:.2f} passes/chunk {a[6]/nch:.2f}]")
        h.close()
    print(" | ".join(out), flush=True)
