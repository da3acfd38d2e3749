#!/usr/bin/env python
"""Benchmark of the ECP separation round (BASELINE.json metric: nonlinear constraints linearised per second
per separation round) on synthetic instances of SURVEY.md section 8d.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (N > 1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle restatement)

A step is one separation round over the workload: evaluate g(x*) and the Jacobian rows of every nonlinear row,
test violation, emit the violated rows' cuts as CSR (and, for N > 1, combine all ranks' cuts over NCCL).
`value`: rounds with x* already resident in HBM, CUDA-event timed on the launching stream.
`e2e`:   the same round through the reference-facing separator call with HOST buffers (x* upload and cut
         download inside the timed region).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (synth kind, seed, description)
    "lse": (1, 20260002, "configs[2]: synthetic log-sum-exp rows log sum_k exp(a_k x_jk + b_k), K in 4..16"),
    "qcqp": (0, 20260001, "configs[1] family: sparse convex QCQP rows sum a x^2 + sum b x, 8 columns per row"),
    "soc": (2, 20260003, "configs[3] NL rows: sqrt(sum (s x)^2) - t"),
    "portfolio": (3, 20260003, "configs[3]: portfolio, one SOC-like NL row sqrt(sum_k (s_k x_jk)^2) - t per nine sparse linear rows (linear rows are never separated)"),
}
METRIC = "nonlinear cons linearised/sec per separation round"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Polls SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.stop = [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4, "hw_power_brake": 0x80}
        while not self.stop:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.004)

    def __enter__(self):
        if self.nv:
            try:                            # first NVML queries are slow (lazy paths): take them before the timed region starts
                nv = self.nv
                nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM)
                (nv.nvmlDeviceGetCurrentClocksEventReasons if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons)(self.dev)
            except Exception:
                pass
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.nv:
            self.t.join(timeout=1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_instance(lib, kind, seed, nv, row_begin, nrows):
    w = lib.synth_rows(kind, seed, nv, row_begin, nrows)
    x0 = lib.synth_point(kind, seed, nv)
    return w, x0


def cpu_rounds(oracle, w, x0, nv, ub, threads, seconds_budget, min_rounds=2):
    """Times the oracle's separation round (the reference algorithm, all rows, full Jacobian) on host cores."""
    os.environ["KTN_ORACLE_THREADS"] = str(threads)
    h = oracle.create()
    h.load(nv, w)
    h.set_bounds(w.lb, ub)
    h.separate(x0, fetch=False)
    ts = []
    t_end = time.perf_counter() + seconds_budget
    while len(ts) < min_rounds or (time.perf_counter() < t_end and len(ts) < 50):
        t0 = time.perf_counter()
        st, nc, nz, er = h.separate(x0, fetch=False)
        ts.append(time.perf_counter() - t0)
    h.close()
    return float(np.median(ts)), len(ts), nc


def workload_string(args, desc):
    """config.workload: the SAME string in both arms (the driver compares them)."""
    return f"{args.workload}: {desc}; m={args.rows} rows per GPU, n={args.vars} vars, violated fraction {args.v}, f_tol 1e-6"


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for the path.  The reference is Julia over un-vendored
    packages and cannot be built or run here, so this arm times the oracle (its C restatement) with all host threads.
    It loads the oracle and the synthetic generators only (libktn_synth.so), never the product library."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import katana_jl_b200  # noqa: F401
    from katana_jl_b200.binding import KtnLibrary, SynthLibrary
    synth = SynthLibrary()
    oracle = KtnLibrary(os.path.join(ROOT, "oracle", "libktn_oracle.so"))
    kind, seed, desc = WORKLOADS[args.workload]
    nv = args.vars
    rows = args.rows                      # a step is a bounded sample of the job: one GPU's share of the weak-scaling workload
    w, x0 = make_instance(synth, kind, seed, nv, 0, rows)
    h = oracle.create(); h.load(nv, w); g = h.eval_g(x0); h.close()
    nl_mask = (w.flags & 1) != 0
    nl_rows = int(nl_mask.sum())
    ub = np.full(rows, np.quantile(g[nl_mask], 1 - args.v))
    threads = os.cpu_count() or 1
    os.environ["KTN_ORACLE_THREADS"] = str(threads)
    h = oracle.create(); h.load(nv, w); h.set_bounds(w.lb, ub)
    for _ in range(args.warmup):
        h.separate(x0, fetch=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st, nc, nz, er = h.separate(x0, fetch=False)
    dt = time.perf_counter() - t0
    value = nl_rows * args.steps / dt
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "constraints/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(args, desc), "rows_per_step": rows, "nl_rows_per_step": nl_rows,
                   "note": "the CPU rate does not depend on the GPU count: every step separates one GPU's share of the workload"},
        "cpu_baseline": {"value": value, "unit": "constraints/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} full rounds over {rows} rows (C restatement of the Katana.jl separator, OpenMP over rows; not Julia)"},
        "e2e": {"value": value, "unit": "constraints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def sub_batch(b, keep):
    """The cuts of `b` whose row ids are in the sorted array `keep`, as (row_id, col, val, lo, hi) with per-cut entry slices concatenated."""
    sel = np.nonzero(np.isin(b.row_id, keep))[0]
    rp = np.asarray(b.row_ptr)
    idx = np.concatenate([np.arange(rp[c], rp[c + 1]) for c in sel]) if len(sel) else np.empty(0, np.int64)
    return (np.asarray(b.row_id)[sel].copy(), (rp[sel + 1] - rp[sel]).copy(), np.asarray(b.col)[idx].copy(), np.asarray(b.val)[idx].copy(),
            np.asarray(b.lo)[sel].copy(), np.asarray(b.hi)[sel].copy())


def same_cuts(a, b):
    return all(x.shape == y.shape and x.tobytes() == y.tobytes() for x, y in zip(a, b))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lse", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=1000000, help="nonlinear rows per GPU (weak scaling)")
    ap.add_argument("--vars", type=int, default=100000)
    ap.add_argument("--violated", dest="v", type=float, default=0.1, help="violated fraction of the rows at x* (not --v: torchrun's own parser claims that prefix)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--pipeline", type=int, default=0, help="e2e leg: shards per device whose cut transfers overlap the following shards' kernels (KTN_FLAG_EAGER_VIEW); 0 = KatanaGPUSeparator's own measured default (one device: two shards between 750 000 and 4 000 000 rows -- 10^6 LSE rows: 1 shard 0.604 ms, 2: 0.575, 3: 0.587, 4: 0.632, 8: 0.76; several devices: 1)")
    ap.add_argument("--direct", type=int, default=0, help="e2e leg at one GPU, one shard: 1 = the round's kernels store the cuts straight into the pinned host buffer (KTN_FLAG_DIRECT_VIEW; measured 0.622 ms against 0.604); 0 = the separator's default: device blob + one download after the round")
    ap.add_argument("--skip-e2e", action="store_true", help="exchange sweeps only: skip the end-to-end leg and the sharded parity check (the line then carries no e2e)")
    ap.add_argument("--topk", type=int, default=0, help="build extension: keep only the k most violated rows per round (0 = reference behaviour: all)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import katana_jl_b200  # noqa: F401
    from katana_jl_b200.binding import FLAG_LEAN_VIEW, KtnLibrary, comm_unique_id, load_cuda_library

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")     # host-side barriers: an NCCL barrier would keep the waiting ranks' GPUs busy
    lib = load_cuda_library()              # fails loudly if libktn.so is missing: there is no CPU fallback
    kind, seed, desc = WORKLOADS[args.workload]
    nv, rows = args.vars, args.rows
    row_begin = rank * rows               # weak scaling: every GPU owns `rows` rows of an instance with world*rows rows
    w, x0 = make_instance(lib, kind, seed, nv, row_begin, rows)
    if args.steps < 64:      # the library samples the K1 | K2 timing event (one round in eight): a short run needs it in every round
        os.environ.setdefault("KTN_K1_EVENT_EVERY", "1")
    h = lib.create(device=local, flags=FLAG_LEAN_VIEW)        # as KatanaGPUSeparator creates it: cut views carry what the LP needs
    h.load(nv, w)
    h.set_row_offset(row_begin)
    g = h.eval_g(x0)
    nl_mask = (w.flags & 1) != 0          # the rows the loop of src/model.jl:272 tests: nlconstr_ixs
    nl_rows = int(nl_mask.sum())
    ub_scalar = float(np.quantile(g[nl_mask], 1 - args.v))
    ub = np.full(rows, ub_scalar)
    h.set_bounds(w.lb, ub)
    if args.topk > 0:
        h.set_params(1e-6, 1e9, args.topk)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(comm_unique_id(lib)), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        h.comm_init(world, rank, bytes(idt.cpu().numpy().tobytes()))
    stream = torch.cuda.Stream()           # a real (non-default) stream: handle 0 would mean "the library's own stream"
    torch.cuda.set_stream(stream)
    h.set_stream(stream.cuda_stream)
    dx = torch.from_numpy(x0).cuda()

    def step():
        h.separate_device_async(dx.data_ptr())
        if world > 1:
            h.allgather_cuts_async()       # the push kernel (or NCCL) ships the round's cut blob beside the next round

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident rounds -------------------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    st, n_cuts, nnz, err = h.sync_counts()
    if world > 1:
        h.sync_gathered()                  # warm-up covers the exchange's completion path too
    import gc
    gc.collect(); gc.disable()             # no collector pauses inside the timed loops
    t_before = h.timings()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record(stream)
        for _ in range(args.steps):
            step()
        if world > 1:
            tot_cuts, tot_nnz = h.sync_gathered()      # the exchange is pipelined one round deep: the last payload belongs to the timed region
        e1.record(stream)
        barrier()
    ms_total = e0.elapsed_time(e1)
    t_after = h.timings()
    exchange_desc = {"peer-push": "peer push: ktn_push_kernel stores every rank's packed cut blob into all ranks' receive arenas (CUDA IPC) over NVLink, "
                                  "on its own stream beside the next round; NCCL bootstraps only",
                     "nccl": "NCCL on its own stream, pipelined: sizes allgather, then ONE ncclAllGather of the packed cut blobs (slots of the largest blob)",
                     "none": "none"}[h.exchange_transport() if world > 1 else "none"]
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * nl_rows / (ms_per_step * 1e-3)      # nonlinear constraints linearised per second: all ranks' NL rows per round
    launches = t_after["launches"] - t_before["launches"]
    timed = max(1, t_after["rounds_timed"] - t_before["rounds_timed"])
    k1_ms = (t_after["eval_ms_sum"] - t_before["eval_ms_sum"]) / timed
    k2_ms = (t_after["compact_ms_sum"] - t_before["compact_ms_sum"]) / timed       # K2 + K3 (no event is recorded between them)
    gc.enable()

    # ---- sharded runs: the gathered batch of one more round against a single handle, outside the timed region ------
    sharded = None
    if world > 1 and args.skip_e2e:
        xt = max(1, t_after["exchanges_timed"] - t_before["exchanges_timed"])
        exchange_ms = (t_after["exchange_ms_sum"] - t_before["exchange_ms_sum"]) / xt
        blob_bytes = 64 + 28 * n_cuts + 12 * nnz + 8
        sharded = {"exchange_ms": exchange_ms, "inbound_bytes_per_gpu": (world - 1) * blob_bytes,
                   "inbound_gbs_per_gpu": (world - 1) * blob_bytes / (exchange_ms * 1e-3) / 1e9 if exchange_ms > 0 else None}
    elif world > 1:
        xt = max(1, t_after["exchanges_timed"] - t_before["exchanges_timed"])
        exchange_ms = (t_after["exchange_ms_sum"] - t_before["exchange_ms_sum"]) / xt
        ubs = [None] * world
        dist.all_gather_object(ubs, ub_scalar, group=cpu_group)
        step()
        gathered = h.fetch_gathered()      # every rank downloads the combined batch: ascending global rows
        blob_bytes = 64 + 28 * n_cuts + 12 * nnz + 8
        sharded = {"exchange_ms": exchange_ms, "inbound_bytes_per_gpu": (world - 1) * blob_bytes,
                   "inbound_gbs_per_gpu": (world - 1) * blob_bytes / (exchange_ms * 1e-3) / 1e9 if exchange_ms > 0 else None,
                   "gathered_cuts": int(gathered.n_cuts)}
        piece = max(1, min(rows, 100000 // world))
        keep = np.concatenate([np.arange(r * rows, r * rows + piece) for r in range(world)])
        if rank == 0:
            chk = lib.create(device=local, flags=0)
            chk.load_begin(nv, world * piece)
            for r in range(world):
                chk.add_rows(r * piece, lib.synth_rows(kind, seed, nv, r * rows, piece))
            chk.load_end()
            chk.set_bounds(np.full(world * piece, -np.inf), np.repeat(np.array(ubs), piece))
            ref = chk.separate(x0)
            ref_rows = keep[ref.row_id]                 # the check handle numbers its rows 0 .. world * piece - 1
            want = (ref_rows, np.diff(ref.row_ptr), ref.col, ref.val, ref.lo, ref.hi)
            sharded["sharded_parity_exchange"] = bool(same_cuts(sub_batch(gathered, keep), want))
            chk.close()

    if args.skip_e2e:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "n_gpus": world, "ms_per_step": ms_per_step, "k1_ms": k1_ms, "k2_k3_ms": k2_ms, "sharded": sharded,
                              "push_blocks": os.environ.get("KTN_PUSH_BLOCKS"), "exchange": os.environ.get("KTN_EXCHANGE"), "e2e": None}), flush=True)
        if world > 1:
            dist.barrier(group=cpu_group); dist.destroy_process_group()
        return
    # ---- e2e: the separator call a Katana user makes, host buffers in and out ------------------------------------
    h.set_stream(0)
    torch.cuda.synchronize()
    from katana_jl_b200.separators import KatanaGPUSeparator
    x_host = x0.copy()
    e2e_steps = max(3, min(args.steps, 20))
    if world == 1:
        # the separator as a Katana user creates it for a large model: the rows in `args.pipeline` consecutive shards on the one device, every
        # shard's cuts downloaded as soon as it has finished (KTN_FLAG_EAGER_VIEW)
        sep = KatanaGPUSeparator(devices=[local], pipeline=args.pipeline or None, direct=bool(args.direct))
        per = sep.shards_per_device(nl_rows)       # the separator decides on the rows a round tests (nlconstr_ixs)
        if per > 1 or args.direct:
            hp = lib.create(**sep.handle_options(nl_rows)); hp.load(nv, w); hp.set_bounds(w.lb, ub)
        else:
            hp = h
        sep.handle = hp; sep.num_var, sep.num_constr = nv, rows
        e2e_call = "KatanaGPUSeparator.separate(xstar) -> CutBatch (ktn_separate + ktn_fetch_cuts_view: x* from host memory, cuts into the library's pinned buffer)"
        if args.direct and per <= 1:
            e2e_call = ("KatanaGPUSeparator.separate(xstar) -> CutBatch (ktn_separate + ktn_fetch_cuts_view: x* from host memory; KTN_FLAG_DIRECT_VIEW: the cut kernels "
                        "store the batch over PCIe straight into the library's pinned host buffer while they build it, no copy afterwards)")
        if per > 1:
            e2e_call += (f"; pipeline={per}{'' if args.pipeline else ' (the separator default at this size)'}: the rows run as consecutive shards on one stream, a small kernel per shard "
                         "stores its cuts into the pinned batch while the next shard computes, one host synchronisation per round")
    elif rank == 0:
        # ONE separator in ONE process over all world*rows rows, as the reference owns it (src/Katana.jl:18): ktn_options.ngpus
        sepg = KatanaGPUSeparator(ngpus=world, pipeline=args.pipeline or None)
        hg = lib.create(**sepg.handle_options(world * nl_rows))
        hg.load_begin(nv, world * rows)
        for r in range(world):
            hg.add_rows(r * rows, w if r == 0 else lib.synth_rows(kind, seed, nv, r * rows, rows))
        hg.load_end()
        hg.set_bounds(np.full(world * rows, -np.inf), np.repeat(np.array(ubs), rows))
        sep = sepg; sep.handle = hg; sep.num_var, sep.num_constr = nv, world * rows
        e2e_call = (f"KatanaGPUSeparator(ngpus={world}).separate(xstar) -> ONE CutBatch of all {world} devices' cuts in host memory (single process: x* uploaded to every "
                    "device, rounds side by side, every device downloads its cuts over its own PCIe link into one pinned buffer)")
    e2e_dt, batch = None, None
    if world == 1 or rank == 0:
        for _ in range(3):
            batch = sep.separate(x_host)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            batch = sep.separate(x_host)
        e2e_dt = time.perf_counter() - t0
        if world > 1:
            sharded["sharded_parity"] = bool(same_cuts(sub_batch(batch, keep), want)) and bool(sharded["sharded_parity_exchange"])
            sharded["e2e_cuts"] = int(batch.n_cuts)
    if world > 1:
        dist.barrier(group=cpu_group)      # the other ranks wait here, GPUs idle, while rank 0 times the single-process separator
    if rank != 0:
        dist.destroy_process_group()
        return
    e2e_value = world * nl_rows * e2e_steps / e2e_dt
    h2d_bytes = 8 * nv * world
    d2h_bytes = 64 + sum(int(getattr(batch, f).nbytes) for f in ("row_id", "row_ptr", "col", "val", "lo", "hi", "g", "viol", "bconst"))

    # ---- roofline of the dominant kernel (K1) and of the whole round ---------------------------------------------
    peak, peak_src = load_peaks()
    alg_round = h.algorithmic_bytes()                      # sum_NL(4 nnz + 8 C + 16) + 8 n + sum_sel(12 nnz + 28)
    alg_k1 = alg_round - 12 * nnz - 28 * n_cuts            # K1 reads every row's columns, constants and bounds and x*; the cuts' CSR is K2 / K3's share
    achieved = alg_k1 / (k1_ms * 1e-3) / 1e9
    round_ms = ms_per_step if world == 1 else k1_ms + k2_ms
    family = True      # every benchmark form is a family shape (ktn_family.h)
    out = {
        "metric": METRIC, "value": value, "unit": "constraints/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(args, desc),
                   "rows_per_gpu": rows, "nl_rows_per_gpu": nl_rows, "num_var": nv, "violated_fraction": args.v, "topk": args.topk, "cuts_per_round_per_gpu": n_cuts, "cut_nnz_per_round_per_gpu": nnz,
                   "l2": f"no flush needed: one round streams {alg_round / 1e6:.0f} MB of inputs > 126 MB L2",
                   "exchange": "none" if world == 1 else exchange_desc},
        "roofline": {"bound": "hbm", "kernel": "ktn_family_kernel (K1: evaluate g of every row, violation test; one launch per round)" if family else "ktn_round_kernel (K1, tape interpreter: evaluate, test, cut rows)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_k1, "ms_per_launch": k1_ms,
                     # whole round: one GPU: the timed step IS K1 + K2 + K3 back to back (CUDA events around the timed rounds); several GPUs: the
                     # step also holds the exchange, so the round is the sum of the kernel times of the sampled rounds (ktn_timings)
                     "round": {"algorithmic_bytes": alg_round, "ms": round_ms, "frac": alg_round / (round_ms * 1e-3) / 1e9 / peak,
                               "k2_k3_ms": k2_ms, "sampled_k1_k2_k3_ms": k1_ms + k2_ms,
                               "note": "K1 / K2+K3 times come from one round in eight (the event between K1 and K2 costs 2-4 us and is sampled)", "kernels": "K1 ktn_family_kernel, K2 ktn_compact_kernel (ordered compaction), K3 ktn_cut_kernel (cuts of the selected rows)"}},
        "e2e": {"value": e2e_value, "unit": "constraints/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": 1e3 * e2e_dt / e2e_steps, "steps": e2e_steps, "call": e2e_call},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
    }
    if sharded is not None:
        out["sharded"] = sharded
        out["sharded_parity"] = sharded.get("sharded_parity")
    traffic_file = os.path.join(ROOT, "profiles", "k1_dram_traffic.json")
    if os.path.exists(traffic_file):
        tr = json.load(open(traffic_file)).get(args.workload)
        if tr and tr.get("rows") == rows and abs(tr.get("v", -1) - args.v) < 1e-9:
            out["roofline"]["traffic"] = tr["dram_bytes"]
            out["roofline"]["traffic_source"] = tr.get("source")

    # ---- cpu_baseline: the reference algorithm on this box's host cores (N = 1 only) -----------------------------
    if world == 1 and not args.no_cpu:
        oracle = KtnLibrary(os.path.join(ROOT, "oracle", "libktn_oracle.so"))
        sample_rows = min(rows, 1000000)
        ws, _ = (w, None) if sample_rows == rows else make_instance(lib, kind, seed, nv, 0, sample_rows)
        med1, n1, nc1 = cpu_rounds(oracle, ws, x0, nv, ub[:sample_rows], 1, args.cpu_seconds)
        cores = os.cpu_count() or 1
        medn, nn, _ = cpu_rounds(oracle, ws, x0, nv, ub[:sample_rows], cores, args.cpu_seconds / 3)
        nl_sample = int(((ws.flags & 1) != 0).sum())
        out["cpu_baseline"] = {"value": nl_sample / med1, "unit": "constraints/s", "cores": 1, "kind": "port",
                               "sample": f"median of {n1} rounds over {sample_rows} rows of the same workload, 1 thread (the reference is single-threaded); "
                                         "C restatement of the Katana.jl separator, not Julia",
                               "all_cores": {"value": nl_sample / medn, "cores": cores, "rounds": nn}}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
