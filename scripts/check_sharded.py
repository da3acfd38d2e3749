"""Sharded-path check on real GPUs (run under torchrun, one rank per GPU): every rank separates its contiguous slice of the
rows, the cuts are combined over NCCL (pipelined exchange), and the gathered CSR must be bit-identical to one handle separating
the whole instance.  python -m torch.distributed.run --nproc-per-node N scripts/check_sharded.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import katana_jl_b200  # noqa: F401
from katana_jl_b200.binding import comm_unique_id, load_cuda_library

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = load_cuda_library()
ok = True
for kind, nv, rows in ((1, 5000, 40000), (0, 2000, 30000)):
    total = world * rows
    seed = 20260001 + kind
    x0 = lib.synth_point(kind, seed, nv)
    w = lib.synth_rows(kind, seed, nv, rank * rows, rows)
    h = lib.create(device=local); h.load(nv, w); h.set_row_offset(rank * rows)
    wf = lib.synth_rows(kind, seed, nv, 0, total)
    hf = lib.create(device=local); hf.load(nv, wf)
    g = hf.eval_g(x0)
    ub = np.full(total, np.quantile(g, 0.85))
    hf.set_bounds(wf.lb, ub); h.set_bounds(w.lb, ub[rank * rows:(rank + 1) * rows])
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(comm_unique_id(lib)), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    h.comm_init(world, rank, bytes(idt.cpu().numpy().tobytes()))
    dx = [torch.from_numpy(x0 * s).cuda() for s in (1.0, 0.5, 0.75, 1.0, 0.9)]
    refs = [hf.separate(x0 * s) for s in (1.0, 0.5, 0.75, 1.0, 0.9)]
    # several rounds in flight (the exchange is pipelined), the gathered result is the LAST round's
    for upto in (1, 2, 3, 5):
        for i in range(upto):
            h.separate_device_async(dx[i].data_ptr()); h.allgather_cuts_async()
        got = h.fetch_gathered(); ref = refs[upto - 1]
        for f in ("row_id", "row_ptr", "col", "val", "lo", "hi", "g", "viol", "bconst"):
            a, b = getattr(got, f), getattr(ref, f)
            same = a.shape == b.shape and a.tobytes() == b.tobytes()
            if not same:
                ok = False
                print(f"rank {rank} kind {kind} rounds {upto}: MISMATCH in {f} {a.shape} {b.shape}", flush=True)
    # top-k extension: local top-k per rank, exchange, identical merge on every rank == the single handle's top-k
    from katana_jl_b200.sharding import merge_topk
    for k in (1, 500, 10**7):
        h.set_params(1e-6, 1e9, k); hf.set_params(1e-6, 1e9, k)
        h.separate_device_async(dx[0].data_ptr()); h.allgather_cuts_async()
        got = merge_topk(h.fetch_gathered(), k); ref = hf.separate(x0)
        for f in ("row_id", "row_ptr", "col", "val", "lo", "hi", "g", "viol", "bconst"):
            a, b = getattr(got, f), getattr(ref, f)
            if not (a.shape == b.shape and a.tobytes() == b.tobytes()):
                ok = False
                print(f"rank {rank} kind {kind} top-{k}: MISMATCH in {f} {a.shape} {b.shape}", flush=True)
    h.close(); hf.close()
t = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED PARITY", "OK" if int(t.item()) else "FAILED", f"({world} ranks)", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) else 1)
