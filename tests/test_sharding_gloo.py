"""Sharded separation: contiguous row ranges per rank, rank-major concatenation
equals the single-handle result.  Host logic checked with world_size 2 over gloo on the CPU, with the oracle
standing in for each rank's device."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import katana_jl_b200  # noqa: F401
    from katana_jl_b200.binding import CUDA_LIB_PATH, KtnLibrary
    from katana_jl_b200.sharding import combine_rank_major, merge_topk, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    synth = KtnLibrary(CUDA_LIB_PATH)                       # generators only, no device call
    orc = KtnLibrary(os.path.join(ROOT, "oracle", "libktn_oracle.so"))
    kind, seed, nv, m = 1, 20260002, 2000, 5000
    x0 = synth.synth_point(kind, seed, nv)
    r0, r1 = shard_range(m, world, rank)
    w = synth.synth_rows(kind, seed, nv, r0, r1 - r0)        # a row depends only on (seed, global row index)
    h = orc.create(); h.load(nv, w)
    g_local = h.eval_g(x0)
    g_all = [None] * world
    dist.all_gather_object(g_all, g_local)
    ub = np.quantile(np.concatenate(g_all), 0.9)
    h.set_bounds(w.lb, np.full(r1 - r0, ub))
    local = h.separate(x0)
    parts = [None] * world
    dist.all_gather_object(parts, (r0, local))
    combined = combine_rank_major(parts)
    if rank == 0:
        wf = synth.synth_rows(kind, seed, nv, 0, m)
        hf = orc.create(); hf.load(nv, wf); hf.set_bounds(wf.lb, np.full(m, ub))
        full = hf.separate(x0)
        ok = all(np.array_equal(getattr(full, f), getattr(combined, f)) for f in ("row_id", "row_ptr", "col", "val", "lo", "hi", "g", "viol", "bconst"))
        open(os.path.join(out_dir, "result"), "w").write(f"{int(ok)} {full.n_cuts} {combined.n_cuts}")
    # top-k extension: every rank keeps its k most violated rows, the merge of the union is the whole instance's top-k
    for k in (1, 37, 200, 10**6):
        h.set_params(1e-6, 1e9, k)
        parts_k = [None] * world
        dist.all_gather_object(parts_k, (r0, h.separate(x0)))
        merged = merge_topk(combine_rank_major(parts_k), k)
        if rank == 0:
            hf.set_params(1e-6, 1e9, k)
            full_k = hf.separate(x0)
            ok_k = all(np.array_equal(getattr(full_k, f), getattr(merged, f)) for f in ("row_id", "row_ptr", "col", "val", "lo", "hi", "g", "viol", "bconst"))
            open(os.path.join(out_dir, "result"), "a").write(f" {int(ok_k)}:{full_k.n_cuts}")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_whole(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok, n_full, n_comb, *topk = open(tmp_path / "result").read().split()
    assert ok == "1" and n_full == n_comb and int(n_full) > 100
    assert len(topk) == 4 and all(t.startswith("1:") for t in topk), topk          # sharded top-k == whole-instance top-k
    assert [int(t.split(":")[1]) for t in topk[:3]] == [1, 37, 200]


def test_shard_ranges_cover_and_balance():
    from katana_jl_b200.sharding import shard_range
    for m in (0, 1, 7, 1000, 10**6 + 3):
        for world in (1, 2, 4, 8):
            rs = [shard_range(m, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == m
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


def test_weighted_shards_balance_tape_bytes():
    """split on tape bytes.  Contiguous, covering, and no rank heavier than the ideal share by more than one row."""
    from katana_jl_b200.sharding import shard_ranges_by_weight
    rng = np.random.default_rng(7)
    for m in (0, 1, 5, 1000, 20011):
        for world in (1, 2, 3, 8):
            w = rng.integers(18, 67, size=m).astype(float)            # LSE tape lengths 4K + 2, K in 4..16
            if m > 10:
                w[-1] = 5000.0                                         # a dense epigraph row at the end
            rs = shard_ranges_by_weight(w, world)
            assert len(rs) == world and rs[0][0] == 0 and rs[-1][1] == m
            assert all(rs[i][1] == rs[i + 1][0] and rs[i][0] <= rs[i][1] for i in range(world - 1))
            if m >= 1000:
                share = w.sum() / world
                assert max(w[a:b].sum() for a, b in rs) <= share + w.max()
    # equal weights degenerate to (almost) equal counts
    rs = shard_ranges_by_weight(np.ones(1000), 8)
    assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
