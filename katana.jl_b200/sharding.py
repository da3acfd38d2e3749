"""Row sharding across GPUs.

Constraints are independent given x*, so the NL rows are split into contiguous ranges, one per
rank (one process per GPU).  Every rank separates its slice; concatenating the per-rank cut
batches in rank order IS ascending row order, the reference's emission order (src/model.jl:272),
so the combined cut set is identical for any GPU count.
"""
import numpy as np

from .binding import CutBatch


def shard_range(num_rows, world_size, rank):
    """Contiguous, balanced [begin, end) of rank `rank`."""
    base, rem = divmod(num_rows, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_ranges_by_weight(weights, world_size):
    """Contiguous [begin, end) per rank with (nearly) equal total weight: SURVEY.md section 8e splits on tape bytes, not on
    row counts (a rank of long tapes would otherwise finish last).  `weights[i]` >= 0 is row i's cost -- use the tape length
    `np.diff(wire.expr_ptr)` (dense epigraph rows: their n + 1 Jacobian entries).  Order is preserved, so the last row (the
    epigraph row, src/model.jl:148) stays on the last non-empty rank.  Cuts are made where the prefix sum crosses k/world of the
    total; ranks may be empty when there are fewer rows than ranks."""
    w = np.asarray(weights, dtype=np.float64)
    m = len(w)
    if m == 0:
        return [(0, 0)] * world_size
    prefix = np.concatenate([[0.0], np.cumsum(np.maximum(w, 0.0))])
    total = prefix[-1]
    if total <= 0.0:
        return [shard_range(m, world_size, r) for r in range(world_size)]
    cuts = [0]
    for k in range(1, world_size):
        target = total * k / world_size
        i = int(np.searchsorted(prefix, target, side="left"))          # first prefix >= target
        if i > 0 and target - prefix[i - 1] < prefix[i] - target:        # the nearer boundary
            i -= 1
        cuts.append(min(max(i, cuts[-1]), m))
    cuts.append(m)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def combine_rank_major(parts):
    """parts: list over ranks of (row_begin, CutBatch with shard-local row ids).  Returns one CutBatch
    with global row ids; stops at the first rank that reported a non-finite row (src/model.jl:278)."""
    row_id, col, val, lo, hi, g, viol, bc, ptr = [], [], [], [], [], [], [], [], [np.zeros(1, np.int64)]
    status, err_row, off = 0, -1, 0
    for row_begin, b in parts:
        row_id.append(b.row_id + row_begin); col.append(b.col); val.append(b.val)
        lo.append(b.lo); hi.append(b.hi); g.append(b.g); viol.append(b.viol); bc.append(b.bconst)
        ptr.append(b.row_ptr[1:] + off); off += int(b.row_ptr[-1])
        if b.status != 0:
            status, err_row = b.status, b.err_row + row_begin
            break
    cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt)
    return CutBatch(status, err_row, cat(row_id, np.int64), np.concatenate(ptr), cat(col, np.int32), cat(val, np.float64),
                    cat(lo, np.float64), cat(hi, np.float64), cat(g, np.float64), cat(viol, np.float64), cat(bc, np.float64))


def merge_topk(batch, k):
    """Global top-k of a combined batch ("top-k extension").  Every rank keeps its own k most violated
    rows (ktn_options.topk on a sharded handle is a LOCAL selection); the union of those, combined rank-major, contains the
    global top-k, and this is the identical deterministic merge every rank (or the host) runs on it: rank the cuts by
    (NaN first, violation descending, row ascending) -- the oracle's order -- keep k, emit them in ascending row order.
    Rounds that stopped at a non-finite cut are returned as they are (the reference abandons such a solve, src/model.jl:278)."""
    n = len(batch.row_id)
    if k <= 0 or n <= k or batch.status != 0:
        return batch
    v = batch.viol
    nan = np.isnan(v)
    order = np.lexsort((batch.row_id, -np.where(nan, 0.0, v), ~nan))      # last key first: NaN rows, then violation, then row
    keep = np.sort(order[:k])                                                # positions in the batch = ascending rows
    lens = np.diff(batch.row_ptr)[keep]
    ptr = np.concatenate([np.zeros(1, np.int64), np.cumsum(lens)]).astype(np.int64)
    src = np.repeat(batch.row_ptr[keep] - ptr[:-1], lens) + np.arange(int(ptr[-1]), dtype=np.int64)
    return CutBatch(batch.status, batch.err_row, batch.row_id[keep], ptr, batch.col[src], batch.val[src], batch.lo[keep], batch.hi[keep],
                    batch.g[keep], batch.viol[keep], batch.bconst[keep])
