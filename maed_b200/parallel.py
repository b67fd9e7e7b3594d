"""Multi-GPU plumbing for the clip-sharded hot path (one process per GPU, torch.distributed).

The path shards by clip: every rank runs the same model on its own clips and there is no
data-path collective in the forward.  What needs care is the host logic around it — which clips a rank owns,
how per-rank results are re-assembled in order, and how a throughput is reduced (max over ranks) — and that is
what lives here, backend-agnostic so it is covered by world_size-2 `gloo` tests on CPU (tests/test_parallel.py).
"""
import torch
import torch.distributed as dist


def clip_shard(n_clips: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) range of clips owned by `rank` (first n_clips % world ranks get one extra)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %d/%d" % (rank, world))
    base, extra = divmod(n_clips, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device="cpu") -> float:
    """Job time = slowest rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sharded_forward(forward_fn, clips: torch.Tensor, gather: bool = True):
    """Runs `forward_fn` on this rank's shard of `clips` (N, T, ...) and (optionally) all-gathers the per-clip
    outputs back into global clip order on every rank.  `forward_fn(x) -> dict of tensors with leading dim N_local`."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = clip_shard(clips.shape[0], rank, world)
    local = forward_fn(clips[lo:hi]) if hi > lo else {}
    if world == 1 or not gather:
        return local
    # ragged gather: shards differ by at most one clip; pad to the largest shard
    sizes = [clip_shard(clips.shape[0], r, world) for r in range(world)]
    nmax = max(h - l for l, h in sizes)
    keys = sorted(local.keys()) if local else None
    obj = [None] * world
    dist.all_gather_object(obj, keys)
    keys = next(k for k in obj if k is not None)
    out = {}
    for k in keys:
        if local:
            v = local[k]
            shape, dtype, device = (nmax,) + tuple(v.shape[1:]), v.dtype, v.device
            meta = [shape, str(dtype)]
        else:
            meta = None
        metas = [None] * world
        dist.all_gather_object(metas, meta)
        shape, dtype_s = next(m for m in metas if m is not None)
        dtype = getattr(torch, dtype_s.split(".")[-1])
        device = local[k].device if local else clips.device
        buf = torch.zeros(shape, dtype=dtype, device=device)
        if local:
            buf[: hi - lo] = local[k]
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
        out[k] = torch.cat([parts[r][: sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)
    return out
