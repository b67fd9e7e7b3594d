"""CPU, world_size 2, gloo: host logic of the clip-sharded multi-GPU path (maed_b200/parallel.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maed_b200 import parallel


def test_clip_shard_partitions_exactly():
    for n in range(0, 40):
        for w in (1, 2, 3, 4, 8):
            spans = [parallel.clip_shard(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.clip_shard(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clips, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        clips = torch.randn(n_clips, 2, 5, generator=g)              # same global batch on every rank

        def fake_forward(x):                                          # clip-independent stand-in for MAED.forward
            return {"theta": x.sum(dim=(1, 2), keepdim=False).unsqueeze(-1) * torch.arange(1, 4.0),
                    "rot": x.flip(-1)}

        out = parallel.sharded_forward(fake_forward, clips)
        ref = fake_forward(clips)
        ok = all(torch.equal(out[k], ref[k]) for k in ref)
        slow = parallel.max_over_ranks(10.0 + rank)                  # slowest rank defines the job time
        q.put((rank, ok, slow))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [8, 5, 1])
def test_sharded_forward_two_ranks_gloo(n_clips):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [11.0, 11.0]


# ------------------------------------------------------------------ training path: gradient averaging (gloo, CPU)
def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from maed_b200 import train

        class Holder(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.a = torch.nn.Parameter(torch.zeros(3, 5))
                self.b = torch.nn.Parameter(torch.zeros(7))

        m = Holder()
        # (1) flat-buffer path: what the engine's backward leaves behind (p.grad are views of TrainState.flat_grad)
        st = train.TrainState(m)
        views = st.ensure_grads([m.a, m.b])
        for p, v in zip((m.a, m.b), views):
            v.fill_(float(rank + 1))
            p.grad = v
        m._train_state = st
        train.allreduce_gradients(m)
        flat_ok = bool((m.a.grad == 1.5).all() and (m.b.grad == 1.5).all() and m.a.grad.data_ptr() == st.flat_grad.data_ptr())
        # (2) per-parameter fallback (no engine state)
        m2 = Holder()
        m2.a.grad = torch.full((3, 5), float(rank))
        m2.b.grad = torch.full((7,), float(10 * rank))
        train.allreduce_gradients(m2)
        per_ok = bool((m2.a.grad == 0.5).all() and (m2.b.grad == 5.0).all())
        q.put((rank, flat_ok, per_ok))
    finally:
        dist.destroy_process_group()


def test_allreduce_gradients_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] for r in res), res
