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
        st._layout([m.a, m.b], [True, True])
        views = st._views(st.flat_grad, [m.a, m.b])
        for p, v in zip((m.a, m.b), views):
            v.fill_(float(rank + 1))
            p.grad = v
        m._train_state = st
        m._train_param_order = [("a", m.a), ("b", m.b)]
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


# ------------------------------------------------------------- torch DistributedDataParallel around the product module
def _ddp_worker(rank, world, port, q):
    """reference train.py:113: DistributedDataParallel(model, broadcast_buffers=False) around the model in train() mode; the
    engine's single autograd node must feed DDP's per-parameter hooks (every parameter gets a gradient) and the averaged
    gradients must land in the engine's flat gradient buffer."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "emu"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    os.environ.setdefault("MAED_EMU_THREADS", "4")
    torch.set_num_threads(4)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import harness
        from maed_b200.models import MAED
        from oracle import synth
        with harness.product_on_cpu():
            m = MAED("ste", 1, 12, "vanilla", "ktd", 1024)
            synth.fill_module_(m, 21)
            m = m.train().enable_training(True, dropout_p=0.0)
            ddp = torch.nn.parallel.DistributedDataParallel(m, broadcast_buffers=False)
            x = synth.synth_frames(world, 1, 21)[rank:rank + 1]
            A, B, C_ = [synth.synth_tensor("grad_probe.%s" % k, (world, n), 21) for k, n in (("pose", 144), ("shape", 10), ("cam", 3))]
            d = ddp(x)["_debug"]
            ((d["pose6d"] * A[rank:rank + 1]).sum() + (d["shape"] * B[rank:rank + 1]).sum() + (d["cam"] * C_[rank:rank + 1]).sum()).backward()
            st = m._train_state
            ok_all = all(p.grad is not None for p in m.parameters())
            in_flat = all(st.flat_grad.data_ptr() <= p.grad.data_ptr() < st.flat_grad.data_ptr() + 4 * st.flat_grad.numel()
                          for p in m.parameters())
            picks = ["encoder.patch_embed.backbone.stem.conv.weight", "encoder.blocks.0.attn.qkv.weight", "decoder.fc2.bias",
                     "encoder.pos_embed"]
            named = dict(m.named_parameters())
            q.put((rank, ok_all, in_flat, {k: named[k].grad.detach().numpy().copy() for k in picks}, None))   # numpy: pickled by value
    except Exception as e:                                                        # noqa: BLE001
        import traceback
        q.put((rank, False, False, {}, traceback.format_exc()[-2500:] + repr(e)))
    finally:
        dist.destroy_process_group()


def _overlap_worker(rank, world, port, q):
    """No torch DDP: overlap_gradient_allreduce() — the engine's backward progress hook launches one asynchronous all-reduce
    per finished range of the flat gradient buffer; allreduce_gradients() afterwards only waits for them."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "emu"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    os.environ.setdefault("MAED_EMU_THREADS", "4")
    torch.set_num_threads(4)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import harness
        from maed_b200 import train
        from maed_b200.models import MAED
        from oracle import synth
        with harness.product_on_cpu():
            m = MAED("ste", 2, 12, "vanilla", "ktd", 1024)
            synth.fill_module_(m, 21)
            m = m.train().enable_training(True, dropout_p=0.0)
            train.overlap_gradient_allreduce(m)
            x = synth.synth_frames(world, 1, 21)[rank:rank + 1]
            A, B, C_ = [synth.synth_tensor("grad_probe.%s" % k, (world, n), 21) for k, n in (("pose", 144), ("shape", 10), ("cam", 3))]
            d = m(x)["_debug"]
            ((d["pose6d"] * A[rank:rank + 1]).sum() + (d["shape"] * B[rank:rank + 1]).sum() + (d["cam"] * C_[rank:rank + 1]).sum()).backward()
            st = m._train_state
            n_works = len(st._works)                       # 2 blocks + the embeddings / backbone range
            train.allreduce_gradients(m)
            flat = st.grads_are_flat([p for _, p in m._train_param_order])
            picks = ["encoder.patch_embed.backbone.stem.conv.weight", "encoder.blocks.0.attn.qkv.weight",
                     "encoder.blocks.1.mlp.fc2.bias", "decoder.fc2.bias", "encoder.pos_embed", "encoder.cls_token"]
            named = dict(m.named_parameters())
            q.put((rank, n_works, flat, {k: named[k].grad.detach().numpy().copy() for k in picks}, None))
    except Exception as e:                                                        # noqa: BLE001
        import traceback
        q.put((rank, 0, False, {}, traceback.format_exc()[-2500:] + repr(e)))
    finally:
        dist.destroy_process_group()


def test_overlapped_gradient_allreduce_two_ranks_gloo():
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "emu"))
    import harness
    from maed_b200.models import MAED
    from oracle import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    with harness.product_on_cpu():                       # meanwhile: both frames in ONE process = 2 x the average
        m = MAED("ste", 2, 12, "vanilla", "ktd", 1024)
        synth.fill_module_(m, 21)
        m = m.train().enable_training(True, dropout_p=0.0)
        A, B, C_ = [synth.synth_tensor("grad_probe.%s" % k, (2, n), 21) for k, n in (("pose", 144), ("shape", 10), ("cam", 3))]
        d = m(synth.synth_frames(2, 1, 21))["_debug"]
        ((d["pose6d"] * A).sum() + (d["shape"] * B).sum() + (d["cam"] * C_).sum()).backward()
        ref = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    for rank, n_works, flat, grads, err in res:
        assert err is None, "rank %d: %s" % (rank, err)
        assert n_works == 3 and flat, (rank, n_works, flat)
        for k, g in grads.items():
            e = ((2.0 * torch.from_numpy(g) - ref[k]).norm() / ref[k].norm()).item()
            assert e < 2e-3, (rank, k, e)
    for k in res[0][3]:
        assert (res[0][3][k] == res[1][3][k]).all(), k


def test_torch_ddp_wraps_the_training_module_two_ranks_gloo():
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "emu"))
    import harness
    from maed_b200.models import MAED
    from oracle import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    # meanwhile: the same two frames in ONE process -> sum of the per-frame gradients = 2 x DDP's average
    with harness.product_on_cpu():
        m = MAED("ste", 1, 12, "vanilla", "ktd", 1024)
        synth.fill_module_(m, 21)
        m = m.train().enable_training(True, dropout_p=0.0)
        A, B, C_ = [synth.synth_tensor("grad_probe.%s" % k, (2, n), 21) for k, n in (("pose", 144), ("shape", 10), ("cam", 3))]
        d = m(synth.synth_frames(2, 1, 21))["_debug"]
        ((d["pose6d"] * A).sum() + (d["shape"] * B).sum() + (d["cam"] * C_).sum()).backward()
        ref = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    for rank, ok_all, in_flat, grads, err in res:
        assert err is None, "rank %d: %s" % (rank, err)
        assert ok_all and in_flat, (rank, ok_all, in_flat)
        for k, g in grads.items():
            e = ((2.0 * torch.from_numpy(g) - ref[k]).norm() / ref[k].norm()).item()
            assert e < 2e-3, (rank, k, e)
    for k in res[0][3]:
        assert (res[0][3][k] == res[1][3][k]).all(), k                    # both ranks hold the same averaged gradient
