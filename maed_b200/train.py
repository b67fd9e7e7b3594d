"""Training-side glue of the B200 MAED engine: autograd boundary, geometry tail, fused Adam, gradient all-reduce.

The hot path of a train step (reference lib/core/trainer.py:238-255: ``preds = model(inp); loss.backward(); optimizer.step()``)
runs in libmaed_b200.so:

  * ``MaedTrainFunction`` — one autograd node for the whole network up to the decoder outputs pose6d / shape / cam:
    forward = ``maed_train_forward`` (saved-activation tape in a workspace), backward = ``maed_train_backward`` (every
    parameter gradient written into one flat fp32 buffer; the returned gradients are views of it, so DDP hooks and
    ``optimizer.zero_grad()`` behave as with any nn.Module);
  * ``decode_outputs`` — the O(BT*24) geometry tail (rot6d -> rotmat -> angle-axis, projection; reference
    lib/utils/geometry.py:320-334,58-223, lib/models/spin.py:113-157) as two small autograd nodes over CUDA kernels
    (``csrc/decode_bwd.cu``), so that autograd links the reference's ``Loss`` to the engine boundary;
  * ``FusedAdam`` — torch.optim.Adam semantics (reference lib/utils/utils.py:127-131) in one kernel per step over the
    flat parameter / gradient / moment buffers;
  * ``allreduce_gradients`` — data-parallel gradient averaging over ``torch.distributed`` (NCCL over NVLink on the GPU
    box, gloo in the CPU tests): one all-reduce of the flat gradient buffer.

Several forwards may be outstanding before one backward (the reference's stage-2 iteration runs the video batch and the
image batch through the model, then ONE ``loss.backward()``: lib/core/trainer.py:186-202): every forward owns its tape
workspace (``TapePool``), and a backward that finds gradients already present accumulates instead of overwriting
(``TrainState.grad_targets``), so ``zero_grad(set_to_none=False)`` and micro-batch accumulation behave as with any nn.Module.
"""
import ctypes as C

import torch

from . import _lib

_TRAIN_MODES = ("parallel", "series", "vanilla", "temporal", "coupling")


# ----------------------------------------------------------------------------------------------- geometry tail
class _PoseTail(torch.autograd.Function):
    """(pose6d [R,144], shape [R,10], cam [R,3]) -> (theta [R,85] = cam | angle-axis | shape, rotmat [R,24,3,3]).
    reference lib/utils/geometry.py:320-334 (rot6d -> rotation matrix) and :58-223 (-> angle-axis), lib/models/ktd.py:116-118.
    Forward: the inference engine's kernel (maed_op_decode_outputs); backward: csrc/decode_bwd.cu (forward-mode derivative of
    the same arithmetic), one launch."""

    @staticmethod
    def forward(ctx, pose6d, shape, cam):
        if not pose6d.is_cuda:
            raise RuntimeError("maed_b200 geometry tail runs on CUDA only; got a %s tensor — there is no CPU fallback" % pose6d.device)
        pose6d, shape, cam = [t.contiguous().float() for t in (pose6d, shape, cam)]
        R = pose6d.shape[0]
        f32 = dict(dtype=torch.float32, device=pose6d.device)
        rot, theta, dummy = torch.empty(R, 24, 3, 3, **f32), torch.empty(R, 85, **f32), torch.empty(R, 1, 2, **f32)
        with torch.cuda.device(pose6d.device):
            _lib.call("maed_op_decode_outputs", _lib.ptr(pose6d), _lib.ptr(shape), _lib.ptr(cam), R, None, 1, _lib.ptr(rot),
                      _lib.ptr(theta), _lib.ptr(dummy), _lib.stream_ptr())
        ctx.save_for_backward(pose6d)
        return theta, rot

    @staticmethod
    def backward(ctx, d_theta, d_rot):
        (pose6d,) = ctx.saved_tensors
        R = pose6d.shape[0]
        d_theta = d_theta.contiguous().float() if d_theta is not None else None
        d_rot = d_rot.contiguous().float() if d_rot is not None else None
        d_pose = torch.empty_like(pose6d)
        with torch.cuda.device(pose6d.device):
            d_aa = C.c_void_p(d_theta.data_ptr() + 12) if d_theta is not None else None       # theta[:, 3:75]
            _lib.call("maed_decode_pose_backward", _lib.ptr(pose6d), R, _lib.ptr(d_rot), d_aa, 85, _lib.ptr(d_pose),
                      _lib.stream_ptr())
        if d_theta is None:
            return d_pose, None, None
        return d_pose, d_theta[:, 75:], d_theta[:, :3]


class _Project(torch.autograd.Function):
    """kp_2d = weak-perspective projection of kp_3d (or of zero joints) by cam, reference lib/models/spin.py:113-157."""

    @staticmethod
    def forward(ctx, kp3d, cam, n_joints):
        cam = cam.contiguous().float()
        if not cam.is_cuda:
            raise RuntimeError("maed_b200 geometry tail runs on CUDA only; got a %s tensor — there is no CPU fallback" % cam.device)
        kp3d = kp3d.contiguous().float() if kp3d is not None else None
        R = cam.shape[0]
        kp2d = torch.empty(R, n_joints, 2, dtype=torch.float32, device=cam.device)
        with torch.cuda.device(cam.device):
            _lib.call("maed_project_keypoints", _lib.ptr(kp3d), _lib.ptr(cam), R, n_joints, _lib.ptr(kp2d), None, None, None,
                      _lib.stream_ptr())
        ctx.save_for_backward(cam) if kp3d is None else ctx.save_for_backward(cam, kp3d)
        ctx.n_joints = n_joints
        return kp2d

    @staticmethod
    def backward(ctx, d_kp2d):
        saved = ctx.saved_tensors
        cam, kp3d = saved[0], (saved[1] if len(saved) > 1 else None)
        R = cam.shape[0]
        d_kp2d = d_kp2d.contiguous().float()
        d_cam = torch.empty_like(cam)
        d_kp3d = torch.empty_like(kp3d) if kp3d is not None and ctx.needs_input_grad[0] else None
        with torch.cuda.device(cam.device):
            _lib.call("maed_project_keypoints", _lib.ptr(kp3d), _lib.ptr(cam), R, ctx.n_joints, None, _lib.ptr(d_kp2d),
                      _lib.ptr(d_cam), _lib.ptr(d_kp3d), _lib.stream_ptr())
        return d_kp3d, d_cam, None


def _smpl_assets(head, dev):
    if head.v_template.device != dev:
        head.to(dev)
    return _lib.MaedSmplAssets(*[_lib.ptr(getattr(head, k)) for k in (
        "v_template", "shapedirs", "posedirs", "J_template", "J_shapedirs", "lbs_weights", "J_regressor_extra", "parents",
        "extra_vertex_ids", "joint_map")])


class _SmplBody(torch.autograd.Function):
    """(betas [R,10], rotmat [R,24,3,3]) -> (verts [R,6890,3], joints [R,49,3] or J_regressor @ verts): the body model of the
    decoders (reference lib/models/ktd.py:100-114 -> lib/models/smpl.py:84-106 -> smplx.lbs.lbs) with BOTH directions in
    csrc/smpl.cu (maed_smpl_forward / maed_smpl_backward), so the keypoint losses reach pose and shape through CUDA kernels."""

    @staticmethod
    def forward(ctx, betas, rotmat, head, J_regressor):
        if not betas.is_cuda:
            raise RuntimeError("maed_b200 SMPL kernels run on CUDA only; got a %s tensor — there is no CPU fallback" % betas.device)
        betas, rotmat = betas.contiguous().float(), rotmat.contiguous().float()
        dev, R = betas.device, betas.shape[0]
        reg = J_regressor.to(dev, torch.float32).contiguous() if J_regressor is not None else None
        nj = reg.shape[0] if reg is not None else head.n_joints
        f32 = dict(dtype=torch.float32, device=dev)
        verts, joints = torch.empty(R, 6890, 3, **f32), torch.empty(R, nj, 3, **f32)
        with torch.cuda.device(dev):
            assets = _smpl_assets(head, dev)
            nbytes = _lib.load().maed_smpl_scratch_bytes(R)
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.call("maed_smpl_forward", C.byref(assets), _lib.ptr(betas), _lib.ptr(rotmat), R, _lib.ptr(reg),
                      0 if reg is None else reg.shape[0], _lib.ptr(verts), _lib.ptr(joints), _lib.ptr(scratch), C.c_size_t(nbytes),
                      _lib.stream_ptr())
        ctx.save_for_backward(betas, rotmat)
        ctx.head, ctx.reg = head, reg
        return verts, joints

    @staticmethod
    def backward(ctx, d_verts, d_joints):
        betas, rotmat = ctx.saved_tensors
        head, reg = ctx.head, ctx.reg
        dev, R = betas.device, betas.shape[0]
        nj = reg.shape[0] if reg is not None else head.n_joints
        d_verts = d_verts.contiguous().float() if d_verts is not None else None
        d_joints = d_joints.contiguous().float() if d_joints is not None else torch.zeros(R, nj, 3, dtype=torch.float32, device=dev)
        d_betas, d_rot = torch.empty_like(betas), torch.empty_like(rotmat)
        with torch.cuda.device(dev):
            assets = _smpl_assets(head, dev)
            nbytes = _lib.load().maed_smpl_backward_scratch_bytes(R)
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.call("maed_smpl_backward", C.byref(assets), _lib.ptr(betas), _lib.ptr(rotmat), R, _lib.ptr(reg),
                      0 if reg is None else reg.shape[0], _lib.ptr(d_verts), _lib.ptr(d_joints), _lib.ptr(d_betas), _lib.ptr(d_rot),
                      _lib.ptr(scratch), C.c_size_t(nbytes), _lib.stream_ptr())
        return d_betas, d_rot, None, None


def decode_outputs(pose6d, shape, cam, n_joints=49, smpl_head=None, J_regressor=None):
    """reference lib/models/ktd.py:94-124.  Without a body model (see SMPLHead) verts / joints are zeros."""
    nt = pose6d.shape[0]
    theta, rot = _PoseTail.apply(pose6d, shape, cam)
    if smpl_head is not None and smpl_head.has_assets:
        verts, kp3d = _SmplBody.apply(shape, rot, smpl_head, J_regressor)
        kp2d = _Project.apply(kp3d, cam, kp3d.shape[1])
    else:
        verts, kp3d = pose6d.new_zeros(nt, 6890, 3), pose6d.new_zeros(nt, n_joints, 3)
        kp2d = _Project.apply(None, cam, n_joints)
    return {"theta": theta, "verts": verts, "kp_2d": kp2d, "kp_3d": kp3d, "rotmat": rot}


# --------------------------------------------------------------------------------------------- engine state
class TapePool:
    """Tape workspaces of the outstanding training forwards.  A forward takes the smallest free buffer that fits (or
    allocates one); its autograd node holds it through a `_Tape` and returns it after the backward — or when the graph is
    dropped without a backward (`_Tape.__del__`).  At most `keep` idle buffers stay cached."""

    def __init__(self, keep=2):
        self.free, self.keep = [], keep

    def take(self, nbytes, dev):
        best = None
        for i, t in enumerate(self.free):
            if t.numel() >= nbytes and t.device == dev and (best is None or t.numel() < self.free[best].numel()):
                best = i
        if best is not None:
            return self.free.pop(best)
        self.free = [t for t in self.free if t.device == dev]        # stale devices / too small: let them go first
        if len(self.free) >= self.keep:
            self.free.pop(0)
        return torch.empty(nbytes, dtype=torch.uint8, device=dev)

    def give(self, t):
        if len(self.free) < self.keep:
            self.free.append(t)


class _Tape:
    def __init__(self, pool, ws):
        self.pool, self.ws = pool, ws

    def release(self):
        if self.ws is not None:
            self.pool.give(self.ws)
            self.ws = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def _mix64(*vals):
    """splitmix64-style hash of a few integers -> dropout seed (steps, ranks and models get unrelated mask streams)."""
    h = 0x9E3779B97F4A7C15
    for v in vals:
        h = (h ^ (int(v) & 0xFFFFFFFFFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF
        h = (h + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        h = ((h ^ (h >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        h = ((h ^ (h >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        h ^= h >> 31
    return h


class TrainState:
    """Per-model buffers of the training path: flat gradient buffer, derived dgrad weights, tape workspaces."""

    def __init__(self, model):
        self.model = model
        self.flat_grad = None
        self.spares = []                  # [(flat buffer, pointer array)]: targets of backwards that must not touch flat_grad
        self.in_flight = set()            # handed to autograd, not yet accumulated (0 = flat_grad, k = spares[k - 1])
        self._sentinel = None
        self._sentinel_handle = None
        self.grad_offsets = None
        self.grad_ptrs = None
        self.tpack = None
        self.tpack_key = None
        self.tapes = TapePool()
        self.loss_scale = 4096.0
        self.step_seed = 0
        self._exchange = None
        # overlapped data-parallel gradient exchange (overlap_gradient_allreduce)
        self.overlap = False
        self._progress = None             # ctypes thunk of the engine's backward progress hook (kept alive here)
        self._overlap_live = False        # this backward writes flat_grad itself under an initialised process group
        self._works = []                  # outstanding async all-reduces of this backward
        self._avg_in_collective = True

    def set_overlap(self, model, on):
        """Installs / removes the engine's backward progress hook: as soon as the kernels producing a contiguous range of the
        flat gradient buffer have been enqueued, that range is all-reduced asynchronously (NCCL runs it on its own stream
        behind the compute stream's work so far), so the exchange of the STE blocks' gradients overlaps the rest of the
        backward — what DDP's bucketed all-reduce does for the reference (train.py:113)."""
        import torch.distributed as dist
        self.overlap = bool(on)
        if not on:
            if self._progress is not None:
                _lib.call("maed_train_set_progress", model._engine, _lib.PROGRESS_FN(0), None)
                self._progress = None
            return

        def progress(_user, first, end):
            try:
                if not self._overlap_live:
                    return 0
                offs = self.grad_offsets
                lo = next(o for o in offs[first:] if o is not None)
                hi = next((o for o in offs[end:] if o is not None), self.flat_grad.numel())
                if hi > lo:
                    op = dist.ReduceOp.AVG if self._avg_in_collective else dist.ReduceOp.SUM
                    self._works.append(dist.all_reduce(self.flat_grad[lo:hi], op=op, async_op=True))
                return 0
            except Exception:                           # an exception must not unwind through the C frames
                import traceback
                traceback.print_exc()
                return 1

        model._get_engine()
        self._progress = _lib.PROGRESS_FN(progress)
        _lib.call("maed_train_set_progress", model._engine, self._progress, None)

    def ensure_exchange(self, model, dev):
        """SyncBatchNorm for encoder='cnn' under data parallelism (the reference converts every BatchNorm with
        nn.SyncBatchNorm.convert_sync_batchnorm, train.py:95): the engine hands the per-channel sums of each BatchNorm to this
        callback, which adds them up over the ranks with ONE small all-reduce (NCCL over NVLink on the GPU box, stream-ordered;
        gloo in the CPU tests).  `model.sync_batchnorm = False` keeps per-rank statistics."""
        import torch.distributed as dist
        want = (model.encoder_type.lower() == "cnn" and getattr(model, "sync_batchnorm", True) and dist.is_available()
                and dist.is_initialized() and dist.get_world_size() > 1)
        if want and self._exchange is None:
            buf = torch.zeros(2 * 2048 + 1, dtype=torch.float64, device=dev)

            def exchange(_user, n):
                try:
                    dist.all_reduce(buf[:n])
                    return 0
                except Exception:                       # an exception must not unwind through the C frames
                    return 1

            fn = _lib.EXCHANGE_FN(exchange)
            _lib.call("maed_train_set_exchange", model._engine, fn, None, _lib.ptr(buf), buf.numel())
            self._exchange = (buf, fn)                  # keep the buffer and the ctypes thunk alive
        elif not want and self._exchange is not None:
            _lib.call("maed_train_set_exchange", model._engine, _lib.EXCHANGE_FN(0), None, None, 0)
            self._exchange = None

    def _layout(self, tensors, is_param):
        """(Re)allocates the flat gradient buffer: engine table order, every PARAMETER padded to 4 elements; buffers of the
        table — the BatchNorm running statistics of encoder='cnn' — get no slot and a NULL pointer."""
        dev = tensors[0].device
        total = sum((t.numel() + 3) // 4 * 4 for t, p in zip(tensors, is_param) if p)
        if self.flat_grad is None or self.flat_grad.numel() != total or self.flat_grad.device != dev:
            self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
            self.spares = []
            self.in_flight.clear()
            self.grad_offsets, off = [], 0
            for t, p in zip(tensors, is_param):
                self.grad_offsets.append(off if p else None)
                off += (t.numel() + 3) // 4 * 4 if p else 0
            self.grad_ptrs = self._ptr_array(self.flat_grad)

    def _ptr_array(self, flat):
        base = flat.data_ptr()
        return (C.c_void_p * len(self.grad_offsets))(*[None if o is None else base + 4 * o for o in self.grad_offsets])

    def _views(self, flat, tensors):
        return [None if o is None else flat[o:o + t.numel()].view(t.shape) for o, t in zip(self.grad_offsets, tensors)]

    def grad_targets(self, tensors, is_param, params):
        """Where this backward writes, and what it hands to autograd: (pointer array for the engine, gradient tensors in
        table order or None, buffer to add the result into or None).

        Normal step (no gradient present on any parameter, nothing parked): the engine writes the flat buffer itself and
        autograd receives FRESH views of it — AccumulateGrad adopts a gradient tensor nobody else references instead of
        cloning it, so ``p.grad`` aliases the flat buffer (one all-reduce / one Adam launch covers all 72 M parameters).
        Otherwise the engine writes a spare flat buffer:
          * a further forward's node in the same backward pass (the reference's video + image iteration): the first node's
            views are still parked in autograd's input buffers (``in_flight``; a post-accumulate hook on one sentinel
            parameter tells when they have been consumed).  The spare is added to the parked buffer on the device (one
            launch) and autograd receives no gradients from this node — its own out-of-place summation would detach
            ``p.grad`` from the flat buffer;
          * micro-batch accumulation / ``zero_grad(set_to_none=False)``: autograd receives views of the spare and
            AccumulateGrad adds them to the existing ``p.grad`` in place."""
        self._layout(tensors, is_param)
        sentinel = next((p for p in params if p.requires_grad), None)
        if sentinel is not None and self._sentinel is not sentinel:
            if self._sentinel_handle is not None:
                self._sentinel_handle.remove()
            self._sentinel_handle = sentinel.register_post_accumulate_grad_hook(lambda _p: self.in_flight.clear())
            self._sentinel = sentinel

        def buffer(k):
            while k > len(self.spares):
                buf = torch.zeros_like(self.flat_grad)
                self.spares.append((buf, self._ptr_array(buf)))
            return (self.flat_grad, self.grad_ptrs) if k == 0 else self.spares[k - 1]

        if self.in_flight:
            parked = next(iter(self.in_flight))
            buf, ptrs = buffer(1 if parked != 1 else 2)
            return ptrs, None, buffer(parked)[0]
        k = 0 if all(p.grad is None for p in params) else 1
        self.in_flight.add(k)
        buf, ptrs = buffer(k)
        return ptrs, self._views(buf, tensors), None

    def flatten_grads(self, params_in_table_order):
        """Re-binds every .grad to its slot of the flat buffer (copying if it lives elsewhere).  No-op in the normal case;
        needed when autograd had to sum out of place (p.grad is then an ordinary tensor).  False if a gradient is missing."""
        if self.grads_are_flat(params_in_table_order):
            return True
        if self.flat_grad is None or any(p.grad is None for p in params_in_table_order):
            return False
        offs = [o for o in self.grad_offsets if o is not None]
        if len(offs) != len(params_in_table_order):
            return False
        with torch.no_grad():
            for p, o in zip(params_in_table_order, offs):
                v = self.flat_grad[o:o + p.numel()].view(p.shape)
                if p.grad.data_ptr() != v.data_ptr():
                    v.copy_(p.grad)
                    p.grad = v
        return True

    def grads_are_flat(self, params_in_table_order):
        """True when every parameter's .grad is the view of the flat buffer at its slot (what FusedAdam's one-launch step and
        the one-call all-reduce rely on)."""
        if self.flat_grad is None:
            return False
        base = self.flat_grad.data_ptr()
        offs = [o for o in self.grad_offsets if o is not None]
        if len(offs) != len(params_in_table_order):
            return False
        return all(p.grad is not None and p.grad.data_ptr() == base + 4 * o for p, o in zip(params_in_table_order, offs))

    def ensure_tpack(self, eng, params_arr, key, dev):
        lib = _lib.load()
        if key != self.tpack_key or self.tpack is None:
            nbytes = lib.maed_train_pack_bytes(eng)
            if self.tpack is None or self.tpack.numel() < nbytes or self.tpack.device != dev:
                self.tpack = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.call("maed_train_pack", eng, params_arr, _lib.ptr(self.tpack), _lib.stream_ptr())
            self.tpack_key = key


def _rank():
    import torch.distributed as dist
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def _weights_key(model, params):
    """Changes whenever a parameter is written: autograd's version counters (torch optimisers, in-place edits) plus the
    counter FusedAdam bumps (its kernel writes parameter memory behind those counters)."""
    return (getattr(model, "_weights_gen", 0), sum(p._version for p in params))


class MaedTrainFunction(torch.autograd.Function):
    """(x, *params) -> (pose6d, shape, cam) with the whole network in between (engine tape + backward)."""

    @staticmethod
    def forward(ctx, model, x, dropout_p, *params):
        st = model._train_state
        N, T = x.shape[:2]
        dev = x.device
        with torch.cuda.device(dev):
            eng = model._prepare(dev)                   # engine + forward packed weights (version-keyed cache)
            # keyed by the pack generation, not by _packed_key: FusedAdam updates parameters behind autograd's version
            # counters, so the key can repeat although the weights changed (invalidate_cache() forces a re-pack)
            st.ensure_tpack(eng, model._param_ptrs, model._pack_gen, dev)
            st.ensure_exchange(model, dev)
            nbytes = _lib.load().maed_train_workspace_bytes(eng, N * T)
            tape = _Tape(st.tapes, st.tapes.take(nbytes, dev))          # this forward's own tape (see TapePool)
            f32 = dict(dtype=torch.float32, device=dev)
            pose = torch.empty(N * T, 144, **f32)
            shape = torch.empty(N * T, 10, **f32)
            cam = torch.empty(N * T, 3, **f32)
            outs = _lib.MaedTrainOutputs(None, _lib.ptr(pose), _lib.ptr(shape), _lib.ptr(cam))
            st.step_seed += 1
            st.in_flight.clear()                        # no backward pass is running while a forward is being recorded
            seed = _mix64(torch.initial_seed(), st.step_seed, _rank())   # per step AND per data-parallel rank
            _lib.call("maed_train_forward", eng, model._param_ptrs, _lib.ptr(model._packed), _lib.ptr(x), N, T, _lib.ptr(tape.ws),
                      C.c_size_t(tape.ws.numel()), C.c_float(dropout_p), C.c_ulonglong(seed), C.byref(outs), _lib.stream_ptr())
        ctx.model, ctx.x, ctx.dropout_p, ctx.n_params = model, x, dropout_p, len(params)
        ctx.tape, ctx.weights_key = tape, _weights_key(model, params)
        return pose, shape, cam

    @staticmethod
    def backward(ctx, d_pose, d_shape, d_cam):
        model, x, tape = ctx.model, ctx.x, ctx.tape
        st = model._train_state
        if tape is None or tape.ws is None:
            raise RuntimeError("maed_b200: backward() a second time through the same forward — the activation tape was "
                               "released after the first backward (retain_graph is not supported)")
        if ctx.weights_key != _weights_key(model, [p for _, p in model._train_param_order]):
            raise RuntimeError("maed_b200: the parameters changed between this forward and its backward (optimizer.step() or "
                               "an in-place edit): the taped activations no longer match the weights")
        N, T = x.shape[:2]
        dev = x.device
        tensors = model._tensor_table()
        param_names = {n for n, _ in model._train_param_order}
        params = [p for _, p in model._train_param_order]
        ptrs, views, add_into = st.grad_targets(tensors, [n in param_names for n in model._param_names], params)
        st._works = []
        st._overlap_live = False
        if st.overlap and ptrs is st.grad_ptrs:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                st._overlap_live = True
                st._avg_in_collective = dist.get_backend() == "nccl"        # gloo has no AVG: sum, then divide
        z = lambda g, n: (torch.zeros(N * T, n, dtype=torch.float32, device=dev) if g is None  # noqa: E731
                          else g.contiguous().float())
        d_pose, d_shape, d_cam = z(d_pose, 144), z(d_shape, 10), z(d_cam, 3)
        with torch.cuda.device(dev):
            _lib.call("maed_train_backward", model._engine, model._param_ptrs, _lib.ptr(model._packed), _lib.ptr(st.tpack),
                      _lib.ptr(x), N, T, _lib.ptr(tape.ws), C.c_size_t(tape.ws.numel()), _lib.ptr(d_pose),
                      _lib.ptr(d_shape), _lib.ptr(d_cam), C.c_float(st.loss_scale), C.c_float(ctx.dropout_p), ptrs,
                      _lib.stream_ptr())
        st._overlap_live = False
        tape.release()
        ctx.tape = None
        if add_into is not None:                           # a parked buffer of this backward pass collects the sum
            spare = st.spares[0][0] if ptrs is st.spares[0][1] else st.spares[1][0]
            add_into.add_(spare)
            return (None, None, None) + (None,) * ctx.n_params
        by_name = dict(zip(model._param_names, views))
        accumulate = ptrs is not st.grad_ptrs             # a spare buffer: its views must not become a p.grad
        grads = []
        for name, p in model._train_param_order:
            g = by_name[name] if p.requires_grad else None
            if g is not None and accumulate and p.grad is None:
                g = g.clone()                           # would be adopted and then overwritten by the next spare write
            grads.append(g)
        return (None, None, None) + tuple(grads)


def train_forward(model, x, J_regressor=None):
    """MAED.forward in train() mode with autograd enabled (called from maed_b200.models.maed.MAED.forward)."""
    is_cnn = model.encoder_type.lower() == "cnn"
    if not is_cnn and model._cfg.mode not in (_lib.MODES[m] for m in _TRAIN_MODES):
        raise NotImplementedError("maed_b200 training supports st_mode in %s (or encoder='cnn')" % (_TRAIN_MODES,))
    if model.precision != "split":
        raise NotImplementedError("maed_b200 training runs in precision='split'")
    if not x.is_cuda:
        raise RuntimeError("maed_b200.MAED runs on CUDA (sm_100a) only; got a %s tensor — there is no CPU fallback" % x.device)
    N, T = x.shape[:2]
    x = x.to(torch.float32).contiguous()
    model._get_engine()
    if getattr(model, "_train_state", None) is None:
        model._train_state = TrainState(model)
    named = dict(model.named_parameters())
    model._train_param_order = [(n, named[n]) for n in model._param_names if n in named]
    params = [p for _, p in model._train_param_order]
    dropout_p = model._train_dropout_p if model._train_dropout_p is not None else 0.5   # nn.Dropout() default (ktd.py:54-56)
    pose, shape, cam = MaedTrainFunction.apply(model, x, dropout_p, *params)
    if is_cnn:
        # nn.BatchNorm2d.train(): the engine has updated running_mean / running_var in place (momentum 0.1, statistics of
        # THIS rank's batch — no SyncBatchNorm exchange); count the step and drop the eval-mode weight pack that folds them
        with torch.no_grad():
            torch._foreach_add_([b for n, b in model.named_buffers() if n.endswith("num_batches_tracked")], 1)
        model.invalidate_cache()
    nj = 17 if J_regressor is not None else model.decoder.smpl.n_joints
    o = decode_outputs(pose, shape, cam, nj, model.decoder.smpl, J_regressor)
    return {"theta": o["theta"].reshape(N, T, -1), "verts": o["verts"].reshape(N, T, -1, 3),
            "kp_2d": o["kp_2d"].reshape(N, T, -1, 2), "kp_3d": o["kp_3d"].reshape(N, T, -1, 3),
            "rotmat": o["rotmat"].reshape(N, T, -1, 3, 3),
            "_debug": {"pose6d": pose, "shape": shape, "cam": cam}}


# ------------------------------------------------------------------------------------------------- optimiser
class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (L2 weight decay folded into the gradient, bias-corrected moments; reference
    lib/utils/utils.py:127-131) on the engine's Adam kernel.  Accepts the reference's per-parameter groups
    (``[{'params': p, 'name': n} ...]``).  With ``FusedAdam.for_model(model, ...)`` the parameters are flattened into
    one buffer laid out like the engine's flat gradient buffer and the whole step is ONE kernel launch; otherwise one
    launch per parameter tensor.  Pass ``model=`` so the packed tensor-core weights are re-derived after the step
    (the kernel writes parameter memory behind autograd's version counters).

    ``state_dict()`` / ``load_state_dict()`` use torch.optim.Adam's layout (per parameter ``step``, ``exp_avg``,
    ``exp_avg_sq``), so the reference's checkpoints resume here and vice versa (lib/core/trainer.py:335,359); in the flat
    mode the per-parameter moments are views of the flat moment buffers."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, model=None):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._model = model
        self._flat = None

    @classmethod
    def for_model(cls, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        """Flattens the model's parameters (engine order, each padded to a multiple of 4 elements — the layout of
        TrainState.flat_grad) and builds the single-launch optimiser.  Call it AFTER ``model.to(device)``: moving the
        module afterwards re-allocates the parameters outside the flat buffer (``step()`` checks and says so)."""
        model._get_engine()
        tensors = model._tensor_table()
        # parameters only, in engine table order: the table of encoder='cnn' interleaves BatchNorm running buffers with the
        # parameters, and Adam (weight decay!) must not touch those — same filter as TrainState.grad_targets
        named = dict(model.named_parameters())
        names = [n for n in model._param_names if n in named]
        params = [named[n] for n in names]
        total = sum((p.numel() + 3) // 4 * 4 for p in params)
        flat = torch.zeros(total, dtype=torch.float32, device=tensors[0].device)
        offsets, off = [], 0
        with torch.no_grad():
            for p in params:
                n = p.numel()
                flat[off:off + n].copy_(p.reshape(-1))
                p.data = flat[off:off + n].view(p.shape)
                offsets.append(off)
                off += (n + 3) // 4 * 4
        model.invalidate_cache()
        opt = cls([{"params": p, "name": n} for n, p in model.named_parameters()], lr=lr, betas=betas, eps=eps,
                  weight_decay=weight_decay, model=model)
        opt._flat = {"p": flat, "m": torch.zeros_like(flat), "v": torch.zeros_like(flat), "step": 0, "params": params,
                     "offsets": offsets}
        opt._bind_flat_state()
        return opt

    def _bind_flat_state(self):
        """Optimizer.state[p] = views of the flat moment buffers (what state_dict() serialises)."""
        f = self._flat
        for p, off in zip(f["params"], f["offsets"]):
            n = p.numel()
            self.state[p] = {"step": torch.tensor(float(f["step"])), "exp_avg": f["m"][off:off + n].view(p.shape),
                             "exp_avg_sq": f["v"][off:off + n].view(p.shape)}

    def state_dict(self):
        if self._flat is not None:
            for p in self._flat["params"]:
                self.state[p]["step"] = torch.tensor(float(self._flat["step"]))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)             # replaces self.state by copies of the loaded tensors
        f = self._flat
        if f is None:
            for s in self.state.values():
                if "step" in s:
                    s["step"] = int(float(s["step"]))
            return
        steps = set()
        with torch.no_grad():
            for p, off in zip(f["params"], f["offsets"]):
                s = self.state.get(p)
                if not s:
                    continue
                n = p.numel()
                f["m"][off:off + n].copy_(s["exp_avg"].reshape(-1))
                f["v"][off:off + n].copy_(s["exp_avg_sq"].reshape(-1))
                steps.add(int(float(s["step"])))
        if len(steps) > 1:
            raise RuntimeError("FusedAdam (flat): the loaded per-parameter step counts differ: %s" % sorted(steps))
        f["step"] = steps.pop() if steps else 0
        self._bind_flat_state()

    def _kernel(self, p, g, m, v, n, group, step, grad_scale):
        b1, b2 = group["betas"]
        with torch.cuda.device(p.device):
            _lib.call("maed_adam_step", _lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), C.c_longlong(n),
                      C.c_double(group["lr"]), C.c_double(b1), C.c_double(b2), C.c_double(group["eps"]),
                      C.c_double(group["weight_decay"]), step, C.c_float(grad_scale), _lib.stream_ptr())

    def _flat_ready(self, st):
        """The one-launch path needs (a) every parameter still inside the flat parameter buffer, (b) every .grad the view of
        the engine's flat gradient buffer at the same offset, (c) one set of hyper-parameters."""
        f = self._flat
        base = f["p"].data_ptr()
        if any(p.data_ptr() != base + 4 * o for p, o in zip(f["params"], f["offsets"])):
            raise RuntimeError("FusedAdam.for_model: the model's parameters no longer live in the optimiser's flat buffer "
                               "(model.to()/.cuda()/load of new tensors after for_model?) — build the optimiser after moving "
                               "the model")
        if st is None or not st.flatten_grads(f["params"]):
            return False
        g0 = self.param_groups[0]
        key = lambda g: (g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"])  # noqa: E731
        return all(key(g) == key(g0) for g in self.param_groups)

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = closure() if closure is not None else None
        st = getattr(self._model, "_train_state", None) if self._model is not None else None
        if self._flat is not None and self._flat_ready(st):
            f = self._flat
            f["step"] += 1
            self._kernel(f["p"], st.flat_grad, f["m"], f["v"], f["p"].numel(), self.param_groups[0], f["step"], grad_scale)
        else:
            if self._flat is not None:
                self._flat["step"] += 1
            for group in self.param_groups:
                for p in group["params"]:
                    if p.grad is None:
                        continue
                    if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                        raise RuntimeError("FusedAdam: parameters must be contiguous CUDA float32 tensors")
                    s = self.state[p]
                    if not s:
                        s["step"], s["exp_avg"], s["exp_avg_sq"] = 0, torch.zeros_like(p), torch.zeros_like(p)
                    step = self._flat["step"] if self._flat is not None else int(float(s["step"])) + 1
                    s["step"] = step
                    g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                    self._kernel(p, g, s["exp_avg"], s["exp_avg_sq"], p.numel(), group, step, grad_scale)
        if self._model is not None:
            self._model._weights_gen = getattr(self._model, "_weights_gen", 0) + 1
            self._model.invalidate_cache()
        return loss


def overlap_gradient_allreduce(model, on=True):
    """Data-parallel training without torch DDP: exchange the gradients WHILE the backward is still running.  After this
    call every ``loss.backward()`` through the model all-reduces the flat gradient buffer range by range as the engine finishes
    it (STE block by block, then embeddings + backbone); ``allreduce_gradients(model)`` after the backward then only waits for
    those collectives (stream-ordered) instead of reducing 288.5 MB in one exposed call."""
    model._get_engine()
    if getattr(model, "_train_state", None) is None:
        model._train_state = TrainState(model)
    model._train_state.set_overlap(model, on)
    return model


def allreduce_gradients(model, world_size=None):
    """Average the parameter gradients over the data-parallel ranks with ONE all-reduce of the flat gradient buffer
    (reference: DDP's bucketed all-reduce, train.py:113; 288.5 MB fp32 per step) — or, after overlap_gradient_allreduce(),
    wait for the range-wise all-reduces the backward already launched.  No-op without torch.distributed."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return
    ws = world_size or dist.get_world_size()
    st = getattr(model, "_train_state", None)
    if st is not None and st._works:
        for w in st._works:
            w.wait()                                    # stream-ordered on NCCL: later kernels wait, the host does not
        st._works = []
        if not st._avg_in_collective:
            st.flat_grad.div_(ws)
        return
    order = getattr(model, "_train_param_order", None)
    if st is not None and order is not None and st.flatten_grads([p for _, p in order]):
        dist.all_reduce(st.flat_grad)
        st.flat_grad.div_(ws)
        return
    for p in model.parameters():
        if p.grad is not None:
            dist.all_reduce(p.grad)
            p.grad.div_(ws)
