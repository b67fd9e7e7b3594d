#!/usr/bin/env python
"""bench.py — throughput of the MAED per-clip forward hot path on B200 (clips/s), driver contract of the task.

  python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path (default)
  python bench.py --impl reference ...                       # the reference's algorithm on the host CPU cores
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU

Workload (BASELINE.json configs[1]): per GPU 8 clips x T=16 frames of 224x224 synthetic (N(0,1)) frames,
MAED(encoder='ste', 6 blocks, 12 heads, st_mode='parallel', decoder='ktd'), random-init weights, forward only
(eval / inference), split-fp16 tensor-core precision (the mode that passes the 1e-3 parity gate).
A "step" = one forward over the batch of 8 clips.  Weak scaling: each rank runs its own 8 clips, the path
shards by clip, there is no data-path collective (only the timing barrier / max-reduce).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIPS_PER_GPU, T = 8, 16
MODE, DECODER = "parallel", "ktd"
GFLOP_PER_CLIP = 411.35          # algorithmic 2*MACs, forward, parallel mode (SURVEY.md §8d / BASELINE.md §2)
METRIC = "clips/sec (T=16, 224x224, bs=8/gpu), MAED ste-parallel+ktd forward"   # ONE string for both arms (the driver pairs them by it)
WORKLOAD = "configs[1]: bs=8/gpu T=16 224x224 synthetic clips, ResNetV2(3,4,9)+STE-parallel(6x12)+KTD forward"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled every 20 ms DURING the timed region, through NVML in a thread of this process.

    (Round 1-2 spawned `nvidia-smi -lms 200` right before the timed region: its start-up — NVML init, device enumeration under
    the driver lock — landed inside the 0.3 s region, gave 2-3 samples and slowed the launches it was meant to observe: the
    device-timed step read 14.9-16.7 ms on boxes whose end-to-end loop, measured later without it, ran at 14.4 ms.)  NVML is
    initialised in the constructor, i.e. before the warm-up; a sample is three NVML getters.  Falls back to the nvidia-smi
    subprocess — started in the constructor, well before the timed region — when pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path, self.nv, self.h = gpu_index, None, None, None, None
        self.samples, self.on, self.thread, self.t0 = [], False, None, None
        self.period = float(os.environ.get("MAED_BENCH_CLOCK_MS", "20")) / 1000.0      # 0: no sampling (A/B of the sampler's own cost)
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nv, self.h = pynvml, h
        except Exception:
            try:
                fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
                os.close(fd)
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=timestamp," + self.Q,
                                              "--format=csv,noheader,nounits", "-lms", "100"],
                                             stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            except Exception:
                self.proc = None

    def _poll(self):
        nv, h = self.nv, self.h
        while self.on:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.samples.append((mhz, mask))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        self.t0 = time.time()
        if self.nv is not None and self.period > 0:
            import threading
            self.on = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self):
        t1 = time.time()
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm, mx, reasons = [], [], set()
        if self.nv is not None:
            self.on = False
            if self.thread is not None:
                self.thread.join(timeout=1.0)
            for mhz, mask in self.samples:
                sm.append(mhz)
                mx.append(self.max_mhz)
                for name, bit in self.BITS:
                    if mask & bit:
                        reasons.add(name)
            out["source"] = "nvml, 20 ms"
        elif self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            import datetime
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    if self.t0 is not None and not (self.t0 - 0.1 <= ts <= t1 + 0.1):
                        continue                              # sample outside the timed region
                    sm.append(float(f[2])); mx.append(float(f[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
            out["source"] = "nvidia-smi -lms 100 (started before the warm-up)"
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [s.strip() for s in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def _profile(on):
    """MAED_BENCH_PROFILE=1: cudaProfilerStart/Stop around the device-timed region, so that
    `ncu --profile-from-start off` lists exactly the launches of the timed steps (a number printed under ncu is never a
    bench value)."""
    if os.environ.get("MAED_BENCH_PROFILE"):
        torch.cuda.synchronize()
        (torch.cuda.profiler.start if on else torch.cuda.profiler.stop)()


def build_model(device):
    from maed_b200.models import MAED
    from maed_b200 import synth          # deterministic random-init weights of that architecture
    m = MAED("ste", 6, 12, MODE, DECODER, 1024)
    synth.fill_module_(m, 0)
    return m.to(device).eval()


def _cpu_state():
    from maed_b200 import synth
    from maed_b200.models import MAED
    m = MAED("ste", 6, 12, MODE, DECODER, 1024)
    synth.fill_module_(m, 0)
    return {k: v.detach() for k, v in list(m.named_parameters()) + list(m.named_buffers())}


def cpu_forward_fn():
    """(callable x -> output dict, kind, description) of the CPU arm: the reference's OWN modules (unmodified
    lib/models from /root/reference or its verbatim git-ignored copy oracle/_ref/, imported through oracle/ref_shim.py)
    when that tree is present — kind "reference" —, else the pinned oracle port (oracle/maed_oracle.py) — kind "port"."""
    from maed_b200 import synth
    try:
        from oracle import ref_shim
        if ref_shim.reference_available():
            cwd = os.getcwd()
            try:
                model = ref_shim.build_reference_model(MODE, DECODER)
            finally:
                os.chdir(cwd)
            synth.fill_module_(model, 0)
            model.eval()
            return (lambda x: model(x)), "reference", "unmodified reference lib.models.MAED (%s)" % ref_shim.REFERENCE_ROOT
    except Exception as e:                                # missing torchvision etc.: fall back to the port, say why
        sys.stderr.write("bench.py: reference import failed (%s); CPU arm uses the oracle port\n" % e)
    from oracle import maed_oracle as O
    sd = _cpu_state()
    return (lambda x: O.maed_forward(x, sd, MODE, DECODER)), "port", "oracle/maed_oracle.py (CPU restatement pinned to the reference's golden vectors)"


def best_cpu_threads(fwd):
    """PyTorch CPU ops do not scale to every core of a 100+-core host (the first run used all 128 threads and was
    15x slower than 8 threads); pick the thread count that is fastest on a 2-frame probe clip."""
    from maed_b200 import synth
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, cores) if c <= cores})
    x = synth.synth_frames(1, 2, 1)
    best, best_t = cands[0], float("inf")
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            fwd(x)
            t0 = time.perf_counter()
            fwd(x)
            dt = time.perf_counter() - t0
            if dt < best_t:
                best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cpu_clips_per_s(n_clips, reps, warmup=0, budget_s=None):
    """Times `reps` forwards of n_clips x T=16 on the host cores; returns (per-step seconds, kind, description, n_clips).
    With `budget_s` the batch per step is halved (8 -> 4 -> 2 -> 1 clips) until warmup + reps steps fit the budget, judged
    from a one-clip probe."""
    from maed_b200 import synth
    fwd, kind, desc = cpu_forward_fn()
    best_cpu_threads(fwd)
    if budget_s:
        x1 = synth.synth_frames(1, T, 2)
        with torch.no_grad():
            t0 = time.perf_counter()
            fwd(x1)
            t1 = time.perf_counter() - t0
        while n_clips > 1 and n_clips * t1 * (warmup + reps) > budget_s:
            n_clips //= 2
    x = synth.synth_frames(n_clips, T, 0)
    times = []
    with torch.no_grad():
        for i in range(warmup + reps):
            t0 = time.perf_counter()
            fwd(x)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return times, kind, desc, n_clips


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores, same metric /
    config as the CUDA arm.  Every step is the FULL batch of configs[1] (8 clips x T=16) when the run fits 150 s, else a
    bounded sample of it (stated in config.sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, kind, desc, n_clips = cpu_clips_per_s(CLIPS_PER_GPU, args.steps, args.warmup, budget_s=150.0)
    total = sum(times)
    val = n_clips * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": val,
        "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "clips_per_gpu": CLIPS_PER_GPU, "seq_len": T, "st_mode": MODE, "decoder": DECODER,
                   "sample": ("none: every step is the full batch of 8 clips x T=16" if n_clips == CLIPS_PER_GPU else
                              "%d clips x T=16 per step (bounded sample of the 8-clip batch: the full batch would exceed the "
                              "150 s budget of this arm on this host)" % n_clips),
                   "device": "host CPU (%d logical cores), torch %s, %d threads (fastest of 8/16/32/64/all on a probe clip)" % (os.cpu_count() or 1, torch.__version__, torch.get_num_threads())},
        "cpu_baseline": {"value": val, "unit": "clips/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": "%d x (%d clips, T=16) forward, %s" % (len(times), n_clips, desc)},
        "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def time_dominant_gemm(dev, reps=20):
    """Roofline of the dominant kernel (gemm_tc_kernel on the STE MLP fc1 shape: 25216 x 3072 x 768, 119 GFLOP per
    launch = 14.87 GFLOP/clip/block x 8 clips), timed alone with CUDA events on the launch stream."""
    from maed_b200 import ops
    M, N, K = CLIPS_PER_GPU * T * 197, 3072, 768
    a = ops.split(torch.randn(M, K, device=dev))
    b = ops.split(torch.randn(N, K, device=dev) * 0.02)
    bias = torch.zeros(N, device=dev)
    for _ in range(3):
        ops.gemm(a, b, bias=bias, act=ops.ACT_GELU, out_mode=ops.OUT_F16_SPLIT)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.gemm(a, b, bias=bias, act=ops.ACT_GELU, out_mode=ops.OUT_F16_SPLIT)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    return 2.0 * M * N * K / 1e12, ms


def train_measure(args, dev, dist, world, rank, local, st_mode, encoder="ste", loss_kind="mse", steps=None, sampler_cls=None):
    """BASELINE configs[2]/[3]: bs = 8 clips x T = 16 per GPU, forward + backward + Adam, random init, synthetic clips
    (reference lib/core/trainer.py:238-255: preds = model(inp); loss.backward(); optimizer.step()).  N > 1: data parallel over
    clips, the flat gradient buffer (288.5 MB fp32) is all-reduced over NCCL / NVLink every step (reference train.py:113).
    Returns a dict (rank 0) or None."""
    from maed_b200 import ops, train
    from maed_b200.models import MAED
    from maed_b200 import synth
    steps = steps or args.steps
    torch.manual_seed(0)
    cnn = encoder == "cnn"
    # 'cnn': the literal stage-1 shape (configs/config_stage1.yaml: 128 images per GPU, T = 1, torchvision ResNet-50 encoder)
    clips, Tt = (128, 1) if cnn else (CLIPS_PER_GPU, T)
    if getattr(args, "clips", None):
        clips = args.clips
    if getattr(args, "seq_len", None):
        Tt = args.seq_len
    # T = 32 (BASELINE configs[4]): the reference's temp_embed has 16 rows (vision_transformer.py:364); the extension is a
    # 32-row parameter (DESIGN.md / SURVEY.md 8d config 5)
    model = MAED("cnn" if cnn else "ste", 6, 12, st_mode, DECODER, 1024, temp_frames=max(16, Tt))
    if cnn:
        synth.fill_module_(model, 0)              # running statistics / affine parameters of a plausible BatchNorm state
    model = model.to(dev).train()
    opt = train.FusedAdam.for_model(model, lr=1e-4, weight_decay=1e-5)          # configs/config_stage2.yaml:63-66
    overlap = bool(dist) and not cnn and not getattr(args, "no_overlap", False)
    if overlap:
        train.overlap_gradient_allreduce(model)       # range-wise async all-reduce launched from inside the backward
    xs = [synth.synth_frames(clips, Tt, 300 + i).to(dev) for i in range(4)]
    target = torch.zeros(clips, Tt, 85, device=dev)
    target[..., 0] = 1.0
    h_loss = torch.empty(1).pin_memory()
    criterion, target_3d = None, None
    if loss_kind == "fused":
        # the reference's LossVideo (lib/core/loss.py:159-210) with the stage-2 weights (configs/config_stage2.yaml:34-40) on
        # synthetic targets of the trainer's shapes, through the fused CUDA loss (maed_b200/loss.py)
        from maed_b200.loss import Loss
        criterion = Loss(e_loss_weight=300., e_3d_loss_weight=600., e_pose_loss_weight=60., e_shape_loss_weight=0.06,
                         e_smpl_norm_loss=1., e_smpl_accl_loss=0., device=dev)
        g = torch.Generator().manual_seed(5)
        ones = torch.ones(clips, Tt, 49, 1)
        th = 0.2 * torch.randn(clips, Tt, 85, generator=g)
        th[..., :3] = torch.tensor([1.0, 0.0, 0.0])
        target_3d = {"kp_2d": torch.cat([2 * torch.rand(clips, Tt, 49, 2, generator=g) - 1, ones], -1).to(dev),
                     "kp_3d": torch.cat([0.3 * torch.randn(clips, Tt, 49, 3, generator=g), ones], -1).to(dev),
                     "theta": th.to(dev), "w_smpl": torch.ones(clips, Tt, device=dev)}
    ar_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def step(x):
        opt.zero_grad(set_to_none=True)
        if criterion is not None:
            loss, _ = criterion(model(x), target_3d=target_3d, target_2d=None)
        else:
            loss = ((model(x)["theta"] - target) ** 2).mean()
        loss.backward()
        if dist:
            ar_ev[0].record()
            train.allreduce_gradients(model, world)
            ar_ev[1].record()
        opt.step()
        return loss

    sampler = sampler_cls(local) if (sampler_cls and rank == 0) else None     # set up before the warm-up, outside the timed region
    for i in range(max(args.warmup, 3)):
        step(xs[i % 4])
    torch.cuda.synchronize(dev)
    if dist:
        dist.barrier()
    torch.cuda.synchronize(dev)
    if sampler:
        sampler.start()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _profile(True)
    e0.record()
    for i in range(steps):
        loss = step(xs[i % 4])
    e1.record()
    torch.cuda.synchronize(dev)
    _profile(False)
    launches = ops.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    ar_ms = torch.tensor([ar_ev[0].elapsed_time(ar_ev[1]) if dist else 0.0], device=dev)
    if dist:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ar_ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * clips * steps / (ms_total / 1000.0)
    # end to end: pinned host frames -> H2D -> train step -> D2H of the loss, every step
    hx = [synth.synth_frames(clips, Tt, 400 + i).pin_memory() for i in range(2)]
    dxb = torch.empty_like(xs[0])
    if dist:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for i in range(steps):
        dxb.copy_(hx[i % 2], non_blocking=True)
        h_loss.copy_(step(dxb).detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1000.0], device=dev)
    if dist:
        dist.barrier()
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    final_loss = float(loss.item())
    grad_bytes = model._train_state.flat_grad.numel() * 4
    del model, opt, xs, dxb
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peaks, peak_src = load_peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
    gflop_per_clip = 8.174 * Tt if cnn else GFLOP_PER_CLIP * Tt / T            # ResNet-50: 8.17 GFLOP per frame
    step_tflops = 3.0 * clips * gflop_per_clip / 1000.0 / (ms_total / steps / 1000.0)
    return {
        "metric": ("images/sec (224x224, bs=%d/gpu, T=1), MAED cnn(ResNet-50)+ktd train step (fwd+bwd+Adam)" % clips if cnn else
                   "clips/sec (T=%d, 224x224, bs=%d/gpu), MAED ste-%s+ktd train step (fwd+bwd+Adam)" % (Tt, clips, st_mode)),
        "mode": "train", "value": value, "unit": "clips/s", "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 hi/lo split operands for forward, data- and weight-gradient GEMMs; fp32 reductions / Adam",
        "data": "synthetic",
        "config": {"workload": ("configs/config_stage1.yaml shape: 128 images per GPU, encoder='cnn', train step, BatchNorm on "
                                "batch statistics (SyncBatchNorm exchange when N > 1)") if cnn else
                               ("BASELINE configs[4] shape: bs=4/gpu T=32 train step (temp_embed extended to 32 rows), random-init"
                                if Tt == 32 else
                                "BASELINE configs[2] (N=1) / configs[3] shape (N>1): bs=8/gpu T=16 train step (fwd+bwd+Adam), random-init"),
                   "clips_per_gpu": clips, "seq_len": Tt, "st_mode": st_mode, "decoder": DECODER,
                   "loss": ("reference LossVideo, stage-2 weights, fused CUDA loss (keypoint terms act on the zero body model "
                            "unless SMPL assets are loaded)") if loss_kind == "fused" else
                           "MSE on theta (the reference's parameter-space terms; keypoint terms need the SMPL tier)",
                   "parallelism": "data parallel x%d over clips, NCCL all-reduce of the flat fp32 gradient buffer every step (%s)"
                                  % (world, "overlapped with the backward" if overlap else "after the backward")},
        "clocks": clocks, "gpu_launches": int(launches), "final_loss": final_loss,
        "collective": {"kind": "all-reduce (NCCL, flat fp32 gradient buffer)", "bytes_per_step": grad_bytes,
                       "overlapped_with_backward": overlap,
                       "launches_per_step": (6 + 1 if overlap else 1),
                       "exposed_ms_last_step_max_over_ranks": float(ar_ms.item()),
                       "note": ("range-wise async all-reduces (one per STE block as its gradients become final, then embeddings + "
                                "backbone) launched by the engine's backward progress hook; exposed = what the compute stream "
                                "still waits for after the backward") if overlap else
                               "one all-reduce of the whole flat buffer after the backward (fully exposed)"} if dist else None,
        "e2e": {"value": world * clips * steps / (float(e2e_ms.item()) / 1000.0), "unit": "clips/s",
                "h2d_bytes_per_step": clips * Tt * 3 * 224 * 224 * 4, "d2h_bytes_per_step": 4},
        "roofline": {"bound": "tensor", "unit": "TFLOP/s", "peak": peak_tf, "peak_source": peak_src + ", bf16_tflops_sustained",
                     "achieved": step_tflops, "frac": step_tflops / peak_tf, "traffic": None,
                     "note": "whole step per GPU, algorithmic FLOPs = 3 x forward (SURVEY.md 8d)"},
    }


def _init(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the maed_b200 hot path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    from maed_b200 import build
    if local == 0:
        build.build()                                     # one builder per node; the other ranks wait for the .so
    if dist:
        dist.barrier()
    return world, rank, local, dev, dist


def run_train(args):
    """--mode train: the train step as the bench line (configs[2] at N=1, the configs[3] shape under torchrun)."""
    world, rank, local, dev, dist = _init(args)
    line = train_measure(args, dev, dist, world, rank, local, args.st_mode or MODE, args.encoder, args.loss,
                         sampler_cls=ClockSampler)
    if rank == 0:
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="maed_b200", choices=["maed_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="forward mode: skip the nested train-step measurement")
    ap.add_argument("--no-overlap", action="store_true", help="train step at N>1: one exposed all-reduce after the backward "
                    "instead of the overlapped range-wise exchange (A/B)")
    ap.add_argument("--mode", default="forward", choices=["forward", "train"],
                    help="forward: BASELINE configs[1] (the driver's metric, with the train step nested under \"train\"); "
                         "train: the configs[2]/[3] train step (fwd+bwd+Adam, NCCL gradient all-reduce at N>1) as the line")
    ap.add_argument("--loss", default="mse", choices=["mse", "fused"],
                    help="train mode only: 'fused' = the reference's LossVideo through maed_b200.loss")
    ap.add_argument("--encoder", default="ste", choices=["ste", "cnn"],
                    help="train mode only: 'cnn' = the stage-1 shape, 128 images per GPU")
    ap.add_argument("--st-mode", default=None, help="train mode only: parallel (default) or series")
    ap.add_argument("--clips", type=int, default=None, help="train mode only: clips (images for 'cnn') per GPU")
    ap.add_argument("--seq-len", type=int, default=None, help="train mode only: frames per clip (32 = BASELINE configs[4])")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    if args.mode == "train":
        return run_train(args)

    world, rank, local, dev, dist = _init(args)
    from maed_b200 import ops
    model = build_model(dev)
    from maed_b200 import synth
    # 4 distinct device-resident batches (308 MB > 126 MB L2), rotated; a step also streams ~4.5 GB of workspace
    xs = [synth.synth_frames(CLIPS_PER_GPU, T, 100 + i).to(dev) for i in range(4)]
    sampler = ClockSampler(local) if rank == 0 else None     # NVML (or nvidia-smi) comes up here, outside the timed region
    out = None
    for i in range(args.warmup):
        out = model(xs[i % 4])                               # same object lifetimes as the timed loop: the allocator is in steady state
    torch.cuda.synchronize(dev)

    # ---------------------------------------------------------------- device-resident timed region ("value")
    if dist:
        dist.barrier()
    torch.cuda.synchronize(dev)
    if sampler:
        sampler.start()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _profile(True)
    e0.record()
    for i in range(args.steps):
        out = model(xs[i % 4])
    e1.record()
    torch.cuda.synchronize(dev)
    _profile(False)
    launches = ops.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if dist:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * CLIPS_PER_GPU * args.steps / (ms_total / 1000.0)

    # ---------------------------------------------------------------- end-to-end through the public API ("e2e"):
    # host pinned frames -> H2D -> MAED.forward -> D2H of all five outputs the reference's evaluator pulls
    # (lib/core/evaluate.py:80-84: theta, verts, kp_2d, kp_3d, rotmat), every step, copies inside the timed region;
    # the H2D of the next two steps runs on a side stream while a step computes (what a pin_memory DataLoader with
    # prefetch_factor=2 + non_blocking copies does).
    hx = [synth.synth_frames(CLIPS_PER_GPU, T, 200 + i).pin_memory() for i in range(2)]
    NB = 3                                                   # device input / host output buffers in rotation
    dx = [torch.empty_like(xs[0]) for _ in range(NB)]
    h_out = {k: torch.empty(v.shape).pin_memory() for k, v in out.items() if k in ("theta", "verts", "kp_2d", "kp_3d", "rotmat")}
    copy_stream = torch.cuda.Stream(dev)
    main_stream = torch.cuda.current_stream(dev)
    ready = [torch.cuda.Event() for _ in range(NB)]
    consumed = [torch.cuda.Event() for _ in range(NB)]

    d2h_stream = torch.cuda.Stream(dev)
    h_outs = [h_out] + [{k: torch.empty_like(v).pin_memory() for k, v in h_out.items()} for _ in range(NB - 1)]
    done = [torch.cuda.Event() for _ in range(NB)]
    read = [torch.cuda.Event() for _ in range(NB)]

    def e2e_loop(steps):
        """A prefetching loader two steps deep: the H2D of step i+2 (copy stream) and the D2H of the five outputs of steps
        i-1 / i-2 (second copy stream) run under step i's compute; the host reads the results of step i-2 before it launches
        step i+1 and drains the last two at the end — every step's input copy and result read is inside the timed region.
        (Depth 2 instead of 1: the host then never has to launch a step in the shadow of a single running one, so launch
        jitter on a shared host does not show up as GPU idle time.)"""
        def h2d(j):
            b = j % NB
            with torch.cuda.stream(copy_stream):
                if j >= NB:
                    copy_stream.wait_event(consumed[b])
                dx[b].copy_(hx[j % 2], non_blocking=True)
                ready[b].record(copy_stream)
        for j in range(min(2, steps)):
            h2d(j)
        keep = [None] * NB
        for i in range(steps):
            cur = i % NB
            if i + 2 < steps:
                h2d(i + 2)
            main_stream.wait_event(ready[cur])
            o = model(dx[cur])
            consumed[cur].record(main_stream)
            done[cur].record(main_stream)
            keep[cur] = o                                   # the outputs stay alive until their copy has been issued and read
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done[cur])
                for k, h in h_outs[cur].items():
                    h.copy_(o[k], non_blocking=True)
                read[cur].record(d2h_stream)
            if i >= 2:
                read[(i - 2) % NB].synchronize()            # the caller reads the result of step i-2 while steps i-1, i compute
        for j in range(max(0, steps - 2), steps):
            read[j % NB].synchronize()

    e2e_loop(2)
    if dist:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    torch.cuda.synchronize(dev)
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1000.0], device=dev)
    if dist:
        dist.barrier()
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * CLIPS_PER_GPU * args.steps / (float(e2e_ms.item()) / 1000.0)
    h2d = xs[0].numel() * 4
    d2h = sum(h.numel() for h in h_out.values()) * 4
    del xs, dx, out
    model._workspace = None
    torch.cuda.empty_cache()

    # ---------------------------------------------------------------- nested: the train step (configs[2]; configs[3] shape at N>1)
    train_line = None
    if not args.no_train:
        train_line = train_measure(args, dev, dist, world, rank, local, MODE, steps=min(args.steps, 10))

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    peaks, peak_src = load_peaks()
    tf_launch, gemm_ms = time_dominant_gemm(dev)
    achieved = tf_launch / (gemm_ms / 1000.0)
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
    peak_burst = float(peaks.get("bf16_tflops", peak_tf))          # the GEMM is timed alone (20 launches): burst peak
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    step_tflops = world * CLIPS_PER_GPU * GFLOP_PER_CLIP / 1000.0 / (ms_total / args.steps / 1000.0) / world
    line = {
        "metric": METRIC,
        "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 hi/lo split operands (3 tcgen05 MMAs per K step, fp32 accumulate in TMEM); fp32 norms/softmax/tail",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "clips_per_gpu": CLIPS_PER_GPU, "seq_len": T, "st_mode": MODE, "decoder": DECODER,
                   "precision": model.precision, "parallelism": "clip-sharded x%d, no data-path collective" % world,
                   "l2": "4 rotating input batches (308 MB) and ~4.5 GB of workspace streamed per step >> 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel<256> (STE mlp.fc1 25216x3072x768, split)",
                     "achieved": achieved, "peak": peak_burst, "unit": "TFLOP/s", "frac": achieved / peak_burst,
                     "traffic": traffic, "peak_source": peak_src + ", bf16_tflops (burst: the kernel is timed alone); "
                     "whole_step_frac_of_peak uses bf16_tflops_sustained = %.1f" % peak_tf,
                     "algorithmic_tflop_per_launch": tf_launch, "ms_per_launch": gemm_ms,
                     "note": "algorithmic FLOPs (2MNK); the split-precision path issues 3x that on the tensor pipe (DESIGN.md §3)",
                     "issued_tflops": 3.0 * achieved, "issued_frac_of_sustained_peak": 3.0 * achieved / peak_tf,
                     "whole_step_algorithmic_tflops_per_gpu": step_tflops,
                     "whole_step_frac_of_peak": step_tflops / peak_tf},
    }
    if train_line is not None:
        line["train"] = train_line
    if world == 1 and not args.no_cpu_baseline:
        times, kind, desc, _ = cpu_clips_per_s(2, 2, warmup=1)
        line["cpu_baseline"] = {"value": 2 * len(times) / sum(times), "unit": "clips/s", "cores": torch.get_num_threads(),
                                "kind": kind, "sample": "2 x (2 clips, T=16) forward after 1 warm-up, %s" % desc}
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
