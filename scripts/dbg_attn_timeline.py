"""Debug / measurement helper for the spatial attention kernel at the bench shape (128 frames x 12 heads x 197 tokens).
  python scripts/dbg_attn_timeline.py time      -> back-to-back CUDA-event time per launch (inputs rotate over 2 x 232 MB
                                                  buffers, so every launch reads from HBM, not L2)
  MAED_B200_ATTN_DBG=1 python scripts/dbg_attn_timeline.py timeline   -> clock64 event timeline of CTA 0 (stderr)
Not part of the product path."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maed_b200._lib import call, ptr, stream_ptr

BT, ntok, heads = 128, 197, 12
rows = BT * ntok
mode = sys.argv[1] if len(sys.argv) > 1 else "time"
nbuf = 1 if mode == "timeline" else 2
qkvs = []
for i in range(nbuf):
    q = (torch.randn(2, rows, 3 * heads * 64, device="cuda") * 0.5).half()
    q[1] *= 1e-3
    qkvs.append(q)
F32 = os.environ.get("ATTN_OUT") == "f32"      # fp32 output (what the 'parallel' mode consumes) instead of fp16 planes
outs = [torch.empty(rows, heads * 64, dtype=torch.float32, device="cuda") if F32 else
        torch.empty(2, rows, heads * 64, dtype=torch.float16, device="cuda") for _ in range(nbuf)]


def launch(i):
    q, o = qkvs[i % nbuf], outs[i % nbuf]
    if F32:
        call("maed_op_attention", 0, ptr(q), C.c_longlong(q[0].numel()), BT // 16, 16, ntok, heads, C.c_float(0.125), 3, ptr(o),
             None, C.c_longlong(0), stream_ptr())
    else:
        call("maed_op_attention", 0, ptr(q), C.c_longlong(q[0].numel()), BT // 16, 16, ntok, heads, C.c_float(0.125), 3, None,
             ptr(o), C.c_longlong(o[0].numel()), stream_ptr())


for i in range(3):
    launch(i)
torch.cuda.synchronize()
if mode == "time":
    n = 40
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        launch(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    flops = 4.0 * heads * ntok * ntok * 64 * BT
    print("ATTN_TIME variant=%s us_per_launch=%.2f useful_TFLOPs=%.1f" % (
        ("f32_direct" if os.environ.get("MAED_B200_ATTN_DIRECT") else "f32_tma") if F32 else
        ("direct" if os.environ.get("MAED_B200_ATTN_DIRECT") else "tma_store"), us, flops / us * 1e-6))
