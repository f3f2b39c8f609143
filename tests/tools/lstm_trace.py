"""Clock trace of one CTA of the tensor-core LSTM recurrence v2 (GPU box): where does a step go?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from chatterbox_embed_b200 import VoiceEncoder, _lib
from oracle import weights
dev = torch.device("cuda:0")
ctx = _lib.context(0)
ve = VoiceEncoder(); ve.load_state_dict(weights.ve_state_dict("W1")); ve = ve.to(dev).eval()
ctx.set_option("mode", 1)
parts = torch.rand((3072, 160, 40), device=dev) * 0.3
ve(parts); torch.cuda.synchronize()
trace = torch.zeros(160 * 2 * 8, dtype=torch.int64, device=dev)
ctx.set_option("lstm_trace", trace.data_ptr())
ve(parts); torch.cuda.synchronize()
tr = trace.cpu().numpy().reshape(160, 2, 8).astype(np.float64)
names = ["mma:full_ok", "mma:committed", "g0:top", "g0:accum_ok", "g0:math_done", "ex:ready_ok", "ex:free_ok", "ex:tma_issued"]
print("cycles relative to mma:full_ok of sub-tile A of the same step (last layer traced)")
for t in (1, 2, 50, 100, 158):
    base = tr[t, 0, 0]
    for x in (0, 1):
        print(f"  t={t} x={x}: " + "  ".join(f"{n}={tr[t, x, i] - base:7.0f}" for i, n in enumerate(names)))
print(f"cycles per step (t=50..150): {(tr[150, 0, 0] - tr[50, 0, 0]) / 100:.0f}")
d = tr[50:150]
for x in (0, 1):
    print(f"  x={x} mean: mma issue {np.mean(d[:, x, 1] - d[:, x, 0]):.0f}  commit->accum_ok {np.mean(d[:, x, 3] - d[:, x, 1]):.0f}  wait accum {np.mean(d[:, x, 3] - d[:, x, 2]):.0f}"
          f"  math {np.mean(d[:, x, 4] - d[:, x, 3]):.0f}  math_done->ready_ok {np.mean(d[:, x, 5] - d[:, x, 4]):.0f}  free wait {np.mean(d[:, x, 6] - d[:, x, 5]):.0f}"
          f"  issue {np.mean(d[:, x, 7] - d[:, x, 6]):.0f}  issued->next full_ok {np.mean(tr[51:151, x, 0] - d[:, x, 7]):.0f}")
ctx.set_option("lstm_trace", 0)
