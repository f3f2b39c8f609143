"""GPU-box check of the tensor-core LSTM recurrence against the oracle, with timing.  python tests/tools/lstm_check.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from chatterbox_embed_b200 import VoiceEncoder, _lib
from oracle import nets, weights
dev = torch.device("cuda:0")
ctx = _lib.context(0)
for kind in ("W1", "W2"):
    sdv = weights.ve_state_dict(kind)
    ve = VoiceEncoder(); ve.load_state_dict(sdv); ve = ve.to(dev).eval()
    for n in (3, 200):
        g = torch.Generator().manual_seed(n)
        parts = (torch.rand((n, 160, 40), generator=g) * 0.3)
        with torch.inference_mode():
            want = nets.ve_forward(sdv, parts.numpy()).numpy()
        for mode in (0, 1):
            ctx.set_option("mode", mode)
            got = ve(parts.to(dev)).cpu().numpy()
            d = np.abs(got - want)
            cs = min(float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b))) for a, b in zip(got, want))
            print(f"{kind} n={n} mode={mode}: max-abs {d.max():.3e} mean {d.mean():.3e} min-cos {cs:.8f} nan={np.isnan(got).sum()}", flush=True)
# timing: 3072 partials
ctx.set_option("mode", 1)
parts = torch.rand((3072, 160, 40), device=dev) * 0.3
for _ in range(2): ve(parts)
ctx.profile_enable(True)
for _ in range(3): ve(parts)
torch.cuda.synchronize()
for k, v in sorted(ctx.profile_report().items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:24s} {v['ms']/3:8.3f} ms/iter  n={v['launches']/3:.0f}  {v['flops']/max(v['ms'],1e-9)/1e9:8.1f} TF/s")
ctx.profile_enable(False)
