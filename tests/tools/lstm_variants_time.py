"""A/B timing of the recurrence kernel's variants (GPU box): gate warps per quadrant 2 / 4, fp32 / bf16 xw, on the bench's own
shape (3072 partials = 14 clusters) -- interleaved repeats so that clock drift hits every variant alike.
    python tests/tools/lstm_variants_time.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from chatterbox_embed_b200 import VoiceEncoder, _lib
from oracle import weights
dev = torch.device("cuda:0")
ctx = _lib.context(0)
ve = VoiceEncoder(); ve.load_state_dict(weights.ve_state_dict("W1")); ve = ve.to(dev).eval()
ctx.set_option("mode", 1)
n = 3072
parts = torch.rand((n, 160, 40), device=dev) * 0.3
variants = [(2, 0), (4, 0), (2, 1), (4, 1)]
res = {v: [] for v in variants}
for rep in range(6):
    for gw, x16 in variants:
        ctx.set_option("lstm_gate_warps", gw); ctx.set_option("xw_bf16", x16)
        ve(parts); torch.cuda.synchronize()
        ctx.profile_enable(True)
        for _ in range(3): ve(parts)
        torch.cuda.synchronize()
        r = ctx.profile_report()
        ctx.profile_enable(False)
        res[(gw, x16)].append((r["lstm_rec_tc_kernel"]["ms"] / 3, r["lstm_xw_gemm"]["ms"] / 3))
ctx.set_option("lstm_gate_warps", 4); ctx.set_option("xw_bf16", 0)
for v in variants:
    a = np.array(res[v])
    print(f"gate warps/quadrant {v[0]}, xw bf16 {v[1]}: recurrence (3 layers) min {a[:,0].min():.3f} median {np.median(a[:,0]):.3f} ms; xw GEMMs (2) min {a[:,1].min():.3f} ms")
