"""Timing experiments on the tensor-core LSTM recurrence (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from chatterbox_embed_b200 import VoiceEncoder, _lib
from oracle import weights
dev = torch.device("cuda:0")
ctx = _lib.context(0)
print("max active clusters:", _lib.lib().cbx_lstm_max_clusters(ctx._h))
ve = VoiceEncoder(); ve.load_state_dict(weights.ve_state_dict("W1")); ve = ve.to(dev).eval()
ctx.set_option("mode", 1)
for n in [192, 192 * 4, 192 * 8, 192 * 9, 192 * 12, 192 * 16, 192 * 17, 192 * 32]:
    parts = torch.rand((n, 160, 40), device=dev) * 0.3
    for _ in range(2): ve(parts)
    ctx.profile_enable(True)
    for _ in range(3): ve(parts)
    torch.cuda.synchronize()
    r = ctx.profile_report()["lstm_rec_tc_kernel"]
    print(f"n={n:5d} ({n//192:2d} clusters): lstm_rec_tc_kernel {r['ms']/r['launches']:.3f} ms/launch", flush=True)
    ctx.profile_enable(False)
