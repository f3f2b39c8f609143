"""CTA-pair GEMM (gemm_pair=1) against the single-CTA kernels: same embeddings on a ragged batch, then timing.
    python tools/pair_check.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chatterbox_embed_b200 import CAMPPlus, VoiceEncoder, _lib, scheduler, synth
from oracle import weights

dev = torch.device("cuda:0")
ctx = _lib.context(0)
sdv, sdc = weights.ve_state_dict("W1"), weights.campplus_state_dict("W1")
ve = VoiceEncoder(); ve.load_state_dict(sdv); ve = ve.to(dev).eval()
cp = CAMPPlus(); cp.load_state_dict(sdc); cp = cp.to(dev).eval()
emb = scheduler.SpeakerEmbedder(ve, cp)
lens = [16000, 48000, 50000, 25599, 37760, 720, 160000, 99999, 64000]
wavs = [synth.clip(i, n) for i, n in enumerate(lens)]
out = {}
for pair in (0, 1, 0, 1):
    ctx.set_option("gemm_pair", pair)
    ve_o, xv_o = emb.embed_wavs(wavs)
    out.setdefault(pair, []).append(xv_o.copy())
    print("pair", pair, "xv[0,:4]", xv_o[0, :4], flush=True)
d = np.abs(out[0][0] - out[1][0]).max()
print("max |pair - single| =", d, " run-to-run single", np.abs(out[0][0] - out[0][1]).max(), " pair", np.abs(out[1][0] - out[1][1]).max())
assert d < 1e-5 * max(1.0, np.abs(out[0][0]).max()), "CTA-pair GEMM differs"
print("OK")
