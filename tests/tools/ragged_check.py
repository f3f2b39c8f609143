"""Repeatability of the mode-1 x-vector error on the ragged config-3 sample (run-to-run, per build).  python tests/tools/ragged_check.py [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from chatterbox_embed_b200 import CAMPPlus, VoiceEncoder, _lib, scheduler, synth
from oracle import nets, weights
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
ctx = _lib.context(0)
sdv, sdc = weights.ve_state_dict("W1"), weights.campplus_state_dict("W1")
ve = VoiceEncoder(); ve.load_state_dict(sdv); ve = ve.to(dev).eval()
cp = CAMPPlus(); cp.load_state_dict(sdc); cp = cp.to(dev).eval()
emb = scheduler.SpeakerEmbedder(ve, cp)
lens = [int(x) for x in synth.ragged_lengths(40)]
wavs = [synth.clip(i, n) for i, n in enumerate(lens)]
pick = sorted({int(np.argmin(lens)), int(np.argmax(lens)), 7, 23})
want = nets.campplus_embed_wavs(sdc, [wavs[i] for i in pick])
scale = max(1.0, float(np.abs(want).max()))
ctx.set_option("mode", 0)
_, xv0 = emb.embed_wavs(wavs)
print("strict fp32 mode: max err", np.abs(xv0[pick] - want).max(), "scale", scale)
ctx.set_option("mode", 1)
for opt in ([("pdl", 1)], [("pdl", 0)], [("pdl", 1), ("overlap", 0)]):
    for k, v in opt: ctx.set_option(k, v)
    errs = []
    for r in range(reps):
        _, xv = emb.embed_wavs(wavs)
        errs.append(float(np.abs(xv[pick] - want).max()))
    print(opt, "mode 1 max err per run:", ["%.2e" % e for e in errs], "vs strict run", "%.2e" % float(np.abs(xv[pick] - xv0[pick]).max()))
    ctx.set_option("pdl", 1); ctx.set_option("overlap", 1)
