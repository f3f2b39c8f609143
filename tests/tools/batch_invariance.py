"""Does a clip's embedding depend on what else is in the batch?  (mode 1)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from chatterbox_embed_b200 import CAMPPlus, VoiceEncoder, _lib, scheduler, synth
from oracle import weights
dev = torch.device("cuda:0")
sdv, sdc = weights.ve_state_dict("W1"), weights.campplus_state_dict("W1")
ve = VoiceEncoder(); ve.load_state_dict(sdv); ve = ve.to(dev).eval()
cp = CAMPPlus(); cp.load_state_dict(sdc); cp = cp.to(dev).eval()
emb = scheduler.SpeakerEmbedder(ve, cp)
lens = [int(x) for x in synth.ragged_lengths(12)] + [16000, 25599, 160000]
wavs = [synth.clip(i, n) for i, n in enumerate(lens)]
vb, xb = emb.embed_wavs(wavs)
dv, dx = [], []
for i, w in enumerate(wavs):
    v1, x1 = emb.embed_wavs([w])
    dv.append(float(np.abs(v1[0] - vb[i]).max())); dx.append(float(np.abs(x1[0] - xb[i]).max()))
rev_v, rev_x = emb.embed_wavs(wavs[::-1])
print("VE  single vs batch: max", max(dv), " exact clips", sum(d == 0 for d in dv), "/", len(dv), " reversed batch:", float(np.abs(rev_v[::-1] - vb).max()))
print("XV  single vs batch: max", max(dx), " exact clips", sum(d == 0 for d in dx), "/", len(dx), " reversed batch:", float(np.abs(rev_x[::-1] - xb).max()), " scale", float(np.abs(xb).max()))
