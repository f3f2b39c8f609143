"""Stress: many batches of varied size through the HOST path (two streams, programmatic dependent launch) must give the same
bits as the fully serialised configuration (pdl = 0, overlap = 0).  python tests/tools/pdl_stress.py [rounds]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from chatterbox_embed_b200 import CAMPPlus, VoiceEncoder, _lib, scheduler, synth
from oracle import weights
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 24
dev = torch.device("cuda:0")
ctx = _lib.context(0)
sdv, sdc = weights.ve_state_dict("W1"), weights.campplus_state_dict("W1")
ve = VoiceEncoder(); ve.load_state_dict(sdv); ve = ve.to(dev).eval()
cp = CAMPPlus(); cp.load_state_dict(sdc); cp = cp.to(dev).eval()
emb = scheduler.SpeakerEmbedder(ve, cp)
rng = np.random.RandomState(7)
bad = 0
for r in range(rounds):
    n = int(rng.choice([1, 2, 3, 5, 8, 13, 40, 96]))
    lens = [int(16000 * rng.uniform(0.5, 12.0)) for _ in range(n)]
    wavs = [synth.clip(100 * r + i, L) for i, L in enumerate(lens)]
    ctx.set_option("pdl", 0); ctx.set_option("overlap", 0)
    ref = emb.embed_wavs(wavs)
    ctx.set_option("pdl", 1); ctx.set_option("overlap", 1)
    for rep in range(3):
        got = emb.embed_wavs(wavs)
        ok = np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
        if not ok:
            bad += 1
            print(f"round {r} rep {rep}: n={n} MISMATCH ve {np.abs(got[0]-ref[0]).max():.3e} xv {np.abs(got[1]-ref[1]).max():.3e}", flush=True)
print(f"{rounds} rounds x 3 repetitions: {bad} mismatches")
