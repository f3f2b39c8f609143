"""Stage-wise parity report (GPU box): every tap of both encoders against the golden tensors of the verbatim reference,
for the weight sets W0 / W1 / W2 in the strict-fp32 mode (0) and the tensor-core mode (1).  W2 on the golden clips is the
"W2 x pure chirp" case of VERDICT r01 (clips 1 and 3 are pure chirps): the per-stage errors show where it diverges.
    python tests/tools/stage_report.py [out.json]
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chatterbox_embed_b200 import CAMPPlus, VoiceEncoder, _lib, scheduler
from oracle import weights
import stage_taps

dev = "cuda:0"
ctx = _lib.context(0)
gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden")
res = {}
for kind in ("W0", "W1", "W2"):
    g = np.load(os.path.join(gold, f"ref_{kind}.npz"))
    ve = VoiceEncoder(); ve.load_state_dict(weights.ve_state_dict(kind)); ve = ve.to(dev).eval()
    cp = CAMPPlus(); cp.load_state_dict(weights.campplus_state_dict(kind)); cp = cp.to(dev).eval()
    emb = scheduler.SpeakerEmbedder(ve, cp)
    for mode in (0, 1):
        ctx.set_option("mode", mode)
        e = stage_taps.stage_errors(emb, g)
        res[f"{kind}_m{mode}"] = e
        print(kind, "mode", mode, " ".join(f"{k}={v:.3g}" for k, v in e.items()), flush=True)
        if kind == "W2":
            # per clip: which clips carry the x-vector error (1, 3 = pure chirps; 0, 2, 4 = noise)
            wavs = stage_taps.make_golden.golden_wavs()
            ve_o, xv_o = emb.embed_wavs(wavs)
            for i in range(len(wavs)):
                a, b = xv_o[i].astype(np.float64), g["xv_emb"][i].astype(np.float64)
                print(f"   clip {i}: max|x|={np.abs(b).max():.1f} abs err={np.abs(a - b).max():.3g} cos={a @ b / np.linalg.norm(a) / np.linalg.norm(b):.7f}")
ctx.set_option("mode", 1)
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
