"""Stage buffers with programmatic dependent launch on vs off on a SMALL batch (tiny kernels overlap deepest)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from chatterbox_embed_b200 import CAMPPlus, VoiceEncoder, _lib, synth
from oracle import weights, make_golden
dev = torch.device("cuda:0")
ctx = _lib.context(0)
sdv, sdc = weights.ve_state_dict("W1"), weights.campplus_state_dict("W1")
ve = VoiceEncoder(); ve.load_state_dict(sdv); ve = ve.to(dev).eval(); ve._ctx()
cp = CAMPPlus(); cp.load_state_dict(sdc); cp = cp.to(dev).eval(); cp._ctx()
wavs = make_golden.golden_wavs()
lens = [len(w) for w in wavs]
flat = np.concatenate(wavs); off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
pcm = torch.from_numpy(flat).to(dev)
n = len(lens)
flags = _lib.DO_VE | _lib.DO_XV
ws = torch.zeros(ctx.workspace_bytes(lens, 77, 0.8, flags), dtype=torch.uint8, device=dev)
veo = torch.empty(n, 256, device=dev); xvo = torch.empty(n, 192, device=dev); st = torch.zeros(n, dtype=torch.int32, device=dev)
names = ["xv_fcm", "xv_cat1", "xv_cat2", "xv_cat3", "xv_tr3", "xv_stats"]
def tap(name):
    o, r, c, ld = ctx.locate(name)
    return ws[o:o + r * ld * 4].view(torch.float32).view(r, ld)[:, :c].clone()
res = {}
stream = torch.cuda.Stream()           # a non-blocking stream: the legacy default stream serialises everything
ovl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for pdl in (0, 1, 1, 1):
    ctx.set_option("pdl", pdl); ctx.set_option("overlap", ovl)
    torch.cuda.synchronize()
    ctx.embed(pcm.data_ptr(), off, 20.0, 77, 0.8, veo.data_ptr(), xvo.data_ptr(), st.data_ptr(), ws.data_ptr(), ws.numel(), stream.cuda_stream, flags)
    torch.cuda.synchronize()
    res.setdefault(pdl, []).append({k: tap(k) for k in names} | {"xv_out": xvo.clone()})
for k in names + ["xv_out"]:
    a = res[0][0][k]
    for j, r in enumerate(res[1]):
        b = r[k]
        neq = (a != b) & ~(torch.isnan(a) & torch.isnan(b))
        cols = neq.any(0).nonzero().flatten().tolist()
        rows = neq.any(1).nonzero().flatten().tolist()
        print(f"{k:8s} pdl run {j}: mismatches {int(neq.sum())} maxabs {(a - b).abs().nan_to_num(0).max().item():.3e} cols {cols[:6]}..{cols[-2:]} rows {rows[:6]}..{rows[-2:]}")
