"""Run the device path twice on the same ragged batch and report, stage by stage, whether the intermediate buffers are
bit-identical.  python tests/tools/determinism.py [mode] [overlap] [pdl]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from chatterbox_embed_b200 import CAMPPlus, VoiceEncoder, _lib, synth
from oracle import weights
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
overlap = int(sys.argv[2]) if len(sys.argv) > 2 else 0
pdl = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda:0")
ctx = _lib.context(0)
for k, v in (("mode", mode), ("overlap", overlap), ("pdl", pdl)): ctx.set_option(k, v)
sdv, sdc = weights.ve_state_dict("W1"), weights.campplus_state_dict("W1")
ve = VoiceEncoder(); ve.load_state_dict(sdv); ve = ve.to(dev).eval(); ve._ctx()
cp = CAMPPlus(); cp.load_state_dict(sdc); cp = cp.to(dev).eval(); cp._ctx()
lens = [int(x) for x in synth.ragged_lengths(40)]
wavs = [synth.clip(i, n) for i, n in enumerate(lens)]
flat = np.concatenate(wavs); off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
pcm = torch.from_numpy(flat).to(dev)
n = len(lens)
flags = _lib.DO_VE | _lib.DO_XV
ws = torch.zeros(ctx.workspace_bytes(lens, 77, 0.8, flags), dtype=torch.uint8, device=dev)
veo = torch.empty(n, 256, device=dev); xvo = torch.empty(n, 192, device=dev); st = torch.zeros(n, dtype=torch.int32, device=dev)
names = ["ve_mel", "ve_partial_emb", "xv_fbank", "xv_cmn_mean", "xv_fcm_b0", "xv_fcm_b1", "xv_fcm_b2", "xv_fcm_b3", "xv_fcm_b4", "xv_fcm_b5", "xv_fcm_b6", "xv_fcm", "xv_cat1", "xv_cat2", "xv_cat3", "xv_tr3", "xv_stats"]
def tap(name):
    o, r, c, ld = ctx.locate(name)
    return ws[o:o + r * ld * 4].view(torch.float32).view(r, ld)[:, :c].clone()
runs = []
for rep in range(4):
    ctx.embed(pcm.data_ptr(), off, 20.0, 77, 0.8, veo.data_ptr(), xvo.data_ptr(), st.data_ptr(), ws.data_ptr(), ws.numel(), 0, flags)
    torch.cuda.synchronize()
    runs.append({k: tap(k) for k in names} | {"ve_out": veo.clone(), "xv_out": xvo.clone()})
print(f"mode {mode} overlap {overlap} pdl {pdl}")
for k in names + ["ve_out", "xv_out"]:
    diffs = []
    for r in range(1, 4):
        a, b = runs[0][k], runs[r][k]
        neq = (a != b) & ~(torch.isnan(a) & torch.isnan(b))
        d = (a - b).abs().nan_to_num(0).max().item()
        diffs.append((int(neq.sum().item()), d))
    first = None
    a, b = runs[0][k], runs[1][k]
    neq = ((a != b) & ~(torch.isnan(a) & torch.isnan(b)))
    if neq.any():
        idx = neq.nonzero()[0].tolist()
        cols = neq.any(0).nonzero().flatten()
        first = (idx, "cols", cols[:4].tolist(), "..", cols[-2:].tolist(), "ncols", len(cols))
    print(f"  {k:16s} shape {tuple(runs[0][k].shape)}  mismatches vs run0 (count, maxabs): {diffs}  first {first}")

# where exactly does the first nondeterministic FCM buffer differ?
for k in ("xv_fcm_b6", "xv_fcm"):
    a, b = runs[0][k], runs[1][k]
    neq = ((a != b) & ~(torch.isnan(a) & torch.isnan(b)))
    rows = neq.any(1).nonzero().flatten().cpu().numpy()
    print(k, "rows with mismatches:", len(rows), rows[:40], "row % 3:", np.bincount(rows % 3, minlength=3))
    for r in rows[:6]:
        cols = neq[r].nonzero().flatten().cpu().numpy()
        f = np.unique(cols // 32)
        print(f"   row {r}: {len(cols)} cols, freqs {f[:12]}, run0 {a[r, cols[0]].item():.4f} run1 {b[r, cols[0]].item():.4f} run2 {runs[2][k][r, cols[0]].item():.4f}")
