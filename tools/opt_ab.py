"""In-process A/B of library options on the bench step (256 x 10 s clips, device-resident PCM, CUDA events): the arms are run in turn,
`reps` times, so that the clock drift of the box hits every arm alike.

    python tools/opt_ab.py bn_prefetch=0 bn_prefetch=296 [--what both,xv,ve] [--reps 5] [--steps 10]

An arm is `key=value[,key=value...]`; after each run the options are put back to the values they had at start."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from chatterbox_embed_b200 import CAMPPlus, VoiceEncoder, _lib, scheduler, synth

argv = sys.argv[1:]
arms, what, reps, K = [], ["both", "xv"], 5, 10
i = 0
while i < len(argv):
    if argv[i] == "--what":
        what = argv[i + 1].split(","); i += 2
    elif argv[i] == "--reps":
        reps = int(argv[i + 1]); i += 2
    elif argv[i] == "--steps":
        K = int(argv[i + 1]); i += 2
    else:
        arms.append(argv[i]); i += 1
N, L = 256, 160000
dev = torch.device("cuda:0")
torch.manual_seed(0)
ve = VoiceEncoder().to(dev).eval(); cp = CAMPPlus().to(dev).eval()
emb = scheduler.SpeakerEmbedder(ve, cp)
ctx = emb.ctx()
ctx.set_option("mode", 1)
off = np.arange(N + 1, dtype=np.int64) * L
pcm = torch.from_numpy(np.concatenate([synth.clip(i, L) for i in range(N)])).to(dev)
ve_o = torch.empty((N, 256), device=dev); xv_o = torch.empty((N, 192), device=dev); status = torch.empty(N, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream
FLAGS = {"both": _lib.DO_VE | _lib.DO_XV, "xv": _lib.DO_XV, "ve": _lib.DO_VE}


def run(flags):
    ws = emb._ws.get(ctx.workspace_bytes(np.diff(off), 77, 0.8, flags), dev)
    def once():
        ctx.embed(pcm.data_ptr(), off, 20.0, 77, 0.8, ve_o.data_ptr(), xv_o.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), stream, flags)
    for _ in range(2): once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): once()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


res = {(a, w): [] for a in arms for w in what}
outs = {}
for r in range(reps):
    for a in arms:
        kv = [x.split("=") for x in a.split(",")]
        old = {k: int(ctx.get_option(k)) for k, _ in kv}
        for k, v in kv: ctx.set_option(k, int(v))
        for w in what:
            res[(a, w)].append(run(FLAGS[w]))
        if "both" in what:
            outs[a] = (ve_o.cpu().numpy().copy(), xv_o.cpu().numpy().copy())
        for k, v in old.items(): ctx.set_option(k, v)
for w in what:
    for a in arms:
        x = np.array(res[(a, w)])
        print(f"{w:5s} {a:40s} min {x.min():7.3f}  median {np.median(x):7.3f} ms   " + " ".join(f"{v:.3f}" for v in x))
if len(outs) > 1:
    base = outs[arms[0]]
    for a in arms[1:]:
        print(f"results {a} vs {arms[0]}: VE max-abs diff {np.abs(outs[a][0] - base[0]).max():.3e}, x-vector {np.abs(outs[a][1] - base[1]).max():.3e}")
