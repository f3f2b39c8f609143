"""BASELINE config 0: ONE 10 s clip through the drop-in calls (the reference's own usage): wall-clock latency per call."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chatterbox_embed_b200 import CAMPPlus, VoiceEncoder, SpeakerConditioner, scheduler, synth
dev = torch.device("cuda:0")
torch.manual_seed(0)
ve = VoiceEncoder().to(dev).eval(); cp = CAMPPlus().to(dev).eval()
emb = scheduler.SpeakerEmbedder(ve, cp)
w = synth.clip(0, 160000)
def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
wt = torch.from_numpy(w).to(dev)[None]
print(f"VoiceEncoder.embeds_from_wavs([10 s clip])        {t(lambda: ve.embeds_from_wavs([w], sample_rate=16000)):.2f} ms  (host array in, numpy out)")
print(f"CAMPPlus.inference(10 s clip on device)           {t(lambda: cp.inference(wt).cpu()):.2f} ms")
print(f"both encoders, one call (SpeakerEmbedder)         {t(lambda: emb.embed_wavs([w])):.2f} ms")
for n in (8, 32):
    ws = [synth.clip(i, 160000) for i in range(n)]
    print(f"both encoders, {n:2d} x 10 s clips in one call         {t(lambda: emb.embed_wavs(ws), 10):.2f} ms")

# where does a single-clip call spend its time: host enqueue (cbx_embed returns) vs GPU completion
from chatterbox_embed_b200 import _lib
ctx = _lib.context(0)
lens = [160000]; off = np.array([0, 160000], np.int64)
pcm = torch.from_numpy(w).to(dev)
flags = _lib.DO_VE | _lib.DO_XV
ws = torch.empty(ctx.workspace_bytes(lens, 77, 0.8, flags), dtype=torch.uint8, device=dev)
veo = torch.empty(1, 256, device=dev); xvo = torch.empty(1, 192, device=dev); st_ = torch.zeros(1, dtype=torch.int32, device=dev)
stream = torch.cuda.Stream()
for key, val in (("overlap", 1), ("overlap", 0)):
    ctx.set_option(key, val)
    enq, tot = [], []
    for _ in range(25):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.embed(pcm.data_ptr(), off, 20.0, 77, 0.8, veo.data_ptr(), xvo.data_ptr(), st_.data_ptr(), ws.data_ptr(), ws.numel(), stream.cuda_stream, flags)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        enq.append((t1 - t0) * 1e3); tot.append((t2 - t0) * 1e3)
    print(f"cbx_embed, one clip, {key}={val}: host enqueue {np.median(enq[5:]):.2f} ms, until done {np.median(tot[5:]):.2f} ms")
ctx.set_option("overlap", 1)
