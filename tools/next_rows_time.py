"""Timings of the next-row kernels at BASELINE size (256 clips x 10 s): resampler, prompt mel, S3 log-mel, consumer projections."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chatterbox_embed_b200 import _lib, Resample, SpeakerProjections
from chatterbox_embed_b200 import mel as pmel, s3tokenizer as s3
dev = "cuda:0"
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
n = 256
ctx = _lib.context(0)
st = torch.cuda.current_stream().cuda_stream
# resampler 24 kHz -> 16 kHz and 44.1 kHz -> 16 kHz (device API, preallocated)
for src in (24000, 44100, 48000):
    L = src * 10
    x = 0.1 * torch.randn(n * L, device=dev)
    Lo = _lib.resample_out_len(src, 16000, L)
    y = torch.empty(n * Lo, device=dev)
    io = np.arange(n + 1, dtype=np.int64) * L; oo = np.arange(n + 1, dtype=np.int64) * Lo
    ms = timeit(lambda: ctx.resample(x.data_ptr(), io, src, 16000, y.data_ptr(), oo, st))
    print(f"resample {src} -> 16000, {n} x 10 s: {ms:.3f} ms  {(x.numel() + y.numel()) * 4 / ms / 1e6:.0f} GB/s in+out")
    del x, y
# S3 log-mel
L = 160000
x = 0.1 * torch.randn(n * L, device=dev)
off = np.arange(n + 1, dtype=np.int64) * L
out = torch.empty(n * 128 * (L // 160), device=dev)
ms = timeit(lambda: ctx.s3_log_mel(x.data_ptr(), off, out.data_ptr(), st))
rows = n * (L // 160)
print(f"s3 log-mel {n} x 10 s: {ms:.3f} ms  ({3 * 2 * rows * 400 * 400 / ms / 1e9:.0f} TF/s of 3xTF32 work)")
# consumer projections on 100k embeddings
m = SpeakerProjections().to(dev)
ve = torch.nn.functional.normalize(torch.randn(100000, 256, device=dev), dim=1); xv = torch.randn(100000, 192, device=dev)
ms1 = timeit(lambda: m.t3_speaker_cond(ve)); ms2 = timeit(lambda: m.flow_speaker_cond(xv))
print(f"projections of 1e5 embeddings: spkr_enc 256->1024 {ms1:.3f} ms ({100000 * (256 + 1024) * 4 / ms1 / 1e6:.0f} GB/s), normalize + 192->80 {ms2:.3f} ms")
