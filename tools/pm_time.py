"""Time the prompt-mel kernel at BASELINE size (256 clips x 10 s at 24 kHz).  python tools/pm_time.py [n_clips]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chatterbox_embed_b200 import _lib, mel as pmel

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = 240000
dev = "cuda:0"
x = 0.1 * torch.randn(n_clips * n, device=dev)
ctx = _lib.context(0)
off = np.arange(n_clips + 1, dtype=np.int64) * n
rows = n_clips * _lib.prompt_mel_frames(n)
out = torch.empty(rows, 80, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    ctx.prompt_mel(x.data_ptr(), off, out.data_ptr(), st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
e0.record()
for _ in range(K):
    ctx.prompt_mel(x.data_ptr(), off, out.data_ptr(), st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
fl = 3 * 2 * rows * 1280 * 1920
print(f"prompt_mel {n_clips} clips x 10 s: {ms:.3f} ms  {fl / ms / 1e9:.1f} TF/s (3xTF32 flops)  {n_clips / ms * 1e3:.0f} clips/s  rows {rows}")
