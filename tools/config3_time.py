"""BASELINE config 3: ragged batch of 1024 clips, 3-30 s each, both encoders, one call (device-resident PCM)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chatterbox_embed_b200 import CAMPPlus, VoiceEncoder, _lib, scheduler, synth
dev = torch.device("cuda:0")
torch.manual_seed(0)
ve = VoiceEncoder().to(dev).eval(); cp = CAMPPlus().to(dev).eval()
emb = scheduler.SpeakerEmbedder(ve, cp)
lens = [int(x) for x in synth.ragged_lengths(1024)]
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
pcm = 0.1 * torch.randn(int(off[-1]), device=dev)
secs = off[-1] / 16000.0
for _ in range(2): out = emb.embed_device(pcm, off)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 5
for _ in range(K): ve_o, xv_o, st = emb.embed_device(pcm, off)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"config 3: 1024 clips, {secs:.0f} s of audio ({min(lens)/16000:.1f}-{max(lens)/16000:.1f} s): {ms:.1f} ms per call = {1024/ms*1e3:.0f} clips/s, {secs/ms*1e3:.0f} audio-s/s; "
      f"status bits {int(st.max())}, finite {bool(torch.isfinite(ve_o).all() and torch.isfinite(xv_o).all())}, peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
