"""Option cat_bf16 (the D-TDNN GEMMs read a bf16 copy of the concatenation buffers): accuracy against the fp32 oracle for the
three weight sets, and what it buys (256 x 10 s clips, CAMPPlus alone and both encoders).

    python tests/tools/cat_bf16_check.py [settings, default: 0 1 2]

cat_bf16 = 1: bf16 copy of the concatenation buffers, tf32 operands; 2: bf16 operands as well (kind::f16, bf16 weights).
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from chatterbox_embed_b200 import CAMPPlus, VoiceEncoder, _lib, scheduler, synth
from oracle import nets, weights

dev = torch.device("cuda:0")
OPTS = [int(a) for a in sys.argv[1:]] or [0, 1, 2]
wavs = [synth.mixed(i, n) for i, n in enumerate((48000, 25600, 64000, 16000 * 7 + 123))]
for kind in ("W0", "W1", "W2"):
    sdc = weights.campplus_state_dict(kind)
    want = nets.campplus_embed_wavs(sdc, wavs)
    cp = CAMPPlus(); cp.load_state_dict(sdc); cp = cp.to(dev).eval()
    ctx = cp._ctx()
    for opt in OPTS:
        ctx.set_option("cat_bf16", opt)
        got = cp.inference([torch.from_numpy(w) for w in wavs]).cpu().numpy()
        err = float(np.abs(got - want).max())
        cos = min(float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b))) for a, b in zip(got, want))
        print(f"{kind} cat_bf16={opt}: max-abs {err:.2e} (max |x| {np.abs(want).max():.2f}), min cos {cos:.6f}", flush=True)
    ctx.set_option("cat_bf16", 0)

N, L, K = 256, 160000, 10
torch.manual_seed(0)
ve = VoiceEncoder().to(dev).eval(); cp = CAMPPlus().to(dev).eval()
emb = scheduler.SpeakerEmbedder(ve, cp)
ctx = emb.ctx()
off = np.arange(N + 1, dtype=np.int64) * L
pcm = torch.from_numpy(np.concatenate([synth.clip(i, L) for i in range(N)])).to(dev)
ve_o = torch.empty((N, 256), device=dev); xv_o = torch.empty((N, 192), device=dev); status = torch.empty(N, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream


def run(flags):
    ws = emb._ws.get(ctx.workspace_bytes(np.diff(off), 77, 0.8, flags), dev)
    def once():
        ctx.embed(pcm.data_ptr(), off, 20.0, 77, 0.8, ve_o.data_ptr(), xv_o.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), stream, flags)
    for _ in range(3): once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): once()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K, xv_o.clone()


ref = None
for opt in OPTS:
    ctx.set_option("cat_bf16", opt)
    ms_x, xv = run(_lib.DO_XV)
    ms_b, _ = run(_lib.DO_VE | _lib.DO_XV)
    if ref is None: ref = xv
    print(f"cat_bf16={opt}: CAMPPlus alone {ms_x:.3f} ms, both encoders {ms_b:.3f} ms per step = {N / ms_b * 1e3:.0f} clips/s; "
          f"x-vector vs cat_bf16=0 max-abs {float((xv - ref).abs().max()):.2e}, finite {bool(torch.isfinite(xv).all())}", flush=True)
ctx.set_option("cat_bf16", 0)
