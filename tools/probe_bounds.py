"""Where the step's time is: both encoders together, each alone, and with kernels removed from the CAMPPlus chain
(`probe` option: timing only, results wrong while set).  256 x 10 s clips, device-resident PCM, CUDA events.

    python tools/probe_bounds.py [steps]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chatterbox_embed_b200 import CAMPPlus, VoiceEncoder, _lib, scheduler, synth

K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
N, L = 256, 160000
dev = torch.device("cuda:0")
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
    return e0.elapsed_time(e1) / K


BOTH, VE, XV = _lib.DO_VE | _lib.DO_XV, _lib.DO_VE, _lib.DO_XV
DEFAULTS = {"overlap": 1, "pdl": 1, "probe": 0, "lstm_gate_warps": 4, "fcm_fuse": 1, "mode": 1, "transit_n256": 1, "lstm_late": 1, "dft_eo": 1}
rows = []
for name, flags, opts in (("both encoders (the bench step)", BOTH, {}), ("VoiceEncoder alone", VE, {}), ("CAMPPlus alone", XV, {}),
                          ("both, one stream (overlap 0)", BOTH, {"overlap": 0}), ("CAMPPlus alone, no dependent launch (pdl 0)", XV, {"pdl": 0}),
                          ("both, 8 gate warps in the recurrence (lstm_gate_warps 2)", BOTH, {"lstm_gate_warps": 2}),
                          ("VoiceEncoder alone, 8 gate warps", VE, {"lstm_gate_warps": 2}),
                          ("both, FCM identity blocks as two convolutions (fcm_fuse 0)", BOTH, {"fcm_fuse": 0}),
                          ("CAMPPlus alone, fcm_fuse 0", XV, {"fcm_fuse": 0}),
                          ("CAMPPlus alone, transit GEMMs with 128-wide tiles (transit_n256 0)", XV, {"transit_n256": 0}),
                          ("CAMPPlus alone again", XV, {}),
                          ("both, recurrence beside the FCM phase (lstm_late 0)", BOTH, {"lstm_late": 0}),
                          ("both encoders (lstm_late 1 again)", BOTH, {}),
                          ("both, lstm_late 0 again", BOTH, {"lstm_late": 0}),
                          ("both, bf16 mode (mode 2)", BOTH, {"mode": 2}),
                          ("both encoders again (drift check)", BOTH, {}),
                          ("CAMPPlus alone, CAM gate kernel removed (probe 1; dev build only)", XV, {"probe": 1})):
    try:
        for k, v in opts.items(): ctx.set_option(k, v)
    except Exception as e:
        print(f"{name:<66s} skipped ({e})"); continue
    ms = run(flags)
    for k in opts: ctx.set_option(k, DEFAULTS[k])
    rows.append((name, ms))
    print(f"{name:<66s} {ms:7.3f} ms", flush=True)
