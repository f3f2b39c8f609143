"""GPU-box check of the tensor-core front-ends against the oracle (stage level) + timing.  python tests/tools/fe_check.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from chatterbox_embed_b200 import CAMPPlus, VoiceEncoder, _lib, synth
from oracle import frontend, weights
dev = torch.device("cuda:0")
ctx = _lib.context(0)
sdv, sdc = weights.ve_state_dict("W1"), weights.campplus_state_dict("W1")
ve = VoiceEncoder(); ve.load_state_dict(sdv); ve = ve.to(dev).eval(); ve._ctx()
cp = CAMPPlus(); cp.load_state_dict(sdc); cp = cp.to(dev).eval(); cp._ctx()
lens = [16000, 48000, 50000, 25599, 37760, 720, 160000]
wavs = [synth.clip(i, n) for i, n in enumerate(lens)]
wavs[2] = synth.with_silence(2, 50000, 6000, 9000)
flat = np.concatenate(wavs); off = np.concatenate([[0], np.cumsum(lens)])
pcm = torch.from_numpy(flat).to(dev)
flags = _lib.DO_VE | _lib.DO_XV
n = len(lens)
for mode in (0, 1):
    ctx.set_option("mode", mode)
    ws = torch.zeros(ctx.workspace_bytes(lens, 77, 0.8, flags), dtype=torch.uint8, device=dev)
    veo = torch.empty(n, 256, device=dev); xvo = torch.empty(n, 192, device=dev); st = torch.zeros(n, dtype=torch.int32, device=dev)
    ctx.embed(pcm.data_ptr(), off, 20.0, 77, 0.8, veo.data_ptr(), xvo.data_ptr(), st.data_ptr(), ws.data_ptr(), ws.numel(), 0, flags)
    torch.cuda.synchronize()
    def tap(name):
        o, r, c, ld = ctx.locate(name)
        return ws[o:o + r * ld * 4].view(torch.float32).view(r, ld)[:, :c].cpu().numpy()
    dyn = ws[ctx.locate("ve_dyn")[0]:][:n * 24].view(torch.int32).view(n, 6).cpu().numpy()
    mel = tap("ve_mel"); fb = tap("xv_fbank")
    for i in range(n):
        rows = ctx.clip_rows(i)
        s, e = frontend.trim_bounds(wavs[i], 20)
        m_ref = frontend.ve_melspectrogram(wavs[i][s:e])
        k = min(len(m_ref), dyn[i][3])
        g = mel[rows["mel_row"]:rows["mel_row"] + k]
        d = np.abs(g - m_ref[:k])
        f_ref = frontend.kaldi_fbank_torchaudio(wavs[i])
        gf = fb[rows["fb_row"]:rows["fb_row"] + len(f_ref)]
        df = np.abs(gf - f_ref)
        print(f"mode {mode} clip {i} len {lens[i]}: mel max-rel {d.max() / np.abs(m_ref).max():.2e} (max {np.abs(m_ref).max():.2e}) | fbank log-domain max {df.max():.2e} mean {df.mean():.2e}", flush=True)
# timing at bench size
ctx.set_option("mode", 1)
lens = [160000] * 256
pcm = (torch.randn(256 * 160000, device=dev) * 0.1)
off = np.arange(257) * 160000
ws = torch.zeros(ctx.workspace_bytes(lens, 77, 0.8, flags), dtype=torch.uint8, device=dev)
veo = torch.empty(256, 256, device=dev); xvo = torch.empty(256, 192, device=dev); st = torch.zeros(256, dtype=torch.int32, device=dev)
for it in range(3):
    if it == 1: ctx.profile_enable(True)
    ctx.embed(pcm.data_ptr(), off, 20.0, 77, 0.8, veo.data_ptr(), xvo.data_ptr(), st.data_ptr(), ws.data_ptr(), ws.numel(), 0, flags)
torch.cuda.synchronize()
for k, v in sorted(ctx.profile_report().items(), key=lambda kv: -kv[1]["ms"]):
    if "dft" in k or "mel" in k or "fbank" in k:
        print(f"  {k:26s} {v['ms']/2:8.3f} ms/step  {v['flops']/max(v['ms'],1e-9)/1e9:8.1f} TF/s")
ctx.profile_enable(False)
