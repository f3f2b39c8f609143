"""Stage-wise parity report of the CUDA path against the oracle (run on the GPU box).
    python tests/tools/gpu_check.py [W0|W1|W2] [mode]
"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from chatterbox_embed_b200 import CAMPPlus, VoiceEncoder, _lib, synth
from oracle import frontend, nets, weights

kind = sys.argv[1] if len(sys.argv) > 1 else "W1"
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
torch.manual_seed(0)
dev = torch.device("cuda:0")
ctx = _lib.context(0)
ctx.set_option("mode", mode)
sdv, sdc = weights.ve_state_dict(kind), weights.campplus_state_dict(kind)
ve = VoiceEncoder(); ve.load_state_dict(sdv); ve = ve.to(dev).eval()
cp = CAMPPlus(); cp.load_state_dict(sdc); cp = cp.to(dev).eval(); cp._ctx(); ve._ctx()

lens = [16000, 48000, 50000, 25599, 37760, 720]
wavs = [synth.clip(i, n) for i, n in enumerate(lens)]
wavs[2] = synth.with_silence(2, 50000, 6000, 9000)

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max()), float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30)), \
        float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))

# ---- VE ----
t = time.time(); got = ve.embeds_from_wavs(wavs[:5], 16000); torch.cuda.synchronize(); print("VE time", time.time() - t)
want = nets.ve_embed_wavs(sdv, wavs[:5])
for i in range(5):
    print("VE clip", i, lens[i], "maxabs/rel/cos", rel(got[i], want[i]))
# stage taps through the device path
flat = np.concatenate(wavs[:5]); off = np.concatenate([[0], np.cumsum(lens[:5])])
pcm = torch.from_numpy(flat).to(dev)
flags = _lib.DO_VE | _lib.DO_XV
ws = torch.empty(ctx.workspace_bytes(lens[:5], 77, 0.8, flags), dtype=torch.uint8, device=dev)
veo = torch.empty(5, 256, device=dev); xvo = torch.empty(5, 192, device=dev); st = torch.zeros(5, dtype=torch.int32, device=dev)
ctx.embed(pcm.data_ptr(), off, 20.0, 77, 0.8, veo.data_ptr(), xvo.data_ptr(), st.data_ptr(), ws.data_ptr(), ws.numel(), 0, flags)
torch.cuda.synchronize()
print("status", st.cpu().numpy(), "launches", ctx.launch_count())

def tap(name):
    o, r, c, ld = ctx.locate(name)
    return ws[o:o + r * ld * 4].view(torch.float32).view(r, ld)[:, :c].cpu().numpy()

dyn = ws[ctx.locate("ve_dyn")[0]:][:5 * 24].view(torch.int32).view(5, 6).cpu().numpy()
print("trim dyn\n", dyn, "\noracle", [frontend.trim_bounds(w, 20) for w in wavs[:5]])
mel = tap("ve_mel")
for i in range(5):
    rows = ctx.clip_rows(i)
    s, e = frontend.trim_bounds(wavs[i], 20)
    m_ref = frontend.ve_melspectrogram(wavs[i][s:e])
    n = min(len(m_ref), dyn[i][3])
    print("mel clip", i, m_ref.shape, rel(mel[rows["mel_row"]:rows["mel_row"] + n], m_ref[:n]))
# ---- XV ----
fb = tap("xv_fbank"); mean = tap("xv_cmn_mean")
for i in range(5):
    rows = ctx.clip_rows(i)
    f_ref = frontend.kaldi_fbank_torchaudio(wavs[i])
    g = fb[rows["fb_row"]:rows["fb_row"] + len(f_ref)]
    print("fbank clip", i, f_ref.shape, rel(g, f_ref), "mean", rel(mean[i], f_ref.mean(0)))
taps = {}
xv_ref = nets.campplus_embed_wavs(sdc, [wavs[1]], taps)
rows = ctx.clip_rows(1)
T = taps["fcm"].shape[-1]; Tp = taps["tdnn"].shape[-1]
fcm = tap("xv_fcm")[rows["fb_row"]:rows["fb_row"] + T]                # [T][f*32+c]
fcm_ref = taps["fcm"][0].numpy().reshape(32, 10, T).transpose(2, 1, 0).reshape(T, 320)
print("fcm", rel(fcm, fcm_ref))
for name, key, C in (("xv_cat1", "block1", 512), ("xv_cat2", "block2", 1024), ("xv_cat3", "block3", 1024)):
    g = tap(name)[rows["td_row"]:rows["td_row"] + Tp]
    r_ = taps[key][0].numpy().T
    print(name, "first128", rel(g[:, :128], r_[:, :128]), "all", rel(g, r_))
print("tr3", rel(tap("xv_tr3")[rows["td_row"]:rows["td_row"] + Tp], taps["transit3"][0].numpy().T))
print("stats", rel(tap("xv_stats")[1], taps["stats"][0].numpy()))
xv_all = nets.campplus_embed_wavs(sdc, wavs[:5])
got_x = xvo.cpu().numpy()
for i in range(5):
    print("XV clip", i, rel(got_x[i], xv_all[i]))
print("VE(dev path)", rel(veo.cpu().numpy(), want))
# guard rows must be zero
g = tap("xv_cat3"); print("guard rows zero:", float(np.abs(g[:2]).max()), float(np.abs(g[rows['td_row'] + Tp: rows['td_row'] + Tp + 2]).max()))
# CAMPPlus.inference through the class + 720-sample clip
out = cp.inference([torch.from_numpy(w) for w in wavs]).cpu().numpy()
print("class inference", rel(out[:5], xv_all), "720-sample finite:", np.isfinite(out[5]).all(), rel(out[5], nets.campplus_embed_wavs(sdc, [wavs[5]])[0]))
