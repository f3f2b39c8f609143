"""Stage-wise comparison of the CUDA path's taps (cbx_locate) with the golden tensors the VERBATIM reference modules produced
(tests/golden/ref_W*.npz, oracle/make_golden.py).  Shared by tests/test_gpu_parity.py::test_golden_stages (both modes) and the
report tool tests/tools/stage_report.py.

The embedding gate of the north_star has almost no power with default-init weights (SURVEY.md 8d hazard 1), so every stage
of both encoders is compared on its own: trim indices, mel, the last hidden state of each LSTM layer, partial embeddings,
fbank, FCM head, TDNN, the three dense blocks (whose leading channels are the transit outputs), transit 3, the pooled
statistics."""
import numpy as np
import torch

from oracle import make_golden

DEV = "cuda:0"
CLIP = 1                      # the 3 s clip of make_golden.CLIPS whose stage tensors are in the fixture


def relerr(a, b):
    """max |a - b| relative to max |b| (a per-stage scale: the stages span 1e-2 .. 1e2)."""
    return float(np.abs(np.asarray(a, np.float64) - b).max() / (np.abs(b).max() + 1e-30))


def stage_errors(emb, g):
    """Run the golden clips through ``emb`` (scheduler.SpeakerEmbedder) in the library's current mode and return
    {stage: error} against the fixture ``g``; 'trim' is the number of mismatching trim indices (must be 0)."""
    wavs = make_golden.golden_wavs()
    ctx = emb.ctx()
    flat = np.concatenate(wavs)
    off = np.concatenate([[0], np.cumsum([len(w) for w in wavs])]).astype(np.int64)
    pcm = torch.from_numpy(flat).to(DEV)
    ve_o, xv_o, _ = emb.embed_device(pcm, off)
    torch.cuda.synchronize()
    ws = emb._ws.buf

    def tap(name):
        o, r, c, ld = ctx.locate(name)
        return ws[o:o + r * ld * 4].view(torch.float32).view(r, ld)[:, :c].cpu().numpy()

    n = len(wavs)
    out = {}
    dyn = ws[ctx.locate("ve_dyn")[0]:][:n * 24].view(torch.int32).view(n, 6).cpu().numpy()
    out["trim"] = int((dyn[:, :2] != g["trim"]).sum())
    rows = ctx.clip_rows(CLIP)
    n_mel = g["mel_1"].shape[0]
    out["mel"] = relerr(tap("ve_mel")[rows["mel_row"]:rows["mel_row"] + n_mel], g["mel_1"])
    n_p = g["partial_emb_1"].shape[0]
    slots = ctx.locate("ve_partial_emb")[1]
    hl = tap("ve_hlast").reshape(3, slots, 256)[:, rows["slot"]:rows["slot"] + n_p]
    for l in range(3):
        out[f"lstm_h{l}"] = float(np.abs(hl[l] - g["lstm_h_1"][l]).max())          # |h| <= 1: absolute
    out["partial_emb"] = float(np.abs(tap("ve_partial_emb")[rows["slot"]:rows["slot"] + n_p] - g["partial_emb_1"]).max())
    t_fb = g["fbank_cmn_1"].shape[0]
    fb = tap("xv_fbank")[rows["fb_row"]:rows["fb_row"] + t_fb] - tap("xv_cmn_mean")[CLIP]
    d = np.abs(fb - g["fbank_cmn_1"])
    out["fbank_mean"], out["fbank_max"] = float(d.mean()), float(d.max())      # log domain; floor bins of a chirp
    fcm = tap("xv_fcm")[rows["fb_row"]:rows["fb_row"] + t_fb]                   # [t][f*32+c] -> reference channel c*10+f
    fcm = fcm.reshape(t_fb, 10, 32).transpose(2, 1, 0).reshape(320, t_fb)[:, ::16]
    out["fcm"] = relerr(fcm, g["fcm_1"])
    t_td = (t_fb - 1) // 2 + 1
    cat1 = tap("xv_cat1")[rows["td_row"]:rows["td_row"] + t_td].T
    out["tdnn"] = relerr(cat1[:128, ::8], g["tdnn_1"])
    out["block1"] = relerr(cat1[:, ::16], g["block1_1"])
    out["block2"] = relerr(tap("xv_cat2")[rows["td_row"]:rows["td_row"] + t_td].T[:, ::16], g["block2_1"])
    out["block3"] = relerr(tap("xv_cat3")[rows["td_row"]:rows["td_row"] + t_td].T[:, ::16], g["block3_1"])
    out["transit3"] = relerr(tap("xv_tr3")[rows["td_row"]:rows["td_row"] + t_td].T[:, ::16], g["transit3_1"])
    out["stats"] = relerr(tap("xv_stats")[CLIP], g["stats_1"])
    ve = ve_o.cpu().numpy(); xv = xv_o.cpu().numpy()
    out["ve_emb"] = float(np.abs(ve - g["ve_emb"]).max())
    out["xv_emb_abs"] = float(np.abs(xv - g["xv_emb"]).max())
    out["xv_emb"] = relerr(xv, g["xv_emb"])
    cosv = lambda a, b: float(np.dot(a.astype(np.float64), b) / (np.linalg.norm(a.astype(np.float64)) * np.linalg.norm(b.astype(np.float64))))
    out["ve_min_cos"] = min(cosv(a, b) for a, b in zip(ve, g["ve_emb"]))
    out["xv_min_cos"] = min(cosv(a, b) for a, b in zip(xv, g["xv_emb"]))
    return out
