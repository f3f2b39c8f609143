"""Oracle (TEST INFRASTRUCTURE): generate tests/golden/*.npz by running the
VERBATIM reference modules from /root/reference (build container only).

    python -m oracle.make_golden

Inputs are synthetic (chatterbox_embed_b200.synth) and weights are the seeded
sets of oracle/weights.py, loaded into the reference classes with
``load_state_dict(strict=True)``; so the GPU box can regenerate the same inputs
and weights and compare the CUDA path against what the reference itself produced.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from chatterbox_embed_b200 import synth
from . import refload, weights

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (clip index, samples, leading quiet samples, trailing quiet samples)
CLIPS = [(0, 16000, 0, 0), (1, 48000, 0, 0), (2, 50000, 6000, 9000), (3, 160000, 0, 0), (4, 25599, 0, 0)]


def golden_wavs():
    return [synth.with_silence(i, n, a, b) if (a or b) else synth.clip(i, n) for i, n, a, b in CLIPS]


def prompt_mel_wavs():
    """24 kHz inputs of the prompt-mel fixture: synthetic voice, white noise (every bin above the clamp), a chirp with
    near-empty bins (exercises the clamp and the 3xTF32 split), the shortest legal clip and a length that is not a hop multiple."""
    rng = np.random.RandomState(24)
    t = np.arange(36000) / 24000.0
    chirp = (0.5 * np.sin(2 * np.pi * (200.0 * t + 0.5 * 6000.0 * t * t / t[-1]))).astype(np.float32)
    return [synth.clip(7, 72000), (0.2 * rng.randn(30001)).astype(np.float32), chirp, synth.clip(8, 721), synth.clip(9, 4799)]


def main_prompt_mel():
    """tests/golden/ref_prompt_mel.npz from the verbatim s3gen/utils/mel.py (torch fp32 stft)."""
    assert refload.available(), "reference tree not found"
    mel = refload.mel_module()
    out = {}
    for i, w in enumerate(prompt_mel_wavs()):
        out[f"mel_{i}"] = mel.mel_spectrogram(torch.from_numpy(w))[0].T.contiguous().numpy()     # (T, 80), as embed_ref hands it on
    batch = np.stack([prompt_mel_wavs()[0][:24000], prompt_mel_wavs()[1][:24000]])
    out["mel_batch"] = mel.mel_spectrogram(batch).numpy()                                        # (2, 80, 50): the numpy / batch branch
    np.savez_compressed(os.path.join(OUT, "ref_prompt_mel.npz"), **out)
    print("prompt_mel", {k: v.shape for k, v in out.items()})


def s3_wavs():
    """16 kHz inputs of the S3Tokenizer log-mel fixture: voice-like, white noise, a chirp (bins 8 decades below the peak: the
    max - 8 floor), the shortest legal clip, a length that is not a hop multiple."""
    rng = np.random.RandomState(16)
    t = np.arange(40000) / 16000.0
    chirp = (0.5 * np.sin(2 * np.pi * (100.0 * t + 0.5 * 7000.0 * t * t / t[-1]))).astype(np.float32)
    return [synth.clip(11, 48000), (0.1 * rng.randn(16001)).astype(np.float32), chirp, synth.clip(12, 201), synth.clip(13, 25599)]


def main_s3():
    """tests/golden/ref_s3_log_mel.npz from the verbatim s3tokenizer.py (its third-party base class stubbed, refload.py)."""
    assert refload.available(), "reference tree not found"
    tok = refload.s3tokenizer_module().S3Tokenizer()
    out = {}
    for i, w in enumerate(s3_wavs()):
        out[f"mel_{i}"] = tok.log_mel_spectrogram(torch.from_numpy(w)[None])[0].numpy()           # (128, L // 160)
    batch = np.stack([s3_wavs()[0][:16000], s3_wavs()[1][:16000] * 1e-3])
    out["mel_batch"] = tok.log_mel_spectrogram(torch.from_numpy(batch)).numpy()                  # one call: the floor is global
    lens = [1, 639, 640, 641, 16000, 16001, 25599, 160000]
    out["pad_in"] = np.array(lens)
    out["pad_out"] = np.array([tok.pad([np.zeros(n, np.float32)], 16000)[0].shape[1] for n in lens])
    np.savez_compressed(os.path.join(OUT, "ref_s3_log_mel.npz"), **out)
    print("s3_log_mel", {k: v.shape for k, v in out.items()})


def main():
    assert refload.available(), "reference tree not found"
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    rve = refload.voice_encoder_module()
    hp = rve.VoiceEncConfig()

    # integer known answers from the reference's own get_num_wins / get_frame_step
    n_frames = np.arange(1, 4001)
    wins = np.array([rve.get_num_wins(int(n), 77, 0.8, hp) for n in n_frames], dtype=np.int64)
    wins80 = np.array([rve.get_num_wins(int(n), 80, 0.8, hp) for n in n_frames], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "ints.npz"), n_frames=n_frames, wins77=wins, wins80=wins80,
                        step_rate13=rve.get_frame_step(0.5, 1.3, hp), step_none=rve.get_frame_step(0.5, None, hp))

    wavs = golden_wavs()
    import librosa  # the oracle shim registered by refload
    trims = np.array([librosa.effects.trim(w, top_db=20)[1] for w in wavs], dtype=np.int64)
    for kind in ("W0", "W1", "W2"):
        ve = refload.make_voice_encoder(weights.ve_state_dict(kind))
        cp = refload.make_campplus(weights.campplus_state_dict(kind))
        out = {"trim": trims}
        out["ve_emb"] = ve.embeds_from_wavs(wavs, sample_rate=16000)
        out["ve_emb_notrim"] = ve.embeds_from_wavs(wavs, sample_rate=16000, trim_top_db=None)
        with torch.inference_mode():
            out["xv_emb"] = np.concatenate([cp.inference(torch.from_numpy(w)[None]).numpy() for w in wavs])
            # stage tensors for the 3 s clip (index 1)
            mel = rve.melspectrogram(wavs[1], hp).T.astype(np.float32)
            out["mel_1"] = mel
            n_p, _ = rve.get_num_wins(len(mel), 77, 0.8, hp)
            padded = np.concatenate([mel, np.zeros((400, 40), np.float32)])
            parts = np.stack([padded[77 * p: 77 * p + 160] for p in range(n_p)])
            out["partial_emb_1"] = ve(torch.from_numpy(parts)).numpy()
            _, (h_n, _) = ve.lstm(torch.from_numpy(parts))     # last hidden state of each of the three layers (3, P, 256)
            out["lstm_h_1"] = h_n.numpy()
            feat, _, _ = refload.xvector_module().extract_feature([torch.from_numpy(wavs[1])])
            out["fbank_cmn_1"] = feat[0].numpy()
            fcm = cp.head(feat.permute(0, 2, 1))
            out["fcm_1"] = fcm[0, :, ::16].numpy()            # (320, T/16) subsample
            td = cp.xvector.tdnn(fcm)
            out["tdnn_1"] = td[0, :, ::8].numpy()
            b1 = cp.xvector.block1(td)
            out["block1_1"] = b1[0, :, ::16].numpy()
            xvm = cp.xvector
            b2 = xvm.block2(xvm.transit1(b1))
            out["block2_1"] = b2[0, :, ::16].numpy()          # (1024, T'/16): first 256 rows = transit1 output
            b3 = xvm.block3(xvm.transit2(b2))
            out["block3_1"] = b3[0, :, ::16].numpy()          # (1024, T'/16): first 512 rows = transit2 output
            t3 = xvm.transit3(b3)
            out["transit3_1"] = t3[0, :, ::16].numpy()        # (512, T'/16), before out_nonlinear
            out["stats_1"] = xvm.stats(xvm.out_nonlinear(t3))[0].numpy()   # (1024,) mean | unbiased std
        np.savez_compressed(os.path.join(OUT, f"ref_{kind}.npz"), **out)
        print(kind, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    import sys
    if "--prompt-mel" in sys.argv:          # only this fixture (the others stay byte-identical in git)
        main_prompt_mel()
    elif "--s3" in sys.argv:
        main_s3()
    else:
        main()
        main_prompt_mel()
        main_s3()
