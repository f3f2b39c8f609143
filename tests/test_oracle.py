"""CPU tests: the oracle against the golden vectors (produced by the verbatim reference, oracle/make_golden.py), against
torchaudio / torch.stft cross-checks, and -- when /root/reference is present -- against the reference modules live."""
import os

import numpy as np
import pytest
import torch

from chatterbox_embed_b200 import synth
from oracle import frontend, make_golden, nets, refload, weights


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_num_wins_matches_reference_table(golden_dir):
    g = _load(golden_dir, "ints.npz")
    for n, (w, t) in zip(g["n_frames"], g["wins77"]):
        assert nets.num_wins(int(n), 77, 0.8) == (int(w), int(t))
    for n, (w, t) in zip(g["n_frames"], g["wins80"]):
        assert nets.num_wins(int(n), 80, 0.8) == (int(w), int(t))
    assert nets.frame_step(0.5, 1.3) == int(g["step_rate13"]) == 77
    assert nets.frame_step(0.5, None) == int(g["step_none"]) == 80


def test_survey_known_answers():
    table = {1: (1, 160), 159: (1, 160), 160: (1, 160), 161: (1, 160), 236: (2, 237), 237: (2, 237), 300: (3, 314),
             301: (3, 314), 1001: (12, 1007), 3001: (38, 3009)}
    for n, want in table.items():
        assert nets.num_wins(n, 77, 0.8) == want


def test_mel_basis_matches_torchaudio_slaney():
    import torchaudio
    fb = torchaudio.functional.melscale_fbanks(201, 0., 8000., 40, 16000, norm="slaney", mel_scale="slaney").T.numpy()
    assert np.abs(fb - frontend.ve_mel_basis()).max() < 1e-7


def test_stft_matches_torch_stft():
    w = synth.clip(1, 20000)
    a = frontend.stft(w, 400, 160, 400, True, "reflect")
    b = torch.stft(torch.from_numpy(w), 400, 160, 400, torch.hann_window(400), center=True, pad_mode="reflect",
                   return_complex=True).numpy()
    assert a.shape == b.shape == (201, 126)
    assert np.abs(a - b).max() < 2e-5 * np.abs(b).max()


@pytest.mark.parametrize("idx", [0, 1])
def test_kaldi_restatement_matches_torchaudio(idx):
    w = synth.clip(idx, 32000)
    a, b = frontend.kaldi_fbank_numpy(w), frontend.kaldi_fbank_torchaudio(w)
    assert a.shape == b.shape == (198, 80)
    assert np.abs(a - b).mean() < 5e-4      # log domain; near-empty bins of a chirp differ by up to ~1e-2
    assert np.abs(a - b).max() < 5e-2


def test_trim_bounds_golden(golden_dir):
    g = _load(golden_dir, "ref_W0.npz")
    wavs = make_golden.golden_wavs()
    got = np.array([frontend.trim_bounds(w, 20) for w in wavs])
    assert (got == g["trim"]).all()
    assert (got[2] != [0, 50000]).all()         # the silent-edged clip is really trimmed


@pytest.mark.parametrize("kind", ["W0", "W1"])
def test_oracle_vs_golden_embeddings(golden_dir, kind):
    g = _load(golden_dir, f"ref_{kind}.npz")
    wavs = make_golden.golden_wavs()
    sdv, sdc = weights.ve_state_dict(kind), weights.campplus_state_dict(kind)
    ve = nets.ve_embed_wavs(sdv, wavs)
    assert np.abs(ve - g["ve_emb"]).max() < 2e-6
    ve2 = nets.ve_embed_wavs(sdv, wavs, trim_top_db=None)
    assert np.abs(ve2 - g["ve_emb_notrim"]).max() < 2e-6
    sel = [0, 1, 4]
    xv = nets.campplus_embed_wavs(sdc, [wavs[i] for i in sel])
    assert np.abs(xv - g["xv_emb"][sel]).max() < 2e-5 * max(1.0, np.abs(g["xv_emb"]).max())


@pytest.mark.parametrize("kind", ["W0", "W1"])
def test_oracle_vs_golden_stages(golden_dir, kind):
    g = _load(golden_dir, f"ref_{kind}.npz")
    w = make_golden.golden_wavs()[1]
    mel = frontend.ve_melspectrogram(w)
    assert np.abs(mel - g["mel_1"]).max() <= 1e-6 * np.abs(g["mel_1"]).max()
    sdv, sdc = weights.ve_state_dict(kind), weights.campplus_state_dict(kind)
    with torch.inference_mode():
        pe = nets.ve_forward(sdv, nets.ve_partials(mel)).numpy()
    assert np.abs(pe - g["partial_emb_1"]).max() < 2e-6
    feat = frontend.campplus_features(w)
    assert np.abs(feat - g["fbank_cmn_1"]).max() < 1e-5
    taps = {}
    nets.campplus_embed_wavs(sdc, [w], taps)
    for key, stride in (("fcm", 16), ("tdnn", 8), ("block1", 16)):
        a = taps[key][0, :, ::stride].numpy()
        b = g[f"{key}_1"]
        assert np.abs(a - b).max() <= 1e-4 * max(1.0, np.abs(b).max()), key


def test_npy_fixture_format_roundtrip(tmp_path):
    """audio_test/reference_voice_clone.npy pins the FORMAT (NPY v1, <f4, (1,192), C order): np.save reproduces it."""
    emb = np.arange(192, dtype=np.float32)[None] / 7
    p = tmp_path / "clone.npy"
    np.save(p, emb)
    raw = p.read_bytes()
    assert len(raw) == 896 and raw[:8] == b"\x93NUMPY\x01\x00"
    assert b"'descr': '<f4'" in raw[:128] and b"'fortran_order': False" in raw[:128] and b"'shape': (1, 192)" in raw[:128]
    ref_fixture = os.path.join(refload.REF_ROOT, "audio_test", "reference_voice_clone.npy")
    if os.path.exists(ref_fixture):
        rb = open(ref_fixture, "rb").read()
        assert rb[:128] == raw[:128] and len(rb) == 896


@pytest.mark.skipif(not refload.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("kind", ["W0", "W1"])
def test_oracle_vs_live_reference(kind):
    sdv, sdc = weights.ve_state_dict(kind), weights.campplus_state_dict(kind)
    ve = refload.make_voice_encoder(sdv)
    cp = refload.make_campplus(sdc)
    wavs = [synth.clip(10, 20000), synth.with_silence(11, 30000, 4000, 3000)]
    assert np.abs(ve.embeds_from_wavs(wavs, 16000) - nets.ve_embed_wavs(sdv, wavs)).max() < 2e-6
    with torch.inference_mode():
        r = np.concatenate([cp.inference(torch.from_numpy(w)[None]).numpy() for w in wavs])
    assert np.abs(r - nets.campplus_embed_wavs(sdc, wavs)).max() < 2e-5 * max(1.0, np.abs(r).max())


def test_w2_is_sensitive():
    """The sensitised set must tell a correct front-end from a zeroed one (SURVEY.md 8d hazard 1)."""
    sdv = weights.ve_state_dict("W2")
    a = nets.ve_embed_wavs(sdv, [synth.clip(0, 32000)])[0]
    b = nets.ve_embed_mels(sdv, [np.zeros((201, 40), np.float32)])[0]
    assert float(a @ b) < 0.99


@pytest.mark.parametrize("src,dst", [(24000, 16000), (44100, 16000), (22050, 16000), (48000, 16000), (16000, 24000)])
def test_resample_restatement_matches_torchaudio(src, dst):
    """get_resampler (s3gen.py:41-44) = torchaudio.transforms.Resample with defaults: the oracle's restatement against the
    library installed here (torchaudio is importable in the container and on the GPU box)."""
    rng = np.random.RandomState(7)
    wav = (0.3 * rng.randn(src // 3 + 17)).astype(np.float32)
    a = frontend.resample_numpy(wav, src, dst)
    b = frontend.resample_torchaudio(wav, src, dst)
    # the filter banks are identical (checked below); what differs is the accumulation: torch's CPU conv1d sums the up to 475
    # taps in float32 (noise ~1e-5 on O(1) signals, not even stable between two conv1d call shapes), the restatement in float64
    assert a.shape == b.shape and np.abs(a - b).max() < 3e-5
    import math
    import torch
    import torchaudio.functional.functional as TF
    g = math.gcd(src, dst)
    bank, width = TF._get_sinc_resample_kernel(src, dst, g)
    assert np.abs(frontend.resample_bank(src, dst)[0] - bank.numpy()[:, 0, :]).max() == 0.0 and frontend.resample_bank(src, dst)[1] == width


# ---- S3Gen prompt mel (s3gen/utils/mel.py:33-81): restatements against the fixture the verbatim reference produced ----------
def prompt_mel_close(got, ref):
    """The reference computes the 1920-point FFT in fp32: near the 1e-5 clamp its own log-mels sit up to ~1e-2 from the exact
    value (2.6e-6 in the linear domain), so parity is stated in the linear domain: |d mel| <= 1e-5 + 5e-4 * mel."""
    g, r = np.exp(np.asarray(got, np.float64)), np.exp(np.asarray(ref, np.float64))
    return got.shape == ref.shape and bool(np.all(np.abs(g - r) <= 1e-5 + 5e-4 * r))


def test_prompt_mel_oracle_vs_reference_fixture(golden_dir):
    g = _load(golden_dir, "ref_prompt_mel.npz")
    wavs = make_golden.prompt_mel_wavs()
    for i, w in enumerate(wavs):
        ref = g[f"mel_{i}"]
        assert ref.shape == (frontend.prompt_mel_num_frames(len(w)), 80)
        assert np.abs(frontend.prompt_mel_torch(w) - ref).max() < 1e-5          # same torch ops: equal up to thread-count effects
        assert prompt_mel_close(frontend.prompt_mel_numpy(w), ref)             # float64 DFT restatement
    assert np.abs(g["mel_batch"][0].T - frontend.prompt_mel_torch(wavs[0][:24000])).max() < 1e-5
    with pytest.raises(ValueError):
        frontend.prompt_mel_num_frames(720)
    assert [frontend.prompt_mel_num_frames(n) for n in (721, 959, 960, 240000)] == [1, 1, 2, 500]


def test_prompt_mel_basis_is_two_sparse():
    """The kernel's epilogue relies on every DFT bin feeding at most two of the 80 filters and on bins 0 and >= 640 having no weight."""
    B = frontend.prompt_mel_basis()
    assert B.shape == (80, 961) and (B != 0).sum(0).max() == 2
    used = np.flatnonzero((B != 0).any(0))
    assert used[0] == 1 and used[-1] == 639


# ---- S3Tokenizer front-end (s3tokenizer/s3tokenizer.py:52-74, 128-168) ------------------------------------------------------
def test_s3_log_mel_oracle_vs_reference_fixture(golden_dir):
    g = _load(golden_dir, "ref_s3_log_mel.npz")
    for i, w in enumerate(make_golden.s3_wavs()):
        ref = g[f"mel_{i}"]
        assert ref.shape == (128, len(w) // 160)
        assert np.abs(frontend.s3_log_mel_torch(w) - ref).max() < 1e-5           # the reference's own torch ops
        assert np.abs(frontend.s3_log_mel_numpy(w) - ref).max() < 2e-4           # float64 DFT; the reference's fp32 FFT is ~6e-5 away
    B = frontend.s3_mel_basis()
    assert B.shape == (128, 201) and (B != 0).sum(0).max() == 2                  # what the kernel's 2-sparse epilogue relies on
    used = np.flatnonzero((B != 0).any(0))
    assert used[0] == 1 and used[-1] == 199
