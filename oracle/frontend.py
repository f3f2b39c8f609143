"""Oracle (TEST INFRASTRUCTURE): CPU restatement of the two audio front-ends.

VoiceEncoder front-end = librosa 0.11.0 calls made by the reference
(``melspec.py:9-16`` filters.mel, ``melspec.py:54-64`` stft,
``voice_encoder.py:267`` effects.trim).  librosa is a third-party dependency
pinned at 0.11.0 in the reference's ``pyproject.toml:13`` and is absent from
/root/reference and from this image, so its published algorithm is restated
here (parity unpinned at that boundary, see ``oracle/__init__.py``).

CAMPPlus front-end = ``torchaudio.compliance.kaldi.fbank(num_mel_bins=80)``
followed by per-utterance mean subtraction (``xvector.py:45-58``).  torchaudio
is installed, so ``kaldi_fbank_numpy`` is validated against it directly.
"""
from __future__ import annotations

import math
import types

import numpy as np

# ----------------------------------------------------------------------------
# librosa.filters.mel  (Slaney scale, slaney area-norm)  -- melspec.py:11-16
# ----------------------------------------------------------------------------
_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = math.log(6.4) / 27.0


def hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    mel = f / _F_SP
    big = f >= _MIN_LOG_HZ
    return np.where(big, _MIN_LOG_MEL + np.log(np.maximum(f, 1e-30) / _MIN_LOG_HZ) / _LOGSTEP, mel)


def mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    hz = m * _F_SP
    big = m >= _MIN_LOG_MEL
    return np.where(big, _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL)), hz)


def filters_mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
    """librosa.filters.mel(htk=False, norm='slaney', dtype=float32)."""
    if fmax is None:
        fmax = sr / 2.0
    n_bins = 1 + n_fft // 2
    fft_hz = np.linspace(0.0, sr / 2.0, n_bins)
    edges = mel_to_hz_slaney(np.linspace(hz_to_mel_slaney(fmin), hz_to_mel_slaney(fmax), n_mels + 2))
    widths = np.diff(edges)
    dist = edges[:, None] - fft_hz[None, :]
    bank = np.zeros((n_mels, n_bins), dtype=np.float32)
    for m in range(n_mels):
        rising = -dist[m] / widths[m]
        falling = dist[m + 2] / widths[m + 1]
        bank[m] = np.maximum(0.0, np.minimum(rising, falling))
    area = 2.0 / (edges[2:] - edges[:-2])
    bank *= area[:, None]
    return bank


# ----------------------------------------------------------------------------
# librosa.stft  -- melspec.py:57-64 (center=True, reflect, periodic hann)
# ----------------------------------------------------------------------------
def hann_periodic(n):
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def stft(y, n_fft=2048, hop_length=None, win_length=None, center=True, pad_mode="constant"):
    y = np.asarray(y)
    win_length = n_fft if win_length is None else win_length
    hop_length = win_length // 4 if hop_length is None else hop_length
    win = hann_periodic(win_length)
    if win_length < n_fft:
        lp = (n_fft - win_length) // 2
        win = np.pad(win, (lp, n_fft - win_length - lp))
    if center:
        y = np.pad(y, n_fft // 2, mode=pad_mode)
    n_frames = 1 + (len(y) - n_fft) // hop_length
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(n_frames)[None, :]
    frames = y[idx].astype(np.float64) * win[:, None]
    out = np.fft.rfft(frames, axis=0)
    return out.astype(np.complex64 if y.dtype == np.float32 else np.complex128)


# ----------------------------------------------------------------------------
# librosa.effects.trim -- voice_encoder.py:267
# ----------------------------------------------------------------------------
def trim_bounds(y, top_db=60.0, frame_length=2048, hop_length=512):
    """[start, end) sample range kept by librosa.effects.trim(y, top_db)."""
    y = np.asarray(y)
    n = len(y)
    yp = np.pad(y, frame_length // 2, mode="constant")
    n_frames = 1 + (len(yp) - frame_length) // hop_length
    idx = np.arange(frame_length)[None, :] + hop_length * np.arange(n_frames)[:, None]
    power = np.mean(np.abs(yp[idx]) ** 2, axis=1)
    rms = np.sqrt(power)
    amin = 1e-5
    ref = rms.max() if n_frames else 0.0
    db = 10.0 * np.log10(np.maximum(amin ** 2, rms.astype(np.float64) ** 2))
    db -= 10.0 * np.log10(max(amin ** 2, float(ref) ** 2))
    keep = np.flatnonzero(db > -top_db)
    if keep.size == 0:
        return 0, 0
    return int(keep[0]) * hop_length, min(n, (int(keep[-1]) + 1) * hop_length)


def effects_trim(y, top_db=60.0, frame_length=2048, hop_length=512):
    s, e = trim_bounds(y, top_db, frame_length, hop_length)
    return y[s:e], np.asarray([s, e])


def _no_resample(*a, **k):  # voice_encoder.py:262 is dead code for 16 kHz input
    raise NotImplementedError("oracle shim: librosa.resample (kaiser_fast) is out of scope")


def make_librosa_shim():
    """Module object standing in for ``import librosa`` when the verbatim
    reference files are imported."""
    mod = types.ModuleType("librosa")
    mod.stft = stft
    mod.resample = _no_resample
    mod.filters = types.ModuleType("librosa.filters")
    mod.filters.mel = filters_mel
    mod.effects = types.ModuleType("librosa.effects")
    mod.effects.trim = effects_trim
    mod.__version__ = "0.11.0-oracle-shim"
    return mod


def register_librosa_shim():
    """``import librosa`` and ``from librosa.filters import mel`` (s3gen/utils/mel.py:2) both resolve to the shim."""
    import sys
    mod = sys.modules.get("librosa")
    if mod is None or not str(getattr(mod, "__version__", "")).endswith("oracle-shim"):
        mod = make_librosa_shim()
        sys.modules["librosa"] = mod
    sys.modules.setdefault("librosa.filters", mod.filters)
    sys.modules.setdefault("librosa.effects", mod.effects)
    return mod


# ----------------------------------------------------------------------------
# VoiceEncoder mel -- melspec.py:26-51 with hp of config.py:1-18
# ----------------------------------------------------------------------------
VE_SR, VE_NFFT, VE_HOP, VE_NMEL = 16000, 400, 160, 40
_ve_basis = None


def ve_mel_basis():
    global _ve_basis
    if _ve_basis is None:
        _ve_basis = filters_mel(VE_SR, VE_NFFT, VE_NMEL, 0.0, 8000.0)
    return _ve_basis


def ve_melspectrogram(wav):
    """(L,) float32 -> (T, 40) float32, T = 1 + L // 160.  Power mel, no log."""
    spec = stft(np.asarray(wav, dtype=np.float32), VE_NFFT, VE_HOP, VE_NFFT, True, "reflect")
    mag = np.abs(spec)
    mag **= 2.0
    mel = np.dot(ve_mel_basis(), mag)
    return np.ascontiguousarray(mel.T.astype(np.float32))


# ----------------------------------------------------------------------------
# Kaldi fbank (torchaudio.compliance.kaldi.fbank defaults, 80 bins) + CMN
# ----------------------------------------------------------------------------
KALDI_EPS = float(np.finfo(np.float32).eps)


def kaldi_mel_banks(num_bins=80, padded=512, sr=16000.0, low=20.0, high=0.0):
    nyq = 0.5 * sr
    if high <= 0:
        high += nyq
    mel = lambda f: 1127.0 * np.log(1.0 + np.asarray(f, dtype=np.float64) / 700.0)
    lo, hi = mel(low), mel(high)
    delta = (hi - lo) / (num_bins + 1)
    b = np.arange(num_bins, dtype=np.float64)[:, None]
    left, center, right = lo + b * delta, lo + (b + 1) * delta, lo + (b + 2) * delta
    m = mel((sr / padded) * np.arange(padded // 2))[None, :]
    up = (m - left) / (center - left)
    down = (right - m) / (right - center)
    bank = np.maximum(0.0, np.minimum(up, down))
    return np.pad(bank, ((0, 0), (0, 1)))  # Nyquist bin gets weight 0


def povey_window(n=400):
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / (n - 1))) ** 0.85


def kaldi_num_frames(n_samples, win=400, hop=160):
    return 0 if n_samples < win else 1 + (n_samples - win) // hop


def kaldi_fbank_numpy(wav, num_bins=80):
    """float64 restatement of Kaldi.fbank(wav[None], num_mel_bins=80) -> (T_k, 80)."""
    y = np.asarray(wav, dtype=np.float64)
    tk = kaldi_num_frames(len(y))
    assert tk > 0, "choose a window size 400 that is [2, len]"
    fr = y[np.arange(400)[None, :] + 160 * np.arange(tk)[:, None]]
    fr = fr - fr.mean(axis=1, keepdims=True)
    prev = np.concatenate([fr[:, :1], fr[:, :-1]], axis=1)
    fr = (fr - 0.97 * prev) * povey_window()[None, :]
    fr = np.pad(fr, ((0, 0), (0, 112)))
    pw = np.abs(np.fft.rfft(fr, axis=1)) ** 2
    en = pw @ kaldi_mel_banks(num_bins).T
    return np.log(np.maximum(en, KALDI_EPS)).astype(np.float32)


def kaldi_fbank_torchaudio(wav, num_bins=80):
    import torch
    import torchaudio.compliance.kaldi as K
    return K.fbank(torch.as_tensor(np.asarray(wav, dtype=np.float32))[None], num_mel_bins=num_bins).numpy()


def campplus_features(wav, use_torchaudio=True):
    """fbank minus its own column mean (xvector.py:50-51)."""
    f = kaldi_fbank_torchaudio(wav) if use_torchaudio else kaldi_fbank_numpy(wav)
    return (f - f.mean(axis=0, keepdims=True)).astype(np.float32)


# ----------------------------------------------------------------------------
# torchaudio.transforms.Resample (get_resampler, s3gen/s3gen.py:41-44): restatement + the library itself
# ----------------------------------------------------------------------------
def resample_bank(src_sr, dst_sr, lowpass_filter_width=6, rolloff=0.99):
    """(bank [new][2*width+orig] float32, width) exactly as torchaudio functional.py:_get_sinc_resample_kernel builds it."""
    import math
    g = math.gcd(int(src_sr), int(dst_sr))
    orig, new = int(src_sr) // g, int(dst_sr) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    # torchaudio divides an int64 arange by new_freq: that quotient is float32 (default dtype) before it meets the float64 idx
    t = ((np.arange(0, -new, -1).astype(np.float32) / np.float32(new)).astype(np.float64)[:, None] + idx) * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    kern = np.where(t == 0, 1.0, np.sin(t) / np.where(t == 0, 1.0, t)) * window * (base / orig)
    return kern.astype(np.float32), width


def resample_numpy(wav, src_sr, dst_sr):
    """sinc_interp_hann polyphase FIR as torchaudio applies it (functional.py:_apply_sinc_resample_kernel), accumulated in
    float64 (torch's conv1d accumulates in float32)."""
    import math
    wav = np.asarray(wav, dtype=np.float32)
    if src_sr == dst_sr:
        return wav
    g = math.gcd(int(src_sr), int(dst_sr))
    orig, new = int(src_sr) // g, int(dst_sr) // g
    kern, width = resample_bank(src_sr, dst_sr)
    L = len(wav)
    xp = np.concatenate([np.zeros(width, np.float32), wav, np.zeros(width + orig, np.float32)])
    n_frames = (len(xp) - kern.shape[1]) // orig + 1
    frames = np.lib.stride_tricks.as_strided(xp, shape=(n_frames, kern.shape[1]), strides=(xp.strides[0] * orig, xp.strides[0]))
    out = (frames.astype(np.float64) @ kern.T.astype(np.float64)).astype(np.float32).reshape(-1)
    return out[: int(math.ceil(new * L / orig))]


def resample_torchaudio(wav, src_sr, dst_sr):
    import torch
    import torchaudio
    return torchaudio.functional.resample(torch.as_tensor(np.asarray(wav, dtype=np.float32)), int(src_sr), int(dst_sr)).numpy()


# ----------------------------------------------------------------------------
# S3Gen prompt mel (24 kHz) -- s3gen/utils/mel.py:33-81 with the defaults of :20-29
# ----------------------------------------------------------------------------
PM_SR, PM_NFFT, PM_HOP, PM_NMEL, PM_FMAX = 24000, 1920, 480, 80, 8000.0
PM_PAD = (PM_NFFT - PM_HOP) // 2
_pm_basis = None


def prompt_mel_basis():
    global _pm_basis
    if _pm_basis is None:
        _pm_basis = filters_mel(PM_SR, PM_NFFT, PM_NMEL, 0.0, PM_FMAX)
    return _pm_basis


def prompt_mel_num_frames(n_samples):
    """mel.py:56-74: reflect pad by 720 each side, then non-centred frames of 1920 every 480."""
    if n_samples <= PM_PAD:
        raise ValueError("reflect padding needs more than 720 samples")
    return 1 + (n_samples + 2 * PM_PAD - PM_NFFT) // PM_HOP


def prompt_mel_numpy(wav):
    """(L,) float32 at 24 kHz -> (T, 80) float32 = mel_spectrogram(wav)[0].T, float64 DFT."""
    y = np.pad(np.asarray(wav, dtype=np.float32), PM_PAD, mode="reflect")
    spec = stft(y, PM_NFFT, PM_HOP, PM_NFFT, False)                      # hann_window(1920) is periodic, like ours
    mag = np.sqrt(spec.real.astype(np.float64) ** 2 + spec.imag.astype(np.float64) ** 2 + 1e-9).astype(np.float32)
    mel = prompt_mel_basis().astype(np.float32) @ mag
    return np.ascontiguousarray(np.log(np.maximum(mel, 1e-5)).T.astype(np.float32))


def prompt_mel_torch(wav):
    """The same through torch.stft (fp32), operation for operation what mel.py does."""
    import torch
    y = torch.as_tensor(np.asarray(wav, dtype=np.float32))[None]
    y = torch.nn.functional.pad(y.unsqueeze(1), (PM_PAD, PM_PAD), mode="reflect").squeeze(1)
    spec = torch.view_as_real(torch.stft(y, PM_NFFT, hop_length=PM_HOP, win_length=PM_NFFT, window=torch.hann_window(PM_NFFT),
                                         center=False, pad_mode="reflect", normalized=False, onesided=True, return_complex=True))
    spec = torch.sqrt(spec.pow(2).sum(-1) + 1e-9)
    mel = torch.matmul(torch.from_numpy(prompt_mel_basis()).float(), spec)
    return np.ascontiguousarray(torch.log(torch.clamp(mel, min=1e-5))[0].T.numpy())


# ----------------------------------------------------------------------------
# S3Tokenizer log-mel (16 kHz, 128 bins) -- s3tokenizer/s3tokenizer.py:128-168 (filters :39-47)
# ----------------------------------------------------------------------------
_s3_basis = None


def s3_mel_basis():
    global _s3_basis
    if _s3_basis is None:
        _s3_basis = filters_mel(16000, 400, 128)
    return _s3_basis


def s3_log_mel_numpy(wav):
    """(L,) float32 -> (128, L // 160) float32, float64 DFT."""
    spec = stft(np.asarray(wav, dtype=np.float32), 400, 160, 400, True, "reflect")
    mag = (np.abs(spec.astype(np.complex128))[:, :-1] ** 2).astype(np.float32)
    mel = s3_mel_basis().astype(np.float32) @ mag
    log_spec = np.log10(np.maximum(mel, 1e-10))
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)
    return ((log_spec + 4.0) / 4.0).astype(np.float32)


def s3_log_mel_torch(wav):
    """The same with the reference's torch ops (fp32 torch.stft)."""
    import torch
    audio = torch.as_tensor(np.asarray(wav, dtype=np.float32))
    st = torch.stft(audio, 400, 160, window=torch.hann_window(400), return_complex=True)
    magnitudes = st[..., :-1].abs() ** 2
    mel_spec = torch.from_numpy(s3_mel_basis()).float() @ magnitudes
    log_spec = torch.clamp(mel_spec, min=1e-10).log10()
    log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
    return ((log_spec + 4.0) / 4.0).numpy()
