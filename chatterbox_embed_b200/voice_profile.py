"""The voice-profile container of the reference and its batched creation on the B200.

``VoiceProfile`` (s3gen.py:427-470) and ``save_voice_profile`` / ``load_voice_profile`` (tts.py:510-586, vc.py:606-708): a
``.npy`` file holding ONE pickled dict -- ``embedding`` (1,192) CAMPPlus x-vector, ``ve_embedding`` (1,256) VoiceEncoder
embedding, ``prompt_feat`` (1,T,80) 24 kHz log-mel and, when a tokenizer is available, ``prompt_token`` (1,T/2) /
``prompt_token_len`` (1,) -- which ``generate(voice_profile_path=...)`` consumes downstream.

``VoiceProfiler.save_voice_profiles`` builds many profiles per call: every clip of the batch goes through one ragged
resampler launch per target rate, one prompt-mel launch and one pass of both encoders, instead of the reference's
one-clip-at-a-time loop.  Deviation: the reference resamples the VoiceEncoder's input with ``librosa.resample`` (soxr_hq, a
third-party library absent here) and the CAMPPlus input with torchaudio; both encoders get the torchaudio-style resampler
here (identical for 16 kHz input, where neither resamples)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _host, _lib
from .campplus import CAMPPlus
from .mel import mel_spectrogram_ragged
from .resample import get_resampler
from .scheduler import SpeakerEmbedder
from .voice_encoder import VoiceEncoder

S3_SR, S3GEN_SR = 16000, 24000


class VoiceProfile:
    """s3gen.py:427-470; ``ve_embedding`` is attached by ``load_voice_profile`` like tts.py:572-578 does."""

    def __init__(self, embedding: torch.Tensor, prompt_feat: Optional[torch.Tensor] = None, prompt_feat_len: Optional[int] = None,
                 prompt_token: Optional[torch.Tensor] = None, prompt_token_len: Optional[torch.Tensor] = None):
        self.embedding = embedding
        self.prompt_feat = prompt_feat
        self.prompt_feat_len = prompt_feat_len
        self.prompt_token = prompt_token
        self.prompt_token_len = prompt_token_len

    @classmethod
    def load(cls, path: str, device: str = "cpu") -> "VoiceProfile":
        """s3gen.py:447-456: the pickled dict back into a profile (optional keys stay None)."""
        data = np.load(path, allow_pickle=True).item()
        def tensor(key):
            return torch.tensor(data[key]).to(device) if key in data else None
        return cls(embedding=tensor("embedding"), prompt_feat=tensor("prompt_feat"), prompt_feat_len=data.get("prompt_feat_len"),
                   prompt_token=tensor("prompt_token"), prompt_token_len=tensor("prompt_token_len"))

    def save(self, path: str):
        np.save(path, profile_dict(self))


def profile_dict(profile: VoiceProfile, ve_embedding: Optional[torch.Tensor] = None) -> dict:
    """The dict the reference pickles (key order as tts.py:537-549)."""
    data = {"embedding": profile.embedding.detach().cpu().numpy()}
    if ve_embedding is not None:
        data["ve_embedding"] = ve_embedding.detach().cpu().numpy()
    if profile.prompt_feat is not None:
        data["prompt_feat"] = profile.prompt_feat.detach().cpu().numpy()
    if profile.prompt_feat_len is not None:
        data["prompt_feat_len"] = profile.prompt_feat_len
    if profile.prompt_token is not None:
        data["prompt_token"] = profile.prompt_token.detach().cpu().numpy()
    if profile.prompt_token_len is not None:
        data["prompt_token_len"] = profile.prompt_token_len.detach().cpu().numpy()
    return data


def load_voice_profile(path: str, device="cpu") -> VoiceProfile:
    """tts.py:555-586 / vc.py:676-708: like VoiceProfile.load, plus the ``ve_embedding`` attribute (None for old files)."""
    data = np.load(path, allow_pickle=True).item()
    def tensor(key):
        return torch.from_numpy(data[key]).to(device) if key in data else None
    profile = VoiceProfile(embedding=tensor("embedding"), prompt_feat=tensor("prompt_feat"), prompt_feat_len=data.get("prompt_feat_len"),
                           prompt_token=tensor("prompt_token"), prompt_token_len=tensor("prompt_token_len"))
    profile.ve_embedding = tensor("ve_embedding")
    return profile


def load_audio(path: str) -> Tuple[np.ndarray, int]:
    """What ``librosa.load(path, sr=None)`` returns for a PCM / float WAV file: mono float32 in [-1, 1] and the native rate."""
    from scipy.io import wavfile
    sr, x = wavfile.read(path)
    if x.dtype == np.int16:
        y = x.astype(np.float32) / 32768.0
    elif x.dtype == np.int32:
        y = x.astype(np.float32) / 2147483648.0
    elif x.dtype == np.uint8:
        y = (x.astype(np.float32) - 128.0) / 128.0
    else:
        y = x.astype(np.float32)
    if y.ndim == 2:
        y = y.mean(axis=1)
    return np.ascontiguousarray(y), int(sr)


class VoiceProfiler:
    """Owner of the two encoders (and optionally an S3 tokenizer callable) that writes voice profiles."""

    def __init__(self, ve: VoiceEncoder, speaker_encoder: CAMPPlus, tokenizer=None):
        self.embedder = SpeakerEmbedder(ve, speaker_encoder)
        self.tokenizer = tokenizer
        self.device = ve.device

    @torch.inference_mode()
    def build_profiles(self, wavs: Sequence[Union[np.ndarray, torch.Tensor]], srs: Union[int, Sequence[int]]) -> List[dict]:
        """One profile dict per clip; clips may differ in length and sample rate."""
        n = len(wavs)
        srs = [int(srs)] * n if np.isscalar(srs) else [int(s) for s in srs]
        dev = self.device
        clips = [(torch.from_numpy(w) if isinstance(w, np.ndarray) else w).detach().reshape(-1).to(dev, torch.float32) for w in wavs]
        w16: List[Optional[torch.Tensor]] = [None] * n
        w24: List[Optional[torch.Tensor]] = [None] * n
        for sr in sorted(set(srs)):
            idx = [i for i in range(n) if srs[i] == sr]
            for target, dst in ((S3_SR, w16), (S3GEN_SR, w24)):
                outs = [clips[i] for i in idx] if sr == target else get_resampler(sr, target, dev).ragged([clips[i] for i in idx])
                for i, o in zip(idx, outs):
                    dst[i] = o
        mels = mel_spectrogram_ragged(w24)                                        # prompt_feat, (T_i, 80) each
        off = np.concatenate([[0], np.cumsum([int(w.numel()) for w in w16])]).astype(np.int64)
        ve, xv, status = self.embedder.embed_device(torch.cat(w16) if n > 1 else w16[0].contiguous(), off)
        _host.raise_for_status(status.cpu().numpy(), _lib.DO_VE | _lib.DO_XV)
        ve, xv = ve.cpu(), xv.cpu()
        out = []
        for i in range(n):
            prof = VoiceProfile(embedding=xv[i:i + 1], prompt_feat=mels[i][None])
            if self.tokenizer is not None:                                        # s3gen.py:189-200
                tokens, lens = self.tokenizer(w16[i][None])
                if prof.prompt_feat.shape[1] != 2 * tokens.shape[1]:
                    tokens = tokens[:, :prof.prompt_feat.shape[1] // 2]
                    lens = lens.clone().detach()
                    lens[0] = tokens.shape[1]
                prof.prompt_token, prof.prompt_token_len = tokens, lens
            out.append(profile_dict(prof, ve_embedding=ve[i:i + 1]))
        return out

    def save_voice_profiles(self, wavs, srs, save_paths: Sequence[str]):
        for data, path in zip(self.build_profiles(wavs, srs), save_paths):
            np.save(path, data)

    def save_voice_profile(self, audio: Union[str, Tuple[np.ndarray, int]], save_path: str):
        """tts.py:510-553: ``audio`` is a WAV path (as in the reference) or an ``(array, sample_rate)`` pair."""
        wav, sr = load_audio(audio) if isinstance(audio, str) else audio
        self.save_voice_profiles([wav], [sr], [save_path])

    def load_voice_profile(self, path: str) -> VoiceProfile:
        return load_voice_profile(path, self.device)
