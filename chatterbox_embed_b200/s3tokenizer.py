"""Front-end of the reference's ``S3Tokenizer`` (s3tokenizer/s3tokenizer.py:22-168) on the B200: ``pad``, ``_prepare_audio`` and
``log_mel_spectrogram`` -- the 128-bin log-mel at 100 frames/s that ``quantize()`` consumes.  The tokenizer network itself is the
third-party ``s3tokenizer`` package (S3TokenizerV2), which is not part of this library: pass its ``quantize`` in as
``quantizer`` to get tokens out of ``forward``."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _host, _lib

S3_SR = 16_000
S3_HOP = 160            # 100 frames/sec
S3_TOKEN_HOP = 640      # 25 tokens/sec
S3_TOKEN_RATE = 25
SPEECH_VOCAB_SIZE = 6561
N_MELS = 128


def log_mel_spectrogram_ragged(clips: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """16 kHz clips of different lengths in one launch pair; one ``(128, T_i)`` tensor per clip (views of one buffer)."""
    dev = clips[0].device
    ctx = _lib.context(_host.device_index(dev))
    lens = [int(c.numel()) for c in clips]
    for n in lens:
        if n <= 200:
            raise RuntimeError(f"Padding size should be less than the corresponding input dimension, but got: padding (200, 200) "
                               f"at dimension 2 of input [1, 1, {n}]")          # what torch.stft(center=True) raises
    frames = [_lib.s3_log_mel_frames(n) for n in lens]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    row = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
    x = torch.cat([c.reshape(-1).to(dev, torch.float32) for c in clips]) if len(clips) > 1 else clips[0].reshape(-1).to(dev, torch.float32).contiguous()
    out = torch.empty(int(row[-1]) * N_MELS, dtype=torch.float32, device=dev)
    if row[-1] > 0:
        ctx.s3_log_mel(x.data_ptr(), off, out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    return [out[a * N_MELS:b * N_MELS].view(N_MELS, b - a) for a, b in zip(row[:-1], row[1:])]


class S3TokenizerFrontend:
    def __init__(self, device="cuda", quantizer: Optional[Callable] = None):
        self.device = torch.device(device)
        self.n_fft = 400
        self.quantizer = quantizer

    @staticmethod
    def _as_batch(wav) -> torch.Tensor:
        """numpy or tensor, (L,) or (B, L) -> tensor (B, L)."""
        t = torch.from_numpy(wav) if isinstance(wav, np.ndarray) else wav
        return t.unsqueeze(0) if t.dim() == 1 else t

    def pad(self, wavs, sr) -> List[torch.Tensor]:
        """s3tokenizer.py:52-74: every wav zero-padded on the right to a whole number of 40 ms tokens (25 tokens/s), with the
        reference's float arithmetic (ceil of length / sr * 25, then int of tokens * (sr / 25))."""
        out = []
        for wav in wavs:
            t = self._as_batch(wav)
            tokens = np.ceil((t.shape[1] / sr) * S3_TOKEN_RATE)
            want = int(tokens * (sr / S3_TOKEN_RATE))
            out.append(torch.nn.functional.pad(t, (0, want - t.shape[-1]), mode="constant", value=0))
        return out

    def _prepare_audio(self, wavs):
        """s3tokenizer.py:76-87: a list of (1, L) tensors."""
        return [self._as_batch(w) for w in wavs]

    @torch.no_grad()
    def log_mel_spectrogram(self, audio: torch.Tensor, padding: int = 0) -> torch.Tensor:
        """s3tokenizer.py:128-168: ``(*, L)`` -> ``(*, 128, L // 160)``; the ``max - 8`` floor is over the whole call, as there."""
        if not torch.is_tensor(audio):
            audio = torch.from_numpy(audio)
        audio = audio.to(self.device)
        if padding > 0:
            audio = torch.nn.functional.pad(audio, (0, padding))
        lead = audio.shape[:-1]
        rows = audio.reshape(-1, audio.shape[-1])
        mels = torch.stack(log_mel_spectrogram_ragged([r for r in rows]))
        if rows.shape[0] > 1:
            # the kernel floors per clip; a multi-row call of the reference floors at the maximum of the whole tensor
            mels = torch.maximum(mels, mels.max() - 2.0)            # (x + 4) / 4 is monotone: "max - 8" becomes "max - 2"
        return mels.reshape(lead + mels.shape[-2:])

    @torch.no_grad()
    def mels(self, wavs, max_len: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """The mel half of ``forward`` (s3tokenizer.py:104-115): all clips in one ragged launch, truncated to ``4 * max_len``
        frames and zero-padded to the longest like ``s3tokenizer.utils.padding`` -> ``(B, 128, T_max)``, ``(B,)`` lengths."""
        clips = [w.reshape(-1) for w in self._prepare_audio(wavs)]
        mels = log_mel_spectrogram_ragged([c.to(self.device) for c in clips])
        if max_len is not None:
            mels = [m[..., :max_len * 4] for m in mels]
        lens = torch.tensor([m.shape[-1] for m in mels], dtype=torch.int32)
        out = torch.zeros((len(mels), N_MELS, int(lens.max())), dtype=torch.float32, device=self.device)
        for i, m in enumerate(mels):
            out[i, :, :m.shape[-1]] = m
        return out, lens

    @torch.no_grad()
    def forward(self, wavs, accelerator=None, max_len: Optional[int] = None):
        if self.quantizer is None:
            raise NotImplementedError("the S3TokenizerV2 network (third-party s3tokenizer package) is not part of this library; "
                                      "construct S3TokenizerFrontend(quantizer=model.quantize) to get tokens")
        mels, mel_lens = self.mels(wavs, max_len)
        speech_tokens, speech_token_lens = self.quantizer(mels, mel_lens.to(self.device))
        return speech_tokens.long().detach(), speech_token_lens.long().detach()

    __call__ = forward
