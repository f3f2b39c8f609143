"""``mel_spectrogram`` of the reference's ``s3gen/utils/mel.py:33-81`` (the Matcha-TTS extractor with CosyVoice's settings,
mel.py:20-29) on the B200: the ``prompt_feat`` of ``S3Token2Mel.embed_ref`` (s3gen.py:177).  Same signature and return
value -- ``(B, 80, T)`` log-mel, ``T = 1 + (L - 480) // 480`` -- computed by one tensor-core kernel (csrc/promptmel_tc.cu).
Only the reference's default parameters are offered; there is no CPU fallback."""
from __future__ import annotations

from typing import List, Sequence, Union

import numpy as np
import torch

from . import _host, _lib

N_FFT, NUM_MELS, SAMPLING_RATE, HOP_SIZE, WIN_SIZE, FMIN, FMAX = 1920, 80, 24000, 480, 1920, 0, 8000


def _default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.CbxError("mel_spectrogram needs a CUDA device (B200, sm_100); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def mel_spectrogram_ragged(clips: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """24 kHz clips of different lengths in one launch; one ``(T_i, 80)`` tensor per clip (views of one buffer)."""
    dev = clips[0].device
    if dev.type != "cuda":
        dev = _default_device()
    ctx = _lib.context(_host.device_index(dev))
    lens = [int(c.numel()) for c in clips]
    for n in lens:
        if n <= (N_FFT - HOP_SIZE) // 2:
            # torch.nn.functional.pad(mode="reflect") in the reference raises for these (mel.py:56-58)
            raise RuntimeError(f"Padding size should be less than the corresponding input dimension, but got: padding (720, 720) "
                               f"at dimension 2 of input [1, 1, {n}]")
    frames = [_lib.prompt_mel_frames(n) for n in lens]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    row = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
    x = torch.cat([c.reshape(-1).to(dev, torch.float32) for c in clips]) if len(clips) > 1 else clips[0].reshape(-1).to(dev, torch.float32).contiguous()
    out = torch.empty((int(row[-1]), NUM_MELS), dtype=torch.float32, device=dev)
    ctx.prompt_mel(x.data_ptr(), off, out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    return [out[a:b] for a, b in zip(row[:-1], row[1:])]


def mel_spectrogram(y: Union[torch.Tensor, np.ndarray], n_fft=N_FFT, num_mels=NUM_MELS, sampling_rate=SAMPLING_RATE, hop_size=HOP_SIZE,
                    win_size=WIN_SIZE, fmin=FMIN, fmax=FMAX, center=False) -> torch.Tensor:
    if (n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center) != (N_FFT, NUM_MELS, SAMPLING_RATE, HOP_SIZE, WIN_SIZE, FMIN, FMAX, False):
        raise NotImplementedError("only the reference's default mel settings (mel.py:20-29) are built")
    if isinstance(y, np.ndarray):
        y = torch.tensor(y).float()
    if len(y.shape) == 1:
        y = y[None, ]
    if torch.min(y) < -1.0:
        print("min value is ", torch.min(y))
    if torch.max(y) > 1.0:
        print("max value is ", torch.max(y))
    mels = mel_spectrogram_ragged([row for row in y])
    return torch.stack(mels).transpose(1, 2)           # (B, 80, T), as the reference returns it
