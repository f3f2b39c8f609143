"""``torchaudio.transforms.Resample(src_sr, dst_sr)`` (defaults: sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99) on the
B200: the resampler the reference gets from ``get_resampler`` (s3gen/s3gen.py:41-44) for native-rate -> 16 kHz / 24 kHz.
Same call surface: an object called on a ``(..., L)`` float tensor on the CUDA device, returning ``(..., ceil(L*dst/src))``."""
from __future__ import annotations

from functools import lru_cache
from typing import List, Sequence

import numpy as np
import torch

from . import _host, _lib


class Resample:
    def __init__(self, orig_freq: int = 16000, new_freq: int = 16000):
        if int(orig_freq) != orig_freq or int(new_freq) != new_freq:
            raise Exception("Frequencies must be of integer type to ensure quality resampling computation.")   # as torchaudio
        self.orig_freq, self.new_freq = int(orig_freq), int(new_freq)

    def to(self, device):          # API compatibility with nn.Module.to(); the filter bank lives in the libcbx context
        return self

    def __call__(self, waveform: torch.Tensor) -> torch.Tensor:
        if self.orig_freq == self.new_freq:
            return waveform
        if not waveform.is_floating_point():
            raise TypeError(f"Expected floating point type for waveform tensor, but received {waveform.dtype}.")
        shape = waveform.shape
        x = waveform.detach().to(torch.float32).reshape(-1, shape[-1]).contiguous()
        outs = self.ragged([row for row in x])
        return torch.stack(outs).reshape(shape[:-1] + (outs[0].shape[-1],))

    def ragged(self, clips: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        """Clips of different lengths in one launch; returns one tensor per clip (views of one buffer)."""
        dev = clips[0].device
        ctx = _lib.context(_host.device_index(dev))
        lens = [int(c.numel()) for c in clips]
        in_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        out_lens = [_lib.resample_out_len(self.orig_freq, self.new_freq, n) for n in lens]
        out_off = np.concatenate([[0], np.cumsum(out_lens)]).astype(np.int64)
        x = torch.cat([c.reshape(-1).to(dev, torch.float32) for c in clips]) if len(clips) > 1 else clips[0].reshape(-1).to(dev, torch.float32).contiguous()
        y = torch.empty(int(out_off[-1]), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        ctx.resample(x.data_ptr(), in_off, self.orig_freq, self.new_freq, y.data_ptr(), out_off, stream)
        return [y[a:b] for a, b in zip(out_off[:-1], out_off[1:])]


@lru_cache(100)
def get_resampler(src_sr, dst_sr, device=None):
    """s3gen.py:41-44."""
    return Resample(src_sr, dst_sr)
