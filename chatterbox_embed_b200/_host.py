"""Host-side plumbing shared by the drop-in classes: flattening clips, workspaces, status -> exceptions."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def device_index(device: torch.device) -> int:
    if device.type != "cuda":
        raise _lib.CbxError(f"chatterbox_embed_b200 runs on a B200 only (no CPU fallback); module is on {device}. "
                            "Move it with .to('cuda').")
    return device.index if device.index is not None else torch.cuda.current_device()


def flatten_host(wavs: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    lens = np.array([len(w) for w in wavs], dtype=np.int64)
    off = np.zeros(len(wavs) + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    flat = np.empty(int(off[-1]), np.float32)
    for w, a, b in zip(wavs, off[:-1], off[1:]):
        flat[a:b] = np.asarray(w, dtype=np.float32).reshape(-1)
    return flat, off


class Workspace:
    """Grow-only torch-owned device scratch (torch is the allocator; libcbx only carves it)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes: int, device: torch.device) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = None
            self.buf = torch.empty(int(nbytes * 1.1) + 4096, dtype=torch.uint8, device=device)
        return self.buf


def raise_for_status(status: np.ndarray, flags: int) -> None:
    if flags & _lib.DO_VE:
        bad = np.flatnonzero(status & _lib.CLIP_VE_TOO_SHORT)
        if bad.size:
            raise ValueError(f"clips {bad.tolist()[:8]} have fewer than 201 samples after trim: the reflect-padded STFT "
                             "(melspec.py:57-64) cannot be formed")
    if flags & _lib.DO_XV:
        bad = np.flatnonzero(status & _lib.CLIP_XV_TOO_SHORT)
        if bad.size:
            raise AssertionError(f"choose a window size 400 that is [2, len]: clips {bad.tolist()[:8]} are shorter than one "
                                 "Kaldi frame (torchaudio kaldi.py:142-144)")


class WeightSync:
    """Mixin for the drop-in modules: push the module's tensors into the libcbx context when they change, without walking the
    module tree on every call (``state_dict()`` of CAMPPlus is ~940 tensors: 2-3 ms of Python per call, more than the GPU
    work of a single clip).  The tensor list is cached and dropped whenever the module is moved / cast (``_apply``) or loaded
    (``load_state_dict``); in-place edits are seen through the tensors' version counters."""

    def _cbx_tensors(self):
        cache = self.__dict__.get("_cbx_cache")
        if cache is None:
            cache = [(k, v) for k, v in self.state_dict().items() if self._cbx_wants(k)]
            self.__dict__["_cbx_cache"] = cache
        return cache

    def _cbx_fingerprint(self, dev):
        ts = self._cbx_tensors()
        # every tensor's storage address and version: `p.data = ...` / `set_()` of any tensor is seen, not only in-place edits
        return (dev, id(self), len(ts), hash(tuple((v.data_ptr(), v._version) for _, v in ts)))

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_cbx_cache", None)
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.__dict__.pop("_cbx_cache", None)
        return super().load_state_dict(*args, **kwargs)

    def _cbx_sync(self, which: int, slot: str):
        dev = device_index(self.device)
        ctx = _lib.context(dev)
        key = self._cbx_fingerprint(dev)
        if ctx.__dict__.get(slot) != key:
            ctx.load_weights(which, {k: v.detach().float().cpu().numpy() for k, v in self._cbx_tensors()})
            ctx.__dict__[slot] = key
        return ctx
