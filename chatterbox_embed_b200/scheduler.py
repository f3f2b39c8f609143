"""Ragged-length batch scheduler: shards clips over the GPUs of one box and returns all embeddings on every rank
with ONE all-gather of a (N_pad/R, 448) fp32 block (256 VoiceEncoder | 192 CAMPPlus).  Clips are independent in both
encoders (CMN, CAM means, stats pooling and partial means are all per clip), so there is no collective inside the path
(SURVEY.md section 8e)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _host, _lib
from .campplus import CAMPPlus
from .voice_encoder import VoiceEncoder

EMB = 256 + 192


def partition(lengths: Sequence[int], world: int) -> List[np.ndarray]:
    """Deal clips to ranks so that every rank gets an equal (+-1) clip count and near-equal work under the FLOP cost
    model c(L) (cbx_clip_cost): sorted by cost, dealt out and back over the ranks -- natively, in cbx_partition
    (1e5 clips over 8 ranks: ~2 ms with the radix sort; the Python loop it replaces took 240 ms).  Returns the clip indices of every rank, ascending."""
    rank_of, _, _ = _lib.partition(lengths, world)
    return [np.flatnonzero(rank_of == r).astype(np.int64) for r in range(world)]


def inverse_permutation(shards: Sequence[np.ndarray], n: int) -> np.ndarray:
    """Row of clip i inside the rank-major, per-rank padded gathered array."""
    per = max(len(s) for s in shards)
    inv = np.full(n, -1, dtype=np.int64)
    for r, s in enumerate(shards):
        inv[s] = r * per + np.arange(len(s))
    assert (inv >= 0).all()
    return inv


class SpeakerEmbedder:
    """Both encoders over one PCM upload.  ``ve`` / ``cp`` are the drop-in modules (they own the weights)."""

    def __init__(self, ve: VoiceEncoder, cp: CAMPPlus):
        assert ve.device == cp.device
        self.ve, self.cp = ve, cp
        self.device = ve.device
        self._ws = _host.Workspace()
        self._ctx = None

    def ctx(self) -> _lib.Context:
        self.ve._ctx()
        return self.cp._ctx()

    def embed_host(self, flat: np.ndarray, offsets: np.ndarray, trim_top_db: Optional[float] = 20.0, step: int = 77,
                   min_coverage: float = 0.8) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """HOST buffers in, HOST arrays out; host<->device copies happen inside libcbx (cbx_embed_host)."""
        flags = _lib.DO_VE | _lib.DO_XV | (0 if trim_top_db else _lib.NO_TRIM)
        return self.ctx().embed_host(flat, offsets, float(trim_top_db or 0.0), step, min_coverage, flags)

    def embed_stream(self, batches, trim_top_db: Optional[float] = 20.0, step: int = 77, min_coverage: float = 0.8, pinned: bool = False):
        """Voice-bank extraction over many batches: ``batches`` yields (flat_pcm, offsets); yields (ve, xv, status) per
        batch, in order.  Two batches are in flight (cbx_embed_host_submit / _wait), so the host<->device copies of one
        overlap the kernels of the other."""
        ctx = self.ctx()
        flags = _lib.DO_VE | _lib.DO_XV | (0 if trim_top_db else _lib.NO_TRIM) | (_lib.PCM_PINNED if pinned else 0)
        pending = []
        for k, (flat, offsets) in enumerate(batches):
            ctx.embed_host_submit(k & 1, flat, offsets, float(trim_top_db or 0.0), step, min_coverage, flags)
            pending.append(k & 1)
            if len(pending) == 2:
                yield ctx.embed_host_wait(pending.pop(0))
        while pending:
            yield ctx.embed_host_wait(pending.pop(0))

    def embed_wavs(self, wavs: Sequence[np.ndarray], **kw):
        flat, off = _host.flatten_host(wavs)
        ve, xv, status = self.embed_host(flat, off, **kw)
        _host.raise_for_status(status, _lib.DO_VE | _lib.DO_XV)
        return ve, xv

    def embed_device(self, pcm: torch.Tensor, offsets: np.ndarray, out: Optional[torch.Tensor] = None,
                     trim_top_db: Optional[float] = 20.0, step: int = 77, min_coverage: float = 0.8):
        """Device-resident PCM (flat fp32) -> (n, 448) device tensor [VE | XV] and an int32 status vector.
        Stream-ordered on torch's current stream, no synchronisation."""
        ctx = self.ctx()
        n = len(offsets) - 1
        lens = np.diff(offsets)
        flags = _lib.DO_VE | _lib.DO_XV | (0 if trim_top_db else _lib.NO_TRIM)
        if out is None:
            out = torch.empty((2, n, 256), dtype=torch.float32, device=self.device)   # [0]=VE, [1][:, :192]=XV
        status = torch.empty(n, dtype=torch.int32, device=self.device)
        ws = self._ws.get(ctx.workspace_bytes(lens, step, min_coverage, flags), self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        ve_out = out[0]
        xv_out = out[1].view(-1)[: n * 192].view(n, 192)
        ctx.embed(pcm.data_ptr(), offsets, float(trim_top_db or 0.0), step, min_coverage, ve_out.data_ptr(),
                  xv_out.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), stream, flags)
        return ve_out, xv_out, status


def batch_bounds(lengths: Sequence[int], max_clips: int = 256, max_samples: int = 256 * 160000) -> List[Tuple[int, int]]:
    """Cut a shard (clip lengths in processing order) into consecutive batches of at most ``max_clips`` clips and
    ``max_samples`` PCM samples (a single longer clip is a batch of its own): the unit of one cbx_embed_host_submit."""
    out, i, n = [], 0, len(lengths)
    while i < n:
        j, acc = i, 0
        while j < n and j - i < max_clips and (j == i or acc + int(lengths[j]) <= max_samples):
            acc += int(lengths[j]); j += 1
        out.append((i, j))
        i = j
    return out


def embed_shard(emb: "SpeakerEmbedder", fetch, lengths: Sequence[int], max_clips: int = 256, max_samples: int = 256 * 160000,
                pinned: bool = True, out: Optional[np.ndarray] = None, **kw) -> Tuple[np.ndarray, np.ndarray]:
    """Voice-bank extraction of one rank's shard (BASELINE config 4): ``fetch(b, i0, i1)`` returns the flat host PCM of
    clips [i0, i1) of the shard (batch number b); batches stream through ``embed_stream`` (two in flight).  Returns the
    (n, 448) float32 block [VE | XV] in shard order (written into ``out`` if given, e.g. the numpy view of a pinned tensor that
    is then handed to the all-gather) and the per-clip status words."""
    lengths = np.asarray(lengths, dtype=np.int64)
    bounds = batch_bounds(lengths, max_clips, max_samples)
    local = out if out is not None else np.empty((len(lengths), EMB), dtype=np.float32)
    assert local.shape == (len(lengths), EMB) and local.dtype == np.float32
    status = np.empty(len(lengths), dtype=np.int32)

    def gen():
        for b, (i0, i1) in enumerate(bounds):
            off = np.concatenate([[0], np.cumsum(lengths[i0:i1])]).astype(np.int64)
            yield fetch(b, i0, i1), off

    for (i0, i1), (ve, xv, st) in zip(bounds, emb.embed_stream(gen(), pinned=pinned, **kw)):
        local[i0:i1, :256] = ve
        local[i0:i1, 256:] = xv
        status[i0:i1] = st
    return local, status


def gather_embeddings(local: torch.Tensor, shards: Sequence[np.ndarray], n_total: int, group=None) -> torch.Tensor:
    """local: (len(shards[rank]), 448) on this rank's device (or CPU with gloo).  Returns (n_total, 448) in the
    original clip order on every rank, via one all_gather_into_tensor of the padded per-rank block."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    per = max(len(s) for s in shards)
    block = torch.zeros((per, local.shape[1]), dtype=local.dtype, device=local.device)
    block[: local.shape[0]] = local
    full = torch.empty((world * per, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full, block, group=group)
    inv = torch.as_tensor(inverse_permutation(shards, n_total), device=local.device)
    return full[inv]
