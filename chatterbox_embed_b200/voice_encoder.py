"""Drop-in for ``chatterbox.models.voice_encoder.VoiceEncoder`` (voice_encoder.py:119-274) whose compute runs in
libcbx.so on a B200.  Same constructor, same state_dict keys, same method signatures and return types."""
from __future__ import annotations

import warnings

from typing import List, Optional, Union

import numpy as np
import torch
from torch import Tensor, nn

from . import _host, _lib
from .config import VoiceEncConfig, check_baked


def get_frame_step(overlap: float, rate: Optional[float], hp: VoiceEncConfig = None) -> int:
    """voice_encoder.py:69-81"""
    assert 0 <= overlap < 1
    return _lib.frame_step(overlap, rate)


def get_num_wins(n_frames: int, step: int, min_coverage: float, hp: VoiceEncConfig = None):
    """voice_encoder.py:54-66"""
    return _lib.num_wins(n_frames, step, min_coverage)


class VoiceEncoder(_host.WeightSync, nn.Module):
    def __init__(self, hp=VoiceEncConfig()):
        super().__init__()
        check_baked(hp)
        self.hp = hp
        # parameter containers only (same state_dict keys as the reference); never called
        self.lstm = nn.LSTM(hp.num_mels, hp.ve_hidden_size, num_layers=3, batch_first=True)
        self.proj = nn.Linear(hp.ve_hidden_size, hp.speaker_embed_size)
        self.similarity_weight = nn.Parameter(torch.tensor([10.]), requires_grad=True)
        self.similarity_bias = nn.Parameter(torch.tensor([-5.]), requires_grad=True)
        self._ws = _host.Workspace()
        self._loaded_key = None

    @property
    def device(self):
        return next(self.parameters()).device

    # -- weights -> libcbx --------------------------------------------------------------------------------------
    @staticmethod
    def _cbx_wants(key: str) -> bool:
        return key.startswith(("lstm.", "proj."))

    def _ctx(self) -> _lib.Context:
        return self._cbx_sync(0, "_ve_key")

    # -- reference API ---------------------------------------------------------------------------------------------
    def forward(self, mels: torch.FloatTensor):
        """(B, 160, 40) partial mels -> (B, 256) L2-normed embeddings on the module's device."""
        ctx = self._ctx()
        assert mels.dim() == 3 and mels.shape[1] == self.hp.ve_partial_frames and mels.shape[2] == self.hp.num_mels
        if self.hp.normalized_mels and (mels.min() < 0 or mels.max() > 1):
            raise Exception(f"Mels outside [0, 1]. Min={mels.min()}, Max={mels.max()}")
        x = mels.detach().to(self.device, torch.float32).contiguous()
        n = x.shape[0]
        out = torch.empty((n, self.hp.speaker_embed_size), dtype=torch.float32, device=self.device)
        ws = self._ws.get(ctx.ve_forward_workspace_bytes(n), self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        ctx.ve_forward_partials(x.data_ptr(), n, out.data_ptr(), ws.data_ptr(), ws.numel(), stream)
        return out

    def inference(self, mels: torch.Tensor, mel_lens, overlap=0.5, rate: float = None, min_coverage=0.8, batch_size=None):
        """(B, T, 40) padded mels + lengths -> (B, 256) embeddings on CPU (voice_encoder.py:162-199)."""
        mel_lens = mel_lens.tolist() if torch.is_tensor(mel_lens) else list(mel_lens)
        step = get_frame_step(overlap, rate, self.hp)
        plans = [get_num_wins(int(l), step, min_coverage, self.hp) for l in mel_lens]
        n_partials = [p[0] for p in plans]
        mels = mels.to(self.device, torch.float32)
        # index arithmetic only: partial p of clip b covers rows [step*p, step*p+160), rows >= its length read as 0
        b_idx = torch.repeat_interleave(torch.arange(len(mel_lens)), torch.tensor(n_partials))
        p_idx = torch.cat([torch.arange(n) for n in n_partials])
        rows = (p_idx * step)[:, None] + torch.arange(self.hp.ve_partial_frames)[None, :]
        lens = torch.tensor(mel_lens)[b_idx][:, None]
        valid = (rows < lens) & (rows < mels.shape[1])
        rows_c = rows.clamp(max=mels.shape[1] - 1).to(self.device)
        parts = mels[b_idx.to(self.device)[:, None], rows_c] * valid.to(self.device)[..., None]
        partial_embeds = self.forward(parts)
        seg = torch.zeros(len(mel_lens), partial_embeds.shape[1], device=self.device)
        seg.index_add_(0, b_idx.to(self.device), partial_embeds)
        raw = seg / torch.tensor(n_partials, device=self.device, dtype=torch.float32)[:, None]
        return (raw / torch.linalg.norm(raw, dim=1, keepdim=True)).cpu()

    @staticmethod
    def utt_to_spk_embed(utt_embeds: np.ndarray):
        assert utt_embeds.ndim == 2
        utt_embeds = np.mean(utt_embeds, axis=0)
        return utt_embeds / np.linalg.norm(utt_embeds, 2)

    @staticmethod
    def voice_similarity(embeds_x: np.ndarray, embeds_y: np.ndarray):
        embeds_x = embeds_x if embeds_x.ndim == 1 else VoiceEncoder.utt_to_spk_embed(embeds_x)
        embeds_y = embeds_y if embeds_y.ndim == 1 else VoiceEncoder.utt_to_spk_embed(embeds_y)
        return embeds_x @ embeds_y

    def embeds_from_mels(self, mels: Union[Tensor, List[np.ndarray]], mel_lens=None, as_spk=False, batch_size=32, **kwargs):
        if isinstance(mels, List):
            mels = [np.asarray(mel) for mel in mels]
            assert all(m.shape[1] == mels[0].shape[1] for m in mels), "Mels aren't in (B, T, M) format"
            mel_lens = [mel.shape[0] for mel in mels]
            t_max = max(mel_lens)
            packed = torch.zeros((len(mels), t_max, mels[0].shape[1]), dtype=torch.float32)
            for i, m in enumerate(mels):
                packed[i, :len(m)] = torch.as_tensor(m, dtype=torch.float32)
            mels = packed
        with torch.inference_mode():
            utt_embeds = self.inference(mels.to(self.device), mel_lens, batch_size=batch_size, **kwargs).numpy()
        return self.utt_to_spk_embed(utt_embeds) if as_spk else utt_embeds

    def embeds_from_wavs(self, wavs: List[np.ndarray], sample_rate, as_spk=False, batch_size=32,
                         trim_top_db: Optional[float] = 20, **kwargs):
        """List of float waveforms -> (B, 256) float32 numpy (voice_encoder.py:246-274).  The whole chain (trim, mel,
        partials, LSTM, mean) runs in one libcbx call on HOST buffers.

        ``sample_rate != 16000``: the reference resamples with ``librosa.resample(res_type="kaiser_fast")``
        (voice_encoder.py:260-264) -- resampy's Kaiser-windowed sinc table, a third-party algorithm that is neither in the
        reference tree nor installed here, so it has no oracle.  DOCUMENTED SUBSTITUTE: the library's own polyphase
        windowed-sinc resampler (``cbx_resample``, torchaudio ``Resample`` defaults, bit-pinned to torchaudio by
        tests/test_oracle.py) -- the resampler the reference itself uses in front of CAMPPlus (s3gen.py:41-44).  Both are
        linear-phase low-pass interpolators; the embeddings agree to the extent the pass bands do (not bit-level parity,
        and not covered by the parity claim).  A warning says so once.  Integer sample rates only."""
        if sample_rate != self.hp.sample_rate:
            from .resample import Resample
            if not getattr(VoiceEncoder, "_warned_resample", False):
                warnings.warn("VoiceEncoder.embeds_from_wavs: resampling with the torchaudio-style windowed-sinc resampler of libcbx "
                              "instead of librosa's kaiser_fast (third-party, no oracle): close, not bit-level, parity for non-16 kHz input")
                VoiceEncoder._warned_resample = True
            dev = self.device
            rs = Resample(int(sample_rate), self.hp.sample_rate)
            outs = rs.ragged([torch.as_tensor(np.asarray(w, dtype=np.float32)).to(dev) for w in wavs])
            wavs = [o.cpu().numpy() for o in outs]
        rate = kwargs.pop("rate", 1.3)          # Resemble's default value (voice_encoder.py:269-270)
        overlap = kwargs.pop("overlap", 0.5)
        min_coverage = kwargs.pop("min_coverage", 0.8)
        if kwargs:
            raise TypeError(f"unexpected arguments {sorted(kwargs)}")
        step = get_frame_step(overlap, rate, self.hp)
        ctx = self._ctx()
        flat, off = _host.flatten_host(wavs)
        flags = _lib.DO_VE | (0 if trim_top_db else _lib.NO_TRIM)
        ve, _, status = ctx.embed_host(flat, off, float(trim_top_db or 0.0), step, float(min_coverage), flags)
        _host.raise_for_status(status, flags)
        return self.utt_to_spk_embed(ve) if as_spk else ve
