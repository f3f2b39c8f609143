"""The first layer each generator applies to a speaker embedding, so a voice bank can store generator-ready conditioning:

* ``spkr_enc`` -- ``T3CondEnc.spkr_enc`` Linear(256 -> 1024) on the VoiceEncoder embedding (t3/modules/cond_enc.py:50,70);
* ``spk_embed_affine_layer`` -- ``F.normalize`` + Linear(192 -> 80) on the CAMPPlus x-vector (s3gen/flow.py:73,252-253).

Parameter names match the reference modules, so ``load_state_dict`` takes the ``cond_enc.spkr_enc.*`` tensors of
``t3_cfg.safetensors`` and the ``flow.spk_embed_affine_layer.*`` tensors of ``s3gen.safetensors`` with their prefixes stripped."""
from __future__ import annotations

import torch
from torch import nn

from . import _host, _lib


class SpeakerProjections(nn.Module):
    def __init__(self, speaker_embed_size: int = 256, n_channels: int = 1024, spk_embed_dim: int = 192, output_size: int = 80):
        super().__init__()
        self.spkr_enc = nn.Linear(speaker_embed_size, n_channels)
        self.spk_embed_affine_layer = nn.Linear(spk_embed_dim, output_size)

    def _run(self, x: torch.Tensor, lin: nn.Linear, normalize: bool) -> torch.Tensor:
        dev = lin.weight.device
        ctx = _lib.context(_host.device_index(dev))
        x = x.detach().to(dev, torch.float32).reshape(-1, lin.in_features).contiguous()
        y = torch.empty((x.shape[0], lin.out_features), dtype=torch.float32, device=dev)
        w = lin.weight.detach().contiguous()
        b = lin.bias.detach().contiguous() if lin.bias is not None else None
        ctx.project(x.data_ptr(), x.shape[0], lin.in_features, w.data_ptr(), b.data_ptr() if b is not None else None, lin.out_features,
                    normalize, y.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        return y

    @torch.inference_mode()
    def t3_speaker_cond(self, speaker_emb: torch.Tensor) -> torch.Tensor:
        """cond_enc.py:70: ``spkr_enc(speaker_emb.view(-1, 256))[:, None]`` -> (B, 1, 1024)."""
        return self._run(speaker_emb, self.spkr_enc, False)[:, None]

    @torch.inference_mode()
    def flow_speaker_cond(self, embedding: torch.Tensor) -> torch.Tensor:
        """flow.py:252-253: ``spk_embed_affine_layer(F.normalize(embedding, dim=1))`` -> (B, 80)."""
        return self._run(embedding, self.spk_embed_affine_layer, True)
