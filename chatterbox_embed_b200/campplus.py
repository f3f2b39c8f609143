"""Drop-in for ``chatterbox.models.s3gen.xvector.CAMPPlus`` (xvector.py:340-428): same constructor defaults, same
state_dict keys (937 tensors), ``inference(audio_list) -> (B,192)`` on the module's device.  All compute (Kaldi fbank,
CMN, FCM head, D-TDNN blocks with CAM, statistics pooling, dense) runs in libcbx.so.

Ragged batches: the reference zero-pads features and ignores lengths (xvector.py:56, 425-428), so its result for a
short clip depends on its batch mates; every in-repo caller passes one clip.  Here each clip's embedding equals its
own B=1 result regardless of what else is in the batch."""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple, Union

import numpy as np
import torch
from torch import nn

from . import _host, _lib

_BLOCKS = ((12, 3, 1), (24, 3, 2), (16, 3, 2))   # (layers, kernel, dilation)  xvector.py:376-378


def _spec() -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) for every state_dict entry; kind in {conv2d, conv1d, bias, bn_w, bn_b, bn_mean, bn_var, bn_n}."""
    out: List[Tuple[str, Tuple[int, ...], str]] = []

    def bn(prefix, c, affine=True):
        if affine:
            out.append((prefix + ".weight", (c,), "bn_w"))
            out.append((prefix + ".bias", (c,), "bn_b"))
        out.append((prefix + ".running_mean", (c,), "bn_mean"))
        out.append((prefix + ".running_var", (c,), "bn_var"))
        out.append((prefix + ".num_batches_tracked", (), "bn_n"))

    out.append(("head.conv1.weight", (32, 1, 3, 3), "conv2d"))
    bn("head.bn1", 32)
    for layer in ("head.layer1", "head.layer2"):
        for blk in (0, 1):
            p = f"{layer}.{blk}"
            out.append((p + ".conv1.weight", (32, 32, 3, 3), "conv2d"))
            bn(p + ".bn1", 32)
            out.append((p + ".conv2.weight", (32, 32, 3, 3), "conv2d"))
            bn(p + ".bn2", 32)
            if blk == 0:
                out.append((p + ".shortcut.0.weight", (32, 32, 1, 1), "conv2d"))
                bn(p + ".shortcut.1", 32)
    out.append(("head.conv2.weight", (32, 32, 3, 3), "conv2d"))
    bn("head.bn2", 32)
    out.append(("xvector.tdnn.linear.weight", (128, 320, 5), "conv1d"))
    bn("xvector.tdnn.nonlinear.batchnorm", 128)
    c = 128
    for b, (n_layers, k, _d) in enumerate(_BLOCKS, start=1):
        for i in range(1, n_layers + 1):
            p = f"xvector.block{b}.tdnnd{i}"
            cin = c + 32 * (i - 1)
            bn(p + ".nonlinear1.batchnorm", cin)
            out.append((p + ".linear1.weight", (128, cin, 1), "conv1d"))
            bn(p + ".nonlinear2.batchnorm", 128)
            out.append((p + ".cam_layer.linear_local.weight", (32, 128, k), "conv1d"))
            out.append((p + ".cam_layer.linear1.weight", (64, 128, 1), "conv1d"))
            out.append((p + ".cam_layer.linear1.bias", (64,), "bias"))
            out.append((p + ".cam_layer.linear2.weight", (32, 64, 1), "conv1d"))
            out.append((p + ".cam_layer.linear2.bias", (32,), "bias"))
        c += 32 * n_layers
        bn(f"xvector.transit{b}.nonlinear.batchnorm", c)
        out.append((f"xvector.transit{b}.linear.weight", (c // 2, c, 1), "conv1d"))
        c //= 2
    bn("xvector.out_nonlinear.batchnorm", c)
    out.append(("xvector.dense.linear.weight", (192, 2 * c, 1), "conv1d"))
    bn("xvector.dense.nonlinear.batchnorm", 192, affine=False)
    return out


class _Node(nn.Module):
    """Parameter container; children are created on demand so dotted keys map onto a module tree."""

    def child(self, name: str) -> "_Node":
        if name not in self._modules:
            self.add_module(name, _Node())
        return self._modules[name]


def _init(shape, kind):
    if kind == "conv2d":
        bound = 1.0 / math.sqrt(shape[1] * shape[2] * shape[3])
        return torch.empty(shape).uniform_(-bound, bound)
    if kind == "conv1d":                       # kaiming_normal_ (xvector.py:407-411)
        return torch.randn(shape) * math.sqrt(2.0 / (shape[1] * shape[2]))
    if kind in ("bias", "bn_b", "bn_mean"):
        return torch.zeros(shape)
    if kind in ("bn_w", "bn_var"):
        return torch.ones(shape)
    return torch.tensor(0, dtype=torch.long)


class CAMPPlus(_host.WeightSync, nn.Module):
    def __init__(self, feat_dim=80, embedding_size=192, growth_rate=32, bn_size=4, init_channels=128,
                 config_str="batchnorm-relu", memory_efficient=True, output_level="segment", **kwargs):
        super().__init__()
        baked = dict(feat_dim=80, embedding_size=192, growth_rate=32, bn_size=4, init_channels=128,
                     config_str="batchnorm-relu", output_level="segment")
        given = dict(feat_dim=feat_dim, embedding_size=embedding_size, growth_rate=growth_rate, bn_size=bn_size,
                     init_channels=init_channels, config_str=config_str, output_level=output_level)
        for k, v in baked.items():
            if given[k] != v:
                raise ValueError(f"CAMPPlus({k}={given[k]!r}) differs from the configuration the sm_100a kernels are built for ({v!r})")
        self.output_level = output_level
        for key, shape, kind in _spec():
            *path, leaf = key.split(".")
            node = self
            for part in path:
                node = node.child(part) if isinstance(node, _Node) else CAMPPlus._child(node, part)
            t = _init(shape, kind)
            if kind in ("bn_mean", "bn_var", "bn_n"):
                node.register_buffer(leaf, t)
            else:
                node.register_parameter(leaf, nn.Parameter(t))
        self._ws = _host.Workspace()

    @staticmethod
    def _child(mod: nn.Module, name: str) -> _Node:
        if name not in mod._modules:
            mod.add_module(name, _Node())
        return mod._modules[name]

    @property
    def device(self):
        return next(self.parameters()).device

    @staticmethod
    def _cbx_wants(key: str) -> bool:
        return not key.endswith("num_batches_tracked")

    def _ctx(self) -> _lib.Context:
        return self._cbx_sync(1, "_xv_key")

    def _clips(self, audio_list) -> List[torch.Tensor]:
        if torch.is_tensor(audio_list):
            assert audio_list.dim() == 2, "expected (B, L) waveforms"
            return [row for row in audio_list]
        return [torch.as_tensor(a).reshape(-1) for a in audio_list]

    @torch.inference_mode()
    def inference(self, audio_list):
        """List of (L_i,) 16 kHz waveforms or a (B, L) tensor -> (B, 192) float32 on the module's device."""
        ctx = self._ctx()
        dev = self.device
        clips = [c.to(dev, torch.float32) for c in self._clips(audio_list)]
        lens = [int(c.numel()) for c in clips]
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        pcm = torch.cat(clips) if len(clips) > 1 else clips[0].contiguous()
        n = len(clips)
        out = torch.empty((n, 192), dtype=torch.float32, device=dev)
        status = torch.zeros(n, dtype=torch.int32, device=dev)
        flags = _lib.DO_XV
        ws = self._ws.get(ctx.workspace_bytes(lens, 77, 0.8, flags), dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        ctx.embed(pcm.data_ptr(), off, 0.0, 77, 0.8, 0, out.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), stream, flags)
        _host.raise_for_status(status.cpu().numpy(), flags)
        return out

    @torch.inference_mode()
    def forward(self, x):
        """(B, T, 80) features (what extract_feature returns: log-fbank, mean-normalised) -> (B, 192)  (xvector.py:417-423).
        A list of (T_i, 80) tensors is taken as a ragged batch (each clip pooled over its own frames)."""
        ctx = self._ctx()
        dev = self.device
        if torch.is_tensor(x):
            assert x.dim() == 3 and x.shape[2] == 80, "expected (B, T, 80) features"
            feats = x.to(dev, torch.float32).contiguous()
            frames = [int(x.shape[1])] * int(x.shape[0])
        else:
            rows = [torch.as_tensor(f).to(dev, torch.float32).reshape(-1, 80) for f in x]
            frames = [int(r.shape[0]) for r in rows]
            feats = torch.cat(rows).contiguous()
        n = len(frames)
        assert n > 0 and min(frames) > 0, "expected at least one frame per clip"
        off = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
        out = torch.empty((n, 192), dtype=torch.float32, device=dev)
        status = torch.zeros(n, dtype=torch.int32, device=dev)
        ws = self._ws.get(ctx.campplus_forward_workspace_bytes(off), dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        ctx.campplus_forward_feats(feats.data_ptr(), off, out.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), stream)
        return out
