"""Oracle (TEST INFRASTRUCTURE): seeded weight sets with the reference's
``state_dict`` key names (SURVEY.md section 8b "Weights", section 8d "Weight sets").

The tensors are a pure function of the seed (CPU ``torch.Generator``), so the
GPU box regenerates exactly what ``make_golden.py`` fed to the verbatim
reference modules in the build container.

W0  default-init-like: same distributions as the reference constructors
    (``voice_encoder.py:126-133`` nn.LSTM / nn.Linear defaults; ``xvector.py:407-411``
    kaiming_normal_ on every Conv1d, zero biases; default Conv2d init in the FCM
    head; BatchNorm identity).
W1  W0 with every BatchNorm's affine and running stats randomised, so that BN
    folding / prologue paths cannot hide behind an identity BN.
W2  sensitised: VoiceEncoder LSTM ~U(-0.15,0.15) with weight_ih_l0 x30; CAMPPlus
    BN running stats calibrated on a synthetic batch (output rms ~1, the scale
    of the real fixture ``audio_test/reference_voice_clone.npy``).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import nets


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _uniform(g, shape, bound):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound


def _normal(g, shape, std):
    return torch.randn(shape, generator=g, dtype=torch.float32) * std


# ----------------------------------------------------------------------------
def ve_state_dict(kind="W0", seed=0):
    g = _gen(1000 + seed)
    sd = {}
    bound = 0.15 if kind == "W2" else 1.0 / math.sqrt(256)
    for l in range(3):
        k = 40 if l == 0 else 256
        sd[f"lstm.weight_ih_l{l}"] = _uniform(g, (1024, k), bound)
        sd[f"lstm.weight_hh_l{l}"] = _uniform(g, (1024, 256), bound)
        sd[f"lstm.bias_ih_l{l}"] = _uniform(g, (1024,), bound)
        sd[f"lstm.bias_hh_l{l}"] = _uniform(g, (1024,), bound)
    if kind == "W2":
        sd["lstm.weight_ih_l0"] = sd["lstm.weight_ih_l0"] * 30.0
    sd["proj.weight"] = _uniform(g, (256, 256), 1.0 / 16)
    sd["proj.bias"] = _uniform(g, (256,), 1.0 / 16)
    sd["similarity_weight"] = torch.tensor([10.0])
    sd["similarity_bias"] = torch.tensor([-5.0])
    return sd


# ----------------------------------------------------------------------------
def _bn_entries(sd, g, prefix, c, kind, affine=True):
    if kind == "W0" or kind == "W2":
        if affine:
            sd[prefix + ".weight"] = torch.ones(c)
            sd[prefix + ".bias"] = torch.zeros(c)
        sd[prefix + ".running_mean"] = torch.zeros(c)
        sd[prefix + ".running_var"] = torch.ones(c)
    else:
        if affine:
            sd[prefix + ".weight"] = torch.rand(c, generator=g) + 0.5
            sd[prefix + ".bias"] = _normal(g, (c,), 0.1)
        sd[prefix + ".running_mean"] = _normal(g, (c,), 0.1)
        sd[prefix + ".running_var"] = torch.rand(c, generator=g) + 0.5
    sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def _conv2d(g, co, ci, kh, kw):
    return _uniform(g, (co, ci, kh, kw), 1.0 / math.sqrt(ci * kh * kw))


def _conv1d(g, co, ci, k):
    return _normal(g, (co, ci, k), math.sqrt(2.0 / (ci * k)))


def campplus_state_dict(kind="W0", seed=0, calib_wavs=None):
    g = _gen(2000 + seed)
    sd = {}
    sd["head.conv1.weight"] = _conv2d(g, 32, 1, 3, 3)
    _bn_entries(sd, g, "head.bn1", 32, kind)
    for layer in ("head.layer1", "head.layer2"):
        for blk in (0, 1):
            p = f"{layer}.{blk}"
            sd[p + ".conv1.weight"] = _conv2d(g, 32, 32, 3, 3)
            _bn_entries(sd, g, p + ".bn1", 32, kind)
            sd[p + ".conv2.weight"] = _conv2d(g, 32, 32, 3, 3)
            _bn_entries(sd, g, p + ".bn2", 32, kind)
            if blk == 0:
                sd[p + ".shortcut.0.weight"] = _conv2d(g, 32, 32, 1, 1)
                _bn_entries(sd, g, p + ".shortcut.1", 32, kind)
    sd["head.conv2.weight"] = _conv2d(g, 32, 32, 3, 3)
    _bn_entries(sd, g, "head.bn2", 32, kind)

    sd["xvector.tdnn.linear.weight"] = _conv1d(g, 128, 320, 5)
    _bn_entries(sd, g, "xvector.tdnn.nonlinear.batchnorm", 128, kind)
    c = 128
    for b, (n_layers, _dil) in enumerate(nets.BLOCKS, start=1):
        for i in range(1, n_layers + 1):
            p = f"xvector.block{b}.tdnnd{i}"
            cin = c + 32 * (i - 1)
            _bn_entries(sd, g, p + ".nonlinear1.batchnorm", cin, kind)
            sd[p + ".linear1.weight"] = _conv1d(g, 128, cin, 1)
            _bn_entries(sd, g, p + ".nonlinear2.batchnorm", 128, kind)
            sd[p + ".cam_layer.linear_local.weight"] = _conv1d(g, 32, 128, 3)
            sd[p + ".cam_layer.linear1.weight"] = _conv1d(g, 64, 128, 1)
            sd[p + ".cam_layer.linear1.bias"] = torch.zeros(64) if kind == "W0" else _normal(g, (64,), 0.1)
            sd[p + ".cam_layer.linear2.weight"] = _conv1d(g, 32, 64, 1)
            sd[p + ".cam_layer.linear2.bias"] = torch.zeros(32) if kind == "W0" else _normal(g, (32,), 0.1)
        c += 32 * n_layers
        _bn_entries(sd, g, f"xvector.transit{b}.nonlinear.batchnorm", c, kind)
        sd[f"xvector.transit{b}.linear.weight"] = _conv1d(g, c // 2, c, 1)
        c //= 2
    _bn_entries(sd, g, "xvector.out_nonlinear.batchnorm", c, kind)
    sd["xvector.dense.linear.weight"] = _conv1d(g, 192, 2 * c, 1)
    _bn_entries(sd, g, "xvector.dense.nonlinear.batchnorm", 192, kind, affine=False)
    if kind == "W2":
        calibrate_campplus(sd, calib_wavs)
    return sd


def calibrate_campplus(sd, wavs=None):
    """Set every BN's running stats to the batch statistics of a synthetic batch
    (what one train-mode forward with momentum=None would store)."""
    from chatterbox_embed_b200 import synth
    from . import frontend
    if wavs is None:
        wavs = [synth.mixed(100 + i, 24000) for i in range(16)]
    feats = torch.from_numpy(np.stack([frontend.campplus_features(w) for w in wavs]))
    orig = nets._bn

    def calib_bn(sd_, prefix, x, affine=True):
        dims = [0] + list(range(2, x.dim()))
        sd_[prefix + ".running_mean"] = x.mean(dim=dims).detach().clone()
        sd_[prefix + ".running_var"] = x.var(dim=dims, unbiased=True).detach().clone()
        return orig(sd_, prefix, x, affine)

    nets._bn = calib_bn
    try:
        with torch.inference_mode():
            nets.campplus_forward(sd, feats)
    finally:
        nets._bn = orig
    for k in list(sd):
        if k.endswith("running_mean") or k.endswith("running_var"):
            sd[k] = sd[k].clone()
    return sd
