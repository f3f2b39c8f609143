"""Oracle (TEST INFRASTRUCTURE): CPU restatement of the two speaker encoders,
written against a plain ``state_dict`` with the reference's key names.

VoiceEncoder: ``voice_encoder.py:54-81`` (window arithmetic), ``:139-160``
(LSTM + proj + ReLU + L2), ``:162-199`` (partials, per-clip mean, L2).
CAMPPlus: ``xvector.py:61-127`` (FCM head), ``:160-337`` (TDNN / CAM dense
layers / transit / dense), ``:146-152`` (statistics pooling), ``:417-423``.

Pinned against the verbatim reference modules by tests/test_oracle.py (when
/root/reference is present) and through tests/golden/*.npz everywhere else.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import frontend

PARTIAL_FRAMES = 160
BN_EPS = 1e-5


# ----------------------------------------------------------------------------
# integer window arithmetic  (voice_encoder.py:54-81)
# ----------------------------------------------------------------------------
def frame_step(overlap=0.5, rate=1.3, sample_rate=16000, partial=PARTIAL_FRAMES):
    if rate is None:
        step = int(np.round(partial * (1 - overlap)))
    else:
        step = int(np.round((sample_rate / rate) / partial))
    assert 0 < step <= partial
    return step


def num_wins(n_frames, step=77, min_coverage=0.8, win=PARTIAL_FRAMES):
    assert n_frames > 0
    n, rem = divmod(max(n_frames - win + step, 0), step)
    if n == 0 or (rem + (win - step)) / win >= min_coverage:
        n += 1
    return n, win + step * (n - 1)


# ----------------------------------------------------------------------------
# VoiceEncoder
# ----------------------------------------------------------------------------
def lstm_layer(x, w_ih, w_hh, b_ih, b_hh, return_seq=True):
    """x (N,T,I) -> h sequence (N,T,H); PyTorch gate order i,f,g,o; h0=c0=0."""
    n, t, _ = x.shape
    hdim = w_hh.shape[1]
    h = x.new_zeros(n, hdim)
    c = x.new_zeros(n, hdim)
    xin = x @ w_ih.T + (b_ih + b_hh)
    out = []
    for s in range(t):
        g = xin[:, s] + h @ w_hh.T
        i, f, gg, o = g.split(hdim, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out.append(h)
    return torch.stack(out, dim=1)


def ve_forward(sd, partials, return_layers=False):
    """(N,160,40) -> (N,256) L2-normed partial embeddings (voice_encoder.py:139-160)."""
    x = torch.as_tensor(partials, dtype=torch.float32)
    layers = []
    for l in range(3):
        x = lstm_layer(x, sd[f"lstm.weight_ih_l{l}"], sd[f"lstm.weight_hh_l{l}"],
                       sd[f"lstm.bias_ih_l{l}"], sd[f"lstm.bias_hh_l{l}"])
        layers.append(x[:, -1])
    raw = F.relu(x[:, -1] @ sd["proj.weight"].T + sd["proj.bias"])
    emb = raw / torch.linalg.norm(raw, dim=1, keepdim=True)
    return (emb, layers) if return_layers else emb


def ve_partials(mel, step=77, min_coverage=0.8):
    """(T,40) -> (P,160,40) with zero rows past T (voice_encoder.py:172-187)."""
    mel = np.asarray(mel, dtype=np.float32)
    p, target = num_wins(len(mel), step, min_coverage)
    if target > len(mel):
        mel = np.concatenate([mel, np.zeros((target - len(mel), mel.shape[1]), np.float32)])
    return np.stack([mel[i * step: i * step + PARTIAL_FRAMES] for i in range(p)])


def ve_embed_mels(sd, mels, rate=1.3, overlap=0.5, min_coverage=0.8):
    """list of (T_i,40) -> (B,256) float32 (voice_encoder.py:162-199)."""
    step = frame_step(overlap, rate)
    out = []
    with torch.inference_mode():
        for mel in mels:
            pe = ve_forward(sd, ve_partials(mel, step, min_coverage))
            raw = pe.mean(dim=0)
            out.append((raw / torch.linalg.norm(raw)).numpy())
    return np.stack(out).astype(np.float32)


def ve_embed_wavs(sd, wavs, trim_top_db=20, **kw):
    """voice_encoder.py:246-274 for 16 kHz input."""
    if trim_top_db:
        wavs = [frontend.effects_trim(np.asarray(w), top_db=trim_top_db)[0] for w in wavs]
    mels = [frontend.ve_melspectrogram(w) for w in wavs]
    return ve_embed_mels(sd, mels, **kw)


# ----------------------------------------------------------------------------
# CAMPPlus
# ----------------------------------------------------------------------------
def _bn(sd, prefix, x, affine=True):
    shape = [1, -1] + [1] * (x.dim() - 2)
    y = (x - sd[prefix + ".running_mean"].view(shape)) / torch.sqrt(sd[prefix + ".running_var"].view(shape) + BN_EPS)
    if affine:
        y = y * sd[prefix + ".weight"].view(shape) + sd[prefix + ".bias"].view(shape)
    return y


def _res_block(sd, p, x, stride):
    y = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], stride=(stride, 1), padding=1)))
    y = _bn(sd, p + ".bn2", F.conv2d(y, sd[p + ".conv2.weight"], padding=1))
    if (p + ".shortcut.0.weight") in sd:
        x = _bn(sd, p + ".shortcut.1", F.conv2d(x, sd[p + ".shortcut.0.weight"], stride=(stride, 1)))
    return F.relu(y + x)


def fcm_head(sd, feats):
    """(B,80,T) -> (B,320,T)  (xvector.py:94-127)."""
    x = feats.unsqueeze(1)
    x = F.relu(_bn(sd, "head.bn1", F.conv2d(x, sd["head.conv1.weight"], padding=1)))
    for layer in ("head.layer1", "head.layer2"):
        x = _res_block(sd, layer + ".0", x, 2)
        x = _res_block(sd, layer + ".1", x, 1)
    x = F.relu(_bn(sd, "head.bn2", F.conv2d(x, sd["head.conv2.weight"], stride=(2, 1), padding=1)))
    return x.reshape(x.shape[0], x.shape[1] * x.shape[2], x.shape[3])


def _seg_mean(u, seg=100):
    t = u.shape[-1]
    s = F.avg_pool1d(u, kernel_size=seg, stride=seg, ceil_mode=True)
    return s.repeat_interleave(seg, dim=-1)[..., :t]


def cam_dense_layer(sd, p, x, dilation):
    a = F.relu(_bn(sd, p + ".nonlinear1.batchnorm", x))
    h = F.conv1d(a, sd[p + ".linear1.weight"])
    u = F.relu(_bn(sd, p + ".nonlinear2.batchnorm", h))
    y = F.conv1d(u, sd[p + ".cam_layer.linear_local.weight"], padding=dilation, dilation=dilation)
    ctx = u.mean(-1, keepdim=True) + _seg_mean(u)
    ctx = F.relu(F.conv1d(ctx, sd[p + ".cam_layer.linear1.weight"], sd[p + ".cam_layer.linear1.bias"]))
    m = torch.sigmoid(F.conv1d(ctx, sd[p + ".cam_layer.linear2.weight"], sd[p + ".cam_layer.linear2.bias"]))
    return y * m


BLOCKS = ((12, 1), (24, 2), (16, 2))  # (layers, dilation)  xvector.py:376-378


def campplus_forward(sd, feats, taps=None):
    """feats (B,T,80) CMN'd fbank -> (B,192)  (xvector.py:417-423).  ``taps`` (dict)
    receives stage outputs for stage-wise parity checks."""
    x = torch.as_tensor(feats, dtype=torch.float32).permute(0, 2, 1)
    x = fcm_head(sd, x)
    if taps is not None:
        taps["fcm"] = x
    x = F.relu(_bn(sd, "xvector.tdnn.nonlinear.batchnorm",
                   F.conv1d(x, sd["xvector.tdnn.linear.weight"], stride=2, padding=2)))
    if taps is not None:
        taps["tdnn"] = x
    for b, (n_layers, dil) in enumerate(BLOCKS, start=1):
        for i in range(1, n_layers + 1):
            x = torch.cat([x, cam_dense_layer(sd, f"xvector.block{b}.tdnnd{i}", x, dil)], dim=1)
        if taps is not None:
            taps[f"block{b}"] = x
        x = F.conv1d(F.relu(_bn(sd, f"xvector.transit{b}.nonlinear.batchnorm", x)),
                     sd[f"xvector.transit{b}.linear.weight"])
        if taps is not None:
            taps[f"transit{b}"] = x
    x = F.relu(_bn(sd, "xvector.out_nonlinear.batchnorm", x))
    stats = torch.cat([x.mean(dim=-1), x.std(dim=-1, unbiased=True)], dim=-1)
    if taps is not None:
        taps["stats"] = stats
    e = F.conv1d(stats.unsqueeze(-1), sd["xvector.dense.linear.weight"]).squeeze(-1)
    return _bn(sd, "xvector.dense.nonlinear.batchnorm", e, affine=False)


def campplus_embed_wavs(sd, wavs, taps=None):
    """One clip at a time (the reference's real usage)."""
    out = []
    with torch.inference_mode():
        for w in wavs:
            f = torch.from_numpy(frontend.campplus_features(w))[None]
            out.append(campplus_forward(sd, f, taps)[0].numpy())
    return np.stack(out).astype(np.float32)
