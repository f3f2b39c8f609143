"""What a bf16 mode (BASELINE config 5) would cost in accuracy, family by family, before a kernel is written (CPU only).

The product path computes in TF32 with fp32 storage.  A bf16 mode would keep the fp32 accumulators and store activations
(and weights) in bf16, which halves the HBM traffic of exactly the HBM-bound kernels (FCM convolutions, the D-TDNN
bottleneck GEMM over the growing concatenation, the LSTM input projections).  With tcgen05 kind::f16 BOTH operands of an
MMA are bf16, so the emulation rounds the input and the weight of every convolution / matmul of a family to bf16 and
accumulates in fp32 (torch CPU); everything else stays fp32.  TF32 rounding (10-bit mantissa, what the product path does
today) is reported beside it as the yardstick.

    python tests/tools/bf16_study.py [n_clips] [clip_seconds]
    python tests/tools/bf16_study.py settings      # the two built settings of the cat_bf16 option, emulated on the clips of
                                                   # cat_bf16_check.py (prediction beside the GPU measurement)

Output: per weight set (W0 default init, W1 randomised BN, W2 sensitised / calibrated) and per family set, the max-abs
error and minimum cosine of the 192-d x-vector and of the 256-d VoiceEncoder embedding against the fp32 oracle.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.nn.functional as F

from chatterbox_embed_b200 import synth
from oracle import frontend, nets, weights


def to_bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def to_tf32(x):
    """Round to nearest even on the 13 dropped mantissa bits (what cvt.rna.tf32 / the tensor core's operand read does
    up to the tie rule)."""
    i = x.contiguous().view(torch.int32)
    i = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32)


class Rounding:
    """Patches F.conv1d / F.conv2d as seen by oracle.nets; the family of a call is decided from the weight's shape."""

    def __init__(self, families, rnd, rnd_w=None, others=None):
        """families: rounded with ``rnd`` (activations) / ``rnd_w`` (weights; default = rnd); every other family with
        ``others`` (both operands) if given."""
        self.families, self.rnd, self.rnd_w, self.others = set(families), rnd, rnd_w or rnd, others

    @staticmethod
    def family(w):
        s = tuple(w.shape)
        if len(s) == 4:
            return "fcm"
        co, ci, k = s
        if k == 5:
            return "tdnn"
        if k == 3:
            return "local"
        if co == 128 and k == 1:
            return "bottleneck"
        if co in (64, 32) and ci in (128, 64):
            return "cam"
        if co == 192:
            return "dense"
        return "transit"

    def __enter__(self):
        self.c1, self.c2 = F.conv1d, F.conv2d
        me = self

        def operands(x, w):
            if me.family(w) in me.families:
                return me.rnd(x), me.rnd_w(w)
            return (me.others(x), me.others(w)) if me.others else (x, w)

        def conv1d(x, w, *a, **k):
            return me.c1(*operands(x, w), *a, **k)

        def conv2d(x, w, *a, **k):
            return me.c2(*operands(x, w), *a, **k)

        nets.F.conv1d, nets.F.conv2d = conv1d, conv2d
        return self

    def __exit__(self, *exc):
        nets.F.conv1d, nets.F.conv2d = self.c1, self.c2


def ve_rounded(sd, wavs, rnd, which):
    """VoiceEncoder with the operands of the chosen matmuls rounded: 'xw' = input projections (one dense GEMM per layer in
    the product path), 'hh' = the recurrent product, 'state' = h stored rounded between steps and layers."""
    out = []
    for w in wavs:
        mel = torch.from_numpy(frontend.ve_melspectrogram(w))
        parts = torch.from_numpy(nets.ve_partials(mel.numpy())) if not torch.is_tensor(nets.ve_partials(mel.numpy())) else nets.ve_partials(mel.numpy())
        x = parts.float()
        for layer in range(3):
            w_ih, w_hh = sd[f"lstm.weight_ih_l{layer}"], sd[f"lstm.weight_hh_l{layer}"]
            b = sd[f"lstm.bias_ih_l{layer}"] + sd[f"lstm.bias_hh_l{layer}"]
            xi, wi = (rnd(x), rnd(w_ih)) if "xw" in which else (x, w_ih)
            xw = xi @ wi.T + b
            wh = rnd(w_hh) if "hh" in which else w_hh
            h = torch.zeros(x.shape[0], 256)
            c = torch.zeros(x.shape[0], 256)
            hs = []
            for t in range(x.shape[1]):
                hh = rnd(h) if "hh" in which else h
                g = xw[:, t] + hh @ wh.T
                i, f, gg, o = g.chunk(4, dim=1)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
                h = torch.sigmoid(o) * torch.tanh(c)
                hs.append(h)
            x = torch.stack(hs, dim=1)
        e = F.relu(x[:, -1] @ sd["proj.weight"].T + sd["proj.bias"])
        e = e / e.norm(dim=1, keepdim=True)
        m = e.mean(0)
        out.append((m / m.norm()).numpy())
    return np.stack(out)


def report(name, got, want):
    err = float(np.abs(got - want).max())
    cos = min(float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b))) for a, b in zip(got, want))
    print(f"  {name:<44s} max-abs {err:9.2e}   min cos {cos:.6f}", flush=True)


def settings():
    """cat_bf16 = 1: the bottleneck / transit GEMMs read bf16 activations (then BN + ReLU, tf32 stage), weights tf32;
    cat_bf16 = 2: their operands are bf16.  Everything else TF32, as on the GPU.  The BN + ReLU between the stored activation
    and the MMA operand is not re-rounded here (setting 1 rounds the stored value; the tf32 rounding after BN is implied by
    rounding the convolution input), which is what the kernels do up to the order of two roundings."""
    wavs = [synth.mixed(i, n) for i, n in enumerate((48000, 25600, 64000, 16000 * 7 + 123))]
    fam = ("bottleneck", "transit")
    with torch.inference_mode():
        for kind in ("W0", "W1", "W2"):
            sdc = weights.campplus_state_dict(kind)
            want = nets.campplus_embed_wavs(sdc, wavs)
            print(f"CAMPPlus {kind}: |x-vector| max {np.abs(want).max():.3f}")
            with Rounding((), to_tf32, others=to_tf32):
                report("tf32 everywhere (cat_bf16 = 0)", nets.campplus_embed_wavs(sdc, wavs), want)
            with Rounding(fam, to_bf16, rnd_w=to_tf32, others=to_tf32):
                report("cat_bf16 = 1 (bf16 activations into those GEMMs)", nets.campplus_embed_wavs(sdc, wavs), want)
            with Rounding(fam, to_bf16, others=to_tf32):
                report("cat_bf16 = 2 (bf16 operands)", nets.campplus_embed_wavs(sdc, wavs), want)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "settings":
        return settings()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
    torch.set_num_threads(os.cpu_count() or 1)
    wavs = [synth.mixed(i, int(secs * 16000)) for i in range(n)]
    cam_sets = [("fcm",), ("bottleneck",), ("fcm", "bottleneck"), ("fcm", "bottleneck", "tdnn", "transit"),
                ("fcm", "tdnn", "bottleneck", "local", "cam", "transit", "dense")]
    with torch.inference_mode():
        for kind in ("W0", "W1", "W2"):
            sdc = weights.campplus_state_dict(kind)
            want = nets.campplus_embed_wavs(sdc, wavs)
            print(f"CAMPPlus {kind}: |x-vector| max {np.abs(want).max():.3f}, rms {np.sqrt((want ** 2).mean()):.3f}")
            with Rounding(cam_sets[-1], to_tf32):
                report("tf32 everywhere (today's mode)", nets.campplus_embed_wavs(sdc, wavs), want)
            for fams in cam_sets:
                with Rounding(fams, to_bf16):
                    report("bf16: " + "+".join(fams) if len(fams) < 7 else "bf16 everywhere", nets.campplus_embed_wavs(sdc, wavs), want)
        for kind in ("W0", "W2"):
            sdv = weights.ve_state_dict(kind)
            want = ve_rounded(sdv, wavs, lambda x: x, ())
            ref = nets.ve_embed_wavs(sdv, wavs, trim_top_db=None)
            print(f"VoiceEncoder {kind}: restated loop vs oracle max-abs {np.abs(want - ref).max():.1e}")
            report("tf32 xw+hh (today's mode)", ve_rounded(sdv, wavs, to_tf32, ("xw", "hh")), want)
            report("bf16 xw (input projections)", ve_rounded(sdv, wavs, to_bf16, ("xw",)), want)
            report("bf16 xw+hh", ve_rounded(sdv, wavs, to_bf16, ("xw", "hh")), want)


if __name__ == "__main__":
    main()
