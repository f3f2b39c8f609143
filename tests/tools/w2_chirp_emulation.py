"""CPU emulation of the tensor-core mode on the golden clips with W2 (the "W2 x pure chirp" case):
the fp32 oracle with the operands of every convolution rounded to TF32 (fp32 accumulation), stage by stage against the
unrounded oracle.  If the emulation shows the same per-stage error growth as the B200 (tests/tools/stage_report.py), the
loss of cosine is operand rounding amplified by the network, not a kernel fault.
    python tests/tools/w2_chirp_emulation.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
from oracle import make_golden, nets, weights, frontend
from bf16_study import Rounding, to_tf32

torch.set_num_threads(os.cpu_count() or 1)
wavs = make_golden.golden_wavs()
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
with torch.inference_mode():
    for kind in ("W1", "W2"):
        sdc = weights.campplus_state_dict(kind)
        for ci, w in enumerate(wavs):
            f = torch.from_numpy(frontend.campplus_features(w))[None]
            t0, t1 = {}, {}
            want = nets.campplus_forward(sdc, f, t0)[0]
            with Rounding((), to_tf32, others=to_tf32):
                got = nets.campplus_forward(sdc, f, t1)[0]
            # an input perturbation of the size of the fbank floor-bin noise instead (1e-3 in the log domain), fp32 arithmetic
            g = torch.Generator().manual_seed(ci)
            t2 = {}
            pert = nets.campplus_forward(sdc, f + 1e-3 * torch.randn(f.shape, generator=g), t2)[0]
            cs = lambda a, b: float(a.double() @ b.double() / a.double().norm() / b.double().norm())
            print(f"{kind} clip {ci} ({'chirp' if ci % 2 else 'noise'}): max|x|={float(want.abs().max()):.1f}  tf32-emulated: "
                  + " ".join(f"{k}={rel(t1[k], t0[k]):.2e}" for k in ("fcm", "tdnn", "block1", "block2", "block3", "transit3", "stats"))
                  + f" xv_rel={rel(got, want):.2e} cos={cs(got, want):.6f} | fp32 with 1e-3 input noise: block3={rel(t2['block3'], t0['block3']):.2e} cos={cs(pert, want):.6f}", flush=True)
