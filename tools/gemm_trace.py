"""Per-CTA timeline of the pre-activation (bottleneck) GEMM: where do the ~47 us of K-independent time per launch go?
    python tools/gemm_trace.py            (GPU box)"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chatterbox_embed_b200 import _lib
L = _lib.lib(); ctx = _lib.context(0)
fn = L.cbx_test_tgemm
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
dev = "cuda:0"
M = 130048
tiles = (M + 127) // 128
A = torch.randn(M + 8, 1024, device=dev)
bias = torch.randn(128, device=dev); a = torch.rand(1024, device=dev) + 0.5; b = torch.randn(1024, device=dev) * 0.3
Cc = torch.empty(M, 128, device=dev)
trace = torch.zeros(tiles * 16, dtype=torch.int64, device=dev)
VAR = int(sys.argv[1]) if len(sys.argv) > 1 else 7          # 7: production epilogue, 4: plain bias epilogue
for K in (128, 512, 1024):
    W = torch.randn(128, K, device=dev) / K ** 0.5
    args = (ctx._h, A.data_ptr(), 1024, W.data_ptr(), K, Cc.data_ptr(), 128, M, 128, K, bias.data_ptr(), a.data_ptr(), b.data_ptr())
    for _ in range(3):
        fn(*args, VAR, 0, None)
    torch.cuda.synchronize()
    fn(ctx._h, None, 0, None, 0, trace.data_ptr(), 0, 0, 0, 0, None, None, None, 6, 0, None)
    rc = fn(*args, VAR, 0, None)
    assert rc == 0, L.cbx_last_error(ctx._h)
    torch.cuda.synchronize()
    fn(ctx._h, None, 0, None, 0, None, 0, 0, 0, 0, None, None, None, 6, 0, None)
    t = trace.cpu().numpy().reshape(tiles, 16).astype(np.int64)
    t0 = t[:, 1].min()
    sm = t[:, 0]
    ent, setup, first, pdone, acc, epi, end = [(t[:, i] - t0) / 1e3 for i in range(1, 8)]
    print(f"K={K}: kernel span {end.max():.1f} us; CTAs {tiles}; SMs used {len(set(sm.tolist()))}")
    print(f"  setup (entry -> after sync)        median {np.median(setup - ent):6.2f}  p90 {np.quantile(setup - ent, .9):6.2f} us")
    print(f"  first MMA (after sync -> A+B full) median {np.median(first - setup):6.2f}  p90 {np.quantile(first - setup, .9):6.2f}")
    print(f"  K loop (first MMA -> producers done) median {np.median(pdone - first):6.2f}  p90 {np.quantile(pdone - first, .9):6.2f}")
    print(f"  accum wait (producers done -> accum) median {np.median(acc - pdone):6.2f}  p90 {np.quantile(acc - pdone, .9):6.2f}")
    print(f"  epilogue (accum -> last store issued) median {np.median(epi - acc):6.2f}  p90 {np.quantile(epi - acc, .9):6.2f}")
    print(f"  drain (store issued -> exit)         median {np.median(end - epi):6.2f}  p90 {np.quantile(end - epi, .9):6.2f}")
    print(f"  CTA lifetime                         median {np.median(end - ent):6.2f}  p90 {np.quantile(end - ent, .9):6.2f}")
    # group 0's two chunks (0 and 2): TMEM load, epilogue functor, staging + barrier, store issue
    for j, c in enumerate((0, 2)):
        ld, ep, stg, sto = [(t[:, 8 + 4 * j + i] - t0) / 1e3 for i in range(4)]
        prev = acc if j == 0 else (t[:, 11] - t0) / 1e3
        print(f"  chunk {c}: tmem ld {np.median(ld - prev):5.2f}  functor {np.median(ep - ld):5.2f}  stage+bar {np.median(stg - ep):5.2f}  store issue {np.median(sto - stg):5.2f} us")
    # per-SM occupancy over time: how many CTA-slots are busy on average
    busy = (end - ent).sum() / (end.max() * 148 * 2)
    print(f"  slot utilisation (sum of lifetimes / (span x 296 slots)) {busy:.2f}")
    order = np.argsort(ent)
    print("  entry times of CTAs #0, #295, #296, #591, #592, #887, #888, last:", [round(float(ent[order[i]]), 1) for i in (0, 295, 296, 591, 592, 887, 888, tiles - 1)])
