"""Unit test of the CTA-pair pre-activation GEMM (variant 5) against fp64 and against the single-CTA kernel (variant 4)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chatterbox_embed_b200 import _lib
L = _lib.lib(); ctx = _lib.context(0)
fn = L.cbx_test_tgemm
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
dev = "cuda:0"
torch.manual_seed(0)
def run(M, N, K, variant, lda):
    A = torch.randn(M + 8, lda, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    bias = torch.randn(N, device=dev)
    a = torch.rand(lda, device=dev) + 0.5; b = torch.randn(lda, device=dev) * 0.3
    Cc = torch.full((M, N), float("nan"), device=dev)
    rc = fn(ctx._h, A.data_ptr(), lda, W.data_ptr(), K, Cc.data_ptr(), N, M, N, K, bias.data_ptr(), a.data_ptr(), b.data_ptr(), variant, 0, None)
    torch.cuda.synchronize()
    assert rc == 0, L.cbx_last_error(ctx._h)
    ref = torch.relu(A[:M, :K].double() * a[:K].double() + b[:K].double()) @ W.double().T + bias.double()
    err = (Cc.double() - ref).abs()
    print(f"variant {variant} M={M} N={N} K={K}: max err {err.max().item():.3e} nan={torch.isnan(Cc).sum().item()}", flush=True)
    if err.max().item() > 1e-2 or torch.isnan(Cc).any():
        bad = (err > 1e-2) | torch.isnan(Cc)
        rows = bad.any(1).nonzero().flatten(); cols = bad.any(0).nonzero().flatten()
        print("   bad rows", rows[:6].tolist(), "...", rows[-3:].tolist(), "n", len(rows), " bad cols", cols[:6].tolist(), "...", cols[-3:].tolist(), "n", len(cols))
        print("   got", Cc[rows[0], :4].tolist(), "want", ref[rows[0], :4].tolist())
    return Cc
for (M, N, K) in [(256, 128, 32), (256, 128, 128), (128, 128, 64), (300, 128, 96), (129, 128, 160), (1000, 128, 480), (512, 256, 512), (100000, 128, 1024)]:
    c4 = run(M, N, K, 4, 1024)
    c5 = run(M, N, K, 5, 1024)

# timing: single-CTA (4) vs CTA-pair (5) at the bench's row count
M = 130048
for K in (128, 256, 512, 1024):
    A = torch.randn(M + 8, 1024, device=dev); W = torch.randn(128, K, device=dev) / K ** 0.5
    bias = torch.randn(128, device=dev); a = torch.rand(1024, device=dev) + 0.5; b = torch.randn(1024, device=dev) * 0.3
    Cc = torch.empty(M, 128, device=dev)
    for v in (4, 9):
        for _ in range(3):
            fn(ctx._h, A.data_ptr(), 1024, W.data_ptr(), K, Cc.data_ptr(), 128, M, 128, K, bias.data_ptr(), a.data_ptr(), b.data_ptr(), v, 0, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn(ctx._h, A.data_ptr(), 1024, W.data_ptr(), K, Cc.data_ptr(), 128, M, 128, K, bias.data_ptr(), a.data_ptr(), b.data_ptr(), v, 0, None)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"K={K} variant {v}: {ms*1e3:.1f} us  {M*K*4/ms/1e6:.0f} GB/s of X", flush=True)
