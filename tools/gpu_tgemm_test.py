"""Unit test of the tcgen05 GEMM engine against torch fp64 matmul (GPU box)."""
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
def tf32(x):  # round-to-nearest-even emulation is close enough for a tolerance check
    return (x.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32) if False else x
def run(M, N, K, variant, lda=None, shift=0):
    lda = lda or K
    A = torch.randn(M + 8, lda, device=dev)
    ntaps = 2 if variant == 3 else 1
    W = torch.randn(N, K * ntaps, device=dev) / K ** 0.5
    bias = torch.randn(N, device=dev)
    a = torch.rand(lda, device=dev) + 0.5; b = torch.randn(lda, device=dev) * 0.3
    Cc = torch.full((M, N), float("nan"), device=dev)
    rc = fn(ctx._h, A.data_ptr(), lda, W.data_ptr(), K * ntaps, Cc.data_ptr(), N, M, N, K, bias.data_ptr(), a.data_ptr(), b.data_ptr(), variant, shift, None)
    torch.cuda.synchronize()
    assert rc == 0, L.cbx_last_error(ctx._h)
    Ad = A[:M, :K].double()
    if variant == 1:
        Ad = torch.relu(Ad * a[:K].double() + b[:K].double())
    ref = Ad @ W[:, :K].double().T + bias.double()
    if variant == 3:
        A2 = torch.zeros(M, K, device=dev, dtype=torch.float64)
        src = torch.arange(M, device=dev) + shift
        ok = (src >= 0) & (src < M)
        A2[ok] = A[src[ok], :K].double()
        ref = ref + A2 @ W[:, K:].double().T
    err = (Cc.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"variant {variant} M={M} N={N} K={K} lda={lda}: max err {err:.3e} (scale {scale:.2f}, rel {err/scale:.2e}) nan={torch.isnan(Cc).sum().item()}")
    return err / scale
ok = True
for (M, N, K, v, lda, sh) in [(128, 128, 32, 0, None, 0), (128, 128, 64, 0, None, 0), (256, 128, 256, 0, None, 0), (1000, 1024, 256, 0, None, 0), (777, 1024, 40, 0, None, 0),
                          (512, 128, 480, 1, 1024, 0), (300, 128, 992, 1, 1024, 0), (640, 32, 384, 2, None, 0), (515, 128, 128, 3, None, 2), (515, 128, 128, 3, None, -1)]:
    try:
        r = run(M, N, K, v, lda, sh)
        ok &= r < 3e-3
    except Exception as e:
        print("FAILED", (M, N, K, v), e); ok = False
print("ALL OK" if ok else "SOME FAILED")
# timing of a big plain GEMM
M, N, K = 491520, 1024, 256
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); Cc = torch.empty(M, N, device=dev); bias = torch.zeros(N, device=dev)
for _ in range(2):
    fn(ctx._h, A.data_ptr(), K, W.data_ptr(), K, Cc.data_ptr(), N, M, N, K, bias.data_ptr(), None, None, 0, 0, None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    fn(ctx._h, A.data_ptr(), K, W.data_ptr(), K, Cc.data_ptr(), N, M, N, K, bias.data_ptr(), None, None, 0, 0, None)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"big GEMM {M}x{N}x{K}: {ms:.3f} ms, {2*M*N*K/ms/1e9:.1f} TFLOP/s, write {M*N*4/ms/1e6:.0f} GB/s")
torch.backends.cuda.matmul.allow_tf32 = True
for _ in range(2): torch.matmul(A, W.T, out=Cc)
e0.record()
for _ in range(5): torch.matmul(A, W.T, out=Cc)
e1.record(); torch.cuda.synchronize()
print(f"cuBLAS tf32 same shape: {e0.elapsed_time(e1)/5:.3f} ms")
