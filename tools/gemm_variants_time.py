"""Single-launch timing of the pre-activation GEMM variants of cbx_test_tgemm at the bench's row count.
    python tools/gemm_variants_time.py 1 4 ..."""
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
variants = [int(v) for v in sys.argv[1:]] or [4]
M = 130048
for K in (128, 256, 512, 1024):
    A = torch.randn(M + 8, 1024, device=dev); W = torch.randn(128, K, device=dev) / K ** 0.5
    bias = torch.randn(128, device=dev); a = torch.rand(1024, device=dev) + 0.5; b = torch.randn(1024, device=dev) * 0.3
    Cc = torch.empty(M, 128, device=dev)
    for v in variants:
        args = (ctx._h, A.data_ptr(), 1024, W.data_ptr(), K, Cc.data_ptr(), 128, M, 128, K, bias.data_ptr(), a.data_ptr(), b.data_ptr(), v, 0, None)
        for _ in range(3):
            assert fn(*args) == 0, L.cbx_last_error(ctx._h)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn(*args)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"K={K} variant {v}: {ms*1e3:.1f} us  {M*K*4/ms/1e6:.0f} GB/s of X", flush=True)
