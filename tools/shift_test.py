import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chatterbox_embed_b200 import _lib
L = _lib.lib(); ctx = _lib.context(0)
fn = L.cbx_test_shift_gemm; fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
torch.manual_seed(0)
A = torch.randn(256, 32, device="cuda"); W = torch.randn(32, 32, device="cuda"); Cc = torch.zeros(128, 32, device="cuda")
for ubo in (0, 1):
    for shift in (0, 1, 2, 5, 7, 8, 9, 42, 86, 127):
        Cc.zero_()
        rc = fn(ctx._h, A.data_ptr(), W.data_ptr(), Cc.data_ptr(), shift, ubo)
        assert rc == 0, L.cbx_last_error(ctx._h)
        ref = A[shift:shift + 128].double() @ W.double().T
        err = (Cc.double() - ref).abs().max().item()
        print(f"base_offset={ubo} shift={shift:3d}: max err {err:.3e} {'OK' if err < 2e-2 else 'WRONG'}", flush=True)
