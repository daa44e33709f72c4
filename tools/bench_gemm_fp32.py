"""Times s2l_wgrad_rows_fp32 / s2l_dx_rows_fp32 at the shapes of the exact per-call backward against torch (cuBLAS fp32, TF32 off)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech2lip_b200 import _cabi

torch.backends.cuda.matmul.allow_tf32 = False
lib = _cabi.lib()
st = torch.cuda.current_stream().cuda_stream


def timeit(f, n=20):
    for _ in range(3):
        f()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for N in (9600, 65536):
    for (L, A, B, share) in ((5, 256, 256, False), (2, 256, 256, False), (2, 256, 42, True), (2, 256, 64, True), (1, 4, 256, False)):
        dy = torch.randn(L, N, A, device="cuda")
        h = torch.randn(1 if share else L, N, B, device="cuda")
        out = torch.empty(L, A, B, device="cuda")
        scr = torch.empty(max(16, lib.s2l_wgrad_rows_scratch_bytes(N, L, A, B)), dtype=torch.uint8, device="cuda")
        ours = timeit(lambda: lib.s2l_wgrad_rows_fp32(dy.data_ptr(), h.data_ptr(), N, L, A, B, N * A, 0 if share else N * B, out.data_ptr(),
                                                      scr.data_ptr(), st))
        he = h.expand(L, -1, -1)
        lib_ms = timeit(lambda: torch.bmm(dy.transpose(1, 2), he))
        fl = 2.0 * L * N * A * B
        print("wgrad N=%6d L=%d %3dx%3d: ours %.3f ms (%.1f TFLOP/s)  torch.bmm %.3f ms (%.1f)" % (N, L, A, B, ours, fl / ours / 1e9, lib_ms,
                                                                                                  fl / lib_ms / 1e9))
    a1, a2 = torch.randn(N, 256, device="cuda"), torch.randn(N, 256, device="cuda")
    for B in (64, 42):
        w1, w2 = torch.randn(256, B, device="cuda"), torch.randn(256, B, device="cuda")
        out = torch.empty(N, B, device="cuda")
        ours = timeit(lambda: lib.s2l_dx_rows_fp32(a1.data_ptr(), w1.data_ptr(), a2.data_ptr(), w2.data_ptr(), N, B, out.data_ptr(), B, st))
        lib_ms = timeit(lambda: a1 @ w1 + a2 @ w2)
        print("dx    N=%6d B=%d: ours %.3f ms  torch %.3f ms" % (N, B, ours, lib_ms))
