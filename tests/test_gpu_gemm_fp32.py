"""The two fp32 GEMM kernels of the exact per-call backward (s2l_wgrad_rows_fp32 / s2l_dx_rows_fp32) against float64 products:
ragged row counts, narrow and non-multiple-of-64 widths, shared (stride 0) operands, strided output rows."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from speech2lip_b200 import _cabi
    return _cabi, _cabi.lib()


@pytest.mark.parametrize("N,L,A,B,share", [(1, 1, 3, 256, False), (1000, 5, 256, 256, False), (9600, 2, 256, 42, True),
                                           (70001, 2, 256, 64, True), (4097, 1, 4, 256, False)])
def test_wgrad_rows_matches_float64(N, L, A, B, share):
    cabi, lib = _lib()
    g = torch.Generator().manual_seed(N + A)
    dy = torch.randn(L, N, A, generator=g).cuda()
    h = torch.randn(1 if share else L, N, B, generator=g).cuda()
    out = torch.empty(L, A, B, device="cuda")
    scratch = torch.empty(max(16, lib.s2l_wgrad_rows_scratch_bytes(N, L, A, B)), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    cabi.check(lib.s2l_wgrad_rows_fp32(dy.data_ptr(), h.data_ptr(), N, L, A, B, N * A, 0 if share else N * B, out.data_ptr(),
                                       scratch.data_ptr(), s), "wgrad")
    ref = torch.einsum("lna,lnb->lab", dy.double(), h.double().expand(L, -1, -1))
    scale = ref.abs().max().item()
    assert (out.double() - ref).abs().max().item() <= 1e-5 * scale      # fp32 sums of N terms, slab-wise
    out2 = torch.empty_like(out)
    cabi.check(lib.s2l_wgrad_rows_fp32(dy.data_ptr(), h.data_ptr(), N, L, A, B, N * A, 0 if share else N * B, out2.data_ptr(),
                                       scratch.data_ptr(), s), "wgrad")
    assert torch.equal(out, out2)            # fixed summation order


@pytest.mark.parametrize("N,B,two,pad", [(1, 64, True, 2), (9600, 42, True, 0), (777, 63, False, 0), (5000, 64, True, 3)])
def test_dx_rows_matches_float64(N, B, two, pad):
    cabi, lib = _lib()
    g = torch.Generator().manual_seed(N + B)
    a1, a2 = torch.randn(N, 256, generator=g).cuda(), torch.randn(N, 256, generator=g).cuda()
    w1, w2 = torch.randn(256, B, generator=g).cuda(), torch.randn(256, B, generator=g).cuda()
    out = torch.full((N, B + pad), 7.0, device="cuda")
    cabi.check(lib.s2l_dx_rows_fp32(a1.data_ptr(), w1.data_ptr(), a2.data_ptr() if two else None, w2.data_ptr() if two else None, N, B,
                                    out.data_ptr() + 4 * pad, B + pad, torch.cuda.current_stream().cuda_stream), "dx")
    ref = a1.double() @ w1.double() + (a2.double() @ w2.double() if two else 0)
    assert (out[:, pad:].double() - ref).abs().max().item() <= 2e-6 * ref.abs().max().item()
    assert (out[:, :pad] == 7.0).all()       # columns outside the written block untouched


def test_gemm_argument_errors():
    cabi, lib = _lib()
    assert lib.s2l_wgrad_rows_fp32(None, None, 4, 1, 4, 4, 0, 0, None, None, None) == 1
    t = torch.zeros(64, device="cuda")
    assert lib.s2l_dx_rows_fp32(t.data_ptr(), t.data_ptr(), None, None, 1, 8, t.data_ptr(), 4, None) == 2
