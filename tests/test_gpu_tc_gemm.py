"""The tcgen05 3xTF32 GEMM building block against an fp64 matmul: fp32-grade accuracy is the point
(single-pass TF32 would be ~1e-3)."""
import pytest
import torch

from ab_opt_b200 import _capi as C

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('M,N,K', [(128, 128, 32), (128, 128, 128), (300, 128, 1824), (1000, 256, 128), (77, 64, 64)])
def test_gemm3x_matches_fp64(M, N, K):
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g, device=DEV)
    B = torch.randn(N, K, generator=g, device=DEV) / K ** 0.5
    bias = torch.randn(N, generator=g, device=DEV)
    D = torch.full((M, N), float('nan'), device=DEV)
    C.check(C.lib().abopt_debug_gemm3x(0, M, N, K, C.ptr(A), C.ptr(B), C.ptr(bias), C.ptr(D), C.stream_ptr(torch.device(DEV))))
    ref64 = A.double() @ B.double().T + bias.double()
    ref32 = (A @ B.T + bias).double()            # cuBLAS fp32 (allow_tf32 off by default)
    e_tc = (D.double() - ref64).abs().max().item()
    e_32 = (ref32 - ref64).abs().max().item()
    assert torch.isfinite(D).all()
    assert e_tc <= 4 * e_32 + 2e-6, (e_tc, e_32)
