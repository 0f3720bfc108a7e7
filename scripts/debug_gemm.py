import sys, torch, time
sys.path.insert(0, '.')
from ab_opt_b200 import _capi as C
DEV='cuda:0'
for (M,N,K) in [(1024,128,32),(1024,128,128),(1024,128,512),(1024,128,1824),(1024,128,4096)]:
    g = torch.Generator(device=DEV).manual_seed(1)
    A = torch.randn(M, K, generator=g, device=DEV); B = torch.randn(N, K, generator=g, device=DEV) / K ** 0.5
    D = torch.empty(M, N, device=DEV)
    C.check(C.lib().abopt_debug_gemm3x(0, M, N, K, C.ptr(A), C.ptr(B), None, C.ptr(D), C.stream_ptr(torch.device(DEV))))
    r64 = A.double() @ B.double().T
    r32 = (A @ B.T).double()
    e = D.double() - r64
    print(f'K={K:5d} tc max {e.abs().max():.2e} rms {e.pow(2).mean().sqrt():.2e} signed-bias {(e*torch.sign(r64)).mean():+.2e} | fp32 max {(r32-r64).abs().max():.2e} rms {(r32-r64).pow(2).mean().sqrt():.2e}')
    # positive-only operands: truncation shows as a systematic negative bias
    A2, B2 = A.abs(), B.abs()
    C.check(C.lib().abopt_debug_gemm3x(0, M, N, K, C.ptr(A2), C.ptr(B2), None, C.ptr(D), C.stream_ptr(torch.device(DEV))))
    r64 = A2.double() @ B2.double().T
    e = (D.double() - r64) / r64
    print(f'        positive operands: rel err mean {e.mean():+.2e} max {e.abs().max():.2e}')
