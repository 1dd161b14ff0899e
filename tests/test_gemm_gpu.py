"""The tcgen05/TMA mainloop on its own: clipk_gemm16 against torch fp32 matmul, every operand-major combination.

These are the three GEMM shapes of the path: S = X Y^T (K-major, K-major), dX = G Y (K-major, MN-major) and
dY = G^T X (MN-major, MN-major).  bf16 products are exact in fp32, so only the accumulation order differs.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(A, B, M, N, K, a_mn, b_mn, D=None, accumulate=0, f16=0):
    from clipk import _lib
    lib = _lib.load()
    if D is None:
        D = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32)
    rc = lib.clipk_gemm16(A.data_ptr(), B.data_ptr(), D.data_ptr(), M, N, K, A.stride(0), B.stride(0), D.stride(0),
                             a_mn, b_mn, f16, accumulate, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "clipk_gemm16")
    torch.cuda.synchronize()
    return D


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 512), (384, 256, 1024), (100, 72, 200), (4096, 512, 4096)])
def test_gemm_matches_torch(M, N, K, a_mn, b_mn):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + a_mn * 2 + b_mn)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    ref = A.float() @ B.float().T
    # MN-major operands are stored [K, mn]; pad mn to a multiple of 8 so the row pitch is 16-byte aligned
    def store(x, mn_major):
        if not mn_major:
            if K % 8:
                buf = torch.zeros(x.shape[0], (K + 7) // 8 * 8, device="cuda", dtype=torch.bfloat16)
                buf[:, :K] = x
                return buf[:, :K]
            return x.contiguous()
        mn = x.shape[0]
        buf = torch.zeros(K, (mn + 7) // 8 * 8, device="cuda", dtype=torch.bfloat16)
        buf[:, :mn] = x.T
        return buf[:, :mn]
    Ds = torch.full((M, (N + 3) // 4 * 4), float("nan"), device="cuda")
    D = _gemm(store(A, a_mn), store(B, b_mn), M, N, K, a_mn, b_mn, D=Ds[:, :N] if N % 4 == 0 else None)
    err = (D[:, :N] - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 1e-4 * scale + 1e-3, (err, scale)


def test_gemm_accumulate():
    M, N, K = 256, 256, 128
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn(N, K, device="cuda").bfloat16()
    D = torch.ones(M, N, device="cuda")
    _gemm(A, B, M, N, K, 0, 0, D=D, accumulate=1)
    ref = A.float() @ B.float().T + 1.0
    assert (D - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("a_mn", [0, 1])
def test_gemm_fp16_operands(a_mn):
    """The gradient GEMMs multiply the fp16 G panel by fp16 copies of the features (kind::f16 with both formats F16;
    mixing fp16 with bf16 in one MMA is an illegal instruction on sm_100a)."""
    M, N, K = 256, 512, 320
    A = (torch.randn(M, K, device="cuda") * 3).half()
    B = torch.randn(N, K, device="cuda").half()
    ref = A.float() @ B.float().T
    Ast = A.T.contiguous() if a_mn else A.contiguous()       # MN-major A is stored [K, M]
    Bst = B.T.contiguous()                                  # B is MN-major in both gradient GEMMs: stored [K, N]
    D = _gemm(Ast, Bst, M, N, K, a_mn, 1, f16=1)
    assert (D - ref).abs().max().item() <= 1e-4 * ref.abs().max().item() + 1e-3
