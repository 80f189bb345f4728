"""GPU: the tcgen05 descriptor / operand-layout conventions (csrc/ufo_umma.cuh) against torch matmul."""
import pytest
import torch

from uforecon_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bf16", [0, 1])
@pytest.mark.parametrize("mode,N,K", [(0, 240, 80), (0, 80, 80), (0, 160, 160), (0, 80, 160), (0, 16, 96), (0, 144, 96),
                                      (0, 176, 176), (0, 96, 176), (0, 256, 256), (1, 112, 128), (1, 96, 64)])
def test_umma_gemm(mode, N, K, bf16):
    lib = _lib.load()
    g = torch.Generator().manual_seed(N * 1000 + K + mode)
    dt = torch.bfloat16 if bf16 else torch.float16
    if mode == 0:
        A = torch.randn(128, K, generator=g)
        B = torch.randn(N, K, generator=g)
        ref = A.to(dt).double() @ B.to(dt).double().t()
    else:
        A = torch.randn(K, 128, generator=g)
        B = torch.randn(K, N, generator=g)
        ref = A.to(dt).double().t() @ B.to(dt).double()
    Ad, Bd = A.cuda().contiguous(), B.cuda().contiguous()
    D = torch.full((128, N), float("nan"), device="cuda")
    _lib.check(lib.ufo_debug_umma_selftest(Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), N, K, mode, bf16,
                                           torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    err = (D.cpu().double() - ref).abs().max().item()
    assert err < 1e-3 * (K ** 0.5), f"max err {err}"


@pytest.mark.parametrize("bf16", [0, 1])
@pytest.mark.parametrize("N,K", [(80, 80), (160, 80), (80, 160), (16, 96), (96, 176), (160, 176)])
def test_umma_gemm_ts_two_halves(N, K, bf16):
    """A operand written to TMEM by its row owners (tcgen05.st), two halves of one CTA issuing independently."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(N * 1000 + K + 7)
    dt = torch.bfloat16 if bf16 else torch.float16
    A = torch.randn(2, 128, K, generator=g)
    B = torch.randn(N, K, generator=g)
    ref = A.to(dt).double() @ B.to(dt).double().t()
    Ad, Bd = A.cuda().contiguous(), B.cuda().contiguous()
    D = torch.full((2, 128, N), float("nan"), device="cuda")
    _lib.check(lib.ufo_debug_umma_selftest(Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), N, K, 2, bf16,
                                           torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    err = (D.cpu().double() - ref).abs().max().item()
    assert err < 1e-3 * (K ** 0.5), f"max err {err}"
