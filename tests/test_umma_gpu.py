"""tcgen05 building blocks (csrc/umma.cuh): the 3-term tf32 split GEMM on the tensor cores must be
fp32-accurate (plain tf32 fails the 1e-3 parity bar, the split matches fp32)."""
import ctypes

import pytest
import torch

from tests import helpers as H
from tfnas_b200 import _lib

pytestmark = pytest.mark.gpu


def run_selftest(M, N, K, variant=0, seed=0):
    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(K, M, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    C = torch.full((N, M), float('nan'), device='cuda')
    npad = (N + 15) // 16 * 16
    nbytes = ((K + 31) // 32) * 2 * npad * 128
    ws = torch.zeros(nbytes, dtype=torch.uint8, device='cuda')
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.tfnas_umma_selftest(M, N, K, vp(A), vp(B), vp(C), vp(ws), nbytes, variant,
                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    ref = B.double() @ A.double()
    return C, ref


@pytest.mark.parametrize('M,N,K', [(128, 128, 32), (128, 16, 8), (256, 96, 64), (1000, 112, 200), (130, 256, 576), (49, 24, 1152)])
def test_split_tf32_gemm_is_fp32_accurate(M, N, K):
    C, ref = run_selftest(M, N, K)
    # fp32 accumulation over K terms: ~sqrt(K) * 2^-24, plus the dropped lo*lo term (2^-22)
    e = H.rel_l2(C, ref)
    print('umma selftest', M, N, K, e)
    assert e < 1e-6 + 4e-7 * K ** 0.5


def test_single_tf32_is_not_enough():
    C, ref = run_selftest(256, 64, 256, variant=1)
    e = H.rel_l2(C, ref)
    assert 1e-5 < e < 5e-3     # ~2^-11 relative per product: why the split exists
