"""Second-stem depthwise convolution (tfnas_dwconv_fwd/_bwd) against the CPU oracle arithmetic (torch CPU conv2d, the
op the reference's nn.Conv2d(groups=C) delegates to: models/layers.py:486-489)."""
import pytest
import torch
import torch.nn.functional as F

from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('N,C,HH,WW,K', [(2, 32, 16, 16, 3), (3, 5, 7, 7, 3), (2, 8, 9, 14, 5), (1, 32, 112, 112, 3)])
def test_dwconv_matches_cpu_conv(N, C, HH, WW, K):
    from tfnas_b200.ops import dwconv
    g = torch.Generator().manual_seed(N * 100 + C)
    x = torch.randn(N, C, HH, WW, generator=g)
    w = torch.randn(C, 1, K, K, generator=g) * 0.3
    G = torch.randn(N, C, HH, WW, generator=g)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, 1, K // 2, 1, C)
    (yr * G).sum().backward()
    xg, wg = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    y = dwconv(xg, wg)
    (y * G.cuda()).sum().backward()
    assert H.rel_l2(y, yr) < 1e-5
    assert H.rel_l2(xg.grad, xr.grad) < 1e-5
    assert H.rel_l2(wg.grad, wr.grad) < 1e-4


def test_dwconv_rejects_cpu_tensor():
    from tfnas_b200 import _lib
    from tfnas_b200.ops import dwconv
    with pytest.raises(_lib.TfnasError):
        dwconv(torch.randn(1, 4, 8, 8), torch.randn(4, 1, 3, 3))
