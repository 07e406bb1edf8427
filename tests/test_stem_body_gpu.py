"""The stems in the library (tfnas_stem_fwd/_bwd: direct 3x3/2 convolution + the second stem on the MixedOP phase kernels)
against torch fp64 on the GPU: output and all seven weight gradients."""
import pytest
import torch
import torch.nn.functional as F

from tests import helpers as H
from tfnas_b200.ops import ArenaPool, StemFn

pytestmark = pytest.mark.gpu


def _ref(img, w, dtype):
    conv_w, dw, rw, rb, ew, eb, pw = [t.detach().to(dtype).requires_grad_(True) for t in w]
    bn = lambda t: F.batch_norm(t, None, None, None, None, True, 0.0, 1e-5)
    x = F.relu(bn(F.conv2d(img.to(dtype), conv_w, None, 2, 1)))
    x = F.relu(bn(F.conv2d(x, dw, None, 1, 1, 1, dw.shape[0])))
    g = F.adaptive_avg_pool2d(x, 1)
    g = F.conv2d(F.relu(F.conv2d(g, rw, rb)), ew, eb)
    x = x * torch.sigmoid(g)
    return bn(F.conv2d(x, pw)), (conv_w, dw, rw, rb, ew, eb, pw)


@pytest.mark.parametrize('N,size', [(2, 64), (3, 224), (16, 224)])
def test_stem_matches_torch_fp64(N, size):
    g = torch.Generator().manual_seed(N * 1000 + size)
    img = torch.randn(N, 3, size, size, generator=g).cuda()
    mk = lambda *s: (torch.randn(*s, generator=g) * (1.0 / max(1, s[1] if len(s) > 1 else 1) ** 0.5)).cuda().requires_grad_(True)
    w = [mk(32, 3, 3, 3), mk(32, 1, 3, 3), mk(8, 32, 1, 1), (torch.randn(8, generator=g) * 0.1).cuda().requires_grad_(True),
         mk(32, 8, 1, 1), (torch.randn(32, generator=g) * 0.1).cuda().requires_grad_(True), mk(16, 32, 1, 1)]
    G = torch.randn(N, 16, size // 2, size // 2, generator=g).cuda()
    pool = ArenaPool()
    out = StemFn.apply(img, pool, *w)
    (out * G).sum().backward()
    ref, rw = _ref(img, w, torch.float64)
    (ref * G.double()).sum().backward()
    e = dict(out=H.rel_l2(out, ref))
    names = ['conv_w', 'dw', 'se_rw', 'se_rb', 'se_ew', 'se_eb', 'pw']
    for n, a, b in zip(names, w, rw):
        e[n] = H.rel_l2(a.grad, b.grad)
    print('stem N=%d %dx%d errors' % (N, size, size), {k: '%.1e' % v for k, v in e.items()})
    assert all(v < 1e-3 for v in e.values()), e
    assert len(pool.free) == 1


def test_stem_no_grad_forward():
    g = torch.Generator().manual_seed(5)
    img = torch.randn(2, 3, 64, 64, generator=g).cuda()
    w = [torch.randn(32, 3, 3, 3, generator=g).cuda(), torch.randn(32, 1, 3, 3, generator=g).cuda(),
         torch.randn(8, 32, 1, 1, generator=g).cuda(), torch.zeros(8).cuda(), torch.randn(32, 8, 1, 1, generator=g).cuda(),
         torch.zeros(32).cuda(), torch.randn(16, 32, 1, 1, generator=g).cuda()]
    with torch.no_grad():
        out = StemFn.apply(img, ArenaPool(), *w)
    ref, _ = _ref(img, w, torch.float64)
    assert H.rel_l2(out, ref) < 1e-4


def _head_ref(x, fm_w, fc_w, fc_b, dtype):
    ws = [t.detach().to(dtype).requires_grad_(True) for t in (fm_w, fc_w, fc_b)]
    xr = x.detach().to(dtype).requires_grad_(True)
    y = F.batch_norm(F.conv2d(xr, ws[0]), None, None, None, None, True, 0.0, 1e-5)
    y = y * torch.sigmoid(y)
    p = F.adaptive_avg_pool2d(y, 1).flatten(1)
    return F.linear(p, ws[1], ws[2]), xr, ws


@pytest.mark.parametrize('N,hw,cin,cmid,ncls', [(4, 7, 320, 1280, 100), (128, 7, 320, 1280, 100), (3, 5, 72, 200, 10)])
def test_head_matches_torch_fp64(N, hw, cin, cmid, ncls):
    """tfnas_head_fwd/_bwd (feature-mix 1x1 conv on the tcgen05 project / dc / weight-gradient kernels with the identity
    activation, BN + Swish + pooling, classifier) against torch fp64: logits, dx and the three weight gradients."""
    from tfnas_b200.ops import HeadFn
    g = torch.Generator().manual_seed(N + hw)
    x = torch.randn(N, cin, hw, hw, generator=g).cuda().requires_grad_(True)
    fm_w = (torch.randn(cmid, cin, 1, 1, generator=g) / cin ** 0.5).cuda().requires_grad_(True)
    fc_w = (torch.randn(ncls, cmid, generator=g) / cmid ** 0.5).cuda().requires_grad_(True)
    fc_b = (torch.randn(ncls, generator=g) * 0.1).cuda().requires_grad_(True)
    G = torch.randn(N, ncls, generator=g).cuda()
    pool = ArenaPool()
    logits = HeadFn.apply(x, pool, fm_w, fc_w, fc_b)
    (logits * G).sum().backward()
    ref, xr, ws = _head_ref(x, fm_w, fc_w, fc_b, torch.float64)
    (ref * G.double()).sum().backward()
    e = dict(logits=H.rel_l2(logits, ref), dx=H.rel_l2(x.grad, xr.grad), fm_w=H.rel_l2(fm_w.grad, ws[0].grad),
             fc_w=H.rel_l2(fc_w.grad, ws[1].grad), fc_b=H.rel_l2(fc_b.grad, ws[2].grad))
    print('head N=%d %dx%d %d->%d->%d errors' % (N, hw, hw, cin, cmid, ncls), {k: '%.1e' % v for k, v in e.items()})
    assert all(v < 1e-3 for v in e.values()), e
    assert len(pool.free) == 1
    # weights frozen (alpha step): dx only
    x2 = x.detach().clone().requires_grad_(True)
    l2 = HeadFn.apply(x2, pool, fm_w.detach(), fc_w.detach(), fc_b.detach())
    (l2 * G).sum().backward()
    assert H.rel_l2(x2.grad, xr.grad) < 1e-3
