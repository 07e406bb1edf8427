"""BASELINE-size (N=128) checks of individual MixedOPs: parity against the oracle port running on the
same GPU in fp32 (cuDNN, TF32 off — the CPU oracle would need minutes and ~70 GB here), plus
size-independent properties of the path: zero channel mean of the BN'd mixture, linearity of the
backward in dL/dout, and sum_j dL/dlog_alpha_j = 0 (softmax Jacobian)."""
import pytest
import torch

from oracle import port
from tests import helpers as H
from tfnas_b200.config import CAND_SPEC, lut_key
from tfnas_b200.model_search import MixedOP, NoisePlan, injected

pytestmark = pytest.mark.gpu

SHAPES = [  # ic, oc, s, act, H  (a stride-2 relu, a residual swish, a 14x14 and a 7x7 layer)
    (16, 24, 2, 'relu', 112),
    (40, 40, 1, 'swish', 28),
    (112, 112, 1, 'swish', 14),
    (192, 320, 1, 'swish', 7),
]


@pytest.mark.parametrize('ic,oc,s,act,size', SHAPES)
def test_fullsize_parity_and_properties(ic, oc, s, act, size):
    _fullsize(ic, oc, s, act, size, H.default_mcs(ic))


def _fullsize(ic, oc, s, act, size, mcs):
    N = 128
    P, x, gum, lats = H.make_problem(ic, oc, s, size, N, mcs, seed=size)
    lut = {}
    for i, (k, _e, sm) in enumerate(CAND_SPEC):
        lut.setdefault(lut_key(size, ic, sm * ic, oc, k, s, act), {})[mcs[i]] = float(lats[i])
    lat_list = [lut[lut_key(size, ic, sm * ic, oc, k, s, act)][mcs[i]] for i, (k, _e, sm) in enumerate(CAND_SPEC)]
    op = MixedOP(ic, oc, s, False, act, 8, {i: mcs[i] for i in range(8)}, lut)
    op.load_state_dict({k[2:]: v for k, v in P.items()})
    op.set_temperature(5.0)
    op.cuda()
    for n, p in op.named_parameters():
        p.requires_grad_(n == 'log_alphas')
    xg = x.cuda().requires_grad_(True)
    g = torch.Generator().manual_seed(1)

    def run(G):
        xg.grad = None
        op.log_alphas.grad = None
        with injected(NoisePlan(noise=[gum])):
            out, lat = op(xg, False, 'max')
        (out * G).sum().add(lat * 0.37).backward()
        return out.detach(), xg.grad.clone(), op.log_alphas.grad.clone()

    out0, _, _ = run(torch.zeros(1, device='cuda'))
    G1 = torch.randn(out0.shape, generator=g).cuda()
    G2 = torch.randn(out0.shape, generator=g).cuda()
    out, dx1, da1 = run(G1)
    _, dx2, da2 = run(G2)
    _, dx12, da12 = run(G1 + G2)
    # properties
    mix = out - (xg.detach() if (ic == oc and s == 1) else 0)
    assert float(mix.mean((0, 2, 3)).abs().max()) < 1e-4
    assert H.rel_l2(dx12, dx1 + dx2) < 1e-4
    lat_term = op.log_alphas.grad * 0  # latency part enters all three runs once
    assert abs(float(da12.sum())) < 1e-3 * float(da12.abs().max()) + 1e-6
    # parity against the oracle port on the same device (fp32, TF32 off)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        Pd = {k: v.cuda().requires_grad_(k.endswith('log_alphas')) for k, v in P.items()}
        xr = x.cuda().requires_grad_(True)
        o_ref, l_ref = port.mixedop_alpha(xr, Pd, 'b.', ic, oc, s, act, 5.0, gum.cuda(), lat_list)
        ((o_ref * G1).sum() + l_ref * 0.37).backward()
    e = dict(out=H.rel_l2(out, o_ref.detach()), dx=H.rel_l2(dx1, xr.grad), dalpha=H.rel_l2(da1, Pd['b.log_alphas'].grad))
    print(ic, oc, s, act, size, e)
    # ReLU layers: two fp32 implementations disagree on the gate of the few hundred (out of ~1e9)
    # pre-activations that sit within rounding of 0, which alone gives ~3e-4 l2 on dx; the contract
    # tolerance is 1e-3.  Swish layers are smooth and agree to ~1e-6.
    dx_tol = 1e-3 if act == 'relu' else 2e-4
    assert e['out'] < 1e-4 and e['dx'] < dx_tol and e['dalpha'] < 1e-3
