"""The fused algebra the CUDA kernels implement (oracle/fused_math.py) == autograd of the literal
port, in fp64, including odd sizes, stride 2, ragged widths and the analytic BN1 / dW1 folds."""
import pytest
import torch

from oracle import fused_math as fm, port
from tests import helpers as H


@pytest.mark.parametrize('ic,oc,s,act,size,N', [(8, 8, 1, 'swish', 6, 3), (8, 12, 2, 'relu', 8, 2),
                                                 (6, 6, 1, 'relu', 7, 2), (6, 10, 2, 'swish', 9, 2)])
def test_fused_alpha_and_single(ic, oc, s, act, size, N):
    mcs = H.default_mcs(ic, ragged=True)
    P, x, gum, lats = H.make_problem(ic, oc, s, size, N, mcs, seed=ic + size, dtype=torch.float64)
    g = torch.Generator().manual_seed(9)
    T, dlat = 5.0, 0.37
    ref = H.oracle_alpha(P, x, gum, lats, ic, oc, s, act, T, None)
    G = torch.randn(ref['out'].shape, generator=g, dtype=torch.float64)
    ref = H.oracle_alpha(P, x, gum, lats, ic, oc, s, act, T, G, dlat)
    cands = [fm.cand_weights(P, 'b.', i) for i in range(8)]
    w = port.gumbel_weights(P['b.log_alphas'], gum, T)
    out, S = fm.forward(x, cands, list(range(8)), s, act, w)
    dx, dmix, _ = fm.backward(x, cands, list(range(8)), s, act, S, G, w)
    da = fm.alpha_grad(dmix, w, lats, dlat, T)
    assert H.rel_max(out, ref['out']) < 1e-12
    assert H.rel_max(dx, ref['dx']) < 1e-11
    assert H.rel_max(da, ref['dalpha']) < 1e-10
    for idx in (0, 3, 5, 6):
        r1 = H.oracle_single(P, x, ic, oc, s, act, idx, G)
        o, S1 = fm.forward(x, cands, [idx], s, act, None)
        dx1, _, wg = fm.backward(x, cands, [idx], s, act, S1, G, None, True)
        assert H.rel_max(o, r1['out']) < 1e-12
        assert H.rel_max(dx1, r1['dx']) < 1e-11
        m = dict(w1='w1', dw='dw', w3='w3', se_rw='rw', se_rb='rb', se_ew='ew', se_eb='eb')
        for (i, sname), gref in r1['wgrads'].items():
            assert H.rel_max(wg[idx][m[sname]].reshape(gref.shape), gref) < 1e-10, sname
