"""GPU debug driver: compare every saved intermediate of the CUDA path with oracle/fused_math.py.

    python tools/debug_phases.py            (on a GPU box)
Prints one line per (shape, mode, tensor) with l2-relative error; exits non-zero above tolerance.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fused_math as fm, port  # noqa: E402
from tests import helpers as H  # noqa: E402
from tfnas_b200.config import CAND_SPEC  # noqa: E402

TOL = 2e-4
bad = []


def report(tag, name, a, b, tol=TOL):
    e = H.rel_l2(a, b)
    flag = '' if e < tol else '   <<<<<< FAIL'
    if not (e < tol):
        bad.append((tag, name, e))
    print('%-34s %-10s l2rel %.2e  maxrel %.2e%s' % (tag, name, e, H.rel_max(a, b), flag), flush=True)


def run(ic, oc, stride, act, Hs, N, ragged, seed, Wd=None):
    mcs = H.default_mcs(ic, ragged)
    P, x, gum, lats = H.make_problem(ic, oc, stride, Hs, N, mcs, seed, W=Wd)
    tag = 'ic%d oc%d s%d %s %dx%d N%d%s' % (ic, oc, stride, act, Hs, Wd or Hs, N, ' rag' if ragged else '')
    dt = torch.float64
    Pd = {k: v.to(dt) for k, v in P.items()}
    cands = [fm.cand_weights(Pd, 'b.', i) for i in range(8)]
    T = 5.0
    w = port.gumbel_weights(Pd['b.log_alphas'], gum.to(dt), T)
    out_ref, S = fm.forward(x.to(dt), cands, list(range(8)), stride, act, w)
    g = torch.Generator().manual_seed(seed + 1)
    G = torch.randn(out_ref.shape, generator=g)
    dlat = 0.37
    dx_ref, dmix, _ = fm.backward(x.to(dt), cands, list(range(8)), stride, act, S, G.to(dt), w)
    da_ref = fm.alpha_grad(dmix, w, lats.to(dt), dlat, T)
    t0 = time.time()
    r = H.raw_call(P, x, gum, lats, ic, oc, stride, act, mcs, 0xFF, T, G, dlat)
    # forward intermediates
    report(tag, 'mean_x', r['xmom'][:ic], S['mu_x'])
    report(tag, 'cov_x', r['xmom'][ic:].view(ic, ic), S['cov'])
    MC = sum(mcs)
    off = 0
    zs = 0
    for i in range(8):
        c = S['c'][i]
        mc = mcs[i]
        report(tag, 'mu1[%d]' % i, r['bn1'][off:off + mc], c['mu1'], 1e-3)
        report(tag, 'r1[%d]' % i, r['bn1'][MC + off:MC + off + mc], c['r1'])
        uh = (torch.einsum('ck,nkhw->nchw', cands[i]['w1'], x.to(dt)) - c['mu1'][None, :, None, None]) * c['r1'][None, :, None, None]
        report(tag, 'UH[%d]' % i, r['UH'][:, off:off + mc], uh)
        report(tag, 'D[%d]' % i, r['D'][:, off:off + mc], c['d'])
        report(tag, 'r2[%d]' % i, r['bn2'][MC + off:MC + off + mc], c['r2'])
        if 'g' in c:
            report(tag, 'segate[%d]' % i, r['seg'][:, zs:zs + mc], c['g'])
            zs += mc
        report(tag, 'Z[%d]' % i, r['Z'][:, i], c['z'])
        report(tag, 'r3[%d]' % i, r['bn3'][8 * oc + i * oc:8 * oc + (i + 1) * oc], c['r3'])
        off += mc
    report(tag, 'mixw', r['mixw'], w)
    report(tag, 'out', r['out'], out_ref)
    print('%-34s %-10s cuda %.6f ref %.6f' % (tag, 'lat', r['out_lat'], float((w * lats.to(dt)).sum())))
    report(tag, 'dx', r['dx'], dx_ref)
    report(tag, 'dalpha', r['dalpha'], da_ref, 1e-3)
    # sampled mode with weight grads, two candidates
    for idx in (1, 6):
        o_ref, S1 = fm.forward(x.to(dt), cands, [idx], stride, act, None)
        dx1, _, wg = fm.backward(x.to(dt), cands, [idx], stride, act, S1, G.to(dt), None, True)
        r1 = H.raw_call(P, x, gum, lats, ic, oc, stride, act, mcs, 1 << idx, T, G, 0.0, want_wgrad=True)
        t = tag + ' op%d' % idx
        report(t, 'out', r1['out'], o_ref)
        report(t, 'dx', r1['dx'], dx1)
        m = dict(w1='w1', dw='dw', w3='w3', se_rw='rw', se_rb='rb', se_ew='ew', se_eb='eb')
        for (i, s), gt in r1['wgrads'].items():
            report(t, 'd' + s, gt.reshape(wg[idx][m[s]].shape), wg[idx][m[s]], 5e-4)
    print('%s done in %.1fs' % (tag, time.time() - t0), flush=True)


if __name__ == '__main__':
    torch.manual_seed(0)
    cases = [
        (8, 8, 1, 'swish', 7, 3, False, 1, None),
        (8, 16, 2, 'relu', 12, 2, True, 2, None),
        (16, 24, 2, 'relu', 20, 2, False, 3, None),
        (24, 24, 1, 'relu', 14, 2, True, 4, None),
        (40, 40, 1, 'swish', 9, 3, False, 5, 11),
        (40, 80, 2, 'swish', 13, 2, True, 6, None),
    ]
    if len(sys.argv) > 1:
        cases = cases[:int(sys.argv[1])]
    for c in cases:
        run(*c)
    print('FAILURES:', len(bad))
    for b in bad:
        print('  ', b)
    sys.exit(1 if bad else 0)
