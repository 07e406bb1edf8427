"""CPU ORACLE — test infrastructure.  Phase-by-phase restatement of the FUSED algebra.

``oracle/port.py`` restates the reference literally (autograd does the backward).
This file restates the same MixedOP as the CUDA path computes it — analytic BN1
statistics from the input moments, materialised D/Z, residual folded out of the
candidate sum, explicit hand-derived backward with the BN1 backward folded into
an ic x ic correction — so every CUDA phase (F0..F4, B1..B4 in DESIGN.md) has a
CPU twin whose intermediates can be compared one by one.  It is validated
against ``port.py`` autograd in fp64 (tests/test_fused_math.py).

Math: SURVEY.md Appendix C; reference models/layers.py:539-561,
models/model_search.py:86-91.
"""
import torch
import torch.nn.functional as F

from tfnas_b200.config import CAND_SPEC

EPS = 1e-5


def act_f(x, act):
    return torch.relu(x) if act == 'relu' else x * torch.sigmoid(x)


def act_df(x, act):
    if act == 'relu':
        return (x > 0).to(x.dtype)
    s = torch.sigmoid(x)
    return s * (1 + x * (1 - s))


def cand_weights(P, prefix, i):
    pre = '%sm_ops.%d.' % (prefix, i)
    w = dict(w1=P[pre + 'inverted_bottleneck.conv.weight'].flatten(1),      # [mc, ic]
             dw=P[pre + 'depth_conv.conv.weight'],                            # [mc,1,k,k]
             w3=P[pre + 'point_linear.conv.weight'].flatten(1))              # [oc, mc]
    if CAND_SPEC[i][2]:
        w['rw'] = P[pre + 'squeeze_excite.conv_reduce.weight'].flatten(1)    # [se, mc]
        w['rb'] = P[pre + 'squeeze_excite.conv_reduce.bias']
        w['ew'] = P[pre + 'squeeze_excite.conv_expand.weight'].flatten(1)    # [mc, se]
        w['eb'] = P[pre + 'squeeze_excite.conv_expand.bias']
    return w


def forward(x, cands, active, stride, act, mix_w=None):
    """x [N,ic,H,W]; cands: list of 8 weight dicts (None if inactive); active: list of ids.

    mix_w: tensor [8] of mixing weights (alpha mode) or None (single-candidate mode,
    weight 1).  Returns (out, saved) where saved holds what the CUDA path keeps.
    """
    N, ic, H, W = x.shape
    # F0: input moments
    xf = x.permute(1, 0, 2, 3).reshape(ic, -1)
    Pn = xf.shape[1]
    mu_x = xf.mean(1)
    xc = xf - mu_x[:, None]
    cov = (xc @ xc.t()) / Pn
    S = dict(mu_x=mu_x, cov=cov, c={})
    out = None
    for i in active:
        wt = cands[i]
        k = CAND_SPEC[i][0]
        mc = wt['w1'].shape[0]
        # analytic BN1 statistics
        mu1 = wt['w1'] @ mu_x
        v1 = ((wt['w1'] @ cov) * wt['w1']).sum(1)
        r1 = torch.rsqrt(v1 + EPS)
        # F1: expand + BN1 + act + depthwise -> D, BN2 stats
        u = torch.einsum('ck,nkhw->nchw', wt['w1'], x)
        uh = (u - mu1[None, :, None, None]) * r1[None, :, None, None]
        a = act_f(uh, act)
        d = F.conv2d(a, wt['dw'], None, stride, k // 2, 1, mc)
        mu2 = d.mean((0, 2, 3))
        v2 = (d * d).mean((0, 2, 3)) - mu2 * mu2
        r2 = torch.rsqrt(v2 + EPS)
        dh = (d - mu2[None, :, None, None]) * r2[None, :, None, None]
        b = act_f(dh, act)
        c = dict(mu1=mu1, r1=r1, d=d, mu2=mu2, r2=r2)
        # F2: SE
        if 'rw' in wt:
            p = b.mean((2, 3))                              # [N, mc]
            t = p @ wt['rw'].t() + wt['rb']                 # [N, se]
            h = act_f(t, act)
            e = h @ wt['ew'].t() + wt['eb']                 # [N, mc]
            g = torch.sigmoid(e)
            cc = b * g[:, :, None, None]
            c.update(p=p, t=t, g=g)
        else:
            cc = b
        # F3: project + BN3 stats
        z = torch.einsum('oc,nchw->nohw', wt['w3'], cc)
        mu3 = z.mean((0, 2, 3))
        v3 = (z * z).mean((0, 2, 3)) - mu3 * mu3
        r3 = torch.rsqrt(v3 + EPS)
        c.update(z=z, mu3=mu3, r3=r3)
        S['c'][i] = c
        # F4: combine
        yh = (z - mu3[None, :, None, None]) * r3[None, :, None, None]
        term = yh if mix_w is None else mix_w[i] * yh
        out = term if out is None else out + term
    oc = out.shape[1]
    S['residual'] = (ic == oc and stride == 1)
    if S['residual']:
        out = out + x          # sum_i w_i == 1 folds the residual (single mode: weight 1)
    return out, S


def backward(x, cands, active, stride, act, S, G, mix_w=None, want_wgrad=False):
    """Explicit backward.  Returns dx, dmix (dL/dw_i without the latency term; None in
    single mode), and dict of weight grads per candidate if want_wgrad."""
    N, ic, H, W = x.shape
    Pn = N * H * W
    Ho, Wo = G.shape[2], G.shape[3]
    Q = N * Ho * Wo
    mu_x, cov = S['mu_x'], S['cov']
    dx_main = torch.zeros_like(x)
    cvec = torch.zeros(ic, dtype=x.dtype)
    Mm = torch.zeros(ic, ic, dtype=x.dtype)
    dmix = torch.zeros(8, dtype=x.dtype) if mix_w is not None else None
    wg = {}
    sG = G.sum((0, 2, 3))
    for i in active:
        wt, c = cands[i], S['c'][i]
        k = CAND_SPEC[i][0]
        mc = wt['w1'].shape[0]
        wi = 1.0 if mix_w is None else mix_w[i]
        # B1: BN3-backward statistics (+ dL/dw_i)
        yh = (c['z'] - c['mu3'][None, :, None, None]) * c['r3'][None, :, None, None]
        sGY = (G * yh).sum((0, 2, 3))
        if dmix is not None:
            dmix[i] = sGY.sum()
        # B2: dz, dc = W3^T dz, SE partial
        dz = (wi * c['r3'])[None, :, None, None] * (G - (sG / Q)[None, :, None, None] - yh * (sGY / Q)[None, :, None, None])
        dc = torch.einsum('oc,nohw->nchw', wt['w3'], dz)
        dh = (c['d'] - c['mu2'][None, :, None, None]) * c['r2'][None, :, None, None]
        b = act_f(dh, act)
        g_ = {}
        if 'rw' in wt:
            g = c['g']
            cc = b * g[:, :, None, None]
            dg = (dc * b).sum((2, 3))                                   # [N, mc]
            de = dg * g * (1 - g)
            h = act_f(c['t'], act)
            dhid = de @ wt['ew']                                        # [N, se]
            dt = dhid * act_df(c['t'], act)
            dp = dt @ wt['rw']                                          # [N, mc]
            db = dc * g[:, :, None, None] + dp[:, :, None, None] / (Ho * Wo)
            if want_wgrad:
                g_.update(ew=de.t() @ h, eb=de.sum(0), rw=dt.t() @ c['p'], rb=dt.sum(0))
        else:
            cc = b
            db = dc
        if want_wgrad:
            g_['w3'] = torch.einsum('nohw,nchw->oc', dz, cc)
        # B2b: BN2-backward statistics
        ddh = db * act_df(dh, act)
        sD1 = ddh.sum((0, 2, 3))
        sD2 = (ddh * dh).sum((0, 2, 3))
        # B3: dd, transposed depthwise, recompute u-hat, du-hat, stats, main dx GEMM
        dd = c['r2'][None, :, None, None] * (ddh - (sD1 / Q)[None, :, None, None] - dh * (sD2 / Q)[None, :, None, None])
        u = torch.einsum('ck,nkhw->nchw', wt['w1'], x)
        uh = (u - c['mu1'][None, :, None, None]) * c['r1'][None, :, None, None]
        a = act_f(uh, act)
        pad = k // 2
        opad = (H + 2 * pad - k) % stride  # output_padding so that the transposed conv returns H
        da = F.conv_transpose2d(dd, wt['dw'], None, stride, pad, opad, mc)
        duh = da * act_df(uh, act)
        sU1 = duh.sum((0, 2, 3))
        sU2 = (duh * uh).sum((0, 2, 3))
        dx_main += torch.einsum('ck,nchw->nkhw', wt['w1'] * c['r1'][:, None], duh)
        m1, m2 = sU1 / Pn, sU2 / Pn
        cvec += wt['w1'].t() @ (c['r1'] * m1)
        Mm += wt['w1'].t() @ ((c['r1'] ** 2 * m2)[:, None] * wt['w1'])
        if want_wgrad:
            # depthwise weight grad: correlate dd with a
            ap = F.pad(a, (pad, pad, pad, pad))
            gdw = torch.zeros_like(wt['dw'])
            for ky in range(k):
                for kx in range(k):
                    sl = ap[:, :, ky:ky + (Ho - 1) * stride + 1:stride, kx:kx + (Wo - 1) * stride + 1:stride]
                    gdw[:, 0, ky, kx] = (dd * sl).sum((0, 2, 3))
            g_['dw'] = gdw
            Smat = torch.einsum('nchw,nkhw->ck', duh, x)
            g_['w1'] = c['r1'][:, None] * (Smat - (m1 * Pn)[:, None] * mu_x[None, :]
                                           - (m2 * c['r1'] * Pn)[:, None] * (wt['w1'] @ cov))
            wg[i] = g_
    # B4: finalize dx
    xc = x - mu_x[None, :, None, None]
    dx = dx_main - cvec[None, :, None, None] - torch.einsum('kj,njhw->nkhw', Mm, xc)
    if S['residual']:
        dx = dx + G
    return dx, dmix, wg


def alpha_grad(dmix, mix_w, lats, dlat, T):
    """dL/dlog_alpha from dL/dw (softmax((log_alpha+g)/T) Jacobian) incl. the latency term."""
    dw = dmix + dlat * lats
    return mix_w * (dw - (dw * mix_w).sum()) / T
