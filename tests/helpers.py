"""Shared helpers for the parity tests: random MixedOP problems, raw C-ABI calls, error metrics."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tfnas_b200.config import CAND_SPEC  # noqa: E402

NAMES = dict(w1='inverted_bottleneck.conv.weight', dw='depth_conv.conv.weight', w3='point_linear.conv.weight',
             se_rw='squeeze_excite.conv_reduce.weight', se_rb='squeeze_excite.conv_reduce.bias',
             se_ew='squeeze_excite.conv_expand.weight', se_eb='squeeze_excite.conv_expand.bias')
SLOTS = ('w1', 'dw', 'w3', 'se_rw', 'se_rb', 'se_ew', 'se_eb')


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def rel_max(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def make_problem(ic, oc, stride, H, N, mcs, seed=0, dtype=torch.float32, wscale=None, W=None):
    """Random MixedOP problem with reference-named parameter dict (prefix 'b.')."""
    g = torch.Generator().manual_seed(seed)
    P = {}
    pre = 'b.'
    for i, (k, _e, sm) in enumerate(CAND_SPEC):
        mc = mcs[i]
        q = pre + 'm_ops.%d.' % i
        P[q + NAMES['w1']] = torch.randn(mc, ic, 1, 1, generator=g) * (1.0 / ic ** 0.5)
        P[q + NAMES['dw']] = torch.randn(mc, 1, k, k, generator=g) * (1.0 / k)
        P[q + NAMES['w3']] = torch.randn(oc, mc, 1, 1, generator=g) * (1.0 / mc ** 0.5)
        if sm:
            se = sm * ic
            P[q + NAMES['se_rw']] = torch.randn(se, mc, 1, 1, generator=g) * (1.0 / mc ** 0.5)
            P[q + NAMES['se_rb']] = torch.randn(se, generator=g) * 0.1
            P[q + NAMES['se_ew']] = torch.randn(mc, se, 1, 1, generator=g) * (1.0 / se ** 0.5)
            P[q + NAMES['se_eb']] = torch.randn(mc, generator=g) * 0.1
    P[pre + 'log_alphas'] = F.log_softmax(torch.randn(8, generator=g) * 0.5, -1)
    P = {k: v.to(dtype) for k, v in P.items()}
    Wd = H if W is None else W
    x = (torch.randn(N, ic, H, Wd, generator=g) * 1.3 + 0.2 * torch.randn(1, ic, 1, 1, generator=g)).to(dtype)
    gum = -torch.empty(8).exponential_(generator=g).log().to(dtype)
    lats = torch.rand(8, generator=g).to(dtype) * 2
    return P, x, gum, lats


def relu_margin(P, x, ic, stride, act, active):
    """Smallest |pre-activation| feeding a ReLU (fp64).  Two correct fp32 implementations may disagree on the
    gate of an element that sits within rounding of 0, which perturbs BN-backward sums visibly on tiny
    problems; parity tests pick seeds whose margin is comfortably above fp32 rounding."""
    if act != 'relu':
        return 1.0
    from oracle import fused_math as fm
    Pd = {k: v.double() for k, v in P.items()}
    cands = [fm.cand_weights(Pd, 'b.', i) for i in range(8)]
    _o, S = fm.forward(x.double(), cands, list(active), stride, act, torch.full((8,), 0.125, dtype=torch.float64))
    m = 1.0
    for i in active:
        c = S['c'][i]
        u = torch.einsum('ck,nkhw->nchw', cands[i]['w1'], x.double())
        uh = (u - c['mu1'][None, :, None, None]) * c['r1'][None, :, None, None]
        dh = (c['d'] - c['mu2'][None, :, None, None]) * c['r2'][None, :, None, None]
        m = min(m, float(uh.abs().min()), float(dh.abs().min()))
    return m


def make_conditioned_problem(ic, oc, stride, H, N, mcs, seed, act, active=range(8), W=None, margin=5e-6):
    for t in range(50):
        P, x, gum, lats = make_problem(ic, oc, stride, H, N, mcs, seed + 1000 * t, W=W)
        if relu_margin(P, x, ic, stride, act, active) > margin:
            return P, x, gum, lats
    raise RuntimeError('no well-conditioned seed found')


def default_mcs(ic, ragged=False):
    m = [3 * ic, 6 * ic, 3 * ic, 6 * ic, 3 * ic, 6 * ic, 3 * ic, 6 * ic]
    if ragged:
        m = [v - d for v, d in zip(m, (1, 5, 0, 7, 3, 2, 9, 11))]
    return m


# ------------------------------------------------------------------------------------------
# raw C-ABI driver (what a non-Python client would do), used by the GPU parity tests
# ------------------------------------------------------------------------------------------
def raw_call(P, x, gum, lats, ic, oc, stride, act, mcs, mask, T=5.0, G=None, dlat=0.0, want_wgrad=False,
             need_dx=True):
    """Run tfnas_mixedop_fwd (+ _bwd if G given) through ctypes on cuda:0.  Returns dict of results
    incl. the parsed saved-buffer regions."""
    from tfnas_b200 import _lib
    lib = _lib.load()
    dev = torch.device('cuda:0')
    N, _, H, W = x.shape
    d = _lib.MixedOpDesc()
    d.N, d.ic, d.oc, d.H, d.W, d.stride, d.act, d.num_ops = N, ic, oc, H, W, stride, _lib.ACT_CODE[act], 8
    for i, (k, _e, sm) in enumerate(CAND_SPEC):
        d.mc[i], d.k[i], d.se[i] = mcs[i], k, sm * ic
    xd = x.float().to(dev).contiguous()
    wd = {}
    arr = _lib.CandArray()
    garr = _lib.CandArray()
    gd = {}
    active = [i for i in range(8) if mask >> i & 1]
    for i in active:
        for s in SLOTS:
            key = 'b.m_ops.%d.%s' % (i, NAMES[s])
            if key in P:
                t = P[key].float().to(dev).contiguous()
                wd[(i, s)] = t
                setattr(arr[i], s, t.data_ptr())
                if want_wgrad:
                    gt = torch.full_like(t, float('nan'))
                    gd[(i, s)] = gt
                    setattr(garr[i], s, gt.data_ptr())
    la = P['b.log_alphas'].float().to(dev)
    gumd, latd = gum.float().to(dev), lats.float().to(dev)
    nsaved = lib.tfnas_mixedop_saved_bytes(ctypes.byref(d), mask)
    nws = lib.tfnas_mixedop_workspace_bytes(ctypes.byref(d), mask, 1 if want_wgrad else 0)
    assert nsaved > 0, lib.tfnas_last_error()
    saved = torch.zeros(nsaved, dtype=torch.uint8, device=dev)
    ws = torch.zeros(nws, dtype=torch.uint8, device=dev)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    out = torch.full((N, oc, Ho, Wo), float('nan'), device=dev)
    out_lat = torch.zeros((), device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    vp = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    _lib.check(lib.tfnas_mixedop_fwd(ctypes.byref(d), mask, vp(xd), arr, vp(la), vp(gumd), vp(latd), T, vp(out),
                                     vp(out_lat), vp(saved), nsaved, vp(ws), nws, st))
    torch.cuda.synchronize()
    r = dict(out=out.cpu(), out_lat=float(out_lat))
    offs = (ctypes.c_size_t * 13)()
    _lib.check(lib.tfnas_debug_saved_layout(ctypes.byref(d), mask, offs))
    na = len(active)
    MC = sum(mcs[i] for i in active)
    MCse = sum(mcs[i] for i in active if CAND_SPEC[i][2])
    SEH = sum(CAND_SPEC[i][2] * ic for i in active)

    def region(idx, n, dt=torch.float32):
        nb = n * (8 if dt == torch.float64 else 4)
        return saved[offs[idx]:offs[idx] + nb].view(dt).cpu()
    r['xmom'] = region(0, ic + ic * ic, torch.float64)
    r['bn1'] = region(1, 2 * MC)
    r['bn2'] = region(2, 2 * MC)
    r['bn3'] = region(3, 2 * na * oc)
    r['mixw'] = region(4, 8)
    r['sep'] = region(6, N * MCse).view(N, MCse) if MCse else None
    r['set'] = region(7, N * SEH).view(N, SEH) if SEH else None
    r['seg'] = region(8, N * MCse).view(N, MCse) if MCse else None
    r['UH'] = region(9, N * MC * H * W).view(N, MC, H, W)
    r['D'] = region(10, N * MC * Ho * Wo).view(N, MC, Ho, Wo)
    r['Z'] = region(11, N * na * oc * Ho * Wo).view(N, na, oc, Ho, Wo)
    r['active'] = active
    if G is not None:
        Gd = G.float().to(dev).contiguous()
        dx = torch.full_like(xd, float('nan')) if need_dx else None
        dal = torch.full((8,), float('nan'), device=dev)
        dl = torch.tensor(float(dlat), device=dev)
        ws.zero_()
        _lib.check(lib.tfnas_mixedop_bwd(ctypes.byref(d), mask, vp(xd), arr, vp(Gd), vp(dl), T, vp(saved), nsaved,
                                         vp(dx), vp(dal), garr if want_wgrad else None, vp(ws), nws, st))
        torch.cuda.synchronize()
        boffs = (ctypes.c_size_t * 10)()
        _lib.check(lib.tfnas_debug_bwd_layout(ctypes.byref(d), mask, 1 if want_wgrad else 0, boffs))

        def wreg(idx, n, dt=torch.float32):
            nb = n * (8 if dt == torch.float64 else 4)
            return ws[boffs[idx]:boffs[idx] + nb].view(dt).cpu()
        r['sD'] = wreg(2, 2 * MC, torch.float64).view(MC, 2)
        r['sU'] = wreg(3, 2 * MC, torch.float64).view(MC, 2)
        r['DC'] = wreg(7, N * MC * Ho * Wo).view(N, MC, Ho, Wo)
        r['DA'] = wreg(8, N * MC * H * W).view(N, MC, H, W)
        r['dx'] = dx.cpu() if need_dx else None
        r['dalpha'] = dal.cpu()
        r['wgrads'] = {k: v.cpu() for k, v in gd.items()}
    return r


def oracle_alpha(P, x, gum, lats, ic, oc, stride, act, T=5.0, G=None, dlat=0.0, dtype=torch.float64):
    """oracle/port.py (literal restatement) in `dtype`, autograd backward."""
    from oracle import port
    Pd = {k: v.to(dtype).clone().requires_grad_(True) for k, v in P.items()}
    xd = x.to(dtype).clone().requires_grad_(True)
    out, lat = port.mixedop_alpha(xd, Pd, 'b.', ic, oc, stride, act, T, gum.to(dtype), [float(v) for v in lats])
    r = dict(out=out.detach(), out_lat=float(lat.detach()))
    if G is not None:
        ((out * G.to(dtype)).sum() + lat * dlat).backward()
        r['dx'] = xd.grad
        r['dalpha'] = Pd['b.log_alphas'].grad
    return r


def oracle_single(P, x, ic, oc, stride, act, idx, G=None, dtype=torch.float64):
    from oracle import port
    Pd = {k: v.to(dtype).clone().requires_grad_(True) for k, v in P.items()}
    xd = x.to(dtype).clone().requires_grad_(True)
    out = port.mixedop_single(xd, Pd, 'b.', ic, oc, stride, act, idx)
    r = dict(out=out.detach())
    if G is not None:
        (out * G.to(dtype)).sum().backward()
        r['dx'] = xd.grad
        r['wgrads'] = {(idx, s): Pd['b.m_ops.%d.%s' % (idx, NAMES[s])].grad for s in SLOTS
                       if 'b.m_ops.%d.%s' % (idx, NAMES[s]) in Pd}
    return r
