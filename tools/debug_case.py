import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import fused_math as fm
from tests import helpers as H
ic, oc, s, act, size, N, idx = 16, 24, 2, 'relu', 16, 2, int(sys.argv[1]) if len(sys.argv) > 1 else 3
mcs = H.default_mcs(ic, False)
P, x, gum, lats = H.make_problem(ic, oc, s, size, N, mcs, seed=3 * ic + size)
dt = torch.float64
Pd = {k: v.to(dt) for k, v in P.items()}
cands = [fm.cand_weights(Pd, 'b.', i) for i in range(8)]
G = torch.randn(N, oc, size // s, size // s, generator=torch.Generator().manual_seed(6))
o_ref, S1 = fm.forward(x.to(dt), cands, [idx], s, act, None)
dx1, _, wg = fm.backward(x.to(dt), cands, [idx], s, act, S1, G.to(dt), None, True)
for rep in range(3):
    r1 = H.raw_call(P, x, gum, lats, ic, oc, s, act, mcs, 1 << idx, 5.0, G, 0.0, want_wgrad=True)
    print('GEMM=%s idx %d rep %d: out %.2e dx %.2e' % (os.environ.get('TFNAS_GEMM', 'umma'), idx, rep, H.rel_l2(r1['out'], o_ref), H.rel_l2(r1['dx'], dx1)),
          {k[1]: '%.1e' % H.rel_l2(v.reshape(-1), wg[idx][dict(w1='w1', dw='dw', w3='w3', se_rw='rw', se_rb='rb', se_ew='ew', se_eb='eb')[k[1]]].reshape(-1)) for k, v in r1['wgrads'].items()})
d = (r1['dx'].double() - dx1)
print('dx diff per channel mean', d.mean((0, 2, 3)).tolist()[:6], 'std', d.std().item())
g1 = r1['wgrads'][(idx, 'w1')].double().reshape(wg[idx]['w1'].shape)
rowerr = (g1 - wg[idx]['w1']).norm(dim=1) / wg[idx]['w1'].norm(dim=1)
print('dW1 row rel err:', ['%.0e' % v for v in rowerr.tolist()])

# recompute the reference intermediates of the backward (fp64) for candidate idx
wt, c = cands[idx], S1['c'][idx]
Gd, xd = G.to(dt), x.to(dt)
import torch.nn.functional as F
Q = N * (size // s) ** 2; Pn = N * size * size
sG = Gd.sum((0, 2, 3))
yh = (c['z'] - c['mu3'][None, :, None, None]) * c['r3'][None, :, None, None]
sGY = (Gd * yh).sum((0, 2, 3))
dz = c['r3'][None, :, None, None] * (Gd - (sG / Q)[None, :, None, None] - yh * (sGY / Q)[None, :, None, None])
dc = torch.einsum('oc,nohw->nchw', wt['w3'], dz)
dh = (c['d'] - c['mu2'][None, :, None, None]) * c['r2'][None, :, None, None]
ddh = dc * fm.act_df(dh, act)
sD = torch.stack([ddh.sum((0, 2, 3)), (ddh * dh).sum((0, 2, 3))], 1)
dd = c['r2'][None, :, None, None] * (ddh - (sD[:, 0] / Q)[None, :, None, None] - dh * (sD[:, 1] / Q)[None, :, None, None])
k = 5 if idx in (2, 3, 6, 7) else 3
mc = wt['w1'].shape[0]
da = F.conv_transpose2d(dd, wt['dw'], None, s, k // 2, (size + 2 * (k // 2) - k) % s, mc)
u = torch.einsum('ck,nkhw->nchw', wt['w1'], xd)
uh = (u - c['mu1'][None, :, None, None]) * c['r1'][None, :, None, None]
duh = da * fm.act_df(uh, act)
sU = torch.stack([duh.sum((0, 2, 3)), (duh * uh).sum((0, 2, 3))], 1)
def rowerr(a, b):
    a = a.double().reshape(a.shape[0] if a.dim() == 2 else -1, -1) if a.dim() == 2 else a.double().permute(1, 0, 2, 3).reshape(a.shape[1], -1)
    b = b.double().reshape(b.shape[0] if b.dim() == 2 else -1, -1) if b.dim() == 2 else b.double().permute(1, 0, 2, 3).reshape(b.shape[1], -1)
    return ((a - b).norm(dim=1) / (b.norm(dim=1) + 1e-30))
for name, got, ref in [('DC(ddh)', r1['DC'], ddh), ('sD', r1['sD'], sD), ('DA', r1['DA'], da), ('sU', r1['sU'], sU)]:
    e = rowerr(got, ref)
    bad = (e > 1e-4).nonzero().flatten().tolist()
    print('%-8s max row err %.1e  bad rows %s' % (name, e.max().item(), bad[:10]), [('%.2e' % e[i]) for i in bad[:5]])
