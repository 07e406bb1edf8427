"""Generate the golden fixtures by running the REAL reference (read-only, /root/reference).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Runs only in the build container.  Outputs (committed):
  lut_gpu.npz            the latency LUT (reference latency_pkl/latency_gpu.pkl) as flat arrays
  mixedop_cfg1.npz       BASELINE config 1: stage2.block1 MixedOP, bs=2, 32x32, alpha mode fwd+bwd
  network_alpha.npz      full supernet alpha-step (bs=2, 224x224): logits, lat, alpha/beta grads
  network_wstep.npz      bi-sampled w-step (gumbel + random path): logits, indices, grad norms and four seeded random
                         projections of every live weight gradient
Inputs are regenerated in the tests from the recorded seeds with the same seeded generators.
"""
import os
import random
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import ref_shim, port  # noqa: E402
from tfnas_b200 import config  # noqa: E402
from tests import golden_inputs as gi  # noqa: E402


def dump_lut():
    lut = ref_shim.load_lut('latency_gpu.pkl')
    keys = [k for k in lut if k != 'base']
    lens = np.array([len(lut[k]) for k in keys], dtype=np.int32)
    vals = np.concatenate([np.array([lut[k][m] for m in range(1, len(lut[k]) + 1)], dtype=np.float64) for k in keys])
    for k in keys:
        assert list(lut[k].keys()) == list(range(1, len(lut[k]) + 1))
    np.savez_compressed(os.path.join(HERE, 'lut_gpu.npz'), base=np.float64(lut['base']), keys=np.array(keys),
                        lens=lens, vals=vals)
    return lut


def mixedop_cfg1(lut):
    ms, _cfg, _pm = ref_shim._import()
    ic, oc, s, act, size, N = gi.CFG1['ic'], gi.CFG1['oc'], gi.CFG1['stride'], gi.CFG1['act'], gi.CFG1['size'], gi.CFG1['N']
    mcd = config.get_mc_num_dddict(config.mc_mask_dddict)['stage2']['block1']
    lut1 = gi.patched_lut_cfg1(lut)
    P, x, tgt_G, seed_noise = gi.cfg1_inputs()
    op = ms.MixedOP(ic, oc, s, False, act, 8, mcd, lut1)
    op.set_temperature(5.0)
    op.train()
    sd = {k[len('b.'):]: v for k, v in P.items()}
    op.load_state_dict(sd)
    xr = x.clone().requires_grad_(True)
    with ref_shim.InjectedNoise(seed_noise):
        out, lat = op(xr, False, 'max')
    ((out * tgt_G).sum() + lat * 0.37).backward()
    np.savez_compressed(os.path.join(HERE, 'mixedop_cfg1.npz'), out=out.detach().numpy(), lat=np.float32(lat.item()),
                        dx=xr.grad.numpy(), dalpha=op.log_alphas.grad.numpy())
    print('cfg1 lat', float(lat))


def network_steps(lut):
    mcs = config.get_mc_num_dddict(config.mc_mask_dddict)
    P, x, tgt = gi.network_inputs()
    net = ref_shim.build_network(mcs, lut, seed=2)
    net.load_state_dict(P)
    # alpha step
    for p in net.weight_parameters():
        p.requires_grad_(False)
    with ref_shim.InjectedNoise(gi.NET['noise_seed']):
        logits, lat = net(x, sampling=False)
    loss = F.cross_entropy(logits, tgt) + torch.abs(lat / gi.NET['target_lat'] - 1.) * gi.NET['lambda_lat']
    loss.backward()
    names = [n for n, _ in net.named_parameters() if n.endswith('log_alphas')]
    bnames = [n for n, _ in net.named_parameters() if n.endswith('betas')]
    npar = dict(net.named_parameters())
    np.savez_compressed(os.path.join(HERE, 'network_alpha.npz'), logits=logits.detach().numpy(),
                        lat=np.float32(lat.item()), loss=np.float32(loss.item()),
                        dalpha=np.stack([npar[n].grad.numpy() for n in names]),
                        dbeta=np.concatenate([npar[n].grad.numpy() for n in bnames]),
                        alpha_names=np.array(names), beta_names=np.array(bnames))
    print('alpha step lat', float(lat), 'loss', float(loss))
    # bi-sampled w step
    for p in net.parameters():
        p.grad = None
    for p in net.weight_parameters():
        p.requires_grad_(True)
    for p in net.arch_parameters():
        p.requires_grad_(False)
    random.seed(gi.NET['py_seed'])
    with ref_shim.InjectedNoise(gi.NET['wstep_noise_seed']):
        lg, _ = net(x, sampling=True, mode='gumbel')
        idx_g = [int(np.argmin(m.switches)) for m in net.modules() if isinstance(m, type(net.stage1.block1))]
        lr, _ = net(x, sampling=True, mode='random')
    loss_w = F.cross_entropy(lg, tgt) + F.cross_entropy(lr, tgt)
    loss_w.backward()
    wn = [n for n, _ in net.named_parameters() if not (n.endswith('log_alphas') or n.endswith('betas'))]
    gnorm = np.array([float(npar[n].grad.norm()) if npar[n].grad is not None else -1.0 for n in wn], dtype=np.float64)
    sel = ['first_stem.conv.weight', 'classifier.linear.weight']
    # element-level pin of EVERY live gradient without committing 35 MB: four seeded random projections per tensor
    # (a permuted, transposed or sign-flipped dW changes them by O(|g|))
    gproj = np.stack([gi.grad_projections(npar[n].grad, j) if npar[n].grad is not None else np.zeros(gi.NPROJ)
                      for j, n in enumerate(wn)])
    np.savez_compressed(os.path.join(HERE, 'network_wstep.npz'), logits_g=lg.detach().numpy(), logits_r=lr.detach().numpy(),
                        idx_g=np.array(idx_g), loss=np.float32(loss_w.item()), gnorm=gnorm, gproj=gproj, wnames=np.array(wn),
                        g_first_stem=npar[sel[0]].grad.numpy(), g_classifier=npar[sel[1]].grad.numpy())
    print('w step loss', float(loss_w), 'live grads', int((gnorm >= 0).sum()))


if __name__ == '__main__':
    assert ref_shim.available()
    lut = dump_lut()
    mixedop_cfg1(lut)
    network_steps(lut)
    print('golden fixtures written to', HERE)
