"""GPU parity of the full supernet (drop-in Network) against the golden vectors produced by the
real reference: alpha-step logits / latency / alpha- and beta-grads, and the bi-sampled w-step."""
import os
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_shim
from tests import golden_inputs as gi
from tests import helpers as H
from tfnas_b200 import config
from tfnas_b200.model_search import Network, NoisePlan, injected

pytestmark = pytest.mark.gpu
TOL = 1e-3   # north-star tolerance on logits and alpha-grads (norm-relative)


def _net():
    mcs = config.get_mc_num_dddict(config.mc_mask_dddict)
    P, x, tgt = gi.network_inputs()
    net = Network(100, mcs, gi.load_lut())
    net.load_state_dict(P)
    net.set_temperature(5.0)
    return net.cuda().train(), x.cuda(), tgt.cuda()


def test_alpha_step_matches_reference_golden():
    z = np.load(os.path.join(gi.GOLDEN_DIR, 'network_alpha.npz'))
    net, x, tgt = _net()
    for p in net.weight_parameters():
        p.requires_grad_(False)
    with injected(NoisePlan(noise=ref_shim.draw_plan_noise(gi.NET['noise_seed']))):
        logits, lat = net(x, sampling=False)
    loss = F.cross_entropy(logits, tgt) + torch.abs(lat / gi.NET['target_lat'] - 1.) * gi.NET['lambda_lat']
    loss.backward()
    npar = dict(net.named_parameters())
    da = torch.stack([npar[str(n)].grad for n in z['alpha_names']])
    db = torch.cat([npar[str(n)].grad for n in z['beta_names']])
    e = dict(logits=H.rel_l2(logits, torch.from_numpy(z['logits'])), lat=abs(float(lat) - float(z['lat'])),
             loss=abs(float(loss) - float(z['loss'])), dalpha=H.rel_l2(da, torch.from_numpy(z['dalpha'])),
             dalpha_max=H.rel_max(da, torch.from_numpy(z['dalpha'])), dbeta=H.rel_l2(db, torch.from_numpy(z['dbeta'])))
    print('alpha-step errors', e)
    assert e['logits'] < TOL and e['lat'] < 1e-4 and e['loss'] < 1e-4
    assert e['dalpha'] < TOL and e['dalpha_max'] < TOL and e['dbeta'] < TOL


def _port_wstep_grads(idx_g, idx_r):
    """Every live weight gradient of the same bi-sampled w-step from the CPU oracle port (bit-exact with the reference)."""
    from oracle import port
    mcs = config.get_mc_num_dddict(config.mc_mask_dddict)
    P, x, tgt = gi.network_inputs()
    Pg = {k: v.clone().requires_grad_(not port.is_arch_key(k)) for k, v in P.items()}
    lg, _ = port.network_forward(x, Pg, mcs, gi.load_lut(), True, indices=idx_g)
    lr, _ = port.network_forward(x, Pg, mcs, gi.load_lut(), True, indices=idx_r)
    (F.cross_entropy(lg, tgt) + F.cross_entropy(lr, tgt)).backward()
    return {k: v.grad for k, v in Pg.items() if v.grad is not None}


def test_bisampled_wstep_matches_reference_golden():
    """Bi-sampled w-step at bs 2 against the REAL reference's vectors: logits, sampled indices, and EVERY live weight
    gradient element-wise -- through four seeded random projections of the reference's gradients (golden fixture) and
    directly against the CPU port (which the CPU suite pins to the same fixture).  Tolerance 1e-3 (north star), fp32
    everywhere incl. the cuDNN/cuBLAS backward of the stems / head (TF32 off, tfnas_b200/__init__.py)."""
    z = np.load(os.path.join(gi.GOLDEN_DIR, 'network_wstep.npz'))
    assert not torch.backends.cudnn.allow_tf32 and not torch.backends.cuda.matmul.allow_tf32
    net, x, tgt = _net()
    for p in net.arch_parameters():
        p.requires_grad_(False)
    random.seed(gi.NET['py_seed'])
    noise = ref_shim.draw_plan_noise(gi.NET['wstep_noise_seed'])
    mops = [m for m in net.modules() if hasattr(m, 'switches')]
    with injected(NoisePlan(noise=noise)):
        lg, zero = net(x, sampling=True, mode='gumbel')
    idx_g = [m.switches.index(False) for m in mops]
    sampled = []
    from tfnas_b200 import model_search as ms_mod
    orig = ms_mod.MixedOP._sample_index

    def spy(self, mode):
        i = orig(self, mode)
        sampled.append(i)
        return i
    ms_mod.MixedOP._sample_index = spy
    try:
        lr, _ = net(x, sampling=True, mode='random')
    finally:
        ms_mod.MixedOP._sample_index = orig
    idx_r = list(sampled)
    assert zero == 0.0 and idx_g == [int(v) for v in z['idx_g']]
    (F.cross_entropy(lg, tgt) + F.cross_entropy(lr, tgt)).backward()
    assert H.rel_l2(lg, torch.from_numpy(z['logits_g'])) < TOL
    assert H.rel_l2(lr, torch.from_numpy(z['logits_r'])) < TOL
    npar = dict(net.named_parameters())
    ref = _port_wstep_grads(idx_g, idx_r)
    worst_n = worst_p = worst_e = 0.0
    for j, (n, gn) in enumerate(zip(z['wnames'], z['gnorm'])):
        g = npar[str(n)].grad
        if gn < 0:
            assert g is None, n
            continue
        assert g is not None, n
        worst_n = max(worst_n, abs(float(g.norm()) - gn) / (gn + 1e-12))
        worst_p = max(worst_p, float(np.abs(gi.grad_projections(g, j) - z['gproj'][j]).max()) / (gn + 1e-12))
        worst_e = max(worst_e, H.rel_l2(g, ref[str(n)]))
    print('w-step worst over %d live tensors: grad-norm %.2e, projection %.2e, element-wise vs port %.2e'
          % (int((z['gnorm'] >= 0).sum()), worst_n, worst_p, worst_e))
    assert worst_n < TOL and worst_p < TOL and worst_e < TOL
    assert H.rel_l2(npar['first_stem.conv.weight'].grad, torch.from_numpy(z['g_first_stem'])) < TOL
    assert H.rel_l2(npar['classifier.linear.weight'].grad, torch.from_numpy(z['g_classifier'])) < TOL
    assert all(all(m.switches) for m in net.modules() if hasattr(m, 'switches'))


def test_two_stream_wstep_equals_single_stream():
    """search_loop.w_step with the two sampled passes on two CUDA streams gives the same loss, sampled paths and weight
    update as the sequential single-stream step."""
    import torch.nn as nn
    from tfnas_b200 import model_search
    from tfnas_b200.search_loop import make_optimizers, w_step

    class Wrap(nn.Module):          # the loops address the network as model.module (nn.DataParallel in the reference)
        def __init__(self, m):
            super().__init__()
            self.module = m

        def forward(self, *a, **k):
            return self.module(*a, **k)

    res = []
    for overlap in (False, True):
        net, x, tgt = _net()
        model = Wrap(net)
        opt_w, _ = make_optimizers(net)
        random.seed(7)
        model_search.seed_noise(13)
        loss, logits = w_step(model, x, tgt, nn.CrossEntropyLoss().cuda(), opt_w, 5.0, None, bisample=True, overlap=overlap)
        torch.cuda.synchronize()
        flat = torch.cat([p.detach().reshape(-1) for p in net.weight_parameters()])
        res.append((float(loss), logits.detach().clone(), flat.clone()))
    model_search.seed_noise(None)
    (l0, g0, w0), (l1, g1, w1) = res
    e = dict(loss=abs(l0 - l1), logits=H.rel_l2(g1, g0), weights=H.rel_l2(w1, w0))
    print('two-stream vs single-stream w-step', e)
    assert e['loss'] < 1e-4 and e['logits'] < 1e-4 and e['weights'] < 1e-5
