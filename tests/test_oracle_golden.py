"""The oracle port must reproduce the committed golden vectors (made by the REAL reference,
tests/golden/make_golden.py) and, where the reference is mounted, the live reference bit-exactly."""
import os
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import port, ref_shim
from tests import golden_inputs as gi
from tests import helpers as H
from tfnas_b200 import config

G = gi.GOLDEN_DIR
TOL = 2e-5   # same torch ops; only the host CPU's kernel selection may differ


def _names():
    return ['%s.%s.' % (s, b) for s, b, *_ in config.block_shapes()]


def test_lut_fixture_shape():
    lut = gi.load_lut()
    assert len(lut) == 67 and abs(lut['base'] - 1.985807514190674) < 1e-12
    for st in config.lat_lookup_key_dddict.values():
        for bl in st.values():
            for key in bl.values():
                assert key in lut


def test_cfg1_mixedop_matches_golden():
    z = np.load(os.path.join(G, 'mixedop_cfg1.npz'))
    lut = gi.patched_lut_cfg1(gi.load_lut())
    P, x, Gt, nseed = gi.cfg1_inputs()
    c = gi.CFG1
    mcd = config.get_mc_num_dddict(config.mc_mask_dddict)['stage2']['block1']
    lats = port.mixedop_lats(lut, c['size'], c['ic'], c['oc'], c['stride'], c['act'], mcd)
    noise = ref_shim.draw_plan_noise(nseed, 1)[0]
    r = H.oracle_alpha(P, x, noise, torch.tensor(lats), c['ic'], c['oc'], c['stride'], c['act'], 5.0, Gt, 0.37,
                       dtype=torch.float32)
    assert H.rel_max(r['out'], torch.from_numpy(z['out'])) < TOL
    assert abs(r['out_lat'] - float(z['lat'])) < 1e-5
    assert H.rel_max(r['dx'], torch.from_numpy(z['dx'])) < TOL
    assert H.rel_max(r['dalpha'], torch.from_numpy(z['dalpha'])) < 1e-4


def test_network_alpha_step_matches_golden():
    z = np.load(os.path.join(G, 'network_alpha.npz'))
    lut = gi.load_lut()
    mcs = config.get_mc_num_dddict(config.mc_mask_dddict)
    P, x, tgt = gi.network_inputs()
    Pg = {k: v.clone().requires_grad_(port.is_arch_key(k)) for k, v in P.items()}
    noise = ref_shim.draw_plan_noise(gi.NET['noise_seed'])
    logits, lat = port.network_forward(x, Pg, mcs, lut, False, 5.0, noise=noise)
    loss, _, _ = port.arch_loss(logits, lat, tgt, gi.NET['target_lat'], gi.NET['lambda_lat'])
    loss.backward()
    assert H.rel_max(logits.detach(), torch.from_numpy(z['logits'])) < TOL
    assert abs(float(lat) - float(z['lat'])) < 1e-4
    da = torch.stack([Pg[str(n)].grad for n in z['alpha_names']])
    db = torch.cat([Pg[str(n)].grad for n in z['beta_names']])
    assert H.rel_max(da, torch.from_numpy(z['dalpha'])) < 1e-3
    assert H.rel_max(db, torch.from_numpy(z['dbeta'])) < 1e-3


def test_network_wstep_matches_golden():
    z = np.load(os.path.join(G, 'network_wstep.npz'))
    lut = gi.load_lut()
    mcs = config.get_mc_num_dddict(config.mc_mask_dddict)
    P, x, tgt = gi.network_inputs()
    noise = ref_shim.draw_plan_noise(gi.NET['wstep_noise_seed'])
    idx_g = [port.sample_gumbel_index(P[n + 'log_alphas'], noise[i]) for i, n in enumerate(_names())]
    assert idx_g == [int(v) for v in z['idx_g']]
    rnd = random.Random(gi.NET['py_seed'])
    idx_r = []
    for ig in idx_g:
        rest = [j for j in range(8) if j != ig]
        idx_r.append(rest[rnd.choice(range(7))])
    Pg = {k: v.clone().requires_grad_(not port.is_arch_key(k)) for k, v in P.items()}
    lg, _ = port.network_forward(x, Pg, mcs, lut, True, indices=idx_g)
    lr, _ = port.network_forward(x, Pg, mcs, lut, True, indices=idx_r)
    (F.cross_entropy(lg, tgt) + F.cross_entropy(lr, tgt)).backward()
    assert H.rel_max(lg.detach(), torch.from_numpy(z['logits_g'])) < TOL
    assert H.rel_max(lr.detach(), torch.from_numpy(z['logits_r'])) < TOL
    for j, (n, gn) in enumerate(zip(z['wnames'], z['gnorm'])):
        g = Pg[str(n)].grad
        if gn < 0:
            assert g is None or float(g.abs().max()) == 0.0
        else:
            assert abs(float(g.norm()) - gn) <= 1e-3 * gn + 1e-7, n
            # element-level pin: seeded random projections of the reference's gradient (scale of <g, r> is |g|)
            assert np.abs(gi.grad_projections(g, j) - z['gproj'][j]).max() <= 1e-4 * gn + 1e-9, n
    assert H.rel_max(Pg['first_stem.conv.weight'].grad, torch.from_numpy(z['g_first_stem'])) < 1e-3


@pytest.mark.skipif(not ref_shim.available(), reason='reference not mounted')
def test_port_bit_exact_vs_live_reference():
    lut = ref_shim.load_lut()
    mcs = config.get_mc_num_dddict(config.mc_mask_dddict)
    P, x, tgt = gi.network_inputs()
    x, tgt = x[:1], tgt[:1]
    net = ref_shim.build_network(mcs, lut, seed=2)
    P0 = port.init_params(mcs, seed=2)
    assert all(torch.equal(v, P0[k]) for k, v in net.state_dict().items())
    net.load_state_dict(P)
    with ref_shim.InjectedNoise(11):
        lo, lat = net(x, sampling=False)
    noise = ref_shim.draw_plan_noise(11)
    lo2, lat2 = port.network_forward(x, P, mcs, lut, False, 5.0, noise=noise)
    assert torch.equal(lo, lo2) and float(lat) == float(lat2)


@pytest.mark.skipif(not ref_shim.available(), reason='reference not mounted')
def test_lut_fixture_equals_reference_pickle():
    ref = ref_shim.load_lut()
    fix = gi.load_lut()
    assert list(ref.keys()) == list(fix.keys())
    for k in ref:
        if k == 'base':
            assert ref[k] == fix[k]
        else:
            assert list(ref[k].items()) == list(fix[k].items())
