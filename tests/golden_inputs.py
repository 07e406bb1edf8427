"""Seeded inputs shared by tests/golden/make_golden.py (which runs the real reference) and the
tests that replay the same inputs through the oracle port and the CUDA path."""
import os
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from tfnas_b200 import config
from tfnas_b200.config import CAND_SPEC, lut_key

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# BASELINE.json configs[0]: single MixedOP cell (stage2 block1, 8 candidates), bs=2, 32x32
CFG1 = dict(ic=24, oc=40, stride=2, act='swish', size=32, N=2, seed=123, noise_seed=321)
NET = dict(N=2, size=224, seed=7, noise_seed=123, wstep_noise_seed=77, py_seed=5, target_lat=15.0, lambda_lat=0.1)


def load_lut():
    """LUT fixture -> the same dict-of-dicts structure as the reference pickle."""
    from tfnas_b200.lut import load_lut as _load
    return _load(os.path.join(GOLDEN_DIR, 'lut_gpu.npz'))


def patched_lut_cfg1(lut):
    """the shipped LUT has no 32x32 keys; alias the 56x56 stage2.block1 rows."""
    out = dict(lut)
    for (k, _e, sm) in CAND_SPEC:
        out[lut_key(32, 24, sm * 24, 40, k, 2, 'swish')] = lut[lut_key(56, 24, sm * 24, 40, k, 2, 'swish')]
    return out


def cfg1_inputs():
    from tests import helpers as H
    mcd = config.get_mc_num_dddict(config.mc_mask_dddict)['stage2']['block1']
    mcs = [mcd[i] for i in range(8)]
    P, x, _gum, _lats = H.make_problem(CFG1['ic'], CFG1['oc'], CFG1['stride'], CFG1['size'], CFG1['N'], mcs, CFG1['seed'])
    g = torch.Generator().manual_seed(CFG1['seed'] + 1)
    G = torch.randn(CFG1['N'], CFG1['oc'], CFG1['size'] // 2, CFG1['size'] // 2, generator=g)
    return P, x, G, CFG1['noise_seed']


def network_inputs():
    from oracle import port
    mcs = config.get_mc_num_dddict(config.mc_mask_dddict)
    P = port.init_params(mcs, seed=2)
    g = torch.Generator().manual_seed(NET['seed'])
    for k in P:
        if k.endswith('log_alphas'):
            P[k] = F.log_softmax(P[k] + 0.3 * torch.randn(8, generator=g), -1)
        elif k.endswith('betas'):
            P[k] = 0.2 * torch.randn(P[k].shape, generator=g)
    x = torch.randn(NET['N'], 3, NET['size'], NET['size'], generator=g)
    tgt = torch.randint(0, 100, (NET['N'],), generator=g)
    return P, x, tgt


NPROJ = 4


def grad_projections(g, j):
    """<g, r_k> for NPROJ seeded N(0,1) tensors r_k (seed derived from the tensor's index j in the weight list); fp64."""
    gen = torch.Generator().manual_seed(100003 + int(j))
    r = torch.randn(NPROJ, g.numel(), generator=gen, dtype=torch.float64)
    return (r @ g.detach().double().cpu().reshape(-1)).numpy()


def derived_arch():
    """A parsed architecture for the derived-network tests: stage2 and stage4 one block short, all eight candidates used."""
    from collections import OrderedDict
    depth = dict(stage1=2, stage2=2, stage3=4, stage4=3, stage5=4, stage6=1)
    arch, i = OrderedDict(), 0
    for s, k in depth.items():
        arch[s] = OrderedDict()
        for j in range(k):
            arch[s]['block%d' % (j + 1)] = (5 * i + 3) % 8
            i += 1
    return arch


def derived_inputs():
    g = torch.Generator().manual_seed(11)
    return torch.randn(4, 3, 64, 64, generator=g), torch.randn(3, 3, 64, 64, generator=g)
