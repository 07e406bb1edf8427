"""train_search.py end to end on the GPU with synthetic data (reference CLI surface, train_search.py:155-315): warm-up epochs
(train_wo_arch), search epochs (train_w_arch + elastic width rescaling), validation in the last epochs, one checkpoint per
epoch in the reference's format."""
import glob
import importlib.util
import os

import numpy as np
import pytest
import torch

from tests import golden_inputs as gi
from tfnas_b200 import config, parsing

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_train_search_cli_synthetic(tmp_path):
    spec = importlib.util.spec_from_file_location('ts_cli_run', os.path.join(ROOT, 'train_search.py'))
    ts = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ts)
    lut = os.path.join(gi.GOLDEN_DIR, 'lut_gpu.npz')
    ts.main(['--synthetic', '4', '--epochs', '12', '--warm_epochs', '10', '--batch_size', '8', '--print_freq', '2',
             '--save', str(tmp_path), '--lookup_path', lut, '--note', 'pytest'])
    runs = glob.glob(os.path.join(str(tmp_path), 'search-*-pytest'))
    assert len(runs) == 1
    ckpts = sorted(glob.glob(os.path.join(runs[0], 'searched_model_*.pth.tar')))
    assert [os.path.basename(c) for c in ckpts] == ['searched_model_%02d.pth.tar' % i for i in range(13)]
    first = torch.load(ckpts[0], weights_only=False)
    last = torch.load(ckpts[-1], weights_only=False)
    # reference format: max-width tensors under 'module.'-prefixed names + the channel masks
    assert set(last) == {'state_dict', 'mc_mask_dddict'} and all(k.startswith('module.') for k in last['state_dict'])
    assert list(first['state_dict']) == list(last['state_dict'])
    mx = config.get_mc_num_dddict(config.mc_mask_dddict, is_max=True)
    assert last['state_dict']['module.stage3.block2.m_ops.1.inverted_bottleneck.conv.weight'].shape[0] == mx['stage3']['block2'][1]
    # weights moved, architecture parameters moved only in the two search epochs and stay normalised (log_softmax renorm)
    w0, w1 = first['state_dict']['module.first_stem.conv.weight'], last['state_dict']['module.first_stem.conv.weight']
    assert float((w0 - w1).abs().max()) > 0
    la = last['state_dict']['module.stage2.block1.log_alphas']
    assert abs(float(torch.exp(la).sum()) - 1.0) < 1e-5 and float((la - first['state_dict']['module.stage2.block1.log_alphas']).abs().max()) > 0
    assert all(torch.isfinite(v).all() for v in last['state_dict'].values())
    # the checkpoint parses into a legal architecture with our parser (the reference parser is checked on CPU, test_search_host)
    op_w, depth_w = parsing.get_op_and_depth_weights(ckpts[-1])
    arch = parsing.parse_architecture(op_w, depth_w)
    assert list(arch) == ['stage%d' % i for i in range(1, 7)] and all(0 <= v < 8 for st in arch.values() for v in st.values())
    log = open(os.path.join(runs[0], 'log.txt')).read()
    assert 'TRAIN wo_Arch' in log and 'TRAIN w_Arch' in log and 'VALIDATE' in log and 'Now shrinking or expanding the arch' in log
