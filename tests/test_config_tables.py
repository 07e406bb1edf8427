import hashlib

import pytest
import torch

from oracle import ref_shim
from tfnas_b200 import config


def _digest():
    h = hashlib.sha256()
    for st, bl, ic, oc, s, act, size in config.block_shapes():
        for i in range(8):
            m = config.mc_mask_dddict[st][bl][i]
            h.update(('%s %s %d %d %d %s|' % (st, bl, i, m.numel(), int(m.sum()), config.lat_lookup_key_dddict[st][bl][i])).encode())
    return h.hexdigest()


def test_tables_shape_and_digest():
    shapes = list(config.block_shapes())
    assert len(shapes) == 18
    assert shapes[0][2:] == (16, 24, 2, 'relu', 112) and shapes[-1][2:] == (192, 320, 1, 'swish', 7)
    mcs = config.get_mc_num_dddict(config.mc_mask_dddict)
    mx = config.get_mc_num_dddict(config.mc_mask_dddict, is_max=True)
    assert mcs['stage1']['block1'][0] == 48 and mx['stage1']['block1'][1] == 128
    assert config.lat_lookup_key_dddict['stage6']['block1'][7] == 'MBInvertedResBlock_7_192_384_320_k5_s1_swish'
    assert _digest() == '01d8a2155c38e1de7ead0d503f35c9eb5b3bc5ea8854682b53626a88e0b4a0f8'


@pytest.mark.skipif(not ref_shim.available(), reason='reference not mounted')
def test_tables_equal_reference():
    _ms, cfg, pm = ref_shim._import()
    for st in cfg.mc_mask_dddict:
        for bl in cfg.mc_mask_dddict[st]:
            for i in cfg.mc_mask_dddict[st][bl]:
                assert torch.equal(cfg.mc_mask_dddict[st][bl][i], config.mc_mask_dddict[st][bl][i])
                assert cfg.lat_lookup_key_dddict[st][bl][i] == config.lat_lookup_key_dddict[st][bl][i]
    assert pm.get_mc_num_dddict(cfg.mc_mask_dddict) == config.get_mc_num_dddict(config.mc_mask_dddict)
    assert pm.get_mc_num_dddict(cfg.mc_mask_dddict, is_max=True) == config.get_mc_num_dddict(config.mc_mask_dddict, True)
