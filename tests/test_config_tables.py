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


def test_lut_builder_covers_exactly_the_reference_table():
    """tools/make_lat_lut.py (B200 latency table) measures the same 66 keys with the same width ranges as the
    reference's shipped latency_gpu.pkl."""
    import importlib.util
    import os
    from tests import golden_inputs as gi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('make_lat_lut', os.path.join(root, 'tools', 'make_lat_lut.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    table = m.key_table()
    lut = gi.load_lut()
    assert list(table) == [k for k in lut if k != 'base']
    assert all(table[k][-1] == len(lut[k]) for k in table)


def test_b200_latency_table_loads_and_covers_the_search_space():
    """profiles/latency_b200.npz (built by tools/make_lat_lut.py on two B200s) has the reference table's structure, so
    `train_search.py --lookup_path profiles/latency_b200.npz` works, and every candidate of every MixedOP finds its row."""
    import os
    from tests import golden_inputs as gi
    from tfnas_b200 import config
    from tfnas_b200.lut import load_lut
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    b200, ref = load_lut(os.path.join(root, 'profiles', 'latency_b200.npz')), gi.load_lut()
    assert list(b200) == list(ref) and b200['base'] > 0
    assert all(list(b200[k]) == list(ref[k]) for k in ref if k != 'base')
    mcs = config.get_mc_num_dddict(config.mc_mask_dddict)
    for stage, block, *_ in config.block_shapes():
        for i in range(8):
            assert b200[config.lat_lookup_key_dddict[stage][block][i]][mcs[stage][block][i]] > 0
    lat_e6 = b200['base'] + sum(b200[config.lat_lookup_key_dddict[s][b][7]][mcs[s][b][7]] for s, b, *_ in config.block_shapes())
    assert 0.5 < lat_e6 < 50.0          # ms at batch 32 for the widest path; the Titan-RTX table gives ~31
