"""Host-side pieces of the search controller (no GPU): elastic width scaling and architecture parsing
against the reference's own functions (when mounted), mask-sliced weight movement round trip, and the
lr schedule quirk."""
import copy
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import ref_shim
from tests import golden_inputs as gi
from tfnas_b200 import config, elastic, parsing
from tfnas_b200.model_search import Network
from tfnas_b200.parallel import SearchParallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ts():
    spec = importlib.util.spec_from_file_location('ts_cli', os.path.join(ROOT, 'train_search.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_cli_flags_match_reference_defaults():
    args = _ts().build_parser().parse_args([])
    ref = dict(epochs=90, batch_size=32, w_lr=0.025, w_mom=0.9, w_wd=1e-5, a_lr=0.01, a_wd=5e-4, a_beta1=0.5,
               a_beta2=0.999, grad_clip=5.0, T=5.0, T_decay=0.96, num_classes=100, seed=2, note='try', lambda_lat=0.1,
               target_lat=15.0, print_freq=100, workers=4, lookup_path='./latency_pkl/latency_gpu.pkl', save='./checkpoints')
    for k, v in ref.items():
        assert getattr(args, k) == v, k


def test_lr_list_has_the_reference_quirk():
    lr = _ts().cosine_lr_list(0.025, 90)
    assert len(lr) == 90 and abs(lr[0] - 0.025007616982291203) < 1e-12 and lr[-1] < 1e-5


def _random_arch(seed):
    g = np.random.RandomState(seed)
    op_w = [g.rand(8) for _ in range(18)]
    depth_w = [g.rand(n) for n in (2, 3, 4, 4, 4, 1)]
    return op_w, depth_w


def test_parse_and_rescale_match_reference_or_invariants():
    lut = gi.load_lut()
    keys = config.lat_lookup_key_dddict
    mx = config.get_mc_num_dddict(config.mc_mask_dddict, is_max=True)
    for seed in range(4):
        op_w, depth_w = _random_arch(seed)
        arch = parsing.parse_architecture(op_w, depth_w)
        for si, st in enumerate(arch):
            assert len(arch[st]) == int(np.argmax(depth_w[si])) + 1
        mc = config.get_mc_num_dddict(config.mc_mask_dddict)
        for target in (9.0, 15.0, 18.0):
            new, before, after = elastic.rescale_widths(arch, copy.deepcopy(mc), mx, keys, lut, target)
            assert abs(elastic.get_lookup_latency(arch, new, keys, lut) - after) < 1e-9
            for st in arch:
                for bl, op in arch[st].items():
                    assert mx[st][bl][op] // 2 <= new[st][bl][op] <= mx[st][bl][op]
            if ref_shim.available():
                ref_ts = _ref_functions()
                pm = ref_shim._import()[2]
                assert pm.parse_architecture(op_w, depth_w) == arch
                m2 = copy.deepcopy(mc)
                stages = ['stage%d' % i for i in range(1, 7)]
                if before > target:
                    m2, lat2 = ref_ts['fit'](arch, m2, mx, keys, lut, target, stages, -1)
                else:
                    m2, lat2 = ref_ts['fit'](arch, m2, mx, keys, lut, target, stages, 1)
                for start in range(2, 7):
                    m2, lat2 = ref_ts['fit'](arch, m2, mx, keys, lut, target, ['stage%d' % i for i in range(start, 7)], 1)
                assert m2 == new and abs(lat2 - after) < 1e-12


def _ref_functions():
    """Pull fit_mc_num_by_latency / get_lookup_latency / bound_clip out of the reference train_search.py
    without executing its argparse / mkdir side effects."""
    src = open(os.path.join(ref_shim.REF_ROOT, 'train_search.py')).read()
    start = src.index('def get_lookup_latency(')
    end = src.index("if __name__ == '__main__':")
    ns = {'copy': copy}
    exec(compile(src[start:end], 'ref_train_search_tail', 'exec'), ns)
    return {'fit': ns['fit_mc_num_by_latency'], 'lat': ns['get_lookup_latency']}


def test_master_copy_round_trip_and_channel_reselection():
    masks = config.make_mc_mask_dddict()
    mx = config.get_mc_num_dddict(masks, is_max=True)
    torch.manual_seed(0)
    master = SearchParallel(Network(10, mx, {'base': 0.0})).state_dict()
    master = {k: v.clone() for k, v in master.items()}
    # shrink one op and re-pick its channels by depthwise L1 norm
    arch = parsing.parse_architecture(*_random_arch(1))
    st, bl = 'stage3', 'block1'
    op = arch[st][bl]
    mc = config.get_mc_num_dddict(masks)
    mc[st][bl][op] -= 7
    elastic.reselect_channels(masks, mc, arch, master)
    m = masks[st][bl][op]
    assert int(m.sum()) == mc[st][bl][op]
    w = master['module.%s.%s.m_ops.%d.depth_conv.conv.weight' % (st, bl, op)].abs().sum((1, 2, 3))
    assert float(w[m.bool()].min()) >= float(w[~m.bool()].max())
    narrow = SearchParallel(Network(10, config.get_mc_num_dddict(masks), {'base': 0.0}))
    elastic.load_from_master(narrow, master, masks)
    key = 'module.%s.%s.m_ops.%d.point_linear.conv.weight' % (st, bl, op)
    idx = torch.nonzero(m).view(-1)
    assert torch.equal(narrow.state_dict()[key], master[key][:, idx])
    with torch.no_grad():
        for p in narrow.parameters():
            p.add_(1.0)
    before = {k: v.clone() for k, v in master.items()}
    elastic.store_to_master(master, narrow, masks)
    assert torch.equal(master[key][:, idx], before[key][:, idx] + 1.0)
    off = torch.nonzero(1 - m).view(-1)
    assert torch.equal(master[key][:, off], before[key][:, off])        # masked-out channels untouched
    op_w, depth_w = parsing.get_op_and_depth_weights(narrow)
    assert len(op_w) == 18 and len(depth_w) == 6 and abs(float(op_w[0].sum()) - 1.0 * np.exp(1.0) * 1.0) >= 0


@pytest.mark.skipif(not ref_shim.available(), reason='reference not mounted')
def test_checkpoint_is_consumed_by_the_reference_parser_and_model(tmp_path):
    """A checkpoint written the way train_search.py writes it (max-width state_dict under 'module.' + channel masks) goes
    through the REFERENCE's own parsing_model.get_op_and_depth_weights / parse_architecture / get_mc_num_dddict and loads
    (strict) into the reference's own search Network."""
    import torch.nn.functional as F
    ms, _cfg, pm = ref_shim._import()
    lut = gi.load_lut()
    mx = config.get_mc_num_dddict(config.mc_mask_dddict, is_max=True)
    model = SearchParallel(Network(100, mx, lut))
    g = torch.Generator().manual_seed(0)
    sd = model.state_dict()
    for k in sd:
        if k.endswith('log_alphas'):
            sd[k] = F.log_softmax(torch.randn(8, generator=g), -1)
        elif k.endswith('betas'):
            sd[k] = torch.randn(sd[k].shape, generator=g)
    masks = config.make_mc_mask_dddict()
    masks['stage3']['block2'][1][-17:] = 0
    path = str(tmp_path / 'searched_model_03.pth.tar')
    torch.save({'state_dict': sd, 'mc_mask_dddict': masks}, path)
    op_w, depth_w = pm.get_op_and_depth_weights(path)
    ours_op, ours_depth = parsing.get_op_and_depth_weights(path)
    assert len(op_w) == 18 and len(depth_w) == 6
    assert all(np.array_equal(a, b) for a, b in zip(op_w, ours_op)) and all(np.array_equal(a, b) for a, b in zip(depth_w, ours_depth))
    assert pm.parse_architecture(op_w, depth_w) == parsing.parse_architecture(ours_op, ours_depth)
    ref_masks = torch.load(path, weights_only=False)['mc_mask_dddict']
    assert pm.get_mc_num_dddict(ref_masks) == config.get_mc_num_dddict(masks)
    ref_net = ms.Network(100, mx, lut)
    missing = ref_net.load_state_dict({k[len('module.'):]: v for k, v in sd.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
