"""Host-side behaviour of the drop-in classes (no GPU): constructor surface, state_dict names,
switch / sampling logic, error behaviour, and the 'no CPU fallback' rule."""
import random

import pytest
import torch

from tests import golden_inputs as gi
from tfnas_b200 import _lib, config
from tfnas_b200.model_search import MixedOP, MixedStage, Network, NoisePlan, OPS, PRIMITIVES, injected


def _net():
    mcs = config.get_mc_num_dddict(config.mc_mask_dddict)
    torch.manual_seed(2)
    return Network(100, mcs, gi.load_lut())


def test_surface_and_state_dict():
    net = _net()
    assert len(PRIMITIVES) == 8 and set(OPS) == set(PRIMITIVES)
    sd = net.state_dict()
    assert len(sd) == 754 and len(list(net.buffers())) == 0
    assert len(net.weight_parameters()) == 730 and len(net.arch_parameters()) == 24
    assert len(net.log_alphas_parameters()) == 18 and len(net.betas_parameters()) == 6
    keys = list(sd)
    assert keys.index('stage1.betas') < keys.index('stage1.block1.log_alphas')
    op = net.stage2.block1.m_ops[5]
    assert tuple(op.inverted_bottleneck.conv.weight.shape) == (144, 24, 1, 1)
    assert tuple(op.depth_conv.conv.weight.shape) == (144, 1, 3, 3)
    assert tuple(op.squeeze_excite.conv_reduce.weight.shape) == (48, 144, 1, 1)
    assert (op.name, op.in_channels, op.se_channels, op.out_channels, op.kernel_size, op.stride, op.act_func,
            op.mid_channels) == ('MBInvertedResBlock', 24, 48, 40, 3, 2, 'swish', 144)
    assert torch.allclose(net.stage1.block1.log_alphas.exp().sum(), torch.tensor(1.0))
    # exec()-style surgery used by train_search.py:172-193
    exec('net.stage1.block1.m_ops[0].inverted_bottleneck.conv.weight.data = torch.zeros(40, 16, 1, 1)')
    assert net.stage1.block1.m_ops[0].inverted_bottleneck.conv.weight.shape[0] == 40


def test_lookup_latency_and_errors():
    net = _net()
    lats = net.stage1.block1.get_lookup_latency(112)
    lut = gi.load_lut()
    assert lats[0] == lut['MBInvertedResBlock_112_16_0_24_k3_s2_relu'][48]
    assert lats[7] == lut['MBInvertedResBlock_112_16_32_24_k5_s2_relu'][96]
    with pytest.raises(KeyError):
        net.stage2.block1.get_lookup_latency(32)
    with pytest.raises(ValueError):
        MixedStage([16], [24], [2], [False], ['relu'], {}, {}, 7)
    op = net.stage1.block1
    with pytest.raises(ValueError):
        op._sample_index('max')
    with pytest.raises(AttributeError):
        MixedOP(16, 24, 2, False, 'relu', 8, {i: 48 for i in range(8)}, {}).T   # T is not set by the ctor


def test_sampling_indices_follow_reference_semantics():
    net = _net()
    op = net.stage3.block2
    noise = -torch.empty(8).exponential_().log()
    with injected(NoisePlan(noise=[noise, noise])):
        ig = op._sample_index('gumbel')
        assert ig == int(torch.argmax(op.log_alphas.detach() + noise))
        assert op.switches[ig] is False and sum(op.switches) == 7
        random.seed(3)
        ir = op._sample_index('random')
    assert ir != ig and all(op.switches)
    random.seed(3)
    rest = [j for j in range(8) if j != ig]
    assert ir == rest[random.choice(range(7))]
    with injected(NoisePlan(indices=[4])):
        assert op._sample_index('random') == 4


def test_no_cpu_fallback():
    net = _net()
    net.set_temperature(5.0)
    x = torch.randn(1, 3, 32, 32)
    with pytest.raises(_lib.TfnasError):
        net(x, sampling=True, mode='gumbel')


def test_parameter_list_cache_matches_reference_filters():
    """Network.weight_parameters / arch_parameters / log_alphas_parameters / betas_parameters are cached (the search
    loop asks a dozen times per step) but must equal the reference's name-suffix filters (models/model_search.py:306-350),
    hand out fresh lists, and survive re-wrapping (_apply)."""
    import torch
    from tfnas_b200 import config
    from tfnas_b200.model_search import Network
    from tests import golden_inputs as gi
    net = Network(10, config.get_mc_num_dddict(config.mc_mask_dddict), gi.load_lut())

    def ref():
        named = list(net.named_parameters())
        return ([v for k, v in named if not (k.endswith('log_alphas') or k.endswith('betas'))],
                [v for k, v in named if k.endswith('log_alphas') or k.endswith('betas')],
                [v for k, v in named if k.endswith('log_alphas')], [v for k, v in named if k.endswith('betas')])

    def same(a, b):
        return len(a) == len(b) and all(x is y for x, y in zip(a, b))

    for _ in range(2):
        w, a, la, be = ref()
        assert same(net.weight_parameters(), w) and same(net.arch_parameters(), a)
        assert same(net.log_alphas_parameters(), la) and same(net.betas_parameters(), be)
        assert len(w) == 730 and len(a) == 24
        lst = net.weight_parameters()
        lst.clear()                                   # callers own the list they get
        assert len(net.weight_parameters()) == 730
        net = net.double()                            # _apply drops the cache
    net.extra = torch.nn.Linear(2, 2)
    net.invalidate_param_cache()
    assert len(net.weight_parameters()) == 732


def test_refresh_host_alphas_batches_and_tracks_updates():
    """Network.refresh_host_alphas mirrors every MixedOP's log_alphas on the host in one copy, follows in-place updates
    and `.data = ...` re-assignments (train_search.py:421-422), and gives _sample_index the values it would fetch itself."""
    import torch
    import torch.nn.functional as F
    from tfnas_b200 import config, model_search
    from tfnas_b200.model_search import MixedOP, Network
    from tests import golden_inputs as gi
    net = Network(10, config.get_mc_num_dddict(config.mc_mask_dddict), gi.load_lut())
    ops = [m for m in net.modules() if isinstance(m, MixedOP)]
    assert len(ops) == 18
    g = torch.Generator().manual_seed(0)
    for m in ops:
        m.log_alphas.data = F.log_softmax(torch.randn(8, generator=g), dim=-1)
    net.refresh_host_alphas()
    for m in ops:
        assert torch.equal(m._host_alpha[1], m.log_alphas.detach()) and m._host_alpha[1].dtype == torch.float32
        assert m._alphas_on_host() is m._host_alpha[1]                     # the per-module path sees a fresh mirror
    held = [m._host_alpha[1] for m in ops]
    net.refresh_host_alphas()                                              # nothing stale: nothing replaced
    assert all(a is m._host_alpha[1] for a, m in zip(held, ops))
    with torch.no_grad():
        ops[3].log_alphas.add_(0.25)                                       # in-place update bumps the version
    ops[7].log_alphas.data = F.log_softmax(torch.randn(8, generator=g), dim=-1)   # re-assignment changes the storage
    net.refresh_host_alphas()
    for m in ops:
        assert torch.equal(m._host_alpha[1], m.log_alphas.detach())
    # same sampled indices as the lazy per-module mirrors
    model_search.seed_noise(5)
    a = [m._sample_index('gumbel') for m in ops]
    net.reset_switches()
    for m in ops:
        m._host_alpha = None
    model_search.seed_noise(5)
    b = [m._sample_index('gumbel') for m in ops]
    net.reset_switches()
    model_search.seed_noise(None)
    assert a == b
