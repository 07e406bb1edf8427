"""Derived-network path without a GPU: tfnas_b200.model_eval against the golden fixture generated from the
real reference (tests/golden/make_golden_eval.py), against the live reference when it is mounted (state_dict, config,
latency, drop-connect / dropout draws), checkpoint interchange, the epoch lr rule and a two-rank gloo training step."""
import copy
import importlib.util
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_shim
from tests import golden_inputs as gi
from tfnas_b200 import config, eval_loop, model_eval
from tfnas_b200.parallel import GradSync

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden', 'derived_net.npz')


def _ours(num_classes=10, lut=None, dropout=0.0, drop_connect=0.0, seed=2):
    torch.manual_seed(seed)
    return model_eval.Network(num_classes, gi.derived_arch(), config.get_mc_num_dddict(config.mc_mask_dddict), lut,
                              dropout, drop_connect)


def test_golden_fixture_from_the_reference():
    """Same seed => same initial weights as the reference's Network (identical parameter creation order), so the train-mode
    logits, the running statistics it leaves behind (through the eval-mode logits) and the table latency all reproduce."""
    g = np.load(GOLD)
    net = _ours(lut=gi.load_lut())
    sd = net.state_dict()
    assert len(sd) == int(g['n_state'])
    for k, s in zip(g['probe_keys'], g['probe_sums']):
        assert abs(sd[str(k)].double().sum().item() - float(s)) <= 1e-6 * max(1.0, abs(float(s))), k
    xa, xb = gi.derived_inputs()
    net.train()
    la = net(xa).detach()
    net.eval()
    with torch.no_grad():
        lb = net(xb)
    sd = net.state_dict()
    for k, s in zip(g['running_keys'], g['running_sums']):          # running statistics after one train-mode forward
        assert abs(sd[str(k)].double().sum().item() - float(s)) <= 1e-5 * abs(float(s)), k
    for got, want in ((la, g['logits_train']), (lb, g['logits_eval'])):
        want = torch.from_numpy(want)
        assert (got - want).abs().max() <= 1e-4 * want.abs().max() + 1e-6
    assert abs(net.get_lookup_latency(torch.zeros(1, 3, 224, 224)) - float(g['lat'])) < 1e-9
    assert abs(net.get_lookup_latency(224) - float(g['lat'])) < 1e-9


def test_structure_config_round_trip_and_drop_rates():
    net = _ours(drop_connect=0.2)
    arch = gi.derived_arch()
    nblocks = 1 + sum(len(v) for v in arch.values())
    assert net.block_count == nblocks
    blocks = [net.second_stem] + list(net.blocks())
    assert [round(b.drop_connect_rate, 6) for b in blocks] == [round(0.2 * (i + 1) / nblocks, 6) for i in range(nblocks)]
    b = net.stage2[0]                      # op (5*2+3)%8 = 5: k3 e6 SE, 24 -> 40 stride 2
    assert (b.kernel_size, b.se_channels, b.in_channels, b.out_channels, b.stride, b.act_func) == (3, 48, 24, 40, 2, 'swish')
    assert not b.has_residual and net.stage3[1].has_residual
    cfg = net.config
    json.dumps(cfg)                        # train_eval.py:118-119 writes it as JSON
    assert cfg['first_stem']['name'] == 'ConvLayer' and cfg['stage1'][0]['name'] == 'MBInvertedResBlock'
    assert cfg['classifier'] == dict(name='LinearLayer', in_features=1280, out_features=10, bias=True, use_bn=False,
                                     affine=False, act_func=None, ops_order='weight_bn_act')
    clone = model_eval.NetworkCfg(7, copy.deepcopy(cfg), None, 0.1, 0.2)
    assert clone.classifier.linear.out_features == 7 and clone.config['stage4'] == cfg['stage4']
    a, b2 = net.state_dict(), clone.state_dict()
    assert list(a) == list(b2) and all(a[k].shape == b2[k].shape for k in a if not k.startswith('classifier'))
    assert cfg['stage1'][0]['name'] == 'MBInvertedResBlock'        # building from a config leaves the caller's dict intact
    with pytest.raises(KeyError):
        model_eval.set_layer_from_config(dict(name='IdentityLayer'))


def test_drop_connect_and_eval_mode():
    torch.manual_seed(0)
    x = torch.randn(64, 3, 2, 2)
    assert model_eval.drop_connect(x, False, 0.5) is x and model_eval.drop_connect(x, True, 0.0) is x
    y = model_eval.drop_connect(x, True, 0.25)
    kept = (y.flatten(1).abs().sum(1) > 0)
    assert 30 <= int(kept.sum()) <= 62
    assert torch.allclose(y[kept], x[kept] / 0.75)
    net = _ours(dropout=0.5, drop_connect=0.5)
    net.eval()
    xa, _ = gi.derived_inputs()
    with torch.no_grad():
        assert torch.equal(net(xa), net(xa))          # nothing stochastic in eval mode


@pytest.mark.skipif(not ref_shim.available(), reason='reference not mounted')
def test_against_the_live_reference():
    ref_shim._import()
    from models import model_eval as rme
    lut = gi.load_lut()
    mc = config.get_mc_num_dddict(config.mc_mask_dddict)
    torch.manual_seed(5)
    ref = rme.Network(10, gi.derived_arch(), mc, lut, 0.3, 0.2)
    ours = _ours(lut=lut, dropout=0.3, drop_connect=0.2, seed=5)
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    assert ref.config == ours.config
    x = torch.zeros(1, 3, 224, 224)
    assert ref.get_lookup_latency(x) == pytest.approx(ours.get_lookup_latency(x), abs=1e-9)
    xa, xb = gi.derived_inputs()
    ref.train(), ours.train()
    torch.manual_seed(9)
    la = ref(xa)
    torch.manual_seed(9)                    # same drop-connect gates and dropout mask from the same generator state
    lb = ours(xa)
    assert (la - lb).abs().max() <= 1e-5 * la.abs().max()
    # forward + backward in float64 (fp32 round-off of the two activation formulations aside, the math is identical)
    ref64, ours64 = copy.deepcopy(ref).double(), copy.deepcopy(ours).double()
    torch.manual_seed(9)
    la = ref64(xa.double())
    torch.manual_seed(9)
    lb = ours64(xa.double())
    assert (la - lb).abs().max() <= 1e-12 * la.abs().max()
    la.square().sum().backward()
    lb.square().sum().backward()
    scale = max(float(p.grad.abs().max()) for p in ref64.parameters())     # (a beta feeding a BN has a zero gradient)
    for (k, p), q in zip(ref64.named_parameters(), ours64.parameters()):
        assert (p.grad - q.grad).abs().max() <= 1e-9 * p.grad.abs().max() + 1e-12 * scale, k
    # a checkpoint of ours loads into the reference (strict) and the other way round; NetworkCfg from the reference's config
    ref.load_state_dict(ours.state_dict(), strict=True)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref_cfg = rme.NetworkCfg(10, copy.deepcopy(ours.config), None, 0.0, 0.0)
    ours_cfg = model_eval.NetworkCfg(10, copy.deepcopy(ref.config), None, 0.0, 0.0)
    assert list(ref_cfg.state_dict()) == list(ours_cfg.state_dict())
    ref_cfg.load_state_dict(ours.state_dict(), strict=True)
    ours_cfg.load_state_dict(ours.state_dict(), strict=True)
    ref_cfg.eval(), ours_cfg.eval()
    with torch.no_grad():
        assert (ref_cfg(xb) - ours_cfg(xb)).abs().max() <= 1e-5 * ref_cfg(xb).abs().max()


def test_label_smoothing_reference_and_lr_rule():
    """The unfused criterion equals CrossEntropyLabelSmooth (train_eval.py:72-84); warm-up rule of train_eval.py:201-208."""
    torch.manual_seed(1)
    logits, target = torch.randn(6, 10), torch.randint(0, 10, (6,))
    smooth, plain = eval_loop.make_criteria(0.1, fused=False)
    lp = torch.log_softmax(logits, 1)
    t = torch.zeros_like(lp).scatter_(1, target.unsqueeze(1), 1) * 0.9 + 0.1 / 10
    assert torch.allclose(smooth(logits, target), (-t * lp).mean(0).sum(), atol=1e-6)
    assert torch.allclose(plain(logits, target), torch.nn.functional.cross_entropy(logits, target))
    lrs = [0.2, 0.19, 0.18, 0.17, 0.16, 0.15]
    assert [eval_loop.epoch_lr(lrs, e, 512) for e in range(6)] == pytest.approx([0.04, 0.076, 0.108, 0.136, 0.16, 0.15])
    assert [eval_loop.epoch_lr(lrs, e, 256) for e in range(6)] == lrs


def test_cli_flags_match_reference_defaults(tmp_path):
    spec = importlib.util.spec_from_file_location('te_cli', os.path.join(ROOT, 'train_eval.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    args = m.build_parser().parse_args([])
    ref = dict(epochs=250, batch_size=512, lr=0.2, momentum=0.9, weight_decay=1e-5, grad_clip=5.0, label_smooth=0.1,
               num_classes=1000, dropout_rate=0.2, drop_connect_rate=0.2, seed=2, print_freq=100, workers=16)
    for k, v in ref.items():
        assert getattr(args, k) == v, k
    # model from a config file and from a search checkpoint (train_eval.py:103-115)
    cfg_path = str(tmp_path / 'model.config')
    with open(cfg_path, 'w') as f:
        json.dump(_ours().config, f)
    a = m.build_parser().parse_args(['--config_path', cfg_path, '--num_classes', '10'])
    assert isinstance(m.build_model(a), model_eval.NetworkCfg)
    from tfnas_b200.model_search import Network as SearchNetwork
    from tfnas_b200.parallel import SearchParallel
    mx = config.get_mc_num_dddict(config.mc_mask_dddict, is_max=True)
    ck = str(tmp_path / 'searched_model_01.pth.tar')
    sd = SearchParallel(SearchNetwork(10, mx, gi.load_lut())).state_dict()
    torch.save({'state_dict': sd, 'mc_mask_dddict': config.make_mc_mask_dddict()}, ck)
    a = m.build_parser().parse_args(['--model_path', ck, '--num_classes', '10'])
    net = m.build_model(a)
    assert isinstance(net, model_eval.Network) and net.block_count == 7        # betas = 0: the sink keeps one block per stage
    for k in sd:
        if k.endswith('betas'):
            sd[k] = torch.arange(sd[k].numel(), dtype=torch.float32)            # deepest sink wins: all 18 blocks kept
    torch.save({'state_dict': sd, 'mc_mask_dddict': config.make_mc_mask_dddict()}, ck)
    assert m.build_model(a).block_count == 19
    with pytest.raises(SystemExit):
        m.build_model(m.build_parser().parse_args([]))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        net = _ours(seed=3 + rank)                   # ranks start different ...
        eval_loop.broadcast_model(net)                # ... and are made identical
        ref = copy.deepcopy(net)
        xa, _ = gi.derived_inputs()
        target = torch.tensor([1, 3, 5, 7])
        smooth, _ = eval_loop.make_criteria(0.1, fused=False)
        opt = eval_loop.make_optimizer(net, 0.1, 0.9, 1e-5, fused=False)
        sl = slice(2 * rank, 2 * rank + 2)
        eval_loop.train_step(net, xa[sl], target[sl], smooth, opt, 5.0, GradSync())
        # expected: mean over ranks of the per-shard gradients, clipped, one SGD step from the common start
        grads = []
        for r in range(world):
            m = copy.deepcopy(ref)
            m.train()
            s2 = slice(2 * r, 2 * r + 2)
            smooth(m(xa[s2]), target[s2]).backward()
            grads.append([p.grad for p in m.parameters()])
        mean = [sum(g) / world for g in zip(*grads)]
        for p, g in zip(ref.parameters(), mean):
            p.grad = g
        torch.nn.utils.clip_grad_norm_(list(ref.parameters()), 5.0)
        torch.optim.SGD(ref.parameters(), 0.1, momentum=0.9, weight_decay=1e-5).step()
        err = max(float((p - q).abs().max() / (q.abs().max() + 1e-12)) for p, q in zip(net.parameters(), ref.parameters()))
        q.put((rank, err))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_training_step_equals_mean_of_shard_gradients():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(err < 1e-5 for _, err in res), res


def test_fused_sgd_state_dict_round_trip_on_host():
    """The checkpoint 'optimizer' entry of train_eval.py: hyper-parameters plus the momentum buffers as one flat tensor in
    parameter order (no kernel call involved, so this runs without a GPU)."""
    from tfnas_b200.step import FusedSGD
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5))]
    a = FusedSGD(params, 0.2, momentum=0.9, weight_decay=1e-5)
    assert a.state_dict()['momentum_flat'] is None                 # nothing stepped yet
    a._state(params[0].device)
    flat = torch.arange(17, dtype=torch.float32)
    a.load_state_dict(dict(momentum_flat=flat, lr=0.05, momentum=0.8, weight_decay=1e-4))
    assert (a.lr, a.momentum, a.weight_decay, a.param_groups[0]['lr']) == (0.05, 0.8, 1e-4, 0.05)
    assert torch.equal(a._bufs[id(params[0])], flat[:12].view(3, 4)) and torch.equal(a._bufs[id(params[1])], flat[12:])
    b = FusedSGD(params, 0.2)
    b.load_state_dict(a.state_dict())
    assert torch.equal(b.state_dict()['momentum_flat'], flat) and b.lr == 0.05


@pytest.mark.skipif(not ref_shim.available(), reason='reference not mounted')
def test_label_smoothing_against_the_reference_class():
    """CrossEntropyLabelSmooth lifted out of the reference's train_eval.py (:72-84; the script itself cannot be imported:
    argparse with required arguments and mkdir at import) against the unfused criterion of eval_loop and, on the same
    numbers, the closed form the kernel implements: lse - (1 - eps) l[target] - eps / C * sum(l)."""
    src = open(os.path.join(ref_shim.REF_ROOT, 'train_eval.py')).read()
    start, end = src.index('class CrossEntropyLabelSmooth'), src.index('def set_seed')
    ns = {'nn': torch.nn, 'torch': torch}
    exec(compile(src[start:end], 'ref_train_eval_criterion', 'exec'), ns)
    torch.manual_seed(4)
    logits, target = 3 * torch.randn(16, 100, dtype=torch.float64), torch.randint(0, 100, (16,))
    for eps in (0.0, 0.1, 0.3):
        ref = ns['CrossEntropyLabelSmooth'](100, eps)(logits, target)
        ours = torch.nn.CrossEntropyLoss(label_smoothing=eps)(logits, target)
        lse = torch.logsumexp(logits, 1)
        closed = (lse - (1 - eps) * logits.gather(1, target[:, None])[:, 0] - eps / 100 * logits.sum(1)).mean()
        assert abs(float(ref) - float(ours)) < 1e-12 and abs(float(ref) - float(closed)) < 1e-12
