"""Derived-network re-training on the GPU: the label-smoothing cross-entropy kernel against
CrossEntropyLabelSmooth (train_eval.py:72-84), a fused training trajectory against torch.optim + clip_grad_norm_ +
torch's criterion, the bf16 / channels-last mode, and train_eval.py end to end on synthetic data (from a config file and
from a checkpoint written by train_search.py's format), including a snapshot restart."""
import copy
import glob
import importlib.util
import json
import os

import pytest
import torch

from tests import golden_inputs as gi
from tfnas_b200 import config, eval_loop, model_eval
from tfnas_b200.step import FusedSGD, softmax_ce

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _net(drop=0.0, seed=2, classes=10):
    torch.manual_seed(seed)
    return model_eval.Network(classes, gi.derived_arch(), config.get_mc_num_dddict(config.mc_mask_dddict), None, drop, drop)


@pytest.mark.parametrize('eps', [0.0, 0.1, 0.35])
def test_label_smoothing_ce_kernel(eps):
    torch.manual_seed(3)
    N, C = 96, 1000
    logits = (3 * torch.randn(N, C)).cuda().requires_grad_(True)
    target = torch.randint(0, C, (N,)).cuda()
    loss = softmax_ce(logits, target, eps)
    loss.backward()
    l64 = logits.detach().double().requires_grad_(True)
    lp = torch.log_softmax(l64, 1)                      # CrossEntropyLabelSmooth, in float64
    t = torch.zeros_like(lp).scatter_(1, target.unsqueeze(1), 1) * (1 - eps) + eps / C
    want = (-t * lp).mean(0).sum()
    want.backward()
    assert abs(float(loss.detach()) - float(want.detach())) <= 2e-6 * abs(float(want.detach()))
    assert float((logits.grad.double() - l64.grad).abs().max()) <= 1e-6 * float(l64.grad.abs().max())
    with pytest.raises(Exception):
        softmax_ce(logits, target, 1.0)


def test_fused_trajectory_matches_torch_optim():
    """FusedCrossEntropy(0.1) + FusedSGD(clip 5) vs nn.CrossEntropyLoss(label_smoothing) + clip_grad_norm_ +
    torch.optim.SGD: the update rule on identical gradients, then one whole step from identical weights."""
    a = _net().cuda()
    b = copy.deepcopy(a)
    xa, _ = gi.derived_inputs()
    g = torch.Generator().manual_seed(5)
    batches = [(torch.randn(8, 3, 64, 64, generator=g).cuda(), torch.randint(0, 10, (8,), generator=g).cuda()) for _ in range(4)]
    smooth_f, _ = eval_loop.make_criteria(0.1, fused=True)
    smooth_t, _ = eval_loop.make_criteria(0.1, fused=False)
    opt_f = eval_loop.make_optimizer(a, 0.1, 0.9, 1e-5, fused=True)
    opt_t = eval_loop.make_optimizer(b, 0.1, 0.9, 1e-5, fused=False)
    assert isinstance(opt_f, FusedSGD)
    a.train(), b.train()

    def worst():
        # (+1e-5: a BN shift that feeds a 1x1 conv + batch-stat BN has a mathematically zero gradient, its value is round-off)
        return max(float((p.detach() - q.detach()).abs().max() / (q.detach().abs().max() + 1e-5))
                   for p, q in zip(a.parameters(), b.parameters()))

    # the update rule by itself (clip, weight decay, momentum, lr): two updates from identical weights AND gradients
    for x, t in batches[:2]:
        smooth_f(a(x), t).backward()
        for p, q in zip(a.parameters(), b.parameters()):
            q.grad = p.grad.clone()
        opt_f.step(max_norm=5.0)
        torch.nn.utils.clip_grad_norm_(list(b.parameters()), 5.0)
        opt_t.step()
        opt_f.zero_grad(), opt_t.zero_grad()
        b.load_state_dict({k: v for k, v in a.state_dict().items() if 'running' in k or 'tracked' in k}, strict=False)
        assert worst() < 2e-6, worst()
    # one whole step from identical weights, each arm with its own criterion and gradients.  (Further steps are not
    # compared: the early layers' fp32 gradients of this net differ by ~1e-3 between two valid evaluations, and the two
    # trajectories drift apart chaotically -- 0.4 % to 3 % after four steps at lr 0.1 from run to run.)
    b.load_state_dict(a.state_dict())
    x, t = batches[2]
    lf, _ = eval_loop.train_step(a, x, t, smooth_f, opt_f, 5.0)
    lt, _ = eval_loop.train_step(b, x, t, smooth_t, opt_t, 5.0)
    assert abs(float(lf) - float(lt)) <= 1e-5 * abs(float(lt))
    assert worst() < 2e-3, worst()
    for (k, u), v in zip(a.state_dict().items(), b.state_dict().values()):
        if 'running' in k:
            assert float((u - v).abs().max()) <= 1e-4 * float(v.abs().max()) + 1e-6, k
    # optimiser state round trip (checkpoint 'optimizer' entry)
    st = opt_f.state_dict()
    c = copy.deepcopy(a)
    opt_c = eval_loop.make_optimizer(c, 0.3, 0.0, 0.0, fused=True)
    opt_c.load_state_dict(st)
    assert (opt_c.lr, opt_c.momentum, opt_c.weight_decay) == (opt_f.lr, opt_f.momentum, opt_f.weight_decay)
    x, t = batches[0]
    eval_loop.train_step(a, x, t, smooth_t, opt_f, 5.0)
    eval_loop.train_step(c, x, t, smooth_t, opt_c, 5.0)
    assert max(float((p - q).abs().max() / (q.abs().max() + 1e-5)) for p, q in zip(a.parameters(), c.parameters())) < 2e-3


def test_bf16_channels_last_training_reduces_the_loss():
    net = _net(drop=0.1).cuda()
    smooth, plain = eval_loop.make_criteria(0.1)
    opt = eval_loop.make_optimizer(net, 0.05, 0.9, 1e-5)
    g = torch.Generator().manual_seed(6)
    x, t = torch.randn(16, 3, 64, 64, generator=g).cuda(), torch.randint(0, 10, (16,), generator=g).cuda()
    net.train()
    losses = [float(eval_loop.train_step(net, x, t, smooth, opt, 5.0, None, 'bf16', True)[0]) for _ in range(12)]
    assert all(l == l for l in losses) and losses[-1] < 0.8 * losses[0], losses
    assert all(p.dtype == torch.float32 for p in net.parameters())          # master weights stay fp32 under autocast
    net.eval()
    with torch.no_grad(), eval_loop.autocast('bf16'):
        out = net(x.contiguous(memory_format=torch.channels_last))
    assert out.dtype == torch.bfloat16 and torch.isfinite(out.float()).all()
    with pytest.raises(ValueError):
        eval_loop.autocast('fp8')


def test_train_eval_cli_synthetic(tmp_path):
    spec = importlib.util.spec_from_file_location('te_cli_run', os.path.join(ROOT, 'train_eval.py'))
    te = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(te)
    cfg_path = str(tmp_path / 'arch.config')
    with open(cfg_path, 'w') as f:
        json.dump(_net().config, f)
    lut = os.path.join(gi.GOLDEN_DIR, 'lut_gpu.npz')
    common = ['--synthetic', '3', '--batch_size', '8', '--num_classes', '10', '--image_size', '64', '--print_freq', '1',
              '--save', str(tmp_path), '--config_path', cfg_path, '--lookup_path', lut]
    te.main(common + ['--epochs', '2', '--amp', 'bf16', '--channels_last', '--note', 'bf16'])
    run = glob.glob(os.path.join(str(tmp_path), 'eval-*-bf16'))
    assert len(run) == 1
    ck = torch.load(os.path.join(run[0], 'checkpoint.pth.tar'), weights_only=False)
    assert ck['epoch'] == 2 and set(ck) == {'epoch', 'state_dict', 'best_acc_top1', 'best_acc_top5', 'optimizer'}
    assert all(k.startswith('module.') for k in ck['state_dict'])
    with open(os.path.join(run[0], 'model.config')) as f:
        written = json.load(f)
    clone = model_eval.NetworkCfg(10, written)
    clone.load_state_dict({k[7:]: v for k, v in ck['state_dict'].items()}, strict=True)
    assert int(ck['state_dict']['module.first_stem.bn.num_batches_tracked']) == 6       # 2 epochs x 3 batches
    log = open(os.path.join(run[0], 'log.txt')).read()
    assert 'TRAIN Step' in log and 'VALID Step' in log and 'Val_acc_top1' in log and 'table latency' in log
    # restart from the snapshot (train_eval.py:169-190): continues at epoch 2 in fp32
    te.main(common + ['--epochs', '3', '--snapshot', os.path.join(run[0], 'checkpoint.pth.tar'), '--note', 'resume'])
    run2 = glob.glob(os.path.join(str(tmp_path), 'eval-*-resume'))
    ck2 = torch.load(os.path.join(run2[0], 'checkpoint.pth.tar'), weights_only=False)
    assert ck2['epoch'] == 3 and int(ck2['state_dict']['module.first_stem.bn.num_batches_tracked']) == 9
    assert 'Epoch: 2 lr' in open(os.path.join(run2[0], 'log.txt')).read()
