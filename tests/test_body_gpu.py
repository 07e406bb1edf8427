"""The C++ body executor (tfnas_body_fwd/_bwd: all MixedStages in one call per direction) against the per-MixedOP autograd
path (MixedStage.forward over tfnas_mixedop_* / tfnas_stage_sink_*): same kernels, so results agree to atomics-order noise;
plus the fused step glue (tfnas_sgd_step / tfnas_adam_step / tfnas_softmax_ce) against torch.optim / torch CE."""
import random

import pytest
import torch
import torch.nn.functional as F

from oracle import ref_shim
from tests import golden_inputs as gi
from tests import helpers as H
from tfnas_b200 import config, model_search
from tfnas_b200.model_search import Network, NoisePlan, injected

pytestmark = pytest.mark.gpu


def _net(use_body, N=4, size=224, seed=3):
    mcs = config.get_mc_num_dddict(config.mc_mask_dddict)
    P, _x, _t = gi.network_inputs()
    net = Network(100, mcs, gi.load_lut())
    net.load_state_dict(P)
    net.set_temperature(5.0)
    net.use_body = use_body
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, 3, size, size, generator=g)
    t = torch.randint(0, 100, (N,), generator=g)
    return net.cuda().train(), x.cuda(), t.cuda()


def test_body_alpha_step_equals_per_op_path():
    res = []
    for use_body in (False, True):
        net, x, t = _net(use_body)
        for p in net.weight_parameters():
            p.requires_grad_(False)
        with injected(NoisePlan(noise=ref_shim.draw_plan_noise(5))):
            logits, lat = net(x, sampling=False)
        (F.cross_entropy(logits, t) + torch.abs(lat / 15.0 - 1.) * 0.1).backward()
        npar = dict(net.named_parameters())
        da = torch.stack([npar[k].grad for k in npar if k.endswith('log_alphas')])
        db = torch.cat([npar[k].grad for k in npar if k.endswith('betas')])
        res.append((logits.detach().clone(), float(lat), da.clone(), db.clone()))
    (l0, t0, a0, b0), (l1, t1, a1, b1) = res
    e = dict(logits=H.rel_l2(l1, l0), lat=abs(t1 - t0), dalpha=H.rel_l2(a1, a0), dbeta=H.rel_l2(b1, b0))
    print('body vs per-op alpha step', e)
    # use_body also moves stems and head from torch (cuDNN / cuBLAS) onto the library: different fp32 algorithms there
    assert e['logits'] < 1e-5 and e['lat'] < 1e-5 and e['dalpha'] < 1e-3 and e['dbeta'] < 1e-3


def test_body_w_step_equals_per_op_path():
    res = []
    for use_body in (False, True):
        net, x, t = _net(use_body, N=8, size=224)
        for p in net.arch_parameters():
            p.requires_grad_(False)
        random.seed(4)
        with injected(NoisePlan(noise=ref_shim.draw_plan_noise(6))):
            lg, z = net(x, sampling=True, mode='gumbel')
        lr, _ = net(x, sampling=True, mode='random')
        assert z == 0.0
        (F.cross_entropy(lg, t) + F.cross_entropy(lr, t)).backward()
        res.append((lg.detach().clone(), lr.detach().clone(),
                    {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}))
        assert all(all(m.switches) for m in net.modules() if hasattr(m, 'switches'))
    (g0, r0, w0), (g1, r1, w1) = res
    assert set(w0) == set(w1)
    gmax = max(float(v.norm()) for v in w0.values())
    err = {k: float((w1[k] - w0[k]).norm() / max(float(w0[k].norm()), 1e-6 * gmax)) for k in w0}
    # use_body also moves the stems from torch (cuDNN) onto the library's kernels: two different fp32 algorithms there, the
    # same kernels everywhere else
    worst = max(v for k, v in err.items() if 'stem' not in k)
    worst_stem = max(v for k, v in err.items() if 'stem' in k)
    print('body vs per-op w step: logits %.2e %.2e, worst grad %.2e (stems %.2e) over %d tensors'
          % (H.rel_l2(g1, g0), H.rel_l2(r1, r0), worst, worst_stem, len(w0)))
    assert H.rel_l2(g1, g0) < 1e-5 and H.rel_l2(r1, r0) < 1e-5 and worst < 2e-3 and worst_stem < 3e-3


def test_body_no_grad_forward_and_arena_reuse():
    net, x, t = _net(True)
    with torch.no_grad():
        a, _ = net(x, sampling=True, mode='gumbel')
        net.reset_switches()
        model_search.seed_noise(1)
        b, _ = net(x, sampling=True, mode='random')
    model_search.seed_noise(None)
    assert torch.isfinite(a).all() and torch.isfinite(b).all()
    assert len(net._arena_pool.free) == 2          # one stem arena and one body arena served both passes


def test_fused_sgd_matches_torch_three_steps():
    from tfnas_b200.step import FusedSGD
    g = torch.Generator().manual_seed(0)
    shapes = [(32, 3, 3, 3), (7,), (1152, 192, 1, 1), (100, 1280), (5, 1, 5, 5)] * 60      # > one kernel-parameter chunk
    ref = [torch.randn(s, generator=g).cuda().requires_grad_(True) for s in shapes]
    mine = [p.detach().clone().requires_grad_(True) for p in ref]
    opt = torch.optim.SGD(ref, lr=0.025, momentum=0.9, weight_decay=1e-5)
    fused = FusedSGD(mine, lr=0.025, momentum=0.9, weight_decay=1e-5)
    for step in range(3):
        live = [i for i in range(len(ref)) if (i + step) % 3 != 0]          # tensors without a gradient are skipped (Q5)
        for p in ref + mine:
            p.grad = None
        for i in live:
            gr = torch.randn(shapes[i], generator=g).cuda() * (3.0 if step == 1 else 0.01)
            ref[i].grad = gr.clone()
            mine[i].grad = gr.clone()
        torch.nn.utils.clip_grad_norm_(ref, 5.0)
        opt.step()
        fused.step(max_norm=5.0)
    worst = max(H.rel_l2(a, b) for a, b in zip(mine, ref))
    print('fused SGD vs torch.optim.SGD after 3 steps: worst rel-l2 %.2e' % worst)
    assert worst < 1e-6


def test_fused_adam_matches_torch_three_steps():
    from tfnas_b200.step import FusedArchAdam
    g = torch.Generator().manual_seed(1)
    shapes = [(8,)] * 18 + [(2,), (3,), (4,), (4,), (4,), (1,)]
    ref = [F.log_softmax(torch.randn(s, generator=g), -1).cuda().requires_grad_(True) for s in shapes]
    mine = [p.detach().clone().requires_grad_(True) for p in ref]
    opt = torch.optim.Adam(ref, lr=0.01, betas=(0.5, 0.999), weight_decay=5e-4)
    fused = FusedArchAdam(mine, lr=0.01, betas=(0.5, 0.999), weight_decay=5e-4)
    for step in range(3):
        for a, b in zip(ref, mine):
            gr = torch.randn(a.shape, generator=g).cuda() * (4.0 if step == 0 else 0.05)
            a.grad, b.grad = gr.clone(), gr.clone()
        torch.nn.utils.clip_grad_norm_(ref, 5.0)
        opt.step()
        for p in ref:
            p.data = F.log_softmax(p.detach().data, dim=-1)
        fused.step(max_norm=5.0)
    worst = max(H.rel_l2(a, b) for a, b in zip(mine, ref) if a.numel() > 1)
    print('fused Adam + renorm vs torch after 3 steps: worst rel-l2 %.2e' % worst)
    assert worst < 1e-5 and all(float(a.abs().max()) < 1e-6 for a in mine if a.numel() == 1)


def test_softmax_ce_matches_torch():
    from tfnas_b200.step import softmax_ce
    g = torch.Generator().manual_seed(2)
    logits = (torch.randn(128, 100, generator=g) * 3).cuda().requires_grad_(True)
    tgt = torch.randint(0, 100, (128,), generator=g).cuda()
    ref = logits.detach().clone().requires_grad_(True)
    (F.cross_entropy(ref, tgt) * 1.7).backward()
    loss = softmax_ce(logits, tgt)
    (loss * 1.7).backward()
    assert abs(float(loss) - float(F.cross_entropy(ref, tgt))) < 1e-5
    assert H.rel_l2(logits.grad, ref.grad) < 1e-5


def test_graphed_alpha_step_equals_eager():
    """search_loop.alpha_step replayed as a CUDA graph (forward + backward of the whole network captured once) gives the
    same architecture-parameter trajectory as the eager step over three updates with fresh batches and fresh noise."""
    from tfnas_b200.parallel import SearchParallel
    from tfnas_b200.search_loop import alpha_step, make_optimizers
    from tfnas_b200.step import FusedCrossEntropy
    res = []
    for graph in (False, True):
        net, _x, _t = _net(True, N=4)
        model = SearchParallel(net)
        _ow, opt_a = make_optimizers(net)
        crit = FusedCrossEntropy()
        model_search.seed_noise(21)
        g = torch.Generator().manual_seed(9)
        losses = []
        for step in range(3):
            x = torch.randn(4, 3, 224, 224, generator=g).cuda()
            t = torch.randint(0, 100, (4,), generator=g).cuda()
            la, ll = alpha_step(model, x, t, crit, opt_a, 15.0, 0.1, 5.0, None, graph=graph)
            losses.append((float(la), float(ll)))
        res.append((losses, torch.cat([p.detach().reshape(-1) for p in net.arch_parameters()]).clone()))
        assert ('_alpha_graph' in net.__dict__) == graph
    model_search.seed_noise(None)
    (l0, a0), (l1, a1) = res
    print('graphed vs eager alpha steps: losses', l0, l1, 'arch params rel-l2 %.2e' % H.rel_l2(a1, a0))
    assert all(abs(p[0] - q[0]) < 1e-4 and abs(p[1] - q[1]) < 1e-5 for p, q in zip(l0, l1))
    assert H.rel_l2(a1, a0) < 1e-5
