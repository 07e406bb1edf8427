"""CPU ORACLE — test infrastructure, not product code.

A functional restatement (torch CPU ops, fp32 or fp64) of the reference's
supernet search path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package; the product (``tfnas_b200``) never does.

What it follows, line by line (paths relative to /root/reference):
  * MBConv candidate      models/layers.py:539-561 (ctor :433-537)
  * MixedOP               models/model_search.py:58-91, LUT :93-111
  * MixedStage            models/model_search.py:157-206
  * Network               models/model_search.py:281-304 (ctor :214-279)
  * Gumbel-softmax        torch.nn.functional.gumbel_softmax (noise injected:
                          g = -log(Exp(1)), softmax((logits+g)/tau))
  * search losses         train_search.py:375-379, 409-412

Pinning: ``tests/golden/make_golden.py`` runs the real reference from
/root/reference in the build container and commits inputs/outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this port against
those vectors (and against the live reference when it is mounted).  The
reference itself has no tests or golden vectors, so these
generated vectors are the pin.

Parameters are passed as a flat dict keyed like the reference ``state_dict``
(no ``module.`` prefix), so the same dict drives the reference, this port and
the CUDA path.
"""
import random as _pyrandom
from collections import OrderedDict

import torch
import torch.nn.functional as F

from tfnas_b200.config import CAND_SPEC, NUM_OPS, STAGE_SPEC, block_shapes, lut_key

BN_EPS = 1e-5


def _act(x, act):
    if act == 'relu':
        return F.relu(x)
    if act == 'swish':
        return x * torch.sigmoid(x)  # layers.py:26-35
    raise ValueError(act)


def _bn(x):
    # nn.BatchNorm2d(affine=False, track_running_stats=False): batch stats, biased var
    return F.batch_norm(x, None, None, None, None, True, 0.0, BN_EPS)


def mbconv(x, P, prefix, k, stride, act, se):
    """models/layers.py:539-561 for one candidate whose tensors live at P[prefix + ...]."""
    ic = x.shape[1]
    res = x
    w1 = P.get(prefix + 'inverted_bottleneck.conv.weight')
    if w1 is not None:                                                    # :542-545
        x = _act(_bn(F.conv2d(x, w1)), act)
    dw = P[prefix + 'depth_conv.conv.weight']
    x = _act(_bn(F.conv2d(x, dw, None, stride, k // 2, 1, dw.shape[0])), act)   # :547
    if se > 0:                                                            # :548-550
        g = F.adaptive_avg_pool2d(x, 1)
        g = F.conv2d(g, P[prefix + 'squeeze_excite.conv_reduce.weight'], P[prefix + 'squeeze_excite.conv_reduce.bias'])
        g = _act(g, act)
        g = F.conv2d(g, P[prefix + 'squeeze_excite.conv_expand.weight'], P[prefix + 'squeeze_excite.conv_expand.bias'])
        x = x * torch.sigmoid(g)
    w3 = P[prefix + 'point_linear.conv.weight']
    x = _bn(F.conv2d(x, w3))                                              # :552
    if ic == w3.shape[0] and stride == 1:                                 # :556-559
        x = x + res
    return x


def gumbel_weights(logits, g, T):
    """F.gumbel_softmax(logits, T, hard=False) with the noise g injected."""
    return F.softmax((logits + g) / T, dim=-1)


def mixedop_lats(lut, size, ic, oc, stride, act, mcs):
    """models/model_search.py:93-111."""
    return [lut[lut_key(size, ic, se_mult * ic, oc, k, stride, act)][mcs[i]]
            for i, (k, _e, se_mult) in enumerate(CAND_SPEC)]


def mixedop_alpha(x, P, prefix, ic, oc, stride, act, T, g, lats):
    """MixedOP.forward(sampling=False), models/model_search.py:86-91."""
    w = gumbel_weights(P[prefix + 'log_alphas'], g, T)
    out = 0
    out_lat = 0
    for i, (k, _e, se_mult) in enumerate(CAND_SPEC):
        out = out + w[i] * mbconv(x, P, '%sm_ops.%d.' % (prefix, i), k, stride, act, se_mult * ic)
        out_lat = out_lat + w[i] * lats[i]
    return out, out_lat


def mixedop_single(x, P, prefix, ic, oc, stride, act, idx):
    """MixedOP.forward(sampling=True): one candidate, models/model_search.py:84-85."""
    k, _e, se_mult = CAND_SPEC[idx]
    return mbconv(x, P, '%sm_ops.%d.' % (prefix, idx), k, stride, act, se_mult * ic)


def sample_gumbel_index(log_alphas, g):
    """'gumbel' mode index (:61-63): argmax softmax((log_softmax(a)+g)/T) == argmax(log_softmax(a)+g)."""
    return int(torch.argmax(F.log_softmax(log_alphas.detach().float().cpu(), dim=-1) + g.float().cpu()).item())


def draw_gumbel(n=NUM_OPS, generator=None):
    """One F.gumbel_softmax noise draw from the CPU generator."""
    return -torch.empty(n).exponential_(generator=generator).log()


class SearchPlan(object):
    """Everything random in one supernet forward, drawn up front in forward order."""

    def __init__(self, noise=None, indices=None):
        self.noise = noise        # list of 18 tensors [8] (alpha mode / gumbel sampling)
        self.indices = indices    # list of 18 ints (sampling modes)


def network_forward(x, P, mcs, lut, sampling, T=5.0, noise=None, indices=None, return_feats=False):
    """Network.forward, models/model_search.py:281-304.

    ``mcs[stage][block][op]`` are the mid widths; ``noise`` = 18 Gumbel draws
    (alpha mode) or ``indices`` = 18 candidate ids (sampling modes).
    """
    out_lat = lut['base'] if not sampling else 0.0
    x = F.relu(_bn(F.conv2d(x, P['first_stem.conv.weight'], None, 2, 1)))          # ConvLayer(3,32,k3,s2)
    x = mbconv(x, P, 'second_stem.', 3, 1, 'relu', 8)                              # MBConv(32,32,8,16)
    feats = []
    bi = 0
    shapes = list(block_shapes(x.shape[-1]))
    for stage, sp in STAGE_SPEC.items():
        res_list, lat_list, cum = [], [], 0.0
        nb = len(sp['ics'])
        for j in range(nb):
            _st, block, ic, oc, s, act, _size = shapes[bi]
            prefix = '%s.%s.' % (stage, block)
            if sampling:
                x = mixedop_single(x, P, prefix, ic, oc, s, act, indices[bi])
                lat = 0
            else:
                lats = mixedop_lats(lut, x.shape[-1], ic, oc, s, act, mcs[stage][block])
                x, lat = mixedop_alpha(x, P, prefix, ic, oc, s, act, T, noise[bi], lats)
            cum = cum + lat
            res_list.append(x)
            lat_list.append(cum)
            bi += 1
        # sink-connecting, models/model_search.py:202-204 (start_res == 1 for all six stages)
        beta = F.softmax(P[stage + '.betas'], dim=-1)
        out = 0
        slat = 0
        for j in range(nb):
            out = out + beta[j] * res_list[j]
            slat = slat + beta[j] * lat_list[j]
        x = out
        out_lat = out_lat + slat
        feats.append(x)
    x = _act(_bn(F.conv2d(x, P['feature_mix_layer.conv.weight'])), 'swish')
    x = F.adaptive_avg_pool2d(x, 1).flatten(1)
    x = F.linear(x, P['classifier.linear.weight'], P['classifier.linear.bias'])
    if return_feats:
        return x, out_lat, feats
    return x, out_lat


def arch_loss(logits, lat, target, target_lat, lambda_lat):
    """train_search.py:409-412."""
    loss_a = F.cross_entropy(logits, target)
    loss_l = torch.abs(lat / target_lat - 1.) * lambda_lat
    return loss_a + loss_l, loss_a, loss_l


def init_params(mcs, num_classes=100, seed=2, dtype=torch.float32):
    """Random-init parameter dict with the reference's names/shapes and torch default inits.

    Follows the constructor order of models/model_search.py:214-279 and
    models/layers.py:464-537 so that, under the same seed, the values equal the
    reference's ``Network(...)`` (checked in tests when the reference is mounted).
    """
    import torch.nn as nn
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    P = OrderedDict()

    def conv(name, oc, ic, k, groups=1, bias=False):
        m = nn.Conv2d(ic, oc, k, groups=groups, bias=bias)
        P[name + '.weight'] = m.weight.detach().to(dtype)
        if bias:
            P[name + '.bias'] = torch.zeros(oc, dtype=dtype)   # Network._initialization zeroes biases

    conv('first_stem.conv', 32, 3, 3)
    conv('second_stem.depth_conv.conv', 32, 32, 3, groups=32)
    conv('second_stem.squeeze_excite.conv_reduce', 8, 32, 1, bias=True)
    conv('second_stem.squeeze_excite.conv_expand', 32, 8, 1, bias=True)
    conv('second_stem.point_linear.conv', 16, 32, 1)
    for stage, sp in STAGE_SPEC.items():
        blocks = []
        for j, (ic, oc) in enumerate(zip(sp['ics'], sp['ocs']), start=1):
            block = 'block%d' % j
            for i, (k, _e, se_mult) in enumerate(CAND_SPEC):
                mc = mcs[stage][block][i]
                pre = '%s.%s.m_ops.%d.' % (stage, block, i)
                conv(pre + 'inverted_bottleneck.conv', mc, ic, 1)
                conv(pre + 'depth_conv.conv', mc, mc, k, groups=mc)
                if se_mult:
                    conv(pre + 'squeeze_excite.conv_reduce', se_mult * ic, mc, 1, bias=True)
                    conv(pre + 'squeeze_excite.conv_expand', mc, se_mult * ic, 1, bias=True)
                conv(pre + 'point_linear.conv', oc, mc, 1)
            blocks.append(block)
        # register order in the reference: blockN.log_alphas after each block's ops; betas last.
        for block in blocks:
            P['%s.%s.log_alphas' % (stage, block)] = F.log_softmax(torch.zeros(NUM_OPS), dim=-1).to(dtype)
        P[stage + '.betas'] = torch.zeros(len(blocks), dtype=dtype)
    conv('feature_mix_layer.conv', 1280, 320, 1)
    lin = nn.Linear(1280, num_classes)
    P['classifier.linear.weight'] = lin.weight.detach().to(dtype)
    P['classifier.linear.bias'] = torch.zeros(num_classes, dtype=dtype)
    torch.random.set_rng_state(g)
    return P


def is_arch_key(k):
    return k.endswith('log_alphas') or k.endswith('betas')


def search_unit_cpu(P, mcs, lut, batches, T, target_lat, lambda_lat, seed=2, state=None):
    """One 'search unit' = 2 iterations of train_w_arch (train_search.py:366-426): 2 bi-sampled w-steps + 1 alpha-step,
    each forward + backward + global-norm clip (5.0) + optimiser step -- SGD(lr .025, momentum .9, wd 1e-5) on the live
    weights (tensors without a gradient are skipped, as torch.optim does), Adam(lr .01, betas (.5, .999), wd 5e-4) on the
    architecture parameters followed by the log_softmax renormalisation (:421-422).  ``state`` (a dict kept by the caller)
    carries the momentum buffers / Adam moments between units.  Runs on whatever device P and the batches live on (the
    timed CPU baseline of bench.py, and its same-box GPU baseline).
    """
    state = state if state is not None else {}
    mom = state.setdefault('mom', {})
    adam = state.setdefault('adam', {})
    dev = next(iter(P.values())).device
    gen = torch.Generator().manual_seed(seed)
    rnd = _pyrandom.Random(seed)
    wkeys = [k for k in P if not is_arch_key(k)]
    akeys = [k for k in P if is_arch_key(k)]
    names = [('%s.%s.' % (st, bl)) for st, bl, *_ in block_shapes()]
    n_img = 0
    for it in range(2):
        x_w, t_w = batches[it % len(batches)]
        for k in wkeys:
            P[k].requires_grad_(True)
        for k in akeys:
            P[k].requires_grad_(False)
        noise = [draw_gumbel(generator=gen) for _ in range(18)]
        la_host = torch.stack([P[n + 'log_alphas'].detach() for n in names]).cpu()
        idx_g = [sample_gumbel_index(la_host[i], noise[i]) for i in range(len(names))]
        idx_r = []
        for ig in idx_g:
            rest = [j for j in range(NUM_OPS) if j != ig]
            idx_r.append(rest[rnd.choice(range(len(rest)))])
        lg, _ = network_forward(x_w, P, mcs, lut, True, T, indices=idx_g)
        lr, _ = network_forward(x_w, P, mcs, lut, True, T, indices=idx_r)
        loss = F.cross_entropy(lg, t_w) + F.cross_entropy(lr, t_w)
        grads = torch.autograd.grad(loss, [P[k] for k in wkeys], allow_unused=True)
        live = [(k, P[k], g) for k, g in zip(wkeys, grads) if g is not None]
        tot = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for _, _, g in live]))
        coef = torch.clamp(5.0 / (tot + 1e-6), max=1.0)
        with torch.no_grad():
            for k, p, g in live:
                d = g * coef + 1e-5 * p
                b = mom.get(k)
                b = d.clone() if b is None else b.mul_(0.9).add_(d)
                mom[k] = b
                p.add_(b, alpha=-0.025)
        n_img += x_w.shape[0]
        if it % 2 == 0:
            x_a, t_a = batches[(it + 1) % len(batches)]
            for k in wkeys:
                P[k].requires_grad_(False)
            for k in akeys:
                P[k].requires_grad_(True)
            noise = [draw_gumbel(generator=gen).to(dev) for _ in range(18)]
            la, lat = network_forward(x_a, P, mcs, lut, False, T, noise=noise)
            loss, _, _ = arch_loss(la, lat, t_a, target_lat, lambda_lat)
            grads = torch.autograd.grad(loss, [P[k] for k in akeys])
            tot = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads]))
            coef = torch.clamp(5.0 / (tot + 1e-6), max=1.0)
            state['t'] = t = state.get('t', 0) + 1
            with torch.no_grad():
                for k, g in zip(akeys, grads):
                    p = P[k]
                    g = g * coef + 5e-4 * p
                    m, v = adam.get(k, (torch.zeros_like(p), torch.zeros_like(p)))
                    m = m.lerp(g, 0.5)
                    v = v.mul(0.999).addcmul(g, g, value=0.001)
                    adam[k] = (m, v)
                    denom = v.sqrt() / (1 - 0.999 ** t) ** 0.5 + 1e-8
                    p.copy_(F.log_softmax(p - (0.01 / (1 - 0.5 ** t)) * m / denom, dim=-1))
    for k in P:
        P[k].requires_grad_(False)
    return n_img
