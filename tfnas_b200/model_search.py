"""Drop-in ``MixedOP`` / ``MixedStage`` / ``Network`` for the TF-NAS supernet search path.

Same constructors, attribute tree, parameter / state_dict names and ``forward(x, sampling, mode)``
contract as the reference ``models/model_search.py`` (MixedOP :32-122, MixedStage :125-210,
Network :213-364), so ``train_search.py``'s ``exec``-based weight surgery (:164-193, :235-258)
and ``parsing_model.get_op_and_depth_weights`` work unchanged.  The arithmetic of every MixedOP
and stage sink runs in libtfnas_b200.so through ``tfnas_b200.ops``; the candidate sub-modules
below are parameter holders only (they own ``nn.Conv2d`` objects for their default init and
names, and are never called).

Differences that are deliberate and documented in DESIGN.md:
  * Gumbel noise is drawn on the HOST (CPU generator), in forward order, so a seeded run is
    reproducible against the CPU oracle (the reference draws on the logits' device).
  * The 'gumbel' sampling index is computed from a host mirror of log_alphas (one D2H copy per
    arch-parameter version instead of one ``.item()`` sync per MixedOP per pass).
"""
import ctypes
import os
import random
import time
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .config import CAND_SPEC, PRIMITIVES, lut_key
from .ops import ArenaPool, BodyCall, BodyFn, HeadFn, StemFn, MixedOpCall, MixedOpFn, StageSinkFn, bn_act, dwconv

__all__ = ['PRIMITIVES', 'OPS', 'MixedOP', 'MixedStage', 'Network', 'MBInvertedResBlock', 'ConvLayer',
           'LinearLayer', 'NoisePlan', 'injected', 'seed_noise', 'draw_gumbel']


class _Holder(nn.Sequential):
    """nn.Sequential of named children; used only to reproduce the reference's attribute paths."""


class MBInvertedResBlock(nn.Module):
    """Parameter holder mirroring models/layers.py:431-537 (names, shapes, default inits, metadata)."""

    name = 'MBInvertedResBlock'

    def __init__(self, in_channels, mid_channels, se_channels, out_channels, kernel_size=3, stride=1,
                 groups=1, has_shuffle=False, bias=False, use_bn=True, affine=True, act_func='relu6'):
        super(MBInvertedResBlock, self).__init__()
        if groups != 1 or has_shuffle or bias or affine or not use_bn:
            raise NotImplementedError('search path uses groups=1, no shuffle, no bias, BN affine=False')
        self.in_channels, self.mid_channels, self.se_channels = in_channels, mid_channels, se_channels
        self.out_channels, self.kernel_size, self.stride = out_channels, kernel_size, stride
        self.groups, self.has_shuffle, self.bias, self.use_bn = groups, has_shuffle, bias, use_bn
        self.affine, self.act_func, self.drop_connect_rate = affine, act_func, 0.0
        if mid_channels > in_channels:
            self.inverted_bottleneck = _Holder(OrderedDict([
                ('conv', nn.Conv2d(in_channels, mid_channels, 1, 1, 0, bias=False))]))
        else:
            self.inverted_bottleneck = None
            self.mid_channels = mid_channels = in_channels
        self.depth_conv = _Holder(OrderedDict([
            ('conv', nn.Conv2d(mid_channels, mid_channels, kernel_size, stride, kernel_size // 2,
                               groups=mid_channels, bias=False))]))
        if se_channels > 0:
            self.squeeze_excite = _Holder(OrderedDict([
                ('conv_reduce', nn.Conv2d(mid_channels, se_channels, 1, 1, 0, bias=True)),
                ('conv_expand', nn.Conv2d(se_channels, mid_channels, 1, 1, 0, bias=True))]))
        else:
            self.squeeze_excite = None
            self.se_channels = 0
        self.point_linear = _Holder(OrderedDict([
            ('conv', nn.Conv2d(mid_channels, out_channels, 1, 1, 0, bias=False))]))
        self.has_residual = (in_channels == out_channels) and (stride == 1)

    def weight_list(self):
        w = [self.inverted_bottleneck.conv.weight, self.depth_conv.conv.weight, self.point_linear.conv.weight]
        if self.squeeze_excite is not None:
            w += [self.squeeze_excite.conv_reduce.weight, self.squeeze_excite.conv_reduce.bias,
                  self.squeeze_excite.conv_expand.weight, self.squeeze_excite.conv_expand.bias]
        return w

    def forward(self, x):
        raise RuntimeError('candidate blocks are evaluated inside the fused MixedOP kernels, not called')


OPS = {
    'MBI_k3_e3': lambda ic, mc, oc, s, aff, act: MBInvertedResBlock(ic, mc, 0, oc, 3, s, affine=aff, act_func=act),
    'MBI_k3_e6': lambda ic, mc, oc, s, aff, act: MBInvertedResBlock(ic, mc, 0, oc, 3, s, affine=aff, act_func=act),
    'MBI_k5_e3': lambda ic, mc, oc, s, aff, act: MBInvertedResBlock(ic, mc, 0, oc, 5, s, affine=aff, act_func=act),
    'MBI_k5_e6': lambda ic, mc, oc, s, aff, act: MBInvertedResBlock(ic, mc, 0, oc, 5, s, affine=aff, act_func=act),
    'MBI_k3_e3_se': lambda ic, mc, oc, s, aff, act: MBInvertedResBlock(ic, mc, ic, oc, 3, s, affine=aff, act_func=act),
    'MBI_k3_e6_se': lambda ic, mc, oc, s, aff, act: MBInvertedResBlock(ic, mc, ic * 2, oc, 3, s, affine=aff, act_func=act),
    'MBI_k5_e3_se': lambda ic, mc, oc, s, aff, act: MBInvertedResBlock(ic, mc, ic, oc, 5, s, affine=aff, act_func=act),
    'MBI_k5_e6_se': lambda ic, mc, oc, s, aff, act: MBInvertedResBlock(ic, mc, ic * 2, oc, 5, s, affine=aff, act_func=act),
}


def ctypes_copy(dst, src):
    """memberwise copy of a ctypes Structure (dst may be an element of an array inside another Structure)."""
    ctypes.memmove(ctypes.byref(dst), ctypes.byref(src), ctypes.sizeof(src))


class NoisePlan(object):
    """Optional injection of everything random in one supernet forward (tests / parity runs).

    ``noise``: list of [num_ops] CPU tensors consumed in forward order by alpha-mode and 'gumbel'
    sampling; ``indices``: list of candidate ids consumed by any sampling mode (overrides noise).
    """

    def __init__(self, noise=None, indices=None, device_noise=None):
        self.noise = list(noise) if noise is not None else None
        self.indices = list(indices) if indices is not None else None
        # [num_blocks, 8] device tensor the alpha-mode body reads in place (CUDA-graph replay: the caller refreshes its
        # contents before every replay, search_loop.GraphedAlphaStep)
        self.device_noise = device_noise
        self.pos = 0

    def next(self):
        i = self.pos
        self.pos += 1
        return (self.noise[i] if self.noise is not None else None,
                self.indices[i] if self.indices is not None else None)


_ACTIVE_PLAN = [None]


class injected(object):
    """with injected(NoisePlan(...)): model(x, ...)"""

    def __init__(self, plan):
        self.plan = plan

    def __enter__(self):
        self.prev = _ACTIVE_PLAN[0]
        _ACTIVE_PLAN[0] = self.plan
        return self.plan

    def __exit__(self, *a):
        _ACTIVE_PLAN[0] = self.prev


_NOISE = dict(gen=None, rnd=random)


def seed_noise(seed):
    """Give the search its own CPU streams for Gumbel noise and the 'random' mode (needed under
    data parallelism so every rank samples the same sub-network).  ``None`` restores the reference
    behaviour: global torch CPU generator + global python ``random``."""
    if seed is None:
        _NOISE['gen'], _NOISE['rnd'] = None, random
    else:
        _NOISE['gen'] = torch.Generator().manual_seed(int(seed))
        _NOISE['rnd'] = random.Random(int(seed))


def draw_gumbel(n, generator=None):
    """Same draw as F.gumbel_softmax's noise, from the CPU generator (reference :62,:87)."""
    generator = generator if generator is not None else _NOISE['gen']
    return -torch.empty(n).exponential_(generator=generator).log()


class MixedOP(nn.Module):
    def __init__(self, in_channels, out_channels, stride, affine, act_func, num_ops, mc_num_dict, lat_lookup):
        super(MixedOP, self).__init__()
        if affine:
            raise NotImplementedError('the search path uses affine=False BatchNorm only')
        self.num_ops = num_ops
        self.lat_lookup = lat_lookup
        self.mc_num_dict = mc_num_dict
        self.in_channels, self.out_channels, self.stride, self.act_func = in_channels, out_channels, stride, act_func
        self.m_ops = nn.ModuleList()
        for i in range(num_ops):
            self.m_ops.append(OPS[PRIMITIVES[i]](in_channels, self.mc_num_dict[i], out_channels, stride, affine, act_func))
        self._initialize_log_alphas()
        self.reset_switches()
        self._lat_cache = {}
        self._host_alpha = None   # (version, cpu tensor)

    # --- reference API --------------------------------------------------------------------
    def fink_ori_idx(self, idx):
        count = 0
        for ori_idx in range(len(self.switches)):
            if self.switches[ori_idx]:
                count += 1
                if count == (idx + 1):
                    break
        return ori_idx

    def _initialize_log_alphas(self):
        alphas = torch.zeros((self.num_ops,))
        self.register_parameter('log_alphas', nn.Parameter(F.log_softmax(alphas, dim=-1)))

    def reset_switches(self):
        self.switches = [True] * self.num_ops

    def set_temperature(self, T):
        self.T = T

    def get_lookup_latency(self, size):
        lats = []
        for op in self.m_ops:
            key = lut_key(size, op.in_channels, op.se_channels, op.out_channels, op.kernel_size, op.stride, op.act_func)
            lats.append(self.lat_lookup[key][op.mid_channels])
        return lats

    # --- host-side sampling ---------------------------------------------------------------
    def _alphas_on_host(self):
        v = self.log_alphas._version
        if self._host_alpha is None or self._host_alpha[0] != v or self._host_alpha[2] != self.log_alphas.data_ptr():
            self._host_alpha = (v, self.log_alphas.detach().float().cpu(), self.log_alphas.data_ptr())
        return self._host_alpha[1]

    def _sample_index(self, mode):
        noise, forced = _ACTIVE_PLAN[0].next() if _ACTIVE_PLAN[0] is not None else (None, None)
        live = [i for i, s in enumerate(self.switches) if s]
        if mode == 'gumbel':
            if forced is None:
                la = self._alphas_on_host()[live]
                g = noise[:len(live)] if noise is not None else draw_gumbel(len(live))
                # argmax softmax((log_softmax(a)+g)/T) == argmax(log_softmax(a)+g)
                forced = int(torch.argmax(F.log_softmax(la, dim=-1) + g).item())
            self.switches[forced] = False      # un-mapped, as in the reference (quirk Q2)
            return forced
        if mode == 'gumbel_2':
            if forced is None:
                la = self._alphas_on_host()[live]
                g = noise[:len(live)] if noise is not None else draw_gumbel(len(live))
                forced = self.fink_ori_idx(int(torch.argmax(F.log_softmax(la, dim=-1) + g).item()))
        elif mode == 'min_alphas':
            if forced is None:
                forced = self.fink_ori_idx(int(torch.argmin(self._alphas_on_host()[live]).item()))
        elif mode == 'max_alphas':
            if forced is None:
                forced = self.fink_ori_idx(int(torch.argmax(self._alphas_on_host()[live]).item()))
        elif mode == 'random':
            if forced is None:
                forced = self.fink_ori_idx(_NOISE['rnd'].choice(range(len(live))))
        else:
            raise ValueError('invalid sampling mode...')
        self.reset_switches()
        return forced

    # --- forward --------------------------------------------------------------------------
    def _call(self, x, mask, gumbel=None, lat8=None):
        N, _c, H, W = x.shape
        ops = self.m_ops
        return MixedOpCall(N, self.in_channels, self.out_channels, H, W, self.stride, self.act_func,
                           [op.mid_channels for op in ops], [op.kernel_size for op in ops],
                           [op.se_channels for op in ops], mask, getattr(self, 'T', 1.0), gumbel, lat8)

    def forward(self, x, sampling, mode):
        if sampling:
            idx = self._sample_index(mode)
            out, _ = MixedOpFn.apply(x, None, self._call(x, 1 << idx), *self.m_ops[idx].weight_list())
            return out, 0
        noise, _ = _ACTIVE_PLAN[0].next() if _ACTIVE_PLAN[0] is not None else (None, None)
        g = noise if noise is not None else draw_gumbel(self.num_ops)
        size = x.size(-1)
        key = (size, x.device)
        if key not in self._lat_cache:
            self._lat_cache[key] = torch.tensor(self.get_lookup_latency(size), dtype=torch.float32, device=x.device)
        gd = g.to(device=x.device, dtype=torch.float32, non_blocking=True)
        ws = []
        for op in self.m_ops:
            ws += op.weight_list()
        call = self._call(x, (1 << self.num_ops) - 1, gd, self._lat_cache[key])
        return MixedOpFn.apply(x, self.log_alphas, call, *ws)


class MixedStage(nn.Module):
    def __init__(self, ics, ocs, ss, affs, acts, mc_num_ddict, lat_lookup, stage_type):
        super(MixedStage, self).__init__()
        self.lat_lookup = lat_lookup
        self.mc_num_ddict = mc_num_ddict
        self.stage_type = stage_type  # 0 for stage6 || 1 for stage1 || 2 for stage2 || 3 for stage3/4/5
        self.start_res = 0 if ((ics[0] == ocs[0]) and (ss[0] == 1)) else 1
        self.num_res = len(ics) - self.start_res + 1
        nblocks = {0: 1, 1: 2, 2: 3, 3: 4}.get(stage_type)
        if nblocks is None:
            raise ValueError('invalid stage_type...')
        for j in range(nblocks):
            setattr(self, 'block%d' % (j + 1),
                    MixedOP(ics[j], ocs[j], ss[j], affs[j], acts[j], len(PRIMITIVES), mc_num_ddict['block%d' % (j + 1)], lat_lookup))
        self.nblocks = nblocks
        self._initialize_betas()

    def forward(self, x, sampling, mode):
        res_list = [x]
        lat_list = []
        cum = None
        out = x
        for j in range(self.nblocks):
            out, lat = getattr(self, 'block%d' % (j + 1))(out, sampling, mode)
            res_list.append(out)
            if not sampling:
                cum = lat if cum is None else cum + lat
                lat_list.append(cum)
        res = res_list[self.start_res:]
        if sampling:
            out, _ = StageSinkFn.apply(self.betas, None, *res)
            return out, 0
        lats = ([torch.zeros_like(lat_list[0])] if self.start_res == 0 else []) + lat_list
        return StageSinkFn.apply(self.betas, torch.stack(lats), *res)

    def _initialize_betas(self):
        self.register_parameter('betas', nn.Parameter(torch.zeros((self.num_res))))


class ConvLayer(nn.Module):
    """conv -> BN(batch stats, no affine) -> act, models/layers.py:190-256 with ops_order 'weight_bn_act'."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, affine=False, act_func='relu'):
        super(ConvLayer, self).__init__()
        self.in_channels, self.out_channels, self.kernel_size, self.stride = in_channels, out_channels, kernel_size, stride
        self.act_func = act_func
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, kernel_size // 2, bias=False)

    def forward(self, x):
        return bn_act(self.conv(x), self.act_func)


class LinearLayer(nn.Module):
    def __init__(self, in_features, out_features):
        super(LinearLayer, self).__init__()
        self.linear = nn.Linear(in_features, out_features, True)

    def forward(self, x):
        return self.linear(x)


class _StemBlock(MBInvertedResBlock):
    """second_stem: MBConv(32,32,se 8,16,k3,s1,relu) without expand conv (models/model_search.py:220).
    Depthwise conv and both BN(+act) run in the library; the tiny SE FCs and the 32->16 1x1 conv stay on torch."""

    def forward(self, x):
        x = bn_act(dwconv(x, self.depth_conv.conv.weight), 'relu')
        g = F.adaptive_avg_pool2d(x, 1)
        g = self.squeeze_excite.conv_expand(F.relu(self.squeeze_excite.conv_reduce(g)))
        x = x * torch.sigmoid(g)
        return bn_act(self.point_linear.conv(x), None)


class Network(nn.Module):
    def __init__(self, num_classes, mc_num_dddict, lat_lookup):
        super(Network, self).__init__()
        self.lat_lookup = lat_lookup
        self.mc_num_dddict = mc_num_dddict
        from .config import STAGE_SPEC
        self.first_stem = ConvLayer(3, 32, kernel_size=3, stride=2, affine=False, act_func='relu')
        self.second_stem = _StemBlock(32, 32, 8, 16, kernel_size=3, stride=1, affine=False, act_func='relu')
        for stage, sp in STAGE_SPEC.items():
            n = len(sp['ics'])
            setattr(self, stage, MixedStage(ics=sp['ics'], ocs=sp['ocs'], ss=sp['ss'], affs=[False] * n,
                                            acts=[sp['act']] * n, mc_num_ddict=mc_num_dddict[stage],
                                            lat_lookup=lat_lookup, stage_type=sp['stage_type']))
        self.feature_mix_layer = ConvLayer(320, 1280, kernel_size=1, stride=1, affine=False, act_func='swish')
        self.global_avg_pooling = nn.AdaptiveAvgPool2d(1)
        self.classifier = LinearLayer(1280, num_classes)
        # TFNAS_BODY=0: evaluate the stages through the per-MixedOP autograd Functions (MixedStage.forward) instead of the
        # C++ body executor -- same kernels, same results; kept for A/B tests
        self.use_body = os.environ.get('TFNAS_BODY', '1') != '0'
        self._initialization()

    def forward(self, x, sampling, mode='max'):
        return self.forward_from_stem(self.forward_stems(x), sampling, mode)

    def forward_stems(self, x):
        """first_stem + second_stem (models/model_search.py:283-284).  Split out so a bi-sampled w-step can evaluate the
        stems ONCE for its two sampled sub-networks: both see the same batch and the same stem weights, so the stem output
        (and, by linearity, one backward over the summed gradient) is shared -- identical to two full forwards."""
        if self.use_body:
            ss = self.second_stem
            return StemFn.apply(x, self.__dict__.setdefault('_arena_pool', ArenaPool()), self.first_stem.conv.weight,
                                ss.depth_conv.conv.weight, ss.squeeze_excite.conv_reduce.weight,
                                ss.squeeze_excite.conv_reduce.bias, ss.squeeze_excite.conv_expand.weight,
                                ss.squeeze_excite.conv_expand.bias, ss.point_linear.conv.weight)
        # stems / head on cuDNN / cuBLAS run in true fp32 (TF32 is switched off package-wide, tfnas_b200/__init__.py)
        return self.second_stem(self.first_stem(x))

    def forward_from_stem(self, x, sampling, mode='max'):
        """stage1..6, feature mix, pooling, classifier (models/model_search.py:285-304) from the stem output."""
        out_lat = self.lat_lookup['base'] if not sampling else 0.0
        if sampling and mode in ('gumbel', 'gumbel_2', 'min_alphas', 'max_alphas'):
            self.refresh_host_alphas()         # one device->host copy for all 18 MixedOPs instead of one each
        if self.use_body:
            x, lat = self._body(x, sampling, mode)
            out_lat = out_lat + lat
        else:
            for s in range(1, 7):
                x, lat = getattr(self, 'stage%d' % s)(x, sampling, mode)
                out_lat = out_lat + lat
        if self.use_body:
            x = HeadFn.apply(x, self.__dict__.setdefault('_arena_pool', ArenaPool()), self.feature_mix_layer.conv.weight,
                             self.classifier.linear.weight, self.classifier.linear.bias)
            return x, out_lat
        x = self.feature_mix_layer(x)
        x = self.global_avg_pooling(x)
        x = x.view(x.size(0), -1)
        x = self.classifier(x)
        return x, out_lat

    # --- the six MixedStages in one library call per direction (csrc/body.cu) ----------------------------------------
    def _stages(self):
        return [getattr(self, 'stage%d' % s) for s in range(1, 7)]

    def _body_desc(self, x):
        key = (tuple(x.shape), x.device)
        c = self.__dict__.setdefault('_body_cache', {})
        if key not in c:
            d = _lib.BodyDesc()
            stages = self._stages()
            d.num_stages = len(stages)
            N, _c, H, W = x.shape
            bi = 0
            sizes_in = []
            for si, st in enumerate(stages):
                if st.start_res != 1:
                    raise _lib.TfnasError('the body executor sums the outputs of all blocks of a stage (start_res == 1)')
                d.stage_blocks[si] = st.nblocks
                for j in range(st.nblocks):
                    m = getattr(st, 'block%d' % (j + 1))
                    call = m._call(torch.empty((N, m.in_channels, H, W), device='meta'), (1 << m.num_ops) - 1)
                    ctypes_copy(d.op[bi], call.desc)
                    sizes_in.append(H)
                    _n, _oc, H, W = call.out_shape()
                    bi += 1
            d.num_blocks = bi
            out_shape = (N, d.op[bi - 1].oc, H, W)
            lib = _lib.load()
            full = _lib.BodyMasks(*[(1 << d.op[i].num_ops) - 1 for i in range(bi)])
            # the largest candidate of every MixedOP (k5, e6, SE: the last primitive) bounds the arena of any sampled pass
            big = _lib.BodyMasks(*[1 << (d.op[i].num_ops - 1) for i in range(bi)])
            sizes = dict(alpha=lib.tfnas_body_arena_bytes(ctypes.byref(d), full, 0),
                         sampled=lib.tfnas_body_arena_bytes(ctypes.byref(d), big, 1))
            if not sizes['alpha'] or not sizes['sampled']:
                _lib.check(-1)
            sizes['grad_numel'] = sum(max(sum(w.numel() for w in op.weight_list()) for op in m.m_ops)
                                      for m in self._param_lists()[4])
            c[key] = [d, bi, out_shape, None, sizes, sizes_in]
        return c[key]

    def draw_alpha_noise(self, plan=None):
        """[num_blocks, 8] CPU tensor of the Gumbel draws of one alpha-mode forward, in forward order (one draw of num_ops
        values per MixedOP: the same consumption of the CPU generator as 18 separate F.gumbel_softmax calls)."""
        noise = []
        for m in self._param_lists()[4]:
            g = plan.next()[0] if plan is not None else None
            noise.append(g if g is not None else draw_gumbel(m.num_ops))
        return torch.stack([F.pad(g.float(), (0, _lib.MAX_OPS - g.numel())) for g in noise])

    def _body_lat_table(self, entry, device):
        """[num_blocks, 8] LUT latencies of every candidate at this input size (alpha mode only: the sampled passes never
        touch the LUT, models/model_search.py:84-85, so an image size the LUT does not cover is fine there)."""
        if entry[3] is None:
            mops = self._param_lists()[4]
            rows = [m.get_lookup_latency(h) + [0.0] * (_lib.MAX_OPS - m.num_ops) for m, h in zip(mops, entry[5])]
            entry[3] = torch.tensor(rows, dtype=torch.float32, device=device)
        return entry[3]

    def _body(self, x, sampling, mode):
        entry = self._body_desc(x)
        d, nb, out_shape, _lat, sizes = entry[:5]
        mops = self._param_lists()[4]
        pool = self.__dict__.setdefault('_arena_pool', ArenaPool())
        betas = [st.betas for st in self._stages()]
        T = getattr(mops[0], 'T', 1.0)
        if sampling:
            idx = [m._sample_index(mode) for m in mops]
            tensors = []
            n_per = []
            for m, i in zip(mops, idx):
                wl = m.m_ops[i].weight_list()
                tensors += wl
                n_per.append([len(wl)])
            masks = [1 << i for i in idx]
            # the pool hands out blocks of the "widest candidates" size so every step leases the same block; after elastic
            # rescaling another candidate can need more (wider, or SE state): the exact size of this pass decides
            need = max(sizes['sampled'], _lib.load().tfnas_body_arena_bytes(ctypes.byref(d), _lib.BodyMasks(*masks), 1))
            call = BodyCall(d, nb, len(betas), masks, False, T, None, None, [[i] for i in idx], n_per, pool, need, out_shape)
            call.grad_numel = sizes['grad_numel']
            out, _ = BodyFn.apply(x, call, *(tensors + betas))
            return out, 0.0
        plan = _ACTIVE_PLAN[0]
        if plan is not None and plan.device_noise is not None:
            gd = plan.device_noise
        else:
            gd = self.draw_alpha_noise(plan).to(x.device, non_blocking=True)
        tensors, n_per, active = [], [], []
        for m in mops:
            per = []
            for op in m.m_ops:
                wl = op.weight_list()
                tensors += wl
                per.append(len(wl))
            n_per.append(per)
            active.append(list(range(m.num_ops)))
        masks = [(1 << m.num_ops) - 1 for m in mops]
        lat = self._body_lat_table(entry, x.device)
        call = BodyCall(d, nb, len(betas), masks, True, T, gd, lat, active, n_per, pool, sizes['alpha'], out_shape)
        return BodyFn.apply(x, call, *(tensors + [m.log_alphas for m in mops] + betas))

    def set_temperature(self, T):
        for m in self.modules():
            if isinstance(m, MixedOP):
                m.set_temperature(T)

    def refresh_host_alphas(self):
        """Bring the host mirrors of every MixedOP's log_alphas up to date with ONE device->host copy.

        The 'gumbel' index is drawn on the host (MixedOP._sample_index), so after an alpha update every MixedOP needs its
        new log_alphas on the host.  Fetched lazily per MixedOP that is 18 synchronising copies per pass, each of which
        drains the stream (57 ms of a 124 ms search unit in profiles/host_profile_r1.txt); stacked, it is one."""
        stale = []
        for m in self._param_lists()[4]:
            h = m._host_alpha
            if h is None or h[0] != m.log_alphas._version or h[2] != m.log_alphas.data_ptr():
                stale.append(m)
        if not stale:
            return
        if len(set(m.log_alphas.numel() for m in stale)) == 1 and len(set(m.log_alphas.device for m in stale)) == 1:
            t0 = time.perf_counter()
            flat = torch.stack([m.log_alphas.detach().float().reshape(-1) for m in stale]).cpu()
            # the copy drains the stream (it waits for the alpha update it reads): host time spent BLOCKED here is GPU time,
            # not host work -- bench.py subtracts it from the enqueue time it reports
            self.__dict__['host_blocked_s'] = self.__dict__.get('host_blocked_s', 0.0) + (time.perf_counter() - t0)
            for m, row in zip(stale, flat):
                m._host_alpha = (m.log_alphas._version, row.clone(), m.log_alphas.data_ptr())
        else:                                   # mixed widths / devices: per-module copies
            for m in stale:
                m._alphas_on_host()

    # The reference filters named_parameters() by suffix on every call (models/model_search.py:306-350); walking the
    # 1282-module tree costs ~2.7 ms of host time and the search loop asks a dozen times per step, so the four lists are
    # built once and handed out as copies.  The set of Parameter objects only changes when modules are added or the
    # tensors are re-wrapped (_apply: .cuda() / .to()), which drops the cache; invalidate_param_cache() does it by hand.
    def _param_lists(self):
        c = self.__dict__.get('_plists')
        if c is None:
            w, la, be = [], [], []
            for k, v in self.named_parameters():
                if k.endswith('log_alphas'):
                    la.append((k, v))
                elif k.endswith('betas'):
                    be.append((k, v))
                else:
                    w.append(v)
            arch = [v for k, v in self.named_parameters() if k.endswith('log_alphas') or k.endswith('betas')]
            c = self.__dict__['_plists'] = (w, arch, [v for _k, v in la], [v for _k, v in be],
                                            [m for m in self.modules() if isinstance(m, MixedOP)])
        return c

    def invalidate_param_cache(self):
        self.__dict__.pop('_plists', None)

    def _apply(self, fn, *args, **kwargs):
        self.invalidate_param_cache()
        r = super(Network, self)._apply(fn, *args, **kwargs)
        self.invalidate_param_cache()
        return r

    def weight_parameters(self):
        return list(self._param_lists()[0])

    def arch_parameters(self):
        return list(self._param_lists()[1])

    def log_alphas_parameters(self):
        return list(self._param_lists()[2])

    def betas_parameters(self):
        return list(self._param_lists()[3])

    def reset_switches(self):
        for m in self._param_lists()[4]:
            m.reset_switches()

    def _initialization(self):
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)) and m.bias is not None:
                nn.init.constant_(m.bias, 0)
