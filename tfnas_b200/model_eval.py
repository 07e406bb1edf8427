"""Derived (discrete) network for re-training a searched architecture — SURVEY §8 row f-4.

Host-side mirror of the reference's `models/model_eval.py`: `Network` (:31, built from a parsed architecture and the
per-op mid widths) and `NetworkCfg` (:247, built from a `model.config` JSON).  Same constructor arguments, the same
module tree (hence the same `state_dict` keys, so checkpoints move between the two implementations in either
direction), the same `config` dictionary, `get_lookup_latency` and drop-connect / dropout behaviour
(`models/layers.py:539-561`, `tools/utils.py:77-86`).

Scope note: the layers here are plain `torch.nn` modules (BatchNorm with affine parameters and running statistics is
not what the search-path kernels of this library implement), meant to run under `torch.autocast(bfloat16)` in
channels-last layout — BASELINE config 5.  What IS native on this path is the step glue: the label-smoothing
cross-entropy (`step.FusedCrossEntropy(label_smooth=...)`) and the clip + momentum-SGD update (`step.FusedSGD`).
"""
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from .model_search import PRIMITIVES

# (stage, ((ic, oc, stride), ...), activation) — models/model_eval.py:46-81
STAGE_TABLE = (
    ('stage1', ((16, 24, 2), (24, 24, 1)), 'relu'),
    ('stage2', ((24, 40, 2), (40, 40, 1), (40, 40, 1)), 'swish'),
    ('stage3', ((40, 80, 2), (80, 80, 1), (80, 80, 1), (80, 80, 1)), 'swish'),
    ('stage4', ((80, 112, 1), (112, 112, 1), (112, 112, 1), (112, 112, 1)), 'swish'),
    ('stage5', ((112, 192, 2), (192, 192, 1), (192, 192, 1), (192, 192, 1)), 'swish'),
    ('stage6', ((192, 320, 1),), 'swish'),
)
STAGE_NAMES = tuple(s[0] for s in STAGE_TABLE)

_ACTS = {'relu': nn.ReLU, 'relu6': nn.ReLU6, 'swish': nn.SiLU, 'h-swish': nn.Hardswish}


def _act(name):
    return _ACTS[name](inplace=True) if name in _ACTS else None


def _bn(ch, affine):
    return nn.BatchNorm2d(ch, affine=affine, track_running_stats=affine)


def _seq(**named):
    return nn.Sequential(OrderedDict((k, v) for k, v in named.items() if v is not None))


def drop_connect(x, training, rate):
    """Stochastic depth per image: keep a sample's branch with probability 1-rate and rescale (tools/utils.py:77-86)."""
    if not training or rate <= 0.0:
        return x
    keep = 1.0 - rate
    gate = torch.floor(keep + torch.rand(x.shape[0], 1, 1, 1, dtype=x.dtype, device=x.device))
    return x / keep * gate


class ConvLayer(nn.Module):
    """conv -> BN -> act (models/layers.py:190-265; only the default 'weight_bn_act' order is used by the derived net)."""
    name = 'ConvLayer'

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, groups=1, has_shuffle=False, bias=False,
                 use_bn=True, affine=True, act_func='relu6', ops_order='weight_bn_act'):
        super().__init__()
        if ops_order != 'weight_bn_act' or has_shuffle or groups != 1:
            raise NotImplementedError('derived networks use plain weight_bn_act convolutions')
        self._cfg = dict(kernel_size=kernel_size, stride=stride, groups=groups, has_shuffle=has_shuffle, bias=bias,
                         in_channels=in_channels, out_channels=out_channels, use_bn=use_bn, affine=affine,
                         act_func=act_func, ops_order=ops_order)
        self.stride, self.kernel_size = stride, kernel_size
        self.bn = _bn(out_channels, affine) if use_bn else None      # registered before the conv, as BasicLayer does:
        self.act = _act(act_func)                                    # state_dict order is part of the checkpoint format
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, kernel_size // 2, bias=bias)

    def forward(self, x):
        x = self.conv(x)
        if self.bn is not None:
            x = self.bn(x)
        return x if self.act is None else self.act(x)

    @property
    def config(self):
        return dict(name=self.name, **self._cfg)


class LinearLayer(nn.Module):
    """models/layers.py:322-422 without the unused BN / activation options."""
    name = 'LinearLayer'

    def __init__(self, in_features, out_features, bias=True, use_bn=False, affine=False, act_func=None,
                 ops_order='weight_bn_act'):
        super().__init__()
        if use_bn or act_func is not None:
            raise NotImplementedError('the classifier of a derived network is a plain linear layer')
        self._cfg = dict(in_features=in_features, out_features=out_features, bias=bias, use_bn=use_bn, affine=affine,
                         act_func=act_func, ops_order=ops_order)
        self.linear = nn.Linear(in_features, out_features, bias)

    def forward(self, x):
        return self.linear(x)

    @property
    def config(self):
        return dict(name=self.name, **self._cfg)


class MBInvertedResBlock(nn.Module):
    """MBConv with affine BN and running statistics (models/layers.py:431-561)."""
    name = 'MBInvertedResBlock'

    def __init__(self, in_channels, mid_channels, se_channels, out_channels, kernel_size=3, stride=1, groups=1,
                 has_shuffle=False, bias=False, use_bn=True, affine=True, act_func='relu6'):
        super().__init__()
        if groups != 1 or has_shuffle:
            raise NotImplementedError('grouped / shuffled MBConv blocks are not part of the TF-NAS space')
        expand = mid_channels > in_channels
        if not expand:
            mid_channels = in_channels
        self.in_channels, self.mid_channels, self.out_channels = in_channels, mid_channels, out_channels
        self.se_channels = max(se_channels, 0)
        self.kernel_size, self.stride, self.act_func = kernel_size, stride, act_func
        self.groups, self.has_shuffle, self.bias, self.use_bn, self.affine = groups, has_shuffle, bias, use_bn, affine
        self.drop_connect_rate = 0.0
        bn = (lambda c: _bn(c, affine)) if use_bn else (lambda c: None)
        self.inverted_bottleneck = _seq(conv=nn.Conv2d(in_channels, mid_channels, 1, bias=bias), bn=bn(mid_channels),
                                        act=_act(act_func)) if expand else None
        self.depth_conv = _seq(conv=nn.Conv2d(mid_channels, mid_channels, kernel_size, stride, kernel_size // 2,
                                              groups=mid_channels, bias=bias), bn=bn(mid_channels), act=_act(act_func))
        self.squeeze_excite = _seq(conv_reduce=nn.Conv2d(mid_channels, self.se_channels, 1), act=_act(act_func),
                                   conv_expand=nn.Conv2d(self.se_channels, mid_channels, 1)) if self.se_channels else None
        self.point_linear = _seq(conv=nn.Conv2d(mid_channels, out_channels, 1, bias=bias), bn=bn(out_channels))
        self.has_residual = in_channels == out_channels and stride == 1

    def forward(self, x):
        y = x if self.inverted_bottleneck is None else self.inverted_bottleneck(x)
        y = self.depth_conv(y)
        if self.squeeze_excite is not None:
            y = y * torch.sigmoid(self.squeeze_excite(F.adaptive_avg_pool2d(y, 1)))
        y = self.point_linear(y)
        if self.has_residual:
            y = drop_connect(y, self.training, self.drop_connect_rate) + x
        return y

    @property
    def config(self):
        keys = ('in_channels', 'mid_channels', 'se_channels', 'out_channels', 'kernel_size', 'stride', 'groups',
                'has_shuffle', 'bias', 'use_bn', 'affine', 'act_func')
        return dict(name=self.name, **{k: getattr(self, k) for k in keys})

    def lut_key(self, size):
        """Key grammar of the latency table (models/model_eval.py:143-152)."""
        return '{}_{}_{}_{}_{}_k{}_s{}_{}'.format(self.name, size, self.in_channels, self.se_channels, self.out_channels,
                                                  self.kernel_size, self.stride, self.act_func)


_LAYERS = {c.name: c for c in (ConvLayer, LinearLayer, MBInvertedResBlock)}


def set_layer_from_config(layer_config):
    """models/layers.py:10-23 (does not consume the caller's dictionary)."""
    if layer_config is None:
        return None
    cfg = dict(layer_config)
    return _LAYERS[cfg.pop('name')](**cfg)


def _se_channels(op_idx, ic):
    """SE width of candidate `op_idx` (models/model_eval.py:21-28): none for 0-3, ic for the e3 ops, 2*ic for e6."""
    return 0 if op_idx < 4 else ic * (2 if op_idx % 2 else 1)


class _Derived(nn.Module):
    """Shared body of Network / NetworkCfg: the forward pass, the latency lookup, `config`, initialisation."""

    def _finish(self, drop_connect_rate):
        blocks = [self.second_stem] + [b for s in STAGE_NAMES for b in getattr(self, s)]
        self.block_count = len(blocks)
        for i, b in enumerate(blocks, 1):                # models/model_eval.py:44,96: rate grows linearly with depth
            b.drop_connect_rate = drop_connect_rate * i / self.block_count
        self.global_avg_pooling = nn.AdaptiveAvgPool2d(1)
        for m in self.modules():                          # models/model_eval.py:229-242
            if isinstance(m, (nn.Conv2d, nn.Linear)) and m.bias is not None:
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d) and m.weight is not None:
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def blocks(self):
        for s in STAGE_NAMES:
            for b in getattr(self, s):
                yield b

    def forward(self, x):
        x = self.second_stem(self.first_stem(x))
        for b in self.blocks():
            x = b(x)
        x = self.global_avg_pooling(self.feature_mix_layer(x)).flatten(1)
        if self.dropout_rate > 0.0:
            x = F.dropout(x, p=self.dropout_rate, training=self.training)
        return self.classifier(x)

    def get_lookup_latency(self, x):
        """Sum of the table entries of the blocks at the feature-map size each one sees (models/model_eval.py:133-208);
        the sizes follow from the strides, so nothing is executed."""
        if not self.lat_lookup:
            return 0.0
        size = x.size(-1) if torch.is_tensor(x) else int(x)
        size = (size + 2 * (self.first_stem.kernel_size // 2) - self.first_stem.kernel_size) // self.first_stem.stride + 1
        size = (size - 1) // self.second_stem.stride + 1
        lat = self.lat_lookup['base']
        for b in self.blocks():
            lat += self.lat_lookup[b.lut_key(size)][b.mid_channels]
            size = (size - 1) // b.stride + 1
        return lat

    @property
    def config(self):
        cfg = OrderedDict(first_stem=self.first_stem.config, second_stem=self.second_stem.config)
        for s in STAGE_NAMES:
            cfg[s] = [b.config for b in getattr(self, s)]
        cfg['feature_mix_layer'] = self.feature_mix_layer.config
        cfg['classifier'] = self.classifier.config
        return dict(cfg)


class Network(_Derived):
    """Derived network of a parsed architecture: `parsed_arch[stage][block] = op index`,
    `mc_num_dddict[stage][block][op] = mid width` (models/model_eval.py:31-110)."""

    def __init__(self, num_classes, parsed_arch, mc_num_dddict, lat_lookup=None, dropout_rate=0.0, drop_connect_rate=0.0):
        super().__init__()
        self.lat_lookup, self.mc_num_dddict, self.parsed_arch = lat_lookup, mc_num_dddict, parsed_arch
        self.dropout_rate, self.drop_connect_rate = dropout_rate, drop_connect_rate
        self.first_stem = ConvLayer(3, 32, kernel_size=3, stride=2, affine=True, act_func='relu')
        self.second_stem = MBInvertedResBlock(32, 32, 8, 16, kernel_size=3, stride=1, affine=True, act_func='relu')
        for stage, shapes, act in STAGE_TABLE:
            blocks = nn.ModuleList()
            for (ic, oc, stride), (block, op_idx) in zip(shapes, parsed_arch[stage].items()):
                k = 5 if 'k5' in PRIMITIVES[op_idx] else 3
                blocks.append(MBInvertedResBlock(ic, mc_num_dddict[stage][block][op_idx], _se_channels(op_idx, ic), oc,
                                                 k, stride, affine=True, act_func=act))
            setattr(self, stage, blocks)
        self.feature_mix_layer = ConvLayer(320, 1280, kernel_size=1, stride=1, affine=True, act_func='swish')
        self.classifier = LinearLayer(1280, num_classes)
        self._finish(drop_connect_rate)


class NetworkCfg(_Derived):
    """Derived network from a `model.config` dictionary as `Network.config` / the reference write it
    (models/model_eval.py:247-300)."""

    def __init__(self, num_classes, model_config, lat_lookup=None, dropout_rate=0.0, drop_connect_rate=0.0):
        super().__init__()
        self.lat_lookup, self.model_config = lat_lookup, model_config
        self.dropout_rate, self.drop_connect_rate = dropout_rate, drop_connect_rate
        self.first_stem = set_layer_from_config(model_config['first_stem'])
        self.second_stem = set_layer_from_config(model_config['second_stem'])
        for stage in STAGE_NAMES:
            setattr(self, stage, nn.ModuleList(set_layer_from_config(c) for c in model_config.get(stage, ())))
        self.feature_mix_layer = set_layer_from_config(model_config['feature_mix_layer'])
        self.classifier = set_layer_from_config(dict(model_config['classifier'], out_features=num_classes))
        self._finish(drop_connect_rate)
