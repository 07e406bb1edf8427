"""Latency look-up table of the TF-NAS search space measured on THIS GPU (reference
latency_pkl/make_lat_lut_example.py:44-492): for each of the 66 block keys `MBInvertedResBlock_{size}_{ic}_{se}_{oc}_k{k}_s{s}_{act}`
and every mid width mc = 1 .. max, the inference latency in ms of that MBConv block at batch 32 (fp32 NCHW, eval-mode
BatchNorm with affine, cudnn.benchmark, as the reference measures it), plus 'base' = first conv + second stem + feature mix +
pooling + classifier.  The shipped latency_gpu.pkl is a Titan-RTX table: `--target_lat 15 / 18` only means something against a
table of the hardware the search runs on.

Differences from the reference script, on purpose:
  * CUDA events around the timed iterations instead of time.time() without a device synchronisation (which times the launch);
  * the mid width is sampled every `--mc_step` channels and the table is completed by a least-squares line per key (the shipped
    pickle is visibly such a fit: it is linear in mc and goes negative at mc = 1); `--mc_step 1 --fit none` measures every width;
  * keys are striped over ranks (torchrun / RANK, WORLD_SIZE): embarrassingly parallel over the GPUs of a box.

    python tools/make_lat_lut.py --out profiles/latency_b200.npz [--pickle latency_b200.pkl] [--mc_step 8] [--iters 30]
    python -m torch.distributed.run --nproc-per-node 8 tools/make_lat_lut.py --out profiles/latency_b200.npz
"""
import argparse
import os
import pickle
import sys
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfnas_b200 import config  # noqa: E402
from tfnas_b200.config import CAND_SPEC  # noqa: E402


class Swish(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(x)


def act(name):
    return nn.ReLU(inplace=True) if name == 'relu' else Swish()


class MBConvEval(nn.Module):
    """Inference form of models/layers.py:431-561 (affine BN with running statistics)."""

    def __init__(self, ic, mc, se, oc, k, s, act_func):
        super().__init__()
        # the search-space blocks always carry the expand conv, whatever the mid width
        self.expand = nn.Sequential(nn.Conv2d(ic, mc, 1, bias=False), nn.BatchNorm2d(mc), act(act_func))
        self.dw = nn.Sequential(nn.Conv2d(mc, mc, k, s, k // 2, groups=mc, bias=False), nn.BatchNorm2d(mc), act(act_func))
        self.se = None
        if se > 0:
            self.se = nn.ModuleList([nn.Conv2d(mc, se, 1), act(act_func), nn.Conv2d(se, mc, 1)])
        self.project = nn.Sequential(nn.Conv2d(mc, oc, 1, bias=False), nn.BatchNorm2d(oc))
        self.res = ic == oc and s == 1

    def forward(self, x):
        y = self.dw(self.expand(x))
        if self.se is not None:
            g = F.adaptive_avg_pool2d(y, 1)
            y = y * torch.sigmoid(self.se[2](self.se[1](self.se[0](g))))
        y = self.project(y)
        return y + x if self.res else y


def measure(model, shape, iters, warm, dev):
    model = model.to(dev).eval()
    x = torch.randn(shape, device=dev)
    with torch.no_grad():
        for _ in range(warm):
            model(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            model(x)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def key_table():
    """key -> (size, ic, se, oc, k, s, act, max mc) for the 66 distinct keys, in the reference's order."""
    mx = config.get_mc_num_dddict(config.mc_mask_dddict, is_max=True)
    out = OrderedDict()
    for stage, block, ic, oc, s, a, size in config.block_shapes():
        for i, (k, _e, sm) in enumerate(CAND_SPEC):
            key = config.lat_lookup_key_dddict[stage][block][i]
            m = mx[stage][block][i]
            if key not in out or out[key][-1] < m:
                out[key] = (size, ic, sm * ic, oc, k, s, a, m)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='profiles/latency_b200.npz')
    ap.add_argument('--pickle', default=None, help='also write the reference pickle format here')
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--mc_step', type=int, default=8)
    ap.add_argument('--fit', default='linear', choices=['linear', 'interp', 'none'])
    ap.add_argument('--iters', type=int, default=30)
    ap.add_argument('--warm', type=int, default=10)
    ap.add_argument('--keys', type=int, default=0, help='only the first N keys (smoke runs)')
    a = ap.parse_args()
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True          # the deployed (derived) network runs on PyTorch defaults
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29541')
        dist.init_process_group('gloo', rank=rank, world_size=world)
    table = key_table()
    keys = list(table)[:a.keys] if a.keys else list(table)
    mine = OrderedDict()
    for key in keys[rank::world]:
        size, ic, se, oc, k, s, actf, mmax = table[key]
        grid = sorted(set([1] + list(range(a.mc_step, mmax + 1, a.mc_step)) + [mmax])) if a.fit != 'none' or a.mc_step > 1 \
            else list(range(1, mmax + 1))
        lat = [measure(MBConvEval(ic, mc, se, oc, k, s, actf), (a.batch, ic, size, size), a.iters, a.warm, dev) for mc in grid]
        allm = np.arange(1, mmax + 1, dtype=np.float64)
        if a.fit == 'linear' and len(grid) > 1:
            slope, icpt = np.polyfit(np.array(grid, dtype=np.float64), np.array(lat), 1)
            vals = slope * allm + icpt
        else:
            vals = np.interp(allm, np.array(grid, dtype=np.float64), np.array(lat))
        mine[key] = vals
        print('[rank %d] %-48s mc 1..%-4d %7.3f .. %7.3f ms (%d widths measured)' % (rank, key, mmax, vals[0], vals[-1], len(grid)),
              flush=True)
    base = None
    if rank == 0:
        stem1 = nn.Sequential(nn.Conv2d(3, 32, 3, 2, 1, bias=False), nn.BatchNorm2d(32), nn.ReLU(inplace=True))

        class Stem2(MBConvEval):
            def forward(self, x):            # mid == in channels: no expand conv (models/layers.py:479-482)
                y = self.dw(x)
                g = F.adaptive_avg_pool2d(y, 1)
                y = y * torch.sigmoid(self.se[2](self.se[1](self.se[0](g))))
                return self.project(y)
        fm = nn.Sequential(nn.Conv2d(320, 1280, 1, bias=False), nn.BatchNorm2d(1280), Swish())
        B = a.batch
        base = (measure(stem1, (B, 3, 224, 224), a.iters, a.warm, dev) + measure(Stem2(32, 32, 8, 16, 3, 1, 'relu'), (B, 32, 112, 112), a.iters, a.warm, dev)
                + measure(fm, (B, 320, 7, 7), a.iters, a.warm, dev) + measure(nn.AdaptiveAvgPool2d(1), (B, 1280, 7, 7), a.iters, a.warm, dev)
                + measure(nn.Linear(1280, 1000), (B, 1280), a.iters, a.warm, dev))
    if world > 1:
        import torch.distributed as dist
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        merged = {}
        for p in parts:
            merged.update(p)
        dist.destroy_process_group()
    else:
        merged = mine
    if rank != 0:
        return
    lut = OrderedDict([('base', float(base))])
    for key in keys:
        lut[key] = OrderedDict((m + 1, float(v)) for m, v in enumerate(merged[key]))
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    np.savez_compressed(a.out, base=np.float64(lut['base']), keys=np.array(keys), lens=np.array([len(lut[k]) for k in keys], dtype=np.int32),
                        vals=np.concatenate([np.array(list(lut[k].values()), dtype=np.float64) for k in keys]),
                        gpu=np.array(torch.cuda.get_device_name(local)), batch=np.int32(a.batch), mc_step=np.int32(a.mc_step), fit=np.array(a.fit))
    if a.pickle:
        with open(a.pickle, 'wb') as f:
            pickle.dump(lut, f)
    # latency of the initial-width supernet's argmax-free "all e6 k5 se" path, for orientation against target_lat
    print('base %.3f ms; %d keys, %d entries -> %s' % (lut['base'], len(keys), sum(len(lut[k]) for k in keys), a.out))


if __name__ == '__main__':
    main()
