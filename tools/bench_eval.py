#!/usr/bin/env python
"""Throughput of derived-network re-training (BASELINE config 5 shape: 3x224x224, bf16 autocast, channels-last) on one GPU:
images/s of eval_loop.train_step with the library's fused step glue (label-smoothing CE + clip + SGD) and with the stock
torch glue, same network, synthetic data resident on the device.  The architecture is a stand-in (the TF-NAS-A config is
not in the reference repository): every stage at full depth, k5-e6-SE everywhere unless --arch gives op indices.

    python tools/bench_eval.py [--bs 128] [--steps 30] [--amp bf16|none] [--arch 7]
"""
import argparse
import json
import os
import sys
from collections import OrderedDict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfnas_b200 import config, eval_loop, model_eval  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--bs', type=int, default=128)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=8)
    ap.add_argument('--amp', default='bf16')
    ap.add_argument('--arch', type=int, default=7, help='candidate index used for every block')
    args = ap.parse_args()
    arch = OrderedDict((s, OrderedDict(('block%d' % (j + 1), args.arch) for j in range(len(shapes))))
                       for s, shapes, _ in model_eval.STAGE_TABLE)
    mc = config.get_mc_num_dddict(config.mc_mask_dddict)
    out = dict(workload='derived network (all blocks, op %d) train step, 3x224x224 bs %d, amp %s, channels_last' %
               (args.arch, args.bs, args.amp))
    x = torch.randn(args.bs, 3, 224, 224, device='cuda')
    t = torch.randint(0, 1000, (args.bs,), device='cuda')
    for fused in (True, False):
        torch.manual_seed(2)
        net = model_eval.Network(1000, arch, mc, None, 0.2, 0.2).cuda().train()
        crit, _ = eval_loop.make_criteria(0.1, fused=fused)
        opt = eval_loop.make_optimizer(net, 0.2, 0.9, 1e-5, fused=fused)
        for _ in range(args.warmup):
            eval_loop.train_step(net, x, t, crit, opt, 5.0, None, args.amp, True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            eval_loop.train_step(net, x, t, crit, opt, 5.0, None, args.amp, True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out['fused_glue' if fused else 'torch_glue'] = dict(ms_per_step=ms, images_per_s=args.bs / ms * 1e3)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
