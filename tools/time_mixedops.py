"""Per-kernel timing of the 18 MixedOP shapes (N=128 by default) using the library's event profiler.

    python tools/time_mixedops.py [--N 128] [--mode alpha|single] [--out gpurun_out/kernels.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H  # noqa: E402
from tfnas_b200 import _lib, config  # noqa: E402
from tfnas_b200.config import CAND_SPEC, lut_key  # noqa: E402
from tfnas_b200.model_search import MixedOP, NoisePlan, injected  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--N', type=int, default=128)
    ap.add_argument('--mode', default='alpha')
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--out', default='gpurun_out/kernels.json')
    ap.add_argument('--only', type=int, default=-1)
    a = ap.parse_args()
    rows = []
    tot = {}
    for bi, (st, bl, ic, oc, s, act, size) in enumerate(config.block_shapes()):
        if a.only >= 0 and bi != a.only:
            continue
        mcs = H.default_mcs(ic)
        P, x, gum, lats = H.make_problem(ic, oc, s, size, a.N, mcs, seed=bi)
        lut = {}
        for i, (k, _e, sm) in enumerate(CAND_SPEC):
            lut.setdefault(lut_key(size, ic, sm * ic, oc, k, s, act), {})[mcs[i]] = float(lats[i])
        op = MixedOP(ic, oc, s, False, act, 8, {i: mcs[i] for i in range(8)}, lut)
        op.load_state_dict({k[2:]: v for k, v in P.items()})
        op.set_temperature(5.0)
        op.cuda()
        alpha = a.mode == 'alpha'
        for n, p in op.named_parameters():
            p.requires_grad_((n == 'log_alphas') == alpha)
        xg = x.cuda().requires_grad_(True)
        G = torch.randn(a.N, oc, (size - 1) // s + 1, (size - 1) // s + 1, device='cuda')

        def step():
            xg.grad = None
            if alpha:
                with injected(NoisePlan(noise=[gum])):
                    out, lat = op(xg, False, 'max')
                (out * G).sum().add(lat).backward()
            else:
                with injected(NoisePlan(indices=[5])):
                    out, _ = op(xg, True, 'random')
                (out * G).sum().backward()
        step()
        torch.cuda.synchronize()
        _lib.prof_enable(True)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(a.reps):
            step()
        e1.record()
        torch.cuda.synchronize()
        recs = _lib.prof_collect()
        _lib.prof_enable(False)
        wall = e0.elapsed_time(e1) / a.reps
        ksum = sum(r['ms'] for r in recs) / a.reps
        print('%s.%s ic%d oc%d s%d %dx%d: step %.3f ms, kernels %.3f ms' % (st, bl, ic, oc, s, size, size, wall, ksum))
        for r in recs:
            ms = r['ms'] / r['launches']
            gbs = r['bytes'] / r['launches'] / ms / 1e6
            tf = r['flops'] / r['launches'] / ms / 1e9
            print('   %-10s %8.3f ms  %8.1f GB/s  %7.2f TF/s' % (r['name'], ms * r['launches'] / a.reps, gbs, tf))
            t = tot.setdefault(r['name'], dict(ms=0.0, bytes=0.0, flops=0.0))
            t['ms'] += r['ms'] / a.reps
            t['bytes'] += r['bytes'] / a.reps
            t['flops'] += r['flops'] / a.reps
        rows.append(dict(block='%s.%s' % (st, bl), step_ms=wall, kernel_ms=ksum, kernels=recs))
        del op, xg, G
        torch.cuda.empty_cache()
    print('=== totals over shapes (%s mode, N=%d) ===' % (a.mode, a.N))
    allms = sum(t['ms'] for t in tot.values())
    for k, t in sorted(tot.items(), key=lambda kv: -kv[1]['ms']):
        print('%-10s %8.3f ms %5.1f%%  %8.1f GB/s  %7.2f TF/s' % (k, t['ms'], 100 * t['ms'] / allms, t['bytes'] / t['ms'] / 1e6, t['flops'] / t['ms'] / 1e9))
    print('total kernel ms %.3f; sum of step wall ms %.3f' % (allms, sum(r['step_ms'] for r in rows)))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(dict(N=a.N, mode=a.mode, rows=rows, totals=tot), open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
