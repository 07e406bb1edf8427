"""Host-side cost of one MixedOP call: tiny tensors (GPU time negligible), wall time per fwd+bwd."""
import os
import sys
import time
import cProfile
import pstats

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H  # noqa: E402
from tfnas_b200 import _lib  # noqa: E402
from tfnas_b200.config import CAND_SPEC, lut_key  # noqa: E402
from tfnas_b200.model_search import MixedOP, NoisePlan, injected  # noqa: E402

ic, oc, s, act, size, N = 24, 40, 1, 'swish', 8, 2
mcs = H.default_mcs(ic)
P, x, gum, lats = H.make_problem(ic, oc, s, size, N, mcs, seed=1)
lut = {}
for i, (k, _e, sm) in enumerate(CAND_SPEC):
    lut.setdefault(lut_key(size, ic, sm * ic, oc, k, s, act), {})[mcs[i]] = float(lats[i])
op = MixedOP(ic, oc, s, False, act, 8, {i: mcs[i] for i in range(8)}, lut)
op.load_state_dict({k[2:]: v for k, v in P.items()})
op.set_temperature(5.0)
op.cuda()
xg = x.cuda().requires_grad_(True)
G = torch.randn(N, oc, size, size, device='cuda')


def run(alpha, iters):
    for n, p in op.named_parameters():
        p.requires_grad_((n == 'log_alphas') == alpha)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    t0 = time.perf_counter()
    for _ in range(iters):
        xg.grad = None
        if alpha:
            with injected(NoisePlan(noise=[gum])):
                out, lat = op(xg, False, 'max')
            (out * G).sum().add(lat).backward()
        else:
            with injected(NoisePlan(indices=[5])):
                out, _ = op(xg, True, 'random')
            (out * G).sum().backward()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    n = (_lib.launch_count() - l0) / iters
    print('%s: host %.1f us per fwd+bwd (%.0f library launches, %.2f us/launch if all launch cost), drain %.1f us'
          % ('alpha' if alpha else 'single', 1e6 * (t1 - t0) / iters, n, 1e6 * (t1 - t0) / iters / n, 1e6 * (t2 - t1) / iters))


for alpha in (False, True):
    run(alpha, 20)
    run(alpha, 200)
pr = cProfile.Profile()
pr.enable()
run(False, 200)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
