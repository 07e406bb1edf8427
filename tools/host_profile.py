"""cProfile of the HOST side of the bench's search unit (2 bi-sampled w-steps + 1 alpha-step) on cuda:0: where the
enqueue time goes (Python / torch dispatch / the library's ctypes calls).  Time inside tfnas_mixedop_fwd / _bwd includes
waiting for a full launch queue, i.e. the part of the step that is GPU-bound.
    python tools/host_profile.py [units] > profiles/host_profile_rN.txt"""
import cProfile
import io
import os
import pstats
import sys
import time

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import golden_inputs as gi  # noqa: E402
from tfnas_b200 import _lib, config, model_search  # noqa: E402
from tfnas_b200.model_search import Network  # noqa: E402
from tfnas_b200.parallel import GradSync, SearchParallel  # noqa: E402
from tfnas_b200.search_loop import alpha_step, make_optimizers, w_step  # noqa: E402


def main():
    units = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    _lib.load()
    dev = torch.device('cuda', 0)
    torch.manual_seed(2)
    model_search.seed_noise(2)
    net = Network(100, config.get_mc_num_dddict(config.mc_mask_dddict), gi.load_lut())
    net.set_temperature(5.0)
    model = SearchParallel(net).to(dev).train()
    from tfnas_b200.step import FusedCrossEntropy
    crit = FusedCrossEntropy().to(dev)
    opt_w, opt_a = make_optimizers(net)
    sync = GradSync()
    g = torch.Generator().manual_seed(2)
    pool = [(torch.randn(128, 3, 224, 224, generator=g).to(dev), torch.randint(0, 100, (128,), generator=g).to(dev)) for _ in range(3)]

    def unit(i):
        for it in range(2):
            x, t = pool[(2 * i + it) % 3]
            w_step(model, x, t, crit, opt_w, 5.0, sync, bisample=True)
            if it % 2 == 0:
                xa, ta = pool[(2 * i + it + 1) % 3]
                alpha_step(model, xa, ta, crit, opt_a, 15.0, 0.1, 5.0, sync)

    for i in range(3):
        unit(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(units):
        unit(i)
    host = (time.perf_counter() - t0) / units
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / units
    print('unprofiled: host enqueue %.1f ms / unit, wall %.1f ms / unit' % (host * 1e3, wall * 1e3))
    pr = cProfile.Profile()
    pr.enable()
    for i in range(units):
        unit(i)
    pr.disable()
    torch.cuda.synchronize()
    for key in ('tottime', 'cumulative'):
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats(key).print_stats(28)
        print('==== sorted by %s (%d units) ====' % (key, units))
        print('\n'.join(l for l in s.getvalue().splitlines() if l.strip())[:6000])


if __name__ == '__main__':
    main()
