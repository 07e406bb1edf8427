"""Launch timeline of one search unit from the library's event profiler: per stream the busy time, the idle gaps between
consecutive launches and the largest kernels; answers "is the step bound by kernel time or by launch gaps".
    python tools/timeline.py [units] > profiles/timeline_rN.txt"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import golden_inputs as gi  # noqa: E402
from tfnas_b200 import _lib, config, model_search  # noqa: E402
from tfnas_b200.model_search import Network  # noqa: E402
from tfnas_b200.parallel import GradSync, SearchParallel  # noqa: E402
from tfnas_b200.search_loop import alpha_step, make_optimizers, w_step  # noqa: E402
from tfnas_b200.step import FusedCrossEntropy  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    torch.manual_seed(2)
    model_search.seed_noise(2)
    net = Network(100, config.get_mc_num_dddict(config.mc_mask_dddict), gi.load_lut())
    net.set_temperature(5.0)
    model = SearchParallel(net).to(dev).train()
    crit = FusedCrossEntropy().to(dev)
    opt_w, opt_a = make_optimizers(net)
    sync = GradSync()
    g = torch.Generator().manual_seed(2)
    pool = [(torch.randn(128, 3, 224, 224, generator=g).to(dev), torch.randint(0, 100, (128,), generator=g).to(dev)) for _ in range(3)]

    def unit(i, mark=None):
        for it in range(2):
            x, t = pool[(2 * i + it) % 3]
            w_step(model, x, t, crit, opt_w, 5.0, sync, bisample=True)
            if mark is not None:
                mark.append(('w%d' % it, len(_lib.prof_timeline())))
            if it % 2 == 0:
                xa, ta = pool[(2 * i + it + 1) % 3]
                alpha_step(model, xa, ta, crit, opt_a, 15.0, 0.1, 5.0, sync)
                if mark is not None:
                    mark.append(('alpha', len(_lib.prof_timeline())))

    for i in range(4):
        unit(i)
    torch.cuda.synchronize()
    _lib.prof_enable(True)
    marks = []
    unit(4, marks)
    torch.cuda.synchronize()
    tl = _lib.prof_timeline()
    _lib.prof_enable(False)
    print('%d library launches in one unit; span %.2f ms (library kernels only: torch ops appear as gaps)' % (len(tl), tl[-1][3] - tl[0][2]))
    lo = 0
    for name, hi in marks:
        seg = tl[lo:hi]
        lo = hi
        if not seg:
            continue
        t0, t1 = min(s[2] for s in seg), max(s[3] for s in seg)
        print('\n== phase %s: %d launches, %.2f ms wall (%.2f .. %.2f)' % (name, len(seg), t1 - t0, t0, t1))
        by = collections.OrderedDict()
        for s in seg:
            by.setdefault(s[1], []).append(s)
        for st, rows in by.items():
            busy = sum(r[3] - r[2] for r in rows)
            gaps = [rows[i][2] - rows[i - 1][3] for i in range(1, len(rows))]
            big = sorted(gaps, reverse=True)[:5]
            print('  stream %x: %4d launches, span %.2f ms, busy %.2f ms, gaps total %.2f ms (median %.1f us, >20us: %d, top %s)' % (
                st & 0xffffff, len(rows), rows[-1][3] - rows[0][2], busy, sum(gaps),
                1e3 * sorted(gaps)[len(gaps) // 2] if gaps else 0, sum(1 for v in gaps if v > 0.02),
                ' '.join('%.0fus' % (1e3 * v) for v in big)))
            agg = collections.Counter()
            for r in rows:
                agg[r[0]] += r[3] - r[2]
            print('     top: ' + ', '.join('%s %.2f' % kv for kv in agg.most_common(12)))


if __name__ == '__main__':
    main()
