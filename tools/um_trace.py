"""In-kernel phase trace of the tcgen05 project / dx kernels (tfnas_debug_um_trace) for one MixedOP shape.

    python -m tfnas_b200.build --trace
    TFNAS_B200_LIB=tfnas_b200/lib/libtfnas_b200_trace.so python tools/um_trace.py --only 1 [--mode alpha]
Prints, per traced kernel, the mean clock64() cycles thread 0 of a CTA spent in each phase.
"""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H  # noqa: E402
from tfnas_b200 import _lib, config  # noqa: E402
from tfnas_b200.config import CAND_SPEC, lut_key  # noqa: E402
from tfnas_b200.model_search import MixedOP, NoisePlan, injected  # noqa: E402

SLOTS = 16
NAMES = ['total', 'setup', 'wait_mma', 'tma_issue', 'cp_wait', 'emit', 'barrier', 'mma_issue', 'epilogue', 'chunks', 'kid']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--N', type=int, default=128)
    ap.add_argument('--mode', default='alpha')
    ap.add_argument('--only', type=int, default=1)
    ap.add_argument('--ctas', type=int, default=4096)
    a = ap.parse_args()
    lib = _lib.load()
    raw = ctypes.CDLL(_lib.LIB_PATH)
    raw.tfnas_debug_um_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    st, bl, ic, oc, s, act, size = list(config.block_shapes())[a.only]
    mcs = H.default_mcs(ic)
    P, x, gum, lats = H.make_problem(ic, oc, s, size, a.N, mcs, seed=a.only)
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
    buf = torch.zeros(a.ctas * SLOTS, dtype=torch.int64, device='cuda')

    def fwd():
        if alpha:
            with injected(NoisePlan(noise=[gum])):
                out, lat = op(xg, False, 'max')
            return (out * G).sum().add(lat)
        with injected(NoisePlan(indices=[5])):
            out, _ = op(xg, True, 'random')
        return (out * G).sum()

    def report(tag):
        torch.cuda.synchronize()
        t = buf.view(a.ctas, SLOTS).cpu().double()
        t = t[t[:, 0] > 0]
        print('%s.%s ic%d oc%d s%d %dx%d  [%s]  traced CTAs %d' % (st, bl, ic, oc, s, size, size, tag, t.shape[0]))
        if t.shape[0]:
            m = t.mean(0)
            print('   ' + '  '.join('%s=%.0f' % (NAMES[i], m[i]) for i in range(len(NAMES))))
            per = m.clone()
            print('   per chunk: ' + '  '.join('%s=%.0f' % (NAMES[i], m[i] / max(m[9], 1)) for i in (2, 3, 4, 5, 6, 7)))
        buf.zero_()

    loss = fwd()                 # warm-up, untraced
    loss.backward()
    xg.grad = None
    torch.cuda.synchronize()
    raw.tfnas_debug_um_trace(ctypes.c_void_p(buf.data_ptr()), a.ctas)
    loss = fwd()
    report('forward: project')
    loss.backward()
    report('backward: dx')
    raw.tfnas_debug_um_trace(None, 0)


if __name__ == '__main__':
    main()
