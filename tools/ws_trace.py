"""Phase trace of the persistent GEMM kernels (tfnas_debug_ws_trace) for one MixedOP shape.

    python -m tfnas_b200.build --trace
    TFNAS_B200_LIB=tfnas_b200/lib/libtfnas_b200_trace.so python tools/ws_trace.py --only 1 [--kernels project,dx]
For each kernel (run alone through TFNAS_WS) prints the mean clock64() cycles per K chunk that producer warp 0 and the
MMA thread spent in each phase.
"""
import argparse
import ctypes
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SLOTS = 16
PN = ['iters', 'advance+loads', '-', '-', 'wait_empty', 'emit', 'fence', 'arrive']
MN = ['chunks', 'wait_w', 'wait_operands', 'issue+commit', 'wait_acc']


def child(a):
    from tests import helpers as H
    from tfnas_b200 import _lib, config
    from tfnas_b200.config import CAND_SPEC, lut_key
    from tfnas_b200.model_search import MixedOP, NoisePlan, injected
    _lib.load()
    raw = ctypes.CDLL(_lib.LIB_PATH)
    raw.tfnas_debug_ws_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
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
    for n, p in op.named_parameters():
        p.requires_grad_(n == 'log_alphas')
    xg = x.cuda().requires_grad_(True)
    G = torch.randn(a.N, oc, (size - 1) // s + 1, (size - 1) // s + 1, device='cuda')
    buf = torch.zeros(148 * SLOTS, dtype=torch.int64, device='cuda')

    def step():
        with injected(NoisePlan(noise=[gum])):
            out, lat = op(xg, False, 'max')
        (out * G).sum().add(lat).backward()
        xg.grad = None
        torch.cuda.synchronize()

    step()
    raw.tfnas_debug_ws_trace(ctypes.c_void_p(buf.data_ptr()), 148)
    step()
    t = buf.view(148, SLOTS).cpu().double()
    t = t[t[:, 0] > 0]
    m = t.mean(0)
    print('%s.%s ic%d oc%d s%d %dx%d  TFNAS_WS=%s  CTAs %d' % (st, bl, ic, oc, s, size, size, os.environ.get('TFNAS_WS'), t.shape[0]))
    if t.shape[0]:
        it = max(m[0], 1)
        print('   producer: iters=%.0f  per iter: ' % m[0] + '  '.join('%s=%.0f' % (PN[i], m[i] / it) for i in range(1, 8)) +
              '  total=%.0f' % (sum(m[1:8]) / it))
        ch = max(m[8], 1)
        print('   mma:      chunks=%.0f  per chunk: ' % m[8] + '  '.join('%s=%.0f' % (MN[i], m[8 + i] / ch) for i in range(1, 5)) +
              '  total=%.0f' % (sum(m[9:13]) / ch))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--N', type=int, default=128)
    ap.add_argument('--only', type=int, default=1)
    ap.add_argument('--kernels', default='project,dx,dc,expand')
    ap.add_argument('--child', action='store_true')
    a = ap.parse_args()
    if a.child:
        return child(a)
    for k in a.kernels.split(','):
        env = dict(os.environ, TFNAS_WS=k)
        subprocess.run([sys.executable, __file__, '--child', '--N', str(a.N), '--only', str(a.only)], env=env, timeout=300)


if __name__ == '__main__':
    main()
