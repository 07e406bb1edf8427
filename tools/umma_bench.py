"""Cycles per 12-MMA round (one K=32 chunk of the tf32x3 split) for MN-major vs K-major A operands."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfnas_b200 import _lib  # noqa: E402

raw = ctypes.CDLL(_lib.LIB_PATH)
raw.tfnas_umma_bench.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
for ctas in (148, 296):
    for N in (16, 32, 48, 112, 192, 256):
        row = []
        for mode in (0, 1):
            out = torch.zeros(ctas, dtype=torch.int64, device='cuda')
            rc = raw.tfnas_umma_bench(mode, N, 200, ctas, ctypes.c_void_p(out.data_ptr()), None)
            torch.cuda.synchronize()
            assert rc == 0
            row.append(float(out.double().mean()))
        print('ctas %3d  N %3d   MN-major A: %7.0f cyc/round   K-major A: %7.0f cyc/round' % (ctas, N, row[0], row[1]))
