"""Latency look-up tables: the reference's pickle (latency_pkl/latency_{gpu,cpu}.pkl: OrderedDict with 'base' ->
float and 66 block keys -> OrderedDict{mid_channels -> ms}) or the compact .npz form used by the fixtures."""
import pickle
from collections import OrderedDict

import numpy as np


def load_lut(path):
    if str(path).endswith('.npz'):
        z = np.load(path)
        lut = OrderedDict()
        lut['base'] = float(z['base'])
        pos = 0
        for k, n in zip(z['keys'], z['lens']):
            vals = z['vals'][pos:pos + int(n)]
            lut[str(k)] = OrderedDict((m + 1, float(v)) for m, v in enumerate(vals))
            pos += int(n)
        return lut
    with open(path, 'rb') as f:
        return pickle.load(f)
