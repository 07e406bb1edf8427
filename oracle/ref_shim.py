"""CPU ORACLE — test infrastructure.  Import shim for the REAL reference.

Only usable where ``/root/reference`` is mounted (the build container).  It is
used to (a) pin ``oracle/port.py`` against the reference's own executable code
and (b) generate the golden vectors committed under ``tests/golden/``.
Nothing on the GPU box may import this (the reference does not travel).
"""
import os
import pickle
import random
import sys

import torch

REF_ROOT = os.environ.get('TFNAS_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'models', 'model_search.py'))


def _import():
    if not available():
        raise RuntimeError('reference not mounted at %s' % REF_ROOT)
    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from models import model_search as ms           # noqa
    from tools import config as cfg                 # noqa
    import parsing_model as pm                      # noqa
    return ms, cfg, pm


def load_lut(name='latency_gpu.pkl'):
    with open(os.path.join(REF_ROOT, 'latency_pkl', name), 'rb') as f:
        return pickle.load(f)


class InjectedNoise(object):
    """Context manager: make F.gumbel_softmax consume a prepared list of Gumbel draws.

    The reference draws ``-empty_like(logits).exponential_().log()`` inside
    ``F.gumbel_softmax`` (models/model_search.py:62,87); seeding the CPU generator
    and drawing ``torch.empty(8).exponential_()`` in forward order reproduces it
    bit-exactly, so here we simply reseed so the reference draws the
    same numbers the plan holds.
    """

    def __init__(self, seed):
        self.seed = seed

    def __enter__(self):
        self.state = torch.random.get_rng_state()
        torch.manual_seed(self.seed)
        return self

    def __exit__(self, *a):
        torch.random.set_rng_state(self.state)


def draw_plan_noise(seed, n_blocks=18, n_ops=8):
    """The same 18x8 draws the reference will make under torch.manual_seed(seed)."""
    st = torch.random.get_rng_state()
    torch.manual_seed(seed)
    noise = [-torch.empty(n_ops).exponential_().log() for _ in range(n_blocks)]
    torch.random.set_rng_state(st)
    return noise


def build_network(mcs, lut, num_classes=100, seed=2, T=5.0):
    ms, _cfg, _pm = _import()
    st = torch.random.get_rng_state()
    torch.manual_seed(seed)
    net = ms.Network(num_classes, mcs, lut)
    torch.random.set_rng_state(st)
    net.set_temperature(T)
    net.train()
    return net


def build_mixedop(ic, oc, stride, act, mc_dict, lut, seed=2, T=5.0):
    ms, _cfg, _pm = _import()
    st = torch.random.get_rng_state()
    torch.manual_seed(seed)
    op = ms.MixedOP(ic, oc, stride, False, act, 8, mc_dict, lut)
    torch.random.set_rng_state(st)
    op.set_temperature(T)
    op.train()
    return op


def seed_python_random(seed):
    random.seed(seed)
