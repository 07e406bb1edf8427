"""world_size-2 gloo tests of the data-parallel plumbing (tfnas_b200/parallel.py): the flat-bucket
gradient all-reduce averages exactly the live gradients, un-sampled candidates stay grad-None on every
rank, seeded sampling is identical across ranks, and the SearchParallel wrapper keeps the reference's
``module.`` state_dict prefix."""
import os
import random
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import golden_inputs as gi
from tfnas_b200 import config, model_search
from tfnas_b200.model_search import Network
from tfnas_b200.parallel import GradSync, SearchParallel, assert_in_sync


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)        # ranks start from DIFFERENT global RNG states
        random.seed(100 + rank)
        model_search.seed_noise(7)           # ... but share the sampling streams
        mcs = config.get_mc_num_dddict(config.mc_mask_dddict)
        torch.manual_seed(2)
        net = Network(10, mcs, gi.load_lut())
        idx = []
        for m in net.modules():
            if isinstance(m, model_search.MixedOP):
                ig = m._sample_index('gumbel')
                ir = m._sample_index('random')
                idx += [ig, ir]
        ok_sync = assert_in_sync(idx)
        # fake gradients: rank-dependent on the "sampled" tensors, None elsewhere
        params = net.weight_parameters()
        live = [p for i, p in enumerate(params) if i % 3 == 0]
        for j, p in enumerate(live):
            p.grad = torch.full_like(p, float(rank + 1) * (j % 5 + 1))
        nbytes = GradSync()(params)
        ok_avg = all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (j % 5 + 1))) for j, p in enumerate(live))
        ok_none = all(p.grad is None for i, p in enumerate(params) if i % 3 != 0)
        ok_bytes = nbytes == 4 * sum(p.numel() for p in live)
        sd = SearchParallel(net).state_dict()
        ok_prefix = all(k.startswith('module.') for k in sd) and len(sd) == 754
        q.put((rank, ok_sync, ok_avg, ok_none, ok_bytes, ok_prefix, idx[:6]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_grad_sync_and_shared_sampling():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert all(r[1:6]), r
    assert res[0][6] == res[1][6]
