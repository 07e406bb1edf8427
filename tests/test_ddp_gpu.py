"""2-GPU NCCL check of the data-parallel search step (skipped with < 2 GPUs).  SURVEY 8e parity definition: after the
all-reduce every rank holds the arithmetic MEAN over ranks of the ORACLE's gradients, each computed independently (oracle
port, fp32, on the GPU) on one rank's shard with the same weights / sampled indices; ranks sample identical sub-networks;
weights stay bit-identical across ranks after an optimiser step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port_no), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world)
    try:
        _worker_body(rank, world, q)
    except BaseException as e:          # a failing rank must not leave its peer waiting in a collective until the timeout
        import traceback
        q.put((rank, 'error', traceback.format_exc()))
        os._exit(1)
    finally:
        dist.destroy_process_group()


def _worker_body(rank, world, q):
    if True:
        import random
        import torch.nn.functional as F
        from oracle import port
        from tests import golden_inputs as gi
        from tfnas_b200 import config, model_search
        from tfnas_b200.model_search import MixedOP, Network
        from tfnas_b200.parallel import GradSync, SearchParallel
        from tfnas_b200.search_loop import make_optimizers, w_step
        mcs = config.get_mc_num_dddict(config.mc_mask_dddict)
        lut = gi.load_lut()
        P, _x, _t = gi.network_inputs()
        net = Network(100, mcs, lut)
        net.load_state_dict(P)
        net.set_temperature(5.0)
        model = SearchParallel(net).cuda().train()
        crit = nn.CrossEntropyLoss().cuda()
        g = torch.Generator().manual_seed(5)
        BSR, SZ = 16, 96                                   # per-rank shard: 16 images of 96x96 (3x3 planes in the last stage)
        xs = torch.randn(world, BSR, 3, SZ, SZ, generator=g)
        ts = torch.randint(0, 100, (world, BSR), generator=g)

        # --- this rank's shard through the CUDA path (same seeds on every rank => same sampled sub-networks) ------------
        model_search.seed_noise(11)
        sampled = []
        orig = MixedOP._sample_index

        def spy(self, mode):
            i = orig(self, mode)
            sampled.append(i)
            return i
        MixedOP._sample_index = spy
        try:
            for p in net.weight_parameters():
                p.requires_grad = True
            for p in net.arch_parameters():
                p.requires_grad = False
            lg, _ = model(xs[rank].cuda(), sampling=True, mode='gumbel')
            lr, _ = model(xs[rank].cuda(), sampling=True, mode='random')
        finally:
            MixedOP._sample_index = orig
        idx_g, idx_r = sampled[:18], sampled[18:]
        (crit(lg, ts[rank].cuda()) + crit(lr, ts[rank].cuda())).backward()
        from tfnas_b200.parallel import assert_in_sync
        same_idx = assert_in_sync(idx_g + idx_r)
        nbytes = GradSync()(net.weight_parameters())
        got = {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}

        # --- SURVEY 8e parity definition: mean over ranks of ORACLE gradients, each computed on one rank's shard --------
        ref = None
        for s_ in range(world):
            Pg = {k: v.cuda().requires_grad_(not port.is_arch_key(k)) for k, v in P.items()}
            rg, _ = port.network_forward(xs[s_].cuda(), Pg, mcs, lut, True, indices=idx_g)
            rr, _ = port.network_forward(xs[s_].cuda(), Pg, mcs, lut, True, indices=idx_r)
            (F.cross_entropy(rg, ts[s_].cuda()) + F.cross_entropy(rr, ts[s_].cuda())).backward()
            gs = {k: v.grad.double() for k, v in Pg.items() if v.grad is not None}
            ref = gs if ref is None else {k: ref[k] + gs[k] for k in ref}
        ref = {k: v / world for k, v in ref.items()}
        same_keys = set(ref) == set(got)
        gmax = max(float(v.norm()) for v in ref.values())
        worst = max(float((got[k].double() - ref[k]).norm() / max(float(ref[k].norm()), 1e-6 * gmax)) for k in ref)

        # one real optimiser step through the public loop, then compare weights across ranks
        model_search.seed_noise(3)
        opt_w, _ = make_optimizers(net)
        w_step(model, xs[rank].cuda(), ts[rank].cuda(), crit, opt_w, 5.0, GradSync(), bisample=True)
        flat = torch.cat([p.detach().reshape(-1) for p in net.weight_parameters()])
        lo, hi = flat.clone(), flat.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        q.put((rank, worst, same_keys and same_idx, nbytes, bool((lo == hi).all().item())))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_gpu_nccl_grad_mean_and_weight_sync():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = []
    try:
        for _ in procs:
            r = q.get(timeout=300)
            assert r[1] != 'error', r[2]
            res.append(r)
    finally:
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.kill()
    res.sort()
    assert all(p.exitcode == 0 for p in procs)
    for rank, worst, same_keys, nbytes, synced in res:
        print('rank', rank, 'worst element-wise rel-l2 of the all-reduced grads vs the mean of per-shard ORACLE grads', worst,
              'bucket bytes', nbytes)
        assert worst < 1e-3 and same_keys and nbytes > 1e6 and synced
