"""2-GPU NCCL check of the data-parallel search step (skipped with < 2 GPUs): after GradSync every rank holds
the MEAN of the per-shard gradients (SURVEY 8e parity definition), ranks sample identical sub-networks, and
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


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world)
    try:
        from tests import golden_inputs as gi
        from tfnas_b200 import config, model_search
        from tfnas_b200.model_search import Network
        from tfnas_b200.parallel import GradSync, SearchParallel
        from tfnas_b200.search_loop import make_optimizers, w_step
        mcs = config.get_mc_num_dddict(config.mc_mask_dddict)
        torch.manual_seed(2)
        net = Network(100, mcs, gi.load_lut())
        net.set_temperature(5.0)
        model = SearchParallel(net).cuda().train()
        crit = nn.CrossEntropyLoss().cuda()
        g = torch.Generator().manual_seed(5)
        xs = torch.randn(world, 4, 3, 64, 64, generator=g)
        ts = torch.randint(0, 100, (world, 4), generator=g)

        def grads_for(shard):
            model_search.seed_noise(11)
            for p in net.parameters():
                p.grad = None
            for p in net.weight_parameters():
                p.requires_grad = True
            for p in net.arch_parameters():
                p.requires_grad = False
            lg, _ = model(xs[shard].cuda(), sampling=True, mode='gumbel')
            lr, _ = model(xs[shard].cuda(), sampling=True, mode='random')
            (crit(lg, ts[shard].cuda()) + crit(lr, ts[shard].cuda())).backward()
            return {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}

        per = [grads_for(s) for s in range(world)]                 # every rank computes all shards (reference for the mean)
        # The collective is checked on the SAME gradients that enter the reference mean: the kernels accumulate their
        # batch statistics with atomics, and at this tiny batch (4 images, 2x2 planes in the last stages) the BN
        # backward is ill-conditioned enough that a second evaluation differs by ~1e-4 (checked separately below).
        mine = {n: g.clone() for n, g in per[rank].items()}
        again = grads_for(rank)
        rerun = max(float((again[n] - mine[n]).norm() / (mine[n].norm() + 1e-20)) for n in mine)
        for n, p in net.named_parameters():
            p.grad = mine.get(n)
        # reference for the collective: gather what every rank computed for ITS shard and average
        names = sorted(mine)
        flat_mine = torch.cat([mine[n].reshape(-1) for n in names])
        gathered = [torch.empty_like(flat_mine) for _ in range(world)]
        dist.all_gather(gathered, flat_mine)
        ref_flat = sum(gathered) / world
        nbytes = GradSync()(net.weight_parameters())
        npar = dict(net.named_parameters())
        got_flat = torch.cat([npar[n].grad.reshape(-1) for n in names])
        worst = float((got_flat - ref_flat).norm() / (ref_flat.norm() + 1e-20))
        # cross-rank reproducibility: the other shards recomputed HERE vs what their owners computed (atomics-order noise)
        cross = 0.0
        off = 0
        for n in names:
            k = mine[n].numel()
            for s_ in range(world):
                other = gathered[s_][off:off + k].view_as(mine[n])
                cross = max(cross, float((per[s_][n] - other).norm() / (other.norm() + 1e-20)))
            off += k
        rerun = max(rerun, cross)
        same_keys = all(set(per[0]) == set(per[s]) for s in range(world))
        # one real optimiser step through the public loop, then compare weights across ranks
        model_search.seed_noise(3)
        opt_w, _ = make_optimizers(net)
        w_step(model, xs[rank].cuda(), ts[rank].cuda(), crit, opt_w, 5.0, GradSync(), bisample=True)
        flat = torch.cat([p.detach().reshape(-1) for p in net.weight_parameters()])
        lo, hi = flat.clone(), flat.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        q.put((rank, worst, same_keys, nbytes, bool((lo == hi).all().item()), rerun))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_gpu_nccl_grad_mean_and_weight_sync():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, worst, same_keys, nbytes, synced, rerun in res:
        print('rank', rank, 'worst rel err of all-reduced grads vs mean of shard grads', worst, 'bucket bytes', nbytes,
              'run-to-run', rerun)
        assert worst < 1e-5 and same_keys and nbytes > 1e6 and synced
        assert rerun < 2e-3        # atomics-order noise of one rank's own gradients at bs 4 (north-star tolerance 1e-3 is at bs 128)
