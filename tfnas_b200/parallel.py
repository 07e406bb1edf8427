"""Data-parallel plumbing for the search loop: one process per GPU, torch.distributed (NCCL on
GPUs, gloo in the CPU tests), one flat gradient bucket per optimiser step.

Semantics (the reference's own nn.DataParallel search cannot run on >1 GPU):
  * every rank draws the SAME Gumbel noise and the SAME sampled candidate indices each step
    (shared seeds, see ``model_search.seed_noise``), so all ranks train the same sub-network;
  * BatchNorm statistics are per-rank (local batch), exactly as the per-GPU batch of the reference;
  * after backward the live gradients are averaged over ranks in ONE all-reduce of a flat fp32
    bucket (w-step: only the sampled candidates' tensors + stems/head, about 35 MB; alpha-step:
    162 floats), then clipping and the optimiser step run identically on every rank.
"""
import os

import torch
import torch.distributed as dist


class SearchParallel(torch.nn.Module):
    """Stand-in for the reference's ``nn.DataParallel(model)`` wrapper (train_search.py:95,158):
    exposes ``.module`` and prefixes state_dict keys with ``module.``; one device per process."""

    def __init__(self, module):
        super(SearchParallel, self).__init__()
        self.module = module

    def forward(self, *a, **kw):
        return self.module(*a, **kw)


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


class GradSync(object):
    """Sum (or average) the live gradients of ``params`` over ranks.

    Gradients that are adjacent views of one storage -- the body executor writes all weight gradients of a sampled pass
    into ONE flat buffer (ops.BodyFn.backward) -- are all-reduced in place as one range: no ``cat`` and no copy back.  What is
    left (the dozen stem / head tensors) goes through one small packed bucket."""

    SMALL = 1 << 18      # ranges below 1 MB are packed together

    def __init__(self):
        self.bytes_last = 0
        self.calls_last = 0

    def __call__(self, params, average=True):
        if world_size() == 1:
            return 0
        grads = [p.grad for p in params if p.grad is not None]
        if not grads:
            return 0
        w = world_size()
        # coalesce adjacent views of the same storage
        by_store = {}
        for g in grads:
            if not g.is_contiguous():
                raise RuntimeError('GradSync needs contiguous gradients')
            by_store.setdefault(g.untyped_storage().data_ptr(), []).append(g)
        big, small = [], []
        for gs in by_store.values():
            gs.sort(key=lambda t: t.storage_offset())
            start, end, first = gs[0].storage_offset(), gs[0].storage_offset() + gs[0].numel(), gs[0]
            members = [gs[0]]
            runs = []
            for g in gs[1:]:
                if g.storage_offset() == end:
                    end += g.numel()
                    members.append(g)
                else:
                    runs.append((first, start, end, members))
                    start, end, first, members = g.storage_offset(), g.storage_offset() + g.numel(), g, [g]
            runs.append((first, start, end, members))
            for first, start, end, members in runs:
                if end - start >= self.SMALL:
                    big.append(first.new_empty(0).set_(first.untyped_storage(), start, (end - start,), (1,)))
                else:
                    small += members
        nbytes = 0
        calls = 0
        for flat in big:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            if average:
                flat.div_(w)
            nbytes += flat.numel() * 4
            calls += 1
        if small:
            flat = torch.cat([g.reshape(-1) for g in small])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            if average:
                flat.div_(w)
            torch._foreach_copy_(small, [o.view_as(g) for o, g in zip(flat.split([g.numel() for g in small]), small)])
            nbytes += flat.numel() * 4
            calls += 1
        self.bytes_last, self.calls_last = nbytes, calls
        return nbytes


def assert_in_sync(values):
    """Debug/test helper: all ranks must hold identical integer lists (sampled indices)."""
    if world_size() == 1:
        return True
    t = torch.tensor(values, dtype=torch.int64)
    if dist.get_backend() == 'nccl':
        t = t.cuda()
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool((lo == hi).all().item())
