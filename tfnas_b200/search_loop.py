"""The search inner loops: ``train_wo_arch`` / ``train_w_arch`` / ``validate`` with the reference's
names, argument meaning and update rules (train_search.py:318-462), over the B200 MixedOP kernels.

Deliberate differences: meters are accumulated on the device and read back only at
``print_freq`` (the reference calls ``.item()`` three times per step); gradients are averaged over
ranks by ``parallel.GradSync`` when torch.distributed is initialised.
"""
import logging

import torch
import torch.nn as nn
import torch.nn.functional as F

from .parallel import GradSync, world_size
from .step import FusedArchAdam, FusedCrossEntropy, FusedSGD


class DeviceMeter(object):
    """AverageMeter (tools/utils.py:37-58) whose sum lives on the device."""

    def __init__(self):
        self.sum = None
        self.cnt = 0

    def update(self, val, n=1):
        v = val.detach().float() * n
        self.sum = v if self.sum is None else self.sum + v
        self.cnt += n

    @property
    def avg(self):
        """Mean over everything seen so far, over ALL ranks when torch.distributed is initialised (every rank reads its
        meters at the same steps, so the all-reduce is collective-safe)."""
        if self.sum is None:
            return 0.0
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            t = torch.stack([self.sum.reshape(()).double(), torch.tensor(float(self.cnt), dtype=torch.float64,
                                                                         device=self.sum.device)])
            dist.all_reduce(t)
            return float(t[0]) / max(float(t[1]), 1.0)
        return float(self.sum) / max(self.cnt, 1)


class DevicePrefetcher(object):
    """Iterate a queue of (pinned) host batches with the host->device copy of batch i+1 running on a side stream while step i
    computes (the reference copies on the compute stream right before use, train_search.py:367-368).  Yields device
    tensors that are safe to use on the current stream."""

    _STREAMS = {}

    def __init__(self, queue, device=None):
        self.queue = queue
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)

    def __len__(self):
        return len(self.queue)

    def _load(self, batch):
        key = self.device.index
        if key not in DevicePrefetcher._STREAMS:
            DevicePrefetcher._STREAMS[key] = torch.cuda.Stream(device=self.device)
        cs = DevicePrefetcher._STREAMS[key]
        with torch.cuda.stream(cs):
            dev = tuple(t.to(self.device, non_blocking=True) for t in batch)
            ev = cs.record_event()
        return dev, ev

    def __iter__(self):
        it = iter(self.queue)
        try:
            nxt = self._load(next(it))
        except StopIteration:
            return
        while nxt is not None:
            batch, ev = nxt
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for t in batch:
                t.record_stream(cur)
            try:
                nxt = self._load(next(it))
            except StopIteration:
                nxt = None
            yield batch


def accuracy(output, target, topk=(1,)):
    """tools/utils.py:61-74, returning device tensors (no host sync)."""
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.view(1, -1).expand_as(pred.t()))
    return [correct[:k].reshape(-1).float().sum(0).mul_(100.0 / target.size(0)) for k in topk]


def _set_requires_grad(net, weights, arch):
    for p in net.weight_parameters():
        p.requires_grad = weights
    for p in net.arch_parameters():
        p.requires_grad = arch


_SIDE_STREAMS = {}
# the two passes of a bi-sampled step accumulate into the shared stem / head parameters from two streams on purpose
if hasattr(torch.autograd.graph, 'set_warn_on_accumulate_grad_stream_mismatch'):
    torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)


def _side_streams(device):
    """Two pass streams per device for the bi-sampled w-step (created once).  They are HIGH-priority streams: the library's
    weight-gradient side streams (csrc/bwd.cu) sit at the default, lowest priority, so when SM slots free up the pending
    CTAs of a pass's dx critical path are scheduled before weight-gradient tiles."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = (torch.cuda.Stream(device=key, priority=-1), torch.cuda.Stream(device=key, priority=-1))
    return _SIDE_STREAMS[key]


def w_step(model, x_w, target_w, criterion, optimizer_w, grad_clip, sync=None, bisample=True, overlap=True):
    """One weight step (train_search.py:370-385): loss = CE(gumbel path) [+ CE(random other path)].

    The two sampled sub-networks of a bi-sampled step are independent until their losses are added, and a single
    sampled path leaves most SMs idle in the late stages (49 pixel tiles at 7x7), so with ``overlap`` the two passes are
    enqueued on two CUDA streams (forward and, through autograd's stream tracking, backward).  The host-side order of
    the sampling decisions (gumbel path first, then a random OTHER path) is unchanged."""
    net = model.module
    _set_requires_grad(net, True, False)
    if bisample and overlap and x_w.is_cuda:
        cur = torch.cuda.current_stream(x_w.device)
        s1, s2 = _side_streams(x_w.device)
        # the stems see the same batch and the same weights in both sampled sub-networks: evaluate them once (and, in
        # backward, once over the sum of the two gradients) -- same result as two full forwards
        shared = hasattr(net, 'forward_stems')
        x0 = net.forward_stems(x_w) if shared else None
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            logits_g, _ = net.forward_from_stem(x0, True, 'gumbel') if shared else model(x_w, sampling=True, mode='gumbel')
            loss_g = criterion(logits_g, target_w)
        with torch.cuda.stream(s2):
            logits_r, _ = net.forward_from_stem(x0, True, 'random') if shared else model(x_w, sampling=True, mode='random')
            loss_r = criterion(logits_r, target_w)
        cur.wait_stream(s1)
        cur.wait_stream(s2)
        for t in (x_w, target_w) + ((x0,) if shared else ()):
            t.record_stream(s1)
            t.record_stream(s2)
        for t in (logits_g, loss_g):
            t.record_stream(cur)
        loss_r.record_stream(cur)
        loss = loss_g + loss_r
    else:
        logits_g, _ = model(x_w, sampling=True, mode='gumbel')
        loss = criterion(logits_g, target_w)
        if bisample:
            logits_r, _ = model(x_w, sampling=True, mode='random')
            loss = loss + criterion(logits_r, target_w)
        else:
            net.reset_switches()
    optimizer_w.zero_grad()
    loss.backward()
    _apply_update(optimizer_w, net.weight_parameters(), grad_clip, sync)
    return loss, logits_g


def _apply_update(optimizer, params, grad_clip, sync):
    """all-reduce (data parallel) -> global-norm clip -> optimiser step.  With the fused optimisers the three are two
    library launches over a table of the live tensors (the 1/world of the gradient mean rides in the same kernels)."""
    if isinstance(optimizer, (FusedSGD, FusedArchAdam)):
        scale = 1.0
        if sync is not None and world_size() > 1:
            sync(params, average=False)
            scale = 1.0 / world_size()
        optimizer.step(max_norm=grad_clip, grad_scale=scale)
        return
    if sync is not None:
        sync(params)
    if grad_clip > 0:
        nn.utils.clip_grad_norm_(params, grad_clip)
    optimizer.step()


GRAPH_LAUNCHES = [0]      # library kernels launched through CUDA-graph replays (tfnas_launch_count only sees direct launches)


class GraphedAlphaStep(object):
    """The alpha step's forward + backward (stems -> 18 MixedOPs + 6 sinks -> head -> loss -> backward) captured ONCE as a
    CUDA graph and replayed: its launch sequence is fixed (all candidates active, static shapes, every kernel in the library,
    no host synchronisation), only the batch and the Gumbel noise change -- they are copied into static buffers before each
    replay.  The optimiser update stays outside the graph (Adam's bias correction changes every step).  Re-captured when the
    temperature, the batch shape or the loss constants change (once per epoch in train_search.py)."""

    def __init__(self, model, criterion, x_a, target_a, target_lat, lambda_lat):
        from .model_search import NoisePlan, injected
        net = model.module
        dev = x_a.device
        self.key = (tuple(x_a.shape), float(getattr(net._param_lists()[4][0], 'T', 1.0)), float(target_lat), float(lambda_lat))
        self.x = torch.empty_like(x_a)
        self.t = torch.empty_like(target_a)
        self.noise_host = torch.empty((len(net._param_lists()[4]), 8), dtype=torch.float32).pin_memory()
        self.noise = torch.zeros_like(self.noise_host, device=dev)
        self.x.copy_(x_a)
        self.t.copy_(target_a)

        def run():
            with injected(NoisePlan(device_noise=self.noise)):
                logits, lat = model(self.x, sampling=False)
            loss_a = criterion(logits, self.t)
            loss_l = torch.abs(lat / target_lat - 1.) * lambda_lat
            (loss_a + loss_l).backward()
            return loss_a, loss_l
        # warm-up on a side stream (allocator, lazily initialised kernels), as torch.cuda.graphs asks, then capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                for p in net.arch_parameters():
                    p.grad = None
                run()
        torch.cuda.current_stream(dev).wait_stream(side)
        for p in net.arch_parameters():
            p.grad = None
        from . import _lib
        self.graph = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.loss_a, self.loss_l = run()
        self.launches = _lib.launch_count() - l0          # kernel nodes of the library in the graph

    def replay(self, net, x_a, target_a):
        self.x.copy_(x_a, non_blocking=True)
        self.t.copy_(target_a, non_blocking=True)
        self.noise_host.copy_(net.draw_alpha_noise())
        self.noise.copy_(self.noise_host, non_blocking=True)
        self.graph.replay()
        GRAPH_LAUNCHES[0] += self.launches
        return self.loss_a, self.loss_l


def alpha_step(model, x_a, target_a, criterion, optimizer_a, target_lat, lambda_lat, grad_clip, sync=None, graph=None):
    """One architecture step (train_search.py:404-422) incl. the log_softmax renormalisation.  ``graph`` (default: on for
    the library-only network with the fused Adam, TFNAS_GRAPH_ALPHA=0 turns it off) replays the step's forward + backward
    as a CUDA graph (GraphedAlphaStep)."""
    net = model.module
    _set_requires_grad(net, False, True)
    if graph is None:
        import os
        graph = (os.environ.get('TFNAS_GRAPH_ALPHA', '1') != '0' and isinstance(optimizer_a, FusedArchAdam)
                 and getattr(net, 'use_body', False) and x_a.is_cuda)
    if graph:
        from . import model_search as _ms
        if _ms._ACTIVE_PLAN[0] is not None:
            graph = False            # injected noise / indices (tests): the eager path consumes the plan
    if graph:
        key = (tuple(x_a.shape), float(getattr(net._param_lists()[4][0], 'T', 1.0)), float(target_lat), float(lambda_lat))
        g = net.__dict__.get('_alpha_graph')
        if g is None or g.key != key or g.criterion is not criterion:
            g = GraphedAlphaStep(model, criterion, x_a, target_a, target_lat, lambda_lat)
            g.criterion = criterion
            net.__dict__['_alpha_graph'] = g
        loss_a, loss_l = g.replay(net, x_a, target_a)
        _apply_update(optimizer_a, net.arch_parameters(), grad_clip, sync)
        return loss_a, loss_l
    logits_a, lat = model(x_a, sampling=False)
    loss_a = criterion(logits_a, target_a)
    loss_l = torch.abs(lat / target_lat - 1.) * lambda_lat
    loss = loss_a + loss_l
    optimizer_a.zero_grad()
    loss.backward()
    _apply_update(optimizer_a, net.arch_parameters(), grad_clip, sync)
    if not isinstance(optimizer_a, FusedArchAdam):      # the fused kernel renormalises in the same launch
        for p in net.arch_parameters():      # applies to betas too (quirk Q4)
            p.data = F.log_softmax(p.detach().data, dim=-1)
    return loss_a, loss_l


def train_wo_arch(train_queue, model, criterion, optimizer_w, args, sync=None):
    objs_w, top1, top5 = DeviceMeter(), DeviceMeter(), DeviceMeter()
    model.train()
    sync = sync if sync is not None else GradSync()
    for step, (x_w, target_w) in enumerate(DevicePrefetcher(train_queue)):
        loss, logits = w_step(model, x_w, target_w, criterion, optimizer_w, args.grad_clip, sync, bisample=False)
        prec1, prec5 = accuracy(logits, target_w, topk=(1, 5))
        n = x_w.size(0)
        objs_w.update(loss, n)
        top1.update(prec1, n)
        top5.update(prec5, n)
        if step % args.print_freq == 0:
            logging.info('TRAIN wo_Arch Step: %04d Objs: %f R1: %f R5: %f', step, objs_w.avg, top1.avg, top5.avg)
    return top1.avg


def train_w_arch(train_queue, val_queue, model, criterion, optimizer_w, optimizer_a, args, sync=None):
    objs_a, objs_l, objs_w = DeviceMeter(), DeviceMeter(), DeviceMeter()
    top1, top5 = DeviceMeter(), DeviceMeter()
    model.train()
    sync = sync if sync is not None else GradSync()
    val_queue_iter = None
    for step, (x_w, target_w) in enumerate(DevicePrefetcher(train_queue)):
        loss_w, logits = w_step(model, x_w, target_w, criterion, optimizer_w, args.grad_clip, sync, bisample=True)
        prec1, prec5 = accuracy(logits, target_w, topk=(1, 5))
        n = x_w.size(0)
        objs_w.update(loss_w, n)
        top1.update(prec1, n)
        top5.update(prec5, n)
        if step % 2 == 0:
            try:
                x_a, target_a = next(val_queue_iter)
            except (StopIteration, TypeError):
                val_queue_iter = iter(DevicePrefetcher(val_queue))
                x_a, target_a = next(val_queue_iter)
            loss_a, loss_l = alpha_step(model, x_a, target_a, criterion, optimizer_a, args.target_lat,
                                        args.lambda_lat, args.grad_clip, sync)
            n = x_a.size(0)
            objs_a.update(loss_a, n)
            objs_l.update(loss_l, n)
        if step % args.print_freq == 0:
            logging.info('TRAIN w_Arch Step: %04d Objs_W: %f R1: %f R5: %f Objs_A: %f Objs_L: %f',
                         step, objs_w.avg, top1.avg, top5.avg, objs_a.avg, objs_l.avg)
    return top1.avg


def validate(val_queue, model, criterion, args):
    objs, top1, top5 = DeviceMeter(), DeviceMeter(), DeviceMeter()
    model.train()    # batch statistics on purpose: no running stats exist (quirk Q7)
    for step, (x, target) in enumerate(DevicePrefetcher(val_queue)):
        with torch.no_grad():
            logits, _ = model(x, sampling=True, mode='gumbel')
            loss = criterion(logits, target)
        model.module.reset_switches()
        prec1, prec5 = accuracy(logits, target, topk=(1, 5))
        n = x.size(0)
        objs.update(loss, n)
        top1.update(prec1, n)
        top5.update(prec5, n)
        if step % args.print_freq == 0:
            logging.info('VALIDATE Step: %04d Objs: %f R1: %f R5: %f', step, objs.avg, top1.avg, top5.avg)
    return top1.avg


def make_optimizers(net, w_lr=0.025, w_mom=0.9, w_wd=1e-5, a_lr=0.01, a_beta1=0.5, a_beta2=0.999, a_wd=5e-4, fused=None):
    """train_search.py:196-206.  ``fused`` (default: on for CUDA parameters, TFNAS_FUSED_OPT=0 turns it off) selects the
    library's fused clip + update kernels (tfnas_b200/step.py) instead of torch.optim; same update rules."""
    if fused is None:
        import os
        fused = os.environ.get('TFNAS_FUSED_OPT', '1') != '0' and next(net.parameters()).is_cuda
    if fused:
        return (FusedSGD(net.weight_parameters(), lr=w_lr, momentum=w_mom, weight_decay=w_wd),
                FusedArchAdam(net.arch_parameters(), lr=a_lr, betas=(a_beta1, a_beta2), weight_decay=a_wd))
    optimizer_w = torch.optim.SGD(net.weight_parameters(), lr=w_lr, momentum=w_mom, weight_decay=w_wd)
    optimizer_a = torch.optim.Adam(net.arch_parameters(), lr=a_lr, betas=(a_beta1, a_beta2), weight_decay=a_wd)
    return optimizer_w, optimizer_a
