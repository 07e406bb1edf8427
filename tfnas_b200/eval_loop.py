"""Re-training loop of a derived network: what the reference's ``train_eval.py:243-300`` (fp32,
DataParallel) and ``train_eval_amp.py`` (apex AMP + DDP) do per batch and per epoch, as one process per GPU.

forward (optionally under ``torch.autocast(bfloat16)`` with channels-last activations — the apex ``--opt_level`` of the
reference has no counterpart in this image) -> label-smoothing cross-entropy and its gradient in ONE library launch
(csrc/optim.cu::k_softmax_ce) -> backward -> gradient all-reduce (parallel.GradSync, NCCL) -> global-norm clip + momentum
SGD + weight decay in two library launches over a pointer table (step.FusedSGD).  Meters stay on the device; the host
reads them every ``print_freq`` steps only."""
import contextlib
import logging
import time

import torch

from .search_loop import DeviceMeter, DevicePrefetcher, _apply_update, accuracy
from .step import FusedCrossEntropy, FusedSGD


def autocast(amp):
    """``amp``: None / 'none' = fp32 (train_eval.py), 'bf16' = mixed precision (train_eval_amp.py's role)."""
    if amp in (None, '', 'none', 'fp32'):
        return contextlib.nullcontext()
    if amp != 'bf16':
        raise ValueError('invalid amp mode: %s' % amp)
    return torch.autocast('cuda', dtype=torch.bfloat16)


def make_optimizer(model, lr=0.2, momentum=0.9, weight_decay=1e-5, fused=True):
    """torch.optim.SGD(model.parameters(), lr, momentum, weight_decay) (train_eval.py:129-131); fused by default."""
    params = list(model.parameters())
    if fused:
        return FusedSGD(params, lr, momentum=momentum, weight_decay=weight_decay)
    return torch.optim.SGD(params, lr, momentum=momentum, weight_decay=weight_decay)


def make_criteria(label_smooth, fused=True):
    """(training criterion with label smoothing, plain validation criterion) — train_eval.py:123-127."""
    if fused:
        return FusedCrossEntropy(label_smooth), FusedCrossEntropy()
    return torch.nn.CrossEntropyLoss(label_smoothing=label_smooth), torch.nn.CrossEntropyLoss()


def set_lr(optimizer, lr):
    if isinstance(optimizer, FusedSGD):
        optimizer.lr = float(lr)
    for g in optimizer.param_groups:
        g['lr'] = float(lr)


def epoch_lr(lr_list, epoch, batch_size):
    """The lr an epoch trains with: the cosine value, linearly warmed up over the first five epochs when the global batch
    exceeds 256 (train_eval.py:201-208)."""
    lr = lr_list[epoch]
    return lr * (epoch + 1) / 5.0 if (epoch < 5 and batch_size > 256) else lr


def train_step(model, x, target, criterion, optimizer, grad_clip=5.0, sync=None, amp=None, channels_last=False):
    """One optimiser step (train_eval.py:259-270).  Returns (loss, logits) as device tensors."""
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    with autocast(amp):
        logits = model(x)
    loss = criterion(logits.float(), target)
    optimizer.zero_grad()
    loss.backward()
    _apply_update(optimizer, list(model.parameters()), grad_clip, sync)
    return loss.detach(), logits.detach()


def train(train_queue, model, criterion, optimizer, args, sync=None):
    """One epoch (train_eval.py:243-283): returns (top-1 average, loss average) over all ranks."""
    objs, top1, top5 = DeviceMeter(), DeviceMeter(), DeviceMeter()
    model.train()
    t_print = time.time()
    for step, (x, target) in enumerate(DevicePrefetcher(train_queue)):
        loss, logits = train_step(model, x, target, criterion, optimizer, args.grad_clip, sync,
                                  getattr(args, 'amp', None), getattr(args, 'channels_last', False))
        n = x.size(0)
        prec1, prec5 = accuracy(logits.float(), target, topk=(1, 5))
        objs.update(loss, n)
        top1.update(prec1, n)
        top5.update(prec5, n)
        if step % args.print_freq == 0:
            now = time.time()
            logging.info('TRAIN Step: %03d Objs: %e R1: %f R5: %f Duration: %ds', step, objs.avg, top1.avg, top5.avg,
                         0 if step == 0 else now - t_print)
            t_print = now
    return top1.avg, objs.avg


def validate(val_queue, model, criterion, args):
    """train_eval.py:286-312: running statistics, no dropout / drop-connect, plain cross-entropy."""
    objs, top1, top5 = DeviceMeter(), DeviceMeter(), DeviceMeter()
    model.eval()
    t_print = time.time()
    with torch.no_grad():
        for step, (x, target) in enumerate(DevicePrefetcher(val_queue)):
            if getattr(args, 'channels_last', False):
                x = x.contiguous(memory_format=torch.channels_last)
            with autocast(getattr(args, 'amp', None)):
                logits = model(x)
            loss = criterion(logits.float(), target)
            n = x.size(0)
            prec1, prec5 = accuracy(logits.float(), target, topk=(1, 5))
            objs.update(loss, n)
            top1.update(prec1, n)
            top5.update(prec5, n)
            if step % args.print_freq == 0:
                now = time.time()
                logging.info('VALID Step: %03d Objs: %e R1: %f R5: %f Duration: %ds', step, objs.avg, top1.avg, top5.avg,
                             0 if step == 0 else now - t_print)
                t_print = now
    return top1.avg, top5.avg, objs.avg


def broadcast_model(model, src=0):
    """All ranks start from rank ``src``'s parameters and buffers (what wrapping in DDP does at construction)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src)
