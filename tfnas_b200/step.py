"""Fused step glue of the search loop over the C ABI (csrc/optim.cu): what the reference does with
``nn.utils.clip_grad_norm_`` + ``torch.optim.SGD`` (train_search.py:196-199, :381-385), ``torch.optim.Adam`` + the
``log_softmax`` renormalisation of every architecture parameter (:200-206, :414-422) and ``nn.CrossEntropyLoss`` (:121).

Per optimiser step the live tensors (parameters whose ``.grad`` exists -- torch.optim skips the others)
go to the library as one pointer table: two launches for clip + SGD over ~200 tensors instead of ~650 ATen launches, one
launch for clip + Adam + renormalisation of the 24 architecture tensors.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .ops import _ptr, _stream


class FusedSGD(object):
    """clip_grad_norm_ + SGD(momentum, weight_decay); same update rule and skip-None semantics as torch.optim.SGD.

    Momentum buffers are zero-initialised views of one flat tensor: ``buf = momentum * 0 + d`` on a tensor's first live step
    is exactly torch's "clone the first d_p"."""

    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0):
        self.params = list(params)
        self.lr, self.momentum, self.weight_decay = float(lr), float(momentum), float(weight_decay)
        self.param_groups = [dict(params=self.params, lr=self.lr)]       # read-only mirror for callers that print the lr
        self._bufs = None
        self._ws = None
        self.last_norm = None

    def _state(self, dev):
        if self._bufs is None:
            flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=dev)
            self._bufs = {id(p): b.view(p.shape) for p, b in zip(self.params, flat.split([p.numel() for p in self.params]))}
            self._ws = torch.zeros(16, dtype=torch.uint8, device=dev)
            self.last_norm = torch.zeros((), dtype=torch.float32, device=dev)
        return self._bufs

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            p.grad = None

    def state_dict(self):
        """Momentum buffers as ONE flat tensor in parameter order (None before the first step) plus the hyper-parameters."""
        flat = None
        if self._bufs is not None:
            flat = torch.cat([self._bufs[id(p)].reshape(-1) for p in self.params]).clone()
        return dict(momentum_flat=flat, lr=self.lr, momentum=self.momentum, weight_decay=self.weight_decay)

    def load_state_dict(self, state):
        self.lr, self.momentum, self.weight_decay = float(state['lr']), float(state['momentum']), float(state['weight_decay'])
        self.param_groups[0]['lr'] = self.lr
        if state.get('momentum_flat') is not None:
            self._state(self.params[0].device)
            torch.cat([self._bufs[id(p)].reshape(-1) for p in self.params])       # shape check
            off = 0
            for p in self.params:
                self._bufs[id(p)].copy_(state['momentum_flat'][off:off + p.numel()].view(p.shape))
                off += p.numel()

    def step(self, max_norm=0.0, grad_scale=1.0):
        live = [p for p in self.params if p.grad is not None]
        if not live:
            return
        dev = live[0].device
        bufs = self._state(dev)
        rows = []
        for p in live:
            g = p.grad
            if not (g.is_contiguous() and p.is_contiguous() and g.dtype == torch.float32 and g.device == dev):
                raise _lib.TfnasError('FusedSGD needs contiguous float32 parameters / gradients on one CUDA device')
            rows.append((p.data_ptr(), g.data_ptr(), bufs[id(p)].data_ptr(), p.numel()))
        tab = np.array(rows, dtype=np.int64)
        _lib.check(_lib.load().tfnas_sgd_step(len(live), tab.ctypes.data_as(ctypes.POINTER(_lib.SgdTensor)), self.lr,
                                              self.momentum, self.weight_decay, float(max_norm), float(grad_scale),
                                              _ptr(self.last_norm), _ptr(self._ws), 16, _stream()))


class FusedArchAdam(object):
    """clip_grad_norm_ + Adam(betas, eps, weight_decay) + ``p = log_softmax(p)`` for every architecture tensor
    (log_alphas AND betas, reference quirk Q4), one launch."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, renorm=True):
        self.params = list(params)
        self.lr, self.betas, self.eps, self.weight_decay, self.renorm = float(lr), betas, float(eps), float(weight_decay), renorm
        self.param_groups = [dict(params=self.params, lr=self.lr)]
        self.t = 0
        self._m = self._v = None

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            p.grad = None

    def step(self, max_norm=0.0, grad_scale=1.0):
        if any(p.grad is None for p in self.params):
            raise _lib.TfnasError('FusedArchAdam: every architecture parameter must have a gradient')
        dev = self.params[0].device
        if self._m is None:
            n = sum(p.numel() for p in self.params)
            self._m = list(torch.zeros(n, dtype=torch.float32, device=dev).split([p.numel() for p in self.params]))
            self._v = list(torch.zeros(n, dtype=torch.float32, device=dev).split([p.numel() for p in self.params]))
        self.t += 1
        tab = (_lib.AdamTensor * len(self.params))()
        for e, p, m, v in zip(tab, self.params, self._m, self._v):
            e.p, e.g, e.m, e.v = p.data_ptr(), p.grad.data_ptr(), m.data_ptr(), v.data_ptr()
            e.numel, e.renorm = p.numel(), 1 if self.renorm else 0
        _lib.check(_lib.load().tfnas_adam_step(len(self.params), tab, self.t, self.lr, self.betas[0], self.betas[1], self.eps,
                                               self.weight_decay, float(max_norm), float(grad_scale), _stream()))
        bump_versions(self.params)      # host mirrors of log_alphas (MixedOP._host_alpha) key on the version counter


def bump_versions(params):
    """The kernels update parameters behind autograd's back; caches keyed on ``tensor._version`` (the host mirrors of
    log_alphas used for Gumbel sampling) must see the change."""
    params = list(params)
    torch._C._autograd._unsafe_set_version_counter(params, [p._version + 1 for p in params])


class _SoftmaxCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, label_smooth=0.0):
        if not (logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == 2 and target.dtype == torch.int64):
            raise _lib.TfnasError('softmax_ce: logits [N, C] float32 CUDA, target int64')
        logits = logits.contiguous()
        target = target.contiguous()
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        dl = torch.empty_like(logits)
        _lib.check(_lib.load().tfnas_softmax_ce_smooth(logits.shape[0], logits.shape[1], _ptr(logits), _ptr(target),
                                                       float(label_smooth), _ptr(loss), _ptr(dl), _stream()))
        ctx.save_for_backward(dl)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g, None, None


def softmax_ce(logits, target, label_smooth=0.0):
    """nn.CrossEntropyLoss()(logits, target) (mean reduction) with its gradient produced in the same launch;
    ``label_smooth`` > 0 gives the derived-network criterion CrossEntropyLabelSmooth (train_eval.py:72-84)."""
    return _SoftmaxCE.apply(logits, target, label_smooth)


class FusedCrossEntropy(torch.nn.Module):
    """Drop-in for ``nn.CrossEntropyLoss()`` (train_search.py:121) and, with ``label_smooth``, for
    ``CrossEntropyLabelSmooth(num_classes, epsilon)`` (train_eval.py:72-84), on the library kernel."""

    def __init__(self, label_smooth=0.0):
        super().__init__()
        self.label_smooth = float(label_smooth)

    def forward(self, logits, target):
        return softmax_ce(logits.float(), target, self.label_smooth)
