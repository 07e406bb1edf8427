"""torch.autograd bridges onto the C ABI: one Function per MixedOP call, one per stage sink.

PyTorch is plumbing here (device memory from its caching allocator, the current CUDA stream,
autograd bookkeeping between MixedOPs); every FLOP of the MixedOP / sink runs in
libtfnas_b200.so.  Nothing in this file falls back to torch math.
"""
import ctypes

import torch

from . import _lib
from ._lib import CandArray, CandPtrs, MixedOpDesc, check

_SLOTS = ('w1', 'dw', 'w3', 'se_rw', 'se_rb', 'se_ew', 'se_eb')


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    # raw handle of the current stream of the current device (torch.cuda.current_stream() costs ~30 us per call)
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


def _require_cuda_f32(t, name):
    if not (t.is_cuda and t.dtype == torch.float32):
        raise _lib.TfnasError('%s must be a CUDA float32 tensor (tfnas_b200 has no CPU path)' % name)


class MixedOpCall(object):
    """Static description of one MixedOP evaluation (shape, candidate set, noise, LUT row)."""

    __slots__ = ('desc', 'mask', 'alpha_mode', 'T', 'gumbel', 'lat8', 'active', 'n_per')

    def __init__(self, N, ic, oc, H, W, stride, act, mcs, ks, ses, mask, T=1.0, gumbel=None, lat8=None):
        d = MixedOpDesc()
        d.N, d.ic, d.oc, d.H, d.W, d.stride = N, ic, oc, H, W, stride
        d.act = _lib.ACT_CODE[act]
        d.num_ops = len(mcs)
        for i in range(len(mcs)):
            d.mc[i], d.k[i], d.se[i] = mcs[i], ks[i], ses[i]
        self.desc = d
        self.mask = mask
        full = (1 << len(mcs)) - 1
        self.alpha_mode = (mask & full) == full and len(mcs) > 1
        self.T = float(T)
        self.gumbel = gumbel
        self.lat8 = lat8
        self.active = [i for i in range(len(mcs)) if mask >> i & 1]
        self.n_per = [7 if ses[i] > 0 else 3 for i in self.active]

    def out_shape(self):
        d = self.desc
        return (d.N, d.oc, (d.H - 1) // d.stride + 1, (d.W - 1) // d.stride + 1)


def _cand_array(call, tensors):
    """tensors: flat list over active candidates (3 or 7 each) -> CandArray indexed by candidate id."""
    arr = CandArray()
    pos = 0
    for i, n in zip(call.active, call.n_per):
        for name, t in zip(_SLOTS[:n], tensors[pos:pos + n]):
            setattr(arr[i], name, t.data_ptr())
        pos += n
    return arr


class MixedOpFn(torch.autograd.Function):
    """forward(x, log_alphas_or_None, call, *weights) -> (out, out_lat)."""

    @staticmethod
    def forward(ctx, x, log_alphas, call, *weights):
        lib = _lib.load()
        _require_cuda_f32(x, 'x')
        x = x.contiguous()
        weights = tuple(w.contiguous() for w in weights)
        for w in weights:
            _require_cuda_f32(w, 'weight')
        d = call.desc
        if tuple(x.shape) != (d.N, d.ic, d.H, d.W):
            raise _lib.TfnasError('x shape %s does not match descriptor' % (tuple(x.shape),))
        nsaved = lib.tfnas_mixedop_saved_bytes(ctypes.byref(d), call.mask)
        nws = lib.tfnas_mixedop_workspace_bytes(ctypes.byref(d), call.mask, 0)
        if nsaved == 0:
            check(-1)
        saved = torch.empty(nsaved, dtype=torch.uint8, device=x.device)
        ws = torch.empty(nws, dtype=torch.uint8, device=x.device)
        out = torch.empty(call.out_shape(), dtype=torch.float32, device=x.device)
        out_lat = torch.zeros((), dtype=torch.float32, device=x.device)
        arr = _cand_array(call, weights)
        la = log_alphas.contiguous() if call.alpha_mode else None
        check(lib.tfnas_mixedop_fwd(ctypes.byref(d), call.mask, _ptr(x), arr, _ptr(la),
                                    _ptr(call.gumbel) if call.alpha_mode else None,
                                    _ptr(call.lat8) if call.alpha_mode else None, call.T,
                                    _ptr(out), _ptr(out_lat), _ptr(saved), nsaved, _ptr(ws), nws, _stream()))
        ctx.call = call
        ctx.saved_buf = saved
        ctx.save_for_backward(x, *weights)
        return out, out_lat

    @staticmethod
    def backward(ctx, gout, glat):
        lib = _lib.load()
        call = ctx.call
        d = call.desc
        if ctx.saved_buf is None:
            raise _lib.TfnasError('MixedOP backward ran twice: the buffers kept from forward are released after the '
                                  'first backward (retain_graph / double backward are not supported)')
        x = ctx.saved_tensors[0]
        weights = ctx.saved_tensors[1:]
        need_dx = ctx.needs_input_grad[0]
        need_da = call.alpha_mode and ctx.needs_input_grad[1]
        need_dw = any(ctx.needs_input_grad[3:])
        gout = gout.contiguous()
        dx = torch.empty_like(x) if (need_dx or need_dw) else None
        dalpha = torch.zeros(d.num_ops, dtype=torch.float32, device=x.device) if need_da else None
        arr = _cand_array(call, weights)
        grads = None
        garr = None
        if need_dw:
            grads = [torch.empty_like(w) for w in weights]
            garr = _cand_array(call, grads)
        nws = lib.tfnas_mixedop_workspace_bytes(ctypes.byref(d), call.mask, 1 if need_dw else 0)
        ws = torch.empty(nws, dtype=torch.uint8, device=x.device)
        gl = glat.contiguous() if (glat is not None and call.alpha_mode) else None
        check(lib.tfnas_mixedop_bwd(ctypes.byref(d), call.mask, _ptr(x), arr, _ptr(gout), _ptr(gl), call.T,
                                    _ptr(ctx.saved_buf), ctx.saved_buf.numel(), _ptr(dx), _ptr(dalpha), garr,
                                    _ptr(ws), nws, _stream()))
        ctx.saved_buf = None
        return (dx if need_dx else None, dalpha, None) + (tuple(grads) if grads else (None,) * len(weights))


class StageSinkFn(torch.autograd.Function):
    """forward(betas, cumlat_or_None, *res) -> (out, out_lat): models/model_search.py:202-204."""

    @staticmethod
    def forward(ctx, betas, cumlat, *res):
        lib = _lib.load()
        K = len(res)
        res = tuple(r.contiguous() for r in res)
        for r in res:
            _require_cuda_f32(r, 'res')
        out = torch.empty_like(res[0])
        out_lat = torch.zeros((), dtype=torch.float32, device=out.device)
        ptrs = (ctypes.c_void_p * K)(*[r.data_ptr() for r in res])
        cl = cumlat.contiguous() if cumlat is not None else None
        check(lib.tfnas_stage_sink_fwd(K, out.numel(), ptrs, _ptr(betas), _ptr(cl), _ptr(out),
                                       _ptr(out_lat) if cl is not None else None, _stream()))
        ctx.has_lat = cl is not None
        ctx.save_for_backward(betas, cl if cl is not None else betas, *res)
        return out, out_lat

    @staticmethod
    def backward(ctx, gout, glat):
        lib = _lib.load()
        betas, cl = ctx.saved_tensors[0], ctx.saved_tensors[1]
        res = ctx.saved_tensors[2:]
        K = len(res)
        if not ctx.has_lat:
            cl = None
        gout = gout.contiguous()
        dres = [torch.empty_like(r) for r in res]
        dbetas = torch.empty_like(betas)
        dcum = torch.empty(K, dtype=torch.float32, device=gout.device) if cl is not None else None
        ws = torch.empty(64, dtype=torch.uint8, device=gout.device)
        ptrs = (ctypes.c_void_p * K)(*[r.data_ptr() for r in res])
        dptrs = (ctypes.c_void_p * K)(*[r.data_ptr() for r in dres])
        gl = glat.contiguous() if (glat is not None and cl is not None) else None
        check(lib.tfnas_stage_sink_bwd(K, gout.numel(), ptrs, _ptr(betas), _ptr(cl), _ptr(gout), _ptr(gl), dptrs,
                                       _ptr(dbetas), _ptr(dcum), _ptr(ws), 64, _stream()))
        return (dbetas, dcum) + tuple(dres)


class BnActFn(torch.autograd.Function):
    """y = act(BN_batchstats(x)) (no affine, eps 1e-5): the BN + activation of the stems / feature-mix layer."""

    @staticmethod
    def forward(ctx, x, act):
        lib = _lib.load()
        _require_cuda_f32(x, 'x')
        x = x.contiguous()
        N, C = x.shape[0], x.shape[1]
        HW = x.numel() // (N * C)
        y = torch.empty_like(x)
        mr = torch.empty(2 * C, dtype=torch.float32, device=x.device)
        ws = torch.empty(16 * C, dtype=torch.uint8, device=x.device)
        check(lib.tfnas_bn_act_fwd(N, C, HW, _lib.ACT_CODE[act], _ptr(x), _ptr(y), _ptr(mr), _ptr(ws), 16 * C, _stream()))
        ctx.act = act
        ctx.save_for_backward(x, mr)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x, mr = ctx.saved_tensors
        gy = gy.contiguous()
        N, C = x.shape[0], x.shape[1]
        HW = x.numel() // (N * C)
        dx = torch.empty_like(x)
        ws = torch.empty(16 * C, dtype=torch.uint8, device=x.device)
        check(lib.tfnas_bn_act_bwd(N, C, HW, _lib.ACT_CODE[ctx.act], _ptr(x), _ptr(mr), _ptr(gy), _ptr(dx), _ptr(ws),
                                   16 * C, _stream()))
        return dx, None


def bn_act(x, act):
    return BnActFn.apply(x, act)


class DwConvFn(torch.autograd.Function):
    """Depthwise KxK, stride 1, padding K//2, no bias: the second stem's depth_conv (models/layers.py:486-489)."""

    @staticmethod
    def forward(ctx, x, weight):
        lib = _lib.load()
        _require_cuda_f32(x, 'x')
        _require_cuda_f32(weight, 'weight')
        x = x.contiguous()
        w = weight.contiguous()
        N, C, H, W = x.shape
        K = w.shape[-1]
        if w.shape[0] != C or w.shape[1] != 1 or w.shape[2] != K:
            raise ValueError('depthwise weight must be [C, 1, K, K]')
        y = torch.empty_like(x)
        check(lib.tfnas_dwconv_fwd(N, C, H, W, K, 1, _ptr(x), _ptr(w), _ptr(y), _stream()))
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x, w = ctx.saved_tensors
        gy = gy.contiguous()
        N, C, H, W = x.shape
        K = w.shape[-1]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        check(lib.tfnas_dwconv_bwd(N, C, H, W, K, 1, _ptr(x), _ptr(w), _ptr(gy), _ptr(dx), _ptr(dw), _stream()))
        return dx, dw


def dwconv(x, weight):
    return DwConvFn.apply(x, weight)


# ------------------------------------------------------------------------------------------------------------------
# Supernet body: all MixedStages of Network.forward in one C-ABI call per direction (csrc/body.cu)
# ------------------------------------------------------------------------------------------------------------------
class ArenaPool(object):
    """Device buffers for tfnas_body_fwd/_bwd, kept across steps (no allocation in the steady state).  A buffer is leased
    for one forward pass and handed back when the pass's autograd context dies (after backward, or right away under
    no_grad); best-fit reuse keeps one large (alpha-step) and two small (bi-sampled w-step) buffers alive."""

    def __init__(self):
        self.free = []

    def lease(self, nbytes, device):
        best = None
        for i, b in enumerate(self.free):
            if b.numel() >= nbytes and b.device == device and (best is None or b.numel() < self.free[best].numel()):
                best = i
        buf = self.free.pop(best) if best is not None else torch.empty(nbytes, dtype=torch.uint8, device=device)
        return _Lease(self, buf)

    def clear(self):
        self.free = []


class _Lease(object):
    __slots__ = ('pool', 'buf')

    def __init__(self, pool, buf):
        self.pool, self.buf = pool, buf

    def release(self):
        if self.buf is not None:
            self.pool.free.append(self.buf)
            self.buf = None

    def __del__(self):
        self.release()


class BodyCall(object):
    """One pass through the body: static shapes (BodyDesc), this pass's candidate masks and the flat tensor layout
    [live weights in (block, candidate, slot) order | log_alphas per block (alpha mode) | betas per stage]."""

    def __init__(self, desc, nblocks, nstages, masks, alpha_mode, T, gumbel, lat, active, n_per, pool, arena_bytes, out_shape):
        self.desc, self.nblocks, self.nstages = desc, nblocks, nstages
        self.masks = _lib.BodyMasks(*masks)
        self.alpha_mode, self.T, self.gumbel, self.lat = alpha_mode, float(T), gumbel, lat
        self.active = active          # per block: list of active candidate ids
        self.n_per = n_per            # per block: tensors per active candidate (3 or 7)
        self.nweights = sum(sum(n) for n in n_per)
        self.pool, self.arena_bytes, self.out_shape = pool, arena_bytes, out_shape
        self.grad_numel = 0           # floats to allocate for the flat weight-gradient buffer (0: the exact sum)

    def cand_array(self, tensors):
        arr = _lib.BodyCandArray()
        pos = 0
        for b in range(self.nblocks):
            base = b * _lib.MAX_OPS
            for i, n in zip(self.active[b], self.n_per[b]):
                c = arr[base + i]
                for name, t in zip(_SLOTS[:n], tensors[pos:pos + n]):
                    setattr(c, name, t.data_ptr())
                pos += n
        return arr


def _ptr_array(cls, tensors):
    return cls(*[t.data_ptr() for t in tensors])


class BodyFn(torch.autograd.Function):
    """forward(x, call, *tensors) -> (out, out_lat)."""

    @staticmethod
    def forward(ctx, x, call, *tensors):
        lib = _lib.load()
        _require_cuda_f32(x, 'x')
        x = x.contiguous()
        for t in tensors:
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise _lib.TfnasError('body parameters must be contiguous CUDA float32 tensors')
        nw = call.nweights
        weights = tensors[:nw]
        la = tensors[nw:nw + call.nblocks] if call.alpha_mode else ()
        betas = tensors[nw + len(la):]
        lease = call.pool.lease(call.arena_bytes, x.device)
        out = torch.empty(call.out_shape, dtype=torch.float32, device=x.device)
        out_lat = torch.zeros((), dtype=torch.float32, device=x.device)
        warr = call.cand_array(weights)
        check(lib.tfnas_body_fwd(ctypes.byref(call.desc), call.masks, _ptr(x), warr,
                                 _ptr_array(_lib.BlockPtrs, la) if call.alpha_mode else None,
                                 _ptr_array(_lib.StagePtrs, betas),
                                 _ptr(call.gumbel) if call.alpha_mode else None, _ptr(call.lat) if call.alpha_mode else None,
                                 call.T, _ptr(out), _ptr(out_lat), _ptr(lease.buf), lease.buf.numel(), _stream()))
        ctx.call = call
        ctx.lease = lease
        ctx.warr = warr
        ctx.save_for_backward(x, *tensors)
        return out, out_lat

    @staticmethod
    def backward(ctx, gout, glat):
        lib = _lib.load()
        call = ctx.call
        lease = ctx.lease
        if lease is None or lease.buf is None:
            raise _lib.TfnasError('body backward ran twice: the arena of the pass is released after the first backward')
        x = ctx.saved_tensors[0]
        tensors = ctx.saved_tensors[1:]
        nw, nb = call.nweights, call.nblocks
        la = tensors[nw:nw + nb] if call.alpha_mode else ()
        betas = tensors[nw + len(la):]
        nig = ctx.needs_input_grad
        need_dw = any(nig[2:2 + nw])
        need_da = call.alpha_mode and any(nig[2 + nw:2 + nw + nb])
        need_db = any(nig[2 + nw + len(la):])
        need_dx = nig[0] or need_dw
        dev = x.device
        gout = gout.contiguous()
        dx = torch.empty_like(x) if need_dx else None
        grads, garr = (), None
        if need_dw:
            # one flat buffer for all weight gradients of the pass (GradSync all-reduces it in place); sized for the largest
            # candidate set so that every step asks the caching allocator for the same block
            sizes = [t.numel() for t in tensors[:nw]]
            tot = sum(sizes)
            flat = torch.empty(max(tot, call.grad_numel), dtype=torch.float32, device=dev)
            grads = tuple(g.view(t.shape) for g, t in zip(flat[:tot].split(sizes), tensors[:nw]))
            garr = call.cand_array(grads)
        dla = dla_rows = None
        if need_da:
            dla = torch.zeros((nb, _lib.MAX_OPS), dtype=torch.float32, device=dev)
            dla_rows = [dla[i, :la[i].numel()] for i in range(nb)]
        dbe = None
        if need_db:
            dbe = [torch.zeros_like(b) for b in betas]
        gl = glat.contiguous() if (glat is not None and call.alpha_mode) else None
        check(lib.tfnas_body_bwd(ctypes.byref(call.desc), call.masks, _ptr(x), ctx.warr, _ptr_array(_lib.StagePtrs, betas),
                                 _ptr(gout), _ptr(gl), call.T, _ptr(dx),
                                 _ptr_array(_lib.BlockPtrs, dla_rows) if need_da else None,
                                 _ptr_array(_lib.StagePtrs, dbe) if need_db else None, garr,
                                 _ptr(lease.buf), lease.buf.numel(), _stream()))
        lease.release()
        ctx.lease = None
        out = [dx if nig[0] else None, None]
        out += list(grads) if need_dw else [None] * nw
        if call.alpha_mode:
            out += dla_rows if need_da else [None] * nb
        out += dbe if need_db else [None] * len(betas)
        return tuple(out)


# ------------------------------------------------------------------------------------------------------------------
# Stems: first_stem (conv 3x3/2 + BN + ReLU) and second_stem (MBConv without expand) in one call per direction
# ------------------------------------------------------------------------------------------------------------------
_STEM_SLOTS = ('conv_w', 'dw', 'se_rw', 'se_rb', 'se_ew', 'se_eb', 'pw')


def _stem_ptrs(tensors):
    p = _lib.StemPtrs()
    for name, t in zip(_STEM_SLOTS, tensors):
        setattr(p, name, t.data_ptr())
    return p


class StemFn(torch.autograd.Function):
    """forward(img, pool, conv_w, dw, se_rw, se_rb, se_ew, se_eb, pw) -> x0 [N, c_out, H/2, W/2]."""

    @staticmethod
    def forward(ctx, img, pool, *weights):
        lib = _lib.load()
        _require_cuda_f32(img, 'img')
        img = img.contiguous()
        for w in weights:
            if not (w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()):
                raise _lib.TfnasError('stem parameters must be contiguous CUDA float32 tensors')
        d = _lib.StemDesc()
        d.N, d.c_in, d.H, d.W = img.shape
        d.c_mid, d.se, d.c_out = weights[0].shape[0], weights[2].shape[0], weights[6].shape[0]
        need_grad = any(ctx.needs_input_grad[2:])
        nbytes = lib.tfnas_stem_arena_bytes(ctypes.byref(d), 1 if need_grad else 0)
        if not nbytes:
            check(-1)
        lease = pool.lease(nbytes, img.device)
        out = torch.empty((d.N, d.c_out, (d.H - 1) // 2 + 1, (d.W - 1) // 2 + 1), dtype=torch.float32, device=img.device)
        wp = _stem_ptrs(weights)
        check(lib.tfnas_stem_fwd(ctypes.byref(d), _ptr(img), ctypes.byref(wp), _ptr(out), _ptr(lease.buf), lease.buf.numel(),
                                 _stream()))
        ctx.desc, ctx.lease, ctx.wp = d, lease, wp
        ctx.save_for_backward(img, *weights)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        lease = ctx.lease
        if lease is None or lease.buf is None:
            raise _lib.TfnasError('stem backward ran twice: the arena of the pass is released after the first backward')
        img = ctx.saved_tensors[0]
        weights = ctx.saved_tensors[1:]
        gout = gout.contiguous()
        flat = torch.empty(sum(w.numel() for w in weights), dtype=torch.float32, device=img.device)
        grads = tuple(g.view(w.shape) for g, w in zip(flat.split([w.numel() for w in weights]), weights))
        gp = _stem_ptrs(grads)
        check(lib.tfnas_stem_bwd(ctypes.byref(ctx.desc), _ptr(img), ctypes.byref(ctx.wp), _ptr(gout), ctypes.byref(gp),
                                 _ptr(lease.buf), lease.buf.numel(), _stream()))
        lease.release()
        ctx.lease = None
        return (None, None) + grads


class HeadFn(torch.autograd.Function):
    """forward(x, pool, fm_w, fc_w, fc_b) -> logits: feature-mix 1x1 conv + BN + Swish + global average pooling + classifier."""

    @staticmethod
    def forward(ctx, x, pool, fm_w, fc_w, fc_b):
        lib = _lib.load()
        _require_cuda_f32(x, 'x')
        x = x.contiguous()
        for w in (fm_w, fc_w, fc_b):
            if not (w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()):
                raise _lib.TfnasError('head parameters must be contiguous CUDA float32 tensors')
        d = _lib.HeadDesc()
        d.N, d.c_in, d.H, d.W = x.shape
        d.c_mid, d.num_classes = fm_w.shape[0], fc_w.shape[0]
        nbytes = lib.tfnas_head_arena_bytes(ctypes.byref(d))
        if not nbytes:
            check(-1)
        lease = pool.lease(nbytes, x.device)
        logits = torch.empty((d.N, d.num_classes), dtype=torch.float32, device=x.device)
        wp = _lib.HeadPtrs(fm_w.data_ptr(), fc_w.data_ptr(), fc_b.data_ptr())
        check(lib.tfnas_head_fwd(ctypes.byref(d), _ptr(x), ctypes.byref(wp), _ptr(logits), _ptr(lease.buf), lease.buf.numel(),
                                 _stream()))
        ctx.desc, ctx.lease, ctx.wp = d, lease, wp
        ctx.save_for_backward(x, fm_w, fc_w, fc_b)
        return logits

    @staticmethod
    def backward(ctx, glogits):
        lib = _lib.load()
        lease = ctx.lease
        if lease is None or lease.buf is None:
            raise _lib.TfnasError('head backward ran twice: the arena of the pass is released after the first backward')
        x, fm_w, fc_w, fc_b = ctx.saved_tensors
        glogits = glogits.contiguous()
        dx = torch.empty_like(x)
        grads = (None, None, None)
        gp = None
        if any(ctx.needs_input_grad[2:]):
            grads = (torch.empty_like(fm_w), torch.empty_like(fc_w), torch.empty_like(fc_b))
            gpv = _lib.HeadPtrs(grads[0].data_ptr(), grads[1].data_ptr(), grads[2].data_ptr())
            gp = ctypes.byref(gpv)
        check(lib.tfnas_head_bwd(ctypes.byref(ctx.desc), _ptr(x), ctypes.byref(ctx.wp), _ptr(glogits), _ptr(dx), gp,
                                 _ptr(lease.buf), lease.buf.numel(), _stream()))
        lease.release()
        ctx.lease = None
        return (dx if ctx.needs_input_grad[0] else None, None) + grads
