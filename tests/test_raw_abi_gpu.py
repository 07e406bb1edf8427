"""A client with NO Python model classes: tests/helpers.raw_call drives tfnas_mixedop_fwd / _bwd through ctypes with plain
device pointers (what a non-PyTorch binding would do) and the results are checked against the CPU oracle."""
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('ic,oc,s,act,size,N', [(24, 40, 2, 'swish', 14, 3), (16, 16, 1, 'relu', 12, 2)])
def test_raw_abi_alpha_mode(ic, oc, s, act, size, N):
    mcs = H.default_mcs(ic)
    P, x, gum, lats = H.make_conditioned_problem(ic, oc, s, size, N, mcs, 77 + ic, act)
    G = torch.randn(N, oc, (size - 1) // s + 1, (size - 1) // s + 1, generator=torch.Generator().manual_seed(3))
    r = H.raw_call(P, x, gum, lats, ic, oc, s, act, mcs, 0xFF, T=5.0, G=G, dlat=0.41)
    ref = H.oracle_alpha(P, x, gum, lats, ic, oc, s, act, 5.0, G, 0.41)
    e = dict(out=H.rel_l2(r['out'], ref['out']), lat=abs(r['out_lat'] - ref['out_lat']), dx=H.rel_l2(r['dx'], ref['dx']),
             dalpha=H.rel_l2(r['dalpha'], ref['dalpha']))
    print('raw ABI alpha mode', e)
    assert e['out'] < 1e-4 and e['lat'] < 1e-5 and e['dx'] < 1e-3 and e['dalpha'] < 1e-3


def test_raw_abi_sampled_mode_with_weight_grads():
    ic, oc, s, act, size, N, idx = 40, 40, 1, 'swish', 14, 4, 6
    mcs = H.default_mcs(ic, ragged=True)
    P, x, gum, lats = H.make_conditioned_problem(ic, oc, s, size, N, mcs, 5, act, active=[idx])
    G = torch.randn(N, oc, size, size, generator=torch.Generator().manual_seed(4))
    r = H.raw_call(P, x, gum, lats, ic, oc, s, act, mcs, 1 << idx, G=G, want_wgrad=True)
    ref = H.oracle_single(P, x, ic, oc, s, act, idx, G)
    assert H.rel_l2(r['out'], ref['out']) < 1e-4 and H.rel_l2(r['dx'], ref['dx']) < 1e-3
    for k, g in ref['wgrads'].items():
        assert H.rel_l2(r['wgrads'][k].reshape(g.shape), g) < 1e-3, k


def test_raw_abi_rejects_misaligned_and_bad_arguments():
    import ctypes
    from tfnas_b200 import _lib
    lib = _lib.load()
    d = _lib.MixedOpDesc()
    d.N, d.ic, d.oc, d.H, d.W, d.stride, d.act, d.num_ops = 2, 8, 8, 8, 8, 1, 0, 8
    for i in range(8):
        d.mc[i], d.k[i], d.se[i] = 24, 3 if i % 4 < 2 else 5, 0
    ns, nw = lib.tfnas_mixedop_saved_bytes(ctypes.byref(d), 0x01), lib.tfnas_mixedop_workspace_bytes(ctypes.byref(d), 0x01, 0)
    buf = torch.zeros(ns + nw + 64, dtype=torch.uint8, device='cuda')
    x = torch.zeros(2 * 8 * 64 + 1, device='cuda')
    out = torch.zeros(2 * 8 * 64, device='cuda')
    w = [torch.zeros(24 * 8, device='cuda'), torch.zeros(24 * 9, device='cuda'), torch.zeros(8 * 24, device='cuda')]
    arr = _lib.CandArray()
    arr[0].w1, arr[0].dw, arr[0].w3 = [t.data_ptr() for t in w]
    vp = ctypes.c_void_p
    args = lambda xp: (ctypes.byref(d), 0x01, vp(xp), arr, None, None, None, 1.0, vp(out.data_ptr()), None,
                       vp(buf.data_ptr()), ns, vp(buf.data_ptr() + ((ns + 255) // 256) * 256), nw, None)
    assert lib.tfnas_mixedop_fwd(*args(x.data_ptr() + 4)) == -1 and b'aligned' in lib.tfnas_last_error()
    assert lib.tfnas_mixedop_fwd(*args(x.data_ptr())) == 0
    d.stride = 3
    assert lib.tfnas_mixedop_fwd(*args(x.data_ptr())) == -1
    torch.cuda.synchronize()
