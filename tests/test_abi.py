"""The C-ABI library loads without a GPU, exports every symbol include/tfnas_b200.h declares, and
rejects bad descriptors with error codes (no compute calls here)."""
import ctypes
import os
import re

import pytest

from tfnas_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'tfnas_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(tfnas_[a-z0-9_]+)\s*\(', src)))


def test_exports_match_header():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.EXPORTS) == names
    assert lib.tfnas_version() == 1


def _desc(**kw):
    d = _lib.MixedOpDesc()
    d.N, d.ic, d.oc, d.H, d.W, d.stride, d.act, d.num_ops = 2, 16, 24, 8, 8, 2, 0, 8
    for i in range(8):
        d.mc[i], d.k[i], d.se[i] = 48, (3, 3, 5, 5)[i % 4], (16 if i >= 4 else 0)
    for k, v in kw.items():
        setattr(d, k, v)
    return d


def test_sizes_and_validation_without_gpu():
    lib = _lib.load()
    d = _desc()
    full = lib.tfnas_mixedop_saved_bytes(ctypes.byref(d), 0xFF)
    one = lib.tfnas_mixedop_saved_bytes(ctypes.byref(d), 0x02)
    assert full > one > 0
    # D + UH + Z dominate: N*MC*(HW + HWo)*4 + N*8*oc*HWo*4
    assert full >= 2 * 384 * (64 + 16) * 4 + 2 * 8 * 24 * 16 * 4
    assert lib.tfnas_mixedop_workspace_bytes(ctypes.byref(d), 0xFF, 0) > 0
    # max(forward scratch, backward scratch): the weight-gradient regions can only grow it
    assert lib.tfnas_mixedop_workspace_bytes(ctypes.byref(d), 0x02, 1) >= lib.tfnas_mixedop_workspace_bytes(ctypes.byref(d), 0x02, 0)
    o0, o1 = (ctypes.c_size_t * 10)(), (ctypes.c_size_t * 10)()
    assert lib.tfnas_debug_bwd_layout(ctypes.byref(d), 0x02, 0, o0) == 0 and lib.tfnas_debug_bwd_layout(ctypes.byref(d), 0x02, 1, o1) == 0
    assert o1[9] > o0[9]
    bad = _desc(stride=3)
    assert lib.tfnas_mixedop_saved_bytes(ctypes.byref(bad), 0xFF) == 0
    assert b'stride' in lib.tfnas_last_error()
    bad = _desc(ic=400)
    assert lib.tfnas_mixedop_saved_bytes(ctypes.byref(bad), 0xFF) == 0
    arr = _lib.CandArray()
    rc = lib.tfnas_mixedop_fwd(ctypes.byref(_desc(num_ops=9)), 0xFF, None, arr, None, None, None, 1.0, None, None,
                               None, 0, None, 0, None)
    assert rc == -1
    rc = lib.tfnas_mixedop_fwd(ctypes.byref(d), 0xFF, None, arr, None, None, None, 1.0, None, None, None, 0, None, 0, None)
    assert rc == -1 and b'null' in lib.tfnas_last_error()
    rc = lib.tfnas_stage_sink_fwd(7, 16, None, None, None, None, None, None)
    assert rc == -1
    with pytest.raises(_lib.TfnasError):
        _lib.check(rc)


def test_persistent_gemm_configs_fit_the_sm():
    """Host-side check of the pipeline configurations the persistent tcgen05 GEMMs pick (umma_ws.cu::ws_fit) for every N
    chunk width the supernet produces and for the extremes: at most 227 KB of shared memory, operand stages a multiple of
    the producer groups, at least two weight slots up to 176 columns, at most 1024 threads."""
    import ctypes
    from tfnas_b200 import _lib
    lib = _lib.load()
    out = (ctypes.c_uint32 * 5)()
    widths = sorted(set([16, 32, 48, 80, 112, 128, 144, 160, 176, 192, 208, 256]))
    for which in range(4):
        for nc in widths:
            rc = lib.tfnas_debug_ws_config(which, nc, out)
            if rc != 0:
                assert nc > 192, (which, nc)          # only chunks wider than any prep cap may be refused
                continue
            S, NB, smem, G, nthr = list(out)
            assert smem <= 232448 and S >= G and S % G == 0 and S <= 8 and 1 <= NB <= 8 and nthr <= 1024, (which, nc, list(out))
            if nc <= 176:        # (project at 192 columns runs with a single weight slot: 4 stages + 48 KB slots + sums)
                assert NB >= 2, (which, nc, list(out))
    assert lib.tfnas_debug_ws_config(7, 32, out) != 0 and lib.tfnas_debug_ws_config(1, 24, out) != 0


def test_network_level_descriptors_validate_without_gpu():
    """tfnas_{stem,body,head}_arena_bytes: sizes for valid descriptors, 0 (with a message) for invalid ones; host only."""
    lib = _lib.load()
    sd = _lib.StemDesc(4, 224, 224, 3, 32, 8, 16)
    n0, n1 = lib.tfnas_stem_arena_bytes(ctypes.byref(sd), 0), lib.tfnas_stem_arena_bytes(ctypes.byref(sd), 1)
    assert n1 >= n0 > 4 * 32 * 112 * 112 * 4 * 2            # at least UH + D
    sd.c_mid = 48
    assert lib.tfnas_stem_arena_bytes(ctypes.byref(sd), 0) == 0 and b'3 -> 32' in lib.tfnas_last_error()
    hd = _lib.HeadDesc(4, 7, 7, 320, 1280, 100)
    assert lib.tfnas_head_arena_bytes(ctypes.byref(hd)) > 4 * 1280 * 49 * 4 * 2
    hd.num_classes = 0
    assert lib.tfnas_head_arena_bytes(ctypes.byref(hd)) == 0
    bd = _lib.BodyDesc()
    bd.num_stages, bd.num_blocks = 2, 3
    bd.stage_blocks[0], bd.stage_blocks[1] = 2, 1
    shapes = [(16, 24, 2, 16), (24, 24, 1, 8), (24, 40, 2, 8)]
    for i, (ic, oc, s, hw) in enumerate(shapes):
        d = bd.op[i]
        d.N, d.ic, d.oc, d.H, d.W, d.stride, d.act, d.num_ops = 2, ic, oc, hw, hw, s, 1, 8
        for j in range(8):
            d.mc[j], d.k[j], d.se[j] = 3 * ic, (3, 3, 5, 5)[j % 4], (ic if j >= 4 else 0)
    full = _lib.BodyMasks(*([0xFF] * 3))
    one = _lib.BodyMasks(*([0x04] * 3))
    a, b = lib.tfnas_body_arena_bytes(ctypes.byref(bd), full, 0), lib.tfnas_body_arena_bytes(ctypes.byref(bd), one, 1)
    assert a > b > 0
    bd.op[1].ic = 32                                          # does not chain with op 0's 24 output channels
    assert lib.tfnas_body_arena_bytes(ctypes.byref(bd), full, 0) == 0 and b'chain' in lib.tfnas_last_error()
    bd.op[1].ic = 24
    bd.num_blocks = 4
    assert lib.tfnas_body_arena_bytes(ctypes.byref(bd), full, 0) == 0
