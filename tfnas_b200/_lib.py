"""ctypes binding of libtfnas_b200.so (the C ABI declared in include/tfnas_b200.h).

There is NO fallback: if the shared library is missing or fails to load, importing the ops
raises.  Build it with ``python -m tfnas_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes
import os

MAX_OPS = 8
ACT_RELU, ACT_SWISH = 0, 1
ACT_NONE = 2
ACT_CODE = {'relu': ACT_RELU, 'swish': ACT_SWISH, None: ACT_NONE, 'none': ACT_NONE}

_HERE = os.path.dirname(os.path.abspath(__file__))
# TFNAS_B200_LIB selects another build of the same ABI (e.g. the phase-trace debug variant)
LIB_PATH = os.environ.get('TFNAS_B200_LIB') or os.path.join(_HERE, 'lib', 'libtfnas_b200.so')


class MixedOpDesc(ctypes.Structure):
    _fields_ = [('N', ctypes.c_int32), ('ic', ctypes.c_int32), ('oc', ctypes.c_int32),
                ('H', ctypes.c_int32), ('W', ctypes.c_int32), ('stride', ctypes.c_int32),
                ('act', ctypes.c_int32), ('num_ops', ctypes.c_int32),
                ('mc', ctypes.c_int32 * MAX_OPS), ('k', ctypes.c_int32 * MAX_OPS),
                ('se', ctypes.c_int32 * MAX_OPS)]


class CandPtrs(ctypes.Structure):
    _fields_ = [('w1', ctypes.c_void_p), ('dw', ctypes.c_void_p), ('w3', ctypes.c_void_p),
                ('se_rw', ctypes.c_void_p), ('se_rb', ctypes.c_void_p),
                ('se_ew', ctypes.c_void_p), ('se_eb', ctypes.c_void_p)]


class ProfEntry(ctypes.Structure):
    _fields_ = [('name', ctypes.c_char * 32), ('ms', ctypes.c_double), ('bytes', ctypes.c_double),
                ('flops', ctypes.c_double), ('launches', ctypes.c_int64)]


MAX_BLOCKS, MAX_STAGES = 32, 8


class BodyDesc(ctypes.Structure):
    _fields_ = [('num_stages', ctypes.c_int32), ('num_blocks', ctypes.c_int32),
                ('stage_blocks', ctypes.c_int32 * MAX_STAGES), ('op', MixedOpDesc * MAX_BLOCKS)]


class StemDesc(ctypes.Structure):
    _fields_ = [('N', ctypes.c_int32), ('H', ctypes.c_int32), ('W', ctypes.c_int32), ('c_in', ctypes.c_int32),
                ('c_mid', ctypes.c_int32), ('se', ctypes.c_int32), ('c_out', ctypes.c_int32)]


class StemPtrs(ctypes.Structure):
    _fields_ = [('conv_w', ctypes.c_void_p), ('dw', ctypes.c_void_p), ('se_rw', ctypes.c_void_p),
                ('se_rb', ctypes.c_void_p), ('se_ew', ctypes.c_void_p), ('se_eb', ctypes.c_void_p), ('pw', ctypes.c_void_p)]


class HeadDesc(ctypes.Structure):
    _fields_ = [('N', ctypes.c_int32), ('H', ctypes.c_int32), ('W', ctypes.c_int32), ('c_in', ctypes.c_int32),
                ('c_mid', ctypes.c_int32), ('num_classes', ctypes.c_int32)]


class HeadPtrs(ctypes.Structure):
    _fields_ = [('fm_w', ctypes.c_void_p), ('fc_w', ctypes.c_void_p), ('fc_b', ctypes.c_void_p)]


class SgdTensor(ctypes.Structure):
    _fields_ = [('p', ctypes.c_void_p), ('g', ctypes.c_void_p), ('buf', ctypes.c_void_p), ('numel', ctypes.c_int64)]


class AdamTensor(ctypes.Structure):
    _fields_ = [('p', ctypes.c_void_p), ('g', ctypes.c_void_p), ('m', ctypes.c_void_p), ('v', ctypes.c_void_p),
                ('numel', ctypes.c_int32), ('renorm', ctypes.c_int32)]


class ProfLaunch(ctypes.Structure):
    _fields_ = [('name', ctypes.c_char * 32), ('stream', ctypes.c_uint64), ('start_ms', ctypes.c_double),
                ('end_ms', ctypes.c_double)]


CandArray = CandPtrs * MAX_OPS
BodyCandArray = CandPtrs * (MAX_OPS * MAX_BLOCKS)
BodyMasks = ctypes.c_uint32 * MAX_BLOCKS
BlockPtrs = ctypes.c_void_p * MAX_BLOCKS
StagePtrs = ctypes.c_void_p * MAX_STAGES
EXPORTS = ['tfnas_version', 'tfnas_last_error', 'tfnas_launch_count',
           'tfnas_mixedop_saved_bytes', 'tfnas_mixedop_workspace_bytes',
           'tfnas_mixedop_fwd', 'tfnas_mixedop_bwd',
           'tfnas_stage_sink_fwd', 'tfnas_stage_sink_bwd', 'tfnas_debug_saved_layout', 'tfnas_debug_bwd_layout',
           'tfnas_prof_enable', 'tfnas_prof_collect', 'tfnas_prof_timeline', 'tfnas_config_side_stream', 'tfnas_umma_selftest', 'tfnas_bn_act_fwd', 'tfnas_bn_act_bwd',
           'tfnas_dwconv_fwd', 'tfnas_dwconv_bwd', 'tfnas_debug_ws_config',
           'tfnas_body_arena_bytes', 'tfnas_body_fwd', 'tfnas_body_bwd',
           'tfnas_stem_arena_bytes', 'tfnas_stem_fwd', 'tfnas_stem_bwd',
           'tfnas_head_arena_bytes', 'tfnas_head_fwd', 'tfnas_head_bwd',
           'tfnas_sgd_step', 'tfnas_adam_step', 'tfnas_softmax_ce', 'tfnas_softmax_ce_smooth']

_lib = None


class TfnasError(RuntimeError):
    pass


def load():
    """Load the library once; raise loudly when it is absent (no CPU / torch fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TfnasError('%s not built: run `python -m tfnas_b200.build` (needs nvcc); '
                         'tfnas_b200 has no fallback path' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, sz, u32, f32, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_float, ctypes.c_int
    dp = ctypes.POINTER(MixedOpDesc)
    cp = ctypes.POINTER(CandPtrs)
    lib.tfnas_version.restype = i32
    lib.tfnas_last_error.restype = ctypes.c_char_p
    lib.tfnas_launch_count.restype = ctypes.c_uint64
    lib.tfnas_mixedop_saved_bytes.restype = sz
    lib.tfnas_mixedop_saved_bytes.argtypes = [dp, u32]
    lib.tfnas_mixedop_workspace_bytes.restype = sz
    lib.tfnas_mixedop_workspace_bytes.argtypes = [dp, u32, i32]
    lib.tfnas_mixedop_fwd.restype = i32
    lib.tfnas_mixedop_fwd.argtypes = [dp, u32, vp, cp, vp, vp, vp, f32, vp, vp, vp, sz, vp, sz, vp]
    lib.tfnas_mixedop_bwd.restype = i32
    lib.tfnas_mixedop_bwd.argtypes = [dp, u32, vp, cp, vp, vp, f32, vp, sz, vp, vp, cp, vp, sz, vp]
    lib.tfnas_stage_sink_fwd.restype = i32
    lib.tfnas_stage_sink_fwd.argtypes = [i32, sz, ctypes.POINTER(vp), vp, vp, vp, vp, vp]
    lib.tfnas_stage_sink_bwd.restype = i32
    lib.tfnas_stage_sink_bwd.argtypes = [i32, sz, ctypes.POINTER(vp), vp, vp, vp, vp, ctypes.POINTER(vp), vp, vp,
                                         vp, sz, vp]
    lib.tfnas_debug_saved_layout.restype = i32
    lib.tfnas_debug_saved_layout.argtypes = [dp, u32, ctypes.POINTER(sz)]
    lib.tfnas_debug_ws_config.restype = i32
    lib.tfnas_debug_ws_config.argtypes = [i32, i32, ctypes.POINTER(u32)]
    lib.tfnas_debug_bwd_layout.restype = i32
    lib.tfnas_debug_bwd_layout.argtypes = [dp, u32, i32, ctypes.POINTER(sz)]
    lib.tfnas_bn_act_fwd.restype = i32
    lib.tfnas_bn_act_fwd.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, sz, vp]
    lib.tfnas_bn_act_bwd.restype = i32
    lib.tfnas_bn_act_bwd.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, vp, sz, vp]
    lib.tfnas_dwconv_fwd.restype = i32
    lib.tfnas_dwconv_fwd.argtypes = [i32, i32, i32, i32, i32, i32, vp, vp, vp, vp]
    lib.tfnas_dwconv_bwd.restype = i32
    lib.tfnas_dwconv_bwd.argtypes = [i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.tfnas_prof_enable.restype = i32
    lib.tfnas_prof_enable.argtypes = [i32]
    lib.tfnas_prof_collect.restype = i32
    lib.tfnas_prof_collect.argtypes = [ctypes.POINTER(ProfEntry), i32]
    lib.tfnas_prof_timeline.restype = i32
    lib.tfnas_prof_timeline.argtypes = [ctypes.POINTER(ProfLaunch), i32]
    lib.tfnas_config_side_stream.restype = i32
    lib.tfnas_config_side_stream.argtypes = [i32]
    lib.tfnas_umma_selftest.restype = i32
    lib.tfnas_umma_selftest.argtypes = [i32, i32, i32, vp, vp, vp, vp, sz, i32, vp]
    bp = ctypes.POINTER(BodyDesc)
    mp = ctypes.POINTER(u32)
    pp = ctypes.POINTER(vp)
    lib.tfnas_body_arena_bytes.restype = sz
    lib.tfnas_body_arena_bytes.argtypes = [bp, mp, i32]
    lib.tfnas_body_fwd.restype = i32
    lib.tfnas_body_fwd.argtypes = [bp, mp, vp, cp, pp, pp, vp, vp, f32, vp, vp, vp, sz, vp]
    lib.tfnas_body_bwd.restype = i32
    lib.tfnas_body_bwd.argtypes = [bp, mp, vp, cp, pp, vp, vp, f32, vp, pp, pp, cp, vp, sz, vp]
    sdp, spp = ctypes.POINTER(StemDesc), ctypes.POINTER(StemPtrs)
    lib.tfnas_stem_arena_bytes.restype = sz
    lib.tfnas_stem_arena_bytes.argtypes = [sdp, i32]
    lib.tfnas_stem_fwd.restype = i32
    lib.tfnas_stem_fwd.argtypes = [sdp, vp, spp, vp, vp, sz, vp]
    lib.tfnas_stem_bwd.restype = i32
    lib.tfnas_stem_bwd.argtypes = [sdp, vp, spp, vp, spp, vp, sz, vp]
    hdp, hpp = ctypes.POINTER(HeadDesc), ctypes.POINTER(HeadPtrs)
    lib.tfnas_head_arena_bytes.restype = sz
    lib.tfnas_head_arena_bytes.argtypes = [hdp]
    lib.tfnas_head_fwd.restype = i32
    lib.tfnas_head_fwd.argtypes = [hdp, vp, hpp, vp, vp, sz, vp]
    lib.tfnas_head_bwd.restype = i32
    lib.tfnas_head_bwd.argtypes = [hdp, vp, hpp, vp, vp, hpp, vp, sz, vp]
    lib.tfnas_sgd_step.restype = i32
    lib.tfnas_sgd_step.argtypes = [i32, ctypes.POINTER(SgdTensor), f32, f32, f32, f32, f32, vp, vp, sz, vp]
    lib.tfnas_adam_step.restype = i32
    lib.tfnas_adam_step.argtypes = [i32, ctypes.POINTER(AdamTensor), i32, f32, f32, f32, f32, f32, f32, f32, vp]
    lib.tfnas_softmax_ce.restype = i32
    lib.tfnas_softmax_ce.argtypes = [i32, i32, vp, vp, vp, vp, vp]
    lib.tfnas_softmax_ce_smooth.restype = i32
    lib.tfnas_softmax_ce_smooth.argtypes = [i32, i32, vp, vp, ctypes.c_float, vp, vp, vp]
    if lib.tfnas_version() != 1:
        raise TfnasError('ABI version mismatch: %d' % lib.tfnas_version())
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise TfnasError('tfnas_b200 error %d: %s' % (rc, load().tfnas_last_error().decode()))


def launch_count():
    return int(load().tfnas_launch_count())


def prof_enable(on=True):
    check(load().tfnas_prof_enable(1 if on else 0))


def prof_collect(max_entries=64):
    """-> list of dicts {name, ms, bytes, flops, launches}, aggregated per kernel name."""
    buf = (ProfEntry * max_entries)()
    n = load().tfnas_prof_collect(buf, max_entries)
    if n < 0:
        check(n)
    return [dict(name=buf[i].name.decode(), ms=buf[i].ms, bytes=buf[i].bytes, flops=buf[i].flops,
                 launches=int(buf[i].launches)) for i in range(n)]


def prof_timeline(max_entries=20000):
    """-> list of (name, stream, start_ms, end_ms) per recorded launch, in launch order."""
    buf = (ProfLaunch * max_entries)()
    n = load().tfnas_prof_timeline(buf, max_entries)
    if n < 0:
        check(n)
    return [(buf[i].name.decode(), int(buf[i].stream), buf[i].start_ms, buf[i].end_ms) for i in range(n)]
