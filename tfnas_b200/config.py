"""Search-space tables of the TF-NAS supernet, generated (not transcribed).

The reference ships these as two hand-unrolled literals in ``tools/config.py``:
``mc_mask_dddict`` (:4-197, per stage/block/op 0/1 channel masks of length
4*ic / 8*ic whose first 3*ic / 6*ic entries are one) and
``lat_lookup_key_dddict`` (:200-393, the LUT key of every candidate).  Both are
fully determined by the stage table of ``models/model_search.py:221-274`` and
the candidate list ``:7-29``; this module derives them from that table.
``tests/test_config_tables.py`` checks them entry by entry against the
reference when ``/root/reference`` is mounted, and against a committed digest
otherwise.
"""
from collections import OrderedDict

import torch

# (kernel, expand, se_mult) per candidate, reference models/model_search.py:7-29
PRIMITIVES = [
    'MBI_k3_e3',
    'MBI_k3_e6',
    'MBI_k5_e3',
    'MBI_k5_e6',
    'MBI_k3_e3_se',
    'MBI_k3_e6_se',
    'MBI_k5_e3_se',
    'MBI_k5_e6_se',
]
#            k  e  se_mult (se_channels = se_mult * ic)
CAND_SPEC = [(3, 3, 0), (3, 6, 0), (5, 3, 0), (5, 6, 0),
             (3, 3, 1), (3, 6, 2), (5, 3, 1), (5, 6, 2)]
NUM_OPS = len(PRIMITIVES)

# stage -> (ics, ocs, strides, act, stage_type); models/model_search.py:221-274
STAGE_SPEC = OrderedDict([
    ('stage1', dict(ics=[16, 24], ocs=[24, 24], ss=[2, 1], act='relu', stage_type=1)),
    ('stage2', dict(ics=[24, 40, 40], ocs=[40, 40, 40], ss=[2, 1, 1], act='swish', stage_type=2)),
    ('stage3', dict(ics=[40, 80, 80, 80], ocs=[80, 80, 80, 80], ss=[2, 1, 1, 1], act='swish', stage_type=3)),
    ('stage4', dict(ics=[80, 112, 112, 112], ocs=[112, 112, 112, 112], ss=[1, 1, 1, 1], act='swish', stage_type=3)),
    ('stage5', dict(ics=[112, 192, 192, 192], ocs=[192, 192, 192, 192], ss=[2, 1, 1, 1], act='swish', stage_type=3)),
    ('stage6', dict(ics=[192], ocs=[320], ss=[1], act='swish', stage_type=0)),
])
# spatial size entering stage1 for a 224x224 image (two stride-2 stems... the
# first stem is stride 2, the second stride 1): 112
STAGE1_INPUT_SIZE = 112


def block_shapes(input_size=STAGE1_INPUT_SIZE):
    """Yield (stage, block, ic, oc, stride, act, in_size) for the 18 MixedOPs in forward order."""
    size = input_size
    for stage, sp in STAGE_SPEC.items():
        for j, (ic, oc, s) in enumerate(zip(sp['ics'], sp['ocs'], sp['ss']), start=1):
            yield stage, 'block%d' % j, ic, oc, s, sp['act'], size
            size = size // s


def make_mc_mask_dddict():
    """Fresh copy of the initial channel masks (reference tools/config.py:4-197)."""
    d = OrderedDict()
    for stage, block, ic, _oc, _s, _act, _size in block_shapes():
        d.setdefault(stage, OrderedDict())[block] = OrderedDict()
        for op_idx, (_k, e, _se) in enumerate(CAND_SPEC):
            ones, zeros = (3, 1) if e == 3 else (6, 2)
            d[stage][block][op_idx] = torch.cat((torch.ones(ic * ones), torch.zeros(ic * zeros)))
    return d


def lut_key(size, ic, se, oc, k, s, act):
    """LUT key grammar, reference models/model_search.py:99-107."""
    return 'MBInvertedResBlock_{}_{}_{}_{}_k{}_s{}_{}'.format(size, ic, se, oc, k, s, act)


def make_lat_lookup_key_dddict():
    """LUT key per stage/block/op (reference tools/config.py:200-393)."""
    d = OrderedDict()
    for stage, block, ic, oc, s, act, size in block_shapes():
        d.setdefault(stage, OrderedDict())[block] = OrderedDict()
        for op_idx, (k, _e, se_mult) in enumerate(CAND_SPEC):
            d[stage][block][op_idx] = lut_key(size, ic, se_mult * ic, oc, k, s, act)
    return d


mc_mask_dddict = make_mc_mask_dddict()
lat_lookup_key_dddict = make_lat_lookup_key_dddict()


def get_mc_num_dddict(mc_mask_dddict, is_max=False):
    """Mask -> width table; same contract as reference parsing_model.py:76-88."""
    out = OrderedDict()
    for stage in mc_mask_dddict:
        out[stage] = OrderedDict()
        for block in mc_mask_dddict[stage]:
            out[stage][block] = OrderedDict()
            for op_idx, mask in mc_mask_dddict[stage][block].items():
                out[stage][block][op_idx] = mask.size(0) if is_max else int(mask.sum().item())
    return out
