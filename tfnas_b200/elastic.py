"""Elasticity scaling of the search (reference train_search.py:262-307, 465-532) and the
mask-sliced movement of weights between the max-width master copy and the per-epoch supernet
(:164-193, :235-258), written against tensors instead of ``exec`` strings."""
import copy

import numpy as np
import torch

# (parameter suffix, dimension sliced by the channel mask or None, candidate must have SE)
_SLICED = (('inverted_bottleneck.conv.weight', 0, False), ('depth_conv.conv.weight', 0, False),
           ('point_linear.conv.weight', 1, False), ('squeeze_excite.conv_reduce.weight', 1, True),
           ('squeeze_excite.conv_reduce.bias', None, True), ('squeeze_excite.conv_expand.weight', 0, True),
           ('squeeze_excite.conv_expand.bias', 0, True))


def get_lookup_latency(parsed_arch, mc_num_dddict, lat_lookup_key_dddict, lat_lookup):
    lat = lat_lookup['base']
    for stage in parsed_arch:
        for block, op_idx in parsed_arch[stage].items():
            lat += lat_lookup[lat_lookup_key_dddict[stage][block][op_idx]][mc_num_dddict[stage][block][op_idx]]
    return lat


def bound_clip(mc_num, max_mc_num):
    lo = max_mc_num // 2
    if mc_num <= lo:
        return lo, False
    if mc_num >= max_mc_num:
        return max_mc_num, False
    return mc_num, True


def fit_mc_num_by_latency(parsed_arch, mc_num_dddict, mc_maxnum_dddict, lat_lookup_key_dddict, lat_lookup,
                          target_lat, stages, sign):
    """Grow (sign=+1) or shrink (sign=-1) the chosen ops' widths, in proportion to their current
    ratios, until the LUT latency crosses target_lat or every width hits its bound."""
    assert sign in (-1, 1)
    lat = get_lookup_latency(parsed_arch, mc_num_dddict, lat_lookup_key_dddict, lat_lookup)
    picks = [(st, bl, parsed_arch[st][bl]) for st in stages for bl in parsed_arch[st]]
    cur = [mc_num_dddict[st][bl][op] for st, bl, op in picks]
    mx = [mc_maxnum_dddict[st][bl][op] for st, bl, op in picks]
    step = [int(round(c / min(cur))) for c in cur]
    free = [True] * len(picks)
    new = copy.deepcopy(mc_num_dddict)
    new_lat = lat
    while any(free) and sign * new_lat <= sign * target_lat:
        mc_num_dddict, lat = copy.deepcopy(new), new_lat
        for i, (st, bl, op) in enumerate(picks):
            new[st][bl][op], free[i] = bound_clip(mc_num_dddict[st][bl][op] + sign * step[i], mx[i])
        new_lat = get_lookup_latency(parsed_arch, new, lat_lookup_key_dddict, lat_lookup)
    if sign == -1:
        return copy.deepcopy(new), new_lat
    return mc_num_dddict, lat


def rescale_widths(parsed_arch, mc_num_dddict, mc_maxnum_dddict, keys, lut, target_lat):
    """Shrink-then-re-expand or expand schedule of train_search.py:270-287."""
    before = get_lookup_latency(parsed_arch, mc_num_dddict, keys, lut)
    all_stages = ['stage%d' % i for i in range(1, 7)]
    if before == target_lat:
        return mc_num_dddict, before, before
    first = -1 if before > target_lat else 1
    mc, after = fit_mc_num_by_latency(parsed_arch, mc_num_dddict, mc_maxnum_dddict, keys, lut, target_lat, all_stages, first)
    for start in range(2, 7):
        mc, after = fit_mc_num_by_latency(parsed_arch, mc, mc_maxnum_dddict, keys, lut, target_lat,
                                          ['stage%d' % i for i in range(start, 7)], 1)
    return mc, before, after


def reselect_channels(mask_dddict, mc_num_dddict, parsed_arch, state_dict, prefix='module.'):
    """For the CHOSEN op of each block whose width changed: keep the channels with the largest L1
    norm of their depthwise filters in the max-width master copy (train_search.py:293-305)."""
    for stage in parsed_arch:
        for block, op_idx in parsed_arch[stage].items():
            mask = mask_dddict[stage][block][op_idx]
            want = mc_num_dddict[stage][block][op_idx]
            if want == int(mask.sum().item()):
                continue
            w = state_dict['%s%s.%s.m_ops.%d.depth_conv.conv.weight' % (prefix, stage, block, op_idx)]
            order = np.argsort(np.sum(np.abs(w.detach().cpu().numpy()), axis=(1, 2, 3)))[::-1][:want]
            mask.zero_()
            mask[order.tolist()] = 1.0


def _cand_keys(prefix, stage, block, op_idx):
    base = '%s%s.%s.m_ops.%d.' % (prefix, stage, block, op_idx)
    return [(base + sfx, dim, se) for sfx, dim, se in _SLICED]


def load_from_master(model, state_dict, mask_dddict, prefix='module.'):
    """Fill the (narrow) supernet from the max-width master state_dict (train_search.py:164-193)."""
    own = dict(model.named_parameters())
    with torch.no_grad():
        for key, val in state_dict.items():
            if 'm_ops' not in key:
                own[key].data = val.data.to(own[key].device)
        for stage in mask_dddict:
            for block in mask_dddict[stage]:
                for op_idx, mask in mask_dddict[stage][block].items():
                    index = torch.nonzero(mask).view(-1)
                    for key, dim, se in _cand_keys(prefix, stage, block, op_idx):
                        if se and op_idx < 4:
                            continue
                        src = state_dict[key]
                        idx = index.to(src.device)
                        own[key].data = (src if dim is None else torch.index_select(src, dim, idx)).data.to(own[key].device)


def store_to_master(state_dict, model, mask_dddict, prefix='module.'):
    """Scatter the trained narrow tensors back into the master copy (train_search.py:235-258)."""
    cur = model.state_dict()
    with torch.no_grad():
        for key in state_dict:
            if 'm_ops' not in key:
                state_dict[key].data = cur[key].data
        for stage in mask_dddict:
            for block in mask_dddict[stage]:
                for op_idx, mask in mask_dddict[stage][block].items():
                    index = torch.nonzero(mask).view(-1)
                    for key, dim, se in _cand_keys(prefix, stage, block, op_idx):
                        if se and op_idx < 4:
                            continue
                        dst, src = state_dict[key].data, cur[key].to(state_dict[key].device)
                        idx = index.to(dst.device)
                        if dim is None:
                            dst[:] = src
                        elif dim == 0:
                            dst[idx] = src
                        else:
                            dst[:, idx] = src
