"""alpha/beta -> discrete architecture.  Same contracts as the reference parsing_model.py:23-88
(``get_op_and_depth_weights``, ``parse_architecture``, ``get_mc_num_dddict``), which
train_search.py imports for its per-epoch shrink/expand step."""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from .config import STAGE_SPEC, get_mc_num_dddict  # noqa: F401  (re-exported)


def get_op_and_depth_weights(model_or_path):
    """exp(log_alphas) per MixedOP and softmax(betas) per stage, in state_dict order."""
    if isinstance(model_or_path, str):
        state_dict = torch.load(model_or_path, map_location='cpu')['state_dict']
    else:
        state_dict = model_or_path.state_dict()
    op_weights, depth_weights = [], []
    for key, val in state_dict.items():
        if key.endswith('log_alphas'):
            op_weights.append(np.exp(val.detach().cpu().numpy()))
        elif key.endswith('betas'):
            depth_weights.append(F.softmax(val.detach().cpu(), dim=-1).numpy())
    return op_weights, depth_weights


def parse_architecture(op_weights, depth_weights):
    """argmax candidate per block; keep blocks 1..argmax(beta)+1 of every stage."""
    arch = OrderedDict()
    it = iter(op_weights)
    for stage, sp in STAGE_SPEC.items():
        arch[stage] = OrderedDict(('block%d' % (j + 1), int(np.argmax(next(it)))) for j in range(len(sp['ics'])))
    for stage, dw in zip(arch, depth_weights):
        keep = int(np.argmax(dw)) + 1
        for block in [b for b in arch[stage] if int(b[5:]) > keep]:
            del arch[stage][block]
    return arch
