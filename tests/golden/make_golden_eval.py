"""Generates tests/golden/derived_net.npz by running the REAL reference (models/model_eval.py) in the build container:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_eval.py

A derived network of a fixed parsed architecture (one block fewer in stages 2 and 4, every candidate type used), default
init under torch.manual_seed(2), then: train-mode logits of batch A (batch statistics; updates the running statistics),
eval-mode logits of batch B (running statistics), the table latency and a few state_dict probes.  The fixture travels to
the GPU box; the reference does not."""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from tests import golden_inputs as gi  # noqa: E402


def main():
    ref_shim._import()
    from models import model_eval as rme
    import parsing_model as pm
    from tools import config as rcfg
    arch, mc = gi.derived_arch(), pm.get_mc_num_dddict(rcfg.mc_mask_dddict)
    lut = ref_shim.load_lut()
    torch.manual_seed(2)
    net = rme.Network(10, arch, mc, lut, 0.0, 0.0)
    sd = net.state_dict()
    probes = OrderedDict((k, sd[k].double().sum().item()) for k in list(sd)[::13])       # initial state
    xa, xb = gi.derived_inputs()
    net.train()
    la = net(xa).detach()
    net.eval()
    with torch.no_grad():
        lb = net(xb)
        lat = float(net.get_lookup_latency(torch.zeros(1, 3, 224, 224)))
    sd = net.state_dict()
    running = OrderedDict((k, sd[k].double().sum().item()) for k in sd if k.endswith('running_var'))    # after one train forward
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'derived_net.npz'), logits_train=la.numpy(),
                        logits_eval=lb.numpy(), lat=lat, probe_keys=np.array(list(probes)),
                        probe_sums=np.array(list(probes.values())), n_state=len(sd),
                        running_keys=np.array(list(running)), running_sums=np.array(list(running.values())))
    print('wrote derived_net.npz: lat %.5f, %d state entries' % (lat, len(sd)))


if __name__ == '__main__':
    main()
