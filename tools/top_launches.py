import sys, os, collections, torch
sys.path.insert(0, '/root/repo')
from tests import golden_inputs as gi
from tfnas_b200 import _lib, config, model_search
from tfnas_b200.model_search import Network
from tfnas_b200.parallel import GradSync, SearchParallel
from tfnas_b200.search_loop import alpha_step, make_optimizers, w_step
from tfnas_b200.step import FusedCrossEntropy
dev = torch.device('cuda', 0)
torch.manual_seed(2); model_search.seed_noise(2)
net = Network(100, config.get_mc_num_dddict(config.mc_mask_dddict), gi.load_lut()); net.set_temperature(5.0)
model = SearchParallel(net).to(dev).train(); crit = FusedCrossEntropy().to(dev)
opt_w, opt_a = make_optimizers(net); sync = GradSync()
g = torch.Generator().manual_seed(2)
pool = [(torch.randn(128, 3, 224, 224, generator=g).to(dev), torch.randint(0, 100, (128,), generator=g).to(dev)) for _ in range(3)]
for i in range(3):
    w_step(model, *pool[i % 3], crit, opt_w, 5.0, sync, bisample=True)
torch.cuda.synchronize()
_lib.prof_enable(True)
w_step(model, *pool[0], crit, opt_w, 5.0, sync, bisample=True)
torch.cuda.synchronize()
tl = _lib.prof_timeline()
_lib.prof_enable(False)
for name in ('wgrad_w3', 'wgrad_w1', 'dx', 'dc', 'b4mm', 'um_prep_w', 'se_bwd', 'w1fin'):
    d = sorted([(e - s) for n, st, s, e in tl if n == name], reverse=True)
    print('%-10s n=%3d total %.2f ms top: %s' % (name, len(d), sum(d), ' '.join('%.0f' % (1e3 * v) for v in d[:12])))
