import sys, torch
sys.path.insert(0, '/root/repo')
from tfnas_b200.ops import ArenaPool, HeadFn
g = torch.Generator().manual_seed(0)
x = torch.randn(128, 320, 7, 7, generator=g).cuda().requires_grad_(True)
fm = (torch.randn(1280, 320, 1, 1, generator=g) / 18).cuda().requires_grad_(True)
fw = (torch.randn(100, 1280, generator=g) / 36).cuda().requires_grad_(True)
fb = torch.zeros(100).cuda().requires_grad_(True)
G = torch.randn(128, 100, generator=g).cuda()
pool = ArenaPool()
for i in range(3):
    out = HeadFn.apply(x, pool, fm, fw, fb)
    (out * G).sum().backward()
torch.cuda.synchronize()
