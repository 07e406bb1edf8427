import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import helpers as H
from tfnas_b200 import _lib

def run(M, N, K, variant, A, B):
    lib = _lib.load()
    C = torch.full((N, M), float('nan'), device='cuda')
    npad = (N + 15) // 16 * 16
    nbytes = ((K + 31) // 32) * 2 * npad * 128
    ws = torch.zeros(nbytes, dtype=torch.uint8, device='cuda')
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.tfnas_umma_selftest(M, N, K, vp(A), vp(B), vp(C), vp(ws), nbytes, variant, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return C

M, N, K = 128, 32, 32
ones = lambda *s: torch.ones(*s, device='cuda')
C = run(M, N, K, 2, ones(K, M), ones(N, K))
print('variant2 (tcgen05.st pattern): C[0,:4]', C[0, :4].tolist(), 'C[3,:4]', C[3,:4].tolist(), 'C[31,124:]', C[31,124:].tolist())
C = run(M, N, K, 0, ones(K, M), ones(N, K))
print('ones x ones K=32: unique', torch.unique(C).tolist()[:10])
C = run(M, N, 8, 0, ones(8, M), ones(N, 8))
print('ones x ones K=8: unique', torch.unique(C).tolist()[:10])
g = torch.Generator().manual_seed(0)
# A = ones, B random: C[n][p] = sum_k B[n][k] -> checks B layout independent of A layout
B = torch.randn(N, K, generator=g).cuda()
C = run(M, N, K, 0, ones(K, M), B)
print('A=1, B rand: err', H.rel_l2(C, (B.double().sum(1, keepdim=True)).expand(N, M)))
A = torch.randn(K, M, generator=g).cuda()
C = run(M, N, K, 0, A, ones(N, K))
print('A rand, B=1: err', H.rel_l2(C, A.double().sum(0, keepdim=True).expand(N, M)))
C = run(M, N, K, 0, A, B)
print('A rand, B rand: err', H.rel_l2(C, B.double() @ A.double()))
