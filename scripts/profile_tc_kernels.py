"""Run the tensor-core kernels once at BASELINE sizes (for an ncu capture): decode + the projection GEMM shapes."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from selfocc_b200 import ops, synth
from selfocc_b200.mapping import GridMeterMapping
dev = torch.device('cuda:0')
m = GridMeterMapping(**synth.NUSC_MAPPING)
desc = m.volume_desc(0)
planes = [p.to(dev) for p in synth.random_planes(m, 96)]
w1, b1, w2, b2 = (t.to(dev) for t in synth.random_mlp(96, 1))
for _ in range(3):
    ops.tpv_decode(*planes, w1, b1, w2, b2, desc)
for M, N in ((153000, 288), (81983, 648), (81983, 96), (7967, 3456)):
    x = torch.randn(M, 96, device=dev); w = 0.1 * torch.randn(N, 96, device=dev); b = torch.randn(N, device=dev)
    hi, lo = ops.split_tf32(w)
    for _ in range(3):
        ops.linear_3xtf32(x, hi, lo, b)
torch.cuda.synchronize()
print('done')
