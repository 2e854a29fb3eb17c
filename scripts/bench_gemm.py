"""Device timing of so_linear_3xtf32 vs cuBLAS fp32 for the projection shapes of one encoder layer (cfg 3)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from selfocc_b200 import ops

dev = torch.device('cuda:0')
shapes = [('value_proj img', 153000, 96, 96), ('self offsets', 81983, 432, 96), ('self weights', 81983, 216, 96),
          ('out_proj', 81983, 96, 96), ('hw offsets', 66049, 384, 96), ('zh offsets', 7967, 2304, 96), ('zh weights', 7967, 1152, 96),
          ('ffn1', 81983, 192, 96), ('ffn2', 81983, 96, 192)]
res = []
for name, M, N, K in shapes:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.1; b = torch.randn(N, device=dev)
    hi, lo = ops.split_tf32(w)
    def t(fn, it=20):
        for _ in range(3): fn()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(it): fn()
        e.record(); torch.cuda.synchronize()
        return a.elapsed_time(e) / it
    t_mine = t(lambda: ops.linear_3xtf32(x, hi, lo, b))
    t_cublas = t(lambda: torch.nn.functional.linear(x, w, b))
    gb = (M * K + M * N + 2 * N * K) * 4 / 1e9
    res.append(dict(name=name, M=M, N=N, K=K, ms_3xtf32=round(t_mine, 4), ms_cublas=round(t_cublas, 4),
                    gbs=round(gb / (t_mine * 1e-3), 1), tflops_eff=round(2 * M * N * K / (t_mine * 1e-3) / 1e12, 2)))
    print(res[-1])
print(json.dumps(res))
