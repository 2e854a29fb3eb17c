#!/usr/bin/env python
"""Render-only timing at BASELINE configs[2] size (6 x 900 x 1600 rays x 256 samples, 257 x 257 x 31 volume): plain vs packed
kernels, depth-only and colour (color_dims = 3), on the analytic scene (rays terminate) and on a decoded random-plane scene
(free space everywhere: no ray terminates, the early exit never fires).  CUDA events on the launch stream, L2 flushed
before every launch.  Prints one JSON line per variant."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from selfocc_b200 import ops, synth, _lib  # noqa: E402
from selfocc_b200.mapping import GridMeterMapping  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    iters = int(os.environ.get('ITERS', 5))
    m = GridMeterMapping(**synth.NUSC_MAPPING)
    _, i2l = synth.camera_rig()
    cams = torch.tensor(i2l, dtype=torch.float32, device=dev)
    ny, nx, S = 900, 1600, 256
    rd = ops.make_ray_desc(6, grid=(ny, nx, 1.0, 0.0, 1.0, 0.0))
    pr = ops.make_render_params(synth.NUSC_RANGE, S, 20.0855)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    H, W, Z = m.size_h, m.size_w, m.size_d
    for scene in ('analytic', 'decoded'):
        for cd in (0, 3):
            desc = m.volume_desc(cd)
            if scene == 'analytic':
                sdf = synth.analytic_sdf_volume(m, noise=0.02)
                vs = synth.pack_sdf_volume(sdf, desc.zpitch).to(dev)
                vf = synth.pack_feat_volume(torch.randn(cd, H, W, Z), desc.feat_pitch).to(dev) if cd else None
            else:
                planes = synth.random_planes(m, 96, scale=1.0, seed=5)
                w1, b1, w2, b2 = synth.random_mlp(96, 1 + cd, seed=2)
                vs, vf = ops.tpv_decode(*[p.to(dev) for p in planes], w1.to(dev), b1.to(dev), w2.to(dev), b2.to(dev), desc)
            want = ['depth', 'max_depth', 'acc', 'normal_vis'] + (['rgb'] if cd else [])
            out = {k: torch.empty((6 * ny * nx,) + ((3,) if k in ('normal_vis', 'rgb') else ()), device=dev) for k in want}
            for kind in ('plain', 'packed'):
                pack = ops.render_pack(vs, vf, desc) if kind == 'packed' else None
                ts = []
                for it in range(iters + 2):
                    flush.zero_()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    r = ops.render_infer(vs, vf, desc, cams, rd, pr, want=want, out=out, pack=pack)
                    b.record()
                    torch.cuda.synchronize()
                    if it >= 2:
                        ts.append(a.elapsed_time(b))
                ts.sort()
                print(json.dumps({'scene': scene, 'color_dims': cd, 'kernel': kind, 'ms_median': ts[len(ts) // 2], 'ms_min': ts[0],
                                  'rays_per_s': 6 * ny * nx / (ts[len(ts) // 2] * 1e-3), 'acc_mean': float(r['acc'].mean()),
                                  'terminated_frac': float((r['acc'] > 0.999).float().mean())}), flush=True)


if __name__ == '__main__':
    main()
