"""Generate golden vectors from the importable pure-torch pieces of the reference.

Run ONCE in the build container (``/root/reference`` is absent on the GPU box):
    python tests/golden/make_golden.py
Writes ``tests/golden/reference_golden.npz``.  Only reference modules that import without
mmcv/mmengine/nerfstudio are used; they are loaded by file path and
executed unmodified.  Inputs are seeded and stored beside the outputs.
"""
import importlib.util
import os
import sys
import numpy as np
import torch

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))


def load(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    out = {}
    mp = load('model/encoder/bevformer/mappings.py', 'ref_mappings')
    # (i) the reference's own __main__ smoke inputs (mappings.py:300-329)
    m = mp.GridMeterMapping(nonlinear_mode='linear', h_size=[2, 2], h_range=[2, 4], h_half=False, w_size=[2, 2],
                            w_range=[2, 4], w_half=False, d_size=[2, 2], d_range=[-1., 1., 5.])
    grid = torch.tensor([[4, 0, 0], [0, 4, 1], [4, 4, 2], [5, 6, 4], [1, 0, 1.5], [7.5, 8, 2.5]])
    meter = torch.tensor([[-6., 0., -1.], [0., -6., 0.], [0., 0., 1.], [2., 1., 5.], [-6., -3.6667, 0.5],
                          [6., 4.8333, 1.8333]])
    out['map_small_grid'] = grid.numpy()
    out['map_small_g2m'] = m.grid2meter(grid).numpy()
    out['map_small_meter'] = meter.numpy()
    out['map_small_m2g'] = m.meter2grid(meter).numpy()
    # (ii) the nuScenes depth mapping (config/nuscenes/nuscenes_depth.py:188-198) on random points
    args = dict(nonlinear_mode='linear', h_size=[128, 0], h_range=[51.2, 0], h_half=False, w_size=[128, 0],
                w_range=[51.2, 0], w_half=False, d_size=[30, 0], d_range=[-4.0, 5.0, 5.0])
    m2 = mp.GridMeterMapping(**args)
    pts = (torch.rand(64, 3) - 0.5) * torch.tensor([120., 120., 12.])
    out['map_nus_meter'] = pts.numpy()
    out['map_nus_m2g'] = m2.meter2grid(pts).numpy()
    out['map_nus_m2g_norm'] = m2.meter2grid(pts, True).numpy()
    g = torch.rand(64, 3) * torch.tensor([256., 256., 30.])
    out['map_nus_grid'] = g.numpy()
    out['map_nus_g2m'] = m2.grid2meter(g).numpy()
    # (iii) a mapping with outer rings + half axes
    args3 = dict(nonlinear_mode='linear', h_size=[8, 4], h_range=[10., 20.], h_half=True, w_size=[6, 2],
                 w_range=[12., 8.], w_half=False, d_size=[4, 2], d_range=[-2.0, 2.0, 6.0])
    m3 = mp.GridMeterMapping(**args3)
    pts3 = (torch.rand(64, 3) - 0.3) * torch.tensor([40., 30., 10.])
    out['map_ring_meter'] = pts3.numpy()
    out['map_ring_m2g'] = m3.meter2grid(pts3, True).numpy()
    g3 = torch.rand(64, 3) * torch.tensor([12., 16., 6.])
    out['map_ring_grid'] = g3.numpy()
    out['map_ring_g2m'] = m3.grid2meter(g3).numpy()

    # cross-view reference points (tpvformer/utils.py:5-71)
    tu = load('model/encoder/tpvformer/utils.py', 'ref_tpv_utils')
    out['cvref_5_7_3_p4'] = tu.get_cross_view_ref_points(5, 7, 3, [4, 4, 4]).numpy()
    out['cvref_4_3_6_p3'] = tu.get_cross_view_ref_points(4, 3, 6, [3, 3, 3]).numpy()

    # point_sampling (bevformer/utils.py:116-206)
    bu = load('model/encoder/bevformer/utils.py', 'ref_bev_utils')
    ref3d = (torch.rand(1, 5, 40, 3) - 0.5) * torch.tensor([80., 80., 8.])
    K = np.array([[1266., 0, 800, 0], [0, 1266., 450, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    l2i = []
    for yaw in (0., -55., 110.):
        a = np.deg2rad(yaw)
        c2l = np.eye(4)
        # camera axes (x right, y down, z forward) expressed in the lidar frame (x right, y fwd, z up)
        fwd = np.array([np.sin(-a), np.cos(-a), 0.])
        right = np.array([np.cos(-a), -np.sin(-a), 0.])
        down = np.array([0., 0., -1.])
        c2l[:3, 0], c2l[:3, 1], c2l[:3, 2], c2l[:3, 3] = right, down, fwd, [0.3, 0.5, 1.5]
        l2i.append(K @ np.linalg.inv(c2l))
    l2i = np.stack(l2i)[None]
    metas = [dict(lidar2img=list(l2i[0]), img_shape=(900, 1600))]
    rc, mk = bu.point_sampling(ref3d, metas)
    out['ps_ref3d'] = ref3d.numpy()
    out['ps_lidar2img'] = l2i
    out['ps_uv'] = rc.numpy()
    out['ps_mask'] = mk.numpy()

    # ray sampler (ray_sampler.py) fixed + cellular
    rs = load('model/head/nerfacc_head/ray_sampler.py', 'ref_ray_sampler')
    out['rays_fixed_6x10_90x160'] = rs.RaySampler('fixed', [6, 10], [90, 160])().numpy()
    out['rays_fixed_450x800'] = rs.RaySampler('fixed', [450, 800], [900, 1600])()[::997].numpy()
    np.random.seed(123)
    u4 = np.random.uniform(size=4)
    np.random.seed(123)
    out['rays_cell_u4'] = u4
    out['rays_cell_6x10_90x160'] = rs.RaySampler('cellular', [6, 10], [90, 160], ray_upper_crop=8)().numpy()

    # SH bases (sh_render.py)
    sh = load('model/head/utils/sh_render.py', 'ref_sh')
    dirs = torch.nn.functional.normalize(torch.randn(16, 3), dim=-1)
    feats = torch.randn(16, 3)
    out['sh_dirs'] = dirs.numpy()
    out['sh_feats'] = feats.numpy()
    out['sh_deg0_relu'] = sh.SHRender(None, dirs, feats, 0, 'relu').numpy()
    out['sh_deg0_sigmoid'] = sh.SHRender(None, dirs, feats, 0, 'sigmoid').numpy()

    # depth metric arithmetic (utils/metric_util.py:247-279; module import needs mmengine, so the
    # function source is exec'd on its own)
    src = open(os.path.join(REF, 'utils/metric_util.py')).read().split('\n')
    s = next(i for i, l in enumerate(src) if l.startswith('def cal_depth_metric'))
    e = next(i for i in range(s + 1, len(src)) if src[i].startswith('class '))
    ns = {'torch': torch}
    exec('\n'.join(src[s:e]), ns)
    gt = torch.rand(200) * 60 + 1
    pred = gt * (1 + 0.2 * torch.randn(200))
    md = ns['cal_depth_metric'](pred, gt)
    out['dm_gt'], out['dm_pred'] = gt.numpy(), pred.numpy()
    out['dm_vals'] = np.array([float(md[k]) for k in ('abs_rel', 'sq_rel', 'rmse', 'rmse_log', 'a1', 'a2', 'a3')])

    np.savez_compressed(os.path.join(HERE, 'reference_golden.npz'), **out)
    print('wrote', len(out), 'arrays')




def dump_reference_model_configs():
    """The lifter / encoder / head config dicts of every shipped TPV experiment config, as the reference's own config
    files define them (plain `exec` of config/<set>/<name>.py: they are self-contained python apart from `_base_`).
    Written to tests/golden/reference_model_cfgs.json; tests/test_modules_cpu.py builds the B200 modules from them."""
    import glob
    import json
    out = {}
    for f in sorted(glob.glob(os.path.join(REF, 'config', '*', '*.py'))):
        if '_base_' in f:
            continue
        ns = {}
        exec(compile(open(f).read(), f, 'exec'), ns)
        m = ns['model']
        if m['encoder']['type'] != 'TPVFormerEncoder':
            continue                      # nuscenes_occ_bev.py: the BEV-only variant is out of scope
        out[os.path.relpath(f, os.path.join(REF, 'config'))] = {k: m[k] for k in ('lifter', 'encoder', 'head')}
    json.dump(out, open(os.path.join(HERE, 'reference_model_cfgs.json'), 'w'), indent=1, sort_keys=True)
    print('wrote model configs of', sorted(out))


if __name__ == '__main__':
    main()
    dump_reference_model_configs()
