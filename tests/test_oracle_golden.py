"""Pin the oracle against golden vectors produced by the reference's own importable code
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import torch
from oracle.mapping import GridMeterMappingRef
from oracle import lifting, rays, render, metric

NUS = dict(nonlinear_mode='linear', h_size=[128, 0], h_range=[51.2, 0], h_half=False, w_size=[128, 0],
           w_range=[51.2, 0], w_half=False, d_size=[30, 0], d_range=[-4.0, 5.0, 5.0])
RING = dict(nonlinear_mode='linear', h_size=[8, 4], h_range=[10., 20.], h_half=True, w_size=[6, 2],
            w_range=[12., 8.], w_half=False, d_size=[4, 2], d_range=[-2.0, 2.0, 6.0])
SMALL = dict(nonlinear_mode='linear', h_size=[2, 2], h_range=[2, 4], h_half=False, w_size=[2, 2],
             w_range=[2, 4], w_half=False, d_size=[2, 2], d_range=[-1., 1., 5.])
T = lambda a: torch.from_numpy(np.asarray(a))


def test_mapping_reference_smoke_vectors(golden):
    m = GridMeterMappingRef(**SMALL)
    assert torch.equal(m.grid2meter(T(golden['map_small_grid'])), T(golden['map_small_g2m']))
    assert torch.equal(m.meter2grid(T(golden['map_small_meter'])), T(golden['map_small_m2g']))
    # SURVEY section 4 known answers (rows 0-3 round-trip exactly)
    exp = torch.tensor([[-6., 0, -1], [0, -6, 0], [0, 0, 1], [2, 1, 5], [-6, -4, 0.5], [6, 5, 2]])
    assert torch.allclose(m.grid2meter(T(golden['map_small_grid'])), exp)


def test_mapping_nuscenes_and_ring(golden):
    m = GridMeterMappingRef(**NUS)
    assert (m.size_h, m.size_w, m.size_d) == (257, 257, 31)
    assert torch.equal(m.meter2grid(T(golden['map_nus_meter'])), T(golden['map_nus_m2g']))
    assert torch.equal(m.meter2grid(T(golden['map_nus_meter']), True), T(golden['map_nus_m2g_norm']))
    assert torch.equal(m.grid2meter(T(golden['map_nus_grid'])), T(golden['map_nus_g2m']))
    r = GridMeterMappingRef(**RING)
    assert torch.equal(r.meter2grid(T(golden['map_ring_meter']), True), T(golden['map_ring_m2g']))
    assert torch.equal(r.grid2meter(T(golden['map_ring_grid'])), T(golden['map_ring_g2m']))


def test_cross_view_ref_points(golden):
    assert torch.equal(lifting.cross_view_ref_points(5, 7, 3, [4, 4, 4]), T(golden['cvref_5_7_3_p4']))
    assert torch.equal(lifting.cross_view_ref_points(4, 3, 6, [3, 3, 3]), T(golden['cvref_4_3_6_p3']))


def test_point_sampling(golden):
    uv, mk = lifting.point_sampling_ref(T(golden['ps_ref3d']), T(golden['ps_lidar2img']), (900, 1600))
    assert torch.equal(mk, T(golden['ps_mask']))
    assert torch.equal(uv, T(golden['ps_uv']))
    assert 0 < mk.sum() < mk.numel()


def test_ray_sampler(golden):
    assert torch.equal(rays.fixed_ray_grid([6, 10], [90, 160]), T(golden['rays_fixed_6x10_90x160']))
    assert torch.equal(rays.fixed_ray_grid([450, 800], [900, 1600])[::997], T(golden['rays_fixed_450x800']))
    cell = rays.cellular_ray_grid([6, 10], [90, 160], golden['rays_cell_u4'], ray_upper_crop=8)
    assert torch.equal(cell, T(golden['rays_cell_6x10_90x160']))


def test_sh_colour(golden):
    f = T(golden['sh_feats'])
    assert torch.allclose(torch.relu(f * render.C0 + 0.5), T(golden['sh_deg0_relu']), atol=0, rtol=0)
    assert torch.allclose(torch.sigmoid(f * render.C0), T(golden['sh_deg0_sigmoid']), atol=0, rtol=0)


def test_depth_metric(golden):
    md = metric.cal_depth_metric_ref(T(golden['dm_pred']), T(golden['dm_gt']))
    got = np.array([float(md[k]) for k in ('abs_rel', 'sq_rel', 'rmse', 'rmse_log', 'a1', 'a2', 'a3')])
    assert np.array_equal(got, golden['dm_vals'])


def test_ref3d_tables_shapes():
    m = GridMeterMappingRef(**NUS)
    hw, zh, wz = lifting.ref_3d_tables(m, [48, 48, 8])
    assert hw.shape == (8, 257 * 257, 3) and zh.shape == (48, 31 * 257, 3) and wz.shape == (48, 257 * 31, 3)
    # pillar of the hw plane spans z in [-4, 5]
    assert torch.allclose(hw[:, 0, 2], torch.linspace(-4, 5, 8))
