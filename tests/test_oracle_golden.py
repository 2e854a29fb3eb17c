"""Pin the oracle against golden vectors produced by the reference's own importable code
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
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


# ---- B5 / B7 pinned to the reference's in-tree statement (BEVNeRF, bev_nerf.py:62-175) -----------------------------
@pytest.mark.parametrize('case', ['a', 'b'])
def test_decode_and_field_query_match_bevnerf(case):
    import json
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden_bevnerf.npz'))
    margs = json.loads(str(g[case + '_margs']))
    C, cd, sd, H, W, Z = (int(v) for v in g[case + '_dims'])
    m = GridMeterMappingRef(**margs)
    assert (m.size_h, m.size_w, m.size_d) == (H, W, Z)
    planes = [T(g['%s_plane%d' % (case, i)])[0] for i in range(3)]
    w1, b1, w2, b2 = (T(g[case + k]) for k in ('_w1', '_b1', '_w2', '_b2'))
    ref_vol = T(g[case + '_vol'])[0]                                          # [Cf, H, W, Z] from pre_compute_density_color
    # B5: same fp32 arithmetic (broadcast sum, Softplus, Linear); only the GEMM blocking differs -> a few ulp
    vol = render.tpv_decode_ref(planes[0], planes[1], planes[2], (H, W, Z), w1, b1, w2, b2)
    assert vol.shape == ref_vol.shape
    assert torch.allclose(vol, ref_vol, rtol=0, atol=2e-6)
    vol64 = render.tpv_decode_ref(*[p.double() for p in planes], (H, W, Z), w1.double(), b1.double(), w2.double(), b2.double())
    assert (vol64 - ref_vol.double()).abs().max() < 5e-6
    # B7: the oracle's field query on the REFERENCE's volume must reproduce BEVNeRF.forward / forward_geo / query_density
    x = T(g[case + '_x'])
    h, _ = render.field_query_ref(ref_vol, m, x, with_grad=False)
    assert torch.allclose(torch.nn.functional.softplus(h[:, :1]), T(g[case + '_sigma']), rtol=0, atol=1e-6)
    assert torch.equal(T(g[case + '_sigma']), T(g[case + '_sigma_geo'])) and torch.equal(T(g[case + '_sigma']), T(g[case + '_dens']))
    if cd:
        rgb = torch.relu(h[:, 1:4] * render.C0 + 0.5)                         # sh_render.py:84-94, degree 0
        assert torch.allclose(rgb, T(g[case + '_rgb']), rtol=0, atol=1e-6)
    if sd:
        assert torch.allclose(torch.softmax(h[:, 1 + cd:], -1), T(g[case + '_sems']), rtol=0, atol=1e-6)
    # the explicit 8-corner form used by the training-parity tests states the same function (incl. zeros padding outside)
    hm, _ = render.field_query_manual(ref_vol.double(), m, x.double())
    assert (hm - h.double()).abs().max() < 1e-5
    outside = ((x[:, 0].abs() > margs['w_range'][0]) | (x[:, 2] > margs['d_range'][1] + 0.31)).sum()
    assert outside > 10                                                       # the padding branch is exercised
