"""CPU: analytic self-checks of the (unpinned) parts of the oracle (SURVEY.md 8c iii)."""
import math
import torch
from oracle.mapping import GridMeterMappingRef
from oracle import render as orender, lifting as ol, rays as orays
from selfocc_b200 import synth


def _map(hw=10, d=6):
    margs, aabb = synth.small_mapping(hw, d)
    return GridMeterMappingRef(**margs), aabb


def test_manual_field_query_equals_grid_sample():
    m, aabb = _map()
    g = torch.Generator().manual_seed(0)
    vol = torch.randn(4, m.size_h, m.size_w, m.size_d, generator=g, dtype=torch.float64)
    x = (torch.rand(500, 3, generator=g, dtype=torch.float64) - 0.5) * torch.tensor([30., 30., 7.]) + torch.tensor([0., 0., 0.5])
    h1, g1 = orender.field_query_ref(vol, m, x)
    h2, g2 = orender.field_query_manual(vol, m, x)
    assert torch.allclose(h1, h2, atol=1e-10) and torch.allclose(g1, g2, atol=1e-9)


def test_planar_sdf_renders_plane_depth():
    """sdf = z - z0 (exactly representable by trilinear interpolation): as inv_s grows the rendered depth tends to
    the camera-z depth of the plane hit, and normals to +z."""
    m, aabb = _map(16, 8)
    H, W, Z = m.size_h, m.size_w, m.size_d
    g = torch.stack(torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64),
                                   torch.arange(Z, dtype=torch.float64), indexing='ij'), -1)
    z0 = -1.0
    vol = (m.grid2meter(g)[..., 2] - z0)[None]
    l2i, i2l = synth.camera_rig((0.,), f=126.6, cx=80., cy=45., height=0.5, radius=0.0)
    pix = torch.tensor([[80., 70.], [60., 80.], [100., 85.]])
    origin, direction = orays.img2lidar_rays(torch.tensor(i2l, dtype=torch.float64)[None].float(), pix)
    out = orender.head_render_ref(vol, m, origin.double(), direction.double(), aabb, inv_s=400.0, S=2048)
    # analytic: ray o + t*dir (dir has camera depth 1) hits z = z0 at t = (z0 - o_z) / dir_z
    t_hit = (z0 - origin[0, 0, 2].double()) / direction[0, 0, :, 2].double()
    # the +1e-5 in the alpha formula leaks ~1e-5 of weight per sample in front of the surface: a known ~0.5% pull
    assert torch.allclose(out['depth'][0, 0], t_hit, rtol=1e-2)
    assert (out['depth'][0, 0] < t_hit).all()
    assert torch.allclose(out['vis_normal'][0, 0], torch.tensor([0.5, 0.5, 1.0], dtype=torch.float64).expand(3, 3), atol=2e-2)
    assert torch.allclose(out['acc'][0, 0], torch.ones(3, dtype=torch.float64), atol=2e-3)


def test_msda_integer_centres_and_uniform_weights():
    shapes = [(4, 6)]
    value = torch.arange(24 * 2 * 4, dtype=torch.float64).reshape(1, 24, 2, 4)
    loc = torch.tensor([(2 + 0.5) / 6, (1 + 0.5) / 4], dtype=torch.float64).reshape(1, 1, 1, 1, 1, 2).repeat(1, 1, 2, 1, 1, 1)
    out = ol.msda_ref(value, shapes, loc, torch.ones(1, 1, 2, 1, 1, dtype=torch.float64))
    assert torch.equal(out.view(2, 4), value[0, 1 * 6 + 2])
    # a location on the border between 4 pixels averages them
    loc2 = torch.tensor([3.0 / 6, 2.0 / 4], dtype=torch.float64).reshape(1, 1, 1, 1, 1, 2).repeat(1, 1, 2, 1, 1, 1)
    out2 = ol.msda_ref(value, shapes, loc2, torch.ones(1, 1, 2, 1, 1, dtype=torch.float64))
    exp = (value[0, 1 * 6 + 2] + value[0, 1 * 6 + 3] + value[0, 2 * 6 + 2] + value[0, 2 * 6 + 3]) / 4
    assert torch.allclose(out2.view(2, 4), exp)


def test_point_sampling_identity_camera():
    """lidar2img = identity: uv = (x/z/w, y/z/h), mask = in front & inside the unit square."""
    ref = torch.tensor([[[0.5, 0.25, 1.0], [2.0, 1.0, 4.0], [1.0, 1.0, -1.0], [3.0, 0.1, 1.0]]]).reshape(1, 1, 4, 3)
    uv, mask = ol.point_sampling_ref(ref, torch.eye(4)[None, None], (1.0, 2.0))
    assert torch.allclose(uv[0, 0, :, 0], torch.tensor([[0.25, 0.25], [0.25, 0.25], [1.0 / 1e-5 / 2, 1.0 / 1e-5], [1.5, 0.1]]))
    assert mask[0, 0, :, 0].tolist() == [True, True, False, False]


def test_max_depth_first_maximum():
    w = torch.tensor([[0.1, 0.4, 0.4, 0.1]])
    ts = torch.tensor([[1., 2., 3., 4.]])
    d = torch.full((1, 4), 0.5)
    md, idx = orender.max_depth_ref(w, ts, d)
    assert idx.item() == 1 and md.item() == 2.0
