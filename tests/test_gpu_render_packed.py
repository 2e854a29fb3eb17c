"""GPU parity: the packed-volume render kernels (so_render_pack + so_render_infer_packed) vs the fp64 oracle and vs the
plain kernel, on small scenes the oracle finishes in seconds (full-size gates: test_gpu_fullsize.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from selfocc_b200 import synth
from test_gpu_render import _dev, _scene, _cams, _M64


def _render_both(n_feat, aabb_scale=1.0, S=64, inv_s=20.0, ny=18, nx=32, want_extra=()):
    dev = _dev()
    from oracle import rays as orays
    from selfocc_b200 import ops
    m, mref, aabb, sdf, feat, vol_ref = _scene(n_feat=n_feat)
    if aabb_scale != 1.0:          # a ROI larger than the volume: samples outside take the zero-padding path
        aabb = [a * aabb_scale for a in aabb]
    _, i2l = _cams(3)
    ih, iw = 90, 160
    pix = orays.fixed_ray_grid([ny, nx], [ih, iw])
    origin, direction = orays.img2lidar_rays(i2l[None], pix)
    desc = m.volume_desc(n_feat)
    vs = synth.pack_sdf_volume(sdf, desc.zpitch).to(dev)
    vf = synth.pack_feat_volume(feat, desc.feat_pitch).to(dev) if n_feat else None
    pack = ops.render_pack(vs, vf, desc)
    assert pack is not None and pack.numel() == (2 * desc.H * desc.W * desc.zpitch if n_feat == 0 else 4 * desc.H * desc.W * desc.Z)
    rd = ops.make_ray_desc(3, grid=(ny, nx, iw / nx, 0.0, ih / ny, 0.0))
    pr = ops.make_render_params(aabb, S, inv_s, bkgd='white')
    want = ['depth', 'max_depth', 'max_idx', 'acc', 'normal_vis'] + (['rgb'] if n_feat else [])
    plain = ops.render_infer(vs, vf, desc, i2l.to(dev), rd, pr, want=want)
    packed = ops.render_infer(vs, vf, desc, i2l.to(dev), rd, pr, want=want, pack=pack)
    probed = ops.render_infer(vs, vf, desc, i2l.to(dev), rd, pr, want=want, pack=pack, probe_grid=True)
    cpu = lambda d: {k: v.cpu() for k, v in d.items()}
    return cpu(plain), cpu(packed), cpu(probed), dict(vol=vol_ref.double(), mref=_M64(mref), origin=origin, direction=direction,
                                                       aabb=aabb, inv_s=inv_s, S=S, n_feat=n_feat)


@pytest.mark.parametrize('n_feat,aabb_scale', [(0, 1.0), (3, 1.0), (0, 1.3), (3, 1.3)])
def test_packed_render_matches_fp64_oracle(n_feat, aabb_scale):
    from oracle.parity import render_parity
    plain, packed, probed, ctx = _render_both(n_feat, aabb_scale)
    rep = render_parity(probed, ctx['vol'], ctx['mref'], ctx['origin'], ctx['direction'], ctx['aabb'], ctx['inv_s'], ctx['S'],
                        color_dims=3 if n_feat else 0, bkgd='white')
    print(rep)
    assert rep['ok'], rep
    # the probe build (no early exit) and the production build agree to rounding, and exactly on the index
    assert torch.allclose(packed['depth'], probed['depth'], rtol=2e-6, atol=0)
    assert torch.equal(packed['max_idx'], probed['max_idx'])
    if aabb_scale > 1.0:    # the padded path was really taken: some sample coordinates lie outside the volume
        g = probed['grid']
        assert (g.min() < 0) or (g[..., 0].max() > ctx['vol'].shape[1] - 1)


@pytest.mark.parametrize('n_feat', [0, 3])
def test_packed_and_plain_kernels_agree(n_feat):
    plain, packed, _, ctx = _render_both(n_feat)
    rel = ((packed['depth'] - plain['depth']).abs() / plain['depth'].abs().clamp_min(1e-6)).max().item()
    assert rel < 2e-5, rel
    assert torch.allclose(packed['acc'], plain['acc'], atol=5e-6)
    assert torch.allclose(packed['normal_vis'], plain['normal_vis'], atol=2e-5)
    assert (packed['max_idx'] == plain['max_idx']).float().mean() > 0.995
    same = packed['max_idx'] == plain['max_idx']
    assert torch.allclose(packed['max_depth'][same], plain['max_depth'][same], rtol=1e-6)
    if n_feat:
        assert torch.allclose(packed['rgb'], plain['rgb'], atol=2e-5)


def test_packed_early_exit_is_invisible_on_a_dense_scene():
    """A scene whose rays all terminate (solid half-space right in front of the cameras): warps stop marching early; outputs
    must equal the probe build (which never exits early) to < 1e-7 relative and exactly on the index."""
    dev = _dev()
    from selfocc_b200 import ops
    from selfocc_b200.mapping import GridMeterMapping
    margs, aabb = synth.small_mapping(16, 8)
    m = GridMeterMapping(**margs)
    sdf = synth.analytic_sdf_volume(m, ground_z=10.0, spheres=(), boxes=())          # sdf = z - 10 < 0 everywhere: inside
    desc = m.volume_desc(0)
    vs = synth.pack_sdf_volume(sdf, desc.zpitch).to(dev)
    pack = ops.render_pack(vs, None, desc)
    _, i2l = _cams(3)
    rd = ops.make_ray_desc(3, grid=(18, 32, 5.0, 0.0, 5.0, 0.0))
    pr = ops.make_render_params(aabb, 256, 50.0)
    want = ('depth', 'max_idx', 'acc', 'normal_vis')
    a = ops.render_infer(vs, None, desc, i2l.to(dev), rd, pr, want=want, pack=pack)
    b = ops.render_infer(vs, None, desc, i2l.to(dev), rd, pr, want=want, pack=pack, probe_grid=True)
    assert (a['acc'] > 0.999).all()
    assert torch.allclose(a['depth'], b['depth'], rtol=1e-7, atol=0) and torch.equal(a['max_idx'], b['max_idx'])
    assert torch.allclose(a['acc'], b['acc'], atol=1e-7) and torch.allclose(a['normal_vis'], b['normal_vis'], atol=1e-7)


def test_packed_entry_point_routes_unsupported_configs_to_the_plain_kernel():
    dev = _dev()
    from selfocc_b200 import ops
    m, _, aabb, sdf, feat, _ = _scene(n_feat=3)
    desc = m.volume_desc(3)
    vs = synth.pack_sdf_volume(sdf, desc.zpitch).to(dev)
    vf = synth.pack_feat_volume(feat, desc.feat_pitch).to(dev)
    pack = ops.render_pack(vs, vf, desc)
    _, i2l = _cams(3)
    rd = ops.make_ray_desc(3, grid=(6, 8, 20.0, 0.0, 15.0, 0.0))
    for pr in (ops.make_render_params(aabb, 48, 20.0),                               # S not a power of two
               ops.make_render_params(aabb, 64, 20.0, anchor_mid=False),             # start anchor
               ops.make_render_params(aabb, 64, 20.0, sh_act='sigmoid')):            # sigmoid colour
        a = ops.render_infer(vs, vf, desc, i2l.to(dev), rd, pr, want=('depth', 'rgb', 'max_idx'))
        b = ops.render_infer(vs, vf, desc, i2l.to(dev), rd, pr, want=('depth', 'rgb', 'max_idx'), pack=pack)
        assert torch.equal(a['depth'], b['depth']) and torch.equal(a['rgb'], b['rgb']) and torch.equal(a['max_idx'], b['max_idx'])
    # 8 decoded channels (semantics): no packed form
    desc8 = m.volume_desc(8)
    assert ops.render_pack(vs, synth.pack_feat_volume(torch.randn(8, *sdf.shape), desc8.feat_pitch).to(dev), desc8) is None
