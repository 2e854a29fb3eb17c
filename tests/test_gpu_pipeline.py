"""GPU parity of the whole hot path through the reference-facing module API:
TPVQueryLifter -> TPVFormerEncoder -> NeuSHead.prepare/render/forward_occ vs the CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from selfocc_b200 import synth, configs
from selfocc_b200.registry import build_head
import selfocc_b200.segmentor  # noqa: F401


def _setup(color_dims=0, return_sem=False, seed=0):
    if not torch.cuda.is_available():
        pytest.skip('needs CUDA')
    torch.manual_seed(seed)
    margs, rng = synth.small_mapping(8, 4, rng=20.0, z0=-2.0, z1=4.0)
    cfg = configs.hot_path_config(mapping_args=margs, pc_range=rng, num_cams=6, num_layers=2, num_points_cross=(6, 6, 4),
                                  num_points_self=4, num_samples=48, ray_number=(9, 16), ray_img_size=(90, 160),
                                  color_dims=color_dims, return_sem=return_sem)
    model = build_head(cfg)
    model.encoder.init_weights()
    # perturb the zero-initialised offset / weight projections so the test is not the trivial uniform-softmax case
    with torch.no_grad():
        for n, p in model.named_parameters():
            if 'sampling_offsets.weight' in n or 'attention_weights.weight' in n:
                p.normal_(0, 0.05)
        model.lifter.tpv_hw.mul_(0.5); model.lifter.tpv_zh.mul_(0.5); model.lifter.tpv_wz.mul_(0.5)
        model.head.model.field.deviation_network.variance.fill_(0.25)
    model.eval()
    l2i, i2l = synth.camera_rig(synth.NUSC_YAWS, f=126.6, cx=80., cy=45., height=0.5, radius=0.2)
    metas = [dict(lidar2img=list(l2i), img2lidar=list(i2l), img_shape=(90, 160))]
    feats = [torch.randn(1, 6, 96, h, w) for h, w in [(12, 20), (6, 10), (3, 5), (2, 3)]]
    return model, cfg, margs, rng, metas, feats, torch.tensor(l2i, dtype=torch.float32), torch.tensor(i2l, dtype=torch.float32)


def _oracle_planes(model, cfg, margs, feats, l2i):
    from oracle.mapping import GridMeterMappingRef
    from oracle import lifting as ol
    p = {k[len('encoder.'):]: v.detach().cpu().double() for k, v in model.state_dict().items() if k.startswith('encoder.')}
    mref = GridMeterMappingRef(**margs)
    enc = cfg['encoder']
    ocfg = dict(num_freqs=[12] * 3, tot_range=enc['positional_encoding']['tot_range'], num_points_cross=enc['num_points_cross'],
                num_points_self=enc['num_points_self'][0], num_layers=enc['num_layers'], num_heads=6, num_cams=6)
    planes = [model.lifter.tpv_hw, model.lifter.tpv_zh, model.lifter.tpv_wz]
    planes = [q.detach().cpu().double() for q in planes]
    # fp64 oracle: positional features/tables are fp32 constants promoted to fp64
    out = ol.tpv_encoder_ref(_P64(p), mref, planes, [f.double() for f in feats], l2i[None], (90, 160), ocfg)
    return mref, out


class _P64(dict):
    """parameter dict whose fp32 constants get promoted when they meet fp64 activations."""
    def __init__(self, d):
        super().__init__(d)


def test_encoder_matches_oracle():
    model, cfg, margs, rng, metas, feats, l2i, i2l = _setup()
    dev = torch.device('cuda:0')
    model.to(dev)
    import oracle.lifting as ol
    # run the oracle in fp64 (tables stay fp32-valued): monkeypatch-free -- the oracle promotes via torch type promotion
    orig = ol.tpv_pos_features
    ol.tpv_pos_features = lambda *a, **k: [f.double() for f in orig(*a, **k)]
    orig_ps = ol.point_sampling_ref
    ol.point_sampling_ref = lambda r, m, s: tuple(t.double() if t.dtype.is_floating_point else t for t in orig_ps(r, m, s))
    orig_cv = ol.cross_view_ref_points
    ol.cross_view_ref_points = lambda *a: orig_cv(*a).double()
    try:
        mref, ref = _oracle_planes(model, cfg, margs, feats, l2i)
    finally:
        ol.tpv_pos_features, ol.point_sampling_ref, ol.cross_view_ref_points = orig, orig_ps, orig_cv
    with torch.no_grad():
        res = model.lifter(ms_img_feats=[f.to(dev) for f in feats])
        out = model.encoder(representation=res['representation'], ms_img_feats=[f.to(dev) for f in feats], metas=metas)
    for a, b in zip(out['representation'], ref):
        err = (a.cpu() - b.float()).abs().max().item()
        print('encoder plane max abs err %.3e (|x| max %.2f)' % (err, b.abs().max().item()))
        assert err < 2e-4
    # the autograd (training) path of the same modules must agree with the fused inference path
    res2 = model.lifter(ms_img_feats=[f.to(dev) for f in feats])
    out2 = model.encoder(representation=res2['representation'], ms_img_feats=[f.to(dev) for f in feats], metas=metas)
    assert out2['representation'][0].requires_grad
    for a, b in zip(out['representation'], out2['representation']):
        assert torch.allclose(a, b.detach(), atol=2e-5, rtol=1e-5)
    out2['representation'][0].sum().backward()
    assert model.lifter.tpv_hw.grad is not None and torch.isfinite(model.lifter.tpv_hw.grad).all()


@pytest.mark.parametrize('color_dims,return_sem,batch', [(0, False, 0), (7, True, 300)])
def test_head_prepare_render_matches_oracle(color_dims, return_sem, batch):
    model, cfg, margs, rng, metas, feats, l2i, i2l = _setup(color_dims, return_sem)
    dev = torch.device('cuda:0')
    model.to(dev)
    from oracle.mapping import GridMeterMappingRef
    from oracle import render as orender, rays as orays, metric
    mref = GridMeterMappingRef(**margs)
    planes = [0.5 * torch.randn_like(p) for p in (model.lifter.tpv_hw, model.lifter.tpv_zh, model.lifter.tpv_wz)]
    with torch.no_grad():
        model.head.prepare(representation=planes, metas=metas)
        out = model.head.render(metas=metas, batch=batch)
    f = model.head.model.field
    w1, b1, w2, b2 = (t.detach().cpu().double() for t in (f.density_net[1].weight, f.density_net[1].bias,
                                                         f.density_net[3].weight, f.density_net[3].bias))
    H, W, Z = mref.size_h, mref.size_w, mref.size_d
    vol = orender.tpv_decode_ref(*[p[0].cpu().double() for p in planes], (H, W, Z), w1, b1, w2, b2)
    pix = orays.fixed_ray_grid([9, 16], [90, 160])
    assert torch.equal(out['ms_rays'].cpu(), pix)                         # ray order / pixel coords: bit exact
    origin, direction = orays.img2lidar_rays(i2l[None], pix)
    inv_s = float(f.deviation_network.get_variance())
    ref = orender.head_render_ref(vol, mref, origin.double(), direction.double(), rng, inv_s, batch=batch, S=48,
                                  color_dims=3 if color_dims else 0, bkgd='white')
    d, dref = out['ms_depths'][0].cpu(), ref['depth'].float()
    assert d.shape == (1, 6, 144)
    rel = ((d - dref).abs() / dref.abs().clamp_min(1e-6)).max().item()
    absrel = float(metric.cal_depth_metric_ref(d.reshape(-1).double(), dref.reshape(-1).double().clamp(1e-3, 80))['abs_rel'])
    print('pipeline depth max rel err %.3e, AbsRel vs oracle %.3e' % (rel, absrel))
    assert rel < 1e-4 and absrel < 1e-5
    assert torch.allclose(out['ms_accs'][0].cpu(), ref['acc'].float(), atol=2e-5)
    # max-depth (neus_head.py:430-438): the sample the kernel picked must be the oracle's first maximum of w / delta on every
    # ray whose two best scores are not a rounding-level tie (relative gap < 1e-5); ties are counted, not waved through
    from oracle.parity import _idx_report
    md = out['ms_max_depths'][0].cpu().double().reshape(-1, 1)
    idx_k = (ref['ts'].reshape(md.shape[0], -1) - md).abs().argmin(-1)
    rep = _idx_report(idx_k, ref, 48)
    print('max-depth index:', rep)
    assert rep['mismatch_not_tie'] == 0 and rep['mismatch_score_off'] == 0 and rep['tie_rays'] <= 0.02 * md.shape[0]
    agree = idx_k == ref['max_idx'].reshape(-1)
    assert torch.allclose(md[agree, 0].float(), ref['max_depth'].reshape(-1)[agree].float(), rtol=1e-5, atol=1e-6)
    if color_dims:
        assert torch.allclose(out['ms_colors'][0].cpu(), ref['rgb'].float(), atol=5e-5)
        assert torch.allclose(out['sem'][0].cpu(), ref['sem'].float(), atol=5e-5)
    else:
        assert out['ms_colors'][0].shape[-1] == 0


def test_forward_occ_matches_oracle():
    model, cfg, margs, rng, metas, feats, l2i, i2l = _setup(7, True)
    dev = torch.device('cuda:0')
    model.to(dev)
    from oracle.mapping import GridMeterMappingRef
    from oracle import render as orender
    mref = GridMeterMappingRef(**margs)
    planes = [0.5 * torch.randn_like(p) for p in (model.lifter.tpv_hw, model.lifter.tpv_zh, model.lifter.tpv_wz)]
    out = model.head.forward_occ(representation=planes, metas=metas, aabb=rng, resolution=1.3)
    f = model.head.model.field
    w1, b1, w2, b2 = (t.detach().cpu().double() for t in (f.density_net[1].weight, f.density_net[1].bias,
                                                         f.density_net[3].weight, f.density_net[3].bias))
    vol = orender.tpv_decode_ref(*[p[0].cpu().double() for p in planes], (mref.size_h, mref.size_w, mref.size_d), w1, b1, w2, b2)
    sdf, sem, xyz = orender.uniform_sdf_ref(vol, mref, rng, 1.3)
    assert out['sdf'].shape == sdf.shape
    assert torch.allclose(out['xyz'].cpu(), xyz, atol=1e-5)
    assert torch.allclose(out['sdf'].cpu(), sdf.float(), atol=3e-5)
    assert torch.allclose(out['logits'].cpu(), sem.float(), atol=3e-5)
    assert out['sem'].dtype == torch.int64


def test_render_sharded_single_process_equals_render():
    """dist.render_sharded without a process group degenerates to the plain render (the 2-rank collective itself is
    covered on CPU/gloo in tests/test_dist_cpu.py; slices of the flat ray order are covered in test_gpu_render.py)."""
    model, cfg, margs, rng, metas, feats, l2i, i2l = _setup()
    dev = torch.device('cuda:0')
    model.to(dev)
    from selfocc_b200.dist import render_sharded
    planes = [0.5 * torch.randn_like(p) for p in (model.lifter.tpv_hw, model.lifter.tpv_zh, model.lifter.tpv_wz)]
    with torch.no_grad():
        model.head.prepare(representation=planes, metas=metas)
        a = model.head.render(metas=metas, batch=100)
        b = render_sharded(model.head, metas, batch=100)
    assert torch.equal(a['ms_depths'][0], b['ms_depths'][0]) and torch.equal(a['ms_max_depths'][0], b['ms_max_depths'][0])
    assert torch.equal(a['ms_accs'][0], b['ms_accs'][0])


def test_kitti_like_mono_config_with_half_axis_and_colour():
    """BASELINE configs[3] shape class (config/kitti/kitti_occ.py:166-176,188,322): ONE camera, h axis not mirrored
    (h_half=True), colour decoded (color_dims=3), Z with a padded pitch; render + occupancy lattice vs the oracle."""
    if not torch.cuda.is_available():
        pytest.skip('needs CUDA')
    from oracle.mapping import GridMeterMappingRef
    from oracle import render as orender, rays as orays
    from selfocc_b200.dist import uniform_sdf_sharded
    torch.manual_seed(1)
    dev = torch.device('cuda:0')
    margs = dict(nonlinear_mode='linear', h_size=[16, 0], h_range=[25.6, 0], h_half=True, w_size=[8, 0], w_range=[12.8, 0],
                 w_half=False, d_size=[8, 0], d_range=[-2.0, 4.4, 4.4])
    rng = [-12.8, 0.0, -2.0, 12.8, 25.6, 4.4]
    cfg = configs.hot_path_config(mapping_args=margs, pc_range=rng, num_cams=1, num_layers=1, num_points_cross=(6, 6, 4),
                                  num_points_self=4, num_samples=64, ray_number=(11, 38), ray_img_size=(88, 304), color_dims=3)
    model = build_head(cfg).eval().to(dev)
    l2i, i2l = synth.camera_rig((0.,), f=180., cx=152., cy=44., height=0.3, radius=0.1)
    metas = [dict(lidar2img=list(l2i), img2lidar=list(i2l), img_shape=(88, 304))]
    feats = [torch.randn(1, 1, 96, h, w, device=dev) for h, w in [(11, 38), (6, 19), (3, 10), (2, 5)]]
    with torch.no_grad():
        res = model(ms_img_feats=feats, metas=metas, prepare=True)
        out = model.head.render(metas=metas, batch=200)
        sdf_grid, xyz = uniform_sdf_sharded(model.head, rng, 0.8)
    mref = GridMeterMappingRef(**margs)
    f = model.head.model.field
    w1, b1, w2, b2 = (t.detach().cpu().double() for t in (f.density_net[1].weight, f.density_net[1].bias,
                                                         f.density_net[3].weight, f.density_net[3].bias))
    planes = [p.detach().cpu().double() for p in res['representation']]
    vol = orender.tpv_decode_ref(planes[0][0], planes[1][0], planes[2][0], (mref.size_h, mref.size_w, mref.size_d), w1, b1, w2, b2)
    origin, direction = orays.img2lidar_rays(torch.tensor(i2l, dtype=torch.float32)[None], orays.fixed_ray_grid([11, 38], [88, 304]))
    ref = orender.head_render_ref(vol, mref, origin.double(), direction.double(), rng, float(f.deviation_network.get_variance()),
                                  batch=200, S=64, color_dims=3, bkgd='white')
    d, dref = out['ms_depths'][0].cpu(), ref['depth'].float()
    assert d.shape == (1, 1, 418)
    assert ((d - dref).abs() / dref.abs().clamp_min(1e-6)).max().item() < 1e-4
    assert torch.allclose(out['ms_colors'][0].cpu(), ref['rgb'].float(), atol=5e-5)
    sdf_ref, _, xyz_ref = orender.uniform_sdf_ref(vol, mref, rng, 0.8)
    assert torch.allclose(xyz.cpu(), xyz_ref, atol=1e-5) and torch.allclose(sdf_grid.cpu(), sdf_ref.float(), atol=3e-5)


def test_render_poses_equals_separate_renders_and_novel_view_matches_oracle():
    """8f-3: K source poses in one launch == K head.render calls (eval_novel_depth.py:159-172); Img2LiDAR's novel_view /
    trans_kw_eval branches (img2lidar.py:32-61) against the oracle's ray generator."""
    model, cfg, margs, rng, metas, feats, l2i, i2l = _setup(color_dims=3)
    dev = torch.device('cuda:0')
    model.to(dev)
    head = model.head
    head.num_samples = 64                                  # power of two: the packed kernels
    head.render_bkgd = 'white'
    planes = [0.5 * torch.randn_like(p) for p in (model.lifter.tpv_hw, model.lifter.tpv_zh, model.lifter.tpv_wz)]
    K = 3
    poses = []
    for k in range(K):                                     # temImg2lidars: the rig moved along y by k * 0.7 m
        T = np.eye(4); T[1, 3] = 0.7 * k
        poses.append([T @ m for m in metas[0]['img2lidar']])
    metas[0]['temImg2lidars'] = poses
    with torch.no_grad():
        head.prepare(representation=planes, metas=metas)
        multi = head.render_poses(metas=metas)
        singles = []
        head.img2lidar.trans_kw = head.img2lidar.trans_kw_eval = ['render_img2lidar']
        for k in range(K):
            metas[0]['render_img2lidar'] = poses[k]
            singles.append(head.render(metas=metas))
    assert multi['ms_depths'][0].shape == (K, 6, 144)
    for key in ('ms_depths', 'ms_accs', 'ms_max_depths', 'ms_colors', 'vis_normal'):
        assert torch.equal(multi[key][0], torch.cat([s[key][0] for s in singles], 0)), key
    # chunked (batch > 0, chunks straddle poses -> per-pose launches) keeps the reference's clip groups
    with torch.no_grad():
        multi_b = head.render_poses(metas=metas, batch=500)
        single_b = head.render(metas=metas, batch=500)
    assert torch.equal(multi_b['ms_depths'][0][K - 1:], single_b['ms_depths'][0])
    # novel view + eval-time key selection
    from oracle.mapping import GridMeterMappingRef
    from oracle import render as orender, rays as orays
    import os
    head.img2lidar.trans_kw, head.img2lidar.trans_kw_eval = ['img2lidar'], ['render_img2lidar']
    head.img2lidar.novel_view = [0.4, -0.3, 0.2, 12.0]
    os.environ['eval'] = 'true'
    try:
        with torch.no_grad():
            out = head.render(metas=metas)                 # eval -> trans_kw_eval -> poses[K-1], then the novel view
    finally:
        os.environ['eval'] = 'false'
    mref = GridMeterMappingRef(**margs)
    f = head.model.field
    w1, b1, w2, b2 = (t.detach().cpu().double() for t in (f.density_net[1].weight, f.density_net[1].bias,
                                                         f.density_net[3].weight, f.density_net[3].bias))
    vol = orender.tpv_decode_ref(*[p[0].cpu().double() for p in planes], (mref.size_h, mref.size_w, mref.size_d), w1, b1, w2, b2)
    pix = orays.fixed_ray_grid([9, 16], [90, 160])
    M = torch.tensor(np.asarray(poses[K - 1]), dtype=torch.float32)
    origin, direction = orays.img2lidar_rays(M[None], pix, novel_view=[0.4, -0.3, 0.2, 12.0])
    ref = orender.head_render_ref(vol, mref, origin.double(), direction.double(), rng, float(f.deviation_network.get_variance()),
                                  S=64, color_dims=3, bkgd='white')
    d, dref = out['ms_depths'][0].cpu().double(), ref['depth']
    good = ((d - dref).abs() / dref.abs().clamp_min(1e-6)) <= 1e-4
    assert good.float().mean() > 0.99                               # the rest: fp32 cell-face flips (oracle/parity.py)
    assert torch.allclose(out['ms_colors'][0].cpu().double()[good], ref['rgb'][good], atol=1e-4)


def test_device_depth_metric_matches_reference_arithmetic():
    """8f-3: selfocc_b200.metric.DepthMetric vs DepthMetric._after_step restated on the CPU (utils/metric_util.py:247-349)."""
    if not torch.cuda.is_available():
        pytest.skip('needs CUDA')
    from oracle.metric import depth_metric_step_ref
    from selfocc_b200.metric import DepthMetric, depth_sample
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(0)
    N, n, h, w = 6, 3000, 45, 80
    pred = torch.rand(N, h, w, generator=g) * 60 + 0.5
    loc = torch.rand(N, n, 2, generator=g) * 1.04 - 0.02          # a few points beyond the border
    gt = torch.rand(N, n, generator=g) * 70 + 1.0
    mask = torch.rand(N, n, generator=g) < 0.7
    ref, pred_s = depth_metric_step_ref(loc, gt, mask, pred)
    assert torch.allclose(depth_sample(pred.to(dev), loc.to(dev)).cpu(), pred_s, rtol=1e-6, atol=1e-6)
    dm = DepthMetric(camera_names=['c%d' % i for i in range(N)]).to(dev)
    for _ in range(2):
        dm._after_step(loc.to(dev), gt.to(dev), mask.to(dev), pred.to(dev))
    res = dm._after_epoch()
    for ti, typ in enumerate(dm.eval_types):
        for k in ('abs_rel', 'sq_rel', 'rmse', 'rmse_log', 'a1', 'a2', 'a3', 'scaling'):
            assert torch.allclose(res[k][ti].cpu(), ref[typ][k], rtol=2e-5, atol=1e-6), (typ, k, res[k][ti].cpu(), ref[typ][k])
    assert float(dm.count) == 2.0
