"""GPU parity: lifting kernels (MSDA op, fused cross/self attention cores, point sampling, index lists)
through the C ABI vs the CPU oracle."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from selfocc_b200 import synth
from selfocc_b200.mapping import GridMeterMapping


def _dev():
    if not torch.cuda.is_available():
        pytest.skip('needs CUDA')
    return torch.device('cuda:0')


def _levels(shapes, dev):
    ss = torch.tensor(shapes, dtype=torch.int64)
    lsi = torch.cat([ss.new_zeros(1), ss.prod(1).cumsum(0)[:-1]])
    return ss.to(dev), lsi.to(dev)


@pytest.mark.parametrize('B,Hd,Dh,Q,P,shapes', [
    (2, 6, 16, 37, 8, [(14, 25), (7, 13), (4, 7), (2, 4)]),
    (1, 4, 32, 65, 3, [(9, 9), (5, 9), (9, 5)]),
    (1, 6, 16, 1, 48, [(6, 10)]),
])
def test_msda_forward_backward(B, Hd, Dh, Q, P, shapes):
    dev = _dev()
    from oracle.lifting import msda_ref
    from selfocc_b200 import ops
    g = torch.Generator().manual_seed(1)
    L = len(shapes)
    Nv = sum(h * w for h, w in shapes)
    value = torch.randn(B, Nv, Hd, Dh, generator=g)
    loc = torch.rand(B, Q, Hd, L, P, 2, generator=g) * 1.3 - 0.15   # some samples fall outside [0,1]
    w = torch.softmax(torch.randn(B, Q, Hd, L * P, generator=g), -1).view(B, Q, Hd, L, P)
    v64, l64, w64 = (t.double().requires_grad_(True) for t in (value, loc, w))
    ref = msda_ref(v64, shapes, l64, w64)
    ss, lsi = _levels(shapes, dev)
    vd, ld, wd = (t.to(dev).requires_grad_(True) for t in (value, loc, w))
    out = ops.MultiScaleDeformableAttnFunction.apply(vd, ss, lsi, ld, wd, 64)
    assert out.shape == (B, Q, Hd * Dh)
    assert torch.allclose(out.detach().cpu(), ref.detach().float(), atol=2e-5, rtol=1e-5)
    go = torch.randn(B, Q, Hd * Dh, generator=g)
    ref.backward(go.double())
    out.backward(go.to(dev))
    assert torch.allclose(vd.grad.cpu(), v64.grad.float(), atol=5e-5, rtol=1e-4)
    assert torch.allclose(wd.grad.cpu(), w64.grad.float(), atol=5e-5, rtol=1e-4)
    # location gradients are discontinuous exactly on pixel edges; random locations never sit there
    assert torch.allclose(ld.grad.cpu(), l64.grad.float(), atol=5e-4, rtol=1e-3)


def test_msda_one_hot_integer_centres():
    """Analytic known answer (SURVEY.md 8c iii): one-hot weight at a pixel centre returns that value row."""
    dev = _dev()
    from selfocc_b200 import ops
    shapes = [(5, 7)]
    value = torch.arange(5 * 7 * 2 * 16, dtype=torch.float32).reshape(1, 35, 2, 16)
    loc = torch.tensor([[(3 + 0.5) / 7, (2 + 0.5) / 5]]).reshape(1, 1, 1, 1, 1, 2).repeat(1, 1, 2, 1, 1, 1)
    w = torch.ones(1, 1, 2, 1, 1)
    ss, lsi = _levels(shapes, dev)
    out = ops.msda_forward(value.to(dev), ss, lsi, loc.to(dev), w.to(dev)).cpu()
    assert torch.equal(out.view(2, 16), value[0, 2 * 7 + 3])


def _rig(n_cam):
    l2i, i2l = synth.camera_rig(synth.NUSC_YAWS[:n_cam], f=126.6, cx=80., cy=45., height=0.5, radius=0.2)
    return torch.tensor(l2i, dtype=torch.float32)


def test_point_sampling_and_index_lists_bit_exact():
    dev = _dev()
    from oracle.mapping import GridMeterMappingRef
    from oracle import lifting as ol
    from selfocc_b200 import ops
    margs, _ = synth.small_mapping(20, 6, rng=30.0)
    mref = GridMeterMappingRef(**margs)
    l2i = _rig(6)
    for r3 in ol.ref_3d_tables(mref, [5, 7, 4]):
        uv_ref, mask_ref = ol.point_sampling_ref(r3[None], l2i[None], (90, 160))
        uv, mask, vis = ops.point_sampling(r3.contiguous().to(dev), l2i.to(dev), (90, 160))
        assert torch.equal(mask.cpu().bool(), mask_ref[:, 0])        # index-generating: bit exact
        assert torch.equal(uv.cpu(), uv_ref[:, 0])                   # same fp32 op order as the reference
        assert torch.equal(vis.cpu().bool(), mask_ref[:, 0].any(-1))
        assert 0 < mask_ref.sum() < mask_ref.numel()
        lists, lens = ops.visible_index_lists(mask)
        idx_ref = ol.visible_index_lists(mask_ref)
        for c in range(6):
            assert int(lens[c]) == len(idx_ref[c])
            assert torch.equal(lists[c, :int(lens[c])].cpu(), idx_ref[c])  # int64 indices, bit exact


def _attn_params(C, Hd, L, P, g, pre):
    p = {}
    p[pre + 'sampling_offsets.weight'] = 0.3 * torch.randn(Hd * L * P * 2, C, generator=g)
    p[pre + 'sampling_offsets.bias'] = 2.0 * torch.randn(Hd * L * P * 2, generator=g)
    p[pre + 'attention_weights.weight'] = 0.5 * torch.randn(Hd * L * P, C, generator=g)
    p[pre + 'attention_weights.bias'] = 0.5 * torch.randn(Hd * L * P, generator=g)
    p[pre + 'value_proj.weight'] = torch.eye(C)
    p[pre + 'value_proj.bias'] = torch.zeros(C)
    return p


def test_fused_cross_attention_core_matches_rebatch_reference():
    dev = _dev()
    from oracle.mapping import GridMeterMappingRef
    from oracle import lifting as ol
    from selfocc_b200 import ops
    g = torch.Generator().manual_seed(5)
    C, Hd, N = 96, 6, 6
    shapes = [(12, 20), (6, 10), (3, 5), (2, 3)]
    Nv = sum(h * w for h, w in shapes)
    margs, _ = synth.small_mapping(10, 4, rng=30.0)
    mref = GridMeterMappingRef(**margs)
    l2i = _rig(N)
    tables = ol.ref_3d_tables(mref, [48, 20, 8])      # D = 8 (vec4, 1 group) / 20 (scalar, 2 groups) / 48 (vec4, 4 groups)
    for r3 in tables:
        D, Q = r3.shape[:2]
        uv_ref, mask_ref = ol.point_sampling_ref(r3[None], l2i[None], (90, 160))
        query = torch.randn(1, Q, C, generator=g)
        feat = torch.randn(N, Nv, 1, C, generator=g)
        p = _attn_params(C, Hd, len(shapes), D, g, 'x.deformable_attention.')
        p['x.output_proj.weight'] = torch.eye(C)
        p['x.output_proj.bias'] = torch.zeros(C)
        ref, _ = ol.image_cross_attn_ref({k: v.double() for k, v in p.items()}, 'x.', query.double(), feat.double(), shapes,
                                         uv_ref.double(), mask_ref, Hd, N)
        slots_ref = (ref - query.double())[0].float()
        off = F.linear(query[0], p['x.deformable_attention.sampling_offsets.weight'], p['x.deformable_attention.sampling_offsets.bias'])
        lg = F.linear(query[0], p['x.deformable_attention.attention_weights.weight'], p['x.deformable_attention.attention_weights.bias'])
        ss, lsi = _levels(shapes, dev)
        uv, mask, vis = ops.point_sampling(r3.contiguous().to(dev), l2i.to(dev), (90, 160))
        slots, count = ops.tpv_cross_attn_forward(feat[:, :, 0].reshape(N, Nv, Hd, C // Hd).contiguous().to(dev), ss, lsi,
                                                  off.view(Q, Hd, len(shapes), D, 2).contiguous().to(dev),
                                                  lg.view(Q, Hd, len(shapes), D).contiguous().to(dev), uv, vis, want_count=True)
        assert torch.equal(count.cpu().long(), mask_ref[:, 0].any(-1).sum(0))
        err = (slots.cpu() - slots_ref).abs().max().item()
        print('cross-attn core max abs err %.3e' % err)
        assert err < 5e-5
        # the two kernel generations (shared set-up, D % (4 * groups) == 0, vs one set-up per lane) state the same function
        from selfocc_b200 import _lib
        _lib.load().so_attn_force_v1(1)
        try:
            slots1, count1 = ops.tpv_cross_attn_forward(feat[:, :, 0].reshape(N, Nv, Hd, C // Hd).contiguous().to(dev), ss, lsi,
                                                        off.view(Q, Hd, len(shapes), D, 2).contiguous().to(dev),
                                                        lg.view(Q, Hd, len(shapes), D).contiguous().to(dev), uv, vis, want_count=True)
        finally:
            _lib.load().so_attn_force_v1(0)
        assert torch.allclose(slots1, slots, atol=2e-5) and torch.equal(count1, count)


@pytest.mark.parametrize('P', [5, 12])       # 12 (every shipped config): the shared-set-up kernel; 5: the per-lane kernel
def test_fused_self_attention_core(P):
    dev = _dev()
    from oracle import lifting as ol
    from selfocc_b200 import ops
    g = torch.Generator().manual_seed(9)
    C, Hd = 96, 6
    H, W, Z = 9, 7, 4
    shapes = [(H, W), (Z, H), (W, Z)]
    Q = H * W + Z * H + W * Z
    ref2d = ol.cross_view_ref_points(H, W, Z, [P, P, P])
    query = torch.randn(1, Q, C, generator=g)
    p = _attn_params(C, Hd, 3, P, g, 'a.')
    p['a.output_proj.weight'] = torch.eye(C)
    p['a.output_proj.bias'] = torch.zeros(C)
    out_ref = ol.cross_view_self_attn_ref({k: v.double() for k, v in p.items()}, 'a.', query.double(),
                                          torch.zeros_like(query).double(), ref2d[None].double(), shapes, Hd, P)
    core_ref = (out_ref - query.double())[0].float()
    off = F.linear(query[0], p['a.sampling_offsets.weight'], p['a.sampling_offsets.bias']).view(Q, Hd, 3, P, 2)
    lg = F.linear(query[0], p['a.attention_weights.weight'], p['a.attention_weights.bias']).view(Q, Hd, 3, P)
    ss, lsi = _levels(shapes, dev)
    out = ops.tpv_self_attn_forward(query[0].view(Q, Hd, C // Hd).contiguous().to(dev), ss, lsi, off.contiguous().to(dev),
                                    lg.contiguous().to(dev), ref2d.contiguous().to(dev))
    err = (out.cpu() - core_ref).abs().max().item()
    print('self-attn core max abs err %.3e' % err)
    assert err < 5e-5
    # the two kernel generations state the same function
    from selfocc_b200 import _lib
    _lib.load().so_attn_force_v1(1)
    try:
        out1 = ops.tpv_self_attn_forward(query[0].view(Q, Hd, C // Hd).contiguous().to(dev), ss, lsi, off.contiguous().to(dev),
                                         lg.contiguous().to(dev), ref2d.contiguous().to(dev))
    finally:
        _lib.load().so_attn_force_v1(0)
    assert torch.allclose(out1, out, atol=2e-5)


def test_fused_flatten_equals_the_reference_sequence():
    """A3 (tpvformer_encoder.py:261-277): so_flatten_level == flatten(3).permute + cams_embeds + level_embeds + cat, bit for bit."""
    if not torch.cuda.is_available():
        pytest.skip('needs CUDA')
    dev = torch.device('cuda:0')
    from selfocc_b200 import ops
    g = torch.Generator().manual_seed(0)
    N, C = 6, 96
    shapes = [(12, 25), (6, 13), (3, 7), (2, 3)]
    feats = [torch.randn(1, N, C, h, w, generator=g).to(dev) for h, w in shapes]
    cams, lvls = torch.randn(N, C, generator=g).to(dev), torch.randn(4, C, generator=g).to(dev)
    ref = []
    for l, f in enumerate(feats):
        t = f.flatten(3).permute(1, 0, 3, 2)
        t = t + cams[:, None, None, :]
        ref.append(t + lvls[None, None, l:l + 1, :])
    ref = torch.cat(ref, 2).permute(0, 2, 1, 3).contiguous()
    got = ops.flatten_levels(feats, cams, lvls)
    assert got.shape == ref.shape and torch.equal(got, ref)
