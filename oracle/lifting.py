"""Oracle: image -> tri-plane lifting (test infrastructure, see oracle/__init__.py).

In-repo pieces restated from the reference (file:line in each docstring).  The deformable
attention core is mmcv==2.0.1 ``multi_scale_deform_attn`` -- NOT vendored, NOT installed:
PARITY UNPINNED for that function; its published semantics (Deformable-DETR) are restated
in ``msda_ref`` and anchored on the reference call sites
model/encoder/bevformer/attention/image_cross_attention.py:338-345 and
model/encoder/tpvformer/attention/cross_view_hybrid_attention.py:109-116.

Parameters are passed as a flat dict keyed like the reference modules' ``state_dict``
(e.g. ``layers.0.attentions.1.attn_hw.deformable_attention.value_proj.weight``) so a
product module's ``state_dict()`` can be fed straight in.
"""
import math
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- geometry tables
def tpv_plane_meters(mapping):
    """tpvformer_encoder.py:84-101: metre coordinates of each plane's cells (for the pos-embed)."""
    H, W, Z = mapping.size_h, mapping.size_w, mapping.size_d
    ar = lambda n: torch.arange(n, dtype=torch.float)
    hw = torch.stack([ar(H)[:, None].expand(-1, W), ar(W)[None].expand(H, -1), torch.zeros(H, W)], -1)
    zh = torch.stack([ar(H)[None].expand(Z, -1), torch.zeros(Z, H), ar(Z)[:, None].expand(-1, H)], -1)
    wz = torch.stack([torch.zeros(W, Z), ar(W)[:, None].expand(-1, Z), ar(Z)[None].expand(W, -1)], -1)
    return (mapping.grid2meter(hw)[..., [0, 1]], mapping.grid2meter(zh)[..., [1, 2]],
            mapping.grid2meter(wz)[..., [0, 2]])


def ref_3d_tables(mapping, num_points_cross):
    """tpvformer_encoder.py:131-154: per-plane pillars of 3-D reference points in metres,
    returned as [P, Q, 3] (P = points per pillar along the plane's missing axis)."""
    H, W, Z = mapping.size_h, mapping.size_w, mapping.size_d
    ar = lambda n: torch.arange(n, dtype=torch.float)
    P0, P1, P2 = num_points_cross[2], num_points_cross[1], num_points_cross[0]
    hw = torch.stack([ar(H)[:, None, None].expand(H, W, P0), ar(W)[None, :, None].expand(H, W, P0),
                      torch.linspace(0, Z - 1, P0)[None, None].expand(H, W, P0)], -1)
    zh = torch.stack([ar(H)[None, :, None].expand(Z, H, P1), torch.linspace(0, W - 1, P1)[None, None].expand(Z, H, P1),
                      ar(Z)[:, None, None].expand(Z, H, P1)], -1)
    wz = torch.stack([torch.linspace(0, H - 1, P2)[None, None].expand(W, Z, P2), ar(W)[:, None, None].expand(W, Z, P2),
                      ar(Z)[None, :, None].expand(W, Z, P2)], -1)
    return [mapping.grid2meter(t).flatten(0, 1).transpose(0, 1).contiguous() for t in (hw, zh, wz)]


def cross_view_ref_points(H, W, Z, P):
    """tpvformer/utils.py:5-71 (offset=0): [HW+ZH+WZ, 3 (level = plane hw/zh/wz), P, 2 (x,y)].
    ``P`` = [p_wz, p_zh, p_hw] like ``num_points_self``; all three are equal in shipped configs."""
    lin = lambda n: torch.linspace(0, n - 1, n) / n
    hs, ws, zs = lin(H), lin(W), lin(Z)
    pil = lambda n, p: torch.linspace(0, n - 1, p) / n

    def block(dims, fx, fy, p):
        # fx/fy: callables giving the normalised x / y coordinate of shape dims+[p]
        return torch.stack([fx.expand(*dims, p), fy.expand(*dims, p)], -1).flatten(0, 1)

    p = P[2]  # queries of the hw plane [H, W]
    hw = torch.stack([
        block((H, W), ws[None, :, None], hs[:, None, None], p),
        block((H, W), hs[:, None, None], pil(Z, p)[None, None], p),
        block((H, W), pil(Z, p)[None, None], ws[None, :, None], p)], 1)
    p = P[1]  # queries of the zh plane [Z, H]
    zh = torch.stack([
        block((Z, H), pil(W, p)[None, None], hs[None, :, None], p),
        block((Z, H), hs[None, :, None], zs[:, None, None], p),
        block((Z, H), zs[:, None, None], pil(W, p)[None, None], p)], 1)
    p = P[0]  # queries of the wz plane [W, Z]
    wz = torch.stack([
        block((W, Z), ws[:, None, None], pil(H, p)[None, None], p),
        block((W, Z), pil(H, p)[None, None], zs[None, :, None], p),
        block((W, Z), zs[None, :, None], ws[:, None, None], p)], 1)
    return torch.cat([hw, zh, wz], 0)


def pos_freq_features(num_freqs, meter01):
    """tpvformer_pos_embed.py:6-14: meter01 [A,B,2] in [0,1] -> [A*B, 4*num_freqs] (sin,cos interleaved)."""
    freqs = math.pi * (2 ** torch.arange(-1, num_freqs - 1, dtype=torch.float))
    mf = meter01.unsqueeze(-1) * freqs
    return torch.stack([torch.sin(mf), torch.cos(mf)], -1).flatten(-3).flatten(0, 1)


def tpv_pos_features(mapping, num_freqs, tot_range):
    """tpvformer_pos_embed.py:25-51: normalise plane metres by the point-cloud range then encode."""
    hw, zh, wz = [m.clone() for m in tpv_plane_meters(mapping)]
    r = tot_range
    nx = lambda v: (v - r[0]) / (r[3] - r[0])
    ny = lambda v: (v - r[1]) / (r[4] - r[1])
    nz = lambda v: (v - r[2]) / (r[5] - r[2])
    hw = torch.stack([nx(hw[..., 0]), ny(hw[..., 1])], -1)
    zh = torch.stack([ny(zh[..., 0]), nz(zh[..., 1])], -1)
    wz = torch.stack([nx(wz[..., 0]), nz(wz[..., 1])], -1)
    return [pos_freq_features(n, m) for n, m in zip(num_freqs, (hw, zh, wz))]


def point_sampling_ref(ref_3d, lidar2img, img_shape):
    """bevformer/utils.py:116-206 without the optional post_rots / focal_ratios branches.
    ref_3d [B, D, Q, 3] metres, lidar2img [B, N, 4, 4], img_shape (h, w)
    -> reference_points_cam [N, B, Q, D, 2] in image-normalised (x, y), mask [N, B, Q, D] bool."""
    p = torch.cat([ref_3d.float(), torch.ones_like(ref_3d[..., :1])], -1).permute(1, 0, 2, 3)
    D, B, Q = p.shape[:3]
    N = lidar2img.shape[1]
    cam = torch.matmul(lidar2img.float().view(1, B, N, 1, 4, 4), p.view(D, B, 1, Q, 4, 1)).squeeze(-1)
    eps = 1e-5
    mask = cam[..., 2:3] > eps
    uv = cam[..., 0:2] / torch.maximum(cam[..., 2:3], torch.ones_like(cam[..., 2:3]) * eps)
    uv[..., 0] /= img_shape[1]
    uv[..., 1] /= img_shape[0]
    mask = mask & (uv[..., 1:2] > 0.0) & (uv[..., 1:2] < 1.0) & (uv[..., 0:1] < 1.0) & (uv[..., 0:1] > 0.0)
    return uv.permute(2, 1, 3, 0, 4).contiguous(), mask.permute(2, 1, 3, 0, 4).squeeze(-1).contiguous()


# ----------------------------------------------------------------------------- attention core
def msda_ref(value, spatial_shapes, sampling_locations, attention_weights):
    """Multi-scale deformable attention, mmcv 2.0.1 semantics restated from public knowledge
    (source unavailable offline; PARITY UNPINNED).

    value [B, sum(h_l*w_l), Hd, Dh]; spatial_shapes [[h_l, w_l]]; sampling_locations
    [B, Q, Hd, L, P, 2] normalised (x, y); attention_weights [B, Q, Hd, L, P] -> [B, Q, Hd*Dh].
    out[b,q,h,:] = sum_{l,p} w * bilinear(value_l[b,:,h,:], loc), pixel = loc*size - 0.5
    (align_corners=False), zero padding."""
    B, _, Hd, Dh = value.shape
    _, Q, _, L, P, _ = sampling_locations.shape
    out = value.new_zeros(B * Hd, Dh, Q)
    start = 0
    for l, (h, w) in enumerate([(int(a), int(b)) for a, b in spatial_shapes]):
        v = value[:, start:start + h * w].permute(0, 2, 3, 1).reshape(B * Hd, Dh, h, w)
        start += h * w
        g = (2 * sampling_locations[:, :, :, l] - 1).permute(0, 2, 1, 3, 4).reshape(B * Hd, Q, P, 2)
        s = F.grid_sample(v, g, mode='bilinear', padding_mode='zeros', align_corners=False)  # BHd,Dh,Q,P
        wl = attention_weights[:, :, :, l].permute(0, 2, 1, 3).reshape(B * Hd, 1, Q, P)
        out = out + (s * wl).sum(-1)
    return out.view(B, Hd * Dh, Q).transpose(1, 2).contiguous()


def _lin(p, key, x):
    return F.linear(x, p[key + '.weight'], p[key + '.bias'])


def deform_locations(ref, offsets, spatial_shapes, per_level_ref):
    """image_cross_attention.py:323-328 (ref [B,Q,D,2] broadcast over heads+levels, P == D) and
    cross_view_hybrid_attention.py:93-99 (ref [B,Q,L,P,2] broadcast over heads only)."""
    norm = torch.tensor([[float(w), float(h)] for h, w in spatial_shapes], dtype=offsets.dtype)
    r = ref[:, :, None, :, :, :] if per_level_ref else ref[:, :, None, None, :, :]
    return r + offsets / norm[None, None, None, :, None, :]


def cross_view_self_attn_ref(p, pre, query, query_pos, ref_2d, spatial_shapes, num_heads, num_points):
    """CrossViewHybridAttention.forward (cross_view_hybrid_attention.py:63-124), eval mode,
    called as in tpvformer_encoder_layer.py:169-179: value = query (no pos), identity = query."""
    B, Q, C = query.shape
    L = len(spatial_shapes)
    q = query + query_pos
    value = _lin(p, pre + 'value_proj', query).view(B, Q, num_heads, -1)
    off = _lin(p, pre + 'sampling_offsets', q).view(B, Q, num_heads, L, num_points, 2)
    aw = _lin(p, pre + 'attention_weights', q).view(B, Q, num_heads, L * num_points).softmax(-1)
    aw = aw.view(B, Q, num_heads, L, num_points)
    loc = deform_locations(ref_2d, off, spatial_shapes, per_level_ref=True)
    out = msda_ref(value, spatial_shapes, loc, aw)
    return _lin(p, pre + 'output_proj', out) + query


def visible_index_lists(mask):
    """image_cross_attention.py:90-94: per camera, int64 indices of queries with any in-frustum
    point (batch element 0 only, as in the reference).  mask [N,B,Q,D] bool."""
    return [m[0].sum(-1).nonzero().squeeze(-1) for m in mask]


def image_cross_attn_ref(p, pre, query, feat, spatial_shapes, ref_cam, mask, num_heads, num_cams):
    """BEVCrossAttention.forward + BEVDeformableAttention.forward
    (image_cross_attention.py:84-139, 293-351), eval mode, residual = query.
    query [B,Q,C]; feat [N, sum(hw), B, C]; ref_cam [N,B,Q,D,2]; mask [N,B,Q,D]."""
    B, Q, C = query.shape
    D = ref_cam.shape[3]
    L = len(spatial_shapes)
    idx = visible_index_lists(mask)
    max_len = max(len(i) for i in idx)
    q_re = query.new_zeros(B * num_cams, max_len, C)
    r_re = ref_cam.new_zeros(B * num_cams, max_len, D, 2)
    for i in range(num_cams):
        for j in range(B):
            q_re[j * num_cams + i, :len(idx[i])] = query[j, idx[i]]
            r_re[j * num_cams + i, :len(idx[i])] = ref_cam[i, j, idx[i]]
    value = feat.permute(2, 0, 1, 3).reshape(num_cams * B, -1, C)
    dpre = pre + 'deformable_attention.'
    value = _lin(p, dpre + 'value_proj', value).view(B * num_cams, -1, num_heads, C // num_heads)
    off = _lin(p, dpre + 'sampling_offsets', q_re).view(B * num_cams, max_len, num_heads, L, D, 2)
    aw = _lin(p, dpre + 'attention_weights', q_re).view(B * num_cams, max_len, num_heads, L * D).softmax(-1)
    aw = aw.view(B * num_cams, max_len, num_heads, L, D)
    loc = deform_locations(r_re, off, spatial_shapes, per_level_ref=False)
    out = msda_ref(value, spatial_shapes, loc, aw)
    slots = torch.zeros_like(query)
    for i in range(num_cams):
        for j in range(B):
            slots[j, idx[i]] += out[j * num_cams + i, :len(idx[i])]
    count = (mask.sum(-1) > 0).permute(1, 2, 0).sum(-1).clamp(min=1.0)
    slots = slots / count[..., None]
    return _lin(p, pre + 'output_proj', slots) + query, idx


def ffn_ref(p, pre, x):
    """mmcv FFN(num_fcs=2, ReLU, add_identity=True), eval mode (tpvformer_encoder_layer.py:198-206)."""
    return x + _lin(p, pre + 'layers.1', F.relu(_lin(p, pre + 'layers.0.0', x)))


def _ln(p, key, x):
    return F.layer_norm(x, x.shape[-1:], p[key + '.weight'], p[key + '.bias'])


def tpv_layer_ref(p, pre, planes, tpv_pos, feat, img_shapes, ref_2d, ref_cams, masks, tpv_size, cfg):
    """TPVFormerLayer.forward (tpvformer_encoder_layer.py:158-219) with
    operation_order = ('self_attn','norm','cross_attn','norm','ffn','norm'), post-norm."""
    H, W, Z = tpv_size
    split = [H * W, Z * H, W * Z]
    ss = [(H, W), (Z, H), (W, Z)]
    q = cross_view_self_attn_ref(p, pre + 'attentions.0.', torch.cat(planes, 1), torch.cat(tpv_pos, 1), ref_2d, ss,
                                 cfg['num_heads'], cfg['num_points_self'])
    q = _ln(p, pre + 'norms.0', q)
    planes = list(torch.split(q, split, 1))
    names = ['attn_hw', 'attn_zh', 'attn_wz']
    planes = [image_cross_attn_ref(p, pre + 'attentions.1.%s.' % names[i], planes[i], feat, img_shapes, ref_cams[i],
                                   masks[i], cfg['num_heads'], cfg['num_cams'])[0] for i in range(3)]
    q = _ln(p, pre + 'norms.1', torch.cat(planes, 1))
    q = _ln(p, pre + 'norms.2', ffn_ref(p, pre + 'ffns.0.', q))
    return list(torch.split(q, split, 1))


def flatten_img_feats(p, ms_img_feats):
    """tpvformer_encoder.py:261-277: [B,N,C,h,w] x L -> feat [N, sum(hw), B, C] with camera and
    level embeddings added; spatial shapes [(h,w)]."""
    feats, shapes = [], []
    for lvl, f in enumerate(ms_img_feats):
        B, N, C, h, w = f.shape
        f = f.flatten(3).permute(1, 0, 3, 2)
        f = f + p['cams_embeds'][:, None, None, :] + p['level_embeds'][None, None, lvl:lvl + 1, :]
        feats.append(f)
        shapes.append((h, w))
    return torch.cat(feats, 2).permute(0, 2, 1, 3).contiguous(), shapes


def tpv_encoder_ref(p, mapping, planes, ms_img_feats, lidar2img, img_shape, cfg):
    """TPVFormerEncoder.forward (tpvformer_encoder.py:192-290), eval mode, camera_aware=False.
    planes: list of 3 [B,Q_i,C]; returns list of 3 [B,Q_i,C]."""
    B = planes[0].shape[0]
    H, W, Z = mapping.size_h, mapping.size_w, mapping.size_d
    feats = tpv_pos_features(mapping, cfg['num_freqs'], cfg['tot_range'])
    tpv_pos = [_lin(p, 'positional_encoding.position_layer_' + n, f)[None].repeat(B, 1, 1)
               for n, f in zip(('hw', 'zh', 'wz'), feats)]
    feat, shapes = flatten_img_feats(p, ms_img_feats)
    ref_cams, masks = [], []
    for r3 in ref_3d_tables(mapping, cfg['num_points_cross']):
        rc, m = point_sampling_ref(r3[None].repeat(B, 1, 1, 1), lidar2img, img_shape)
        ref_cams.append(rc)
        masks.append(m)
    ref_2d = cross_view_ref_points(H, W, Z, [cfg['num_points_self']] * 3)[None].expand(B, -1, -1, -1, -1)
    for i in range(cfg['num_layers']):
        planes = tpv_layer_ref(p, 'layers.%d.' % i, planes, tpv_pos, feat, shapes, ref_2d, ref_cams, masks,
                               (H, W, Z), cfg)
    return planes
