"""Model config dicts in the reference's format (config/nuscenes/nuscenes_depth.py:185-350 and siblings),
restricted to the hot-path stages (lifter / encoder / head).  ``hot_path_config`` scales the same
structure down for tests."""
import copy
from .synth import NUSC_MAPPING, NUSC_RANGE


def hot_path_config(mapping_args=None, pc_range=None, dim=96, num_heads=6, num_cams=6, num_levels=4, num_layers=4,
                    num_points_cross=(48, 48, 8), num_points_self=12, num_samples=256, ray_number=(450, 800),
                    ray_img_size=(900, 1600), color_dims=0, return_max_depth=True, return_sem=False, trans_kw='img2lidar',
                    ray_sample_mode='fixed', render_bkgd='white', dropout=0.1):
    mapping_args = copy.deepcopy(mapping_args or NUSC_MAPPING)
    pc_range = list(pc_range or NUSC_RANGE)
    tpv_h = 1 + (1 if mapping_args['h_half'] else 2) * sum(mapping_args['h_size'])
    tpv_w = 1 + (1 if mapping_args['w_half'] else 2) * sum(mapping_args['w_size'])
    tpv_z = 1 + sum(mapping_args['d_size'])
    layer = dict(
        type='TPVFormerLayer',
        attn_cfgs=[
            dict(type='CrossViewHybridAttention', embed_dims=dim, num_heads=num_heads, num_levels=3,
                 num_points=num_points_self, dropout=dropout, batch_first=True),
            dict(type='TPVCrossAttention', embed_dims=dim, num_cams=num_cams, dropout=dropout, batch_first=True,
                 num_heads=num_heads, num_levels=num_levels, num_points=list(num_points_cross))],
        feedforward_channels=2 * dim, ffn_dropout=dropout,
        operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'))
    return dict(
        type='TPVHotPath',
        lifter=dict(type='TPVQueryLifter', tpv_h=tpv_h, tpv_w=tpv_w, tpv_z=tpv_z, dim=dim),
        encoder=dict(type='TPVFormerEncoder', mapping_args=mapping_args, embed_dims=dim, num_cams=num_cams,
                     num_feature_levels=num_levels,
                     positional_encoding=dict(type='TPVPositionalEncoding', num_freqs=[12] * 3, embed_dims=dim,
                                              tot_range=pc_range),
                     num_points_cross=list(num_points_cross), num_points_self=[num_points_self] * 3,
                     transformerlayers=[copy.deepcopy(layer) for _ in range(num_layers)], num_layers=num_layers),
        head=dict(type='NeuSHead', roi_aabb=pc_range, resolution=0.4, near_plane=0.0, far_plane=1e10,
                  num_samples=num_samples, num_samples_importance=0, num_up_sample_steps=0, base_variance=4,
                  beta_init=0.1, beta_max=0.195, total_iters=3516 * 11, beta_hand_tune=False,
                  use_numerical_gradients=False, sample_gradient=True, return_uniform_sdf=False, return_second_grad=False,
                  return_max_depth=return_max_depth, return_sem=return_sem,
                  ray_sample_mode=ray_sample_mode, ray_number=list(ray_number), ray_img_size=list(ray_img_size),
                  ray_upper_crop=0, trans_kw=trans_kw, novel_view=None, render_bkgd=render_bkgd,
                  mapping_args=mapping_args, embed_dims=dim, color_dims=color_dims, density_layers=2, sh_deg=0,
                  sh_act='relu', two_split=False, tpv=True))
