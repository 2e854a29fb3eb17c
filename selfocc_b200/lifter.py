"""A1: TPVQueryLifter (reference model/lifter/tpv_query_lifter.py:6-36) -- same name, ctor kwargs,
parameter names (``tpv_hw``, ``tpv_zh``, ``tpv_wz``) and output dict."""
import torch
import torch.nn as nn
from .registry import MODELS


@MODELS.register_module()
class TPVQueryLifter(nn.Module):
    def __init__(self, tpv_h, tpv_w, tpv_z, dim, init_cfg=None, **kwargs):
        super().__init__()
        self.tpv_h, self.tpv_w, self.tpv_z, self.dim = tpv_h, tpv_w, tpv_z, dim
        self.tpv_hw = nn.Parameter(torch.randn(1, tpv_h * tpv_w, dim))
        self.tpv_zh = nn.Parameter(torch.randn(1, tpv_z * tpv_h, dim))
        self.tpv_wz = nn.Parameter(torch.randn(1, tpv_w * tpv_z, dim))

    def forward(self, ms_img_feats, *args, **kwargs):
        bs = ms_img_feats[0].shape[0]
        if bs == 1:  # the planes are read-only downstream: no 30 MB copy per frame
            return {'representation': [self.tpv_hw, self.tpv_zh, self.tpv_wz]}
        return {'representation': [p.repeat(bs, 1, 1) for p in (self.tpv_hw, self.tpv_zh, self.tpv_wz)]}
