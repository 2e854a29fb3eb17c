"""Stage pipeline of the reference's TPVSegmentor (model/segmentor/tpv_segmentor.py:87-125) restricted
to the hot path: lifter -> encoder -> head, each called with the whole accumulated dict.  The image
backbone / neck (third-party mmseg ResNet + FPN, out of scope) is replaced by ``ms_img_feats`` inputs."""
import torch.nn as nn
from . import lifter as _lifter, encoder as _encoder, head as _head  # noqa: F401  (import = registration, like `import model`)
from .registry import MODELS, build_head


@MODELS.register_module()
class TPVHotPath(nn.Module):
    def __init__(self, lifter=None, encoder=None, head=None, **kwargs):
        super().__init__()
        self.lifter, self.encoder, self.head = build_head(lifter), build_head(encoder), build_head(head)

    def forward(self, ms_img_feats=None, metas=None, occ_only=False, prepare=False, **kwargs):
        results = {'ms_img_feats': ms_img_feats, 'metas': metas}
        results.update(kwargs)
        results.update(self.lifter(**results))
        results.update(self.encoder(**results))
        if occ_only and hasattr(self.head, 'forward_occ'):
            outs = self.head.forward_occ(**results)
        elif prepare and hasattr(self.head, 'prepare'):
            outs = self.head.prepare(**results)
        else:
            outs = self.head(**results)
        results.update(outs)
        return results
