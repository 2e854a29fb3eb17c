"""8f-3: DepthMetric on the device (reference utils/metric_util.py:282-397) -- same buffers, same
``_reset / _after_step / _after_epoch`` surface and the same numbers, but the per-camera boolean-mask indexing
(``depth_gt_i[depth_mask_i]``: a device->host sync per camera per frame) is replaced by two small kernels
(``so_depth_metric_sample`` + ``so_depth_metric_sums``) and a sort-based masked median, so a frame's metric step
enqueues without synchronising."""
import ctypes as C
import torch
import torch.nn as nn

from . import _lib
from .ops import _chk, _p, _stream

KEYS = ('abs_rel', 'sq_rel', 'rmse', 'rmse_log', 'a1', 'a2', 'a3')


def depth_sample(depth_pred, depth_loc):
    """depth_pred [N,h,w], depth_loc [N,n,2] in [0,1] -> [N,n] (grid_sample bilinear / border / align_corners=True)."""
    lib = _lib.load()
    _chk(depth_pred, name='depth_pred'); _chk(depth_loc, name='depth_loc')
    N, h, w = depth_pred.shape
    n = depth_loc.shape[1]
    out = torch.empty(N, n, device=depth_pred.device)
    _lib.check(lib.so_depth_metric_sample(_p(depth_pred), _p(depth_loc), N, n, h, w, _p(out), _stream()), 'so_depth_metric_sample')
    return out


def depth_metric_sums(sampled, depth_gt, mask_u8, scale=None):
    lib = _lib.load()
    _chk(sampled, name='sampled'); _chk(depth_gt, name='depth_gt'); _chk(mask_u8, torch.uint8, 'depth_mask'); _chk(scale, name='scale')
    N, n = sampled.shape
    sums = torch.empty(N, 8, device=sampled.device)
    _lib.check(lib.so_depth_metric_sums(_p(sampled), _p(depth_gt), _p(mask_u8), _p(scale), N, n, _p(sums), _stream()),
               'so_depth_metric_sums')
    return sums


def masked_median(x, mask):
    """torch.median(x_i[mask_i]) per row (the LOWER median, like torch.median) without boolean indexing."""
    filled = torch.where(mask, x, torch.full_like(x, float('inf')))
    srt = filled.sort(dim=1).values
    cnt = mask.sum(1)
    idx = ((cnt - 1).clamp_min(0) // 2).unsqueeze(1)
    return srt.gather(1, idx).squeeze(1)


def metrics_from_sums(sums):
    """[N,8] error sums -> dict of [N] metrics (cal_depth_metric, metric_util.py:247-279)."""
    c = sums[:, 7].clamp_min(1.0)
    return {'abs_rel': sums[:, 0] / c, 'sq_rel': sums[:, 1] / c, 'rmse': (sums[:, 2] / c).sqrt(), 'rmse_log': (sums[:, 3] / c).sqrt(),
            'a1': sums[:, 4] / c, 'a2': sums[:, 5] / c, 'a3': sums[:, 6] / c}


class DepthMetric(nn.Module):
    def __init__(self, camera_names=['front'], eval_types=['raw', 'median']):
        super().__init__()
        self.num_cams, self.camera_names = len(camera_names), camera_names
        self.num_types, self.eval_types = len(eval_types), eval_types
        for k in KEYS + ('scaling',):
            self.register_buffer(k, torch.zeros(self.num_types, self.num_cams))
        self.register_buffer('count', torch.zeros(1))

    def _reset(self):
        for k in KEYS + ('scaling', 'count'):
            getattr(self, k).zero_()

    @torch.no_grad()
    def _after_step(self, depth_loc, depth_gt, depth_mask, depth_pred):
        """depth_loc [N,n,2], depth_gt [N,n], depth_mask bool [N,n], depth_pred [N,h,w] (metric_util.py:311-349)."""
        depth_loc, depth_gt, depth_pred = depth_loc.float().contiguous(), depth_gt.float().contiguous(), depth_pred.float().contiguous()
        mask = depth_mask.bool()
        mask_u8 = mask.to(torch.uint8).contiguous()
        sampled = depth_sample(depth_pred, depth_loc)
        for ti, typ in enumerate(self.eval_types):
            if typ == 'raw':
                scale = torch.ones(self.num_cams, device=sampled.device)
            elif typ == 'median':
                scale = masked_median(depth_gt, mask) / masked_median(sampled, mask)
            else:
                raise NotImplementedError(typ)
            m = metrics_from_sums(depth_metric_sums(sampled, depth_gt, mask_u8, scale.contiguous()))
            self.scaling[ti] += scale
            for k in KEYS:
                getattr(self, k)[ti] += m[k]
        self.count += 1

    def _after_epoch(self, logger=None):
        """metric_util.py:351-397: all-reduce over ranks when torch.distributed is initialised, then average."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.barrier()
            for k in ('count',) + KEYS + ('scaling',):
                dist.all_reduce(getattr(self, k))
            dist.barrier()
        res = {k: getattr(self, k) / self.count for k in KEYS + ('scaling',)}
        if logger is not None and (not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0):
            logger.info('Averaging over %s samples.' % self.count.item())
            for ti, typ in enumerate(self.eval_types):
                logger.info('%s evaluation:' % typ)
                for cam, name in enumerate(self.camera_names):
                    logger.info('%12s | ' % name + ' '.join('%s %.3f' % (k, res[k][ti, cam]) for k in KEYS + ('scaling',)))
                logger.info('%12s | ' % 'All' + ' '.join('%s %.3f' % (k, res[k][ti].mean()) for k in KEYS + ('scaling',)))
        return res
