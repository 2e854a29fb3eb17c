"""Host-side grid <-> metre mapping (product code; mirrors the interface of the reference's
``GridMeterMapping`` -- model/encoder/bevformer/mappings.py:153-196 -- for nonlinear_mode='linear').

Used at module-construction time (reference-point tables, positional-encoding features) and to fill
the ``so_volume_desc.axis`` table consumed by the CUDA kernels.  ``'linear_upscale'`` (the quadratic
NonLinearMapping, mappings.py:199-287) is used by no shipped config and is rejected.
"""
import torch
from . import _lib


class GridMeterMapping:
    def __init__(self, nonlinear_mode='linear_upscale', h_size=(128, 32), h_range=(51.2, 28.8), h_half=False,
                 w_size=(128, 32), w_range=(51.2, 28.8), w_half=False, d_size=(20, 10), d_range=(-4.0, 4.0, 12.0)):
        if nonlinear_mode != 'linear':
            raise NotImplementedError(
                "selfocc_b200 implements nonlinear_mode='linear' only (every shipped SelfOcc config); got %r"
                % (nonlinear_mode,))
        self.nonlinear_mode = nonlinear_mode
        self._ax = {
            'h': dict(size=[float(v) for v in h_size], rng=[float(v) for v in h_range], half=bool(h_half), start=0.0),
            'w': dict(size=[float(v) for v in w_size], rng=[float(v) for v in w_range], half=bool(w_half), start=0.0),
            'd': dict(size=[float(v) for v in d_size], rng=[float(d_range[1] - d_range[0]), float(d_range[2] - d_range[1])],
                      half=True, start=float(d_range[0])),
        }
        for a in self._ax.values():
            n = int(a['size'][0] + a['size'][1])
            a['len'] = 1 + n if a['half'] else 1 + 2 * n
            a['offset'] = 0.0 if a['half'] else float(n)
        self.size_h, self.size_w, self.size_d = (self._ax[k]['len'] for k in 'hwd')

    # -- tensor maps (used for construction-time tables only; the hot path does this in-kernel) --
    @staticmethod
    def _seg(v, a0, b0, a1, b1):
        """Two-segment piecewise-linear map a -> b.  Operation order (divide, then multiply) follows
        mappings.py:53-60 / :101-109 so construction-time tables are bit-identical to the reference's."""
        inner = v / a0 * b0
        if a1 <= 0:
            return inner
        return torch.where(v > a0, b0 + (v - a0) / a1 * b1, inner)

    def _axis_g2m(self, k, g):
        a = self._ax[k]
        c = g - a['offset'] if a['offset'] != 0 else g
        m = self._seg(c.abs(), a['size'][0], a['rng'][0], a['size'][1], a['rng'][1])
        m = torch.sign(c) * m
        return m + a['start'] if k == 'd' else m

    def _axis_m2g(self, k, m):
        a = self._ax[k]
        c = m - a['start'] if k == 'd' else m
        g = torch.sign(c) * self._seg(c.abs(), a['rng'][0], a['size'][0], a['rng'][1], a['size'][1])
        if a['offset'] != 0:
            g = g + a['size'][0] + a['size'][1]
        return g

    def grid2meter(self, grid):
        """grid[..., (h, w[, d])] -> metres[..., (x, y[, z])]"""
        y = self._axis_g2m('h', grid[..., 0])
        x = self._axis_g2m('w', grid[..., 1])
        if grid.shape[-1] == 3:
            return torch.stack([x, y, self._axis_g2m('d', grid[..., 2])], -1)
        return torch.stack([x, y], -1)

    def meter2grid(self, meter, normalize=False):
        """metres[..., (x, y, z)] -> grid[..., (h, w, d)]"""
        h = self._axis_m2g('h', meter[..., 1])
        w = self._axis_m2g('w', meter[..., 0])
        d = self._axis_m2g('d', meter[..., 2])
        if normalize:
            h, w, d = h / (self.size_h - 1), w / (self.size_w - 1), d / (self.size_d - 1)
        return torch.stack([h, w, d], -1)

    # -- C-ABI descriptor --
    def volume_desc(self, n_feat=0):
        d = _lib.VolumeDesc()
        d.H, d.W, d.Z = self.size_h, self.size_w, self.size_d
        d.zpitch = (self.size_d + 7) // 8 * 8
        d.n_feat = int(n_feat)
        d.feat_pitch = (int(n_feat) + 3) // 4 * 4
        for i, k in enumerate('hwd'):
            a = self._ax[k]
            d.axis[i].start = a['start']
            d.axis[i].range0, d.axis[i].range1 = a['rng']
            d.axis[i].size0, d.axis[i].size1 = a['size']
            d.axis[i].offset = a['offset']
        return d
