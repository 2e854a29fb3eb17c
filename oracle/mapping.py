"""Oracle: grid <-> metre mapping (test infrastructure, see oracle/__init__.py).

Follows reference model/encoder/bevformer/mappings.py:4-150 (LinearMapping) and
:153-196 (GridMeterMapping, nonlinear_mode='linear').  Each axis is a two-segment
piecewise-linear map: ``size[0]`` cells span ``range[0]`` metres (inner), a further
``size[1]`` cells span ``range[1]`` metres (outer); h/w are mirrored about the centre
unless ``*_half``.  All shipped configs use an outer size of 0 (pure affine).
"""
import torch


class GridMeterMappingRef:
    def __init__(self, nonlinear_mode='linear', h_size=(128, 32), h_range=(51.2, 28.8), h_half=False,
                 w_size=(128, 32), w_range=(51.2, 28.8), w_half=False, d_size=(20, 10),
                 d_range=(-4.0, 4.0, 12.0)):
        if nonlinear_mode != 'linear':
            raise NotImplementedError("oracle covers nonlinear_mode='linear' only (all shipped configs)")
        self.h_size, self.h_range, self.h_half = list(h_size), list(h_range), h_half
        self.w_size, self.w_range, self.w_half = list(w_size), list(w_range), w_half
        self.d_size = list(d_size)
        self.d_range = [d_range[1] - d_range[0], d_range[2] - d_range[1]]
        self.d_start = d_range[0]
        self.size_h = (1 + self.h_size[0] + self.h_size[1]) if h_half else (1 + 2 * (self.h_size[0] + self.h_size[1]))
        self.size_w = (1 + self.w_size[0] + self.w_size[1]) if w_half else (1 + 2 * (self.w_size[0] + self.w_size[1]))
        self.size_d = 1 + self.d_size[0] + self.d_size[1]

    @staticmethod
    def _g2m(c_abs, size, rng):
        # mappings.py:53-60
        if size[1] == 0:
            return c_abs / size[0] * rng[0]
        return torch.where(c_abs > size[0], rng[0] + (c_abs - size[0]) / size[1] * rng[1], c_abs / size[0] * rng[0])

    @staticmethod
    def _m2g(m_abs, size, rng):
        # mappings.py:101-109
        if size[1] == 0:
            return m_abs / rng[0] * size[0]
        return torch.where(m_abs > rng[0], size[0] + (m_abs - rng[0]) / rng[1] * size[1], m_abs / rng[0] * size[0])

    def grid2meter(self, grid):
        """grid[..., (h, w[, d])] -> metres[..., (x, y[, z])]  (mappings.py:39-95)."""
        h, w = grid[..., 0], grid[..., 1]
        h_ctr = h if self.h_half else h - (self.h_size[0] + self.h_size[1])
        y = torch.sign(h_ctr) * self._g2m(h_ctr.abs(), self.h_size, self.h_range)
        w_ctr = w if self.w_half else w - (self.w_size[0] + self.w_size[1])
        x = torch.sign(w_ctr) * self._g2m(w_ctr.abs(), self.w_size, self.w_range)
        if grid.shape[-1] == 3:
            d = grid[..., 2]
            z = torch.sign(d) * self._g2m(d.abs(), self.d_size, self.d_range) + self.d_start
            return torch.stack([x, y, z], -1)
        return torch.stack([x, y], -1)

    def meter2grid(self, meter, normalize=False):
        """metres[..., (x, y, z)] -> grid[..., (h, w, d)]  (mappings.py:97-150)."""
        x, y, z = meter[..., 0], meter[..., 1], meter[..., 2]
        w = torch.sign(x) * self._m2g(x.abs(), self.w_size, self.w_range)
        if not self.w_half:
            w = w + self.w_size[0] + self.w_size[1]
        h = torch.sign(y) * self._m2g(y.abs(), self.h_size, self.h_range)
        if not self.h_half:
            h = h + self.h_size[0] + self.h_size[1]
        zc = z - self.d_start
        d = torch.sign(zc) * self._m2g(zc.abs(), self.d_size, self.d_range)
        if normalize:
            h = h / (self.size_h - 1)
            w = w / (self.size_w - 1)
            d = d / (self.size_d - 1)
        return torch.stack([h, w, d], -1)
