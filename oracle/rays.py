"""Oracle: pixel-ray generation (test infrastructure, see oracle/__init__.py).

Follows reference model/head/nerfacc_head/ray_sampler.py:23-68 (RaySampler),
model/head/nerfacc_head/img2lidar.py:25-70 (Img2LiDAR) and the ray preparation in
model/head/neus_head/neus_head.py:321-327.
"""
import math
import numpy as np
import torch


def fixed_ray_grid(ray_number, ray_img_size):
    """'fixed' sampler table (ray_sampler.py:23-31): row-major [ny*nx, 2] (x, y) pixels."""
    ny, nx = ray_number
    x = torch.arange(nx, dtype=torch.float) * (1.0 * ray_img_size[1] / nx)
    y = torch.arange(ny, dtype=torch.float) * (1.0 * ray_img_size[0] / ny)
    return torch.stack([x[None, :].expand(ny, -1), y[:, None].expand(-1, nx)], -1).flatten(0, 1)


def cellular_ray_grid(ray_number, ray_img_size, u4, ray_upper_crop=0, ray_x_dsr_max=None, ray_y_dsr_max=None):
    """'cellular' sampler (ray_sampler.py:32-46,58-68).  ``u4`` = the four host uniforms the
    reference draws from ``np.random.uniform`` in order (x stride, y stride, x offset, y offset)."""
    ny, nx = ray_number
    xm = 1.0 * ray_img_size[1] / nx if ray_x_dsr_max is None else ray_x_dsr_max
    ym = 1.0 * (ray_img_size[0] - ray_upper_crop) / ny if ray_y_dsr_max is None else ray_y_dsr_max
    assert xm > 1 and ym > 1
    x_dsr = u4[0] * (xm - 1) + 1
    y_dsr = u4[1] * (ym - 1) + 1
    x_emp = u4[2] * (ray_img_size[1] - nx * x_dsr)
    y_emp = u4[3] * (ray_img_size[0] - ray_upper_crop - ny * y_dsr)
    x = torch.arange(nx, dtype=torch.float)
    y = torch.arange(ny, dtype=torch.float)
    rays = torch.stack([x[None, :].expand(ny, -1), y[:, None].expand(-1, nx)], -1).clone()
    rays[..., 0] = rays[..., 0] * x_dsr + x_emp
    rays[..., 1] = rays[..., 1] * y_dsr + y_emp + ray_upper_crop
    return rays.flatten(0, 1)


def rot_z(deg):
    """dataset/utils.py:4-23 get_rm(angle, 'z', deg=True)."""
    a = np.deg2rad(deg)
    rm = np.eye(3)
    rm[0, 0] = rm[1, 1] = np.cos(a)
    rm[0, 1] = -np.sin(a)
    rm[1, 0] = np.sin(a)
    return rm


def img2lidar_rays(img2lidar, rays, novel_view=None):
    """img2lidar [B,N,4,4] fp32, rays [R,2] -> origin [B,N,3], direction [B,N,R,3] (un-normalised).
    img2lidar.py:51-70: optional novel view = z-rotation (deg) of the 3x3 block + xyz translation."""
    M = img2lidar.float().clone()
    rays = rays.float()
    if novel_view is not None:
        R = torch.as_tensor(rot_z(novel_view[3]), dtype=torch.float)
        M[..., :3, :3] = R[None, None] @ M[..., :3, :3]
    origin = M[..., :3, 3].clone()
    if novel_view is not None:
        origin[..., 0] += novel_view[0]
        origin[..., 1] += novel_view[1]
        origin[..., 2] += novel_view[2]
    pad = torch.cat([rays.reshape(1, 1, -1, 2), torch.ones(1, 1, rays.shape[0], 1, device=rays.device)], -1)
    direction = torch.matmul(M[..., :3, :3].unsqueeze(2), pad.unsqueeze(-1)).squeeze(-1)
    return origin, direction


def flatten_rays(origin, direction):
    """neus_head.py:322-327: (cam, ray)-major flattening, unit directions + their norms."""
    bs, n_cam, n_ray = direction.shape[:3]
    assert bs == 1
    o = origin.unsqueeze(2).repeat(1, 1, n_ray, 1).flatten(0, 2)
    d = direction.flatten(0, 2)
    nrm = torch.norm(d, dim=-1, keepdim=True)
    return o, d / nrm, nrm


def num_chunks(n_rays_total, batch):
    """neus_head.py:341-345: chunk count of the serial render loop (torch.chunk sizes follow)."""
    return int(math.ceil(n_rays_total * 1.0 / batch)) if batch > 0 else 1
