"""Multi-GPU plumbing for the render path: rays are independent given the decoded
volume, so the flat (cam, ray) order of neus_head.py:324-325 is split into contiguous per-rank slices
(the same split ``torch.chunk`` produces) and the rendered maps are put back together with ONE
all_gather.  Backend-agnostic (``nccl`` on the GPUs, ``gloo`` in the CPU tests)."""
import torch
import torch.distributed as dist


def ray_slice(total, world_size, rank):
    """(begin, count) of this rank's contiguous slice; identical to torch.chunk(arange(total), world_size)."""
    per = -(-total // world_size)
    begin = min(rank * per, total)
    return begin, max(0, min(per, total - begin))


def all_gather_rays(local, total, group=None):
    """local [count, ...] of this rank -> [total, ...] in flat ray order on every rank.  One collective:
    slices are padded to the common per-rank length so a single all_gather_into_tensor suffices."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    per = -(-total // world)
    tail = local.shape[1:]
    buf = local
    if local.shape[0] != per:
        buf = local.new_zeros((per,) + tuple(tail))
        buf[:local.shape[0]] = local
    out = local.new_empty((world * per,) + tuple(tail))
    try:
        dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):        # older gloo builds: list form, still one collective
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf.contiguous(), group=group)
        out = torch.cat(parts, 0)
    return out[:total]


def render_sharded(head, metas, batch=0, group=None):
    """NeuSHead.render with the frame's rays sharded over the process group; every rank returns the
    full maps (bit-identical to the single-GPU render: same kernel, same per-ray arithmetic)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    sampler = head._sampler()
    n_cam = head.img2lidar.matrices(metas, torch.device('cpu')).shape[1]
    total = n_cam * sampler.ray_number
    begin, count = ray_slice(total, world, rank)
    out = head.render(metas=metas, batch=batch, ray_range=(begin, count))
    keys = ['ms_depths', 'ms_accs'] + (['ms_max_depths'] if head.return_max_depth else [])
    res = {'ms_rays': out['ms_rays']}
    # one collective: pack the per-ray scalars side by side
    packed = torch.stack([out[k][0] for k in keys], -1)
    full = all_gather_rays(packed, total, group)
    for i, k in enumerate(keys):
        res[k] = [full[:, i].reshape(1, n_cam, sampler.ray_number)]
    return res


def uniform_sdf_sharded(head, aabb, resolution, group=None):
    """Occupancy lattice of NeuSHead.forward_occ / get_uniform_sdf (neus_head.py:265-293) with the lattice points
    sharded over the process group (BASELINE configs[3]): each rank queries a contiguous slice of the
    flattened [H, W, D] lattice and ONE all_gather assembles the sdf.  ``head.prepare()`` must have run on every rank."""
    from . import ops
    f = head.model.field
    dev = f.vol_sdf.device
    xs = torch.linspace(aabb[0], aabb[3], int((aabb[3] - aabb[0]) / resolution), device=dev)
    ys = torch.linspace(aabb[1], aabb[4], int((aabb[4] - aabb[1]) / resolution), device=dev)
    zs = torch.linspace(aabb[2], aabb[5], int((aabb[5] - aabb[2]) / resolution), device=dev)
    W, H, D = len(xs), len(ys), len(zs)
    xyz = torch.stack([xs[None, :, None].expand(H, W, D), ys[:, None, None].expand(H, W, D),
                       zs[None, None, :].expand(H, W, D)], dim=-1).flatten(0, 2)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    total = xyz.shape[0]
    begin, count = ray_slice(total, world, rank)
    local = ops.field_query(f.vol_sdf, f.vol_feat, f.desc, xyz[begin:begin + count].contiguous())[0]
    return all_gather_rays(local, total, group).reshape(H, W, D), xyz.reshape(H, W, D, 3)
