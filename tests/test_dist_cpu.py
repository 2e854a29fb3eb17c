"""CPU / gloo, world_size 2: the ray-sharding host logic (slice arithmetic + the single all_gather)."""
import os
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from selfocc_b200.dist import ray_slice, all_gather_rays, all_gather_planar


def test_ray_slice_equals_torch_chunk():
    for total in (1, 7, 100, 2160000, 8640000, 13):
        for world in (1, 2, 3, 4, 8):
            chunks = torch.chunk(torch.arange(total), world)
            for r in range(world):
                b, c = ray_slice(total, world, r)
                if r < len(chunks):
                    assert (b, c) == (int(chunks[r][0]), len(chunks[r]))
                else:
                    assert c == 0


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    b, c = ray_slice(total, world, rank)
    local = torch.stack([torch.arange(b, b + c, dtype=torch.float32), -torch.arange(b, b + c, dtype=torch.float32)], -1)
    full = all_gather_rays(local, total)
    ok = torch.equal(full[:, 0], torch.arange(total, dtype=torch.float32)) and torch.equal(full[:, 1], -full[:, 0])
    # planar multi-tensor gather: a [count] and a [count, 3] tensor in one collective
    d = torch.arange(b, b + c, dtype=torch.float32)
    rgb = torch.stack([d, 2 * d, 3 * d], -1)
    fd, frgb = all_gather_planar([d, rgb], total)
    ok = ok and torch.equal(fd, torch.arange(total, dtype=torch.float32)) and torch.equal(frgb[:, 2], 3 * fd) and frgb.shape == (total, 3)
    q.put((rank, bool(ok), tuple(full.shape)))
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_rays_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    total = 1001  # odd: ranks hold 501 and 500 rays
    procs = [ctx.Process(target=_worker, args=(r, 2, 29731, total, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert res == [(0, True, (total, 2)), (1, True, (total, 2))]


def test_sharded_lifter_slices_round_trip():
    """Host logic of the query-sharded lifting (selfocc_b200/dist.py:ShardedLifter): per-plane contiguous slices, padding to the
    common per-rank length, and the reassembly of what all_gather_into_tensor delivers -- on the CPU, no kernels involved."""
    from selfocc_b200 import synth, configs
    from selfocc_b200.registry import build_head
    import selfocc_b200.segmentor  # noqa: F401
    from selfocc_b200.dist import ShardedLifter
    margs, rng = synth.small_mapping(6, 3)
    cfg = configs.hot_path_config(mapping_args=margs, pc_range=rng, num_cams=3, num_layers=1, num_points_cross=(4, 4, 4),
                                  num_points_self=4, num_samples=32, ray_number=(6, 10), ray_img_size=(90, 160))
    enc = build_head(cfg).encoder
    sl = ShardedLifter(enc)
    H, W, Z = enc.tpv_size
    assert sl.sizes == [H * W, Z * H, W * Z]
    q = torch.randn(sum(sl.sizes), 8)
    for world in (1, 2, 3, 5, 8):
        bufs = []
        covered = torch.zeros(sum(sl.sizes), dtype=torch.int32)
        for r in range(world):
            local = sl._local_rows(q, r, world)
            assert local.shape[0] == sum(c for _, c in sl.slices(r, world))
            bufs.append(sl.pad_local(local, r, world))
            off = 0
            for (b, c), n in zip(sl.slices(r, world), sl.sizes):
                covered[off + b:off + b + c] += 1
                off += n
        assert (covered == 1).all()                               # every query owned by exactly one rank
        assert all(b.shape[0] == sl.per_rank_rows(world) for b in bufs)
        assert torch.equal(sl.assemble(torch.stack(bufs, 0), world), q)
