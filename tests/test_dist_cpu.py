"""CPU / gloo, world_size 2: the ray-sharding host logic (slice arithmetic + the single all_gather)."""
import os
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from selfocc_b200.dist import ray_slice, all_gather_rays


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
