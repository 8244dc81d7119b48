"""world_size-2 gloo test of the N>1 host path (ray partition + tile all-gather) with a stand-in per-ray renderer.
The CUDA renderer itself is per-ray deterministic (GPU test), so exactness of the partition is a host-side property."""
import os
import socket
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import nerfart_b200  # noqa: F401
from nerfart_b200 import parallel


def _fake_render(ro, rd, **kw):
    # any per-ray function: result depends only on that ray
    rgb = torch.sin(ro * 3.0 + rd * 5.0)
    depth = (ro * rd).sum(-1)
    return rgb, depth, {'rgb': rgb, 'depth_volume': depth, 'mask_volume': torch.cos(depth)}


def _worker(rank, world, port, n, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    ro = torch.randn(n, 3, generator=g); rd = torch.randn(n, 3, generator=g)
    out = parallel.render_sharded(_fake_render, ro, rd, keys=('rgb', 'depth_volume'))
    ref = _fake_render(ro, rd)[2]
    ok = torch.equal(out['rgb'], ref['rgb']) and torch.equal(out['depth_volume'][:, 0], ref['depth_volume'])
    lo, hi, per = parallel.ray_block(n, rank, world)
    q.put((rank, bool(ok), lo, hi, per))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def test_ray_blocks_cover_everything_once():
    for n in (1, 7, 129600, 518400):
        for world in (1, 2, 3, 4, 8):
            seen = 0
            for r in range(world):
                lo, hi, per = parallel.ray_block(n, r, world)
                assert 0 <= lo <= hi <= n and hi - lo <= per
                assert lo == min(r * per, n)
                seen += hi - lo
            assert seen == n


def test_sharded_render_equals_single_rank_gloo():
    world, n = 2, 1001                      # odd: the last block is short
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, *_ in res), res
