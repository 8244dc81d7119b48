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


class _FakeEngine:
    """stands in for NetEngine in backward_patches: the 'gradient' of a patch is a per-ray function summed over the patch"""
    def __init__(self):
        self.calls = []

    def hold_pack(self):
        import contextlib
        return contextlib.nullcontext(self)

    def grad_zero(self):
        self._gpack = torch.zeros(4, dtype=torch.float64); self._gscal = torch.zeros(2, dtype=torch.float64)

    def render_bwd(self, ro, rd, scal, fwd, g, **kw):
        self.calls.append((int(ro.shape[0]), kw['eikonal_count']))
        self._gpack += torch.stack([ro.double().sum(), rd.double().sum(), (g.double() * ro.double()).sum(), torch.tensor(float(ro.shape[0]), dtype=torch.float64)])
        # second scalar: the number of reference patches this launch stands for (rays x samples / the per-patch eikonal normaliser)
        self._gscal += torch.tensor([g.double().sum(), ro.shape[0] * fwd['d_vals'].shape[-1] / kw['eikonal_count']], dtype=torch.float64)

    def unpack_grads(self, ts, tr):
        return [], self._gscal


class _FakeModel:
    def __init__(self):
        self.implicit_surface = torch.nn.Linear(1, 1); self.radiance_net = torch.nn.Linear(1, 1)
        self.ln_beta = torch.nn.Parameter(torch.zeros(1)); self.speed_factor = 1.0
        self._e = _FakeEngine()

    def engine(self):
        return self._e


def _train_worker(rank, world, port, n, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from nerfart_b200.models.frameworks import _finetune
    g = torch.Generator().manual_seed(5)
    ro = torch.randn(1, n, 3, generator=g); rd = torch.randn(1, n, 3, generator=g); G = torch.randn(1, n, 3, generator=g)
    m = _FakeModel()
    fake_patch = lambda a, b: ({'d_vals': torch.zeros(a.shape[0], 6)}, None)
    scal, _ = _finetune.backward_patches(m, 'volsdf', ro, rd, G, fake_patch, w_eikonal=0.1, white_bkgd=False, batch_size=100)
    q.put((rank, m._e._gpack.tolist(), scal.tolist(), m._e.calls, m.ln_beta.grad.tolist()))
    dist.destroy_process_group()


def test_training_patches_round_robin_and_gradient_allreduce_gloo():
    """N>1 training path: patches dealt round-robin, packed gradient + scalars summed once; every rank ends with the
    single-process result (the per-patch eikonal normaliser is unchanged by the partition)."""
    world, n = 2, 730                       # 8 patches of 100 rays, the last one short
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    g = torch.Generator().manual_seed(5)
    ro = torch.randn(1, n, 3, generator=g); rd = torch.randn(1, n, 3, generator=g); G = torch.randn(1, n, 3, generator=g)
    want = [ro.double().sum().item(), rd.double().sum().item(), (G.double() * ro.double()).sum().item(), float(n)]
    for rank, gpack, scal, calls, lnb in res:
        assert all(abs(a - b) < 1e-9 for a, b in zip(gpack, want)), (gpack, want)
        assert abs(scal[0] - G.double().sum().item()) < 1e-9 and scal[1] == 8.0
        # launch groups (default NA_PATCH_GROUP = 6): rank 0 holds four full patches -> one launch; rank 1 three full + the short one
        assert calls == ([(400, 600)] if rank == 0 else [(300, 600), (30, 180)]), calls
        assert abs(lnb[0] - G.double().sum().item()) < 1e-4


def test_patch_groups():
    """_finetune.patch_groups: even groups of at most NA_PATCH_GROUP full patches, a short last patch alone, group 1 = the
    reference's launch structure (volsdf.py:754-783)."""
    from nerfart_b200.models.frameworks._finetune import patch_groups
    starts = list(range(0, 129600, 1200))
    g6 = patch_groups(starts, 129600, 1200, group=6)
    assert len(g6) == 18 and all(len(g) == 6 for g in g6) and sum(g6, []) == starts
    assert patch_groups(starts, 129600, 1200, group=1) == [[i] for i in starts]
    assert [len(g) for g in patch_groups(starts[3::8], 129600, 1200, group=6)] == [5, 5, 4]      # rank 3 of 8: 14 patches
    assert [len(g) for g in patch_groups(starts[4::8], 129600, 1200, group=6)] == [5, 4, 4]      # rank 4 of 8: 13 patches
    assert patch_groups([0, 100, 200], 250, 100, group=6) == [[0, 100], [200]]
    assert patch_groups([], 250, 100, group=6) == []


def _sync_worker(rank, world, port, q):
    import random
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(100 + rank); random.seed(200 + rank)               # ranks start out of step, as after a DistributedSampler
    c2w = torch.eye(4)[None] * (rank + 1); gt = torch.full((1, 12, 3), float(rank))
    parallel.sync_step_inputs([c2w, gt])
    draws = (random.choice(range(1000)), random.sample(range(1000), 8), torch.randint(0, 1000, (5,)).tolist())
    q.put((rank, c2w[0, 0, 0].item(), gt.mean().item(), draws))
    dist.destroy_process_group()


def test_finetune_step_inputs_and_rng_are_identical_on_every_rank_gloo():
    """the multi-rank fine-tune step equals the single-GPU step only if every rank scores the same image with the same
    random draws: rank 0's camera / target are broadcast and all ranks are seeded from one value."""
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sync_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] == res[1][1] == 1.0 and res[0][2] == res[1][2] == 0.0          # rank 0's tensors everywhere
    assert res[0][3] == res[1][3]                                                    # identical draws


def test_tile_partition_covers_every_ray_once_and_is_balanced():
    for (H, W, world) in ((480, 270, 8), (960, 540, 8), (64, 64, 3), (37, 9, 2), (5, 5, 4)):
        order, off = parallel.tile_partition(H, W, world)
        assert sorted(order.tolist()) == list(range(H * W)) and off[0] == 0 and off[-1] == H * W
        counts = [off[r + 1] - off[r] for r in range(world)]
        if H * W >= 64 * 64:
            assert max(counts) - min(counts) <= 16 * 16 * 2 + 16 * max(H, W) // 4, counts     # within ~2 tiles (+ ragged edge tiles)
        assert parallel.rank_rays(H * W, 0, world, H, W, mode='block') is None
        if world > 1:
            got = parallel.rank_rays(H * W, 1, world, H, W, mode='tiles')
            assert torch.equal(got, order[off[1]:off[2]])
            # a rank's rays are spread over the whole image (what balances the sampler load), not one band
            rows = (got // W).float()
            if H >= 64:
                assert rows.min() < H * 0.25 and rows.max() > H * 0.75


def _tiles_worker(rank, world, port, H, W, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    n = H * W
    g = torch.Generator().manual_seed(5)
    ro = torch.randn(n, 3, generator=g); rd = torch.randn(n, 3, generator=g)
    idx = parallel.rank_rays(n, rank, world, H, W, mode='tiles')
    mine = _fake_render(ro[idx], rd[idx])[0]
    full = parallel.gather_rays(mine, n, idx, H, W)
    q.put((rank, bool(torch.equal(full, _fake_render(ro, rd)[0]))))
    dist.destroy_process_group()


def test_interleaved_tile_gather_equals_single_rank_gloo():
    world, H, W = 2, 37, 21
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tiles_worker, args=(r, world, port, H, W, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
