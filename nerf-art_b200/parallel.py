"""Multi-GPU plumbing for the render path: one process per GPU, rays block-partitioned by rank, one all-gather of the
rendered tiles.  Rays are independent units -- every op of the renderer is per ray and the kernels are
deterministic per ray (tests/test_gpu_parity.py::test_volsdf_render_is_deterministic_and_ray_independent) -- so the
partition needs no data-path collective other than the final gather, and the gathered image is bit-identical to the
single-GPU image.  Replaces the reference's `nn.DataParallel(self.renderer, dim=1)` (models/frameworks/volsdf.py:632-633).
"""
import torch
import torch.distributed as dist


def world_rank(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def allreduce_sum(*tensors, group=None):
    """Sum the packed gradient accumulators over the ranks (one collective per tensor per training step)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in tensors:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def ray_block(n_rays, rank, world):
    """Contiguous block [lo, hi) of rank `rank` in row-major pixel order; every block has ceil(n/world) slots (the last
    ones may be short or empty) so the gathered buffer *is* the image."""
    per = (n_rays + world - 1) // world
    lo = min(rank * per, n_rays)
    hi = min(lo + per, n_rays)
    return lo, hi, per


def gather_tiles(tile, n_rays, group=None):
    """All-gather per-rank tiles [hi-lo, C] into the full [n_rays, C] tensor on every rank (NCCL on GPUs; gloo in the
    CPU tests).  Padding rows of short blocks are dropped."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return tile
    rank = dist.get_rank(group)
    lo, hi, per = ray_block(n_rays, rank, world)
    send = tile.new_zeros((per,) + tuple(tile.shape[1:]))
    send[:hi - lo] = tile
    out = tile.new_empty((world * per,) + tuple(tile.shape[1:]))
    dist.all_gather_into_tensor(out, send, group=group)
    return out[:n_rays]


def render_sharded(render_fn, rays_o, rays_d, keys=('rgb',), group=None, **kwargs):
    """Render this rank's block of `rays_o/rays_d` [N,3] with `render_fn` (the reference-shaped `(rgb, depth, extras)`
    callable) and return the gathered full-frame tensors for `keys` of the extras dict."""
    n = rays_o.shape[0]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi, _ = ray_block(n, rank, world)
    _, _, extras = render_fn(rays_o[lo:hi], rays_d[lo:hi], **kwargs)
    return {k: gather_tiles(extras[k].reshape(hi - lo, -1), n, group) for k in keys}


def sync_step_inputs(tensors, group=None, seed_rng=True):
    """Make one fine-tune step identical on every rank (world > 1): broadcast the step's input tensors (camera matrices, target
    image) from rank 0 in place, and seed Python's `random` and torch's CPU / CUDA generators on EVERY rank with one value drawn
    by rank 0, so that the style losses' draws (random.choice / random.sample of negative prompts, volsdf.py:903-909; PatchNCE's
    torch.randint crops, patchnce_loss.py:196-212) are the same everywhere.

    Why: ranks render disjoint ray blocks of ONE image and score the gathered frame; the patch gradients are then summed as if
    they came from one loss.  That only equals the single-GPU step when every rank holds the same camera, the same target and the
    same image gradient.  The reference's `train.py --ddp` gives each rank a different image through DistributedSampler
    (train.py:83-87) -- under nerfart_b200's ray partition the sampler must not shard (rank 0's image wins here).
    No-op at world size 1 (the RNG streams of a single-GPU run are left untouched)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tensors
    for t in tensors:
        dist.broadcast(t, src=0, group=group)
    if seed_rng:
        import random
        seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64)
        dev = tensors[0].device if tensors else torch.device('cpu')
        seed_d = seed.to(dev)
        dist.broadcast(seed_d, src=0, group=group)
        s = int(seed_d.item())
        random.seed(s)
        torch.manual_seed(s)                       # CPU and all CUDA generators
    return tensors


# ---------------------------------------------------------------------------------------------------------------------
# Interleaved partition for data-dependent sampler load: the error-bound sampler spends 1..7x the work on a ray
# depending on where it converges (iter_usage 0..6 / -1), and that varies smoothly over the image -- a contiguous block can be
# all background or all silhouette.  Pixel tiles of `tile` x `tile` are dealt round-robin to the ranks instead; the gathered
# buffer is un-permuted into the image.
# ---------------------------------------------------------------------------------------------------------------------
_TILE_CACHE = {}


def tile_partition(H, W, world, tile=16, device='cpu'):
    """-> (order [H*W] int64: ray indices grouped by owning rank, row-major inside a rank; offsets [world+1])."""
    key = (int(H), int(W), int(world), int(tile), str(device))
    hit = _TILE_CACHE.get(key)
    if hit is None:
        tx = (W + tile - 1) // tile
        tid = (torch.arange(H) // tile)[:, None] * tx + (torch.arange(W) // tile)[None, :]
        owner = (tid % world).reshape(-1)
        order = torch.argsort(owner, stable=True)
        counts = torch.bincount(owner, minlength=world)
        offsets = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(counts, 0)])
        hit = (order.to(device), [int(o) for o in offsets])
        _TILE_CACHE[key] = hit
    return hit


def rank_rays(n_rays, rank, world, H=None, W=None, mode=None, tile=16, device='cpu'):
    """Ray indices this rank renders: None (= the contiguous block of ray_block) or an index tensor (interleaved tiles).
    mode: 'block' | 'tiles' (default: $NA_PARTITION, else 'block')."""
    import os
    mode = mode or os.environ.get('NA_PARTITION', 'block')
    if mode == 'tiles' and world > 1 and H is not None and W is not None and H * W == n_rays:
        order, off = tile_partition(H, W, world, tile, device)
        return order[off[rank]:off[rank + 1]]
    return None


def gather_rays(vals, n_rays, idx, H=None, W=None, tile=16, group=None):
    """All-gather per-rank values [len(idx) or block, C] into [n_rays, C]; `idx` as returned by rank_rays (None = block)."""
    if idx is None:
        return gather_tiles(vals, n_rays, group)
    world = dist.get_world_size(group)
    order, off = tile_partition(H, W, world, tile, vals.device)
    per = max(off[r + 1] - off[r] for r in range(world))
    send = vals.new_zeros((per,) + tuple(vals.shape[1:]))
    send[:vals.shape[0]] = vals
    got = vals.new_empty((world * per,) + tuple(vals.shape[1:]))
    dist.all_gather_into_tensor(got, send, group=group)
    out = vals.new_empty((n_rays,) + tuple(vals.shape[1:]))
    for r in range(world):
        out[order[off[r]:off[r + 1]]] = got[r * per:r * per + off[r + 1] - off[r]]
    return out
