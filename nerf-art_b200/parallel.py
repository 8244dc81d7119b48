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
