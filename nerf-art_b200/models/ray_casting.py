"""Surface rendering behind the reference's API (reference models/ray_casting.py), computed by the sm_100a kernels.

Kept from the reference: `surface_render(rays_o, rays_d, model, ...) -> (colors, depths, extras)` with extras keys
`implicit_nablas`, `mask_surface`, `normals_surface` (ray_casting.py:187-263), `root_finding_surface_points` (35-160) and
`sphere_tracing_surface_points` (163-184) with their return tuples, batched `[B,N,3]` / unbatched `[N,3]` ray layouts.
The surface query is the model's own ImplicitSurface (what render.py passes); arbitrary Python callables are not supported --
there is no PyTorch fallback.
"""
from collections import OrderedDict

import torch

from .base import ImplicitSurface


def _engine_of(surface):
    if isinstance(surface, ImplicitSurface):
        owner = surface._owner() if getattr(surface, '_owner', None) is not None else None
        if owner is not None:
            return owner.engine()
        return surface.engine()
    if hasattr(surface, 'engine'):
        return surface.engine()
    raise NotImplementedError('surface_query_fn must be the ImplicitSurface of a nerfart_b200 model (no Python-callable path)')


def _flat(rays_o, rays_d, batched):
    if batched:
        if rays_o.shape[0] != 1:
            raise NotImplementedError('batch size > 1 is not used by any shipped config')
        return rays_o[0].float().contiguous(), rays_d[0].float().contiguous()
    return rays_o.float().contiguous(), rays_d.float().contiguous()


def root_finding_surface_points(surface_query_fn, rays_o, rays_d, near=0.0, far=6.0, batched=True, batched_info={}, N_steps=256,
                                logit_tau=0.0, method='secant', N_secant_steps=8, fill_inf=True):
    """-> (d_pred_out [(B),N], pt_pred [(B),N,3], mask, mask_sign_change).  rays_d must already be normalised (ray_casting.py:58)."""
    if method != 'secant':
        raise NotImplementedError("only method='secant' is implemented (the reference's default)")
    ro, rd = _flat(rays_o, rays_d, batched)
    with torch.no_grad():
        d, pt, m, msc = _engine_of(surface_query_fn).ray_cast(ro, rd, 'root_finding', near=near, far=far, N_steps=N_steps,
                                                              N_secant_steps=N_secant_steps, logit_tau=logit_tau, fill_inf=fill_inf)
    if batched:
        d, pt, m, msc = d[None], pt[None], m[None], msc[None]
    return d, pt, m, msc


def sphere_tracing_surface_points(implicit_surface, rays_o, rays_d, near=0.0, far=6.0, batched=True, batched_info={}, N_iters=20):
    """-> (d_preds, pts, mask)  (ray_casting.py:163-184)."""
    ro, rd = _flat(rays_o, rays_d, batched)
    with torch.no_grad():
        d, pt, m, _ = _engine_of(implicit_surface).ray_cast(ro, rd, 'sphere_tracing', near=near, far=far, N_iters=N_iters)
    if batched:
        d, pt, m = d[None], pt[None], m[None]
    return d, pt, m


def surface_render(rays_o, rays_d, model, calc_normal=True, rayschunk=8192, netchunk=1048576, batched=True, use_view_dirs=True,
                   show_progress=False, ray_casting_algo='', ray_casting_cfgs={}, **not_used_kwargs):
    """Same contract as the reference's surface_render (ray_casting.py:187-263); `rayschunk` / `netchunk` are accepted and
    ignored (the fused kernels have no activation memory to bound)."""
    if not use_view_dirs:
        raise NotImplementedError('use_view_dirs=False is not a shipped configuration')
    if ray_casting_algo not in ('root_finding', 'sphere_tracing'):
        raise NotImplementedError                                   # ray_casting.py:233
    if batched:
        B = rays_d.shape[0]
        ro, rd = torch.reshape(rays_o, [B, -1, 3]), torch.reshape(rays_d, [B, -1, 3])
    else:
        ro, rd = torch.reshape(rays_o, [-1, 3]), torch.reshape(rays_d, [-1, 3])
    ro, rd = _flat(ro, rd, batched)
    with torch.no_grad():
        o = model.engine().surface_render(ro, rd, ray_casting_algo, calc_normal=calc_normal, **ray_casting_cfgs)
    lead = (lambda t: t[None]) if batched else (lambda t: t)
    extras = OrderedDict([('implicit_nablas', lead(o['nablas'])), ('mask_surface', lead(o['mask']))])
    if calc_normal:
        extras['normals_surface'] = lead(o['normals'])
    return lead(o['rgb']), lead(o['depth']), extras
