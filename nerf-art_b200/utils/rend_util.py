"""Ray utilities with the reference's names and signatures (reference utils/rend_util.py)."""
import ctypes as C
import torch

from .. import _lib
from .._lib import check, ptr, stream_ptr


def get_rays(c2w, intrinsics, H, W, N_rays=-1):
    """reference rend_util.get_rays (112-165): c2w [(B,)4,4], intrinsics [(B,)4,4] -> rays_o, rays_d [(B,)H*W,3] (directions
    un-normalised), select_inds.  The pixel grid / lift / c2w product run in one kernel (csrc/api.cu:get_rays_kernel);
    N_rays > 0 selects a random subset afterwards exactly like the reference (randint h, randint w; lines 137-141)."""
    if c2w.shape[-1] == 7:
        raise NotImplementedError('quaternion poses are not used by the shipped data loaders')
    L = _lib.lib()
    prefix = c2w.shape[:-2]
    dev = c2w.device
    c2w_f = c2w.reshape(-1, 4, 4).float().contiguous()
    K_f = intrinsics.to(dev).reshape(-1, 4, 4).float().contiguous()
    nb = c2w_f.shape[0]
    ro = torch.empty(nb, H * W, 3, device=dev, dtype=torch.float32)
    rd = torch.empty(nb, H * W, 3, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        for b in range(nb):
            check(L.na_get_rays(ptr(c2w_f[b]), ptr(K_f[b if K_f.shape[0] > 1 else 0]), int(H), int(W), ptr(ro[b]), ptr(rd[b]),
                                stream_ptr(dev)), 'na_get_rays')
    if N_rays > 0:
        N_rays = min(N_rays, H * W)
        select_hs = torch.randint(0, H, size=[N_rays]).to(dev)
        select_ws = torch.randint(0, W, size=[N_rays]).to(dev)
        select_inds = select_hs * W + select_ws
        ro, rd = ro[:, select_inds], rd[:, select_inds]
        select_inds = select_inds.expand([*prefix, N_rays])
    else:
        select_inds = torch.arange(H * W, device=dev).expand([*prefix, H * W])
    return ro.reshape(*prefix, -1, 3), rd.reshape(*prefix, -1, 3), select_inds


def near_far_from_sphere(ray_origins, ray_directions, r=1.0, keepdim=True):
    """reference rend_util.near_far_from_sphere (168-186); inside the NeuS renderer this is fused (csrc/neus_render.cu)."""
    mid = -torch.sum(ray_origins * ray_directions, dim=-1, keepdim=keepdim)
    return (mid - r).clamp_min(0.0), (mid + r).clamp_min(r)


def lin2img(tensor, H, W, batched=False, B=None):
    """reference rend_util.lin2img (238-248)."""
    *_, num_samples, channels = tensor.shape
    assert num_samples == H * W
    if batched:
        if B is None:
            B = tensor.shape[0]
        else:
            tensor = tensor.view([B, num_samples // B, channels])
        return tensor.permute(0, 2, 1).view([B, channels, H, W])
    return tensor.permute(1, 0).view([channels, H, W])
