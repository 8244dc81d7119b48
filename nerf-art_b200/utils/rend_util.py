"""Ray utilities with the reference's names and signatures (reference utils/rend_util.py) -- every public symbol of that module, so
that `utils/rend_util.py` of a reference checkout can forward here (INTEGRATION.md level 1) without breaking `dataio/`.

Hot-path functions (get_rays, sample_pdf, sample_cdf) run as CUDA kernels through the C ABI and raise on CPU tensors.  The camera /
pose helpers the data loaders import (load_K_Rt_from_P, rot_to_quat, look_at ...; dataio/DTU.py:8, dataio/custom.py:10) are host-side
set-up code outside the path and stay plain numpy / torch.
"""
import ctypes as C
import numpy as np
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


# ------------------------------------------------------------------------------------------------------------------
# inverse-CDF samplers (reference rend_util.py:256-328) on the sampler kernels (csrc/sampler.cuh through na_sample_pdf / na_sample_cdf)
# ------------------------------------------------------------------------------------------------------------------
def _invert(fn_name, bins, wc, N_importance, det, u):
    if not bins.is_cuda:
        raise RuntimeError(f'nerfart_b200.utils.rend_util.{fn_name}: CUDA tensors only (there is no CPU path)')
    L = _lib.lib()
    dev = bins.device
    prefix = bins.shape[:-1]
    n = bins.shape[-1]
    b = bins.detach().reshape(-1, n).float().contiguous()
    w = wc.detach().reshape(-1, n - 1).float().contiguous()
    rows = b.shape[0]
    if u is None:
        if det:                                              # torch.linspace(0, 1, N) as the reference's oracle side computes it
            from ..engine import cpu_linspace
            u, per_row = cpu_linspace(N_importance, dev), 0
        else:
            u, per_row = torch.rand(rows, N_importance, device=dev, dtype=torch.float32), 1      # rend_util.py:272 / 307
    else:
        u = u.to(dev).float().contiguous()
        per_row = 0 if u.dim() == 1 else 1
        u = u.reshape(rows, N_importance) if per_row else u
    out = torch.empty(rows, N_importance, device=dev, dtype=torch.float32)
    fn = getattr(L, 'na_' + fn_name)
    with torch.cuda.device(dev):
        check(fn(ptr(b), ptr(w), rows, int(n), ptr(u), per_row, int(N_importance), ptr(out), None, stream_ptr(dev)), 'na_' + fn_name)
    return out.reshape(*prefix, N_importance)


def sample_pdf(bins, weights, N_importance, det=False, eps=1e-5, u=None):
    """reference rend_util.sample_pdf (256-293): bins [..., N], weights [..., N-1] -> samples [..., N_importance].
    `u` (extension) injects the uniform draws; eps is the reference's fixed 1e-5."""
    assert eps == 1e-5, 'the kernels implement the reference default eps=1e-5'
    return _invert('sample_pdf', bins, weights, N_importance, det, u)


def sample_cdf(bins, cdf, N_importance, det=False, eps=1e-5, u=None):
    """reference rend_util.sample_cdf (295-328): bins [..., N], (un-normalised) cdf [..., N-1] -> samples [..., N_importance]."""
    assert eps == 1e-5, 'the kernels implement the reference default eps=1e-5'
    return _invert('sample_cdf', bins, cdf, N_importance, det, u)


# ------------------------------------------------------------------------------------------------------------------
# geometry helpers of the callers (outside the hot path; tensor expressions)
# ------------------------------------------------------------------------------------------------------------------
def lift(x, y, z, intrinsics):
    """reference rend_util.lift (95-109): pixel (x, y) at depth z -> homogeneous camera-space point [..., 4] for intrinsics with
    skew.  (get_rays above does this inside its kernel; this stand-alone form is kept for callers that import it.)"""
    K = intrinsics.to(x.device)
    fx, fy = K[..., 0, 0].unsqueeze(-1), K[..., 1, 1].unsqueeze(-1)
    cx, cy = K[..., 0, 2].unsqueeze(-1), K[..., 1, 2].unsqueeze(-1)
    sk = K[..., 0, 1].unsqueeze(-1)
    x_cam = (x - cx + cy * sk / fy - sk * y / fy) / fx * z
    y_cam = (y - cy) / fy * z
    return torch.stack((x_cam, y_cam, z, torch.ones_like(z)), dim=-1)


def _ray_sphere_terms(ray_origins, ray_directions):
    oo = torch.sum(ray_origins ** 2, dim=-1, keepdim=True)
    od = torch.sum(ray_origins * ray_directions, dim=-1, keepdim=True)
    return oo, od


def get_sphere_intersection(ray_origins, ray_directions, r=1.0):
    """reference rend_util.get_sphere_intersection (189-211): (near, far, mask) [..., 1] of unit-direction rays with the
    origin-centred sphere of radius r; rays that miss get near = far = 0."""
    oo, od = _ray_sphere_terms(ray_origins, ray_directions)
    disc = od ** 2 + r ** 2 - oo
    hit = disc > 0
    root = torch.sqrt(torch.where(hit, disc, torch.zeros_like(disc)))
    zero = torch.zeros_like(disc)
    near = torch.where(hit, -root - od, zero).clamp_min(0.0)
    far = torch.where(hit, root - od, zero).clamp_min(0.0)
    return near, far, hit


def get_dvals_from_radius(ray_origins, ray_directions, rs, far_end=True):
    """reference rend_util.get_dvals_from_radius (214-235): depth at which a unit-direction ray is at distance rs from the origin."""
    oo, od = _ray_sphere_terms(ray_origins, ray_directions)
    disc = rs ** 2 - (oo - od ** 2)
    assert (disc > 0).all()
    root = torch.sqrt(disc)
    return -od + root if far_end else torch.clamp_min(-od - root, 0.)


# ------------------------------------------------------------------------------------------------------------------
# camera / pose helpers used by dataio/ and render.py's camera paths (host-side numpy / torch; reference rend_util.py:8-92)
# ------------------------------------------------------------------------------------------------------------------
def load_K_Rt_from_P(P):
    """reference rend_util.load_K_Rt_from_P (8-25): 3x4 projection matrix -> (4x4 intrinsics, 4x4 camera-to-world pose) through
    cv2.decomposeProjectionMatrix, K normalised by K[2,2]."""
    import cv2
    K, R, t = cv2.decomposeProjectionMatrix(P)[:3]
    intrinsics = np.eye(4)
    intrinsics[:3, :3] = K / K[2, 2]
    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3] = R.T
    pose[:3, 3] = (t[:3] / t[3])[:, 0]
    return intrinsics, pose


def normalize(vec):
    """reference rend_util.normalize (27-28)."""
    return vec / (np.linalg.norm(vec, axis=-1, keepdims=True) + 1e-9)


def view_matrix(forward, up, cam_location):
    """reference rend_util.view_matrix (30-42): camera-to-world from a forward / up pair (columns x, y, z, location)."""
    z = normalize(forward)
    x = normalize(np.cross(up, z))
    y = normalize(np.cross(z, x))
    mat = np.stack((x, y, z, cam_location), axis=-1)
    last = np.array([[0., 0., 0., 1.]])
    if mat.ndim > 2:
        last = np.tile(last, [mat.shape[0], 1, 1])
    return np.concatenate((mat, last), axis=-2)


def look_at(cam_location, point, up=np.array([0., -1., 0.])):
    """reference rend_util.look_at (44-53): OpenCV convention, the camera looks along +z."""
    return view_matrix(normalize(point - cam_location), up, cam_location)


def rot_to_quat(R):
    """reference rend_util.rot_to_quat (55-73): [B,3,3] rotation -> [B,4] quaternion (w, x, y, z); the trace-based branch only,
    like the reference."""
    w = torch.sqrt(1.0 + R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]) / 2
    q = torch.ones((R.shape[0], 4)).to(R.device)
    q[..., 0] = w
    q[..., 1] = (R[..., 2, 1] - R[..., 1, 2]) / (4 * w)
    q[..., 2] = (R[..., 0, 2] - R[..., 2, 0]) / (4 * w)
    q[..., 3] = (R[..., 1, 0] - R[..., 0, 1]) / (4 * w)
    return q


def quat_to_rot(q):
    """reference rend_util.quat_to_rot (76-92): [B,4] quaternion (normalised here) -> [B,3,3]."""
    q = torch.nn.functional.normalize(q, dim=-1)
    r, i, j, k = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    rows = [torch.stack([1 - 2 * (j ** 2 + k ** 2), 2 * (j * i - k * r), 2 * (i * k + r * j)], -1),
            torch.stack([2 * (j * i + k * r), 1 - 2 * (i ** 2 + k ** 2), 2 * (j * k - i * r)], -1),
            torch.stack([2 * (k * i - j * r), 2 * (j * k + i * r), 1 - 2 * (i ** 2 + j ** 2)], -1)]
    return torch.stack(rows, -2)
