"""Generate tests/golden/train_*.npz: parameter gradients of the fine-tune step's second pass, from the UNMODIFIED reference.

Run ONLY in the build container (needs /root/reference):   python tests/golden/make_golden_train.py

What is executed is exactly what `Trainer.forward` does per ray patch in its fine-tune branch
(models/frameworks/volsdf.py:769-783, models/frameworks/neus.py:551-563):

    rgb_pred, _, extras = volume_render(rays_o_patch, rays_d_patch, model, detailed_output=True, require_nablas=True, ...)
    rgb_pred.backward(gradient_patch, retain_graph=True)
    eikonal_loss = w_eikonal * mse(||extras['implicit_nablas']||, 1)          # calc_eikonal_loss, volsdf.py:917-939
    eikonal_loss.backward()

on seeded synthetic state (tests/fixtures.py) and a seeded image gradient.  The Trainer class itself cannot be constructed
offline (it loads CLIP and VGG weights in __init__, volsdf.py:639-642), so its loop body is restated here line by line.
Stored: the sample depths the reference used (so that the backward can be checked independently of the sampler), the image
gradient, rgb, the eikonal loss value, and d loss / d parameter for every parameter (state-dict key -> array).
"""
import os, sys, warnings
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
import ref_shim
ref_shim.install()
warnings.filterwarnings('ignore')
import numpy as np
import torch
import torch.nn.functional as F
import fixtures as fx
from make_golden import build_volsdf, build_neus, rays_for, npy     # noqa: E402  (reference-side builders)
from models.frameworks import volsdf as rvolsdf, neus as rneus       # noqa: E402

W_EIKONAL = 0.1          # configs/volsdf_fangzhou_vangogh.yaml:47, configs/neus_fangzhou_vangogh.yaml:94


def seeded_gradient(n, seed=7):
    g = torch.Generator(device='cpu'); g.manual_seed(seed)
    return (0.05 * torch.randn(1, n, 3, generator=g, dtype=torch.float32))


def eikonal(extras):
    nablas = extras['implicit_nablas'].flatten(-3, -2)
    nn_ = torch.norm(nablas, dim=-1)
    return W_EIKONAL * F.mse_loss(nn_, nn_.new_ones(nn_.shape), reduction='mean')


def grads_of(m, full=False):
    """Small tensors in full; 256x256-class weight_v gradients as every 4th row plus row / column sums (keeps the
    fixtures small; `full=True` keeps everything)."""
    out = {}
    for k, p in m.named_parameters():
        g = npy(p.grad) if p.grad is not None else np.zeros(tuple(p.shape), np.float32)
        if full or g.size <= 16384:
            out['grad.' + k] = g
        else:
            out['gradrows4.' + k] = g[::4].copy()
            out['gradsum0.' + k] = g.sum(0, dtype=np.float64).astype(np.float32)
            out['gradsum1.' + k] = g.sum(1, dtype=np.float64).astype(np.float32)
    return out


def volsdf_case(name, beta, bump, H, W, N_samples, N_importance, white_bkgd=False, use_eikonal=True, full=False):
    m = build_volsdf(beta, bump).train()
    _, _, ro, rd = rays_for('tilted', H, W)
    G = seeded_gradient(H * W)
    m.zero_grad()
    rgb, _, ex = rvolsdf.volume_render(ro, rd, m, batched=True, near=0.0, far=6.0, obj_bounding_radius=3.0, perturb=False,
                                       white_bkgd=white_bkgd, max_upsample_steps=6, N_samples=N_samples, N_importance=N_importance,
                                       detailed_output=True, require_nablas=True, use_view_dirs=True, rayschunk=4096)
    rgb.backward(G, retain_graph=True)
    eik = torch.zeros(())
    if use_eikonal:
        eik = eikonal(ex)
        eik.backward()
    out = dict(rays_o=npy(ro[0]), rays_d=npy(rd[0]), G=npy(G[0]), rgb=npy(rgb[0]), d_vals=npy(ex['d_vals'][0]),
               sdf=npy(ex['implicit_surface'][0]), nablas=npy(ex['implicit_nablas'][0]), radiance=npy(ex['radiance'][0]),
               eikonal_loss=np.float32(eik.item()), w_eikonal=np.float32(W_EIKONAL if use_eikonal else 0.0),
               white_bkgd=np.int32(white_bkgd), N_samples=np.int32(N_samples), N_importance=np.int32(N_importance),
               beta_init=np.float32(beta), bump=np.float32(bump))
    out.update(grads_of(m, full))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    gn = {k: float(np.abs(v).max()) for k, v in out.items() if k.startswith('grad')}
    print(name, 'eik', float(eik), 'max|grad|', max(gn.values()), 'ln_beta grad', out['grad.ln_beta'])


def _neus_render(ro, rd, m, N_samples, N_importance):
    return rneus.volume_render(ro, rd, m, batched=True, obj_bounding_radius=1.0, perturb=False, white_bkgd=False,
                               upsample_algo='official_solution', N_nograd_samples=2048, N_upsample_iters=4, N_outside=0,
                               N_samples=N_samples, N_importance=N_importance, detailed_output=True, require_nablas=True,
                               use_view_dirs=True, rayschunk=4096)


def neus_case(name, var, bump, H, W, N_samples, N_importance):
    m = build_neus(var, bump).train()
    m.fix_module('radiance_net')                                     # neus.py:28,455-456
    c2w, K = fx.tilted_camera(H, W)
    c2w = c2w.clone(); c2w[:3, 3] *= 0.6                             # keep the unit sphere in view
    from utils import rend_util
    ro, rd, _ = rend_util.get_rays(c2w[None], K[None], H, W, -1)
    G = seeded_gradient(H * W)
    m.zero_grad()
    sorted_rec = []                                                  # d_all is not among the extras: record the last torch.sort
    orig_sort = torch.sort

    def rec_sort(*a, **k):
        r = orig_sort(*a, **k)
        sorted_rec.append(r[0].detach().clone())
        return r
    torch.sort = rec_sort
    try:
        rgb, _, ex = _neus_render(ro, rd, m, N_samples, N_importance)
    finally:
        torch.sort = orig_sort
    d_all = npy(sorted_rec[-1][0])
    assert np.array_equal(0.5 * (d_all[:, 1:] + d_all[:, :-1]), npy(ex['d_final'][0]))
    rgb.backward(G, retain_graph=True)
    eik = eikonal(ex)
    eik.backward()
    d_mid = npy(ex['d_final'][0])
    out = dict(rays_o=npy(ro[0]), rays_d=npy(rd[0]), G=npy(G[0]), rgb=npy(rgb[0]), d_mid=d_mid, d_all=d_all,
               sdf=npy(ex['implicit_surface'][0]), nablas=npy(ex['implicit_nablas'][0]), radiance=npy(ex['radiance'][0]),
               alpha=npy(ex['alpha'][0]), eikonal_loss=np.float32(eik.item()), w_eikonal=np.float32(W_EIKONAL),
               N_samples=np.int32(N_samples), N_importance=np.int32(N_importance), variance_init=np.float32(var),
               bump=np.float32(bump))
    out.update(grads_of(m))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    gn = {k: float(np.abs(v).max()) for k, v in out.items() if k.startswith('grad')}
    print(name, 'eik', float(eik), 'max|grad|', max(gn.values()), 'ln_s grad', out['grad.ln_s'])


if __name__ == '__main__':
    torch.set_num_threads(os.cpu_count())
    volsdf_case('train_volsdf_b0.1', 0.1, 0.5, 6, 6, 32, 16, full=True)
    volsdf_case('train_volsdf_b0.01_white', 0.01, 0.5, 5, 5, 32, 16, white_bkgd=True)
    volsdf_case('train_volsdf_noeik', 0.1, 0.5, 4, 4, 32, 16, use_eikonal=False)
    neus_case('train_neus', 0.05, 0.5, 6, 6, 32, 16)
