"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU (fp32).

Run ONLY in the build container (needs /root/reference):   python tests/golden/make_golden.py
The GPU box has no /root/reference; tests there read the committed .npz files.

What is pinned (the reference has no golden vectors of its own, so its executed
output on seeded synthetic state is the pin):
  stages.npz          per-stage in/out: Embedder, ImplicitSurface(+nablas), RadianceNet,
                      sdf_to_sigma, error_bound, sample_pdf, sample_cdf, get_rays, near_far_from_sphere,
                      NeuS sdf_to_alpha / alpha_to_w
  volsdf_*.npz        volume_render end to end (basic + detailed extras), three betas, two cameras
  neus_*.npz          NeuS volume_render end to end
  state_digest.json   sha256 of every tensor of the seeded reference models (init parity of the
                      product's nn.Module mirror) + key/shape layout
"""
import os, sys, json, hashlib, warnings
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
import ref_shim
ref_shim.install()
warnings.filterwarnings('ignore')
import numpy as np
import torch
import fixtures as fx

from models.base import get_embedder                                  # reference modules
from models.frameworks import volsdf as rvolsdf, neus as rneus
from utils import rend_util

torch.set_num_threads(os.cpu_count())


def npy(t):
    return t.detach().cpu().numpy()


def sha(t):
    return hashlib.sha256(np.ascontiguousarray(npy(t)).tobytes()).hexdigest()


def build_volsdf(beta_init, bump, seed=0):
    torch.manual_seed(seed)
    m = rvolsdf.VolSDF(**fx.volsdf_kwargs(beta_init))
    sd = m.state_dict(); fx.perturb_state_dict(sd, bump=bump); m.load_state_dict(sd)
    return m.eval()


def build_neus(variance_init, bump, seed=0):
    torch.manual_seed(seed)
    m = rneus.NeuS(**fx.neus_kwargs(variance_init))
    sd = m.state_dict(); fx.perturb_state_dict(sd, bump=bump); m.load_state_dict(sd)
    return m.eval()


def rays_for(cam, H, W):
    c2w, K = (fx.closed_form_camera if cam == 'closed' else fx.tilted_camera)(H, W)
    ro, rd, _ = rend_util.get_rays(c2w[None], K[None], H, W, -1)
    return c2w, K, ro, rd


def digest(out):
    d = {}
    torch.manual_seed(0); m = rvolsdf.VolSDF(**fx.volsdf_kwargs(0.1))
    d['volsdf_seed0_raw'] = {k: [list(v.shape), sha(v)] for k, v in m.state_dict().items()}
    m = build_volsdf(0.01, 0.5)
    d['volsdf_seed0_beta0.01_bump0.5'] = {k: [list(v.shape), sha(v)] for k, v in m.state_dict().items()}
    torch.manual_seed(0); m = rneus.NeuS(**fx.neus_kwargs(0.05))
    d['neus_seed0_raw'] = {k: [list(v.shape), sha(v)] for k, v in m.state_dict().items()}
    with open(os.path.join(out, 'state_digest.json'), 'w') as f:
        json.dump(d, f, indent=0)


def stages(out):
    g = torch.Generator().manual_seed(7)
    S = {}
    # Embedder (models/base.py:46-64)
    x = (torch.rand(64, 3, generator=g) * 6 - 3).float()
    e6, _ = get_embedder(6); e4, _ = get_embedder(4)
    S['emb_x'] = npy(x); S['emb6'] = npy(e6(x)); S['emb4'] = npy(e4(x))
    # networks
    for tag, m in [('v', build_volsdf(0.01, 0.5)), ('n', build_neus(0.05, 0.5))]:
        x = (torch.rand(512, 3, generator=g) * 3.4 - 1.7).float()
        v = torch.nn.functional.normalize(torch.randn(512, 3, generator=g), dim=-1)
        with torch.no_grad():
            sdf, feat = m.implicit_surface.forward(x, return_h=True)
        sdf2, nab, feat2 = m.implicit_surface.forward_with_nablas(x.clone())
        with torch.no_grad():
            rad = m.radiance_net.forward(x, v, nab, feat2)
        S[f'net_{tag}_x'] = npy(x); S[f'net_{tag}_v'] = npy(v)
        S[f'net_{tag}_sdf'] = npy(sdf); S[f'net_{tag}_feat'] = npy(feat); S[f'net_{tag}_nabla'] = npy(nab)
        S[f'net_{tag}_rad'] = npy(rad)
        if tag == 'v':
            with torch.no_grad():
                S['net_v_surface'] = npy(m.forward_surface(x)[0])
                r, s, nb = m.forward(x, v)
                S['net_v_fwd_rad'] = npy(r); S['net_v_fwd_sdf'] = npy(s)
                a, b = m.forward_ab(); S['net_v_ab'] = np.array([a.item(), b.item()], dtype=np.float32)
        else:
            S['net_n_s'] = np.array([m.forward_s().item()], dtype=np.float32)
    # sdf_to_sigma / error_bound (volsdf.py:34-94)
    M, N = 12, 96
    d = torch.sort(torch.rand(M, N, generator=g) * 6, dim=-1).values.float()
    sdf = (torch.rand(M, N, generator=g) * 2 - 0.7).float() * torch.linspace(0.02, 1.0, M)[:, None]
    for i, (a, b) in enumerate([(10.0, 0.1), (100.0, 0.01), (500.0, 0.002)]):
        S[f'sigma_{i}'] = npy(rvolsdf.sdf_to_sigma(sdf, a, b))
        S[f'ebound_{i}'] = npy(rvolsdf.error_bound(d, sdf, a, b))
    beta_row = (torch.rand(M, 1, generator=g) * 0.2 + 0.003).float()
    S['eb_d'] = npy(d); S['eb_sdf'] = npy(sdf); S['eb_beta_row'] = npy(beta_row)
    S['ebound_row'] = npy(rvolsdf.error_bound(d, sdf, 1. / beta_row, beta_row))
    # sample_pdf / sample_cdf (rend_util.py:256-328)
    w = rvolsdf.error_bound(d, sdf, 1. / beta_row, beta_row).clamp(0, 1e5)
    S['spdf_w'] = npy(w)
    S['spdf_det'] = npy(rend_util.sample_pdf(d, w, 50, det=True))
    cdf = 1 - torch.exp(-torch.cumsum(torch.rand(M, N - 1, generator=g) * 0.08 * torch.linspace(0.05, 1, M)[:, None], -1))
    S['scdf_cdf'] = npy(cdf)
    S['scdf_det'] = npy(rend_util.sample_cdf(d, cdf, 16, det=True))
    u = torch.rand(M, 16, generator=g)
    orig = torch.rand
    torch.rand = lambda *a, **k: u.clone()
    try:
        S['scdf_rand'] = npy(rend_util.sample_cdf(d, cdf, 16, det=False))
        S['spdf_rand'] = npy(rend_util.sample_pdf(d, w, 16, det=False))
    finally:
        torch.rand = orig
    S['samp_u'] = npy(u)
    # rays (rend_util.py:95-186)
    for cam in ['closed', 'tilted']:
        c2w, K, ro, rd = rays_for(cam, 6, 5)
        S[f'rays_{cam}_c2w'] = npy(c2w); S[f'rays_{cam}_K'] = npy(K)
        S[f'rays_{cam}_o'] = npy(ro[0]); S[f'rays_{cam}_d'] = npy(rd[0])
    o = (torch.randn(32, 3, generator=g) * 0.8).float(); dd = torch.nn.functional.normalize(torch.randn(32, 3, generator=g), dim=-1)
    n_, f_ = rend_util.near_far_from_sphere(o, dd, r=1.0)
    S['nf_o'] = npy(o); S['nf_d'] = npy(dd); S['nf_near'] = npy(n_); S['nf_far'] = npy(f_)
    # NeuS alpha / weights (neus.py:36-78)
    cdf_, al = rneus.sdf_to_alpha(sdf, 20.0)
    S['neus_alpha_cdf'] = npy(cdf_); S['neus_alpha'] = npy(al); S['neus_w'] = npy(rneus.alpha_to_w(al))
    np.savez_compressed(os.path.join(out, 'stages.npz'), **S)


def run_volsdf(out, name, beta_init, bump, cam, H, W, Ns, Ni, detailed, perturb_u=None):
    m = build_volsdf(beta_init, bump)
    c2w, K, ro, rd = rays_for(cam, H, W)
    kw = dict(batched=True, near=0.0, far=6.0, obj_bounding_radius=3.0, perturb=perturb_u is not None,
              white_bkgd=False, max_upsample_steps=6, N_samples=Ns, N_importance=Ni, epsilon=0.1,
              max_bisection_steps=10, require_nablas=True, calc_normal=True, detailed_output=detailed, rayschunk=2048)
    orig = torch.rand
    if perturb_u is not None:
        u0 = torch.tensor(perturb_u)
        torch.rand = lambda shape, **k: u0.expand(list(shape)).clone()
    try:
        with torch.no_grad():
            rgb, depth, ex = rvolsdf.volume_render(ro, rd, m, **kw)
    finally:
        torch.rand = orig
    S = dict(rays_o=npy(ro[0]), rays_d=npy(rd[0]),
             meta=np.array([beta_init, bump, H, W, Ns, Ni], dtype=np.float64))
    if perturb_u is not None:
        S['u0'] = np.asarray(perturb_u, dtype=np.float32)
    for k, v in ex.items():
        S[k] = npy(v[0])
    np.savez_compressed(os.path.join(out, name + '.npz'), **S)
    iu = S.get('iter_usage')
    print(name, 'rgb', S['rgb'].min(), S['rgb'].max(),
          '' if iu is None else {int(k): int((iu == k).sum()) for k in np.unique(iu)})


def run_neus(out, name, var_init, bump, cam, H, W, detailed):
    m = build_neus(var_init, bump)
    c2w, K, ro, rd = rays_for(cam, H, W)
    ro = ro * 0.3                      # bring the camera near the r=1 bounding sphere (NeuS scenes are unit-sphere scaled)
    with torch.no_grad():
        rgb, depth, ex = rneus.volume_render(ro, rd, m, batched=True, obj_bounding_radius=1.0, perturb=False,
                                             white_bkgd=False, calc_normal=True, detailed_output=detailed,
                                             rayschunk=2048, upsample_algo='official_solution', N_nograd_samples=2048,
                                             N_upsample_iters=4, N_outside=0)
    S = dict(rays_o=npy(ro[0]), rays_d=npy(rd[0]), meta=np.array([var_init, bump, H, W], dtype=np.float64))
    for k, v in ex.items():
        S[k] = npy(v[0])
    np.savez_compressed(os.path.join(out, name + '.npz'), **S)
    print(name, 'rgb', S['rgb'].min(), S['rgb'].max(), 'acc', S['mask_volume'].min(), S['mask_volume'].max())


if __name__ == "__main__":
    out = HERE
    digest(out)
    stages(out)
    # BASELINE config 1 (plumbing): 64x64, 32+16, beta 0.1, basic outputs only
    run_volsdf(out, 'volsdf_cfg1_b0.1', 0.1, 0.0, 'closed', 64, 64, 32, 16, detailed=False)
    # detailed small cases, three betas (sampler takes 0..6 iterations / never converges)
    run_volsdf(out, 'volsdf_det_b0.1', 0.1, 0.5, 'tilted', 20, 20, 32, 16, detailed=True)
    run_volsdf(out, 'volsdf_det_b0.01', 0.01, 0.5, 'tilted', 20, 20, 32, 16, detailed=True)
    run_volsdf(out, 'volsdf_det_b0.002', 0.002, 0.5, 'tilted', 20, 20, 32, 16, detailed=True)
    # the real sample counts of BASELINE config 2 (128+64, d_init 512) on a few rays
    run_volsdf(out, 'volsdf_n128_b0.01', 0.01, 0.5, 'tilted', 12, 12, 128, 64, detailed=True)
    run_volsdf(out, 'volsdf_n128_b0.1', 0.1, 0.0, 'closed', 12, 12, 128, 64, detailed=False)
    # stratified (perturb=True) final sampling with an injected uniform row
    u0 = np.random.RandomState(3).rand(16).astype(np.float32)
    run_volsdf(out, 'volsdf_perturb_b0.01', 0.01, 0.5, 'tilted', 16, 16, 32, 16, detailed=True, perturb_u=u0)
    run_neus(out, 'neus_det', 0.05, 0.5, 'tilted', 20, 20, detailed=True)
    run_neus(out, 'neus_basic', 0.05, 0.0, 'closed', 40, 40, detailed=False)
