"""Surface rendering (reference models/ray_casting.py): oracle vs the reference-generated golden vectors (CPU) and the CUDA path
vs the same vectors through the reference-shaped API (GPU).

The first-sign-change search and the secant bracket updates are threshold decisions on sdf values: a ray whose marched sdf passes
within the arithmetic error of zero may flip, so masks must agree on >= 99 % of the rays and values are compared on the rays whose
masks agree (the fixtures have no such ray for the oracle; the single-accumulator tensor-core mode carries a 1.5e-5 sdf error)."""
import numpy as np
import pytest
import torch

from helpers import golden, make_volsdf, make_neus, oracle_net, orc, linf

CASES = [('surface_volsdf', 'volsdf', dict(near=0.0, far=6.0, N_steps=256, N_secant_steps=8, logit_tau=0.0, fill_inf=True), dict(near=0.0, far=6.0, N_iters=20)),
         ('surface_neus', 'neus', dict(near=0.0, far=2.5, N_steps=128, N_secant_steps=8, logit_tau=0.0, fill_inf=False), dict(near=0.0, far=2.5, N_iters=20)),
         ('surface_volsdf_inside', 'volsdf', dict(near=0.0, far=4.0, N_steps=64, N_secant_steps=4, logit_tau=0.05, fill_inf=True), dict(near=0.2, far=4.0, N_iters=8))]


def _model(G, framework, device='cpu'):
    bump = float(G['meta'][3])
    return make_volsdf(0.1, bump, device=device) if framework == 'volsdf' else make_neus(0.05, bump, device=device)


def _check(G, tag, d, pt, mask, tol_d, min_agree=1.0):
    gm = G[tag + '_mask'].astype(bool)
    agree = (np.asarray(mask, bool) == gm)
    assert agree.mean() >= min_agree, (tag, 'mask agreement', agree.mean())
    both = agree & gm
    gd = G[tag + '_d']
    if tag == 'rf':
        same_class = agree & ~gm
        assert np.array_equal(np.asarray(d)[same_class], gd[same_class])              # inf / far / 0 are exact
        if both.any():
            assert linf(np.asarray(d)[both], gd[both]) < tol_d
            assert linf(np.asarray(pt)[both], G[tag + '_pt'][both]) < tol_d
    else:
        # sphere tracing is an iterated map: rays that converge on the surface (mask true) contract errors, rays that leave the
        # volume (mask false, frozen at an arbitrary step) amplify last-bit sdf differences -- compared loosely
        if both.any():
            assert linf(np.asarray(d)[both], gd[both]) < tol_d * 20
        dead = agree & ~gm
        if dead.any():
            assert linf(np.asarray(d)[dead], gd[dead]) < max(2e-3, tol_d * 100)
    return agree


@pytest.mark.parametrize('name,framework,rcfg,scfg', CASES)
def test_oracle_ray_casting_vs_reference(name, framework, rcfg, scfg):
    G = golden(name)
    net = oracle_net(_model(G, framework), framework)
    dirs = orc._normalize(G['rays_d'])
    q = lambda x: orc.sdf_net(net, x)[0]
    d, pt, mask, msc = orc.root_finding_surface_points(q, G['rays_o'], dirs, **rcfg)
    _check(G, 'rf', d, pt, mask, 1e-5)           # grazing rays: depth error = sdf error / slope
    assert np.array_equal(msc, G['rf_mask_sign_change'].astype(bool))
    d, pt, mask = orc.sphere_tracing_surface_points(q, G['rays_o'], dirs, **scfg)
    _check(G, 'st', d, pt, mask, 1e-5)
    for algo, cfg, tag in (('root_finding', rcfg, 'sr_rf'), ('sphere_tracing', scfg, 'sr_st')):
        out = orc.surface_render(net, G['rays_o'], G['rays_d'], algo, **cfg)
        gm = G[tag + '_mask'].astype(bool)
        assert np.array_equal(out['mask_surface'], gm)
        tol = 3e-5 if algo == 'root_finding' else 1e-3          # sphere tracing: rays still marching after N_iters sit on an amplifying map
        assert linf(out['rgb'], G[tag + '_rgb']) < tol
        assert linf(out['normals_surface'], G[tag + '_normals']) < tol * 3
        sel = slice(None) if algo == 'root_finding' else gm       # sphere tracing: dead rays stop at arbitrary points
        assert linf(out['implicit_nablas'][sel], G[tag + '_nablas'][sel]) < tol * 3


@pytest.mark.gpu
@pytest.mark.parametrize('name,framework,rcfg,scfg', CASES)
def test_cuda_ray_casting_vs_reference(name, framework, rcfg, scfg):
    import nerfart_b200
    from nerfart_b200.models import ray_casting as rc
    dev = 'cuda:0'
    G = golden(name)
    m = _model(G, framework, device=dev)
    tc = nerfart_b200.default_precision() != 'fp32'
    tol = 2e-4 if tc else 2e-5
    ro = torch.tensor(G['rays_o'], device=dev); rd = torch.tensor(G['rays_d'], device=dev)
    rdn = torch.nn.functional.normalize(rd, dim=-1)
    d, pt, mask, msc = rc.root_finding_surface_points(m.implicit_surface, ro[None], rdn[None], batched=True, **rcfg)
    assert d.shape == (1, ro.shape[0]) and pt.shape == (1, ro.shape[0], 3) and mask.dtype == torch.bool
    _check(G, 'rf', d[0].cpu().numpy(), pt[0].cpu().numpy(), mask[0].cpu().numpy(), tol, min_agree=0.99)
    d, pt, mask = rc.sphere_tracing_surface_points(m.implicit_surface, ro, rdn, batched=False, **scfg)
    _check(G, 'st', d.cpu().numpy(), pt.cpu().numpy(), mask.cpu().numpy(), tol, min_agree=0.99)
    for algo, cfg, tag in (('root_finding', rcfg, 'sr_rf'), ('sphere_tracing', scfg, 'sr_st')):
        col, dep, ex = rc.surface_render(ro[None], rd[None], m, calc_normal=True, batched=True, use_view_dirs=True,
                                         ray_casting_algo=algo, ray_casting_cfgs=dict(cfg))
        assert list(ex.keys()) == ['implicit_nablas', 'mask_surface', 'normals_surface']
        gm = G[tag + '_mask'].astype(bool)
        agree = ex['mask_surface'][0].cpu().numpy() == gm
        assert agree.mean() >= 0.99
        frac = agree.mean()
        if algo == 'sphere_tracing':
            agree = agree & gm
        rep = {k: linf(v[0].cpu().numpy()[agree], G[tag + '_' + g][agree]) for k, v, g in
               (('rgb', col, 'rgb'), ('normals', ex['normals_surface'], 'normals'), ('nablas', ex['implicit_nablas'], 'nablas'))}
        print(name, algo, rep, 'mask agreement', frac)
        lim = (3e-3 if tc else 3e-4) * (1 if algo == 'root_finding' else 3)   # hit points move by the depth tolerance; colours / normals follow
        assert rep['rgb'] < lim and rep['normals'] < lim * 3 and rep['nablas'] < lim * 3
        assert (col[0][~ex['mask_surface'][0]] == 0).all()


@pytest.mark.gpu
def test_extract_mesh_sdf_grid_matches_pointwise_evaluation():
    """mesh_util.sdf_grid: the N^3 SDF samples of extract_mesh equal implicit_surface.forward at the reference's
    lattice points, and the oracle's SDF network on the same points."""
    import numpy as np
    from nerfart_b200.utils import mesh_util as mu
    from helpers import make_volsdf
    m = make_volsdf(0.1, 0.0, device='cuda:0')
    N, s = 24, 2.0
    grid = mu.sdf_grid(m.implicit_surface, s, N, chunk=5000)                 # several chunks, the last one short
    pts = mu.grid_points(0, N ** 3, N, s, 'cuda:0')
    with torch.no_grad():
        direct = m.implicit_surface.forward(pts).cpu().numpy().reshape(N, N, N)
    assert np.array_equal(grid, direct)
    ref = np.asarray(orc.sdf_net(oracle_net(m, 'volsdf'), pts.cpu().numpy())[0]).reshape(N, N, N)      # numpy restatement of base.py:217-262
    assert np.abs(grid - ref).max() < 1e-4


@pytest.mark.gpu
def test_frame_sink_writes_the_frames_of_the_synchronous_path(tmp_path):
    """utils/frame_sink.py: frames rendered through render_views + FrameSink equal `(rgb * 255).astype(uint8)` of a
    synchronous render of the same views (render.py:508-509,530-536), in order, as 8-bit RGB PNGs."""
    import cv2
    import numpy as np
    from helpers import make_volsdf, fx
    from nerfart_b200.utils.frame_sink import FrameSink, render_views
    from nerfart_b200.utils import rend_util
    from nerfart_b200.models.frameworks.volsdf import volume_render
    import functools
    dev = 'cuda:0'
    m = make_volsdf(0.1, 0.5, device=dev)
    H, W = 40, 24
    cams = []
    for k in range(4):
        c2w, K = fx.tilted_camera(H, W)
        c2w = c2w.clone(); c2w[0, 3] += 0.1 * k
        cams.append(c2w)
    kw = dict(batched=True, near=0.0, far=6.0, perturb=False, max_upsample_steps=6, N_samples=32, N_importance=16, require_nablas=True,
              calc_normal=True, detailed_output=False)
    render_fn = functools.partial(volume_render, model=m)
    sink = FrameSink(str(tmp_path / 'rgb'), keep_frames=True)
    assert render_views(render_fn, [c.numpy() for c in cams], K.to(dev), H, W, sink, **kw) == 4
    sink.close()
    for k, c2w in enumerate(cams):
        with torch.no_grad():
            ro, rd, _ = rend_util.get_rays(c2w[None].to(dev), K[None].to(dev), H, W)
            rgb, _, _ = volume_render(ro, rd, m, **kw)
        want = (rgb.data.cpu().reshape(H, W, 3).numpy() * 255.).astype(np.uint8)
        got = cv2.imread(str(tmp_path / 'rgb' / ('%05d.png' % (k + 1))))[..., ::-1]
        assert np.array_equal(got, want) and np.array_equal(sink.frames[k + 1], want)
