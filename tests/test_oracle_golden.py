"""Pin the numpy oracle (oracle/nerfart_oracle.py) against outputs of the UNMODIFIED reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py in the build container).  CPU only."""
import json
import hashlib
import os
import numpy as np
import pytest
import torch

from helpers import golden, make_volsdf, make_neus, oracle_net, linf, orc, GOLDEN, fx

S = golden('stages')


def test_linspace_matches_torch():
    for n in (2, 16, 17, 32, 64, 128, 130, 512, 514, 1024):
        assert np.array_equal(orc._linspace(0, 1, n), torch.linspace(0, 1, n).numpy()), n


def test_embedder():
    assert linf(orc.embed(S['emb_x'], 6), S['emb6']) < 5e-6
    assert linf(orc.embed(S['emb_x'], 4), S['emb4']) < 5e-6


def test_state_digest_product_init_equals_reference_init():
    d = json.load(open(os.path.join(GOLDEN, 'state_digest.json')))
    def dig(m):
        return {k: [list(v.shape), hashlib.sha256(np.ascontiguousarray(v.numpy()).tobytes()).hexdigest()]
                for k, v in m.state_dict().items()}
    from helpers import VolSDF
    torch.manual_seed(0)
    assert dig(VolSDF(**fx.volsdf_kwargs(0.1))) == d['volsdf_seed0_raw']
    assert dig(make_volsdf(0.01, 0.5)) == d['volsdf_seed0_beta0.01_bump0.5']
    from nerfart_b200.models.frameworks.neus import NeuS
    torch.manual_seed(0)
    assert dig(NeuS(**fx.neus_kwargs(0.05))) == d['neus_seed0_raw']


@pytest.mark.parametrize('tag', ['v', 'n'])
def test_networks(tag):
    m = make_volsdf(0.01, 0.5) if tag == 'v' else make_neus(0.05, 0.5)
    net = oracle_net(m, 'volsdf' if tag == 'v' else 'neus')
    x, v = S[f'net_{tag}_x'], S[f'net_{tag}_v']
    sdf, feat, nab = orc.sdf_net(net, x, with_nablas=True)
    assert linf(sdf, S[f'net_{tag}_sdf']) < 2e-5
    assert linf(feat, S[f'net_{tag}_feat']) < 5e-5
    assert linf(nab, S[f'net_{tag}_nabla']) < 2e-4
    rad = orc.radiance_net(net, x, v, S[f'net_{tag}_nabla'], S[f'net_{tag}_feat'])
    assert linf(rad, S[f'net_{tag}_rad']) < 2e-5
    if tag == 'v':
        assert linf(orc.volsdf_forward_surface(net, x), S['net_v_surface']) < 2e-5
        r, s, _ = orc.volsdf_forward(net, x, v)
        assert linf(r, S['net_v_fwd_rad']) < 5e-5 and linf(s, S['net_v_fwd_sdf']) < 2e-5
        a, b = net.alpha_beta()
        assert abs(a - S['net_v_ab'][0]) / a < 1e-6 and abs(b - S['net_v_ab'][1]) / b < 1e-6
    else:
        assert abs(net.s() - S['net_n_s'][0]) / net.s() < 1e-6


def test_sigma_and_error_bound():
    d, sdf = S['eb_d'], S['eb_sdf']
    for i, (a, b) in enumerate([(10.0, 0.1), (100.0, 0.01), (500.0, 0.002)]):
        np.testing.assert_allclose(orc.sdf_to_sigma(sdf, np.float32(a), np.float32(b)), S[f'sigma_{i}'], rtol=2e-6, atol=1e-30)
        ref = S[f'ebound_{i}']; got = orc.error_bound(d, sdf, a, b)
        assert np.array_equal(np.isinf(ref), np.isinf(got))
        fin = np.isfinite(ref)
        np.testing.assert_allclose(got[fin], ref[fin], rtol=3e-4, atol=5e-7)
    br = S['eb_beta_row']
    ref = S['ebound_row']; got = orc.error_bound(d, sdf, (1.0 / br).astype(np.float32), br)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isinf(ref), np.isinf(got))
    np.testing.assert_allclose(got[fin], ref[fin], rtol=3e-4, atol=5e-7)


def test_samplers():
    d = S['eb_d']
    # the u == 1.0 sample (last column, det=True) depends on whether fp32 sum(weights) rounds cdf[-1] to >= 1: excluded
    assert linf(orc.sample_pdf(d, S['spdf_w'], 50, det=True)[:, :-1], S['spdf_det'][:, :-1]) < 2e-5
    assert linf(orc.sample_cdf(d, S['scdf_cdf'], 16, det=True), S['scdf_det']) < 2e-6
    assert linf(orc.sample_cdf(d, S['scdf_cdf'], 16, det=False, u=S['samp_u']), S['scdf_rand']) < 2e-6
    assert linf(orc.sample_pdf(d, S['spdf_w'], 16, det=False, u=S['samp_u']), S['spdf_rand']) < 2e-5


def test_rays_and_sphere():
    for cam in ('closed', 'tilted'):
        ro, rd = orc.get_rays(S[f'rays_{cam}_c2w'], S[f'rays_{cam}_K'], 6, 5)
        assert linf(ro, S[f'rays_{cam}_o']) == 0
        assert linf(rd, S[f'rays_{cam}_d']) < 1e-6
    n, f = orc.near_far_from_sphere(S['nf_o'], S['nf_d'], 1.0)
    assert linf(n, S['nf_near']) < 1e-6 and linf(f, S['nf_far']) < 1e-6


def test_neus_alpha_weights():
    cdf, al = orc.sdf_to_alpha(S['eb_sdf'], 20.0)
    assert linf(cdf, S['neus_alpha_cdf']) < 1e-6 and linf(al, S['neus_alpha']) < 1e-5
    assert linf(orc.alpha_to_w(S['neus_alpha']), S['neus_w']) < 1e-6


VOLSDF_CASES = [('volsdf_cfg1_b0.1', 0.0), ('volsdf_det_b0.1', 0.5), ('volsdf_det_b0.01', 0.5), ('volsdf_det_b0.002', 0.5),
                ('volsdf_n128_b0.01', 0.5), ('volsdf_n128_b0.1', 0.0), ('volsdf_perturb_b0.01', 0.5)]


def compare_volsdf(out, G, name, scale=1.0, min_same=0.985):
    """Shared by the oracle-vs-reference and the CUDA-vs-oracle tests.

    The error-bound sampler is discontinuous in its inputs: the 10-step bisection on `max bound <= eps`
    (volsdf.py:266-273) and the converged test (162/240) are threshold decisions, so two fp32-correct evaluations whose
    sdf values differ in the last bit can settle on different beta+ (different roots of a non-monotone function) for a
    minority of *non-converged* rays, which moves their fine samples by O(1) (measured between the oracle and the
    reference themselves: ~10 % of the rays of the beta=0.01 fixture, none at the BASELINE beta=0.1).  Rays are therefore
    split into path-consistent rays (same iter_usage, same beta_map), which must agree tightly, and path-divergent rays,
    whose share is bounded and whose images must still be close.  `scale` widens the value tolerances for the reduced-precision
    arithmetic modes of the CUDA path (the sampler-path requirements stay as they are)."""
    rep = {k: linf(out[k], G[k]) for k in ('rgb', 'depth_volume', 'mask_volume', 'normals_volume') if k in out and k in G}
    n = G['rgb'].shape[0]
    if 'iter_usage' in G and 'iter_usage' in out:
        bm_o = np.asarray(out['beta_map']).reshape(n); bm_g = G['beta_map'].reshape(n)
        same = (np.asarray(out['iter_usage']).reshape(n) == G['iter_usage'].reshape(n)) & (np.abs(bm_o - bm_g) <= 2e-6 * np.abs(bm_g))
    else:
        same = np.ones(n, dtype=bool)
    frac_div = 1.0 - same.mean()
    print(name, rep, 'path-divergent rays: %.3f' % frac_div)
    conv_g = (G['iter_usage'].reshape(n) >= 0) if 'iter_usage' in G else np.ones(n, dtype=bool)
    # rays the reference itself converged on must follow the same path almost always
    assert (same | ~conv_g).mean() > min_same, 'converged rays took a different sampler path'
    assert frac_div < 0.35, frac_div          # reference vs oracle: 25 %; fp32 kernels 28 %; default tc mode 29.5 % (profiles/r3b_tc_accumulation.md)
    for k, tol_med, tol_max in (('rgb', 3e-6, 3e-3), ('depth_volume', 1e-5, 3e-2), ('mask_volume', 2e-6, 2e-4), ('normals_volume', 3e-5, 3e-2)):
        if k not in out or k not in G:
            continue
        err = np.abs(np.asarray(out[k]) - G[k]).reshape(n, -1).max(axis=1)
        assert np.median(err[same]) < tol_med * scale, (k, 'median', np.median(err[same]))
        assert np.quantile(err[same], 0.98) < min(tol_max * scale, 0.05), (k, 'q98', np.quantile(err[same], 0.98))
    # path-divergent rays still render nearly the same image: the integral is insensitive to where the samples sit
    assert np.abs(np.asarray(out['rgb']) - G['rgb']).max() < 0.15
    if 'd_vals' in G and 'd_vals' in out:
        assert np.median(np.abs(np.asarray(out['d_vals'])[same] - G['d_vals'][same])) < 2e-6
    return same


@pytest.mark.parametrize('name,bump', VOLSDF_CASES)
def test_volsdf_render_end_to_end(name, bump):
    G = golden(name)
    beta_init, _, H, W, Ns, Ni = G['meta']
    net = oracle_net(make_volsdf(float(beta_init), bump), 'volsdf')
    M = G['rays_o'].shape[0]
    uf = np.broadcast_to(G['u0'], (M, int(Ni))).copy() if 'u0' in G else None
    out = orc.volsdf_render(net, G['rays_o'], G['rays_d'], N_samples=int(Ns), N_importance=int(Ni), perturb=uf is not None,
                            u_final=uf, detailed_output='d_vals' in G)
    same = compare_volsdf(out, G, name)
    if float(beta_init) >= 0.1:
        assert same.all()            # BASELINE configs 1/2: every ray converges at once; no divergence allowed


@pytest.mark.parametrize('name,bump', [('neus_det', 0.5), ('neus_basic', 0.0)])
def test_neus_render_end_to_end(name, bump):
    G = golden(name)
    net = oracle_net(make_neus(float(G['meta'][0]), bump), 'neus')
    out = orc.neus_render(net, G['rays_o'], G['rays_d'], detailed_output='d_final' in G)
    report = {k: linf(out[k], G[k]) for k in ('rgb', 'depth_volume', 'mask_volume', 'normals_volume')}
    print(name, report)
    if 'd_final' in G:
        assert np.median(np.abs(out['d_final'] - G['d_final'])) < 1e-6
    for k, tol_med, tol_max in (('rgb', 2e-6, 2e-3), ('depth_volume', 5e-6, 2e-2), ('mask_volume', 2e-6, 2e-3), ('normals_volume', 2e-5, 2e-2)):
        err = np.abs(out[k] - G[k])
        assert np.median(err) < tol_med, (k, np.median(err))
        assert err.max() < tol_max, (k, err.max())
