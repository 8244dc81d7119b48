"""GPU: backward of the render (csrc/train.cu through the C ABI) against
  (1) parameter gradients from the UNMODIFIED reference's autograd (tests/golden/train_*.npz) and
  (2) the float64 backward oracle on larger seeded inputs (several tiles, several sample splits, ragged last tile),
and the fine-tune step `Trainer.forward` end to end with an injected style loss.

Tolerances.  The kernels compute in fp32 (CUDA cores) and sum over samples with atomics.  Two effects set the floor, both
shared with the reference's own fp32 run: (i) softplus'(z) = sigmoid(100 z) amplifies pre-activation rounding 25x, and a 1-ulp
difference in the normalised ray direction moves sin(32 x) by 1e-5, so the forward nabla differs from the reference's by ~4e-5;
(ii) the radiance ReLUs are kinks: among the ~4e5 layer-0 pre-activations of a golden case a handful sit within that noise of
zero and flip, each moving one row of the layer-0 gradient.  Measured on B200 (scripts/train_debug.py): radiance layers 1-4
agree to 2e-6 of the tensor scale, SDF layers to 1e-4..5e-4, radiance layer 0 to 5e-3 (L-inf) / 1e-3 (L2).  Asserted: every
gradient tensor within 1e-2 of its own largest entry and 5e-3 in relative L2; scalars within 2e-3 relative.
"""
import types
from collections import OrderedDict

import numpy as np
import pytest
import torch

from helpers import make_volsdf, make_neus, golden, fx
import nerfart_oracle_train as ot
from test_oracle_train import compare_grads

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
REL_TOL = 1e-2
L2_TOL = 5e-3
# tensor-core modes ("loaded" backward: the activations come from the tcgen05 forward kernel, whose pre-activations carry the
# 1.5e-5 error of split-fp16 operands + truncating fp32 accumulation).  softplus' = sigmoid(100 z) amplifies it 25x and the
# second-order (eikonal) term 100 g-bar g (1 - s) another 100x, so the eikonal-on cases land at 0.5-2.3 % (Linf) / 0.4-0.8 %
# (L2) on the most sensitive tensors (measured on B200: SDF layer 7 bias, radiance layer 0 bias); without the eikonal term and
# for NeuS the tensor-core mode is as close as fp32 (5e-4 / 1e-3).  fp32 mode keeps the all-fp32 recompute and the tight bound.
TOL = {'fp32': (REL_TOL, L2_TOL), 'tc': (3e-2, 1.2e-2), 'tc_mixed': (3e-2, 1.2e-2)}


def worst_errors(grads, g):
    """(worst Linf / max|ref|, worst relative L2, its tensor) over the tensors the golden file holds in full"""
    worst = (0.0, 0.0, '')
    for k, v in grads.items():
        if 'grad.' + k not in g:
            continue
        ref = g['grad.' + k]; a = np.asarray(v).reshape(ref.shape)
        li = float(np.abs(a - ref).max() / (np.abs(ref).max() + 1e-30))
        l2 = float(np.linalg.norm((a - ref).ravel()) / (np.linalg.norm(ref.ravel()) + 1e-30))
        if l2 > worst[1]:
            worst = (max(li, worst[0]), l2, k)
        else:
            worst = (max(li, worst[0]), worst[1], worst[2])
    return worst


def t(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=DEV)


def state(m):
    return {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}


def normalize(d):
    return d / d.norm(dim=-1, keepdim=True).clamp_min(1e-12)


def volsdf_fwd_at(m, ro, rd, d_all):
    """the detailed forward outputs at given depths, from the product's own network kernel (VolSDF.forward, volsdf.py:359-370)"""
    n, P = d_all.shape
    dn = normalize(rd)
    pts = ro[:, None, :] + dn[:, None, :] * d_all[:, :, None]
    with torch.no_grad():
        rad, sdf, nab = m.forward(pts.reshape(-1, 3), dn[:, None, :].expand(n, P, 3).reshape(-1, 3))
    return dict(d_vals=d_all.contiguous(), sdf=sdf.reshape(n, P).contiguous(), nablas=nab.reshape(n, P, 3).contiguous(),
                radiance=rad.reshape(n, P, 3).contiguous())


def neus_fwd_at(m, ro, rd, d_all):
    n, P = d_all.shape
    dn = normalize(rd)
    pts = ro[:, None, :] + dn[:, None, :] * d_all[:, :, None]
    d_mid = 0.5 * (d_all[:, 1:] + d_all[:, :-1])
    pm = ro[:, None, :] + dn[:, None, :] * d_mid[:, :, None]
    with torch.no_grad():
        _, sdf, nab = m.forward(pts.reshape(-1, 3), dn[:, None, :].expand(n, P, 3).reshape(-1, 3))
        rad = m.forward_radiance(pm.reshape(-1, 3), dn[:, None, :].expand(n, P - 1, 3).reshape(-1, 3))
    return dict(d_all=d_all.contiguous(), sdf=sdf.reshape(n, P).contiguous(), nablas=nab.reshape(n, P, 3).contiguous(),
                radiance=rad.reshape(n, P - 1, 3).contiguous())


def product_grads(m, framework, ro, rd, fwd, G, w_eik, white, train_radiance=True, eik_count=None):
    eng = m.engine()
    eng.pack()
    eng.grad_zero()
    if framework == 'volsdf':
        a, b = m.forward_ab()
        scal = torch.cat([a.detach().reshape(1), b.detach().reshape(1)]).float().contiguous()
        P = fwd['d_vals'].shape[1]
    else:
        scal = m.forward_s().detach().reshape(1).float().contiguous()
        P = fwd['d_all'].shape[1]
    eng.render_bwd(ro.contiguous(), rd.contiguous(), scal, fwd, G, w_eikonal=w_eik, eikonal_count=eik_count or ro.shape[0] * P,
                   white_bkgd=white, speed_factor=m.speed_factor, train_surface=True, train_radiance=train_radiance)
    pairs, scal_out = eng.unpack_grads(True, train_radiance)
    torch.cuda.synchronize()
    names = {id(p): k for k, p in m.named_parameters()}
    grads = OrderedDict((names[id(p)], g.cpu().numpy()) for p, g in pairs)
    return grads, scal_out.cpu().numpy()


@pytest.mark.parametrize('precision', ['fp32', 'tc'])
@pytest.mark.parametrize('name', ['train_volsdf_b0.1', 'train_volsdf_b0.01_white', 'train_volsdf_noeik'])
def test_volsdf_backward_matches_reference_autograd(name, precision):
    g = golden(name)
    m = make_volsdf(float(g['beta_init']), float(g['bump']), device=DEV)
    m.engine().precision = precision
    ro, rd = t(g['rays_o']), t(g['rays_d'])
    fwd = volsdf_fwd_at(m, ro, rd, t(g['d_vals']))
    grads, scal = product_grads(m, 'volsdf', ro, rd, fwd, t(g['G']), float(g['w_eikonal']), bool(g['white_bkgd']))
    grads['ln_beta'] = np.array([scal[0]], np.float32)
    print(name, precision, 'worst (Linf, L2, tensor) vs reference autograd:', worst_errors(grads, g))
    assert compare_grads(grads, g, *TOL[precision]) == 9 * 3 + 5 * 3 + 1
    assert abs(scal[1] - float(g['eikonal_loss'])) <= 1e-5 + 1e-4 * float(g['eikonal_loss'])
    print(name, 'ln_beta grad', scal[0], float(g['grad.ln_beta'][0]), 'eik', scal[1], float(g['eikonal_loss']))


@pytest.mark.parametrize('precision', ['fp32', 'tc'])
def test_neus_backward_matches_reference_autograd(precision):
    g = golden('train_neus')
    m = make_neus(float(g['variance_init']), float(g['bump']), device=DEV)
    m.engine().precision = precision
    ro, rd = t(g['rays_o']), t(g['rays_d'])
    fwd = neus_fwd_at(m, ro, rd, t(g['d_all']))
    grads, scal = product_grads(m, 'neus', ro, rd, fwd, t(g['G']), float(g['w_eikonal']), False, train_radiance=False)
    grads['ln_s'] = np.array([scal[0]], np.float32)
    print('neus', precision, 'worst (Linf, L2, tensor) vs reference autograd:', worst_errors(grads, g))
    assert compare_grads(grads, g, *TOL[precision]) == 9 * 3 + 1
    assert abs(scal[1] - float(g['eikonal_loss'])) <= 1e-5 + 1e-4 * float(g['eikonal_loss'])


def rel_err(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / (np.abs(b).max() + 1e-30))


def l2_err(a, b):
    return float(np.linalg.norm((np.asarray(a, np.float64) - b).ravel()) / (np.linalg.norm(np.asarray(b).ravel()) + 1e-30))


@pytest.mark.parametrize('precision', ['fp32', 'tc'])
def test_volsdf_backward_vs_oracle_many_tiles(precision):
    """333 rays x 48 points = 15 984 samples = 124.9 tiles (ragged), sample depths from the product's own sampler."""
    m = make_volsdf(0.05, 0.5, device=DEV)
    m.engine().precision = precision
    from nerfart_b200.models.frameworks.volsdf import render_patch
    from nerfart_b200.utils import rend_util
    c2w, K = fx.tilted_camera(37, 9)
    with torch.no_grad():
        ro, rd, _ = rend_util.get_rays(c2w[None].to(DEV), K[None].to(DEV), 37, 9)
    ro, rd = ro[0].contiguous(), rd[0].contiguous()
    fwd, _ = render_patch(m, ro, rd, N_samples=32, N_importance=16, max_upsample_steps=6)
    gen = torch.Generator(device='cpu'); gen.manual_seed(3)
    G = (0.05 * torch.randn(ro.shape[0], 3, generator=gen)).to(DEV)
    grads, scal = product_grads(m, 'volsdf', ro, rd, fwd, G, 0.1, False)
    net = ot.TrainNet(state(m), 'volsdf')
    og, oeik, orgb = ot.volsdf_backward(net, ro.cpu().numpy(), rd.cpu().numpy(), fwd['d_vals'].cpu().numpy(), G.cpu().numpy(), 0.1, False)
    worst = max(rel_err(grads[k], np.asarray(og[k]).reshape(grads[k].shape)) for k in grads)
    worst2 = max(l2_err(grads[k], np.asarray(og[k]).reshape(grads[k].shape)) for k in grads)
    print(precision, 'worst relative gradient error vs oracle (Linf, L2)', worst, worst2, 'ln_beta', scal[0], float(og['ln_beta'][0]), 'eik', scal[1], oeik)
    assert worst2 < TOL[precision][1], worst2
    assert worst < TOL[precision][0]
    assert abs(scal[0] - float(og['ln_beta'][0])) <= 2e-3 * abs(float(og['ln_beta'][0])) + 1e-6
    assert abs(scal[1] - oeik) <= 1e-4 * oeik + 1e-6
    assert np.abs(fwd['rgb'].cpu().numpy() - orgb).max() < (1e-4 if precision == 'fp32' else 1e-3)


@pytest.mark.parametrize('precision', ['fp32', 'tc'])
def test_backward_reads_no_uninitialised_workspace(precision):
    """Ragged patch (333 rays x 48 points = 124.9 tiles) through a workspace poisoned with NaN: the stash rows of the last tile's
    padding samples are read by the weight-gradient kernels, so every one of them must have been written (with zero gradient)."""
    m = make_volsdf(0.05, 0.5, device=DEV)
    m.engine().precision = precision
    from nerfart_b200.models.frameworks.volsdf import render_patch
    from nerfart_b200.utils import rend_util
    c2w, K = fx.tilted_camera(37, 9)
    with torch.no_grad():
        ro, rd, _ = rend_util.get_rays(c2w[None].to(DEV), K[None].to(DEV), 37, 9)
    ro, rd = ro[0].contiguous(), rd[0].contiguous()
    fwd, _ = render_patch(m, ro, rd, N_samples=32, N_importance=16, max_upsample_steps=6)
    G = torch.full((ro.shape[0], 3), 0.02, device=DEV)
    clean, sc = product_grads(m, 'volsdf', ro, rd, fwd, G, 0.1, False)
    eng = m.engine()
    eng._tws.view(torch.float32).fill_(float('nan'))
    dirty, sd = product_grads(m, 'volsdf', ro, rd, fwd, G, 0.1, False)
    for k in clean:
        assert np.isfinite(dirty[k]).all(), k
        assert np.abs(dirty[k] - clean[k]).max() <= 2e-5 * np.abs(clean[k]).max() + 1e-9, k      # atomics: summation order only
    assert np.isfinite(sd).all()


def test_backward_accumulates_over_patches_and_is_deterministic_in_structure():
    """two half patches accumulate to the gradient of the whole patch (same eikonal normaliser): the GradPack is additive"""
    g = golden('train_volsdf_b0.1')
    m = make_volsdf(float(g['beta_init']), float(g['bump']), device=DEV)
    m.engine().precision = 'fp32'
    ro, rd, G = t(g['rays_o']), t(g['rays_d']), t(g['G'])
    fwd = volsdf_fwd_at(m, ro, rd, t(g['d_vals']))
    n, P = fwd['d_vals'].shape
    whole, sw = product_grads(m, 'volsdf', ro, rd, fwd, G, 0.1, False)
    eng = m.engine(); eng.grad_zero()
    a, b = m.forward_ab()
    scal = torch.cat([a.detach().reshape(1), b.detach().reshape(1)]).float().contiguous()
    for sl in (slice(0, 20), slice(20, n)):
        part = {k: v[sl].contiguous() for k, v in fwd.items()}
        eng.render_bwd(ro[sl].contiguous(), rd[sl].contiguous(), scal, part, G[sl], w_eikonal=0.1, eikonal_count=n * P,
                       white_bkgd=False, speed_factor=m.speed_factor)
    pairs, sp = eng.unpack_grads()
    names = {id(p): k for k, p in m.named_parameters()}
    for p, gr in pairs:
        ref = whole[names[id(p)]]
        assert np.abs(gr.cpu().numpy() - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-7, names[id(p)]
    np.testing.assert_allclose(sp.cpu().numpy(), sw, rtol=1e-5, atol=1e-8)


class _Args(dict):
    __getattr__ = dict.__getitem__


@pytest.mark.parametrize('precision', ['fp32', 'tc', 'tc_mixed'])
def test_trainer_forward_finetune_step_volsdf(monkeypatch, precision):
    """Trainer.forward (fine-tune branch) with an injected differentiable style loss: same protocol as the reference
    (loss already back-propagated, .grad populated, optimizer.zero_grad() called inside), gradients equal to the oracle's.
    fp32 mode, 'tc' and the DEFAULT mode 'tc_mixed' (renders in tc_mixed; the backward program always re-evaluates the forward with tc operands)."""
    from nerfart_b200.models.frameworks import volsdf as pv, _finetune
    PB = 200                                                         # 576 rays: two full patches (ONE launch group) + a short one
    monkeypatch.setattr(_finetune, 'BATCH_SIZE', PB)
    m = make_volsdf(0.1, 0.5, device=DEV).train()
    m.engine().precision = precision
    H = W = 24
    target = torch.full((1, H * W, 3), 0.25, device=DEV)
    wts = torch.linspace(0.5, 1.5, H * W * 3, device=DEV).reshape(1, 3, H, W)

    def clip_like(gt, s_text, pred, t_text):
        return ((pred - gt) ** 2 * wts).mean()
    zero = lambda *a, **k: torch.zeros((), device=DEV)
    trainer = pv.Trainer(m, is_finetune=True, target_hw=[H, W],
                         loss_dict={'clip': clip_like, 'perceptual': None, 'contrastive': zero, 'patchnce': zero})
    trainer.neg_texts = ['a'] * 10
    args = _Args(training=_Args(is_finetune=True), data=_Args(downscale=2),
                 model=_Args(radiance=_Args(use_view_dirs=True)),
                 finetune=_Args(use_eikonal=True, w_eikonal=0.1, w_clip=1.0, w_perceptual=2.0, w_contrastive=0.2, w_patchnce=0.1,
                                src_text='photo', target_text='painting'))
    c2w, K = fx.tilted_camera(H, W)
    kw = dict(near=0.0, far=6.0, batched=True, perturb=False, white_bkgd=False, max_upsample_steps=6, use_nerfplusplus=False,
              obj_bounding_radius=3.0, H=H, W=W, N_samples=32, N_importance=16)
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    for p in m.parameters():
        p.grad = torch.ones_like(p)                                  # must be cleared by the zero_grad() inside
    ret = trainer(args, None, {'intrinsics': K[None], 'c2w': c2w[None]}, {'rgb': target.cpu()}, kw, 0, optimizer=opt)
    assert ret['losses'].dim() == 0 and set(ret['extras']) == {'scalars', 'select_inds'}
    # oracle: same image gradient, same patches, depths from the product's deterministic sampler
    from nerfart_b200.utils import rend_util
    with torch.no_grad():
        ro, rd, _ = rend_util.get_rays(c2w[None].to(DEV), K[None].to(DEV), H, W)
        rgb, _, _ = trainer.renderer(ro, rd, detailed_output=False, require_nablas=True, **kw)
    rgb = rgb.detach().requires_grad_(True)
    loss = clip_like(target.reshape(1, H, W, 3).permute(0, 3, 1, 2), '', rgb.reshape(1, H, W, 3).permute(0, 3, 1, 2), '')
    loss.backward()
    assert abs(float(loss) - float(ret['losses'])) < 1e-6
    G = rgb.grad[0]
    net = ot.TrainNet(state(m), 'volsdf')
    total = None
    for i in range(0, H * W, PB):                                    # the oracle goes patch by patch, like the reference
        fwd, _ = pv.render_patch(m, ro[0, i:i + PB].contiguous(), rd[0, i:i + PB].contiguous(), **kw)
        og, _, _ = ot.volsdf_backward(net, ro[0, i:i + PB].cpu().numpy(), rd[0, i:i + PB].cpu().numpy(), fwd['d_vals'].cpu().numpy(),
                                      G[i:i + PB].cpu().numpy(), 0.1, False)
        total = og if total is None else OrderedDict((k, total[k] + og[k]) for k in og)
    worst = 0.0
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        worst = max(worst, rel_err(p.grad.cpu().numpy(), np.asarray(total[k]).reshape(tuple(p.shape))))
    print(f'Trainer.forward [{precision}]: worst relative parameter-gradient error vs oracle', worst)
    assert worst < TOL[precision][0]
    opt.step()


@pytest.mark.parametrize('precision', ['fp32', 'tc', 'tc_mixed'])
def test_trainer_forward_finetune_step_neus(monkeypatch, precision):
    from nerfart_b200.models.frameworks import neus as pn, _finetune
    monkeypatch.setattr(_finetune, 'BATCH_SIZE', 300)
    m = make_neus(0.05, 0.5, device=DEV).train()
    m.engine().precision = precision
    H = W = 20
    target = torch.full((1, H * W, 3), 0.25, device=DEV)

    def clip_like(gt, s_text, pred, t_text):
        return ((pred - gt) ** 2).mean()
    zero = lambda *a, **k: torch.zeros((), device=DEV)
    trainer = pn.Trainer(m, is_finetune=True, target_hw=[H, W],
                         loss_dict={'clip': clip_like, 'perceptual': None, 'contrastive': zero, 'patchnce': zero})
    trainer.neg_texts = ['a'] * 10
    assert not any(p.requires_grad for p in m.radiance_net.parameters())            # neus.py:28
    args = _Args(training=_Args(is_finetune=True), data=_Args(downscale=2),
                 model=_Args(radiance=_Args(use_view_dirs=True)),
                 finetune=_Args(use_eikonal=True, w_eikonal=0.1, w_clip=1.0, w_perceptual=2.0, w_contrastive=0.2, w_patchnce=0.1,
                                src_text='photo', target_text='painting'))
    c2w, K = fx.tilted_camera(H, W)
    c2w = c2w.clone(); c2w[:3, 3] *= 0.6
    kw = dict(upsample_algo='official_solution', N_nograd_samples=2048, N_upsample_iters=4, N_outside=0, obj_bounding_radius=1.0,
              batched=True, perturb=False, white_bkgd=False, H=H, W=W, N_samples=32, N_importance=16)
    opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-4)
    ret = trainer(args, None, {'intrinsics': K[None], 'c2w': c2w[None]}, {'rgb': target.cpu()}, kw, 0, optimizer=opt)
    from nerfart_b200.utils import rend_util
    with torch.no_grad():
        ro, rd, _ = rend_util.get_rays(c2w[None].to(DEV), K[None].to(DEV), H, W)
        rgb, _, _ = trainer.renderer(ro, rd, detailed_output=False, **kw)
    rgb = rgb.detach().requires_grad_(True)
    ((rgb - target) ** 2).mean().backward()
    G = rgb.grad[0]
    net = ot.TrainNet(state(m), 'neus')
    total = None
    for i in range(0, H * W, 300):
        fwd, _ = pn.render_patch(m, ro[0, i:i + 300].contiguous(), rd[0, i:i + 300].contiguous(), **kw)
        og, _, _ = ot.neus_backward(net, ro[0, i:i + 300].cpu().numpy(), rd[0, i:i + 300].cpu().numpy(), fwd['d_all'].cpu().numpy(),
                                    G[i:i + 300].cpu().numpy(), 0.1, False)
        total = og if total is None else OrderedDict((k, total[k] + og[k]) for k in og)
    worst = 0.0
    for k, p in m.named_parameters():
        if not p.requires_grad:
            assert p.grad is None, k
            continue
        worst = max(worst, rel_err(p.grad.cpu().numpy(), np.asarray(total[k]).reshape(tuple(p.shape))))
    print(f'NeuS Trainer.forward [{precision}]: worst relative parameter-gradient error vs oracle', worst)
    assert worst < TOL[precision][0]
    opt.step()


@pytest.mark.parametrize('l_bf16,r_bf16', [(0, 0), (1, 1)])          # (tcgen05 kind::f16 rejects an fp16 operand against a bf16 one)
@pytest.mark.parametrize('m_rows', [64, 128 * 37 + 50, 230400])
def test_wgrad_f16_kernel_vs_torch(l_bf16, r_bf16, m_rows):
    """csrc/wgrad_f16.cu alone (TMA-fed MN-major tcgen05 GEMM over sample-major 16-bit planes): out = L^T R and the column sums of L
    against torch in float64 on the same 16-bit values; fp16 and bf16 operands; ragged and full-patch sample counts."""
    import ctypes as C
    import nerfart_b200
    from nerfart_b200._lib import check
    lib = nerfart_b200.lib()
    m_pad = (m_rows + 127) // 128 * 128
    g = torch.Generator(device=DEV); g.manual_seed(7)
    dt = [torch.bfloat16 if l_bf16 else torch.float16, torch.bfloat16 if r_bf16 else torch.float16]
    Lp = (torch.randn(m_pad, 256, device=DEV, generator=g) * torch.rand(m_pad, 1, device=DEV, generator=g)).to(dt[0])
    Rp = torch.randn(m_pad, 256, device=DEV, generator=g).to(dt[1])
    Lp[m_rows:] = 0                                                    # padding rows carry zero gradient planes
    planes = torch.empty(2, m_pad, 256, dtype=torch.int16, device=DEV)
    planes[0] = Lp.view(torch.int16); planes[1] = Rp.view(torch.int16)
    out = torch.zeros(256, 256, device=DEV); bias = torch.zeros(256, device=DEV)
    check(lib.na_debug_wgrad_f16(C.c_void_p(planes.data_ptr()), m_pad, m_rows, l_bf16, r_bf16, C.c_void_p(out.data_ptr()),
                                 C.c_void_p(bias.data_ptr()), None), 'na_debug_wgrad_f16')
    torch.cuda.synchronize()
    ref = Lp[:m_rows].double().t() @ Rp[:m_rows].double()
    bref = Lp[:m_rows].double().sum(0)
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    berr = float((bias.double() - bref).abs().max() / (bref.abs().max() + 1e-30))
    print(f'wgrad_f16 m={m_rows} fmt=({l_bf16},{r_bf16}): rel err {err:.2e}, bias {berr:.2e}')
    assert err < 2e-5 and berr < 2e-5                                   # products are exact in fp32; only the summation order differs


@pytest.mark.parametrize('precision,white,w_eik', [('tc_mixed', False, 0.1), ('tc', True, 0.0), ('tc_mixed', True, 0.1)])
def test_split_training_program_equals_the_one_launch_program(precision, white, w_eik):
    """na_volsdf_render_fwd_train + na_volsdf_render_bwd_stashed (the patch's forward render IS the forward half of the training
    program, the backward launch runs GEMMs 21..40 only) against na_volsdf_render_fwd + na_volsdf_render_bwd (the backward launch
    re-evaluates the forward pass): same forward outputs bit for bit, same gradients up to the order of the atomic sums.
    Several tiles, a ragged last tile (523 rays x 48 samples = 196.1 tiles) and the sphere-background override on part of the rays."""
    from nerfart_b200.models.frameworks.volsdf import render_patch
    from nerfart_b200.utils import rend_util
    m = make_volsdf(0.1, 0.5, device=DEV).train()
    eng = m.engine(); eng.precision = precision
    H, W = 24, 24
    c2w, K = fx.tilted_camera(H, W)
    with torch.no_grad():
        ro, rd, _ = rend_util.get_rays(c2w[None].to(DEV), K[None].to(DEV), H, W)
    n = 523
    ro, rd = ro[0, :n].contiguous(), rd[0, :n].contiguous()
    g = torch.Generator(device=DEV); g.manual_seed(11)
    G = 1e-2 * torch.randn(n, 3, device=DEV, generator=g)
    kw = dict(near=0.0, far=6.0, N_samples=32, N_importance=16, max_upsample_steps=6, perturb=False, white_bkgd=white)
    res = []
    for stash in (False, True):
        fwd, ab = render_patch(m, ro, rd, train_stash=stash, **kw)
        assert (eng._stash_key is not None) == stash
        eng.grad_zero()
        eng.render_bwd(ro, rd, ab, fwd, G, w_eikonal=w_eik, eikonal_count=n * 48, white_bkgd=white, speed_factor=m.speed_factor)
        assert eng._stash_key is None                                  # a stash is consumed by the backward of ITS render only
        pairs, scal = eng.unpack_grads()
        torch.cuda.synchronize()
        res.append((fwd, [(p, gr.clone()) for p, gr in pairs], scal.clone()))
    (f0, p0, s0), (f1, p1, s1) = res
    for k in ('rgb', 'd_vals', 'sdf', 'radiance', 'nablas'):
        assert torch.equal(f0[k], f1[k]), k
    names = {id(p): k for k, p in m.named_parameters()}
    worst = 0.0
    for (pa, ga), (pb, gb) in zip(p0, p1):
        assert pa is pb
        e = float((ga - gb).abs().max() / (ga.abs().max() + 1e-30))
        worst = max(worst, e)
        assert e < 2e-3, (names[id(pa)], e)
    print(f'split vs one-launch training program [{precision}, white={white}, eik={w_eik}]: worst gradient difference {worst:.2e} of the tensor scale')
    assert torch.allclose(s0, s1, rtol=1e-5, atol=1e-12)
    # a render without the stash in between invalidates it: the backward falls back to the one-launch program
    fwd, ab = render_patch(m, ro, rd, train_stash=True, **kw)
    render_patch(m, ro[:50].contiguous(), rd[:50].contiguous(), **kw)
    assert eng._stash_key is None


@pytest.mark.parametrize('precision,w_eik', [('tc_mixed', 0.1), ('tc', 0.0)])
def test_split_training_program_equals_the_one_launch_program_neus(precision, w_eik):
    """NeuS (neus.py:520-576; radiance net frozen): na_neus_render_fwd_train + na_neus_render_bwd_stashed -- the two final evaluations of
    the forward render (P points: sdf + nabla, P - 1 midpoints: radiance) are the forward halves, their stashes one behind the other in
    the workspace; the backward launches run the second-order sweep + trunk (points) and GEMMs 21..40 (midpoints) only -- against the
    one-launch programs: same forward outputs bit for bit, same gradients up to the order of the atomic sums."""
    from nerfart_b200.models.frameworks.neus import render_patch
    from nerfart_b200.utils import rend_util
    m = make_neus(0.05, 0.5, device=DEV).train()
    eng = m.engine(); eng.precision = precision
    H, W = 24, 24
    c2w, K = fx.tilted_camera(H, W)
    c2w = c2w.clone(); c2w[:3, 3] = c2w[:3, 3] * 0.3                               # inside NeuS' unit bounding sphere
    with torch.no_grad():
        ro, rd, _ = rend_util.get_rays(c2w[None].to(DEV), K[None].to(DEV), H, W)
    n = 401
    ro, rd = ro[0, :n].contiguous(), rd[0, :n].contiguous()
    g = torch.Generator(device=DEV); g.manual_seed(12)
    G = 1e-2 * torch.randn(n, 3, device=DEV, generator=g)
    kw = dict(obj_bounding_radius=1.0, N_samples=16, N_importance=16, N_upsample_iters=4, perturb=False, white_bkgd=False)
    res = []
    for stash in (False, True):
        fwd, s = render_patch(m, ro, rd, train_stash=stash, **kw)
        assert (eng._stash_key is not None) == stash
        eng.grad_zero()
        eng.render_bwd(ro, rd, s, fwd, G, w_eikonal=w_eik, eikonal_count=n * 32, white_bkgd=False, speed_factor=m.speed_factor,
                       train_radiance=False)
        assert eng._stash_key is None
        pairs, scal = eng.unpack_grads(True, False)
        torch.cuda.synchronize()
        res.append((fwd, [(p, gr.clone()) for p, gr in pairs], scal.clone()))
    (f0, p0, s0), (f1, p1, s1) = res
    for k in ('rgb', 'd_all', 'sdf', 'radiance', 'nablas'):
        assert torch.equal(f0[k], f1[k]), k
    names = {id(p): k for k, p in m.named_parameters()}
    worst = 0.0
    for (pa, ga), (pb, gb) in zip(p0, p1):
        assert pa is pb
        e = float((ga - gb).abs().max() / (ga.abs().max() + 1e-30))
        worst = max(worst, e)
        assert e < 2e-3, (names[id(pa)], e)
    print(f'NeuS split vs one-launch training program [{precision}, eik={w_eik}]: worst gradient difference {worst:.2e} of the tensor scale')
    assert torch.allclose(s0, s1, rtol=1e-5, atol=1e-12)
