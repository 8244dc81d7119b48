"""GPU debug aid: per-tensor gradient errors of the CUDA backward vs the float64 oracle on a train golden case."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import numpy as np, torch
from helpers import make_volsdf, golden
import nerfart_oracle_train as ot
import test_gpu_train as T
name = sys.argv[1] if len(sys.argv) > 1 else 'train_volsdf_b0.1'
g = golden(name)
m = make_volsdf(float(g['beta_init']), float(g['bump']), device='cuda:0'); m.engine().precision = 'fp32'
ro, rd = T.t(g['rays_o']), T.t(g['rays_d'])
fwd = T.volsdf_fwd_at(m, ro, rd, T.t(g['d_vals']))
grads, scal = T.product_grads(m, 'volsdf', ro, rd, fwd, T.t(g['G']), float(g['w_eikonal']), bool(g['white_bkgd']))
net = ot.TrainNet(T.state(m), 'volsdf')
og, oeik, orgb = ot.volsdf_backward(net, g['rays_o'], g['rays_d'], g['d_vals'], g['G'], float(g['w_eikonal']), bool(g['white_bkgd']))
print('scal', scal, float(og['ln_beta'][0]), oeik)
for k in grads:
    a = grads[k]; b = np.asarray(og[k]).reshape(a.shape)
    e = np.abs(a - b)
    line = f'{k:50s} max|ref| {np.abs(b).max():.3e} err {e.max():.3e} rel {e.max()/(np.abs(b).max()+1e-30):.2e}'
    if a.ndim == 2 and a.shape[1] > 1:
        col = e.max(0); row = e.max(1)
        line += f'  worst cols {np.argsort(-col)[:6].tolist()} rows {np.argsort(-row)[:4].tolist()}'
    print(line)
for k in ('sdf', 'nablas', 'radiance'):
    ref = {'sdf': g['sdf'], 'nablas': g['nablas'], 'radiance': g['radiance']}[k]
    print('fwd', k, np.abs(fwd[k].cpu().numpy() - ref).max())
if len(sys.argv) > 2:
    # bigger seeded case: H x W rays from the tilted camera, depths from the product's sampler
    from helpers import fx
    from nerfart_b200.models.frameworks.volsdf import render_patch
    from nerfart_b200.utils import rend_util
    H, W = int(sys.argv[2]), int(sys.argv[3])
    c2w, K = fx.tilted_camera(H, W)
    with torch.no_grad():
        ro, rd, _ = rend_util.get_rays(c2w[None].cuda(), K[None].cuda(), H, W)
    ro, rd = ro[0].contiguous(), rd[0].contiguous()
    fwd, _ = render_patch(m, ro, rd, N_samples=32, N_importance=16, max_upsample_steps=6)
    gen = torch.Generator(device='cpu'); gen.manual_seed(3)
    G = (0.05 * torch.randn(ro.shape[0], 3, generator=gen)).cuda()
    grads, scal = T.product_grads(m, 'volsdf', ro, rd, fwd, G, 0.1, False)
    og, oeik, orgb = ot.volsdf_backward(net, ro.cpu().numpy(), rd.cpu().numpy(), fwd['d_vals'].cpu().numpy(), G.cpu().numpy(), 0.1, False)
    print('BIG', H * W, 'rays; scal', scal, float(og['ln_beta'][0]), oeik)
    for k in grads:
        a = grads[k]; b = np.asarray(og[k]).reshape(a.shape)
        e = np.abs(a - b)
        print(f'{k:50s} max|ref| {np.abs(b).max():.3e} err {e.max():.3e} rel {e.max()/(np.abs(b).max()+1e-30):.2e}')
