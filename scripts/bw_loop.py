"""Stress loop: N training patches back to back (forward render + tcgen05 backward + weight gradients), progress printed per patch."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import torch
from helpers import make_volsdf, fx
import nerfart_b200
from nerfart_b200.models.frameworks.volsdf import render_patch
from nerfart_b200.utils import rend_util
dev = 'cuda:0'
n_patch = int(sys.argv[1]) if len(sys.argv) > 1 else 30
m = make_volsdf(0.1, 0.0, device=dev)
H, W = 480, 270
c2w, K = fx.closed_form_camera(H, W)
with torch.no_grad():
    ro, rd, _ = rend_util.get_rays(c2w[None].to(dev), K[None].to(dev), H, W)
eng = m.engine(); eng.grad_zero()
gen = torch.Generator(device=dev); gen.manual_seed(5)
import contextlib
hold = contextlib.nullcontext()       # (NetEngine.pack() is version-tracked: the weights are folded once)
t0 = time.time()
with hold:
  for k in range(n_patch):
    i = 1200 * (k % 108)
    rop, rdp = ro[0, i:i + 1200].contiguous(), rd[0, i:i + 1200].contiguous()
    fwd, ab = render_patch(m, rop, rdp, N_samples=128, N_importance=64, max_upsample_steps=6, perturb=False)
    G = 1e-3 * torch.randn(1200, 3, device=dev, generator=gen)
    eng.render_bwd(rop, rdp, ab, fwd, G, w_eikonal=0.1, eikonal_count=1200 * 192, white_bkgd=False, speed_factor=m.speed_factor)
    if k % 50 == 49:
        torch.cuda.synchronize(); print('patch', k, f'{time.time() - t0:.2f} s', flush=True)
torch.cuda.synchronize()
pairs, scal = eng.unpack_grads()
print('done', n_patch, f'{time.time() - t0:.2f} s', float(scal[1]), flush=True)
gp = eng._gpack.detach().cpu()
print('gradpack: finite', bool(torch.isfinite(gp).all()), 'abs max', float(gp.abs().max()), 'sum', float(gp.double().sum()), flush=True)
if os.environ.get('BW_LOOP_SAVE'): torch.save(gp, os.environ['BW_LOOP_SAVE'])
