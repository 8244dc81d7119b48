"""Small fixed workload for `ncu --set full` on the backward kernels: two patches of 1200 rays x 192 samples.
PROF_SPLIT=1 (default): the split training program (forward render with the stash + backward-only launch); 0: the one-launch program."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import torch
from helpers import make_volsdf, fx
import nerfart_b200
from nerfart_b200.models.frameworks.volsdf import render_patch
from nerfart_b200.utils import rend_util
dev = 'cuda:0'
m = make_volsdf(0.1, 0.0, device=dev)
H, W = 480, 270
c2w, K = fx.closed_form_camera(H, W)
with torch.no_grad():
    ro, rd, _ = rend_util.get_rays(c2w[None].to(dev), K[None].to(dev), H, W)
eng = m.engine(); eng.grad_zero()
for i in (60000, 61200):
    rop, rdp = ro[0, i:i + 1200].contiguous(), rd[0, i:i + 1200].contiguous()
    fwd, ab = render_patch(m, rop, rdp, N_samples=128, N_importance=64, max_upsample_steps=6, train_stash=os.environ.get('PROF_SPLIT', '1') != '0')
    G = torch.full((1200, 3), 1e-3, device=dev)
    eng.render_bwd(rop, rdp, ab, fwd, G, w_eikonal=0.1, eikonal_count=1200 * 192, white_bkgd=False, speed_factor=m.speed_factor)
torch.cuda.synchronize()
print('done')
