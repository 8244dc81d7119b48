"""Per-GEMM clock trace of ONE tile of the training program (NA_TM_TRACE build, scripts/build_variant.sh trace -DNA_TM_TRACE):
runs two patches, saves CTA 0's second-tile stamps of the last BW launch; read with scripts/trace_show.py <npy> 41."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import numpy as np, torch
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
dbg = torch.zeros(1024, dtype=torch.int64, device=dev)
for i in (60000, 61200):
    rop, rdp = ro[0, i:i + 1200].contiguous(), rd[0, i:i + 1200].contiguous()
    fwd, ab = render_patch(m, rop, rdp, N_samples=128, N_importance=64, max_upsample_steps=6)
    G = torch.full((1200, 3), 1e-3, device=dev)
    torch.cuda.synchronize(); dbg.zero_()
    nerfart_b200.lib().na_debug_set_buffer(C.c_void_p(dbg.data_ptr()))
    eng.render_bwd(rop, rdp, ab, fwd, G, w_eikonal=0.1, eikonal_count=1200 * 192, white_bkgd=False, speed_factor=m.speed_factor)
    torch.cuda.synchronize()
    nerfart_b200.lib().na_debug_set_buffer(None)
d = dbg.cpu().numpy()
np.save(sys.argv[1] if len(sys.argv) > 1 else 'bw_trace.npy', d)
print('CTA0: mma total', d[0], 'wait a', d[1], 'wait weights', d[2], '| epilogue total', d[3], 'wait d', d[4])
