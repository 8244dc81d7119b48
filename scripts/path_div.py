"""Sampler-path divergence and SDF error of one arithmetic mode (run one process per NA_TM_ORDER / NA_TM_DEBIAS setting: the
library reads them once).  usage: python scripts/path_div.py <precision>"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import numpy as np, torch
from helpers import golden, make_volsdf, linf
import test_gpu_parity as tp
prec = sys.argv[1] if len(sys.argv) > 1 else 'tc'
out = {'precision': prec, 'order': os.environ.get('NA_TM_ORDER', '1'), 'debias': os.environ.get('NA_TM_DEBIAS', '0')}
S = golden('stages')
m = make_volsdf(0.01, 0.5, device='cuda:0'); m.engine().precision = prec
x = torch.tensor(S['net_v_x'], device='cuda:0')
with torch.no_grad():
    sdf, feat = m.implicit_surface.forward(x, return_h=True)
out['sdf_linf_vs_ref'] = linf(sdf.cpu().numpy(), S['net_v_sdf'])
out['sdf_rms_vs_ref'] = float(np.sqrt(np.mean((sdf.cpu().numpy().astype(np.float64) - S['net_v_sdf']) ** 2)))
out['sdf_mean_signed'] = float(np.mean(sdf.cpu().numpy().astype(np.float64) - S['net_v_sdf']))
for name in ('volsdf_det_b0.01', 'volsdf_det_b0.002', 'volsdf_n128_b0.01'):
    G, o = tp._render_volsdf(name, 0.5, prec=prec)
    n = G['rgb'].shape[0]
    bm_o = np.asarray(o['beta_map']).reshape(n); bm_g = G['beta_map'].reshape(n)
    same = (np.asarray(o['iter_usage']).reshape(n) == G['iter_usage'].reshape(n)) & (np.abs(bm_o - bm_g) <= 2e-6 * np.abs(bm_g))
    conv = G['iter_usage'].reshape(n) >= 0
    out[name] = {'divergent': round(float(1 - same.mean()), 4), 'conv_same': round(float((same | ~conv).mean()), 4),
                 'rgb_linf': round(linf(o['rgb'], G['rgb']), 5)}
print(json.dumps(out))
