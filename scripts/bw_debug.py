"""Compare the gradients of the tcgen05 backward program (default) with the loaded-mode SIMT backward (NA_BWD_TMEM=0), per tensor and per NeuS pass."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import numpy as np, torch
from helpers import make_volsdf, make_neus, golden
from test_gpu_train import product_grads, neus_fwd_at, volsdf_fwd_at, t
which = sys.argv[1] if len(sys.argv) > 1 else 'neus'
if which == 'neus':
    g = golden('train_neus'); m = make_neus(float(g['variance_init']), float(g['bump']), device='cuda:0')
    m.engine().precision = 'tc'
    ro, rd = t(g['rays_o']), t(g['rays_d']); fwd = neus_fwd_at(m, ro, rd, t(g['d_all']))
    run = lambda: product_grads(m, 'neus', ro, rd, fwd, t(g['G']), float(g['w_eikonal']), False, train_radiance=False)
else:
    g = golden('train_volsdf_b0.1'); m = make_volsdf(float(g['beta_init']), float(g['bump']), device='cuda:0')
    m.engine().precision = 'tc'
    ro, rd = t(g['rays_o']), t(g['rays_d']); fwd = volsdf_fwd_at(m, ro, rd, t(g['d_vals']))
    run = lambda: product_grads(m, 'volsdf', ro, rd, fwd, t(g['G']), float(g['w_eikonal']), bool(g['white_bkgd']))
for ps in ((None, 'A', 'B') if which == 'neus' else (None,)):
    if ps: os.environ['NA_BWD_DEBUG_PASS'] = ps
    else: os.environ.pop('NA_BWD_DEBUG_PASS', None)
    os.environ['NA_BWD_TMEM'] = '0'; ga, _ = run()
    os.environ['NA_BWD_TMEM'] = '1'; gb, _ = run()
    print('=== pass', ps or 'both')
    for k in ga:
        a, b = ga[k], gb[k]
        sc = np.abs(a).max() + 1e-30
        print(f'  {k:55s} scale {sc:.3e}  Linf {np.abs(a - b).max() / sc:.3e}  L2 {np.linalg.norm((a - b).ravel()) / (np.linalg.norm(a.ravel()) + 1e-30):.3e}  finite {np.isfinite(b).all()}')
