"""rgb / depth / normals L-inf of the tensor-core modes on every beta = 0.1 (BASELINE-shaped) golden fixture of the reference, and on
small-beta fixtures for the path-consistent rays: the evidence for (not) making tc_mixed the default render mode."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import numpy as np, torch
from helpers import linf
import test_gpu_parity as tp
for prec in ('fp32', 'tc', 'tc_mixed'):
    for name, bump in (('volsdf_cfg1_b0.1', 0.0), ('volsdf_det_b0.1', 0.5), ('volsdf_n128_b0.1', 0.0), ('volsdf_n128_b0.01', 0.5), ('volsdf_det_b0.01', 0.5)):
        G, o = tp._render_volsdf(name, bump, prec=prec)
        n = G['rgb'].shape[0]
        same = np.ones(n, bool)
        if 'iter_usage' in G and 'iter_usage' in o:
            bm_o = np.asarray(o['beta_map']).reshape(n); bm_g = G['beta_map'].reshape(n)
            same = (np.asarray(o['iter_usage']).reshape(n) == G['iter_usage'].reshape(n)) & (np.abs(bm_o - bm_g) <= 2e-6 * np.abs(bm_g))
        rec = {'precision': prec, 'fixture': name, 'path_consistent': round(float(same.mean()), 4)}
        for k in ('rgb', 'depth_volume', 'normals_volume'):
            if k in o and k in G:
                e = np.abs(np.asarray(o[k]) - G[k]).reshape(n, -1).max(1)
                rec[k] = {'linf_consistent_rays': float(e[same].max()) if same.any() else None, 'median': float(np.median(e[same])) if same.any() else None}
        print(json.dumps(rec), flush=True)
