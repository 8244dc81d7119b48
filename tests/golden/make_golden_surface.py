"""Generate tests/golden/surface_*.npz by running the UNMODIFIED reference's surface-rendering path on CPU (fp32).

Run ONLY in the build container (needs /root/reference):   python tests/golden/make_golden_surface.py
Pins models/ray_casting.py: root_finding_surface_points (35-160) + run_secant_method (11-30),
sphere_tracing_surface_points (163-184) and surface_render (187-263) on the seeded synthetic models of
make_golden.py (VolSDF bump 0.5 and NeuS bump 0 (the bumped NeuS fixture has no zero crossing), tilted 20x20 camera, plus a case whose near plane starts inside).
"""
import os, sys, warnings
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
import ref_shim
ref_shim.install()
warnings.filterwarnings('ignore')
import numpy as np
import torch
import torch.nn.functional as F
import fixtures as fx
from make_golden import build_volsdf, build_neus, rays_for, npy
from models import ray_casting as rc

torch.set_num_threads(os.cpu_count())


def run(out, name, framework, bump, H, W, scale_o, root_cfg, sphere_cfg):
    m = build_volsdf(0.1, bump) if framework == 'volsdf' else build_neus(0.05, bump)
    c2w, K, ro, rd = rays_for('tilted', H, W)
    ro = ro * scale_o
    S = dict(rays_o=npy(ro[0]), rays_d=npy(rd[0]), meta=np.array([H, W, scale_o, bump], dtype=np.float64))
    rdn = F.normalize(rd, dim=-1)
    with torch.no_grad():
        d, pt, mask, msc = rc.root_finding_surface_points(m.implicit_surface, ro.clone(), rdn.clone(), batched=True, **root_cfg)
        S['rf_d'] = npy(d[0]); S['rf_pt'] = npy(pt[0]); S['rf_mask'] = npy(mask[0]); S['rf_mask_sign_change'] = npy(msc[0])
        d, pt, mask = rc.sphere_tracing_surface_points(m.implicit_surface, ro.clone(), rdn.clone(), batched=True, **sphere_cfg)
        S['st_d'] = npy(d[0]); S['st_pt'] = npy(pt[0]); S['st_mask'] = npy(mask[0])
        for algo, cfg in (('root_finding', root_cfg), ('sphere_tracing', sphere_cfg)):
            col, dep, ex = rc.surface_render(ro.clone(), rd.clone(), m, calc_normal=True, rayschunk=8192, batched=True,
                                             use_view_dirs=True, ray_casting_algo=algo, ray_casting_cfgs=dict(cfg))
            tag = 'sr_rf' if algo == 'root_finding' else 'sr_st'
            S[tag + '_rgb'] = npy(col[0]); S[tag + '_depth'] = npy(dep[0])
            S[tag + '_nablas'] = npy(ex['implicit_nablas'][0]); S[tag + '_mask'] = npy(ex['mask_surface'][0])
            S[tag + '_normals'] = npy(ex['normals_surface'][0])
    np.savez_compressed(os.path.join(out, name + '.npz'), **S)
    print(name, 'root-finding hits', int(S['rf_mask'].sum()), '/', H * W, ' sphere-tracing alive', int(S['st_mask'].sum()),
          ' depth range', (float(S['rf_d'][S['rf_mask']].min()), float(S['rf_d'][S['rf_mask']].max())) if S['rf_mask'].any() else None,
          ' first point occupied', int((S['rf_d'] == 0).sum()))


if __name__ == '__main__':
    out = HERE
    # render.py's surface-render call uses the defaults of root_finding_surface_points (N_steps 256, 8 secant steps, near 0, far 6)
    run(out, 'surface_volsdf', 'volsdf', 0.5, 20, 20, 1.0, dict(near=0.0, far=6.0, N_steps=256, N_secant_steps=8, logit_tau=0.0, fill_inf=True),
        dict(near=0.0, far=6.0, N_iters=20))
    # NeuS-scaled scene; camera close enough that some rays start inside / graze (mask_0_not_occupied, pos->neg checks), fill_inf=False
    run(out, 'surface_neus', 'neus', 0.0, 20, 20, 0.3, dict(near=0.0, far=2.5, N_steps=128, N_secant_steps=8, logit_tau=0.0, fill_inf=False),
        dict(near=0.0, far=2.5, N_iters=20))
    run(out, 'surface_volsdf_inside', 'volsdf', 0.5, 12, 12, 0.3, dict(near=0.0, far=4.0, N_steps=64, N_secant_steps=4, logit_tau=0.05, fill_inf=True),
        dict(near=0.2, far=4.0, N_iters=8))
