"""Numerics + speed check of the tensor-core MLP path against the fp32 path and the reference golden vectors."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import numpy as np, torch
from helpers import golden, make_volsdf, make_neus, linf
import nerfart_b200

dev = 'cuda:0'
S = golden('stages')
small = len(sys.argv) > 1 and sys.argv[1] == 'small'
TC_MODES = tuple(os.environ.get('NA_CHECK_MODES', 'tc,tc_mixed,tc2acc').split(','))
for tag in ('v', 'n'):
    m = make_volsdf(0.01, 0.5, device=dev) if tag == 'v' else make_neus(0.05, 0.5, device=dev)
    x = torch.tensor(S[f'net_{tag}_x'], device=dev); v = torch.tensor(S[f'net_{tag}_v'], device=dev)
    res = {}
    for prec in ('fp32',) + TC_MODES:
        m.engine().precision = prec
        with torch.no_grad():
            sdf, feat = m.implicit_surface.forward(x, return_h=True)
            torch.cuda.synchronize(); print(tag, prec, 'sdf_eval ok', flush=True)
            rad, sdf2, nab = m.forward(x, v)
            torch.cuda.synchronize(); print(tag, prec, 'full_eval ok', flush=True)
        res[prec] = dict(sdf=sdf.cpu().numpy(), feat=feat.cpu().numpy(), rad=rad.cpu().numpy(), sdf2=sdf2.cpu().numpy(), nab=nab.cpu().numpy())
    G = dict(sdf=S[f'net_{tag}_sdf'], feat=S[f'net_{tag}_feat'], nab=S[f'net_{tag}_nabla'])
    for prec in ('fp32',) + TC_MODES:
        R = res[prec]
        print(f'{tag} {prec:9s} Linf vs reference: sdf {linf(R["sdf"], G["sdf"]):.3e} feat {linf(R["feat"], G["feat"]):.3e} nab {linf(R["nab"], G["nab"]):.3e}'
              f' | vs fp32 kernel: rad {linf(R["rad"], res["fp32"]["rad"]):.3e} sdf(full) {linf(R["sdf2"], res["fp32"]["sdf2"]):.3e}', flush=True)
if small:
    sys.exit(0)
# speed
m = make_volsdf(0.1, 0.0, device=dev)
g = torch.Generator(device=dev); g.manual_seed(1)
n = 4 * 1024 * 1024
x = torch.rand(n, 3, device=dev, generator=g) * 4 - 2
v = torch.nn.functional.normalize(torch.randn(n // 4, 3, device=dev, generator=g), dim=-1)
F_SDF = 2 * (39 * 256 + 256 * 256 * 2 + 256 * 217 + 256 * 256 * 4 + 256); F_FULL = 2 * (524544 + 459008 + 265216)
import ctypes as C
dbg = torch.zeros(1024, dtype=torch.int64, device=dev)
trace_dir = os.environ.get('NA_TRACE_DIR')
for prec in ('fp32',) + TC_MODES:
    m.engine().precision = prec
    eng = m.engine()
    nerfart_b200.lib().na_debug_set_buffer(C.c_void_p(dbg.data_ptr()) if prec != 'fp32' else None)
    for fn, cnt, fl, name in ((lambda: eng.sdf_eval(x, apply_bg=True), n, F_SDF, 'sdf-only'), (lambda: eng.full_eval(x[:n // 4], v, want_feat=False), n // 4, F_FULL, 'full')):
        fn(); torch.cuda.synchronize(); dbg.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); fn(); e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3 / 2
        print(f'{prec:9s} {name:9s} {cnt/t/1e6:9.2f} Msamples/s  {cnt*fl/t/1e12:8.2f} TFLOP/s (algorithmic)  {t*1e3:8.2f} ms')
        if prec != 'fp32':
            if trace_dir: np.save(os.path.join(trace_dir, f'trace_{prec}_{name}.npy'), dbg.cpu().numpy())
            d = dbg.cpu().tolist(); ntile = (cnt + 127) // 128 / 148          # (cycle counters: -DNA_TM_CYCLES builds only)
            if d[0] > 0: print(f'      CTA0 cycles/tile: mma-thread total {d[0]/ntile:9.0f}  wait a_ready {d[1]/ntile:9.0f}  wait weights {d[2]/ntile:9.0f} | epilogue total {d[3]/ntile:9.0f}  wait d_ready {d[4]/ntile:9.0f}')
