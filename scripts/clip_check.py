"""CLIP ViT-B/32 image tower: tf32 (tcgen05) vs fp32 (CUDA cores) linear layers -- agreement and time per encode (fwd, fwd+bwd)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nerfart_b200
from nerfart_b200.criteria.clip_vit import ClipVisionB32

dev = 'cuda:0'
towers = {p: ClipVisionB32.random(0, dev, precision=p) for p in ('fp32', 'tf32')}
torch.cuda.synchronize()
for B in (1, 2, 12, 14):
    g = torch.Generator(device='cpu'); g.manual_seed(B)
    x = torch.rand(B, 3, 224, 224, generator=g).to(dev)
    gf = torch.randn(B, 512, generator=g).to(dev)
    res = {}
    for p, tw in towers.items():
        def fwd():
            with torch.no_grad():
                return tw.encode_image(x)
        def fwdbwd():
            xi = x.clone().requires_grad_(True)
            f = tw.encode_image(xi); (f * gf).sum().backward()
            return f.detach(), xi.grad
        t = {}
        for name, fn in (('fwd', fwd), ('fwd+bwd', fwdbwd)):
            fn(); fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = nerfart_b200.launch_count()
            e0.record()
            for _ in range(5): out = fn()
            e1.record(); torch.cuda.synchronize()
            t[name] = (e0.elapsed_time(e1) / 5, (nerfart_b200.launch_count() - n0) // 5)
        res[p] = (out, t)
    (f32, g32), (f19, g19) = res['fp32'][0], res['tf32'][0]
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    print(f'B={B:2d}  fp32: fwd {res["fp32"][1]["fwd"][0]:7.3f} ms, fwd+bwd {res["fp32"][1]["fwd+bwd"][0]:7.3f} ms | '
          f'tf32: fwd {res["tf32"][1]["fwd"][0]:7.3f} ms ({res["tf32"][1]["fwd"][1]} launches), fwd+bwd {res["tf32"][1]["fwd+bwd"][0]:7.3f} ms '
          f'({res["tf32"][1]["fwd+bwd"][1]} launches) | tf32 vs fp32: feats {rel(f19, f32):.2e}, image grad {rel(g19, g32):.2e}', flush=True)
