"""One warm-up + one profiled forward/backward of the CLIP tower (for ncu launch lists): python scripts/prof_clip.py [precision] [B]."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nerfart_b200.criteria.clip_vit import ClipVisionB32
prec = sys.argv[1] if len(sys.argv) > 1 else 'tf32'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
tw = ClipVisionB32.random(0, 'cuda:0', precision=prec)
x = torch.rand(B, 3, 224, 224, device='cuda:0')
for _ in range(2):
    xi = x.clone().requires_grad_(True)
    f = tw.encode_image(xi); f.square().sum().backward()
    torch.cuda.synchronize()
