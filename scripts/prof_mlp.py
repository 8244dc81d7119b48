"""Small fixed workload for `ncu --set full`: one SDF-only and one full launch of the MLP kernel (1 Mi / 256 Ki points)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import torch
from helpers import make_volsdf
import nerfart_b200
prec = sys.argv[1] if len(sys.argv) > 1 else 'tc'
dev = 'cuda:0'
m = make_volsdf(0.1, 0.0, device=dev)
m.engine().precision = prec
g = torch.Generator(device=dev); g.manual_seed(1)
n = 1024 * 1024
x = torch.rand(n, 3, device=dev, generator=g) * 4 - 2
v = torch.nn.functional.normalize(torch.randn(n // 4, 3, device=dev, generator=g), dim=-1)
with torch.no_grad():
    m.engine().sdf_eval(x, apply_bg=True)
    m.engine().full_eval(x[:n // 4], v, want_feat=False)
torch.cuda.synchronize()
print('done')
