"""Where the style-loss time of a fine-tune step goes: wall clock and CUDA-event time of calc_style_loss forward / backward, the image
tower alone, and a torch.profiler op table (host-bound glue shows up as many tiny kernels)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import torch
import nerfart_b200
from nerfart_b200.criteria import make_loss_dict, TextFeatures
from nerfart_b200.criteria.clip_vit import ClipVisionB32
from nerfart_b200.models.frameworks import _finetune
dev = torch.device('cuda:0')
H, W = 480, 270


def fake_text(strings):
    out = []
    for s_ in strings:
        gg = torch.Generator(device='cpu'); gg.manual_seed(sum(ord(c) * (i + 1) for i, c in enumerate(s_)) % (2 ** 31))
        out.append(torch.randn(512, generator=gg))
    return torch.stack(out).to(dev)


class _A(dict):
    __getattr__ = dict.__getitem__


tower = ClipVisionB32.random(0, dev)
loss_dict = make_loss_dict(tower, TextFeatures(fake_text, templates=['a photo of a {}.'] * 79), [H, W])


class T:                       # the attributes calc_style_loss reads from the Trainer
    pass


t = T(); t.loss_dict = loss_dict; t.neg_texts = [f'negative prompt {i}' for i in range(40)]
for k in ('clip', 'perceptual', 'contrastive', 'patchnce'):
    setattr(t, k + '_loss', loss_dict.get(k))
args = _A(training=_A(is_finetune=True), data=_A(downscale=2), model=_A(radiance=_A(use_view_dirs=True)),
          finetune=_A(use_eikonal=True, w_eikonal=0.1, w_clip=1.0, w_perceptual=2.0, w_contrastive=0.2, w_patchnce=0.1,
                      src_text='photo', target_text='painting'))
g = torch.Generator(device='cpu'); g.manual_seed(1)
rgb0 = torch.rand(1, H * W, 3, generator=g).to(dev); gt = torch.rand(1, H * W, 3, generator=g).to(dev)


def one(profile=False):
    rgb = rgb0.clone().requires_grad_(True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    loss = _finetune.calc_style_loss(t, rgb, gt, args, H)
    t1h = time.perf_counter(); torch.cuda.synchronize(); t1 = time.perf_counter()
    loss.backward()
    t2h = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1h - t0, t1 - t0, t2h - t1, t2 - t1, float(loss))


for _ in range(3):
    one()
rs = [one() for _ in range(5)]
for r in rs:
    print('style fwd: host %.1f ms, done %.1f ms | bwd: host %.1f ms, done %.1f ms | loss %.5f' % (1e3 * r[0], 1e3 * r[1], 1e3 * r[2], 1e3 * r[3], r[4]))
# the tower alone
for B in (1, 12):
    x = torch.randn(B, 3, 224, 224, device=dev, requires_grad=True)
    for _ in range(3):
        tower.encode_image(x).sum().backward()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        f = tower.encode_image(x)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    for _ in range(10):
        tower.encode_image(x).sum().backward()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print('tower B=%d: fwd %.2f ms, fwd+bwd %.2f ms' % (B, 1e2 * (t1 - t0), 1e2 * (t2 - t1)))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    one()
print(prof.key_averages().table(sort_by='self_cpu_time_total', row_limit=30, max_name_column_width=50))
