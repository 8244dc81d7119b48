"""GPU: the CLIP ViT-B/32 image tower kernels (csrc/clip_vit.cu) and the three style losses built on them.

Parity is UNPINNED against the reference's own model: openai/CLIP and its weights are not available offline.
The stand-in oracle is `transformers.CLIPVisionModelWithProjection(CLIPVisionConfig())` -- the same architecture
(768/12/12/3072, patch 32, quick_gelu, proj 512) -- with seeded random weights, fp32, on the same GPU; forward features and
d loss / d image must agree with its autograd.  Tolerances: 2e-4 relative to the largest entry (fp32 kernels; the stand-in's
cuBLAS GEMMs may use a different summation order).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def towers():
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    import nerfart_b200  # noqa: F401
    from nerfart_b200.criteria.clip_vit import ClipVisionB32
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    hf = CLIPVisionModelWithProjection(CLIPVisionConfig())
    with torch.no_grad():                                   # HF's init is tiny (std 0.02): scale up so every block matters
        for n, p in hf.named_parameters():
            if p.dim() >= 2:
                p.mul_(3.0)
            elif 'bias' in n:
                p.normal_(0, 0.1)
    hf = hf.to(DEV).float().eval()
    for p in hf.parameters():
        p.requires_grad_(False)
    return hf, ClipVisionB32.from_hf(hf, DEV, precision='fp32')


@pytest.fixture(scope='module')
def tower_tf32(towers):
    from nerfart_b200.criteria.clip_vit import ClipVisionB32
    return ClipVisionB32.from_hf(towers[0], DEV, precision='tf32')


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.mark.parametrize('B', [1, 3, 14])
def test_encode_image_forward_and_image_gradient(towers, B):
    hf, tower = towers
    g = torch.Generator(device='cpu'); g.manual_seed(B)
    x = torch.randn(B, 3, 224, 224, generator=g).to(DEV)
    gf = torch.randn(B, 512, generator=g).to(DEV)
    x1 = x.clone().requires_grad_(True)
    ref = hf(pixel_values=x1).image_embeds
    (ref * gf).sum().backward()
    x2 = x.clone().requires_grad_(True)
    out = tower.encode_image(x2)
    (out * gf).sum().backward()
    e_f, e_g = rel(out, ref), rel(x2.grad, x1.grad)
    print(f'B={B}: features rel err {e_f:.2e} (max |f| {float(ref.abs().max()):.3f}), image-gradient rel err {e_g:.2e}')
    assert e_f < 2e-4 and e_g < 2e-4


@pytest.mark.parametrize('B', [1, 2, 3, 14])
def test_tf32_tensor_core_tower_forward_and_image_gradient(towers, tower_tf32, B):
    """The default arithmetic: every linear layer on tcgen05 kind::tf32 (csrc/tgemm.cu; 11-bit operands, fp32 accumulation in
    TMEM; the reference runs the tower in fp16).  Tolerances are those of 10-bit-mantissa operands through 12 blocks, against
    the fp32 stand-in: 1e-2 of the largest feature, 3e-2 of the largest image-gradient entry, cosine > 0.9999."""
    import nerfart_b200
    hf, _ = towers
    g = torch.Generator(device='cpu'); g.manual_seed(100 + B)
    x = torch.randn(B, 3, 224, 224, generator=g).to(DEV)
    gf = torch.randn(B, 512, generator=g).to(DEV)
    x1 = x.clone().requires_grad_(True)
    ref = hf(pixel_values=x1).image_embeds
    (ref * gf).sum().backward()
    x2 = x.clone().requires_grad_(True)
    n0 = nerfart_b200.launch_count()
    out = tower_tf32.encode_image(x2)
    (out * gf).sum().backward()
    assert nerfart_b200.launch_count() > n0
    e_f, e_g = rel(out, ref), rel(x2.grad, x1.grad)
    cos_f = float(F.cosine_similarity(out.flatten(), ref.flatten(), dim=0))
    cos_g = float(F.cosine_similarity(x2.grad.flatten(), x1.grad.flatten(), dim=0))
    print(f'tf32 B={B}: features rel err {e_f:.2e} cos {cos_f:.6f}, image-gradient rel err {e_g:.2e} cos {cos_g:.6f}')
    assert e_f < 1e-2 and e_g < 3e-2 and cos_f > 0.9999 and cos_g > 0.9995


def test_two_encodes_before_backward_keep_their_own_activations(towers):
    hf, tower = towers
    g = torch.Generator(device='cpu'); g.manual_seed(9)
    xa = torch.randn(2, 3, 224, 224, generator=g).to(DEV).requires_grad_(True)
    xb = torch.randn(1, 3, 224, 224, generator=g).to(DEV).requires_grad_(True)
    fa, fb = tower.encode_image(xa), tower.encode_image(xb)
    (fa.sum() * 2 + (fb ** 2).sum()).backward()
    xa2, xb2 = xa.detach().clone().requires_grad_(True), xb.detach().clone().requires_grad_(True)
    ra, rb = hf(pixel_values=xa2).image_embeds, hf(pixel_values=xb2).image_embeds
    (ra.sum() * 2 + (rb ** 2).sum()).backward()
    assert rel(xa.grad, xa2.grad) < 2e-4 and rel(xb.grad, xb2.grad) < 2e-4


class _FakeText:
    """deterministic stand-in for the text tower: class string -> normalised [5,512] features; counts encodes"""
    def __init__(self):
        self.n = 0

    def encode(self, strings):
        self.n += 1
        out = []
        for s in strings:
            gen = torch.Generator(device='cpu'); gen.manual_seed(abs(hash(s)) % (2 ** 31))
            out.append(torch.randn(512, generator=gen))
        return torch.stack(out).to(DEV)


def _restated_losses(hf, tf, target_hw, src_img, pred, s_text, t_text, negs, crops, is_full_res):
    """the reference's formulas (clip_loss.py:248-254, contrastive_loss.py:139-153, patchnce_loss.py:146-220) written out with
    plain torch ops and the stand-in model as `encode_image`"""
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073], device=DEV).view(1, 3, 1, 1)
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711], device=DEV).view(1, 3, 1, 1)
    enc = lambda x: F.normalize(hf(pixel_values=(x - mean) / std).image_embeds, dim=-1)
    # directional
    pre = lambda x: F.interpolate(x, size=(224, 224), mode='bicubic', align_corners=False)
    tdir = F.normalize((tf(t_text) - tf(s_text)).mean(0, keepdim=True), dim=-1)
    ed = F.normalize(enc(pre(pred)) - enc(pre(src_img)), dim=-1)
    l_clip = (1 - F.cosine_similarity(ed, tdir)).mean()
    # global contrastive (inputs are treated as [-1,1] images: (x+1)/2; resize short side to 224 bicubic; centre crop)
    def pre2(x):
        x = (x + 1) / 2
        h, w = x.shape[-2:]
        nh, nw = (int(224 * h / w), 224) if w <= h else (224, int(224 * w / h))
        x = F.interpolate(x, size=(nh, nw), mode='bicubic', align_corners=False)
        t, l = int(round((nh - 224) / 2.)), int(round((nw - 224) / 2.))
        return x[..., t:t + 224, l:l + 224]
    te, se = enc(pre2(pred)), enc(pre2(src_img))
    d = lambda a, b: F.pairwise_distance(a, b, keepdim=True)
    l_con = torch.mean(d(te, tf(t_text)) ** 2 + torch.clamp(2.0 - d(te, tf(negs[0])), min=0) ** 2 + torch.clamp(2.0 - d(te, se.detach()), min=0) ** 2)
    # patch NCE
    img = F.interpolate(F.pad(pred, (270, 270, 480, 480)), size=tuple(target_hw), mode='bicubic', align_corners=False)
    th = 224 if is_full_res else 112
    l_nce = 0
    for (i, j) in crops:
        c = img[..., i:i + th, j:j + th]
        if not is_full_res:
            c = F.interpolate(c, size=(224, 224), mode='bicubic', align_corners=False)
        c = F.interpolate((c + 1) / 2, size=(224, 224), mode='bilinear', align_corners=False)
        e = enc(c)
        pos = torch.exp(F.cosine_similarity(e, tf(t_text)) / 0.07)
        neg = sum(torch.exp(F.cosine_similarity(e, tf(n)) / 0.07) for n in negs)
        l_nce = l_nce + torch.mean(-torch.log(pos / (pos + neg)))
    return l_clip, l_con, l_nce


def test_style_losses_match_restated_reference_formulas_and_cache_text(towers):
    hf, tower = towers
    from nerfart_b200.criteria import make_loss_dict, TextFeatures
    fake = _FakeText()
    text = TextFeatures(fake.encode, templates=['a photo of a {}.', 'a painting of a {}.', '{}', 'art of the {}.', 'the {} in a game.'])
    H, W = 96, 54
    target_hw = [480, 270]
    ld = make_loss_dict(tower, text, target_hw)
    g = torch.Generator(device='cpu'); g.manual_seed(4)
    src = torch.rand(1, 3, H, W, generator=g).to(DEV)
    pred = torch.rand(1, 3, H, W, generator=g).to(DEV).requires_grad_(True)
    negs = [f'neg{i}' for i in range(8)]
    torch.manual_seed(11)
    crops = ld['patchnce'].sample_crops(480, 270, 112, 112, False)
    torch.manual_seed(11)
    l1 = ld['clip'](src, 'photo', pred, 'painting')
    l2 = ld['contrastive'](src, negs[0], pred, 'painting')
    l3 = ld['patchnce'](negs, pred, 'painting', False)
    (l1 + 0.2 * l2 + 0.1 * l3).backward()
    n_enc = fake.n
    assert n_enc == 2 + 8                                   # 'photo', 'painting', 8 negatives: each class string encoded once
    ld['patchnce'](negs, pred.detach(), 'painting', False); ld['contrastive'](src, negs[0], pred.detach(), 'painting')
    assert fake.n == n_enc                                   # nothing re-encoded (the reference would run 2*79 + 12*9*79 text passes)
    pred2 = pred.detach().clone().requires_grad_(True)
    r1, r2, r3 = _restated_losses(hf, lambda s: text(s), target_hw, src, pred2, 'photo', 'painting', negs, crops, False)
    (r1 + 0.2 * r2 + 0.1 * r3).backward()
    print('losses', float(l1), float(r1), float(l2), float(r2), float(l3), float(r3), 'grad rel err', rel(pred.grad, pred2.grad))
    for a, b in ((l1, r1), (l2, r2), (l3, r3)):
        assert abs(float(a) - float(b)) <= 2e-4 * max(1.0, abs(float(b)))
    assert rel(pred.grad, pred2.grad) < 1e-3


def test_from_openai_state_dict_layout_and_build_loss_dict():
    """`ClipVisionB32.from_openai` on a model in the UPSTREAM openai/CLIP module / state-dict layout (tests/stubs/clip: the public
    ViT-B/32 definition in plain PyTorch, seeded weights with non-trivial LayerNorm affines and biases): key mapping, the
    `visual.proj` orientation ([768, 512], x @ proj), in_proj packing and the conv1 patch order are all exercised -- features and the
    image gradient must equal the upstream module's autograd.  Then `criteria.build_loss_dict` (what Trainer.__init__ calls,
    volsdf.py:639-645) is built from the same `clip` module: one shared tower, cached text features, the four losses."""
    import os
    import sys
    stubs = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'stubs')
    sys.path.append(stubs)
    try:
        import clip
        if not hasattr(clip, 'FakeCLIP'):
            pytest.skip('a real `clip` package is installed; this test pins the layout against the synthetic one')
        from nerfart_b200.criteria.clip_vit import ClipVisionB32
        torch.backends.cuda.matmul.allow_tf32 = False
        model, _ = clip.load('ViT-B/32', device=DEV)
        model = model.float().eval()
        with torch.no_grad():
            for n, p in model.visual.named_parameters():           # upstream init is tiny: scale the matrices so every block matters
                if p.dim() >= 2 and 'positional' not in n:
                    p.mul_(2.0)
        tower = ClipVisionB32.from_openai(model, DEV, precision='fp32')
        g = torch.Generator(device='cpu'); g.manual_seed(3)
        x = torch.randn(3, 3, 224, 224, generator=g).to(DEV)
        gf = torch.randn(3, 512, generator=g).to(DEV)
        x1 = x.clone().requires_grad_(True)
        ref = model.encode_image(x1)
        (ref * gf).sum().backward()
        x2 = x.clone().requires_grad_(True)
        out = tower.encode_image(x2)
        (out * gf).sum().backward()
        e_f, e_g = rel(out, ref.detach()), rel(x2.grad, x1.grad)
        print('from_openai: features', e_f, 'image gradient', e_g)
        assert e_f < 2e-4 and e_g < 2e-4
        # the factory Trainer uses, on the same module (VGG16 weights are not available offline: seeded stand-in, criteria/perceptual.py)
        os.environ['NA_VGG16_WEIGHTS'] = 'random:0'
        from nerfart_b200.criteria import build_loss_dict
        ld = build_loss_dict([480, 270], DEV)
        assert set(ld) == {'clip', 'contrastive', 'patchnce', 'perceptual'} and ld['perceptual'] is not None
        img = torch.rand(1, 3, 480, 270, device=DEV)
        pred = (img * 0.9 + 0.05).requires_grad_(True)
        from criteria_templates import TEMPLATES                     # any template list: the cache is keyed by class string
        for k in ('clip', 'contrastive', 'patchnce'):
            ld[k].text.templates = TEMPLATES
        loss = ld['clip'](img, 'photo', pred, 'painting') + ld['perceptual'](pred, img) + ld['contrastive'](img, 'sketch', pred, 'painting')
        loss.backward()
        assert torch.isfinite(loss) and torch.isfinite(pred.grad).all() and float(pred.grad.abs().max()) > 0
    finally:
        sys.path.remove(stubs)
        os.environ.pop('NA_VGG16_WEIGHTS', None)
