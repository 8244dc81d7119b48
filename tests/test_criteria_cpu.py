"""CPU: host logic of the criteria package -- text-feature cache, crop sampling order, no CPU fallback of the image tower."""
import pytest
import torch

import nerfart_b200  # noqa: F401
from nerfart_b200.criteria import TextFeatures
from nerfart_b200.criteria.losses import PatchNCELoss, _resize_short_side, _center_crop
from nerfart_b200.criteria.clip_vit import ClipVisionB32


def test_text_features_are_encoded_once_per_class_string_and_normalised():
    calls = []

    def enc(strings):
        calls.append(list(strings))
        return torch.arange(len(strings) * 4, dtype=torch.float32).reshape(len(strings), 4) + 1
    t = TextFeatures(enc, templates=['a {}', 'the {}!'])
    a = t('cat'); b = t('cat'); c = t('dog')
    assert a is b and len(calls) == 2 and calls[0] == ['a cat', 'the cat!']
    assert torch.allclose(a.norm(dim=-1), torch.ones(2)) and a.shape == c.shape == (2, 4)
    raw = t('cat', norm=False)
    assert len(calls) == 3 and not torch.allclose(raw.norm(dim=-1), torch.ones(2))


def test_patchnce_crop_draws_follow_the_reference_rng_order():
    # patchnce_loss.py:196-212: per crop randint(0,H-th+1) [discarded], randint(100,H-th+1-100), randint(0,W-tw+1)
    p = PatchNCELoss(None, None, [480, 270])
    torch.manual_seed(5)
    got = p.sample_crops(480, 270, 112, 112, False)
    torch.manual_seed(5)
    want = []
    for _ in range(12):
        torch.randint(0, 480 - 112 + 1, size=(1,))
        i = torch.randint(100, 480 - 112 + 1 - 100, size=(1,)).item()
        j = torch.randint(0, 270 - 112 + 1, size=(1,)).item()
        want.append((i, j))
    assert got == want and all(100 <= i < 269 and 0 <= j <= 158 for i, j in got)


def test_resize_short_side_and_center_crop_shapes():
    x = torch.rand(1, 3, 96, 54)
    y = _resize_short_side(x, 224, 'bicubic')
    assert y.shape[-2:] == (int(224 * 96 / 54), 224)
    assert _center_crop(y, 224).shape[-2:] == (224, 224)


def test_image_tower_refuses_cpu():
    with pytest.raises(RuntimeError):
        ClipVisionB32({}, 'cpu')


def test_build_loss_dict_fails_loudly_without_clip():
    from nerfart_b200.criteria import build_loss_dict
    try:
        import clip  # noqa: F401
        pytest.skip('openai clip is installed here')
    except ImportError:
        pass
    with pytest.raises(RuntimeError, match='clip'):
        build_loss_dict([480, 270], 'cpu')


def test_perceptual_loss_never_drops_silently(monkeypatch, tmp_path):
    """criteria/perp_loss.py is part of the optimised objective (volsdf.py:898-900): without the VGG16 weights construction raises."""
    from nerfart_b200.criteria import perceptual
    monkeypatch.delenv('NA_VGG16_WEIGHTS', raising=False)
    monkeypatch.setattr(torch.hub, 'get_dir', lambda: str(tmp_path))
    import torchvision.models as tvm

    def offline(*a, **k):
        raise OSError('no network')
    monkeypatch.setattr(tvm, 'vgg16', offline)
    with pytest.raises(RuntimeError, match='not dropped silently'):
        perceptual.VGGPerceptualLoss()
    assert perceptual.VGGPerceptualLoss(weights='random:1') is not None


REF_PERP = '/root/reference/criteria/perp_loss.py'


@pytest.mark.skipif(not __import__('os').path.exists(REF_PERP), reason='reference tree not present')
def test_perceptual_loss_equals_the_reference_module_on_the_same_weights(monkeypatch):
    """VGGPerceptualLoss against the UNMODIFIED criteria/perp_loss.py holding the same (seeded) VGG16 weights: loss and image
    gradient bit-identical on CPU (same ops in the same order; the reference's unused conv4 block is skipped here)."""
    import importlib.util
    import torchvision.models as tvm
    from nerfart_b200.criteria.perceptual import VGGPerceptualLoss, _load_state
    sd = _load_state('random:3')
    real = tvm.vgg16

    def seeded_vgg16(pretrained=False, **kw):
        m = real(weights=None)
        full = m.state_dict(); full.update(sd); m.load_state_dict(full)
        return m
    monkeypatch.setattr(tvm, 'vgg16', seeded_vgg16)
    spec = importlib.util.spec_from_file_location('ref_perp_loss', REF_PERP)
    ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
    R, M = ref.VGGPerceptualLoss().eval(), VGGPerceptualLoss(weights='random:3').eval()
    torch.manual_seed(0)
    a = torch.rand(1, 3, 60, 34, requires_grad=True); b = torch.rand(1, 3, 60, 34)
    lr = R(a, b); gr, = torch.autograd.grad(lr, a)
    lm = M(a, b); gm, = torch.autograd.grad(lm, a)
    assert float(lr.detach()) == float(lm.detach()) and torch.equal(gr, gm) and float(gr.abs().max()) > 0


def test_patchnce_batched_arithmetic_equals_the_reference_loop():
    """PatchNCELoss.forward evaluates all crops / classes / templates at once when B == 1; value and image gradient must equal the
    reference's per-crop loop (patchnce_loss.py:146-160,196-220), here restated literally on a stand-in image tower."""
    g = torch.Generator().manual_seed(3)
    P = torch.randn(3 * 8 * 8, 512, generator=g) * 0.1

    class Tower:
        def encode_image(self, x):                       # [B,3,224,224] -> [B,512]: a fixed linear map of the 8x8 average-pooled image
            return torch.nn.functional.adaptive_avg_pool2d(x, 8).flatten(1) @ P
    feats = {}

    def text(s, norm=True):
        if s not in feats:
            gg = torch.Generator().manual_seed(len(feats) + 10)
            f = torch.randn(7, 512, generator=gg)
            feats[s] = f / f.norm(dim=-1, keepdim=True)
        return feats[s]
    loss = PatchNCELoss(Tower(), text, [96, 54])
    negs = [f'neg {i}' for i in range(8)]
    img = torch.rand(1, 3, 96, 54, generator=g)
    img.requires_grad_(True)
    loss.ZeroPad = torch.nn.ZeroPad2d((54, 54, 96, 96))
    loss.sample_crops = lambda H, W, th, tw, full: [(3 * k, 2 * k) for k in range(12)]
    import nerfart_b200.criteria.losses as LM

    def run(fn):
        if img.grad is not None:
            img.grad = None
        v = fn(); v.backward()
        return float(v), img.grad.clone()
    # (a 96x54 frame cannot hold the reference's 112-pixel crops: shrink them for this arithmetic check)
    orig_forward = LM.PatchNCELoss.forward

    def small_crops(self, source_classes, target_img, target_class, is_full_res, batched=True):
        target_img = self.ZeroPad(target_img)
        target_img = torch.nn.functional.interpolate(target_img, size=tuple(self.target_hw), mode='bicubic', align_corners=False)
        crops = [torch.nn.functional.interpolate(target_img[..., i:i + 40, j:j + 30], size=(224, 224), mode='bicubic', align_corners=False)
                 for i, j in self.sample_crops(96, 54, 40, 30, False)]
        enc = self.get_image_features(torch.cat(crops, 0))
        sfl = [self.get_text_features(s, norm=True) for s in source_classes]
        tf = self.get_text_features(target_class, norm=True)
        if batched:
            return self.crop_losses(enc, sfl, tf, 1)
        total = 0
        for c in range(12):                              # the reference's loop, one crop at a time
            e = enc[c:c + 1]
            near = self.cos(e, tf.detach())
            neg = 0
            for sf in sfl:
                neg = neg + torch.exp(self.cos(e, sf.detach()) / self.temperature)
            pos = torch.exp(near / self.temperature)
            total = total + torch.mean(-torch.log(pos / (pos + neg)))
        return total
    va, ga = run(lambda: small_crops(loss, negs, img, 'painting', False, True))
    vb, gb = run(lambda: small_crops(loss, negs, img, 'painting', False, False))
    assert abs(va - vb) < 1e-5 * abs(vb) and (ga - gb).abs().max() < 1e-5 * gb.abs().max()
    assert orig_forward is LM.PatchNCELoss.forward
