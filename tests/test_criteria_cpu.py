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
