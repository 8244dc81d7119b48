"""Style losses of the CLIP fine-tune step (reference criteria/): one shared CLIP ViT-B/32 image tower on hand-written
kernels (clip_vit.py), one text-feature cache (text.py), the three CLIP losses (losses.py).

`build_loss_dict(target_hw, device)` returns what the reference's Trainer builds at volsdf.py:639-645.  It needs the
third-party `clip` package and its ViT-B/32 weights (README.md:27; fetched by `clip.load` on first use) for the TEXT tower
and as the source of the image-tower weights; offline neither exists, and this function raises instead of substituting
anything.  The VGG16 perceptual term (criteria/perp_loss.py) is `perceptual.VGGPerceptualLoss` (cuDNN convolutions, SURVEY.md 8f
rank 3); it raises when the ImageNet weights cannot be found -- the term is never dropped silently.
"""
from .clip_vit import ClipVisionB32
from .perceptual import VGGPerceptualLoss                                         # noqa: F401
from .text import TextFeatures
from .losses import CLIPLoss, ContrastiveLoss, PatchNCELoss, DirectionLoss       # noqa: F401


def make_loss_dict(tower, text, target_hw, perceptual=None):
    return {'contrastive': ContrastiveLoss(tower, text), 'patchnce': PatchNCELoss(tower, text, target_hw),
            'clip': CLIPLoss(tower, text), 'perceptual': perceptual}


def build_loss_dict(target_hw, device):
    try:
        import clip                                            # openai/CLIP
    except Exception as e:
        raise RuntimeError('nerfart_b200.criteria: the CLIP losses need the `clip` package (pip install '
                           'git+https://github.com/openai/CLIP.git, reference README.md:27) for the text tower and the '
                           'ViT-B/32 weights; pass `loss_dict=` to Trainer to supply your own losses') from e
    model, _ = clip.load("ViT-B/32", device=device)
    model = model.float().eval()
    tower = ClipVisionB32.from_openai(model, device)
    text = TextFeatures(lambda strings: model.encode_text(clip.tokenize(strings).to(device)))
    perceptual = VGGPerceptualLoss().to(device)                 # raises when the VGG16 weights are unavailable (see perceptual.py)
    return make_loss_dict(tower, text, target_hw, perceptual)
