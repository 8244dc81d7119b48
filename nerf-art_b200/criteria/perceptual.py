"""VGG16 perceptual loss of the fine-tune step (reference criteria/perp_loss.py:9-57; weighted by finetune.w_perceptual in
calc_style_loss, volsdf.py:898-900 / neus.py:648-649): L1 distance between the relu3_3 feature maps of the rendered image and the
ground-truth image, both ImageNet-normalised and bilinearly resized to 224 x 224.

Library call, stated: the seven 3x3 convolutions run through `torch.nn.functional.conv2d` (cuDNN) -- SURVEY.md 8f rank 3 keeps this
loss out of the hand-written-kernel scope (10 GFLOP per step against 1e14 for the renders).  Two differences from the reference,
neither changes a result: the reference instantiates torchvision's vgg16 FOUR times and also evaluates conv4 (features[16:23]),
whose output the loss never reads (perp_loss.py:13-18,49-55); here one set of conv1_1..conv3_3 weights is held and conv4 is skipped.

The term is NEVER dropped silently.  Weights come from, in order: the `weights=` argument / $NA_VGG16_WEIGHTS (a torchvision vgg16
state-dict file), torchvision's hub cache (`vgg16-397923af.pth`), a torchvision download; if none is available construction
raises.  `random:<seed>` (argument or environment) builds seeded random weights -- for tests and benchmarks in sandboxes without the
checkpoint, where only the arithmetic and its cost matter.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ['VGGPerceptualLoss']

# torchvision vgg16().features indices of conv1_1, conv1_2, conv2_1, conv2_2, conv3_1, conv3_2, conv3_3 and their (out, in) channels;
# a max-pool (features[4], features[9]) precedes conv2_1 and conv3_1
CONVS = ((0, 64, 3), (2, 64, 64), (5, 128, 64), (7, 128, 128), (10, 256, 128), (12, 256, 256), (14, 256, 256))
POOL_BEFORE = {5, 10}
HUB_FILE = 'vgg16-397923af.pth'


def _load_state(spec):
    """torchvision vgg16 state dict restricted to the seven convolutions of features[:16]."""
    if spec is None:
        spec = os.environ.get('NA_VGG16_WEIGHTS')
    if isinstance(spec, dict):
        sd = spec
    elif isinstance(spec, str) and spec.startswith('random'):
        seed = int(spec.split(':', 1)[1]) if ':' in spec else 0
        g = torch.Generator(device='cpu'); g.manual_seed(seed)
        sd = {}
        for idx, co, ci in CONVS:
            sd[f'features.{idx}.weight'] = torch.randn(co, ci, 3, 3, generator=g) * (2.0 / (9 * ci)) ** 0.5
            sd[f'features.{idx}.bias'] = torch.randn(co, generator=g) * 0.05
    else:
        path = spec
        if path is None:
            cand = os.path.join(torch.hub.get_dir(), 'checkpoints', HUB_FILE)
            path = cand if os.path.exists(cand) else None
        if path is not None:
            sd = torch.load(path, map_location='cpu')
        else:
            try:
                from torchvision.models import vgg16, VGG16_Weights
                sd = vgg16(weights=VGG16_Weights.IMAGENET1K_V1).state_dict()          # what vgg16(pretrained=True) loads
            except Exception as e:
                raise RuntimeError(
                    'nerfart_b200.criteria.VGGPerceptualLoss: the ImageNet VGG16 weights are not available (no $NA_VGG16_WEIGHTS, no '
                    f'{HUB_FILE} in the torch hub cache, download failed).  The reference adds w_perceptual * this loss to the style '
                    'objective (volsdf.py:898-900); it is not dropped silently: provide the checkpoint, or set finetune.w_perceptual '
                    'to 0 to train without the term.') from e
    return {k: sd[k].detach().float() for idx, _, _ in CONVS for k in (f'features.{idx}.weight', f'features.{idx}.bias')}


class VGGPerceptualLoss(nn.Module):
    def __init__(self, resize=True, weights=None):
        super().__init__()
        sd = _load_state(weights)
        for idx, _, _ in CONVS:
            self.register_buffer(f'w{idx}', sd[f'features.{idx}.weight'].contiguous())
            self.register_buffer(f'b{idx}', sd[f'features.{idx}.bias'].contiguous())
        self.register_buffer('mean', torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer('std', torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))
        self.resize = resize

    def features(self, x):
        """relu3_3 of the normalised (and resized) image batch."""
        x = (x - self.mean) / self.std
        if self.resize:
            x = F.interpolate(x, mode='bilinear', size=(224, 224), align_corners=False)
        for idx, _, _ in CONVS:
            if idx in POOL_BEFORE:
                x = F.max_pool2d(x, kernel_size=2, stride=2)
            x = F.relu(F.conv2d(x, getattr(self, f'w{idx}'), getattr(self, f'b{idx}'), padding=1))
        return x

    def forward(self, input, target, feature_layers=None):
        """input / target [B,3,H,W] in [0,1] (grey images are repeated to 3 channels, perp_loss.py:27-29) -> scalar."""
        if input.shape[1] != 3:
            input, target = input.repeat(1, 3, 1, 1), target.repeat(1, 3, 1, 1)
        both = self.features(torch.cat([input, target], 0))          # one pass for the two images
        n = input.shape[0]
        return F.l1_loss(both[:n], both[n:])
