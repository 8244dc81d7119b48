"""Synthetic openai/CLIP for tests (see ../README.md): `load`, `tokenize`, a model with `.visual` (ViT-B/32 image tower in the
upstream module / state-dict layout, plain PyTorch), `encode_image`, `encode_text`.  Written from the public model definition
(SURVEY.md App. C): conv1 32x32/32 without bias -> [cls | 49 patches] + positional embedding -> ln_pre -> 12 x
(x += attn(ln_1 x); x += mlp(ln_2 x)) with nn.MultiheadAttention(768, 12) and c_fc / QuickGELU / c_proj -> ln_post(cls) @ proj.
Weights are seeded random numbers: nothing here approximates the trained model; it pins layout and arithmetic only."""
from collections import OrderedDict
import hashlib

import torch
import torch.nn as nn

_CONTEXT = 77


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([('c_fc', nn.Linear(d_model, d_model * 4)), ('gelu', QuickGELU()),
                                              ('c_proj', nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = nn.LayerNorm(d_model)

    def forward(self, x):                                   # x: [tokens, batch, width]
        h = self.ln_1(x)
        x = x + self.attn(h, h, h, need_weights=False)[0]
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads) for _ in range(layers)])

    def forward(self, x):
        return self.resblocks(x)


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution=224, patch_size=32, width=768, layers=12, heads=12, output_dim=512):
        super().__init__()
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

    def forward(self, x):
        x = self.conv1(x)                                   # [B, width, 7, 7]
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
        cls = self.class_embedding.to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype, device=x.device)
        x = torch.cat([cls, x], dim=1) + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        x = self.transformer(x.permute(1, 0, 2)).permute(1, 0, 2)
        return self.ln_post(x[:, 0, :]) @ self.proj


class FakeCLIP(nn.Module):
    def __init__(self, seed=0):
        super().__init__()
        g = torch.random.get_rng_state()
        torch.manual_seed(seed)
        self.visual = VisionTransformer()
        # non-trivial LayerNorm affine and biases, so that a swapped weight/bias or a wrong orientation cannot pass unnoticed
        with torch.no_grad():
            for n, p in self.visual.named_parameters():
                if n.endswith('bias'):
                    p.normal_(0.0, 0.02)
                elif 'ln_' in n and n.endswith('weight'):
                    p.normal_(1.0, 0.05)
        self.text_table = nn.Parameter(torch.randn(4096, 512) * 0.05, requires_grad=False)
        torch.random.set_rng_state(g)

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image):
        return self.visual(image.type(self.dtype))

    def encode_text(self, tokens):
        """Deterministic stand-in for the text tower: mean of per-(token, position) table rows -> [n, 512]."""
        idx = (tokens.long() * 31 + torch.arange(tokens.shape[1], device=tokens.device)[None] * 7) % self.text_table.shape[0]
        w = (tokens != 0).to(self.text_table.dtype)[..., None]
        return (self.text_table[idx] * w).sum(1) / w.sum(1).clamp_min(1)


def tokenize(texts, context_length=_CONTEXT, truncate=False):
    if isinstance(texts, str):
        texts = [texts]
    out = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, t in enumerate(texts):
        words = t.lower().split()[:context_length - 2]
        ids = [49406] + [int(hashlib.md5(w.encode()).hexdigest()[:6], 16) % 49000 + 256 for w in words] + [49407]
        out[i, :len(ids)] = torch.tensor(ids)
    return out


def _transform(n_px=224):
    from torchvision import transforms as T
    return T.Compose([T.Resize(n_px, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(n_px), lambda im: im.convert('RGB'),
                      T.ToTensor(), T.Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))])


def available_models():
    return ['ViT-B/32']


def load(name='ViT-B/32', device='cpu', jit=False, download_root=None):
    assert name == 'ViT-B/32', name
    model = FakeCLIP(0).to(device).eval()
    if str(device).startswith('cuda'):
        model = model.half()                                # upstream clip.load keeps fp16 weights on CUDA
    return model, _transform(224)
