"""CLIP ViT-B/32 image tower on the hand-written kernels of csrc/clip_vit.cu (C ABI na_clip_vitb32_encode_fwd / _bwd).

Replaces `self.model.encode_image(images)` of the reference's three CLIP losses (criteria/clip_loss.py:206-208,
contrastive_loss.py:113-115, patchnce_loss.py:127-129).  The reference holds THREE copies of the model (clip_loss.py:165,
contrastive_loss.py:97, patchnce_loss.py:97); here one `ClipVisionB32` is shared.  Weights are frozen (never trained by
NeRF-Art), so the backward produces d loss / d image only.
"""
import ctypes as C

import torch

from .. import _lib
from .._lib import check, ptr, stream_ptr

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)          # clip_preprocess.transforms[-1] (openai/CLIP clip.py _transform)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


class _Layer(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ('ln_1_w', 'ln_1_b', 'in_proj_w', 'in_proj_b', 'out_proj_w', 'out_proj_b',
                                          'ln_2_w', 'ln_2_b', 'c_fc_w', 'c_fc_b', 'c_proj_w', 'c_proj_b')]


class NaClipWeights(C.Structure):
    _fields_ = [('conv1', C.c_void_p), ('class_embedding', C.c_void_p), ('positional_embedding', C.c_void_p),
                ('ln_pre_w', C.c_void_p), ('ln_pre_b', C.c_void_p), ('layers', _Layer * 12),
                ('ln_post_w', C.c_void_p), ('ln_post_b', C.c_void_p), ('proj', C.c_void_p),
                ('packed', C.c_void_p), ('precision', C.c_int32), ('reserved', C.c_int32)]


NA_CLIP_FP32, NA_CLIP_TF32 = 0, 1
CLIP_PRECISIONS = {'fp32': NA_CLIP_FP32, 'tf32': NA_CLIP_TF32}


def _bind():
    L = _lib.lib()
    if not getattr(L, '_clip_bound', False):
        L.na_clip_workspace_bytes.restype = C.c_size_t
        L.na_clip_workspace_bytes.argtypes = [C.c_int32]
        L.na_clip_vitb32_encode_fwd.argtypes = [C.POINTER(NaClipWeights), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.na_clip_vitb32_encode_bwd.argtypes = [C.POINTER(NaClipWeights), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.na_clip_packed_bytes.restype = C.c_size_t
        L.na_clip_packed_bytes.argtypes = []
        L.na_clip_pack_weights.argtypes = [C.POINTER(NaClipWeights), C.c_void_p, C.c_void_p]
        L._clip_bound = True
    return L


class ClipVisionB32:
    """Weights of openai/CLIP `model.visual` (ViT-B/32) as fp32 CUDA tensors, keyed by the openai state-dict names."""

    KEYS = ['conv1.weight', 'class_embedding', 'positional_embedding', 'ln_pre.weight', 'ln_pre.bias', 'ln_post.weight',
            'ln_post.bias', 'proj'] + [f'transformer.resblocks.{i}.{k}' for i in range(12) for k in (
                'ln_1.weight', 'ln_1.bias', 'attn.in_proj_weight', 'attn.in_proj_bias', 'attn.out_proj.weight', 'attn.out_proj.bias',
                'ln_2.weight', 'ln_2.bias', 'mlp.c_fc.weight', 'mlp.c_fc.bias', 'mlp.c_proj.weight', 'mlp.c_proj.bias')]

    def __init__(self, state_dict, device, precision=None):
        """precision: 'tf32' (default; tcgen05 kind::tf32 linear layers, csrc/tgemm.cu) or 'fp32' (CUDA cores); env NA_CLIP overrides
        the default.  The reference runs this tower in fp16 (clip.load on CUDA), so either is at least its precision."""
        import os
        self.device = torch.device(device)
        self.precision = precision or os.environ.get('NA_CLIP', 'tf32')
        if self.precision not in CLIP_PRECISIONS:
            raise RuntimeError(f'unknown CLIP precision {self.precision!r} (expected one of {sorted(CLIP_PRECISIONS)})')
        if self.device.type != 'cuda':
            raise RuntimeError('nerfart_b200: the CLIP image tower runs on CUDA only (no CPU path exists)')
        self.w = {k: state_dict[k].detach().to(self.device, torch.float32).contiguous() for k in self.KEYS}
        assert self.w['conv1.weight'].shape == (768, 3, 32, 32) and self.w['proj'].shape == (768, 512), 'not a ViT-B/32 image tower'
        W = NaClipWeights()
        g = lambda k: self.w[k].data_ptr()
        W.conv1, W.class_embedding, W.positional_embedding = g('conv1.weight'), g('class_embedding'), g('positional_embedding')
        W.ln_pre_w, W.ln_pre_b, W.ln_post_w, W.ln_post_b, W.proj = g('ln_pre.weight'), g('ln_pre.bias'), g('ln_post.weight'), g('ln_post.bias'), g('proj')
        for i in range(12):
            p = f'transformer.resblocks.{i}.'
            Lw = W.layers[i]
            Lw.ln_1_w, Lw.ln_1_b, Lw.in_proj_w, Lw.in_proj_b = g(p + 'ln_1.weight'), g(p + 'ln_1.bias'), g(p + 'attn.in_proj_weight'), g(p + 'attn.in_proj_bias')
            Lw.out_proj_w, Lw.out_proj_b, Lw.ln_2_w, Lw.ln_2_b = g(p + 'attn.out_proj.weight'), g(p + 'attn.out_proj.bias'), g(p + 'ln_2.weight'), g(p + 'ln_2.bias')
            Lw.c_fc_w, Lw.c_fc_b, Lw.c_proj_w, Lw.c_proj_b = g(p + 'mlp.c_fc.weight'), g(p + 'mlp.c_fc.bias'), g(p + 'mlp.c_proj.weight'), g(p + 'mlp.c_proj.bias')
        W.packed, W.precision, W.reserved = None, CLIP_PRECISIONS[self.precision], 0
        self.cw = W
        self.packed = None
        if self.precision == 'tf32':
            # TF32 operand images of the (frozen) weights, both orientations: built once
            L = _bind()
            self.packed = torch.empty(int(L.na_clip_packed_bytes()), dtype=torch.uint8, device=self.device)
            with torch.cuda.device(self.device):
                check(L.na_clip_pack_weights(C.byref(W), ptr(self.packed), stream_ptr(self.device)), 'na_clip_pack_weights')
            W.packed = self.packed.data_ptr()

    # -- constructors ---------------------------------------------------------------------------------------------
    @classmethod
    def from_openai(cls, clip_model, device, precision=None):
        """`clip_model` = the object `clip.load("ViT-B/32")` returns (clip_loss.py:165)."""
        sd = {k[len('visual.'):]: v for k, v in clip_model.state_dict().items() if k.startswith('visual.')}
        return cls(sd, device, precision)

    @classmethod
    def from_hf(cls, hf_model, device, precision=None):
        """transformers.CLIPVisionModelWithProjection (the offline stand-in oracle) -> openai layout."""
        s = hf_model.state_dict()
        v = 'vision_model.'
        sd = {'conv1.weight': s[v + 'embeddings.patch_embedding.weight'], 'class_embedding': s[v + 'embeddings.class_embedding'],
              'positional_embedding': s[v + 'embeddings.position_embedding.weight'],
              'ln_pre.weight': s[v + 'pre_layrnorm.weight'], 'ln_pre.bias': s[v + 'pre_layrnorm.bias'],
              'ln_post.weight': s[v + 'post_layernorm.weight'], 'ln_post.bias': s[v + 'post_layernorm.bias'],
              'proj': s['visual_projection.weight'].t().contiguous()}
        for i in range(12):
            a, b = f'{v}encoder.layers.{i}.', f'transformer.resblocks.{i}.'
            sd[b + 'ln_1.weight'], sd[b + 'ln_1.bias'] = s[a + 'layer_norm1.weight'], s[a + 'layer_norm1.bias']
            sd[b + 'ln_2.weight'], sd[b + 'ln_2.bias'] = s[a + 'layer_norm2.weight'], s[a + 'layer_norm2.bias']
            sd[b + 'attn.in_proj_weight'] = torch.cat([s[a + f'self_attn.{n}_proj.weight'] for n in 'qkv'], 0)
            sd[b + 'attn.in_proj_bias'] = torch.cat([s[a + f'self_attn.{n}_proj.bias'] for n in 'qkv'], 0)
            sd[b + 'attn.out_proj.weight'], sd[b + 'attn.out_proj.bias'] = s[a + 'self_attn.out_proj.weight'], s[a + 'self_attn.out_proj.bias']
            sd[b + 'mlp.c_fc.weight'], sd[b + 'mlp.c_fc.bias'] = s[a + 'mlp.fc1.weight'], s[a + 'mlp.fc1.bias']
            sd[b + 'mlp.c_proj.weight'], sd[b + 'mlp.c_proj.bias'] = s[a + 'mlp.fc2.weight'], s[a + 'mlp.fc2.bias']
        return cls(sd, device, precision)

    @classmethod
    def random(cls, seed, device, precision=None):
        """Seeded random weights of the right shapes (benchmarks / smoke tests when no CLIP weights are available offline)."""
        g = torch.Generator(device='cpu'); g.manual_seed(seed)
        shapes = {'conv1.weight': (768, 3, 32, 32), 'class_embedding': (768,), 'positional_embedding': (50, 768), 'proj': (768, 512)}
        per = {'attn.in_proj_weight': (2304, 768), 'attn.in_proj_bias': (2304,), 'attn.out_proj.weight': (768, 768),
               'attn.out_proj.bias': (768,), 'mlp.c_fc.weight': (3072, 768), 'mlp.c_fc.bias': (3072,), 'mlp.c_proj.weight': (768, 3072),
               'mlp.c_proj.bias': (768,)}
        sd = {}
        for k in cls.KEYS:
            if k.endswith('.weight') and ('ln_' in k):
                sd[k] = torch.ones(768)
            elif k.endswith('.bias') and ('ln_' in k):
                sd[k] = torch.zeros(768)
            else:
                shp = shapes.get(k) or per[k.split('.', 3)[3]]
                fan_in = shp[-1] if len(shp) == 2 else (3072 if len(shp) == 4 else 768)
                sd[k] = torch.randn(*shp, generator=g) * (fan_in ** -0.5 if len(shp) > 1 else 0.02)
        return cls(sd, device, precision)

    # -- the op ----------------------------------------------------------------------------------------------------
    def encode_image(self, images):
        """[B,3,224,224] fp32 (resized, CLIP-normalised) -> [B,512]; differentiable w.r.t. `images`."""
        return _Encode.apply(images, self)


class _Encode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, images, tower):
        L = _bind()
        if images.dim() != 4 or tuple(images.shape[1:]) != (3, 224, 224):
            raise RuntimeError(f'CLIP ViT-B/32 expects [B,3,224,224], got {tuple(images.shape)}')
        x = images.detach().to(tower.device, torch.float32).contiguous()
        B = x.shape[0]
        nbytes = L.na_clip_workspace_bytes(B)
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=tower.device)
        feats = torch.empty(B, 512, dtype=torch.float32, device=tower.device)
        with torch.cuda.device(tower.device):
            check(L.na_clip_vitb32_encode_fwd(C.byref(tower.cw), ptr(x), B, ptr(feats), ptr(ws), ws.numel(), stream_ptr(tower.device)),
                  'na_clip_vitb32_encode_fwd')
        ctx.tower, ctx.B = tower, B
        ctx.ws = ws if ctx.needs_input_grad[0] else None          # the activations live until this encode's backward
        ctx.in_dtype = images.dtype
        return feats

    @staticmethod
    def backward(ctx, grad_feats):
        L = _bind()
        tower = ctx.tower
        g = grad_feats.detach().to(torch.float32).contiguous()
        gi = torch.empty(ctx.B, 3, 224, 224, dtype=torch.float32, device=tower.device)
        with torch.cuda.device(tower.device):
            check(L.na_clip_vitb32_encode_bwd(C.byref(tower.cw), ptr(g), ctx.B, ptr(gi), ptr(ctx.ws), ctx.ws.numel(), stream_ptr(tower.device)),
                  'na_clip_vitb32_encode_bwd')
        ctx.ws = None
        return gi.to(ctx.in_dtype), None
