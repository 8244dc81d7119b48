"""Parameter containers mirroring the reference's `models/base.py` public surface.

Same class names, constructor arguments, attribute names and -- the parity constraint -- the same `state_dict()`
key layout as the reference (`bias`, `weight_g`, `weight_v` per layer: the old-style `nn.utils.weight_norm`
names, /root/reference/models/base.py:226-227,365-366) and the same seeded initialisation
(geometric sphere init, base.py:207-224), so reference checkpoints load unchanged and a seeded `get_model` gives
bit-identical weights (pinned by tests/golden/state_digest.json).

The arithmetic is NOT here: `forward*` dispatch to the sm_100a kernels through the C ABI (`_lib`).  There is no
PyTorch fallback; calling them on CPU tensors raises.
"""
import math
import weakref
import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..engine import NetEngine


class Embedder(nn.Module):
    """Positional encoding [x, sin(2^0 x), cos(2^0 x), ...] (reference models/base.py:14-64).  Inside the renderers the
    encoding is fused into the MLP kernel; this module only exists so `get_embedder` keeps its contract
    (`embed_fn, out_dim`) for callers that size layers from it."""

    def __init__(self, input_dim, N_freqs):
        super().__init__()
        self.input_dim, self.N_freqs = input_dim, N_freqs
        self.out_dim = input_dim * (1 + 2 * N_freqs)
        self.freq_bands = [2.0 ** k for k in range(N_freqs)]

    def forward(self, x):
        parts = [x]
        for f in self.freq_bands:
            parts += [torch.sin(x * f), torch.cos(x * f)]
        return torch.cat(parts, dim=-1)


def get_embedder(multires, input_dim=3):
    """reference models/base.py:67-81: multires < 0 -> identity."""
    if multires < 0:
        return nn.Identity(), input_dim
    e = Embedder(input_dim, multires)
    return e, e.out_dim


class WNLinear(nn.Module):
    """A weight-normalised linear layer *as stored in a reference checkpoint*: parameters `bias`, `weight_g` [out,1],
    `weight_v` [out,in], registered in that order.  W_eff = g * v / ||v||_row is folded by the pack kernel
    (csrc/api.cu) every time the engine repacks; nothing is computed in Python."""

    def __init__(self, in_dim, out_dim, init_fn=None):
        super().__init__()
        lin = nn.Linear(in_dim, out_dim)            # consumes the RNG exactly like the reference's nn.Linear
        if init_fn is not None:
            init_fn(lin)
        w = lin.weight.detach()
        self.in_features, self.out_features = in_dim, out_dim
        self.bias = nn.Parameter(lin.bias.detach().clone())
        self.weight_g = nn.Parameter(torch.norm_except_dim(w, 2, 0).clone())
        self.weight_v = nn.Parameter(w.clone())

    def extra_repr(self):
        return f'in={self.in_features}, out={self.out_features}, weight_norm'


class ImplicitSurface(nn.Module):
    """SDF network (reference models/base.py:131-282).  Supported geometry = what every shipped config builds:
    D=8, W=256, skips=[4], embed_multires=6, W_geo_feat=256, weight_norm, no SIREN."""

    def __init__(self, W=256, D=8, skips=[4], W_geo_feat=256, input_ch=3, radius_init=1.0, obj_bounding_size=2.0,
                 geometric_init=True, embed_multires=6, weight_norm=True, use_siren=False):
        super().__init__()
        if use_siren or not weight_norm or W != 256 or D != 8 or list(skips) != [4] or embed_multires != 6 or W_geo_feat != 256:
            raise NotImplementedError('nerfart_b200 kernels implement the shipped geometry only '
                                      '(D=8, W=256, skips=[4], embed_multires=6, W_geo_feat=256, weight_norm, no SIREN)')
        self.radius_init = radius_init
        self.register_buffer('obj_bounding_size', torch.tensor([obj_bounding_size]).float())
        self.geometric_init = geometric_init
        self.D, self.W, self.W_geo_feat = D, W, W_geo_feat
        self.skips, self.use_siren = skips, use_siren
        self.embed_fn, emb_ch = get_embedder(embed_multires)
        layers = []
        for l in range(D + 1):
            out_dim = (1 + W_geo_feat) if l == D else (W - emb_ch if (l + 1) in skips else W)
            in_dim = emb_ch if l == 0 else W
            layers.append(WNLinear(in_dim, out_dim, self._geometric_init_fn(l, D, in_dim, out_dim, emb_ch)
                                   if geometric_init else None))
        self.surface_fc_layers = nn.ModuleList(layers)
        self._owner = None                      # set by VolSDF / NeuS so the engine sees the radiance net too
        self._engine = None

    def _geometric_init_fn(self, l, D, in_dim, out_dim, emb_ch):
        # sphere init of SAL / IDR, the reference's base.py:207-224 (call order matters for RNG parity)
        skips, r0 = self.skips, self.radius_init

        def fn(lin):
            with torch.no_grad():
                std = math.sqrt(2) / math.sqrt(out_dim)
                if l == D:
                    nn.init.normal_(lin.weight, mean=math.sqrt(math.pi) / math.sqrt(in_dim), std=0.0001)
                    nn.init.constant_(lin.bias, -r0)
                elif l == 0:
                    nn.init.constant_(lin.bias, 0.0)
                    nn.init.constant_(lin.weight[:, 3:], 0.0)
                    nn.init.normal_(lin.weight[:, :3], 0.0, std)
                elif l in skips:
                    nn.init.constant_(lin.bias, 0.0)
                    nn.init.normal_(lin.weight, 0.0, std)
                    nn.init.constant_(lin.weight[:, -(emb_ch - 3):], 0.0)
                else:
                    nn.init.constant_(lin.bias, 0.0)
                    nn.init.normal_(lin.weight, 0.0, std)
        return fn

    # -- engine plumbing -------------------------------------------------------------------------
    def engine(self):
        owner = self._owner() if self._owner is not None else None
        if owner is not None:
            return owner.engine()
        if self._engine is None:
            self._engine = NetEngine(self, None, framework='neus', multires_view=-1, bounding_radius=float(self.obj_bounding_size[0]))
        return self._engine

    def pretrain_hook(self, configs={}):
        configs['target_radius'] = self.radius_init
        configs['obj_bounding_size'] = self.obj_bounding_size.item()
        return False                                      # only the (unsupported) SIREN variant pretrains, base.py:233-241

    def forward(self, x: torch.Tensor, return_h=False, **kwargs):
        """sdf (and the 256-d geometry feature) at x[..., 3]; reference base.py:243-263."""
        _no_autograd(self, x)
        sdf, h = self.engine().sdf_eval(x, apply_bg=False, want_feat=return_h)
        return (sdf, h) if return_h else sdf

    def forward_with_nablas(self, x: torch.Tensor, has_grad_bypass: bool = None, **kwargs):
        """(sdf, d sdf/dx, feature); reference base.py:265-282 (there via autograd, here the fused reverse sweep)."""
        _no_autograd(self, x, has_grad_bypass)
        _, sdf, nab, feat = self.engine().full_eval(x, None, want_radiance=False, apply_bg=False)
        return sdf, nab, feat


def _no_autograd(module, x, bypass=None):
    want = torch.is_grad_enabled() if bypass is None else bypass
    if want and (x.requires_grad or any(p.requires_grad for p in module.parameters())):
        if torch.is_grad_enabled() and bypass is None and not x.requires_grad:
            # parameters require grad by default; a forward under grad mode without a backward consumer is common
            # (e.g. validation without no_grad).  The kernels are forward-only: results carry no graph.
            return
        raise NotImplementedError('nerfart_b200: backward kernels are not built yet (SURVEY.md 8a row a17 is a "next" row); '
                                  'run under torch.no_grad()')


class RadianceNet(nn.Module):
    """Radiance network (reference models/base.py:312-391): cat[x, embed(view), normals, feature] -> 4x(256, ReLU) -> (3, Sigmoid)."""

    def __init__(self, D=4, W=256, skips=[], W_geo_feat=256, embed_multires=6, embed_multires_view=4,
                 use_view_dirs=True, weight_norm=True, use_siren=False):
        super().__init__()
        if use_siren or not weight_norm or D != 4 or W != 256 or list(skips) or embed_multires >= 0 or not use_view_dirs \
                or embed_multires_view not in (-1, 4) or W_geo_feat != 256:
            raise NotImplementedError('nerfart_b200 kernels implement the shipped radiance geometry only '
                                      '(D=4, W=256, no skips, embed_multires=-1, embed_multires_view in {-1,4}, view dirs)')
        self.skips, self.D, self.W, self.use_view_dirs = skips, D, W, use_view_dirs
        self.embed_fn, ch_pts = get_embedder(embed_multires)
        self.embed_fn_view, ch_view = get_embedder(embed_multires_view)
        self.embed_multires_view = embed_multires_view
        in0 = ch_pts + ch_view + 3 + W_geo_feat
        self.layers = nn.ModuleList([WNLinear(in0 if l == 0 else W, 3 if l == D else W) for l in range(D + 1)])
        self._owner = None

    def forward(self, x, view_dirs=None, normals=None, geometry_feature=None):
        raise NotImplementedError('the radiance net only runs fused behind VolSDF.forward / NeuS.forward / the renderers')


# ---------------------------------------------------------------------------------------------------
# optimiser / scheduler helpers the reference's train.py imports from models.base (base.py:486-584)
# ---------------------------------------------------------------------------------------------------
def get_optimizer(args, model):
    """reference base.py:486-521.  args.training.lr is a number (one Adam group) or a dict: 'default' plus keys naming a direct
    parameter of the model (e.g. ln_beta / ln_s) or a child module; the default group comes FIRST in optimizer.param_groups (the
    order reference optimizer checkpoints are saved with); an unknown key raises RuntimeError('wrong lr key:', name)."""
    import numbers
    from torch import optim
    lr = args.training.lr
    if isinstance(lr, numbers.Number):
        return optim.Adam(model.parameters(), lr=lr)
    if not isinstance(lr, dict):
        raise NotImplementedError
    default_lr = lr.pop('default')
    groups, taken = [], set()
    for key, value in lr.items():
        if key in model._parameters:
            taken.add(key)
            groups.append({'params': getattr(model, key), 'lr': value})
        elif key in model._modules:
            child = getattr(model, key)
            taken.update(f'{key}.{n}' for n, _ in child.named_parameters())
            groups.append({'params': child.parameters(), 'lr': value})
        else:
            raise RuntimeError('wrong lr key:', key)
    rest = [p for n, p in model.named_parameters() if n not in taken]
    return optim.Adam(params=[{'params': rest, 'lr': default_lr}] + groups, lr=default_lr)


def _warmup_cosine(total_steps, warmup_steps, min_factor):
    """reference CosineAnnealWarmUpSchedulerLambda, base.py:524-535."""
    assert 0 <= min_factor < 1

    def factor(it):
        if it < warmup_steps:
            return it / warmup_steps
        return min_factor + (1 - min_factor) * 0.5 * (np.cos(np.pi * (it - warmup_steps) / (total_steps - warmup_steps)) + 1.0)
    return factor


def _exponential_step(total_steps, min_factor):
    """reference ExponentialSchedulerLambda, base.py:538-544: min_factor ** clip(it / total, 0, 1)."""
    assert 0 <= min_factor < 1
    return lambda it: np.exp(np.clip(it / total_steps, 0, 1) * np.log(min_factor))


def get_scheduler(args, optimizer, last_epoch=-1):
    """reference base.py:547-584: 'multistep' | 'warmupcosine' | 'exponential_step' from args.training.scheduler; min_factor defaults
    to 0.1 through setdefault (so it also appears in the saved config); like the reference, 'exponential_step' ignores last_epoch."""
    from torch.optim import lr_scheduler
    sc = args.training.scheduler
    if sc.type == 'multistep':
        return lr_scheduler.MultiStepLR(optimizer, sc.milestones, gamma=sc.gamma, last_epoch=last_epoch)
    if sc.type == 'warmupcosine':
        return lr_scheduler.LambdaLR(optimizer, _warmup_cosine(args.training.num_iters, sc.warmup_steps, sc.setdefault('min_factor', 0.1)),
                                     last_epoch=last_epoch)
    if sc.type == 'exponential_step':
        return lr_scheduler.LambdaLR(optimizer, _exponential_step(args.training.num_iters, sc.setdefault('min_factor', 0.1)))
    raise NotImplementedError
