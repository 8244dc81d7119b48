"""NeuS behind the reference's API (reference models/frameworks/neus.py).  See volsdf.py in this package for the
conventions; the renderer call is `NetEngine.neus_render` (csrc/neus_render.cu)."""
import copy
import weakref
from collections import OrderedDict
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from ..base import ImplicitSurface, RadianceNet
from ...engine import NetEngine

FIX_MODULE = "radiance_net"          # neus.py:28


class NeuS(nn.Module):
    def __init__(self, variance_init=0.05, speed_factor=1.0, input_ch=3, W_geo_feat=-1, use_outside_nerf=False,
                 obj_bounding_radius=1.0, surface_cfg=dict(), radiance_cfg=dict()):
        super().__init__()
        if use_outside_nerf:
            raise NotImplementedError('the outside NeRF++ is unreachable with the shipped configs (with_mask: True, neus.py:702)')
        self.ln_s = nn.Parameter(data=torch.Tensor([-np.log(variance_init) / speed_factor]), requires_grad=True)
        self.speed_factor = speed_factor
        self.obj_bounding_radius = obj_bounding_radius
        self.implicit_surface = ImplicitSurface(W_geo_feat=W_geo_feat, input_ch=input_ch,
                                                obj_bounding_size=obj_bounding_radius, **surface_cfg)
        if W_geo_feat < 0:
            W_geo_feat = self.implicit_surface.W
        self.radiance_net = RadianceNet(W_geo_feat=W_geo_feat, **radiance_cfg)
        self.implicit_surface._owner = weakref.ref(self)
        self._engine = None

    def engine(self) -> NetEngine:
        if self._engine is None:
            self._engine = NetEngine(self.implicit_surface, self.radiance_net, 'neus',
                                     self.radiance_net.embed_multires_view, self.obj_bounding_radius)
        return self._engine

    def forward_radiance(self, x: torch.Tensor, view_dirs: torch.Tensor, return_nablas=False):
        return self.engine().full_eval(x, view_dirs, want_radiance=True, apply_bg=False)[0]

    def forward_s(self):
        return torch.exp(self.ln_s * self.speed_factor)

    def forward(self, x: torch.Tensor, view_dirs: torch.Tensor, return_nablas=False):
        rad, sdf, nab, _ = self.engine().full_eval(x, view_dirs, want_radiance=True, apply_bg=False)
        return rad, sdf, nab

    def fix_module(self, module_name):
        if module_name is None or module_name == "":
            return
        if module_name not in ('implicit_surface', 'radiance_net'):
            raise NotImplementedError(f"{module_name} is not a valid module.")
        for p in getattr(self, module_name).parameters():
            p.requires_grad = False


def volume_render(rays_o, rays_d, model: NeuS, obj_bounding_radius=1.0, batched=False, batched_info={}, calc_normal=False,
                  use_view_dirs=True, rayschunk=65536, netchunk=1048576, white_bkgd=False,
                  near_bypass: Optional[float] = None, far_bypass: Optional[float] = None, detailed_output=True,
                  show_progress=False, perturb=False, fixed_s_recp=1 / 64., N_samples=64, N_importance=64, N_outside=0,
                  upsample_algo='official_solution', N_nograd_samples=2048, N_upsample_iters=4, u_rand=None, **dummy_kwargs):
    """Same contract as the reference's volume_render (neus.py:142-424) for upsample_algo='official_solution', N_outside=0."""
    if upsample_algo != 'official_solution' or N_outside > 0 or near_bypass is not None or far_bypass is not None or not use_view_dirs:
        raise NotImplementedError("only upsample_algo='official_solution', N_outside=0, no near/far bypass (the shipped configs)")
    B = rays_d.shape[0] if batched else None
    ro = rays_o.reshape(-1, 3).float().contiguous()
    rd = rays_d.reshape(-1, 3).float().contiguous()
    s = model.forward_s().detach().reshape(1).float().contiguous()
    o = model.engine().neus_render(ro, rd, s, obj_bounding_radius=obj_bounding_radius, N_samples=N_samples,
                                   N_importance=N_importance, N_upsample_iters=N_upsample_iters, white_bkgd=white_bkgd,
                                   perturb=perturb, detailed_output=detailed_output, u_rand=u_rand)

    def shp(t, *tail):
        return t.reshape(*((B, -1) if batched else (-1,)), *tail)

    P = N_samples + N_importance
    ret = OrderedDict([('rgb', shp(o['rgb'], 3)), ('depth_volume', shp(o['depth'])), ('mask_volume', shp(o['acc']))])
    if calc_normal:
        ret['normals_volume'] = shp(o['normals'], 3)
    if detailed_output:
        d_all = shp(o['d_all'], P)
        sdf = shp(o['sdf'], P)
        ret['implicit_nablas'] = shp(o['nablas'], P, 3)
        ret['implicit_surface'] = sdf
        ret['radiance'] = shp(o['radiance'], P - 1, 3)
        ret['alpha'] = shp(o['alpha'], P - 1)
        ret['cdf'] = torch.sigmoid(sdf * s)                                     # logging-only extra (neus.py:29-33)
        ret['visibility_weights'] = shp(o['weights'], P - 1)
        ret['d_final'] = 0.5 * (d_all[..., 1:] + d_all[..., :-1])
        ret['d_all'] = d_all
    return ret['rgb'], ret['depth_volume'], ret


class SingleRenderer(nn.Module):
    def __init__(self, model: NeuS):
        super().__init__()
        self.model = model

    def forward(self, rays_o, rays_d, **kwargs):
        return volume_render(rays_o, rays_d, self.model, **kwargs)


def render_patch(model, ro, rd, obj_bounding_radius=1.0, perturb=False, white_bkgd=False, N_samples=64, N_importance=64,
                 N_upsample_iters=4, u_rand=None, train_stash=False, **dummy_kwargs):
    """Forward render of one flat ray patch with the detailed outputs the backward needs (defaults of neus.py:142-170).
    Returns (flat outputs, {s} on the device).  train_stash: see volsdf.render_patch (split training program)."""
    s = model.forward_s().detach().reshape(1).float().contiguous()
    o = model.engine().neus_render(ro, rd, s, obj_bounding_radius=obj_bounding_radius, N_samples=N_samples,
                                   N_importance=N_importance, N_upsample_iters=N_upsample_iters, white_bkgd=white_bkgd,
                                   perturb=perturb, detailed_output=True, u_rand=u_rand, train_stash=train_stash)
    return o, s


class Trainer(nn.Module):
    """The reference's Trainer (neus.py:427-628): `forward` implements the CLIP fine-tune branch (520-576) on the CUDA
    backward kernels (radiance_net frozen, neus.py:28,455-456); the reconstruction branch raises.  See volsdf.Trainer."""

    def __init__(self, model: NeuS, device_ids=[0], batched=True, is_finetune=False, target_hw: list = None, loss_dict=None):
        super().__init__()
        self.model = model
        self.renderer = SingleRenderer(model)
        self.device = device_ids[0] if isinstance(device_ids, (list, tuple)) and len(device_ids) else 0
        self.is_finetune = is_finetune
        self.neg_texts = None
        self.target_hw = target_hw if target_hw is not None else [960, 540]
        self.loss_dict = loss_dict
        if is_finetune:
            self.model.fix_module(FIX_MODULE)

    def _losses(self):
        if self.loss_dict is None:
            from ...criteria import build_loss_dict
            self.loss_dict = build_loss_dict(self.target_hw, next(self.model.parameters()).device)
        return self.loss_dict

    def forward(self, args, indices, model_input, ground_truth, render_kwargs_train: dict, it: int, device='cuda', optimizer=None):
        if not args.training.is_finetune:
            raise NotImplementedError('nerfart_b200: only the CLIP fine-tune branch of Trainer.forward (neus.py:520-576) is built')
        from ._finetune import finetune_forward
        self._losses()
        losses, select_inds = finetune_forward(self, 'neus', args, model_input, ground_truth, render_kwargs_train, optimizer,
                                               lambda ro, rd, **kw: render_patch(self.model, ro, rd, **kw))
        extras = {}
        extras['scalars'] = {'1/s': 1. / self.model.forward_s().data}
        extras['select_inds'] = select_inds
        return OrderedDict([('losses', losses), ('extras', extras)])


def get_model(args, render_target=None):
    """reference neus.get_model, neus.py:693-750."""
    if not args.training.setdefault('with_mask', True) and False:
        pass
    model_config = {
        'obj_bounding_radius': args.model.obj_bounding_radius,
        'W_geo_feat': args.model.setdefault('W_geometry_feature', 256),
        'use_outside_nerf': not args.training.with_mask,
        'speed_factor': args.training.setdefault('speed_factor', 1.0),
        'variance_init': args.model.setdefault('variance_init', 0.05),
    }
    s, r = args.model.surface, args.model.radiance
    model_config['surface_cfg'] = {
        'embed_multires': s.setdefault('embed_multires', 6), 'radius_init': s.setdefault('radius_init', 1.0),
        'geometric_init': s.setdefault('geometric_init', True), 'D': s.setdefault('D', 8), 'W': s.setdefault('W', 256),
        'skips': s.setdefault('skips', [4]),
    }
    model_config['radiance_cfg'] = {
        'embed_multires': r.setdefault('embed_multires', -1), 'embed_multires_view': r.setdefault('embed_multires_view', -1),
        'use_view_dirs': r.setdefault('use_view_dirs', True), 'D': r.setdefault('D', 4), 'W': r.setdefault('W', 256),
        'skips': r.setdefault('skips', []),
    }
    model = NeuS(**model_config)
    render_kwargs_train = {
        'upsample_algo': args.model.setdefault('upsample_algo', 'official_solution'),
        'N_nograd_samples': args.model.setdefault('N_nograd_samples', 2048),
        'N_upsample_iters': args.model.setdefault('N_upsample_iters', 4),
        'N_outside': args.model.setdefault('N_outside', 0),
        'obj_bounding_radius': args.data.setdefault('obj_bounding_radius', 1.0),
        'batched': args.data.batch_size is not None,
        'perturb': args.model.setdefault('perturb', True),
        'white_bkgd': args.model.setdefault('white_bkgd', False),
    }
    render_kwargs_test = copy.deepcopy(render_kwargs_train)
    render_kwargs_test['rayschunk'] = args.data.val_rayschunk
    render_kwargs_test['perturb'] = False
    trainer = Trainer(model, args.device_ids, batched=render_kwargs_train['batched'], is_finetune=args.training.is_finetune,
                      target_hw=render_target)
    return model, trainer, render_kwargs_train, render_kwargs_test, trainer.renderer
