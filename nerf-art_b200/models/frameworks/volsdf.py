"""VolSDF behind the reference's API (reference models/frameworks/volsdf.py), computed by the sm_100a kernels.

Kept from the reference, because train.py / render.py / checkpoints depend on it:
  * class / function names and signatures: VolSDF, volume_render, SingleRenderer, Trainer, get_model
  * state_dict keys (ln_beta, implicit_surface.*, radiance_net.*), `forward_ab`, `forward_surface`, `forward`, `fix_module`
  * volume_render's kwargs (volsdf.py:389-424), batched / unbatched ray layouts and the `(rgb, depth, extras)` contract
    with the extras keys of volsdf.py:566-594.
Everything numerical happens in csrc/ (one C-ABI call per render); nothing here falls back to PyTorch math.
"""
import copy
import weakref
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from ..base import ImplicitSurface, RadianceNet
from ...engine import NetEngine

FIX_MODULE = None          # volsdf.py:7-8


class VolSDF(nn.Module):
    def __init__(self, beta_init=0.1, speed_factor=1.0, input_ch=3, W_geo_feat=-1, obj_bounding_radius=3.0,
                 use_nerfplusplus=False, surface_cfg=dict(), radiance_cfg=dict()):
        super().__init__()
        if use_nerfplusplus:
            raise NotImplementedError("outside_scene 'nerf++' is not used by any shipped config and is out of scope (SURVEY.md section 2 row 1)")
        self.speed_factor = speed_factor
        self.ln_beta = nn.Parameter(data=torch.Tensor([np.log(beta_init) / self.speed_factor]), requires_grad=True)
        self.use_sphere_bg = True
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
            self._engine = NetEngine(self.implicit_surface, self.radiance_net, 'volsdf',
                                     self.radiance_net.embed_multires_view, self.obj_bounding_radius)
        return self._engine

    def forward_ab(self):
        beta = torch.exp(self.ln_beta * self.speed_factor)
        return 1. / beta, beta

    def forward_surface(self, x: torch.Tensor):
        """(min(sdf, R-||x||), feature) -- reference volsdf.py:341-347."""
        return self.engine().sdf_eval(x, apply_bg=True, want_feat=True)

    def forward_surface_with_nablas(self, x: torch.Tensor):
        _, sdf, nab, feat = self.engine().full_eval(x, None, want_radiance=False, apply_bg=True)
        return sdf, nab, feat

    def forward(self, x: torch.Tensor, view_dirs: torch.Tensor = None, return_nablas=False):
        """(radiance, sdf, nablas) -- reference volsdf.py:359-370 (view_dirs is always given on the shipped configs)."""
        if view_dirs is None:
            raise NotImplementedError('use_view_dirs=False is not a shipped configuration')
        rad, sdf, nab, _ = self.engine().full_eval(x, view_dirs, want_radiance=True, apply_bg=True)
        return rad, sdf, nab

    def fix_module(self, module_name):
        if module_name is None or module_name == "":
            return
        if module_name not in ('implicit_surface', 'radiance_net'):
            raise NotImplementedError(f"{module_name} is not a valid module.")
        for p in getattr(self, module_name).parameters():
            p.requires_grad = False


def volume_render(rays_o, rays_d, model: VolSDF, near=0.0, far=6.0, obj_bounding_radius=3.0, batched=False, batched_info={},
                  require_nablas=False, calc_normal=True, use_view_dirs=True, rayschunk=4000, netchunk=1048576,
                  white_bkgd=False, use_nerfplusplus=False, detailed_output=True, show_progress=False, perturb=False,
                  N_samples=128, N_importance=64, N_outside=32, max_upsample_steps=5, max_bisection_steps=10, epsilon=0.1,
                  u_final=None, **dummy_kwargs):
    """Same contract as the reference's volume_render (volsdf.py:389-615).  `rayschunk` / `netchunk` are accepted and
    ignored: the fused kernels keep activations on chip, so there is nothing to chunk for memory.
    `u_final` [N_rays, N_importance] optionally injects the uniform draws used when perturb=True (rend_util.py:307)."""
    if use_nerfplusplus or not use_view_dirs:
        raise NotImplementedError('nerf++ background / use_view_dirs=False are not shipped configurations')
    B = rays_d.shape[0] if batched else None
    ro = rays_o.reshape(-1, 3).float().contiguous()
    rd = rays_d.reshape(-1, 3).float().contiguous()
    alpha, beta = model.forward_ab()
    ab = torch.cat([alpha.detach().reshape(1), beta.detach().reshape(1)]).float().contiguous()
    o = model.engine().volsdf_render(
        ro, rd, ab, near=near, far=far, N_samples=N_samples, N_importance=N_importance,
        max_upsample_steps=max_upsample_steps, max_bisection_steps=max_bisection_steps, epsilon=epsilon,
        white_bkgd=white_bkgd, perturb=perturb, calc_normal=bool(calc_normal and require_nablas),
        detailed_output=detailed_output, u_final=None if u_final is None else u_final.reshape(-1, N_importance).float())

    def shp(t, *tail):
        return t.reshape(*((B, -1) if batched else (-1,)), *tail)

    P = N_samples + N_importance
    ret = OrderedDict([('rgb', shp(o['rgb'], 3)), ('depth_volume', shp(o['depth'])), ('mask_volume', shp(o['acc']))])
    if calc_normal and require_nablas:
        ret['normals_volume'] = shp(o['normals'], 3)
    if detailed_output:
        ret['implicit_surface'] = shp(o['sdf'], P)
        if require_nablas:
            ret['implicit_nablas'] = shp(o['nablas'], P, 3)
        ret['radiance'] = shp(o['radiance'], P, 3)
        sigma = shp(o['sigma'], P)
        tau = shp(o['tau'], P - 1)
        d_vals = shp(o['d_vals'], P)
        p_i = torch.exp(-torch.relu(sigma[..., :-1] * (d_vals[..., 1:] - d_vals[..., :-1])))   # logging-only extras
        ret['alpha'] = 1.0 - p_i
        ret['p_i'] = p_i
        ret['visibility_weights'] = tau
        ret['d_vals'] = d_vals
        ret['sigma'] = sigma
        ret['beta_map'] = shp(o['beta_map'], 1)
        ret['iter_usage'] = shp(o['iter_usage'])
    return ret['rgb'], ret['depth_volume'], ret


class SingleRenderer(nn.Module):
    def __init__(self, model: VolSDF):
        super().__init__()
        self.model = model

    def forward(self, rays_o, rays_d, **kwargs):
        return volume_render(rays_o, rays_d, self.model, **kwargs)


def render_patch(model: VolSDF, ro, rd, near=0.0, far=6.0, perturb=False, white_bkgd=False, max_upsample_steps=5, N_samples=128,
                 N_importance=64, max_bisection_steps=10, epsilon=0.1, u_final=None, train_stash=False, **dummy_kwargs):
    """Forward render of one flat ray patch with the detailed per-sample outputs the backward needs (volume_render's
    defaults, volsdf.py:389-424).  Returns (flat outputs, {alpha, beta} on the device).
    train_stash: this render is the forward half of the training program (NetEngine.volsdf_render): the reference renders the patch
    with grad and back-propagates through that one evaluation (volsdf.py:760-783); `engine.render_bwd` on these outputs then runs the
    backward half only instead of re-evaluating the networks."""
    alpha, beta = model.forward_ab()
    ab = torch.cat([alpha.detach().reshape(1), beta.detach().reshape(1)]).float().contiguous()
    o = model.engine().volsdf_render(ro, rd, ab, near=near, far=far, N_samples=N_samples, N_importance=N_importance,
                                     max_upsample_steps=max_upsample_steps, max_bisection_steps=max_bisection_steps,
                                     epsilon=epsilon, white_bkgd=white_bkgd, perturb=perturb, calc_normal=False,
                                     detailed_output=True, u_final=u_final, train_stash=train_stash)
    return o, ab


class Trainer(nn.Module):
    """The reference's Trainer (volsdf.py:627-837).  `forward` implements the CLIP fine-tune branch (719-786) on the CUDA
    backward kernels; the from-scratch reconstruction branch (787-823) is outside the hot path and raises.
    `loss_dict` (extension): the four style losses; by default they are built from nerfart_b200.criteria like the
    reference builds them from criteria/ (volsdf.py:639-645)."""

    def __init__(self, model: VolSDF, device_ids=[0], batched=True, is_finetune=False, target_hw: list = None, loss_dict=None):
        super().__init__()
        self.model = model
        self.renderer = SingleRenderer(model)
        self.device = device_ids[0] if isinstance(device_ids, (list, tuple)) and len(device_ids) else 0
        self.is_finetune = is_finetune
        self.neg_texts = None
        self.target_hw = target_hw if target_hw is not None else [960, 540]
        self.loss_dict = loss_dict
        if is_finetune:
            self.model.fix_module(FIX_MODULE)                           # volsdf.py:8,646-647: nothing is frozen for VolSDF

    def _losses(self):
        if self.loss_dict is None:
            from ...criteria import build_loss_dict
            self.loss_dict = build_loss_dict(self.target_hw, next(self.model.parameters()).device)
        return self.loss_dict

    def forward(self, args, indices, model_input, ground_truth, render_kwargs_train: dict, it: int, optimizer=None):
        if not args.training.is_finetune:
            raise NotImplementedError('nerfart_b200: only the CLIP fine-tune branch of Trainer.forward (volsdf.py:719-786) is '
                                      'built; from-scratch reconstruction (787-823) is outside the accelerated hot path')
        from ._finetune import finetune_forward
        self._losses()
        losses, select_inds = finetune_forward(self, 'volsdf', args, model_input, ground_truth, render_kwargs_train, optimizer,
                                               lambda ro, rd, **kw: render_patch(self.model, ro, rd, **kw))
        extras = {}
        alpha, beta = self.model.forward_ab()
        extras['scalars'] = {'beta': beta.data, 'alpha': alpha.data}
        extras['select_inds'] = select_inds
        return OrderedDict([('losses', losses), ('extras', extras)])


    def val(self, logger, ret, to_img_fn, it, render_kwargs_test):
        """reference Trainer.val (volsdf.py:840-876): two TensorBoard figures from the detailed validation render -- the per-ray
        beta heat map and the number of upsampling iterations each ray used (iter_usage == -1, never converged, is shown as
        max_upsample_steps + 1).  Host-side logging of the caller (train.py:205-206); needs matplotlib and the reference's
        utils.io_util.gallery (train.py's own tree)."""
        import matplotlib.pyplot as plt
        from utils import io_util

        def tiled(per_ray):                                     # [B,N,1] -> gallery of the B validation images, [H',W',1]
            maps = to_img_fn(per_ray).permute(0, 2, 3, 1).data.cpu().numpy()
            return io_util.gallery(maps, int(np.sqrt(maps.shape[0])))

        def heat_map(img, lo, hi, ticks, labels, tag):
            fig = plt.figure(figsize=(5, 3), dpi=100)
            ax = fig.add_subplot(111)
            bar = fig.colorbar(ax.imshow(img, vmin=lo, vmax=hi), ticks=ticks)
            bar.ax.set_yticklabels(labels)
            logger.add_figure(fig, tag, it)

        beta_map = tiled(ret['beta_map'])
        beta = self.model.forward_ab()[1].data.cpu().numpy().item()
        beta_max = beta_map.max().item()
        ticks = np.linspace(beta, beta_max, 10).tolist() if beta_max != beta else [beta]
        labels = ['{:.4f}'.format(b) for b in ticks]
        labels[0] = 'beta={:.4f}'.format(beta)
        heat_map(beta_map, beta, beta_max, ticks, labels, 'val/beta_heat_map')

        max_iter = render_kwargs_test['max_upsample_steps']
        usage = tiled(ret['iter_usage'].unsqueeze(-1))
        usage[usage == -1] = max_iter + 1
        ticks = list(range(max_iter + 2))
        labels = ['{:d}'.format(b) for b in ticks]
        labels[-1] = 'not converged'
        heat_map(usage, 0, max_iter + 1, ticks, labels, 'val/upsample_iters')


def get_model(args, render_target=None):
    """reference volsdf.get_model, volsdf.py:943-994: same defaults injected into `args`, same five-tuple."""
    model_config = {
        'use_nerfplusplus': args.model.setdefault('outside_scene', 'builtin') == 'nerf++',
        'obj_bounding_radius': args.model.obj_bounding_radius,
        'W_geo_feat': args.model.setdefault('W_geometry_feature', 256),
        'speed_factor': args.training.setdefault('speed_factor', 1.0),
        'beta_init': args.training.setdefault('beta_init', 0.1),
    }
    s, r = args.model.surface, args.model.radiance
    use_siren = args.model.setdefault('use_siren', False)
    model_config['surface_cfg'] = {
        'use_siren': s.setdefault('use_siren', use_siren), 'embed_multires': s.setdefault('embed_multires', 6),
        'radius_init': s.setdefault('radius_init', 1.0), 'geometric_init': s.setdefault('geometric_init', True),
        'D': s.setdefault('D', 8), 'W': s.setdefault('W', 256), 'skips': s.setdefault('skips', [4]),
    }
    model_config['radiance_cfg'] = {
        'use_siren': r.setdefault('use_siren', use_siren), 'embed_multires': r.setdefault('embed_multires', -1),
        'embed_multires_view': r.setdefault('embed_multires_view', -1), 'use_view_dirs': r.setdefault('use_view_dirs', True),
        'D': r.setdefault('D', 4), 'W': r.setdefault('W', 256), 'skips': r.setdefault('skips', []),
    }
    model = VolSDF(**model_config)
    render_kwargs_train = {
        'near': args.data.near, 'far': args.data.far, 'batched': True,
        'perturb': args.model.setdefault('perturb', True), 'white_bkgd': args.model.setdefault('white_bkgd', False),
        'max_upsample_steps': args.model.setdefault('max_upsample_iter', 5),
        'use_nerfplusplus': model_config['use_nerfplusplus'], 'obj_bounding_radius': args.model.obj_bounding_radius,
    }
    render_kwargs_test = copy.deepcopy(render_kwargs_train)
    render_kwargs_test['rayschunk'] = args.data.val_rayschunk
    render_kwargs_test['perturb'] = False
    trainer = Trainer(model, args.device_ids, batched=True, is_finetune=args.training.is_finetune, target_hw=render_target)
    return model, trainer, render_kwargs_train, render_kwargs_test, trainer.renderer
