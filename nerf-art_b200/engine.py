"""Host-side engine: owns the packed weights + workspace of one model and issues the C-ABI calls.

Everything here is pointer plumbing around `libnerfart_b200.so`; tensors are allocated with torch so that the
caching allocator and the current stream are shared with the caller (train.py / render.py of the reference).
"""
import ctypes as C
import os
import torch

from . import _lib
from ._lib import (NaTrainCfg, NaRawGrads, NaNetDesc, NaRawParams, NaVolsdfCfg, NaVolsdfOut, NaNeusCfg, NaNeusOut, NaSurfaceCfg, NaSurfaceOut, check, ptr,
                   stream_ptr, NA_FRAMEWORK_VOLSDF, NA_FRAMEWORK_NEUS, NA_RAYCAST_ROOT_FINDING, NA_RAYCAST_SPHERE_TRACING, PRECISIONS)

_LINSPACE_CACHE = {}


def cpu_linspace(n, device):
    """torch.linspace(0, 1, n) evaluated on the CPU (the reference's oracle side) and shipped to `device` once.
    The reference builds these tables at volsdf.py:472,483 / rend_util.py:269,304 / neus.py:235."""
    key = (int(n), str(device))
    t = _LINSPACE_CACHE.get(key)
    if t is None:
        t = torch.linspace(0.0, 1.0, int(n), dtype=torch.float32).to(device)
        _LINSPACE_CACHE[key] = t
    return t


class NetEngine:
    def __init__(self, implicit_surface, radiance_net, framework, multires_view, bounding_radius):
        self.surface, self.radiance = implicit_surface, radiance_net
        self.framework = framework
        self.desc = NaNetDesc(NA_FRAMEWORK_VOLSDF if framework == 'volsdf' else NA_FRAMEWORK_NEUS,
                              int(multires_view), float(bounding_radius), 0.0)
        self.packed = None
        self._pack_key = None
        self._ws = None
        self._dummy = None
        from . import default_precision
        self.precision = default_precision()

    # ------------------------------------------------------------------------------------------
    def _device(self):
        return self.surface.surface_fc_layers[0].weight_v.device

    def _raw_params(self):
        dev = self._device()
        if dev.type != 'cuda':
            raise RuntimeError('nerfart_b200: the model must live on a CUDA device (no CPU path exists)')
        raw = NaRawParams()
        keep = []
        layers = list(self.surface.surface_fc_layers)
        if self.radiance is not None:
            layers += list(self.radiance.layers)
        else:                                           # stand-alone ImplicitSurface: zero radiance net of the right shapes
            if self._dummy is None:
                sd = 9 if self.desc.multires_view < 0 else 33
                shapes = [(256, 256 + sd), (256, 256), (256, 256), (256, 256), (3, 256)]
                self._dummy = [(torch.zeros(o, device=dev), torch.ones(o, 1, device=dev), torch.ones(o, i, device=dev))
                               for o, i in shapes]
            layers += self._dummy
        assert len(layers) == 14
        for i, l in enumerate(layers):
            b, g, v = (l.bias, l.weight_g, l.weight_v) if not isinstance(l, tuple) else l
            for t in (b, g, v):
                if t.dtype != torch.float32 or not t.is_contiguous():
                    raise RuntimeError('parameters must be contiguous fp32')
            raw.bias[i], raw.weight_g[i], raw.weight_v[i] = b.data_ptr(), g.data_ptr(), v.data_ptr()
            keep += [b, g, v]
        return raw, keep

    def pack(self):
        """Fold weight-norm etc. into the GEMM-ready planes (na_pack_weights, 7 launches).  Called at the start of every render /
        eval, but the launches are issued only when the parameters changed since the last pack: every parameter tensor's
        (data_ptr, torch version counter) and the launch stream are compared -- optimiser steps, load_state_dict and .to() all bump one --
        so a 90-view render.py run or the 109 renders of one fine-tune step pack once.  `invalidate()` forces a re-pack after
        writes torch cannot see (e.g. a foreign kernel writing through data_ptr)."""
        L = _lib.lib()
        dev = self._device()
        if os.environ.get('NA_PRELOAD', '1') != '0':
            _lib.preload_kernels(dev)              # once per device: every kernel image is resident before the first launch
        raw, keep = self._raw_params()
        key = (str(dev), torch.cuda.current_stream(dev).cuda_stream) + tuple((t.data_ptr(), t._version) for t in keep)
        if self.packed is not None and self.packed.device == dev and key == self._pack_key and os.environ.get('NA_PACK_ALWAYS') != '1':
            return self.packed
        nbytes = L.na_packed_weights_bytes(C.byref(self.desc))
        if self.packed is None or self.packed.device != dev or self.packed.numel() * 4 < nbytes:
            self.packed = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(L.na_pack_weights(C.byref(self.desc), C.byref(raw), ptr(self.packed), stream_ptr(dev)), 'na_pack_weights')
        self._pack_key = key
        return self.packed

    def invalidate(self):
        """Force the next pack() to re-fold the weights."""
        self._pack_key = None

    def workspace(self, nbytes):
        dev = self._device()
        if self._ws is None or self._ws.device != dev or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        return self._ws

    # ------------------------------------------------------------------------------------------
    def sdf_eval(self, x, apply_bg, want_feat=False):
        L = _lib.lib()
        shape = x.shape[:-1]
        xf = x.detach().reshape(-1, 3).float().contiguous()
        m = xf.shape[0]
        dev = xf.device
        sdf = torch.empty(m, device=dev, dtype=torch.float32)
        feat = torch.empty(m, 256, device=dev, dtype=torch.float32) if want_feat else None
        self.pack()
        ws = self.workspace(L.na_eval_workspace_bytes(m))
        with torch.cuda.device(dev):
            check(L.na_sdf_eval(C.byref(self.desc), ptr(self.packed), ptr(xf), m, int(bool(apply_bg)), PRECISIONS[self.precision],
                                ptr(sdf), ptr(feat), ptr(ws), ws.numel(), stream_ptr(dev)), 'na_sdf_eval')
        return sdf.reshape(shape), (feat.reshape(*shape, 256) if want_feat else None)

    def full_eval(self, x, view_dirs, want_radiance=True, apply_bg=None, want_feat=True):
        L = _lib.lib()
        shape = x.shape[:-1]
        xf = x.detach().reshape(-1, 3).float().contiguous()
        m = xf.shape[0]
        dev = xf.device
        vf = view_dirs.detach().reshape(-1, 3).float().contiguous() if view_dirs is not None else None
        rad = torch.empty(m, 3, device=dev, dtype=torch.float32) if want_radiance else None
        sdf = torch.empty(m, device=dev, dtype=torch.float32)
        nab = torch.empty(m, 3, device=dev, dtype=torch.float32)
        feat = torch.empty(m, 256, device=dev, dtype=torch.float32) if want_feat else None
        self.pack()
        ws = self.workspace(L.na_eval_workspace_bytes(m))
        desc = self.desc
        if apply_bg is not None and bool(apply_bg) != (desc.framework == NA_FRAMEWORK_VOLSDF):
            desc = NaNetDesc(NA_FRAMEWORK_VOLSDF if apply_bg else NA_FRAMEWORK_NEUS, desc.multires_view, desc.bounding_radius, 0.0)
        with torch.cuda.device(dev):
            check(L.na_full_eval(C.byref(desc), ptr(self.packed), ptr(xf), ptr(vf), m, PRECISIONS[self.precision],
                                 ptr(rad), ptr(sdf), ptr(nab), ptr(feat), ptr(ws), ws.numel(), stream_ptr(dev)), 'na_full_eval')
        return (rad.reshape(*shape, 3) if want_radiance else None), sdf.reshape(shape), nab.reshape(*shape, 3), (feat.reshape(*shape, 256) if want_feat else None)

    # ------------------------------------------------------------------------------------------
    def volsdf_render(self, rays_o, rays_d, alpha_beta, *, near, far, N_samples, N_importance, max_upsample_steps,
                      max_bisection_steps, epsilon, white_bkgd, perturb, calc_normal, detailed_output, u_final=None, train_stash=False):
        """rays [N,3] (un-normalised directions) -> dict of flat outputs.  volsdf.volume_render, volsdf.py:389-615.
        train_stash (tensor-core modes, detailed_output): the final full evaluation also is the forward half of the training program
        (na_volsdf_render_fwd_train): `render_bwd` on the returned outputs then runs the backward half only."""
        L = _lib.lib()
        dev = rays_o.device
        n = rays_o.shape[0]
        P = N_samples + N_importance
        cfg = NaVolsdfCfg(int(N_samples), int(N_importance), int(max_upsample_steps), int(max_bisection_steps), float(near),
                          float(far), float(epsilon), int(bool(white_bkgd)), int(bool(perturb)), PRECISIONS[self.precision],
                          int(bool(detailed_output)), 0)
        f32 = dict(device=dev, dtype=torch.float32)
        o = dict(rgb=torch.empty(n, 3, **f32), depth=torch.empty(n, **f32), acc=torch.empty(n, **f32),
                 normals=torch.empty(n, 3, **f32) if calc_normal else None,
                 beta_map=torch.empty(n, **f32), iter_usage=torch.empty(n, **f32))
        if detailed_output:
            o.update(d_vals=torch.empty(n, P, **f32), sdf=torch.empty(n, P, **f32), nablas=torch.empty(n, P, 3, **f32),
                     radiance=torch.empty(n, P, 3, **f32), sigma=torch.empty(n, P, **f32), tau=torch.empty(n, P - 1, **f32))
        out = NaVolsdfOut(*[ptr(o.get(k)) for k in ('rgb', 'depth', 'acc', 'normals', 'beta_map', 'iter_usage',
                                                     'd_vals', 'sdf', 'nablas', 'radiance', 'sigma', 'tau')])
        if perturb and u_final is None:
            u_final = torch.rand(n, N_importance, **f32)           # rend_util.py:307 draws these with torch.rand
        self.pack()
        ws = self.workspace(L.na_volsdf_workspace_bytes(C.byref(cfg), n))
        tc, ti = cpu_linspace(N_samples, dev), cpu_linspace(4 * N_samples, dev)
        uu, ui = cpu_linspace(4 * N_samples + 2, dev), cpu_linspace(N_importance, dev)
        self._stash_key = None
        if train_stash:
            if not detailed_output or self.precision not in ('tc', 'tc_mixed'):
                raise RuntimeError("train_stash needs detailed_output and precision 'tc' / 'tc_mixed'")
            tws = self._train_workspace(n, P, dev)
        with torch.cuda.device(dev):
            if train_stash:
                check(L.na_volsdf_render_fwd_train(C.byref(self.desc), ptr(self.packed), C.byref(cfg), ptr(rays_o), ptr(rays_d), n,
                                                   ptr(alpha_beta), ptr(tc), ptr(ti), ptr(uu), ptr(ui),
                                                   ptr(u_final.contiguous()) if u_final is not None else None,
                                                   C.byref(out), ptr(ws), ws.numel(), ptr(tws), tws.numel(), stream_ptr(dev)),
                      'na_volsdf_render_fwd_train')
                # what the stash belongs to: the backward half is only valid for these outputs, these weights and this precision
                self._stash_key = (n, P, o['d_vals'].data_ptr(), o['radiance'].data_ptr(), self.precision, self._pack_key)
            else:
                check(L.na_volsdf_render_fwd(C.byref(self.desc), ptr(self.packed), C.byref(cfg), ptr(rays_o), ptr(rays_d), n,
                                             ptr(alpha_beta), ptr(tc), ptr(ti), ptr(uu), ptr(ui),
                                             ptr(u_final.contiguous()) if u_final is not None else None,
                                             C.byref(out), ptr(ws), ws.numel(), stream_ptr(dev)), 'na_volsdf_render_fwd')
        return o

    def _train_workspace(self, n, P, dev):
        L = _lib.lib()
        nbytes = L.na_train_workspace_bytes_mode(C.byref(self.desc), n, P, PRECISIONS[self.precision])
        if getattr(self, '_tws', None) is None or self._tws.device != dev or self._tws.numel() < nbytes:
            self._tws = None
            self._tws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        return self._tws

    def neus_render(self, rays_o, rays_d, s_dev, *, obj_bounding_radius, N_samples, N_importance, N_upsample_iters,
                    white_bkgd, perturb, detailed_output, u_rand=None, train_stash=False):
        """neus.volume_render ('official_solution', N_outside=0), neus.py:142-424.  train_stash: see volsdf_render."""
        L = _lib.lib()
        dev = rays_o.device
        n = rays_o.shape[0]
        P = N_samples + N_importance
        cfg = NaNeusCfg(int(N_samples), int(N_importance), int(N_upsample_iters), float(obj_bounding_radius),
                        int(bool(white_bkgd)), int(bool(perturb)), PRECISIONS[self.precision], int(bool(detailed_output)))
        f32 = dict(device=dev, dtype=torch.float32)
        o = dict(rgb=torch.empty(n, 3, **f32), depth=torch.empty(n, **f32), acc=torch.empty(n, **f32),
                 normals=torch.empty(n, 3, **f32))
        if detailed_output:
            o.update(d_all=torch.empty(n, P, **f32), sdf=torch.empty(n, P, **f32), nablas=torch.empty(n, P, 3, **f32),
                     radiance=torch.empty(n, P - 1, 3, **f32), alpha=torch.empty(n, P - 1, **f32),
                     weights=torch.empty(n, P - 1, **f32))
        out = NaNeusOut(*[ptr(o.get(k)) for k in ('rgb', 'depth', 'acc', 'normals', 'd_all', 'sdf', 'nablas', 'radiance',
                                                   'alpha', 'weights')])
        n_new = N_importance // N_upsample_iters
        if perturb and u_rand is None:
            u_rand = torch.rand(N_upsample_iters, n, n_new, **f32)
        self.pack()
        ws = self.workspace(L.na_neus_workspace_bytes(C.byref(cfg), n))
        tc, ui = cpu_linspace(N_samples, dev), cpu_linspace(n_new, dev)
        self._stash_key = None
        if train_stash:
            if not detailed_output or self.precision not in ('tc', 'tc_mixed'):
                raise RuntimeError("train_stash needs detailed_output and precision 'tc' / 'tc_mixed'")
            tws = self._train_workspace(n, P, dev)
        with torch.cuda.device(dev):
            if train_stash:
                check(L.na_neus_render_fwd_train(C.byref(self.desc), ptr(self.packed), C.byref(cfg), ptr(rays_o), ptr(rays_d), n,
                                                 ptr(s_dev), ptr(tc), ptr(ui), ptr(u_rand.contiguous()) if u_rand is not None else None,
                                                 C.byref(out), ptr(ws), ws.numel(), ptr(tws), tws.numel(), stream_ptr(dev)),
                      'na_neus_render_fwd_train')
                self._stash_key = (n, P, o['d_all'].data_ptr(), o['radiance'].data_ptr(), self.precision, self._pack_key)
            else:
                check(L.na_neus_render_fwd(C.byref(self.desc), ptr(self.packed), C.byref(cfg), ptr(rays_o), ptr(rays_d), n,
                                           ptr(s_dev), ptr(tc), ptr(ui), ptr(u_rand.contiguous()) if u_rand is not None else None,
                                           C.byref(out), ptr(ws), ws.numel(), stream_ptr(dev)), 'na_neus_render_fwd')
        return o

    # ------------------------------------------------------------------------------------------
    # backward of the render (second pass of the fine-tune step; include/nerfart_b200.h "Backward of the render")
    def grad_zero(self):
        """Fresh gradient accumulators: GradPack (effective-weight gradients) and the float64 scalars
        [d loss/d ln_beta or ln_s, eikonal loss].  The counterpart of optimizer.zero_grad() (volsdf.py:753)."""
        L = _lib.lib()
        dev = self._device()
        n = L.na_grad_pack_bytes(C.byref(self.desc)) // 4
        if getattr(self, '_gpack', None) is None or self._gpack.device != dev:
            self._gpack = torch.zeros(n, dtype=torch.float32, device=dev)
            self._gscal = torch.zeros(2, dtype=torch.float64, device=dev)
        else:
            self._gpack.zero_(); self._gscal.zero_()

    def render_bwd(self, rays_o, rays_d, scal, fwd, grad_rgb, *, w_eikonal, eikonal_count, white_bkgd, speed_factor,
                   train_surface=True, train_radiance=True):
        """Accumulate the gradients of one ray patch.  `fwd` = the flat detailed outputs of `volsdf_render` / `neus_render`
        for the same rays; `scal` = {alpha, beta} (VolSDF) or {s} (NeuS) on the device; grad_rgb [n,3]."""
        L = _lib.lib()
        dev = rays_o.device
        n = rays_o.shape[0]
        neus = self.framework != 'volsdf'
        d_all = fwd['d_all'] if neus else fwd['d_vals']
        P = d_all.shape[-1]
        cfg = NaTrainCfg(int(P), float(w_eikonal), int(eikonal_count), int(bool(white_bkgd)), float(speed_factor),
                         int(bool(train_surface)), int(bool(train_radiance)), PRECISIONS[self.precision])
        # the forward render of exactly these outputs left its stash in the workspace (volsdf_render(train_stash=True)): backward half only
        stashed = getattr(self, '_stash_key', None) is not None and \
            self._stash_key == (n, P, d_all.data_ptr(), fwd['radiance'].data_ptr(), self.precision, self._pack_key)
        self._stash_key = None
        self._train_workspace(n, P, dev)
        g = grad_rgb.reshape(-1, 3).float().contiguous()
        fn = (L.na_neus_render_bwd_stashed if stashed else L.na_neus_render_bwd) if neus else \
            (L.na_volsdf_render_bwd_stashed if stashed else L.na_volsdf_render_bwd)
        with torch.cuda.device(dev):
            check(fn(C.byref(self.desc), ptr(self.packed), C.byref(cfg), ptr(rays_o), ptr(rays_d), n, ptr(scal), ptr(d_all),
                     ptr(fwd['sdf']), ptr(fwd['radiance']), ptr(fwd['nablas']), ptr(g), ptr(self._gpack),
                     C.c_void_p(self._gscal.data_ptr()), ptr(self._tws), self._tws.numel(), stream_ptr(dev)),
                  'na_neus_render_bwd' if neus else 'na_volsdf_render_bwd')

    def unpack_grads(self, train_surface=True, train_radiance=True):
        """GradPack -> list of (parameter, gradient) for bias / weight_g / weight_v of the trained layers, plus the scalars."""
        L = _lib.lib()
        dev = self._device()
        raw, keep = self._raw_params()
        out = NaRawGrads()
        pairs = []
        layers = list(self.surface.surface_fc_layers) + (list(self.radiance.layers) if self.radiance is not None else [])
        for i, l in enumerate(layers):
            if (i < 9 and not train_surface) or (i >= 9 and not train_radiance):
                continue
            gb, gg, gv = torch.zeros_like(l.bias), torch.zeros_like(l.weight_g), torch.zeros_like(l.weight_v)
            out.bias[i], out.weight_g[i], out.weight_v[i] = gb.data_ptr(), gg.data_ptr(), gv.data_ptr()
            pairs += [(l.bias, gb), (l.weight_g, gg), (l.weight_v, gv)]
        with torch.cuda.device(dev):
            check(L.na_unpack_grads(C.byref(self.desc), C.byref(raw), ptr(self._gpack), C.byref(out), stream_ptr(dev)), 'na_unpack_grads')
        del keep
        return pairs, self._gscal

    # ------------------------------------------------------------------------------------------
    def _surface_cfg(self, algo, near, far, N_steps, N_secant_steps, N_iters, logit_tau, fill_inf):
        if isinstance(near, torch.Tensor) or isinstance(far, torch.Tensor):
            raise NotImplementedError('per-ray near / far tensors are not used by render.py and are not implemented')
        if algo not in ('root_finding', 'sphere_tracing'):
            raise NotImplementedError                       # ray_casting.py:233
        return NaSurfaceCfg(NA_RAYCAST_ROOT_FINDING if algo == 'root_finding' else NA_RAYCAST_SPHERE_TRACING, int(N_steps),
                            int(N_secant_steps), int(N_iters), float(near), float(far), float(logit_tau), int(bool(fill_inf)), 1,
                            PRECISIONS[self.precision])

    def ray_cast(self, rays_o, rays_d_unit, algo, *, near=0.0, far=6.0, N_steps=256, N_secant_steps=8, N_iters=20, logit_tau=0.0,
                 fill_inf=True):
        """root_finding_surface_points / sphere_tracing_surface_points (ray_casting.py:35-184) on flat [N,3] rays with unit directions."""
        L = _lib.lib()
        dev = rays_o.device
        n = rays_o.shape[0]
        cfg = self._surface_cfg(algo, near, far, N_steps, N_secant_steps, N_iters, logit_tau, fill_inf)
        depth = torch.empty(n, device=dev, dtype=torch.float32)
        pts = torch.empty(n, 3, device=dev, dtype=torch.float32)
        mask = torch.empty(n, device=dev, dtype=torch.uint8)
        msc = torch.empty(n, device=dev, dtype=torch.uint8)
        self.pack()
        ws = self.workspace(L.na_surface_workspace_bytes(C.byref(cfg), n))
        ts = cpu_linspace(N_steps, dev) if algo == 'root_finding' else None
        with torch.cuda.device(dev):
            check(L.na_ray_cast(C.byref(self.desc), ptr(self.packed), C.byref(cfg), ptr(rays_o), ptr(rays_d_unit), n, ptr(ts),
                                ptr(depth), ptr(pts), C.c_void_p(mask.data_ptr()), C.c_void_p(msc.data_ptr()), ptr(ws), ws.numel(),
                                stream_ptr(dev)), 'na_ray_cast')
        return depth, pts, mask.bool(), msc.bool()

    def surface_render(self, rays_o, rays_d, algo, calc_normal=True, **cfgs):
        """ray_casting.surface_render (187-263) on flat [N,3] rays (directions un-normalised)."""
        L = _lib.lib()
        dev = rays_o.device
        n = rays_o.shape[0]
        kw = dict(near=0.0, far=6.0, N_steps=256, N_secant_steps=8, N_iters=20, logit_tau=0.0, fill_inf=True)
        for k, v in cfgs.items():
            if k not in kw:
                raise TypeError(f'unexpected ray-casting option {k!r}')
            kw[k] = v
        cfg = self._surface_cfg(algo, kw['near'], kw['far'], kw['N_steps'], kw['N_secant_steps'], kw['N_iters'], kw['logit_tau'], kw['fill_inf'])
        f32 = dict(device=dev, dtype=torch.float32)
        o = dict(rgb=torch.empty(n, 3, **f32), depth=torch.empty(n, **f32), mask=torch.empty(n, device=dev, dtype=torch.uint8),
                 nablas=torch.empty(n, 3, **f32), normals=torch.empty(n, 3, **f32) if calc_normal else None)
        out = NaSurfaceOut(ptr(o['rgb']), ptr(o['depth']), C.c_void_p(o['mask'].data_ptr()), ptr(o['nablas']), ptr(o['normals']))
        self.pack()
        ws = self.workspace(L.na_surface_workspace_bytes(C.byref(cfg), n))
        ts = cpu_linspace(kw['N_steps'], dev) if algo == 'root_finding' else None
        with torch.cuda.device(dev):
            check(L.na_surface_render_fwd(C.byref(self.desc), ptr(self.packed), C.byref(cfg), ptr(rays_o), ptr(rays_d), n, ptr(ts),
                                          C.byref(out), ptr(ws), ws.numel(), stream_ptr(dev)), 'na_surface_render_fwd')
        o['mask'] = o['mask'].bool()
        return o
