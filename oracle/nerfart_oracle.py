"""CPU oracle (numpy, fp32) for the NeRF-Art volumetric-render hot path.

TEST INFRASTRUCTURE ONLY.  This file restates, in plain numpy, the algorithm of the reference
(cassiePython/NeRF-Art, a fork of ventusff/neurecon) for the path BASELINE.json names.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import it, and only as the checker or as the CPU baseline -- never as a product code path.  The
product (`nerf-art_b200`) fails loudly when its CUDA library is missing; it never falls back here.

Parity pinning: every function below is checked against outputs of the UNMODIFIED reference run
on CPU in this repository's build container (tests/golden/make_golden.py -> tests/golden/*.npz;
tests/test_oracle_golden.py).  The reference ships no tests / golden vectors of its own
(SURVEY.md section 4), so executing it on seeded synthetic state is the only pin available.

All citations are file:line in /root/reference.
Arrays are float32 unless stated; prefix sums follow torch's CPU kernels, which accumulate
float32 inputs in float64 (`at::acc_type<float, /*is_cuda=*/false>` = double) and round each
prefix back to float32.
"""
from collections import OrderedDict
from concurrent.futures import ThreadPoolExecutor
import os
import numpy as np

F32 = np.float32

# Row-parallel evaluation of the per-sample networks: numpy's elementwise kernels are single threaded, the reference's
# torch CPU kernels are not.  Splitting the rows over a thread pool (numpy releases the GIL inside ufuncs and BLAS) lets the
# oracle use every host core when it is timed as the CPU baseline; results do not depend on the split (row-wise functions).
_POOL = None
_CHUNK = 8192


def _pool():
    global _POOL
    if _POOL is None:
        _POOL = ThreadPoolExecutor(max_workers=max(1, os.cpu_count() or 1))
    return _POOL


def _rowwise(fn, *arrays):
    m = arrays[0].shape[0]
    if m <= _CHUNK:
        return fn(*arrays)
    try:
        from threadpoolctl import threadpool_limits
    except Exception:                                   # pragma: no cover
        threadpool_limits = None
    cuts = list(range(0, m, _CHUNK))

    def run(c0):
        return fn(*[a[c0:c0 + _CHUNK] for a in arrays])
    if threadpool_limits is not None:
        with threadpool_limits(limits=1, user_api='blas'):
            parts = list(_pool().map(run, cuts))
    else:
        parts = list(_pool().map(run, cuts))
    if isinstance(parts[0], tuple):
        return tuple(np.concatenate([p[i] for p in parts], axis=0) for i in range(len(parts[0])))
    return np.concatenate(parts, axis=0)


# ----------------------------------------------------------------------------------------------
# small torch-CPU-compatible helpers
# ----------------------------------------------------------------------------------------------
def _cumsum(x):
    """torch.cumsum(float32, dim=-1) on CPU: double accumulator, float32 outputs."""
    return np.cumsum(x.astype(np.float64), axis=-1).astype(F32)


def _cumprod(x):
    """torch.cumprod(float32, dim=-1) on CPU: double accumulator, float32 outputs."""
    return np.cumprod(x.astype(np.float64), axis=-1).astype(F32)


def _linspace(a, b, n):
    """torch.linspace(a, b, n) float32 on CPU: step=(b-a)/(n-1); first half a+step*i, second half b-step*(n-1-i)
    (aten/src/ATen/native/cpu/RangeFactoriesKernel.cpp).  The vectorised kernel fuses multiply and add (one rounding);
    emulated by doing the multiply-add in float64 (fp32 x small int is exact there) and rounding once."""
    a = F32(a); b = F32(b)
    if n == 1:
        return np.array([a], dtype=F32)
    step = F32((b - a) / F32(n - 1))
    i = np.arange(n)
    lo = (np.float64(a) + np.float64(step) * i).astype(F32)
    hi = (np.float64(b) - np.float64(step) * (n - 1 - i)).astype(F32)
    return np.where(i < n // 2, lo, hi).astype(F32)


def _normalize(v, eps=1e-12):
    """F.normalize(v, dim=-1): v / max(||v||, eps)."""
    n = np.sqrt(np.sum(v * v, axis=-1, keepdims=True, dtype=F32))
    return (v / np.maximum(n, F32(eps))).astype(F32)


def _searchsorted_left(cdf, u):
    """torch.searchsorted(cdf, u, right=False) row-wise: first i with cdf[i] >= u."""
    out = np.empty(u.shape, dtype=np.int64)
    for r in range(cdf.shape[0]):
        out[r] = np.searchsorted(cdf[r], u[r], side='left')
    return out


# ----------------------------------------------------------------------------------------------
# networks  (models/base.py)
# ----------------------------------------------------------------------------------------------
def embed(x, multires):
    """Embedder.forward, models/base.py:46-64 with get_embedder, 67-81:
    [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]; multires<0 -> identity."""
    if multires < 0:
        return x.astype(F32)
    out = [x]
    for k in range(multires):
        f = F32(2.0 ** k)
        out.append(np.sin(x * f))
        out.append(np.cos(x * f))
    return np.concatenate(out, axis=-1).astype(F32)


def weight_norm_eff(g, v):
    """nn.utils.weight_norm (dim=0), models/base.py:226-227,365-366: W[o,:] = g[o] * v[o,:]/||v[o,:]||."""
    n = np.sqrt(np.sum(v.astype(F32) ** 2, axis=1, keepdims=True, dtype=F32))
    return (v * (g.reshape(-1, 1) / n)).astype(F32)


def softplus100(z):
    """nn.Softplus(beta=100) (models/base.py:202), threshold 20: z if 100 z > 20 else log1p(exp(100 z))/100."""
    bz = z * F32(100.0)
    with np.errstate(over='ignore'):
        soft = np.log1p(np.exp(np.minimum(bz, F32(30.0)))) / F32(100.0)
    return np.where(bz > F32(20.0), z, soft).astype(F32)


def _sigmoid(x):
    with np.errstate(over='ignore'):
        return (F32(1.0) / (F32(1.0) + np.exp(-x))).astype(F32)


class Net:
    """Effective (weight-norm folded) parameters of one model, from a reference-layout state dict
    (key layout: SURVEY.md section 5 / models/base.py:226,365)."""

    def __init__(self, sd, framework):
        sd = {k: np.asarray(v, dtype=F32) for k, v in sd.items()}
        self.framework = framework
        self.sW, self.sb = [], []
        i = 0
        while f'implicit_surface.surface_fc_layers.{i}.weight_v' in sd:
            p = f'implicit_surface.surface_fc_layers.{i}.'
            self.sW.append(weight_norm_eff(sd[p + 'weight_g'], sd[p + 'weight_v']))
            self.sb.append(sd[p + 'bias'])
            i += 1
        self.D = len(self.sW) - 1
        self.rW, self.rb = [], []
        i = 0
        while f'radiance_net.layers.{i}.weight_v' in sd:
            p = f'radiance_net.layers.{i}.'
            self.rW.append(weight_norm_eff(sd[p + 'weight_g'], sd[p + 'weight_v']))
            self.rb.append(sd[p + 'bias'])
            i += 1
        self.skip = 4                        # models/base.py:135 skips=[4] in every shipped config
        self.multires = 6                    # embed_multires (configs/*.yaml)
        if framework == 'volsdf':
            self.speed = F32(10.0)
            self.ln_beta = F32(sd['ln_beta'].reshape(-1)[0])
            self.multires_view = -1
            self.bound = F32(3.0)
        else:
            self.speed = F32(10.0)
            self.ln_s = F32(sd['ln_s'].reshape(-1)[0])
            self.multires_view = 4
            self.bound = F32(1.0)

    # VolSDF.forward_ab, models/frameworks/volsdf.py:337-339
    def alpha_beta(self):
        beta = F32(np.exp(F32(self.ln_beta * self.speed)))
        return F32(F32(1.0) / beta), beta

    # NeuS.forward_s, models/frameworks/neus.py:116-117
    def s(self):
        return F32(np.exp(F32(self.ln_s * self.speed)))


def sdf_net(net, x, with_nablas=False):
    return _rowwise(lambda xx: _sdf_net(net, xx, with_nablas), x)


def _sdf_net(net, x, with_nablas=False):
    """ImplicitSurface.forward (models/base.py:243-263) and forward_with_nablas (265-282).
    The gradient is the closed-form reverse sweep autograd performs.
    x [M,3] -> sdf [M], feat [M,256] (, nabla [M,3])."""
    x = x.astype(F32)
    e = embed(x, net.multires)
    h = e
    zs = []
    inv_sqrt2 = F32(1.0) / F32(np.sqrt(2))       # base.py:250 divides by np.sqrt(2) (a float64 -> f32 tensor op)
    for i in range(net.D):
        if i == net.skip:
            h = (np.concatenate([h, e], axis=-1) / F32(np.sqrt(2))).astype(F32)
        z = h @ net.sW[i].T + net.sb[i]
        zs.append(z)
        h = softplus100(z)
    out = h @ net.sW[net.D].T + net.sb[net.D]
    sdf = out[:, 0].copy()
    feat = out[:, 1:].copy()
    if not with_nablas:
        return sdf, feat
    g = np.broadcast_to(net.sW[net.D][0], (x.shape[0], net.sW[net.D].shape[1])).astype(F32)
    ge = np.zeros_like(e)
    for i in range(net.D - 1, -1, -1):
        bz = zs[i] * F32(100.0)
        dsp = np.where(bz > F32(20.0), F32(1.0), _sigmoid(bz)).astype(F32)   # softplus' with torch's threshold
        g = (g * dsp) @ net.sW[i]
        if i == net.skip:
            g = (g / F32(np.sqrt(2))).astype(F32)
            n_h = net.sW[i].shape[1] - e.shape[1]
            ge += g[:, n_h:]
            g = g[:, :n_h]
    ge += g
    nab = ge[:, 0:3].copy()
    for k in range(net.multires):
        f = F32(2.0 ** k)
        gs = ge[:, 3 + 6 * k: 6 + 6 * k]
        gc = ge[:, 6 + 6 * k: 9 + 6 * k]
        nab += f * (gs * np.cos(x * f) - gc * np.sin(x * f))
    return sdf, feat, nab.astype(F32)


def radiance_net(net, x, view_dirs, nablas, feat):
    return _rowwise(lambda a, b, c, d: _radiance_net(net, a, b, c, d), x, view_dirs, nablas, feat)


def _radiance_net(net, x, view_dirs, nablas, feat):
    """RadianceNet.forward, models/base.py:372-391: cat[x, embed_view(v), nablas, feat] -> 4x(Linear,ReLU)
    -> Linear,Sigmoid."""
    h = np.concatenate([embed(x, -1), embed(view_dirs, net.multires_view), nablas, feat], axis=-1).astype(F32)
    n = len(net.rW)
    for i in range(n):
        h = h @ net.rW[i].T + net.rb[i]
        h = np.maximum(h, F32(0)) if i < n - 1 else _sigmoid(h)
    return h.astype(F32)


def volsdf_forward_surface(net, x):
    """VolSDF.forward_surface, models/frameworks/volsdf.py:341-347: min(sdf, R - ||x||)."""
    sdf, _ = sdf_net(net, x)
    r = np.sqrt(np.sum(x * x, axis=-1, dtype=F32))
    return np.minimum(sdf, net.bound - r).astype(F32)


def volsdf_forward(net, x, view_dirs):
    """VolSDF.forward (359-370) via forward_surface_with_nablas (349-357): sdf overridden by the
    sphere background where R-||x|| < sdf, nablas NOT replaced, radiance from the raw nablas."""
    sdf, feat, nab = sdf_net(net, x, with_nablas=True)
    d_bg = net.bound - np.sqrt(np.sum(x * x, axis=-1, dtype=F32))
    sdf = np.where(d_bg < sdf, d_bg, sdf).astype(F32)
    rad = radiance_net(net, x, view_dirs, nab, feat)
    return rad, sdf, nab


# ----------------------------------------------------------------------------------------------
# rays  (utils/rend_util.py)
# ----------------------------------------------------------------------------------------------
def get_rays(c2w, K, H, W):
    """rend_util.get_rays (112-165) + lift (95-109), N_rays=-1: integer pixel coords (no +0.5),
    ray index = h*W + w, rays_d = c2w @ [x_lift, y_lift, 1, 1] - cam_loc (un-normalised)."""
    c2w = c2w.astype(F32); K = K.astype(F32)
    i = np.tile(np.arange(W, dtype=F32), H)          # x = w
    j = np.repeat(np.arange(H, dtype=F32), W)        # y = h
    fx, fy, cx, cy, sk = K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[0, 1]
    z = np.ones_like(i)
    x_lift = (i - cx + cy * sk / fy - sk * j / fy) / fx * z
    y_lift = (j - cy) / fy * z
    pix = np.stack([x_lift, y_lift, z, np.ones_like(z)], axis=0).astype(F32)   # [4, HW]
    world = (c2w @ pix).T[:, :3]
    cam = c2w[:3, 3]
    rays_d = (world - cam[None, :]).astype(F32)
    rays_o = np.broadcast_to(cam[None, :], rays_d.shape).astype(F32).copy()
    return rays_o, rays_d


def near_far_from_sphere(rays_o, rays_d, r=1.0):
    """rend_util.near_far_from_sphere, 168-186."""
    mid = -np.sum(rays_o * rays_d, axis=-1, keepdims=True, dtype=F32)
    near = np.maximum(mid - F32(r), F32(0.0))
    far = np.maximum(mid + F32(r), F32(r))
    return near.astype(F32), far.astype(F32)


# ----------------------------------------------------------------------------------------------
# inverse-CDF samplers  (utils/rend_util.py)
# ----------------------------------------------------------------------------------------------
def _invert_cdf(bins, cdf, u, eps=1e-5):
    """Shared tail of sample_pdf / sample_cdf, rend_util.py:276-293 / 311-328."""
    inds = _searchsorted_left(cdf, u)
    below = np.maximum(inds - 1, 0)
    above = np.minimum(inds, cdf.shape[-1] - 1)
    cdf0 = np.take_along_axis(cdf, below, axis=-1); cdf1 = np.take_along_axis(cdf, above, axis=-1)
    b0 = np.take_along_axis(bins, below, axis=-1); b1 = np.take_along_axis(bins, above, axis=-1)
    denom = (cdf1 - cdf0).astype(F32)
    denom = np.where(denom < F32(eps), F32(1.0), denom)
    t = ((u - cdf0) / denom).astype(F32)
    return (b0 + t * (b1 - b0)).astype(F32), inds


def sample_pdf(bins, weights, n, det=True, u=None, return_inds=False):
    """rend_util.sample_pdf, 256-293.  bins [M,N], weights [M,N-1].  `u` [M,n] must be given when
    det=False (the caller draws it; the reference uses torch.rand at 272)."""
    weights = (weights + F32(1e-5)).astype(F32)
    # torch.sum(float32) on CPU is a vectorised cascade sum whose association order depends on the host ISA; the exactly
    # rounded sum (float64 accumulate, one rounding) is within 1 ulp of it.  Only a sample drawn at u == 1.0 exactly can
    # see the difference (it decides whether cdf[-1] >= 1).
    pdf = (weights / np.sum(weights, axis=-1, keepdims=True, dtype=np.float64).astype(F32)).astype(F32)
    cdf = _cumsum(pdf)
    cdf = np.concatenate([np.zeros_like(cdf[:, :1]), cdf], axis=-1)
    if det:
        u = np.broadcast_to(_linspace(0.0, 1.0, n), (cdf.shape[0], n)).copy()
    s, inds = _invert_cdf(bins, cdf, u.astype(F32))
    return (s, inds) if return_inds else s


def sample_cdf(bins, cdf, n, det=True, u=None, return_inds=False):
    """rend_util.sample_cdf, 295-328.  cdf [M,N-1] is used as-is (NOT normalised)."""
    cdf = np.concatenate([np.zeros_like(cdf[:, :1]), cdf], axis=-1).astype(F32)
    if det:
        u = np.broadcast_to(_linspace(0.0, 1.0, n), (cdf.shape[0], n)).copy()
    s, inds = _invert_cdf(bins, cdf, u.astype(F32))
    return (s, inds) if return_inds else s


# ----------------------------------------------------------------------------------------------
# VolSDF  (models/frameworks/volsdf.py)
# ----------------------------------------------------------------------------------------------
def sdf_to_sigma(sdf, alpha, beta):
    """volsdf.sdf_to_sigma, 34-53."""
    e = (F32(0.5) * np.exp(-np.abs(sdf) / beta)).astype(F32)
    psi = np.where(sdf >= 0, e, F32(1.0) - e)
    return (alpha * psi).astype(F32)


def _R_t(d_vals, sdf, alpha, beta):
    """exclusive prefix sum of sigma_i * delta_i (volsdf.py:77-81, 128-132). [M,N] -> [M,N-1]."""
    sigma = sdf_to_sigma(sdf, alpha, beta)
    delta = (d_vals[:, 1:] - d_vals[:, :-1]).astype(F32)
    cs = _cumsum((sigma[:, :-1] * delta).astype(F32))
    return np.concatenate([np.zeros_like(cs[:, :1]), cs], axis=-1)[:, :-1], delta


def error_bound(d_vals, sdf, alpha, beta):
    """volsdf.error_bound, 56-94.  alpha/beta scalars or [M,1].  [M,N] -> [M,N-1]."""
    alpha = np.asarray(alpha, dtype=F32); beta = np.asarray(beta, dtype=F32)
    R_t, delta = _R_t(d_vals, sdf, alpha, beta)
    sabs = np.abs(sdf)
    d_star = np.maximum(F32(0.5) * (sabs[:, :-1] + sabs[:, 1:] - delta), F32(0.0)).astype(F32)
    with np.errstate(over='ignore', invalid='ignore'):
        errors = (alpha / (F32(4) * beta) * (delta ** 2) * np.exp(-d_star / beta)).astype(F32)
        errors_t = _cumsum(errors)
        bounds = (np.exp(-R_t) * (np.exp(errors_t) - F32(1.0))).astype(F32)
    bounds[np.isnan(bounds)] = np.inf
    return bounds


def opacity_invert_cdf_sample(d_vals, sdf, alpha, beta, n, det=True, u=None):
    """fine_sample.opacity_invert_cdf_sample, volsdf.py:122-136."""
    R_t, _ = _R_t(d_vals, sdf, np.asarray(alpha, dtype=F32), np.asarray(beta, dtype=F32))
    opacity = (F32(1.0) - np.exp(-R_t)).astype(F32)
    return sample_cdf(d_vals, opacity, n, det=det, u=u)


def fine_sample(sdf_fn, d_init, rays_o, rays_d, alpha_net, beta_net, far, eps=0.1, max_iter=6,
                max_bisection=10, n_importance=64, n_up=512, det=True, u_final=None):
    """volsdf.fine_sample, 97-302 (VolSDF section 3.4), restated with the same masks.
    sdf_fn(pts [P,3]) -> sdf [P] is `VolSDF.forward_surface`.  rays [M,3], d_init [M,N0].
    Returns d_fine [M,n_importance], beta_map [M,1], iter_usage [M] (float, like the reference)."""
    M, N0 = d_init.shape
    eps = F32(eps)

    def query(d, o, dd):
        pts = (o[:, None, :] + dd[:, None, :] * d[:, :, None]).astype(F32)
        return sdf_fn(pts.reshape(-1, 3)).reshape(d.shape)

    def final(mask, d, s, a, b):
        uu = None if det else u_final[mask]
        return opacity_invert_cdf_sample(d, s, a, b, n_importance, det=det, u=uu)

    d_vals = d_init.astype(F32)
    d_fine = np.zeros((M, n_importance), dtype=F32)
    iter_usage = np.zeros((M,), dtype=F32)
    far = np.full((M, 1), far, dtype=F32) if np.isscalar(far) else far.astype(F32)
    beta = np.sqrt((far ** 2) / F32(4 * (N0 - 1) * np.log(1 + eps))).astype(F32)      # 149
    alpha = (F32(1.0) / beta).astype(F32)
    sdf = query(d_vals, rays_o, rays_d)                                                 # 159
    net_max = error_bound(d_vals, sdf, alpha_net, beta_net).max(axis=-1)                # 162
    mask = net_max > eps
    bounds = error_bound(d_vals, sdf, alpha, beta)                                      # 168
    bounds_masked = bounds[mask]
    converged = np.zeros((M,), dtype=bool)
    if (~mask).sum() > 0:                                                               # 174-177
        d_fine[~mask] = final(~mask, d_vals[~mask], sdf[~mask], alpha_net, beta_net)
        iter_usage[~mask] = 0
    converged[~mask] = True
    cur_N = N0
    it = 0
    while it < max_iter:                                                                # 184
        it += 1
        if mask.sum() == 0:
            break
        up = sample_pdf(d_vals[mask], bounds_masked, n_up + 2, det=True)[:, 1:-1]       # 196
        d_vals = np.concatenate([d_vals, np.zeros((M, n_up), dtype=F32)], axis=-1)      # 211
        sdf = np.concatenate([sdf, np.zeros((M, n_up), dtype=F32)], axis=-1)
        d_m = d_vals[mask]; s_m = sdf[mask]
        d_m[:, cur_N:cur_N + n_up] = up
        order = np.argsort(d_m, axis=-1, kind='stable')                                 # 220 (torch.sort)
        d_m = np.take_along_axis(d_m, order, axis=-1)
        s_m[:, cur_N:cur_N + n_up] = query(up, rays_o[mask], rays_d[mask])              # 222
        s_m = np.take_along_axis(s_m, order, axis=-1)                                   # 226
        d_vals[mask] = d_m; sdf[mask] = s_m
        cur_N += n_up
        net_max[mask] = error_bound(d_vals[mask], sdf[mask], alpha_net, beta_net).max(axis=-1)   # 240
        sub = net_max[mask] > eps
        conv_mask = mask.copy(); conv_mask[mask] = ~sub
        if conv_mask.sum() > 0:                                                         # 248-251
            converged[conv_mask] = True
            d_fine[conv_mask] = final(conv_mask, d_vals[conv_mask], sdf[conv_mask], alpha_net, beta_net)
            iter_usage[conv_mask] = it
        if sub.sum() == 0:
            break
        new_mask = mask.copy(); new_mask[mask] = sub
        b_right = beta[new_mask].copy()                                                 # 260
        b_left = (beta_net * np.ones_like(b_right)).astype(F32)
        d_t = d_vals[new_mask]; s_t = sdf[new_mask]
        for _ in range(max_bisection):                                                  # 266-273
            b_tmp = (F32(0.5) * (b_left + b_right)).astype(F32)
            a_tmp = (F32(1.0) / b_tmp).astype(F32)
            bm = error_bound(d_t, s_t, a_tmp, b_tmp).max(axis=-1)
            ok = bm <= eps
            b_right[ok] = b_tmp[ok]
            b_left[~ok] = b_tmp[~ok]
        beta[new_mask] = b_right
        alpha[new_mask] = (F32(1.0) / beta[new_mask]).astype(F32)
        bounds_masked = np.clip(error_bound(d_t, s_t, alpha[new_mask], beta[new_mask]), F32(0), F32(1e5))  # 280-282
        mask = new_mask
    nc = ~converged
    if nc.sum() > 0:                                                                    # 294-300
        b_plus = beta[nc]; a_plus = (F32(1.0) / b_plus).astype(F32)
        d_fine[nc] = final(nc, d_vals[nc], sdf[nc], a_plus, b_plus)
        iter_usage[nc] = -1
    beta[converged] = beta_net
    return d_fine, beta, iter_usage


def composite_volsdf(d_all, sigma, radiances, nablas=None, white_bkgd=False):
    """Ray integration, volsdf.py:540-576 (the last sample is dropped at 558, +1e-10 at 551/560)."""
    delta = (d_all[:, 1:] - d_all[:, :-1]).astype(F32)
    p = np.exp(-np.maximum(sigma[:, :-1] * delta, F32(0))).astype(F32)
    shifted = np.concatenate([np.ones_like(p[:, :1]), p], axis=-1)
    tau = ((F32(1) - p + F32(1e-10)) * _cumprod(shifted)[:, :-1]).astype(F32)
    rgb = np.sum(tau[:, :, None] * radiances[:, :-1, :], axis=-2, dtype=F32)
    depth = np.sum(tau / (tau.sum(-1, keepdims=True, dtype=F32) + F32(1e-10)) * d_all[:, :-1], axis=-1, dtype=F32)
    acc = np.sum(tau, axis=-1, dtype=F32)
    if white_bkgd:
        rgb = rgb + (F32(1.0) - acc[:, None])
    out = OrderedDict(rgb=rgb.astype(F32), depth_volume=depth.astype(F32), mask_volume=acc.astype(F32))
    if nablas is not None:
        nn_ = _normalize(nablas)
        P = min(tau.shape[-1], nn_.shape[-2])
        out['normals_volume'] = np.sum(nn_[:, :P, :] * tau[:, :P, None], axis=-2, dtype=F32).astype(F32)
    out['p_i'] = p; out['visibility_weights'] = tau
    return out


def volsdf_render(net, rays_o, rays_d, near=0.0, far=6.0, N_samples=128, N_importance=64,
                  max_upsample_steps=6, max_bisection_steps=10, epsilon=0.1, white_bkgd=False,
                  perturb=False, u_final=None, detailed_output=False, rayschunk=2048):
    """volsdf.volume_render, 389-615, with use_view_dirs=True, require_nablas=True, calc_normal=True,
    batched rays flattened to [M,3].  `perturb=True` needs `u_final` [M,N_importance]."""
    rays_o = rays_o.reshape(-1, 3).astype(F32)
    rays_d = _normalize(rays_d.reshape(-1, 3).astype(F32))                             # 442
    M = rays_o.shape[0]
    alpha, beta = net.alpha_beta()
    outs = []
    for c0 in range(0, M, rayschunk):                                                   # 599
        o = rays_o[c0:c0 + rayschunk]; d = rays_d[c0:c0 + rayschunk]
        R = o.shape[0]
        nears = np.full((R, 1), near, dtype=F32); fars = np.full((R, 1), far, dtype=F32)
        t = _linspace(0, 1, N_samples)
        d_coarse = (nears * (F32(1) - t) + fars * t).astype(F32)                        # 472-474
        t4 = _linspace(0, 1, N_samples * 4)
        d_init = (nears * (F32(1) - t4) + fars * t4).astype(F32)                        # 483-484
        uf = None if u_final is None else u_final[c0:c0 + rayschunk]
        d_fine, beta_map, iter_usage = fine_sample(
            lambda p: volsdf_forward_surface(net, p), d_init, o, d, alpha, beta, fars,
            eps=epsilon, max_iter=max_upsample_steps, max_bisection=max_bisection_steps,
            n_importance=N_importance, n_up=N_samples * 4, det=not perturb, u_final=uf)
        d_all = np.sort(np.concatenate([d_coarse, d_fine], axis=-1), axis=-1)           # 501-502
        P = d_all.shape[1]
        pts = (o[:, None, :] + d[:, None, :] * d_all[:, :, None]).astype(F32).reshape(-1, 3)
        vd = np.broadcast_to(d[:, None, :], (R, P, 3)).reshape(-1, 3)
        rad, sdf, nab = volsdf_forward(net, pts, vd)                                    # 510
        rad = rad.reshape(R, P, 3); sdf = sdf.reshape(R, P); nab = nab.reshape(R, P, 3)
        sigma = sdf_to_sigma(sdf, alpha, beta)                                          # 514
        ret = composite_volsdf(d_all, sigma, rad, nab, white_bkgd)
        if detailed_output:                                                             # 578-594
            ret['implicit_surface'] = sdf; ret['implicit_nablas'] = nab; ret['radiance'] = rad
            ret['alpha'] = (F32(1.0) - ret['p_i']).astype(F32)
            ret['d_vals'] = d_all; ret['sigma'] = sigma
            ret['beta_map'] = beta_map; ret['iter_usage'] = iter_usage
        else:
            del ret['p_i'], ret['visibility_weights']
        outs.append(ret)
    return OrderedDict((k, np.concatenate([o_[k] for o_ in outs], axis=0)) for k in outs[0])


# ----------------------------------------------------------------------------------------------
# NeuS  (models/frameworks/neus.py)
# ----------------------------------------------------------------------------------------------
def sdf_to_alpha(sdf, s):
    """neus.sdf_to_alpha, 36-43 with cdf_Phi_s (29-33)."""
    cdf = _sigmoid((sdf * F32(s)).astype(F32))
    a = ((cdf[:, :-1] - cdf[:, 1:]) / (cdf[:, :-1] + F32(1e-10))).astype(F32)
    return cdf, np.maximum(a, F32(0)).astype(F32)


def alpha_to_w(alpha):
    """neus.alpha_to_w, 65-78."""
    shifted = np.concatenate([np.ones_like(alpha[:, :1]), (F32(1.0) - alpha + F32(1e-10)).astype(F32)], axis=-1)
    return (alpha * _cumprod(shifted)[:, :-1]).astype(F32)


def neus_upsample(sdf_fn, d_coarse, rays_o, rays_d, N_importance=64, N_upsample_iters=4, det=True, u=None):
    """'official_solution' upsampling, neus.py:275-303.  sdf_fn = ImplicitSurface.forward (no bg)."""
    def query(d):
        pts = (rays_o[:, None, :] + d[:, :, None] * rays_d[:, None, :]).astype(F32)
        return sdf_fn(pts.reshape(-1, 3)).reshape(d.shape)
    _d = d_coarse.astype(F32)
    _sdf = query(_d)
    n_new = N_importance // N_upsample_iters
    for i in range(N_upsample_iters):
        prev_sdf, next_sdf = _sdf[:, :-1], _sdf[:, 1:]
        prev_z, next_z = _d[:, :-1], _d[:, 1:]
        mid_sdf = ((prev_sdf + next_sdf) * F32(0.5)).astype(F32)
        dot = ((next_sdf - prev_sdf) / (next_z - prev_z + F32(1e-5))).astype(F32)
        prev_dot = np.concatenate([np.zeros_like(dot[:, :1]), dot[:, :-1]], axis=-1)
        dot = np.clip(np.minimum(prev_dot, dot), F32(-10.0), F32(0.0)).astype(F32)
        dist = (next_z - prev_z).astype(F32)
        prev_e = (mid_sdf - dot * dist * F32(0.5)).astype(F32)
        next_e = (mid_sdf + dot * dist * F32(0.5)).astype(F32)
        s = F32(64 * (2 ** i))
        prev_cdf = _sigmoid((prev_e * s).astype(F32)); next_cdf = _sigmoid((next_e * s).astype(F32))
        alpha = ((prev_cdf - next_cdf + F32(1e-5)) / (prev_cdf + F32(1e-5))).astype(F32)
        w = alpha_to_w(alpha)
        uu = None if det else u[i]
        d_new = sample_pdf(_d, w, n_new, det=det, u=uu)
        _d = np.concatenate([_d, d_new], axis=-1)
        _sdf = np.concatenate([_sdf, query(d_new)], axis=-1)
        order = np.argsort(_d, axis=-1, kind='stable')
        _d = np.take_along_axis(_d, order, axis=-1)
        _sdf = np.take_along_axis(_sdf, order, axis=-1)
    return _d


def neus_render(net, rays_o, rays_d, obj_bounding_radius=1.0, N_samples=64, N_importance=64,
                N_upsample_iters=4, white_bkgd=False, perturb=False, u=None, detailed_output=False,
                rayschunk=65536):
    """neus.volume_render, 142-424 (upsample_algo='official_solution', N_outside=0, calc_normal=True)."""
    rays_o = rays_o.reshape(-1, 3).astype(F32)
    rays_d = _normalize(rays_d.reshape(-1, 3).astype(F32))
    outs = []
    for c0 in range(0, rays_o.shape[0], rayschunk):
        o = rays_o[c0:c0 + rayschunk]; d = rays_d[c0:c0 + rayschunk]
        R = o.shape[0]
        near, far = near_far_from_sphere(o, d, r=obj_bounding_radius)
        t = _linspace(0, 1, N_samples)
        d_coarse = (near * (F32(1) - t) + far * t).astype(F32)                          # 235-236
        d_all = neus_upsample(lambda p: sdf_net(net, p)[0], d_coarse, o, d, N_importance, N_upsample_iters,
                              det=not perturb, u=u)
        P = d_all.shape[1]
        pts = (o[:, None, :] + d[:, None, :] * d_all[:, :, None]).astype(F32)
        d_mid = (F32(0.5) * (d_all[:, 1:] + d_all[:, :-1])).astype(F32)                 # 312
        pts_mid = (o[:, None, :] + d[:, None, :] * d_mid[:, :, None]).astype(F32)
        sdf, _, nab = sdf_net(net, pts.reshape(-1, 3), with_nablas=True)                # 320
        sdf = sdf.reshape(R, P); nab = nab.reshape(R, P, 3)
        cdf, alpha = sdf_to_alpha(sdf, net.s())                                         # 322
        pm = pts_mid.reshape(-1, 3)
        _, feat_m, nab_m = sdf_net(net, pm, with_nablas=True)                           # 324 forward_radiance (111-114)
        vd = np.broadcast_to(d[:, None, :], (R, P - 1, 3)).reshape(-1, 3)
        rad = radiance_net(net, pm, vd, nab_m, feat_m).reshape(R, P - 1, 3)
        w = alpha_to_w(alpha)                                                           # 373
        rgb = np.sum(w[:, :, None] * rad, axis=-2, dtype=F32)
        depth = np.sum(w / (w.sum(-1, keepdims=True, dtype=F32) + F32(1e-10)) * d_mid, axis=-1, dtype=F32)
        acc = np.sum(w, axis=-1, dtype=F32)
        if white_bkgd:
            rgb = rgb + (F32(1.0) - acc[:, None])
        nn_ = _normalize(nab)
        Pn = min(w.shape[-1], nn_.shape[-2])
        ret = OrderedDict(rgb=rgb.astype(F32), depth_volume=depth.astype(F32), mask_volume=acc.astype(F32),
                          normals_volume=np.sum(nn_[:, :Pn, :] * w[:, :Pn, None], axis=-2, dtype=F32).astype(F32))
        if detailed_output:                                                             # 397-407
            ret['implicit_nablas'] = nab; ret['implicit_surface'] = sdf; ret['radiance'] = rad
            ret['alpha'] = alpha; ret['cdf'] = cdf; ret['visibility_weights'] = w; ret['d_final'] = d_mid
            ret['d_all'] = d_all
        outs.append(ret)
    return OrderedDict((k, np.concatenate([o_[k] for o_ in outs], axis=0)) for k in outs[0])


# ----------------------------------------------------------------------------------------------
# surface rendering  (models/ray_casting.py)
# ----------------------------------------------------------------------------------------------
def _secant_point(f_low, f_high, d_low, d_high):
    return (-f_low * (d_high - d_low) / (f_high - f_low) + d_low).astype(F32)          # ray_casting.py:16,29


def root_finding_surface_points(query_fn, rays_o, rays_d, near=0.0, far=6.0, N_steps=256, logit_tau=0.0, N_secant_steps=8, fill_inf=True):
    """root_finding_surface_points (ray_casting.py:35-160) + run_secant_method (11-30) on flat rays [N,3] with unit directions.
    query_fn: points [M,3] -> sdf [M] (ImplicitSurface.forward).  -> d_pred_out [N], pt_pred [N,3], mask [N], mask_sign_change [N]."""
    o = rays_o.astype(F32); dd = rays_d.astype(F32)
    N = o.shape[0]
    t = _linspace(0.0, 1.0, N_steps)[None, :]
    near_ = np.full((N, 1), near, F32); far_ = np.full((N, 1), far, F32)
    d = (near_ * (F32(1.0) - t) + far_ * t).astype(F32)                                 # 77
    p = (o[:, None, :] + d[:, :, None] * dd[:, None, :]).astype(F32)                   # 80
    tau = F32(logit_tau)
    val = (query_fn(p.reshape(-1, 3)).reshape(N, N_steps) - tau).astype(F32)            # 87-89
    m0 = val[:, 0] > 0                                                                  # 93
    sign = np.concatenate([np.sign(val[:, :-1] * val[:, 1:]), np.ones((N, 1), F32)], axis=-1).astype(F32)   # 96-101
    cost = sign * np.arange(N_steps, 0, -1).astype(F32)                                 # 104
    idx = np.argmin(cost, axis=-1); values = cost[np.arange(N), idx]                    # 106
    msc = values < 0                                                                    # 109
    mp2n = val[np.arange(N), idx] > 0                                                   # 112
    mask = msc & mp2n & m0                                                              # 114
    idx2 = np.minimum(idx + 1, N_steps - 1)                                             # 126
    d_high = d[np.arange(N), idx][mask]; f_high = val[np.arange(N), idx][mask]
    d_low = d[np.arange(N), idx2][mask]; f_low = val[np.arange(N), idx2][mask]
    om, dm = o[mask], dd[mask]
    if mask.sum() > 0:
        d_pred = _secant_point(f_low, f_high, d_low, d_high)
        for _ in range(N_secant_steps):
            p_mid = (om + d_pred[:, None] * dm).astype(F32)
            f_mid = (query_fn(p_mid) - tau).astype(F32)
            low = f_mid < 0
            d_low = np.where(low, d_pred, d_low); f_low = np.where(low, f_mid, f_low)
            d_high = np.where(~low, d_pred, d_high); f_high = np.where(~low, f_mid, f_high)
            d_pred = _secant_point(f_low, f_high, d_low, d_high)
    else:
        d_pred = np.ones(0, F32)
    pt = np.ones((N, 3), F32)
    pt[mask] = (om + d_pred[:, None] * dm).astype(F32)                                  # 142
    d_out = np.ones(N, F32)
    d_out[mask] = d_pred
    d_out[~mask] = np.inf if fill_inf else F32(far)                                     # 150
    d_out[~m0] = 0                                                                      # 151
    return d_out, pt, mask, msc


def sphere_tracing_surface_points(query_fn, rays_o, rays_d, near=0.0, far=6.0, N_iters=20):
    """sphere_tracing_surface_points (ray_casting.py:163-184)."""
    o = rays_o.astype(F32); dd = rays_d.astype(F32)
    d = (np.ones(o.shape[0], F32) * F32(near)).astype(F32)
    mask = np.ones(o.shape[0], bool)
    for _ in range(N_iters):
        val = query_fn((o + dd * d[:, None]).astype(F32))
        d[mask] = (d[mask] + val[mask]).astype(F32)
        mask[d > far] = False
        mask[d < 0] = False
    return d, (o + dd * d[:, None]).astype(F32), mask


def surface_render(net, rays_o, rays_d, algo, **cfg):
    """surface_render (ray_casting.py:187-263) on flat rays; rays_d un-normalised."""
    dirs = _normalize(rays_d.astype(F32))
    q = lambda x: sdf_net(net, x)[0]
    if algo == 'root_finding':
        d, pt, mask, _ = root_finding_surface_points(q, rays_o, dirs, **cfg)
    elif algo == 'sphere_tracing':
        d, pt, mask = sphere_tracing_surface_points(q, rays_o, dirs, **cfg)
    else:
        raise NotImplementedError
    if net.framework == 'volsdf':
        col, _, nab = volsdf_forward(net, pt, dirs)
    else:
        sdf, feat, nab = sdf_net(net, pt, with_nablas=True)
        col = radiance_net(net, pt, dirs, nab, feat)
    col = col.copy(); col[~mask] = 0
    normals = _normalize(nab); normals[~mask] = 0
    return OrderedDict(rgb=col, depth=d, implicit_nablas=nab, mask_surface=mask, normals_surface=normals)
