"""CPU oracle (numpy) for the BACKWARD of the NeRF-Art volumetric render: what PyTorch autograd computes in the second
pass of the fine-tune step, `Trainer.forward` (models/frameworks/volsdf.py:769-783, models/frameworks/neus.py:551-563):

    rgb_pred.backward(gradient_patch);  (w_eikonal * mse(||implicit_nablas||, 1)).backward()

TEST INFRASTRUCTURE ONLY (same rules as nerfart_oracle.py): imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product.

The reference has no hand-written backward: autograd differentiates the forward graph, including the graph of
`autograd.grad(sdf, x, create_graph=True)` (models/base.py:265-282), i.e. a second-order path through the SDF network.
This file restates that derivative in closed form (DESIGN.md section 9 has the derivation); it is pinned by
tests/golden/train_*.npz, which hold the parameter gradients the UNMODIFIED reference produces on seeded state
(tests/golden/make_golden_train.py).  Arithmetic is float64 by default (checker accuracy; inputs are the fp32 parameters),
float32 on request (used when this port is timed as the CPU baseline).

Notation (per sample):  in_i input of SDF layer i, z_i = W_i in_i + b_i, h_i = softplus_100(z_i), s_i = softplus'(z_i);
reverse sweep u_7 = W_8[0,:], g_i = u_i * s_i, v_i = W_i^T g_i (= d sdf / d in_i), u_{i-1} = v_i; the skip layer consumes
cat(h_3, emb)/sqrt2; ge = d sdf / d emb; nabla = J^T ge with J = d emb / d x.
"""
from collections import OrderedDict
import numpy as np

import nerfart_oracle as orc

SQRT2 = np.sqrt(2.0)


def _embed_and_jac(x, multires, dt):
    """emb [M,3+6L] and the diagonal Jacobian entries d emb_k / d x_{c(k)} [M,3+6L] (each embedding entry depends on
    exactly one coordinate, models/base.py:46-64)."""
    x = x.astype(dt)
    out, jac = [x], [np.ones_like(x)]
    for k in range(multires):
        f = dt(2.0 ** k)
        out += [np.sin(x * f), np.cos(x * f)]
        jac += [f * np.cos(x * f), -f * np.sin(x * f)]
    return np.concatenate(out, -1), np.concatenate(jac, -1)


def _fold_coords(v, multires):
    """sum the 3+6L embedding-gradient columns onto their source coordinate: [M,3+6L] -> [M,3]."""
    out = v[:, 0:3].copy()
    for k in range(2 * multires):
        out += v[:, 3 + 3 * k: 6 + 3 * k]
    return out


def _softplus_parts(z, dt):
    """h, s = softplus', c = softplus'' of nn.Softplus(beta=100, threshold=20) as torch differentiates it
    (softplus_backward: z/(z+1) with z = exp(100 x) below the threshold, 1 above; its derivative 100 s (1-s) / 0)."""
    bz = z * dt(100.0)
    lin = bz > dt(20.0)
    ez = np.exp(np.minimum(bz, dt(30.0)))
    h = np.where(lin, z, np.log1p(ez) / dt(100.0))
    s = np.where(lin, dt(1.0), ez / (ez + dt(1.0)))
    c = np.where(lin, dt(0.0), dt(100.0) * s * (dt(1.0) - s))
    return h, s, c


def weight_norm_backward(g, v, dW):
    """nn.utils.weight_norm(dim=0): W[o,:] = g[o] v[o,:] / ||v[o,:]||  ->  (dg [out,1], dv [out,in])."""
    dt = dW.dtype
    g = g.reshape(-1, 1).astype(dt); v = v.astype(dt)
    n = np.sqrt(np.sum(v * v, axis=1, keepdims=True))
    dg = np.sum(dW * v, axis=1, keepdims=True) / n
    dv = (g / n) * (dW - dg * v / n)
    return dg, dv


class TrainNet:
    """Raw (weight_g, weight_v, bias) parameters + effective weights, from a reference-layout state dict."""

    def __init__(self, sd, framework, dtype=np.float64):
        self.dt = dt = dtype
        self.framework = framework
        self.sd = {k: np.asarray(v, dtype=np.float32) for k, v in sd.items()}
        self.base = orc.Net(sd, framework)

        def eff(prefix):
            g = self.sd[prefix + 'weight_g'].astype(dt); v = self.sd[prefix + 'weight_v'].astype(dt)
            return v * (g.reshape(-1, 1) / np.sqrt(np.sum(v * v, axis=1, keepdims=True)))
        self.sW = [eff(f'implicit_surface.surface_fc_layers.{i}.') for i in range(9)]
        self.sb = [self.sd[f'implicit_surface.surface_fc_layers.{i}.bias'].astype(dt) for i in range(9)]
        self.rW = [eff(f'radiance_net.layers.{i}.') for i in range(5)]
        self.rb = [self.sd[f'radiance_net.layers.{i}.bias'].astype(dt) for i in range(5)]
        self.multires = 6
        self.multires_view = self.base.multires_view
        self.bound = dt(self.base.bound)
        self.speed = dt(self.base.speed)


def mlp_backward(net: TrainNet, x, view, g_sdf, g_nab, g_rad, apply_bg, with_radiance=True, keep=None):
    """Backward of (sdf, nabla, radiance)(x, view) w.r.t. the effective weights.
    x, view [M,3]; upstream gradients g_sdf [M], g_nab [M,3], g_rad [M,3] (None = zero).
    Returns dict dW_s[0..8] ([out,in], reference layout), db_s, dW_r[0..4], db_r (radiance entries only if with_radiance).
    `keep`: optional dict that receives the per-sample intermediates the CUDA kernel stashes (for stage-wise tests)."""
    dt = net.dt
    M = x.shape[0]
    x = x.astype(dt)
    emb, jac = _embed_and_jac(x, net.multires, dt)
    E = emb.shape[1]
    NH = net.sW[4].shape[1] - E                                   # 217
    # ---- forward (models/base.py:243-263) -------------------------------------------------------
    ins, ss, cs = [], [], []
    h = emb
    for i in range(8):
        if i == 4:
            h = np.concatenate([h, emb], -1) / dt(SQRT2)
        ins.append(h)
        z = h @ net.sW[i].T + net.sb[i]
        h, s, c = _softplus_parts(z, dt)
        ss.append(s); cs.append(c)
    h7 = h
    out = h7 @ net.sW[8].T + net.sb[8]
    sdf_raw, feat = out[:, 0], out[:, 1:]
    # ---- reverse sweep = autograd.grad(sdf, x) (base.py:271-277) -------------------------------
    us, gs = [None] * 8, [None] * 8
    u = np.broadcast_to(net.sW[8][0], (M, 256)).astype(dt)
    ge = np.zeros_like(emb)
    for i in range(7, -1, -1):
        us[i] = u
        gs[i] = u * ss[i]
        v = gs[i] @ net.sW[i]
        if i == 4:
            v = v / dt(SQRT2)
            ge = ge + v[:, NH:]
            v = v[:, :NH]
        u = v
    ge = ge + u
    nab = _fold_coords(ge * jac, net.multires)
    # ---- radiance net (base.py:372-391) ----------------------------------------------------------
    g_sdf = np.zeros(M, dt) if g_sdf is None else g_sdf.astype(dt)
    n_bar = np.zeros((M, 3), dt) if g_nab is None else g_nab.astype(dt).copy()
    feat_bar = np.zeros((M, 256), dt)
    res = OrderedDict(dW_s=[None] * 9, db_s=[None] * 9, dW_r=[None] * 5, db_r=[None] * 5)
    if with_radiance and g_rad is not None:
        vemb = orc.embed(view.astype(np.float32), net.multires_view).astype(dt)
        r_in = np.concatenate([x, vemb, nab, feat], -1)
        ys = [r_in]
        y = r_in
        for j in range(4):
            y = np.maximum(y @ net.rW[j].T + net.rb[j], 0)
            ys.append(y)
        rgb = 1.0 / (1.0 + np.exp(-(y @ net.rW[4].T + net.rb[4])))
        d = g_rad.astype(dt) * rgb * (1 - rgb)
        for j in range(4, -1, -1):
            res['dW_r'][j] = d.T @ ys[j]
            res['db_r'][j] = d.sum(0)
            d = d @ net.rW[j]
            if j > 0:
                d = d * (ys[j] > 0)
        nv = vemb.shape[1]
        n_bar += d[:, 3 + nv: 6 + nv]
        feat_bar = d[:, 6 + nv:]
        if keep is not None:
            keep['rgb'] = rgb
    # ---- SDF head -------------------------------------------------------------------------------
    if apply_bg:                                                   # volsdf.py:349-357: overridden samples pass no sdf gradient
        d_bg = net.bound - np.sqrt(np.sum(x * x, -1))
        g_sdf = np.where(d_bg < sdf_raw, 0.0, g_sdf)
    out_bar = np.concatenate([g_sdf[:, None], feat_bar], -1)
    dW8 = out_bar.T @ h7
    res['db_s'][8] = out_bar.sum(0)
    h_bar = out_bar @ net.sW[8]
    # ---- second-order path: gradient of nabla w.r.t. the weights (forward-like sweep, layers 0..7) ----------------
    ge_bar = np.concatenate([n_bar] * (1 + 2 * net.multires), -1) * jac            # d nabla / d ge = J
    s_bars = [None] * 8
    dW2 = [None] * 8
    v_bar = ge_bar                                                 # gradient w.r.t. v_0 (39)
    v_bars = [None] * 8
    for i in range(8):
        if i == 4:
            v_bar = np.concatenate([v_bar, ge_bar], -1) / dt(SQRT2)    # v_4 feeds u_3 (first 217) and ge (last 39), both /sqrt2
        v_bars[i] = v_bar
        g_bar = v_bar @ net.sW[i].T                                # v_i = W_i^T g_i
        dW2[i] = gs[i].T @ v_bar
        s_bars[i] = g_bar * us[i]
        v_bar = g_bar * ss[i]                                      # = u_bar_i, which is v_bar_{i+1}
    dW8[0] += v_bar.sum(0)                                         # u_7 = W_8[0,:]
    res['dW_s'][8] = dW8
    # ---- first-order backward through the trunk, layers 7..0 ------------------------------------------------
    z_bars = [None] * 8
    for i in range(7, -1, -1):
        z_bar = h_bar * ss[i] + s_bars[i] * cs[i]
        z_bars[i] = z_bar
        res['dW_s'][i] = z_bar.T @ ins[i] + dW2[i]
        res['db_s'][i] = z_bar.sum(0)
        in_bar = z_bar @ net.sW[i]
        if i == 4:
            in_bar = in_bar[:, :NH] / dt(SQRT2)
        h_bar = in_bar
    if keep is not None:
        keep.update(sdf_raw=sdf_raw, feat=feat, nab=nab, ins=ins, ss=ss, us=us, gs=gs, v_bars=v_bars, z_bars=z_bars,
                    feat_bar=feat_bar, n_bar=n_bar, g_sdf=g_sdf, ge_bar=ge_bar)
    return res


def _accumulate(total, part):
    for k in total:
        for i in range(len(total[k])):
            if part[k][i] is not None:
                total[k][i] = part[k][i] if total[k][i] is None else total[k][i] + part[k][i]


def param_grads(net: TrainNet, eff, extra):
    """effective-weight gradients -> gradients of the reference's parameters (state-dict keys)."""
    out = OrderedDict()
    for name, dWs, dbs, n in (('implicit_surface.surface_fc_layers', eff['dW_s'], eff['db_s'], 9),
                              ('radiance_net.layers', eff['dW_r'], eff['db_r'], 5)):
        for i in range(n):
            if dWs[i] is None:
                continue
            p = f'{name}.{i}.'
            dg, dv = weight_norm_backward(net.sd[p + 'weight_g'], net.sd[p + 'weight_v'], dWs[i])
            out[p + 'weight_g'] = dg; out[p + 'weight_v'] = dv; out[p + 'bias'] = dbs[i]
    out.update(extra)
    return out


def eikonal_grad(nab, w_eik, count):
    """d/d nabla of  w * mean((||nabla|| - 1)^2)  over `count` points (calc_eikonal_loss, volsdf.py:917-939) and the loss."""
    nn_ = np.sqrt(np.sum(nab * nab, -1))
    loss = w_eik * np.sum((nn_ - 1.0) ** 2) / count
    g = (w_eik * 2.0 / count) * ((nn_ - 1.0) / np.maximum(nn_, 1e-30))[:, None] * nab
    return g, loss


def volsdf_backward(net: TrainNet, rays_o, rays_d, d_all, G, w_eik=0.0, white_bkgd=False, keep=None):
    """Gradients of  sum(rgb * G) + w_eik * mean((||nabla||-1)^2)  for VolSDF's volume_render at fixed sample depths
    d_all [R,P] (the sampler runs under no_grad, volsdf.py:113).  Returns (param-grad dict, eikonal loss, rgb)."""
    dt = net.dt
    o = rays_o.reshape(-1, 3).astype(dt)
    d = rays_d.reshape(-1, 3).astype(dt)
    d = d / np.maximum(np.sqrt(np.sum(d * d, -1, keepdims=True)), 1e-12)             # volsdf.py:442
    R, P = d_all.shape
    t = d_all.astype(dt)
    x = (o[:, None, :] + d[:, None, :] * t[:, :, None]).reshape(-1, 3)
    view = np.broadcast_to(d[:, None, :], (R, P, 3)).reshape(-1, 3)
    # forward values needed by the compositing backward (float64 re-evaluation of the reference forward)
    k0 = {}
    mlp_backward(net, x, view, None, None, np.zeros((R * P, 3), dt), apply_bg=True, keep=k0)
    sdf = k0['sdf_raw']
    d_bg = net.bound - np.sqrt(np.sum(x * x, -1))
    sdf = np.where(d_bg < sdf, d_bg, sdf).reshape(R, P)
    rad = k0['rgb'].reshape(R, P, 3)
    nab = k0['nab']
    ln_beta = dt(net.sd['ln_beta'].reshape(-1)[0])
    beta = np.exp(ln_beta * net.speed); alpha = 1.0 / beta
    e = 0.5 * np.exp(-np.abs(sdf) / beta)
    psi = np.where(sdf >= 0, e, 1 - e)
    sigma = alpha * psi                                                               # volsdf.py:34-53
    delta = t[:, 1:] - t[:, :-1]
    xx = sigma[:, :-1] * delta
    p = np.exp(-np.maximum(xx, 0))
    T = np.cumprod(np.concatenate([np.ones((R, 1), dt), p], -1), -1)                  # T[:, i] = prod_{j<i} p_j, i = 0..P-1
    tau = (1 - p + 1e-10) * T[:, :-1]
    rgb = np.sum(tau[:, :, None] * rad[:, :-1], -2)
    if white_bkgd:
        rgb = rgb + (1 - tau.sum(-1))[:, None]
    G = G.reshape(R, 3).astype(dt)
    c_bar = np.zeros((R, P, 3), dt)
    c_bar[:, :-1] = tau[:, :, None] * G[:, None, :]
    tau_bar = np.sum(rad[:, :-1] * G[:, None, :], -1) - (G.sum(-1, keepdims=True) if white_bkgd else 0.0)
    tt = tau_bar * tau
    S = np.concatenate([np.cumsum(tt[:, ::-1], -1)[:, ::-1][:, 1:], np.zeros((R, 1), dt)], -1)   # S_i = sum_{k>i} tau_bar_k tau_k
    x_bar = (tau_bar * T[:, 1:] - S) * (xx > 0)
    sigma_bar = np.zeros((R, P), dt)
    sigma_bar[:, :-1] = x_bar * delta
    s_bar = sigma_bar * (-(alpha / beta) * e) * (sdf != 0)
    dpsi_dbeta = np.where(sdf >= 0, 1.0, -1.0) * e * np.abs(sdf) / beta ** 2
    beta_bar = np.sum(sigma_bar * (psi * (-1.0 / beta ** 2) + alpha * dpsi_dbeta))
    ln_beta_bar = beta_bar * net.speed * beta
    g_nab, eik = (eikonal_grad(nab, w_eik, R * P) if w_eik else (None, 0.0))
    eff = mlp_backward(net, x, view, s_bar.reshape(-1), g_nab, c_bar.reshape(-1, 3), apply_bg=True, keep=keep)
    if keep is not None:
        keep.update(s_bar=s_bar, c_bar=c_bar, tau=tau, sigma=sigma)
    grads = param_grads(net, eff, {'ln_beta': np.array([ln_beta_bar])})
    return grads, eik, rgb


def neus_backward(net: TrainNet, rays_o, rays_d, d_all, G, w_eik=0.0, white_bkgd=False, train_radiance=False):
    """Same for NeuS' volume_render (neus.py:305-395) at fixed d_all [R,P]: alpha from the SDF at d_all, radiance at the
    P-1 midpoints through a second evaluation of the SDF network (forward_radiance, neus.py:111-114)."""
    dt = net.dt
    o = rays_o.reshape(-1, 3).astype(dt)
    d = rays_d.reshape(-1, 3).astype(dt)
    d = d / np.maximum(np.sqrt(np.sum(d * d, -1, keepdims=True)), 1e-12)
    R, P = d_all.shape
    t = d_all.astype(dt)
    t_mid = 0.5 * (t[:, 1:] + t[:, :-1])
    x = (o[:, None, :] + d[:, None, :] * t[:, :, None]).reshape(-1, 3)
    xm = (o[:, None, :] + d[:, None, :] * t_mid[:, :, None]).reshape(-1, 3)
    view_m = np.broadcast_to(d[:, None, :], (R, P - 1, 3)).reshape(-1, 3)
    kA, kB = {}, {}
    mlp_backward(net, x, None, None, None, None, apply_bg=False, with_radiance=False, keep=kA)
    mlp_backward(net, xm, view_m, None, None, np.zeros((R * (P - 1), 3), dt), apply_bg=False, keep=kB)
    sdf = kA['sdf_raw'].reshape(R, P)
    nab = kA['nab']
    rad = kB['rgb'].reshape(R, P - 1, 3)
    ln_s = dt(net.sd['ln_s'].reshape(-1)[0])
    s = np.exp(ln_s * net.speed)
    cdf = 1.0 / (1.0 + np.exp(-sdf * s))
    raw = (cdf[:, :-1] - cdf[:, 1:]) / (cdf[:, :-1] + 1e-10)
    a = np.maximum(raw, 0)
    q = 1 - a + 1e-10
    T = np.cumprod(np.concatenate([np.ones((R, 1), dt), q], -1), -1)[:, :-1]
    w = a * T
    rgb = np.sum(w[:, :, None] * rad, -2)
    if white_bkgd:
        rgb = rgb + (1 - w.sum(-1))[:, None]
    G = G.reshape(R, 3).astype(dt)
    c_bar = w[:, :, None] * G[:, None, :]
    w_bar = np.sum(rad * G[:, None, :], -1) - (G.sum(-1, keepdims=True) if white_bkgd else 0.0)
    ww = w_bar * w
    S = np.concatenate([np.cumsum(ww[:, ::-1], -1)[:, ::-1][:, 1:], np.zeros((R, 1), dt)], -1)
    a_bar = (w_bar * T - S / q) * (raw >= 0)
    cdf_bar = np.zeros((R, P), dt)
    cdf_bar[:, :-1] += a_bar * (cdf[:, 1:] + 1e-10) / (cdf[:, :-1] + 1e-10) ** 2
    cdf_bar[:, 1:] += -a_bar / (cdf[:, :-1] + 1e-10)
    pre_bar = cdf_bar * cdf * (1 - cdf)
    sdf_bar = pre_bar * s
    s_bar = np.sum(pre_bar * sdf)
    ln_s_bar = s_bar * net.speed * s
    g_nab, eik = (eikonal_grad(nab, w_eik, R * P) if w_eik else (None, 0.0))
    effA = mlp_backward(net, x, None, sdf_bar.reshape(-1), g_nab, None, apply_bg=False, with_radiance=False)
    effB = mlp_backward(net, xm, view_m, None, None, c_bar.reshape(-1, 3), apply_bg=False)
    _accumulate(effA, effB)
    if not train_radiance:                                                            # neus.py:28 FIX_MODULE = "radiance_net"
        effA['dW_r'] = [None] * 5
    grads = param_grads(net, effA, {'ln_s': np.array([ln_s_bar])})
    return grads, eik, rgb
