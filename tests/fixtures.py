"""Deterministic synthetic fixtures shared by the golden generator (reference side, build
container only) and the tests / smoke / bench (product side, anywhere).

Nothing here touches /root/reference.  Weights come from the model class' own seeded init
(the reference's `geometric_init`, models/base.py:207-224, mirrored by the product's module)
followed by `perturb_state_dict`, a seeded perturbation that turns the trivial sphere SDF
into a bumpy, view-dependent scene so that the error-bound sampler takes 0..6 upsample
iterations and some rays never converge.
"""
import numpy as np
import torch

VOLSDF_SURFACE_CFG = dict(D=8, W=256, skips=[4], embed_multires=6, radius_init=1.0,
                          geometric_init=True, use_siren=False)
VOLSDF_RADIANCE_CFG = dict(D=4, W=256, skips=[], embed_multires=-1, embed_multires_view=-1,
                           use_view_dirs=True, use_siren=False)
NEUS_SURFACE_CFG = dict(D=8, W=256, skips=[4], embed_multires=6, radius_init=0.5,
                        geometric_init=True, use_siren=False)
NEUS_RADIANCE_CFG = dict(D=4, W=256, skips=[], embed_multires=-1, embed_multires_view=4,
                         use_view_dirs=True, use_siren=False)


def volsdf_kwargs(beta_init=0.1):
    """configs/volsdf_fangzhou_nature.yaml:21-44 through volsdf.get_model (volsdf.py:943-975)."""
    return dict(beta_init=beta_init, speed_factor=10.0, W_geo_feat=256, obj_bounding_radius=3.0,
                use_nerfplusplus=False,
                surface_cfg=dict(VOLSDF_SURFACE_CFG), radiance_cfg=dict(VOLSDF_RADIANCE_CFG))


def neus_kwargs(variance_init=0.05):
    """configs/neus_fangzhou_vangogh.yaml:18-44 through neus.get_model (neus.py:693-731)."""
    return dict(variance_init=variance_init, speed_factor=10.0, W_geo_feat=256,
                obj_bounding_radius=1.0, use_outside_nerf=False,
                surface_cfg=dict(NEUS_SURFACE_CFG), radiance_cfg=dict(NEUS_RADIANCE_CFG))


def perturb_state_dict(sd, seed=1234, bump=0.0, radiance_gain=3.0):
    """In-place, seeded.  `bump` scales noise written into the (zero-initialised) positional-
    encoding columns of SDF layer 0 and the skip columns of layer 4, `radiance_gain` multiplies
    every radiance weight_g (sphere init alone renders 0.5 grey)."""
    g = torch.Generator(device='cpu')
    g.manual_seed(seed)
    for k in sorted(sd.keys()):
        v = sd[k]
        if k.startswith('radiance_net') and k.endswith('weight_g'):
            v.mul_(radiance_gain)
        if bump > 0 and k == 'implicit_surface.surface_fc_layers.0.weight_v':
            noise = torch.randn(v.shape, generator=g, dtype=torch.float32)
            scale = v[:, :3].std().item()
            # lower weight on the high octaves so the field stays roughly 1-Lipschitz
            octave = torch.ones(v.shape[1])
            for f in range(6):
                octave[3 + 6 * f: 9 + 6 * f] = 0.5 ** f
            octave[:3] = 0.0
            v.add_(noise * octave[None, :] * (bump * scale))
        if bump > 0 and k == 'implicit_surface.surface_fc_layers.4.weight_v':
            noise = torch.randn(v.shape, generator=g, dtype=torch.float32)
            scale = v[:, :217].std().item()
            mask = torch.zeros(v.shape[1]); mask[217 + 3:] = 1.0
            v.add_(noise * mask[None, :] * (0.3 * bump * scale))
        if bump > 0 and k.endswith('.bias') and k.startswith('radiance_net'):
            v.add_(0.2 * torch.randn(v.shape, generator=g, dtype=torch.float32))
    return sd


def closed_form_camera(H, W):
    """SURVEY.md 8d closed-form fixture: c2w = I with translation (0,0,-2.5), fx=fy=1.25*H."""
    c2w = torch.eye(4, dtype=torch.float32)
    c2w[2, 3] = -2.5
    K = torch.eye(4, dtype=torch.float32)
    K[0, 0] = K[1, 1] = 1.25 * H
    K[0, 2] = W / 2.0
    K[1, 2] = H / 2.0
    return c2w, K


def tilted_camera(H, W):
    """A second, off-axis pose (camera at (1.2,-0.8,-2.2) looking at the origin) so rays are not
    symmetric about the optical axis."""
    eye = np.array([1.2, -0.8, -2.2], dtype=np.float64)
    fwd = -eye / np.linalg.norm(eye)
    up = np.array([0.0, -1.0, 0.0])
    right = np.cross(fwd, up); right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    c2w = np.eye(4)
    c2w[:3, 0] = right; c2w[:3, 1] = down; c2w[:3, 2] = fwd; c2w[:3, 3] = eye
    K = np.eye(4)
    K[0, 0] = K[1, 1] = 1.1 * H
    K[0, 2] = W / 2.0 - 0.5
    K[1, 2] = H / 2.0 + 0.25
    K[0, 1] = 0.01
    return torch.tensor(c2w, dtype=torch.float32), torch.tensor(K, dtype=torch.float32)
