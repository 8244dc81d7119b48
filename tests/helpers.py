"""Shared test helpers: golden loading, seeded models (product classes), oracle nets."""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import fixtures as fx                      # noqa: E402
import nerfart_oracle as orc               # noqa: E402
import nerfart_b200                        # noqa: E402,F401
from nerfart_b200.models.frameworks.volsdf import VolSDF          # noqa: E402


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


def make_volsdf(beta_init, bump, seed=0, device='cpu'):
    """Product-side model with the same seeded state as tests/golden/make_golden.build_volsdf (pinned by state_digest.json)."""
    torch.manual_seed(seed)
    m = VolSDF(**fx.volsdf_kwargs(beta_init))
    sd = m.state_dict(); fx.perturb_state_dict(sd, bump=bump); m.load_state_dict(sd)
    return m.to(device).eval()


def make_neus(variance_init, bump, seed=0, device='cpu'):
    from nerfart_b200.models.frameworks.neus import NeuS
    torch.manual_seed(seed)
    m = NeuS(**fx.neus_kwargs(variance_init))
    sd = m.state_dict(); fx.perturb_state_dict(sd, bump=bump); m.load_state_dict(sd)
    return m.to(device).eval()


def oracle_net(model, framework):
    return orc.Net({k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}, framework)


def linf(a, b):
    if np.asarray(a).size == 0:
        return 0.0
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))
