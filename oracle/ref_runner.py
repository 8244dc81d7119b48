"""Run the UNMODIFIED reference (cassiePython/NeRF-Art) for timing and parity -- TEST / BENCH INFRASTRUCTURE, never the product.

Only tests/, __graft_entry__.smoke() and bench.py's reference legs (`--impl reference`, `cpu_baseline`, `reference_gpu`) may use this
module.  It imports the reference's own `models.frameworks.volsdf` / `neus` from /root/reference (build container) or from
oracle/_ref/reference (staged by oracle/build_ref.sh; that copy travels to the GPU box, /root/reference does not), with import-time
shims only: stand-in modules for packages the render path never calls (tests/stubs) and `inspect.ArgSpec`
(models/frameworks/volsdf.py:9 imports a name removed in Python 3.11).  No arithmetic of the reference is touched.
"""
import inspect
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CANDIDATES = ['/root/reference', os.path.join(HERE, '_ref', 'reference')]


def reference_root():
    for p in CANDIDATES:
        if os.path.exists(os.path.join(p, 'models', 'frameworks', 'volsdf.py')):
            return p
    return None


_mods = {}


def load():
    """Import the reference's modules once.  Returns a dict of modules, or None when no reference tree is present."""
    if _mods:
        return _mods
    root = reference_root()
    if root is None:
        return None
    if not hasattr(inspect, 'ArgSpec'):
        inspect.ArgSpec = tuple
    stubs = os.path.join(ROOT, 'tests', 'stubs')
    # the reference's top-level package names (models, utils, criteria ...) are generic: import them with the reference root
    # first on sys.path, then restore the path; the stubs only fill packages that are not installed
    saved = list(sys.path)
    sys.path.insert(0, root)
    sys.path.append(stubs)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            from models.frameworks import volsdf as rvolsdf, neus as rneus          # noqa: E402
            from utils import rend_util as rrend                                       # noqa: E402
    finally:
        sys.path[:] = saved
    _mods.update(volsdf=rvolsdf, neus=rneus, rend_util=rrend, root=root)
    return _mods


def build_model(framework, state_dict, kwargs, device='cpu'):
    """The reference's own VolSDF / NeuS module holding `state_dict` (same key layout as the product's mirror)."""
    import torch
    import warnings
    m = load()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        model = (m['volsdf'].VolSDF if framework == 'volsdf' else m['neus'].NeuS)(**kwargs)
    model.load_state_dict({k: v.detach().cpu() for k, v in state_dict.items()}, strict=True)
    return model.to(device).eval()


def volume_render(framework, model, rays_o, rays_d, **kw):
    """The reference's volume_render under no_grad; rays [1,N,3] on the model's device.  Returns (rgb, depth, extras, seconds)."""
    import torch
    m = load()
    fn = m['volsdf'].volume_render if framework == 'volsdf' else m['neus'].volume_render
    dev = rays_o.device
    if dev.type == 'cuda':
        torch.cuda.synchronize(dev)
    t0 = time.time()
    with torch.no_grad():
        rgb, depth, ex = fn(rays_o, rays_d, model, **kw)
    if dev.type == 'cuda':
        torch.cuda.synchronize(dev)
    return rgb, depth, ex, time.time() - t0
