"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the host mirror keeps the reference's
checkpoint layout, and the product refuses to run without CUDA (no silent fallback)."""
import os
import re
import ctypes
import pytest
import torch

from helpers import ROOT, make_volsdf, fx


def test_shared_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from nerfart_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    L = ctypes.CDLL(_lib.LIB_PATH)
    hdr = open(os.path.join(ROOT, 'include', 'nerfart_b200.h')).read()
    names = set(re.findall(r'\b(na_[a-z0-9_]+)\s*\(', hdr))
    assert len(names) >= 15
    for n in sorted(names):
        assert hasattr(L, n), f'{n} declared in include/nerfart_b200.h but not exported'
    L.na_version.restype = ctypes.c_int
    assert L.na_version() >= 100
    L.na_error_string.restype = ctypes.c_char_p
    assert b'workspace' in L.na_error_string(-2)


def test_checkpoint_roundtrip_keeps_reference_layout(tmp_path):
    m = make_volsdf(0.1, 0.5)
    keys = list(m.state_dict().keys())
    assert keys[0] == 'ln_beta' and keys[1] == 'implicit_surface.obj_bounding_size'
    assert 'implicit_surface.surface_fc_layers.3.weight_v' in keys and m.state_dict()['implicit_surface.surface_fc_layers.3.weight_v'].shape == (217, 256)
    assert m.state_dict()['radiance_net.layers.0.weight_v'].shape == (256, 265)
    assert sum(p.numel() for p in m.parameters()) == 796347
    p = tmp_path / 'latest.pt'
    torch.save({'model': m.state_dict(), 'global_step': 3, 'epoch_idx': 0}, p)     # utils/checkpoints.py:42-45 layout
    from helpers import VolSDF
    m2 = VolSDF(**fx.volsdf_kwargs(0.1))
    m2.load_state_dict(torch.load(p)['model'])
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


def test_no_cpu_fallback():
    m = make_volsdf(0.1, 0.0)
    with torch.no_grad(), pytest.raises(RuntimeError):
        m.implicit_surface.forward(torch.zeros(4, 3))
    from nerfart_b200.models.frameworks.volsdf import volume_render
    with torch.no_grad(), pytest.raises(RuntimeError):
        volume_render(torch.zeros(4, 3), torch.ones(4, 3), m, N_samples=8, N_importance=4)


def test_get_model_contract():
    class D(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__
    args = D(model=D(framework='VolSDF', obj_bounding_radius=3.0, max_upsample_iter=6, surface=D(radius_init=1.0, D=8, skips=[4], embed_multires=6),
                     radiance=D(D=4, skips=[], embed_multires=-1, embed_multires_view=-1, use_view_dirs=True), W_geometry_feature=256),
             training=D(speed_factor=10.0, is_finetune=False), data=D(near=0.0, far=6.0, val_rayschunk=1024), device_ids=[0])
    from nerfart_b200.models.frameworks import get_model
    model, trainer, kw_train, kw_test, render_fn = get_model(args, [480, 270])
    assert set(kw_train) == {'near', 'far', 'batched', 'perturb', 'white_bkgd', 'max_upsample_steps', 'use_nerfplusplus', 'obj_bounding_radius'}
    assert kw_test['rayschunk'] == 1024 and kw_test['perturb'] is False and kw_train['perturb'] is True
    assert render_fn is trainer.renderer and hasattr(model, 'implicit_surface') and hasattr(model, 'radiance_net')
    assert args.model.outside_scene == 'builtin' and args.training.beta_init == 0.1   # defaults injected like the reference


def test_mesh_grid_points_follow_the_reference_lattice_and_ply_writer_round_trips(tmp_path):
    """utils/mesh_util.py:87-100 (float divisions included) on a small grid, bit-identical; binary PLY header / payload sizes."""
    import numpy as np
    import torch
    from nerfart_b200.utils import mesh_util as mu
    for N, s in ((7, 2.0), (16, 2.4)):
        idx = np.arange(0, N ** 3, 1).astype(np.int64)
        xyz = np.zeros([N ** 3, 3])
        xyz[:, 2] = idx % N; xyz[:, 1] = (idx / N) % N; xyz[:, 0] = ((idx / N) / N) % N
        o = -s / 2.
        xyz[:, 0] = xyz[:, 0] * (s / (N - 1)) + o; xyz[:, 1] = xyz[:, 1] * (s / (N - 1)) + o; xyz[:, 2] = xyz[:, 2] * (s / (N - 1)) + o
        assert torch.equal(torch.from_numpy(xyz).float(), mu.grid_points(0, N ** 3, N, s, 'cpu'))
        assert torch.equal(torch.from_numpy(xyz[5:40]).float(), mu.grid_points(5, 40, N, s, 'cpu'))       # chunks are consistent
    p = str(tmp_path / 't.ply')
    mu.write_ply(p, np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], float), np.array([[0, 1, 2], [0, 2, 3]]))
    raw = open(p, 'rb').read()
    head, body = raw.split(b'end_header\n')
    assert b'element vertex 4' in head and b'element face 2' in head and b'property list uchar int vertex_indices' in head
    assert len(body) == 4 * 12 + 2 * 13
