"""'train.py / render.py run unchanged' -- the import-and-contract half, on CPU (INTEGRATION.md level 1).

A scratch copy of the reference checkout gets the three forwards of scripts/apply_level1.py; a fresh interpreter then imports the
reference's OWN `train`, `render` and `dataio` modules on top of them, loads the shipped YAML configs through the reference's
`io_util.load_config`, builds the model with `get_model`, loads a checkpoint that the UNMODIFIED reference classes saved, and builds
the optimizer / scheduler the way train.py does (train.py:100-158, render.py:257-268, dataio/DTU.py:8).  No kernel is launched.
Needs the reference tree (/root/reference in the build container, or oracle/_ref/reference where oracle/build_ref.sh put it)."""
import os
import shutil
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CANDIDATES = ['/root/reference', os.path.join(ROOT, 'oracle', '_ref', 'reference')]
REF = next((p for p in REF_CANDIDATES if os.path.exists(os.path.join(p, 'train.py'))), None)
pytestmark = pytest.mark.skipif(REF is None, reason='reference tree not present')


def make_checkout(tmp_path):
    dst = str(tmp_path / 'NeRF-Art')
    shutil.copytree(REF, dst, ignore=shutil.ignore_patterns('data', '.git', '*.png', '*.jpg', '*.mp4', '*.gif', '__pycache__'))
    sys.path.insert(0, os.path.join(ROOT, 'scripts'))
    import apply_level1
    apply_level1.apply(dst)
    return dst


def run_py(code, cwd, extra_path=()):
    env = dict(os.environ)
    env['PYTHONPATH'] = os.pathsep.join([cwd, ROOT, os.path.join(ROOT, 'tests', 'stubs'), *extra_path])
    env['CUDA_VISIBLE_DEVICES'] = ''
    r = subprocess.run([sys.executable, '-c', textwrap.dedent(code)], cwd=cwd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + '\n' + r.stderr[-3000:]
    return r.stdout


def test_level1_forwards_are_the_documented_ones():
    sys.path.insert(0, os.path.join(ROOT, 'scripts'))
    import apply_level1
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    for rel in apply_level1.FORWARDS:
        assert rel in doc, f'INTEGRATION.md does not mention {rel}'
    assert 'scripts/apply_level1.py' in doc


@pytest.mark.parametrize('framework,config', [('volsdf', 'configs/volsdf_fangzhou_vangogh.yaml'), ('neus', 'configs/neus_fangzhou_vangogh.yaml')])
def test_reference_entry_points_import_and_build_on_the_mirror(tmp_path, framework, config):
    co = make_checkout(tmp_path)
    # 1. a checkpoint written by the UNMODIFIED reference classes (reference layout: {'model': state_dict, ...}, checkpoints.py:42-45)
    ck = str(tmp_path / 'ref_ckpt.pt')
    run_py(f'''
        import sys, inspect, torch
        if not hasattr(inspect, 'ArgSpec'): inspect.ArgSpec = tuple          # models/frameworks/volsdf.py:9 (removed in py3.11)
        from utils import io_util
        import argparse
        args = io_util.load_config(argparse.Namespace(config={config!r}, resume_dir=None), [])
        if {framework!r} == 'volsdf':
            from models.frameworks.volsdf import get_model
        else:
            from models.frameworks.neus import get_model
        args.training.is_finetune = False                                      # no CLIP / VGG downloads for this fixture
        args.device_ids = [0]
        torch.manual_seed(3)
        model = get_model(args, [480, 270])[0]
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        torch.save({{'model': model.state_dict(), 'optimizer': opt.state_dict(), 'global_step': 7, 'epoch_idx': 1}}, {ck!r})
        print(len(model.state_dict()))
    ''', cwd=REF)
    # 2. the reference's own entry modules on top of the forwards
    out = run_py(f'''
        import argparse, torch, collections
        import train, render, dataio                                             # the reference's files, unchanged
        from dataio import DTU, custom                                           # dataio/DTU.py:8 imports rot_to_quat, load_K_Rt_from_P
        import models.frameworks, models.base, utils.rend_util
        assert train.get_model.__module__.startswith('nerfart_b200'), train.get_model.__module__
        assert render.get_model is train.get_model
        assert train.rend_util.get_rays.__module__.startswith('nerfart_b200')
        from utils import io_util
        args = io_util.load_config(argparse.Namespace(config={config!r}, resume_dir=None, ddp=False), [])
        args.training.is_finetune = True                                         # the shipped *_vangogh configs: fine-tune branch
        args.device_ids = [0]
        model, trainer, kw_train, kw_test, render_fn = train.get_model(args, [480, 270])
        assert type(model).__module__.startswith('nerfart_b200')
        state = torch.load({ck!r}, map_location='cpu')
        missing = model.load_state_dict(state['model'], strict=True)             # render.py:266-268
        assert not missing.missing_keys and not missing.unexpected_keys
        # train.py:114,158
        opt = train.get_optimizer(args, model)
        opt.load_state_dict(state['optimizer'])                                  # CheckpointIO.load restores it the same way
        sched = train.get_scheduler(args, opt, last_epoch=6)
        assert hasattr(trainer, 'val') == ({framework!r} == 'volsdf')            # train.py:205 probes with hasattr
        assert callable(render_fn) and kw_test['perturb'] is False and 'rayschunk' in kw_test
        assert hasattr(model, 'implicit_surface') and hasattr(model.implicit_surface, 'pretrain_hook')     # train.py:150
        n = sum(p.numel() for p in model.parameters())
        print('OK', n, sorted(kw_train)[:3], type(sched).__name__)
    ''', cwd=co)
    assert 'OK' in out
