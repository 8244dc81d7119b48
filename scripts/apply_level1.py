#!/usr/bin/env python
"""INTEGRATION.md level 1, as a script: turn four modules of a NeRF-Art checkout into one-line forwards to nerfart_b200.

    python scripts/apply_level1.py /path/to/NeRF-Art

After this, `train.py` / `render.py` of that checkout run unchanged on the B200 kernels (put this repository on PYTHONPATH).
The original files are kept beside the forwards as `*.py.orig`.  tests/test_dropin_*.py apply exactly this to a scratch copy.
"""
import os
import shutil
import sys

FORWARDS = {
    'models/frameworks/__init__.py':
        '# forwarded to nerfart_b200 (reference: models/frameworks/__init__.py:1-11)\n'
        'from nerfart_b200.models.frameworks import get_model            # noqa: F401\n',
    'models/base.py':
        '# forwarded to nerfart_b200 (reference: get_optimizer 486-521 / get_scheduler 547-584 are imported by train.py)\n'
        'from nerfart_b200.models.base import *                           # noqa: F401,F403\n'
        'from nerfart_b200.models.base import get_optimizer, get_scheduler, ImplicitSurface, RadianceNet, get_embedder   # noqa: F401\n',
    'utils/rend_util.py':
        '# forwarded to nerfart_b200 (reference: get_rays 112-165, lin2img 238-248 are used by render.py / train.py;\n'
        '# rot_to_quat / load_K_Rt_from_P by dataio/)\n'
        'from nerfart_b200.utils.rend_util import *                       # noqa: F401,F403\n',
    'utils/mesh_util.py':
        '# forwarded to nerfart_b200 (reference: extract_mesh 82-112 is called by train.py:213-222 and tools/extract_surface.py)\n'
        'from nerfart_b200.utils.mesh_util import *                       # noqa: F401,F403\n',
}


def apply(checkout):
    for rel, text in FORWARDS.items():
        path = os.path.join(checkout, rel)
        if not os.path.exists(path):
            raise FileNotFoundError(f'{path}: not a NeRF-Art checkout?')
        if not os.path.exists(path + '.orig'):
            shutil.copy(path, path + '.orig')
        with open(path, 'w') as f:
            f.write(text)
    return sorted(FORWARDS)


if __name__ == '__main__':
    if len(sys.argv) != 2:
        sys.exit(__doc__)
    for rel in apply(sys.argv[1]):
        print('forwarded', rel)
