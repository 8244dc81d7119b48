"""Import shim for the UNMODIFIED reference (cassiePython/NeRF-Art) on CPU.

Only used by tests/golden/make_golden.py, in the build container where
/root/reference is mounted.  Nothing under tests/ (-m gpu), bench.py or
__graft_entry__.smoke() imports this file: /root/reference does not exist on the GPU box.

The shims touch import-time names only, never arithmetic:
  * stub modules for packages the reference imports but the render path never calls
    (addict, imageio, skimage, matplotlib, plyfile, clip)
  * inspect.ArgSpec, removed in Python 3.11 but imported by models/frameworks/volsdf.py:9
"""
import sys, types, inspect

REF_ROOT = '/root/reference'


def install():
    for name in ['addict', 'imageio', 'skimage', 'skimage.transform', 'skimage.measure',
                 'matplotlib', 'matplotlib.pyplot', 'clip', 'plyfile']:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)

    class _D(dict):
        __getattr__ = lambda s, k: s[k]
        __setattr__ = lambda s, k, v: s.__setitem__(k, v)
    sys.modules['addict'].Dict = _D
    sys.modules['skimage.transform'].rescale = None
    sys.modules['skimage'].transform = sys.modules['skimage.transform']
    sys.modules['skimage'].measure = sys.modules['skimage.measure']
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    if not hasattr(inspect, 'ArgSpec'):
        inspect.ArgSpec = tuple
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
