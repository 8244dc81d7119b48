"""Import alias: `import nerfart_b200` loads the package that lives in ./nerf-art_b200/ (the contract's directory name has
a hyphen, which Python cannot import by name) *as* `nerfart_b200`, so every submodule has exactly one identity."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'nerf-art_b200')
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, '__init__.py'), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
