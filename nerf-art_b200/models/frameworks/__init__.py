"""`get_model(args, target_hw)` dispatcher -- reference models/frameworks/__init__.py:1-11."""


def get_model(args, target_hw=None):
    fw = args.model.framework
    if fw == 'VolSDF':
        from .volsdf import get_model as _gm
    elif fw == 'NeuS':
        from .neus import get_model as _gm
    else:
        raise NotImplementedError(fw)          # the reference raises for UNISURF too (frameworks/__init__.py:2-4)
    return _gm(args, target_hw)
