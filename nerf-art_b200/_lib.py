"""ctypes binding of libnerfart_b200.so (the C ABI in include/nerfart_b200.h).

PyTorch is plumbing here: it owns device memory and streams; every tensor crosses the boundary as a raw
device pointer.  There is NO fallback: if the shared library is missing or a call fails, a RuntimeError
is raised.
"""
import ctypes as C
import os
import subprocess
import sys
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('NA_LIB_PATH') or os.path.join(_HERE, 'libnerfart_b200.so')    # NA_LIB_PATH: diagnostic builds (scripts/build_trace.sh)
CSRC = os.path.join(_HERE, 'csrc')
SOURCES = ['api.cu', 'mlp_simt.cu', 'volsdf_render.cu', 'neus_render.cu', 'surface_render.cu', 'mlp_tc.cu', 'mlp_tmem.cu', 'train.cu', 'clip_vit.cu', 'tgemm.cu', 'wgrad_f16.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--compiler-options', '-fPIC']
LINK_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '--compiler-options', '-fPIC']

NA_FRAMEWORK_VOLSDF, NA_FRAMEWORK_NEUS = 0, 1
NA_PRECISION_FP32, NA_PRECISION_TC, NA_PRECISION_TC2ACC, NA_PRECISION_TC_MIXED = 0, 1, 2, 3
PRECISIONS = {'fp32': NA_PRECISION_FP32, 'tc': NA_PRECISION_TC, 'tc2acc': NA_PRECISION_TC2ACC, 'tc_mixed': NA_PRECISION_TC_MIXED}


class NaNetDesc(C.Structure):
    _fields_ = [('framework', C.c_int32), ('multires_view', C.c_int32), ('bounding_radius', C.c_float), ('reserved', C.c_float)]


class NaRawParams(C.Structure):
    _fields_ = [('bias', C.c_void_p * 14), ('weight_g', C.c_void_p * 14), ('weight_v', C.c_void_p * 14)]


class NaVolsdfCfg(C.Structure):
    _fields_ = [('n_samples', C.c_int32), ('n_importance', C.c_int32), ('max_upsample_steps', C.c_int32),
                ('max_bisection_steps', C.c_int32), ('near', C.c_float), ('far', C.c_float), ('epsilon', C.c_float),
                ('white_bkgd', C.c_int32), ('perturb', C.c_int32), ('precision', C.c_int32), ('detailed', C.c_int32),
                ('reserved', C.c_int32)]


class NaVolsdfOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ('rgb', 'depth', 'acc', 'normals', 'beta_map', 'iter_usage',
                                          'd_vals', 'sdf', 'nablas', 'radiance', 'sigma', 'tau')]


class NaNeusCfg(C.Structure):
    _fields_ = [('n_samples', C.c_int32), ('n_importance', C.c_int32), ('n_upsample_iters', C.c_int32),
                ('bounding_radius', C.c_float), ('white_bkgd', C.c_int32), ('perturb', C.c_int32),
                ('precision', C.c_int32), ('detailed', C.c_int32)]


class NaNeusOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ('rgb', 'depth', 'acc', 'normals', 'd_all', 'sdf', 'nablas', 'radiance',
                                          'alpha', 'weights')]


NA_RAYCAST_ROOT_FINDING, NA_RAYCAST_SPHERE_TRACING = 0, 1


class NaSurfaceCfg(C.Structure):
    _fields_ = [('algo', C.c_int32), ('n_steps', C.c_int32), ('n_secant_steps', C.c_int32), ('n_iters', C.c_int32),
                ('near', C.c_float), ('far', C.c_float), ('logit_tau', C.c_float), ('fill_inf', C.c_int32),
                ('use_view_dirs', C.c_int32), ('precision', C.c_int32)]


class NaSurfaceOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ('rgb', 'depth', 'mask', 'nablas', 'normals')]


class NaTrainCfg(C.Structure):
    _fields_ = [('points_per_ray', C.c_int32), ('w_eikonal', C.c_float), ('eikonal_count', C.c_int32), ('white_bkgd', C.c_int32),
                ('speed_factor', C.c_float), ('train_surface', C.c_int32), ('train_radiance', C.c_int32), ('precision', C.c_int32)]


class NaRawGrads(C.Structure):
    _fields_ = [('bias', C.c_void_p * 14), ('weight_g', C.c_void_p * 14), ('weight_v', C.c_void_p * 14)]


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into libnerfart_b200.so (nvcc cross-compiles without a GPU).  One object per source,
    compiled in parallel and rebuilt only when the source or a header changed; objects live in nerf-art_b200/build/ (git-ignored)."""
    from concurrent.futures import ThreadPoolExecutor
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')] + \
        [os.path.join(_HERE, '..', 'include', 'nerfart_b200.h')]
    hdr_time = max(os.path.getmtime(h) for h in hdrs)
    nvcc = os.environ.get('NVCC', 'nvcc')
    extra = os.environ.get('NA_NVCC_EXTRA', '').split()
    objdir = os.path.join(_HERE, 'build' + ('_' + '_'.join(e.strip('-').replace('=', '_') for e in extra) if extra else ''))
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        if os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_time):
            return obj, False
        cmd = [nvcc] + NVCC_FLAGS + extra + ['-c', '-o', obj, src]
        if verbose:
            print(' '.join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True, cwd=CSRC)
        return obj, True
    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        res = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in res]
    if any(c for _, c in res) or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(o) for o in objs):
        cmd = [nvcc] + LINK_FLAGS + ['-o', LIB_PATH] + objs
        if verbose:
            print(' '.join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB_PATH


_lib = None


def lib():
    """Load the shared library (never builds implicitly on a GPU box: the .so ships in-tree)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"`. '
                               'nerfart_b200 has no CPU / PyTorch fallback.')
        L = C.CDLL(LIB_PATH)
        L.na_version.restype = C.c_int
        L.na_error_string.restype = C.c_char_p
        L.na_error_string.argtypes = [C.c_int]
        L.na_last_cuda_error.restype = C.c_int
        L.na_kernel_launch_count.restype = C.c_int64
        L.na_preload_kernels.restype = C.c_int
        L.na_diag_enable.argtypes = [C.c_int]
        L.na_diag_dump.argtypes = [C.c_char_p, C.c_int]
        L.na_diag_dump.restype = C.c_int
        L.na_packed_weights_bytes.restype = C.c_size_t
        L.na_packed_weights_bytes.argtypes = [C.POINTER(NaNetDesc)]
        L.na_pack_weights.argtypes = [C.POINTER(NaNetDesc), C.POINTER(NaRawParams), C.c_void_p, C.c_void_p]
        L.na_eval_workspace_bytes.restype = C.c_size_t
        L.na_eval_workspace_bytes.argtypes = [C.c_int64]
        L.na_sdf_eval.argtypes = [C.POINTER(NaNetDesc), C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.na_full_eval.argtypes = [C.POINTER(NaNetDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.na_get_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.na_error_bound.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_float, C.c_float,
                                     C.c_void_p, C.c_void_p]
        for fn in (L.na_sample_pdf, L.na_sample_cdf):
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                           C.c_void_p, C.c_void_p]
        L.na_volsdf_workspace_bytes.restype = C.c_size_t
        L.na_volsdf_workspace_bytes.argtypes = [C.POINTER(NaVolsdfCfg), C.c_int64]
        L.na_volsdf_render_fwd.argtypes = [C.POINTER(NaNetDesc), C.c_void_p, C.POINTER(NaVolsdfCfg), C.c_void_p, C.c_void_p,
                                           C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.POINTER(NaVolsdfOut), C.c_void_p, C.c_size_t, C.c_void_p]
        L.na_volsdf_render_fwd_train.argtypes = [C.POINTER(NaNetDesc), C.c_void_p, C.POINTER(NaVolsdfCfg), C.c_void_p, C.c_void_p,
                                                 C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.POINTER(NaVolsdfOut), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        if hasattr(L, 'na_neus_render_fwd'):
            L.na_neus_workspace_bytes.restype = C.c_size_t
            L.na_neus_workspace_bytes.argtypes = [C.POINTER(NaNeusCfg), C.c_int64]
            L.na_neus_render_fwd.argtypes = [C.POINTER(NaNetDesc), C.c_void_p, C.POINTER(NaNeusCfg), C.c_void_p, C.c_void_p,
                                             C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.POINTER(NaNeusOut), C.c_void_p, C.c_size_t, C.c_void_p]
            L.na_neus_render_fwd_train.argtypes = [C.POINTER(NaNetDesc), C.c_void_p, C.POINTER(NaNeusCfg), C.c_void_p, C.c_void_p,
                                                   C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.POINTER(NaNeusOut), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        L.na_surface_workspace_bytes.restype = C.c_size_t
        L.na_surface_workspace_bytes.argtypes = [C.POINTER(NaSurfaceCfg), C.c_int64]
        L.na_ray_cast.argtypes = [C.POINTER(NaNetDesc), C.c_void_p, C.POINTER(NaSurfaceCfg), C.c_void_p, C.c_void_p, C.c_int64,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.na_surface_render_fwd.argtypes = [C.POINTER(NaNetDesc), C.c_void_p, C.POINTER(NaSurfaceCfg), C.c_void_p, C.c_void_p,
                                            C.c_int64, C.c_void_p, C.POINTER(NaSurfaceOut), C.c_void_p, C.c_size_t, C.c_void_p]
        L.na_grad_pack_bytes.restype = C.c_size_t
        L.na_grad_pack_bytes.argtypes = [C.POINTER(NaNetDesc)]
        L.na_train_workspace_bytes.restype = C.c_size_t
        L.na_train_workspace_bytes.argtypes = [C.POINTER(NaNetDesc), C.c_int64, C.c_int32]
        L.na_debug_wgrad_f16.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.na_train_workspace_bytes_mode.restype = C.c_size_t
        L.na_train_workspace_bytes_mode.argtypes = [C.POINTER(NaNetDesc), C.c_int64, C.c_int32, C.c_int32]
        for fn in (L.na_volsdf_render_bwd, L.na_neus_render_bwd, L.na_volsdf_render_bwd_stashed, L.na_neus_render_bwd_stashed):
            fn.argtypes = [C.POINTER(NaNetDesc), C.c_void_p, C.POINTER(NaTrainCfg), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                           C.c_size_t, C.c_void_p]
        L.na_unpack_grads.argtypes = [C.POINTER(NaNetDesc), C.POINTER(NaRawParams), C.c_void_p, C.POINTER(NaRawGrads), C.c_void_p]
        _lib = L
    return _lib


def check(code, what):
    if code != 0:
        L = lib()
        msg = L.na_error_string(code).decode()
        if code == -3:
            msg += f' [cudaError {L.na_last_cuda_error()}]'
        raise RuntimeError(f'nerfart_b200: {what} failed: {msg}')


def ptr(t):
    """Device pointer of a contiguous fp32 CUDA tensor (or NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('nerfart_b200 kernels need CUDA tensors (there is no CPU path)')
    assert t.is_contiguous(), 'tensor must be contiguous'
    return C.c_void_p(t.data_ptr())



def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def diag_enable(level=1):
    """Stall diagnostics (csrc/common.cuh).  1: per-launch CTA counters + a record of every timed-out mbarrier wait in
    host-mapped memory; 2: also an event per kernel launch (names the first launch that never finished); 0: off."""
    check(lib().na_diag_enable(int(level)), 'na_diag_enable')


def diag_dump():
    """Text report of na_diag_dump: callable from a watchdog thread while the stream is stuck, and after a launch failure."""
    buf = C.create_string_buffer(1 << 16)
    n = lib().na_diag_dump(buf, len(buf))
    return buf.raw[:n].decode(errors='replace')


_preloaded = set()


def preload_kernels(device):
    """Force-load every kernel image of the library on `device` once (na_preload_kernels)."""
    idx = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if idx in _preloaded:
        return
    with torch.cuda.device(idx):
        check(lib().na_preload_kernels(), 'na_preload_kernels')
    _preloaded.add(idx)


def launch_count():
    return int(lib().na_kernel_launch_count())
