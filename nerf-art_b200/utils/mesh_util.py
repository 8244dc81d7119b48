"""Mesh extraction with the reference's names and signatures (reference utils/mesh_util.py).

`extract_mesh` evaluates the SDF on an N^3 grid (512^3 = 134 M points by default) and runs marching cubes.  The reference builds
the 3.2 GB float64 coordinate table on the host and pushes 8192 chunks of 16 K points through `implicit_surface.forward`
(mesh_util.py:82-110); here the coordinates are generated on the device per chunk and evaluated by the fused SDF-only kernel
(`na_sdf_eval`, 134 M points in well under a second), and only the N^3 float32 grid crosses to the host for
`skimage.measure.marching_cubes` (host-side, outside the hot path, exactly as in the reference).

Reproduced on purpose: the reference indexes the grid with *float* divisions (`(i / N) % N`, mesh_util.py:92-94 -- the Python-2
integer divisions of the code it was adapted from), so the y / x coordinates of a point carry the fractions z/N and y/N + z/N^2 of a
cell; the sampled lattice is sheared by up to one cell.  `np.int` (mesh_util.py:87, removed from numpy 1.24) is not reproduced.
"""
import time

import numpy as np
import torch

__all__ = ['extract_mesh', 'convert_sigma_samples_to_ply', 'grid_points', 'sdf_grid']


def grid_points(i0, i1, N, s, device):
    """Coordinates of grid points i0 .. i1-1 as the reference computes them (mesh_util.py:87-100), float64 arithmetic, float32 result
    [i1-i0, 3]."""
    idx = torch.arange(i0, i1, device=device, dtype=torch.float64)
    z = torch.remainder(idx, N)
    y = torch.remainder(idx / N, N)
    x = torch.remainder((idx / N) / N, N)
    step, origin = s / (N - 1), -s / 2.
    return torch.stack((x * step + origin, y * step + origin, z * step + origin), dim=-1).float()


def sdf_grid(implicit_surface, volume_size=2.0, N=512, show_progress=False, chunk=16 * 1024 * 1024):
    """float32 [N,N,N] SDF samples of `implicit_surface.forward` on the reference's lattice (numpy array on the host)."""
    dev = next(implicit_surface.parameters()).device
    out = torch.empty(N ** 3, dtype=torch.float32, device=dev)
    rng = range(0, N ** 3, chunk)
    if show_progress:
        from tqdm import tqdm
        rng = tqdm(rng)
    with torch.no_grad():
        for i in rng:
            j = min(i + chunk, N ** 3)
            out[i:j] = implicit_surface.forward(grid_points(i, j, N, volume_size, dev))
    return out.reshape(N, N, N).cpu().numpy()


def write_ply(path, verts, faces):
    """Binary little-endian PLY with float x / y / z vertices and `list uchar int vertex_indices` faces (what plyfile writes for the
    reference's element descriptions, mesh_util.py:60-74)."""
    verts = np.ascontiguousarray(verts, dtype='<f4')
    faces = np.ascontiguousarray(faces, dtype='<i4')
    rec = np.empty(len(faces), dtype=[('n', 'u1'), ('v', '<i4', (3,))])
    rec['n'] = 3
    rec['v'] = faces
    with open(path, 'wb') as f:
        f.write(('ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n'
                 'element face %d\nproperty list uchar int vertex_indices\nend_header\n' % (len(verts), len(faces))).encode())
        f.write(verts.tobytes())
        f.write(rec.tobytes())


def convert_sigma_samples_to_ply(input_3d_sigma_array, voxel_grid_origin, volume_size, ply_filename_out, level=5.0, offset=None,
                                 scale=None):
    """reference mesh_util.convert_sigma_samples_to_ply (12-80): marching cubes at `level` with voxel spacing `volume_size`, vertices
    shifted by the grid origin (then / scale, - offset), written as PLY.  Needs scikit-image on the host."""
    try:
        import skimage.measure
    except Exception as e:
        raise RuntimeError('extract_mesh needs scikit-image for marching cubes (host side; reference utils/mesh_util.py:33)') from e
    t0 = time.time()
    verts, faces, _, _ = skimage.measure.marching_cubes(input_3d_sigma_array, level=level, spacing=volume_size)
    pts = np.asarray(verts, dtype=np.float64) + np.asarray(voxel_grid_origin, dtype=np.float64)[None, :]
    if scale is not None:
        pts = pts / scale
    if offset is not None:
        pts = pts - offset
    write_ply(ply_filename_out, pts, faces)
    return time.time() - t0


def extract_mesh(implicit_surface, volume_size=2.0, level=0.0, N=512, filepath='./surface.ply', show_progress=True, chunk=16 * 1024):
    """reference mesh_util.extract_mesh (82-112), same arguments.  `chunk` is the reference's per-call point count; the fused kernel
    has no per-point activation memory, so at least 16 Mi points go into one launch."""
    s = volume_size
    grid = sdf_grid(implicit_surface, s, N, show_progress, max(int(chunk), 16 * 1024 * 1024))
    convert_sigma_samples_to_ply(grid, [-s / 2., -s / 2., -s / 2.], [float(s) / N] * 3, filepath, level=level)
