"""CPU: the quad layout of the training stash (csrc/common.cuh: stash_quad_index) is a bijection of every 128-sample tile onto
itself, keeps a column quad of one sample contiguous (float4) and 32 consecutive samples of a quad contiguous (512 B) -- the two
properties the tcgen05 epilogues (thread = sample row) and the weight-gradient loaders rely on.  Compiles the header's own
function for the host with nvcc (no GPU needed)."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include "common.cuh"
#include <cstdio>
#include <vector>
int main() {
    const long long M = 384;                       // three tiles
    std::vector<int> seen((size_t)M * 256, 0);
    for (long long m = 0; m < M; ++m)
        for (int c = 0; c < 256; ++c) {
            const size_t i = na::stash_quad_index(m, c);
            if (i >= seen.size() || seen[i]++) { std::printf("not a bijection at m=%lld c=%d\n", m, c); return 1; }
            if (i / (128 * 256) != (size_t)(m / 128)) { std::printf("leaves its tile at m=%lld c=%d\n", m, c); return 2; }
            if ((c & 3) && i != na::stash_quad_index(m, c - 1) + 1) { std::printf("quad not contiguous\n"); return 3; }
            if ((m & 127) && i != na::stash_quad_index(m - 1, c) + 4) { std::printf("samples of a quad not contiguous\n"); return 4; }
        }
    std::printf("ok\n");
    return 0;
}
'''


@pytest.mark.skipif(shutil.which('nvcc') is None, reason='nvcc not on PATH')
def test_stash_quad_index_is_a_tilewise_bijection_with_contiguous_quads():
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, 't.cu'), os.path.join(d, 't')
        with open(src, 'w') as f:
            f.write(SRC)
        subprocess.run(['nvcc', '-std=c++17', '-O1', '-I', os.path.join(ROOT, 'nerf-art_b200', 'csrc'), '-o', exe, src], check=True)
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0 and out.stdout.strip() == 'ok', out.stdout + out.stderr
