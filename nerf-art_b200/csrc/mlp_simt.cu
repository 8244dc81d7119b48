// Fused per-sample network kernel, fp32 CUDA-core path (NA_PRECISION_FP32).
//
// One persistent CTA per SM walks tiles of TM=128 samples.  For every tile the whole network --
// positional encoding, the 8+1 layer SDF MLP (models/base.py:243-263), the closed-form reverse sweep
// that autograd performs for d sdf/dx (base.py:265-282), and the 4+1 layer
// radiance MLP (base.py:372-391) -- runs out of shared memory; activations never touch HBM.
// Weights (weight-norm folded, pre-transposed; pack.cu) stream from L2 in 8-row chunks through a
// 3-stage cp.async ring.  Softplus' values needed by the reverse sweep and the 256-d geometry feature
// are parked in a per-CTA, L2-resident scratch; each thread re-reads only what it wrote itself.
//
// Shared-memory layout (floats):  A[296][128] activations, k-major, float4-granular XOR swizzle
//   A[k][m] lives at  k*128 + (((m>>2) ^ ((k>>2)&7))<<2) + (m&3)
// so that both the GEMM reads (fixed k, 8 consecutive m) and the epilogue writes (16 lanes with
// k = 4*tx+c, 8 consecutive m) are bank-conflict free.  Rows 256..295 hold the 39-d embedding during
// the SDF forward pass and the small radiance inputs [x | embed(view) | nabla] afterwards.
#include "simt_tile.cuh"

namespace na {

__global__ void __launch_bounds__(NT, 1)
mlp_simt_kernel(const EvalJob job, const float* __restrict__ pk, const PackF32 L, float* __restrict__ scratch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MlpSmem& S = *reinterpret_cast<MlpSmem*>(smem_raw);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const bool explicit_pts = job.x != nullptr;
    const long long total = explicit_pts ? job.m
                          : (long long)(job.n_rows_dev ? min(*job.n_rows_dev, job.n_rows) : job.n_rows) * job.P;
    const long long n_tiles = (total + TM - 1) / TM;
    float* sp = scratch + (size_t)blockIdx.x * (9 * 256 * TM);      // 8 softplus' planes + 1 feature plane
    float* fplane = sp + 8 * 256 * TM;
    const int nv = job.multires_view < 0 ? 3 : 3 + 6 * job.multires_view;
    const int spad = small_pad(job.multires_view);
    float acc[8][16];

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // ---- 1. points, positional encoding (models/base.py:46-64) -> tail rows -----------------
        if (tid < TM) {
            const int m = tid;
            const long long w = tile * TM + m;
            float x0 = 0.f, x1 = 0.f, x2 = 0.f, v0 = 0.f, v1 = 0.f, v2 = 1.f;
            long long oidx = -1;
            if (w < total) {
                if (explicit_pts) {
                    x0 = job.x[w * 3 + 0]; x1 = job.x[w * 3 + 1]; x2 = job.x[w * 3 + 2];
                    if (job.view) { v0 = job.view[w * 3 + 0]; v1 = job.view[w * 3 + 1]; v2 = job.view[w * 3 + 2]; }
                    oidx = w;
                } else {
                    const long long row = w / job.P; const int j = (int)(w - row * job.P);
                    const long long ray = job.row_ids ? job.row_ids[row] : row;
                    const float* tp = job.t + ray * job.t_stride + job.t_off + j;
                    float t = tp[0];
                    if (job.midpoints) t = __fmul_rn(0.5f, __fadd_rn(tp[1], t));
                    v0 = job.rays_d[ray * 3 + 0]; v1 = job.rays_d[ray * 3 + 1]; v2 = job.rays_d[ray * 3 + 2];
                    // pts = rays_o + rays_d * d  (separately rounded mul and add, like the reference's tensor ops)
                    x0 = __fadd_rn(job.rays_o[ray * 3 + 0], __fmul_rn(v0, t));
                    x1 = __fadd_rn(job.rays_o[ray * 3 + 1], __fmul_rn(v1, t));
                    x2 = __fadd_rn(job.rays_o[ray * 3 + 2], __fmul_rn(v2, t));
                    oidx = ray * job.o_stride + job.o_off + j;
                }
            }
            S.OIDX[m] = oidx;
            S.X[m] = x0; S.X[TM + m] = x1; S.X[2 * TM + m] = x2;
            S.V[m] = v0; S.V[TM + m] = v1; S.V[2 * TM + m] = v2;
            const float xs[3] = {x0, x1, x2};
#pragma unroll
            for (int c = 0; c < 3; ++c) S.A[a_index(TAIL0 + c, m)] = xs[c];
#pragma unroll
            for (int f = 0; f < 6; ++f) {
                const float fr = (float)(1 << f);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float sn, cs; sincosf(__fmul_rn(xs[c], fr), &sn, &cs);
                    S.A[a_index(TAIL0 + 3 + 6 * f + c, m)] = sn;
                    S.A[a_index(TAIL0 + 6 + 6 * f + c, m)] = cs;
                }
            }
            S.A[a_index(TAIL0 + 39, m)] = 0.f;
        }
        // (gemm_tile starts with a __syncthreads after its first wait, which also publishes the writes above)

        // ---- 2. SDF forward, layers 0..7 -------------------------------------------------------
        for (int layer = 0; layer < N_SDF_HID; ++layer) {
            if (layer == 0) gemm_tile<4>(acc, pk + L.sdf_wt[0], EMB_PAD, TAIL0, S.A, S.Ws, tid);
            else            gemm_tile<4>(acc, pk + L.sdf_wt[layer], W, 0, S.A, S.Ws, tid);
            const float* bias = pk + L.sdf_b[layer];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int k = 64 * j + 4 * tx + c;
                    const float b = __ldg(bias + k);
                    float h[8], dh[8];
                    if (layer == 3 && k >= SKIP_H) {
                        // skip connection: h = cat([h(217), emb(39)]) ; the 1/sqrt2 lives in layer 4's packed weights
#pragma unroll
                        for (int i = 0; i < 8; ++i) { h[i] = S.A[a_index(TAIL0 + (k - SKIP_H), 8 * ty + i)]; dh[i] = 0.f; }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) softplus100(acc[i][4 * j + c] + b, h[i], dh[i]);
                    }
                    store_col_A(S.A, k, ty, tx, h);
                    if (job.want_full) store_col_plane(sp + layer * 256 * TM, k, ty, dh);
                }
            }
        }
        __syncthreads();
        // ---- 3. sdf = <h8, W8[0]> + b8[0]  (+ VolSDF sphere background, volsdf.py:341-357) -------
        narrow_layer<1>(S.A, pk + L.w8_sdf, S.RED, tid);
        __syncthreads();
        if (tid < TM) {
            const int m = tid;
            float sdf = S.RED[m] + S.RED[3 * TM + m] + __ldg(pk + L.b8_sdf);
            if (job.apply_bg) {
                const float x0 = S.X[m], x1 = S.X[TM + m], x2 = S.X[2 * TM + m];
                const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x1, x1)), __fmul_rn(x2, x2)));
                sdf = fminf(sdf, job.bound_r - nrm);
            }
            S.SDF[m] = sdf;
            if (S.OIDX[m] >= 0 && job.sdf) job.sdf[S.OIDX[m]] = sdf;
        }
        if (!job.want_full && !job.feat) { __syncthreads(); continue; }

        // ---- 4. geometry feature = h8 @ W8[1:257]^T + b8[1:]  -> scratch plane (and feat output) --
        gemm_tile<4>(acc, pk + L.w8t_feat, W, 0, S.A, S.Ws, tid);
        {
            const float* bias = pk + L.b8_feat;
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int k = 64 * j + 4 * tx + c;
                    const float b = __ldg(bias + k);
                    float f[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { f[i] = acc[i][4 * j + c] + b; acc[i][4 * j + c] = f[i]; }
                    store_col_plane(fplane, k, ty, f);
                }
            if (job.feat) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const long long o = S.OIDX[8 * ty + i];
                    if (o >= 0) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<float4*>(job.feat + o * 256 + 64 * j + 4 * tx) =
                                make_float4(acc[i][4 * j], acc[i][4 * j + 1], acc[i][4 * j + 2], acc[i][4 * j + 3]);
                    }
                }
            }
        }
        if (!job.want_full) { __syncthreads(); continue; }

        // ---- 5. reverse sweep: d sdf / d x  (what autograd.grad does at base.py:271-277) ---------
        // A <- d sdf/d z7 = W8[0,:] * softplus'(z7)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int k = 64 * j + 4 * tx + c;
                const float w8 = __ldg(pk + L.w8_sdf + k);
                float d[8];
                load_col_plane(sp + 7 * 256 * TM, k, ty, d);
#pragma unroll
                for (int i = 0; i < 8; ++i) d[i] *= w8;
                store_col_A(S.A, k, ty, tx, d);
            }
        for (int layer = 7; layer >= 1; --layer) {
            gemm_tile<4>(acc, pk + L.sdf_w[layer], W, 0, S.A, S.Ws, tid);      // -> d sdf / d (input of `layer`)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int k = 64 * j + 4 * tx + c;
                    float d[8], g[8];
                    load_col_plane(sp + (layer - 1) * 256 * TM, k, ty, d);
#pragma unroll
                    for (int i = 0; i < 8; ++i) g[i] = acc[i][4 * j + c];
                    if (layer == 4 && k >= SKIP_H) {
                        // embedding branch of the skip connection
#pragma unroll
                        for (int i = 0; i < 8; ++i) S.GE[(k - SKIP_H) * TM + 8 * ty + i] = g[i];
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) g[i] *= d[i];               // softplus'(z_{layer-1}); 0 on the skip columns
                    store_col_A(S.A, k, ty, tx, g);
                }
        }
        {
            float acc1[8][4];
            gemm_tile<1>(acc1, pk + L.sdf_w[0], W, 0, S.A, S.Ws, tid);      // d sdf / d emb, 39 useful columns
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int k = 4 * tx + c;
                if (k < EMB) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) S.GE[k * TM + 8 * ty + i] += acc1[i][c];
                }
            }
        }
        __syncthreads();
        // ---- 6. nabla and the small radiance inputs -> tail rows ---------------
        if (tid < TM) {
            const int m = tid;
            const float xs[3] = {S.X[m], S.X[TM + m], S.X[2 * TM + m]};
            float nb[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float n = S.GE[c * TM + m];
#pragma unroll
                for (int f = 0; f < 6; ++f) {
                    const float fr = (float)(1 << f);
                    float sn, cs; sincosf(__fmul_rn(xs[c], fr), &sn, &cs);
                    n += fr * (S.GE[(3 + 6 * f + c) * TM + m] * cs - S.GE[(6 + 6 * f + c) * TM + m] * sn);
                }
                nb[c] = n; S.NAB[c * TM + m] = n;
            }
            if (job.rad) {
                int q = 0;
#pragma unroll
                for (int c = 0; c < 3; ++c) S.A[a_index(TAIL0 + q++, m)] = xs[c];
                const float vs[3] = {S.V[m], S.V[TM + m], S.V[2 * TM + m]};
#pragma unroll
                for (int c = 0; c < 3; ++c) S.A[a_index(TAIL0 + q++, m)] = vs[c];
                if (job.multires_view >= 0) {
                    for (int f = 0; f < job.multires_view; ++f) {
                        const float fr = (float)(1 << f);
                        float sn[3], cs[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) sincosf(__fmul_rn(vs[c], fr), &sn[c], &cs[c]);
#pragma unroll
                        for (int c = 0; c < 3; ++c) S.A[a_index(TAIL0 + q++, m)] = sn[c];
#pragma unroll
                        for (int c = 0; c < 3; ++c) S.A[a_index(TAIL0 + q++, m)] = cs[c];
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) S.A[a_index(TAIL0 + q++, m)] = nb[c];
                for (; q < spad; ++q) S.A[a_index(TAIL0 + q, m)] = 0.f;
            }
            const long long o = S.OIDX[m];
            if (o >= 0 && job.nab) { job.nab[o * 3 + 0] = nb[0]; job.nab[o * 3 + 1] = nb[1]; job.nab[o * 3 + 2] = nb[2]; }
        }
        if (!job.rad) { __syncthreads(); continue; }
        (void)nv;
        // ---- 7. radiance net (base.py:372-391) ---------------------------------------------------
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int k = 64 * j + 4 * tx + c;
                float f[8];
                load_col_plane(fplane, k, ty, f);
                store_col_A(S.A, k, ty, tx, f);
            }
        for (int layer = 0; layer < 4; ++layer) {
            gemm_tile<4>(acc, pk + L.rad_wt[layer], layer == 0 ? W + spad : W, 0, S.A, S.Ws, tid);
            const float* bias = pk + L.rad_b[layer];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int k = 64 * j + 4 * tx + c;
                    const float b = __ldg(bias + k);
                    float h[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) h[i] = fmaxf(acc[i][4 * j + c] + b, 0.f);
                    store_col_A(S.A, k, ty, tx, h);
                }
        }
        __syncthreads();
        narrow_layer<3>(S.A, pk + L.rad_w4, S.RED, tid);
        __syncthreads();
        if (tid < TM) {
            const int m = tid;
            const long long o = S.OIDX[m];
            if (o >= 0) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    job.rad[o * 3 + c] = sigmoidf_(S.RED[c * TM + m] + S.RED[(3 + c) * TM + m] + __ldg(pk + L.rad_b4 + c));
            }
        }
        __syncthreads();
    }
}

size_t mlp_simt_scratch_bytes(int grid) { return (size_t)grid * 9 * 256 * TM * sizeof(float); }

int launch_mlp_simt(const EvalJob& job, const float* packed, const PackF32& L, float* scratch, size_t scratch_bytes,
                    cudaStream_t stream) {
    static bool attr_done[64] = {false};
    const size_t smem = sizeof(MlpSmem);
    if (first_on_device(attr_done)) {
        NA_TRY(check_cuda(cudaFuncSetAttribute(mlp_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    }
    const long long total = job.x ? job.m : (long long)job.n_rows * job.P;
    if (total <= 0) return NA_OK;
    long long tiles = (total + TM - 1) / TM;
    int grid = (int)(tiles < (long long)num_sms() ? tiles : (long long)num_sms());
    if (scratch_bytes < mlp_simt_scratch_bytes(grid)) return NA_ERR_WORKSPACE;
    mlp_simt_kernel<<<grid, NT, smem, stream>>>(job, packed, L, scratch);
    NA_CHECK_LAUNCH();
    return NA_OK;
}

int preload_mlp_simt() {
    NA_PRELOAD(mlp_simt_kernel);
    return NA_OK;
}

}  // namespace na
