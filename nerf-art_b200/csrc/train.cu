// Backward of the volumetric render: the second pass of the fine-tune step (`Trainer.forward`,
// models/frameworks/volsdf.py:769-783, models/frameworks/neus.py:551-563):
//     rgb_pred.backward(gradient_patch);  (w_eikonal * mse(||implicit_nablas||, 1)).backward()
// The reference leaves this to autograd, including the second-order path through
// `autograd.grad(sdf, x, create_graph=True)` (models/base.py:265-282).  Here it is three hand-written stages:
//
//   1. `volsdf_composite_bwd_kernel` / `neus_composite_bwd_kernel` (one thread per ray): dL/d rgb -> per-sample
//      dL/d radiance, dL/d sdf, dL/d nabla (eikonal term), and the scalar dL/d ln_beta (VolSDF) / dL/d ln_s (NeuS).
//   2. `mlp_bwd_kernel` (fp32 CUDA cores; persistent, one 128-sample tile per CTA at a time, activations in shared memory):
//      recomputes the forward pass of the tile, then runs radiance backward, the forward-like second-order sweep and
//      the first-order trunk backward (oracle/nerfart_oracle_train.py::mlp_backward is the same program
//      in numpy).  Every (delta, input) pair a weight gradient needs is written once to sample-major planes in HBM.
//   3. `wgrad_kernel` (split over samples, atomics into the packed gradient) + `colsum_kernel` (biases):
//      dW = sum_samples delta^T input, accumulated in the GradPack buffer; `unpack_grads_kernel` maps GradPack to the
//      reference's parameters (weight_g, weight_v, bias) through the weight-norm Jacobian.
#include "simt_tile.cuh"
#include <cstdlib>
#include <cstring>

namespace na {

// ---------------------------------------------------------------------------------------------------------------------
// stash planes (sample-major [Mpad][256] fp32 unless noted), written by mlp_bwd_kernel, read by wgrad / colsum
// ---------------------------------------------------------------------------------------------------------------------
enum {
    PL_IN = 0,      // IN_1..IN_8  (IN_i = input of SDF layer i = h_{i-1}; IN_4 = [h_3 | emb]; IN_8 = h_7)      planes 0..7
    PL_ZB = 8,      // z-bar_0..7: dL/d(pre-activation) of SDF layer i                                             planes 8..15
    PL_G = 16,      // g_0..7: reverse-sweep values u_i * softplus'(z_i)                                           planes 16..23
    PL_VB = 24,     // v-bar_1..8: gradient w.r.t. the reverse-sweep products (v-bar_8 = u-bar_7)                  planes 24..31
    PL_FEAT = 32,   // geometry feature (radiance layer 0 input)
    PL_FB = 33,     // dL/d feature
    PL_YS = 34,     // ys_1..4: outputs of radiance layers 0..3                                                    planes 34..37
    PL_D = 38,      // delta_0..3: dL/d(pre-activation) of radiance layers 0..3                                    planes 38..41
    N_WIDE = 42
};
static_assert(PL_IN == ST_IN && PL_G == ST_G && PL_FEAT == ST_FEAT && PL_YS == ST_YS && PL_ZB == ST_ZB && PL_VB == ST_VB &&
              PL_FB == ST_FB && PL_D == ST_D && N_WIDE == ST_N_WIDE, "stash plane numbering (common.cuh)");
constexpr int NLD = 40;     // row stride of the narrow planes EMB, VB0, SMALL
enum { NP_EMB = 0, NP_VB0 = 1, NP_SMALL = 2 };
struct Stash {
    float* wide;        // [N_WIDE][Mpad][256]
    float* narrow;      // [3][Mpad][40]
    float* tiny;        // [2][Mpad][4]:  0 = delta_4 (radiance output layer), 1 = masked dL/d sdf
    size_t mpad;
    __host__ __device__ float* w(int p) const { return wide + (size_t)p * mpad * 256; }
    __host__ __device__ float* n(int p) const { return narrow + (size_t)p * mpad * NLD; }
    __host__ __device__ float* t(int p) const { return tiny + (size_t)p * mpad * 4; }
};
static size_t stash_floats(size_t mpad) { return mpad * ((size_t)N_WIDE * 256 + 3 * NLD + 2 * 4); }

struct BwdJob {
    const float* rays_o; const float* rays_d;      // [n_rows,3], d normalised
    int n_rows, P;                                 // points per ray evaluated by this launch
    const float* t; long long t_stride; int midpoints;
    const float* g_sdf; const float* g_nab; const float* g_rad;     // dense per sample [n_rows*P](,3); nullable = 0
    int apply_bg; float bound_r;
    int has_rad;                                    // 0: NeuS pass A (sdf + nabla only)
    int multires_view;
    // loaded mode: raw sdf [n_rows*P] and radiance [n_rows*P,3] of the forward launch that filled the IN / S / G / FEAT / YS planes
    const float* f_sdf; const float* f_rad;
};

struct __align__(16) TrainSmem {
    MlpSmem s;
    float GSDF[TM];        // upstream dL/d sdf
    float GNAB[3 * TM];    // upstream dL/d nabla (eikonal)
    float GRAD[3 * TM];    // upstream dL/d radiance
    float NB[3 * TM];      // dL/d nabla arriving through the radiance net
    float D4[3 * TM];      // delta of the radiance output layer
    float GS[TM];          // dL/d sdf after the sphere-background mask
};

// fragment helpers: thread (ty, tx) owns rows 8ty..8ty+7 and, for j = 0..3, columns 64j + 4tx + {0..3}
__device__ __forceinline__ void ld_frag4(const float* __restrict__ plane, int ty, int tx, int j, float4 (&p)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = *reinterpret_cast<const float4*>(plane + (size_t)(8 * ty + i) * 256 + 64 * j + 4 * tx);
}
__device__ __forceinline__ void st_frag4(float* __restrict__ plane, int ty, int tx, int j, const float (&acc)[8][16]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(plane + (size_t)(8 * ty + i) * 256 + 64 * j + 4 * tx) =
            make_float4(acc[i][4 * j], acc[i][4 * j + 1], acc[i][4 * j + 2], acc[i][4 * j + 3]);
}
__device__ __forceinline__ void st_frag4v(float* __restrict__ plane, int ty, int tx, int j, const float4 (&p)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(plane + (size_t)(8 * ty + i) * 256 + 64 * j + 4 * tx) = p[i];
}
__device__ __forceinline__ void st_A4(float* A, int ty, int tx, int j, const float (&acc)[8][16]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = acc[i][4 * j + c];
        store_col_A(A, 64 * j + 4 * tx + c, ty, tx, v);
    }
}
__device__ __forceinline__ void load_col_A(const float* A, int k, int ty, int tx, float (&v)[8]) {
    const float* base = A + k * TM;
    const int swz = tx & 7;
    const float4 a = *reinterpret_cast<const float4*>(base + (((2 * ty) ^ swz) << 2));
    const float4 b = *reinterpret_cast<const float4*>(base + (((2 * ty + 1) ^ swz) << 2));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ float f4c(const float4& v, int c) { return c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w; }

__device__ __forceinline__ float to_tf32(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Tensor-core tile GEMM for the backward-data products (linear in the upstream gradient, so TF32 operands are enough):
// OUT[m][c] = sum_{r<R} A[a_row0 + r][m] * B[r][c], c < 256, `mma.sync.m16n8k8` TF32, fp32 accumulate.  B planes are pre-rounded
// to TF32 (PackTrain.*_r); A is rounded (cvt.rna) at fragment load.  8 warps = 4 (m) x 2 (n); warp (wm, wn) owns rows 32 wm..+31
// and columns 128 wn..+127; acc[mt][nt][f] is the m16n8 C fragment: row 32 wm + 16 mt + g + 8 (f>>1), column 128 wn + 8 nt + 2 t + (f&1)
// with g = lane/4, t = lane%4.  Weight ring rows are 264 floats apart so a fragment load touches 32 distinct banks.
constexpr int WSTR = 264;
__device__ __forceinline__ void gemm_tile_tf32(float (&acc)[2][16][4], const float* __restrict__ Bg, int R, int a_row0,
                                               const float* __restrict__ As, float* __restrict__ Ws, int tid) {
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int m0 = (warp >> 1) * 32, n0 = (warp & 1) * 128;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j)
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[i][j][f] = 0.f;
    const int nch = R / KC;
    auto issue = [&](int c) {
        float* dst = Ws + (c % 3) * (KC * WSTR);
        const float* src = Bg + (size_t)c * KC * 256;
        for (int idx = tid; idx < KC * 64; idx += NT) {
            const int row = idx >> 6, c4 = idx & 63;
            cp_async16(dst + row * WSTR + c4 * 4, src + row * 256 + c4 * 4);
        }
        cp_async_commit();
    };
    issue(0);
    if (nch > 1) issue(1);
    for (int c = 0; c < nch; ++c) {
        if (c + 1 < nch) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();
        if (c + 2 < nch) issue(c + 2);
        const float* wb = Ws + (c % 3) * (KC * WSTR);
        const int k0 = a_row0 + c * KC;
        unsigned af[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int m = m0 + 16 * mt + g;
            af[mt][0] = __float_as_uint(to_tf32(As[a_index(k0 + t, m)]));
            af[mt][1] = __float_as_uint(to_tf32(As[a_index(k0 + t, m + 8)]));
            af[mt][2] = __float_as_uint(to_tf32(As[a_index(k0 + t + 4, m)]));
            af[mt][3] = __float_as_uint(to_tf32(As[a_index(k0 + t + 4, m + 8)]));
        }
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) {
            unsigned bf[2];
            bf[0] = __float_as_uint(wb[t * WSTR + n0 + 8 * nt + g]);
            bf[1] = __float_as_uint(wb[(t + 4) * WSTR + n0 + 8 * nt + g]);
            mma_tf32(acc[0][nt], af[0], bf);
            mma_tf32(acc[1][nt], af[1], bf);
        }
    }
    __syncthreads();
}
// visit the thread's C fragment as (row, even column, value pair)
template <typename F>
__device__ __forceinline__ void mma_each(float (&acc)[2][16][4], int tid, F f) {
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int m0 = (warp >> 1) * 32, n0 = (warp & 1) * 128;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 16; ++nt)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf)
                f(m0 + 16 * mt + g + 8 * hf, n0 + 8 * nt + 2 * t, acc[mt][nt][2 * hf], acc[mt][nt][2 * hf + 1]);
}
__device__ __forceinline__ float2 ld2(const float* __restrict__ plane, int row, int col) {
    return *reinterpret_cast<const float2*>(plane + (size_t)row * 256 + col);
}
__device__ __forceinline__ void st2(float* __restrict__ plane, int row, int col, float a, float b) {
    *reinterpret_cast<float2*>(plane + (size_t)row * 256 + col) = make_float2(a, b);
}

template <bool TF32>
__global__ void __launch_bounds__(NT, 1)
mlp_bwd_kernel(const BwdJob job, const float* __restrict__ pk, const PackF32 L, const float* __restrict__ tp, const PackTrain T,
               const Stash st, float* __restrict__ scratch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TrainSmem& Q = *reinterpret_cast<TrainSmem*>(smem_raw);
    MlpSmem& S = Q.s;
    constexpr bool LOADED = false;      // (a former mode that took the forward activations from a tcgen05 launch; the tensor-core
                                        // modes now run the whole backward in csrc/mlp_tmem.cu + csrc/wgrad_f16.cu)
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long total = (long long)job.n_rows * job.P;
    const long long n_tiles = (total + TM - 1) / TM;
    float* SP = scratch + (size_t)blockIdx.x * (16 * 256 * TM);     // per-CTA: 8 softplus' planes, then 8 second-order planes
    float* QP = SP + 8 * 256 * TM;
    const int nv = job.multires_view < 0 ? 3 : 3 + 6 * job.multires_view;
    const int spad = small_pad(job.multires_view);
    float acc[8][16];
    float (&acc2)[2][16][4] = reinterpret_cast<float (&)[2][16][4]>(acc);      // the same registers as an mma C fragment

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const size_t row0 = (size_t)tile * TM;
        auto WP = [&](int p) { return st.w(p) + row0 * 256; };
        // softplus' of SDF layer i as a [TM][256] tile: the forward launch's stash plane (loaded mode) or this CTA's scratch
        auto SPL = [&](int layer) -> float* { return SP + layer * 256 * TM; };
        // ---- 0. points, upstream gradients, positional encoding ---------------------------------------------------
        if (tid < TM) {
            const int m = tid;
            const long long w = tile * TM + m;
            float x0 = 0.f, x1 = 0.f, x2 = 0.f, v0 = 0.f, v1 = 0.f, v2 = 1.f;
            float gs = 0.f, gn[3] = {0.f, 0.f, 0.f}, gr[3] = {0.f, 0.f, 0.f};
            if (w < total) {
                const long long ray = w / job.P; const int j = (int)(w - ray * job.P);
                const float* tp_ = job.t + ray * job.t_stride + j;
                float t = tp_[0];
                if (job.midpoints) t = __fmul_rn(0.5f, __fadd_rn(tp_[1], t));
                v0 = job.rays_d[ray * 3 + 0]; v1 = job.rays_d[ray * 3 + 1]; v2 = job.rays_d[ray * 3 + 2];
                x0 = __fadd_rn(job.rays_o[ray * 3 + 0], __fmul_rn(v0, t));
                x1 = __fadd_rn(job.rays_o[ray * 3 + 1], __fmul_rn(v1, t));
                x2 = __fadd_rn(job.rays_o[ray * 3 + 2], __fmul_rn(v2, t));
                if (job.g_sdf) gs = job.g_sdf[w];
                if (job.g_nab) { gn[0] = job.g_nab[w * 3]; gn[1] = job.g_nab[w * 3 + 1]; gn[2] = job.g_nab[w * 3 + 2]; }
                if (job.g_rad) { gr[0] = job.g_rad[w * 3]; gr[1] = job.g_rad[w * 3 + 1]; gr[2] = job.g_rad[w * 3 + 2]; }
            }
            S.X[m] = x0; S.X[TM + m] = x1; S.X[2 * TM + m] = x2;
            S.V[m] = v0; S.V[TM + m] = v1; S.V[2 * TM + m] = v2;
            Q.GSDF[m] = gs;
#pragma unroll
            for (int c = 0; c < 3; ++c) { Q.GNAB[c * TM + m] = gn[c]; Q.GRAD[c * TM + m] = gr[c]; Q.NB[c * TM + m] = 0.f; }
            const float xs[3] = {x0, x1, x2};
            float* emb_row = st.n(NP_EMB) + (row0 + m) * NLD;
#pragma unroll
            for (int c = 0; c < 3; ++c) { S.A[a_index(TAIL0 + c, m)] = xs[c]; emb_row[c] = xs[c]; }
#pragma unroll
            for (int f = 0; f < 6; ++f) {
                const float fr = (float)(1 << f);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float sn, cs; sincosf(__fmul_rn(xs[c], fr), &sn, &cs);
                    S.A[a_index(TAIL0 + 3 + 6 * f + c, m)] = sn;  emb_row[3 + 6 * f + c] = sn;
                    S.A[a_index(TAIL0 + 6 + 6 * f + c, m)] = cs;  emb_row[6 + 6 * f + c] = cs;
                }
            }
            S.A[a_index(TAIL0 + 39, m)] = 0.f; emb_row[39] = 0.f;
        }
        if constexpr (!LOADED) {
        // ---- 1. SDF forward, layers 0..7: h_i -> A and IN_{i+1}; softplus' -> SP_i -------------------------------------
        for (int layer = 0; layer < N_SDF_HID; ++layer) {
            if (layer == 0) gemm_tile<4>(acc, pk + L.sdf_wt[0], EMB_PAD, TAIL0, S.A, S.Ws, tid);
            else            gemm_tile<4>(acc, pk + L.sdf_wt[layer], W, 0, S.A, S.Ws, tid);
            const float* bias = pk + L.sdf_b[layer];
            float* sp = SP + layer * 256 * TM;
            float* inp = WP(PL_IN + layer);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 sv[8];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int k = 64 * j + 4 * tx + c;
                    const float b = __ldg(bias + k);
                    if (layer == 3 && k >= SKIP_H) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            acc[i][4 * j + c] = S.A[a_index(TAIL0 + (k - SKIP_H), 8 * ty + i)];
                            (c == 0 ? sv[i].x : c == 1 ? sv[i].y : c == 2 ? sv[i].z : sv[i].w) = 0.f;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float h, dh; softplus100(acc[i][4 * j + c] + b, h, dh);
                            acc[i][4 * j + c] = h;
                            (c == 0 ? sv[i].x : c == 1 ? sv[i].y : c == 2 ? sv[i].z : sv[i].w) = dh;
                        }
                    }
                }
                st_frag4v(sp, ty, tx, j, sv);
                st_A4(S.A, ty, tx, j, acc);
                st_frag4(inp, ty, tx, j, acc);
            }
        }
                }
        __syncthreads();
        // ---- 2. sdf head, sphere-background mask of the upstream gradient (volsdf.py:349-357) ---------------------------
        if constexpr (!LOADED) {
            narrow_layer<1>(S.A, pk + L.w8_sdf, S.RED, tid);
            __syncthreads();
        }
        if (tid < TM) {
            const int m = tid;
            float sdf;
            if constexpr (LOADED) { const long long w = tile * TM + m; sdf = w < total ? job.f_sdf[w] : 0.f; }
            else sdf = S.RED[m] + S.RED[3 * TM + m] + __ldg(pk + L.b8_sdf);
            float gs = Q.GSDF[m];
            if (job.apply_bg) {
                const float x0 = S.X[m], x1 = S.X[TM + m], x2 = S.X[2 * TM + m];
                const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x1, x1)), __fmul_rn(x2, x2)));
                if (job.bound_r - nrm < sdf) gs = 0.f;
            }
            Q.GS[m] = gs;
            *reinterpret_cast<float4*>(st.t(1) + (row0 + m) * 4) = make_float4(gs, 0.f, 0.f, 0.f);
        }
        if constexpr (!LOADED) {
        // ---- 3. geometry feature -> FEAT plane -------------------------------------------------------------------------
        if (job.has_rad) {
            gemm_tile<4>(acc, pk + L.w8t_feat, W, 0, S.A, S.Ws, tid);
            const float* bias = pk + L.b8_feat;
            float* fp = WP(PL_FEAT);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float b = __ldg(bias + 64 * j + 4 * tx + c);
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i][4 * j + c] += b;
                }
                st_frag4(fp, ty, tx, j, acc);
            }
        }
        __syncthreads();
        // ---- 4. reverse sweep (autograd.grad(sdf, x), base.py:271-277): g_i -> A and G_i --------------------------------
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float4 sv[8];
            ld_frag4(SP + 7 * 256 * TM, ty, tx, j, sv);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float w8 = __ldg(pk + L.w8_sdf + 64 * j + 4 * tx + c);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i][4 * j + c] = w8 * f4c(sv[i], c);
            }
            st_A4(S.A, ty, tx, j, acc);
            st_frag4(WP(PL_G + 7), ty, tx, j, acc);
        }
        for (int layer = 7; layer >= 1; --layer) {
            gemm_tile<4>(acc, pk + L.sdf_w[layer], W, 0, S.A, S.Ws, tid);      // u_{layer-1}
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 sv[8];
                ld_frag4(SP + (layer - 1) * 256 * TM, ty, tx, j, sv);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int k = 64 * j + 4 * tx + c;
                    if (layer == 4 && k >= SKIP_H) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) S.GE[(k - SKIP_H) * TM + 8 * ty + i] = acc[i][4 * j + c];
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i][4 * j + c] *= f4c(sv[i], c);
                }
                st_A4(S.A, ty, tx, j, acc);
                st_frag4(WP(PL_G + layer - 1), ty, tx, j, acc);
            }
        }
        {
            float acc1[8][4];
            gemm_tile<1>(acc1, pk + L.sdf_w[0], W, 0, S.A, S.Ws, tid);          // d sdf / d emb
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
        }
        if constexpr (!LOADED) {
        // ---- 5. nabla; radiance forward ---------------------------------------------------------------------------------
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
                nb[c] = n;
            }
            if (job.has_rad) {
                float* srow = st.n(NP_SMALL) + (row0 + m) * NLD;
                int q = 0;
                auto put = [&](float v) { S.A[a_index(TAIL0 + q, m)] = v; srow[q] = v; ++q; };
#pragma unroll
                for (int c = 0; c < 3; ++c) put(xs[c]);
                const float vs[3] = {S.V[m], S.V[TM + m], S.V[2 * TM + m]};
#pragma unroll
                for (int c = 0; c < 3; ++c) put(vs[c]);
                if (job.multires_view >= 0) {
                    for (int f = 0; f < job.multires_view; ++f) {
                        const float fr = (float)(1 << f);
                        float sn[3], cs[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) sincosf(__fmul_rn(vs[c], fr), &sn[c], &cs[c]);
#pragma unroll
                        for (int c = 0; c < 3; ++c) put(sn[c]);
#pragma unroll
                        for (int c = 0; c < 3; ++c) put(cs[c]);
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) put(nb[c]);
                while (q < NLD) { if (q < spad) S.A[a_index(TAIL0 + q, m)] = 0.f; srow[q] = 0.f; ++q; }
            }
        }
        }
        if (job.has_rad) {
            if constexpr (LOADED) {
                // ys_4 (radiance layer 3 output, for the ReLU mask of delta_3) -> A; delta_4 from the forward launch's rgb
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 yv[8];
                    ld_frag4(WP(PL_YS + 3), ty, tx, j, yv);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i][4 * j + c] = f4c(yv[i], c);
                    st_A4(S.A, ty, tx, j, acc);
                }
                if (tid < TM) {
                    const int m = tid;
                    const long long w = tile * TM + m;
                    float d4[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float rgb = w < total ? job.f_rad[w * 3 + c] : 0.f;
                        d4[c] = Q.GRAD[c * TM + m] * rgb * (1.f - rgb);
                        Q.D4[c * TM + m] = d4[c];
                    }
                    *reinterpret_cast<float4*>(st.t(0) + (row0 + m) * 4) = make_float4(d4[0], d4[1], d4[2], 0.f);
                }
                __syncthreads();
            } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 fv[8];
                ld_frag4(WP(PL_FEAT), ty, tx, j, fv);
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i][4 * j + c] = f4c(fv[i], c);
                st_A4(S.A, ty, tx, j, acc);
            }
            for (int layer = 0; layer < 4; ++layer) {
                gemm_tile<4>(acc, pk + L.rad_wt[layer], layer == 0 ? W + spad : W, 0, S.A, S.Ws, tid);
                const float* bias = pk + L.rad_b[layer];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float b = __ldg(bias + 64 * j + 4 * tx + c);
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i][4 * j + c] = fmaxf(acc[i][4 * j + c] + b, 0.f);
                    }
                    st_A4(S.A, ty, tx, j, acc);
                    st_frag4(WP(PL_YS + layer), ty, tx, j, acc);
                }
            }
            __syncthreads();
            narrow_layer<3>(S.A, pk + L.rad_w4, S.RED, tid);
            __syncthreads();
            if (tid < TM) {
                const int m = tid;
                float d4[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float rgb = sigmoidf_(S.RED[c * TM + m] + S.RED[(3 + c) * TM + m] + __ldg(pk + L.rad_b4 + c));
                    d4[c] = Q.GRAD[c * TM + m] * rgb * (1.f - rgb);
                    Q.D4[c * TM + m] = d4[c];
                }
                *reinterpret_cast<float4*>(st.t(0) + (row0 + m) * 4) = make_float4(d4[0], d4[1], d4[2], 0.f);
            }
            __syncthreads();
            }
            // ---- 6. radiance backward: delta_3 from the output layer, then layers 3..1, then the layer-0 inputs ---------
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int k = 64 * j + 4 * tx + c;
                    const float r0 = __ldg(pk + L.rad_w4 + k), r1 = __ldg(pk + L.rad_w4 + 256 + k), r2 = __ldg(pk + L.rad_w4 + 512 + k);
                    float y[8];
                    load_col_A(S.A, k, ty, tx, y);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int m = 8 * ty + i;
                        const float g = Q.D4[m] * r0 + Q.D4[TM + m] * r1 + Q.D4[2 * TM + m] * r2;
                        acc[i][4 * j + c] = y[i] > 0.f ? g : 0.f;
                    }
                }
                st_A4(S.A, ty, tx, j, acc);
                st_frag4(WP(PL_D + 3), ty, tx, j, acc);
            }
            if constexpr (TF32) {
                for (int layer = 3; layer >= 1; --layer) {
                    gemm_tile_tf32(acc2, tp + T.rad_w_r[layer], W, 0, S.A, S.Ws, tid);          // dL/d ys[layer]
                    const float* ysp = WP(PL_YS + layer - 1); float* dp = WP(PL_D + layer - 1);
                    mma_each(acc2, tid, [&](int row, int col, float& v0, float& v1) {
                        const float2 y = ld2(ysp, row, col);
                        v0 = y.x > 0.f ? v0 : 0.f; v1 = y.y > 0.f ? v1 : 0.f;
                        st2(dp, row, col, v0, v1);
                        S.A[a_index(col, row)] = v0; S.A[a_index(col + 1, row)] = v1;
                    });
                }
                gemm_tile_tf32(acc2, tp + T.rad_w_r[0], W, 0, S.A, S.Ws, tid);                  // dL/d feature
                float* fbp = WP(PL_FB);
                mma_each(acc2, tid, [&](int row, int col, float& v0, float& v1) { st2(fbp, row, col, v0, v1); });
            } else {
            for (int layer = 3; layer >= 1; --layer) {
                    gemm_tile<4>(acc, tp + T.rad_w[layer], W, 0, S.A, S.Ws, tid);   // dL/d ys[layer]
    #pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 yv[8];
                        ld_frag4(WP(PL_YS + layer - 1), ty, tx, j, yv);
    #pragma unroll
                        for (int c = 0; c < 4; ++c)
    #pragma unroll
                            for (int i = 0; i < 8; ++i) acc[i][4 * j + c] = f4c(yv[i], c) > 0.f ? acc[i][4 * j + c] : 0.f;
                        st_A4(S.A, ty, tx, j, acc);
                        st_frag4(WP(PL_D + layer - 1), ty, tx, j, acc);
                    }
                }
                gemm_tile<4>(acc, tp + T.rad_w[0], W, 0, S.A, S.Ws, tid);           // dL/d feature
    #pragma unroll
                for (int j = 0; j < 4; ++j) st_frag4(WP(PL_FB), ty, tx, j, acc);
            }
            {
                float acc1[8][4];
                gemm_tile<1>(acc1, tp + T.rad_w0_small, W, 0, S.A, S.Ws, tid);   // dL/d (x | view | nabla): keep nabla
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int k = 4 * tx + c - (3 + nv);
                    if (k >= 0 && k < 3) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) Q.NB[k * TM + 8 * ty + i] = acc1[i][c];
                    }
                }
            }
        }
        __syncthreads();
        // ---- 7. dL/d nabla -> dL/d ge = J * n-bar -> tail rows (v-bar_0) ----------------------------------------------------
        if (tid < TM) {
            const int m = tid;
            const float xs[3] = {S.X[m], S.X[TM + m], S.X[2 * TM + m]};
            float* vrow = st.n(NP_VB0) + (row0 + m) * NLD;
            float nbar[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                nbar[c] = Q.NB[c * TM + m] + Q.GNAB[c * TM + m];
                S.A[a_index(TAIL0 + c, m)] = nbar[c]; vrow[c] = nbar[c];
            }
#pragma unroll
            for (int f = 0; f < 6; ++f) {
                const float fr = (float)(1 << f);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float sn, cs; sincosf(__fmul_rn(xs[c], fr), &sn, &cs);
                    const float a = nbar[c] * fr * cs, b = -nbar[c] * fr * sn;
                    S.A[a_index(TAIL0 + 3 + 6 * f + c, m)] = a;  vrow[3 + 6 * f + c] = a;
                    S.A[a_index(TAIL0 + 6 + 6 * f + c, m)] = b;  vrow[6 + 6 * f + c] = b;
                }
            }
            S.A[a_index(TAIL0 + 39, m)] = 0.f; vrow[39] = 0.f;
        }
        // ---- 8. second-order sweep, layers 0..7 (forward-like): g-bar_i = W_i v-bar_i ---------------------------------------
        if constexpr (TF32) {
            for (int layer = 0; layer < N_SDF_HID; ++layer) {
                gemm_tile_tf32(acc2, tp + T.sdf_wt_r[layer], layer == 0 ? EMB_PAD : W, layer == 0 ? TAIL0 : 0, S.A, S.Ws, tid);
                const float* spp = SPL(layer); float* qp = QP + layer * 256 * TM;
                const float* gp_ = WP(PL_G + layer); float* vbp = WP(PL_VB + layer);
                mma_each(acc2, tid, [&](int row, int col, float& v0, float& v1) {
                    const float2 s2 = ld2(spp, row, col), g2 = ld2(gp_, row, col);
                    st2(qp, row, col, 100.f * v0 * g2.x * (1.f - s2.x), 100.f * v1 * g2.y * (1.f - s2.y));
                    v0 *= s2.x; v1 *= s2.y;
                    if (layer == 3) {
                        if (col >= SKIP_H) v0 = S.A[a_index(TAIL0 + (col - SKIP_H), row)];
                        if (col + 1 >= SKIP_H) v1 = S.A[a_index(TAIL0 + (col + 1 - SKIP_H), row)];
                    }
                    st2(vbp, row, col, v0, v1);
                    S.A[a_index(col, row)] = v0; S.A[a_index(col + 1, row)] = v1;
                });
            }
        } else {
        for (int layer = 0; layer < N_SDF_HID; ++layer) {
                if (layer == 0) gemm_tile<4>(acc, pk + L.sdf_wt[0], EMB_PAD, TAIL0, S.A, S.Ws, tid);
                else            gemm_tile<4>(acc, pk + L.sdf_wt[layer], W, 0, S.A, S.Ws, tid);
                float* qp = QP + layer * 256 * TM;
    #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 sv[8], gv[8];
                    ld_frag4(SPL(layer), ty, tx, j, sv);
                    ld_frag4(WP(PL_G + layer), ty, tx, j, gv);
    #pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int k = 64 * j + 4 * tx + c;
    #pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float gb = acc[i][4 * j + c], s = f4c(sv[i], c), g = f4c(gv[i], c);
                            const float q = 100.f * gb * g * (1.f - s);             // s-bar_i * softplus''(z_i)
                            (c == 0 ? sv[i].x : c == 1 ? sv[i].y : c == 2 ? sv[i].z : sv[i].w) = q;
                            acc[i][4 * j + c] = gb * s;                              // u-bar_i = v-bar_{i+1}
                        }
                        if (layer == 3 && k >= SKIP_H) {
    #pragma unroll
                            for (int i = 0; i < 8; ++i) acc[i][4 * j + c] = S.A[a_index(TAIL0 + (k - SKIP_H), 8 * ty + i)];
                        }
                    }
                    st_frag4v(qp, ty, tx, j, sv);
                    st_A4(S.A, ty, tx, j, acc);
                    st_frag4(WP(PL_VB + layer), ty, tx, j, acc);
                }
            }
        }
        // ---- 9. first-order backward through the trunk: z-bar_7 from the head, then layers 7..1 ----------------------------
        if constexpr (TF32) {
            if (job.has_rad) {
                const float* fbp = WP(PL_FB);
                mma_each(acc2, tid, [&](int row, int col, float& v0, float& v1) {
                    const float2 f2 = ld2(fbp, row, col);
                    S.A[a_index(col, row)] = f2.x; S.A[a_index(col + 1, row)] = f2.y;
                });
                gemm_tile_tf32(acc2, tp + T.w8_feat_r, W, 0, S.A, S.Ws, tid);                   // W8[1:,:]^T feat-bar
            } else {
                mma_each(acc2, tid, [&](int, int, float& v0, float& v1) { v0 = 0.f; v1 = 0.f; });
            }
            for (int layer = 8; layer >= 1; --layer) {
                if (layer < 8) gemm_tile_tf32(acc2, tp + T.sdf_w_r[layer], W, 0, S.A, S.Ws, tid); // h-bar_{layer-1}
                const float* spp = SPL(layer - 1); const float* qp = QP + (layer - 1) * 256 * TM;
                float* zbp = WP(PL_ZB + layer - 1);
                mma_each(acc2, tid, [&](int row, int col, float& v0, float& v1) {
                    const float2 s2 = ld2(spp, row, col), q2 = ld2(qp, row, col);
                    if (layer == 8) {
                        const float gs = Q.GS[row];
                        v0 += gs * __ldg(pk + L.w8_sdf + col); v1 += gs * __ldg(pk + L.w8_sdf + col + 1);
                    }
                    v0 = v0 * s2.x + q2.x; v1 = v1 * s2.y + q2.y;
                    st2(zbp, row, col, v0, v1);
                    S.A[a_index(col, row)] = v0; S.A[a_index(col + 1, row)] = v1;
                });
            }
        } else {
        if (job.has_rad) {
    #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 fv[8];
                    ld_frag4(WP(PL_FB), ty, tx, j, fv);
    #pragma unroll
                    for (int c = 0; c < 4; ++c)
    #pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i][4 * j + c] = f4c(fv[i], c);
                    st_A4(S.A, ty, tx, j, acc);
                }
                gemm_tile<4>(acc, tp + T.w8_feat, W, 0, S.A, S.Ws, tid);             // W8[1:,:]^T feat-bar
            } else {
    #pragma unroll
                for (int i = 0; i < 8; ++i)
    #pragma unroll
                    for (int c = 0; c < 16; ++c) acc[i][c] = 0.f;
            }
            for (int layer = 8; layer >= 1; --layer) {
                if (layer < 8) gemm_tile<4>(acc, pk + L.sdf_w[layer], W, 0, S.A, S.Ws, tid);   // h-bar_{layer-1}
    #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 sv[8], qv[8];
                    ld_frag4(SPL(layer - 1), ty, tx, j, sv);
                    ld_frag4(QP + (layer - 1) * 256 * TM, ty, tx, j, qv);
    #pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float w8 = layer == 8 ? __ldg(pk + L.w8_sdf + 64 * j + 4 * tx + c) : 0.f;
    #pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float hb = acc[i][4 * j + c];
                            if (layer == 8) hb += Q.GS[8 * ty + i] * w8;
                            acc[i][4 * j + c] = hb * f4c(sv[i], c) + f4c(qv[i], c);
                        }
                    }
                    st_A4(S.A, ty, tx, j, acc);
                    st_frag4(WP(PL_ZB + layer - 1), ty, tx, j, acc);
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// weight gradients:  out[l][r] += sum_m L[m][l] * R[m][r]  (+ a second pair), split over samples, atomics into GradPack
// ---------------------------------------------------------------------------------------------------------------------
struct WgradTask {
    const float* L; const float* R; const float* L2; const float* R2;
    float* out;
    int ldl, ldr, nl, nr, ldo;
    int blk0;                   // first block index of this task in grid.x
    int nbr;                    // output blocks along r
};
constexpr int MAX_WTASKS = 32;
struct WgradTable { WgradTask t[MAX_WTASKS]; int n; int total_blocks; };
constexpr int WK = 16;          // samples per smem chunk

__global__ void __launch_bounds__(256)
wgrad_kernel(const WgradTable tab, long long m_total, int rows_per_split) {
    __shared__ __align__(16) float Ls[2][WK][128];
    __shared__ __align__(16) float Rs[2][WK][128];
    int ti = 0;
    while (ti + 1 < tab.n && (int)blockIdx.x >= tab.t[ti + 1].blk0) ++ti;
    const WgradTask t = tab.t[ti];
    const int b = blockIdx.x - t.blk0;
    const int lb = (b / t.nbr) * 128, rb = (b % t.nbr) * 128;
    const long long m0 = (long long)blockIdx.y * rows_per_split;
    long long m1 = m0 + rows_per_split; if (m1 > m_total) m1 = m_total;
    if (m0 >= m1) return;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int npair = t.L2 ? 2 : 1;
    // loader mapping: 256 threads x 2 float4 per operand per chunk: row = (tid*2+e)/32, col4 = (tid*2+e)%32
    float4 lreg[2], rreg[2];
    auto gload = [&](const float* Lp, const float* Rp, long long mm) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            // row-major planes: a warp reads 512 B of one sample
            // (64-byte segments; the shared-memory stores below stay conflict-free: each quarter warp hits 8 distinct 4-bank groups)
            const int idx = tid * 2 + e;
            const int row = idx >> 5;
            const int c4 = (idx & 31) * 4;
            const long long m = mm + row;
            lreg[e] = make_float4(0.f, 0.f, 0.f, 0.f); rreg[e] = lreg[e];
            if (m < m1) {
                if (lb + c4 < t.nl) lreg[e] = *reinterpret_cast<const float4*>(Lp + (size_t)m * t.ldl + lb + c4);
                if (rb + c4 < t.nr) rreg[e] = *reinterpret_cast<const float4*>(Rp + (size_t)m * t.ldr + rb + c4);
            }
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int idx = tid * 2 + e;
            const int row = idx >> 5;
            const int c4 = (idx & 31) * 4;
            *reinterpret_cast<float4*>(&Ls[buf][row][c4]) = lreg[e];
            *reinterpret_cast<float4*>(&Rs[buf][row][c4]) = rreg[e];
        }
    };
    for (int pr = 0; pr < npair; ++pr) {
        const float* Lp = pr ? t.L2 : t.L; const float* Rp = pr ? t.R2 : t.R;
        int buf = 0;
        gload(Lp, Rp, m0);
        sstore(0);
        __syncthreads();
        for (long long mm = m0; mm < m1; mm += WK) {
            const bool more = mm + WK < m1;
            if (more) gload(Lp, Rp, mm + WK);
#pragma unroll
            for (int kk = 0; kk < WK; ++kk) {
                const float4 a0 = *reinterpret_cast<const float4*>(&Ls[buf][kk][8 * ty]);
                const float4 a1 = *reinterpret_cast<const float4*>(&Ls[buf][kk][8 * ty + 4]);
                const float4 b0 = *reinterpret_cast<const float4*>(&Rs[buf][kk][4 * tx]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Rs[buf][kk][64 + 4 * tx]);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
            }
            if (more) sstore(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int l = lb + 8 * ty + i;
        if (l >= t.nl) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int r = rb + (j < 4 ? 4 * tx + j : 64 + 4 * tx + (j - 4));
            if (r < t.nr) atomicAdd(t.out + (size_t)l * t.ldo + r, acc[i][j]);
        }
    }
}

// Tensor-core variant (NA_WGRAD=tf32, default): the same tiling with `mma.sync.m16n8k8` TF32 products and fp32 accumulation.
// out = L^T R: the mma A operand (row l, col k = sample) is Ls[k][l], the B operand (row k, col r) is Rs[k][r] -- both are the
// shared-memory tiles as loaded (row stride 136 floats: the 4 k x 8 column lanes of a fragment load hit 32 distinct banks).
// Operands are rounded to TF32 (cvt.rna) once, when they are staged.  8 warps = 2 (l) x 4 (r); a warp owns 64 x 32 outputs.
// Weight gradients are linear in the upstream gradient, so 10-bit operand mantissas cost ~3e-4 relative error per tensor
// (measured, tests/test_gpu_train.py), inside the parity bound; the fp32 kernel above stays selectable (NA_WGRAD=fp32).
constexpr int WLD = 136;
__global__ void __launch_bounds__(256)
wgrad_tf32_kernel(const WgradTable tab, long long m_total, int rows_per_split) {
    __shared__ __align__(16) float Ls[2][WK][WLD];
    __shared__ __align__(16) float Rs[2][WK][WLD];
    int ti = 0;
    while (ti + 1 < tab.n && (int)blockIdx.x >= tab.t[ti + 1].blk0) ++ti;
    const WgradTask t = tab.t[ti];
    const int b = blockIdx.x - t.blk0;
    const int lb = (b / t.nbr) * 128, rb = (b % t.nbr) * 128;
    const long long m0 = (long long)blockIdx.y * rows_per_split;
    long long m1 = m0 + rows_per_split; if (m1 > m_total) m1 = m_total;
    if (m0 >= m1) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
    const int wl = (warp >> 2) * 64, wr = (warp & 3) * 32;
    float acc[4][4][4];                                   // [l tile of 16][r tile of 8][fragment]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
    const int npair = t.L2 ? 2 : 1;
    float4 lreg[2], rreg[2];
    auto gload = [&](const float* Lp, const float* Rp, long long mm) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            // row-major planes: a warp reads 512 B of one sample
            // (64-byte segments; the shared-memory stores below stay conflict-free: each quarter warp hits 8 distinct 4-bank groups)
            const int idx = tid * 2 + e;
            const int row = idx >> 5;
            const int c4 = (idx & 31) * 4;
            const long long m = mm + row;
            lreg[e] = make_float4(0.f, 0.f, 0.f, 0.f); rreg[e] = lreg[e];
            if (m < m1) {
                if (lb + c4 < t.nl) lreg[e] = *reinterpret_cast<const float4*>(Lp + (size_t)m * t.ldl + lb + c4);
                if (rb + c4 < t.nr) rreg[e] = *reinterpret_cast<const float4*>(Rp + (size_t)m * t.ldr + rb + c4);
            }
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int idx = tid * 2 + e;
            const int row = idx >> 5;
            const int c4 = (idx & 31) * 4;
            *reinterpret_cast<float4*>(&Ls[buf][row][c4]) = make_float4(to_tf32(lreg[e].x), to_tf32(lreg[e].y), to_tf32(lreg[e].z), to_tf32(lreg[e].w));
            *reinterpret_cast<float4*>(&Rs[buf][row][c4]) = make_float4(to_tf32(rreg[e].x), to_tf32(rreg[e].y), to_tf32(rreg[e].z), to_tf32(rreg[e].w));
        }
    };
    for (int pr = 0; pr < npair; ++pr) {
        const float* Lp = pr ? t.L2 : t.L; const float* Rp = pr ? t.R2 : t.R;
        int buf = 0;
        gload(Lp, Rp, m0);
        sstore(0);
        __syncthreads();
        for (long long mm = m0; mm < m1; mm += WK) {
            const bool more = mm + WK < m1;
            if (more) gload(Lp, Rp, mm + WK);
#pragma unroll
            for (int k0 = 0; k0 < WK; k0 += 8) {
                unsigned bf[4][2];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    bf[j][0] = __float_as_uint(Rs[buf][k0 + tq][wr + 8 * j + g]);
                    bf[j][1] = __float_as_uint(Rs[buf][k0 + tq + 4][wr + 8 * j + g]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    unsigned af[4];
                    af[0] = __float_as_uint(Ls[buf][k0 + tq][wl + 16 * i + g]);
                    af[1] = __float_as_uint(Ls[buf][k0 + tq][wl + 16 * i + g + 8]);
                    af[2] = __float_as_uint(Ls[buf][k0 + tq + 4][wl + 16 * i + g]);
                    af[3] = __float_as_uint(Ls[buf][k0 + tq + 4][wl + 16 * i + g + 8]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) mma_tf32(acc[i][j], af, bf[j]);
                }
            }
            if (more) sstore(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const int l = lb + wl + 16 * i + g + (f >> 1) * 8;
                const int r = rb + wr + 8 * j + 2 * tq + (f & 1);
                if (l < t.nl && r < t.nr) atomicAdd(t.out + (size_t)l * t.ldo + r, acc[i][j][f]);
            }
}

struct ColsumTask { const float* P; float* out; int ld, n; };
constexpr int MAX_CTASKS = 24;
struct ColsumTable { ColsumTask t[MAX_CTASKS]; int n; };
__global__ void __launch_bounds__(256)
colsum_kernel(const ColsumTable tab, long long m_total, int rows_per_split) {
    const ColsumTask t = tab.t[blockIdx.x];
    const long long m0 = (long long)blockIdx.y * rows_per_split;
    long long m1 = m0 + rows_per_split; if (m1 > m_total) m1 = m_total;
    const int c = threadIdx.x;
    if (c >= t.n) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    long long m = m0;
    for (; m + 3 < m1; m += 4) {
        s0 += t.P[m * t.ld + c]; s1 += t.P[(m + 1) * t.ld + c]; s2 += t.P[(m + 2) * t.ld + c]; s3 += t.P[(m + 3) * t.ld + c];
    }
    for (; m < m1; ++m) s0 += t.P[m * t.ld + c];
    if (m0 < m1) atomicAdd(t.out + c, (s0 + s1) + (s2 + s3));
}

// ---------------------------------------------------------------------------------------------------------------------
// GradPack: gradient w.r.t. the packed effective weights (floats)
// ---------------------------------------------------------------------------------------------------------------------
struct GradPack {
    size_t sdf_w[8];        // [256 out][ld in]  ld = 40 (layer 0) / 256
    size_t sdf_b[8];        // [256]
    size_t w8_sdf;          // [4][256]  (row 0 used: dL/d W8[0,:] from the sdf output; the u-bar_7 column sums are added to it)
    size_t b8_sdf;          // [4]
    size_t w8_feat;         // [256 feat][256]
    size_t b8_feat;         // [256]
    size_t rad_w0f;         // [256][256] feature columns of radiance layer 0
    size_t rad_w0s;         // [256][40]  small-input columns (x | view | nabla)
    size_t rad_w[4];        // [1..3]: [256][256]
    size_t rad_b[4];        // [256]
    size_t rad_w4;          // [4][256]
    size_t rad_b4;          // [4]
    size_t total;
};
static GradPack grad_layout() {
    GradPack g; size_t o = 0;
    for (int i = 0; i < 8; ++i) { g.sdf_w[i] = o; o += (size_t)W * (i == 0 ? NLD : W); }
    for (int i = 0; i < 8; ++i) { g.sdf_b[i] = o; o += W; }
    g.w8_sdf = o; o += 4 * W;  g.b8_sdf = o; o += 4;
    g.w8_feat = o; o += (size_t)W * W;  g.b8_feat = o; o += W;
    g.rad_w0f = o; o += (size_t)W * W;  g.rad_w0s = o; o += (size_t)W * NLD;
    g.rad_w[0] = o;
    for (int i = 1; i < 4; ++i) { g.rad_w[i] = o; o += (size_t)W * W; }
    for (int i = 0; i < 4; ++i) { g.rad_b[i] = o; o += W; }
    g.rad_w4 = o; o += 4 * W;  g.rad_b4 = o; o += 4;
    g.total = o;
    return g;
}

// one CTA per (output row, layer): gradient of the effective weight row -> (bias, weight_g, weight_v) gradients through
// W[o,:] = g[o] v[o,:] / ||v[o,:]||  (nn.utils.weight_norm, models/base.py:226-227,365-366)
__global__ void __launch_bounds__(128)
unpack_grads_kernel(const NaRawParams raw, const NaRawGrads out, const GradPack G, const float* __restrict__ gp, int sdim) {
    const int layer = blockIdx.y, o = blockIdx.x;
    if (!out.weight_v[layer]) return;
    int n_out, n_in;
    if (layer < 8) { n_out = layer == 3 ? SKIP_H : W; n_in = layer == 0 ? EMB : W; }
    else if (layer == 8) { n_out = W + 1; n_in = W; }
    else if (layer == 9) { n_out = W; n_in = W + sdim; }
    else if (layer < 13) { n_out = W; n_in = W; }
    else { n_out = 3; n_in = W; }
    if (o >= n_out) return;
    auto dW = [&](int i) -> float {
        if (layer < 8) return gp[G.sdf_w[layer] + (size_t)o * (layer == 0 ? NLD : W) + i] * (layer == 4 ? 0.70710678118654752440f : 1.f);
        if (layer == 8) return o == 0 ? gp[G.w8_sdf + i] : gp[G.w8_feat + (size_t)(o - 1) * W + i];
        if (layer == 9) return i < sdim ? gp[G.rad_w0s + (size_t)o * NLD + i] : gp[G.rad_w0f + (size_t)o * W + (i - sdim)];
        if (layer < 13) return gp[G.rad_w[layer - 9] + (size_t)o * W + i];
        return gp[G.rad_w4 + (size_t)o * W + i];
    };
    const float* v = raw.weight_v[layer] + (size_t)o * n_in;
    float dot = 0.f, nn = 0.f;
    for (int i = threadIdx.x; i < n_in; i += blockDim.x) { const float vi = v[i]; dot += dW(i) * vi; nn += vi * vi; }
    __shared__ float red[2][4];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) { dot += __shfl_xor_sync(0xffffffffu, dot, s); nn += __shfl_xor_sync(0xffffffffu, nn, s); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = dot; red[1][threadIdx.x >> 5] = nn; }
    __syncthreads();
    dot = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    nn = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    const float nrm = sqrtf(nn), g = raw.weight_g[layer][o];
    const float dg = dot / nrm;
    float* dv = out.weight_v[layer] + (size_t)o * n_in;
    for (int i = threadIdx.x; i < n_in; i += blockDim.x) dv[i] = (g / nrm) * (dW(i) - dg * v[i] / nrm);
    if (threadIdx.x == 0) {
        if (out.weight_g[layer]) out.weight_g[layer][o] = dg;
        if (out.bias[layer]) {
            float b;
            if (layer < 8) b = gp[G.sdf_b[layer] + o];
            else if (layer == 8) b = o == 0 ? gp[G.b8_sdf] : gp[G.b8_feat + o - 1];
            else if (layer < 13) b = gp[G.rad_b[layer - 9] + o];
            else b = gp[G.rad_b4 + o];
            out.bias[layer][o] = b;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// compositing backward, one warp per ray
// ---------------------------------------------------------------------------------------------------------------------
struct CompBwdArgs {
    const float* d_all;      // [n][P]
    const float* sdf;        // [n][P]    (VolSDF: after the sphere-background override)
    const float* rad;        // [n][P][3] (VolSDF) / [n][P-1][3] (NeuS, at the midpoints)
    const float* nab;        // [n][P][3]
    const float* G;          // [n][3]    dL/d rgb
    const float* scal;       // VolSDF: {alpha, beta};  NeuS: {s}
    float* g_sdf; float* g_rad; float* g_nab;      // [n][P], [n][P or P-1][3], [n][P][3]
    double* accum;           // [0] += dL/d ln_beta (ln_s), [1] += eikonal loss
    int n, P, white;
    float w_eik, inv_count, speed;
};

__device__ __forceinline__ void eik_point(const CompBwdArgs& a, size_t idx, double& loss) {
    const float n0 = a.nab[idx * 3], n1 = a.nab[idx * 3 + 1], n2 = a.nab[idx * 3 + 2];
    const float nn = sqrtf(n0 * n0 + n1 * n1 + n2 * n2);
    const float e = nn - 1.f;
    loss += (double)(a.w_eik * e * e * a.inv_count);
    const float k = nn > 0.f ? 2.f * a.w_eik * a.inv_count * e / nn : 0.f;
    a.g_nab[idx * 3] = k * n0; a.g_nab[idx * 3 + 1] = k * n1; a.g_nab[idx * 3 + 2] = k * n2;
}

__device__ __forceinline__ void block_accum(double* accum, double v0, double v1) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) { v0 += __shfl_xor_sync(0xffffffffu, v0, s); v1 += __shfl_xor_sync(0xffffffffu, v1, s); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(accum, v0); atomicAdd(accum + 1, v1); }
}

// inclusive suffix sum over the warp (lane l gets sum of lanes >= l)
__device__ __forceinline__ double warp_suffix_sum(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_down_sync(0xffffffffu, v, o); if (lane + o < 32) v += t; }
    return v;
}
// inclusive prefix product over the warp
__device__ __forceinline__ double warp_prefix_prod(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v *= t; }
    return v;
}

// VolSDF ray integration (volsdf.py:540-564) differentiated; sigma from sdf_to_sigma (volsdf.py:34-53).  One WARP per ray: the
// transmittance is an exclusive product scan over the samples, the "sum over later samples" of cumprod's backward an exclusive suffix
// sum, both carried across 32-sample chunks in double (like torch's CPU cumprod, which the forward reproduces).
__global__ void __launch_bounds__(256) volsdf_composite_bwd_kernel(const CompBwdArgs a) {
    const int lane = threadIdx.x & 31;
    const int ray = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    double beta_bar = 0.0, loss = 0.0;
    if (ray < a.n) {
        const int P = a.P, M = P - 1;
        const float alpha = a.scal[0], beta = a.scal[1];
        const float* d = a.d_all + (size_t)ray * P; const float* s = a.sdf + (size_t)ray * P;
        float* gs = a.g_sdf + (size_t)ray * P;
        const float G0 = a.G[ray * 3], G1 = a.G[ray * 3 + 1], G2 = a.G[ray * 3 + 2];
        const float Gsum = a.white ? (G0 + G1 + G2) : 0.f;
        double carry = 1.0;
        for (int base = 0; base < M; base += 32) {               // T_i = prod_{j<i} p_j, parked in g_sdf
            const int i = base + lane;
            float p = 1.f;
            if (i < M) {
                const float e = 0.5f * expf(-fabsf(s[i]) / beta);
                const float sigma = alpha * (s[i] >= 0.f ? e : 1.f - e);
                p = expf(-fmaxf(sigma * (d[i + 1] - d[i]), 0.f));
            }
            const double incl = warp_prefix_prod((double)p, lane);
            double excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 1.0;
            if (i < M) gs[i] = (float)(carry * excl);
            carry *= __shfl_sync(0xffffffffu, incl, 31);
        }
        double S_hi = 0.0;                                       // sum_{k >= base + 32} tau-bar_k tau_k
        for (int base = ((M - 1) >> 5) << 5; base >= 0; base -= 32) {
            const int i = base + lane;
            const bool live = i < M;
            float Ti = 0.f, si = 0.f, e = 0.f, psi = 0.f, delta = 0.f, x = 0.f, p = 1.f, tau_bar = 0.f;
            double v = 0.0;
            size_t ci = 0;
            if (live) {
                Ti = gs[i]; si = s[i];
                e = 0.5f * expf(-fabsf(si) / beta);
                psi = si >= 0.f ? e : 1.f - e;
                delta = d[i + 1] - d[i];
                x = alpha * psi * delta;
                p = expf(-fmaxf(x, 0.f));
                const float tau = (1.f - p + 1e-10f) * Ti;
                ci = ((size_t)ray * P + i) * 3;
                tau_bar = a.rad[ci] * G0 + a.rad[ci + 1] * G1 + a.rad[ci + 2] * G2 - Gsum;
                a.g_rad[ci] = tau * G0; a.g_rad[ci + 1] = tau * G1; a.g_rad[ci + 2] = tau * G2;
                v = (double)tau_bar * tau;
            }
            const double suf = warp_suffix_sum(v, lane);
            const double S = S_hi + (suf - v);                   // sum_{k>i} tau-bar_k tau_k
            S_hi += __shfl_sync(0xffffffffu, suf, 0);
            if (live) {
                const float x_bar = x > 0.f ? (float)((double)tau_bar * Ti * p - S) : 0.f;
                const float sigma_bar = x_bar * delta;
                gs[i] = si != 0.f ? sigma_bar * (-(alpha / beta) * e) : 0.f;
                const float dpsi_dbeta = (si >= 0.f ? 1.f : -1.f) * e * fabsf(si) / (beta * beta);
                beta_bar += (double)(sigma_bar * (psi * (-1.f / (beta * beta)) + alpha * dpsi_dbeta));
            }
        }
        if (lane == 0) {
            gs[P - 1] = 0.f;
            for (int c = 0; c < 3; ++c) a.g_rad[((size_t)ray * P + P - 1) * 3 + c] = 0.f;
        }
        if (a.w_eik != 0.f) for (int i = lane; i < P; i += 32) eik_point(a, (size_t)ray * P + i, loss);
        beta_bar *= (double)(a.speed * beta);                    // beta = exp(ln_beta * speed_factor), volsdf.py:337-339
    }
    block_accum(a.accum, beta_bar, loss);
}

// NeuS ray integration (neus.py:36-43,65-78,373-381) differentiated, one warp per ray.  alpha_i = max((Phi_i - Phi_{i+1}) / (Phi_i +
// 1e-10), 0) couples sample i to its neighbour: d L / d Phi_{i+1} = A_{i+1} - a-bar_i / (Phi_i + 1e-10), with
// A_i = a-bar_i (Phi_{i+1} + 1e-10) / (Phi_i + 1e-10)^2 the contribution of alpha_i to its own leading cdf value.
__global__ void __launch_bounds__(256) neus_composite_bwd_kernel(const CompBwdArgs a) {
    const int lane = threadIdx.x & 31;
    const int ray = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    double s_bar = 0.0, loss = 0.0;
    if (ray < a.n) {
        const int P = a.P, M = P - 1;
        const float sc = a.scal[0];
        const float* sd = a.sdf + (size_t)ray * P;
        float* gs = a.g_sdf + (size_t)ray * P;
        const float G0 = a.G[ray * 3], G1 = a.G[ray * 3 + 1], G2 = a.G[ray * 3 + 2];
        const float Gsum = a.white ? (G0 + G1 + G2) : 0.f;
        auto Phi = [&](int i) { return __fdiv_rn(1.f, 1.f + expf(-sd[i] * sc)); };
        double carry = 1.0;
        for (int base = 0; base < M; base += 32) {
            const int i = base + lane;
            float qf = 1.f;
            if (i < M) {
                const float c0 = Phi(i), c1 = Phi(i + 1);
                qf = 1.f - fmaxf((c0 - c1) / (c0 + 1e-10f), 0.f) + 1e-10f;
            }
            const double incl = warp_prefix_prod((double)qf, lane);
            double excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 1.0;
            if (i < M) gs[i] = (float)(carry * excl);
            carry *= __shfl_sync(0xffffffffu, incl, 31);
        }
        double S_hi = 0.0;
        float A_hi = 0.f;                                        // A of sample base + 32 (0 beyond the last sample)
        for (int base = ((M - 1) >> 5) << 5; base >= 0; base -= 32) {
            const int i = base + lane;
            const bool live = i < M;
            float Ti = 0.f, c0 = 1.f, c1 = 0.f, raw = 0.f, q = 1.f, w_bar = 0.f;
            double v = 0.0;
            if (live) {
                Ti = gs[i];
                c0 = Phi(i); c1 = Phi(i + 1);
                raw = (c0 - c1) / (c0 + 1e-10f);
                const float al = fmaxf(raw, 0.f);
                q = 1.f - al + 1e-10f;
                const float w = al * Ti;
                const size_t ci = ((size_t)ray * (P - 1) + i) * 3;
                w_bar = a.rad[ci] * G0 + a.rad[ci + 1] * G1 + a.rad[ci + 2] * G2 - Gsum;
                a.g_rad[ci] = w * G0; a.g_rad[ci + 1] = w * G1; a.g_rad[ci + 2] = w * G2;
                v = (double)w_bar * w;
            }
            const double suf = warp_suffix_sum(v, lane);
            const double S = S_hi + (suf - v);
            S_hi += __shfl_sync(0xffffffffu, suf, 0);
            const float a_bar = (live && raw >= 0.f) ? (float)((double)w_bar * Ti - S / (double)q) : 0.f;
            const float A = live ? a_bar * (c1 + 1e-10f) / ((c0 + 1e-10f) * (c0 + 1e-10f)) : 0.f;
            float A_next = __shfl_down_sync(0xffffffffu, A, 1);
            if (lane == 31) A_next = A_hi;
            A_hi = __shfl_sync(0xffffffffu, A, 0);
            __syncwarp();                                        // every lane has read its T_i before g_sdf[i + 1] is overwritten
            if (live) {
                const float pre = (A_next - a_bar / (c0 + 1e-10f)) * c1 * (1.f - c1);
                gs[i + 1] = pre * sc;
                s_bar += (double)(pre * sd[i + 1]);
                if (i == 0) {
                    const float pre0 = A * c0 * (1.f - c0);
                    gs[0] = pre0 * sc;
                    s_bar += (double)(pre0 * sd[0]);
                }
            }
        }
        if (a.w_eik != 0.f) for (int i = lane; i < P; i += 32) eik_point(a, (size_t)ray * P + i, loss);
        s_bar *= (double)(a.speed * sc);                         // s = exp(ln_s * speed_factor), neus.py:116-117
    }
    block_accum(a.accum, s_bar, loss);
}

__global__ void normalize_dirs_train_kernel(const float* __restrict__ d, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = d[i * 3], b = d[i * 3 + 1], c = d[i * 3 + 2];
    // F.normalize, volsdf.py:442 -- the same un-fused operations as the forward render's normalize_dirs_kernel (csrc/volsdf_render.cu):
    // the re-evaluated sample positions are then bit-identical to the render's (a contracted a*a + b*b + c*c differs in the last bit,
    // which moves ReLU / softplus' decisions of a few samples)
    const float nrm = fmaxf(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c))), 1e-12f);
    out[i * 3] = __fdiv_rn(a, nrm); out[i * 3 + 1] = __fdiv_rn(b, nrm); out[i * 3 + 2] = __fdiv_rn(c, nrm);
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }
// Workspace of one render_bwd call.  fp32 mode: the fp32 stash written by mlp_bwd_kernel + its per-CTA scratch.  Tensor-core modes:
// the 16-bit stash written by the BW program of csrc/mlp_tmem.cu (42 wide planes of 512 B per sample + 3 narrow ones), its outputs
// and its per-CTA scratch.
struct TrainWs {
    float* dirs; float* g_sdf; float* g_nab; float* g_rad;
    float* stash; float* scratch;                                                   // fp32 mode
    unsigned short* wide16; unsigned short* narrow16; float* tiny; size_t mpad;     // tensor-core modes
    float* f_sdf; float* f_rad; float* fwd_scratch; size_t fwd_scratch_bytes;
    unsigned char* tile_buf;                                                        // split program: per-tile softplus' codes + ReLU masks
    size_t total;
};
size_t mlp_tmem_tile_buf_bytes(long long n_samples);                                // csrc/mlp_tmem.cu
size_t mlp_scratch_bytes();                                                         // csrc/api.cu
int launch_wgrad_f16(const WgF16Task* tasks, int n_tasks, const unsigned short* wide, int n_wide_planes, const unsigned short* narrow,
                     int n_narrow_planes, long long mpad, long long m_rows, cudaStream_t stream);       // csrc/wgrad_f16.cu
int make_stash_store_map(TmaMap* out, const unsigned short* wide, int n_wide_planes, long long mpad);
int launch_wgrad_tiny(const unsigned short* in7, const unsigned short* vb7, const unsigned short* ys3, const float* t0, const float* t1,
                      float* w8_sdf, float* b8_sdf, float* rad_w4, float* rad_b4, long long m_rows, int fwd_bf16, cudaStream_t stream);
int launch_mlp(const EvalJob& job, const void* packed, int precision, float* scratch, size_t scratch_bytes, cudaStream_t stream);
static bool bwd_on_tensor_cores(int precision) {
    static const bool fp32_bwd = [] { const char* e = getenv("NA_BWD"); return e && strcmp(e, "fp32") == 0; }();
    static const bool force_recompute = [] { const char* e = getenv("NA_BWD_RECOMPUTE"); return e && e[0] == '1'; }();
    return precision != NA_PRECISION_FP32 && !fp32_bwd && !force_recompute;
}
static TrainWs train_ws(void* base, long long n_rays, int P, bool tc) {
    const size_t M = (size_t)n_rays * P, mpad = (M + TM - 1) / TM * TM;
    unsigned char* p = (unsigned char*)base; size_t o = 0;
    TrainWs w = {};
    auto take = [&](size_t bytes) { float* r = (float*)(p + o); o += align256(bytes); return r; };
    w.mpad = mpad;
    w.dirs = take((size_t)n_rays * 3 * 4);
    w.g_sdf = take(M * 4); w.g_nab = take(M * 12); w.g_rad = take(M * 12);
    if (tc) {
        w.wide16 = (unsigned short*)take((size_t)N_WIDE * mpad * 256 * 2);
        w.narrow16 = (unsigned short*)take((size_t)3 * mpad * ST_NLD * 2);
        w.tiny = take((size_t)2 * mpad * 4 * 4);
        w.f_sdf = take(M * 4); w.f_rad = take(M * 12);
        w.fwd_scratch_bytes = mlp_scratch_bytes(); w.fwd_scratch = take(w.fwd_scratch_bytes);
        w.tile_buf = (unsigned char*)take(mlp_tmem_tile_buf_bytes((long long)M));
    } else {
        w.stash = take(stash_floats(mpad) * 4);
        w.scratch = take((size_t)num_sms() * 16 * 256 * TM * 4);
    }
    w.total = o;
    return w;
}

// fp32 mode (cfg.precision == NA_PRECISION_FP32, or NA_BWD_RECOMPUTE=1 / NA_BWD=fp32): mlp_bwd_kernel recomputes the forward pass of every
// tile on the fp32 FMA pipe and runs the backward-data GEMMs on mma.sync TF32 (NA_BWD=fp32: FFMA too), leaving an fp32 stash.
static int launch_mlp_bwd(const BwdJob& job, const float* pk, const PackF32& L, const float* tp, const PackTrain& T, const Stash& st,
                          const TrainWs& w, cudaStream_t stream) {
    static bool attr_set[64] = {false};
    static const bool fp32_bwd = [] { const char* e = getenv("NA_BWD"); return e && strcmp(e, "fp32") == 0; }();
    const size_t smem = sizeof(TrainSmem);
    int dev = 0; cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        NA_TRY(check_cuda(cudaFuncSetAttribute(mlp_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        NA_TRY(check_cuda(cudaFuncSetAttribute(mlp_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        attr_set[dev] = true;
    }
    const long long total = (long long)job.n_rows * job.P;
    if (total <= 0) return NA_OK;
    const long long tiles = (total + TM - 1) / TM;
    const int grid = (int)(tiles < (long long)num_sms() ? tiles : (long long)num_sms());
    if (fp32_bwd) mlp_bwd_kernel<false><<<grid, NT, smem, stream>>>(job, pk, L, tp, T, st, w.scratch);
    else          mlp_bwd_kernel<true><<<grid, NT, smem, stream>>>(job, pk, L, tp, T, st, w.scratch);
    NA_CHECK_LAUNCH();
    return NA_OK;
}

// Tensor-core modes: ONE tcgen05 launch per patch (BW program of csrc/mlp_tmem.cu: forward re-evaluation with three-product operands,
// then the 20 backward-data GEMMs), which leaves every (delta, input) pair in the bf16 stash; then the weight gradients
// (csrc/wgrad_f16.cu).
// stash planes / tables of the training workspace -> the fields of an EvalJob
static int attach_stash(EvalJob& e, const TrainWs& w) {
    const size_t mpad = w.mpad;
    unsigned short* nar = w.narrow16;
    e.st_wide = w.wide16; e.st_mpad = mpad;
    NA_TRY(make_stash_store_map(&e.st_store_map, w.wide16, N_WIDE, (long long)mpad));
    e.st_emb = nar + (size_t)NP_EMB * mpad * ST_NLD; e.st_vb0 = nar + (size_t)NP_VB0 * mpad * ST_NLD; e.st_small = nar + (size_t)NP_SMALL * mpad * ST_NLD;
    e.st_t0 = w.tiny; e.st_t1 = w.tiny + mpad * 4;
    e.tile_buf = w.tile_buf;
    return NA_OK;
}

// bytes of the tensor-core training workspace of one launch of n_rays x P samples (NeuS keeps two: points and midpoints)
size_t train_ws_total(long long n_rays, int P) { return train_ws(nullptr, n_rays, P, true).total; }

// Forward half of the split training program: the final full evaluation of the patch's forward render (csrc/volsdf_render.cu), run
// as program 0..20 with the forward stash planes written and the softplus' codes / ReLU masks persisted per tile.  `fj` is the
// render's own job (its sdf / radiance / nabla outputs, apply_bg); the backward half is na_volsdf_render_bwd_stashed.
int train_forward_stash(const EvalJob& fj, const void* packed, int precision, void* train_workspace, size_t train_ws_bytes,
                        float* scratch, size_t scratch_bytes, cudaStream_t stream) {
    if (precision != NA_PRECISION_TC && precision != NA_PRECISION_TC_MIXED) return NA_ERR_UNSUPPORTED;
    if (!train_workspace || fj.x || !fj.want_full || fj.row_ids || fj.n_rows_dev) return NA_ERR_BAD_ARG;
    const TrainWs w = train_ws(train_workspace, fj.n_rows, fj.P, true);
    if (train_ws_bytes < w.total) return NA_ERR_WORKSPACE;
    EvalJob e = fj;
    NA_TRY(attach_stash(e, w));
    e.bw = 0; e.bw_split = 1;
    return launch_mlp(e, packed, precision, scratch, scratch_bytes, stream);
}

// split != 0: the forward half already ran (train_forward_stash); job.f_rad = its radiance output
static int tc_backward(const BwdJob& job, const void* packed, int precision, const TrainWs& w, int train_surface, int train_radiance,
                       float* gp, cudaStream_t stream, int split = 0) {
    const int fwd_bf16 = 1;
    const long long M = (long long)job.n_rows * job.P;
    if (M <= 0) return NA_OK;
    const size_t mpad = w.mpad;
    unsigned short* nar = w.narrow16;
    auto wp = [&](int p) { return w.wide16 + (size_t)p * mpad * 256; };
    EvalJob e = {};
    e.rays_o = job.rays_o; e.rays_d = job.rays_d; e.n_rows = job.n_rows; e.P = job.P;
    e.t = job.t; e.t_stride = job.t_stride; e.t_off = 0; e.midpoints = job.midpoints;
    e.o_stride = job.P; e.o_off = 0;
    e.sdf = split ? nullptr : w.f_sdf;
    e.rad = !job.has_rad ? nullptr : (split ? const_cast<float*>(job.f_rad) : w.f_rad);      // split: the forward launch's radiance, an input
    e.apply_bg = 0;                                   // raw network sdf: the background mask compares it with R - |x| itself
    e.bound_r = job.bound_r; e.want_full = 1; e.multires_view = job.multires_view;
    NA_TRY(attach_stash(e, w));
    e.bw = 1; e.bw_bg_mask = job.apply_bg; e.bw_split = split ? 2 : 0;
    if (split && job.has_rad && !job.f_rad) return NA_ERR_BAD_ARG;
    e.bw_gsdf = job.g_sdf; e.bw_gnab = job.g_nab; e.bw_grad = job.g_rad;
    // the forward re-evaluation runs in the render's own mode: in tc_mixed the SDF forward pass (what softplus' and the activations come
    // from) keeps its three-product operands, the feature head / reverse sweep / radiance layers use single products exactly as in the
    // patch's forward render (NA_BW_FWD=tc: three products everywhere, the round-2 behaviour)
    static const char* bw_fwd_env = getenv("NA_BW_FWD");
    const bool force_tc = bw_fwd_env && bw_fwd_env[0] == 't' && bw_fwd_env[1] == 'c' && bw_fwd_env[2] == 0;
    const int fwd_precision = precision == NA_PRECISION_TC2ACC ? NA_PRECISION_TC : ((precision == NA_PRECISION_TC_MIXED && force_tc) ? NA_PRECISION_TC : precision);
    NA_TRY(launch_mlp(e, packed, fwd_precision, w.fwd_scratch, w.fwd_scratch_bytes, stream));

    const GradPack G = grad_layout();
    WgF16Task tk[16]; int n = 0;
    // L = a backward plane (bf16) or g_i (forward), R = the matching forward plane or v-bar (bf16)
    auto add = [&](int l0, int l0_bf, int r0, int r0_bf, int l1, int l1_bf, int r1, int r1_bf, int npair, int r_narrow, int n_valid,
                   size_t out, int ldo, size_t bias, bool has_bias) {
        WgF16Task& t = tk[n++];
        t.l_plane[0] = l0; t.r_plane[0] = r0; t.l_plane[1] = l1; t.r_plane[1] = r1;
        t.l_bf16[0] = l0_bf; t.r_bf16[0] = r0_bf; t.l_bf16[1] = l1_bf; t.r_bf16[1] = r1_bf;
        t.npair = npair; t.r_narrow = r_narrow; t.n_valid = n_valid; t.out = gp + out; t.ldo = ldo; t.bias_out = has_bias ? gp + bias : nullptr;
    };
    const int F = fwd_bf16;
    const int sdim = small_dim(job.multires_view);
    if (train_surface) {
        // layer 0: [z-bar_0 x emb] + [g_0 x v-bar_0] (39-wide right operands); bias_0 = colsum(z-bar_0)
        add(PL_ZB + 0, 1, NP_EMB, F, PL_G + 0, F, NP_VB0, 1, 2, 1, EMB, G.sdf_w[0], NLD, G.sdf_b[0], true);
        for (int i = 1; i < 8; ++i)
            add(PL_ZB + i, 1, PL_IN + i - 1, F, PL_G + i, F, PL_VB + i - 1, 1, 2, 0, 256, G.sdf_w[i], 256, G.sdf_b[i], true);
        if (job.has_rad) add(PL_FB, 1, PL_IN + 7, F, 0, 0, 0, 0, 1, 0, 256, G.w8_feat, 256, G.b8_feat, true);
    }
    if (train_radiance && job.has_rad) {
        add(PL_D + 0, 1, PL_FEAT, F, 0, 0, 0, 0, 1, 0, 256, G.rad_w0f, 256, G.rad_b[0], true);
        add(PL_D + 0, 1, NP_SMALL, F, 0, 0, 0, 0, 1, 1, sdim, G.rad_w0s, NLD, 0, false);
        for (int l = 1; l < 4; ++l) add(PL_D + l, 1, PL_YS + l - 1, F, 0, 0, 0, 0, 1, 0, 256, G.rad_w[l], 256, G.rad_b[l], true);
    }
    NA_TRY(launch_wgrad_f16(tk, n, w.wide16, N_WIDE, nar, 3, (long long)mpad, M, stream));
    // rows whose left operand is a tiny fp32 plane: SDF head row 0 (+ the u-bar_7 column sums), radiance output layer
    NA_TRY(launch_wgrad_tiny(train_surface ? wp(PL_IN + 7) : nullptr, wp(PL_VB + 7), (train_radiance && job.has_rad) ? wp(PL_YS + 3) : nullptr,
                             e.st_t0, e.st_t1, gp + G.w8_sdf, gp + G.b8_sdf, gp + G.rad_w4, gp + G.rad_b4, M, fwd_bf16, stream));
    return NA_OK;
}

// weight-gradient GEMMs + bias column sums of one mlp_bwd launch
static int launch_wgrad(const Stash& st, long long m_rows, int has_rad, int train_surface, int train_radiance, float* gp, cudaStream_t stream) {
    const GradPack G = grad_layout();
    WgradTable wt; wt.n = 0; wt.total_blocks = 0;
    ColsumTable ct; ct.n = 0;
    static const int wgrad_mode = [] { const char* e = getenv("NA_WGRAD"); return !e ? 0 : (strcmp(e, "fp32") == 0 ? 2 : (strcmp(e, "mma") == 0 ? 1 : 0)); }();
    auto addw = [&](const float* Lp, int ldl, int nl, const float* Rp, int ldr, int nr, const float* L2, const float* R2, size_t out, int ldo) {
        WgradTask& t = wt.t[wt.n++];
        t.L = Lp; t.R = Rp; t.L2 = L2; t.R2 = R2; t.out = gp + out; t.ldl = ldl; t.ldr = ldr; t.nl = nl; t.nr = nr; t.ldo = ldo;
        t.blk0 = wt.total_blocks; t.nbr = (nr + 127) / 128;
        wt.total_blocks += ((nl + 127) / 128) * t.nbr;
    };
    auto addc = [&](const float* P, int ld, int n, size_t out) { ColsumTask& t = ct.t[ct.n++]; t.P = P; t.ld = ld; t.n = n; t.out = gp + out; };
    if (train_surface) {
        addw(st.w(PL_ZB + 0), 256, 256, st.n(NP_EMB), NLD, NLD, st.w(PL_G + 0), st.n(NP_VB0), G.sdf_w[0], NLD);
        for (int i = 1; i < 8; ++i)
            addw(st.w(PL_ZB + i), 256, 256, st.w(PL_IN + i - 1), 256, 256, st.w(PL_G + i), st.w(PL_VB + i - 1), G.sdf_w[i], 256);
        for (int i = 0; i < 8; ++i) addc(st.w(PL_ZB + i), 256, 256, G.sdf_b[i]);
        addw(st.t(1), 4, 4, st.w(PL_IN + 7), 256, 256, nullptr, nullptr, G.w8_sdf, 256);       // g_sdf^T h_7  -> row 0
        addc(st.w(PL_VB + 7), 256, 256, G.w8_sdf);                                              // + sum u-bar_7
        addc(st.t(1), 4, 4, G.b8_sdf);
        if (has_rad) {
            addw(st.w(PL_FB), 256, 256, st.w(PL_IN + 7), 256, 256, nullptr, nullptr, G.w8_feat, 256);
            addc(st.w(PL_FB), 256, 256, G.b8_feat);
        }
    }
    if (train_radiance && has_rad) {
        addw(st.w(PL_D + 0), 256, 256, st.w(PL_FEAT), 256, 256, nullptr, nullptr, G.rad_w0f, 256);
        addw(st.w(PL_D + 0), 256, 256, st.n(NP_SMALL), NLD, NLD, nullptr, nullptr, G.rad_w0s, NLD);
        for (int l = 1; l < 4; ++l) addw(st.w(PL_D + l), 256, 256, st.w(PL_YS + l - 1), 256, 256, nullptr, nullptr, G.rad_w[l], 256);
        addw(st.t(0), 4, 4, st.w(PL_YS + 3), 256, 256, nullptr, nullptr, G.rad_w4, 256);
        for (int l = 0; l < 4; ++l) addc(st.w(PL_D + l), 256, 256, G.rad_b[l]);
        addc(st.t(0), 4, 4, G.rad_b4);
    }
    const long long tiles = (m_rows + TM - 1) / TM;
    if (wt.n == 0 && ct.n == 0) return NA_OK;
    int splits = (int)((4LL * num_sms() + wt.total_blocks - 1) / wt.total_blocks);
    if (splits > tiles) splits = (int)tiles;
    if (splits < 1) splits = 1;
    const int rows_per_split = (int)(((tiles + splits - 1) / splits) * TM);
    const long long m_total = tiles * TM;
    const bool fp32_wgrad = wgrad_mode == 2;
    if (wt.n == 0) {}
    else if (fp32_wgrad) wgrad_kernel<<<dim3(wt.total_blocks, splits), 256, 0, stream>>>(wt, m_total, rows_per_split);
    else            wgrad_tf32_kernel<<<dim3(wt.total_blocks, splits), 256, 0, stream>>>(wt, m_total, rows_per_split);
    NA_CHECK_LAUNCH();
    int csplits = (int)(tiles < 64 ? tiles : 64);
    const int crows = (int)(((tiles + csplits - 1) / csplits) * TM);
    colsum_kernel<<<dim3(ct.n, csplits), 256, 0, stream>>>(ct, m_total, crows);
    NA_CHECK_LAUNCH();
    return NA_OK;
}

int preload_train() {
    NA_PRELOAD((mlp_bwd_kernel<true>));
    NA_PRELOAD((mlp_bwd_kernel<false>));
    NA_PRELOAD(wgrad_kernel);
    NA_PRELOAD(wgrad_tf32_kernel);
    NA_PRELOAD(colsum_kernel);
    NA_PRELOAD(unpack_grads_kernel);
    NA_PRELOAD(volsdf_composite_bwd_kernel);
    NA_PRELOAD(neus_composite_bwd_kernel);
    NA_PRELOAD(normalize_dirs_train_kernel);
    return NA_OK;
}

}  // namespace na

using namespace na;

extern "C" size_t na_grad_pack_bytes(const NaNetDesc* desc) { (void)desc; return grad_layout().total * sizeof(float); }

extern "C" size_t na_train_workspace_bytes_mode(const NaNetDesc* desc, int64_t n_rays, int32_t points_per_ray, int32_t precision) {
    (void)desc;
    if (n_rays <= 0 || points_per_ray <= 1) return 0;
    const bool tc = bwd_on_tensor_cores(precision);
    size_t b = train_ws(nullptr, n_rays, points_per_ray, tc).total;
    // NeuS, split program: the stash of the P - 1 midpoints (radiance pass) lives behind that of the P points (sdf / nabla pass)
    if (tc && desc && desc->framework == NA_FRAMEWORK_NEUS && points_per_ray > 2) b += train_ws(nullptr, n_rays, points_per_ray - 1, tc).total;
    return b;
}
extern "C" int na_debug_wgrad_f16(const void* planes16, int64_t m_pad, int64_t m_rows, int l_bf16, int r_bf16, float* out, float* bias_out, void* stream) {
    if (!planes16 || !out || m_pad <= 0 || m_rows <= 0 || m_rows > m_pad) return NA_ERR_BAD_ARG;
    WgF16Task t = {};
    t.l_plane[0] = 0; t.r_plane[0] = 1; t.l_bf16[0] = l_bf16; t.r_bf16[0] = r_bf16; t.npair = 1; t.r_narrow = 0; t.n_valid = 256; t.ldo = 256;
    t.out = out; t.bias_out = bias_out;
    return launch_wgrad_f16(&t, 1, (const unsigned short*)planes16, 2, nullptr, 0, m_pad, m_rows, (cudaStream_t)stream);
}
extern "C" size_t na_train_workspace_bytes(const NaNetDesc* desc, int64_t n_rays, int32_t points_per_ray) {
    const size_t a = na_train_workspace_bytes_mode(desc, n_rays, points_per_ray, NA_PRECISION_FP32);
    const size_t b = na_train_workspace_bytes_mode(desc, n_rays, points_per_ray, NA_PRECISION_TC);
    return a > b ? a : b;
}

static int render_bwd(const NaNetDesc* desc, const void* packed, const NaTrainCfg* cfg, const float* rays_o, const float* rays_d,
                      int64_t n, const float* scal, const float* d_all, const float* sdf, const float* rad, const float* nab,
                      const float* grad_rgb, float* gp, double* accum, void* ws_, size_t ws_bytes, cudaStream_t stream, bool neus,
                      bool stashed = false) {
    if (!desc || !packed || !cfg || !rays_o || !rays_d || !scal || !d_all || !sdf || !rad || !nab || !grad_rgb || !gp || !accum || !ws_)
        return NA_ERR_BAD_ARG;
    if (n <= 0) return NA_OK;
    const int P = cfg->points_per_ray;
    if (P < 2 || n * (int64_t)P > 0x7fffffffLL) return NA_ERR_UNSUPPORTED;
    const bool tc = bwd_on_tensor_cores(cfg->precision);
    if (stashed && (!tc || (cfg->precision != NA_PRECISION_TC && cfg->precision != NA_PRECISION_TC_MIXED))) return NA_ERR_UNSUPPORTED;
    const TrainWs w = train_ws(ws_, n, P, tc);
    // NeuS split program: second workspace (midpoints) behind the first
    const TrainWs wB = (stashed && neus) ? train_ws((unsigned char*)ws_ + w.total, n, P - 1, tc) : w;
    if (ws_bytes < w.total + ((stashed && neus) ? wB.total : 0)) return NA_ERR_WORKSPACE;
    const PackF32 L = pack_layout_f32(desc->multires_view);
    const PackTrain T = pack_layout_train();
    const float* pk = (const float*)packed;
    const float* tp = (const float*)((const unsigned char*)packed + train_pack_off(desc->multires_view));
    normalize_dirs_train_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(rays_d, w.dirs, (int)n);
    NA_CHECK_LAUNCH();
    CompBwdArgs a;
    a.d_all = d_all; a.sdf = sdf; a.rad = rad; a.nab = nab; a.G = grad_rgb; a.scal = scal;
    a.g_sdf = w.g_sdf; a.g_rad = w.g_rad; a.g_nab = w.g_nab; a.accum = accum; a.n = (int)n; a.P = P; a.white = cfg->white_bkgd;
    a.w_eik = cfg->w_eikonal; a.inv_count = cfg->eikonal_count > 0 ? 1.f / (float)cfg->eikonal_count : 0.f; a.speed = cfg->speed_factor;
    if (neus) neus_composite_bwd_kernel<<<(int)((n + 7) / 8), 256, 0, stream>>>(a);          // one warp per ray
    else      volsdf_composite_bwd_kernel<<<(int)((n + 7) / 8), 256, 0, stream>>>(a);
    NA_CHECK_LAUNCH();
    const size_t M = (size_t)n * P, mpad = (M + TM - 1) / TM * TM;
    Stash st; st.mpad = mpad; st.wide = w.stash; st.narrow = w.stash + (size_t)N_WIDE * mpad * 256; st.tiny = st.narrow + 3 * mpad * NLD;      // fp32 mode only
    auto backward = [&](const BwdJob& j, long long rows, int ts, int tr, bool second = false) -> int {
        if (tc) return tc_backward(j, packed, cfg->precision, second ? wB : w, ts, tr, gp, stream, stashed ? 1 : 0);
        NA_TRY(launch_mlp_bwd(j, pk, L, tp, T, st, w, stream));
        return launch_wgrad(st, rows, j.has_rad, ts, tr, gp, stream);
    };
    BwdJob job = {};
    job.rays_o = rays_o; job.rays_d = w.dirs; job.n_rows = (int)n; job.t = d_all; job.t_stride = P;
    job.multires_view = desc->multires_view; job.bound_r = desc->bounding_radius;
    const bool has_eik = cfg->w_eikonal != 0.f;
    if (!neus) {
        job.P = P; job.midpoints = 0; job.g_sdf = w.g_sdf; job.g_nab = has_eik ? w.g_nab : nullptr; job.g_rad = w.g_rad;
        job.apply_bg = 1; job.has_rad = 1; job.f_rad = rad;
        NA_TRY(backward(job, (long long)M, cfg->train_surface, cfg->train_radiance));
    } else {
        // pass A: the P points of d_all (sdf -> alpha, nabla -> eikonal); pass B: the P-1 midpoints (radiance), neus.py:320-324
        job.P = P; job.midpoints = 0; job.g_sdf = w.g_sdf; job.g_nab = has_eik ? w.g_nab : nullptr; job.g_rad = nullptr;
        job.apply_bg = 0; job.has_rad = 0;
        const char* dbg_pass = getenv("NA_BWD_DEBUG_PASS");                 // diagnostics: "A" / "B" runs only that pass
        if (cfg->train_surface && !(dbg_pass && dbg_pass[0] == 'B')) {
            NA_TRY(backward(job, (long long)M, 1, 0));
        }
        if (dbg_pass && dbg_pass[0] == 'A') return NA_OK;
        job.P = P - 1; job.midpoints = 1; job.g_sdf = nullptr; job.g_nab = nullptr; job.g_rad = w.g_rad; job.has_rad = 1; job.f_rad = rad;
        NA_TRY(backward(job, (long long)n * (P - 1), cfg->train_surface, cfg->train_radiance, true));
    }
    return NA_OK;
}

extern "C" int na_volsdf_render_bwd(const NaNetDesc* desc, const void* packed, const NaTrainCfg* cfg, const float* rays_o,
                                    const float* rays_d, int64_t n_rays, const float* alpha_beta, const float* d_all, const float* sdf,
                                    const float* radiance, const float* nablas, const float* grad_rgb, void* grad_pack,
                                    double* scalars, void* workspace, size_t workspace_bytes, void* stream) {
    if (desc && desc->framework != NA_FRAMEWORK_VOLSDF) return NA_ERR_BAD_ARG;
    return render_bwd(desc, packed, cfg, rays_o, rays_d, n_rays, alpha_beta, d_all, sdf, radiance, nablas, grad_rgb, (float*)grad_pack,
                      scalars, workspace, workspace_bytes, (cudaStream_t)stream, false);
}

// Backward half of the split training program: `workspace` is the training workspace the forward render of the SAME rays filled
// (na_volsdf_render_fwd_train); sdf / radiance / nablas / d_all are that render's detailed outputs.
extern "C" int na_volsdf_render_bwd_stashed(const NaNetDesc* desc, const void* packed, const NaTrainCfg* cfg, const float* rays_o,
                                            const float* rays_d, int64_t n_rays, const float* alpha_beta, const float* d_all, const float* sdf,
                                            const float* radiance, const float* nablas, const float* grad_rgb, void* grad_pack,
                                            double* scalars, void* workspace, size_t workspace_bytes, void* stream) {
    if (desc && desc->framework != NA_FRAMEWORK_VOLSDF) return NA_ERR_BAD_ARG;
    return render_bwd(desc, packed, cfg, rays_o, rays_d, n_rays, alpha_beta, d_all, sdf, radiance, nablas, grad_rgb, (float*)grad_pack,
                      scalars, workspace, workspace_bytes, (cudaStream_t)stream, false, true);
}

extern "C" int na_neus_render_bwd(const NaNetDesc* desc, const void* packed, const NaTrainCfg* cfg, const float* rays_o,
                                  const float* rays_d, int64_t n_rays, const float* s, const float* d_all, const float* sdf,
                                  const float* radiance, const float* nablas, const float* grad_rgb, void* grad_pack,
                                  double* scalars, void* workspace, size_t workspace_bytes, void* stream) {
    if (desc && desc->framework != NA_FRAMEWORK_NEUS) return NA_ERR_BAD_ARG;
    return render_bwd(desc, packed, cfg, rays_o, rays_d, n_rays, s, d_all, sdf, radiance, nablas, grad_rgb, (float*)grad_pack,
                      scalars, workspace, workspace_bytes, (cudaStream_t)stream, true);
}

// NeuS split program: `workspace` was filled by na_neus_render_fwd_train for the same rays (points stash, then midpoints stash)
extern "C" int na_neus_render_bwd_stashed(const NaNetDesc* desc, const void* packed, const NaTrainCfg* cfg, const float* rays_o,
                                          const float* rays_d, int64_t n_rays, const float* s, const float* d_all, const float* sdf,
                                          const float* radiance, const float* nablas, const float* grad_rgb, void* grad_pack,
                                          double* scalars, void* workspace, size_t workspace_bytes, void* stream) {
    if (desc && desc->framework != NA_FRAMEWORK_NEUS) return NA_ERR_BAD_ARG;
    return render_bwd(desc, packed, cfg, rays_o, rays_d, n_rays, s, d_all, sdf, radiance, nablas, grad_rgb, (float*)grad_pack,
                      scalars, workspace, workspace_bytes, (cudaStream_t)stream, true, true);
}

extern "C" int na_unpack_grads(const NaNetDesc* desc, const NaRawParams* raw, const void* grad_pack, const NaRawGrads* out, void* stream) {
    if (!desc || !raw || !grad_pack || !out) return NA_ERR_BAD_ARG;
    for (int l = 0; l < 14; ++l) if (out->weight_v[l] && (!raw->weight_v[l] || !raw->weight_g[l])) return NA_ERR_BAD_ARG;
    unpack_grads_kernel<<<dim3(257, 14), 128, 0, (cudaStream_t)stream>>>(*raw, *out, grad_layout(), (const float*)grad_pack,
                                                                         small_dim(desc->multires_view));
    NA_CHECK_LAUNCH();
    return NA_OK;
}
