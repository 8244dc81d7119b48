// CLIP ViT-B/32 image tower (`model.encode_image` of openai/CLIP, the third-party model behind criteria/clip_loss.py:206-208,
// contrastive_loss.py:113-115, patchnce_loss.py:127-129): forward and the backward w.r.t. the INPUT IMAGE (the weights are
// frozen in NeRF-Art; only d loss / d rendered pixels is needed, volsdf.py:744-749).
//   conv1 32x32/32 (3->768, no bias) -> [cls | 49 patches] + positional embedding -> ln_pre -> 12 x [x += MHA(ln_1(x)),
//   x += c_proj(QuickGELU(c_fc(ln_2(x))))] -> ln_post(cls) -> @ proj (768->512).          
// First correct version: fp32 on CUDA cores (the reference runs this model in fp16 on GPU, so fp32 is >= its precision).
// Kernels: one tiled SGEMM with fused bias / QuickGELU / residual epilogues (NT for forward, NN for backward-data),
// warp-per-row LayerNorm forward/backward, one CTA per (image, head) attention forward/backward over the 50 tokens,
// im2col / col2im for the non-overlapping patches.  Activations the backward needs are kept in the caller's workspace.
#include "common.cuh"
#undef NA_DIAG_STREAM
#define NA_DIAG_STREAM s

namespace na {
namespace clipv {

constexpr int D = 768, HEADS = 12, HD = 64, TOK = 50, NPATCH = 49, FF = 3072, PATCH_K = 3 * 32 * 32, OUT = 512, LAYERS = 12;
constexpr int IMG = 224;

// ---------------------------------------------------------------------------------------------------------------------
// C[M,N] = epilogue(A[M,K] * B)   B_NT: B is [N][K] (y = x W^T);  else B is [K][N] (dx = dy W).  N % 64 == 0, K % 16 == 0.
// epilogue: + bias[n]; raw copy to C_raw (pre-activation, for the backward); QuickGELU; + residual R[m][n]
// ---------------------------------------------------------------------------------------------------------------------
template <bool B_NT>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N, int K,
             const float* __restrict__ bias, float* __restrict__ C_raw, int act, const float* __restrict__ R) {
    __shared__ __align__(16) float As[2][16][68];
    __shared__ __align__(16) float Bs[2][16][68];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    float4 ra, rb;
    auto gload = [&](int k0) {
        {   // A tile [64 m][16 k]: thread -> row tid/4, k4 = (tid%4)*4
            const int r = tid >> 2, k4 = (tid & 3) * 4;
            ra = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + r < M) ra = *reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * K + k0 + k4);
        }
        if (B_NT) {
            const int r = tid >> 2, k4 = (tid & 3) * 4;
            rb = *reinterpret_cast<const float4*>(B + (size_t)(n0 + r) * K + k0 + k4);
        } else {
            const int r = tid >> 4, n4 = (tid & 15) * 4;
            rb = *reinterpret_cast<const float4*>(B + (size_t)(k0 + r) * N + n0 + n4);
        }
    };
    auto sstore = [&](int buf) {
        {
            const int r = tid >> 2, k4 = (tid & 3) * 4;
            As[buf][k4][r] = ra.x; As[buf][k4 + 1][r] = ra.y; As[buf][k4 + 2][r] = ra.z; As[buf][k4 + 3][r] = ra.w;
        }
        if (B_NT) {
            const int r = tid >> 2, k4 = (tid & 3) * 4;
            Bs[buf][k4][r] = rb.x; Bs[buf][k4 + 1][r] = rb.y; Bs[buf][k4 + 2][r] = rb.z; Bs[buf][k4 + 3][r] = rb.w;
        } else {
            const int r = tid >> 4, n4 = (tid & 15) * 4;
            *reinterpret_cast<float4*>(&Bs[buf][r][n4]) = rb;
        }
    };
    gload(0); sstore(0); __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < K; k0 += 16) {
        const bool more = k0 + 16 < K;
        if (more) gload(k0 + 16);
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) c[i][j] = fmaf(av[i], bv[j], c[i][j]);
        }
        if (more) sstore(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
    const int n = n0 + tx * 4;
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bv = *reinterpret_cast<const float4*>(bias + n);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
        float v[4] = {c[i][0] + bv.x, c[i][1] + bv.y, c[i][2] + bv.z, c[i][3] + bv.w};
        if (C_raw) *reinterpret_cast<float4*>(C_raw + (size_t)m * N + n) = make_float4(v[0], v[1], v[2], v[3]);
        if (act) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = v[j] * __fdiv_rn(1.f, 1.f + expf(-1.702f * v[j]));      // QuickGELU (clip/model.py)
        }
        if (R) {
            const float4 r = *reinterpret_cast<const float4*>(R + (size_t)m * N + n);
            v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
        }
        *reinterpret_cast<float4*>(C + (size_t)m * N + n) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// LayerNorm over D = 768 (eps 1e-5), one warp per row.  x rows are `in_stride` floats apart (ln_post reads the cls rows only).
__global__ void ln_fwd_kernel(const float* __restrict__ x, size_t in_stride, const float* __restrict__ w, const float* __restrict__ b,
                              float* __restrict__ y, float* __restrict__ stats, int rows) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (size_t)row * in_stride;
    float v[24], s = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) { v[i] = xr[lane + 32 * i]; s += v[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) { const float d = v[i] - mean; q += d * d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / D) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 24; ++i) { const int k = lane + 32 * i; y[(size_t)row * D + k] = (v[i] - mean) * rstd * w[k] + b[k]; }
    if (lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
}
// dx = rstd * (dy*w - mean(dy*w) - xhat * mean(dy*w*xhat))  (+ residual gradient `dres`), rows of dx are out_stride apart
__global__ void ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, size_t in_stride, const float* __restrict__ w,
                              const float* __restrict__ stats, const float* __restrict__ dres, float* __restrict__ dx, size_t out_stride,
                              int rows) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    float g[24], xh[24], c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) {
        const int k = lane + 32 * i;
        g[i] = dy[(size_t)row * D + k] * w[k];
        xh[i] = (x[(size_t)row * in_stride + k] - mean) * rstd;
        c1 += g[i]; c2 += g[i] * xh[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { c1 += __shfl_xor_sync(0xffffffffu, c1, o); c2 += __shfl_xor_sync(0xffffffffu, c2, o); }
    c1 *= (1.f / D); c2 *= (1.f / D);
#pragma unroll
    for (int i = 0; i < 24; ++i) {
        const int k = lane + 32 * i;
        float v = rstd * (g[i] - c1 - xh[i] * c2);
        if (dres) v += dres[(size_t)row * D + k];
        dx[(size_t)row * out_stride + k] = v;
    }
}

// multi-head self-attention over the 50 tokens of one image, one CTA per (image, head)
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const float* __restrict__ qkv /*[B*50][2304]*/, float* __restrict__ ctx /*[B*50][768]*/, float* __restrict__ Pm /*[B][12][50][50]*/) {
    __shared__ float q[TOK][HD], k[TOK][HD + 1], v[TOK][HD], S[TOK][TOK + 1];       // 48.8 KB; only k is read with a per-lane row
    const int b = blockIdx.x / HEADS, h = blockIdx.x % HEADS, tid = threadIdx.x;
    for (int idx = tid; idx < TOK * HD; idx += 128) {
        const int t = idx / HD, d = idx % HD;
        const float* r = qkv + (size_t)(b * TOK + t) * (3 * D) + h * HD + d;
        q[t][d] = r[0]; k[t][d] = r[D]; v[t][d] = r[2 * D];
    }
    __syncthreads();
    for (int idx = tid; idx < TOK * TOK; idx += 128) {
        const int i = idx / TOK, j = idx % TOK;
        float s = 0.f;
#pragma unroll 16
        for (int d = 0; d < HD; ++d) s = fmaf(q[i][d], k[j][d], s);
        S[i][j] = s * 0.125f;
    }
    __syncthreads();
    if (tid < TOK) {
        float mx = -1e30f;
        for (int j = 0; j < TOK; ++j) mx = fmaxf(mx, S[tid][j]);
        float sum = 0.f;
        for (int j = 0; j < TOK; ++j) { const float e = expf(S[tid][j] - mx); S[tid][j] = e; sum += e; }
        const float inv = 1.f / sum;
        for (int j = 0; j < TOK; ++j) S[tid][j] *= inv;
    }
    __syncthreads();
    for (int idx = tid; idx < TOK * TOK; idx += 128) Pm[(size_t)blockIdx.x * TOK * TOK + idx] = S[idx / TOK][idx % TOK];
    for (int idx = tid; idx < TOK * HD; idx += 128) {
        const int i = idx / HD, d = idx % HD;
        float o = 0.f;
        for (int j = 0; j < TOK; ++j) o = fmaf(S[i][j], v[j][d], o);
        ctx[(size_t)(b * TOK + i) * D + h * HD + d] = o;
    }
}
__global__ void __launch_bounds__(128)
attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ Pm, const float* __restrict__ dctx, float* __restrict__ dqkv) {
    extern __shared__ __align__(16) float attn_smem[];                              // 72.4 KB (opt-in)
    float (*q)[HD + 1] = reinterpret_cast<float (*)[HD + 1]>(attn_smem);
    float (*k)[HD + 1] = q + TOK;
    float (*v)[HD + 1] = k + TOK;
    float (*dO)[HD + 1] = v + TOK;
    float (*Pp)[TOK + 1] = reinterpret_cast<float (*)[TOK + 1]>(attn_smem + 4 * TOK * (HD + 1));
    float (*dS)[TOK + 1] = Pp + TOK;
    const int b = blockIdx.x / HEADS, h = blockIdx.x % HEADS, tid = threadIdx.x;
    for (int idx = tid; idx < TOK * HD; idx += 128) {
        const int t = idx / HD, d = idx % HD;
        const float* r = qkv + (size_t)(b * TOK + t) * (3 * D) + h * HD + d;
        q[t][d] = r[0]; k[t][d] = r[D]; v[t][d] = r[2 * D];
        dO[t][d] = dctx[(size_t)(b * TOK + t) * D + h * HD + d];
    }
    for (int idx = tid; idx < TOK * TOK; idx += 128) Pp[idx / TOK][idx % TOK] = Pm[(size_t)blockIdx.x * TOK * TOK + idx];
    __syncthreads();
    for (int idx = tid; idx < TOK * TOK; idx += 128) {          // dP
        const int i = idx / TOK, j = idx % TOK;
        float s = 0.f;
#pragma unroll 16
        for (int d = 0; d < HD; ++d) s = fmaf(dO[i][d], v[j][d], s);
        dS[i][j] = s;
    }
    __syncthreads();
    if (tid < TOK) {
        float r = 0.f;
        for (int j = 0; j < TOK; ++j) r = fmaf(dS[tid][j], Pp[tid][j], r);
        for (int j = 0; j < TOK; ++j) dS[tid][j] = Pp[tid][j] * (dS[tid][j] - r);
    }
    __syncthreads();
    for (int idx = tid; idx < TOK * HD; idx += 128) {
        const int t = idx / HD, d = idx % HD;
        float dq = 0.f, dk = 0.f, dv = 0.f;
        for (int j = 0; j < TOK; ++j) {
            dq = fmaf(dS[t][j], k[j][d], dq);
            dk = fmaf(dS[j][t], q[j][d], dk);
            dv = fmaf(Pp[j][t], dO[j][d], dv);
        }
        float* r = dqkv + (size_t)(b * TOK + t) * (3 * D) + h * HD + d;
        r[0] = dq * 0.125f; r[D] = dk * 0.125f; r[2 * D] = dv;
    }
}

__global__ void qgelu_bwd_kernel(const float* __restrict__ h, float* __restrict__ da_inout, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = h[i], s = __fdiv_rn(1.f, 1.f + expf(-1.702f * x));
    da_inout[i] *= s + 1.702f * x * s * (1.f - s);
}
// patches[b*49 + py*7+px][c*1024 + ky*32 + kx] = img[b][c][py*32+ky][px*32+kx]   (conv1 weight [768][3][32][32] flattened)
__global__ void im2col_kernel(const float* __restrict__ img, float* __restrict__ patches, int B, int reverse, float* __restrict__ img_out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)B * 3 * IMG * IMG;
    if (i >= n) return;
    const int x = (int)(i % IMG), y = (int)((i / IMG) % IMG), c = (int)((i / (IMG * IMG)) % 3), b = (int)(i / (3 * IMG * IMG));
    const size_t p = (size_t)(b * NPATCH + (y >> 5) * 7 + (x >> 5)) * PATCH_K + c * 1024 + (y & 31) * 32 + (x & 31);
    if (reverse) img_out[i] = patches[p]; else patches[p] = img[i];
}
// x0[b][t][:] = (t == 0 ? class_embedding : emb[b][t-1][:]) + positional_embedding[t]
__global__ void tokens_kernel(const float* __restrict__ emb, const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ x0, int B) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * TOK * D) return;
    const int k = (int)(i % D), t = (int)((i / D) % TOK), b = (int)(i / ((size_t)D * TOK));
    x0[i] = (t == 0 ? cls[k] : emb[(size_t)(b * NPATCH + t - 1) * D + k]) + pos[t * D + k];
}
__global__ void tokens_bwd_kernel(const float* __restrict__ dx0, float* __restrict__ demb, int B) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * NPATCH * D) return;
    const int k = (int)(i % D), p = (int)((i / D) % NPATCH), b = (int)(i / ((size_t)D * NPATCH));
    demb[i] = dx0[(size_t)(b * TOK + 1 + p) * D + k];
}

constexpr int ATTN_BWD_SMEM = (4 * TOK * (HD + 1) + 2 * TOK * (TOK + 1)) * (int)sizeof(float);

struct Ws {          // offsets in floats
    size_t patches, emb, x0, stats_pre, x[LAYERS + 1], stats1[LAYERS], stats2[LAYERS], qkv[LAYERS], P[LAYERS], xmid[LAYERS], h[LAYERS];
    size_t stats_post, ybuf, ctx, abuf, ypost, d_a, d_b, d_c, total;
};
static Ws layout(int B) {
    Ws w; size_t o = 0; const size_t T = (size_t)B * TOK;
    auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };
    w.patches = take((size_t)B * NPATCH * PATCH_K); w.emb = take((size_t)B * NPATCH * D); w.x0 = take(T * D); w.stats_pre = take(2 * T);
    for (int l = 0; l <= LAYERS; ++l) w.x[l] = take(T * D);
    for (int l = 0; l < LAYERS; ++l) {
        w.stats1[l] = take(2 * T); w.stats2[l] = take(2 * T); w.qkv[l] = take(T * 3 * D); w.P[l] = take((size_t)B * HEADS * TOK * TOK);
        w.xmid[l] = take(T * D); w.h[l] = take(T * FF);
    }
    w.stats_post = take(2 * B); w.ybuf = take(T * D); w.ctx = take(T * D); w.abuf = take(T * FF); w.ypost = take((size_t)B * D);
    w.d_a = take(T * FF); w.d_b = take(T * 3 * D); w.d_c = take(T * D);
    w.total = o;
    return w;
}

}  // namespace clipv
int tgemm(const float* A, const unsigned char* img, float* C, int M, int N, int K, const float* bias, float* raw, int act,
          const float* R, cudaStream_t stream);                                            // csrc/tgemm.cu
int tgemm_pack(const float* src, int rows, int cols, int transpose, unsigned char* img, cudaStream_t stream);
namespace clipv {

// TF32 weight images of one [R][C] matrix (byte offsets into the packed buffer): `nt` serves y = x W^T (operand rows = R,
// contraction = C), `tn` serves dx = dy W (operand rows = C, contraction = R)
struct Img { size_t nt, tn; };
struct Packed { Img conv1, proj, in[LAYERS], out[LAYERS], fc[LAYERS], pr[LAYERS]; size_t total; };
static Packed packed_layout() {
    Packed P; size_t o = 0;
    auto take = [&](size_t r, size_t c) { Img im; im.nt = o; o += r * c * 4; im.tn = o; o += r * c * 4; return im; };
    P.conv1 = take(D, PATCH_K); P.proj = take(D, OUT);
    for (int l = 0; l < LAYERS; ++l) { P.in[l] = take(3 * D, D); P.out[l] = take(D, D); P.fc[l] = take(FF, D); P.pr[l] = take(D, FF); }
    P.total = o;
    return P;
}

// one linear layer: the tcgen05 TF32 kernel when `packed` is set (NA_CLIP_TF32), else the fp32 CUDA-core kernel
struct Gemm {
    const unsigned char* packed; cudaStream_t s;
    int operator()(bool nt, const float* A, const float* Bm, const Img& im, float* C, int M, int N, int K, const float* bias, float* raw,
                   int act, const float* R) const {
        if (N % 64 || K % 16) return NA_ERR_UNSUPPORTED;
        if (packed) return tgemm(A, packed + (nt ? im.nt : im.tn), C, M, N, K, bias, raw, act, R, s);
        dim3 grid(N / 64, (M + 63) / 64);
        if (nt) sgemm_kernel<true><<<grid, 256, 0, s>>>(A, Bm, C, M, N, K, bias, raw, act, R);
        else    sgemm_kernel<false><<<grid, 256, 0, s>>>(A, Bm, C, M, N, K, bias, raw, act, R);
        NA_CHECK_LAUNCH();
        return NA_OK;
    }
};
static int ln_fwd(const float* x, size_t stride, const float* w, const float* b, float* y, float* stats, int rows, cudaStream_t s) {
    ln_fwd_kernel<<<(rows + 3) / 4, 128, 0, s>>>(x, stride, w, b, y, stats, rows);
    NA_CHECK_LAUNCH();
    return NA_OK;
}
static int ln_bwd(const float* dy, const float* x, size_t stride, const float* w, const float* stats, const float* dres, float* dx,
                  size_t ostride, int rows, cudaStream_t s) {
    ln_bwd_kernel<<<(rows + 3) / 4, 128, 0, s>>>(dy, x, stride, w, stats, dres, dx, ostride, rows);
    NA_CHECK_LAUNCH();
    return NA_OK;
}

int preload_clip() {
    NA_PRELOAD((sgemm_kernel<true>));
    NA_PRELOAD((sgemm_kernel<false>));
    NA_PRELOAD(ln_fwd_kernel);
    NA_PRELOAD(ln_bwd_kernel);
    NA_PRELOAD(attn_fwd_kernel);
    NA_PRELOAD(attn_bwd_kernel);
    NA_PRELOAD(qgelu_bwd_kernel);
    NA_PRELOAD(im2col_kernel);
    NA_PRELOAD(tokens_kernel);
    NA_PRELOAD(tokens_bwd_kernel);
    return NA_OK;
}

}  // namespace clipv
}  // namespace na

using namespace na;
using namespace na::clipv;

extern "C" size_t na_clip_packed_bytes(void) { return packed_layout().total; }

extern "C" int na_clip_pack_weights(const NaClipWeights* Wt, void* packed_, void* stream_) {
    if (!Wt || !packed_) return NA_ERR_BAD_ARG;
    cudaStream_t s = (cudaStream_t)stream_;
    unsigned char* pk = (unsigned char*)packed_;
    const Packed P = packed_layout();
    auto both = [&](const float* src, int rows, int cols, const Img& im) {
        NA_TRY(tgemm_pack(src, rows, cols, 0, pk + im.nt, s));
        return tgemm_pack(src, rows, cols, 1, pk + im.tn, s);
    };
    NA_TRY(both(Wt->conv1, D, PATCH_K, P.conv1));
    NA_TRY(both(Wt->proj, D, OUT, P.proj));
    for (int l = 0; l < LAYERS; ++l) {
        const NaClipLayer& Lw = Wt->layers[l];
        NA_TRY(both(Lw.in_proj_w, 3 * D, D, P.in[l]));
        NA_TRY(both(Lw.out_proj_w, D, D, P.out[l]));
        NA_TRY(both(Lw.c_fc_w, FF, D, P.fc[l]));
        NA_TRY(both(Lw.c_proj_w, D, FF, P.pr[l]));
    }
    return NA_OK;
}

extern "C" size_t na_clip_workspace_bytes(int32_t batch) { return batch > 0 ? layout(batch).total * sizeof(float) : 0; }

extern "C" int na_clip_vitb32_encode_fwd(const NaClipWeights* Wt, const float* images, int32_t B, float* feats, void* ws_, size_t ws_bytes,
                                         void* stream_) {
    if (!Wt || !images || !feats || !ws_ || B <= 0) return NA_ERR_BAD_ARG;
    const Ws w = layout(B);
    if (ws_bytes < w.total * sizeof(float)) return NA_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream_;
    float* ws = (float*)ws_;
    const int T = B * TOK;
    const Packed P = packed_layout();
    const Gemm gemm{Wt->precision == NA_CLIP_TF32 ? (const unsigned char*)Wt->packed : nullptr, s};
    if (Wt->precision == NA_CLIP_TF32 && !Wt->packed) return NA_ERR_BAD_ARG;
    const size_t npix = (size_t)B * 3 * IMG * IMG;
    im2col_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, s>>>(images, ws + w.patches, B, 0, nullptr);
    NA_CHECK_LAUNCH();
    NA_TRY(gemm(true, ws + w.patches, Wt->conv1, P.conv1, ws + w.emb, B * NPATCH, D, PATCH_K, nullptr, nullptr, 0, nullptr));
    tokens_kernel<<<(unsigned)(((size_t)T * D + 255) / 256), 256, 0, s>>>(ws + w.emb, Wt->class_embedding, Wt->positional_embedding, ws + w.x0, B);
    NA_CHECK_LAUNCH();
    NA_TRY(ln_fwd(ws + w.x0, D, Wt->ln_pre_w, Wt->ln_pre_b, ws + w.x[0], ws + w.stats_pre, T, s));
    for (int l = 0; l < LAYERS; ++l) {
        const NaClipLayer& Lw = Wt->layers[l];
        NA_TRY(ln_fwd(ws + w.x[l], D, Lw.ln_1_w, Lw.ln_1_b, ws + w.ybuf, ws + w.stats1[l], T, s));
        NA_TRY(gemm(true, ws + w.ybuf, Lw.in_proj_w, P.in[l], ws + w.qkv[l], T, 3 * D, D, Lw.in_proj_b, nullptr, 0, nullptr));
        attn_fwd_kernel<<<B * HEADS, 128, 0, s>>>(ws + w.qkv[l], ws + w.ctx, ws + w.P[l]);
        NA_CHECK_LAUNCH();
        NA_TRY(gemm(true, ws + w.ctx, Lw.out_proj_w, P.out[l], ws + w.xmid[l], T, D, D, Lw.out_proj_b, nullptr, 0, ws + w.x[l]));
        NA_TRY(ln_fwd(ws + w.xmid[l], D, Lw.ln_2_w, Lw.ln_2_b, ws + w.ybuf, ws + w.stats2[l], T, s));
        NA_TRY(gemm(true, ws + w.ybuf, Lw.c_fc_w, P.fc[l], ws + w.abuf, T, FF, D, Lw.c_fc_b, ws + w.h[l], 1, nullptr));
        NA_TRY(gemm(true, ws + w.abuf, Lw.c_proj_w, P.pr[l], ws + w.x[l + 1], T, D, FF, Lw.c_proj_b, nullptr, 0, ws + w.xmid[l]));
    }
    NA_TRY(ln_fwd(ws + w.x[LAYERS], (size_t)TOK * D, Wt->ln_post_w, Wt->ln_post_b, ws + w.ypost, ws + w.stats_post, B, s));
    return gemm(false, ws + w.ypost, Wt->proj, P.proj, feats, B, OUT, D, nullptr, nullptr, 0, nullptr);
}

extern "C" int na_clip_vitb32_encode_bwd(const NaClipWeights* Wt, const float* grad_feats, int32_t B, float* grad_images, void* ws_,
                                         size_t ws_bytes, void* stream_) {
    if (!Wt || !grad_feats || !grad_images || !ws_ || B <= 0) return NA_ERR_BAD_ARG;
    const Ws w = layout(B);
    if (ws_bytes < w.total * sizeof(float)) return NA_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream_;
    float* ws = (float*)ws_;
    const int T = B * TOK;
    static bool attr_done[64] = {false};
    if (first_on_device(attr_done)) {
        NA_TRY(check_cuda(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_BWD_SMEM)));
    }
    const Packed P = packed_layout();
    const Gemm gemm{Wt->precision == NA_CLIP_TF32 ? (const unsigned char*)Wt->packed : nullptr, s};
    if (Wt->precision == NA_CLIP_TF32 && !Wt->packed) return NA_ERR_BAD_ARG;
    float* dx = ws + w.d_c;                       // gradient w.r.t. the residual stream, [T][768]
    float* dy = ws + w.ybuf;                      // scratch [T][768]
    // feats = ln_post(x[:,0,:]) @ proj  ->  d ypost = grad_feats @ proj^T  (proj is [768][512] = "B[N=768][K=512]" in NT form)
    NA_TRY(gemm(true, grad_feats, Wt->proj, P.proj, ws + w.ypost, B, D, OUT, nullptr, nullptr, 0, nullptr));
    NA_TRY(check_cuda(cudaMemsetAsync(dx, 0, (size_t)T * D * sizeof(float), s)));
    NA_TRY(ln_bwd(ws + w.ypost, ws + w.x[LAYERS], (size_t)TOK * D, Wt->ln_post_w, ws + w.stats_post, nullptr, dx, (size_t)TOK * D, B, s));
    for (int l = LAYERS - 1; l >= 0; --l) {
        const NaClipLayer& Lw = Wt->layers[l];
        // x_{l+1} = xmid + c_proj(qgelu(c_fc(ln_2(xmid))))
        NA_TRY(gemm(false, dx, Lw.c_proj_w, P.pr[l], ws + w.d_a, T, FF, D, nullptr, nullptr, 0, nullptr));             // d a = dx @ Wproj
        qgelu_bwd_kernel<<<(unsigned)(((size_t)T * FF + 255) / 256), 256, 0, s>>>(ws + w.h[l], ws + w.d_a, (size_t)T * FF);
        NA_CHECK_LAUNCH();
        NA_TRY(gemm(false, ws + w.d_a, Lw.c_fc_w, P.fc[l], dy, T, D, FF, nullptr, nullptr, 0, nullptr));               // d ln_2 out
        NA_TRY(ln_bwd(dy, ws + w.xmid[l], D, Lw.ln_2_w, ws + w.stats2[l], dx, dx, D, T, s));                     // dx := d xmid
        // xmid = x_l + out_proj(attn(in_proj(ln_1(x_l))))
        NA_TRY(gemm(false, dx, Lw.out_proj_w, P.out[l], ws + w.ctx, T, D, D, nullptr, nullptr, 0, nullptr));            // d ctx
        attn_bwd_kernel<<<B * HEADS, 128, ATTN_BWD_SMEM, s>>>(ws + w.qkv[l], ws + w.P[l], ws + w.ctx, ws + w.d_b);
        NA_CHECK_LAUNCH();
        NA_TRY(gemm(false, ws + w.d_b, Lw.in_proj_w, P.in[l], dy, T, D, 3 * D, nullptr, nullptr, 0, nullptr));         // d ln_1 out
        NA_TRY(ln_bwd(dy, ws + w.x[l], D, Lw.ln_1_w, ws + w.stats1[l], dx, dx, D, T, s));                        // dx := d x_l
    }
    NA_TRY(ln_bwd(dx, ws + w.x0, D, Wt->ln_pre_w, ws + w.stats_pre, nullptr, dy, D, T, s));                       // d x0
    tokens_bwd_kernel<<<(unsigned)(((size_t)B * NPATCH * D + 255) / 256), 256, 0, s>>>(dy, ws + w.emb, B);
    NA_CHECK_LAUNCH();
    NA_TRY(gemm(false, ws + w.emb, Wt->conv1, P.conv1, ws + w.patches, B * NPATCH, PATCH_K, D, nullptr, nullptr, 0, nullptr));   // d patches
    const size_t npix = (size_t)B * 3 * IMG * IMG;
    im2col_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, s>>>(nullptr, ws + w.patches, B, 1, grad_images);
    NA_CHECK_LAUNCH();
    return NA_OK;
}
