// Building blocks of the fp32 CUDA-core tile kernels (mlp_simt.cu: forward; train.cu: backward of the fine-tune step).
// A CTA of 256 threads owns a tile of TM = 128 samples whose activations live in shared memory (k-major, XOR-swizzled);
// weights stream from L2 through a 3-stage cp.async ring.  See mlp_simt.cu for the layout notes.
#pragma once
#include "common.cuh"

namespace na {

constexpr int TM = 128;            // samples per tile
constexpr int NT = 256;            // threads per CTA
constexpr int KC = 8;              // contraction rows per weight chunk
constexpr int A_ROWS = 296;        // 256 + max small_pad (40)
constexpr int TAIL0 = 256;         // first tail row

struct __align__(16) MlpSmem {
    float A[A_ROWS * TM];
    float GE[EMB_PAD * TM];        // d sdf / d embedding accumulator
    float Ws[3 * KC * 264];        // weight chunk ring (row stride 256 for the FFMA tiles, 264 for the TF32 mma tiles of train.cu)
    float X[3 * TM];
    float V[3 * TM];
    float RED[2 * 3 * TM];         // two-half partial sums of the narrow output layers
    float SDF[TM];
    float NAB[3 * TM];
    long long OIDX[TM];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N)); }

__device__ __forceinline__ int a_index(int k, int m) { return k * TM + ((((m >> 2) ^ ((k >> 2) & 7)) << 2) | (m & 3)); }

// OUT[m][c] = sum_{r<R} A[a_row0 + r][m] * B[r][c],  c in the 64*NJ leading columns of B (row stride 256 floats).
// Thread (ty=tid/16, tx=tid%16) owns rows 8ty..8ty+7 and columns 64j+4tx+{0..3}, j<NJ.
template <int NJ>
__device__ __forceinline__ void gemm_tile(float (&acc)[8][4 * NJ], const float* __restrict__ Bg, int R, int a_row0,
                                          const float* __restrict__ As, float* __restrict__ Ws, int tid) {
    constexpr int CW = 64 * NJ;                 // chunk width in floats
    constexpr int F4 = KC * 16 * NJ;            // float4 per chunk
    const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4 * NJ; ++j) acc[i][j] = 0.f;
    const int nch = R / KC;
    auto issue = [&](int c) {
        float* dst = Ws + (c % 3) * (KC * 256);
        const float* src = Bg + (size_t)c * KC * 256;
        for (int idx = tid; idx < F4; idx += NT) {
            int row = idx / (16 * NJ), c4 = idx % (16 * NJ);
            cp_async16(dst + row * CW + c4 * 4, src + row * 256 + c4 * 4);
        }
        cp_async_commit();
    };
    issue(0);
    if (nch > 1) issue(1);
    for (int c = 0; c < nch; ++c) {
        if (c + 1 < nch) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();
        if (c + 2 < nch) issue(c + 2);
        const float* wb = Ws + (c % 3) * (KC * 256);
        const int rbase = a_row0 + c * KC;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            const int r = rbase + kk;
            const int swz = (r >> 2) & 7;
            const float* arow = As + r * TM;
            const float4 a0 = *reinterpret_cast<const float4*>(arow + (((2 * ty) ^ swz) << 2));
            const float4 a1 = *reinterpret_cast<const float4*>(arow + (((2 * ty + 1) ^ swz) << 2));
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const float4 b = *reinterpret_cast<const float4*>(wb + kk * CW + 64 * j + 4 * tx);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    acc[i][4 * j + 0] = fmaf(a[i], b.x, acc[i][4 * j + 0]);
                    acc[i][4 * j + 1] = fmaf(a[i], b.y, acc[i][4 * j + 1]);
                    acc[i][4 * j + 2] = fmaf(a[i], b.z, acc[i][4 * j + 2]);
                    acc[i][4 * j + 3] = fmaf(a[i], b.w, acc[i][4 * j + 3]);
                }
            }
        }
    }
    __syncthreads();        // every thread is done reading A / Ws: the epilogue may overwrite A in place
}

// nn.Softplus(beta=100), threshold 20 (models/base.py:202) and its derivative as torch's softplus_backward computes it.
__device__ __forceinline__ void softplus100(float z, float& h, float& dh) {
    const float bz = z * 100.f;
    if (bz > 20.f) { h = z; dh = 1.f; }
    else { const float e = expf(bz); h = __fdiv_rn(log1pf(e), 100.f); dh = __fdiv_rn(e, e + 1.f); }
}
__device__ __forceinline__ float sigmoidf_(float x) { return __fdiv_rn(1.f, 1.f + expf(-x)); }

// store thread-owned column k (8 rows) of the tile into A (swizzled) / into a [256][TM] scratch plane
__device__ __forceinline__ void store_col_A(float* A, int k, int ty, int tx, const float (&v)[8]) {
    float* base = A + k * TM;
    const int swz = tx & 7;       // == (k>>2)&7 for k = 64j+4tx+c
    *reinterpret_cast<float4*>(base + (((2 * ty) ^ swz) << 2)) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(base + (((2 * ty + 1) ^ swz) << 2)) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store_col_plane(float* plane, int k, int ty, const float (&v)[8]) {
    float* p = plane + k * TM + 8 * ty;
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void load_col_plane(const float* plane, int k, int ty, float (&v)[8]) {
    const float* p = plane + k * TM + 8 * ty;
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// out[o][m] partial sums over half the 256 rows of A; combined by the caller from RED
template <int NO>
__device__ __forceinline__ void narrow_layer(const float* A, const float* __restrict__ Wg /*[NO][256]*/, float* RED, int tid) {
    const int m = tid & (TM - 1), half = tid >> 7;
    float s[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) s[o] = 0.f;
#pragma unroll 8
    for (int kk = 0; kk < 128; ++kk) {
        const int k = half * 128 + kk;
        const float a = A[a_index(k, m)];
#pragma unroll
        for (int o = 0; o < NO; ++o) s[o] = fmaf(a, __ldg(Wg + o * 256 + k), s[o]);
    }
#pragma unroll
    for (int o = 0; o < NO; ++o) RED[(half * 3 + o) * TM + m] = s[o];
}


}  // namespace na
