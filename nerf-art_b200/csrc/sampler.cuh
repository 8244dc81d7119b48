// Per-ray sampler primitives (device side).  One CTA per ray; all per-ray arrays live in shared memory.
//
// Numerics follow the reference's torch-CPU semantics: prefix sums accumulate fp32 inputs in fp64 and round
// every prefix to fp32 (torch.cumsum / cumprod CPU kernels use at::acc_type<float,false> = double); all other
// arithmetic is fp32 with the reference's operation order (no FMA contraction where a product is rounded
// before a sum in the reference).
#pragma once
#include "common.cuh"

namespace na {

constexpr int SNT = 256;                 // threads per sampler CTA
constexpr int MAX_CAP = 8192;            // longest per-ray depth array the shared-memory sampler handles

__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// out[i] = (float) sum_{j<=i} (double) in[j],  i < n.   in/out may alias.  `red` is >= 8 doubles of shared scratch.
__device__ inline void block_cumsum(const float* in, float* out, int n, double* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int items = ((n + SNT - 1) / SNT) | 1;        // odd stride: conflict-free contiguous segments
    const int beg = min(tid * items, n), end = min(beg + items, n);
    double local = 0.0;
    for (int i = beg; i < end; ++i) local += (double)in[i];
    double incl = warp_incl_scan(local, lane);
    if (lane == 31) red[warp] = incl;
    __syncthreads();
    double woff = 0.0;
    for (int w = 0; w < warp; ++w) woff += red[w];
    double run = woff + incl - local;
    __syncthreads();                                   // red[] reusable; also orders in[] reads before out[] writes of others
    for (int i = beg; i < end; ++i) { run += (double)in[i]; out[i] = (float)run; }
    __syncthreads();
}

__device__ inline float block_max(float v, float* redf) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) redf[warp] = v;
    __syncthreads();
    float r = redf[0];
#pragma unroll
    for (int w = 1; w < SNT / 32; ++w) r = fmaxf(r, redf[w]);
    __syncthreads();
    return r;
}

__device__ inline double block_sum(double v, double* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int w = 0; w < SNT / 32; ++w) r += red[w];
    __syncthreads();
    return r;
}

// volsdf.sdf_to_sigma, models/frameworks/volsdf.py:34-53
__device__ __forceinline__ float sdf_to_sigma(float sdf, float alpha, float beta) {
    const float e = __fmul_rn(0.5f, expf(__fdiv_rn(-fabsf(sdf), beta)));
    const float psi = sdf >= 0.f ? e : __fsub_rn(1.f, e);
    return __fmul_rn(alpha, psi);
}

// R_t[i] (i < n-1): exclusive prefix sum of sigma_i * delta_i  (volsdf.py:77-81, 128-132).  Result in Rt[0..n-2].
__device__ inline void compute_Rt(const float* d, const float* s, int n, float alpha, float beta, float* Rt, double* red) {
    for (int i = threadIdx.x; i < n - 1; i += SNT)
        Rt[i] = __fmul_rn(sdf_to_sigma(s[i], alpha, beta), __fsub_rn(d[i + 1], d[i]));
    __syncthreads();
    block_cumsum(Rt, Rt, n - 1, red);
    // inclusive -> exclusive (shift right by one)
    float prev[33];
    const int items = ((n - 1 + SNT - 1) / SNT) | 1;     // <= 33 for n <= MAX_CAP
    const int beg = min((int)threadIdx.x * items, n - 1), end = min(beg + items, n - 1);
    for (int i = beg; i < end; ++i) prev[i - beg] = i == 0 ? 0.f : Rt[i - 1];
    __syncthreads();
    for (int i = beg; i < end; ++i) Rt[i] = prev[i - beg];
    __syncthreads();
}

// volsdf.error_bound, volsdf.py:56-94.  bounds[0..n-2]; returns the block-wide max (NaN -> inf, line 93).
// `Rt` and `bounds` are distinct shared arrays of >= n floats.  If clamp, bounds are clamped to [0,1e5] (volsdf.py:282)
// AFTER the max is taken.
__device__ inline float error_bound(const float* d, const float* s, int n, float alpha, float beta,
                                    float* Rt, float* bounds, bool clamp, double* red, float* redf) {
    compute_Rt(d, s, n, alpha, beta, Rt, red);
    const float coef = __fdiv_rn(alpha, __fmul_rn(4.f, beta));
    for (int i = threadIdx.x; i < n - 1; i += SNT) {
        const float delta = __fsub_rn(d[i + 1], d[i]);
        const float dstar = fmaxf(__fmul_rn(0.5f, __fsub_rn(__fadd_rn(fabsf(s[i]), fabsf(s[i + 1])), delta)), 0.f);
        bounds[i] = __fmul_rn(__fmul_rn(coef, __fmul_rn(delta, delta)), expf(__fdiv_rn(-dstar, beta)));
    }
    __syncthreads();
    block_cumsum(bounds, bounds, n - 1, red);
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < n - 1; i += SNT) {
        float b = __fmul_rn(expf(-Rt[i]), __fsub_rn(expf(bounds[i]), 1.f));
        if (isnan(b)) b = INFINITY;
        mx = fmaxf(mx, b);
        bounds[i] = clamp ? fminf(fmaxf(b, 0.f), 1e5f) : b;
    }
    __syncthreads();
    return block_max(mx, redf);
}

// shared tail of rend_util.sample_pdf / sample_cdf (utils/rend_util.py:276-293 / 311-328): invert `cdf` (n entries,
// cdf[0] = 0) at u.  torch.searchsorted(right=False): first index with cdf[i] >= u.
__device__ __forceinline__ float invert_cdf(const float* bins, const float* cdf, int n, float u, int* ind_out) {
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (cdf[mid] < u) lo = mid + 1; else hi = mid; }   // NaN compares false
    if (ind_out) *ind_out = lo;
    const int below = max(lo - 1, 0), above = min(lo, n - 1);
    const float c0 = cdf[below], c1 = cdf[above], b0 = bins[below], b1 = bins[above];
    float denom = __fsub_rn(c1, c0);
    if (denom < 1e-5f) denom = 1.f;
    const float t = __fdiv_rn(__fsub_rn(u, c0), denom);
    return __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
}

// rend_util.sample_pdf lines 258-265: weights w[0..n-2] (in place) -> cdf[0..n-1] (cdf[0]=0).
__device__ inline void pdf_to_cdf(float* w, float* cdf, int n, double* red) {
    double part = 0.0;
    for (int i = threadIdx.x; i < n - 1; i += SNT) { const float v = __fadd_rn(w[i], 1e-5f); w[i] = v; part += (double)v; }
    __syncthreads();
    const float tot = (float)block_sum(part, red);
    for (int i = threadIdx.x; i < n - 1; i += SNT) w[i] = __fdiv_rn(w[i], tot);
    __syncthreads();
    block_cumsum(w, cdf + 1, n - 1, red);
    if (threadIdx.x == 0) cdf[0] = 0.f;
    __syncthreads();
}

// in-place bitonic sort of keys (ascending), n_pad a power of two; entries >= n must be +inf on entry.
__device__ inline void bitonic_sort_keys(float* key, int n_pad) {
    for (int k = 2; k <= n_pad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pad; i += SNT) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = (i & k) == 0;
                    const float a = key[i], b = key[ixj];
                    if ((a > b) == up) { key[i] = b; key[ixj] = a; }
                }
            }
            __syncthreads();
        }
}
__device__ inline void bitonic_sort_pairs(float* key, float* val, int n_pad) {
    for (int k = 2; k <= n_pad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pad; i += SNT) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = (i & k) == 0;
                    const float a = key[i], b = key[ixj];
                    if ((a > b) == up) { key[i] = b; key[ixj] = a; const float t = val[i]; val[i] = val[ixj]; val[ixj] = t; }
                }
            }
            __syncthreads();
        }
}

}  // namespace na
