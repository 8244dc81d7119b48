// NeuS renderer: neus.volume_render (models/frameworks/neus.py:142-424), upsample_algo='official_solution', N_outside=0.
//   rays_prep(near/far from the bounding sphere) -> [MLP sdf @ 64 coarse] -> 4 x { upsample(it): merge + weights + sample 16
//   -> [MLP sdf @ new] } -> merge -> [MLP sdf+nabla @ 128] -> [MLP radiance @ 127 midpoints] -> composite
#include "sampler.cuh"

namespace na {

int launch_mlp(const EvalJob& job, const void* packed, int precision, float* scratch, size_t scratch_bytes, cudaStream_t stream);
size_t mlp_scratch_bytes();
int launch_normalize_dirs(const float* d_in, float* d_out, long long n, cudaStream_t stream);

// near / far (rend_util.near_far_from_sphere, utils/rend_util.py:168-186) and the coarse depths (neus.py:235-236)
__global__ void neus_init_kernel(const float* __restrict__ ro, const float* __restrict__ dn, long long n_rays, float r,
                                 const float* __restrict__ t_coarse, int n_samples, float* __restrict__ T, int P) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_rays * n_samples) return;
    const long long ray = i / n_samples; const int j = (int)(i - ray * n_samples);
    const float dot = __fadd_rn(__fadd_rn(__fmul_rn(ro[ray * 3], dn[ray * 3]), __fmul_rn(ro[ray * 3 + 1], dn[ray * 3 + 1])),
                                __fmul_rn(ro[ray * 3 + 2], dn[ray * 3 + 2]));
    const float mid = -dot;
    const float near = fmaxf(__fsub_rn(mid, r), 0.f), far = fmaxf(__fadd_rn(mid, r), r);
    const float t = t_coarse[j];
    T[ray * P + j] = __fadd_rn(__fmul_rn(near, __fsub_rn(1.f, t)), __fmul_rn(far, t));
}

// exclusive running product with a double accumulator: out[i] = (float) prod_{j<i} in[j]   (torch.cumprod of [1, in...])
__device__ inline void block_cumprod_excl(const float* in, float* out, int n, double* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int items = ((n + SNT - 1) / SNT) | 1;
    const int beg = min(tid * items, n), end = min(beg + items, n);
    double local = 1.0;
    for (int i = beg; i < end; ++i) local *= (double)in[i];
    double incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double nb = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl *= nb; }
    if (lane == 31) red[warp] = incl;
    double excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.0;
    __syncthreads();
    double woff = 1.0;
    for (int w = 0; w < warp; ++w) woff *= red[w];
    double run = woff * excl;
    float tmp[33];
    for (int i = beg; i < end; ++i) { tmp[i - beg] = (float)run; run *= (double)in[i]; }
    __syncthreads();
    for (int i = beg; i < end; ++i) out[i] = tmp[i - beg];
    __syncthreads();
}

struct NeusArgs {
    float* T; float* S; int P;                    // per-ray arrays, row stride P = n_samples + n_importance
    int n_samples, n_new, n_iters;
    const float* u_det; const float* u_rand; int perturb;   // u_rand [n_iters][n_rays][n_new]
    long long n_rays;
};

// iteration `it` (0..n_iters): merge the n_new depths appended by the previous iteration (torch.sort + gather, neus.py:301-302),
// then -- unless it == n_iters -- build the upsampling weights (neus.py:277-295) and draw n_new depths (296).
__global__ void __launch_bounds__(SNT) neus_upsample_kernel(const NeusArgs a, const int it) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float* D = reinterpret_cast<float*>(smraw); float* Sv = D + a.P; float* X0 = Sv + a.P; float* X1 = X0 + a.P;
    float* UK = X1 + a.P; float* UV = UK + 64;
    __shared__ double red[8];
    const long long ray = blockIdx.x;
    float* Tg = a.T + ray * a.P; float* Sg = a.S + ray * a.P;
    const int n_prev = a.n_samples + (it > 0 ? (it - 1) * a.n_new : 0);
    const int n = a.n_samples + it * a.n_new;
    if (it == 0) {
        for (int i = threadIdx.x; i < n; i += SNT) { D[i] = Tg[i]; Sv[i] = Sg[i]; }
        __syncthreads();
    } else {
        int up_pad = 1; while (up_pad < a.n_new) up_pad <<= 1;
        for (int i = threadIdx.x; i < n_prev; i += SNT) { X0[i] = Tg[i]; X1[i] = Sg[i]; }
        for (int i = threadIdx.x; i < up_pad; i += SNT) { UK[i] = i < a.n_new ? Tg[n_prev + i] : INFINITY; UV[i] = i < a.n_new ? Sg[n_prev + i] : 0.f; }
        __syncthreads();
        bitonic_sort_pairs(UK, UV, up_pad);
        for (int i = threadIdx.x; i < n_prev; i += SNT) {
            const float key = X0[i]; int lo = 0, hi = a.n_new;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (UK[mid] < key) lo = mid + 1; else hi = mid; }
            D[i + lo] = key; Sv[i + lo] = X1[i];
        }
        for (int j = threadIdx.x; j < a.n_new; j += SNT) {
            const float key = UK[j]; int lo = 0, hi = n_prev;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (X0[mid] <= key) lo = mid + 1; else hi = mid; }
            D[j + lo] = key; Sv[j + lo] = UV[j];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += SNT) { Tg[i] = D[i]; Sg[i] = Sv[i]; }
    }
    if (it == a.n_iters) return;
    // ---- weights of the n-1 sections
    const float s = (float)(64 << it);
    for (int i = threadIdx.x; i < n - 1; i += SNT) {
        const float ps = Sv[i], ns = Sv[i + 1], pz = D[i], nz = D[i + 1];
        const float mid_sdf = __fmul_rn(__fadd_rn(ps, ns), 0.5f);
        float dot = __fdiv_rn(__fsub_rn(ns, ps), __fadd_rn(__fsub_rn(nz, pz), 1e-5f));
        float pdot = 0.f;
        if (i > 0) pdot = __fdiv_rn(__fsub_rn(ps, Sv[i - 1]), __fadd_rn(__fsub_rn(pz, D[i - 1]), 1e-5f));
        dot = fminf(fmaxf(fminf(pdot, dot), -10.f), 0.f);
        const float dist = __fsub_rn(nz, pz);
        const float half = __fmul_rn(__fmul_rn(dot, dist), 0.5f);
        const float pe = __fsub_rn(mid_sdf, half), ne = __fadd_rn(mid_sdf, half);
        const float pc = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-__fmul_rn(pe, s))));
        const float nc = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-__fmul_rn(ne, s))));
        const float alpha = __fdiv_rn(__fadd_rn(__fsub_rn(pc, nc), 1e-5f), __fadd_rn(pc, 1e-5f));
        X0[i] = alpha;
        X1[i] = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
    }
    __syncthreads();
    block_cumprod_excl(X1, X1, n - 1, red);                         // alpha_to_w, neus.py:65-78
    for (int i = threadIdx.x; i < n - 1; i += SNT) X1[i] = __fmul_rn(X0[i], X1[i]);
    __syncthreads();
    pdf_to_cdf(X1, X0, n, red);                                     // X0 <- cdf (n entries)
    for (int q = threadIdx.x; q < a.n_new; q += SNT) {
        const float u = a.perturb ? a.u_rand[((long long)it * a.n_rays + ray) * a.n_new + q] : a.u_det[q];
        Tg[n + q] = invert_cdf(D, X0, n, u, nullptr);
    }
}

struct NeusCompositeArgs {
    const float* d_all; const float* sdf; const float* rad; const float* nab; const float* s_dev;
    int P; int white_bkgd; long long n_rays;
    float* rgb; float* depth; float* acc; float* normals; float* alpha_out; float* w_out;
};

// neus.py:322 (sdf_to_alpha), 373-395 (alpha_to_w + integration).  One warp per ray.
__global__ void neus_composite_kernel(const NeusCompositeArgs a) {
    const int lane = threadIdx.x & 31;
    const long long ray = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (ray >= a.n_rays) return;
    const float s = a.s_dev[0];
    const float* d = a.d_all + ray * a.P; const float* sd = a.sdf + ray * a.P;
    const int M = a.P - 1;
    const float* c = a.rad + ray * M * 3; const float* g = a.nab + ray * a.P * 3;
    float r0 = 0.f, r1 = 0.f, r2 = 0.f, wsum = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f;
    for (int pass = 0; pass < 2; ++pass) {
        double carry = 1.0;
        float dsum = 0.f;
        const float den = __fadd_rn(wsum, 1e-10f);
        for (int base = 0; base < M; base += 32) {
            const int i = base + lane;
            float alpha = 0.f;
            if (i < M) {
                const float c0 = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-__fmul_rn(sd[i], s))));
                const float c1 = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-__fmul_rn(sd[i + 1], s))));
                alpha = fmaxf(__fdiv_rn(__fsub_rn(c0, c1), __fadd_rn(c0, 1e-10f)), 0.f);
            }
            double incl = i < M ? (double)__fadd_rn(__fsub_rn(1.f, alpha), 1e-10f) : 1.0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const double nb = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl *= nb; }
            double excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 1.0;
            const float Ti = (float)(carry * excl);
            carry *= __shfl_sync(0xffffffffu, incl, 31);
            if (i < M) {
                const float w = __fmul_rn(alpha, Ti);
                if (pass == 0) {
                    if (a.alpha_out) a.alpha_out[ray * M + i] = alpha;
                    if (a.w_out) a.w_out[ray * M + i] = w;
                    r0 += w * c[i * 3]; r1 += w * c[i * 3 + 1]; r2 += w * c[i * 3 + 2];
                    wsum += w;
                    const float gx = g[i * 3], gy = g[i * 3 + 1], gz = g[i * 3 + 2];
                    const float nrm = fmaxf(sqrtf(gx * gx + gy * gy + gz * gz), 1e-12f);
                    n0 += __fdiv_rn(gx, nrm) * w; n1 += __fdiv_rn(gy, nrm) * w; n2 += __fdiv_rn(gz, nrm) * w;
                } else {
                    const float dm = __fmul_rn(0.5f, __fadd_rn(d[i + 1], d[i]));
                    dsum += __fmul_rn(__fdiv_rn(w, den), dm);
                }
            }
        }
        if (pass == 0) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                r0 += __shfl_xor_sync(0xffffffffu, r0, o); r1 += __shfl_xor_sync(0xffffffffu, r1, o); r2 += __shfl_xor_sync(0xffffffffu, r2, o);
                wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
                n0 += __shfl_xor_sync(0xffffffffu, n0, o); n1 += __shfl_xor_sync(0xffffffffu, n1, o); n2 += __shfl_xor_sync(0xffffffffu, n2, o);
            }
        } else {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
            if (lane == 0) {
                if (a.white_bkgd) { const float w = __fsub_rn(1.f, wsum); r0 += w; r1 += w; r2 += w; }
                a.rgb[ray * 3] = r0; a.rgb[ray * 3 + 1] = r1; a.rgb[ray * 3 + 2] = r2;
                a.depth[ray] = dsum; a.acc[ray] = wsum;
                if (a.normals) { a.normals[ray * 3] = n0; a.normals[ray * 3 + 1] = n1; a.normals[ray * 3 + 2] = n2; }
            }
        }
    }
}

static inline size_t nalign256(size_t x) { return (x + 255) & ~(size_t)255; }
struct NeusWs { size_t dirs, T, S, sdf, nab, rad, scratch, total; };
static NeusWs neus_ws_layout(const NaNeusCfg& c, long long n_rays) {
    const int P = c.n_samples + c.n_importance;
    NeusWs w; size_t o = 0;
    w.dirs = o; o += nalign256((size_t)n_rays * 3 * 4);
    w.T = o; o += nalign256((size_t)n_rays * P * 4);
    w.S = o; o += nalign256((size_t)n_rays * P * 4);
    w.sdf = o; o += nalign256((size_t)n_rays * P * 4);
    w.nab = o; o += nalign256((size_t)n_rays * P * 3 * 4);
    w.rad = o; o += nalign256((size_t)n_rays * (P - 1) * 3 * 4);
    w.scratch = o; o += nalign256(mlp_scratch_bytes());
    w.total = o;
    return w;
}

int preload_neus() {
    NA_PRELOAD(neus_init_kernel);
    NA_PRELOAD(neus_upsample_kernel);
    NA_PRELOAD(neus_composite_kernel);
    return NA_OK;
}

}  // namespace na

using namespace na;

extern "C" size_t na_neus_workspace_bytes(const NaNeusCfg* cfg, int64_t n_rays) {
    if (!cfg || n_rays <= 0) return 0;
    return neus_ws_layout(*cfg, n_rays).total;
}

namespace na {
int train_forward_stash(const EvalJob& fj, const void* packed, int precision, void* train_workspace, size_t train_ws_bytes,
                        float* scratch, size_t scratch_bytes, cudaStream_t stream);                     // csrc/train.cu
size_t train_ws_total(long long n_rays, int P);
}

// train_ws != nullptr: the two final evaluations (P points: sdf + nabla; P - 1 midpoints: radiance) are the forward halves of the split
// training program; their stashes lie one behind the other in the training workspace (csrc/train.cu)
static int neus_render_fwd(const NaNetDesc* desc, const void* packed, const NaNeusCfg* cfg,
                           const float* rays_o, const float* rays_d, int64_t n_rays, const float* s_dev,
                           const float* t_coarse, const float* u_imp, const float* u_rand,
                           const NaNeusOut* out, void* workspace, size_t ws_bytes, void* train_ws, size_t train_ws_bytes, void* stream_) {
    if (!desc || !packed || !cfg || !rays_o || !rays_d || !s_dev || !t_coarse || !u_imp || !out || !workspace || n_rays <= 0) return NA_ERR_BAD_ARG;
    if (!out->rgb || !out->depth || !out->acc) return NA_ERR_BAD_ARG;
    if (cfg->perturb && !u_rand) return NA_ERR_BAD_ARG;
    if (cfg->precision < NA_PRECISION_FP32 || cfg->precision > NA_PRECISION_TC_MIXED) return NA_ERR_UNSUPPORTED;
    if (cfg->n_samples < 2 || cfg->n_upsample_iters < 1 || cfg->n_importance % cfg->n_upsample_iters != 0) return NA_ERR_BAD_ARG;
    const int P = cfg->n_samples + cfg->n_importance, n_new = cfg->n_importance / cfg->n_upsample_iters;
    if (P > 2048 || n_new > 64 || n_new < 1) return NA_ERR_UNSUPPORTED;
    cudaStream_t stream = (cudaStream_t)stream_;
    const NeusWs w = neus_ws_layout(*cfg, n_rays);
    if (ws_bytes < w.total) return NA_ERR_WORKSPACE;
    unsigned char* ws = (unsigned char*)workspace;
    float* dirs = (float*)(ws + w.dirs);
    float* T = out->d_all ? out->d_all : (float*)(ws + w.T);
    float* S = (float*)(ws + w.S);
    float* sdf_f = out->sdf ? out->sdf : (float*)(ws + w.sdf);
    float* nab_f = out->nablas ? out->nablas : (float*)(ws + w.nab);
    float* rad_f = out->radiance ? out->radiance : (float*)(ws + w.rad);
    float* scratch = (float*)(ws + w.scratch);
    const size_t scratch_bytes = w.total - w.scratch;

    NA_TRY(launch_normalize_dirs(rays_d, dirs, n_rays, stream));
    {
        const long long tot = n_rays * cfg->n_samples;
        neus_init_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(rays_o, dirs, n_rays, cfg->bounding_radius, t_coarse,
                                                                              cfg->n_samples, T, P);
        NA_CHECK_LAUNCH();
    }
    EvalJob job = {};
    job.rays_o = rays_o; job.rays_d = dirs; job.n_rows = (int)n_rays; job.P = cfg->n_samples; job.t = T; job.t_stride = P; job.t_off = 0;
    job.o_stride = P; job.o_off = 0; job.sdf = S; job.apply_bg = 0; job.want_full = 0; job.multires_view = desc->multires_view;
    NA_TRY(launch_mlp(job, packed, cfg->precision, scratch, scratch_bytes, stream));                  // neus.py:276

    NeusArgs na_ = {};
    na_.T = T; na_.S = S; na_.P = P; na_.n_samples = cfg->n_samples; na_.n_new = n_new; na_.n_iters = cfg->n_upsample_iters;
    na_.u_det = u_imp; na_.u_rand = u_rand; na_.perturb = cfg->perturb; na_.n_rays = n_rays;
    const size_t smem = ((size_t)4 * P + 128) * sizeof(float);
    NA_TRY(check_cuda(cudaFuncSetAttribute(neus_upsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    for (int it = 0; it <= cfg->n_upsample_iters; ++it) {
        if (it > 0) {
            EvalJob uj = job;
            uj.P = n_new; uj.t_off = cfg->n_samples + (it - 1) * n_new; uj.o_off = uj.t_off;
            NA_TRY(launch_mlp(uj, packed, cfg->precision, scratch, scratch_bytes, stream));          // neus.py:299
        }
        neus_upsample_kernel<<<(unsigned)n_rays, SNT, smem, stream>>>(na_, it);
        NA_CHECK_LAUNCH();
    }
    // sdf + nablas at the P depths (neus.py:320), radiance at the P-1 midpoints through a second SDF pass (324, 111-114)
    EvalJob fj = job;
    fj.P = P; fj.t_off = 0; fj.o_off = 0; fj.sdf = sdf_f; fj.nab = nab_f; fj.rad = nullptr; fj.want_full = 1;
    const size_t ws_a = train_ws ? train_ws_total(n_rays, P) : 0;
    if (train_ws) {
        if (train_ws_bytes < ws_a + train_ws_total(n_rays, P - 1)) return NA_ERR_WORKSPACE;
        NA_TRY(train_forward_stash(fj, packed, cfg->precision, train_ws, ws_a, scratch, scratch_bytes, stream));
    } else NA_TRY(launch_mlp(fj, packed, cfg->precision, scratch, scratch_bytes, stream));
    EvalJob mj = job;
    mj.P = P - 1; mj.t_off = 0; mj.midpoints = 1; mj.o_stride = P - 1; mj.o_off = 0; mj.sdf = nullptr; mj.nab = nullptr; mj.rad = rad_f; mj.want_full = 1;
    if (train_ws) NA_TRY(train_forward_stash(mj, packed, cfg->precision, (unsigned char*)train_ws + ws_a, train_ws_bytes - ws_a, scratch, scratch_bytes, stream));
    else          NA_TRY(launch_mlp(mj, packed, cfg->precision, scratch, scratch_bytes, stream));
    NeusCompositeArgs ca = {};
    ca.d_all = T; ca.sdf = sdf_f; ca.rad = rad_f; ca.nab = nab_f; ca.s_dev = s_dev; ca.P = P; ca.white_bkgd = cfg->white_bkgd; ca.n_rays = n_rays;
    ca.rgb = out->rgb; ca.depth = out->depth; ca.acc = out->acc; ca.normals = out->normals; ca.alpha_out = out->alpha; ca.w_out = out->weights;
    const long long threads = n_rays * 32;
    neus_composite_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(ca);
    NA_CHECK_LAUNCH();
    return NA_OK;
}

extern "C" int na_neus_render_fwd(const NaNetDesc* desc, const void* packed, const NaNeusCfg* cfg,
                                  const float* rays_o, const float* rays_d, int64_t n_rays, const float* s_dev,
                                  const float* t_coarse, const float* u_imp, const float* u_rand,
                                  const NaNeusOut* out, void* workspace, size_t ws_bytes, void* stream_) {
    return neus_render_fwd(desc, packed, cfg, rays_o, rays_d, n_rays, s_dev, t_coarse, u_imp, u_rand, out, workspace, ws_bytes, nullptr, 0, stream_);
}

extern "C" int na_neus_render_fwd_train(const NaNetDesc* desc, const void* packed, const NaNeusCfg* cfg,
                                        const float* rays_o, const float* rays_d, int64_t n_rays, const float* s_dev,
                                        const float* t_coarse, const float* u_imp, const float* u_rand,
                                        const NaNeusOut* out, void* workspace, size_t ws_bytes, void* train_workspace,
                                        size_t train_ws_bytes, void* stream_) {
    if (!train_workspace || !out || !out->d_all || !out->sdf || !out->radiance || !out->nablas) return NA_ERR_BAD_ARG;     // detailed outputs: the backward reads them
    return neus_render_fwd(desc, packed, cfg, rays_o, rays_d, n_rays, s_dev, t_coarse, u_imp, u_rand, out, workspace, ws_bytes,
                           train_workspace, train_ws_bytes, stream_);
}
