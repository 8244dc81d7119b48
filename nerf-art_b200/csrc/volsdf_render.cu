// VolSDF renderer: volsdf.volume_render (models/frameworks/volsdf.py:389-615) as a short sequence of launches
//   rays_prep -> [MLP sdf @ d_init] -> sampler(it=0) -> { [MLP sdf @ new depths of active rays] -> sampler(it) } x max_iter
//   -> [MLP full @ d_all] -> composite
// The per-ray error-bound sampler (fine_sample, volsdf.py:97-302) runs one CTA per ray with every
// per-ray array in shared memory; rays that are still unconverged are compacted into a device-side work list, so the
// host never synchronises.
#include "sampler.cuh"

namespace na {

int launch_mlp(const EvalJob& job, const void* packed, int precision, float* scratch, size_t scratch_bytes, cudaStream_t stream);
size_t mlp_scratch_bytes();

// -----------------------------------------------------------------------------------------------
__global__ void normalize_dirs_kernel(const float* __restrict__ d_in, float* __restrict__ d_out, long long n) {
    // F.normalize(rays_d, dim=-1) (volsdf.py:442, neus.py:199): v / max(||v||, 1e-12)
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = d_in[i * 3], y = d_in[i * 3 + 1], z = d_in[i * 3 + 2];
    const float nrm = fmaxf(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z))), 1e-12f);
    d_out[i * 3] = __fdiv_rn(x, nrm); d_out[i * 3 + 1] = __fdiv_rn(y, nrm); d_out[i * 3 + 2] = __fdiv_rn(z, nrm);
}

int launch_normalize_dirs(const float* d_in, float* d_out, long long n, cudaStream_t stream) {
    normalize_dirs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_in, d_out, n);
    NA_CHECK_LAUNCH();
    return NA_OK;
}

__global__ void init_depths_kernel(float* __restrict__ T, long long cap, int n0, const float* __restrict__ t_init,
                                   float near, float far, long long n_rays) {
    // d_init = nears*(1-t) + fars*t  (volsdf.py:483-484), each op rounded like the tensor expression
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_rays * n0) return;
    const long long ray = i / n0; const int j = (int)(i - ray * n0);
    const float t = t_init[j];
    T[ray * cap + j] = __fadd_rn(__fmul_rn(near, __fsub_rn(1.f, t)), __fmul_rn(far, t));
}

struct SamplerArgs {
    float* T; float* S; long long cap;              // per-ray depth / sdf arrays, row stride cap
    int n0, n_up, n_imp, n_samples, P;
    int max_iter, max_bisect;
    float near, far, eps;
    const float* alpha_beta;                         // device [2]
    const float* t_coarse; const float* u_up; const float* u_imp; const float* u_final; int perturb;
    float* beta_cur;                                 // [n_rays] current beta+
    float* d_all;                                    // [n_rays, P]
    float* beta_map; float* iter_usage;              // [n_rays]
    const int* list_in; const int* count_in;         // active rays of this iteration (it >= 1)
    int* list_out; int* count_out;                   // rays still active afterwards
    long long n_rays;
    int cap_pad, up_pad;                             // shared array lengths
};

// opacity_invert_cdf_sample (volsdf.py:122-136) + d_all = sort(cat(d_coarse, d_fine)) (volsdf.py:501-502)
__device__ inline void finalize_ray(const SamplerArgs& a, long long ray, const float* d, const float* s, int n,
                                    float alpha, float beta, float iter_usage, float* X0, float* X1, float* KEY,
                                    double* red) {
    compute_Rt(d, s, n, alpha, beta, X0, red);
    for (int i = threadIdx.x; i < n; i += SNT) X1[i] = i == 0 ? 0.f : __fsub_rn(1.f, expf(-X0[i - 1]));
    __syncthreads();
    int ppad = 1; while (ppad < a.P) ppad <<= 1;
    for (int q = threadIdx.x; q < ppad; q += SNT) {
        float v = INFINITY;
        if (q < a.n_samples) {
            const float t = a.t_coarse[q];
            v = __fadd_rn(__fmul_rn(a.near, __fsub_rn(1.f, t)), __fmul_rn(a.far, t));
        } else if (q < a.P) {
            const int qi = q - a.n_samples;
            const float u = a.perturb ? a.u_final[ray * a.n_imp + qi] : a.u_imp[qi];
            v = invert_cdf(d, X1, n, u, nullptr);
        }
        KEY[q] = v;
    }
    __syncthreads();
    bitonic_sort_keys(KEY, ppad);
    for (int q = threadIdx.x; q < a.P; q += SNT) a.d_all[ray * a.P + q] = KEY[q];
    if (threadIdx.x == 0) { a.iter_usage[ray] = iter_usage; a.beta_map[ray] = beta; }
    __syncthreads();
}

__global__ void __launch_bounds__(SNT) volsdf_sampler_kernel(const SamplerArgs a, const int it) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float* X0 = reinterpret_cast<float*>(smraw);            // old d  -> R_t
    float* X1 = X0 + a.cap_pad;                              // old s  -> bounds / cdf
    float* D = X1 + a.cap_pad;                               // merged depths
    float* Sv = D + a.cap_pad;                               // merged sdf
    float* UK = Sv + a.cap_pad;                              // new depths (sorted)
    float* UV = UK + a.up_pad;
    float* KEY = UV + a.up_pad;                              // final d_all sort buffer (>= pow2(P))
    __shared__ double red[8];
    __shared__ float redf[8];
    const float alpha_net = a.alpha_beta[0], beta_net = a.alpha_beta[1];
    const long long n_work = it == 0 ? a.n_rays : (long long)*a.count_in;

    for (long long widx = blockIdx.x; widx < n_work; widx += gridDim.x) {
        const long long ray = it == 0 ? widx : a.list_in[widx];
        float* Tg = a.T + ray * a.cap; float* Sg = a.S + ray * a.cap;
        const int n_prev = a.n0 + (it > 0 ? (it - 1) * a.n_up : 0);
        const int n = a.n0 + it * a.n_up;
        float beta_plus;
        if (it == 0) {
            for (int i = threadIdx.x; i < n; i += SNT) { D[i] = Tg[i]; Sv[i] = Sg[i]; }
            // beta+ init, volsdf.py:149:  sqrt(far^2 / (4 (N-1) log(1+eps)))
            const float den = (float)(4.0 * (double)(a.n0 - 1) * log(1.0 + (double)a.eps));
            beta_plus = sqrtf(__fdiv_rn(__fmul_rn(a.far, a.far), den));
            __syncthreads();
        } else {
            // ---- merge the n_up new (depth, sdf) pairs into the sorted arrays: torch.sort + gather, volsdf.py:211-228
            for (int i = threadIdx.x; i < n_prev; i += SNT) { X0[i] = Tg[i]; X1[i] = Sg[i]; }
            for (int i = threadIdx.x; i < a.up_pad; i += SNT) {
                UK[i] = i < a.n_up ? Tg[n_prev + i] : INFINITY; UV[i] = i < a.n_up ? Sg[n_prev + i] : 0.f;
            }
            beta_plus = a.beta_cur[ray];
            __syncthreads();
            bitonic_sort_pairs(UK, UV, a.up_pad);
            for (int i = threadIdx.x; i < n_prev; i += SNT) {            // old element i: after all new elements strictly smaller
                const float key = X0[i];
                int lo = 0, hi = a.n_up;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (UK[mid] < key) lo = mid + 1; else hi = mid; }
                D[i + lo] = key; Sv[i + lo] = X1[i];
            }
            for (int j = threadIdx.x; j < a.n_up; j += SNT) {            // new element j: after all old elements <= it
                const float key = UK[j];
                int lo = 0, hi = n_prev;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (X0[mid] <= key) lo = mid + 1; else hi = mid; }
                D[j + lo] = key; Sv[j + lo] = UV[j];
            }
            __syncthreads();
        }
        // ---- bound with the network's beta: converged?  (volsdf.py:162-163 / 240-251)
        const float mx_net = error_bound(D, Sv, n, alpha_net, beta_net, X0, X1, false, red, redf);
        if (!(mx_net > a.eps)) {
            finalize_ray(a, ray, D, Sv, n, alpha_net, beta_net, (float)it, X0, X1, KEY, red);
            continue;
        }
        if (it > 0) {
            // ---- bisection for the smallest beta+ with bound <= eps (volsdf.py:260-275)
            float b_left = beta_net, b_right = beta_plus;
            for (int step = 0; step < a.max_bisect; ++step) {
                const float b_mid = __fmul_rn(0.5f, __fadd_rn(b_left, b_right));
                const float a_mid = __fdiv_rn(1.f, b_mid);
                const float mx = error_bound(D, Sv, n, a_mid, b_mid, X0, X1, false, red, redf);
                if (mx <= a.eps) b_right = b_mid; else b_left = b_mid;
            }
            beta_plus = b_right;
        }
        if (it == a.max_iter) {
            // never converged: sample with the last beta+ (volsdf.py:294-300)
            finalize_ray(a, ray, D, Sv, n, __fdiv_rn(1.f, beta_plus), beta_plus, -1.f, X0, X1, KEY, red);
            continue;
        }
        // ---- bounds with beta+ (clamped from the second time on, volsdf.py:168 vs 280-282) and upsample proportional to them
        error_bound(D, Sv, n, __fdiv_rn(1.f, beta_plus), beta_plus, X0, X1, it > 0, red, redf);
        pdf_to_cdf(X1, X0, n, red);                                      // X0 <- cdf (n entries)
        // sample_pdf(d, bounds, n_up+2, det=True)[1:-1]  (volsdf.py:196)
        for (int q = threadIdx.x; q < a.n_up; q += SNT) Tg[n + q] = invert_cdf(D, X0, n, a.u_up[q + 1], nullptr);
        if (it > 0) for (int i = threadIdx.x; i < n; i += SNT) { Tg[i] = D[i]; Sg[i] = Sv[i]; }
        if (threadIdx.x == 0) {
            a.beta_cur[ray] = beta_plus;
            const int slot = atomicAdd(a.count_out, 1);
            a.list_out[slot] = (int)ray;
        }
        __syncthreads();
    }
}

// -----------------------------------------------------------------------------------------------
// Ray integration, volsdf.py:540-576.  One warp per ray.
struct CompositeArgs {
    const float* d_all; const float* sdf; const float* rad; const float* nab; const float* alpha_beta;
    int P; int white_bkgd; long long n_rays;
    float* rgb; float* depth; float* acc; float* normals; float* sigma_out; float* tau_out;
};

__global__ void volsdf_composite_kernel(const CompositeArgs a) {
    const int lane = threadIdx.x & 31;
    const long long ray = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (ray >= a.n_rays) return;
    const float alpha = a.alpha_beta[0], beta = a.alpha_beta[1];
    const float* d = a.d_all + ray * a.P; const float* s = a.sdf + ray * a.P;
    const float* c = a.rad + ray * a.P * 3; const float* g = a.nab ? a.nab + ray * a.P * 3 : nullptr;
    double carry = 1.0;                       // prod_{j<i} p_j, double accumulator like torch.cumprod on CPU
    float r0 = 0.f, r1 = 0.f, r2 = 0.f, tsum = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f;
    const int M = a.P - 1;
    // pass 1: tau, rgb, acc, normals
    for (int base = 0; base < M; base += 32) {
        const int i = base + lane;
        float p = 1.f, sg = 0.f;
        if (i < M) {
            sg = sdf_to_sigma(s[i], alpha, beta);
            p = expf(-fmaxf(__fmul_rn(sg, __fsub_rn(d[i + 1], d[i])), 0.f));
        }
        // inclusive product scan over the warp
        double incl = (double)p;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double nb = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl *= nb; }
        double excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0;
        // cumprod([1, p_0, p_1, ...])[i] rounded to fp32 like torch (double accumulate, float store)
        const float Ti = (float)(carry * excl);
        carry *= __shfl_sync(0xffffffffu, incl, 31);
        if (i < M) {
            const float tau = __fmul_rn(__fadd_rn(__fsub_rn(1.f, p), 1e-10f), Ti);
            if (a.tau_out) a.tau_out[ray * M + i] = tau;
            r0 += tau * c[i * 3]; r1 += tau * c[i * 3 + 1]; r2 += tau * c[i * 3 + 2];
            tsum += tau;
            if (g) {
                const float gx = g[i * 3], gy = g[i * 3 + 1], gz = g[i * 3 + 2];
                const float nrm = fmaxf(sqrtf(gx * gx + gy * gy + gz * gz), 1e-12f);
                n0 += __fdiv_rn(gx, nrm) * tau; n1 += __fdiv_rn(gy, nrm) * tau; n2 += __fdiv_rn(gz, nrm) * tau;
            }
        }
        if (a.sigma_out && i < M) a.sigma_out[ray * a.P + i] = sg;
    }
    if (a.sigma_out && lane == 0) a.sigma_out[ray * a.P + M] = sdf_to_sigma(s[M], alpha, beta);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        r0 += __shfl_xor_sync(0xffffffffu, r0, o); r1 += __shfl_xor_sync(0xffffffffu, r1, o); r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        tsum += __shfl_xor_sync(0xffffffffu, tsum, o);
        n0 += __shfl_xor_sync(0xffffffffu, n0, o); n1 += __shfl_xor_sync(0xffffffffu, n1, o); n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    }
    // pass 2: depth = sum tau_i / (sum tau + 1e-10) * d_i  (volsdf.py:560): recompute tau (cheap) to avoid a per-ray buffer
    const float inv_den = __fadd_rn(tsum, 1e-10f);
    carry = 1.0;
    float dsum = 0.f;
    for (int base = 0; base < M; base += 32) {
        const int i = base + lane;
        float p = 1.f;
        if (i < M) p = expf(-fmaxf(__fmul_rn(sdf_to_sigma(s[i], alpha, beta), __fsub_rn(d[i + 1], d[i])), 0.f));
        double incl = (double)p;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double nb = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl *= nb; }
        double excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0;
        const float Ti = (float)(carry * excl);
        carry *= __shfl_sync(0xffffffffu, incl, 31);
        if (i < M) dsum += __fmul_rn(__fdiv_rn(__fmul_rn(__fadd_rn(__fsub_rn(1.f, p), 1e-10f), Ti), inv_den), d[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    if (lane == 0) {
        if (a.white_bkgd) { const float w = __fsub_rn(1.f, tsum); r0 += w; r1 += w; r2 += w; }
        a.rgb[ray * 3] = r0; a.rgb[ray * 3 + 1] = r1; a.rgb[ray * 3 + 2] = r2;
        a.depth[ray] = dsum; a.acc[ray] = tsum;
        if (a.normals) { a.normals[ray * 3] = n0; a.normals[ray * 3 + 1] = n1; a.normals[ray * 3 + 2] = n2; }
    }
}

// -----------------------------------------------------------------------------------------------
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
constexpr long long RAY_CHUNK = 65536;

struct VolsdfWs {
    size_t dirs, T, S, beta_cur, listA, listB, counts, d_all, sdf, rad, nab, scratch, total;
};
static VolsdfWs volsdf_ws_layout(const NaVolsdfCfg& c, long long n_rays) {
    const long long chunk = n_rays < RAY_CHUNK ? n_rays : RAY_CHUNK;
    const long long cap = (long long)4 * c.n_samples * (1 + c.max_upsample_steps);
    const int P = c.n_samples + c.n_importance;
    VolsdfWs w; size_t o = 0;
    w.dirs = o; o += align256((size_t)n_rays * 3 * 4);
    w.T = o; o += align256((size_t)chunk * cap * 4);
    w.S = o; o += align256((size_t)chunk * cap * 4);
    w.beta_cur = o; o += align256((size_t)chunk * 4);
    w.listA = o; o += align256((size_t)chunk * 4);
    w.listB = o; o += align256((size_t)chunk * 4);
    w.counts = o; o += align256(64 * 4);
    w.d_all = o; o += align256((size_t)n_rays * P * 4);
    w.sdf = o; o += align256((size_t)n_rays * P * 4);
    w.rad = o; o += align256((size_t)n_rays * P * 3 * 4);
    w.nab = o; o += align256((size_t)n_rays * P * 3 * 4);
    w.scratch = o; o += align256(mlp_scratch_bytes());
    w.total = o;
    return w;
}

int preload_volsdf() {
    NA_PRELOAD(normalize_dirs_kernel);
    NA_PRELOAD(init_depths_kernel);
    NA_PRELOAD(volsdf_sampler_kernel);
    NA_PRELOAD(volsdf_composite_kernel);
    return NA_OK;
}

}  // namespace na

using namespace na;

extern "C" size_t na_volsdf_workspace_bytes(const NaVolsdfCfg* cfg, int64_t n_rays) {
    if (!cfg || n_rays <= 0) return 0;
    return volsdf_ws_layout(*cfg, n_rays).total;
}

namespace na {
int train_forward_stash(const EvalJob& fj, const void* packed, int precision, void* train_workspace, size_t train_ws_bytes,
                        float* scratch, size_t scratch_bytes, cudaStream_t stream);                     // csrc/train.cu
}

// train_ws != nullptr: the final full evaluation is the forward half of the split training program (it also fills the training
// workspace with the forward stash; csrc/train.cu)
static int volsdf_render_fwd(const NaNetDesc* desc, const void* packed, const NaVolsdfCfg* cfg,
                             const float* rays_o, const float* rays_d, int64_t n_rays, const float* alpha_beta,
                             const float* t_coarse, const float* t_init, const float* u_up, const float* u_imp,
                             const float* u_final, const NaVolsdfOut* out, void* workspace, size_t ws_bytes, void* train_ws,
                             size_t train_ws_bytes, void* stream_) {
    if (!desc || !packed || !cfg || !rays_o || !rays_d || !alpha_beta || !t_coarse || !t_init || !u_up || !u_imp || !out || !workspace)
        return NA_ERR_BAD_ARG;
    if (n_rays <= 0) return NA_ERR_BAD_ARG;
    if (!out->rgb || !out->depth || !out->acc || !out->beta_map || !out->iter_usage) return NA_ERR_BAD_ARG;
    if (cfg->perturb && !u_final) return NA_ERR_BAD_ARG;
    if (desc->framework != NA_FRAMEWORK_VOLSDF) return NA_ERR_BAD_ARG;
    if (cfg->precision < NA_PRECISION_FP32 || cfg->precision > NA_PRECISION_TC_MIXED) return NA_ERR_UNSUPPORTED;
    if (cfg->n_samples < 2 || cfg->n_importance < 1 || cfg->max_upsample_steps < 0 || cfg->max_bisection_steps < 0) return NA_ERR_BAD_ARG;
    const int n0 = 4 * cfg->n_samples, n_up = n0, P = cfg->n_samples + cfg->n_importance;
    const long long cap = (long long)n0 * (1 + cfg->max_upsample_steps);
    if (cap > MAX_CAP || P > 2048) return NA_ERR_UNSUPPORTED;
    cudaStream_t stream = (cudaStream_t)stream_;
    const VolsdfWs w = volsdf_ws_layout(*cfg, n_rays);
    if (ws_bytes < w.total) return NA_ERR_WORKSPACE;
    unsigned char* ws = (unsigned char*)workspace;
    float* dirs = (float*)(ws + w.dirs);
    float* T = (float*)(ws + w.T); float* S = (float*)(ws + w.S);
    float* beta_cur = (float*)(ws + w.beta_cur);
    int* lists[2] = {(int*)(ws + w.listA), (int*)(ws + w.listB)};
    int* counts = (int*)(ws + w.counts);
    float* d_all = out->d_vals ? out->d_vals : (float*)(ws + w.d_all);
    float* sdf_f = out->sdf ? out->sdf : (float*)(ws + w.sdf);
    float* rad_f = out->radiance ? out->radiance : (float*)(ws + w.rad);
    float* nab_f = out->nablas ? out->nablas : (float*)(ws + w.nab);
    float* scratch = (float*)(ws + w.scratch);
    const size_t scratch_bytes = w.total - w.scratch;

    NA_TRY(launch_normalize_dirs(rays_d, dirs, n_rays, stream));

    int cap_pad = (int)cap, up_pad = 1; while (up_pad < n_up) up_pad <<= 1;
    int ppad = 1; while (ppad < P) ppad <<= 1;
    const size_t samp_smem = ((size_t)4 * cap_pad + 2 * up_pad + ppad) * sizeof(float);
    NA_TRY(check_cuda(cudaFuncSetAttribute(volsdf_sampler_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)samp_smem)));

    for (long long c0 = 0; c0 < n_rays; c0 += RAY_CHUNK) {
        const long long nr = (n_rays - c0) < RAY_CHUNK ? (n_rays - c0) : RAY_CHUNK;
        NA_TRY(check_cuda(cudaMemsetAsync(counts, 0, 64 * sizeof(int), stream)));
        {
            const long long tot = nr * n0;
            init_depths_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(T, cap, n0, t_init, cfg->near, cfg->far, nr);
            NA_CHECK_LAUNCH();
        }
        EvalJob job = {};
        job.rays_o = rays_o + c0 * 3; job.rays_d = dirs + c0 * 3;
        job.n_rows = (int)nr; job.P = n0; job.t = T; job.t_stride = cap; job.t_off = 0;
        job.o_stride = cap; job.o_off = 0; job.sdf = S;
        job.apply_bg = 1; job.bound_r = desc->bounding_radius; job.want_full = 0; job.multires_view = desc->multires_view;
        NA_TRY(launch_mlp(job, packed, cfg->precision, scratch, scratch_bytes, stream));

        SamplerArgs sa = {};
        sa.T = T; sa.S = S; sa.cap = cap; sa.n0 = n0; sa.n_up = n_up; sa.n_imp = cfg->n_importance; sa.n_samples = cfg->n_samples; sa.P = P;
        sa.max_iter = cfg->max_upsample_steps; sa.max_bisect = cfg->max_bisection_steps;
        sa.near = cfg->near; sa.far = cfg->far; sa.eps = cfg->epsilon;
        sa.alpha_beta = alpha_beta; sa.t_coarse = t_coarse; sa.u_up = u_up; sa.u_imp = u_imp;
        sa.u_final = u_final ? u_final + c0 * cfg->n_importance : nullptr; sa.perturb = cfg->perturb;
        sa.beta_cur = beta_cur; sa.d_all = d_all + c0 * P; sa.beta_map = out->beta_map + c0; sa.iter_usage = out->iter_usage + c0;
        sa.n_rays = nr; sa.cap_pad = cap_pad; sa.up_pad = up_pad;
        const int grid_loop = (int)(nr < (long long)num_sms() * 8 ? nr : (long long)num_sms() * 8);
        for (int it = 0; it <= cfg->max_upsample_steps; ++it) {
            // counts[it] = number of rays active at iteration it+1 ; lists ping-pong
            sa.list_in = lists[(it + 1) & 1]; sa.count_in = it > 0 ? counts + (it - 1) : nullptr;
            sa.list_out = lists[it & 1]; sa.count_out = counts + it;
            if (it > 0) {
                EvalJob uj = job;
                uj.row_ids = sa.list_in; uj.n_rows_dev = sa.count_in; uj.n_rows = (int)nr; uj.P = n_up;
                uj.t_off = n0 + (it - 1) * n_up; uj.o_off = uj.t_off;
                NA_TRY(launch_mlp(uj, packed, cfg->precision, scratch, scratch_bytes, stream));
            }
            volsdf_sampler_kernel<<<it == 0 ? (unsigned)nr : (unsigned)grid_loop, SNT, samp_smem, stream>>>(sa, it);
            NA_CHECK_LAUNCH();
        }
    }
    // ---- full evaluation at the P merged depths (volsdf.py:503-514) and integration (540-576)
    {
        EvalJob fj = {};
        fj.rays_o = rays_o; fj.rays_d = dirs; fj.n_rows = (int)n_rays; fj.P = P; fj.t = d_all; fj.t_stride = P; fj.t_off = 0;
        fj.o_stride = P; fj.o_off = 0; fj.sdf = sdf_f; fj.rad = rad_f; fj.nab = nab_f;
        fj.apply_bg = 1; fj.bound_r = desc->bounding_radius; fj.want_full = 1; fj.multires_view = desc->multires_view;
        if (train_ws) NA_TRY(train_forward_stash(fj, packed, cfg->precision, train_ws, train_ws_bytes, scratch, scratch_bytes, stream));
        else          NA_TRY(launch_mlp(fj, packed, cfg->precision, scratch, scratch_bytes, stream));
        CompositeArgs ca = {};
        ca.d_all = d_all; ca.sdf = sdf_f; ca.rad = rad_f; ca.nab = out->normals ? nab_f : nullptr; ca.alpha_beta = alpha_beta;
        ca.P = P; ca.white_bkgd = cfg->white_bkgd; ca.n_rays = n_rays;
        ca.rgb = out->rgb; ca.depth = out->depth; ca.acc = out->acc; ca.normals = out->normals;
        ca.sigma_out = out->sigma; ca.tau_out = out->tau;
        const long long threads = n_rays * 32;
        volsdf_composite_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(ca);
        NA_CHECK_LAUNCH();
    }
    return NA_OK;
}

extern "C" int na_volsdf_render_fwd(const NaNetDesc* desc, const void* packed, const NaVolsdfCfg* cfg,
                                    const float* rays_o, const float* rays_d, int64_t n_rays, const float* alpha_beta,
                                    const float* t_coarse, const float* t_init, const float* u_up, const float* u_imp,
                                    const float* u_final, const NaVolsdfOut* out, void* workspace, size_t ws_bytes, void* stream_) {
    return volsdf_render_fwd(desc, packed, cfg, rays_o, rays_d, n_rays, alpha_beta, t_coarse, t_init, u_up, u_imp, u_final, out, workspace,
                             ws_bytes, nullptr, 0, stream_);
}

extern "C" int na_volsdf_render_fwd_train(const NaNetDesc* desc, const void* packed, const NaVolsdfCfg* cfg,
                                          const float* rays_o, const float* rays_d, int64_t n_rays, const float* alpha_beta,
                                          const float* t_coarse, const float* t_init, const float* u_up, const float* u_imp,
                                          const float* u_final, const NaVolsdfOut* out, void* workspace, size_t ws_bytes,
                                          void* train_workspace, size_t train_ws_bytes, void* stream_) {
    if (!train_workspace || !out || !out->d_vals || !out->sdf || !out->radiance || !out->nablas) return NA_ERR_BAD_ARG;   // detailed outputs: the backward reads them
    return volsdf_render_fwd(desc, packed, cfg, rays_o, rays_d, n_rays, alpha_beta, t_coarse, t_init, u_up, u_imp, u_final, out, workspace,
                             ws_bytes, train_workspace, train_ws_bytes, stream_);
}
