// Surface rendering: models/ray_casting.py of the reference as a short sequence of launches around the fused SDF kernel.
//   root finding  (root_finding_surface_points 35-160 + run_secant_method 11-30):
//       march depths -> [MLP sdf @ N_steps points per ray] -> first-sign-change scan (one warp per ray, active rays compacted)
//       -> { [MLP sdf @ the secant point of every active ray] -> secant update } x N_secant_steps -> finalize
//   sphere tracing (sphere_tracing_surface_points 163-184): { [MLP sdf @ current point] -> step } x N_iters
//   surface_render (187-263): normalise dirs -> ray cast -> [MLP full @ the hit points] -> mask colours / normals
// Every arithmetic expression is rounded op by op like the tensor expression it replaces.
#include "common.cuh"

namespace na {

int launch_mlp(const EvalJob& job, const void* packed, int precision, float* scratch, size_t scratch_bytes, cudaStream_t stream);
size_t mlp_scratch_bytes();
int launch_normalize_dirs(const float* d_in, float* d_out, long long n, cudaStream_t stream);

// d_proposal = near * (1 - t) + far * t   (ray_casting.py:77)
__global__ void march_depths_kernel(float* __restrict__ D, int n_steps, const float* __restrict__ t_steps, float near, float far, long long n_rays) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_rays * n_steps) return;
    const float t = t_steps[i % n_steps];
    D[i] = __fadd_rn(__fmul_rn(near, __fsub_rn(1.f, t)), __fmul_rn(far, t));
}

struct RootState {
    float* d_low; float* f_low; float* d_high; float* f_high; float* d_pred;     // [n_rays]
    unsigned char* mask; unsigned char* mask_sign; unsigned char* mask0;        // [n_rays]
    int* list; int* count;                                                       // active (masked) rays
};

__device__ __forceinline__ float secant_point(float f_low, float f_high, float d_low, float d_high) {
    // - f_low * (d_high - d_low) / (f_high - f_low) + d_low   (ray_casting.py:16,29)
    return __fadd_rn(__fdiv_rn(__fmul_rn(-f_low, __fsub_rn(d_high, d_low)), __fsub_rn(f_high, f_low)), d_low);
}

// one warp per ray: val = sdf - tau; the first i with val[i] * val[i+1] < 0 is where min(sign * (N - i)) sits (93-102)
__global__ void root_scan_kernel(const float* __restrict__ D, const float* __restrict__ VAL, int n_steps, float tau, long long n_rays, RootState st) {
    const long long ray = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (ray >= n_rays) return;
    const float* v = VAL + ray * n_steps;
    int first = n_steps;                                                   // index of the first sign change (product < 0)
    for (int base = 0; base < n_steps - 1 && first == n_steps; base += 32) {
        const int i = base + lane;
        bool neg = false;
        if (i < n_steps - 1) neg = __fmul_rn(__fsub_rn(v[i], tau), __fsub_rn(v[i + 1], tau)) < 0.f;
        const unsigned b = __ballot_sync(0xffffffffu, neg);
        if (b) first = base + __ffs(b) - 1;
    }
    if (lane != 0) return;
    const bool sign_change = first < n_steps;
    const bool m0 = __fsub_rn(v[0], tau) > 0.f;                            // the first point is not occupied (85)
    bool m = false;
    if (sign_change) {
        const float fh = __fsub_rn(v[first], tau);
        m = m0 && fh > 0.f;                                                // first change goes from outside (+) to inside (-) (108-110)
        if (m) {
            const int i2 = min(first + 1, n_steps - 1);
            const float dh = D[ray * n_steps + first], dl = D[ray * n_steps + i2], fl = __fsub_rn(v[i2], tau);
            st.d_high[ray] = dh; st.f_high[ray] = fh; st.d_low[ray] = dl; st.f_low[ray] = fl;
            st.d_pred[ray] = secant_point(fl, fh, dl, dh);
            st.list[atomicAdd(st.count, 1)] = (int)ray;
        }
    }
    st.mask[ray] = m; st.mask_sign[ray] = sign_change; st.mask0[ray] = m0;
}

// one secant step for every active ray (17-29)
__global__ void secant_update_kernel(const float* __restrict__ FMID, float tau, RootState st) {
    const int n = *st.count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int ray = st.list[i];
        const float f_mid = __fsub_rn(FMID[ray], tau), dp = st.d_pred[ray];
        float dl = st.d_low[ray], fl = st.f_low[ray], dh = st.d_high[ray], fh = st.f_high[ray];
        if (f_mid < 0.f) { dl = dp; fl = f_mid; st.d_low[ray] = dl; st.f_low[ray] = fl; }
        else { dh = dp; fh = f_mid; st.d_high[ray] = dh; st.f_high[ray] = fh; }
        st.d_pred[ray] = secant_point(fl, fh, dl, dh);
    }
}

// pt_pred / d_pred_out (135-150)
__global__ void root_finalize_kernel(const float* __restrict__ ro, const float* __restrict__ dirs, RootState st, float far, int fill_inf, int has_secant,
                                     long long n_rays, float* __restrict__ depth, float* __restrict__ pts) {
    const long long ray = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    float p[3] = {1.f, 1.f, 1.f}, d;
    if (st.mask[ray]) {
        d = has_secant ? st.d_pred[ray] : 1.f;                           // method != 'secant': d_pred = ones (130)
#pragma unroll
        for (int c = 0; c < 3; ++c) p[c] = __fadd_rn(ro[ray * 3 + c], __fmul_rn(d, dirs[ray * 3 + c]));
    } else {
        d = fill_inf ? __int_as_float(0x7f800000) : far;
    }
    if (!st.mask0[ray]) d = 0.f;                                         // the 0-th point is occupied: depth 0 (149)
    depth[ray] = d;
    pts[ray * 3] = p[0]; pts[ray * 3 + 1] = p[1]; pts[ray * 3 + 2] = p[2];
}

__global__ void sphere_init_kernel(float* __restrict__ d, unsigned char* __restrict__ mask, float near, long long n_rays) {
    const long long ray = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    d[ray] = __fmul_rn(1.f, near); mask[ray] = 1;
}
// d_preds[mask] += sdf[mask]; mask[d > far] = False; mask[d < 0] = False  (177-181)
__global__ void sphere_step_kernel(float* __restrict__ d, unsigned char* __restrict__ mask, const float* __restrict__ val, float far, long long n_rays) {
    const long long ray = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    float dd = d[ray];
    unsigned char m = mask[ray];
    if (m) { dd = __fadd_rn(dd, val[ray]); d[ray] = dd; }
    if (dd > far || dd < 0.f) m = 0;
    mask[ray] = m;
}
__global__ void sphere_points_kernel(const float* __restrict__ ro, const float* __restrict__ dirs, const float* __restrict__ d, long long n_rays,
                                     float* __restrict__ pts) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_rays * 3) return;
    pts[i] = __fadd_rn(ro[i], __fmul_rn(dirs[i], d[i / 3]));            // rays_o + rays_d * d_preds[..., None]  (183)
}

// color[~mask] = 0 ; normals = F.normalize(nablas) ; normals[~mask] = 0   (236, 257-259)
__global__ void surface_finish_kernel(float* __restrict__ rgb, const float* __restrict__ nab, const unsigned char* __restrict__ mask,
                                      float* __restrict__ normals, long long n_rays) {
    const long long ray = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    const bool m = mask[ray] != 0;
    if (!m) { rgb[ray * 3] = 0.f; rgb[ray * 3 + 1] = 0.f; rgb[ray * 3 + 2] = 0.f; }
    if (normals) {
        const float x = nab[ray * 3], y = nab[ray * 3 + 1], z = nab[ray * 3 + 2];
        const float nrm = fmaxf(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z))), 1e-12f);
        normals[ray * 3] = m ? __fdiv_rn(x, nrm) : 0.f; normals[ray * 3 + 1] = m ? __fdiv_rn(y, nrm) : 0.f; normals[ray * 3 + 2] = m ? __fdiv_rn(z, nrm) : 0.f;
    }
}

struct SurfaceWs { size_t dirs, D, VAL, st_f, st_b, list, count, pts, nab, scratch, total; };
static SurfaceWs surface_ws_layout(const NaSurfaceCfg& cfg, long long n_rays) {
    SurfaceWs w; size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~(size_t)255; return r; };
    const long long steps = cfg.algo == NA_RAYCAST_ROOT_FINDING ? cfg.n_steps : 1;
    w.dirs = take((size_t)n_rays * 3 * 4);
    w.D = take((size_t)n_rays * steps * 4);
    w.VAL = take((size_t)n_rays * steps * 4);
    w.st_f = take((size_t)n_rays * 6 * 4);
    w.st_b = take((size_t)n_rays * 4);
    w.list = take((size_t)n_rays * 4);
    w.count = take(256);
    w.pts = take((size_t)n_rays * 3 * 4);
    w.nab = take((size_t)n_rays * 3 * 4);
    w.scratch = take(mlp_scratch_bytes());
    w.total = o;
    return w;
}

// dirs: normalised.  Fills depth [n], pts [n,3], mask [n] (and mask_sign_change [n] for root finding, may be NULL).
static int ray_cast(const NaNetDesc* desc, const void* packed, const NaSurfaceCfg& cfg, const float* ro, const float* dirs, long long n_rays,
                    const float* t_steps, float* depth, float* pts, unsigned char* mask, unsigned char* mask_sign,
                    unsigned char* ws, const SurfaceWs& w, cudaStream_t stream) {
    float* D = (float*)(ws + w.D); float* VAL = (float*)(ws + w.VAL);
    float* scratch = (float*)(ws + w.scratch);
    const size_t scratch_bytes = w.total - w.scratch;
    const unsigned gb = (unsigned)((n_rays + 255) / 256);
    EvalJob job = {};
    job.rays_o = ro; job.rays_d = dirs; job.n_rows = (int)n_rays;
    job.apply_bg = 0; job.bound_r = desc->bounding_radius; job.want_full = 0; job.multires_view = desc->multires_view;
    if (cfg.algo == NA_RAYCAST_ROOT_FINDING) {
        if (!t_steps || cfg.n_steps < 2) return NA_ERR_BAD_ARG;
        RootState st;
        float* f = (float*)(ws + w.st_f);
        st.d_low = f; st.f_low = f + n_rays; st.d_high = f + 2 * n_rays; st.f_high = f + 3 * n_rays; st.d_pred = f + 4 * n_rays;
        float* fmid = f + 5 * n_rays;
        unsigned char* b = ws + w.st_b;
        st.mask = mask; st.mask_sign = mask_sign ? mask_sign : b; st.mask0 = b + n_rays;
        st.list = (int*)(ws + w.list); st.count = (int*)(ws + w.count);
        NA_TRY(check_cuda(cudaMemsetAsync(st.count, 0, sizeof(int), stream)));
        const long long tot = n_rays * cfg.n_steps;
        march_depths_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(D, cfg.n_steps, t_steps, cfg.near, cfg.far, n_rays);
        NA_CHECK_LAUNCH();
        EvalJob mj = job;
        mj.P = cfg.n_steps; mj.t = D; mj.t_stride = cfg.n_steps; mj.t_off = 0; mj.o_stride = cfg.n_steps; mj.o_off = 0; mj.sdf = VAL;
        NA_TRY(launch_mlp(mj, packed, cfg.precision, scratch, scratch_bytes, stream));
        root_scan_kernel<<<(unsigned)((n_rays * 32 + 255) / 256), 256, 0, stream>>>(D, VAL, cfg.n_steps, cfg.logit_tau, n_rays, st);
        NA_CHECK_LAUNCH();
        for (int it = 0; it < cfg.n_secant_steps; ++it) {
            EvalJob sj = job;
            sj.row_ids = st.list; sj.n_rows_dev = st.count; sj.P = 1; sj.t = st.d_pred; sj.t_stride = 1; sj.t_off = 0;
            sj.o_stride = 1; sj.o_off = 0; sj.sdf = fmid;
            NA_TRY(launch_mlp(sj, packed, cfg.precision, scratch, scratch_bytes, stream));
            secant_update_kernel<<<num_sms() * 4, 256, 0, stream>>>(fmid, cfg.logit_tau, st);
            NA_CHECK_LAUNCH();
        }
        root_finalize_kernel<<<gb, 256, 0, stream>>>(ro, dirs, st, cfg.far, cfg.fill_inf, 1, n_rays, depth, pts);
        NA_CHECK_LAUNCH();
    } else if (cfg.algo == NA_RAYCAST_SPHERE_TRACING) {
        sphere_init_kernel<<<gb, 256, 0, stream>>>(depth, mask, cfg.near, n_rays);
        NA_CHECK_LAUNCH();
        for (int it = 0; it < cfg.n_iters; ++it) {
            EvalJob sj = job;
            sj.P = 1; sj.t = depth; sj.t_stride = 1; sj.t_off = 0; sj.o_stride = 1; sj.o_off = 0; sj.sdf = VAL;
            NA_TRY(launch_mlp(sj, packed, cfg.precision, scratch, scratch_bytes, stream));
            sphere_step_kernel<<<gb, 256, 0, stream>>>(depth, mask, VAL, cfg.far, n_rays);
            NA_CHECK_LAUNCH();
        }
        sphere_points_kernel<<<(unsigned)((n_rays * 3 + 255) / 256), 256, 0, stream>>>(ro, dirs, depth, n_rays, pts);
        NA_CHECK_LAUNCH();
    } else {
        return NA_ERR_UNSUPPORTED;
    }
    return NA_OK;
}

int preload_surface() {
    NA_PRELOAD(march_depths_kernel);
    NA_PRELOAD(root_scan_kernel);
    NA_PRELOAD(secant_update_kernel);
    NA_PRELOAD(root_finalize_kernel);
    NA_PRELOAD(sphere_init_kernel);
    NA_PRELOAD(sphere_step_kernel);
    NA_PRELOAD(sphere_points_kernel);
    NA_PRELOAD(surface_finish_kernel);
    return NA_OK;
}

}  // namespace na

using namespace na;

static int check_surface_cfg(const NaSurfaceCfg* cfg) {
    if (cfg->precision < NA_PRECISION_FP32 || cfg->precision > NA_PRECISION_TC_MIXED) return NA_ERR_UNSUPPORTED;
    if (cfg->algo != NA_RAYCAST_ROOT_FINDING && cfg->algo != NA_RAYCAST_SPHERE_TRACING) return NA_ERR_UNSUPPORTED;
    if (cfg->n_secant_steps < 0 || cfg->n_iters < 0) return NA_ERR_BAD_ARG;
    return NA_OK;
}

extern "C" size_t na_surface_workspace_bytes(const NaSurfaceCfg* cfg, int64_t n_rays) {
    if (!cfg || n_rays <= 0) return 0;
    return surface_ws_layout(*cfg, n_rays).total;
}

extern "C" int na_ray_cast(const NaNetDesc* desc, const void* packed, const NaSurfaceCfg* cfg, const float* rays_o, const float* rays_d_unit,
                           int64_t n_rays, const float* t_steps, float* depth, float* pts, uint8_t* mask, uint8_t* mask_sign_change,
                           void* workspace, size_t ws_bytes, void* stream) {
    if (!desc || !packed || !cfg || !rays_o || !rays_d_unit || !depth || !pts || !mask || !workspace || n_rays <= 0) return NA_ERR_BAD_ARG;
    NA_TRY(check_surface_cfg(cfg));
    const SurfaceWs w = surface_ws_layout(*cfg, n_rays);
    if (ws_bytes < w.total) return NA_ERR_WORKSPACE;
    return ray_cast(desc, packed, *cfg, rays_o, rays_d_unit, n_rays, t_steps, depth, pts, mask, mask_sign_change, (unsigned char*)workspace, w,
                    (cudaStream_t)stream);
}

extern "C" int na_surface_render_fwd(const NaNetDesc* desc, const void* packed, const NaSurfaceCfg* cfg, const float* rays_o, const float* rays_d,
                                     int64_t n_rays, const float* t_steps, const NaSurfaceOut* out, void* workspace, size_t ws_bytes, void* stream_) {
    if (!desc || !packed || !cfg || !rays_o || !rays_d || !out || !workspace || n_rays <= 0) return NA_ERR_BAD_ARG;
    if (!out->rgb || !out->depth || !out->mask) return NA_ERR_BAD_ARG;
    NA_TRY(check_surface_cfg(cfg));
    const SurfaceWs w = surface_ws_layout(*cfg, n_rays);
    if (ws_bytes < w.total) return NA_ERR_WORKSPACE;
    cudaStream_t stream = (cudaStream_t)stream_;
    unsigned char* ws = (unsigned char*)workspace;
    float* dirs = (float*)(ws + w.dirs); float* pts = (float*)(ws + w.pts);
    float* nab = out->nablas ? out->nablas : (float*)(ws + w.nab);
    NA_TRY(launch_normalize_dirs(rays_d, dirs, n_rays, stream));                                      // ray_casting.py:212
    NA_TRY(ray_cast(desc, packed, *cfg, rays_o, dirs, n_rays, t_steps, out->depth, pts, out->mask, nullptr, ws, w, stream));
    // color, _, nablas = model.forward(pt_pred, view_dirs)   (ray_casting.py:235)
    EvalJob fj = {};
    fj.x = pts; fj.view = cfg->use_view_dirs ? dirs : nullptr; fj.m = n_rays; fj.rad = out->rgb; fj.nab = nab; fj.sdf = (float*)(ws + w.VAL);
    fj.apply_bg = desc->framework == NA_FRAMEWORK_VOLSDF; fj.bound_r = desc->bounding_radius; fj.want_full = 1; fj.multires_view = desc->multires_view;
    if (!cfg->use_view_dirs) return NA_ERR_UNSUPPORTED;
    NA_TRY(launch_mlp(fj, packed, cfg->precision, (float*)(ws + w.scratch), w.total - w.scratch, stream));
    surface_finish_kernel<<<(unsigned)((n_rays + 255) / 256), 256, 0, stream>>>(out->rgb, nab, out->mask, out->normals, n_rays);
    NA_CHECK_LAUNCH();
    return NA_OK;
}
