/*
 * nerfart_b200.h -- C ABI of libnerfart_b200.so (sm_100a only).
 *
 * Drop-in boundary for the volumetric-render hot path of cassiePython/NeRF-Art.  The reference has
 * no FFI / plugin registry (it is 100 % Python): the "interface" this library replaces
 * is the set of Python call sites listed beside each entry point (file:line in the reference tree).
 * The reference-side binding a maintainer adds is a ctypes stub -- see INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns every buffer,
 *     including the workspace; nothing is allocated or freed inside the library
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden synchronisation
 *   - returns 0 on success, a negative NA_ERR_* code otherwise (no C++ exceptions cross the boundary)
 *   - no global mutable state; calls on different streams / devices may run concurrently
 *   - fp32 tensors, row-major, innermost dimension contiguous
 */
#ifndef NERFART_B200_H
#define NERFART_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NA_OK                 0
#define NA_ERR_BAD_ARG       -1
#define NA_ERR_WORKSPACE     -2   /* workspace too small */
#define NA_ERR_CUDA          -3   /* a CUDA runtime call failed; see na_last_cuda_error() */
#define NA_ERR_UNSUPPORTED   -4   /* shape / option outside what the kernels implement */

#define NA_FRAMEWORK_VOLSDF   0
#define NA_FRAMEWORK_NEUS     1

/* MLP arithmetic mode (DESIGN.md "precision modes"):
 *   FP32     : CUDA-core fp32 FFMA everywhere (bit-for-bit reproducible, tightest parity)
 *   TC       : tcgen05 tensor cores, activations resident in TMEM; every GEMM as three fp16 products of two-term
 *              operands (hi*hi + lo*hi + hi*lo: 22-bit operands, fp32 accumulate in TMEM)
 *   TC2ACC   : tcgen05, activations in shared memory, the correction products in a second TMEM accumulator
 *              (3x lower accumulation error than TC, slower)
 *   TC_MIXED : TC for the SDF forward pass (what sample positions depend on); feature head, reverse sweep and
 *              radiance layers with the hi*hi product only (11-bit operands, TF32-level)                          */
#define NA_PRECISION_FP32     0
#define NA_PRECISION_TC       1
#define NA_PRECISION_TC2ACC   2
#define NA_PRECISION_TC_MIXED 3

/* Network geometry.  Mirrors what models/frameworks/volsdf.py:943-975 / neus.py:693-731 build:
 * SDF net  D=8, W=256, skip at layer 4, embed_multires=6 (39-d), W_geo_feat=256 (models/base.py:131-241)
 * radiance D=4, W=256, embed_multires=-1, embed_multires_view=-1 (VolSDF) or 4 (NeuS)  (base.py:312-369).
 * Other geometries return NA_ERR_UNSUPPORTED. */
typedef struct NaNetDesc {
    int32_t framework;          /* NA_FRAMEWORK_* */
    int32_t multires_view;      /* -1 (identity, 3-d) or 4 (27-d) */
    float   bounding_radius;    /* VolSDF sphere background radius (volsdf.py:341-357); unused for NeuS */
    float   reserved;
} NaNetDesc;

/* Raw (un-folded) parameters in the reference checkpoint layout:
 * for each layer: bias [out], weight_g [out,1], weight_v [out,in].  14 layers: 9 SDF then 5 radiance. */
#define NA_NUM_SDF_LAYERS 9
#define NA_NUM_RAD_LAYERS 5
typedef struct NaRawParams {
    const float* bias[NA_NUM_SDF_LAYERS + NA_NUM_RAD_LAYERS];
    const float* weight_g[NA_NUM_SDF_LAYERS + NA_NUM_RAD_LAYERS];
    const float* weight_v[NA_NUM_SDF_LAYERS + NA_NUM_RAD_LAYERS];
} NaRawParams;

typedef struct NaVolsdfCfg {              /* kwargs of volsdf.volume_render, volsdf.py:389-424 */
    int32_t n_samples;                    /* N_samples (128)            */
    int32_t n_importance;                 /* N_importance (64)          */
    int32_t max_upsample_steps;           /* max_upsample_steps (YAML max_upsample_iter, 6) */
    int32_t max_bisection_steps;          /* max_bisection_steps (10)   */
    float   near, far;                    /* 0.0, 6.0                   */
    float   epsilon;                      /* 0.1                        */
    int32_t white_bkgd;                   /* 0/1                        */
    int32_t perturb;                      /* 0: det linspace u; 1: u_final supplied by the caller */
    int32_t precision;                    /* NA_PRECISION_*             */
    int32_t detailed;                     /* 1: fill the per-sample NaVolsdfDetail arrays */
    int32_t reserved;
} NaVolsdfCfg;

typedef struct NaVolsdfOut {              /* return values of volume_render, volsdf.py:566-594 */
    float* rgb;                           /* [n_rays,3]  */
    float* depth;                         /* [n_rays]    depth_volume */
    float* acc;                           /* [n_rays]    mask_volume  */
    float* normals;                       /* [n_rays,3]  normals_volume (may be NULL) */
    float* beta_map;                      /* [n_rays]    */
    float* iter_usage;                    /* [n_rays]    float like the reference: 0..max, -1 = not converged */
    /* detailed_output (all may be NULL unless cfg.detailed); P = n_samples + n_importance */
    float* d_vals;                        /* [n_rays,P]   */
    float* sdf;                           /* [n_rays,P]   implicit_surface */
    float* nablas;                        /* [n_rays,P,3] implicit_nablas  */
    float* radiance;                      /* [n_rays,P,3] */
    float* sigma;                         /* [n_rays,P]   */
    float* tau;                           /* [n_rays,P-1] visibility_weights */
} NaVolsdfOut;

typedef struct NaNeusCfg {                /* kwargs of neus.volume_render, neus.py:142-185 ('official_solution') */
    int32_t n_samples;                    /* 64 */
    int32_t n_importance;                 /* 64 */
    int32_t n_upsample_iters;             /* 4  */
    float   bounding_radius;              /* obj_bounding_radius (1.0) */
    int32_t white_bkgd;
    int32_t perturb;                      /* 1: u supplied by caller [n_upsample_iters][n_rays][n_importance/iters] */
    int32_t precision;
    int32_t detailed;
} NaNeusCfg;

typedef struct NaNeusOut {                /* neus.py:385-407 */
    float* rgb; float* depth; float* acc; float* normals;
    float* d_all;                         /* [n_rays,P]   */
    float* sdf;                           /* [n_rays,P]   */
    float* nablas;                        /* [n_rays,P,3] */
    float* radiance;                      /* [n_rays,P-1,3] */
    float* alpha;                         /* [n_rays,P-1] */
    float* weights;                       /* [n_rays,P-1] visibility_weights */
} NaNeusOut;

/* Surface rendering (models/ray_casting.py).  algo: NA_RAYCAST_ROOT_FINDING = root_finding_surface_points (35-160, secant refinement
 * 11-30), NA_RAYCAST_SPHERE_TRACING = sphere_tracing_surface_points (163-184).  near / far are scalars (the reference also accepts
 * per-ray tensors; render.py passes floats). */
#define NA_RAYCAST_ROOT_FINDING   0
#define NA_RAYCAST_SPHERE_TRACING 1
typedef struct NaSurfaceCfg {
    int32_t algo;
    int32_t n_steps;                      /* N_steps (256): marching samples per ray (root finding)   */
    int32_t n_secant_steps;               /* N_secant_steps (8)                                       */
    int32_t n_iters;                      /* N_iters (20): sphere-tracing steps                       */
    float   near, far;                    /* 0.0, 6.0                                                 */
    float   logit_tau;                    /* 0.0                                                      */
    int32_t fill_inf;                     /* 1: depth = inf where nothing is hit, 0: far              */
    int32_t use_view_dirs;                /* surface_render(use_view_dirs=True)                       */
    int32_t precision;                    /* NA_PRECISION_*                                           */
} NaSurfaceCfg;

typedef struct NaSurfaceOut {             /* return values of ray_casting.surface_render, 241-263 */
    float*   rgb;                         /* [n_rays,3]  colours, 0 where mask is false */
    float*   depth;                       /* [n_rays]    */
    uint8_t* mask;                        /* [n_rays]    mask_surface */
    float*   nablas;                      /* [n_rays,3]  implicit_nablas (may be NULL)  */
    float*   normals;                     /* [n_rays,3]  normals_surface (may be NULL)  */
} NaSurfaceOut;

/* ---- capability / errors ------------------------------------------------------------------- */
int         na_version(void);
const char* na_error_string(int code);
int         na_last_cuda_error(void);                 /* cudaError_t of the last NA_ERR_CUDA on this thread */
int64_t     na_kernel_launch_count(void);             /* kernels launched by this library since load (bench "gpu_launches") */
int         na_debug_set_buffer(void* dev_int64x8);   /* diagnostics: cycle counters of CTA 0 of the tensor-core MLP kernel */
/* Load every kernel image of the library on the current device now instead of at each kernel's first launch (the driver's
 * default is lazy loading).  The Python engine calls it once per device; C callers should do the same before the first step. */
int         na_preload_kernels(void);
/* Stall diagnostics.  Every mbarrier wait of the tcgen05 kernels is bounded (8 s): a wait that can never complete ends the
 * launch with a trap, i.e. a CUDA error, instead of blocking the stream forever.  na_diag_enable(1) additionally records, in
 * host-mapped memory, per-launch CTA counters and every timed-out wait; na_diag_enable(2) also keeps an event per kernel launch
 * (names the first launch that never finished); na_diag_enable(0) switches both off.  na_diag_dump
 * writes a text report (<= cap bytes; returns the length) and may be called from a watchdog thread while a stream is stuck. */
int         na_diag_enable(int on);
int         na_diag_dump(char* buf_host, int cap);

/* ---- weights: replaces nn.utils.weight_norm's per-forward W = g*v/||v|| (models/base.py:226-227,365-366) */
size_t na_packed_weights_bytes(const NaNetDesc* desc);
int    na_pack_weights(const NaNetDesc* desc, const NaRawParams* raw, void* packed, void* stream);

/* ---- per-sample networks (stage-wise parity + ImplicitSurface / RadianceNet / mesh / ray-casting callers) */
/* ImplicitSurface.forward (models/base.py:243-263): x [m,3] -> sdf [m], feat [m,256] (feat may be NULL).
 * apply_bg!=0: VolSDF.forward_surface, min(sdf, R-||x||) (volsdf.py:341-347). */
int na_sdf_eval(const NaNetDesc* desc, const void* packed, const float* x, int64_t m, int apply_bg,
                int precision, float* sdf, float* feat, void* workspace, size_t ws_bytes, void* stream);
/* VolSDF.forward / NeuS.forward (volsdf.py:359-370, neus.py:120-123): x,view [m,3] -> radiance [m,3], sdf [m], nablas [m,3]
 * (ImplicitSurface.forward_with_nablas base.py:265-282 + RadianceNet.forward base.py:372-391). */
int na_full_eval(const NaNetDesc* desc, const void* packed, const float* x, const float* view, int64_t m,
                 int precision, float* radiance, float* sdf, float* nablas, float* feat,
                 void* workspace, size_t ws_bytes, void* stream);
size_t na_eval_workspace_bytes(int64_t m);

/* ---- ray generation: rend_util.get_rays (utils/rend_util.py:112-165) with N_rays=-1 ------------ */
int na_get_rays(const float* c2w /*[4,4]*/, const float* intrinsics /*[4,4]*/, int H, int W,
                float* rays_o /*[H*W,3]*/, float* rays_d /*[H*W,3]*/, void* stream);

/* ---- sampler stages (parity of rend_util.sample_pdf 256-293, sample_cdf 295-328, volsdf.error_bound 56-94) */
int na_error_bound(const float* d_vals, const float* sdf, int64_t rows, int n, const float* alpha_beta_rows /*[rows,2] or NULL*/,
                   float alpha, float beta, float* bounds /*[rows,n-1]*/, void* stream);
int na_sample_pdf(const float* bins, const float* weights, int64_t rows, int n, const float* u /*[n_out] shared or [rows,n_out]*/,
                  int u_per_row, int n_out, float* samples, int64_t* inds /*may be NULL*/, void* stream);
int na_sample_cdf(const float* bins, const float* cdf, int64_t rows, int n, const float* u, int u_per_row, int n_out,
                  float* samples, int64_t* inds, void* stream);

/* ---- VolSDF renderer: volsdf.volume_render (volsdf.py:389-615) -------------------------------- */
size_t na_volsdf_workspace_bytes(const NaVolsdfCfg* cfg, int64_t n_rays);
/* rays_d un-normalised (normalised inside, volsdf.py:442).  alpha_beta: device [2] = VolSDF.forward_ab().
 * t_coarse [n_samples], t_init [4*n_samples], u_up [4*n_samples+2], u_imp [n_importance]: the torch.linspace(0,1,.)
 * tables the reference builds at volsdf.py:472,483 and rend_util.py:269,304 (passed in so both sides use identical
 * values); u_final [n_rays,n_importance] only when cfg.perturb. */
int na_volsdf_render_fwd(const NaNetDesc* desc, const void* packed, const NaVolsdfCfg* cfg,
                         const float* rays_o, const float* rays_d, int64_t n_rays, const float* alpha_beta,
                         const float* t_coarse, const float* t_init, const float* u_up, const float* u_imp,
                         const float* u_final, const NaVolsdfOut* out, void* workspace, size_t ws_bytes, void* stream);

/* ---- NeuS renderer: neus.volume_render (neus.py:142-424) -------------------------------------- */
size_t na_neus_workspace_bytes(const NaNeusCfg* cfg, int64_t n_rays);
int na_neus_render_fwd(const NaNetDesc* desc, const void* packed, const NaNeusCfg* cfg,
                       const float* rays_o, const float* rays_d, int64_t n_rays, const float* s_dev /*[1] = NeuS.forward_s()*/,
                       const float* t_coarse, const float* u_imp, const float* u_rand,
                       const NaNeusOut* out, void* workspace, size_t ws_bytes, void* stream);

/* ---- surface rendering: models/ray_casting.py ------------------------------------------------------------- */
/* root_finding_surface_points (35-160) / sphere_tracing_surface_points (163-184) on ImplicitSurface.forward: rays_d_unit normalised,
 * t_steps = torch.linspace(0,1,n_steps) (root finding; NULL for sphere tracing) -> depth [n], pts [n,3], mask [n], mask_sign_change [n] (NULL ok) */
size_t na_surface_workspace_bytes(const NaSurfaceCfg* cfg, int64_t n_rays);
int na_ray_cast(const NaNetDesc* desc, const void* packed, const NaSurfaceCfg* cfg, const float* rays_o, const float* rays_d_unit,
                int64_t n_rays, const float* t_steps, float* depth, float* pts, uint8_t* mask, uint8_t* mask_sign_change,
                void* workspace, size_t ws_bytes, void* stream);
/* surface_render (187-263): rays_d un-normalised; normalise -> ray cast -> model.forward at the hit points -> masked colours / normals */
int na_surface_render_fwd(const NaNetDesc* desc, const void* packed, const NaSurfaceCfg* cfg, const float* rays_o, const float* rays_d,
                          int64_t n_rays, const float* t_steps, const NaSurfaceOut* out, void* workspace, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * Backward of the render: the second pass of the CLIP fine-tune step.
 * Replaces what autograd does for   rgb_pred.backward(gradient[:, i:i+batch_size, :], retain_graph=True);
 *                                   eikonal_loss.backward()
 * in Trainer.forward (models/frameworks/volsdf.py:769-783; models/frameworks/neus.py:551-563), with
 * calc_eikonal_loss = w_eikonal * mse(||extras['implicit_nablas']||, 1) over all points of the patch
 * (volsdf.py:917-939; neus.py:667-690).  The sample depths are constants of the backward (the samplers run under
 * torch.no_grad: volsdf.py:113, neus.py:275-303), so the inputs are the detailed outputs of the forward render of the
 * same ray patch (na_volsdf_render_fwd / na_neus_render_fwd with cfg.detailed = 1).
 * Gradients accumulate (+=) into `grad_pack`, a buffer of na_grad_pack_bytes() that the caller zeroes at
 * optimizer.zero_grad() time (volsdf.py:753); na_unpack_grads maps it to d loss / d (bias, weight_g, weight_v) of the
 * reference's parameters through the weight-norm Jacobian (models/base.py:226-227,365-366) once per step.          */
typedef struct NaTrainCfg {
    int32_t points_per_ray;       /* P = N_samples + N_importance of the forward render                                 */
    float   w_eikonal;            /* args.finetune.w_eikonal; 0 disables the eikonal term (finetune.use_eikonal False)  */
    int32_t eikonal_count;        /* number of points the eikonal mean runs over = rays in the reference's patch * P    */
    int32_t white_bkgd;
    float   speed_factor;         /* ln_beta / ln_s speed factor (volsdf.py:337-339, neus.py:116-117)                   */
    int32_t train_surface;        /* 0: implicit_surface is frozen (fix_module): no SDF-net weight gradients            */
    int32_t train_radiance;       /* 0: radiance_net is frozen (NeuS fine-tune, neus.py:28)                             */
    int32_t precision;            /* NA_PRECISION_*: FP32 = the backward kernel recomputes the forward pass of the patch in fp32
                                     on CUDA cores; tensor-core modes = the tcgen05 forward kernel re-evaluates the patch once and
                                     leaves the activations the backward needs in the workspace (no recompute)              */
} NaTrainCfg;

typedef struct NaRawGrads {       /* where na_unpack_grads writes; same layer order as NaRawParams; NULL = skip layer   */
    float* bias[NA_NUM_SDF_LAYERS + NA_NUM_RAD_LAYERS];
    float* weight_g[NA_NUM_SDF_LAYERS + NA_NUM_RAD_LAYERS];
    float* weight_v[NA_NUM_SDF_LAYERS + NA_NUM_RAD_LAYERS];
} NaRawGrads;

size_t na_grad_pack_bytes(const NaNetDesc* desc);
/* workspace of one render_bwd call: for any arithmetic mode (the maximum), or for the mode the call will use (tensor-core modes hold
 * the patch's activation stash as 16-bit planes: 21.6 KB per sample; fp32 mode: 43 KB per sample) */
size_t na_train_workspace_bytes(const NaNetDesc* desc, int64_t n_rays, int32_t points_per_ray);
size_t na_train_workspace_bytes_mode(const NaNetDesc* desc, int64_t n_rays, int32_t points_per_ray, int32_t precision);
/* diagnostics / unit test of the weight-gradient GEMM kernel (csrc/wgrad_f16.cu): planes16 = two 16-bit planes [2][m_pad][256]
 * (plane 0 = L, plane 1 = R; m_pad a multiple of 128), out[256][256] += L^T R over m_rows samples, bias_out[256] += column sums of L */
int na_debug_wgrad_f16(const void* planes16, int64_t m_pad, int64_t m_rows, int l_bf16, int r_bf16, float* out, float* bias_out, void* stream);

/* rays [n,3] (directions un-normalised, as passed to the forward), alpha_beta = {1/beta, beta} (forward_ab),
 * d_all / sdf [n,P], radiance / nablas [n,P,3] = extras d_vals, implicit_surface, radiance, implicit_nablas of the forward
 * (volsdf.py:578-594), grad_rgb [n,3].  scalars (float64[2], accumulated): [0] d loss / d ln_beta, [1] eikonal loss.  */
int na_volsdf_render_bwd(const NaNetDesc* desc, const void* packed, const NaTrainCfg* cfg, const float* rays_o, const float* rays_d,
                         int64_t n_rays, const float* alpha_beta, const float* d_all, const float* sdf, const float* radiance,
                         const float* nablas, const float* grad_rgb, void* grad_pack, double* scalars, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Split form of the VolSDF fine-tune patch (tensor-core modes; same results as na_volsdf_render_fwd + na_volsdf_render_bwd up to
 * rounding order).  The reference renders the patch with grad and then back-propagates through the same graph
 * (volsdf.py:760-783): one network evaluation per sample.  na_volsdf_render_fwd + na_volsdf_render_bwd evaluate the network twice
 * (the backward launch re-evaluates the forward pass it differentiates).  Here the final full evaluation of the forward render is
 * itself the forward half of the training program: it leaves the activations the backward needs in `train_workspace`
 * (na_train_workspace_bytes_mode bytes for the same n_rays / points_per_ray / precision); na_volsdf_render_bwd_stashed, given the
 * SAME workspace and the detailed outputs of that render, runs the backward half only.  `out` must carry d_vals, sdf, radiance and
 * nablas.  NA_ERR_UNSUPPORTED for NA_PRECISION_FP32 / NA_PRECISION_TC2ACC.                                               */
int na_volsdf_render_fwd_train(const NaNetDesc* desc, const void* packed, const NaVolsdfCfg* cfg, const float* rays_o,
                               const float* rays_d, int64_t n_rays, const float* alpha_beta, const float* t_coarse,
                               const float* t_init, const float* u_up, const float* u_imp, const float* u_final,
                               const NaVolsdfOut* out, void* workspace, size_t workspace_bytes, void* train_workspace,
                               size_t train_workspace_bytes, void* stream);
int na_volsdf_render_bwd_stashed(const NaNetDesc* desc, const void* packed, const NaTrainCfg* cfg, const float* rays_o,
                                 const float* rays_d, int64_t n_rays, const float* alpha_beta, const float* d_all, const float* sdf,
                                 const float* radiance, const float* nablas, const float* grad_rgb, void* grad_pack, double* scalars,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* the same split for NeuS (neus.py:520-576): two evaluations per ray sample exist in the reference too (P points for sdf / nabla,
 * P - 1 midpoints for the radiance, neus.py:320-324); each is evaluated once.  The workspace holds both stashes
 * (na_train_workspace_bytes_mode with a NeuS desc accounts for it).                                                       */
int na_neus_render_fwd_train(const NaNetDesc* desc, const void* packed, const NaNeusCfg* cfg, const float* rays_o,
                             const float* rays_d, int64_t n_rays, const float* s, const float* t_coarse, const float* u_imp,
                             const float* u_rand, const NaNeusOut* out, void* workspace, size_t workspace_bytes,
                             void* train_workspace, size_t train_workspace_bytes, void* stream);
int na_neus_render_bwd_stashed(const NaNetDesc* desc, const void* packed, const NaTrainCfg* cfg, const float* rays_o,
                               const float* rays_d, int64_t n_rays, const float* s, const float* d_all, const float* sdf,
                               const float* radiance, const float* nablas, const float* grad_rgb, void* grad_pack, double* scalars,
                               void* workspace, size_t workspace_bytes, void* stream);

/* NeuS: s = {forward_s()}; d_all / sdf [n,P], nablas [n,P,3] at the sample points, radiance [n,P-1,3] at the midpoints
 * (extras of neus.py:397-407).  scalars: [0] d loss / d ln_s, [1] eikonal loss.                                      */
int na_neus_render_bwd(const NaNetDesc* desc, const void* packed, const NaTrainCfg* cfg, const float* rays_o, const float* rays_d,
                       int64_t n_rays, const float* s, const float* d_all, const float* sdf, const float* radiance,
                       const float* nablas, const float* grad_rgb, void* grad_pack, double* scalars, void* workspace,
                       size_t workspace_bytes, void* stream);

int na_unpack_grads(const NaNetDesc* desc, const NaRawParams* raw, const void* grad_pack, const NaRawGrads* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * CLIP ViT-B/32 image tower: replaces `self.model.encode_image(images)` (criteria/clip_loss.py:206-208,
 * criteria/contrastive_loss.py:113-115, criteria/patchnce_loss.py:127-129; third-party openai/CLIP `model.visual`) and
 * its autograd backward w.r.t. the image (the CLIP weights are frozen; NeRF-Art needs d loss / d rendered pixels only).
 * Weights: fp32 device pointers in the openai/CLIP `visual.*` state-dict layout.                                     */
typedef struct NaClipLayer {
    const float *ln_1_w, *ln_1_b;            /* [768]                                                          */
    const float *in_proj_w, *in_proj_b;      /* attn.in_proj_weight [2304,768] (q|k|v), in_proj_bias [2304]    */
    const float *out_proj_w, *out_proj_b;    /* attn.out_proj [768,768], [768]                                 */
    const float *ln_2_w, *ln_2_b;
    const float *c_fc_w, *c_fc_b;            /* mlp.c_fc [3072,768], [3072]  (QuickGELU follows)               */
    const float *c_proj_w, *c_proj_b;        /* mlp.c_proj [768,3072], [768]                                   */
} NaClipLayer;
typedef struct NaClipWeights {
    const float* conv1;                      /* visual.conv1.weight [768,3,32,32] (no bias)                    */
    const float* class_embedding;            /* [768]                                                          */
    const float* positional_embedding;       /* [50,768]                                                       */
    const float *ln_pre_w, *ln_pre_b;
    NaClipLayer layers[12];                  /* visual.transformer.resblocks.{i}                               */
    const float *ln_post_w, *ln_post_b;
    const float* proj;                       /* visual.proj [768,512]                                          */
    const void* packed;                      /* na_clip_pack_weights output (required for NA_CLIP_TF32), else NULL */
    int32_t precision;                       /* NA_CLIP_FP32 | NA_CLIP_TF32                                    */
    int32_t reserved;
} NaClipWeights;
/* arithmetic of the tower's linear layers (conv1-as-GEMM, in_proj, out_proj, c_fc, c_proj, proj; forward and backward-data):
 * NA_CLIP_FP32 = fp32 CUDA cores; NA_CLIP_TF32 = tcgen05 kind::tf32 with fp32 accumulation in TMEM (csrc/tgemm.cu).  The
 * reference runs this model in fp16 on the GPU (clip.load keeps fp16 weights), so both are at least its precision.
 * LayerNorm, softmax attention over the 50 tokens and QuickGELU stay fp32 in both modes.                                  */
#define NA_CLIP_FP32 0
#define NA_CLIP_TF32 1

/* TF32 weight images for NA_CLIP_TF32: every linear weight in both orientations (y = x W^T and dx = dy W), rounded to TF32,
 * pre-tiled as the shared-memory image of the tensor-core operand.  The weights are frozen, so this runs once per tower.  */
size_t na_clip_packed_bytes(void);
int na_clip_pack_weights(const NaClipWeights* weights, void* packed, void* stream);

size_t na_clip_workspace_bytes(int32_t batch);
/* images [B,3,224,224] (already resized + CLIP-normalised) -> feats [B,512].  The workspace keeps the activations of this
 * forward; na_clip_vitb32_encode_bwd must be called with the same workspace and batch before the next forward.           */
int na_clip_vitb32_encode_fwd(const NaClipWeights* weights, const float* images, int32_t batch, float* feats, void* workspace,
                              size_t workspace_bytes, void* stream);
/* grad_feats [B,512] -> grad_images [B,3,224,224] */
int na_clip_vitb32_encode_bwd(const NaClipWeights* weights, const float* grad_feats, int32_t batch, float* grad_images,
                              void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif
