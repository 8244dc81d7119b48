// C-ABI entry points that are not a whole renderer: capability queries, weight packing, per-sample network
// evaluation, ray generation and the stage-wise sampler primitives.  See include/nerfart_b200.h.
#include "sampler.cuh"
#include <atomic>
#include <mutex>
#include <cstdio>
#include <cstring>

namespace na {

thread_local int g_last_cuda_error = 0;
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- stall diagnostics (common.cuh) -------------------------------------------------------------
int g_diag_on = 0;
static HangDiag* g_diag_host = nullptr;          // pinned, mapped
static HangDiag* g_diag_dev = nullptr;
static std::atomic<unsigned long long> g_diag_seq{0};
constexpr int DIAG_RING = 8192;
struct DiagMark { cudaEvent_t ev; const char* file; int line; unsigned long long n; };
static DiagMark g_marks[DIAG_RING];
static std::atomic<unsigned long long> g_mark_n{0};
static std::mutex g_diag_mu;

void diag_mark(const char* file, int line, cudaStream_t stream) {
    std::lock_guard<std::mutex> lk(g_diag_mu);
    const unsigned long long n = g_mark_n.load();
    DiagMark& m = g_marks[n % DIAG_RING];
    if (!m.ev && cudaEventCreateWithFlags(&m.ev, cudaEventDisableTiming) != cudaSuccess) { m.ev = nullptr; cudaGetLastError(); return; }
    if (cudaEventRecord(m.ev, stream) != cudaSuccess) { cudaGetLastError(); return; }
    m.file = file; m.line = line; m.n = n;
    g_mark_n.store(n + 1);
}

SpinCtx diag_next(int kernel, int grid) {
    SpinCtx sc; sc.seq = g_diag_seq.fetch_add(1); sc.kernel = kernel; sc.diag = nullptr;
    if (g_diag_on && g_diag_host) {
        HangSlot& s = g_diag_host->slot[sc.seq % HANG_SLOTS];
        s.seq = sc.seq; s.kernel = kernel; s.grid = grid; s.started = s.ready = s.finished = 0;
        sc.diag = g_diag_dev;
    }
    return sc;
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int launch_mlp_simt(const EvalJob& job, const float* packed, const PackF32& L, float* scratch, size_t scratch_bytes, cudaStream_t stream);
size_t mlp_simt_scratch_bytes(int grid);
struct TcPackLayout { size_t wtc_off, unscale_off, meta_off, total; unsigned stage0[21]; int n_kb[21], n_nh[21]; unsigned n_stages; };
TcPackLayout tc_pack_layout(size_t f32_bytes);
int tc_pack(const float* pk_f32, const PackF32& L, unsigned char* base, const TcPackLayout& T, cudaStream_t stream);
int launch_mlp_tc(const EvalJob& job, const unsigned char* packed_base, size_t f32_bytes, const PackF32& L, float* scratch,
                  size_t scratch_bytes, cudaStream_t stream);
size_t mlp_tc_scratch_bytes(int grid);
extern long long* g_tc_dbg;
size_t mlp_tmem_image_bytes();
int tmem_pack(const float* pk_f32, const size_t* d_offs, const int* d_rows, const float* d_absmax, size_t train_off, const PackTrain& TP,
              unsigned char* image, cudaStream_t stream);
int launch_mlp_tmem(const EvalJob& job, const float* pk_f32, const unsigned char* image, const PackF32& L, int mixed,
                    unsigned char* scratch, size_t scratch_bytes, cudaStream_t stream);
size_t mlp_tmem_scratch_bytes(int grid);

size_t packed_f32_bytes(int multires_view) {
    const PackF32 L = pack_layout_f32(multires_view);
    return ((L.total + 63) / 64 * 64 + 14 * 260 + 64) * sizeof(float);
}
// the TMEM-resident kernel's weight image follows the two-accumulator kernel's image in the packed buffer
size_t tmem_image_off(int multires_view) { return (tc_pack_layout(packed_f32_bytes(multires_view)).total + 1023) & ~(size_t)1023; }
size_t train_pack_off(int multires_view) { return (tmem_image_off(multires_view) + mlp_tmem_image_bytes() + 1023) & ~(size_t)1023; }
// one entry point for all arithmetic modes
int launch_mlp(const EvalJob& job, const void* packed, int precision, float* scratch, size_t scratch_bytes, cudaStream_t stream) {
    const PackF32 L = pack_layout_f32(job.multires_view);
    if (precision == NA_PRECISION_FP32) return launch_mlp_simt(job, (const float*)packed, L, scratch, scratch_bytes, stream);
    if (precision == NA_PRECISION_TC2ACC) return launch_mlp_tc(job, (const unsigned char*)packed, packed_f32_bytes(job.multires_view), L, scratch, scratch_bytes, stream);
    if (precision == NA_PRECISION_TC || precision == NA_PRECISION_TC_MIXED)
        return launch_mlp_tmem(job, (const float*)packed, (const unsigned char*)packed + tmem_image_off(job.multires_view), L,
                               precision == NA_PRECISION_TC_MIXED, (unsigned char*)scratch, scratch_bytes, stream);
    return NA_ERR_UNSUPPORTED;
}
size_t mlp_scratch_bytes() {
    const size_t a = mlp_simt_scratch_bytes(num_sms()), b = mlp_tc_scratch_bytes(num_sms()), c = mlp_tmem_scratch_bytes(num_sms());
    return a > b ? (a > c ? a : c) : (b > c ? b : c);
}

// -----------------------------------------------------------------------------------------------
// weight packing: fold nn.utils.weight_norm (W = g * v / ||v||_row, models/base.py:226-227,365-366), the 1/sqrt(2) of
// the skip connection (base.py:250) and the concat orders into GEMM-ready, zero-padded fp32 planes.
// -----------------------------------------------------------------------------------------------
struct FillDesc {
    unsigned long long dst;     // float offset in the packed buffer
    int R, C;                   // destination plane [R][C]
    int layer;                  // 0..13 source layer
    int mode;                   // 0: dst[r][c] = Weff[c+out0][inmap(r)]   1: dst[r][c] = Weff[r+out0][c]   2: dst[c] = bias[c+out0]
    int out0, out_n;            // valid source rows [out0, out0+out_n)
    int in_dim;                 // source columns
    int map;                    // 0 identity; 1 radiance layer 0 (feat rows first, then the small inputs); 2 identity below `small`, else 0
    int small;                  // small_dim for map 1
    float mult;
    int tf32;                   // 1: round the value to TF32 (cvt.rna) -- operand planes of the mma.sync backward GEMMs
};
constexpr int MAX_FILL = 48;
struct FillTable { FillDesc d[MAX_FILL]; int n; };

__global__ void wn_scale_kernel(const NaRawParams raw, const int* __restrict__ dims /*[14][2] out,in*/, float* __restrict__ scale /*[14][260]*/) {
    const int layer = blockIdx.y;
    const int out = dims[layer * 2], in = dims[layer * 2 + 1];
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= out) return;
    const float* v = raw.weight_v[layer] + (size_t)row * in;
    float s = 0.f;
    for (int i = lane; i < in; i += 32) s += v[i] * v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) scale[layer * 260 + row] = __fdiv_rn(raw.weight_g[layer][row], sqrtf(s));
}

__global__ void pack_fill_kernel(const NaRawParams raw, const FillTable tab, const float* __restrict__ scale, float* __restrict__ packed) {
    const FillDesc f = tab.d[blockIdx.y];
    const int total = f.R * f.C;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int r = idx / f.C, c = idx - r * f.C;
        float val = 0.f;
        if (f.mode == 2) {
            if (c < f.out_n) val = raw.bias[f.layer][c + f.out0];
        } else {
            int o, i;
            if (f.mode == 0) { o = c; i = r; } else { o = r; i = c; }
            int src_in = -1;
            if (f.map == 0) { if (i < f.in_dim) src_in = i; }
            else if (f.map == 2) { if (i < f.small) src_in = i; }
            else { if (i < 256) src_in = f.small + i; else if (i - 256 < f.small) src_in = i - 256; }
            if (o < f.out_n && src_in >= 0) {
                const int so = o + f.out0;
                val = raw.weight_v[f.layer][(size_t)so * f.in_dim + src_in] * scale[f.layer * 260 + so] * f.mult;
                if (f.tf32) { unsigned r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(val)); val = __uint_as_float(r); }
            }
        }
        packed[f.dst + idx] = val;
    }
}

}  // namespace na

using namespace na;

extern "C" int na_version(void) { return 100; }
// diagnostics only: device buffer of >= 8 int64 that mlp_tc_kernel fills with cycle counters of CTA 0 (NULL disables)
namespace na {
struct SmallBlob { unsigned w[256]; };
__global__ void upload_small_kernel(const SmallBlob blob, unsigned* __restrict__ dst, int n_words) {
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) dst[i] = blob.w[i];
}
int upload_small(void* dst, const void* src, size_t n, cudaStream_t stream) {
    if (!dst || !src || n == 0 || n > sizeof(SmallBlob) || (n & 3)) return NA_ERR_BAD_ARG;
    SmallBlob b;
    memcpy(b.w, src, n);
    upload_small_kernel<<<1, 128, 0, stream>>>(b, (unsigned*)dst, (int)(n / 4));
    NA_CHECK_LAUNCH();
    return NA_OK;
}
}  // namespace na

extern "C" int na_debug_set_buffer(void* dev_ptr) { g_tc_dbg = (long long*)dev_ptr; return NA_OK; }
extern "C" const char* na_error_string(int code) {
    switch (code) {
        case NA_OK: return "ok";
        case NA_ERR_BAD_ARG: return "bad argument";
        case NA_ERR_WORKSPACE: return "workspace too small";
        case NA_ERR_CUDA: return "CUDA runtime error (see na_last_cuda_error)";
        case NA_ERR_UNSUPPORTED: return "unsupported shape or option";
        default: return "unknown error";
    }
}
extern "C" int na_last_cuda_error(void) { return g_last_cuda_error; }


// Stall diagnostics (see common.cuh).  na_diag_enable(1): host-mapped record buffer + per-launch CTA counters; (2): also an event per launch.
extern "C" int na_diag_enable(int on) {
    if (on && !g_diag_host) {
        void* h = nullptr; void* d = nullptr;
        NA_TRY(check_cuda(cudaHostAlloc(&h, sizeof(HangDiag), cudaHostAllocMapped)));
        memset(h, 0, sizeof(HangDiag));
        NA_TRY(check_cuda(cudaHostGetDevicePointer(&d, h, 0)));
        g_diag_host = (HangDiag*)h; g_diag_dev = (HangDiag*)d;
    }
    g_diag_on = on < 0 ? 0 : (on > 2 ? 2 : on);
    return NA_OK;
}
// Text report: timed-out waits, tcgen05 launches whose CTAs have not all finished, and the oldest kernel launch (file:line of its
// NA_CHECK_LAUNCH) whose completion event is still pending.  Safe to call from another host thread while the stream is stuck,
// and after a trap (reads host memory and queries events only).  Returns the number of bytes written.
extern "C" int na_diag_dump(char* buf, int cap) {
    if (!buf || cap <= 0) return 0;
    int o = 0;
    auto put = [&](const char* fmt, auto... a) { if (o < cap - 1) { int w = snprintf(buf + o, (size_t)(cap - o), fmt, a...); if (w > 0) o += w < cap - o ? w : cap - o - 1; } };
    if (!g_diag_host) { put("%s", "diagnostics not enabled\n"); return o; }
    const HangDiag& D = *g_diag_host;
    const unsigned nrec = D.n_rec < (unsigned)HANG_RECS ? D.n_rec : (unsigned)HANG_RECS;
    put("timed-out waits: %u\n", D.n_rec);
    for (unsigned i = 0; i < nrec; ++i) {
        const HangRec& r = D.rec[i];
        put("  seq %llu kernel %d block %d warp %d lane %d tag 0x%08x parity %u waited %.2f s\n", r.seq, r.kernel, r.block, r.warp, r.lane, r.tag, r.parity, r.waited_ns * 1e-9);
    }
    put("tcgen05 launches issued: %llu; unfinished:\n", g_diag_seq.load());
    for (int i = 0; i < HANG_SLOTS; ++i) {
        const HangSlot& s = D.slot[i];
        if (s.kernel && s.finished != (unsigned)s.grid && s.seq + HANG_SLOTS > g_diag_seq.load())
            put("  seq %llu kernel %d grid %d started %u ready %u finished %u\n", s.seq, s.kernel, s.grid, s.started, s.ready, s.finished);
    }
    const unsigned long long n = g_mark_n.load();
    const unsigned long long lo = n > DIAG_RING ? n - DIAG_RING : 0;
    put("kernel launches traced: %llu\n", n);
    int shown = 0;
    for (unsigned long long k = lo; k < n && shown < 6; ++k) {
        const DiagMark& m = g_marks[k % DIAG_RING];
        if (!m.ev || m.n != k) continue;
        const cudaError_t q = cudaEventQuery(m.ev);
        if (q == cudaSuccess) continue;
        const char* f = m.file; for (const char* c = m.file; *c; ++c) if (*c == '/') f = c + 1;
        put("  pending launch #%llu at %s:%d (%s)\n", k, f, m.line, q == cudaErrorNotReady ? "not ready" : cudaGetErrorName(q));
        ++shown;
    }
    cudaGetLastError();
    return o;
}
extern "C" int64_t na_kernel_launch_count(void) { return (int64_t)g_launches.load(); }

static size_t pack_scale_off(const PackF32& L) { return (L.total + 63) / 64 * 64; }
static size_t pack_dims_off(const PackF32& L) { return pack_scale_off(L) + 14 * 260; }

extern "C" size_t na_packed_weights_bytes(const NaNetDesc* desc) {
    if (!desc) return 0;
    return train_pack_off(desc->multires_view) + pack_layout_train().total * sizeof(float);
}

extern "C" int na_pack_weights(const NaNetDesc* desc, const NaRawParams* raw, void* packed_, void* stream_) {
    if (!desc || !raw || !packed_) return NA_ERR_BAD_ARG;
    if (desc->multires_view != -1 && desc->multires_view != 4) return NA_ERR_UNSUPPORTED;
    for (int l = 0; l < 14; ++l) if (!raw->bias[l] || !raw->weight_g[l] || !raw->weight_v[l]) return NA_ERR_BAD_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    float* packed = (float*)packed_;
    const PackF32 L = pack_layout_f32(desc->multires_view);
    const int sdim = small_dim(desc->multires_view), spad = small_pad(desc->multires_view);
    int dims[14][2];
    for (int l = 0; l < 8; ++l) { dims[l][0] = (l == 3) ? SKIP_H : W; dims[l][1] = (l == 0) ? EMB : W; }
    dims[8][0] = W + 1; dims[8][1] = W;
    dims[9][0] = W; dims[9][1] = W + sdim;
    for (int l = 10; l < 13; ++l) { dims[l][0] = W; dims[l][1] = W; }
    dims[13][0] = 3; dims[13][1] = W;
    float* scale = packed + pack_scale_off(L);
    int* dims_dev = (int*)(packed + pack_dims_off(L));
    NA_TRY(upload_small(dims_dev, dims, sizeof(dims), stream));
    wn_scale_kernel<<<dim3((257 + 7) / 8, 14), 256, 0, stream>>>(*raw, dims_dev, scale);
    NA_CHECK_LAUNCH();

    FillTable tab; tab.n = 0;
    int fill_tf32 = 0;
    auto add = [&](size_t dst, int R, int C, int layer, int mode, int out0, int out_n, int map, float mult) {
        FillDesc& f = tab.d[tab.n++];
        f.dst = dst; f.R = R; f.C = C; f.layer = layer; f.mode = mode; f.out0 = out0; f.out_n = out_n;
        f.in_dim = dims[layer][1]; f.map = map; f.small = sdim; f.mult = mult; f.tf32 = fill_tf32;
    };
    const float inv_sqrt2 = (float)(1.0 / sqrt(2.0));
    for (int l = 0; l < 8; ++l) {
        const float mult = (l == 4) ? inv_sqrt2 : 1.f;
        add(L.sdf_wt[l], l == 0 ? EMB_PAD : W, W, l, 0, 0, dims[l][0], 0, mult);
        add(L.sdf_w[l], W, W, l, 1, 0, dims[l][0], 0, mult);
        add(L.sdf_b[l], 1, W, l, 2, 0, dims[l][0], 0, 1.f);
    }
    add(L.w8_sdf, 1, W, 8, 1, 0, 1, 0, 1.f);
    add(L.b8_sdf, 1, 4, 8, 2, 0, 1, 0, 1.f);
    add(L.w8t_feat, W, W, 8, 0, 1, W, 0, 1.f);
    add(L.b8_feat, 1, W, 8, 2, 1, W, 0, 1.f);
    add(L.rad_wt[0], W + spad, W, 9, 0, 0, W, 1, 1.f);
    for (int l = 1; l < 4; ++l) add(L.rad_wt[l], W, W, 9 + l, 0, 0, W, 0, 1.f);
    for (int l = 0; l < 4; ++l) add(L.rad_b[l], 1, W, 9 + l, 2, 0, W, 0, 1.f);
    add(L.rad_w4, 3, W, 13, 1, 0, 3, 0, 1.f);
    add(L.rad_b4, 1, 4, 13, 2, 0, 3, 0, 1.f);
    pack_fill_kernel<<<dim3(64, tab.n), 256, 0, stream>>>(*raw, tab, scale, packed);
    NA_CHECK_LAUNCH();
    {   // backward-only planes (train.cu): destination offsets are relative to the train pack
        const PackTrain TP = pack_layout_train();
        float* tp = (float*)((unsigned char*)packed_ + train_pack_off(desc->multires_view));
        tab.n = 0;
        add(TP.rad_w[0], W, W, 9, 1, 0, W, 1, 1.f);                     // dst[r][c] = W0[r][small + c]  (feature columns)
        for (int l = 1; l < 4; ++l) add(TP.rad_w[l], W, W, 9 + l, 1, 0, W, 0, 1.f);
        add(TP.rad_w0_small, W, W, 9, 1, 0, W, 2, 1.f);                 // dst[r][c] = W0[r][c], c < small_dim
        add(TP.w8_feat, W, W, 8, 1, 1, W, 0, 1.f);                      // dst[r][c] = W8[1 + r][c]
        fill_tf32 = 1;
        for (int l = 0; l < 8; ++l) {
            const float mult = (l == 4) ? inv_sqrt2 : 1.f;
            add(TP.sdf_wt_r[l], l == 0 ? EMB_PAD : W, W, l, 0, 0, dims[l][0], 0, mult);
            add(TP.sdf_w_r[l], W, W, l, 1, 0, dims[l][0], 0, mult);
        }
        add(TP.rad_w_r[0], W, W, 9, 1, 0, W, 1, 1.f);
        for (int l = 1; l < 4; ++l) add(TP.rad_w_r[l], W, W, 9 + l, 1, 0, W, 0, 1.f);
        add(TP.w8_feat_r, W, W, 8, 1, 1, W, 0, 1.f);
        fill_tf32 = 0;
        pack_fill_kernel<<<dim3(64, tab.n), 256, 0, stream>>>(*raw, tab, scale, tp);
        NA_CHECK_LAUNCH();
    }
    // tensor-core operand image (hi/lo fp16, UMMA K-major SWIZZLE_128B stages) derived from the fp32 planes
    const TcPackLayout T = tc_pack_layout(packed_f32_bytes(desc->multires_view));
    NA_TRY(tc_pack(packed, L, (unsigned char*)packed_, T, stream));
    // second image (TMEM-resident kernel: 256-row stages) from the same planes / per-plane scales (tables left by tc_pack)
    unsigned char* meta = (unsigned char*)packed_ + T.meta_off;
    return tmem_pack(packed, (const size_t*)meta, (const int*)(meta + 256), (const float*)(meta + 1280),
                     train_pack_off(desc->multires_view) / sizeof(float), pack_layout_train(),
                     (unsigned char*)packed_ + tmem_image_off(desc->multires_view), stream);
}

// -----------------------------------------------------------------------------------------------
extern "C" size_t na_eval_workspace_bytes(int64_t m) {
    (void)m;
    return mlp_scratch_bytes();
}

extern "C" int na_sdf_eval(const NaNetDesc* desc, const void* packed, const float* x, int64_t m, int apply_bg,
                           int precision, float* sdf, float* feat, void* workspace, size_t ws_bytes, void* stream) {
    if (m == 0) return NA_OK;
    if (!desc || !packed || !x || !sdf || !workspace || m < 0) return NA_ERR_BAD_ARG;
    EvalJob job = {};
    job.x = x; job.m = m; job.sdf = sdf; job.feat = feat;
    job.apply_bg = apply_bg; job.bound_r = desc->bounding_radius; job.want_full = 0; job.multires_view = desc->multires_view;
    return launch_mlp(job, packed, precision, (float*)workspace, ws_bytes, (cudaStream_t)stream);
}

extern "C" int na_full_eval(const NaNetDesc* desc, const void* packed, const float* x, const float* view, int64_t m,
                            int precision, float* radiance, float* sdf, float* nablas, float* feat,
                            void* workspace, size_t ws_bytes, void* stream) {
    if (m == 0) return NA_OK;
    if (!desc || !packed || !x || !workspace || m < 0) return NA_ERR_BAD_ARG;
    if (radiance && !view) return NA_ERR_BAD_ARG;
    EvalJob job = {};
    job.x = x; job.view = view; job.m = m; job.sdf = sdf; job.feat = feat; job.rad = radiance; job.nab = nablas;
    job.apply_bg = desc->framework == NA_FRAMEWORK_VOLSDF; job.bound_r = desc->bounding_radius;
    job.want_full = 1; job.multires_view = desc->multires_view;
    return launch_mlp(job, packed, precision, (float*)workspace, ws_bytes, (cudaStream_t)stream);
}

// -----------------------------------------------------------------------------------------------
// rend_util.get_rays (utils/rend_util.py:112-165) + lift (95-109), N_rays = -1
__global__ void get_rays_kernel(const float* __restrict__ c2w, const float* __restrict__ K, int H, int Wd,
                                float* __restrict__ ro, float* __restrict__ rd) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * Wd) return;
    const float i = (float)(idx % Wd), j = (float)(idx / Wd);        // x = w, y = h, integer pixel coordinates
    const float fx = K[0], fy = K[5], cx = K[2], cy = K[6], sk = K[1];
    // x_lift = (x - cx + cy*sk/fy - sk*y/fy) / fx * z ; y_lift = (y - cy)/fy * z ; z = 1
    const float xl = __fmul_rn(__fdiv_rn(__fsub_rn(__fadd_rn(__fsub_rn(i, cx), __fdiv_rn(__fmul_rn(cy, sk), fy)),
                                                    __fdiv_rn(__fmul_rn(sk, j), fy)), fx), 1.f);
    const float yl = __fmul_rn(__fdiv_rn(__fsub_rn(j, cy), fy), 1.f);
    const float p[4] = {xl, yl, 1.f, 1.f};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float w = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) w = fmaf(c2w[r * 4 + q], p[q], w);
        const float cam = c2w[r * 4 + 3];
        rd[idx * 3 + r] = __fsub_rn(w, cam);
        ro[idx * 3 + r] = cam;
    }
}

extern "C" int na_get_rays(const float* c2w, const float* intrinsics, int H, int Wd, float* rays_o, float* rays_d, void* stream) {
    if (!c2w || !intrinsics || !rays_o || !rays_d || H <= 0 || Wd <= 0) return NA_ERR_BAD_ARG;
    get_rays_kernel<<<(H * Wd + 255) / 256, 256, 0, (cudaStream_t)stream>>>(c2w, intrinsics, H, Wd, rays_o, rays_d);
    NA_CHECK_LAUNCH();
    return NA_OK;
}

// -----------------------------------------------------------------------------------------------
// stage-wise sampler entry points (one CTA per row; the renderers use the same device functions)
namespace na {
__global__ void __launch_bounds__(SNT) error_bound_rows_kernel(const float* __restrict__ d, const float* __restrict__ s, int n,
                                                               const float* __restrict__ ab_rows, float alpha, float beta,
                                                               float* __restrict__ bounds) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float* D = reinterpret_cast<float*>(smraw); float* Sv = D + n; float* X0 = Sv + n; float* X1 = X0 + n;
    __shared__ double red[8]; __shared__ float redf[8];
    const long long row = blockIdx.x;
    for (int i = threadIdx.x; i < n; i += SNT) { D[i] = d[row * n + i]; Sv[i] = s[row * n + i]; }
    if (ab_rows) { alpha = ab_rows[row * 2]; beta = ab_rows[row * 2 + 1]; }
    __syncthreads();
    error_bound(D, Sv, n, alpha, beta, X0, X1, false, red, redf);
    for (int i = threadIdx.x; i < n - 1; i += SNT) bounds[row * (n - 1) + i] = X1[i];
}

__global__ void __launch_bounds__(SNT) sample_rows_kernel(const float* __restrict__ bins, const float* __restrict__ wc, int n, int is_pdf,
                                                          const float* __restrict__ u, int u_per_row, int n_out,
                                                          float* __restrict__ samples, long long* __restrict__ inds) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float* B = reinterpret_cast<float*>(smraw); float* Wt = B + n; float* C = Wt + n;
    __shared__ double red[8];
    const long long row = blockIdx.x;
    for (int i = threadIdx.x; i < n; i += SNT) B[i] = bins[row * n + i];
    for (int i = threadIdx.x; i < n - 1; i += SNT) Wt[i] = wc[row * (n - 1) + i];
    __syncthreads();
    if (is_pdf) pdf_to_cdf(Wt, C, n, red);
    else { for (int i = threadIdx.x; i < n; i += SNT) C[i] = i == 0 ? 0.f : Wt[i - 1]; __syncthreads(); }
    for (int q = threadIdx.x; q < n_out; q += SNT) {
        int ind;
        const float uu = u_per_row ? u[row * n_out + q] : u[q];
        samples[row * n_out + q] = invert_cdf(B, C, n, uu, &ind);
        if (inds) inds[row * n_out + q] = ind;
    }
}
}  // namespace na

extern "C" int na_error_bound(const float* d_vals, const float* sdf, int64_t rows, int n, const float* ab_rows,
                              float alpha, float beta, float* bounds, void* stream) {
    if (!d_vals || !sdf || !bounds || rows <= 0 || n < 2) return NA_ERR_BAD_ARG;
    if (n > MAX_CAP) return NA_ERR_UNSUPPORTED;
    const size_t smem = (size_t)4 * n * sizeof(float);
    NA_TRY(check_cuda(cudaFuncSetAttribute(error_bound_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    error_bound_rows_kernel<<<(unsigned)rows, SNT, smem, (cudaStream_t)stream>>>(d_vals, sdf, n, ab_rows, alpha, beta, bounds);
    NA_CHECK_LAUNCH();
    return NA_OK;
}

static int sample_rows(const float* bins, const float* wc, int64_t rows, int n, int is_pdf, const float* u, int u_per_row,
                       int n_out, float* samples, int64_t* inds, void* stream) {
    if (!bins || !wc || !u || !samples || rows <= 0 || n < 2 || n_out <= 0) return NA_ERR_BAD_ARG;
    if (n > MAX_CAP) return NA_ERR_UNSUPPORTED;
    const size_t smem = (size_t)3 * n * sizeof(float);
    NA_TRY(check_cuda(cudaFuncSetAttribute(sample_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    sample_rows_kernel<<<(unsigned)rows, SNT, smem, (cudaStream_t)stream>>>(bins, wc, n, is_pdf, u, u_per_row, n_out, samples, (long long*)inds);
    NA_CHECK_LAUNCH();
    return NA_OK;
}
extern "C" int na_sample_pdf(const float* bins, const float* weights, int64_t rows, int n, const float* u, int u_per_row, int n_out,
                             float* samples, int64_t* inds, void* stream) {
    return sample_rows(bins, weights, rows, n, 1, u, u_per_row, n_out, samples, inds, stream);
}
extern "C" int na_sample_cdf(const float* bins, const float* cdf, int64_t rows, int n, const float* u, int u_per_row, int n_out,
                             float* samples, int64_t* inds, void* stream) {
    return sample_rows(bins, cdf, rows, n, 0, u, u_per_row, n_out, samples, inds, stream);
}

namespace na {
int preload_mlp_simt(); int preload_mlp_tc(); int preload_mlp_tmem(); int preload_wgrad_f16(); int preload_tgemm(); int preload_train();
int preload_volsdf(); int preload_neus(); int preload_surface();
namespace clipv { int preload_clip(); }
}
// Load every kernel image of the library on the current device now (cudaFuncGetAttributes forces the load) instead of at each
// kernel's first launch.  Called once per device when an engine is created: with the driver's default lazy module loading the
// first training step otherwise interleaves dozens of module loads with running persistent kernels.
extern "C" int na_preload_kernels(void) {
    NA_PRELOAD(wn_scale_kernel); NA_PRELOAD(pack_fill_kernel); NA_PRELOAD(get_rays_kernel);
    NA_PRELOAD(error_bound_rows_kernel); NA_PRELOAD(sample_rows_kernel);
    NA_TRY(preload_mlp_simt()); NA_TRY(preload_mlp_tc()); NA_TRY(preload_mlp_tmem()); NA_TRY(preload_wgrad_f16()); NA_TRY(preload_tgemm());
    NA_TRY(preload_train()); NA_TRY(preload_volsdf()); NA_TRY(preload_neus()); NA_TRY(preload_surface()); NA_TRY(clipv::preload_clip());
    return NA_OK;
}
