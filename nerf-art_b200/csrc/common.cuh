// Shared definitions for libnerfart_b200 (sm_100a).  See DESIGN.md for the data layout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/nerfart_b200.h"

namespace na {

// ---------------------------------------------------------------------------------------------
// launch accounting / error plumbing
// ---------------------------------------------------------------------------------------------
extern thread_local int g_last_cuda_error;
void count_launch(int n = 1);
inline int check_cuda(cudaError_t e) {
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return NA_ERR_CUDA; }
    return NA_OK;
}
// NA_DIAG_STREAM: the name of the launch stream at the NA_CHECK_LAUNCH() site (launch-trace diagnostics, below)
#define NA_DIAG_STREAM stream
#define NA_CHECK_LAUNCH()                                                     \
    do { na::count_launch();                                                  \
         cudaError_t _e = cudaGetLastError();                                 \
         if (_e != cudaSuccess) { na::g_last_cuda_error = (int)_e; return NA_ERR_CUDA; } \
         if (na::g_diag_on >= 2) na::diag_mark(__FILE__, __LINE__, (cudaStream_t)(NA_DIAG_STREAM)); } while (0)
#define NA_TRY(x) do { int _r = (x); if (_r != NA_OK) return _r; } while (0)
// small host table -> device memory, carried as a by-value kernel argument (csrc/api.cu): no pageable-memory cudaMemcpyAsync (which
// stages through the driver and synchronises the host) in the weight-packing path; n <= 1024 bytes, a multiple of 4
int upload_small(void* dst, const void* src, size_t n, cudaStream_t stream);
#define NA_PRELOAD(k) do { cudaFuncAttributes _a; NA_TRY(na::check_cuda(cudaFuncGetAttributes(&_a, k))); } while (0)

int num_sms();
// true the first time it is called for (flag array, current device): cudaFuncSetAttribute is per device, not per thread
inline bool first_on_device(bool (&done)[64]) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
}

// ---------------------------------------------------------------------------------------------
// Stall diagnostics.  Every mbarrier wait of the tcgen05 kernels is bounded (SPIN_LIMIT_NS of wall clock, %globaltimer): a wait
// that can never complete ends the launch with a trap -- a CUDA error the caller sees -- instead of blocking the stream forever.
// With na_diag_enable(1) the kernels additionally keep per-launch CTA counters and a record of every timed-out wait in
// host-mapped pinned memory (readable after the trap, and from a watchdog thread while the GPU is stuck), and the host side
// keeps an event per kernel launch so that the first launch that never finished can be named (na_diag_dump).
// ---------------------------------------------------------------------------------------------
struct HangRec { unsigned long long seq; int kernel, block, warp, lane; unsigned tag, parity; unsigned long long waited_ns; };
struct HangSlot { unsigned long long seq; int kernel, grid; unsigned started, ready, finished, pad; };
constexpr int HANG_SLOTS = 1024, HANG_RECS = 48;
struct HangDiag { unsigned n_rec, pad; HangRec rec[HANG_RECS]; HangSlot slot[HANG_SLOTS]; };
struct SpinCtx { HangDiag* diag; unsigned long long seq; int kernel; };
enum DiagKernel { DK_MLP_TMEM = 1, DK_WGRAD_TC = 2, DK_TGEMM = 3, DK_MLP_TC = 4 };
constexpr unsigned long long SPIN_LIMIT_NS = 8000000000ull;
extern int g_diag_on;                                 // 0 off, 1 device-side records + CTA counters, 2 + an event per kernel launch
void diag_mark(const char* file, int line, cudaStream_t stream);
SpinCtx diag_next(int kernel, int grid);               // host: sequence number + (when enabled) a fresh slot for the next launch

#if defined(__CUDACC__) && (!defined(__CUDA_ARCH__) || __CUDA_ARCH__ >= 900)
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
static __device__ __noinline__ void spin_timeout(const SpinCtx* scp, unsigned tag, unsigned parity, unsigned long long waited) {
    const SpinCtx sc = *scp;
    if (sc.diag && (threadIdx.x & 31) == 0) {
        const unsigned i = atomicAdd_system(&sc.diag->n_rec, 1u);
        if (i < (unsigned)HANG_RECS) {
            HangRec& r = sc.diag->rec[i];
            r.seq = sc.seq; r.kernel = sc.kernel; r.block = (int)(blockIdx.x + gridDim.x * blockIdx.y); r.warp = (int)(threadIdx.x >> 5);
            r.lane = (int)(threadIdx.x & 31); r.tag = tag; r.parity = parity; r.waited_ns = waited;
        }
        __threadfence_system();
    }
    // leave the other stuck warps of the grid time to file their own records before the trap takes the context down
    const unsigned long long t0 = global_ns();
    while (global_ns() - t0 < 500000000ull) __nanosleep(1000000);
    __trap();
}
__device__ __forceinline__ unsigned mbar_try_wait_(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
// Bounded mbarrier wait; `tag` names the wait site in the diagnostics (role << 24 | ...).  The context is passed by pointer to the
// kernel's __grid_constant__ parameter: a constant-bank address that costs no register on the hot path (carrying the context by
// value through the 96-register epilogue cost 4.6 % of the SDF-only tile, profiles/r3c_render.md).  try_wait suspends the thread in
// hardware for a system-defined time, so the loop body runs rarely; the wall clock is read every 8192nd failed attempt.
__device__ __forceinline__ void mbar_wait_guarded(unsigned bar, unsigned parity, const SpinCtx* sc, unsigned tag) {
#ifdef NA_NO_SPIN_GUARD
    while (!mbar_try_wait_(bar, parity)) {}
    return;
#endif
    // inline on purpose: an out-of-line wait loop makes every contended wait spill the caller's live registers around the call
    // (measured: +6 % on the SDF-only tile); only the time-out itself is a call
    if (mbar_try_wait_(bar, parity)) return;
    unsigned polls = 0; unsigned long long t0 = 0;
    while (!mbar_try_wait_(bar, parity)) {
        if (++polls == 8192u) {
            polls = 0;
            const unsigned long long t = global_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > SPIN_LIMIT_NS) spin_timeout(sc, tag, parity, t - t0);
        }
    }
}
// plain wait for the throughput-critical, register-starved epilogue warps: any deadlock also blocks the CTA's producer and MMA warps,
// whose waits are the bounded ones (the bounded loop in the 16 epilogue warps cost 3 % of the SDF-only tile)
__device__ __forceinline__ void mbar_wait_plain(unsigned bar, unsigned parity) { while (!mbar_try_wait_(bar, parity)) {} }
__device__ __forceinline__ void diag_count(const SpinCtx& sc, int which) {   // (once per CTA phase: the by-value read is fine here)       // 0 started, 1 ready (TMEM allocated), 2 finished
    if (sc.diag && threadIdx.x == 0) {
        HangSlot& s = sc.diag->slot[sc.seq % HANG_SLOTS];
        atomicAdd_system(which == 0 ? &s.started : (which == 1 ? &s.ready : &s.finished), 1u);
    }
}
#endif

// ---------------------------------------------------------------------------------------------
// network geometry (fixed; NaNetDesc documents where it comes from)
// ---------------------------------------------------------------------------------------------
constexpr int W = 256;              // hidden width
constexpr int EMB = 39;             // 3 + 3*2*6
constexpr int EMB_PAD = 40;
constexpr int SKIP_H = 217;         // 256 - 39: width of layer 3 (models/base.py:186-187)
constexpr int N_SDF_HID = 8;

__host__ __device__ inline int small_dim(int multires_view) {       // x(3) + view + nabla(3)
    return 3 + (multires_view < 0 ? 3 : 3 + 6 * multires_view) + 3;
}
__host__ __device__ inline int small_pad(int multires_view) { return (small_dim(multires_view) + 7) / 8 * 8; }

// Packed fp32 weights ("SIMT pack"): offsets in floats from the start of the packed buffer.
struct PackF32 {
    size_t sdf_wt[N_SDF_HID];       // forward  B[r=in][c=out]  [Kp][256]   (layer 4 pre-scaled by 1/sqrt2)
    size_t sdf_w[N_SDF_HID];        // backward B[r=out][c=in]  [256][256]  (zero padded)
    size_t sdf_b[N_SDF_HID];        // [256]
    size_t w8_sdf;                  // [256]  row 0 of layer 8
    size_t b8_sdf;                  // [4]
    size_t w8t_feat;                // [256][256]  rows 1..256 of layer 8, transposed
    size_t b8_feat;                 // [256]
    size_t rad_wt[4];               // layer 0: [256+small_pad][256] (feat rows first, then x|view|nabla), 1..3: [256][256]
    size_t rad_b[4];                // [256]
    size_t rad_w4;                  // [3][256]
    size_t rad_b4;                  // [4]
    size_t total;                   // floats
};
__host__ inline PackF32 pack_layout_f32(int multires_view) {
    PackF32 L; size_t o = 0;
    for (int i = 0; i < N_SDF_HID; ++i) { L.sdf_wt[i] = o; o += (size_t)(i == 0 ? EMB_PAD : W) * W; }
    for (int i = 0; i < N_SDF_HID; ++i) { L.sdf_w[i] = o; o += (size_t)W * W; }
    for (int i = 0; i < N_SDF_HID; ++i) { L.sdf_b[i] = o; o += W; }
    L.w8_sdf = o; o += W;  L.b8_sdf = o; o += 4;
    L.w8t_feat = o; o += (size_t)W * W;  L.b8_feat = o; o += W;
    L.rad_wt[0] = o; o += (size_t)(W + small_pad(multires_view)) * W;
    for (int i = 1; i < 4; ++i) { L.rad_wt[i] = o; o += (size_t)W * W; }
    for (int i = 0; i < 4; ++i) { L.rad_b[i] = o; o += W; }
    L.rad_w4 = o; o += 3 * W;  L.rad_b4 = o; o += 4;
    L.total = o;
    return L;
}

// Extra fp32 planes only the backward pass needs ("train pack"; follows the tensor-core images in the packed buffer,
// offset train_pack_off()).  Offsets in floats from the start of the train pack.  All planes [256][256], row stride 256.
struct PackTrain {
    size_t rad_w[4];                // [0]: layer 0, feature columns  B[r=out][c=feat]; [1..3]: B[r=out][c=in]
    size_t rad_w0_small;            // layer 0, small-input columns   B[r=out][c=small idx] (zero beyond small_dim)
    size_t w8_feat;                 // SDF head rows 1..256           B[r=feat][c=in]
    // TF32-rounded (cvt.rna) copies for the mma.sync backward-data GEMMs of train.cu:
    size_t sdf_wt_r[N_SDF_HID];     // forward planes  B[r=in][c=out]  ([40][256] for layer 0)   -> second-order sweep
    size_t sdf_w_r[N_SDF_HID];      // backward planes B[r=out][c=in]  (1..7 used)               -> trunk backward
    size_t rad_w_r[4];              // as rad_w
    size_t w8_feat_r;
    size_t total;
};
__host__ inline PackTrain pack_layout_train() {
    PackTrain T; size_t o = 0;
    for (int i = 0; i < 4; ++i) { T.rad_w[i] = o; o += (size_t)W * W; }
    T.rad_w0_small = o; o += (size_t)W * W;
    T.w8_feat = o; o += (size_t)W * W;
    for (int i = 0; i < N_SDF_HID; ++i) { T.sdf_wt_r[i] = o; o += (size_t)(i == 0 ? EMB_PAD : W) * W; }
    for (int i = 0; i < N_SDF_HID; ++i) { T.sdf_w_r[i] = o; o += (size_t)W * W; }
    for (int i = 0; i < 4; ++i) { T.rad_w_r[i] = o; o += (size_t)W * W; }
    T.w8_feat_r = o; o += (size_t)W * W;
    T.total = o;
    return T;
}
size_t train_pack_off(int multires_view);           // bytes from the start of the packed buffer (api.cu)

// ---------------------------------------------------------------------------------------------
// where the points of one MLP launch come from, and where results go
// ---------------------------------------------------------------------------------------------
// opaque storage of a CUtensorMap (cuda.h) inside kernel parameters
struct alignas(64) TmaMap { unsigned long long v[16]; };

struct EvalJob {
    // explicit points (na_sdf_eval / na_full_eval): x != nullptr, m points
    const float* x;  const float* view;  long long m;
    // ray points: point(row, j) = o[ray] + d[ray] * t[ray*t_stride + t_off + j], ray = row_ids ? row_ids[row] : row
    const float* rays_o;  const float* rays_d;      // [n_rays,3], d normalised
    const int*   row_ids;                            // nullable
    const int*   n_rows_dev;                         // nullable: row count lives on the device
    int          n_rows;                             // host row count (upper bound when n_rows_dev != nullptr)
    int          P;                                  // points per row
    const float* t;  long long t_stride;  int t_off;
    int          midpoints;                          // 1: t_j := 0.5*(t[j]+t[j+1])   (NeuS pts_mid, neus.py:312)
    // outputs.  index = explicit ? w : ray*o_stride + o_off + j
    long long o_stride;  int o_off;
    float* sdf;  float* feat;  float* rad;  float* nab;
    // behaviour
    int   apply_bg;  float bound_r;                  // VolSDF sphere background
    int   want_full;                                 // 0: SDF only; 1: + nablas (+ radiance if rad != nullptr)
    int   multires_view;
    long long* dbg;                                  // optional cycle counters (na_debug_set_buffer), else nullptr
    // Stash for the backward pass of a training patch (csrc/mlp_tmem.cu BW program -> csrc/wgrad_f16.cu): every (delta, input) pair a
    // weight gradient needs, as sample-major 16-bit planes [plane][st_mpad][256] indexed by the launch's flat sample number
    // (row-major, 512 B per sample: the layout the weight-gradient kernel's TMA boxes read); nullptr = not written.
    // All planes are bf16: the backward quantities (z-bar, v-bar, deltas, feature-bar) need its range (they follow the upstream
    // gradient, 1e-30 .. 1e4), and tcgen05 kind::f16 does not take an fp16 operand against a bf16 one (illegal instruction).
    // The wide planes are written with TMA tensor stores (st_store_map: boxes of 16 columns x 32 samples, one per epilogue warp and
    // pass): direct stores would be 32-byte pieces at a 512-byte stride, 32 LSU wavefronts per warp instruction.
    unsigned short* st_wide;  size_t st_mpad;
    TmaMap st_store_map;
    unsigned short* st_small;                        // [st_mpad][64] (36 / 12 used): the small radiance inputs x | embed(view) | nabla
    // backward in the same launch (BW program): upstream gradients per sample (nullable = 0), the remaining stash planes, and the
    // sphere-background mask of d L / d sdf (volsdf.py:349-357)
    int bw;  int bw_bg_mask;
    const float* bw_gsdf;  const float* bw_gnab;  const float* bw_grad;
    unsigned short* st_emb;  unsigned short* st_vb0; // [st_mpad][64] (39 used): encoding (fwd format), v-bar_0 (bf16)
    float* st_t0;  float* st_t1;                     // [st_mpad][4] fp32: delta of the radiance output layer; masked d L / d sdf
    // Split training program (csrc/train.cu): 0 = one launch does forward re-evaluation + backward (bw = 1), or a plain render (bw = 0);
    // 1 = the patch's forward render IS the forward half: program 0..20 with the forward stash planes written, softplus' codes and
    // ReLU masks persisted per TILE in tile_buf, the sphere-background flag left in st_t1[.][1] (bw = 0);
    // 2 = backward half: program 21..40 only, reading tile_buf, st_t1 and the forward launch's radiance (`rad` is an INPUT) (bw = 1)
    int bw_split;
    unsigned char* tile_buf;                         // [n_tiles][TILE_BUF_BYTES] (csrc/mlp_tmem.cu), modes 1 / 2
};

// planes of EvalJob::st_wide a forward launch fills (the backward kernels of csrc/train.cu add theirs; see the enum there)
constexpr int ST_IN = 0;        // 0..7   h_i = output of SDF layer i (layer 3: [h_3 | emb])
constexpr int ST_G = 16;        // 16..23 g_i = reverse-sweep value u_i * softplus'(z_i)
constexpr int ST_FEAT = 32;     // geometry feature
constexpr int ST_YS = 34;       // 34..37 outputs of radiance layers 0..3
// ... and the planes the backward program adds (same numbering as the enum in csrc/train.cu)
constexpr int ST_ZB = 8;        // 8..15  z-bar_i
constexpr int ST_VB = 24;       // 24..31 v-bar_{i+1}
constexpr int ST_FB = 33;       // d L / d feature
constexpr int ST_D = 38;        // 38..41 delta_0..3 of the radiance hidden layers

constexpr int ST_N_WIDE = 42;   // wide planes of the 16-bit stash
constexpr int ST_NLD = 64;      // row stride (elements) of the narrow 16-bit planes

// one weight-gradient task of csrc/wgrad_f16.cu: out[l][r] += sum_m L[p][m][l] * R[p][m][r] over p < npair, (+ bias_out[l] += sum_m L[0][m][l]);
// planes by number in the wide stash buffer (R: in the narrow buffer when r_narrow), *_bf16: element format of the plane (else fp16)
struct WgF16Task { int l_plane[2], r_plane[2], l_bf16[2], r_bf16[2]; int npair, r_narrow, n_valid, ldo; float* out; float* bias_out; };

}  // namespace na
