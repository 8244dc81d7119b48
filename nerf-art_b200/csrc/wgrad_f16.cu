// Weight-gradient GEMMs of a training patch on tcgen05 (sm_100a), fed by TMA:   out[l][r] += sum_m L[m][l] * R[m][r]   (+ a second pair)
//
// The operands are the 16-bit stash planes the backward program of csrc/mlp_tmem.cu leaves in HBM, sample-major [m][256] (512 B per
// sample).  The contraction runs over samples, so for the tensor core both operands are "MN-major" (the M / N index is the contiguous
// one): exactly what a plane looks like in memory.  Nothing is transposed or converted on the way:
//   * one elected lane issues TMA tensor loads (cp.async.bulk.tensor.2d, SWIZZLE_128B): a box of 64 columns x 64 samples lands in shared
//     memory as one column of the canonical MN-major SWIZZLE_128B operand layout (128-byte rows = 64 columns of one sample, 8-sample
//     groups of 1024 B); a stage = 64 samples x (256 columns of L + 256 or 64 columns of R) = 64 KB, three stages in flight;
//   * one elected lane issues tcgen05.mma.cta_group::1.kind::f16 with MN-major descriptors (M = 128, N = 256 | 64, K = 16 samples):
//     two accumulators [128 x N] fp32 (L columns 0..127 / 128..255) in TMEM, so every operand byte is read from HBM once per task;
//     forward planes are fp16 (or bf16), backward planes bf16 -- the A / B formats of the instruction descriptor follow the pair;
//   * eight warps sum the columns of the L tile from shared memory while it is resident (the bias gradients: db = sum_m delta[m][:]),
//     then read the accumulators (tcgen05.ld) and add them into the packed gradient with red.global.add.f32.
// CTA = (task, sample split); the number of splits of a task is proportional to its number of (L, R) pairs.
// The kernel is HBM-bound by construction: 64 KB of operands per 16 MMAs (2048 tensor-pipe cycles at most).
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <mutex>

namespace na {
namespace wf {

constexpr int THREADS = 320;                 // warp 0: TMA producer; warp 1: MMA issuer / TMEM owner; warps 2..9: bias sums, then epilogue
constexpr int KS = 64;                       // samples per stage
constexpr int NST = 3;
constexpr int ATOM_BYTES = KS * 128;         // 64 columns x KS samples of a 16-bit plane
constexpr int L_BYTES = 4 * ATOM_BYTES, R_BYTES = 4 * ATOM_BYTES, STAGE_BYTES = L_BYTES + R_BYTES;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(unsigned smem_dst, unsigned ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16_ss(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, unsigned (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
    unsigned pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// 2-D TMA tensor load: box (64 columns x KS rows) at (column c0, row c1) of the plane buffer -> shared memory, completing on `bar`
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// MN-major SWIZZLE_128B UMMA shared-memory descriptor: 128-byte rows along M / N (64 elements), 8 K-rows (samples) per 1024-byte swizzle
// atom; leading byte offset = distance between 64-element blocks along M / N (one TMA box), stride byte offset = distance between
// 8-sample groups (1024 B).  Bit layout as in csrc/mlp_tmem.cu.
__device__ __forceinline__ unsigned long long mn_desc(unsigned smem_addr) {
    return (unsigned long long)((smem_addr & 0x3FFFFu) >> 4) | ((unsigned long long)(ATOM_BYTES >> 4) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = f32 (bit 4), A / B format @7 / @10 (0 = f16, 1 = bf16), A and B MN-major (bits 15, 16), N>>3 @17, M>>4 @24
__host__ __device__ constexpr unsigned idesc_mn(int n, int a_bf16, int b_bf16) {
    return (1u << 4) | ((unsigned)a_bf16 << 7) | ((unsigned)b_bf16 << 10) | (1u << 15) | (1u << 16) | ((unsigned)(n >> 3) << 17) | ((128u >> 4) << 24);
}

struct __align__(1024) Smem {
    unsigned char st[NST][STAGE_BYTES];          // [L: 4 atoms | R: 4 atoms (or 1)]
    unsigned long long full[NST], empty[NST], acc_ready;
    unsigned tmem_base;
};

// rows (samples) of plane p start at p * mpad in the wide map; narrow planes live in their own map
struct Task {
    int l_plane[2], r_plane[2];      // wide-map plane numbers of the pairs (r_plane: narrow-map plane when r_narrow)
    unsigned idesc[2];
    int npair, r_narrow, n_valid;    // n_valid: output columns actually stored (<= N)
    int ldo;
    float* out; float* bias_out;     // bias_out (nullable): += column sums of pair 0's L plane
    int cta0, n_cta;                 // this task's CTAs: [cta0, cta0 + n_cta)
};
constexpr int MAX_TASKS = 16;
struct Table { Task t[MAX_TASKS]; int n; long long mpad; long long n_blk; };      // n_blk: KS-sample blocks that hold samples

__global__ void __launch_bounds__(THREADS, 1)
wgrad_f16_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_n, const Table tab, const __grid_constant__ SpinCtx sc) {
    extern __shared__ unsigned char smem_raw_[];
    Smem& S = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw_) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int ti = 0;
    while (ti + 1 < tab.n && (int)blockIdx.x >= tab.t[ti + 1].cta0) ++ti;
    const Task& t = tab.t[ti];
    const int split = (int)blockIdx.x - t.cta0;
    const long long per = (tab.n_blk + t.n_cta - 1) / t.n_cta;
    const long long blk0 = (long long)split * per;
    long long blk1 = blk0 + per; if (blk1 > tab.n_blk) blk1 = tab.n_blk;
    const int nblk = blk1 > blk0 ? (int)(blk1 - blk0) : 0;
    const int n_stage = nblk * t.npair;
    const int N = t.r_narrow ? 64 : 256;
    const unsigned r_bytes = t.r_narrow ? ATOM_BYTES : R_BYTES;

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(smem_u32(&S.full[s]), 1); mbar_init(smem_u32(&S.empty[s]), 1 + 8); }
        mbar_init(smem_u32(&S.acc_ready), 1);
        fence_barrier_init();
    }
    diag_count(sc, 0);
    if (warp == 1) tmem_alloc(smem_u32(&S.tmem_base), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_d = S.tmem_base;
    diag_count(sc, 1);

    if (n_stage > 0) {
        if (warp == 0) {
            // ---- TMA producer
            if (lane == 0) {
                for (int i = 0; i < n_stage; ++i) {
                    const int pr = i / nblk, b = i - pr * nblk;
                    const unsigned slot = i % NST, ph = (i / NST) & 1;
                    mbar_wait_guarded(smem_u32(&S.empty[slot]), ph ^ 1, &sc, 0x50000000u | (unsigned)(i & 0xffffff));
                    const unsigned bar = smem_u32(&S.full[slot]), base = smem_u32(S.st[slot]);
                    mbar_expect_tx(bar, L_BYTES + r_bytes);
                    const long long m0 = (blk0 + b) * KS;
                    const int lrow = (int)(t.l_plane[pr] * tab.mpad + m0), rrow = (int)(t.r_plane[pr] * tab.mpad + m0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) tma_load_2d(base + j * ATOM_BYTES, &map_w, 64 * j, lrow, bar);
                    if (t.r_narrow) tma_load_2d(base + L_BYTES, &map_n, 0, rrow, bar);
                    else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) tma_load_2d(base + L_BYTES + j * ATOM_BYTES, &map_w, 64 * j, rrow, bar);
                    }
                }
            }
        } else if (warp == 1) {
            // ---- MMA issuer: per stage 4 K-steps of 16 samples x 2 accumulators (L columns 0..127 -> D0, 128..255 -> D1)
            for (int i = 0; i < n_stage; ++i) {
                const int pr = i / nblk;
                const unsigned slot = i % NST, ph = (i / NST) & 1;
                mbar_wait_guarded(smem_u32(&S.full[slot]), ph, &sc, 0x4d000000u | (unsigned)(i & 0xffffff));
                tc_fence_after();
                if (elect_one()) {
                    const unsigned base = smem_u32(S.st[slot]);
                    const unsigned idesc = (t.idesc[pr] & ~(0x3fu << 17)) | ((unsigned)(N >> 3) << 17);
#pragma unroll
                    for (int ks = 0; ks < KS / 16; ++ks) {
                        const unsigned long long b = mn_desc(base + L_BYTES + ks * 2048);
                        umma_f16_ss(tmem_d, mn_desc(base + ks * 2048), b, idesc, (i | ks) != 0);
                        umma_f16_ss(tmem_d + 256u, mn_desc(base + 2 * ATOM_BYTES + ks * 2048), b, idesc, (i | ks) != 0);
                    }
                    umma_commit(smem_u32(&S.empty[slot]));
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(smem_u32(&S.acc_ready));
            __syncwarp();
        } else {
            // ---- bias sums of pair 0's L tile while it is resident: warp w takes samples w-2, w+6, ... of the stage, lane = one 16-byte
            //      chunk (8 columns) of the 256: a warp instruction reads one sample's 512 B, conflict-free in the swizzled layout
            const int rg = warp - 2;
            float bs[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) bs[j] = 0.f;
            const bool l_bf16 = (t.idesc[0] >> 7) & 1u;
            for (int i = 0; i < n_stage; ++i) {
                const unsigned slot = i % NST, ph = (i / NST) & 1;
                // (every warp waits for the stage even when it has nothing to sum: its arrival on `empty` must not run ahead of the
                //  stage, or it would be counted into the previous use of the slot)
                mbar_wait_plain(smem_u32(&S.full[slot]), ph);
                if (t.bias_out && i < nblk) {
                    const unsigned base = smem_u32(S.st[slot]) + (unsigned)(lane >> 3) * ATOM_BYTES;
#pragma unroll
                    for (int k = 0; k < KS / 8; ++k) {
                        const int m = rg + 8 * k;
                        uint4 q;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                                     : "r"(base + (unsigned)(m * 128 + (((lane & 7) ^ (m & 7)) << 4))));
                        const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float2 f;
                            if (l_bf16) f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
                            else f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
                            bs[2 * e] += f.x; bs[2 * e + 1] += f.y;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&S.empty[slot]));
            }
            if (t.bias_out) {
#pragma unroll
                for (int j = 0; j < 8; ++j) atomicAdd(t.bias_out + 8 * lane + j, bs[j]);
            }
            // ---- epilogue: warp w reads TMEM lanes 32 * (w % 4) .. of accumulator (w - 2) / 4
            mbar_wait_plain(smem_u32(&S.acc_ready), 0);
            tc_fence_after();
            const int q = warp & 3, h = (warp - 2) >> 2;
            const int l = 128 * h + 32 * q + lane;
            float* orow = t.out + (size_t)l * t.ldo;
#pragma unroll 1
            for (int c16 = 0; c16 < N / 16; ++c16) {
                unsigned v[16];
                tmem_ld16(tmem_d + ((unsigned)(32 * q) << 16) + (unsigned)(256 * h + 16 * c16), v);
                tmem_wait_ld();
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const int cc = 16 * c16 + 4 * j4;
                    if (cc + 3 < t.n_valid)            // 16-byte vector atomic (rows are 16-byte aligned: ldo is a multiple of 4)
                        atomicAdd(reinterpret_cast<float4*>(orow + cc), make_float4(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1]),
                                                                                    __uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])));
                    else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (cc + j < t.n_valid) atomicAdd(orow + cc + j, __uint_as_float(v[4 * j4 + j]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_d, 512); }
    diag_count(sc, 2);
}

// The weight-gradient rows whose left operand is one of the tiny fp32 planes (1 or 3 used columns) plus the u-bar_7 column sums:
//   w8_sdf[c] += sum_m t1[m] * IN7[m][c] + sum_m VB7[m][c]      b8_sdf += sum_m t1[m]           (SDF head, row 0 of layer 8)
//   rad_w4[j][c] += sum_m t0[m][j] * YS3[m][c]                   rad_b4[j] += sum_m t0[m][j]     (radiance output layer)
// Three 16-bit planes streamed once (coalesced: a warp reads one sample's 512 B as 32 x 16 B); block = 8 warps x a slice of samples.
struct TinyArgs {
    const unsigned short* in7; const unsigned short* vb7; const unsigned short* ys3;     // nullable each
    const float* t0; const float* t1;                                                      // [m][4]
    float* w8_sdf; float* b8_sdf; float* rad_w4; float* rad_b4;
    long long m_rows; int rows_per_block; int fwd_bf16;
};
__device__ __forceinline__ void unpack8(const uint4 q, bool bf16, float (&f)[8]) {
    const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float2 v;
        if (bf16) v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
        else v = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
        f[2 * e] = v.x; f[2 * e + 1] = v.y;
    }
}
__global__ void __launch_bounds__(256)
wgrad_tiny_kernel(const TinyArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m0 = (long long)blockIdx.x * a.rows_per_block;
    long long m1 = m0 + a.rows_per_block; if (m1 > a.m_rows) m1 = a.m_rows;
    float s8[8], r4[3][8], sb = 0.f, rb[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j) { s8[j] = 0.f; r4[0][j] = r4[1][j] = r4[2][j] = 0.f; }
    constexpr int U = 4;                                   // rows per warp iteration: 12 independent 16-byte loads in flight per lane
    for (long long mb = m0 + (long long)warp * U; mb < m1; mb += 8 * U) {
        uint4 qi[U], qv[U], qy[U]; float g[U]; float4 d[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long m = mb + u < m1 ? mb + u : m1 - 1;      // clamped rows are given zero weight below
            const bool live = mb + u < m1;
            if (a.in7) {
                qi[u] = __ldcs(reinterpret_cast<const uint4*>(a.in7 + m * 256) + lane);
                qv[u] = __ldcs(reinterpret_cast<const uint4*>(a.vb7 + m * 256) + lane);
                g[u] = live ? a.t1[m * 4] : 0.f;
            }
            if (a.ys3) {
                qy[u] = __ldcs(reinterpret_cast<const uint4*>(a.ys3 + m * 256) + lane);
                d[u] = live ? *reinterpret_cast<const float4*>(a.t0 + m * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float live = mb + u < m1 ? 1.f : 0.f;
            float f[8];
            if (a.in7) {
                unpack8(qi[u], a.fwd_bf16 != 0, f);
#pragma unroll
                for (int j = 0; j < 8; ++j) s8[j] = fmaf(g[u], f[j], s8[j]);
                unpack8(qv[u], true, f);
#pragma unroll
                for (int j = 0; j < 8; ++j) s8[j] = fmaf(live, f[j], s8[j]);
                sb += g[u];
            }
            if (a.ys3) {
                unpack8(qy[u], a.fwd_bf16 != 0, f);
#pragma unroll
                for (int j = 0; j < 8; ++j) { r4[0][j] = fmaf(d[u].x, f[j], r4[0][j]); r4[1][j] = fmaf(d[u].y, f[j], r4[1][j]); r4[2][j] = fmaf(d[u].z, f[j], r4[2][j]); }
                rb[0] += d[u].x; rb[1] += d[u].y; rb[2] += d[u].z;
            }
        }
    }
    // block-level reduction over the 8 warps, then one vector atomic per 4 columns: 256 x 16-byte atomics per block instead of 8192 scalar ones
    __shared__ float red[8][4][256];
    __shared__ float redb[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[warp][0][8 * lane + j] = s8[j]; red[warp][1][8 * lane + j] = r4[0][j]; red[warp][2][8 * lane + j] = r4[1][j]; red[warp][3][8 * lane + j] = r4[2][j]; }
    if (lane == 0) { redb[warp][0] = sb; redb[warp][1] = rb[0]; redb[warp][2] = rb[1]; redb[warp][3] = rb[2]; }
    __syncthreads();
    {
        const int row = threadIdx.x >> 6, c4 = (threadIdx.x & 63) * 4;          // 4 output rows x 64 column quads
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < 8; ++w) { const float4 t = *reinterpret_cast<const float4*>(&red[w][row][c4]); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
        if (row == 0) { if (a.in7) atomicAdd(reinterpret_cast<float4*>(a.w8_sdf + c4), v); }
        else if (a.ys3) atomicAdd(reinterpret_cast<float4*>(a.rad_w4 + (row - 1) * 256 + c4), v);
        if (threadIdx.x < 4) {
            float b = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) b += redb[w][threadIdx.x];
            if (threadIdx.x == 0) { if (a.in7) atomicAdd(a.b8_sdf, b); }
            else if (a.ys3) atomicAdd(a.rad_b4 + threadIdx.x - 1, b);
        }
    }
}

// ---- host: TMA tensor maps over the plane buffers ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}
// 2-D map over a 16-bit buffer of `rows` x `cols` elements (row-major) with boxes of box_cols x box_rows
static int make_map(CUtensorMap* map, const void* base, unsigned long long rows, unsigned cols, unsigned box_cols, unsigned box_rows, bool swizzle128) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return NA_ERR_UNSUPPORTED;
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? NA_OK : NA_ERR_CUDA;
}

}  // namespace wf

// tasks: see WgF16Task (common.cuh).  wide: [n_wide_planes][mpad][256], narrow: [n_narrow_planes][mpad][64] 16-bit; m_rows samples
// (mpad is a multiple of 128, rows in [m_rows, mpad) carry zero gradient planes).
int launch_wgrad_f16(const WgF16Task* tasks, int n_tasks, const unsigned short* wide, int n_wide_planes, const unsigned short* narrow,
                     int n_narrow_planes, long long mpad, long long m_rows, cudaStream_t stream) {
    using namespace wf;
    if (n_tasks <= 0 || m_rows <= 0) return NA_OK;
    if (n_tasks > MAX_TASKS || mpad % 128 || (unsigned long long)n_wide_planes * (unsigned long long)mpad > 0x7fffffffULL) return NA_ERR_UNSUPPORTED;
    static bool attr_set[64] = {false};
    int dev = 0; cudaGetDevice(&dev);
    const size_t smem = sizeof(Smem) + 1024;
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        NA_TRY(check_cuda(cudaFuncSetAttribute(wgrad_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        attr_set[dev] = true;
    }
    CUtensorMap map_w, map_n;
    // load boxes: 64 columns x KS samples, 128-byte swizzle (one column block of the MN-major operand layout)
    NA_TRY(make_map(&map_w, wide, (unsigned long long)n_wide_planes * mpad, 256, 64, KS, true));
    NA_TRY(make_map(&map_n, narrow ? narrow : wide, narrow ? (unsigned long long)n_narrow_planes * mpad : (unsigned long long)mpad, 64, 64, KS, true));
    Table tab; tab.n = n_tasks; tab.mpad = mpad; tab.n_blk = (m_rows + KS - 1) / KS;
    // CTAs per task proportional to its pair count (narrow-R pairs move 5/8 of the bytes), about one CTA per SM in total
    float weight[MAX_TASKS], wsum = 0.f;
    for (int i = 0; i < n_tasks; ++i) { weight[i] = tasks[i].npair * (tasks[i].r_narrow ? 0.625f : 1.f); wsum += weight[i]; }
    const int sms = num_sms();
    int cta = 0;
    for (int i = 0; i < n_tasks; ++i) {
        Task& t = tab.t[i];
        const WgF16Task& s = tasks[i];
        int n = (int)(weight[i] / wsum * (float)sms + 0.5f);
        if (n < 1) n = 1;
        if ((long long)n > tab.n_blk) n = (int)tab.n_blk;
        for (int p = 0; p < 2; ++p) {
            t.l_plane[p] = s.l_plane[p]; t.r_plane[p] = s.r_plane[p];
            t.idesc[p] = idesc_mn(256, s.l_bf16[p], s.r_bf16[p]);
        }
        t.npair = s.npair; t.r_narrow = s.r_narrow; t.n_valid = s.n_valid; t.ldo = s.ldo; t.out = s.out; t.bias_out = s.bias_out;
        t.cta0 = cta; t.n_cta = n; cta += n;
    }
    wgrad_f16_kernel<<<cta, THREADS, smem, stream>>>(map_w, map_n, tab, diag_next(DK_WGRAD_TC, cta));
    NA_CHECK_LAUNCH();
    return NA_OK;
}

// tensor map for the TMA stores of the backward program (csrc/mlp_tmem.cu): boxes of 16 columns x 32 samples of the wide planes, no swizzle
int make_stash_store_map(TmaMap* out, const unsigned short* wide, int n_wide_planes, long long mpad) {
    static_assert(sizeof(TmaMap) == sizeof(CUtensorMap), "TmaMap must hold a CUtensorMap");
    return wf::make_map(reinterpret_cast<CUtensorMap*>(out), wide, (unsigned long long)n_wide_planes * mpad, 256, 16, 32, false);
}

int launch_wgrad_tiny(const unsigned short* in7, const unsigned short* vb7, const unsigned short* ys3, const float* t0, const float* t1,
                      float* w8_sdf, float* b8_sdf, float* rad_w4, float* rad_b4, long long m_rows, int fwd_bf16, cudaStream_t stream) {
    if (m_rows <= 0 || (!in7 && !ys3)) return NA_OK;
    wf::TinyArgs a;
    a.in7 = in7; a.vb7 = vb7; a.ys3 = ys3; a.t0 = t0; a.t1 = t1; a.w8_sdf = w8_sdf; a.b8_sdf = b8_sdf; a.rad_w4 = rad_w4; a.rad_b4 = rad_b4;
    a.m_rows = m_rows; a.fwd_bf16 = fwd_bf16;
    const int blocks = 2 * num_sms();
    a.rows_per_block = (int)((m_rows + blocks - 1) / blocks);
    a.rows_per_block = (a.rows_per_block + 31) / 32 * 32;
    const int grid = (int)((m_rows + a.rows_per_block - 1) / a.rows_per_block);
    wf::wgrad_tiny_kernel<<<grid, 256, 0, stream>>>(a);
    NA_CHECK_LAUNCH();
    return NA_OK;
}

int preload_wgrad_f16() {
    NA_PRELOAD(wf::wgrad_f16_kernel); NA_PRELOAD(wf::wgrad_tiny_kernel);
    return NA_OK;
}

}  // namespace na
