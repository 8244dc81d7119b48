// Fused per-sample network kernel, tcgen05 tensor-core path (NA_PRECISION_TC).  sm_100a only.
//
// One persistent CTA per SM; each tile is 128 samples = the 128 TMEM lanes of one UMMA accumulator.
// Every layer GEMM  D[128 x 256] = A[128 x K] * B[256 x K]^T  runs on the 5th-gen tensor cores
// (tcgen05.mma.cta_group::1.kind::f16, M=128, N=128 per instruction, fp32 accumulate in TMEM).
//
// Precision: both operands are carried as TWO fp16 terms  v = hi + lo  (22 significant bits, power-of-two
// pre-scaling keeps lo out of the fp16 subnormal range), and three products  hi*hi + hi*lo + lo*hi  are
// accumulated -- the dropped lo*lo term is 2^-22 relative, i.e. fp32-level.  Activation functions, the
// softplus derivative needed by the reverse sweep, the narrow heads (sdf, rgb) and the small radiance
// inputs (x, view, nabla) stay in fp32 on the CUDA cores.
//
// Warp roles (10 warps):  warp 0 lane 0 streams pre-swizzled 16 KB weight stages with cp.async.bulk into a
// 4-deep mbarrier ring; warp 1 lane 0 issues the MMAs and owns TMEM; warps 2..9 are the epilogue: warp w reads TMEM
// lanes 32*(w%4).., i.e. thread = one sample row, two warps per row splitting the 256 columns; they apply bias /
// softplus / relu, split the result into hi/lo fp16 and write it back, in the UMMA K-major SWIZZLE_128B layout,
// as the A operand of the next layer.  Activations never leave the SM.
#include "common.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace na {

constexpr int TC_TM = 128;
constexpr int TC_THREADS = 576;
constexpr int TC_EPI_THREADS = 512;
constexpr int TC_NS = 4;                         // weight stages
constexpr int STAGE_BYTES = 16384;               // 128 rows x 64 fp16
constexpr int A_SPLIT_BYTES = 4 * STAGE_BYTES;   // 128 rows x 256 fp16
constexpr float ACT_SCALE = 16.f;                // activations are stored x16 (keeps the lo term normal in fp16)
constexpr int TC_MAX_GEMM = 24;
#ifndef NA_TC_TWO_ACC
#define NA_TC_TWO_ACC 1      // 1: hi*hi and the two correction products accumulate in separate TMEM accumulators (fewer truncations)
#endif

struct TcGemm { unsigned w_stage0; unsigned char n_kb, n_nh, pad0, pad1; };
struct TcProgram { int n_gemm; TcGemm g[TC_MAX_GEMM]; };

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ unsigned mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
// 32 lanes x 32 columns: thread i of the warp receives lane (base+i), columns c..c+31
__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
    unsigned pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// K-major, SWIZZLE_128B UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (1024 B between 8-row groups) |
//   [46,48) version=1 | [61,64) layout=2 (SWIZZLE_128B)
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr) {
    return (unsigned long long)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D=f32 (bit 4), A=B=f16 (0), both K-major, N=128 (n_dim=16 @17), M=128 (m_dim=8 @24)
constexpr unsigned UMMA_IDESC = (1u << 4) | (16u << 17) | (8u << 24);

// byte offset of element (row r, k) inside one 128-row x 256-k split buffer (four 16 KB K-blocks)
__device__ __forceinline__ unsigned a_off(int r, int k) {
    return (unsigned)((k >> 6) * STAGE_BYTES + (r >> 3) * 1024 + (r & 7) * 128 + ((((k & 63) >> 3) ^ (r & 7)) << 4) + ((k & 7) << 1));
}

constexpr int N_BIAS_ROWS = 13;                  // 0..7 sdf fwd (x ACT_SCALE) | 8 feature (raw) | 9..12 radiance (x ACT_SCALE)

struct __align__(1024) TcSmem {
    unsigned char A1[A_SPLIT_BYTES];
    unsigned char A2[A_SPLIT_BYTES];
    unsigned char Wst[TC_NS * STAGE_BYTES];
    unsigned long long full_bar[TC_NS], empty_bar[TC_NS], d_ready, kb_ready[4];   // kb_ready[k]: A K-block k written AND D columns [64k,64k+64) drained
    unsigned tmem_base;
    __align__(16) float BIAS[N_BIAS_ROWS * 256];
    __align__(16) float W8[256];                       // row 0 of SDF layer 8 (the sdf head)
    __align__(16) float W4[3 * 256];     // radiance output layer
    float X[3 * TC_TM];
    float V[3 * TC_TM];
    float PART[4 * 3 * TC_TM];           // per column-quarter partial sums of the narrow heads
    long long OIDX[TC_TM];
};

// 32 lanes x 16 columns
__device__ __forceinline__ void tmem_ld16(unsigned taddr, unsigned (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr)); return v;
}
__device__ __forceinline__ float lds32(unsigned addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void sts128(unsigned addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// split 8 consecutive fp32 values (already x ACT_SCALE) into hi/lo fp16 and store 16 B + 16 B into the A operand.
// a1 / a2: shared-space byte addresses of the hi / lo split buffers; row_off = (r/8)*1024 + (r%8)*128, sw = r%8
__device__ __forceinline__ void store8s(unsigned a1, unsigned a2, unsigned row_off, unsigned sw, int k0, const float* h) {
    __half2 hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        hi[i] = __floats2half2_rn(h[2 * i], h[2 * i + 1]);
        const float2 back = __half22float2(hi[i]);
        lo[i] = __floats2half2_rn(h[2 * i] - back.x, h[2 * i + 1] - back.y);
    }
    const unsigned off = (unsigned)(k0 >> 6) * STAGE_BYTES + row_off + (((unsigned)((k0 & 63) >> 3) ^ sw) << 4);
    sts128(a1 + off, *reinterpret_cast<const uint4*>(hi));
    sts128(a2 + off, *reinterpret_cast<const uint4*>(lo));
}
// split 8 consecutive fp32 values (already x ACT_SCALE) into hi/lo fp16 and store 16 B + 16 B
__device__ __forceinline__ void store8(unsigned char* A1, unsigned char* A2, int r, int k0, const float* h) {
    __half2 hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        hi[i] = __floats2half2_rn(h[2 * i], h[2 * i + 1]);
        const float2 back = __half22float2(hi[i]);
        lo[i] = __floats2half2_rn(h[2 * i] - back.x, h[2 * i + 1] - back.y);
    }
    const unsigned off = a_off(r, k0);
    *reinterpret_cast<uint4*>(A1 + off) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(A2 + off) = *reinterpret_cast<const uint4*>(lo);
}

// scratch planes: value (plane p, column k, row r) lives at float index ((p*64 + k/4)*128 + r)*4 + k%4
__device__ __forceinline__ float4* plane_ptr(float* sp, int p, int k4, int r) {
    return reinterpret_cast<float4*>(sp) + ((size_t)(p * 64 + k4) * TC_TM + r);
}

enum EpiKind { K_FWD, K_FWD3, K_FWD7, K_FEAT, K_BWD, K_BWD4, K_BWD0, K_RAD0, K_RAD, K_RAD3 };

struct EpiCtx {
    TcSmem* S; float* sp; float* misc; const float* pk; const PackF32* L; const EvalJob* job;
    unsigned t_row; int r, cq, g; float us; int sdim;
    unsigned a1, a2, row_off, sw, bias_s, w8_s, w4_s;      // shared-space byte addresses
    unsigned kb_bar; int signal, lane;                     // kb_ready[0] address; arrive after every 64-column group?
};

// this warp's part of A K-block kb is written and its part of D columns [64kb, 64kb+64) is drained
__device__ __forceinline__ void signal_kb(unsigned kb_bar, int kb, int lane) {
    tc_fence_before();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(kb_bar + 8u * (unsigned)kb);
}

// one GEMM's epilogue for this thread's row and its 64 columns (4 chunks of 16)
template <int KIND, bool FULL, bool TWO_ACC>
__device__ __forceinline__ void epi_gemm(const EpiCtx& c, float& sdf_part, float (&rgb_part)[3], const float (&small_in)[36]) {
    TcSmem& S = *c.S;
    const int r = c.r;
    const float us = c.us, us16 = c.us * ACT_SCALE;
    const unsigned bias = c.bias_s + (unsigned)(KIND == K_FEAT ? 8 : (KIND >= K_RAD0 ? 9 + (c.g - 17) : c.g)) * 1024u;
    const int bl = 16 - c.g;
#pragma unroll 1
    for (int c16 = 0; c16 < 4; ++c16) {
        // thread = (row, column quarter cq): in pass c16 it owns columns 64*c16 + 16*cq .. +16, i.e. every pass completes one
        // 64-wide K-block of the next layer's A operand across the 16 epilogue warps
        const int col0 = c16 * 64 + c.cq * 16;
        if (KIND == K_BWD0 && c16 > 0) break;                       // only 39 useful columns, all in pass 0
        unsigned v[16];
        tmem_ld16(c.t_row + col0, v);
        float acc[16];
        if (TWO_ACC) {
            unsigned v2[16];
            tmem_ld16(c.t_row + 256 + col0, v2);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(v[j]) + __uint_as_float(v2[j]);
        } else {
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(v[j]);
        }
        float o[16];
        if (KIND == K_FWD || KIND == K_FWD3 || KIND == K_FWD7) {
            // z16 = 16 z ; softplus_100(z) = max(z,0) + ln2/100 * log2(1 + 2^(-|100 z| log2 e)).  Written stage by stage over the 16
            // columns so that 16 independent MUFU.EX2 / MUFU.LG2 are in flight per warp.
            float z16[16], t[16];
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 b4 = lds128(bias + (unsigned)(col0 + 4 * j4) * 4u);
                z16[4 * j4] = fmaf(acc[4 * j4], us16, b4.x); z16[4 * j4 + 1] = fmaf(acc[4 * j4 + 1], us16, b4.y);
                z16[4 * j4 + 2] = fmaf(acc[4 * j4 + 2], us16, b4.z); z16[4 * j4 + 3] = fmaf(acc[4 * j4 + 3], us16, b4.w);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) t[j] = ex2_approx(-fabsf(z16[j]) * (100.f * 1.4426950408889634f / ACT_SCALE));
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = lg2_approx(1.f + t[j]);
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = fmaf(o[j], ACT_SCALE * 0.6931471805599453f / 100.f, fmaxf(z16[j], 0.f));
            if (FULL) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    float dh[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { const int j = 4 * j4 + i; const float ru = rcp_approx(1.f + t[j]); dh[i] = z16[j] >= 0.f ? ru : t[j] * ru; }
                    if (KIND == K_FWD3) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) if (col0 + 4 * j4 + i >= SKIP_H) dh[i] = 0.f;
                    }
                    *plane_ptr(c.sp, c.g, (col0 >> 2) + j4, r) = make_float4(dh[0], dh[1], dh[2], dh[3]);
                }
            }
            if (KIND == K_FWD3 && col0 + 15 >= SKIP_H) {
                // skip connection columns: h = emb[k-217] (x16)
                const float xs[3] = {S.X[r], S.X[TC_TM + r], S.X[2 * TC_TM + r]};
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int k = col0 + j;
                    if (k >= SKIP_H) {
                        const int ei = k - SKIP_H;
                        float val;
                        if (ei < 3) val = xs[ei];
                        else { const int f = (ei - 3) / 6, rem = (ei - 3) % 6; float sn, cs; sincosf(__fmul_rn(xs[rem % 3], (float)(1 << f)), &sn, &cs); val = rem < 3 ? sn : cs; }
                        o[j] = val * ACT_SCALE;
                    }
                }
            }
            if (KIND == K_FWD7) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 w4 = lds128(c.w8_s + (unsigned)(col0 + 4 * j4) * 4u);
                    sdf_part = fmaf(o[4 * j4], w4.x, sdf_part); sdf_part = fmaf(o[4 * j4 + 1], w4.y, sdf_part);
                    sdf_part = fmaf(o[4 * j4 + 2], w4.z, sdf_part); sdf_part = fmaf(o[4 * j4 + 3], w4.w, sdf_part);
                }
            }
        } else if (KIND == K_FEAT) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 b4 = lds128(bias + (unsigned)(col0 + 4 * j4) * 4u);
                const float4 f4 = make_float4(fmaf(acc[4 * j4], us, b4.x), fmaf(acc[4 * j4 + 1], us, b4.y),
                                              fmaf(acc[4 * j4 + 2], us, b4.z), fmaf(acc[4 * j4 + 3], us, b4.w));
                if (FULL) *plane_ptr(c.sp, 8, (col0 >> 2) + j4, r) = f4;
                if (c.job->feat && S.OIDX[r] >= 0) *(reinterpret_cast<float4*>(c.job->feat + S.OIDX[r] * 256 + col0) + j4) = f4;
                if (FULL) {
                    // next A: d sdf / d z7 = W8[0,:] * softplus'(z7)
                    const float4 d4 = *plane_ptr(c.sp, 7, (col0 >> 2) + j4, r);
                    const float4 w4 = lds128(c.w8_s + (unsigned)(col0 + 4 * j4) * 4u);
                    o[4 * j4] = w4.x * d4.x * ACT_SCALE; o[4 * j4 + 1] = w4.y * d4.y * ACT_SCALE;
                    o[4 * j4 + 2] = w4.z * d4.z * ACT_SCALE; o[4 * j4 + 3] = w4.w * d4.w * ACT_SCALE;
                }
            }
        } else if (KIND == K_BWD || KIND == K_BWD4) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 d4 = *plane_ptr(c.sp, bl - 1, (col0 >> 2) + j4, r);
                const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = col0 + 4 * j4 + i;
                    if (KIND == K_BWD4 && k >= SKIP_H) c.misc[r * 80 + (k - SKIP_H)] = acc[4 * j4 + i] * us;      // embedding branch of the skip
                    o[4 * j4 + i] = acc[4 * j4 + i] * us16 * dd[i];
                }
            }
        } else if (KIND == K_BWD0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) { const int k = col0 + j; if (k < EMB) c.misc[r * 80 + k] += acc[j] * us; }
        } else {
            // radiance hidden layers: relu(16 z)
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 b4 = lds128(bias + (unsigned)(col0 + 4 * j4) * 4u);
                const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = col0 + 4 * j4 + i;
                    float z = fmaf(acc[4 * j4 + i], us16, bb[i]);
                    if (KIND == K_RAD0) {
                        // small inputs [x | embed(view) | nabla] (x16) in fp32: rows 256.. of the packed layer-0 plane
                        const float* wsm = c.pk + c.L->rad_wt[0] + (size_t)256 * 256 + k;
                        if (c.sdim == 9) {
#pragma unroll
                            for (int j = 0; j < 9; ++j) z = fmaf(small_in[j], __ldg(wsm + j * 256), z);
                        } else {
#pragma unroll
                            for (int j = 0; j < 33; ++j) z = fmaf(small_in[j], __ldg(wsm + j * 256), z);
                        }
                    }
                    o[4 * j4 + i] = fmaxf(z, 0.f);
                }
            }
            if (KIND == K_RAD3) {
#pragma unroll
                for (int cc = 0; cc < 3; ++cc)
#pragma unroll
                    for (int j = 0; j < 16; ++j) rgb_part[cc] = fmaf(o[j], lds32(c.w4_s + (unsigned)(cc * 256 + col0 + j) * 4u), rgb_part[cc]);
            }
        }
        const bool store = !(KIND == K_BWD0 || KIND == K_RAD3 || (KIND == K_FWD7 && !FULL && !c.job->feat) || (KIND == K_FEAT && !FULL));
        if (store) {
            store8s(c.a1, c.a2, c.row_off, c.sw, col0, o);
            store8s(c.a1, c.a2, c.row_off, c.sw, col0 + 8, o + 8);
        }
        if (KIND != K_BWD0 && c.signal) signal_kb(c.kb_bar, c16, c.lane);
    }
}

template <bool FULL, bool TWO_ACC>
__global__ void __launch_bounds__(TC_THREADS, 1)
mlp_tc_kernel(const EvalJob job, const float* __restrict__ pk, const PackF32 L, const unsigned char* __restrict__ wtc,
              const float* __restrict__ unscale, const TcProgram prog, float* __restrict__ scratch) {
    extern __shared__ unsigned char smem_raw_[];
    TcSmem& S = *reinterpret_cast<TcSmem*>((reinterpret_cast<uintptr_t>(smem_raw_) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool explicit_pts = job.x != nullptr;
    const long long total = explicit_pts ? job.m
                          : (long long)(job.n_rows_dev ? min(*job.n_rows_dev, job.n_rows) : job.n_rows) * job.P;
    const long long n_tiles = (total + TC_TM - 1) / TC_TM;

    if (tid == 0) {
        for (int s = 0; s < TC_NS; ++s) { mbar_init(smem_u32(&S.full_bar[s]), 1); mbar_init(smem_u32(&S.empty_bar[s]), 1); }
        mbar_init(smem_u32(&S.d_ready), 1);
        for (int k = 0; k < 4; ++k) mbar_init(smem_u32(&S.kb_ready[k]), TC_EPI_THREADS / 32);   // one arrive per epilogue warp
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(&S.tmem_base), TWO_ACC ? 512 : 256);
    // bias rows and head weights -> shared memory (hidden-layer biases pre-multiplied by ACT_SCALE)
    for (int i = tid; i < N_BIAS_ROWS * 256; i += TC_THREADS) {
        const int row = i >> 8, k = i & 255;
        float b;
        if (row < 8) b = pk[L.sdf_b[row] + k] * ACT_SCALE;
        else if (row == 8) b = pk[L.b8_feat + k];
        else b = pk[L.rad_b[row - 9] + k] * ACT_SCALE;
        S.BIAS[i] = b;
    }
    for (int i = tid; i < 256; i += TC_THREADS) S.W8[i] = pk[L.w8_sdf + i];
    for (int i = tid; i < 768; i += TC_THREADS) S.W4[i] = pk[L.rad_w4 + i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_d = S.tmem_base;

    if (warp == 0) {
        // ================= weight producer =================
        if (lane == 0) {
            unsigned it = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int g = 0; g < prog.n_gemm; ++g) {
                    const unsigned char* src = wtc + (size_t)prog.g[g].w_stage0 * STAGE_BYTES;
                    const int n_st = prog.g[g].n_nh * prog.g[g].n_kb * 2;
                    for (int st = 0; st < n_st; ++st, ++it) {
                        const unsigned slot = it % TC_NS, ph = (it / TC_NS) & 1;
                        mbar_wait(smem_u32(&S.empty_bar[slot]), ph ^ 1);
                        mbar_expect_tx(smem_u32(&S.full_bar[slot]), STAGE_BYTES);
                        bulk_g2s(smem_u32(S.Wst + slot * STAGE_BYTES), src + (size_t)st * STAGE_BYTES, STAGE_BYTES, smem_u32(&S.full_bar[slot]));
                    }
                }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // the whole warp walks the program (converged control flow, waits included); one elected lane issues tcgen05.mma / commit
        {
            unsigned it = 0, a_phase = 0;
            long long t_a = 0, t_full = 0, t_tot0 = clock64();
            const unsigned a1 = smem_u32(S.A1), a2 = smem_u32(S.A2), wst = smem_u32(S.Wst);
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int g = 0; g < prog.n_gemm; ++g) {
                    const int n_kb = prog.g[g].n_kb, n_nh = prog.g[g].n_nh;
                    int waited = 0;
                    for (int nh = 0; nh < n_nh; ++nh)
                        for (int kb = 0; kb < n_kb; ++kb) {
                            // N-half 0 overwrites D columns 0..127: needs K-blocks 0,1 of the previous epilogue (they drain those
                            // columns) plus A K-block kb; N-half 1 needs everything
                            const int need = nh == 0 ? (kb > 1 ? kb : 1) + 1 : 4;
                            if (waited < need) {
                                const long long t0 = clock64();
                                for (; waited < need; ++waited) mbar_wait(smem_u32(&S.kb_ready[waited]), a_phase);
                                t_a += clock64() - t0;
                                tc_fence_after();
                            }
                            for (int s = 0; s < 2; ++s, ++it) {
                                const unsigned slot = it % TC_NS, ph = (it / TC_NS) & 1;
                                { const long long t0 = clock64(); mbar_wait(smem_u32(&S.full_bar[slot]), ph); t_full += clock64() - t0; }
                                tc_fence_after();
                                if (elect_one()) {
                                    const unsigned d = tmem_d + nh * 128;
                                    const unsigned dc = TWO_ACC ? d + 256 : d;        // accumulator of the two correction products
                                    const unsigned long long bd = umma_desc(wst + slot * STAGE_BYTES);
                                    const unsigned long long ad1 = umma_desc(a1 + kb * STAGE_BYTES);
                                    const unsigned long long ad2 = umma_desc(a2 + kb * STAGE_BYTES);
                                    if (s == 0) {
#pragma unroll
                                        for (int ks = 0; ks < 4; ++ks) umma_f16_ss(d, ad1 + 2 * ks, bd + 2 * ks, UMMA_IDESC, (kb | ks) != 0);                    // hi * hi
#pragma unroll
                                        for (int ks = 0; ks < 4; ++ks) umma_f16_ss(dc, ad2 + 2 * ks, bd + 2 * ks, UMMA_IDESC, TWO_ACC ? (kb | ks) != 0 : 1);    // lo * hi
                                    } else {
#pragma unroll
                                        for (int ks = 0; ks < 4; ++ks) umma_f16_ss(dc, ad1 + 2 * ks, bd + 2 * ks, UMMA_IDESC, 1);                                 // hi * lo
                                    }
                                    umma_commit(smem_u32(&S.empty_bar[slot]));
                                }
                                __syncwarp();
                            }
                        }
                    for (; waited < 4; ++waited) mbar_wait(smem_u32(&S.kb_ready[waited]), a_phase);
                    a_phase ^= 1;
                    if (elect_one()) umma_commit(smem_u32(&S.d_ready));
                    __syncwarp();
                }
            if (job.dbg && blockIdx.x == 0 && lane == 0) { job.dbg[0] = clock64() - t_tot0; job.dbg[1] = t_a; job.dbg[2] = t_full; }
        }
    } else {
        // ================= epilogue warps =================
        const int q = warp & 3, cq = (warp - 2) >> 2;
        const int r = 32 * q + lane;                       // sample row == TMEM lane
        float* sp = scratch + (size_t)blockIdx.x * (10 * 256 * TC_TM);     // planes 0..7 softplus', 8 feature, 9 misc rows
        EpiCtx c;
        c.S = &S; c.sp = sp; c.misc = sp + (size_t)9 * 256 * TC_TM;       // misc [row][80]: d sdf/d emb (39) @0 | small radiance inputs (<=33) @40
        c.pk = pk; c.L = &L; c.job = &job; c.t_row = tmem_d + ((unsigned)(32 * q) << 16); c.r = r; c.cq = cq;
        c.sdim = small_dim(job.multires_view);
        c.a1 = smem_u32(S.A1); c.a2 = smem_u32(S.A2); c.row_off = (unsigned)((r >> 3) * 1024 + (r & 7) * 128); c.sw = (unsigned)(r & 7);
        c.bias_s = smem_u32(S.BIAS); c.w8_s = smem_u32(S.W8); c.w4_s = smem_u32(S.W4);
        c.kb_bar = smem_u32(&S.kb_ready[0]); c.lane = lane; c.signal = 0;
        unsigned d_phase = 0;
        long long t_d = 0, t_e0 = clock64();
        const bool has_rad = job.rad != nullptr;

        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            // ---- tile inputs: point, encoding (x ACT_SCALE, hi/lo) into K-block 0 ----------------------------------
            if (cq == 0) {
                const long long w = tile * TC_TM + r;
                float x0 = 0.f, x1 = 0.f, x2 = 0.f, v0 = 0.f, v1 = 0.f, v2 = 1.f;
                long long oidx = -1;
                if (w < total) {
                    if (explicit_pts) {
                        x0 = job.x[w * 3]; x1 = job.x[w * 3 + 1]; x2 = job.x[w * 3 + 2];
                        if (job.view) { v0 = job.view[w * 3]; v1 = job.view[w * 3 + 1]; v2 = job.view[w * 3 + 2]; }
                        oidx = w;
                    } else {
                        const long long row = w / job.P; const int j = (int)(w - row * job.P);
                        const long long ray = job.row_ids ? job.row_ids[row] : row;
                        const float* tp = job.t + ray * job.t_stride + job.t_off + j;
                        float t = tp[0];
                        if (job.midpoints) t = __fmul_rn(0.5f, __fadd_rn(tp[1], t));
                        v0 = job.rays_d[ray * 3]; v1 = job.rays_d[ray * 3 + 1]; v2 = job.rays_d[ray * 3 + 2];
                        x0 = __fadd_rn(job.rays_o[ray * 3], __fmul_rn(v0, t));
                        x1 = __fadd_rn(job.rays_o[ray * 3 + 1], __fmul_rn(v1, t));
                        x2 = __fadd_rn(job.rays_o[ray * 3 + 2], __fmul_rn(v2, t));
                        oidx = ray * job.o_stride + job.o_off + j;
                    }
                }
                S.OIDX[r] = oidx;
                S.X[r] = x0; S.X[TC_TM + r] = x1; S.X[2 * TC_TM + r] = x2;
                S.V[r] = v0; S.V[TC_TM + r] = v1; S.V[2 * TC_TM + r] = v2;
                float emb[64];
                const float xs[3] = {x0, x1, x2};
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) emb[cc] = xs[cc] * ACT_SCALE;
#pragma unroll
                for (int f = 0; f < 6; ++f)
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) {
                        float sn, cs; sincosf(__fmul_rn(xs[cc], (float)(1 << f)), &sn, &cs);
                        emb[3 + 6 * f + cc] = sn * ACT_SCALE; emb[6 + 6 * f + cc] = cs * ACT_SCALE;
                    }
#pragma unroll
                for (int k = EMB; k < 64; ++k) emb[k] = 0.f;
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) store8s(c.a1, c.a2, c.row_off, c.sw, 8 * c8, emb + 8 * c8);
            }
            for (int k = 0; k < 4; ++k) signal_kb(c.kb_bar, k, lane);

            float sdf_part = 0.f, rgb_part[3] = {0.f, 0.f, 0.f};
            float small_in[36];
            for (int g = 0; g < prog.n_gemm; ++g) {
                { const long long t0 = clock64(); mbar_wait(smem_u32(&S.d_ready), d_phase); d_phase ^= 1; t_d += clock64() - t0; }
                tc_fence_after();
                c.g = g; c.us = unscale[g];                      // us = 2^-(weight shift) / ACT_SCALE
                c.signal = g + 1 < prog.n_gemm;
                // program order: 0..7 fwd | 8 feat | 9..15 bwd 7..1 | 16 bwd 0 | 17..20 radiance
                if (g < 8) {
                    if (g == 3) epi_gemm<K_FWD3, FULL, TWO_ACC>(c, sdf_part, rgb_part, small_in);
                    else if (g == 7) epi_gemm<K_FWD7, FULL, TWO_ACC>(c, sdf_part, rgb_part, small_in);
                    else epi_gemm<K_FWD, FULL, TWO_ACC>(c, sdf_part, rgb_part, small_in);
                } else if (g == 8) {
                    epi_gemm<K_FEAT, FULL, TWO_ACC>(c, sdf_part, rgb_part, small_in);
                } else if (FULL) {
                    if (g <= 15) {
                        if (g == 12) epi_gemm<K_BWD4, FULL, TWO_ACC>(c, sdf_part, rgb_part, small_in);
                        else epi_gemm<K_BWD, FULL, TWO_ACC>(c, sdf_part, rgb_part, small_in);
                    } else if (g == 16) {
                        epi_bar_sync();                                   // embedding-branch gradients (written at g == 12 by other threads)
                        epi_gemm<K_BWD0, FULL, TWO_ACC>(c, sdf_part, rgb_part, small_in);
                        epi_bar_sync();                                   // all 39 d sdf/d emb entries complete
                    } else if (g == 17) {
#pragma unroll
                        for (int j = 0; j < 36; ++j) small_in[j] = j < c.sdim ? c.misc[r * 80 + 40 + j] * ACT_SCALE : 0.f;
                        epi_gemm<K_RAD0, FULL, TWO_ACC>(c, sdf_part, rgb_part, small_in);
                    } else if (g == 20) {
                        epi_gemm<K_RAD3, FULL, TWO_ACC>(c, sdf_part, rgb_part, small_in);
                    } else {
                        epi_gemm<K_RAD, FULL, TWO_ACC>(c, sdf_part, rgb_part, small_in);
                    }
                }
                // ---- per-GEMM tails ------------------------------------------------------------------------
                if (g == 7) {
                    // fwd layer 7 stored h8 x16: undo in the head.  sdf = <h8, W8[0]> + b8[0]
                    S.PART[cq * TC_TM + r] = sdf_part * (1.f / ACT_SCALE); sdf_part = 0.f;
                    epi_bar_sync();
                    if (cq == 0) {
                        float sdf = S.PART[r] + S.PART[TC_TM + r] + S.PART[2 * TC_TM + r] + S.PART[3 * TC_TM + r] + __ldg(pk + L.b8_sdf);
                        if (job.apply_bg) {
                            const float x0 = S.X[r], x1 = S.X[TC_TM + r], x2 = S.X[2 * TC_TM + r];
                            const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x1, x1)), __fmul_rn(x2, x2)));
                            sdf = fminf(sdf, job.bound_r - nrm);
                        }
                        if (S.OIDX[r] >= 0 && job.sdf) job.sdf[S.OIDX[r]] = sdf;
                    }
                }
                if (FULL && g == 16) {
                    if (cq == 0) {
                        // nabla and the fp32 small radiance inputs
                        const float xs[3] = {S.X[r], S.X[TC_TM + r], S.X[2 * TC_TM + r]};
                        float nb[3];
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) {
                            float n = c.misc[r * 80 + cc];
#pragma unroll
                            for (int f = 0; f < 6; ++f) {
                                const float fr = (float)(1 << f);
                                float sn, cs; sincosf(__fmul_rn(xs[cc], fr), &sn, &cs);
                                n += fr * (c.misc[r * 80 + 3 + 6 * f + cc] * cs - c.misc[r * 80 + 6 + 6 * f + cc] * sn);
                            }
                            nb[cc] = n;
                        }
                        const long long oo = S.OIDX[r];
                        if (oo >= 0 && job.nab) { job.nab[oo * 3] = nb[0]; job.nab[oo * 3 + 1] = nb[1]; job.nab[oo * 3 + 2] = nb[2]; }
                        if (has_rad) {
                            float* sm = c.misc + r * 80 + 40;
                            int qn = 0;
                            for (int cc = 0; cc < 3; ++cc) sm[qn++] = xs[cc];
                            const float vs[3] = {S.V[r], S.V[TC_TM + r], S.V[2 * TC_TM + r]};
                            for (int cc = 0; cc < 3; ++cc) sm[qn++] = vs[cc];
                            for (int f = 0; f < job.multires_view; ++f) {
                                float sn[3], cs[3];
                                for (int cc = 0; cc < 3; ++cc) sincosf(__fmul_rn(vs[cc], (float)(1 << f)), &sn[cc], &cs[cc]);
                                for (int cc = 0; cc < 3; ++cc) sm[qn++] = sn[cc];
                                for (int cc = 0; cc < 3; ++cc) sm[qn++] = cs[cc];
                            }
                            for (int cc = 0; cc < 3; ++cc) sm[qn++] = nb[cc];
                        }
                    }
                    if (has_rad) {
                        // A <- geometry feature (x16) for radiance layer 0
#pragma unroll 1
                        for (int c16 = 0; c16 < 4; ++c16) {
                            const int col0 = c16 * 64 + cq * 16;
                            float h[16];
#pragma unroll
                            for (int j4 = 0; j4 < 4; ++j4) {
                                const float4 f4 = *plane_ptr(sp, 8, (col0 >> 2) + j4, r);
                                h[4 * j4] = f4.x * ACT_SCALE; h[4 * j4 + 1] = f4.y * ACT_SCALE; h[4 * j4 + 2] = f4.z * ACT_SCALE; h[4 * j4 + 3] = f4.w * ACT_SCALE;
                            }
                            store8s(c.a1, c.a2, c.row_off, c.sw, col0, h);
                            store8s(c.a1, c.a2, c.row_off, c.sw, col0 + 8, h + 8);
                        }
                        epi_bar_sync();                       // small inputs written by cq == 0 threads are read by all four next
                    }
                }
                if (FULL && g == 20) {
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) { S.PART[(cq * 3 + cc) * TC_TM + r] = rgb_part[cc]; rgb_part[cc] = 0.f; }
                    epi_bar_sync();
                    if (cq == 0 && S.OIDX[r] >= 0) {
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) {
                            // radiance layer 3 stored relu x16: undo in the head
                            const float z = (S.PART[cc * TC_TM + r] + S.PART[(3 + cc) * TC_TM + r] + S.PART[(6 + cc) * TC_TM + r] + S.PART[(9 + cc) * TC_TM + r])
                                            * (1.f / ACT_SCALE) + __ldg(pk + L.rad_b4 + cc);
                            job.rad[S.OIDX[r] * 3 + cc] = __fdiv_rn(1.f, 1.f + expf(-z));
                        }
                    }
                }
                if (g + 1 < prog.n_gemm) {
                    if (FULL && g == 16) for (int k = 0; k < 4; ++k) signal_kb(c.kb_bar, k, lane);    // A was (re)written by the tail above
                } else {
                    tc_fence_before();
                    epi_bar_sync();                           // X / OIDX / PART are rewritten by the next tile's input stage
                }
            }
        }
        if (job.dbg && blockIdx.x == 0 && tid == 64) { job.dbg[3] = clock64() - t_e0; job.dbg[4] = t_d; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_d, TWO_ACC ? 512 : 256); }
    (void)0;
}

// ------------------------------------------------------------------------------------------------
// weight packing for the tensor-core path: fp32 plane P[r][c] (r = contraction index, c = output column; exactly the planes
// the fp32 path uses) -> stages [nh][kb][split][128 rows c][64 k r] in the UMMA K-major SWIZZLE_128B smem image, scaled by
// 2^shift so that max|W| lands in [256, 512), split into hi/lo fp16.
// ------------------------------------------------------------------------------------------------
__global__ void plane_absmax_kernel(const float* __restrict__ pk, const size_t* __restrict__ offs, const int* __restrict__ rows, float* __restrict__ out) {
    const int g = blockIdx.x;
    const float* p = pk + offs[g];
    const int n = rows[g] * 256;
    float m = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(p[i]));
    __shared__ float red[256];
    red[threadIdx.x] = m; __syncthreads();
    for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + s]); __syncthreads(); }
    if (threadIdx.x == 0) out[g] = red[0];
}

__global__ void tc_pack_kernel(const float* __restrict__ pk, const size_t* __restrict__ offs, const int* __restrict__ rows,
                               const int* __restrict__ n_kb_, const int* __restrict__ n_nh_, const unsigned* __restrict__ stage0,
                               const float* __restrict__ absmax, unsigned char* __restrict__ wtc, float* __restrict__ unscale) {
    const int g = blockIdx.y;
    const float* p = pk + offs[g];
    const int R = rows[g], n_kb = n_kb_[g], n_nh = n_nh_[g];
    const float mx = absmax[g];
    int shift = 0;
    if (mx > 0.f) { int e; frexpf(mx, &e); shift = 9 - e; }           // mx * 2^shift in [256, 512)
    const float sc = ldexpf(1.f, shift);
    if (blockIdx.x == 0 && threadIdx.x == 0) unscale[g] = ldexpf(1.f, -shift) / ACT_SCALE;
    const int n_el = n_nh * n_kb * 128 * 64;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_el; idx += gridDim.x * blockDim.x) {
        const int kk = idx & 63, nl = (idx >> 6) & 127, blk = idx >> 13;         // blk = nh * n_kb + kb
        const int kb = blk % n_kb, nh = blk / n_kb;
        const int rr = kb * 64 + kk, c = nh * 128 + nl;
        const float w = rr < R ? p[(size_t)rr * 256 + c] * sc : 0.f;
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));
        const size_t st = (size_t)stage0[g] + (size_t)blk * 2;
        const unsigned off = (unsigned)((nl >> 3) * 1024 + (nl & 7) * 128 + (((kk >> 3) ^ (nl & 7)) << 4) + ((kk & 7) << 1));
        *reinterpret_cast<__half*>(wtc + st * STAGE_BYTES + off) = hi;
        *reinterpret_cast<__half*>(wtc + (st + 1) * STAGE_BYTES + off) = lo;
    }
}

// program: 0..7 fwd | 8 feat | 9..15 bwd 7..1 | 16 bwd 0 | 17..20 radiance
constexpr int TC_N_PLANES = 21;
struct TcPackLayout { size_t wtc_off, unscale_off, meta_off, total; unsigned stage0[TC_N_PLANES]; int n_kb[TC_N_PLANES], n_nh[TC_N_PLANES]; unsigned n_stages; };

TcPackLayout tc_pack_layout(size_t f32_bytes) {
    TcPackLayout T;
    unsigned st = 0;
    for (int g = 0; g < TC_N_PLANES; ++g) {
        T.n_kb[g] = (g == 0) ? 1 : 4;
        T.n_nh[g] = (g == 16) ? 1 : 2;
        T.stage0[g] = st; st += (unsigned)(T.n_kb[g] * T.n_nh[g] * 2);
    }
    T.n_stages = st;
    T.wtc_off = (f32_bytes + 1023) & ~(size_t)1023;
    T.unscale_off = T.wtc_off + (size_t)st * STAGE_BYTES;
    T.meta_off = T.unscale_off + 256;
    T.total = T.meta_off + 4096;
    return T;
}

int tc_pack(const float* pk_f32, const PackF32& L, unsigned char* base, const TcPackLayout& T, cudaStream_t stream) {
    size_t offs[TC_N_PLANES]; int rows[TC_N_PLANES];
    for (int i = 0; i < 8; ++i) { offs[i] = L.sdf_wt[i]; rows[i] = i == 0 ? EMB_PAD : 256; }
    offs[8] = L.w8t_feat; rows[8] = 256;
    for (int i = 0; i < 7; ++i) { offs[9 + i] = L.sdf_w[7 - i]; rows[9 + i] = 256; }
    offs[16] = L.sdf_w[0]; rows[16] = 256;
    for (int i = 0; i < 4; ++i) { offs[17 + i] = L.rad_wt[i]; rows[17 + i] = 256; }
    unsigned char* meta = base + T.meta_off;
    size_t* d_offs = (size_t*)meta; int* d_rows = (int*)(meta + 256); int* d_kb = (int*)(meta + 512); int* d_nh = (int*)(meta + 768);
    unsigned* d_st = (unsigned*)(meta + 1024); float* d_absmax = (float*)(meta + 1280);
    NA_TRY(upload_small(d_offs, offs, sizeof(offs), stream));
    NA_TRY(upload_small(d_rows, rows, sizeof(rows), stream));
    NA_TRY(upload_small(d_kb, T.n_kb, sizeof(T.n_kb), stream));
    NA_TRY(upload_small(d_nh, T.n_nh, sizeof(T.n_nh), stream));
    NA_TRY(upload_small(d_st, T.stage0, sizeof(T.stage0), stream));
    plane_absmax_kernel<<<TC_N_PLANES, 256, 0, stream>>>(pk_f32, d_offs, d_rows, d_absmax);
    NA_CHECK_LAUNCH();
    tc_pack_kernel<<<dim3(32, TC_N_PLANES), 256, 0, stream>>>(pk_f32, d_offs, d_rows, d_kb, d_nh, d_st, d_absmax, base + T.wtc_off,
                                                                (float*)(base + T.unscale_off));
    NA_CHECK_LAUNCH();
    return NA_OK;
}

size_t mlp_tc_scratch_bytes(int grid) { return (size_t)grid * 10 * 256 * TC_TM * sizeof(float); }

long long* g_tc_dbg = nullptr;

int launch_mlp_tc(const EvalJob& job_, const unsigned char* packed_base, size_t f32_bytes, const PackF32& L, float* scratch,
                  size_t scratch_bytes, cudaStream_t stream) {
    EvalJob job = job_; job.dbg = g_tc_dbg;
    static bool attr_done[64] = {false};
    static const bool two = []{ const char* e = getenv("NA_TC_TWO_ACC"); return e ? atoi(e) != 0 : (NA_TC_TWO_ACC != 0); }();
    const size_t smem = sizeof(TcSmem) + 1024;
    if (first_on_device(attr_done)) {
        NA_TRY(check_cuda(cudaFuncSetAttribute(mlp_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        NA_TRY(check_cuda(cudaFuncSetAttribute(mlp_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        NA_TRY(check_cuda(cudaFuncSetAttribute(mlp_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        NA_TRY(check_cuda(cudaFuncSetAttribute(mlp_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    }
    const long long total = job.x ? job.m : (long long)job.n_rows * job.P;
    if (total <= 0) return NA_OK;
    const TcPackLayout T = tc_pack_layout(f32_bytes);
    TcProgram prog; prog.n_gemm = 0;
    auto add = [&](int g) { TcGemm t; t.w_stage0 = T.stage0[g]; t.n_kb = (unsigned char)T.n_kb[g]; t.n_nh = (unsigned char)T.n_nh[g]; t.pad0 = t.pad1 = 0; prog.g[prog.n_gemm++] = t; };
    const int last = !job.want_full ? (job.feat ? 8 : 7) : (job.rad ? 20 : 16);
    for (int g = 0; g <= last; ++g) add(g);
    long long tiles = (total + TC_TM - 1) / TC_TM;
    int grid = (int)(tiles < (long long)num_sms() ? tiles : (long long)num_sms());
    if (scratch_bytes < mlp_tc_scratch_bytes(grid)) return NA_ERR_WORKSPACE;
    const float* pkf = (const float*)packed_base; const unsigned char* wtc = packed_base + T.wtc_off;
    const float* usc = (const float*)(packed_base + T.unscale_off);
    if (job.want_full) {
        if (two) mlp_tc_kernel<true, true><<<grid, TC_THREADS, smem, stream>>>(job, pkf, L, wtc, usc, prog, scratch);
        else     mlp_tc_kernel<true, false><<<grid, TC_THREADS, smem, stream>>>(job, pkf, L, wtc, usc, prog, scratch);
    } else {
        if (two) mlp_tc_kernel<false, true><<<grid, TC_THREADS, smem, stream>>>(job, pkf, L, wtc, usc, prog, scratch);
        else     mlp_tc_kernel<false, false><<<grid, TC_THREADS, smem, stream>>>(job, pkf, L, wtc, usc, prog, scratch);
    }
    NA_CHECK_LAUNCH();
    return NA_OK;
}

int preload_mlp_tc() {
    NA_PRELOAD((mlp_tc_kernel<false, false>));
    NA_PRELOAD((mlp_tc_kernel<true, false>));
    NA_PRELOAD((mlp_tc_kernel<false, true>));
    NA_PRELOAD((mlp_tc_kernel<true, true>));
    NA_PRELOAD(plane_absmax_kernel);
    NA_PRELOAD(tc_pack_kernel);
    return NA_OK;
}

}  // namespace na
