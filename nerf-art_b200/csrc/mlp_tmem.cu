// Fused per-sample network kernel, tcgen05 path with TMEM-resident activations (NA_PRECISION_TC / NA_PRECISION_TC_MIXED).
// sm_100a only.
//
// One persistent CTA per SM; a tile is 128 samples = the 128 TMEM lanes.  The activations of a tile never leave tensor
// memory: layer g reads its A operand [128 x K] from TMEM region R[g&1] (tcgen05.mma with the A operand in TMEM), accumulates
// D [128 x 256] fp32 into region R[(g+1)&1], and the epilogue converts D *in place* into the next layer's A operand: each
// thread owns one TMEM lane (= sample row), loads 16 fp32 columns (tcgen05.ld), applies bias / softplus / relu, splits the
// result into two fp16 terms v = hi + lo (22 significant bits) and stores 8 packed hi words + 8 packed lo words back into
// the same 16 columns (tcgen05.st).  K-step s of the next GEMM therefore finds A_hi at column 16 s and A_lo at 16 s + 8.
// Shared memory holds nothing but the weight ring (6 x 32 KB) and the small per-tile tables, so the whole L2 -> SMEM weight
// stream runs six stages ahead, and the MMA operand fetch from shared memory is B only (256-wide MMAs, 64 B/clk).
//
// Products: hi*hi + lo*hi + hi*lo (fp32-level; the dropped lo*lo is 2^-22 relative) for the GEMMs of the SDF forward
// pass; in NA_PRECISION_TC_MIXED the feature head, the reverse sweep and the radiance layers use hi*hi only
// (11-bit operands, i.e. TF32-level), which the reference's own tolerance study allows.
//
// Pipeline per GEMM: the 16 epilogue warps finish one 64-wide K-block of the next A operand per pass and signal it
// (mbarrier); the MMA warp issues the next GEMM's K-block as soon as it is signalled, into the *other* TMEM region, so
// only the last K-block's MMAs are exposed after the epilogue.
//
// Warp roles (18 warps): warp 0 lane 0 = weight producer (cp.async.bulk, full/empty mbarrier ring); warp 1 = MMA issuer
// (one elected lane) and TMEM owner; warps 2..17 = epilogue, warp w owns TMEM lanes 32*(w%4).., four warps per lane
// quadrant take 16 of every 64 columns.
#include "common.cuh"
#include <cuda_fp16.h>
#include <cstdlib>
#include <cstring>

namespace na {
namespace tm {

constexpr int TM = 128;
constexpr int THREADS = 576;
constexpr int EPI_THREADS = 512;
#ifndef NA_TM_NS
#define NA_TM_NS 5
#endif
constexpr int NS = NA_TM_NS;                     // weight stages (32 KB each); 5 leaves room for the encoding / small-weight stashes below
constexpr bool STASH = NS <= 5;
constexpr int STAGE_BYTES = 32768;               // 256 rows x 64 fp16
constexpr float ACT_SCALE = 16.f;                // activations are stored x16 (keeps the lo term normal in fp16)
constexpr int MAX_GEMM = 24;
constexpr int N_PLANES = 21;                     // program: 0..7 fwd | 8 feat | 9..15 bwd 7..1 | 16 bwd 0 | 17..20 radiance

// NA_TM_TRACE (diagnostic build): CTA 0 records clock64() stamps of its second tile into job.dbg[16..]:
//   MMA lane:      slot (g*4+kb)*2 + {0: K-block of A ready, 1: its MMAs issued}
//   epilogue lane: slot 192 + (g*4+pass)*3 + {0: D quarter ready, 1: tcgen05.ld done, 2: A stored}
#ifdef NA_TM_TRACE
#define NA_TRACE_M(tr, g, kb, w) do { if (tr) (tr)[((g) * 4 + (kb)) * 2 + (w)] = clock64(); } while (0)
#define NA_TRACE_E(tr, g, ps, w) do { if (tr) (tr)[192 + ((g) * 4 + (ps)) * 3 + (w)] = clock64(); } while (0)
#else
#define NA_TRACE_M(tr, g, kb, w) do { } while (0)
#define NA_TRACE_E(tr, g, ps, w) do { } while (0)
#endif

struct Gemm { unsigned w_off; unsigned stage_bytes; unsigned char n_kb, prods, n64, pad; };
struct Program { int n_gemm; int nsplit; Gemm g[MAX_GEMM]; };      // nsplit: N-parts (2 or 4) the last K-block of a GEMM is issued in

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
// D[tmem] (+)= A[tmem] * B[smem]^T : A is 128 lanes x 8 columns (16 fp16, two per 32-bit column, low half = even k)
__device__ __forceinline__ void umma_f16_ts(unsigned d_tmem, unsigned a_tmem, unsigned long long b_desc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 16 columns: thread i of the warp <-> lane (base+i), register j <-> column c+j
__device__ __forceinline__ void tmem_ld16(unsigned taddr, unsigned (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(unsigned taddr, const unsigned (&v)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
        :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
    unsigned pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr)); return v;
}
// drop a 128-byte line of per-CTA scratch from L2 without writing it back (its last reader is done with it)
__device__ __forceinline__ void discard_l2(const void* p) { asm volatile("discard.global.L2 [%0], 128;" :: "l"(p) : "memory"); }
__device__ __forceinline__ float lds32(unsigned addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }

// K-major, SWIZZLE_128B UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (1024 B between 8-row groups) |
//   [46,48) version=1 | [61,64) layout=2 (SWIZZLE_128B)
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr) {
    return (unsigned long long)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D=f32 (bit 4), A=B=f16 (0), both K-major, N>>3 @17, M>>4 @24 (M = 128)
constexpr unsigned IDESC_N256 = (1u << 4) | (32u << 17) | (8u << 24);
constexpr unsigned IDESC_N64 = (1u << 4) | (8u << 17) | (8u << 24);
constexpr unsigned IDESC_N128 = (1u << 4) | (16u << 17) | (8u << 24);

constexpr int N_BIAS_ROWS = 13;                  // 0..7 sdf fwd (x ACT_SCALE) | 8 feature (raw) | 9..12 radiance (x ACT_SCALE)

struct __align__(1024) Smem {
    unsigned char Wst[NS * STAGE_BYTES];
    unsigned long long full_bar[NS], empty_bar[NS], d_ready[4], kb_ready[4];   // d_ready[q]: N-quarter q of D is complete; kb_ready[k]: K-block k of the next A operand is in TMEM
    unsigned tmem_base;
    __align__(16) float BIAS[N_BIAS_ROWS * 256];
    __align__(16) float W8[256];                       // row 0 of SDF layer 8 (the sdf head)
    __align__(16) float W4[3 * 256];                   // radiance output layer
    float X[3 * TM];
    float V[3 * TM];
    float PART[4 * 3 * TM];                            // per column-quarter partial sums of the narrow heads
    long long OIDX[TM];
    // STASH: the tile's encoding (x ACT_SCALE; entry k of row r at k*TM + r), reused by the skip connection (layer 3) and the
    // closed-form nabla instead of re-evaluating sincosf; and the 9 small-input rows of radiance layer 0 (VolSDF: x | view | nabla)
    float EMBS[STASH ? EMB * TM : 1];
    __align__(16) float RADW[STASH ? 9 * 256 : 4];
};

// per-CTA global scratch (full mode): 8 softplus' planes (16-bit codes), the geometry feature (fp32), misc rows
constexpr size_t DH_BYTES = (size_t)8 * 64 * TM * 8;          // plane p, column quad k4, row r -> uint2 at (p*64 + k4)*128 + r
constexpr size_t FEAT_BYTES = (size_t)64 * TM * 16;           // float4 at k4*128 + r
constexpr size_t MISC_BYTES = (size_t)80 * TM * 4;            // float at j*128 + r : d sdf/d emb (39) @0 | small radiance inputs (<=33) @40
constexpr size_t SCRATCH_BYTES = DH_BYTES + FEAT_BYTES + MISC_BYTES;

enum EpiKind { K_FWD, K_FWD3, K_FWD7, K_FEAT, K_BWD, K_BWD4, K_BWD0, K_RAD0, K_RAD, K_RAD3 };

struct EpiCtx {
    Smem* S; uint2* dh; float4* featp; float* misc; const float* pk; const PackF32* L; const EvalJob* job;
    unsigned t_lane; int r, cq, g; float us; int sdim;
    unsigned bias_s, w8_s, w4_s, radw_s, kb_bar, d_bar;
    int signal, need_lo, lane;
    float* st_row;                  // ST: st_wide + (flat sample) * 256 of this thread's row, nullptr for padding rows
    size_t st_plane;                // ST: floats per plane
    unsigned d_phase; long long* t_wait; long long* trace;
};

// this warp's part of K-block kb of the next A operand is in TMEM (and its part of D columns [64kb, 64kb+64) is consumed)
__device__ __forceinline__ void signal_kb(unsigned kb_bar, int kb, int lane) {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(kb_bar + 8u * (unsigned)kb);
}

// 16 fp32 values (already x ACT_SCALE) -> 8 packed hi words + 8 packed lo words, stored over the 16 columns at taddr
// (need_lo == 0: the consumer GEMM uses the hi*hi product only, the lo words are neither computed nor stored)
__device__ __forceinline__ void store_a16(unsigned taddr, const float (&o)[16], int need_lo) {
    unsigned hi[8];
    __half2 h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { h[i] = __floats2half2_rn(o[2 * i], o[2 * i + 1]); hi[i] = *reinterpret_cast<const unsigned*>(&h[i]); }
    tmem_st8(taddr, hi);
    if (need_lo) {
        unsigned lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 back = __half22float2(h[i]);
            const __half2 l = __floats2half2_rn(o[2 * i] - back.x, o[2 * i + 1] - back.y);
            lo[i] = *reinterpret_cast<const unsigned*>(&l);
        }
        tmem_st8(taddr + 8, lo);
    }
    tmem_wait_st();
}

// softplus'(z) = sigmoid(100 z) as a 16-bit code: bit 15 = (z >= 0), low 15 bits = round(t * 32767.49), t = exp(-|100 z|);
// decoded as r = 1 / (1 + code/32768), z >= 0 ? r : 1 - r   (absolute error <= 2e-5).  The decode is two bit operations, one
// MUFU.RCP and one add: the 15 bits are dropped straight into the mantissa of a float in [1, 2), which *is* 1 + t.
__device__ __forceinline__ unsigned dh_code(float z16, float t) {
    // round(t * 32767.49) through the 2^23 magic number (FMA pipe; F2I would go to the XU pipe the softplus already saturates)
    return ((~__float_as_uint(z16) >> 16) & 0x8000u) | (__float_as_uint(fmaf(t, 32767.49f, 8388608.f)) & 0x7fffu);
}
__device__ __forceinline__ float dh_decode_lo(unsigned w) {            // code in bits [0,16)
    const float ru = rcp_approx(__uint_as_float(((w & 0x7fffu) << 8) | 0x3f800000u));
    return (w & 0x8000u) ? ru : 1.f - ru;
}
__device__ __forceinline__ float dh_decode_hi(unsigned w) {            // code in bits [16,32)
    const float ru = rcp_approx(__uint_as_float(((w >> 8) & 0x7fff00u) | 0x3f800000u));
    return (w & 0x80000000u) ? ru : 1.f - ru;
}
__device__ __forceinline__ void dh_decode4(const uint2 q, float (&d)[4]) {
    d[0] = dh_decode_lo(q.x); d[1] = dh_decode_hi(q.x); d[2] = dh_decode_lo(q.y); d[3] = dh_decode_hi(q.y);
}

// encoding entries k in [K_LO, K_LO + 16) of [x, sin(2^f x), cos(2^f x)]_f (models/base.py:46-64), x ACT_SCALE; zero outside [0, 39)
template <int K_LO>
__device__ __forceinline__ void emb_range(const float (&xs)[3], float (&e)[16]) {
    constexpr int k_lo = K_LO, k_hi = K_LO + 16;
    float sn[18], cs[18];
#pragma unroll
    for (int pi = 0; pi < 18; ++pi) {
        const int f = pi / 3, cc = pi % 3;
        const int ks = 3 + 6 * f + cc, kc = ks + 3;
        const bool need = (ks >= k_lo && ks < k_hi) || (kc >= k_lo && kc < k_hi);
        sn[pi] = 0.f; cs[pi] = 0.f;
        if (need) sincosf(__fmul_rn(xs[cc], (float)(1 << f)), &sn[pi], &cs[pi]);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int k = k_lo + j;
        float v = 0.f;
        if (k >= 0 && k < 3) v = xs[k];
        else if (k >= 3 && k < EMB) { const int f = (k - 3) / 6, rem = (k - 3) % 6; v = rem < 3 ? sn[f * 3 + rem] : cs[f * 3 + rem - 3]; }
        e[j] = v * ACT_SCALE;
    }
}

// ST: 16 consecutive columns of this thread's row -> stash plane `plane` (values are stored x `scale`)
__device__ __forceinline__ void stash16(const EpiCtx& c, int plane, int col0, const float (&o)[16], float scale) {
    if (!c.st_row) return;
    float4* dst = reinterpret_cast<float4*>(c.st_row + (size_t)plane * c.st_plane + col0);
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) dst[j4] = make_float4(o[4 * j4] * scale, o[4 * j4 + 1] * scale, o[4 * j4 + 2] * scale, o[4 * j4 + 3] * scale);
}

// one GEMM's epilogue for this thread's row and its 64 columns (4 passes of 16)
template <int KIND, bool FULL, bool ST>
__device__ __forceinline__ void epi_gemm(const EpiCtx& c, const unsigned t_d, float& sdf_part, float (&rgb_part)[3], const float (&small_in)[36]) {
    Smem& S = *c.S;
    const int r = c.r;
    const float us = c.us, us16 = c.us * ACT_SCALE;
    const unsigned bias = c.bias_s + (unsigned)(KIND == K_FEAT ? 8 : (KIND >= K_RAD0 ? 9 + (c.g - 17) : c.g)) * 1024u;
    constexpr bool USES_DH = FULL && (KIND == K_FEAT || KIND == K_BWD || KIND == K_BWD4);
    constexpr int N_PASS = KIND == K_BWD0 ? 1 : 4;                  // reverse GEMM 0: only 39 useful columns, all in pass 0
    // softplus' plane this epilogue multiplies by: feature head (g = 8) -> layer 7, reverse GEMM g = 9..15 -> layer 15 - g
    const uint2* dhp = c.dh + (size_t)((15 - c.g) * 64) * TM + r;
    // softplus' codes of this thread's row: thread-private scratch (written by the same thread in the forward GEMM), so the
    // loads can run one pass ahead -- pass 0's are issued before the wait for D, pass c+1's before pass c's arithmetic --
    // which takes the L2 / DRAM latency (the MMA warp was starved 46 % of a full-mode tile, profiles/r1v) off the critical path
    uint2 qn[4];
    if (USES_DH) {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) qn[j4] = dhp[(size_t)(c.cq * 4 + j4) * TM];
    }
#pragma unroll 1
    for (int c16 = 0; c16 < N_PASS; ++c16) {
        // thread = (row, column quarter cq): in pass c16 it owns columns 64*c16 + 16*cq .. +16, i.e. every pass completes one
        // 64-wide K-block of the next layer's A operand across the 16 epilogue warps
        const int col0 = c16 * 64 + c.cq * 16;
        uint2 q[4];
        if (USES_DH) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) q[j4] = qn[j4];
            if (c16 + 1 < N_PASS) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) qn[j4] = dhp[(size_t)(((col0 + 64) >> 2) + j4) * TM];
            }
        }
        {                                                            // pass c16 reads N-quarter c16 of D
            const long long t0 = clock64();
            mbar_wait(c.d_bar + 8u * (unsigned)c16, c.d_phase);
            *c.t_wait += clock64() - t0;
            tc_fence_after();
            NA_TRACE_E(c.trace, c.g, c16, 0);
        }
        float acc[16];
        {
            unsigned v[16];
            tmem_ld16(t_d + col0, v);
            tmem_wait_ld();
            NA_TRACE_E(c.trace, c.g, c16, 1);
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
                    unsigned cd[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int j = 4 * j4 + i;
                        cd[i] = dh_code(z16[j], t[j]);
                        if (KIND == K_FWD3 && col0 + j >= SKIP_H) cd[i] = 0u;          // decodes to 0
                    }
                    c.dh[(size_t)(c.g * 64 + (col0 >> 2) + j4) * TM + r] = make_uint2(cd[0] | (cd[1] << 16), cd[2] | (cd[3] << 16));
                }
            }
            if (KIND == K_FWD3 && c16 == 3 && c.cq >= 1) {
                // skip connection columns (k >= 217): h = emb[k - 217] (x16); this thread's 16 columns are encoding entries
                // 16 cq - 25 .. + 16
                const int e0 = 16 * c.cq - 25;
                if (STASH) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (e0 + j >= 0) o[j] = S.EMBS[(e0 + j) * TM + r];
                } else {
                    // evaluated with one sincosf per (frequency, coordinate) pair that falls in the range
                    const float xs[3] = {S.X[r], S.X[TM + r], S.X[2 * TM + r]};
                    float e[16];
                    if (c.cq == 1) emb_range<-9>(xs, e);
                    else if (c.cq == 2) emb_range<7>(xs, e);
                    else emb_range<23>(xs, e);
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (e0 + j >= 0) o[j] = e[j];
                }
            }
            if (ST) {
                stash16(c, ST_IN + c.g, col0, o, 1.f / ACT_SCALE);
                float sv[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float ru = rcp_approx(1.f + t[j]);
                    sv[j] = z16[j] >= 0.f ? ru : 1.f - ru;
                    if (KIND == K_FWD3 && col0 + j >= SKIP_H) sv[j] = 0.f;
                }
                stash16(c, ST_S + c.g, col0, sv, 1.f);
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
                if (FULL) c.featp[(size_t)((col0 >> 2) + j4) * TM + r] = f4;
                if (ST && c.st_row) reinterpret_cast<float4*>(c.st_row + (size_t)ST_FEAT * c.st_plane + col0)[j4] = f4;
                if (c.job->feat && S.OIDX[r] >= 0) *(reinterpret_cast<float4*>(c.job->feat + S.OIDX[r] * 256 + col0) + j4) = f4;
                if (FULL) {
                    // next A: d sdf / d z7 = W8[0,:] * softplus'(z7)
                    float d4[4];
                    dh_decode4(q[j4], d4);
                    const float4 w4 = lds128(c.w8_s + (unsigned)(col0 + 4 * j4) * 4u);
                    o[4 * j4] = w4.x * d4[0] * ACT_SCALE; o[4 * j4 + 1] = w4.y * d4[1] * ACT_SCALE;
                    o[4 * j4 + 2] = w4.z * d4[2] * ACT_SCALE; o[4 * j4 + 3] = w4.w * d4[3] * ACT_SCALE;
                }
            }
            if (ST && FULL) stash16(c, ST_G + 7, col0, o, 1.f / ACT_SCALE);
        } else if (KIND == K_BWD || KIND == K_BWD4) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                float dd[4];
                dh_decode4(q[j4], dd);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = col0 + 4 * j4 + i;
                    if (KIND == K_BWD4 && k >= SKIP_H) c.misc[(k - SKIP_H) * TM + r] = acc[4 * j4 + i] * us;      // embedding branch of the skip
                    o[4 * j4 + i] = acc[4 * j4 + i] * us16 * dd[i];
                }
            }
            if (ST) stash16(c, ST_G + 15 - c.g, col0, o, 1.f / ACT_SCALE);
        } else if (KIND == K_BWD0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) { const int k = col0 + j; if (k < EMB) c.misc[k * TM + r] += acc[j] * us; }
        } else {
            // radiance hidden layers: relu(16 z)
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 b4 = lds128(bias + (unsigned)(col0 + 4 * j4) * 4u);
                float z[4] = {fmaf(acc[4 * j4], us16, b4.x), fmaf(acc[4 * j4 + 1], us16, b4.y),
                              fmaf(acc[4 * j4 + 2], us16, b4.z), fmaf(acc[4 * j4 + 3], us16, b4.w)};
                if (KIND == K_RAD0) {
                    // small inputs [x | embed(view) | nabla] (x16) in fp32: rows 256.. of the packed layer-0 plane
                    const float4* wsm = reinterpret_cast<const float4*>(c.pk + c.L->rad_wt[0] + (size_t)256 * 256 + col0 + 4 * j4);
                    if (c.sdim == 9) {
#pragma unroll
                        for (int j = 0; j < 9; ++j) {
                            const float4 w = STASH ? lds128(c.radw_s + (unsigned)(j * 256 + col0 + 4 * j4) * 4u) : __ldg(wsm + j * 64);
                            z[0] = fmaf(small_in[j], w.x, z[0]); z[1] = fmaf(small_in[j], w.y, z[1]);
                            z[2] = fmaf(small_in[j], w.z, z[2]); z[3] = fmaf(small_in[j], w.w, z[3]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 33; ++j) {
                            const float4 w = __ldg(wsm + j * 64);
                            z[0] = fmaf(small_in[j], w.x, z[0]); z[1] = fmaf(small_in[j], w.y, z[1]);
                            z[2] = fmaf(small_in[j], w.z, z[2]); z[3] = fmaf(small_in[j], w.w, z[3]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) o[4 * j4 + i] = fmaxf(z[i], 0.f);
            }
            if (ST) stash16(c, ST_YS + c.g - 17, col0, o, 1.f / ACT_SCALE);
            if (KIND == K_RAD3) {
#pragma unroll
                for (int cc = 0; cc < 3; ++cc)
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 w = lds128(c.w4_s + (unsigned)(cc * 256 + col0 + 4 * j4) * 4u);
                        rgb_part[cc] = fmaf(o[4 * j4], w.x, rgb_part[cc]); rgb_part[cc] = fmaf(o[4 * j4 + 1], w.y, rgb_part[cc]);
                        rgb_part[cc] = fmaf(o[4 * j4 + 2], w.z, rgb_part[cc]); rgb_part[cc] = fmaf(o[4 * j4 + 3], w.w, rgb_part[cc]);
                    }
            }
        }
        const bool store = !(KIND == K_BWD0 || KIND == K_RAD3 || (KIND == K_FWD7 && !FULL && !c.job->feat) || (KIND == K_FEAT && !FULL));
        if (store) store_a16(t_d + col0, o, c.need_lo);
        NA_TRACE_E(c.trace, c.g, c16, 2);
        if (USES_DH && (c.lane & 15) == 0) {
            // the 16 lanes' codes of this pass share one line per column quad; they are dead now: keep them out of DRAM
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) discard_l2(dhp + (size_t)((col0 >> 2) + j4) * TM);
        }
        if (KIND != K_BWD0 && c.signal) signal_kb(c.kb_bar, c16, c.lane);
    }
    for (int k = N_PASS; k < 4; ++k) mbar_wait(c.d_bar + 8u * (unsigned)k, c.d_phase);      // keep the other barriers' phases in step
    if (N_PASS < 4) tc_fence_after();
}

template <bool FULL, bool ST>
__global__ void __launch_bounds__(THREADS, 1)
mlp_tmem_kernel(const EvalJob job, const float* __restrict__ pk, const PackF32 L, const unsigned char* __restrict__ wimg,
                const float* __restrict__ unscale, const Program prog, unsigned char* __restrict__ scratch) {
    extern __shared__ unsigned char smem_raw_[];
    Smem& S = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw_) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool explicit_pts = job.x != nullptr;
    const long long total = explicit_pts ? job.m
                          : (long long)(job.n_rows_dev ? min(*job.n_rows_dev, job.n_rows) : job.n_rows) * job.P;
    const long long n_tiles = (total + TM - 1) / TM;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(smem_u32(&S.full_bar[s]), 1); mbar_init(smem_u32(&S.empty_bar[s]), 1); }
        for (int k = 0; k < 4; ++k) mbar_init(smem_u32(&S.d_ready[k]), 1);
        for (int k = 0; k < 4; ++k) mbar_init(smem_u32(&S.kb_ready[k]), EPI_THREADS / 32);   // one arrive per epilogue warp
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(&S.tmem_base), 512);
    // bias rows and head weights -> shared memory (hidden-layer biases pre-multiplied by ACT_SCALE)
    for (int i = tid; i < N_BIAS_ROWS * 256; i += THREADS) {
        const int row = i >> 8, k = i & 255;
        float b;
        if (row < 8) b = pk[L.sdf_b[row] + k] * ACT_SCALE;
        else if (row == 8) b = pk[L.b8_feat + k];
        else b = pk[L.rad_b[row - 9] + k] * ACT_SCALE;
        S.BIAS[i] = b;
    }
    for (int i = tid; i < 256; i += THREADS) S.W8[i] = pk[L.w8_sdf + i];
    for (int i = tid; i < 768; i += THREADS) S.W4[i] = pk[L.rad_w4 + i];
    if (STASH && FULL && job.rad && small_dim(job.multires_view) == 9)
        for (int i = tid; i < 9 * 256; i += THREADS) S.RADW[i] = pk[L.rad_wt[0] + (size_t)256 * 256 + i];
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
                    const unsigned char* src = wimg + prog.g[g].w_off;
                    const unsigned sb = prog.g[g].stage_bytes;
                    const int n_kb = prog.g[g].n_kb, n_sp = prog.g[g].prods == 3 ? 2 : 1;
                    for (int kb = 0; kb < n_kb; ++kb)
                        for (int sp = 0; sp < n_sp; ++sp, ++it) {
                            const unsigned slot = it % NS, ph = (it / NS) & 1;
                            mbar_wait(smem_u32(&S.empty_bar[slot]), ph ^ 1);
                            mbar_expect_tx(smem_u32(&S.full_bar[slot]), sb);
                            bulk_g2s(smem_u32(S.Wst + slot * STAGE_BYTES), src + (size_t)(kb * 2 + sp) * sb, sb, smem_u32(&S.full_bar[slot]));
                        }
                }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // the whole warp walks the program (converged control flow, waits included); one elected lane issues tcgen05.mma / commit
        unsigned it = 0, a_phase = 0;
        long long t_a = 0, t_full = 0, t_tot0 = clock64();
        const unsigned wst = smem_u32(S.Wst);
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (int g = 0; g < prog.n_gemm; ++g) {
#ifdef NA_TM_TRACE
                long long* const tr = (job.dbg && blockIdx.x == 0 && tile == (long long)gridDim.x && lane == 0) ? job.dbg + 16 : nullptr;
#endif
                const int n_kb = prog.g[g].n_kb, prods = prog.g[g].prods;
                const unsigned idesc = prog.g[g].n64 ? IDESC_N64 : IDESC_N256;
                const unsigned t_in = tmem_d + (unsigned)(g & 1) * 256u, t_out = tmem_d + (unsigned)((g + 1) & 1) * 256u;
                for (int kb = 0; kb < n_kb; ++kb) {
                    const unsigned slot0 = it % NS, ph0 = (it / NS) & 1;
                    const unsigned slot1 = (it + 1) % NS, ph1 = ((it + 1) / NS) & 1;
                    const unsigned a_hi = t_in + (unsigned)(kb * 4) * 16u;
                    // the weights first (the ring runs K-blocks ahead, so these return at once), then the A operand: the MMAs go out
                    // right behind the epilogue's signal
                    { const long long t0 = clock64();
                      mbar_wait(smem_u32(&S.full_bar[slot0]), ph0);
                      if (prods == 3) mbar_wait(smem_u32(&S.full_bar[slot1]), ph1);
                      t_full += clock64() - t0; }
                    { const long long t0 = clock64(); mbar_wait(smem_u32(&S.kb_ready[kb]), a_phase); t_a += clock64() - t0; }
                    tc_fence_after();
                    NA_TRACE_M(tr, g, kb, 0);
                    if (elect_one()) {
                        const unsigned long long bd0 = umma_desc(wst + slot0 * STAGE_BYTES), bd1 = umma_desc(wst + slot1 * STAGE_BYTES);
                        if (kb + 1 < n_kb || prog.g[g].n64) {
                            // 256-wide MMAs
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) umma_f16_ts(t_out, a_hi + 16u * ks, bd0 + 2 * ks, idesc, (kb | ks) != 0);              // hi * hi
                            if (prods == 3) {
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) umma_f16_ts(t_out, a_hi + 16u * ks + 8u, bd0 + 2 * ks, idesc, 1);                  // lo * hi
                            }
                            umma_commit(smem_u32(&S.empty_bar[slot0]));
                            if (prods == 3) {
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) umma_f16_ts(t_out, a_hi + 16u * ks, bd1 + 2 * ks, idesc, 1);                       // hi * lo
                                umma_commit(smem_u32(&S.empty_bar[slot1]));
                            }
                        } else {
                            // last K-block: issued in N-parts (quarters, or halves with nsplit == 2), each with its own commit, so that
                            // the epilogue's first passes overlap the MMAs of the remaining parts
                            const int nsp = prog.nsplit;
                            const unsigned ncol = 256u / (unsigned)nsp, idn = nsp == 4 ? IDESC_N64 : IDESC_N128;
                            for (int np = 0; np < nsp; ++np) {
                                const unsigned t_dn = t_out + ncol * np;
                                const unsigned long long boff = (unsigned long long)(np * (int)(ncol * 128u / 16u));     // ncol rows x 128 B
                                const unsigned long long b0 = bd0 + boff, b1 = bd1 + boff;
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) umma_f16_ts(t_dn, a_hi + 16u * ks, b0 + 2 * ks, idn, (kb | ks) != 0);             // hi * hi
                                if (prods == 3) {
#pragma unroll
                                    for (int ks = 0; ks < 4; ++ks) umma_f16_ts(t_dn, a_hi + 16u * ks + 8u, b0 + 2 * ks, idn, 1);                 // lo * hi
#pragma unroll
                                    for (int ks = 0; ks < 4; ++ks) umma_f16_ts(t_dn, a_hi + 16u * ks, b1 + 2 * ks, idn, 1);                      // hi * lo
                                }
                                if (nsp == 4) umma_commit(smem_u32(&S.d_ready[np]));
                                else { umma_commit(smem_u32(&S.d_ready[2 * np])); umma_commit(smem_u32(&S.d_ready[2 * np + 1])); }
                            }
                            umma_commit(smem_u32(&S.empty_bar[slot0]));
                            if (prods == 3) umma_commit(smem_u32(&S.empty_bar[slot1]));
                        }
                    }
                    __syncwarp();
                    it += prods == 3 ? 2 : 1;
                    NA_TRACE_M(tr, g, kb, 1);
                }
                for (int kb = n_kb; kb < 4; ++kb) mbar_wait(smem_u32(&S.kb_ready[kb]), a_phase);     // keep the phases in step
                a_phase ^= 1;
                if (prog.g[g].n64) {
                    if (elect_one()) { for (int k = 0; k < 4; ++k) umma_commit(smem_u32(&S.d_ready[k])); }
                    __syncwarp();
                }
            }
        if (job.dbg && blockIdx.x == 0 && lane == 0) { job.dbg[0] = clock64() - t_tot0; job.dbg[1] = t_a; job.dbg[2] = t_full; }
    } else {
        // ================= epilogue warps =================
        const int q = warp & 3, cq = (warp - 2) >> 2;
        const int r = 32 * q + lane;                       // sample row == TMEM lane
        unsigned char* sp = scratch + (size_t)blockIdx.x * SCRATCH_BYTES;
        EpiCtx c;
        c.S = &S; c.dh = reinterpret_cast<uint2*>(sp); c.featp = reinterpret_cast<float4*>(sp + DH_BYTES);
        c.misc = reinterpret_cast<float*>(sp + DH_BYTES + FEAT_BYTES);
        c.pk = pk; c.L = &L; c.job = &job; c.t_lane = tmem_d + ((unsigned)(32 * q) << 16); c.r = r; c.cq = cq;
        c.sdim = small_dim(job.multires_view);
        c.bias_s = smem_u32(S.BIAS); c.w8_s = smem_u32(S.W8); c.w4_s = smem_u32(S.W4); c.radw_s = smem_u32(S.RADW);
        c.kb_bar = smem_u32(&S.kb_ready[0]); c.d_bar = smem_u32(&S.d_ready[0]); c.lane = lane; c.signal = 0; c.need_lo = 1;
        c.d_phase = 0;
        long long t_d = 0, t_e0 = clock64();
        c.t_wait = &t_d; c.trace = nullptr; c.st_row = nullptr; c.st_plane = job.st_mpad * 256;
        const bool has_rad = job.rad != nullptr;

        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#ifdef NA_TM_TRACE
            c.trace = (job.dbg && blockIdx.x == 0 && tile == (long long)gridDim.x && tid == 64) ? job.dbg + 16 : nullptr;
#endif
            // ---- tile inputs: point, encoding (x ACT_SCALE, hi/lo) into K-block 0 of region 0 ---------------------
            {
                const long long w = tile * TM + r;
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
                c.st_row = (ST && w < total) ? job.st_wide + (size_t)w * 256 : nullptr;
                if (cq == 0) {
                    S.OIDX[r] = oidx;
                    S.X[r] = x0; S.X[TM + r] = x1; S.X[2 * TM + r] = x2;
                    S.V[r] = v0; S.V[TM + r] = v1; S.V[2 * TM + r] = v2;
                }
                const float xs[3] = {x0, x1, x2};
                float e[16];
                if (cq == 0) emb_range<0>(xs, e);
                else if (cq == 1) emb_range<16>(xs, e);
                else if (cq == 2) emb_range<32>(xs, e);
                else emb_range<48>(xs, e);
                if (STASH) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (16 * cq + j < EMB) S.EMBS[(16 * cq + j) * TM + r] = e[j];
                }
                store_a16(c.t_lane + (unsigned)(16 * cq), e, 1);
            }
            for (int k = 0; k < 4; ++k) signal_kb(c.kb_bar, k, lane);

            float sdf_part = 0.f, rgb_part[3] = {0.f, 0.f, 0.f};
            float small_in[36];
            for (int g = 0; g < prog.n_gemm; ++g) {
                c.g = g; c.us = unscale[g];                      // us = 2^-(weight shift) / ACT_SCALE
                c.signal = g + 1 < prog.n_gemm;
                c.need_lo = c.signal ? (prog.g[g + 1].prods == 3) : 0;
                const unsigned t_dd = c.t_lane + (unsigned)((g + 1) & 1) * 256u;       // D of this GEMM == A of the next
                // program order: 0..7 fwd | 8 feat | 9..15 bwd 7..1 | 16 bwd 0 | 17..20 radiance
                if (g < 8) {
                    if (g == 3) epi_gemm<K_FWD3, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                    else if (g == 7) epi_gemm<K_FWD7, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                    else epi_gemm<K_FWD, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                } else if (g == 8) {
                    epi_gemm<K_FEAT, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                } else if (FULL) {
                    if (g <= 15) {
                        if (g == 12) epi_gemm<K_BWD4, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                        else epi_gemm<K_BWD, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                    } else if (g == 16) {
                        epi_bar_sync();                                   // embedding-branch gradients (written at g == 12 by other threads)
                        epi_gemm<K_BWD0, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                    } else if (g == 17) {
                        // small radiance inputs [x | embed(view) | nabla] (x16), kept in registers: every thread rebuilds its row's
                        const float xs[3] = {S.X[r], S.X[TM + r], S.X[2 * TM + r]}, vs[3] = {S.V[r], S.V[TM + r], S.V[2 * TM + r]};
                        const float nb[3] = {S.PART[r], S.PART[TM + r], S.PART[2 * TM + r]};
#pragma unroll
                        for (int j = 0; j < 36; ++j) small_in[j] = 0.f;
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) { small_in[cc] = xs[cc] * ACT_SCALE; small_in[3 + cc] = vs[cc] * ACT_SCALE; }
                        if (c.sdim == 9) {
#pragma unroll
                            for (int cc = 0; cc < 3; ++cc) small_in[6 + cc] = nb[cc] * ACT_SCALE;
                        } else {
#pragma unroll
                            for (int f = 0; f < 4; ++f)
#pragma unroll
                                for (int cc = 0; cc < 3; ++cc) {
                                    float sn, cs; sincosf(__fmul_rn(vs[cc], (float)(1 << f)), &sn, &cs);
                                    small_in[6 + 6 * f + cc] = sn * ACT_SCALE; small_in[9 + 6 * f + cc] = cs * ACT_SCALE;
                                }
#pragma unroll
                            for (int cc = 0; cc < 3; ++cc) small_in[30 + cc] = nb[cc] * ACT_SCALE;
                        }
                        if (ST && cq == 0 && c.st_row && job.st_small) {
                            float* srow = job.st_small + (size_t)((c.st_row - job.st_wide) >> 8) * 40;
#pragma unroll
                            for (int j = 0; j < 36; ++j) srow[j] = small_in[j] * (1.f / ACT_SCALE);
#pragma unroll
                            for (int j = 36; j < 40; ++j) srow[j] = 0.f;
                        }
                        epi_gemm<K_RAD0, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                    } else if (g == 20) {
                        epi_gemm<K_RAD3, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                    } else {
                        epi_gemm<K_RAD, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                    }
                }
                c.d_phase ^= 1;
                // ---- per-GEMM tails ------------------------------------------------------------------------
                if (g == 7) {
                    // fwd layer 7 stored h8 x16: undo in the head.  sdf = <h8, W8[0]> + b8[0]
                    S.PART[cq * TM + r] = sdf_part * (1.f / ACT_SCALE); sdf_part = 0.f;
                    epi_bar_sync();
                    if (cq == 0) {
                        float sdf = S.PART[r] + S.PART[TM + r] + S.PART[2 * TM + r] + S.PART[3 * TM + r] + __ldg(pk + L.b8_sdf);
                        if (job.apply_bg) {
                            const float x0 = S.X[r], x1 = S.X[TM + r], x2 = S.X[2 * TM + r];
                            const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x1, x1)), __fmul_rn(x2, x2)));
                            sdf = fminf(sdf, job.bound_r - nrm);
                        }
                        if (S.OIDX[r] >= 0 && job.sdf) job.sdf[S.OIDX[r]] = sdf;
                    }
                }
                if (FULL && g == 16) {
                    if (has_rad) {
                        // A <- geometry feature (x16) for radiance layer 0, written over this thread's own (consumed) columns of D of
                        // GEMM 16; signalled at once, so radiance GEMM 0 runs under the nabla arithmetic below
                        float4 fn[4];
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) fn[j4] = c.featp[(size_t)(cq * 4 + j4) * TM + r];
#pragma unroll 1
                        for (int c16 = 0; c16 < 4; ++c16) {
                            const int col0 = c16 * 64 + cq * 16;
                            float h[16];
#pragma unroll
                            for (int j4 = 0; j4 < 4; ++j4) {
                                const float4 f4 = fn[j4];
                                h[4 * j4] = f4.x * ACT_SCALE; h[4 * j4 + 1] = f4.y * ACT_SCALE; h[4 * j4 + 2] = f4.z * ACT_SCALE; h[4 * j4 + 3] = f4.w * ACT_SCALE;
                            }
                            if (c16 < 3) {
#pragma unroll
                                for (int j4 = 0; j4 < 4; ++j4) fn[j4] = c.featp[(size_t)(((col0 + 64) >> 2) + j4) * TM + r];
                            }
                            store_a16(t_dd + col0, h, c.need_lo);
                            if ((lane & 7) == 0) {
#pragma unroll
                                for (int j4 = 0; j4 < 4; ++j4) discard_l2(c.featp + (size_t)((col0 >> 2) + j4) * TM + r);
                            }
                            signal_kb(c.kb_bar, c16, lane);
                        }
                    }
                    epi_bar_sync();                                       // all 39 d sdf/d emb entries complete
                    if (cq < 3) {
                        // nabla component cq in closed form: one column quarter per coordinate
                        const int cc = cq;
                        const float xc = S.X[cc * TM + r];
                        float n = c.misc[cc * TM + r];
#pragma unroll
                        for (int f = 0; f < 6; ++f) {
                            const float fr = (float)(1 << f);
                            float sn, cs;
                            if (STASH) { sn = S.EMBS[(3 + 6 * f + cc) * TM + r] * (1.f / ACT_SCALE); cs = S.EMBS[(6 + 6 * f + cc) * TM + r] * (1.f / ACT_SCALE); }
                            else sincosf(__fmul_rn(xc, fr), &sn, &cs);
                            n += fr * (c.misc[(3 + 6 * f + cc) * TM + r] * cs - c.misc[(6 + 6 * f + cc) * TM + r] * sn);
                        }
                        S.PART[cc * TM + r] = n;
                    }
                    epi_bar_sync();                                       // PART[0..2] = nabla of every row
                    if (cq == 0) {
                        const long long oo = S.OIDX[r];
                        if (oo >= 0 && job.nab) { job.nab[oo * 3] = S.PART[r]; job.nab[oo * 3 + 1] = S.PART[TM + r]; job.nab[oo * 3 + 2] = S.PART[2 * TM + r]; }
                    }
                }
                if (FULL && g == 20) {
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) { S.PART[(cq * 3 + cc) * TM + r] = rgb_part[cc]; rgb_part[cc] = 0.f; }
                    epi_bar_sync();
                    if (cq == 0 && S.OIDX[r] >= 0) {
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) {
                            // radiance layer 3 stored relu x16: undo in the head
                            const float z = (S.PART[cc * TM + r] + S.PART[(3 + cc) * TM + r] + S.PART[(6 + cc) * TM + r] + S.PART[(9 + cc) * TM + r])
                                            * (1.f / ACT_SCALE) + __ldg(pk + L.rad_b4 + cc);
                            job.rad[S.OIDX[r] * 3 + cc] = __fdiv_rn(1.f, 1.f + expf(-z));
                        }
                    }
                }
                if (g + 1 >= prog.n_gemm) {
                    tc_fence_before();
                    epi_bar_sync();                           // X / OIDX / PART and TMEM region 0 are rewritten by the next tile's input stage
                }
            }
        }
        if (job.dbg && blockIdx.x == 0 && tid == 64) { job.dbg[3] = clock64() - t_e0; job.dbg[4] = t_d; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_d, 512); }
}

// ------------------------------------------------------------------------------------------------
// weight image: for GEMM g (fp32 plane P[r][c], r = contraction index, c = output column) and K-block kb, two stages
// [hi | lo] of N rows x 64 k fp16 in the UMMA K-major SWIZZLE_128B shared-memory image (N = 256, or 64 for the last
// reverse GEMM), scaled by 2^shift so that max|W| lands in [256, 512).
// ------------------------------------------------------------------------------------------------
__global__ void pack_kernel(const float* __restrict__ pk, const size_t* __restrict__ offs, const int* __restrict__ rows,
                            const int* __restrict__ meta /* [g][4]: n_kb, N, w_off/16, - */, const float* __restrict__ absmax,
                            unsigned char* __restrict__ wimg, float* __restrict__ unscale) {
    const int g = blockIdx.y;
    const float* p = pk + offs[g];
    const int R = rows[g], n_kb = meta[g * 4], N = meta[g * 4 + 1];
    const size_t w_off = (size_t)meta[g * 4 + 2] * 16;
    const size_t sb = (size_t)N * 128;
    const float mx = absmax[g];
    int shift = 0;
    if (mx > 0.f) { int e; frexpf(mx, &e); shift = 9 - e; }           // mx * 2^shift in [256, 512)
    const float sc = ldexpf(1.f, shift);
    if (blockIdx.x == 0 && threadIdx.x == 0) unscale[g] = ldexpf(1.f, -shift) / ACT_SCALE;
    const int n_el = n_kb * N * 64;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_el; idx += gridDim.x * blockDim.x) {
        const int kk = idx & 63, n = (idx >> 6) % N, kb = (idx >> 6) / N;
        const int rr = kb * 64 + kk;
        const float w = rr < R ? p[(size_t)rr * 256 + n] * sc : 0.f;
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));
        const unsigned off = (unsigned)((n >> 3) * 1024 + (n & 7) * 128 + (((kk >> 3) ^ (n & 7)) << 4) + ((kk & 7) << 1));
        *reinterpret_cast<__half*>(wimg + w_off + (size_t)(kb * 2) * sb + off) = hi;
        *reinterpret_cast<__half*>(wimg + w_off + (size_t)(kb * 2 + 1) * sb + off) = lo;
    }
}

struct ImageLayout { unsigned w_off[N_PLANES]; int n_kb[N_PLANES], N[N_PLANES]; size_t image_bytes, unscale_off, meta_off, total; };

static ImageLayout image_layout() {
    ImageLayout T;
    size_t o = 0;
    for (int g = 0; g < N_PLANES; ++g) {
        T.n_kb[g] = (g == 0) ? 1 : 4;
        T.N[g] = (g == 16) ? 64 : 256;
        T.w_off[g] = (unsigned)o;
        o += (size_t)T.n_kb[g] * 2 * T.N[g] * 128;
    }
    T.image_bytes = o;
    T.unscale_off = (o + 1023) & ~(size_t)1023;
    T.meta_off = T.unscale_off + 256;
    T.total = T.meta_off + 1024;
    return T;
}

}  // namespace tm

size_t mlp_tmem_image_bytes() { return tm::image_layout().total; }

// `image` = this kernel's region of the packed buffer; d_offs / d_rows / d_absmax are the per-plane tables tc_pack() left on the device
int tmem_pack(const float* pk_f32, const size_t* d_offs, const int* d_rows, const float* d_absmax, unsigned char* image, cudaStream_t stream) {
    using namespace tm;
    const ImageLayout T = image_layout();
    int meta[N_PLANES * 4];
    for (int g = 0; g < N_PLANES; ++g) { meta[g * 4] = T.n_kb[g]; meta[g * 4 + 1] = T.N[g]; meta[g * 4 + 2] = (int)(T.w_off[g] / 16); meta[g * 4 + 3] = 0; }
    int* d_meta = (int*)(image + T.meta_off);
    NA_TRY(check_cuda(cudaMemcpyAsync(d_meta, meta, sizeof(meta), cudaMemcpyHostToDevice, stream)));
    pack_kernel<<<dim3(32, N_PLANES), 256, 0, stream>>>(pk_f32, d_offs, d_rows, d_meta, d_absmax, image, (float*)(image + T.unscale_off));
    NA_CHECK_LAUNCH();
    return NA_OK;
}

size_t mlp_tmem_scratch_bytes(int grid) { return (size_t)grid * tm::SCRATCH_BYTES; }

extern long long* g_tc_dbg;

// mixed != 0: feature head, reverse sweep and radiance layers use the hi*hi product only
int launch_mlp_tmem(const EvalJob& job_, const float* pk_f32, const unsigned char* image, const PackF32& L, int mixed,
                    unsigned char* scratch, size_t scratch_bytes, cudaStream_t stream) {
    using namespace tm;
    EvalJob job = job_; job.dbg = g_tc_dbg;
    static thread_local bool attr_set = false;
    // NA_TM_PRODS: diagnostics override, a 21-character string of '1'/'3' (products per GEMM of the program)
    static const char* prods_env = getenv("NA_TM_PRODS");
    const size_t smem = sizeof(Smem) + 1024;
    if (!attr_set) {
        NA_TRY(check_cuda(cudaFuncSetAttribute(mlp_tmem_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        NA_TRY(check_cuda(cudaFuncSetAttribute(mlp_tmem_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        NA_TRY(check_cuda(cudaFuncSetAttribute(mlp_tmem_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        attr_set = true;
    }
    const long long total = job.x ? job.m : (long long)job.n_rows * job.P;
    if (total <= 0) return NA_OK;
    const ImageLayout T = image_layout();
    Program prog; prog.n_gemm = 0;
    static const char* nsplit_env = getenv("NA_TM_NSPLIT");                   // diagnostics: "2" = N-halves (the r1n scheme), default quarters
    prog.nsplit = (nsplit_env && nsplit_env[0] == '2') ? 2 : 4;
    const int last = !job.want_full ? (job.feat ? 8 : 7) : (job.rad ? 20 : 16);
    for (int g = 0; g <= last; ++g) {
        Gemm t; t.w_off = T.w_off[g]; t.stage_bytes = (unsigned)T.N[g] * 128u; t.n_kb = (unsigned char)T.n_kb[g];
        t.prods = (unsigned char)((mixed && g >= 8) ? 1 : 3);
        if (prods_env && (int)strlen(prods_env) > g) t.prods = prods_env[g] == '1' ? 1 : 3;
        t.n64 = T.N[g] == 64; t.pad = 0;
        prog.g[prog.n_gemm++] = t;
    }
    long long tiles = (total + TM - 1) / TM;
    int grid = (int)(tiles < (long long)num_sms() ? tiles : (long long)num_sms());
    if (scratch_bytes < mlp_tmem_scratch_bytes(grid)) return NA_ERR_WORKSPACE;
    const float* usc = (const float*)(image + T.unscale_off);
    if (job.st_wide && !job.want_full) return NA_ERR_BAD_ARG;
    if (job.st_wide)        mlp_tmem_kernel<true, true><<<grid, THREADS, smem, stream>>>(job, pk_f32, L, image, usc, prog, scratch);
    else if (job.want_full) mlp_tmem_kernel<true, false><<<grid, THREADS, smem, stream>>>(job, pk_f32, L, image, usc, prog, scratch);
    else                    mlp_tmem_kernel<false, false><<<grid, THREADS, smem, stream>>>(job, pk_f32, L, image, usc, prog, scratch);
    NA_CHECK_LAUNCH();
    return NA_OK;
}

}  // namespace na
