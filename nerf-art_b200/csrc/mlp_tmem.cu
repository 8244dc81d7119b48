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
#include <cuda_bf16.h>
#include <cstdlib>
#include <cstring>

namespace na {
namespace tm {

constexpr int TM = 128;
// Register budget.  Two auxiliary warps (weight producer, MMA issuer) + 16 epilogue warps = 18 warps put FIVE warps on two of the four
// schedulers, so the kernel is compiled for 16384 / (5 x 32) = 102 -> 96 registers per thread.  (Tried: making the auxiliary warps a
// whole warpgroup that gives up registers with setmaxnreg.dec while the epilogue warpgroups take them with setmaxnreg.inc -- ptxas then
// spills 0.5-2.3 KB per kernel instead of 0.1-0.2 KB for every split tried (32 / 112, 40 / 104, 56 / 104, 64 / 104); not kept.)
constexpr int FIRST_EPI_WARP = 2;
constexpr int EPI_THREADS = 512;
constexpr int THREADS = 32 * FIRST_EPI_WARP + EPI_THREADS;
#ifndef NA_TM_NS
#define NA_TM_NS 5
#endif
constexpr int NS = NA_TM_NS;                     // weight stages (32 KB each); 5 leaves room for the encoding / small-weight stashes below
constexpr bool STASH = NS <= 5;
constexpr int STAGE_BYTES = 32768;               // 256 rows x 64 fp16
constexpr float ACT_SCALE = 16.f;                // activations are stored x16 (keeps the lo term normal in fp16)
constexpr int MAX_GEMM = 44;
constexpr int N_PLANES = 26;                     // weight images: 0..7 fwd | 8 feat | 9..15 reverse 7..1 | 16 reverse 0 | 17..20 radiance |
                                                 //                21..24 radiance backward (layers 3,2,1, 0-feature columns) | 25 head backward
// forward program = images 0..20 in order.  BW program (forward + backward of a training patch, 41 GEMMs) appends:
//   21..23 delta_{2,1,0} = (delta_{3,2,1} R_{3,2,1}) * relu'   | 24 feat-bar = delta_0 R_0[:, feature]   | 25 h-bar_7 = feat-bar W8[1:]
//   26..33 second-order sweep g-bar_i = v-bar_i W_i^T (images 0..7) | 34..40 trunk h-bar_{i-1} = z-bar_i W_i, i = 7..1 (images 9..15)

// NA_TM_TRACE (diagnostic build): CTA 0 records clock64() stamps of its second tile into job.dbg[16..]:
//   MMA lane:      slot (g*4+kb)*2 + {0: K-block of A ready, 1: its MMAs issued}                      (44 GEMMs: 352 slots)
//   epilogue lane: slot 352 + (g*4+pass)*3 + {0: D quarter ready, 1: tcgen05.ld done, 2: A stored}    (528 slots; buffer >= 16 + 880 int64)
// NA_TM_CYCLES (diagnostic build, implied by NA_TM_TRACE): CTA 0 accumulates the cycles its MMA lane / one epilogue lane spend in their
// waits into job.dbg[0..4] (scripts/tc_check.py prints them).  Not in the default build: two clock reads and a 64-bit accumulate around
// every wait are ~3 % of an SDF-only epilogue pass.
#if defined(NA_TM_TRACE) && !defined(NA_TM_CYCLES)
#define NA_TM_CYCLES
#endif
#ifdef NA_TM_CYCLES
#define NA_CYC(...) __VA_ARGS__
#else
#define NA_CYC(...)
#endif
#ifdef NA_TM_TRACE
#define NA_TRACE_M(tr, g, kb, w) do { if (tr) (tr)[((g) * 4 + (kb)) * 2 + (w)] = clock64(); } while (0)
#define NA_TRACE_E(tr, g, ps, w) do { if (tr) (tr)[352 + ((g) * 4 + (ps)) * 3 + (w)] = clock64(); } while (0)
#define NA_TRACE_X(tr, i) do { if (tr) (tr)[880 + (i)] = clock64(); } while (0)      // free-form stamps (trace_show.py prints them raw)
#else
#define NA_TRACE_M(tr, g, kb, w) do { } while (0)
#define NA_TRACE_E(tr, g, ps, w) do { } while (0)
#define NA_TRACE_X(tr, i) do { } while (0)
#endif

enum BwOp { OP_FWD = 0, OP_DR, OP_FB, OP_HB, OP_SO, OP_TR };      // OP_FWD: forward program, dispatched on the program index
struct Gemm { unsigned w_off; unsigned stage_bytes; unsigned char n_kb, prods, n64, img, op, lyr, pad0, pad1; float debias; };   // img: weight image (unscale index); debias: see launch_mlp_tmem
struct Program { int n_gemm; int g0; int nsplit; int corr_first; Gemm g[MAX_GEMM]; };      // g0: first GEMM run (21 in the backward-only launch)      // nsplit: N-parts (2 or 4) the last K-block of a GEMM is issued in
// corr_first (product order of a three-product GEMM; the tensor core's fp32 accumulate truncates toward zero, so the error is a BIAS that
// grows linearly with the number of accumulator updates made at full magnitude -- profiles/r3b_tc_accumulation.md):
//   0  K-block-interleaved hi*hi, lo*hi, hi*lo (all 48 updates of a 256-deep layer at full magnitude)
//   1  the two small correction products of ALL K-blocks first, hi*hi last (16 full-magnitude updates; 24 MMAs exposed after the
//      last K-block of A arrives; every W_hi stage is loaded twice)
//   2  corrections of K-blocks 0..n-2 as they arrive, then their hi*hi (runs under the previous epilogue's last pass), then the last
//      K-block as lo*hi, hi*lo, hi*hi per N-quarter (16 + 8 updates at (nearly) full magnitude; same exposed tail as order 0)

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
    float BWV[8 * TM];                                 // BW: per row 0..2 d L/d nabla (eikonal) x rs | 3..5 d L/d radiance x rs, then delta_4 x rs | 6 masked d L/d sdf x rs
    __align__(16) float RADW[STASH ? 9 * 256 : 4];
};

// per-CTA global scratch (full mode): 8 softplus' planes (16-bit codes), the geometry feature (fp32), misc rows
constexpr size_t DH_BYTES = (size_t)8 * 64 * TM * 8;          // plane p, column quad k4, row r -> uint2 at (p*64 + k4)*128 + r
constexpr size_t FEAT_BYTES = (size_t)64 * TM * 16;           // float4 at k4*128 + r
constexpr size_t MISC_BYTES = (size_t)80 * TM * 4;            // float at j*128 + r : d sdf/d emb (39) @0 | small radiance inputs (<=33) @40
// BW program only: q_0..6 = 100 g-bar g (1 - s) parked for the trunk and, as plane 7, the feature part of h-bar_7, as fp16 (both are
// x rs, i.e. O(1), and are added into fp16 operands anyway): uint4 = 8 columns at (plane*32 + k/8)*128 + r; the ReLU masks of the four
// radiance hidden layers
constexpr size_t QP_BYTES = (size_t)8 * 32 * TM * 16;
constexpr size_t GP_BYTES = 0;
constexpr size_t MK_BYTES = (size_t)4 * EPI_THREADS * 8;
constexpr size_t SCRATCH_BYTES = DH_BYTES + FEAT_BYTES + MISC_BYTES + QP_BYTES + GP_BYTES + MK_BYTES;
constexpr size_t TILE_BUF_BYTES = DH_BYTES + MK_BYTES;        // split training program: what the backward half needs from the forward half, per tile

enum EpiKind { K_FWD, K_FWD3, K_FWD7, K_FEAT, K_BWD, K_BWD4, K_BWD0, K_RAD0, K_RAD, K_RAD3,
               K_DR, K_FB, K_HB, K_SO, K_TR,                              // BW program only (K_DR: layer 0 too; K_SO: layers 3 and 7 too)
               // training program: ONE body per group, the layer-specific pieces behind warp-uniform run-time tests.  Its 41 epilogues
               // were 17.5 k SASS instructions (280 KB) walked once per tile: every kind's first pass paid 2-6 k cycles of instruction
               // fetch (profiles/r4a_bw_trace.md).  The render kernels keep the specialised kinds.
               K_FWDX, K_BWDX, K_RADX };

struct EpiCtx {
    // (loop-invariant state is kept small on purpose: the kernels run at their 96-register budget, and whatever does not fit is
    //  re-loaded from local memory inside the epilogue passes -- 56 more bytes of spills were 2.11 -> 2.9 ms per patch in the
    //  backward half, profiles/r4b_split_program.md.  featp / misc are constant offsets from qp; t_wait / trace exist in
    //  diagnostic builds only; the L2 policy of the stash stores is re-created where it is used.)
    Smem* S; uint2* dh; const float* pk; const PackF32* L; const EvalJob* job;
    __device__ __forceinline__ float4* featp() const { return reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(qp) - MISC_BYTES - FEAT_BYTES); }
    __device__ __forceinline__ float* misc() const { return reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(qp) - MISC_BYTES); }
    unsigned t_lane; int r, cq, g; float us; int sdim;
    unsigned s_base;                // shared-space address of the Smem block; the tables / barriers below are constant offsets from it
    __device__ __forceinline__ unsigned bias_s() const { return s_base + (unsigned)offsetof(Smem, BIAS); }
    __device__ __forceinline__ unsigned w8_s() const { return s_base + (unsigned)offsetof(Smem, W8); }
    __device__ __forceinline__ unsigned w4_s() const { return s_base + (unsigned)offsetof(Smem, W4); }
    __device__ __forceinline__ unsigned radw_s() const { return s_base + (unsigned)offsetof(Smem, RADW); }
    __device__ __forceinline__ unsigned kb_bar() const { return s_base + (unsigned)offsetof(Smem, kb_ready); }
    __device__ __forceinline__ unsigned d_bar() const { return s_base + (unsigned)offsetof(Smem, d_ready); }
    int signal, need_lo, lane;
    unsigned short* st_row;         // ST: this thread's row in plane 0 of the 16-bit stash (st_wide + m * 256), nullptr beyond the allocation
    int st_mpad;                    // ST: rows per plane (the TMA row coordinate of plane p, sample m is p * st_mpad + m: below 2^31)
    long long st_m;                 // ST: flat sample index of the row
    __device__ __forceinline__ const TmaMap* st_map() const { return &job->st_store_map; }   // ST: tensor map of the wide planes (store boxes: 16 columns x 32 samples)
    unsigned st_stg;                // ST: this warp's two 1 KB staging buffers in shared memory
    int st_row0;                    // ST: first sample (row of the plane) of this warp's 32 lanes in the current tile
    mutable unsigned st_cnt;        // ST: TMA stores issued by this warp so far (buffer parity)
    // BW: per-row power-of-two scale of the upstream gradient (the backward is linear in it and rows are independent, so every
    // backward quantity of the row is carried x rs in the fp16 operands and stored x irs) and the row's total d L / d nabla (x rs)
    float rs, irs, nbar[3];
    int lyr, has_rad;               // BW: layer index of the running backward GEMM; the program has the radiance part
    unsigned long long* mk;         // BW: this thread's ReLU-mask slots: mk[l * EPI_THREADS], l = radiance hidden layer (64 columns each)
    uint4* qp;                      // BW: per-CTA fp16 scratch planes
    int bw;
    unsigned nx;                    // FULL: scratch planes the NEXT GEMM's epilogue reads, for the prefetch: softplus' plane | parked-term plane << 8 |
                                    // g plane << 16 | (feature rows) << 24; 0xff = none
    unsigned d_phase;
#ifdef NA_TM_CYCLES
    long long* t_wait; long long* trace;
#endif
};

// this warp's part of K-block kb of the next A operand is in TMEM (and its part of D columns [64kb, 64kb+64) is consumed)
__device__ __forceinline__ void signal_kb(unsigned kb_bar, int kb, int lane) {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(kb_bar + 8u * (unsigned)kb);
}

// two floats -> packed fp16x2 (low half = a), round to nearest, saturating to the largest finite fp16 (F2FP.SATFINITE: the clamp is
// part of the conversion instruction)
__device__ __forceinline__ unsigned pack_half2_sat(float a, float b) {
    unsigned r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// 16 fp32 values (already x ACT_SCALE) -> 8 packed hi words + 8 packed lo words, stored over the 16 columns at taddr
// (need_lo == 0: the consumer GEMM uses the hi*hi product only, the lo words are neither computed nor stored)
__device__ __forceinline__ void store_a16(unsigned taddr, const float (&o)[16], int need_lo) {
    unsigned hi[8];
    __half2 h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { hi[i] = pack_half2_sat(o[2 * i], o[2 * i + 1]); h[i] = *reinterpret_cast<const __half2*>(&hi[i]); }
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

// softplus'(z) = sigmoid(100 z) as a 16-bit code.
// NA_TM_CODE15 (the round-2 scheme): bit 15 = (z >= 0), low 15 bits = round(t * 32767.49), t = exp(-|100 z|); decoded as
// r = 1 / (1 + code/32768), z >= 0 ? r : 1 - r (absolute error <= 2e-5): cheap to encode (no MUFU), but every decode is two bit
// operations, one MUFU.RCP, an add, a bit test and a select.
// Default: code = round(s * 65535), s = z >= 0 ? r : t r with r = 1 / (1 + t) (one MUFU.RCP and two more FMA-pipe instructions at the
// ONE encode); decoded as (code + 2^23 as a float) - 2^23, times 1/65535: three instructions, no MUFU, absolute error <= 7.7e-6, exact
// 0 and 1 (saturated units carry no bias).  The codes are decoded once in the render kernel and three times in the training programs
// (reverse sweep, second-order sweep, trunk).
#ifdef NA_TM_CODE15
__device__ __forceinline__ unsigned dh_code(float z16, float t) {
    // round(t * 32767.49) through the 2^23 magic number (FMA pipe; F2I would go to the XU pipe the softplus already saturates)
    return ((~__float_as_uint(z16) >> 16) & 0x8000u) | (__float_as_uint(fmaf(t, 32767.49f, 8388608.f)) & 0x7fffu);
}
__device__ __forceinline__ unsigned dh_pack2(unsigned a, unsigned b) { return a | (b << 16); }
__device__ __forceinline__ float dh_decode_lo(unsigned w) {            // code in bits [0,16)
    const float ru = rcp_approx(__uint_as_float(((w & 0x7fffu) << 8) | 0x3f800000u));
    return (w & 0x8000u) ? ru : 1.f - ru;
}
__device__ __forceinline__ float dh_decode_hi(unsigned w) {            // code in bits [16,32)
    const float ru = rcp_approx(__uint_as_float(((w >> 8) & 0x7fff00u) | 0x3f800000u));
    return (w & 0x80000000u) ? ru : 1.f - ru;
}
#else
__device__ __forceinline__ unsigned dh_code(float z16, float t) {      // returns 2^23-biased float bits: the code is the low 16 bits
    const float r = rcp_approx(1.f + t);
    const float s = z16 >= 0.f ? r : t * r;
    return __float_as_uint(fmaf(s, 65535.f, 8388608.f));
}
__device__ __forceinline__ unsigned dh_pack2(unsigned a, unsigned b) { return __byte_perm(a, b, 0x5410); }      // low halves of a | b << 16
__device__ __forceinline__ float dh_decode_lo(unsigned w) {            // code in bits [0,16)
    return (__uint_as_float((w & 0xffffu) | 0x4b000000u) - 8388608.f) * (1.f / 65535.f);
}
__device__ __forceinline__ float dh_decode_hi(unsigned w) {            // code in bits [16,32)
    return (__uint_as_float(__byte_perm(w, 0x4b000000u, 0x7632)) - 8388608.f) * (1.f / 65535.f);
}
#endif
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

// 16 floats -> 16 packed bf16 values
__device__ __forceinline__ void pack16(const float (&o)[16], float scale, uint4& lo, uint4& hi) {
    unsigned w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(o[2 * i] * scale, o[2 * i + 1] * scale);
        w[i] = *reinterpret_cast<const unsigned*>(&h);
    }
    lo = make_uint4(w[0], w[1], w[2], w[3]); hi = make_uint4(w[4], w[5], w[6], w[7]);
}
__device__ __forceinline__ void sts128(unsigned addr, const uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// ST: 16 consecutive columns of every row of this warp (32 samples x 32 B) -> stash plane `plane`, values x `scale`.  The rows are
// staged in shared memory (lane = row, 32 B each) and written by ONE TMA tensor store per warp (box 16 columns x 32 samples): the
// store engine scatters the 32-byte pieces over the 512-byte-strided rows of the plane, the LSU sees two conflict-free 16-byte
// shared-memory stores per thread.  Two staging buffers per warp; a buffer is reused once the store before last has read it.
// Must be called by all 32 lanes.
__device__ __forceinline__ void stash16(const EpiCtx& c, int plane, int col0, const float (&o)[16], float scale) {
    uint4 lo, hi;
    pack16(o, scale, lo, hi);
    const unsigned buf = c.st_stg + (c.st_cnt & 1u) * 1024u;
    if (c.lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    sts128(buf + (unsigned)c.lane * 32u, lo); sts128(buf + (unsigned)c.lane * 32u + 16u, hi);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (c.lane == 0) {
        // L2 evict-first: the planes stream out to HBM and must not displace the per-CTA scratch (read back three times per tile)
        unsigned long long st_policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(st_policy));
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;"
                     :: "l"(c.st_map()), "r"(col0), "r"(plane * c.st_mpad + c.st_row0), "r"(buf), "l"(st_policy) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    ++c.st_cnt;
}
// all TMA stores of this warp are complete and visible to its later global loads (called by all 32 lanes)
__device__ __forceinline__ void stash_flush(const EpiCtx& c) {
    if (c.lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async.global;" ::: "memory");
    __syncwarp();
}

// BW: entry k of v-bar_0 = (d emb / d x)^T-contracted total d L / d nabla of row r (x rs): x_c -> n_c, sin(f x_c) -> f cos(f x_c) n_c,
// cos(f x_c) -> -f sin(f x_c) n_c; sin / cos come from the tile's encoding stash (x ACT_SCALE)
// entries K_LO .. K_LO + 15 with the index arithmetic done at compile time: straight-line code.  (A run-time-index version,
// sixteen times in a row, was ~750 branchy instructions that run once per tile, i.e. always from a cold instruction cache: 24-38 k
// cycles per tile on the clock trace, profiles/r4a_bw_trace.md.)
template <int K_LO>
__device__ __forceinline__ void vbar0_range(const Smem& S, int r, const float (&nbar)[3], float (&e)[16]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int k = K_LO + j;
        float v = 0.f;
        if (k >= 0 && k < 3) v = nbar[k < 0 ? 0 : (k > 2 ? 2 : k)];
        else if (k >= 3 && k < EMB) {
            const int f = (k - 3) / 6, rem = (k - 3) % 6, cc = rem % 3;
            const float w = (rem < 3 ? 1.f : -1.f) * (float)(1 << f) * (1.f / ACT_SCALE);
            v = nbar[cc] * w * S.EMBS[(rem < 3 ? k + 3 : k - 3) * TM + r];
        }
        e[j] = v;
    }
}
// the 16 entries of column quarter cq (warp-uniform): first entry 16 cq + BASE, BASE = 0 (v-bar_0 itself) or -25 (skip columns of layer 3)
template <int BASE>
__device__ __forceinline__ void vbar0_quarter(const Smem& S, int cq, int r, const float (&nbar)[3], float (&e)[16]) {
    if (cq == 0) vbar0_range<BASE>(S, r, nbar, e);
    else if (cq == 1) vbar0_range<BASE + 16>(S, r, nbar, e);
    else if (cq == 2) vbar0_range<BASE + 32>(S, r, nbar, e);
    else vbar0_range<BASE + 48>(S, r, nbar, e);
}
// BW: 16 consecutive columns of row r in a per-CTA fp16 scratch plane (coalesced: a warp instruction covers 32 rows x 16 B)
// The planes hold value x 2^-8 (QP_SCALE): the parked terms reach 1e4 (100 g-bar g); the saturating conversion keeps an outlier finite
// instead of turning the patch into NaN.  PRESCALED: `v` already carries the factor (folded into the caller's constants).
constexpr float QP_SCALE = 0.00390625f, QP_UNSCALE = 256.f;
template <bool PRESCALED>
__device__ __forceinline__ void qstore16(uint4* base, int plane, int col0, int r, const float (&v)[16]) {
    uint4* p = base + (size_t)(plane * 32 + (col0 >> 3)) * TM + r;
#pragma unroll
    for (int j8 = 0; j8 < 2; ++j8) {
        unsigned w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            w[i] = PRESCALED ? pack_half2_sat(v[8 * j8 + 2 * i], v[8 * j8 + 2 * i + 1])
                             : pack_half2_sat(v[8 * j8 + 2 * i] * QP_SCALE, v[8 * j8 + 2 * i + 1] * QP_SCALE);
        p[(size_t)j8 * TM] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}
// decode of the prefetched rows: parked fp16 terms (x 2^-8), bf16 stash rows
__device__ __forceinline__ void qdecode16(const uint4 (&q)[2], float (&v)[16], float scale = QP_UNSCALE) {       // scale: QP_UNSCALE x the caller's unit
#pragma unroll
    for (int j8 = 0; j8 < 2; ++j8) {
        const unsigned w[4] = {q[j8].x, q[j8].y, q[j8].z, q[j8].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i])); v[8 * j8 + 2 * i] = f.x * scale; v[8 * j8 + 2 * i + 1] = f.y * scale; }
    }
}
__device__ __forceinline__ void bf16x16_to_float(const uint4 (&q)[2], float (&v)[16]) {
#pragma unroll
    for (int j8 = 0; j8 < 2; ++j8) {
        const unsigned w[4] = {q[j8].x, q[j8].y, q[j8].z, q[j8].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i])); v[8 * j8 + 2 * i] = f.x; v[8 * j8 + 2 * i + 1] = f.y; }
    }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
// The per-CTA scratch (1.2 MB x 148 CTAs in the training program) and the g planes of the stash do not stay in L2 between their
// producer and their consumers: 78 % of the epilogue's scratch sectors came from DRAM and their latency was 37 % of all stall samples
// (profiles/r3d_bw_program.md).  A pass therefore requests, one whole GEMM ahead (~10 us), the sectors the same pass of the NEXT GEMM
// will load: softplus' codes (uint2 per row and column quad: 4-row sectors, lane l takes quad l & 3), parked fp16 terms (uint4 per
// row and 8 columns: 2-row sectors, lane l takes half l & 1), the thread's own 32-byte piece of the g plane.  No register is held:
// prefetch.global.L2 has no destination.
template <bool ST>
__device__ __forceinline__ void prefetch_next_gemm(const EpiCtx& c, int col0) {
    const unsigned dh_plane = c.nx & 0xffu, q_plane = (c.nx >> 8) & 0xffu, g_plane = (c.nx >> 16) & 0xffu;
    if (dh_plane != 0xffu) prefetch_l2(c.dh + (size_t)(dh_plane * 64 + (col0 >> 2) + (c.lane & 3)) * TM + c.r);
    if (ST && q_plane != 0xffu) prefetch_l2(c.qp + (size_t)(q_plane * 32 + (col0 >> 3) + (c.lane & 1)) * TM + c.r);
    if (ST && g_plane != 0xffu && c.st_row) prefetch_l2(c.st_row + ((size_t)(ST_G + g_plane) * (size_t)c.st_mpad << 8) + col0);
    if (c.nx >> 24) {
        // geometry feature rows read by the tail of GEMM 16 (float4 per row and column quad: 2-row sectors)
        const float4* f = c.featp() + (size_t)((col0 >> 2) + (c.lane & 3)) * TM;
        prefetch_l2(f + c.r); prefetch_l2(f + (c.r ^ 2));
    }
}
// planes word of GEMM (op, lyr, program index g); see EpiCtx::nx
__device__ __forceinline__ unsigned scratch_planes_of(int op, int lyr, int g, int has_rad) {
    unsigned dh = 0xffu, q = 0xffu, gp = 0xffu, ft = 0u;
    if (op == OP_FWD) { if (g >= 8 && g <= 15) dh = (unsigned)(15 - g); if (g == 16 && has_rad) ft = 1u; }
    else if (op == OP_SO) { dh = (unsigned)lyr; gp = (unsigned)lyr; if (lyr == 7 && has_rad) q = 7u; }
    else if (op == OP_TR) { dh = (unsigned)lyr; q = (unsigned)lyr; }
    return dh | (q << 8) | (gp << 16) | (ft << 24);
}

// one GEMM's epilogue for this thread's row and its 64 columns (4 passes of 16)
// trace build: stamps inside pass 1 of the second-order sweep's third GEMM (program index 28)
#define NA_TRACE_XS(i) do { if (KIND == K_SO && c.g == 28 && c16 == 1) NA_TRACE_X(c.trace, i); } while (0)
template <int KIND, bool FULL, bool ST>
__device__ __forceinline__ void epi_gemm(const EpiCtx& c, const unsigned t_d, float& sdf_part, float (&rgb_part)[3], const float (&small_in)[36]) {
    Smem& S = *c.S;
    const int r = c.r;
    const float us = c.us, us16 = c.us * ACT_SCALE;
    constexpr bool IS_FWD = KIND == K_FWD || KIND == K_FWD3 || KIND == K_FWD7 || KIND == K_FWDX;
    constexpr bool IS_BWD = KIND == K_BWD || KIND == K_BWD4 || KIND == K_BWDX;
    constexpr bool IS_RAD = KIND == K_RAD0 || KIND == K_RAD || KIND == K_RAD3 || KIND == K_RADX;
    constexpr bool IS_BW = KIND >= K_DR && KIND <= K_TR;
    const bool fwd3 = KIND == K_FWD3 || (KIND == K_FWDX && c.g == 3), fwd7 = KIND == K_FWD7 || (KIND == K_FWDX && c.g == 7);
    const bool bwd4 = KIND == K_BWD4 || (KIND == K_BWDX && c.g == 12), rad3 = KIND == K_RAD3 || (KIND == K_RADX && c.g == 20);
    const bool dr0 = KIND == K_DR && c.lyr == 0, so3 = KIND == K_SO && c.lyr == 3, so7 = KIND == K_SO && c.lyr == 7;
    const unsigned bias = c.bias_s() + (unsigned)(KIND == K_FEAT ? 8 : (IS_RAD ? 9 + (c.g - 17) : c.g)) * 1024u;
    constexpr bool USES_DH = FULL && (KIND == K_FEAT || IS_BWD);
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
    // BW (second-order sweep, trunk): the softplus' codes run one pass ahead in registers.  In the backward-only launch D is ready
    // whenever the epilogue asks (the GEMMs are epilogue-bound), so a load issued at the top of its own pass is fully exposed: 55 % of
    // the kernel's stall samples were long-scoreboard waits on these codes (profiles/r4b_split_program.md).  Pass c+1's codes are
    // requested right after pass c's tcgen05.ld: kernel 2.26 -> 2.12 ms per patch.  (Carrying pass 0's codes across the GEMM boundary
    // as well costs eight more live registers through the tails: 400 B of spills and 3.1 ms; the parked terms one pass ahead: no change;
    // the g-plane rows one pass ahead: spills, 2.63 ms -- all measured (profiles/r4b_split_program.md), none kept.)
    constexpr bool PRE_S = KIND == K_SO || KIND == K_TR;
    constexpr bool PRE_G = KIND == K_SO;
    uint2 sn[4];
    if (PRE_S) {
        const uint2* p = c.dh + (size_t)(c.lyr * 64 + c.cq * 4) * TM + r;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) sn[j4] = p[(size_t)j4 * TM];
    }
#pragma unroll 1
    for (int c16 = 0; c16 < N_PASS; ++c16) {
        // thread = (row, column quarter cq): in pass c16 it owns columns 64*c16 + 16*cq .. +16, i.e. every pass completes one
        // 64-wide K-block of the next layer's A operand across the 16 epilogue warps
        const int col0 = c16 * 64 + c.cq * 16;
        NA_TRACE_XS(7);
#ifndef NA_TM_NO_PREFETCH
        // (training program only: in the render kernel, whose 0.7 MB scratch per CTA mostly stays in L2, the extra requests cost 3 %)
        if (FULL && ST) prefetch_next_gemm<ST>(c, col0);
#endif
        uint2 q[4];
        if (USES_DH) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) q[j4] = qn[j4];
            if (c16 + 1 < N_PASS) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) qn[j4] = dhp[(size_t)(((col0 + 64) >> 2) + j4) * TM];
            }
        }
        // BW: the scratch / stash rows this pass needs (softplus' codes, g, parked terms) are requested BEFORE the wait for D, so that
        // their L2 / DRAM latency runs under the GEMM instead of on the epilogue's critical path
        const bool PRE_Q = KIND == K_TR || so7;
        uint2 sraw[4]; uint4 graw[2], qraw[2];
        if (PRE_S) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) sraw[j4] = sn[j4];
        }
        if (PRE_G) {
            if (c.lyr == 0 && c16 == 0) { NA_TRACE_X(c.trace, 5); stash_flush(c); NA_TRACE_X(c.trace, 6); }     // the g planes (TMA-stored during the reverse sweep) are read back from here on
            const uint4* p = reinterpret_cast<const uint4*>(c.st_row + ((size_t)(ST_G + c.lyr) * (size_t)c.st_mpad << 8) + col0);
            graw[0] = __ldcg(p); graw[1] = __ldcg(p + 1);
        }
        if (PRE_Q) {
            const uint4* p = c.qp + (size_t)((KIND == K_TR ? c.lyr : 7) * 32 + (col0 >> 3)) * TM + r;
            if (KIND == K_TR || c.has_rad) { qraw[0] = p[0]; qraw[1] = p[TM]; }
            else { qraw[0] = make_uint4(0u, 0u, 0u, 0u); qraw[1] = qraw[0]; }
        }
        {                                                            // pass c16 reads N-quarter c16 of D
            NA_CYC(const long long t0 = clock64();)
            mbar_wait_plain(c.d_bar() + 8u * (unsigned)c16, c.d_phase);
            if (N_PASS < 4) {
                // a GEMM whose epilogue reads fewer than four N-quarters (the 64-wide reverse GEMM 0: its four d_ready commits are
                // issued together): consume the other phases here, before anything is signalled to the MMA warp
                for (int k = N_PASS; k < 4; ++k) mbar_wait_plain(c.d_bar() + 8u * (unsigned)k, c.d_phase);
            }
            NA_CYC(*c.t_wait += clock64() - t0;)
            tc_fence_after();
            NA_TRACE_E(c.trace, c.g, c16, 0);
        }
        float acc[16];
        {
            unsigned v[16];
            tmem_ld16(t_d + col0, v);
            tmem_wait_ld();
            NA_TRACE_E(c.trace, c.g, c16, 1);
            NA_TRACE_XS(10);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(v[j]);
        }
        if (PRE_S && c16 + 1 < N_PASS) {
            const uint2* p = c.dh + (size_t)(c.lyr * 64 + ((col0 + 64) >> 2)) * TM + r;
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) sn[j4] = p[(size_t)j4 * TM];
        }
        float o[16];
        if (IS_FWD) {
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
                        // skip-connection columns (k >= 217) carry no softplus: their code decodes to 0.  The test is on the pass
                        // (warp-uniform, one branch) before it is on the column: in the merged kind of the training programs `fwd3` is
                        // a run-time value and a per-element test costs 4 % of the forward half (profiles/r4b_split_program.md)
                        if (KIND == K_FWD3) { if (col0 + j >= SKIP_H) cd[i] = 0u; }
                    }
                    if (KIND == K_FWDX && fwd3 && c16 == 3 && c.cq >= 1) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) if (col0 + 4 * j4 + i >= SKIP_H) cd[i] = 0u;
                    }
                    c.dh[(size_t)(c.g * 64 + (col0 >> 2) + j4) * TM + r] = make_uint2(dh_pack2(cd[0], cd[1]), dh_pack2(cd[2], cd[3]));
                }
            }
            if (fwd3 && c16 == 3 && c.cq >= 1) {
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
            }
            if (fwd7) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 w4 = lds128(c.w8_s() + (unsigned)(col0 + 4 * j4) * 4u);
                    sdf_part = fmaf(o[4 * j4], w4.x, sdf_part); sdf_part = fmaf(o[4 * j4 + 1], w4.y, sdf_part);
                    sdf_part = fmaf(o[4 * j4 + 2], w4.z, sdf_part); sdf_part = fmaf(o[4 * j4 + 3], w4.w, sdf_part);
                }
            }
        } else if (KIND == K_FEAT) {
            float fst[16];
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 b4 = lds128(bias + (unsigned)(col0 + 4 * j4) * 4u);
                const float4 f4 = make_float4(fmaf(acc[4 * j4], us, b4.x), fmaf(acc[4 * j4 + 1], us, b4.y),
                                              fmaf(acc[4 * j4 + 2], us, b4.z), fmaf(acc[4 * j4 + 3], us, b4.w));
                if (FULL) c.featp()[(size_t)((col0 >> 2) + j4) * TM + r] = f4;
                if (ST) { fst[4 * j4] = f4.x; fst[4 * j4 + 1] = f4.y; fst[4 * j4 + 2] = f4.z; fst[4 * j4 + 3] = f4.w; }
                if (c.job->feat && S.OIDX[r] >= 0) *(reinterpret_cast<float4*>(c.job->feat + S.OIDX[r] * 256 + col0) + j4) = f4;
                if (FULL) {
                    // next A: d sdf / d z7 = W8[0,:] * softplus'(z7)
                    float d4[4];
                    dh_decode4(q[j4], d4);
                    const float4 w4 = lds128(c.w8_s() + (unsigned)(col0 + 4 * j4) * 4u);
                    o[4 * j4] = w4.x * d4[0] * ACT_SCALE; o[4 * j4 + 1] = w4.y * d4[1] * ACT_SCALE;
                    o[4 * j4 + 2] = w4.z * d4[2] * ACT_SCALE; o[4 * j4 + 3] = w4.w * d4[3] * ACT_SCALE;
                }
            }
            if (ST) stash16(c, ST_FEAT, col0, fst, 1.f);
            if (ST && FULL) { stash16(c, ST_G + 7, col0, o, 1.f / ACT_SCALE); }
        } else if (IS_BWD) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                float dd[4];
                dh_decode4(q[j4], dd);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = col0 + 4 * j4 + i;
                    if (KIND == K_BWD4) { if (k >= SKIP_H) c.misc()[(k - SKIP_H) * TM + r] = acc[4 * j4 + i] * us; }      // embedding branch of the skip
                    o[4 * j4 + i] = acc[4 * j4 + i] * us16 * dd[i];
                }
            }
            if (KIND == K_BWDX && bwd4 && c16 == 3 && c.cq >= 1) {
                // (merged kind: the same stores behind ONE warp-uniform test; columns >= 217 lie in pass 3 of column quarters 1..3)
#pragma unroll
                for (int j = 0; j < 16; ++j) { const int k = col0 + j; if (k >= SKIP_H) c.misc()[(k - SKIP_H) * TM + r] = acc[j] * us; }
            }
            if (ST) { stash16(c, ST_G + 15 - c.g, col0, o, 1.f / ACT_SCALE); }
        } else if (KIND == K_BWD0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) { const int k = col0 + j; if (k < EMB) c.misc()[k * TM + r] += acc[j] * us; }
        } else if (IS_BW) {
            // ---- backward program (BW): every value is carried x rs in the operands; stash planes receive x irs ----------------
            if (KIND == K_DR) {
                // delta_lyr = (delta_{lyr+1} R_{lyr+1}) * [ys_{lyr+1} > 0]   (radiance hidden layers, lyr = 2, 1, 0)
                const unsigned long long m64 = c.mk[(size_t)c.lyr * EPI_THREADS] >> (16 * c16);
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] = ((m64 >> j) & 1ull) ? acc[j] * us16 : 0.f;
                stash16(c, ST_D + c.lyr, col0, o, c.irs * (1.f / ACT_SCALE));
                if (dr0) {
                    // d L / d nabla through radiance layer 0: delta_0 . W0[:, nabla columns]  (partial over this thread's columns)
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc)
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const int srow = c.sdim - 3 + cc;
                            const float4 w = (STASH && c.sdim == 9) ? lds128(c.radw_s() + (unsigned)(srow * 256 + col0 + 4 * j4) * 4u)
                                                                    : __ldg(reinterpret_cast<const float4*>(c.pk + c.L->rad_wt[0] + (size_t)(256 + srow) * 256 + col0) + j4);
                            rgb_part[cc] = fmaf(o[4 * j4], w.x, rgb_part[cc]); rgb_part[cc] = fmaf(o[4 * j4 + 1], w.y, rgb_part[cc]);
                            rgb_part[cc] = fmaf(o[4 * j4 + 2], w.z, rgb_part[cc]); rgb_part[cc] = fmaf(o[4 * j4 + 3], w.w, rgb_part[cc]);
                        }
                }
            } else if (KIND == K_FB) {
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] = acc[j] * us16;                       // d L / d feature
                stash16(c, ST_FB, col0, o, c.irs * (1.f / ACT_SCALE));
            } else if (KIND == K_HB) {
                // feature part of h-bar_7, parked (x rs) in scratch plane 7 until the second-order sweep reaches layer 7
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] = acc[j] * us;
                qstore16<false>(c.qp, 7, col0, r, o);
            } else if (KIND == K_SO) {
                // second-order sweep, layer lyr: g-bar = W v-bar; u-bar = g-bar s = v-bar_{lyr+1}; q = 100 g-bar g (1 - s) joins z-bar_lyr.
                // Units are folded into the constants: o is carried x ACT_SCALE (the A operand's unit; the stash scale undoes it), q is
                // produced x QP_SCALE (the parked plane's unit): t = 16 g-bar, q' = (t g) ((1 - s) 100 / 16 / 256).
                // (written four columns at a time so that neither the decoded softplus' values nor the converted g values stay live: the
                //  kernel runs at its 96-register budget and every spilled value in this loop costs -- 148 -> 204 bytes of spills were
                //  2.11 -> 2.89 ms per patch, profiles/r4b_split_program.md)
                float q[16];                                       // parked term x QP_SCALE; layer 7: z-bar_7 x ACT_SCALE instead
                constexpr float QK = 100.f / ACT_SCALE * QP_SCALE;
                const float gs16 = so7 ? S.BWV[6 * TM + r] * ACT_SCALE : 0.f;
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    float s4[4], g4[4];
                    dh_decode4(sraw[j4], s4);
                    {
                        const uint4 gq = graw[j4 >> 1];
                        const unsigned w0 = (j4 & 1) ? gq.z : gq.x, w1 = (j4 & 1) ? gq.w : gq.y;
                        g4[0] = __uint_as_float(w0 << 16); g4[1] = __uint_as_float(w0 & 0xffff0000u);
                        g4[2] = __uint_as_float(w1 << 16); g4[3] = __uint_as_float(w1 & 0xffff0000u);
                    }
                    float h4[4] = {0.f, 0.f, 0.f, 0.f};
                    if (so7) {
                        // h-bar_7 = (parked feature part) + masked d L / d sdf * W8[0,:]   (x ACT_SCALE)
                        const uint4 hq = qraw[j4 >> 1];
                        const unsigned w0 = (j4 & 1) ? hq.z : hq.x, w1 = (j4 & 1) ? hq.w : hq.y;
                        const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&w0)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&w1));
                        const float4 w4 = lds128(c.w8_s() + (unsigned)(col0 + 4 * j4) * 4u);
                        h4[0] = fmaf(gs16, w4.x, f0.x * (QP_UNSCALE * ACT_SCALE)); h4[1] = fmaf(gs16, w4.y, f0.y * (QP_UNSCALE * ACT_SCALE));
                        h4[2] = fmaf(gs16, w4.z, f1.x * (QP_UNSCALE * ACT_SCALE)); h4[3] = fmaf(gs16, w4.w, f1.y * (QP_UNSCALE * ACT_SCALE));
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int j = 4 * j4 + i;
                        const float t = acc[j] * us16;
                        const float qq = (t * g4[i]) * fmaf(s4[i], -QK, QK);
                        o[j] = t * s4[i];
                        // layer 7: z-bar_7 = h-bar_7 s_7 + q_7
                        q[j] = so7 ? fmaf(h4[i], s4[i], qq * (QP_UNSCALE * ACT_SCALE)) : qq;
                    }
                }
                if (so3 && c16 == 3 && c.cq >= 1) {
                    // skip connection: columns k >= SKIP_H = 217 of v-bar_4 are v-bar_0 entries k - 217 (this thread: 16 cq - 25 ..)
                    float e[16];
                    vbar0_quarter<-25>(S, c.cq, r, c.nbar, e);
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (16 * c.cq - 25 + j >= 0) o[j] = e[j] * ACT_SCALE;
                }
                NA_TRACE_XS(11);
                stash16(c, ST_VB + c.lyr, col0, o, c.irs * (1.f / ACT_SCALE));
                NA_TRACE_XS(12);
                if (so7) {
                    stash16(c, ST_ZB + 7, col0, q, c.irs * (1.f / ACT_SCALE));
#pragma unroll
                    for (int j = 0; j < 16; ++j) o[j] = q[j];                            // the trunk's first A operand is z-bar_7
                } else {
                    qstore16<true>(c.qp, c.lyr, col0, r, q);                             // parked (x rs) until the trunk reaches this layer
                }
                NA_TRACE_XS(13);
            } else {
                // trunk, K_TR: z-bar_lyr = h-bar_lyr s_lyr + q_lyr   (carried x ACT_SCALE, see K_SO)
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    float s4[4];
                    dh_decode4(sraw[j4], s4);
                    const uint4 hq = qraw[j4 >> 1];
                    const unsigned w0 = (j4 & 1) ? hq.z : hq.x, w1 = (j4 & 1) ? hq.w : hq.y;
                    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&w0)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&w1));
                    o[4 * j4] = fmaf(acc[4 * j4] * us16, s4[0], f0.x * (QP_UNSCALE * ACT_SCALE));
                    o[4 * j4 + 1] = fmaf(acc[4 * j4 + 1] * us16, s4[1], f0.y * (QP_UNSCALE * ACT_SCALE));
                    o[4 * j4 + 2] = fmaf(acc[4 * j4 + 2] * us16, s4[2], f1.x * (QP_UNSCALE * ACT_SCALE));
                    o[4 * j4 + 3] = fmaf(acc[4 * j4 + 3] * us16, s4[3], f1.y * (QP_UNSCALE * ACT_SCALE));
                }
                stash16(c, ST_ZB + c.lyr, col0, o, c.irs * (1.f / ACT_SCALE));
            }
            // (single-product fp16 operands: an outlier |x rs| > 4e3 saturates in store_a16's conversion instead of rounding to inf)
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
                            const float4 w = STASH ? lds128(c.radw_s() + (unsigned)(j * 256 + col0 + 4 * j4) * 4u) : __ldg(wsm + j * 64);
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
            if (ST) {
                unsigned long long m16 = 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) m16 |= (unsigned long long)(o[j] > 0.f) << j;
                unsigned long long* slot = c.mk + (size_t)(c.g - 17) * EPI_THREADS;
                *slot = (c16 == 0 ? 0ull : *slot) | (m16 << (16 * c16));
            }
            if (rad3) {
#pragma unroll
                for (int cc = 0; cc < 3; ++cc)
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 w = lds128(c.w4_s() + (unsigned)(cc * 256 + col0 + 4 * j4) * 4u);
                        rgb_part[cc] = fmaf(o[4 * j4], w.x, rgb_part[cc]); rgb_part[cc] = fmaf(o[4 * j4 + 1], w.y, rgb_part[cc]);
                        rgb_part[cc] = fmaf(o[4 * j4 + 2], w.z, rgb_part[cc]); rgb_part[cc] = fmaf(o[4 * j4 + 3], w.w, rgb_part[cc]);
                    }
            }
        }
        const bool store = !(KIND == K_BWD0 || rad3 || KIND == K_HB || (fwd7 && !FULL && !c.job->feat) || (KIND == K_FEAT && !FULL) ||
                             (IS_BW && !c.signal));
        if (store) store_a16(t_d + col0, o, c.need_lo);
        NA_TRACE_E(c.trace, c.g, c16, 2);
        NA_TRACE_XS(14);
        if (USES_DH && (c.lane & 15) == 0 && !ST) {
            // the 16 lanes' codes of this pass share one line per column quad; they are dead now: keep them out of DRAM
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) discard_l2(dhp + (size_t)((col0 >> 2) + j4) * TM);
        }
        if (KIND != K_BWD0 && c.signal) signal_kb(c.kb_bar(), c16, c.lane);
    }
}

// B2: the backward half of the split training program as its own instantiation (program 21.. / 17.. only): without the forward
// epilogues the kernel is a third of the code, and the register allocation of the second-order sweep is not disturbed by them
// (with the prologue below compiled into the one-launch program, that kernel's spills went from 148 to 204 bytes and the
// backward-only launch from 2.11 to 2.89 ms per patch)
// (B2 = 1: with the radiance part, program 21..40; B2 = 2: without it -- NeuS pass A -- program 17..31: has_rad is a compile-time
//  constant in both, so neither carries the other's prologue)
template <bool FULL, bool ST, bool BW, int B2 = 0>
__global__ void __launch_bounds__(THREADS, 1)
mlp_tmem_kernel(const __grid_constant__ EvalJob job, const float* __restrict__ pk, const __grid_constant__ PackF32 L, const unsigned char* __restrict__ wimg,
                const float* __restrict__ unscale, const Program prog, unsigned char* __restrict__ scratch, const __grid_constant__ SpinCtx sc) {
    extern __shared__ unsigned char smem_raw_[];
    Smem& S = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw_) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // ST (training stash): the weight ring is one stage shorter and the fifth 32 KB slot is the staging area of the TMA stores
    constexpr unsigned NSK = ST ? (unsigned)NS - 1u : (unsigned)NS;
    const bool explicit_pts = job.x != nullptr;
    const long long total = explicit_pts ? job.m
                          : (long long)(job.n_rows_dev ? min(*job.n_rows_dev, job.n_rows) : job.n_rows) * job.P;
    const long long n_tiles = (total + TM - 1) / TM;
    // nothing for this CTA (an upsampling iteration whose device-side work list is empty launches the full grid: at beta = 0.1 twelve
    // of the fourteen sampler rounds of a chunk): leave before the barrier / TMEM / table set-up
    if ((long long)blockIdx.x >= n_tiles) { diag_count(sc, 0); diag_count(sc, 1); diag_count(sc, 2); return; }

    if (tid == 0) {
        for (int s = 0; s < (int)NSK; ++s) { mbar_init(smem_u32(&S.full_bar[s]), 1); mbar_init(smem_u32(&S.empty_bar[s]), 1); }
        for (int k = 0; k < 4; ++k) mbar_init(smem_u32(&S.d_ready[k]), 1);
        for (int k = 0; k < 4; ++k) mbar_init(smem_u32(&S.kb_ready[k]), EPI_THREADS / 32);   // one arrive per epilogue warp
        fence_barrier_init();
    }
    diag_count(sc, 0);
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
    diag_count(sc, 1);

    if (warp == 0) {
        // ================= weight producer =================
        if (lane == 0) {
            unsigned it = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int g = prog.g0; g < prog.n_gemm; ++g) {
                    const unsigned char* src = wimg + prog.g[g].w_off;
                    const unsigned sb = prog.g[g].stage_bytes;
                    const int n_kb = prog.g[g].n_kb, n_sp = prog.g[g].prods == 3 ? 2 : 1;
                    // stage sequence: [hi(kb) | lo(kb)] per K-block; with corr_first the hi stages of a three-product GEMM follow once more
#ifndef NA_TM_ORDERS
                    const int order = 0;
#else
                    const int order = n_sp == 2 ? prog.corr_first : 0;
#endif
                    const int n_seq = order == 0 ? n_kb * n_sp : (order == 1 ? 3 * n_kb : 3 * n_kb - 1);
                    for (int q = 0; q < n_seq; ++q, ++it) {
                        // stage index in the image: 2 kb + {0 hi, 1 lo}
                        int st;
                        if (order == 0) st = n_sp == 2 ? q : 2 * q;
                        else if (order == 1) st = q < 2 * n_kb ? q : 2 * (q - 2 * n_kb);
                        else { const int na = 2 * (n_kb - 1); st = q < na ? q : (q < na + n_kb - 1 ? 2 * (q - na) : 2 * (n_kb - 1) + (q - na - (n_kb - 1))); }
                        const unsigned slot = it % NSK, ph = (it / NSK) & 1;
                        mbar_wait_guarded(smem_u32(&S.empty_bar[slot]), ph ^ 1, &sc, 0x50000000u | ((unsigned)g << 8) | slot);
                        mbar_expect_tx(smem_u32(&S.full_bar[slot]), sb);
                        bulk_g2s(smem_u32(S.Wst + slot * STAGE_BYTES), src + (size_t)st * sb, sb, smem_u32(&S.full_bar[slot]));
                    }
                }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // the whole warp walks the program (converged control flow, waits included); one elected lane issues tcgen05.mma / commit
        unsigned it = 0, a_phase = 0;
        NA_CYC(long long t_a = 0; long long t_full = 0; const long long t_tot0 = clock64();)
        const unsigned wst = smem_u32(S.Wst);
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (int g = prog.g0; g < prog.n_gemm; ++g) {
#ifdef NA_TM_TRACE
                long long* const tr = (job.dbg && blockIdx.x == 0 && tile == (long long)gridDim.x && lane == 0) ? job.dbg + 16 : nullptr;
#endif
                const int n_kb = prog.g[g].n_kb, prods = prog.g[g].prods;
                const unsigned idesc = prog.g[g].n64 ? IDESC_N64 : IDESC_N256;
                const unsigned t_in = tmem_d + (unsigned)(g & 1) * 256u, t_out = tmem_d + (unsigned)((g + 1) & 1) * 256u;
                // A GEMM with fewer than four K-blocks (the 39-wide encoding inputs, n_kb == 1) still gets all four kb_ready barriers
                // signalled by its producer stage (tile-input stage / write_vbar0, all four at once).  Those phases must be consumed
                // BEFORE this GEMM's MMAs are issued: once D is committed the epilogue starts signalling the same barriers for the
                // next GEMM, and a second completion before this warp's wait flips the parity back -- the wait would then never
                // return (mbarrier phase aliasing; this was the intermittent first-step stall, profiles/r3a_stall_root_cause.md).
                if (g == 26) NA_TRACE_X(tr, 8);
                for (int kb = n_kb; kb < 4; ++kb) mbar_wait_guarded(smem_u32(&S.kb_ready[kb]), a_phase, &sc, 0x4d040000u | ((unsigned)g << 8) | (unsigned)kb);
                if (g == 26) NA_TRACE_X(tr, 9);
#ifndef NA_TM_ORDERS
                const int order = 0;          // orders 1 / 2 are a build option (-DNA_TM_ORDERS): their code costs 2.4 % of the SDF-only tile even unused
#else
                const int order = prods == 3 ? prog.corr_first : 0;
#endif
                const bool corr_first = order != 0;
                const int n_a = order == 1 ? n_kb : n_kb - 1;          // K-blocks whose corrections / hi*hi go through phases A / B
                if (corr_first) {
                    // ---- phase A: lo*hi and hi*lo of K-blocks 0..n_a-1, as the K-blocks of A arrive (stages [W_hi | W_lo])
                    for (int kb = 0; kb < n_a; ++kb) {
                        const unsigned slot0 = it % NSK, ph0 = (it / NSK) & 1, slot1 = (it + 1) % NSK, ph1 = ((it + 1) / NSK) & 1;
                        const unsigned a_hi = t_in + (unsigned)(kb * 4) * 16u;
                        { NA_CYC(const long long t0 = clock64();)
                          mbar_wait_guarded(smem_u32(&S.full_bar[slot0]), ph0, &sc, 0x4d010000u | ((unsigned)g << 8) | (unsigned)kb);
                          mbar_wait_guarded(smem_u32(&S.full_bar[slot1]), ph1, &sc, 0x4d020000u | ((unsigned)g << 8) | (unsigned)kb);
                          NA_CYC(t_full += clock64() - t0;) }
                        { NA_CYC(const long long t0 = clock64();) mbar_wait_guarded(smem_u32(&S.kb_ready[kb]), a_phase, &sc, 0x4d030000u | ((unsigned)g << 8) | (unsigned)kb); NA_CYC(t_a += clock64() - t0;) }
                        tc_fence_after();
                        NA_TRACE_M(tr, g, kb, 0);
                        if (elect_one()) {
                            const unsigned long long bd0 = umma_desc(wst + slot0 * STAGE_BYTES), bd1 = umma_desc(wst + slot1 * STAGE_BYTES);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) umma_f16_ts(t_out, a_hi + 16u * ks + 8u, bd0 + 2 * ks, idesc, (kb | ks) != 0);     // lo * hi
                            umma_commit(smem_u32(&S.empty_bar[slot0]));
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) umma_f16_ts(t_out, a_hi + 16u * ks, bd1 + 2 * ks, idesc, 1);                       // hi * lo
                            umma_commit(smem_u32(&S.empty_bar[slot1]));
                        }
                        __syncwarp();
                        it += 2;
                        NA_TRACE_M(tr, g, kb, 1);
                    }
                    // ---- phase B: hi*hi of those K-blocks on top (W_hi stages again); order 1: the last one in N-parts with their own commits
                    for (int kb = 0; kb < n_a; ++kb) {
                        const unsigned slot0 = it % NSK, ph0 = (it / NSK) & 1;
                        const unsigned a_hi = t_in + (unsigned)(kb * 4) * 16u;
                        { NA_CYC(const long long t0 = clock64();)
                          mbar_wait_guarded(smem_u32(&S.full_bar[slot0]), ph0, &sc, 0x4d050000u | ((unsigned)g << 8) | (unsigned)kb);
                          NA_CYC(t_full += clock64() - t0;) }
                        tc_fence_after();
                        if (elect_one()) {
                            const unsigned long long bd0 = umma_desc(wst + slot0 * STAGE_BYTES);
                            if (kb + 1 < n_kb || prog.g[g].n64) {
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) umma_f16_ts(t_out, a_hi + 16u * ks, bd0 + 2 * ks, idesc, 1);
                            } else {
                                const int nsp = prog.nsplit;
                                const unsigned ncol = 256u / (unsigned)nsp, idn = nsp == 4 ? IDESC_N64 : IDESC_N128;
                                for (int np = 0; np < nsp; ++np) {
                                    const unsigned t_dn = t_out + ncol * np;
                                    const unsigned long long b0 = bd0 + (unsigned long long)(np * (int)(ncol * 128u / 16u));
#pragma unroll
                                    for (int ks = 0; ks < 4; ++ks) umma_f16_ts(t_dn, a_hi + 16u * ks, b0 + 2 * ks, idn, 1);
                                    if (nsp == 4) umma_commit(smem_u32(&S.d_ready[np]));
                                    else { umma_commit(smem_u32(&S.d_ready[2 * np])); umma_commit(smem_u32(&S.d_ready[2 * np + 1])); }
                                }
                            }
                            umma_commit(smem_u32(&S.empty_bar[slot0]));
                        }
                        __syncwarp();
                        it += 1;
                    }
                }
                if (order == 2) {
                    // ---- phase C: the last K-block: lo*hi, hi*lo, hi*hi per N-part, each part with its own commit
                    const int kb = n_kb - 1;
                    const unsigned slot0 = it % NSK, ph0 = (it / NSK) & 1, slot1 = (it + 1) % NSK, ph1 = ((it + 1) / NSK) & 1;
                    const unsigned a_hi = t_in + (unsigned)(kb * 4) * 16u;
                    { NA_CYC(const long long t0 = clock64();)
                      mbar_wait_guarded(smem_u32(&S.full_bar[slot0]), ph0, &sc, 0x4d010000u | ((unsigned)g << 8) | (unsigned)kb);
                      mbar_wait_guarded(smem_u32(&S.full_bar[slot1]), ph1, &sc, 0x4d020000u | ((unsigned)g << 8) | (unsigned)kb);
                      NA_CYC(t_full += clock64() - t0;) }
                    { NA_CYC(const long long t0 = clock64();) mbar_wait_guarded(smem_u32(&S.kb_ready[kb]), a_phase, &sc, 0x4d030000u | ((unsigned)g << 8) | (unsigned)kb); NA_CYC(t_a += clock64() - t0;) }
                    tc_fence_after();
                    NA_TRACE_M(tr, g, kb, 0);
                    if (elect_one()) {
                        const unsigned long long bd0 = umma_desc(wst + slot0 * STAGE_BYTES), bd1 = umma_desc(wst + slot1 * STAGE_BYTES);
                        const bool whole = prog.g[g].n64;                      // 64-wide GEMM: one part, d_ready committed after the loop
                        const int nsp = whole ? 1 : prog.nsplit;
                        const unsigned ncol = 256u / (unsigned)nsp, idn = whole ? idesc : (nsp == 4 ? IDESC_N64 : IDESC_N128);
                        for (int np = 0; np < nsp; ++np) {
                            const unsigned t_dn = t_out + (whole ? 0u : ncol * np);
                            const unsigned long long boff = whole ? 0ull : (unsigned long long)(np * (int)(ncol * 128u / 16u));
                            const unsigned long long b0 = bd0 + boff, b1 = bd1 + boff;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) umma_f16_ts(t_dn, a_hi + 16u * ks + 8u, b0 + 2 * ks, idn, (kb | ks) != 0);        // lo * hi
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) umma_f16_ts(t_dn, a_hi + 16u * ks, b1 + 2 * ks, idn, 1);                          // hi * lo
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) umma_f16_ts(t_dn, a_hi + 16u * ks, b0 + 2 * ks, idn, 1);                          // hi * hi
                            if (!whole) {
                                if (nsp == 4) umma_commit(smem_u32(&S.d_ready[np]));
                                else { umma_commit(smem_u32(&S.d_ready[2 * np])); umma_commit(smem_u32(&S.d_ready[2 * np + 1])); }
                            }
                        }
                        umma_commit(smem_u32(&S.empty_bar[slot0]));
                        umma_commit(smem_u32(&S.empty_bar[slot1]));
                    }
                    __syncwarp();
                    it += 2;
                    NA_TRACE_M(tr, g, kb, 1);
                }
                for (int kb = 0; kb < (corr_first ? 0 : n_kb); ++kb) {
                    const unsigned slot0 = it % NSK, ph0 = (it / NSK) & 1;
                    const unsigned slot1 = (it + 1) % NSK, ph1 = ((it + 1) / NSK) & 1;
                    const unsigned a_hi = t_in + (unsigned)(kb * 4) * 16u;
                    // the weights first (the ring runs K-blocks ahead, so these return at once), then the A operand: the MMAs go out
                    // right behind the epilogue's signal
                    { NA_CYC(const long long t0 = clock64();)
                      mbar_wait_guarded(smem_u32(&S.full_bar[slot0]), ph0, &sc, 0x4d010000u | ((unsigned)g << 8) | (unsigned)kb);
                      if (prods == 3) mbar_wait_guarded(smem_u32(&S.full_bar[slot1]), ph1, &sc, 0x4d020000u | ((unsigned)g << 8) | (unsigned)kb);
                      NA_CYC(t_full += clock64() - t0;) }
                    { NA_CYC(const long long t0 = clock64();) mbar_wait_guarded(smem_u32(&S.kb_ready[kb]), a_phase, &sc, 0x4d030000u | ((unsigned)g << 8) | (unsigned)kb); NA_CYC(t_a += clock64() - t0;) }
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
                a_phase ^= 1;
                if (prog.g[g].n64) {
                    if (elect_one()) { for (int k = 0; k < 4; ++k) umma_commit(smem_u32(&S.d_ready[k])); }
                    __syncwarp();
                }
            }
        NA_CYC(if (job.dbg && blockIdx.x == 0 && lane == 0) { job.dbg[0] = clock64() - t_tot0; job.dbg[1] = t_a; job.dbg[2] = t_full; })
    } else if (warp >= FIRST_EPI_WARP) {
        // ================= epilogue warps =================
        const int q = warp & 3, cq = (warp - FIRST_EPI_WARP) >> 2;
        const int r = 32 * q + lane;                       // sample row == TMEM lane
        unsigned char* sp = scratch + (size_t)blockIdx.x * SCRATCH_BYTES;
        EpiCtx c;
        c.S = &S; c.dh = reinterpret_cast<uint2*>(sp);
        c.pk = pk; c.L = &L; c.job = &job; c.t_lane = tmem_d + ((unsigned)(32 * q) << 16); c.r = r; c.cq = cq;
        c.sdim = small_dim(job.multires_view);
        c.s_base = smem_u32(&S); c.lane = lane; c.signal = 0; c.need_lo = 1;
        c.d_phase = 0;
        NA_CYC(long long t_d = 0; const long long t_e0 = clock64(); c.t_wait = &t_d;)
        NA_CYC(c.trace = nullptr;)
        c.st_row = nullptr; c.st_mpad = (int)job.st_mpad;
        c.st_m = 0; c.st_cnt = 0; c.st_row0 = 0;
        c.st_stg = smem_u32(S.Wst) + (unsigned)(NS - 1) * STAGE_BYTES + (unsigned)(warp - FIRST_EPI_WARP) * 2048u;
        c.rs = 1.f; c.irs = 1.f; c.nbar[0] = c.nbar[1] = c.nbar[2] = 0.f; c.lyr = 0; c.has_rad = B2 == 1 ? 1 : (B2 == 2 ? 0 : (job.rad != nullptr));
        c.bw = BW ? 1 : 0;
        c.qp = reinterpret_cast<uint4*>(sp + DH_BYTES + FEAT_BYTES + MISC_BYTES);
        c.mk = reinterpret_cast<unsigned long long*>(sp + DH_BYTES + FEAT_BYTES + MISC_BYTES + QP_BYTES + GP_BYTES) + (tid - 32 * FIRST_EPI_WARP);
        const bool has_rad = B2 == 1 ? true : (B2 == 2 ? false : job.rad != nullptr);

        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#ifdef NA_TM_TRACE
            c.trace = (job.dbg && blockIdx.x == 0 && tile == (long long)gridDim.x && tid == 32 * FIRST_EPI_WARP) ? job.dbg + 16 : nullptr;
#endif
            if (ST && job.bw_split != 0) {
                // split training program: the softplus' codes and ReLU masks of a tile outlive the forward launch
                unsigned char* tb = job.tile_buf + (size_t)tile * TILE_BUF_BYTES;
                c.dh = reinterpret_cast<uint2*>(tb);
                c.mk = reinterpret_cast<unsigned long long*>(tb + DH_BYTES) + (tid - 32 * FIRST_EPI_WARP);
            }
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
                // padding rows of the last tile are written too (zero upstream gradient): the weight-gradient kernels read whole tiles
                c.st_m = w;
                c.st_row = (ST && (size_t)w < job.st_mpad) ? job.st_wide + (size_t)w * 256 : nullptr;
                c.st_row0 = (int)(tile * TM) + 32 * q;
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
                if (BW) {
                    // upstream gradients of the row and its power-of-two scale rs: max(|.|) * rs in [1, 2)
                    float gs = 0.f, gn[3] = {0.f, 0.f, 0.f}, gr[3] = {0.f, 0.f, 0.f};
                    if (w < total) {
                        if (job.bw_gsdf) gs = job.bw_gsdf[w];
                        if (job.bw_gnab) { gn[0] = job.bw_gnab[w * 3]; gn[1] = job.bw_gnab[w * 3 + 1]; gn[2] = job.bw_gnab[w * 3 + 2]; }
                        if (job.bw_grad) { gr[0] = job.bw_grad[w * 3]; gr[1] = job.bw_grad[w * 3 + 1]; gr[2] = job.bw_grad[w * 3 + 2]; }
                    }
                    const float mx = fmaxf(fmaxf(fabsf(gs), fmaxf(fabsf(gn[0]), fmaxf(fabsf(gn[1]), fabsf(gn[2])))),
                                           0.25f * fmaxf(fabsf(gr[0]), fmaxf(fabsf(gr[1]), fabsf(gr[2]))));
                    int ex = 0;
                    if (mx > 0.f && mx < 3.0e38f) frexpf(mx, &ex);                       // mx = m * 2^ex, m in [0.5, 1)
                    ex = max(-100, min(100, ex));
                    c.rs = ldexpf(1.f, 1 - ex); c.irs = ldexpf(1.f, ex - 1);
                    if (cq == 0) {
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) { S.BWV[cc * TM + r] = gn[cc] * c.rs; S.BWV[(3 + cc) * TM + r] = gr[cc] * c.rs; }
                        S.BWV[6 * TM + r] = gs;                                          // masked and scaled at the sdf head (g == 7)
                    }
                }
                if (ST && job.bw_split != 2 && c.st_row && job.st_emb) {               // 64 columns: zero beyond the 39 entries
                    uint4 lo, hi;
                    pack16(e, 1.f / ACT_SCALE, lo, hi);
                    uint4* erow = reinterpret_cast<uint4*>(job.st_emb + (size_t)w * ST_NLD + 16 * cq);
                    __stcs(erow, lo); __stcs(erow + 1, hi);
                }
                if (!B2) store_a16(c.t_lane + (unsigned)(16 * cq), e, 1);
            }
            if (!B2) { for (int k = 0; k < 4; ++k) signal_kb(c.kb_bar(), k, lane); }
#ifndef NA_NO_INPUT_BAR
            epi_bar_sync();                                   // EMBS / X / V / BWV of this tile: written above by other warps, read from GEMM 3 on
#endif

            float sdf_part = 0.f, rgb_part[3] = {0.f, 0.f, 0.f};
            float small_in[36];
            // BW: A <- v-bar_0 (K-block 0 of the second-order sweep's first GEMM), over this thread's own (consumed) columns of the
            // previous D; also the narrow v-bar_0 stash plane
            auto write_vbar0 = [&](unsigned t_region) {
                float e[16];
                NA_TRACE_X(c.trace, 0);
                vbar0_quarter<0>(S, cq, r, c.nbar, e);
                NA_TRACE_X(c.trace, 1);
                if (c.st_row && job.st_vb0) {
                    uint4 lo, hi;
                    pack16(e, c.irs, lo, hi);
                    uint4* vrow = reinterpret_cast<uint4*>(job.st_vb0 + (size_t)c.st_m * ST_NLD + 16 * cq);
                    __stcs(vrow, lo); __stcs(vrow + 1, hi);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) e[j] *= ACT_SCALE;
                NA_TRACE_X(c.trace, 2);
                store_a16(t_region + (unsigned)(16 * cq), e, 0);
                NA_TRACE_X(c.trace, 3);
                for (int k = 0; k < 4; ++k) signal_kb(c.kb_bar(), k, lane);
                NA_TRACE_X(c.trace, 4);
            };
            // BW: A <- delta_3 = (delta_4 W4) * [ys_4 > 0] (K-blocks 0..3 of radiance-backward GEMM 21), over this thread's own (consumed)
            // columns of D of GEMM 20; delta_4 (x rs) is in BWV[3..5]
            auto delta3_stage = [&](unsigned t_region) {
                epi_bar_sync();                                   // delta_4 of every row
                const float d4[3] = {S.BWV[3 * TM + r], S.BWV[4 * TM + r], S.BWV[5 * TM + r]};
                const unsigned long long m64 = c.mk[(size_t)3 * EPI_THREADS];
#pragma unroll 1
                for (int c16 = 0; c16 < 4; ++c16) {
                    const int col0 = c16 * 64 + cq * 16;
                    float o[16];
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 w0 = lds128(c.w4_s() + (unsigned)(col0 + 4 * j4) * 4u), w1 = lds128(c.w4_s() + (unsigned)(256 + col0 + 4 * j4) * 4u),
                                     w2 = lds128(c.w4_s() + (unsigned)(512 + col0 + 4 * j4) * 4u);
                        o[4 * j4] = d4[0] * w0.x + d4[1] * w1.x + d4[2] * w2.x; o[4 * j4 + 1] = d4[0] * w0.y + d4[1] * w1.y + d4[2] * w2.y;
                        o[4 * j4 + 2] = d4[0] * w0.z + d4[1] * w1.z + d4[2] * w2.z; o[4 * j4 + 3] = d4[0] * w0.w + d4[1] * w1.w + d4[2] * w2.w;
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) o[j] = ((m64 >> (16 * c16 + j)) & 1ull) ? o[j] * ACT_SCALE : 0.f;
                    stash16(c, ST_D + 3, col0, o, c.irs * (1.f / ACT_SCALE));
                    store_a16(t_region + col0, o, 0);               // GEMM 21 is a single-product GEMM: no lo words
                    signal_kb(c.kb_bar(), c16, lane);
                }
            };
            if (B2) {
                // ---- backward half of the split program: what the tails of GEMMs 7 and 20 do in the one-launch program, from the
                // forward launch's outputs: the masked d L / d sdf (flag in st_t1[.][1]) and delta_4 = d L / d radiance * rgb (1 - rgb)
                if (cq == 0) {
                    const bool live = c.st_row != nullptr && S.OIDX[r] >= 0;
                    float gs = S.BWV[6 * TM + r];
                    if (c.st_row && job.st_t1) {
                        float* t1 = job.st_t1 + (size_t)c.st_m * 4;
                        if (job.bw_bg_mask && __ldcg(t1 + 1) != 0.f) gs = 0.f;
                        *reinterpret_cast<float4*>(t1) = make_float4(gs, 0.f, 0.f, 0.f);
                    }
                    S.BWV[6 * TM + r] = gs * c.rs;
                    if (live && has_rad) {
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) {
                            const float rgb = __ldcg(job.rad + S.OIDX[r] * 3 + cc);
                            S.BWV[(3 + cc) * TM + r] *= rgb * (1.f - rgb);
                        }
                    }
                    if (has_rad && c.st_row && job.st_t0)
                        *reinterpret_cast<float4*>(job.st_t0 + (size_t)c.st_m * 4) =
                            live ? make_float4(S.BWV[3 * TM + r] * c.irs, S.BWV[4 * TM + r] * c.irs, S.BWV[5 * TM + r] * c.irs, 0.f)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (has_rad) {
                    delta3_stage(c.t_lane + (unsigned)(prog.g0 & 1) * 256u);     // (its leading barrier also publishes BWV[6])
                } else {
                    // no radiance part (NeuS pass A): the second-order sweep starts here with the eikonal gradient alone
                    epi_bar_sync();                                               // BWV[6] of every row
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) c.nbar[cc] = S.BWV[cc * TM + r];
                    write_vbar0(c.t_lane + (unsigned)(prog.g0 & 1) * 256u);
                }
            }
            for (int g = prog.g0; g < prog.n_gemm; ++g) {
                const int op = BW ? (int)prog.g[g].op : (int)OP_FWD;
                c.g = g; c.us = unscale[prog.g[g].img];          // us = 2^-(weight shift) / ACT_SCALE
                c.us *= prog.g[g].debias;
                c.lyr = prog.g[g].lyr;
                c.signal = g + 1 < prog.n_gemm;
                c.need_lo = c.signal ? (prog.g[g + 1].prods == 3) : 0;
                c.nx = (FULL && ST && c.signal) ? scratch_planes_of(BW ? (int)prog.g[g + 1].op : (int)OP_FWD, (int)prog.g[g + 1].lyr, g + 1, c.has_rad) : 0x00ffffffu;
                if (BW && (op == OP_HB || (op == OP_FWD && g == 20))) c.signal = 0;     // the next A operand is written by the tail below
                const unsigned t_dd = c.t_lane + (unsigned)((g + 1) & 1) * 256u;       // D of this GEMM == A of the next
                // program order: 0..7 fwd | 8 feat | 9..15 bwd 7..1 | 16 bwd 0 | 17..20 radiance
                if (BW && op != OP_FWD) {
                    if (B2 != 2 && op == OP_DR) {
                        epi_gemm<K_DR, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                    } else if (B2 != 2 && op == OP_FB) {
                        epi_gemm<K_FB, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                    } else if (B2 != 2 && op == OP_HB) {
                        epi_gemm<K_HB, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                    } else if (op == OP_SO) {
                        epi_gemm<K_SO, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                    } else {
                        epi_gemm<K_TR, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                    }
                } else if (B2) {
                    // (backward-only instantiation: the program holds no forward GEMM)
                } else if (g < 8) {
                    if (ST) epi_gemm<K_FWDX, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                    else if (g == 3) epi_gemm<K_FWD3, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                    else if (g == 7) epi_gemm<K_FWD7, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                    else epi_gemm<K_FWD, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                } else if (g == 8) {
                    epi_gemm<K_FEAT, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                } else if (FULL) {
                    if (g <= 15) {
                        if (ST) epi_gemm<K_BWDX, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                        else if (g == 12) epi_gemm<K_BWD4, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
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
                        if (ST && c.st_row && job.st_small) {                           // 64 columns, zero beyond the 36 inputs
                            float part[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) part[j] = 0.f;
#pragma unroll
                            for (int j = 0; j < 36; ++j) if (j >> 4 == cq) part[j & 15] = small_in[j];
                            uint4 lo, hi;
                            pack16(part, 1.f / ACT_SCALE, lo, hi);
                            uint4* srow = reinterpret_cast<uint4*>(job.st_small + (size_t)c.st_m * ST_NLD + 16 * cq);
                            __stcs(srow, lo); __stcs(srow + 1, hi);
                        }
                        epi_gemm<K_RAD0, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                    } else if (ST) {
                        epi_gemm<K_RADX, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                    } else if (g == 20) {
                        epi_gemm<K_RAD3, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                    } else {
                        epi_gemm<K_RAD, FULL, ST>(c, t_dd, sdf_part, rgb_part, small_in);
                    }
                }
                c.d_phase ^= 1;
                // ---- per-GEMM tails ------------------------------------------------------------------------
                if (!B2 && g == 7) {
                    // fwd layer 7 stored h8 x16: undo in the head.  sdf = <h8, W8[0]> + b8[0]
                    S.PART[cq * TM + r] = sdf_part * (1.f / ACT_SCALE); sdf_part = 0.f;
                    epi_bar_sync();
                    if (cq == 0) {
                        float sdf = S.PART[r] + S.PART[TM + r] + S.PART[2 * TM + r] + S.PART[3 * TM + r] + __ldg(pk + L.b8_sdf);
                        if (ST && !BW && job.bw_split == 1 && c.st_row && job.st_t1) {
                            // forward half of the split program: leave the sphere-background flag (volsdf.py:349-357: where
                            // R - |x| < sdf the network's sdf is not the output) for the backward half, next to its d L / d sdf slot
                            const float x0 = S.X[r], x1 = S.X[TM + r], x2 = S.X[2 * TM + r];
                            const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x1, x1)), __fmul_rn(x2, x2)));
                            *reinterpret_cast<float4*>(job.st_t1 + (size_t)c.st_m * 4) = make_float4(0.f, (job.bound_r - nrm < sdf) ? 1.f : 0.f, 0.f, 0.f);
                        }
                        if (job.apply_bg) {
                            const float x0 = S.X[r], x1 = S.X[TM + r], x2 = S.X[2 * TM + r];
                            const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x1, x1)), __fmul_rn(x2, x2)));
                            sdf = fminf(sdf, job.bound_r - nrm);
                        }
                        if (S.OIDX[r] >= 0 && job.sdf) job.sdf[S.OIDX[r]] = sdf;
                        if (BW) {
                            // sphere-background override (volsdf.py:349-357): where R - |x| < sdf the network's sdf is not the output
                            float gs = S.BWV[6 * TM + r];
                            if (job.bw_bg_mask) {
                                const float x0 = S.X[r], x1 = S.X[TM + r], x2 = S.X[2 * TM + r];
                                const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x1, x1)), __fmul_rn(x2, x2)));
                                if (job.bound_r - nrm < sdf) gs = 0.f;
                            }
                            S.BWV[6 * TM + r] = gs * c.rs;
                            if (c.st_row && job.st_t1)
                                *reinterpret_cast<float4*>(job.st_t1 + (size_t)c.st_m * 4) = make_float4(gs, 0.f, 0.f, 0.f);
                        }
                    }
                    if (BW) epi_bar_sync();                   // BWV[6] (masked d L / d sdf) is read by every column quarter in the trunk
                }
                if (!B2 && FULL && g == 16) {
                    if (has_rad) {
                        // A <- geometry feature (x16) for radiance layer 0, written over this thread's own (consumed) columns of D of
                        // GEMM 16; signalled at once, so radiance GEMM 0 runs under the nabla arithmetic below
                        float4 fn[4];
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) fn[j4] = c.featp()[(size_t)(cq * 4 + j4) * TM + r];
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
                                for (int j4 = 0; j4 < 4; ++j4) fn[j4] = c.featp()[(size_t)(((col0 + 64) >> 2) + j4) * TM + r];
                            }
                            store_a16(t_dd + col0, h, c.need_lo);
                            if ((lane & 7) == 0) {
#pragma unroll
                                for (int j4 = 0; j4 < 4; ++j4) discard_l2(c.featp() + (size_t)((col0 >> 2) + j4) * TM + r);
                            }
                            signal_kb(c.kb_bar(), c16, lane);
                        }
                    }
                    epi_bar_sync();                                       // all 39 d sdf/d emb entries complete
                    if (cq < 3) {
                        // nabla component cq in closed form: one column quarter per coordinate
                        const int cc = cq;
                        const float xc = S.X[cc * TM + r];
                        float n = c.misc()[cc * TM + r];
#pragma unroll
                        for (int f = 0; f < 6; ++f) {
                            const float fr = (float)(1 << f);
                            float sn, cs;
                            if (STASH) { sn = S.EMBS[(3 + 6 * f + cc) * TM + r] * (1.f / ACT_SCALE); cs = S.EMBS[(6 + 6 * f + cc) * TM + r] * (1.f / ACT_SCALE); }
                            else sincosf(__fmul_rn(xc, fr), &sn, &cs);
                            n += fr * (c.misc()[(3 + 6 * f + cc) * TM + r] * cs - c.misc()[(6 + 6 * f + cc) * TM + r] * sn);
                        }
                        S.PART[cc * TM + r] = n;
                    }
                    epi_bar_sync();                                       // PART[0..2] = nabla of every row
                    if (cq == 0) {
                        const long long oo = S.OIDX[r];
                        if (oo >= 0 && job.nab) { job.nab[oo * 3] = S.PART[r]; job.nab[oo * 3 + 1] = S.PART[TM + r]; job.nab[oo * 3 + 2] = S.PART[2 * TM + r]; }
                    }
                    if (BW && !has_rad && g + 1 < prog.n_gemm) {
                        // no radiance part (NeuS pass A): the second-order sweep starts here with the eikonal gradient alone
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) c.nbar[cc] = S.BWV[cc * TM + r];
                        write_vbar0(t_dd);
                    }
                }
                if (BW && op == OP_DR && c.lyr == 0) {
                    // d L / d nabla through the radiance net: sum the four column quarters, add the eikonal part
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) { S.PART[(cq * 3 + cc) * TM + r] = rgb_part[cc] * (1.f / ACT_SCALE); rgb_part[cc] = 0.f; }
                    epi_bar_sync();
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc)
                        c.nbar[cc] = S.PART[cc * TM + r] + S.PART[(3 + cc) * TM + r] + S.PART[(6 + cc) * TM + r] + S.PART[(9 + cc) * TM + r] + S.BWV[cc * TM + r];
                }
                if (BW && op == OP_HB) write_vbar0(t_dd);
                if (!B2 && FULL && g == 20 && op == OP_FWD) {
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) { S.PART[(cq * 3 + cc) * TM + r] = rgb_part[cc]; rgb_part[cc] = 0.f; }
                    epi_bar_sync();
                    if (cq == 0 && S.OIDX[r] >= 0) {
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) {
                            // radiance layer 3 stored relu x16: undo in the head
                            const float z = (S.PART[cc * TM + r] + S.PART[(3 + cc) * TM + r] + S.PART[(6 + cc) * TM + r] + S.PART[(9 + cc) * TM + r])
                                            * (1.f / ACT_SCALE) + __ldg(pk + L.rad_b4 + cc);
                            const float rgb = __fdiv_rn(1.f, 1.f + expf(-z));
                            job.rad[S.OIDX[r] * 3 + cc] = rgb;
                            if (BW) S.BWV[(3 + cc) * TM + r] *= rgb * (1.f - rgb);          // delta_4 (x rs) = d L/d radiance * sigmoid'
                        }
                        if (BW && c.st_row && job.st_t0)
                            *reinterpret_cast<float4*>(job.st_t0 + (size_t)c.st_m * 4) =
                                make_float4(S.BWV[3 * TM + r] * c.irs, S.BWV[4 * TM + r] * c.irs, S.BWV[5 * TM + r] * c.irs, 0.f);
                    } else if (BW && cq == 0 && c.st_row && job.st_t0) {
                        // padding row of the last tile: the weight-gradient kernels read whole tiles, its delta_4 is zero
                        *reinterpret_cast<float4*>(job.st_t0 + (size_t)c.st_m * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (BW && g + 1 < prog.n_gemm) delta3_stage(t_dd);
                }
                if (g + 1 >= prog.n_gemm) {
                    tc_fence_before();
                    epi_bar_sync();                           // X / OIDX / PART and TMEM region 0 are rewritten by the next tile's input stage
                }
            }
        }
        if (ST) stash_flush(c);
        NA_CYC(if (job.dbg && blockIdx.x == 0 && tid == 32 * FIRST_EPI_WARP) { job.dbg[3] = clock64() - t_e0; job.dbg[4] = t_d; })
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_d, 512); }
    diag_count(sc, 2);
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
    T.total = T.meta_off + 2048;
    return T;
}

__global__ void absmax_kernel(const float* __restrict__ pk, const size_t* __restrict__ offs, const int* __restrict__ rows, float* __restrict__ out) {
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

}  // namespace tm

size_t mlp_tmem_image_bytes() { return tm::image_layout().total; }

// `image` = this kernel's region of the packed buffer; d_offs / d_rows / d_absmax are the tables of the 21 forward planes tc_pack()
// left on the device; the 5 backward planes (images 21..25) live in the train pack (`train_off` floats from pk_f32)
int tmem_pack(const float* pk_f32, const size_t* d_offs, const int* d_rows, const float* d_absmax, size_t train_off, const PackTrain& TP,
              unsigned char* image, cudaStream_t stream) {
    using namespace tm;
    const ImageLayout T = image_layout();
    int meta[N_PLANES * 4];
    for (int g = 0; g < N_PLANES; ++g) { meta[g * 4] = T.n_kb[g]; meta[g * 4 + 1] = T.N[g]; meta[g * 4 + 2] = (int)(T.w_off[g] / 16); meta[g * 4 + 3] = 0; }
    unsigned char* mbase = image + T.meta_off;
    int* d_meta = (int*)mbase; size_t* offs = (size_t*)(mbase + 512); int* rows = (int*)(mbase + 768); float* absmax = (float*)(mbase + 1024);
    NA_TRY(upload_small(d_meta, meta, sizeof(meta), stream));
    NA_TRY(check_cuda(cudaMemcpyAsync(offs, d_offs, 21 * sizeof(size_t), cudaMemcpyDeviceToDevice, stream)));
    NA_TRY(check_cuda(cudaMemcpyAsync(rows, d_rows, 21 * sizeof(int), cudaMemcpyDeviceToDevice, stream)));
    NA_TRY(check_cuda(cudaMemcpyAsync(absmax, d_absmax, 21 * sizeof(float), cudaMemcpyDeviceToDevice, stream)));
    const size_t offs_bw[5] = {train_off + TP.rad_w[3], train_off + TP.rad_w[2], train_off + TP.rad_w[1], train_off + TP.rad_w[0], train_off + TP.w8_feat};
    const int rows_bw[5] = {256, 256, 256, 256, 256};
    NA_TRY(upload_small(offs + 21, offs_bw, sizeof(offs_bw), stream));
    NA_TRY(upload_small(rows + 21, rows_bw, sizeof(rows_bw), stream));
    absmax_kernel<<<5, 256, 0, stream>>>(pk_f32, offs + 21, rows + 21, absmax + 21);
    NA_CHECK_LAUNCH();
    pack_kernel<<<dim3(32, N_PLANES), 256, 0, stream>>>(pk_f32, offs, rows, d_meta, absmax, image, (float*)(image + T.unscale_off));
    NA_CHECK_LAUNCH();
    return NA_OK;
}

size_t mlp_tmem_scratch_bytes(int grid) { return (size_t)grid * tm::SCRATCH_BYTES; }
size_t mlp_tmem_tile_buf_bytes(long long n_samples) { return (size_t)((n_samples + tm::TM - 1) / tm::TM) * tm::TILE_BUF_BYTES; }

extern long long* g_tc_dbg;

// mixed != 0: feature head, reverse sweep and radiance layers use the hi*hi product only
int launch_mlp_tmem(const EvalJob& job_, const float* pk_f32, const unsigned char* image, const PackF32& L, int mixed,
                    unsigned char* scratch, size_t scratch_bytes, cudaStream_t stream) {
    using namespace tm;
    EvalJob job = job_; job.dbg = g_tc_dbg;
    static bool attr_done[64] = {false};
    // NA_TM_PRODS: diagnostics override, a 21-character string of '1'/'3' (products per GEMM of the program)
    static const char* prods_env = getenv("NA_TM_PRODS");
    const size_t smem = sizeof(Smem) + 1024;
    if (first_on_device(attr_done)) {
        NA_TRY(check_cuda(cudaFuncSetAttribute(mlp_tmem_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        NA_TRY(check_cuda(cudaFuncSetAttribute(mlp_tmem_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        NA_TRY(check_cuda(cudaFuncSetAttribute(mlp_tmem_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        NA_TRY(check_cuda(cudaFuncSetAttribute(mlp_tmem_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        NA_TRY(check_cuda(cudaFuncSetAttribute(mlp_tmem_kernel<true, true, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        NA_TRY(check_cuda(cudaFuncSetAttribute(mlp_tmem_kernel<true, true, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    }
    const long long total = job.x ? job.m : (long long)job.n_rows * job.P;
    if (total <= 0) return NA_OK;
    const ImageLayout T = image_layout();
    Program prog; prog.n_gemm = 0; prog.g0 = 0;
    static const char* nsplit_env = getenv("NA_TM_NSPLIT");                   // diagnostics: "2" = N-halves (the r1n scheme), default quarters
    prog.nsplit = (nsplit_env && nsplit_env[0] == '2') ? 2 : 4;
    // NA_TM_ORDER=0: K-block-interleaved products (the r2 scheme).  NA_TM_DEBIAS=<x>: experimental multiplicative compensation of the
    // accumulate truncation of three-product GEMMs, D *= 1 + x * 2^-24 (0 = off)
    // NA_TM_ORDER=0|1|2: product order of the three-product GEMMs (see Program).  NA_TM_DEBIAS=<x>: the accumulate-truncation bias of a
    // 256-deep three-product GEMM is compensated by D *= 1 + x * 2^-24 (scaled by K-blocks / 4 for shallower GEMMs); defaults are the
    // calibrations of profiles/r3b_tc_accumulation.md for each order; 0 switches the compensation off
    static const char* order_env = getenv("NA_TM_ORDER");
    static const char* debias_env = getenv("NA_TM_DEBIAS");
#ifdef NA_TM_ORDERS
    prog.corr_first = order_env ? (order_env[0] == '0' ? 0 : (order_env[0] == '1' ? 1 : 2)) : 0;
#else
    (void)order_env; prog.corr_first = 0;
#endif
    static const float debias_default[3] = {14.f, 4.f, 8.f};      // (1 + x 2^-24 is representable in steps of 2)
    const float debias_x = debias_env ? (float)atof(debias_env) : debias_default[prog.corr_first];
    const int last = !job.want_full ? (job.feat ? 8 : 7) : (job.rad ? 20 : 16);
    for (int g = 0; g <= last; ++g) {
        Gemm t; t.w_off = T.w_off[g]; t.stage_bytes = (unsigned)T.N[g] * 128u; t.n_kb = (unsigned char)T.n_kb[g];
        t.prods = (unsigned char)((mixed && g >= 8) ? 1 : 3);
        if (prods_env && (int)strlen(prods_env) > g) t.prods = prods_env[g] == '1' ? 1 : 3;
        t.n64 = T.N[g] == 64; t.img = (unsigned char)g; t.op = OP_FWD; t.lyr = 0; t.pad0 = t.pad1 = 0;
        t.debias = t.prods == 3 ? 1.f + debias_x * 5.9604645e-8f * (float)t.n_kb * 0.25f : 1.f;
        prog.g[prog.n_gemm++] = t;
    }
    if (job.bw) {
        // backward program of a training patch (needs the stash): single-product operands, rows scaled per sample
        if (!job.st_wide || !job.want_full || !STASH) return NA_ERR_BAD_ARG;
        auto add = [&](int img, int op, int lyr) {
            Gemm t; t.w_off = T.w_off[img]; t.stage_bytes = (unsigned)T.N[img] * 128u; t.n_kb = (unsigned char)T.n_kb[img];
            t.prods = 1; t.n64 = 0; t.img = (unsigned char)img; t.op = (unsigned char)op; t.lyr = (unsigned char)lyr; t.pad0 = t.pad1 = 0; t.debias = 1.f;
            prog.g[prog.n_gemm++] = t;
        };
        if (job.rad) {
            add(21, OP_DR, 2); add(22, OP_DR, 1); add(23, OP_DR, 0);     // delta_2, delta_1, delta_0
            add(24, OP_FB, 0); add(25, OP_HB, 0);                         // feat-bar; feature part of h-bar_7
        }
        for (int l = 0; l < 8; ++l) add(l, OP_SO, l);                     // second-order sweep
        for (int l = 6; l >= 0; --l) add(15 - l, OP_TR, l);               // trunk: z-bar_l from z-bar_{l+1} W_{l+1}
        if (job.bw_split == 2) {
            // backward half of the split program: the forward launch (bw_split == 1) left the stash planes, tile_buf and (with the
            // radiance part) rad; without it (NeuS pass A: sdf + nabla only) the program starts at the second-order sweep
            if (!job.tile_buf || !job.st_t1) return NA_ERR_BAD_ARG;
            prog.g0 = last + 1;                                          // 21 | 17
        } else if (job.bw_split != 0) return NA_ERR_BAD_ARG;
    } else if (job.bw_split != 0) {
        if (job.bw_split != 1 || !job.st_wide || !job.tile_buf || !job.want_full || !STASH) return NA_ERR_BAD_ARG;
    }
    long long tiles = (total + TM - 1) / TM;
    int grid = (int)(tiles < (long long)num_sms() ? tiles : (long long)num_sms());
    if (scratch_bytes < mlp_tmem_scratch_bytes(grid)) return NA_ERR_WORKSPACE;
    const float* usc = (const float*)(image + T.unscale_off);
    if (job.st_wide && !job.want_full) return NA_ERR_BAD_ARG;
    const SpinCtx sc = diag_next(DK_MLP_TMEM, grid);
    if (job.bw && prog.g0 > 0 && job.rad)
                            mlp_tmem_kernel<true, true, true, 1><<<grid, THREADS, smem, stream>>>(job, pk_f32, L, image, usc, prog, scratch, sc);
    else if (job.bw && prog.g0 > 0)
                            mlp_tmem_kernel<true, true, true, 2><<<grid, THREADS, smem, stream>>>(job, pk_f32, L, image, usc, prog, scratch, sc);
    else if (job.bw)        mlp_tmem_kernel<true, true, true><<<grid, THREADS, smem, stream>>>(job, pk_f32, L, image, usc, prog, scratch, sc);
    else if (job.st_wide && job.bw_split == 1)
                            mlp_tmem_kernel<true, true, false><<<grid, THREADS, smem, stream>>>(job, pk_f32, L, image, usc, prog, scratch, sc);
    else if (job.st_wide)   return NA_ERR_UNSUPPORTED;               // the stash is written by the training programs only
    else if (job.want_full) mlp_tmem_kernel<true, false, false><<<grid, THREADS, smem, stream>>>(job, pk_f32, L, image, usc, prog, scratch, sc);
    else                    mlp_tmem_kernel<false, false, false><<<grid, THREADS, smem, stream>>>(job, pk_f32, L, image, usc, prog, scratch, sc);
    NA_CHECK_LAUNCH();
    return NA_OK;
}

int preload_mlp_tmem() {
    NA_PRELOAD((tm::mlp_tmem_kernel<false, false, false>));
    NA_PRELOAD((tm::mlp_tmem_kernel<true, false, false>));
    NA_PRELOAD((tm::mlp_tmem_kernel<true, true, true>));
    NA_PRELOAD((tm::mlp_tmem_kernel<true, true, false>));
    NA_PRELOAD((tm::mlp_tmem_kernel<true, true, true, 1>));
    NA_PRELOAD((tm::mlp_tmem_kernel<true, true, true, 2>));
    NA_PRELOAD(tm::pack_kernel);
    NA_PRELOAD(tm::absmax_kernel);
    return NA_OK;
}

}  // namespace na
