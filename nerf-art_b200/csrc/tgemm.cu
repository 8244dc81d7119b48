// tcgen05 TF32 GEMM for the skinny linear layers of the CLIP ViT-B/32 image tower (csrc/clip_vit.cu), sm_100a only.
//
//   C[M,N] = epilogue(A[M,K] * Bt[N,K]^T)        A: fp32 row-major activations, Bt: a pre-tiled TF32 weight image
//
// The tower runs at M = 50 * batch <= 700 rows against 88 M frozen weights, so every GEMM is a weight stream: the kernel is
// organised around keeping many bytes of B in flight, not around MMA throughput.
//   * B never needs a layout transform at run time: na_clip_pack_weights stores each weight (in both orientations, for the
//     forward y = x W^T and the backward-data dx = dy W) as 8 KB tiles of 64 rows x 32 k, already rounded to TF32 (cvt.rna) and
//     laid out as the shared-memory image of the UMMA K-major SWIZZLE_128B operand; one elected lane streams the tiles with
//     cp.async.bulk (mbarrier complete_tx) through a 6-stage ring.
//   * A (128 rows x 32 k per stage) is copied global -> shared by four loader warps with 16-byte cp.async straight into the
//     swizzled positions (row r, chunk c -> (r/8)*1024 + (r%8)*128 + ((c ^ r%8) << 4)), four stages ahead
//     (cp.async.wait_group + fence.proxy.async + mbarrier arrive); rows beyond M are zero-filled.  tcgen05 reads fp32 words as
//     TF32 (low 13 mantissa bits ignored).
//   * One elected lane issues tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = 64, K = 8; 4 per stage); the accumulator
//     [128 x 64] fp32 lives in TMEM (64 columns).  tcgen05.commit releases the stage / signals the epilogue.
//   * The loader warps then turn into the epilogue: tcgen05.ld (thread = row), + bias, raw pre-activation copy, QuickGELU,
//     + residual, fp32 stores; with split-K (gridDim.z > 1, used when N alone gives too few CTAs to pull HBM bandwidth) the
//     partial tiles are accumulated with red.global.add into a zeroed C and split 0 adds bias and residual.
#include "common.cuh"

namespace na {
namespace tg {

constexpr int BM = 128, BN = 64, BK = 32;          // BK fp32 = 128 bytes = one swizzle row
constexpr int NST = 6, AHEAD = 4;
constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128;
constexpr int THREADS = 192;

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
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
// D[tmem] (+)= A[smem] * B[smem]^T, TF32 operands (fp32 words, low 13 mantissa bits ignored), fp32 accumulate
__device__ __forceinline__ void umma_tf32_ss(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
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
// K-major SWIZZLE_128B UMMA shared-memory descriptor (same encoding as csrc/mlp_tmem.cu): 1024 B between 8-row groups
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr) {
    return (unsigned long long)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = f32 (bit 4), A = B = tf32 (format 2 @7 and @10), both K-major, N>>3 @17, M>>4 @24
constexpr unsigned IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(BN >> 3) << 17) | ((unsigned)(BM >> 4) << 24);

struct __align__(1024) Smem {
    unsigned char A[NST][A_BYTES];
    unsigned char B[NST][B_BYTES];
    unsigned long long full[NST], empty[NST], acc_ready;
    unsigned tmem_base;
};

__global__ void __launch_bounds__(THREADS, 1)
tgemm_kernel(const float* __restrict__ A, const unsigned char* __restrict__ Bimg, float* __restrict__ C, int M, int N, int K,
             const float* __restrict__ bias, float* __restrict__ C_raw, int act, const float* __restrict__ R, int kb_per_split, const __grid_constant__ SpinCtx sc) {
    extern __shared__ unsigned char smem_raw_[];
    Smem& S = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw_) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tile = blockIdx.x, m0 = blockIdx.y * BM;
    const int nkb_total = K / BK;
    const int kb0 = blockIdx.z * kb_per_split;
    const int nkb = min(nkb_total, kb0 + kb_per_split) - kb0;

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(smem_u32(&S.full[s]), 128 + 1); mbar_init(smem_u32(&S.empty[s]), 1); }
        mbar_init(smem_u32(&S.acc_ready), 1);
        fence_barrier_init();
    }
    diag_count(sc, 0);
    if (warp == 1) tmem_alloc(smem_u32(&S.tmem_base), 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_d = S.tmem_base;
    diag_count(sc, 1);

    if (warp == 0) {
        // ---- weight producer: one 8 KB tile per stage
        if (lane == 0) {
            const unsigned char* src = Bimg + ((size_t)n_tile * nkb_total + kb0) * B_BYTES;
            for (int i = 0; i < nkb; ++i) {
                const unsigned slot = i % NST, ph = (i / NST) & 1;
                mbar_wait_guarded(smem_u32(&S.empty[slot]), ph ^ 1, &sc, 0x50000000u | (unsigned)i);
                mbar_expect_tx(smem_u32(&S.full[slot]), B_BYTES);
                bulk_g2s(smem_u32(S.B[slot]), src + (size_t)i * B_BYTES, B_BYTES, smem_u32(&S.full[slot]));
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer
        for (int i = 0; i < nkb; ++i) {
            const unsigned slot = i % NST, ph = (i / NST) & 1;
            mbar_wait_guarded(smem_u32(&S.full[slot]), ph, &sc, 0x4d000000u | (unsigned)i);
            tc_fence_after();
            if (elect_one()) {
                const unsigned long long ad = umma_desc(smem_u32(S.A[slot])), bd = umma_desc(smem_u32(S.B[slot]));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_tf32_ss(tmem_d, ad + 2 * ks, bd + 2 * ks, IDESC, (i | ks) != 0);
                umma_commit(smem_u32(&S.empty[slot]));
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(smem_u32(&S.acc_ready));
        __syncwarp();
    } else {
        // ---- A loaders (thread = row of the tile), then epilogue (thread = TMEM lane = the same row)
        const int q = warp & 3, row = 32 * q + lane;
        const int gm = m0 + row;
        const bool live = gm < M;
        const float* arow = A + (size_t)(live ? gm : 0) * K + (size_t)kb0 * BK;
        const unsigned dst_row = (unsigned)((row >> 3) * 1024 + (row & 7) * 128);
        for (int i = 0; i < nkb + AHEAD; ++i) {
            if (i < nkb) {
                const unsigned slot = i % NST, ph = (i / NST) & 1;
                mbar_wait_plain(smem_u32(&S.empty[slot]), ph ^ 1);
                const unsigned dst = smem_u32(S.A[slot]) + dst_row;
                if (live) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) cp_async16(dst + (unsigned)((c ^ (row & 7)) << 4), arow + (size_t)i * BK + c * 4);
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" :: "r"(dst + (unsigned)(c << 4)), "f"(0.f) : "memory");
                }
            }
            cp_async_commit();
            if (i >= AHEAD) {
                cp_async_wait<AHEAD>();
                fence_proxy_async();
                mbar_arrive(smem_u32(&S.full[(i - AHEAD) % NST]));
            }
        }
        mbar_wait_plain(smem_u32(&S.acc_ready), 0);
        tc_fence_after();
        const unsigned t_lane = tmem_d + ((unsigned)(32 * q) << 16);
        const bool split = gridDim.z > 1, lead = blockIdx.z == 0;
#pragma unroll 1
        for (int c16 = 0; c16 < BN / 16; ++c16) {
            unsigned v[16];
            tmem_ld16(t_lane + (unsigned)(c16 * 16), v);
            tmem_wait_ld();
            if (!live) continue;
            const int n = n_tile * BN + c16 * 16;
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]);
            if (bias && lead) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n) + j4);
                    o[4 * j4] += b4.x; o[4 * j4 + 1] += b4.y; o[4 * j4 + 2] += b4.z; o[4 * j4 + 3] += b4.w;
                }
            }
            float* crow = C + (size_t)gm * N + n;
            if (C_raw) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4)
                    reinterpret_cast<float4*>(C_raw + (size_t)gm * N + n)[j4] = make_float4(o[4 * j4], o[4 * j4 + 1], o[4 * j4 + 2], o[4 * j4 + 3]);
            }
            if (act) {
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] = o[j] * __fdiv_rn(1.f, 1.f + expf(-1.702f * o[j]));      // QuickGELU (clip/model.py)
            }
            if (R && lead) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 r4 = reinterpret_cast<const float4*>(R + (size_t)gm * N + n)[j4];
                    o[4 * j4] += r4.x; o[4 * j4 + 1] += r4.y; o[4 * j4 + 2] += r4.z; o[4 * j4 + 3] += r4.w;
                }
            }
            if (split) {
#pragma unroll
                for (int j = 0; j < 16; ++j) atomicAdd(crow + j, o[j]);
            } else {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) reinterpret_cast<float4*>(crow)[j4] = make_float4(o[4 * j4], o[4 * j4 + 1], o[4 * j4 + 2], o[4 * j4 + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_d, 64); }
    diag_count(sc, 2);
}

// src: fp32 [rows][cols] row-major.  transpose == 0: operand Bt[n][k] = src[n][k] (N = rows, K = cols);
// transpose == 1: Bt[n][k] = src[k][n] (N = cols, K = rows).  Image: tile (n/64, k/32) at ((n/64) * (K/32) + k/32) * 8 KB.
__global__ void pack_tiles_kernel(const float* __restrict__ src, int rows, int cols, int transpose, unsigned char* __restrict__ img) {
    const int N = transpose ? cols : rows, K = transpose ? rows : cols;
    const size_t total = (size_t)N * K;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        // consecutive threads walk the source row-major (coalesced reads)
        const int r = (int)(idx / cols), c = (int)(idx % cols);
        const int n = transpose ? c : r, k = transpose ? r : c;
        float w = src[idx];
        unsigned bits;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(bits) : "f"(w));
        const int rr = n & 63, e = k & 31;
        const size_t tile = (size_t)(n >> 6) * (K >> 5) + (k >> 5);
        const unsigned off = (unsigned)((rr >> 3) * 1024 + (rr & 7) * 128 + (((e >> 2) ^ (rr & 7)) << 4) + (e & 3) * 4);
        *reinterpret_cast<unsigned*>(img + tile * B_BYTES + off) = bits;
    }
}

}  // namespace tg

int tgemm_pack(const float* src, int rows, int cols, int transpose, unsigned char* img, cudaStream_t stream) {
    const int N = transpose ? cols : rows, K = transpose ? rows : cols;
    if (N % tg::BN || K % tg::BK) return NA_ERR_UNSUPPORTED;
    tg::pack_tiles_kernel<<<592, 256, 0, stream>>>(src, rows, cols, transpose, img);
    NA_CHECK_LAUNCH();
    return NA_OK;
}

// C (+)= epilogue(A[M,K] * image^T).  C must not alias A.  Split-K is chosen here; it needs a linear epilogue (no activation,
// no raw copy) and zeroes C first.
int tgemm(const float* A, const unsigned char* img, float* C, int M, int N, int K, const float* bias, float* raw, int act,
          const float* R, cudaStream_t stream) {
    using namespace tg;
    if (N % BN || K % BK || M <= 0) return NA_ERR_UNSUPPORTED;
    static bool attr_done[64] = {false};
    const size_t smem = sizeof(Smem) + 1024;
    if (first_on_device(attr_done)) {
        NA_TRY(check_cuda(cudaFuncSetAttribute(tgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    }
    const int n_tiles = N / BN, m_tiles = (M + BM - 1) / BM, nkb = K / BK;
    int splits = 1;
    if (!act && !raw) {
        const int ctas = n_tiles * m_tiles;
        splits = (num_sms() + ctas - 1) / ctas;                      // aim at one CTA per SM ...
        if (splits > nkb / 8) splits = nkb / 8;                      // ... with at least 8 K-blocks (64 KB of weights) per split
        if (splits < 1) splits = 1;
    }
    const int kb_per_split = (nkb + splits - 1) / splits;
    splits = (nkb + kb_per_split - 1) / kb_per_split;
    if (splits > 1) NA_TRY(check_cuda(cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), stream)));
    tgemm_kernel<<<dim3(n_tiles, m_tiles, splits), THREADS, smem, stream>>>(A, img, C, M, N, K, bias, raw, act, R, kb_per_split, diag_next(DK_TGEMM, n_tiles * m_tiles * splits));
    NA_CHECK_LAUNCH();
    return NA_OK;
}

int preload_tgemm() {
    NA_PRELOAD(tg::tgemm_kernel);
    NA_PRELOAD(tg::pack_tiles_kernel);
    return NA_OK;
}

}  // namespace na
