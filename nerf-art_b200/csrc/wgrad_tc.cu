// Weight-gradient GEMMs of a training patch on tcgen05 (sm_100a):  out[l][r] += sum_m L[m][l] * R[m][r]  (+ a second (L, R) pair),
// for the 256 x 256 tasks whose operands are wide stash planes in the quad layout (csrc/common.cuh: stash_quad_index).
//
// One CTA = (task, sample split).  The contraction runs over samples, so both operands arrive "sample-major" and have to be
// turned into K-major tensor-core operands on the way into shared memory:
//   * 16 loader warps read the planes as they lie in HBM (one float4 = 4 columns of one sample; a warp instruction covers 32
//     consecutive samples of one column quad = 512 contiguous bytes), two stages of loads in flight per thread, and scatter each
//     float4 as four 4-byte stores into the UMMA K-major SWIZZLE_128B image (row = column l or r, 32 samples = 128 bytes per row;
//     the 32 lanes of a store hit the 32 banks of one row);
//   * a stage is 32 samples: A0 / A1 = columns 0..127 / 128..255 of L (2 x 16 KB), B = all 256 columns of R (32 KB); 3 stages;
//   * one elected lane issues tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = 256, K = 8; fp32 words read as TF32), two
//     accumulators [128 x 256] fp32 = all 512 TMEM columns, so every operand byte is read from HBM exactly once per task;
//   * the loader warps then read the accumulators (tcgen05.ld) and add them into the packed gradient with red.global.add.f32.
// The kernel is HBM-bound by construction (64 KB of operands per 1024 tensor-pipe cycles); the legacy mma.sync version it
// replaces (csrc/train.cu: wgrad_tf32_kernel) stays for the narrow tasks and as NA_WGRAD=mma / fp32.
#include "common.cuh"

namespace na {
namespace wg {

constexpr int THREADS = 576;                 // warp 0: MMA issuer / TMEM owner; warp 1: idle; warps 2..17: loaders, then epilogue
constexpr int LOADERS = 512;
constexpr int NST = 3;
constexpr int KS = 32;                       // samples per stage
constexpr int A_BYTES = 128 * 128, B_BYTES = 256 * 128, STAGE_BYTES = 2 * A_BYTES + B_BYTES;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
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
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr) {
    return (unsigned long long)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// D = f32, A = B = tf32, K-major, N = 256, M = 128
constexpr unsigned IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

struct __align__(1024) Smem {
    unsigned char st[NST][STAGE_BYTES];          // [A0 | A1 | B]
    unsigned long long full[NST], empty[NST], acc_ready;
    unsigned tmem_base;
};

struct Task { const float* L[2]; const float* R[2]; float* out; int ldo; int npair; int pad; };
constexpr int MAX_TASKS = 16;
struct Table { Task t[MAX_TASKS]; int n; int tiles_per_split; long long tiles; };

// element k (sample within the stage) of operand row `row`: 16-byte chunk k/4, XOR-swizzled with row % 8
__device__ __forceinline__ unsigned op_addr(unsigned base, int row, int k) {
    return base + (unsigned)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 2) ^ (row & 7)) << 4) | ((k & 3) << 2)));
}

__global__ void __launch_bounds__(THREADS, 1)
wgrad_tc_kernel(const Table tab, const SpinCtx sc) {
    extern __shared__ unsigned char smem_raw_[];
    Smem& S = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw_) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const Task t = tab.t[blockIdx.x];
    const long long tile0 = (long long)blockIdx.y * tab.tiles_per_split;
    long long tile1 = tile0 + tab.tiles_per_split; if (tile1 > tab.tiles) tile1 = tab.tiles;
    const int n_stage_pair = (int)(tile1 - tile0) * (128 / KS);            // stages per (L, R) pair
    const int n_stage = n_stage_pair * t.npair;

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(smem_u32(&S.full[s]), LOADERS / 32); mbar_init(smem_u32(&S.empty[s]), 1); }
        mbar_init(smem_u32(&S.acc_ready), 1);
        fence_barrier_init();
    }
    diag_count(sc, 0);
    if (warp == 0) tmem_alloc(smem_u32(&S.tmem_base), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_d = S.tmem_base;
    diag_count(sc, 1);
    if (n_stage <= 0) {                                                    // uniform: nothing to do for this split
        __syncthreads();
        if (warp == 0) tmem_dealloc(tmem_d, 512);
        diag_count(sc, 2);
        return;
    }

    if (warp == 0) {
        // ---- MMA issuer: per stage 4 K-steps x 2 accumulators
        for (int i = 0; i < n_stage; ++i) {
            const unsigned slot = i % NST, ph = (i / NST) & 1;
            mbar_wait_guarded(smem_u32(&S.full[slot]), ph, sc, 0x4d000000u | (unsigned)(i & 0xffffff));
            tc_fence_after();
            if (elect_one()) {
                const unsigned base = smem_u32(S.st[slot]);
                const unsigned long long a0 = umma_desc(base), a1 = umma_desc(base + A_BYTES), b = umma_desc(base + 2 * A_BYTES);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    umma_tf32_ss(tmem_d, a0 + 2 * ks, b + 2 * ks, IDESC, (i | ks) != 0);
                    umma_tf32_ss(tmem_d + 256u, a1 + 2 * ks, b + 2 * ks, IDESC, (i | ks) != 0);
                }
                umma_commit(smem_u32(&S.empty[slot]));
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(smem_u32(&S.acc_ready));
        __syncwarp();
    } else if (warp >= 2) {
        // ---- loaders.  A stage = 32 samples x (64 quads of L + 64 quads of R) = 128 (quad, 32-sample) groups; loader warp w takes
        // groups w, w + 16, ...: 8 float4 per thread per stage (lane = sample), kept in registers one stage ahead of the scatter
        const int lw = warp - 2;
        float4 cur[8], nxt[8];
        auto gload = [&](int i, float4 (&v)[8]) {
            const int pr = i / n_stage_pair, is = i - pr * n_stage_pair;
            const long long m = (tile0 * 128) + (long long)is * KS + lane;
            const float* Lp = t.L[pr]; const float* Rp = t.R[pr];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int grp = lw + 16 * j;                               // 0..63: quad of L, 64..127: quad of R
                const float* P = grp < 64 ? Lp : Rp;
                v[j] = __ldcs(reinterpret_cast<const float4*>(P + stash_quad_index(m, 4 * (grp & 63))));
            }
        };
        gload(0, cur);
        for (int i = 0; i < n_stage; ++i) {
            if (i + 1 < n_stage) gload(i + 1, nxt);
            const unsigned slot = i % NST, ph = (i / NST) & 1;
            mbar_wait_guarded(smem_u32(&S.empty[slot]), ph ^ 1, sc, 0x4c000000u | (unsigned)(i & 0xffffff));
            const unsigned base = smem_u32(S.st[slot]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int grp = lw + 16 * j;
                const int row0 = 4 * (grp & 63);                           // operand row of the quad's first column
                // L columns 0..127 -> A0, 128..255 -> A1 (row - 128), R columns -> B
                const unsigned ob = grp < 64 ? (row0 < 128 ? base : base + A_BYTES) : base + 2 * A_BYTES;
                const int r0 = grp < 64 ? (row0 & 127) : row0;
                const float vv[4] = {cur[j].x, cur[j].y, cur[j].z, cur[j].w};
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    asm volatile("st.shared.f32 [%0], %1;" :: "r"(op_addr(ob, r0 + c, lane)), "f"(vv[c]) : "memory");
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&S.full[slot]));
#pragma unroll
            for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
        }
        // ---- epilogue: warp lw reads TMEM lanes 32 * (warp % 4) .., columns 128 * (lw / 4) .. + 128 of the 512
        mbar_wait_guarded(smem_u32(&S.acc_ready), 0, sc, 0x45000000u);
        tc_fence_after();
        const int q = warp & 3, cs = lw >> 2;                              // cs 0,1 -> accumulator 0 (l < 128), 2,3 -> accumulator 1
        const int l = 32 * q + lane + (cs >= 2 ? 128 : 0);
        float* orow = t.out + (size_t)l * t.ldo;
#pragma unroll 1
        for (int c16 = 0; c16 < 8; ++c16) {
            unsigned v[16];
            const int col = cs * 128 + c16 * 16;                           // TMEM column
            tmem_ld16(tmem_d + ((unsigned)(32 * q) << 16) + (unsigned)col, v);
            tmem_wait_ld();
            const int r = col & 255;
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(orow + r + j, __uint_as_float(v[j]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_d, 512); }
    diag_count(sc, 2);
}

}  // namespace wg

// tasks: 256 x 256 outputs (ldo floats per output row), operands = wide stash planes in the quad layout; m_tiles = 128-sample tiles
int launch_wgrad_tc(const WgTcTask* tasks, int n_tasks, long long m_tiles, cudaStream_t stream) {
    using namespace wg;
    if (n_tasks <= 0 || m_tiles <= 0) return NA_OK;
    if (n_tasks > MAX_TASKS) return NA_ERR_UNSUPPORTED;
    static thread_local bool attr_set = false;
    const size_t smem = sizeof(Smem) + 1024;
    if (!attr_set) {
        NA_TRY(check_cuda(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        attr_set = true;
    }
    Table tab; tab.n = n_tasks; tab.tiles = m_tiles;
    for (int i = 0; i < n_tasks; ++i) {
        Task& t = tab.t[i];
        t.L[0] = tasks[i].L[0]; t.L[1] = tasks[i].L[1]; t.R[0] = tasks[i].R[0]; t.R[1] = tasks[i].R[1];
        t.out = tasks[i].out; t.ldo = tasks[i].ldo; t.npair = tasks[i].npair; t.pad = 0;
    }
    int splits = (num_sms() + n_tasks - 1) / n_tasks;                      // one CTA per SM
    if (splits > m_tiles) splits = (int)m_tiles;
    tab.tiles_per_split = (int)((m_tiles + splits - 1) / splits);
    splits = (int)((m_tiles + tab.tiles_per_split - 1) / tab.tiles_per_split);
    wgrad_tc_kernel<<<dim3(n_tasks, splits), THREADS, smem, stream>>>(tab, diag_next(DK_WGRAD_TC, n_tasks * splits));
    NA_CHECK_LAUNCH();
    return NA_OK;
}

int preload_wgrad_tc() {
    NA_PRELOAD(wg::wgrad_tc_kernel);
    return NA_OK;
}

}  // namespace na
