// GraspPointCNN forward on the 5th-generation tensor cores (reference scripts/utils/ml_grasp_optimizer/model.py:102-128).
//
// The six 3x3 convolutions run as bf16 implicit GEMMs: tcgen05.mma (cta_group::1, kind::f16, M=128), fp32 accumulators
// in TMEM, operands brought to shared memory by cp.async.bulk + mbarrier, epilogue (bias + ReLU [+ 2x2 max-pool] + bf16
// pack) from TMEM with tcgen05.ld.  Warp roles: one producer, four MMA-issuing warps (one elected thread each, each owning
// one of the item's four accumulator tiles), eight epilogue warps.  Attention / average / MLP tail: cnn_tail_kernel (fp32),
// which also takes the last max-pool.
//
// Activation layout ("plane-major, shared-halo flat"): a feature map of C channels on S x S pixels for n patches is
// C/8 planes; plane p holds, for every flat position, the 8 channels 8p..8p+7 as one 16-byte unit.  Positions are
//     q = patch * PP + r * pitch + c,     pitch = S + 1,  PP = (S + 1)^2,  pixel (y, x) sits at (r, c) = (y + 1, x + 1)
// Row 0 and column 0 of every patch are zero; they are the left/top halo of that patch and, because the layout is
// flat, also the right halo of the previous row and the bottom halo of the previous patch.  A 3x3 convolution is
// then out[q] = sum_taps W[tap] . in[q + (ky-1)*pitch + (kx-1)] for EVERY q - a 1-D stencil over the flat array -
// so an output tile of 128 consecutive positions needs, for tap (ky, kx), the 128 consecutive input rows starting
// ky*pitch + kx further on.  With the unswizzled K-major UMMA layout (8 rows x 16 B core matrices, rows 16 B
// apart, SBO = 128 B, LBO = plane stride) that is the same shared-memory tile with the descriptor start address
// moved by 16 B per position: the input tile is loaded ONCE per 64 channels and reused by all nine taps.
// Outputs at halo positions are computed and discarded (6 % / 11 % / 21 % of the rows at 32 / 16 / 8 px) and
// written as zeros, which makes the output directly the next layer's input.  LEAD zero rows precede position 0.
//
// Layers followed by a 2x2 max-pool (the second conv of a block at 32 and 16 px) use BLOCKED tiles instead: with the
// stride between groups of 8 rows (SBO) set to one image row (pitch * 16 B) instead of 128 B, the 128 rows of an MMA
// tile are an 8-column x 16-row block of pixels.  A warp of the epilogue then holds 8 columns x 4 rows, the four pixels
// of every pooling window sit in lanes l, l^1, l^8, l^9, and the pool is two shuffles on the packed bf16 values: the
// un-pooled activation (the largest tensor of the network) is never written, and no halo row is computed.
//
// Weights: bf16, BatchNorm folded, packed [n_split][Cin/64][tap][8 planes][NC][8] so that the B operand of one
// (channel chunk, tap) stage is one contiguous bulk copy, again unswizzled K-major (LBO = NC * 16 B).
#include <cuda_bf16.h>

#include <vector>

#include "lg_internal.cuh"

namespace {

constexpr int LEAD = 64;        // zero rows in front of every plane (>= pitch + 1)
constexpr int TILE_M = 512;     // positions per work item: 4 UMMA tiles of 128 rows
constexpr int UMMA_T = 4;
// weight stages in flight (what shared memory allows; the blocked input stage of the pooled layers is a little larger)
__host__ __device__ constexpr int nb_stages(int nc, bool pool) { return nc == 64 ? 8 : (pool ? 4 : 5); }
// layer 0 (KP = 2): one 2 KB stage per tap, so that all nine can stay resident (UmmaConvArgs::wres)
__host__ __device__ constexpr int nb_stages_kp(int kp, int nc, bool pool) { return kp == 2 ? 9 : nb_stages(nc, pool); }
constexpr int CONV_THREADS = 416;   // warp 0 producer, warp 1 MMA issuer + TMEM owner, warps 2-9 epilogue, warps 10-12 further MMA issuers
constexpr int ISSUER2_WARP = 10;
constexpr int EPI_WARPS = 8;        // two warps per TMEM lane quarter, each taking every other 32-column chunk

// Optional pipeline timers (build with -DLG_CNN_TIMING): per CTA, cycles the MMA warp spent waiting for the
// accumulators (0), the input stage (1), the weight stages (2), issuing (3), and the epilogue's wait (4) / work (5).
#ifdef LG_CNN_TIMING
__device__ unsigned long long g_cnn_timing[6][148][8];
#define LG_T0(v) const long long v = clock64()
#define LG_TACC(slot, v) t_acc[slot] += clock64() - (v)
#else
#define LG_T0(v)
#define LG_TACC(slot, v)
#endif

struct UmmaConvArgs {
    const uint4* in;      // [Cin/8 planes][R]
    uint4* out;           // [Cout/8 planes][R]
    const uint4* wt;      // packed weights
    const float* bias;    // [Cout]
    long long R;          // rows per plane
    int pitch, PP;
    const int32_t* n_dev; // device-side patch count (may be null -> n_host)
    int n_host;           // patches the launch was sized for (plane stride R, buffer sizes)
    int KC;               // input channel chunks (of KP planes)
    int n_split;          // Cout / NC
    int rows;             // shared-memory rows per plane of the A stage: TILE_M + 2 * pitch + 2, rounded up to 8
    int cout;
    int layer;            // 0..5 (timers only)
    long long Rout;       // POOL kernels (2x2 max-pool fused into the epilogue, S = 32 or 16): rows per plane of the pooled output
    int issuers;          // MMA-issuing warps (1, 2 or 4): warp 1 and warps 10.., each owning UMMA_T / issuers accumulator tiles
    int wres;             // layer 0 only (KP = 2, one channel chunk, one output split): the nine 2 KB weight stages are loaded once
                          // per CTA and stay in shared memory instead of being streamed again for every item
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
// one non-blocking look at a barrier phase: 1 = complete.  The answer takes ~100 cycles to arrive; whoever asks early and
// reads it late does not pay for them.
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Unswizzled K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
//   bits [0,14) start address >> 4, [16,30) leading byte offset >> 4 (between the two K chunks of 8 elements),
//   [32,46) stride byte offset >> 4 (between groups of 8 rows), bit 46 version.
// the two words of that descriptor: lo = start address >> 4 | LBO >> 4 << 16, hi = SBO >> 4 | version 1 (bit 46)
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr, uint32_t lbo_bytes) {
    return ((addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);      // SBO = 128 B; bit 46 of the descriptor = bit 14 of the high word

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Same wait, but naming the 32 destination registers of a load issued earlier as read-write operands: the compiler
// must then treat their values as produced here, not at the (asynchronous) load, whatever it schedules in between.
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
                   "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                   "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// One work item = TILE_M consecutive output positions x NC output channels.  KP = planes (8 channels each) per
// input-channel chunk: 8 for the 64-channel chunks of layers 1-5, 2 for the 9 (padded to 16) channels of layer 0.
template <int KP, int NC, bool POOL>
__global__ void __launch_bounds__(CONV_THREADS, 1) conv3x3_umma_kernel(UmmaConvArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int ACC_STAGES = 512 / (UMMA_T * NC);          // 2 (NC = 64) or 1 (NC = 128) accumulator sets in TMEM
    constexpr uint32_t B_STAGE = KP * NC * 16;
    constexpr int NB_STAGES = nb_stages_kp(KP, NC, POOL);
    const bool wres = KP == 2 && A.wres != 0;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NC >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_stage_bytes = (uint32_t)KP * A.rows * 16;
    unsigned char* sA = smem;
    unsigned char* sB = smem + 2 * a_stage_bytes;
    float* s_bias = reinterpret_cast<float*>(sB + NB_STAGES * B_STAGE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 256);
    // barrier indices
    uint64_t* a_full = bars;            // [2]
    uint64_t* a_empty = bars + 2;       // [2]
    uint64_t* b_full = bars + 4;        // [NB]
    uint64_t* b_empty = bars + 4 + NB_STAGES;
    uint64_t* acc_full = bars + 4 + 2 * NB_STAGES;   // [2]
    uint64_t* acc_empty = acc_full + 2;              // [2][UMMA_T]: one per accumulator tile, so that the next item's MMAs
                                                     // on tile t start as soon as the epilogue has read tile t
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2 * UMMA_T);

    // The warp index through a shuffle: the compiler then knows it is the same in all lanes, keeps everything derived from it
    // (accumulator tile, descriptors) in uniform registers and feeds tcgen05.mma from them directly - with threadIdx.x >> 5
    // every MMA is preceded by an elect / register-to-uniform "waterfall" of ~20 dependent instructions.
    const int warp = __shfl_sync(0xFFFFFFFFu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int n_patches = A.n_dev ? min(*A.n_dev, A.n_host) : A.n_host;
    const long long Q = (long long)n_patches * A.PP;                       // positions that carry data
    const int n_tiles = (int)((Q + A.pitch + 1 + TILE_M - 1) / TILE_M);      // + the zero row behind the last patch
    // blocked items: half a 32x32 patch (4 column blocks x 16 rows) or two 16x16 patches (2 column blocks each)
    const bool pool32 = POOL && A.pitch == 33;
    const int n_items = POOL ? (pool32 ? 2 * n_patches : (n_patches + 1) / 2) : n_tiles * A.n_split;
    // first position (pixel (y0, 0) of the item's first patch) of a blocked item
    auto pool_base = [&](int item) -> long long {
        return pool32 ? (long long)(item >> 1) * A.PP + (long long)(1 + 16 * (item & 1)) * A.pitch + 1
                      : (long long)(2 * item) * A.PP + A.pitch + 1;
    };

    for (int i = threadIdx.x; i < A.cout; i += CONV_THREADS) s_bias[i] = A.bias[i];
    if (threadIdx.x == 0) {
        // every issuing warp commits once to the "operand consumed" / "accumulators complete" barriers
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&a_full[i]), 1); mbar_init(smem_u32(&a_empty[i]), A.issuers); }
        for (int i = 0; i < NB_STAGES; ++i) { mbar_init(smem_u32(&b_full[i]), 1); mbar_init(smem_u32(&b_empty[i]), A.issuers); }
        for (int i = 0; i < 2; ++i) mbar_init(smem_u32(&acc_full[i]), A.issuers);
        for (int i = 0; i < 2 * UMMA_T; ++i) mbar_init(smem_u32(&acc_empty[i]), EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xFFFFFFFFu, *tmem_slot, 0);      // uniform for the compiler, like the warp index

    if (warp == 0) {
        // ===== producer: bulk copies of the input tile (once per channel chunk) and of the per-tap weights =====
        if (lane == 0) {
            int a_st = 0, a_ph = 0, b_st = 0, b_ph = 0;
            auto issue_a = [&](int item, int kc) {
                const int tile = item / A.n_split;
                mbar_wait(smem_u32(&a_empty[a_st]), a_ph ^ 1);
                const uint32_t bar = smem_u32(&a_full[a_st]);
                mbar_expect_tx(bar, a_stage_bytes);
                const long long rho0 = LEAD + (POOL ? pool_base(item) : (long long)tile * TILE_M) - A.pitch - 1;
                const uint32_t dst = smem_u32(sA + (size_t)a_st * a_stage_bytes);
#pragma unroll 1
                for (int p = 0; p < KP; ++p)
                    bulk_g2s(dst + p * A.rows * 16, A.in + ((long long)(kc * KP + p) * A.R + rho0), A.rows * 16, bar);
                a_st ^= 1;
                if (a_st == 0) a_ph ^= 1;
            };
            const int first = blockIdx.x;
            if (wres && first < n_items) {      // all nine taps' weights, once: stage = tap, each barrier completes a single time
                for (int tap = 0; tap < 9; ++tap) {
                    const uint32_t bar = smem_u32(&b_full[tap]);
                    mbar_expect_tx(bar, B_STAGE);
                    bulk_g2s(smem_u32(sB + (size_t)tap * B_STAGE), A.wt + (size_t)tap * (KP * NC), B_STAGE, bar);
                }
            }
            if (first < n_items) issue_a(first, 0);
            for (int item = first; item < n_items; item += gridDim.x) {
                const int half = item % A.n_split;
                for (int kc = 0; kc < A.KC; ++kc) {
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        // Prefetch the next input stage.  The weight loads run NB_STAGES taps ahead of the MMAs, so at this
                        // tap the MMAs have just left the previous chunk and its input stage is free: asking earlier would
                        // park this thread on a_empty while the weight pipeline behind it drains.
                        if (tap == (NB_STAGES < 8 ? NB_STAGES : 8)) {
                            if (kc + 1 < A.KC) issue_a(item, kc + 1);
                            else if (item + (int)gridDim.x < n_items) issue_a(item + gridDim.x, 0);
                        }
                        if (wres) continue;
                        mbar_wait(smem_u32(&b_empty[b_st]), b_ph ^ 1);
                        const uint32_t bar = smem_u32(&b_full[b_st]);
                        mbar_expect_tx(bar, B_STAGE);
                        const uint4* src = A.wt + ((size_t)(half * A.KC + kc) * 9 + tap) * (KP * NC);
                        bulk_g2s(smem_u32(sB + (size_t)b_st * B_STAGE), src, B_STAGE, bar);
                        if (++b_st == NB_STAGES) { b_st = 0; b_ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1 || warp >= ISSUER2_WARP) {
        // ===== MMA issuer(s) =====
        // One thread needs ~60 cycles per tcgen05.mma, an MMA keeps the tensor pipe busy for 32 (N = 64) or 64 (N = 128)
        // cycles: with two issuing warps (on different schedulers), each owning half of the item's accumulator tiles, the
        // issue cost is no longer the limiter.
        const int issuer = (warp == 1) ? 0 : warp - ISSUER2_WARP + 1;
        if (issuer >= A.issuers) goto done;
        const int t_count = UMMA_T / A.issuers, t_base = issuer * t_count;
        int a_st = 0, a_ph = 0, b_st = 0, b_ph = 0, acc_st = 0, acc_ph = 0;
        bool w_ready = false;                              // resident weights (layer 0): their nine stages have arrived
        uint32_t b_ok = 0;                                 // weight stage b_st is already known to be complete (probed one tap ahead)
        const uint32_t rows2 = 2u * (uint32_t)A.rows;      // two planes (one K = 16 step) in 16-byte units
#ifdef LG_CNN_TIMING
        long long t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const long long t_begin = clock64();
#endif
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            {
                LG_T0(t0);
                for (int tt = 0; tt < t_count; ++tt) mbar_wait(smem_u32(&acc_empty[acc_st * UMMA_T + t_base + tt]), acc_ph ^ 1);
                LG_TACC(0, t0);
            }
            tc_fence_after();
            if (KP == 2 && wres) {
                // Layer 0 with resident weights: nothing to wait for between the taps, so the item's nine MMAs per tile go out
                // back to back (a wait on an mbarrier costs ~100 cycles even when it is already complete, and this layer has
                // one MMA per tap and tile).
                if (!w_ready) {
                    LG_T0(t0);
                    for (int tap = 0; tap < 9; ++tap) mbar_wait(smem_u32(&b_full[tap]), 0);
                    LG_TACC(2, t0);
                    w_ready = true;
                }
                { LG_T0(t0); mbar_wait(smem_u32(&a_full[a_st]), a_ph); LG_TACC(1, t0); }
                tc_fence_after();
                LG_T0(t_issue);
                if (lane == 0) {
                    const uint32_t a_base = smem_u32(sA + (size_t)a_st * a_stage_bytes);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc_st * UMMA_T * NC);
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int ky = tap / 3, kx = tap - ky * 3;
                        const uint32_t a_lo = desc_lo(a_base + (uint32_t)(ky * A.pitch + kx) * 16, A.rows * 16);
                        const uint32_t b_lo = desc_lo(smem_u32(sB + (size_t)tap * B_STAGE), NC * 16);
#pragma unroll
                        for (int tt = 0; tt < UMMA_T; ++tt) {
                            if (tt >= t_count) break;
                            const int t = t_base + tt;
                            const uint64_t ad = ((uint64_t)DESC_HI << 32) | (uint64_t)(a_lo + (uint32_t)(t * 128));
                            const uint64_t bd = ((uint64_t)DESC_HI << 32) | (uint64_t)b_lo;
                            umma_bf16(d_tmem + (uint32_t)(t * NC), ad, bd, IDESC, tap != 0 ? 1u : 0u);
                        }
                    }
                    tc_commit(smem_u32(&a_empty[a_st]));
                    tc_commit(smem_u32(&acc_full[acc_st]));
                }
                __syncwarp();
                LG_TACC(3, t_issue);
                a_st ^= 1;
                if (a_st == 0) a_ph ^= 1;
                if (++acc_st == ACC_STAGES) { acc_st = 0; acc_ph ^= 1; }
                continue;
            }
            for (int kc = 0; kc < A.KC; ++kc) {
                { LG_T0(t0); mbar_wait(smem_u32(&a_full[a_st]), a_ph); LG_TACC(1, t0); }
                const uint32_t a_base = smem_u32(sA + (size_t)a_st * a_stage_bytes);
#pragma unroll 1
                for (int tap = 0; tap < 9; ++tap) {
                    { LG_T0(t0); if (!b_ok) mbar_wait(smem_u32(&b_full[b_st]), b_ph); LG_TACC(2, t0); }
                    tc_fence_after();
                    // Ask for the NEXT weight stage now: the barrier's answer travels while this tap's MMAs are issued, so
                    // the next turn of the loop starts without the round trip when the stage has already landed.
                    int nb_st = b_st + 1, nb_ph = b_ph;
                    if (nb_st == NB_STAGES) { nb_st = 0; nb_ph ^= 1; }
                    b_ok = mbar_test(smem_u32(&b_full[nb_st]), nb_ph);
                    LG_T0(t_issue);
                    if (lane == 0) {
                        // One thread must issue an MMA every 32 (N = 64) / 64 (N = 128) cycles to keep the tensor pipe
                        // busy, so the descriptors are not rebuilt per MMA: the high words are constant and the low
                        // word (start address >> 4 | LBO << 16) just advances by the operand's offset in 16-byte
                        // units (shared-memory addresses stay below 2^18, so the add never carries into the LBO field).
                        const int ky = tap / 3, kx = tap - ky * 3;
                        const uint32_t a_lo = desc_lo(a_base + (uint32_t)(ky * A.pitch + kx) * 16, A.rows * 16);
                        const uint32_t b_lo = desc_lo(smem_u32(sB + (size_t)b_st * B_STAGE), NC * 16);
                        const uint32_t first_acc = (uint32_t)((kc | tap) != 0);
                        const uint32_t desc_hi = POOL ? ((uint32_t)A.pitch | (1u << 14)) : DESC_HI;   // SBO: one image row / 128 B
                        const uint32_t d_tmem = tmem_base + (uint32_t)(acc_st * UMMA_T * NC);
#pragma unroll
                        for (int tt = 0; tt < UMMA_T; ++tt) {
                            if (tt >= t_count) break;
                            const int t = t_base + tt;
#pragma unroll
                            for (int j = 0; j < KP / 2; ++j) {
                                // first row of tile t inside the stage, in 16-byte units
                            const uint32_t t_off = !POOL ? (uint32_t)(t * 128)
                                                   : pool32 ? (uint32_t)(t * 8) : (uint32_t)((t >> 1) * A.PP + (t & 1) * 8);
                            const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + (uint32_t)(j * rows2) + t_off);
                                const uint64_t bd = ((uint64_t)DESC_HI << 32) | (uint64_t)(b_lo + (uint32_t)(j * 2 * NC));
                                umma_bf16(d_tmem + (uint32_t)(t * NC), ad, bd, IDESC, j == 0 ? first_acc : 1u);
                            }
                        }
                        tc_commit(smem_u32(&b_empty[b_st]));
                    }
                    __syncwarp();
                    LG_TACC(3, t_issue);
                    b_st = nb_st; b_ph = nb_ph;
                }
                if (lane == 0) tc_commit(smem_u32(&a_empty[a_st]));
                __syncwarp();
                a_st ^= 1;
                if (a_st == 0) a_ph ^= 1;
            }
            if (lane == 0) tc_commit(smem_u32(&acc_full[acc_st]));
            __syncwarp();
            if (++acc_st == ACC_STAGES) { acc_st = 0; acc_ph ^= 1; }
        }
#ifdef LG_CNN_TIMING
        if (warp == 1 && lane == 0 && blockIdx.x < 148) {
            for (int k = 0; k < 4; ++k) g_cnn_timing[A.layer][blockIdx.x][k] = (unsigned long long)t_acc[k];
            g_cnn_timing[A.layer][blockIdx.x][6] = (unsigned long long)(clock64() - t_begin);
        }
#endif
    } else if (warp < ISSUER2_WARP) {
        // ===== epilogue: TMEM -> registers -> bias + ReLU -> bf16 -> global (plane-major) =====
        const int wq = warp & 3;                 // TMEM lane quarter this warp may read
        const int chalf = (warp - 2) >> 2;       // which of the two warps of that quarter: chunks chalf, chalf + 2, ...
        constexpr int CHUNKS = NC / 64;          // 32-column chunks per warp and tile (1 or 2), loaded together
        if (blockIdx.x == 0) {       // zero rows in front of position 0 of every output plane
            const int et = threadIdx.x - 64;
            const int planes = A.cout / 8;
            const long long Ro = POOL ? A.Rout : A.R;
            for (int i = et; i < planes * LEAD; i += 32 * EPI_WARPS) A.out[(long long)(i / LEAD) * Ro + (i % LEAD)] = make_uint4(0, 0, 0, 0);
        }
        int acc_st = 0, acc_ph = 0;
#ifdef LG_CNN_TIMING
        long long t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int tile = item / A.n_split, half = item % A.n_split;
            { LG_T0(t0); mbar_wait(smem_u32(&acc_full[acc_st]), acc_ph); LG_TACC(4, t0); }
            tc_fence_after();
            LG_T0(t_epi);
            // bias + ReLU + bf16 pack + store of one 128-row tile held in registers
            auto store_tile = [&](int t, uint32_t (&v)[CHUNKS][32]) {
                const long long q = (long long)tile * TILE_M + t * 128 + wq * 32 + lane;
                bool data = q < Q;
                if (data) {      // positions stay below 2^31 (65536 patches x 33 x 33): 32-bit division, no 64-bit subroutine per tile
                    const unsigned ql = (unsigned)q % (unsigned)A.PP;
                    const unsigned r = ql / (unsigned)A.pitch, cc = ql - r * (unsigned)A.pitch;
                    data = (r >= 1u) && (cc >= 1u);
                }
#pragma unroll
                for (int j = 0; j < CHUNKS; ++j) {
                    const int ch = chalf + 2 * j;
                    const float* bs = s_bias + half * NC + ch * 32;
                    uint4* o = A.out + (long long)((half * NC + ch * 32) / 8) * A.R + LEAD + q;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 w = make_uint4(0, 0, 0, 0);
                        if (data) {
                            // the eight biases as two 128-bit shared-memory loads (32-byte aligned: multiples of 8 floats)
                            const float4 b0 = *reinterpret_cast<const float4*>(bs + g * 8), b1 = *reinterpret_cast<const float4*>(bs + g * 8 + 4);
                            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                            float f[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) f[e] = fmaxf(__uint_as_float(v[j][g * 8 + e]) + bb[e], 0.f);
                            w.x = pack_bf16x2(f[0], f[1]); w.y = pack_bf16x2(f[2], f[3]);
                            w.z = pack_bf16x2(f[4], f[5]); w.w = pack_bf16x2(f[6], f[7]);
                        }
                        o[(long long)g * A.R] = w;
                    }
                }
            };
            // blocked tile: bias + ReLU + bf16 pack, 2x2 max over lanes (l, l^1, l^8, l^9), store by the window's first lane
            auto store_tile_pool = [&](int t, uint32_t (&v)[CHUNKS][32]) {
                const int S = A.pitch - 1, po = (S >> 1) + 1, PPo = po * po;
                const int patch = pool32 ? (item >> 1) : 2 * item + (t >> 1);
                const int y = (pool32 ? 16 * (item & 1) : 0) + wq * 4 + (lane >> 3);
                const int x = (pool32 ? 8 * t : 8 * (t & 1)) + (lane & 7);
                const bool writer = ((lane & 9) == 0) && patch < n_patches;
                const long long qo = (long long)patch * PPo + (long long)((y >> 1) + 1) * po + (x >> 1) + 1;
#pragma unroll
                for (int j = 0; j < CHUNKS; ++j) {
                    const int ch = chalf + 2 * j;
                    const float* bs = s_bias + half * NC + ch * 32;
                    uint4* o = A.out + (long long)((half * NC + ch * 32) / 8) * A.Rout + LEAD + qo;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint32_t w[4];
                        const float4 b0 = *reinterpret_cast<const float4*>(bs + g * 8), b1 = *reinterpret_cast<const float4*>(bs + g * 8 + 4);
                        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float f0 = fmaxf(__uint_as_float(v[j][g * 8 + 2 * e]) + bb[2 * e], 0.f);
                            const float f1 = fmaxf(__uint_as_float(v[j][g * 8 + 2 * e + 1]) + bb[2 * e + 1], 0.f);
                            uint32_t m = pack_bf16x2(f0, f1);
                            uint32_t o1 = __shfl_xor_sync(0xFFFFFFFFu, m, 1);
                            __nv_bfloat162 a = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&m), *reinterpret_cast<__nv_bfloat162*>(&o1));
                            m = *reinterpret_cast<uint32_t*>(&a);
                            uint32_t o8 = __shfl_xor_sync(0xFFFFFFFFu, m, 8);
                            a = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&m), *reinterpret_cast<__nv_bfloat162*>(&o8));
                            w[e] = *reinterpret_cast<uint32_t*>(&a);
                        }
                        if (writer) o[(long long)g * A.Rout] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            };
            auto tile_addr = [&](int t, int j) {
                return tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)((acc_st * UMMA_T + t) * NC + (chalf + 2 * j) * 32);
            };
            // a tile that has reached the registers is handed back to the MMA issuers before it is converted and stored
            auto release_tile = [&](int t) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&acc_empty[acc_st * UMMA_T + t]));
            };
            if constexpr (CHUNKS == 1) {
                // N = 64: the epilogue is what bounds these layers, so the TMEM read of tile t + 1 is in flight while
                // tile t is converted and stored (two register buffers)
                uint32_t va[1][32], vb[1][32];
                tmem_ld32(tile_addr(0, 0), va[0]);
#pragma unroll
                for (int t = 0; t < UMMA_T; ++t) {
                    tmem_ld_wait_regs((t & 1) ? vb[0] : va[0]);
                    release_tile(t);
                    if (t + 1 < UMMA_T) tmem_ld32(tile_addr(t + 1, 0), (t & 1) ? va[0] : vb[0]);
                    if (POOL) store_tile_pool(t, (t & 1) ? vb : va); else store_tile(t, (t & 1) ? vb : va);
                }
            } else {
#pragma unroll 1
                for (int t = 0; t < UMMA_T; ++t) {
                    uint32_t v[CHUNKS][32];
#pragma unroll
                    for (int j = 0; j < CHUNKS; ++j) tmem_ld32(tile_addr(t, j), v[j]);
                    tmem_ld_wait();
                    release_tile(t);
                    if (POOL) store_tile_pool(t, v); else store_tile(t, v);
                }
            }
            if (POOL) {
                // zero halo of the pooled layout (row 0 and column 0 of every patch, and the row behind the last patch)
                const int S = A.pitch - 1, So = S >> 1, po = So + 1, PPo = po * po, planes = A.cout / 8;
                const int et = threadIdx.x - 64;
                const int np = pool32 ? 1 : 2;
                for (int k = 0; k < np; ++k) {
                    const int patch = pool32 ? (item >> 1) : 2 * item + k;
                    if (patch >= n_patches) break;
                    const bool top = !pool32 || (item & 1) == 0, bottom = !pool32 || (item & 1) == 1;
                    const int r_lo = pool32 ? 8 * (item & 1) + 1 : 1, r_n = pool32 ? 8 : So;
                    const int n_top = top ? po : 0;
                    const int n_tail = (bottom && patch == n_patches - 1) ? po + 1 : 0;
                    const int cnt = n_top + r_n + n_tail;
                    for (int i = et; i < cnt * planes; i += 32 * EPI_WARPS) {
                        const int pl = i / cnt, z = i - pl * cnt;
                        long long qo;
                        if (z < n_top) qo = (long long)patch * PPo + z;
                        else if (z < n_top + r_n) qo = (long long)patch * PPo + (long long)(r_lo + z - n_top) * po;
                        else qo = (long long)(patch + 1) * PPo + (z - n_top - r_n);
                        A.out[(long long)pl * A.Rout + LEAD + qo] = make_uint4(0, 0, 0, 0);
                    }
                }
            }
            LG_TACC(5, t_epi);
            if (++acc_st == ACC_STAGES) { acc_st = 0; acc_ph ^= 1; }
        }
#ifdef LG_CNN_TIMING
        if (warp == 2 && lane == 0 && blockIdx.x < 148) {
            g_cnn_timing[A.layer][blockIdx.x][4] = (unsigned long long)t_acc[4];
            g_cnn_timing[A.layer][blockIdx.x][5] = (unsigned long long)t_acc[5];
        }
#endif
    }
done:
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// fp32 patches [n][9][32][32] -> layer-0 input: 2 planes (channels 0-7 | 8 + zeros), pitch 33
__global__ void pack_input_kernel(const float* __restrict__ patches, uint4* __restrict__ out, long long R, int n_host,
                                  const int32_t* __restrict__ n_dev) {
    const int n = n_dev ? min(*n_dev, n_host) : n_host;
    const int pitch = 33, PP = 33 * 33;
    const long long total = (long long)n * PP + pitch + 1 + LEAD;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long q = i - LEAD;
        uint4 w0 = make_uint4(0, 0, 0, 0), w1 = w0;
        if (q >= 0 && q < (long long)n * PP) {
            const int pn = (int)(q / PP), ql = (int)(q - (long long)pn * PP);
            const int r = ql / pitch, c = ql - r * pitch;
            if (r >= 1 && c >= 1) {
                const float* p = patches + (size_t)pn * 9 * 1024 + (r - 1) * 32 + (c - 1);
                float f[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) f[k] = p[k * 1024];
                w0.x = pack_bf16x2(f[0], f[1]); w0.y = pack_bf16x2(f[2], f[3]);
                w0.z = pack_bf16x2(f[4], f[5]); w0.w = pack_bf16x2(f[6], f[7]);
                w1.x = pack_bf16x2(f[8], 0.f);
            }
        }
        out[i] = w0;
        out[R + i] = w1;
    }
}

__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
    uint4 r;
    const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
    __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
    return r;
}

// 2x2 max pool between two shared-halo layouts (S -> S/2).  FINAL: write fp32 NHWC [n][S/2][S/2][C] instead
// (the input of cnn_tail_kernel).
template <bool FINAL>
__global__ void pool2x2_kernel(const uint4* __restrict__ in, long long Rin, int S, void* __restrict__ outp, long long Rout,
                               int planes, int n_host, const int32_t* __restrict__ n_dev) {
    const int n = n_dev ? min(*n_dev, n_host) : n_host;
    const int pin = S + 1, PPin = pin * pin, So = S / 2, po = So + 1, PPo = po * po;
    if (FINAL) {
        float* out = reinterpret_cast<float*>(outp);
        const long long total = (long long)n * So * So * planes;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const int p = (int)(i % planes);
            const long long pix = i / planes;
            const int x = (int)(pix % So), y = (int)((pix / So) % So);
            const long long pn = pix / (So * So);
            const uint4* src = in + (long long)p * Rin + LEAD + pn * PPin + (2 * y + 1) * pin + 2 * x + 1;
            uint4 m = bf16x8_max(bf16x8_max(src[0], src[1]), bf16x8_max(src[pin], src[pin + 1]));
            const __nv_bfloat162* pm = reinterpret_cast<const __nv_bfloat162*>(&m);
            float4 lo, hi;
            float2 t;
            t = __bfloat1622float2(pm[0]); lo.x = t.x; lo.y = t.y;
            t = __bfloat1622float2(pm[1]); lo.z = t.x; lo.w = t.y;
            t = __bfloat1622float2(pm[2]); hi.x = t.x; hi.y = t.y;
            t = __bfloat1622float2(pm[3]); hi.z = t.x; hi.w = t.y;
            float4* dst = reinterpret_cast<float4*>(out + (pix * planes + p) * 8);
            dst[0] = lo; dst[1] = hi;
        }
    } else {
        uint4* out = reinterpret_cast<uint4*>(outp);
        const long long per_plane = (long long)n * PPo + po + 1 + LEAD;
        const long long total = per_plane * planes;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const int p = (int)(i / per_plane);
            const long long rho = i - (long long)p * per_plane;
            const long long q = rho - LEAD;
            uint4 m = make_uint4(0, 0, 0, 0);
            if (q >= 0 && q < (long long)n * PPo) {
                const long long pn = q / PPo;
                const int ql = (int)(q - pn * PPo);
                const int r = ql / po, c = ql - r * po;
                if (r >= 1 && c >= 1) {
                    const uint4* src = in + (long long)p * Rin + LEAD + pn * PPin + (2 * (r - 1) + 1) * pin + 2 * (c - 1) + 1;
                    m = bf16x8_max(bf16x8_max(src[0], src[1]), bf16x8_max(src[pin], src[pin + 1]));
                }
            }
            out[(long long)p * Rout + rho] = m;
        }
    }
}

// plane-major bf16 activations -> fp32 NCHW [n][C][S][S] (parity tests: lg_cnn_bf16_features)
__global__ void unpack_features_kernel(const uint4* __restrict__ in, long long R, int S, int C, int n, float* __restrict__ out) {
    const int pitch = S + 1, PP = pitch * pitch;
    const long long total = (long long)n * C * S * S;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % S), y = (int)((i / S) % S), ch = (int)((i / ((long long)S * S)) % C);
        const long long pn = i / ((long long)S * S * C);
        const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(in + (long long)(ch / 8) * R + LEAD + pn * PP + (y + 1) * pitch + x + 1);
        out[i] = __bfloat162float(src[ch % 8]);
    }
}

struct LayerCfg { int cin, cout, S, KP, KC, NC; };
const LayerCfg kLayers[6] = {{9, 64, 32, 2, 1, 64},   {64, 64, 32, 8, 1, 64},   {64, 128, 16, 8, 1, 128},
                             {128, 128, 16, 8, 2, 128}, {128, 256, 8, 8, 2, 128}, {256, 256, 8, 8, 4, 128}};

inline long long rows_per_plane(int S, long long n) {
    const long long pitch = S + 1, Q = n * pitch * pitch;
    const long long tiles = (Q + pitch + 1 + TILE_M - 1) / TILE_M;
    return LEAD + tiles * TILE_M + 640;      // slack: the blocked input stage of the last pooled item reads ~600 rows from its base
}
// rows of one input stage: the item's positions plus one image row and one pixel on either side
inline int a_rows(int S, bool pool) {
    const int span = !pool ? TILE_M : (S == 32 ? 15 * 33 + 32 : 17 * 17 + 15 * 17 + 16);
    return (span + 2 * (S + 1) + 2 + 7) / 8 * 8;
}

template <int KP, int NC, bool POOL>
size_t conv_smem(int S) {
    return 2 * (size_t)KP * a_rows(S, POOL) * 16 + (size_t)nb_stages_kp(KP, NC, POOL) * KP * NC * 16 + 256 * sizeof(float) +
           32 * sizeof(uint64_t) + 16;
}

uint16_t f2bf(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

// packed bf16 weight offsets (in uint4 units) of every layer inside cnn.bf16_blob
size_t layer_w_units(const LayerCfg& l) { return (size_t)(l.cout / l.NC) * l.KC * 9 * l.KP * l.NC; }

template <int KP, int NC, bool POOL>
int launch_conv(const UmmaConvArgs& A, int S, int sms, cudaStream_t st) {
    const size_t smem = conv_smem<KP, NC, POOL>(S);
    LG_ENSURE_SMEM((conv3x3_umma_kernel<KP, NC, POOL>), smem);
    const long long Qmax = (long long)A.n_host * A.PP;
    const int items = POOL ? (S == 32 ? 2 * A.n_host : (A.n_host + 1) / 2)
                           : (int)((Qmax + A.pitch + 1 + TILE_M - 1) / TILE_M) * A.n_split;
    const int grid = items < sms ? items : sms;
    conv3x3_umma_kernel<KP, NC, POOL><<<grid, CONV_THREADS, smem, st>>>(A);
    LG_LAUNCH_CHECK();
    return LG_OK;
}

}  // namespace

uint64_t lg_cnn_blob_floats();
int lg_launch_cnn_tail(const float* feat, const float* blob_tail, float* logits, int n, const int32_t* n_dev, cudaStream_t st);
int lg_launch_cnn_tail_bf16(const void* act, long long act_rows, int lead, const float* blob_tail, float* logits, int n,
                            const int32_t* n_dev, cudaStream_t st);

// Pack the folded fp32 weights (cnn.py:pack_weights layout) into the bf16 operand layout; synchronous.
int lg_cnn_prepare_bf16(lg_context* c) {
    std::vector<float> blob(c->cnn.n_floats);
    LG_CUDA(cudaMemcpy(blob.data(), c->cnn.blob, blob.size() * sizeof(float), cudaMemcpyDeviceToHost));
    size_t units = 0;
    for (const LayerCfg& l : kLayers) units += layer_w_units(l);
    std::vector<uint16_t> packed(units * 8, 0);
    const float* w = blob.data();
    size_t off = 0;
    for (const LayerCfg& l : kLayers) {
        const int n_split = l.cout / l.NC;
        for (int half = 0; half < n_split; ++half)
            for (int kc = 0; kc < l.KC; ++kc)
                for (int tap = 0; tap < 9; ++tap)
                    for (int p = 0; p < l.KP; ++p)
                        for (int n = 0; n < l.NC; ++n)
                            for (int e = 0; e < 8; ++e) {
                                const int ci = kc * l.KP * 8 + p * 8 + e, co = half * l.NC + n;
                                const float v = ci < l.cin ? w[((size_t)tap * l.cin + ci) * l.cout + co] : 0.f;
                                packed[(off + ((((size_t)(half * l.KC + kc) * 9 + tap) * l.KP + p) * l.NC + n)) * 8 + e] = f2bf(v);
                            }
        off += layer_w_units(l);
        w += 9ull * l.cin * l.cout + l.cout;
    }
    if (!c->cnn.bf16_blob) LG_CUDA(cudaMalloc(&c->cnn.bf16_blob, packed.size() * sizeof(uint16_t)));
    LG_CUDA(cudaMemcpy(c->cnn.bf16_blob, packed.data(), packed.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    return LG_OK;
}

static int run_cnn_bf16(lg_context* c, const float* patches, int n, const int32_t* n_dev, float* logits, int stop_layer,
                        float* feat_out, cudaStream_t st, int generic_tail = 0);
int lg_launch_cnn_tail_generic(const float* feat, const float* w, float* logits, int C, int S2, int attention, int m, cudaStream_t st);

// encoder [64, 128, 256] with another attention type (model.py:30-60): the convolutions as for the default architecture,
// then the pooled fp32 features go through the generic tail of the fp32 path
int lg_run_cnn_bf16_variant(lg_context* c, const float* patches, int n, float* logits, cudaStream_t st) {
    return run_cnn_bf16(c, patches, n, nullptr, logits, -1, nullptr, st, 1);
}

// geometry of the layer-0 input, for the gather kernel that writes it directly (lg_score.cu)
long long lg_cnn_input_plane_rows(long long n_patches) { return rows_per_plane(32, n_patches); }
int lg_cnn_input_lead() { return LEAD; }

int lg_run_cnn_bf16(lg_context* c, const float* patches, int n, const int32_t* n_dev, float* logits, cudaStream_t st) {
    if ((n_dev || !patches) && n > c->cnn_cap) { lg_set_error("device-side patch count needs n <= %d", c->cnn_cap); return LG_E_CAPACITY; }
    return run_cnn_bf16(c, patches, n, n_dev, logits, -1, nullptr, st);
}

extern "C" int lg_cnn_bf16_features(lg_context* c, const float* patches, int n, int layer, float* features, void* stream) {
    if (!c || !patches || !features || n < 1 || layer < 0 || layer > 5) return LG_E_ARG;
    if (!c->cnn.loaded) { lg_set_error("lg_cnn_bf16_features: no weights loaded"); return LG_E_ARG; }
    if (n > c->cnn_cap) return LG_E_CAPACITY;
    return run_cnn_bf16(c, patches, n, nullptr, nullptr, layer, features, (cudaStream_t)stream);
}

static int run_cnn_bf16(lg_context* c, const float* patches, int n, const int32_t* n_dev, float* logits, int stop_layer,
                        float* feat_out, cudaStream_t st, int generic_tail) {
    if (!c->cnn.bf16_blob) { lg_set_error("bf16 CNN weights are not prepared"); return LG_E_ARG; }
    static int sms = 0, issuers = 4, fuse_pool = 1, wres = 1;
    if (!sms) {
        int dev = 0;
        LG_CUDA(cudaGetDevice(&dev));
        LG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const char* e = getenv("LG_CNN_ISSUERS");     // measurement switch: 1 = single issuing warp
        if (e && (e[0] == '1' || e[0] == '2' || e[0] == '4')) issuers = e[0] - '0';
        const char* fp = getenv("LG_CNN_FUSE_POOL");  // measurement switch: 0 = separate pool kernels after layers 1 and 3
        if (fp && fp[0] == '0') fuse_pool = 0;
        const char* wr = getenv("LG_CNN_L0_RESIDENT"); // measurement switch: 0 = layer 0 streams its weights per item like the others
        if (wr && wr[0] == '0') wres = 0;
    }
    // bias pointers inside the fp32 blob; bf16 weights inside bf16_blob
    const float* bias[6];
    const uint4* wts[6];
    {
        const float* w = c->cnn.blob;
        const uint4* pw = reinterpret_cast<const uint4*>(c->cnn.bf16_blob);
        for (int l = 0; l < 6; ++l) {
            bias[l] = w + 9ull * kLayers[l].cin * kLayers[l].cout;
            w += 9ull * kLayers[l].cin * kLayers[l].cout + kLayers[l].cout;
            wts[l] = pw;
            pw += layer_w_units(kLayers[l]);
        }
    }
    const float* tail = c->cnn.blob + (lg_cnn_blob_floats() - (256 + 1 + 256 * 256 + 256 + 256 * 128 + 128 + 128 * 64 + 64 + 64 + 1));
    for (int done = 0; done < n; done += c->cnn_cap) {
        const int m = n - done < c->cnn_cap ? n - done : c->cnn_cap;
        uint4* buf[2] = {reinterpret_cast<uint4*>(c->cnn_act0), reinterpret_cast<uint4*>(c->cnn_act1)};
        int cur = 0;   // buffer holding the current layer's input
        if (patches) {     // null: the gather kernel wrote the packed input into buf[0] itself
            const long long R = rows_per_plane(32, m);
            const long long total = (long long)m * 33 * 33 + 34 + LEAD;
            const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
            pack_input_kernel<<<grid, 256, 0, st>>>(patches + (size_t)done * LG_CHANNELS * LG_PATCH * LG_PATCH, buf[0], R, m, n_dev);
            LG_LAUNCH_CHECK();
        }
        for (int l = 0; l < 6; ++l) {
            const LayerCfg& L = kLayers[l];
            UmmaConvArgs A;
            A.in = buf[cur]; A.out = buf[cur ^ 1]; A.wt = wts[l]; A.bias = bias[l];
            A.R = rows_per_plane(L.S, m);
            A.pitch = L.S + 1; A.PP = A.pitch * A.pitch; A.n_dev = n_dev; A.n_host = m;
            const bool fused_pool = fuse_pool && (l == 1 || l == 3);   // second conv of the 32 px and 16 px blocks
            A.KC = L.KC; A.n_split = L.cout / L.NC; A.rows = a_rows(L.S, fused_pool); A.cout = L.cout; A.layer = l; A.issuers = issuers; A.wres = wres;
            A.Rout = rows_per_plane(L.S / 2, m);
            int rc;
            if (L.KP == 2) rc = launch_conv<2, 64, false>(A, L.S, sms, st);
            else if (L.NC == 64) rc = fused_pool ? launch_conv<8, 64, true>(A, L.S, sms, st) : launch_conv<8, 64, false>(A, L.S, sms, st);
            else rc = fused_pool ? launch_conv<8, 128, true>(A, L.S, sms, st) : launch_conv<8, 128, false>(A, L.S, sms, st);
            if (rc) return rc;
            cur ^= 1;
            if (l == 5 && stop_layer != 5 && !generic_tail) break;   // the tail kernel pools the last layer's output while it loads it
            if ((l & 1) && !fused_pool) {   // max-pool after the second conv of each block
                const int planes = L.cout / 8;
                if (l == 5) {
                    const long long total = (long long)m * 16 * planes;
                    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
                    pool2x2_kernel<true><<<grid, 256, 0, st>>>(buf[cur], A.R, L.S, buf[cur ^ 1], 0, planes, m, n_dev);
                } else {
                    const long long Rout = rows_per_plane(L.S / 2, m);
                    const long long total = ((long long)m * (L.S / 2 + 1) * (L.S / 2 + 1) + L.S / 2 + 2 + LEAD) * planes;
                    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
                    pool2x2_kernel<false><<<grid, 256, 0, st>>>(buf[cur], A.R, L.S, buf[cur ^ 1], Rout, planes, m, n_dev);
                }
                LG_LAUNCH_CHECK();
                cur ^= 1;
            }
            if (l == stop_layer) {
                if (l == 5) {   // pooled fp32 NHWC [m][4][4][256] -> NCHW
                    LG_CUDA(cudaMemcpyAsync(feat_out, buf[cur], (size_t)m * 16 * 256 * sizeof(float), cudaMemcpyDeviceToDevice, st));
                } else {
                    const int So = (l & 1) ? L.S / 2 : L.S;
                    unpack_features_kernel<<<148 * 8, 256, 0, st>>>(buf[cur], rows_per_plane(So, m), So, L.cout, m, feat_out);
                    LG_LAUNCH_CHECK();
                }
                return LG_OK;
            }
        }
        int rc;
        if (generic_tail) {     // buf[cur]: pooled fp32 NHWC [m][4][4][256]; the tail's weights follow the six convolutions
            const float* w = c->cnn.blob;
            for (int l = 0; l < 6; ++l) w += 9ull * kLayers[l].cin * kLayers[l].cout + kLayers[l].cout;
            rc = lg_launch_cnn_tail_generic(reinterpret_cast<const float*>(buf[cur]), w, logits + done, 256, 16, c->cnn.cfg.attention, m, st);
        } else {
            rc = lg_launch_cnn_tail_bf16(buf[cur], rows_per_plane(8, m), LEAD, tail, logits + done, m, n_dev, st);
        }
        if (rc) return rc;
    }
    return LG_OK;
}

#ifdef LG_CNN_TIMING
extern "C" int lg_cnn_timing(unsigned long long* out_host) {   // [6][148][8]
    return cudaMemcpyFromSymbol(out_host, g_cnn_timing, sizeof(g_cnn_timing)) == cudaSuccess ? 0 : -2;
}
#endif
