// Stage 1 - optimal-leaf selection (reference scripts/utils/leaf_scorer.py:25-203, 277-306).
//
// Kernels (all batched over frames, no host synchronisation):
//   leaf_band_kernel     ONE pass over labels + depth, 128-bit loads, an 8 x 8 pixel block per thread: per-label pixel count,
//                        coordinate sums, depth sum, sum of ray lengths, bounding box, border contact, depth key range,
//                        first leaf pixel of the frame - and the depth values of the leaf pixels grouped by label inside
//                        every band of 8 rows (what the median needs), the bit mask of the leaf union and the occupancy of
//                        its 8 x 8 blocks (what the distance-transform search needs).  Depth is read once and only under
//                        leaves; nothing re-reads the inputs.
//   leaf_median_kernel   exact np.median per label by radix selection over the label's per-band parts
//   edt_argmax_kernel    the background pixel farthest from every leaf (the only thing leaf_scorer.py:67-71 takes from its
//                        distance field): exact branch and bound on the bit mask, no distance image
//   edt_vcol_kernel / edt_row_kernel   the full exact squared Euclidean distance transform (lg_edt_squared)
//   select_leaf_kernel   the per-leaf scores, tall-leaf rule, Pareto front and weighted pick
//
// Integer sums are exact and order independent, so results do not depend on scheduling: coordinate
// sums are 64-bit integers, depth and ray-length sums are fixed point (2^-28 m and 2^-36).
#include <math_constants.h>
#include <stdlib.h>

#include "lg_internal.cuh"

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr double DEP_SCALE = 268435456.0;       // 2^28
constexpr double DIST_SCALE = 68719476736.0;    // 2^36

__device__ __forceinline__ unsigned f2key(float f) {   // order-preserving float -> uint32 key
    const unsigned u = __float_as_uint(f);
    return u ^ ((unsigned)((int)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
    unsigned u = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
    return __uint_as_float(u);
}

__global__ void clear_tables_kernel(lg_context c, int n) {
    size_t total = (size_t)n * c.L;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        c.cnt[i] = 0; c.sx[i] = 0; c.sy[i] = 0; c.sdep[i] = 0; c.sdist[i] = 0;
        c.bx0[i] = 0xFFFFFFFFu; c.by0[i] = 0xFFFFFFFFu; c.bx1[i] = 0; c.by1[i] = 0; c.border[i] = 0;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        c.first_leaf[i] = 0xFFFFFFFFu; c.edt_best[i] = 0ull; c.status[i] = 0;
        c.krange[2 * i] = 0xFFFFFFFFu; c.krange[2 * i + 1] = 0u;
    }
}

// ray_tab[y * W + x] = sum over rows y' <= y and columns x' <= x of round(2^36 * sqrt(((x' - cx)^2 + (y' - cy)^2) / f^2 + 1)):
// a summed-area table of the length of the viewing ray through a pixel per unit depth (leaf_scorer.py:104-113 with
// X = md (x - cx) / f, Y = md (y - cy) / f, Z = md).  The sum over any rectangle of pixels is then four entries; the
// per-pixel terms are integers (the total of a whole 4096 x 16384 frame stays below 2^63), so any grouping of the pixels
// gives the same total.  The table depends on the camera only: it is built once per camera and read (L2-resident) by every
// frame.  First the row-wise prefix sums (one warp per row: 32-pixel segments, warp scan, carry), then the running sum
// down every column.
__global__ void ray_rows_kernel(unsigned long long* __restrict__ tab, int H, int W, lg_camera cam) {
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (y >= H) return;
    const double inv_f2 = 1.0 / (cam.f * cam.f);
    const double ddy = (double)y - cam.cy;
    unsigned long long carry = 0;
    for (int x0 = 0; x0 < W; x0 += 32) {
        const int x = x0 + lane;
        unsigned long long v = 0;
        if (x < W) {
            const double ddx = (double)x - cam.cx;
            v = (unsigned long long)__double2ll_rn(sqrt((ddx * ddx + ddy * ddy) * inv_f2 + 1.0) * DIST_SCALE);
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(FULL, v, d);
            if (lane >= d) v += t;
        }
        v += carry;
        if (x < W) tab[(size_t)y * W + x] = v;
        carry = __shfl_sync(FULL, v, 31);
    }
}
__global__ void ray_cols_kernel(unsigned long long* __restrict__ tab, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    unsigned long long acc = 0;
    for (int y = 0; y < H; ++y) {
        acc += tab[(size_t)y * W + x];
        tab[(size_t)y * W + x] = acc;
    }
}
// sum of the table's terms over rows [ya, yb] x columns [xa, xb]
__device__ __forceinline__ unsigned long long ray_rect(const unsigned long long* __restrict__ sat, int W, int xa, int xb, int ya, int yb) {
    const unsigned long long* hi = sat + (size_t)yb * W;
    unsigned long long s = hi[xb];
    if (xa > 0) s -= hi[xa - 1];
    if (ya > 0) {
        const unsigned long long* lo = sat + (size_t)(ya - 1) * W;
        s -= lo[xb];
        if (xa > 0) s += lo[xa - 1];
    }
    return s;
}

// ---------------------------------------------------------------------------------------------------
// the one pass over the inputs
// ---------------------------------------------------------------------------------------------------
// A CTA owns one band of LG_BAND (8) image rows of one frame; a thread owns an 8 x 8 block of pixels (8 adjacent columns).
//
// Labels first.  A thread loads the 8 label rows of its block with eight 128-bit loads that are all in flight together and
// classifies the block with a handful of logic operations: UNIFORM (all 64 pixels carry one label - almost every block,
// a leaf is 100-350 px wide and as tall) or BOUNDARY (anything else: a leaf edge, a label outside the table, the ragged
// right / bottom end of an image whose size is not a multiple of 8).  Neighbouring lanes with the same uniform label form a
// run; its first lane adds the run's rectangle to the CTA's table in closed form (pixel count, coordinate sums, bounding
// box, border contact, and the sum of ray lengths as four entries of the summed-area table ray_tab) and reserves the run's
// place among the label's depth values.  Boundary blocks are queued and handled row by row with all lanes busy (one lane
// per row of 8 pixels, rows split into label runs, runs reduced per label with warp redux operations).
//
// Depth second, once, and only where it is needed.  After the label phase the per-label pixel counts of the band are
// known, a warp scans them into offsets, and only then is depth read: uniform leaf blocks issue their sixteen 128-bit
// loads together, add the values to the label's depth sum (2^-28 fixed point: exact, order independent), track the range
// of the depth bit patterns and store the values to their place in `seg`; background blocks never touch depth.
//
// seg: np.median needs the depths of every leaf's pixels.  Inside a band they are stored grouped by label, uniform blocks
// first (64 values each, so every label's block part is 256-byte aligned and written with 128-bit stores), the pixels of
// boundary blocks behind them: [blocks of label 1 | blocks of label 2 | ... | pixels of label 1 | pixels of label 2 | ...].
// band_off holds both offset tables (u16: blocks, pixels).  Positions depend on this band only - no frame-wide count or
// scan - and the median kernel walks the bands of its label's bounding box.  The order inside a label's part is whatever
// the atomics hand out: a median does not depend on it.  Only labels >= 1 are stored: the reference drops the smallest id
// present (the background), which is 0 whenever 0 occurs at all.
//
// The pass also writes the bit mask of the leaf union (ubits, bit x % 8 of byte x / 8) and one occupancy byte per block
// (cellocc) for the distance-transform search, and the frame's first leaf pixel.
constexpr int TS_PX = 8;

struct SmemLeaf {
    unsigned cnt, sx, sy, bx0, bx1, by0, by1, border;
    unsigned long long sdep, sdist;
};

struct BandShared {         // views into the CTA's dynamic shared memory
    uint4* prel;            // [8 * NT] per row of a boundary block: place of each of its 8 pixels in the label's pixel part (8 x u16)
    SmemLeaf* tab;          // [L]
    unsigned* cntF;         // [L] uniform blocks per label; after the scan: offset (in values) of the label's block part
    unsigned* cntP;         // [L] pixels in boundary blocks per label; after the scan: offset of the label's pixel part
    unsigned short* q;      // [NT] threads whose block is a boundary block
};

// 2^28 * depth is exact in float32 (power-of-two scale), so the fixed-point term equals the float64 formulation
__device__ __forceinline__ long long fixed28(float v) { return __float2ll_rn(fminf(fmaxf(v, -2048.f), 2048.f) * 268435456.f); }

// One-row runs, columns [xa, xb] of row y with label il (il < 0: this lane has none), one per lane: reduced per distinct
// label of the warp with full-warp redux operations; one lane per label updates the shared-memory table.  All 32 lanes call.
__device__ __forceinline__ void band_add_runs(const BandShared& S, int il, int xa, int xb, int y, long long sdep,
                                              const unsigned long long* __restrict__ sat, int W, int H, int lane) {
    const unsigned nx = (unsigned)(xb - xa + 1);
    unsigned long long myray = 0;
    if (il >= 1) myray = ray_rect(sat, W, xa, xb, y, y);
    unsigned todo = __ballot_sync(FULL, il >= 1);
    while (todo) {
        const int src = __ffs(todo) - 1;
        const int cl = __shfl_sync(FULL, il, src);
        const bool mine = il == cl;
        todo &= ~__ballot_sync(FULL, mine);
        const unsigned gn = __reduce_add_sync(FULL, mine ? nx : 0u);
        const unsigned gsx = __reduce_add_sync(FULL, mine ? nx * (unsigned)(xa + xb) / 2u : 0u);
        const unsigned gsy = __reduce_add_sync(FULL, mine ? nx * (unsigned)y : 0u);
        const unsigned gxa = __reduce_min_sync(FULL, mine ? (unsigned)xa : 0xFFFFFFFFu);
        const unsigned gxb = __reduce_max_sync(FULL, mine ? (unsigned)xb : 0u);
        const unsigned gya = __reduce_min_sync(FULL, mine ? (unsigned)y : 0xFFFFFFFFu);
        const unsigned gyb = __reduce_max_sync(FULL, mine ? (unsigned)y : 0u);
        // 64-bit sums as two redux operations: value = hi * 2^24 + lo, lo in [0, 2^24)
        const unsigned glo = __reduce_add_sync(FULL, mine ? (unsigned)(sdep & 0xFFFFFFll) : 0u);
        const int ghi = __reduce_add_sync(FULL, mine ? (int)(sdep >> 24) : 0);
        const unsigned rlo = __reduce_add_sync(FULL, mine ? (unsigned)(myray & 0xFFFFFFull) : 0u);
        const unsigned rhi = __reduce_add_sync(FULL, mine ? (unsigned)(myray >> 24) : 0u);
        if (lane == src) {
            SmemLeaf* t = &S.tab[cl];
            atomicAdd(&t->cnt, gn);
            atomicAdd(&t->sx, gsx); atomicAdd(&t->sy, gsy);
            atomicMin(&t->bx0, gxa); atomicMax(&t->bx1, gxb);
            atomicMin(&t->by0, gya); atomicMax(&t->by1, gyb);
            if (gxa == 0 || gxb == (unsigned)(W - 1) || gya == 0 || gyb == (unsigned)(H - 1)) atomicOr(&t->border, 1u);
            atomicAdd(&t->sdep, (unsigned long long)(((long long)ghi << 24) + (long long)glo));
            atomicAdd(&t->sdist, ((unsigned long long)rhi << 24) + rlo);
        }
    }
}

// a hint, no register held: the line starts its way from DRAM to L2 while the CTA does something else
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// the 8 labels of pixels [p0, p0 + np) of a frame (np < 8: the rest reads as -1)
template <bool VEC>
__device__ __forceinline__ void load_labels8(const int16_t* __restrict__ lp, size_t p0, int np, int (&lab)[TS_PX]) {
    if (VEC) {
        const uint4 w = *reinterpret_cast<const uint4*>(lp + p0);
        const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < TS_PX; ++k) lab[k] = (int)(short)((ww[k >> 1] >> (16 * (k & 1))) & 0xFFFFu);
    } else {
#pragma unroll
        for (int k = 0; k < TS_PX; ++k) lab[k] = k < np ? (int)lp[p0 + k] : -1;
    }
}

template <bool VEC, int MAX_NT, int MIN_CTAS>
__global__ void __launch_bounds__(MAX_NT, MIN_CTAS) leaf_band_kernel(lg_context c, const int16_t* __restrict__ labels,
                                                                     const float* __restrict__ depth) {
    extern __shared__ __align__(16) unsigned char bs_smem[];
    __shared__ unsigned s_first, s_bad, s_umin, s_umax, s_nq;
    const int L = c.L, W = c.W, H = c.H, NT = blockDim.x;
    const size_t P = c.P;
    BandShared S;
    S.prel = reinterpret_cast<uint4*>(bs_smem);
    S.tab = reinterpret_cast<SmemLeaf*>(S.prel + (size_t)LG_BAND * NT);
    S.cntF = reinterpret_cast<unsigned*>(S.tab + L);
    S.cntP = S.cntF + L;
    S.q = reinterpret_cast<unsigned short*>(S.cntP + L);
    const int b = blockIdx.y, band = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    for (int l = tid; l < L; l += NT) {
        SmemLeaf z;
        z.cnt = 0; z.sx = 0; z.sy = 0; z.bx0 = 0xFFFFFFFFu; z.by0 = 0xFFFFFFFFu; z.bx1 = 0; z.by1 = 0;
        z.border = 0; z.sdep = 0; z.sdist = 0;
        S.tab[l] = z;
        S.cntF[l] = 0; S.cntP[l] = 0;
    }
    if (tid == 0) { s_first = 0xFFFFFFFFu; s_bad = 0; s_umin = 0xFFFFFFFFu; s_umax = 0; s_nq = 0; }
    __syncthreads();
    const int16_t* lp = labels + (size_t)b * P;
    const float* dp = depth + (size_t)b * P;
    uint8_t* ub = c.ubits + (size_t)b * c.ub_stride;
    uint8_t* cell = c.cellocc + ((size_t)b * c.n_bands + band) * c.ub_pitch;
    const unsigned long long* sat = c.ray_tab;
    const int pitch = c.ub_pitch;
    const int row0 = band * LG_BAND, nrows = min(LG_BAND, H - row0);
    const int x0 = tid * TS_PX;
    const bool active = x0 < W;
    const int npx = active ? min(TS_PX, W - x0) : 0;

    // ================= labels: classify the block =================
    int code = -1;                       // label of a uniform block; -2: boundary block (queued); -1: no pixels
    if (active) {
        bool uni = npx == TS_PX && nrows == LG_BAND;
        unsigned pat = 0;
        if (uni) {
            uint4 lw[LG_BAND];
#pragma unroll
            for (int r = 0; r < LG_BAND; ++r) {
                const size_t p0 = (size_t)(row0 + r) * W + x0;
                if (VEC) {
                    lw[r] = *reinterpret_cast<const uint4*>(lp + p0);
                } else {
                    unsigned w[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int k = 0; k < TS_PX; ++k) w[k >> 1] |= (unsigned)(unsigned short)lp[p0 + k] << (16 * (k & 1));
                    lw[r] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            pat = lw[0].x;
            unsigned diff = pat ^ __byte_perm(pat, 0, 0x1010);           // both halfwords equal?
#pragma unroll
            for (int r = 0; r < LG_BAND; ++r) diff |= (lw[r].x ^ pat) | (lw[r].y ^ pat) | (lw[r].z ^ pat) | (lw[r].w ^ pat);
            uni = diff == 0u && (pat & 0xFFFFu) < (unsigned)L;           // negative labels are >= 0x8000 here
        }
        if (uni) {
            code = (int)(pat & 0xFFFFu);
            if (code >= 1) {             // a leaf block: its depth rows are wanted after the offsets scan - start them towards L2 now
#pragma unroll
                for (int r = 0; r < LG_BAND; ++r) prefetch_l2(dp + (size_t)(row0 + r) * W + x0);
            }
            const uint8_t v = code >= 1 ? 0xFFu : 0u;
#pragma unroll
            for (int r = 0; r < LG_BAND; ++r) ub[(size_t)(row0 + r) * pitch + tid] = v;
            cell[tid] = code >= 1 ? 1 : 0;
        } else {
            code = -2;
            S.q[atomicAdd(&s_nq, 1u)] = (unsigned short)tid;
        }
    }
    // runs of lanes with the same uniform label: the first lane of a run does the run's bookkeeping
    const int prev = __shfl_up_sync(FULL, code, 1);
    const bool starts = lane == 0 || prev != code;
    const unsigned bmask = __ballot_sync(FULL, starts);
    const int head = 31 - __clz(bmask & (0xFFFFFFFFu >> (31 - lane)));
    const unsigned above = lane == 31 ? 0u : (bmask >> (lane + 1)) << (lane + 1);
    const int end = above ? __ffs(above) - 1 : 32;        // one past the last lane of this lane's run
    unsigned relF = 0;
    if (starts && code >= 0) {
        const unsigned n = (unsigned)(end - lane);
        if (code == 0) {
            atomicAdd(&S.tab[0].cnt, 64u * n);
        } else {
            relF = atomicAdd(&S.cntF[code], n);
            const unsigned xa = (unsigned)x0, xb = (unsigned)x0 + 8u * n - 1u, ya = (unsigned)row0, yb = (unsigned)row0 + LG_BAND - 1u;
            SmemLeaf* t = &S.tab[code];
            atomicAdd(&t->cnt, 64u * n);
            atomicAdd(&t->sx, 32u * n * (xa + xb));
            atomicAdd(&t->sy, 8u * n * (8u * ya + 28u));
            atomicMin(&t->bx0, xa); atomicMax(&t->bx1, xb);
            atomicMin(&t->by0, ya); atomicMax(&t->by1, yb);
            if (xa == 0 || xb == (unsigned)(W - 1) || ya == 0 || yb == (unsigned)(H - 1)) atomicOr(&t->border, 1u);
            atomicAdd(&t->sdist, ray_rect(sat, W, (int)xa, (int)xb, (int)ya, (int)yb));
            atomicMin(&s_first, (unsigned)((size_t)row0 * W) + xa);
        }
    }
    relF = __shfl_sync(FULL, relF, head);                 // first block of the run among the label's blocks
    __syncthreads();                                      // the queue is complete

    // ================= boundary blocks, labels: union bits, places of the pixels =================
    const int n_items = (int)s_nq * LG_BAND;              // one item = one row of a queued block; 8 lanes = one block
    for (int it0 = 0; it0 < n_items; it0 += NT) {
        const int it = it0 + tid;
        const bool have = it < n_items;
        const int blk = have ? (int)S.q[it >> 3] : 0, r = it & 7;
        const int xs = blk * TS_PX, y = row0 + r;
        const int np = (have && r < nrows) ? min(TS_PX, W - xs) : 0;
        int lab[TS_PX];
        if (np > 0) load_labels8<VEC>(lp, (size_t)y * W + xs, np, lab);
        else {
#pragma unroll
            for (int k = 0; k < TS_PX; ++k) lab[k] = -1;
        }
        unsigned bits = 0, nbg = 0;
        bool bad = false, ok[TS_PX];
#pragma unroll
        for (int k = 0; k < TS_PX; ++k) {
            const bool in = k < np;
            if (in && lab[k] >= 1) bits |= 1u << k;
            if (in && lab[k] == 0) ++nbg;
            if (in && (lab[k] < 0 || lab[k] >= L)) bad = true;
            ok[k] = in && lab[k] >= 1 && lab[k] < L;
        }
        int after[TS_PX];                                   // pixels of the same run to the right of pixel k
        after[TS_PX - 1] = 0;
#pragma unroll
        for (int k = TS_PX - 2; k >= 0; --k) after[k] = (ok[k] && ok[k + 1] && lab[k + 1] == lab[k]) ? after[k + 1] + 1 : 0;
        unsigned rel[TS_PX], base = 0, off = 0;
#pragma unroll
        for (int k = 0; k < TS_PX; ++k) {
            const bool first = ok[k] && (k == 0 || lab[k] != lab[k - 1]);
            if (first) { base = atomicAdd(&S.cntP[lab[k]], (unsigned)after[k] + 1u); off = 0; }
            rel[k] = ok[k] ? base + off : 0u;
            if (ok[k]) ++off;
        }
        if (have) S.prel[it] = make_uint4(rel[0] | (rel[1] << 16), rel[2] | (rel[3] << 16), rel[4] | (rel[5] << 16), rel[6] | (rel[7] << 16));
        if (np > 0) {
            if (bits) prefetch_l2(dp + (size_t)y * W + xs);      // this row's depth is read in the second boundary phase
            ub[(size_t)y * pitch + blk] = (uint8_t)bits;
            if (bits) atomicMin(&s_first, (unsigned)((size_t)y * W) + (unsigned)xs + (unsigned)(__ffs(bits) - 1));
            if (bad) s_bad = 1;
        }
        const unsigned tot_bg = __reduce_add_sync(FULL, nbg);
        if (lane == 0 && tot_bg) atomicAdd(&S.tab[0].cnt, tot_bg);
        unsigned occ = bits;                                // occupancy of the block: OR over its 8 rows (8 adjacent lanes)
        occ |= __shfl_xor_sync(FULL, occ, 1);
        occ |= __shfl_xor_sync(FULL, occ, 2);
        occ |= __shfl_xor_sync(FULL, occ, 4);
        if (have && r == 0) cell[blk] = occ ? 1 : 0;
    }
    __syncthreads();

    // ================= the band's offsets (warp 0) =================
    if (tid < 32) {
        const int chunk = (L + 31) >> 5;
        const int la = min(lane * chunk, L), lb = min(la + chunk, L);
        uint16_t* offF = c.band_off + ((size_t)b * c.n_bands + band) * 2 * c.lstride;
        uint16_t* offP = offF + c.lstride;
        unsigned mineF = 0, mineP = 0;
        for (int l = max(la, 1); l < lb; ++l) { mineF += S.cntF[l]; mineP += S.cntP[l]; }
        unsigned inclF = mineF, inclP = mineP;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned tf = __shfl_up_sync(FULL, inclF, d), tp = __shfl_up_sync(FULL, inclP, d);
            if (lane >= d) { inclF += tf; inclP += tp; }
        }
        const unsigned totF = __shfl_sync(FULL, inclF, 31), totP = __shfl_sync(FULL, inclP, 31);
        unsigned runF = inclF - mineF, runP = inclP - mineP;
        for (int l = la; l < lb; ++l) {
            const unsigned nf = l >= 1 ? S.cntF[l] : 0u, np2 = l >= 1 ? S.cntP[l] : 0u;
            offF[l] = (uint16_t)runF; offP[l] = (uint16_t)runP;
            S.cntF[l] = 64u * runF;
            S.cntP[l] = 64u * totF + runP;
            runF += nf; runP += np2;
        }
        if (lane == 31) { offF[L] = (uint16_t)totF; offP[L] = (uint16_t)totP; }
    }
    __syncthreads();

    // ================= depth: uniform leaf blocks =================
    float* segb = c.seg + (size_t)b * c.seg_stride + (size_t)row0 * W;
    unsigned umn = 0xFFFFFFFFu, umx = 0u;                   // range of the depth bit patterns of the leaf pixels
    long long sd = 0;
    if (code >= 1) {
        const int n = end - head, j = lane - head;
        float* o = segb + S.cntF[code] + 64u * relF + (unsigned)(TS_PX * j);      // + row * 8 n
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float4 d[4][2];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const size_t p0 = (size_t)(row0 + 4 * half + r) * W + x0;
                if (VEC) {
                    d[r][0] = *reinterpret_cast<const float4*>(dp + p0);
                    d[r][1] = *reinterpret_cast<const float4*>(dp + p0 + 4);
                } else {
                    d[r][0] = make_float4(dp[p0], dp[p0 + 1], dp[p0 + 2], dp[p0 + 3]);
                    d[r][1] = make_float4(dp[p0 + 4], dp[p0 + 5], dp[p0 + 6], dp[p0 + 7]);
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float v[TS_PX] = {d[r][0].x, d[r][0].y, d[r][0].z, d[r][0].w, d[r][1].x, d[r][1].y, d[r][1].z, d[r][1].w};
#pragma unroll
                for (int k = 0; k < TS_PX; ++k) {
                    sd += fixed28(v[k]);
                    const unsigned u = __float_as_uint(v[k]);
                    umn = min(umn, u); umx = max(umx, u);
                }
                float* orow = o + (size_t)(4 * half + r) * (TS_PX * n);
                *reinterpret_cast<float4*>(orow) = d[r][0];
                *reinterpret_cast<float4*>(orow + 4) = d[r][1];
            }
        }
    }
    {   // depth sums of the runs: one 64-bit add per run
        unsigned todo = __ballot_sync(FULL, starts && code >= 1);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int e = __shfl_sync(FULL, end, src);
            const bool mine = lane >= src && lane < e;
            const unsigned glo = __reduce_add_sync(FULL, mine ? (unsigned)(sd & 0xFFFFFFll) : 0u);
            const int ghi = __reduce_add_sync(FULL, mine ? (int)(sd >> 24) : 0);
            if (lane == src) atomicAdd(&S.tab[code].sdep, (unsigned long long)(((long long)ghi << 24) + (long long)glo));
        }
    }

    // ================= depth: boundary blocks, row by row =================
    for (int it0 = 0; it0 < n_items; it0 += NT) {
        const int it = it0 + tid;
        const bool have = it < n_items;
        const int blk = have ? (int)S.q[it >> 3] : 0, r = it & 7;
        const int xs = blk * TS_PX, y = row0 + r;
        const int np = (have && r < nrows) ? min(TS_PX, W - xs) : 0;
        int lab[TS_PX];
        if (np > 0) load_labels8<VEC>(lp, (size_t)y * W + xs, np, lab);
        else {
#pragma unroll
            for (int k = 0; k < TS_PX; ++k) lab[k] = -1;
        }
        bool ok[TS_PX], any_ok = false;
#pragma unroll
        for (int k = 0; k < TS_PX; ++k) { ok[k] = k < np && lab[k] >= 1 && lab[k] < L; any_ok |= ok[k]; }
        float v[TS_PX];
#pragma unroll
        for (int k = 0; k < TS_PX; ++k) v[k] = 0.f;
        if (any_ok) {
            const size_t p0 = (size_t)y * W + xs;
            if (VEC) {
                const float4 d0 = *reinterpret_cast<const float4*>(dp + p0), d1 = *reinterpret_cast<const float4*>(dp + p0 + 4);
                v[0] = d0.x; v[1] = d0.y; v[2] = d0.z; v[3] = d0.w; v[4] = d1.x; v[5] = d1.y; v[6] = d1.z; v[7] = d1.w;
            } else {
#pragma unroll
                for (int k = 0; k < TS_PX; ++k) if (k < np) v[k] = dp[p0 + k];
            }
        }
        const uint4 pr = have ? S.prel[it] : make_uint4(0, 0, 0, 0);
        const unsigned prw[4] = {pr.x, pr.y, pr.z, pr.w};
        int rid[TS_PX], nrun = 0;                           // run index of every stored pixel
#pragma unroll
        for (int k = 0; k < TS_PX; ++k) {
            if (ok[k] && (k == 0 || lab[k] != lab[k - 1])) ++nrun;
            rid[k] = nrun - 1;
            if (ok[k]) {
                segb[S.cntP[lab[k]] + ((prw[k >> 1] >> (16 * (k & 1))) & 0xFFFFu)] = v[k];
                const unsigned u = __float_as_uint(v[k]);
                umn = min(umn, u); umx = max(umx, u);
            }
        }
        const int rounds = __reduce_max_sync(FULL, nrun);
        for (int j = 0; j < rounds; ++j) {
            int il = -1, xa = 0, xb = 0;
            long long s = 0;
#pragma unroll
            for (int k = 0; k < TS_PX; ++k) {
                if (ok[k] && rid[k] == j) {
                    if (il < 0) { il = lab[k]; xa = xs + k; }
                    xb = xs + k;
                    s += fixed28(v[k]);
                }
            }
            band_add_runs(S, il, xa, xb, y, s, sat, W, H, lane);
        }
    }
    {   // range of the depth bit patterns, per CTA
        const unsigned wmn = __reduce_min_sync(FULL, umn), wmx = __reduce_max_sync(FULL, umx);
        if (lane == 0 && wmn <= wmx) { atomicMin(&s_umin, wmn); atomicMax(&s_umax, wmx); }
    }
    __syncthreads();

    // ================= the frame's table =================
    for (int l = tid; l < L; l += NT) {
        const SmemLeaf t = S.tab[l];
        if (t.cnt) {
            const size_t o = (size_t)b * L + l;
            atomicAdd(&c.cnt[o], t.cnt);
            if (l > 0) {
                atomicAdd(&c.sx[o], (unsigned long long)t.sx);
                atomicAdd(&c.sy[o], (unsigned long long)t.sy);
                atomicAdd(&c.sdep[o], t.sdep);
                atomicAdd(&c.sdist[o], t.sdist);
                atomicMin(&c.bx0[o], t.bx0); atomicMax(&c.bx1[o], t.bx1);
                atomicMin(&c.by0[o], t.by0); atomicMax(&c.by1[o], t.by1);
                if (t.border) atomicOr(&c.border[o], 1u);
            }
        }
    }
    if (tid == 0) {
        if (s_first != 0xFFFFFFFFu) atomicMin(&c.first_leaf[b], s_first);
        if (s_bad) atomicOr(&c.status[b], LG_ST_LABEL_RANGE);
        if (s_umin <= s_umax) {
            // Keys of non-negative floats are their bit patterns with the top bit set: the range of the patterns is the
            // range of the keys.  A negative depth (top bit set) would order the other way: give up the bound then - the
            // median's radix search only needs SOME range that holds every key.
            unsigned kmn = s_umin | 0x80000000u, kmx = s_umax | 0x80000000u;
            if (s_umax & 0x80000000u) { kmn = 0u; kmx = 0xFFFFFFFFu; }
            atomicMin(&c.krange[2 * b], kmn); atomicMax(&c.krange[2 * b + 1], kmx);
        }
    }
}

// background id = smallest id present (torch.unique(mask)[1:], leaf_scorer.py:32)
__device__ __forceinline__ int background_id(const uint32_t* cnt, int L) {
    for (int l = 0; l < L; ++l)
        if (cnt[l]) return l;
    return -1;
}

constexpr int MED_NT = 256;
constexpr int MED_CAP = 4096;    // candidate keys kept in shared memory once the search range is this small
constexpr int MED_BITS = 11;     // digit of the radix search: 2048 bins
constexpr int MED_BINS = 1 << MED_BITS;
constexpr int MED_DIRECT = 128;  // candidate sets this small are ranked by counting
constexpr int MED_BATCH = 256;   // bands whose extents are staged in shared memory at a time

struct MedShared {
    unsigned keys[2][MED_CAP];
    unsigned hist[MED_BINS];
    unsigned ext[MED_BATCH][4];  // per band: offset / count of the label's block part and of its pixel part (in values)
    unsigned part[MED_NT / 32];
    unsigned sel[3];             // chosen bin, elements below it, elements in it
    unsigned n_keys;
    unsigned res[2];             // keys of rank r and r + 1
    unsigned mn;
    int bg;
};

// np.median(depth[labels == l]) for every label of every frame (leaf_scorer.py:41-47): radix selection on the label's
// values, which leaf_band_kernel left grouped by label inside every band of rows.  A pass over the values walks the bands
// of the label's bounding box (extents from the band's two offset tables, staged in shared memory for 256 bands at a
// time): the block parts are read by the whole CTA with 128-bit loads, the short pixel parts warp by warp.  Keys are
// ranked relative to the smallest depth key of the frame's leaf pixels.  A label that does not fit shared memory gets one
// streaming pass with a 2048-bin histogram of the top 11 bits of the frame's key range (the median's bin then holds a
// few hundred values at most; were it still more than 4096 the next 11 bits get another pass), a second pass compacts
// the candidates of the median's bin into shared memory, and the selection finishes there: histogram rounds while there
// are more than 128 candidates, then ranking by counting.
__global__ void __launch_bounds__(MED_NT) leaf_median_kernel(lg_context c) {
    extern __shared__ __align__(16) unsigned char med_smem[];
    MedShared& S = *reinterpret_cast<MedShared*>(med_smem);
    const int l = blockIdx.x, b = blockIdx.y, L = c.L, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* cnt = c.cnt + (size_t)b * L;
    if (tid == 0) { S.bg = background_id(cnt, L); S.n_keys = 0; S.mn = 0xFFFFFFFFu; }
    __syncthreads();
    const unsigned n = cnt[l];
    if (n == 0 || l == S.bg || l == 0) {
        if (tid == 0) c.median[(size_t)b * L + l] = CUDART_NAN_F;
        return;
    }
    const size_t o = (size_t)b * L + l;
    const unsigned kmin = c.krange[2 * b], kmax = c.krange[2 * b + 1];
    const int t_first = (int)c.by0[o] / LG_BAND, t_last = (int)c.by1[o] / LG_BAND;
    const float* seg = c.seg + (size_t)b * c.seg_stride;
    const uint16_t* boff = c.band_off + (size_t)b * c.n_bands * 2 * c.lstride;
    const size_t band_px = (size_t)LG_BAND * c.W;

    // f(key, valid) for every value of the label; whole warps call it together (valid = false on padding lanes)
    auto for_each_key = [&](auto f) {
        for (int t0 = t_first; t0 <= t_last; t0 += MED_BATCH) {
            const int nb = min(MED_BATCH, t_last - t0 + 1);
            __syncthreads();
            if (tid < nb) {
                const uint16_t* oF = boff + (size_t)(t0 + tid) * 2 * c.lstride;
                const uint16_t* oP = oF + c.lstride;
                const unsigned f0 = oF[l], f1 = oF[l + 1], p0 = oP[l], p1 = oP[l + 1], totF = oF[L];
                S.ext[tid][0] = 64u * f0; S.ext[tid][1] = 64u * (f1 - f0);
                S.ext[tid][2] = 64u * totF + p0; S.ext[tid][3] = p1 - p0;
            }
            __syncthreads();
            for (int i = warp; i < nb; i += MED_NT / 32) {       // one warp per band: block part (four 128-bit loads in
                const unsigned cf = S.ext[i][1] >> 2;             // flight per lane), then the short pixel part
                const float4* v4 = reinterpret_cast<const float4*>(seg + (size_t)(t0 + i) * band_px + S.ext[i][0]);
                for (unsigned q0 = 0; q0 < cf; q0 += 128) {
                    float4 d[4];
                    bool okq[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const unsigned q = q0 + 32u * u + lane;
                        okq[u] = q < cf;
                        d[u] = okq[u] ? v4[q] : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        f(f2key(d[u].x) - kmin, okq[u]); f(f2key(d[u].y) - kmin, okq[u]);
                        f(f2key(d[u].z) - kmin, okq[u]); f(f2key(d[u].w) - kmin, okq[u]);
                    }
                }
                const unsigned cntv = S.ext[i][3];
                const float* v = seg + (size_t)(t0 + i) * band_px + S.ext[i][2];
                for (unsigned q0 = 0; q0 < cntv; q0 += 32) {
                    const unsigned q = q0 + lane;
                    const bool okq1 = q < cntv;
                    f(okq1 ? f2key(v[q]) - kmin : 0u, okq1);
                }
            }
        }
    };
    auto block_scan_pick = [&](unsigned rank) {
        // S.hist holds MED_BINS counts: finds the bin that holds `rank` -> S.sel = {bin, elements below the bin, elements in it}
        constexpr int PER = MED_BINS / MED_NT;
        unsigned mine = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) mine += S.hist[PER * tid + j];
        unsigned incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) S.part[warp] = incl;
        __syncthreads();
        unsigned before = 0;
        for (int w = 0; w < warp; ++w) before += S.part[w];
        const unsigned excl = before + incl - mine;
        if (rank >= excl && rank < excl + mine) {
            unsigned acc = excl;
            for (int j = 0; j < PER; ++j) {
                const unsigned h = S.hist[PER * tid + j];
                if (rank < acc + h) { S.sel[0] = (unsigned)(PER * tid + j); S.sel[1] = acc; S.sel[2] = h; break; }
                acc += h;
            }
        }
        __syncthreads();
    };
    auto clear_hist = [&]() {
        for (int i = tid; i < MED_BINS; i += MED_NT) S.hist[i] = 0;
        __syncthreads();
    };

    const unsigned k_lo = (n & 1) ? n / 2 : n / 2 - 1;   // rank of the lower middle
    const bool need_hi = !(n & 1);
    // candidates: the keys with (key >> s) == p; `below` elements are smaller than all of them, m are candidates
    int s = kmax > kmin ? 32 - __clz(kmax - kmin) : 0;    // number of significant bits of a relative key
    const int s_top = s;
    unsigned p = 0, below = 0, m = n;
    auto match = [&](unsigned key) { return s >= 32 ? true : (key >> s) == p; };
    while (m > MED_CAP && s > 0) {
        const int nbits = min(MED_BITS, s), s2 = s - nbits;
        const unsigned dmask = (1u << nbits) - 1u;
        clear_hist();
        if (s == s_top)     // first pass: every key is a candidate and has no bit above the digit
            for_each_key([&](unsigned key, bool ok) { if (ok) atomicAdd(&S.hist[key >> s2], 1u); });
        else
            for_each_key([&](unsigned key, bool ok) { if (ok && match(key)) atomicAdd(&S.hist[(key >> s2) & dmask], 1u); });
        __syncthreads();
        block_scan_pick(k_lo - below);
        p = s >= 32 ? S.sel[0] : ((p << nbits) | S.sel[0]);
        below += S.sel[1];
        m = S.sel[2];
        s = s2;
        __syncthreads();
    }
    unsigned klo = 0, khi = 0;
    bool have_hi = false;
    if (s == 0) {                  // every candidate is the key p
        klo = p;
        if (need_hi && below + m > k_lo + 1) { khi = p; have_hi = true; }
    } else {
        // the candidates into shared memory: all values of a small label (one cursor bump per warp), or the few of the
        // median's bin (one in a few hundred values: a plain atomic per hit)
        if (m == n) {
            for_each_key([&](unsigned key, bool ok) {
                const unsigned ball = __ballot_sync(FULL, ok);
                if (ball) {
                    unsigned base = 0;
                    if (lane == 0) base = atomicAdd(&S.n_keys, __popc(ball));
                    base = __shfl_sync(FULL, base, 0);
                    if (ok) S.keys[0][base + __popc(ball & ((1u << lane) - 1u))] = key;
                }
            });
        } else {
            const unsigned pp = p;
            const int sh = s;               // < 32 here: a histogram pass has narrowed the range
            for_each_key([&](unsigned key, bool ok) { if (ok && (key >> sh) == pp) S.keys[0][atomicAdd(&S.n_keys, 1u)] = key; });
        }
        __syncthreads();
        int cur = 0;
        unsigned r = k_lo - below;              // rank among the candidates
        unsigned mm = m;
        bool want_hi = need_hi && (r + 1 < mm); // otherwise the upper middle lies beyond the candidates: streaming pass below
        int ss = s;
        while (mm > MED_DIRECT && ss > 0) {
            const int nbits = min(MED_BITS, ss), s2 = ss - nbits;
            const unsigned dmask = (1u << nbits) - 1u;
            clear_hist();
            for (unsigned i = tid; i < mm; i += MED_NT) atomicAdd(&S.hist[(S.keys[cur][i] >> s2) & dmask], 1u);
            __syncthreads();
            block_scan_pick(r);
            const unsigned bin = S.sel[0], bl = S.sel[1], bm = S.sel[2];
            __syncthreads();
            if (want_hi && r + 1 >= bl + bm) {  // the upper middle is the smallest key of the bins above
                if (tid == 0) S.mn = 0xFFFFFFFFu;
                __syncthreads();
                unsigned mn = 0xFFFFFFFFu;
                for (unsigned i = tid; i < mm; i += MED_NT) { const unsigned key = S.keys[cur][i]; if (((key >> s2) & dmask) > bin) mn = min(mn, key); }
                mn = __reduce_min_sync(FULL, mn);
                if (lane == 0 && mn != 0xFFFFFFFFu) atomicMin(&S.mn, mn);
                __syncthreads();
                khi = S.mn; have_hi = true; want_hi = false;
            }
            if (tid == 0) S.n_keys = 0;
            __syncthreads();
            for (unsigned i0 = tid & ~31u; i0 < mm; i0 += MED_NT) {       // whole warps stay together
                const unsigned i = i0 + lane;
                const unsigned key = i < mm ? S.keys[cur][i] : 0u;
                const bool hit = i < mm && ((key >> s2) & dmask) == bin;
                const unsigned ball = __ballot_sync(FULL, hit);
                if (ball) {
                    unsigned base = 0;
                    if (lane == 0) base = atomicAdd(&S.n_keys, __popc(ball));
                    base = __shfl_sync(FULL, base, 0);
                    if (hit) S.keys[cur ^ 1][base + __popc(ball & ((1u << lane) - 1u))] = key;
                }
            }
            __syncthreads();
            cur ^= 1; r -= bl; mm = bm; ss = s2;
        }
        if (ss == 0 && mm > MED_DIRECT) {       // all remaining candidates are equal
            klo = S.keys[cur][0];
            if (want_hi) { khi = klo; have_hi = true; }
        } else {
            // rank by counting: the key with (#smaller <= rank < #smaller + #equal)
            if (tid == 0) { S.res[0] = 0; S.res[1] = 0; }
            __syncthreads();
            if (tid < (int)mm) {
                const unsigned key = S.keys[cur][tid];
                unsigned smaller = 0, equal = 0;
                for (unsigned j = 0; j < mm; ++j) { const unsigned k2 = S.keys[cur][j]; smaller += k2 < key; equal += k2 == key; }
                if (smaller <= r && r < smaller + equal) S.res[0] = key;
                if (smaller <= r + 1 && r + 1 < smaller + equal) S.res[1] = key;
            }
            __syncthreads();
            klo = S.res[0];
            if (want_hi) { khi = S.res[1]; have_hi = true; }
        }
    }
    float med = key2f(klo + kmin);
    if (need_hi) {
        if (!have_hi) {            // the smallest key larger than every candidate: one more streaming pass (rare)
            unsigned mn = 0xFFFFFFFFu;
            const unsigned pp = p;
            const int sh = s;
            for_each_key([&](unsigned key, bool ok) { if (ok && (sh >= 32 ? false : (key >> sh) > pp)) mn = min(mn, key); });
            mn = __reduce_min_sync(FULL, mn);
            __syncthreads();
            if (tid == 0) S.mn = 0xFFFFFFFFu;
            __syncthreads();
            if (lane == 0) atomicMin(&S.mn, mn);
            __syncthreads();
            khi = S.mn;
        }
        med = __fmul_rn(__fadd_rn(med, key2f(khi + kmin)), 0.5f);   // float32 mean of the two middles
    }
    if (tid == 0) c.median[(size_t)b * L + l] = med;
}

// ---------------------------------------------------------------------------------------------------
// exact squared Euclidean distance transform
// ---------------------------------------------------------------------------------------------------
// (1) The full field (lg_edt_squared with an output image): column pass without a distance image + exact row search.
// For every column the kernel keeps the source pixels as vertical bit words (bit r of word yw = row 32 yw + r is a
// source) plus, per word, the distance from its first row to the nearest source strictly above (vup) and from its last
// row to the nearest source strictly below (vdn).  The column distance g(x, y) of any pixel is then three coalesced loads
// and a count-leading / find-first on the word (g_of), so the row pass computes it where it needs it.
struct SrcMaskZero {        // source = zero pixel of a caller-supplied u8 mask (lg_edt_squared)
    const uint8_t* mask;
    size_t P;
    int W;
    __device__ __forceinline__ bool at(int b, int x, int y) const { return mask[(size_t)b * P + (size_t)y * W + x] == 0; }
};

struct EdtCols {            // one frame's column-pass results
    const uint32_t* vbits;  // [Hw][W]
    const uint16_t* vup;    // [Hw][W]
    const uint16_t* vdn;    // [Hw][W]
    int W;
    // column distance of pixel (x, y): 0 on a source, 0xFFFF when the column has none
    __device__ __forceinline__ unsigned g_of(int x, int y) const {
        const int r = y & 31;
        const size_t o = (size_t)(y >> 5) * W + x;
        const unsigned w = vbits[o];
        if ((w >> r) & 1u) return 0u;
        const unsigned mu = w << (31 - r), md = w >> r;
        const unsigned du = mu ? (unsigned)__clz(mu) : (unsigned)r + vup[o];
        const unsigned dd = md ? (unsigned)(__ffs(md) - 1) : (unsigned)(31 - r) + vdn[o];
        return min(min(du, dd), 0xFFFFu);
    }
};
__device__ __forceinline__ EdtCols edt_cols_of(const lg_context& c, int b) {
    const size_t o = (size_t)b * c.Hw * c.W;
    return EdtCols{c.vbits + o, c.vup + o, c.vdn + o, c.W};
}

constexpr int VC_NT = 128;
template <class SRC>
__global__ void __launch_bounds__(VC_NT) edt_vcol_kernel(lg_context c, SRC src) {
    const int W = c.W, H = c.H, Hw = c.Hw, b = blockIdx.y;
    const int x = blockIdx.x * VC_NT + threadIdx.x;
    if (x >= W) return;
    const size_t fo = (size_t)b * Hw * W;
    uint32_t* vb = c.vbits + fo;
    uint16_t* vu = c.vup + fo;
    uint16_t* vd = c.vdn + fo;
    // downward: the words, and the distance to the nearest source above each word
    unsigned since = 0xFFFFu;                  // distance from the row above the current word to the nearest source at or above it
    for (int yw = 0; yw < Hw; ++yw) {
        unsigned w = 0;
        const int ybase = yw << 5;
        if (ybase + 32 <= H) {
#pragma unroll
            for (int r = 0; r < 32; ++r) w |= (src.at(b, x, ybase + r) ? 1u : 0u) << r;
        } else {
            for (int r = 0; ybase + r < H; ++r) w |= (src.at(b, x, ybase + r) ? 1u : 0u) << r;
        }
        vb[(size_t)yw * W + x] = w;
        vu[(size_t)yw * W + x] = (uint16_t)min(since + 1u, 0xFFFFu);
        since = w ? (unsigned)__clz(w) : min(since + 32u, 0xFFFFu);
    }
    // upward: the distance to the nearest source below each word
    unsigned below = 0xFFFFu;                  // distance from the current word's last row to the nearest source strictly below
    for (int yw = Hw - 1; yw >= 0; --yw) {
        const unsigned w = vb[(size_t)yw * W + x];
        vd[(size_t)yw * W + x] = (uint16_t)below;
        below = w ? (unsigned)__ffs(w) : min(below + 32u, 0xFFFFu);
    }
}

// exact search for pixel x of a row whose squared column distances are in srow (shared memory): d2(x) = min over x' of
// (x - x')^2 + g(x')^2
__device__ __forceinline__ unsigned edt_row_search(const unsigned* srow, int W, int x) {
    unsigned bestd = srow[x];
    for (unsigned k = 1; k * k < bestd; ++k) {
        const int xl = x - (int)k, xr = x + (int)k;
        if (xl < 0 && xr >= W) break;
        const unsigned kk = k * k;
        if (xl >= 0) { const unsigned s = srow[xl]; if (s != 0xFFFFFFFFu) bestd = min(bestd, s + kk); }
        if (xr < W) { const unsigned s = srow[xr]; if (s != 0xFFFFFFFFu) bestd = min(bestd, s + kk); }
    }
    return bestd;
}

constexpr int EDT_NT = 256;
// Row pass of the full field; also keeps the frame's first maximum in best[] (packed d2 << 32 | ~index).
__global__ void __launch_bounds__(EDT_NT) edt_row_kernel(lg_context c, uint32_t* __restrict__ d2out,
                                                          unsigned long long* __restrict__ best) {
    extern __shared__ unsigned srow[];   // g squared, 0xFFFFFFFF = no source in that column
    __shared__ unsigned long long sbest[EDT_NT / 32];
    const int b = blockIdx.x, W = c.W, H = c.H;
    const size_t P = c.P;
    const EdtCols cols = edt_cols_of(c, b);
    unsigned long long mybest = 0;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        int any = 0;
        for (int x = threadIdx.x; x < W; x += EDT_NT) {
            const unsigned v = cols.g_of(x, y);
            srow[x] = (v == 0xFFFFu) ? 0xFFFFFFFFu : v * v;
            any |= (v != 0xFFFFu);
        }
        any = __syncthreads_or(any);     // srow complete; does the row see any source at all?
        if (any) {
            for (int x = threadIdx.x; x < W; x += EDT_NT) {
                const unsigned bestd = edt_row_search(srow, W, x);
                const size_t idx = (size_t)y * W + x;
                d2out[(size_t)b * P + idx] = bestd;
                mybest = max(mybest, ((unsigned long long)bestd << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)idx));
            }
        } else {
            for (int x = threadIdx.x; x < W; x += EDT_NT) d2out[(size_t)b * P + (size_t)y * W + x] = 0xFFFFFFFFu;
        }
        __syncthreads();                 // srow is rewritten by the next row
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mybest = max(mybest, __shfl_xor_sync(FULL, mybest, d));
    if ((threadIdx.x & 31) == 0) sbest[threadIdx.x >> 5] = mybest;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 0; w < EDT_NT / 32; ++w) mybest = max(mybest, sbest[w]);
        if (mybest) atomicMax(&best[b], mybest);
    }
}

// (2) The arg-max alone - the only thing leaf_scorer.py:67-71 takes from its distance field: the background pixel
// farthest from every leaf.  Exact branch and bound on the bit mask of the sources, one CTA per frame, no distance image:
//
//   * Cells.  The frame is cut into square cells of s x s pixels (s = 16 at 1440 x 1080; occupancy from the bytes of
//     leaf_band_kernel); a cell is occupied when it holds a source.  A small two-pass distance transform in shared memory
//     gives every cell its squared distance D (in cells) to the nearest occupied cell.  Every source lies in an occupied
//     cell and every occupied cell holds a source, hence for every pixel of a cell  s (sqrt(D) - sqrt 2) <= d <= s (sqrt(D) + sqrt 2).
//   * Nodes.  The distance field is 1-Lipschitz: a square node whose centre pixel has the exact distance dc holds no pixel
//     farther than dc + rad (rad = distance from the centre to the node's farthest pixel).  A node whose bound is below
//     the best exact distance found so far is dropped - it can neither hold the maximum nor tie with it; otherwise its
//     four quarters are examined, down to single pixels.  Every examined centre is a candidate itself (packed
//     d2 << 32 | ~index, so that among equal distances the first pixel in raster order wins), hence the result is exactly
//     the first maximum of the full field whatever the order of the search.  The order only decides how much is examined:
//     cells go best first (rounds of descending D), the levels below breadth first (all nodes of a level are measured by
//     all warps before any is quartered, so the pruning bound is the best of the whole level).
//   * Exact distance of one pixel (a warp): the parent's exact distance brackets the pixel's own, dc - dd <= d <= dc + dd
//     (dd = distance between the centres), so the nearest source lies in a thin ring around the pixel.  Lanes take the rows
//     dy = 0, +1, -1, +2, ... and look, left and right, at the few words of the row's bit mask that the ring covers
//     (count-leading / find-first; the loads of a row are independent of each other).
//
// Around the maximum the field falls off with slope ~1, so only the nodes within `rad` of it survive on every level: a
// frame costs on the order of a hundred exact evaluations instead of a transform of all its pixels.
constexpr int AM_MAX_NT = 512;
constexpr int AM_MAX_CELLS = 8192;
constexpr int AM_ITEMS = 1536;          // nodes of one level of the search
constexpr int AM_STACK = 48;            // per-warp node stack of the overflow path (depth first: at most 3 entries per level + 4)

struct AmBits {            // one frame's source bits: row y = 32-bit words at bits + y * pitch (bytes), bit x % 32 of word x / 32
    const uint8_t* bits;
    int pitch, W, H;
};

__device__ __forceinline__ unsigned isqrt_floor(unsigned v) {
    unsigned r = (unsigned)__fsqrt_rn((float)v);            // within 1 of the floor for every 32-bit v
    if ((unsigned long long)r * r > v) --r;
    else if ((unsigned long long)(r + 1u) * (r + 1u) <= v) ++r;
    return r;
}

// One row of the exact search below: where to look (issue: computes the ring's two column intervals and loads up to four
// words of the row's bit mask per side - all loads independent) and what was found (resolve).
struct AmRow {
    const uint32_t* rw;
    unsigned a[4], c[4], dy2;
    int hx, xmin, xmax, wl, wr, wmin, wmax;
    bool act_l, act_r;
    unsigned mask_l, mask_r;
};
__device__ __forceinline__ void am_row_issue(AmRow& q, const AmBits& B, int x, int y, int i, unsigned R, unsigned best, unsigned L2) {
    const int ady = (i + 1) >> 1;
    const int yy = (i & 1) ? y + ady : y - ady;
    q.act_l = q.act_r = false;
    q.hx = -1;
#pragma unroll
    for (int u = 0; u < 4; ++u) { q.a[u] = 0; q.c[u] = 0; }
    if ((unsigned)ady > R || yy < 0 || yy >= B.H) return;
    q.dy2 = (unsigned)(ady * ady);
    // dx <= hx  <=  dx^2 + dy^2 < best;  dx < lx  =>  dx^2 + dy^2 < L^2: no source there.  Both roots may be off by
    // one towards the safe side (a few more columns looked at; a candidate that is no improvement changes nothing).
    q.hx = (int)__fsqrt_ru((float)(best - 1u - q.dy2)) + 1;
    const int lx = q.dy2 < L2 ? max((int)__fsqrt_rd((float)(L2 - q.dy2)) - 1, 0) : 0;
    q.rw = reinterpret_cast<const uint32_t*>(B.bits + (size_t)yy * B.pitch);
    q.xmin = max(x - q.hx, 0); q.xmax = min(x + q.hx, B.W - 1);
    const int xl = x - lx, xr = x + lx;                            // first columns worth a look on either side
    q.act_l = xl >= q.xmin; q.act_r = xr <= q.xmax;
    q.wl = q.act_l ? xl >> 5 : 0; q.wr = q.act_r ? xr >> 5 : 0;
    q.mask_l = q.act_l ? 0xFFFFFFFFu >> (31 - (xl & 31)) : 0u; q.mask_r = q.act_r ? 0xFFFFFFFFu << (xr & 31) : 0u;
    q.wmin = q.xmin >> 5; q.wmax = q.xmax >> 5;
    if (q.act_l) {
#pragma unroll
        for (int u = 0; u < 4; ++u) if (q.wl - u >= q.wmin) q.a[u] = q.rw[q.wl - u];
    }
    if (q.act_r) {
#pragma unroll
        for (int u = 0; u < 4; ++u) if (q.wr + u <= q.wmax) q.c[u] = q.rw[q.wr + u];
    }
}
__device__ __forceinline__ unsigned am_row_resolve(AmRow& q, int x) {
    if (q.hx < 0) return 0xFFFFFFFFu;
    int col_l = -1, col_r = 0x7FFFFFFF;
    while (q.act_l || q.act_r) {                       // one turn unless the ring covers more than four words of the row
        if (q.act_l) {
            q.a[0] &= q.mask_l; q.mask_l = 0xFFFFFFFFu;
#pragma unroll
            for (int u = 3; u >= 0; --u) if (q.a[u]) col_l = 32 * (q.wl - u) + 31 - __clz(q.a[u]);
            q.wl -= 4;
            q.act_l = col_l < 0 && q.wl >= q.wmin;
        }
        if (q.act_r) {
            q.c[0] &= q.mask_r; q.mask_r = 0xFFFFFFFFu;
#pragma unroll
            for (int u = 3; u >= 0; --u) if (q.c[u]) col_r = 32 * (q.wr + u) + __ffs(q.c[u]) - 1;
            q.wr += 4;
            q.act_r = col_r == 0x7FFFFFFF && q.wr <= q.wmax;
        }
        if (q.act_l) {
#pragma unroll
            for (int u = 0; u < 4; ++u) q.a[u] = q.wl - u >= q.wmin ? q.rw[q.wl - u] : 0u;
        }
        if (q.act_r) {
#pragma unroll
            for (int u = 0; u < 4; ++u) q.c[u] = q.wr + u <= q.wmax ? q.rw[q.wr + u] : 0u;
        }
    }
    int nd = 0x7FFFFFFF;
    if (col_l >= q.xmin) nd = x - col_l;
    if (col_r <= q.xmax) nd = min(nd, col_r - x);
    return nd <= q.hx ? q.dy2 + (unsigned)nd * (unsigned)nd : 0xFFFFFFFFu;      // nd < W <= 4096: no overflow
}

// Exact squared distance from pixel (x, y) to the nearest source, given that a source exists within distance U and none
// is nearer than L (both in pixels; L = 0 and any valid U always work).  Stops early and returns some value < stop2 - still
// the squared distance to SOME source - as soon as the distance is known to be below stop2.  All lanes of the warp call
// with the same arguments and get the same result.  A lane takes two rows per turn; their loads are in flight together.
template <int ROWS>
__device__ __forceinline__ unsigned am_exact_d2(const AmBits& B, int x, int y, unsigned U, unsigned L, unsigned stop2, int lane) {
    unsigned best = U >= 0xFFFFu ? 0xFFFFFFFFu : U * U + 1u;     // strictly-better search: a source at exactly U must be found
    const unsigned L2 = L >= 0xFFFFu ? 0u : L * L;
    for (int i0 = 0;; i0 += 32 * ROWS) {
        const unsigned R = isqrt_floor(best - 1u);       // rows with |dy| <= R can still improve
        if ((unsigned)((i0 + 1) >> 1) > R) break;
        AmRow q[ROWS];
#pragma unroll
        for (int k = 0; k < ROWS; ++k) am_row_issue(q[k], B, x, y, i0 + 32 * k + lane, R, best, L2);
        unsigned cand = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < ROWS; ++k) cand = min(cand, am_row_resolve(q[k], x));
        cand = __reduce_min_sync(FULL, cand);
        best = min(best, cand);
        if (best < stop2) return best;
    }
    return best;
}

// a node of the search: e0 = x | y << 12 | log2(size) << 26 (corner, W <= 4096, H <= 16384); e1 = U | L << 16 (bounds on the
// distance of its centre); after it was measured: e1 = floor(dc) | 0x80000000 (exact) or 0 (dropped)
struct AmShared {
    unsigned long long best;       // packed (d2 << 32 | ~index) of the best exact pixel so far
    unsigned dmax;                 // largest cell distance (squared, in cells)
    unsigned n_items[2], next_item;
    unsigned items[2][AM_ITEMS][2];
    unsigned stack[AM_MAX_NT / 32][AM_STACK][2];
};

// bits: [n][H][pitch] source bit mask; occ: [n][n_bands][pitch] occupancy of the 8 x 8 blocks; best: [n] packed result
template <int ROWS, int AM_NT>
__global__ void __launch_bounds__(AM_NT, 2) edt_argmax_kernel(const uint8_t* __restrict__ bits_all, size_t bits_stride, int pitch,
                                                            const uint8_t* __restrict__ occ_all, int n_bands, int W, int H,
                                                            int cs /* cell size: 8, 16, 32, ... */, unsigned long long* __restrict__ best_out,
                                                            unsigned* dbg) {
    extern __shared__ __align__(16) unsigned char am_smem[];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cw = (W + cs - 1) / cs, ch = (H + cs - 1) / cs, cells = cw * ch;
    AmShared& S = *reinterpret_cast<AmShared*>(am_smem);
    uint16_t* D = reinterpret_cast<uint16_t*>(am_smem + sizeof(AmShared));      // [cells] squared cell distance (< 0x8000)
    uint8_t* dv = reinterpret_cast<uint8_t*>(D + cells);                        // [cells] vertical cell distance, 0xFF = none
    const AmBits B{bits_all + (size_t)b * bits_stride, pitch, W, H};
    const uint8_t* occ = occ_all + (size_t)b * n_bands * pitch;
    if (tid == 0) { S.best = 0ull; S.dmax = 0; S.n_items[0] = 0; S.n_items[1] = 0; S.next_item = 0; }
    __syncthreads();
    long long t_prev = clock64();
    int t_slot = 8;
    auto stamp = [&]() { if (dbg && tid == 0 && b == 0) { const long long t = clock64(); if (t_slot < 30) dbg[t_slot++] = (unsigned)(t - t_prev); t_prev = t; } };
    // ---- cells: occupancy = OR over the 8 x 8 blocks of the cell, then the vertical pass (one thread per cell column)
    const int k8 = cs >> 3;              // blocks per cell side
    const int bw8 = (W + 7) >> 3;
    int any = 0;
#pragma unroll 4
    for (int cidx = tid; cidx < cells; cidx += AM_NT) {
        const int i = cidx / cw, j = cidx - i * cw;
        unsigned o = 0;
        if (k8 == 2) {                   // 16 x 16 cells: two 16-bit loads (the pitch is even, so is the first block column)
            const int bi = 2 * i, bj = 2 * j;
            o = *reinterpret_cast<const uint16_t*>(occ + (size_t)bi * pitch + bj);
            if (bi + 1 < n_bands) o |= *reinterpret_cast<const uint16_t*>(occ + (size_t)(bi + 1) * pitch + bj);
            if (bj + 1 >= bw8) o &= 0xFFu;
        } else {
            for (int bi = i * k8; bi < min((i + 1) * k8, n_bands); ++bi)
                for (int bj = j * k8; bj < min((j + 1) * k8, bw8); ++bj) o |= occ[(size_t)bi * pitch + bj];
        }
        dv[cidx] = o ? 0u : 0xFFu;
        any |= o != 0;
    }
    any = __syncthreads_or(any);
    if (!any) {                          // no source at all: nothing to measure
        if (tid == 0) best_out[b] = 0ull;
        return;
    }
    for (int j = tid; j < cw; j += AM_NT) {
        unsigned d = 0xFFu;
        for (int i = 0; i < ch; ++i) {
            d = dv[i * cw + j] == 0 ? 0u : min(d + 1u, 0xFFu);
            dv[i * cw + j] = (uint8_t)d;
        }
        d = 0xFFu;
        for (int i = ch - 1; i >= 0; --i) {
            d = min((unsigned)dv[i * cw + j], min(d + 1u, 0xFFu));
            dv[i * cw + j] = (uint8_t)d;
        }
    }
    __syncthreads();
    // ---- cells: row pass, pruned scan of the lower envelope
    unsigned my_dmax = 0;
    for (int cidx = tid; cidx < cells; cidx += AM_NT) {
        const int i = cidx / cw, j = cidx - i * cw;
        const uint8_t* row = dv + i * cw;
        const unsigned v0 = row[j];
        unsigned bestd = v0 == 0xFFu ? 0xFFFFu : v0 * v0;
        for (unsigned k = 1; k * k < bestd; ++k) {
            const int jl = j - (int)k, jr = j + (int)k;
            if (jl < 0 && jr >= cw) break;
            const unsigned kk = k * k;
            if (jl >= 0) { const unsigned v = row[jl]; if (v != 0xFFu) bestd = min(bestd, v * v + kk); }
            if (jr < cw) { const unsigned v = row[jr]; if (v != 0xFFu) bestd = min(bestd, v * v + kk); }
        }
        D[cidx] = (uint16_t)bestd;       // < 0x8000 once the frame has a source: every cell then sees one
        my_dmax = max(my_dmax, bestd);
    }
    my_dmax = __reduce_max_sync(FULL, my_dmax);
    if (lane == 0) atomicMax(&S.dmax, my_dmax);
    __syncthreads();
    stamp();    // coarse done
    // Every pixel of the farthest cell is at least s (sqrt(Dmax) - sqrt 2) away: the search starts with that bound.
    const double SQ2 = 1.41421356237309515;
    const double r_max = sqrt((double)S.dmax);
    const double lb0 = fmax((double)cs * (r_max - SQ2) - 1e-6, 0.0);
    const unsigned lb0_2 = (unsigned)floor(lb0 * lb0);
    if (tid == 0 && lb0_2 > 0) S.best = (unsigned long long)lb0_2 << 32;      // a bound, not a pixel: any exact pixel >= it replaces it
    __syncthreads();
    auto lower_bound = [&]() -> double {   // distance of the best exact pixel so far (or the initial bound)
        return sqrt((double)(unsigned)(*reinterpret_cast<volatile unsigned long long*>(&S.best) >> 32));
    };
    // geometry of a node: centre pixel and the distance from it to the node's farthest pixel
    auto node_centre = [&](int nx0, int ny0, int sz, int& px, int& py, double& rad) {
        const int nx1 = min(nx0 + sz, W), ny1 = min(ny0 + sz, H);              // exclusive
        px = nx0 + ((nx1 - nx0) >> 1); py = ny0 + ((ny1 - ny0) >> 1);
        const int mx = max(px - nx0, nx1 - 1 - px), my = max(py - ny0, ny1 - 1 - py);
        rad = sqrt((double)(mx * mx + my * my));
    };
    // exact distance of a node's centre (bounds U, L on it that are known to hold); the centre becomes a candidate.
    // Returns d2, or a value < stop2 when the node cannot reach the best so far.
    auto eval_node = [&](int px, int py, double rad, unsigned U, unsigned L, unsigned& stop2) -> unsigned {
        const unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&S.best);
        const double lbd = sqrt((double)(unsigned)(cur >> 32));
        stop2 = 0;
        if (lbd - rad - 1e-6 > 0.0) { const double t = lbd - rad - 1e-6; stop2 = (unsigned)floor(t * t); }
        const unsigned d2 = am_exact_d2<ROWS>(B, px, py, U, L, stop2, lane);
        if (d2 >= stop2) {
            const unsigned idx = (unsigned)((size_t)py * W + px);
            const unsigned long long packed = ((unsigned long long)d2 << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
            if (lane == 0 && packed > cur) atomicMax(&S.best, packed);
        }
        return d2;
    };
    // the quarters of a node whose centre (px, py) has a distance in [dlo, dhi]: lane 0 only
    auto push_quarters = [&](int nx0, int ny0, int lg, int px, int py, unsigned dlo, unsigned dhi, unsigned (*out)[2], int& n, int cap) {
        const int h = 1 << (lg - 1);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int qx = nx0 + (q & 1) * h, qy = ny0 + (q >> 1) * h;
            if (qx >= W || qy >= H) continue;
            int cx2, cy2;
            double r2;
            node_centre(qx, qy, h, cx2, cy2, r2);
            const unsigned dd = (unsigned)ceil(sqrt((double)((cx2 - px) * (cx2 - px) + (cy2 - py) * (cy2 - py))) + 1e-9);
            if (n < cap) {
                out[n][0] = (unsigned)qx | ((unsigned)qy << 12) | ((unsigned)(lg - 1) << 26);
                out[n][1] = min(dhi + dd, 0xFFFFu) | ((dlo > dd ? dlo - dd : 0u) << 16);
                ++n;
            }
        }
    };
    unsigned (*stk)[2] = S.stack[warp];
    // depth-first search of one node by this warp (only when a level's list is full)
    auto dfs = [&](unsigned r0, unsigned r1) {
        int sp = 1;
        if (lane == 0) { stk[0][0] = r0; stk[0][1] = r1; }
        __syncwarp();
        while (sp > 0) {
            --sp;
            const unsigned e0 = stk[sp][0], e1 = stk[sp][1];
            __syncwarp();
            const int nx0 = (int)(e0 & 0xFFFu), ny0 = (int)((e0 >> 12) & 0x3FFFu), lg = (int)(e0 >> 26), sz = 1 << lg;
            int px, py;
            double rad;
            node_centre(nx0, ny0, sz, px, py, rad);
            unsigned stop2;
            const unsigned d2 = eval_node(px, py, rad, e1 & 0xFFFFu, e1 >> 16, stop2);
            if (dbg && lane == 0) atomicAdd(&dbg[1], 1u);
            if (d2 < stop2 || lg == 0) continue;
            const unsigned f = isqrt_floor(d2), cdc = f + (f * f != d2);
            if ((double)cdc + rad + 1e-6 < lower_bound()) continue;
            if (lane == 0) push_quarters(nx0, ny0, lg, px, py, f, cdc, stk, sp, AM_STACK);
            sp = __shfl_sync(FULL, sp, 0);
            __syncwarp();
        }
    };
    // ---- level 0, best first: the centres of the cells in rounds of descending cell distance (one cell of distance per
    // round), until the cells that are left cannot reach the best exact distance found.  An examined cell keeps, in place
    // of D, 0x8000 | exact << 14 | ceil(distance of its centre) - when the search stopped early the distance to some
    // source: a bound that holds.
    const int groups = (cells + 31) >> 5;
    unsigned hi2 = 0x8000u;                                   // this round takes the cells with lo2 <= D < hi2
    for (int k = 1;; ++k) {
        const double r_lo = r_max - (double)k;
        const unsigned lo2 = r_lo > 0.0 ? (unsigned)floor(r_lo * r_lo) : 0u;
        // a cell can reach the bound iff s (sqrt(D) + sqrt 2) >= bound, i.e. D >= (bound / s - sqrt 2)^2: one integer
        // threshold per round (the bound at the start of the round; it only grows, so this keeps a few cells too many)
        unsigned thr2 = lo2;
        {
            const double t = lower_bound() / (double)cs - SQ2 - 1e-6;
            if (t > 0.0) thr2 = max(thr2, (unsigned)floor(t * t));
        }
        for (int g = warp; g < groups; g += AM_NT / 32) {
            const int cidx = lane * groups + g;               // neighbouring cells (the top ones are) go to different warps
            const unsigned dcell = cidx < cells ? (unsigned)D[cidx] : 0xFFFFu;
            unsigned todo = __ballot_sync(FULL, dcell >= thr2 && dcell < hi2);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int cc = src * groups + g;
                const int ci = cc / cw, cj = cc - ci * cw;
                const double rc = sqrt((double)__shfl_sync(FULL, dcell, src));
                const unsigned U0 = (unsigned)ceil((double)cs * (rc + SQ2) + 1e-6);
                const unsigned L0 = (unsigned)floor(fmax((double)cs * (rc - SQ2) - 1e-6, 0.0));
                int px, py;
                double rad;
                node_centre(cj * cs, ci * cs, cs, px, py, rad);
                unsigned stop2;
                const unsigned d2 = eval_node(px, py, rad, min(U0, 0xFFFFu), L0, stop2);
                if (dbg && lane == 0) atomicAdd(&dbg[0], 1u);
                const unsigned f = isqrt_floor(d2), cdc = f + (f * f != d2);
                if (lane == 0) D[cc] = (uint16_t)(0x8000u | (d2 >= stop2 && cdc < 0x3FFFu ? 0x4000u : 0u) | min(cdc, 0x3FFFu));
            }
        }
        __syncthreads();
        stamp();    // one round
        if (lo2 == 0u) break;
        if ((double)cs * (sqrt((double)lo2) + SQ2) + 1e-6 < lower_bound()) break;    // uniform: S.best is stable here
        hi2 = lo2;
        __syncthreads();
    }
    __syncthreads();
    // ---- below level 0, breadth first: the quarters of the examined cells that can still reach the best form the first
    // list; a level's nodes are measured by all warps, then the survivors are quartered into the next list
    int lg0 = 0;
    while ((1 << lg0) < cs) ++lg0;
    double rad0;
    { int tx, ty; node_centre(0, 0, cs, tx, ty, rad0); }       // a clipped cell at the frame's edge only has a smaller one
    int cur = 0;
    // an examined cell is quartered iff ceil(dc) + rad0 >= bound (S.best is stable here); a saturated entry (0x3FFF) always is
    const unsigned alive_thr = min((unsigned)fmax(floor(lower_bound() - rad0 - 1e-6), 0.0), 0x3FFFu);
    for (int g = warp; g < groups; g += AM_NT / 32) {
        const int cidx = g * 32 + lane;
        const unsigned e = cidx < cells ? (unsigned)D[cidx] : 0u;
        unsigned alive = __ballot_sync(FULL, (e & 0x8000u) && (e & 0x3FFFu) >= alive_thr);
        while (alive) {
            const int src = __ffs(alive) - 1;
            alive &= alive - 1;
            const int cc = g * 32 + src;
            const int ci = cc / cw, cj = cc - ci * cw;
            const unsigned ee = __shfl_sync(FULL, e, src);
            const unsigned dhi = (ee & 0x3FFFu) >= 0x3FFFu ? 0xFFFFu : (ee & 0x3FFFu);
            const unsigned dlo = (ee & 0x4000u) ? (dhi > 0 ? dhi - 1u : 0u) : 0u;
            int px, py, n = 0;
            double rad;
            node_centre(cj * cs, ci * cs, cs, px, py, rad);
            unsigned q4[4][2];
            if (lane == 0) push_quarters(cj * cs, ci * cs, lg0, px, py, dlo, dhi, q4, n, 4);
            n = __shfl_sync(FULL, n, 0);
            unsigned slot = 0;
            if (lane == 0) slot = atomicAdd(&S.n_items[cur], (unsigned)n);
            slot = __shfl_sync(FULL, slot, 0);
            // The list takes as many of the quarters as still fit - every slot below n_items is then a node that was really
            // written (a node with fewer than four quarters at the frame's edge must not leave a hole of stale entries at
            // the end of the list) - and the rest is searched right here.
            const int fit = slot >= (unsigned)AM_ITEMS ? 0 : min(n, (int)((unsigned)AM_ITEMS - slot));
            if (lane == 0)
                for (int q = 0; q < fit; ++q) { S.items[cur][slot + q][0] = q4[q][0]; S.items[cur][slot + q][1] = q4[q][1]; }
            for (int q = fit; q < n; ++q) {
                const unsigned r0 = __shfl_sync(FULL, q4[q][0], 0), r1 = __shfl_sync(FULL, q4[q][1], 0);
                dfs(r0, r1);
            }
        }
    }
    __syncthreads();
    stamp();    // list built
    for (int lg = lg0 - 1; lg >= 0; --lg) {
        const unsigned n_cur = min(S.n_items[cur], (unsigned)AM_ITEMS);
        if (n_cur == 0) break;                                  // uniform
        // measure
        for (;;) {
            unsigned it = 0;
            if (lane == 0) it = atomicAdd(&S.next_item, 1u);
            it = __shfl_sync(FULL, it, 0);
            if (it >= n_cur) break;
            const unsigned e0 = S.items[cur][it][0], e1 = S.items[cur][it][1];
            const int nx0 = (int)(e0 & 0xFFFu), ny0 = (int)((e0 >> 12) & 0x3FFFu), sz = 1 << lg;
            int px, py;
            double rad;
            node_centre(nx0, ny0, sz, px, py, rad);
            unsigned stop2;
            const unsigned d2 = eval_node(px, py, rad, e1 & 0xFFFFu, e1 >> 16, stop2);
            if (dbg && lane == 0) { atomicAdd(&dbg[1], 1u); atomicAdd(&dbg[2 + min(lg, 5)], 1u); }
            if (lane == 0) S.items[cur][it][1] = d2 >= stop2 ? (0x80000000u | d2 >> 0) : 0u;   // exact d2 (< 2^31: d < 46340) or dropped
        }
        __syncthreads();
        stamp();    // level measured
        if (tid == 0) { S.next_item = 0; S.n_items[cur ^ 1] = 0; }
        __syncthreads();
        if (lg == 0) break;
        // quarter the survivors (the bound is now the best of the whole level)
        for (unsigned i0 = warp * 32; i0 < n_cur; i0 += AM_NT) {
            const unsigned it = i0 + lane;
            const unsigned e0 = it < n_cur ? S.items[cur][it][0] : 0u, e1 = it < n_cur ? S.items[cur][it][1] : 0u;
            bool keep = false;
            int px = 0, py = 0;
            unsigned f = 0, cdc = 0;
            const int nx0 = (int)(e0 & 0xFFFu), ny0 = (int)((e0 >> 12) & 0x3FFFu);
            if (e1 & 0x80000000u) {
                const unsigned d2 = e1 & 0x7FFFFFFFu;
                double rad;
                node_centre(nx0, ny0, 1 << lg, px, py, rad);
                f = isqrt_floor(d2); cdc = f + (f * f != d2);
                keep = (double)cdc + rad + 1e-6 >= lower_bound();
            }
            int fit = 0;                                        // quarters of this lane's node that went to the list
            if (keep) {
                unsigned q4[4][2];
                int n = 0;
                push_quarters(nx0, ny0, lg, px, py, f, cdc, q4, n, 4);
                const unsigned slot = atomicAdd(&S.n_items[cur ^ 1], (unsigned)n);
                // as many as still fit: no slot below n_items stays unwritten (see the first list above)
                fit = slot >= (unsigned)AM_ITEMS ? 0 : min(n, (int)((unsigned)AM_ITEMS - slot));
                for (int q = 0; q < fit; ++q) { S.items[cur ^ 1][slot + q][0] = q4[q][0]; S.items[cur ^ 1][slot + q][1] = q4[q][1]; }
                keep = fit < n;                                 // all placed?
            }
            // quarters that did not fit the list are searched depth first, one after the other by the whole warp
            unsigned left = __ballot_sync(FULL, keep);
            while (left) {
                const int src = __ffs(left) - 1;
                left &= left - 1;
                const int sx0 = __shfl_sync(FULL, nx0, src), sy0 = __shfl_sync(FULL, ny0, src);
                const int spx = __shfl_sync(FULL, px, src), spy = __shfl_sync(FULL, py, src);
                const unsigned sf = __shfl_sync(FULL, f, src), sc = __shfl_sync(FULL, cdc, src);
                const int sfit = __shfl_sync(FULL, fit, src);
                unsigned q4[4][2];
                int n = 0;
                if (lane == 0) push_quarters(sx0, sy0, lg, spx, spy, sf, sc, q4, n, 4);     // same order as on lane `src`
                n = __shfl_sync(FULL, n, 0);
                for (int q = sfit; q < n; ++q) {
                    const unsigned r0 = __shfl_sync(FULL, q4[q][0], 0), r1 = __shfl_sync(FULL, q4[q][1], 0);
                    dfs(r0, r1);
                }
            }
        }
        __syncthreads();
        stamp();    // level quartered
        cur ^= 1;
    }
    __syncthreads();
    if (tid == 0) {
        // If no exact pixel reached the initial bound (cannot happen: the farthest cell's pixels do), the result would be the
        // bound itself with index bits 0: report it as "no pixel" the same way an empty frame is.
        const unsigned long long r = S.best;
        best_out[b] = (r & 0xFFFFFFFFull) ? r : 0ull;
    }
}

// source bits + block occupancy from a caller's u8 mask (source = zero pixel), for the arg-max search of lg_edt_squared
__global__ void mask_bits_kernel(const uint8_t* __restrict__ mask, size_t P, int W, int H, uint8_t* __restrict__ bits,
                                 size_t bits_stride, int pitch, uint8_t* __restrict__ occ, int n_bands) {
    const int b = blockIdx.y;
    const int bw8 = (W + 7) >> 3;
    const int blk = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per 8 x 8 block
    if (blk >= bw8 * n_bands) return;
    const int band = blk / bw8, bj = blk - band * bw8;
    const uint8_t* m = mask + (size_t)b * P;
    unsigned any = 0;
    for (int r = 0; r < 8; ++r) {
        const int y = band * 8 + r;
        if (y >= H) break;
        unsigned v = 0;
        for (int k = 0; k < 8; ++k) {
            const int x = bj * 8 + k;
            if (x < W && m[(size_t)y * W + x] == 0) v |= 1u << k;
        }
        bits[(size_t)b * bits_stride + (size_t)y * pitch + bj] = (uint8_t)v;
        any |= v;
    }
    occ[((size_t)b * n_bands + band) * pitch + bj] = any ? 1 : 0;
}

__global__ void edt_argmax_out_kernel(const unsigned long long* best, int32_t* argmax, int n) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n) argmax[b] = (int32_t)(0xFFFFFFFFu - (unsigned)(best[b] & 0xFFFFFFFFull));
}


// ---------------------------------------------------------------------------------------------------
// the pick
// ---------------------------------------------------------------------------------------------------
// numpy's pairwise float32 summation (np.mean of the medians, leaf_scorer.py:53-54)
__device__ float np_pairwise_sum_f32(const float* a, int n) {
    if (n < 8) {
        float r = 0.f;
        for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
        return r;
    }
    if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(np_pairwise_sum_f32(a, n2), np_pairwise_sum_f32(a + n2, n - n2));
}

struct SelScratch {   // per label, in shared memory
    double s0, s1, s2;
    float med;
    int id;
    int tall;
    int cand;
};

// One warp per frame: lanes take the labels (record reset, id list), then the leaves (scores, Pareto test); the two
// order-sensitive steps keep the reference's order - numpy's pairwise sum of the medians on one lane, and "first maximum
// wins" for the weighted pick (smallest list position among equal scores).
__global__ void __launch_bounds__(32) select_leaf_kernel(lg_context c, lg_camera cam, int32_t* leaf_out, lg_leaf_record* rec_out) {
    extern __shared__ unsigned char sraw[];
    const int b = blockIdx.x, L = c.L, W = c.W, H = c.H, lane = threadIdx.x;
    SelScratch* s = reinterpret_cast<SelScratch*>(sraw);
    float* meds = reinterpret_cast<float*>(sraw + sizeof(SelScratch) * L);
    const size_t o = (size_t)b * L;
    const uint32_t* cnt = c.cnt + o;
    LgRegion reg;
    reg.x0 = reg.y0 = reg.x1 = reg.y1 = 0; reg.ok = 0; reg.sx0 = reg.sy0 = reg.sx1 = reg.sy1 = 0;
    int best_id = -1;
    // background = the smallest id present; the other ids present, in ascending order
    int bg = -1, n_ids = 0;
    for (int l0 = 0; l0 < L; l0 += 32) {
        const int l = l0 + lane;
        const bool present = l < L && cnt[l] != 0;
        unsigned m = __ballot_sync(0xFFFFFFFFu, present);
        if (bg < 0 && m) { bg = l0 + __ffs(m) - 1; m &= m - 1; }      // the background is not listed
        const bool listed = present && l != bg;
        if (rec_out && l < L) {                                        // listed leaves are overwritten below
            lg_leaf_record r;
            memset(&r, 0, sizeof(r));
            r.leaf_id = l;
            rec_out[o + l] = r;
        }
        if (listed) {
            const int k = n_ids + __popc(m & ((1u << lane) - 1u));
            s[k].id = l;
            s[k].med = c.median[o + l];
            meds[k] = s[k].med;
        }
        n_ids += __popc(m);
    }
    __syncwarp();
    if (n_ids > 0 && !(c.status[b] & LG_ST_LABEL_RANGE)) {
        float mean_med = 0.f;
        if (lane == 0) mean_med = __fdiv_rn(np_pairwise_sum_f32(meds, n_ids), (float)n_ids);
        mean_med = __shfl_sync(0xFFFFFFFFu, mean_med, 0);
        const unsigned fl = c.first_leaf[b];
        const double pmin_x = (double)(fl % W), pmin_y = (double)(fl / W);
        const unsigned far = 0xFFFFFFFFu - (unsigned)(c.edt_best[b] & 0xFFFFFFFFull);
        const double pmax_x = (double)(far % W), pmax_y = (double)(far / W);
        int n_tall_c = 0, n_c = 0;
        for (int k = lane; k < n_ids; k += 32) {
            const int l = s[k].id;
            s[k].tall = s[k].med < mean_med;
            s[k].cand = 0;
            const unsigned area = cnt[l];
            lg_leaf_record r;
            memset(&r, 0, sizeof(r));
            r.leaf_id = l; r.area = area; r.median_depth = s[k].med; r.is_tall = s[k].tall;
            const double n = (double)area;
            const double cx = (double)c.sx[o + l] / n, cy = (double)c.sy[o + l] / n;
            r.centroid_x = cx; r.centroid_y = cy;
            const float md = (float)(((double)(long long)c.sdep[o + l] / DEP_SCALE) / n);
            r.mean_depth = md;
            if (area >= LG_MIN_LEAF_AREA) {
                const double dmin = sqrt((cx - pmin_x) * (cx - pmin_x) + (cy - pmin_y) * (cy - pmin_y));
                const double dmax = sqrt((cx - pmax_x) * (cx - pmax_x) + (cy - pmax_y) * (cy - pmax_y));
                const double tot = dmin + dmax;
                const double clutter = tot > 0 ? dmin / tot : 0.0;
                const double mean_dist = (double)md * (((double)c.sdist[o + l] / DIST_SCALE) / n);
                const double dist_score = exp(-mean_dist / 0.3);
                double vis = 0.0;
                if (!c.border[o + l]) {
                    const double hw = W / 2.0, hh = H / 2.0;
                    vis = 1.0 - sqrt((cx - hw) * (cx - hw) + (cy - hh) * (cy - hh)) / sqrt(hw * hw + hh * hh);
                }
                s[k].s0 = clutter; s[k].s1 = dist_score; s[k].s2 = vis; s[k].cand = 1;
                r.clutter = clutter; r.distance = dist_score; r.visibility = vis; r.mean_distance = mean_dist;
                r.is_candidate = 1;
                ++n_c;
                if (s[k].tall) ++n_tall_c;
            }
            if (rec_out) rec_out[o + l] = r;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            n_c += __shfl_xor_sync(0xFFFFFFFFu, n_c, d);
            n_tall_c += __shfl_xor_sync(0xFFFFFFFFu, n_tall_c, d);
        }
        __syncwarp();
        if (n_c > 0) {
            const int want_tall = n_tall_c > 0;
            const double scale = want_tall ? 1.1 : 1.0;
            double best_score = -CUDART_INF;
            int best_k = 0x7FFFFFFF;
            for (int i = lane; i < n_ids; i += 32) {          // ascending i per lane: a later equal score never replaces
                if (!s[i].cand || (want_tall && !s[i].tall)) continue;
                const double a0 = s[i].s0 * scale, a1 = s[i].s1 * scale, a2 = s[i].s2 * scale;
                bool dominated = false;
                for (int j = 0; j < n_ids && !dominated; ++j) {
                    if (j == i || !s[j].cand || (want_tall && !s[j].tall)) continue;
                    const double b0 = s[j].s0 * scale, b1 = s[j].s1 * scale, b2 = s[j].s2 * scale;
                    const bool ge = b0 >= a0 && b1 >= a1 && b2 >= a2;
                    const bool gt = b0 > a0 || b1 > a1 || b2 > a2;
                    if (ge && (gt || j < i)) dominated = true;
                }
                if (dominated) continue;
                const double w = ((0.0 + 0.35 * s[i].s0) + 0.35 * s[i].s1) + 0.3 * s[i].s2;
                if (w > best_score) { best_score = w; best_k = i; }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {                // maximum score, smallest list position among equals
                const double os = __shfl_xor_sync(0xFFFFFFFFu, best_score, d);
                const int ok = __shfl_xor_sync(0xFFFFFFFFu, best_k, d);
                if (os > best_score || (os == best_score && ok < best_k)) { best_score = os; best_k = ok; }
            }
            if (best_k != 0x7FFFFFFF) best_id = s[best_k].id;
        }
    }
    if (lane != 0) return;
    if (best_id >= 0) {
        reg.x0 = (int)c.bx0[o + best_id]; reg.x1 = (int)c.bx1[o + best_id] + 1;
        reg.y0 = (int)c.by0[o + best_id]; reg.y1 = (int)c.by1[o + best_id] + 1;
        reg.ok = 1;
        reg.sx0 = max(0, reg.x0 - LG_REGION_PAD); reg.sy0 = max(0, reg.y0 - LG_REGION_PAD);
        reg.sx1 = min(W, reg.x1 + LG_REGION_PAD); reg.sy1 = min(H, reg.y1 + LG_REGION_PAD);
    } else {
        atomicOr(&c.status[b], LG_ST_NO_LEAF);
    }
    c.leaf_id[b] = best_id;
    c.region[b] = reg;
    if (leaf_out) leaf_out[b] = best_id;
}

}  // namespace

// arg-max search of the distance transform of the sources in c->ubits / c->cellocc for n frames, on `st`
static int run_edt_argmax(lg_context* c, int n, cudaStream_t st) {
    const int cs = c->am_cs;
    const int cells = ((c->W + cs - 1) / cs) * ((c->H + cs - 1) / cs);
    const size_t smem = sizeof(AmShared) + (size_t)cells * 3;
    static unsigned* dbg = nullptr;
    static int dbg_on = -1;
    if (dbg_on < 0) { const char* e = getenv("LG_AM_DEBUG"); dbg_on = e && e[0] == '1'; if (dbg_on) { cudaMalloc(&dbg, 128); } }
    if (dbg_on) cudaMemsetAsync(dbg, 0, 128, st);
    // One row per lane and 16 warps per frame, compiled for two CTAs per SM (64 registers) so that a batch of 256 frames is
    // resident at once.  A/B switch LG_AM_ROWS=2: two rows per lane (124 registers), 8 warps per frame - measured slower
    // (0.54 vs 0.46 ms per 256 frames).
    static int rows = 0;
    if (!rows) { const char* e = getenv("LG_AM_ROWS"); rows = e && e[0] == '2' ? 2 : 1; }
    LG_PREFER_LARGE_SMEM((edt_argmax_kernel<1, 512>));       // runs beside the median kernel
    LG_PREFER_LARGE_SMEM((edt_argmax_kernel<2, 256>));
    if (rows == 2) {
        LG_ENSURE_SMEM((edt_argmax_kernel<2, 256>), smem);
        edt_argmax_kernel<2, 256><<<n, 256, smem, st>>>(c->ubits, c->ub_stride, c->ub_pitch, c->cellocc, c->n_bands, c->W, c->H, cs, c->edt_best, dbg_on ? dbg : nullptr);
    } else {
        LG_ENSURE_SMEM((edt_argmax_kernel<1, 512>), smem);
        edt_argmax_kernel<1, 512><<<n, 512, smem, st>>>(c->ubits, c->ub_stride, c->ub_pitch, c->cellocc, c->n_bands, c->W, c->H, cs, c->edt_best, dbg_on ? dbg : nullptr);
    }
    LG_LAUNCH_CHECK();
    if (dbg_on) {
        unsigned h[32];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, dbg, 128, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[am] frames %d cs %d: level-0 evals %u, below %u (by size 1:%u 2:%u 4:%u 8:%u 16:%u 32+:%u)\n", n, cs, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
        fprintf(stderr, "[am] frame 0 phase cycles:");
        for (int i = 8; i < 30; ++i) fprintf(stderr, " %u", h[i]);
        fprintf(stderr, "\n");
    }
    return LG_OK;
}

// cell size of the arg-max search: few enough cells for a cheap coarse pass, squared cell distances below 2^15
int lg_edt_cell_size(int H, int W) {
    int cs = 8;
    for (;; cs *= 2) {
        const long long cw = (W + cs - 1) / cs, ch = (H + cs - 1) / cs;
        if (cw * ch <= AM_MAX_CELLS && cw * cw + ch * ch < 32768) return cs;
    }
}

int lg_run_stage1(lg_context* c, const int16_t* labels, const float* depth, int n, lg_camera cam, cudaStream_t st) {
    clear_tables_kernel<<<64, 256, 0, st>>>(*c, n);
    LG_LAUNCH_CHECK();
    if (!c->ray_valid || c->ray_cam.f != cam.f || c->ray_cam.cx != cam.cx || c->ray_cam.cy != cam.cy) {
        ray_rows_kernel<<<(c->H + 7) / 8, 256, 0, st>>>(c->ray_tab, c->H, c->W, cam);
        LG_LAUNCH_CHECK();
        ray_cols_kernel<<<(c->W + 127) / 128, 128, 0, st>>>(c->ray_tab, c->H, c->W);
        LG_LAUNCH_CHECK();
        c->ray_cam = cam;
        c->ray_valid = 1;
    }
    // the one pass over labels + depth: per-leaf statistics, grouped depth values, union bit mask
    {
        const int nt = (((c->W + 7) / 8 + 31) / 32) * 32;                    // one thread per 8 columns
        const size_t smem = (size_t)LG_BAND * nt * sizeof(uint4) + c->L * (sizeof(SmemLeaf) + 2 * sizeof(unsigned)) + (size_t)nt * sizeof(unsigned short);
        const dim3 grid((unsigned)c->n_bands, n);
        const bool vec = (c->W % 8 == 0) && ((reinterpret_cast<uintptr_t>(labels) | reinterpret_cast<uintptr_t>(depth)) % 16 == 0);
        static int ctas = 0;      // A/B switch: resident CTAs the register allocation aims at for frames up to 1536 wide
        if (!ctas) { const char* e = getenv("LG_BAND_CTAS"); ctas = e ? atoi(e) : 6; }
        if (vec && nt <= 192 && ctas == 7) {
            LG_ENSURE_SMEM((leaf_band_kernel<true, 192, 7>), smem);
            leaf_band_kernel<true, 192, 7><<<grid, nt, smem, st>>>(*c, labels, depth);
        } else if (vec && nt <= 192 && ctas == 6) {
            LG_ENSURE_SMEM((leaf_band_kernel<true, 192, 6>), smem);
            leaf_band_kernel<true, 192, 6><<<grid, nt, smem, st>>>(*c, labels, depth);
        } else if (vec) {
            LG_ENSURE_SMEM((leaf_band_kernel<true, 512, 2>), smem);
            leaf_band_kernel<true, 512, 2><<<grid, nt, smem, st>>>(*c, labels, depth);
        } else {
            LG_ENSURE_SMEM((leaf_band_kernel<false, 512, 2>), smem);
            leaf_band_kernel<false, 512, 2><<<grid, nt, smem, st>>>(*c, labels, depth);
        }
        LG_LAUNCH_CHECK();
    }
    lg_mark(c, LG_M_STATS, st);
    // the distance-transform search of the union is independent of the medians: it runs beside them
    cudaStream_t aux = lg_fork(c, 0, st);
    int rc = run_edt_argmax(c, n, aux);
    lg_mark(c, LG_M_EDT_ROW, aux);
    if (!rc) {
        rc = lg_ensure_smem_impl((const void*)leaf_median_kernel, sizeof(MedShared));
        if (!rc) rc = lg_prefer_large_smem_impl((const void*)leaf_median_kernel);
        if (!rc) {
            leaf_median_kernel<<<dim3(c->L, n), MED_NT, sizeof(MedShared), st>>>(*c);
            ++g_lg_launches;
            if (cudaGetLastError() != cudaSuccess) { lg_set_error("leaf_median_kernel launch failed"); rc = LG_E_CUDA; }
        }
    }
    lg_mark(c, LG_M_MEDIAN, st);
    const int rcj = lg_join(c, 0, aux, st);        // also on an error path: the side stream must not stay unordered
    return rc ? rc : rcj;
}

int lg_run_select(lg_context* c, int n, lg_camera cam, int32_t* leaf_out, lg_leaf_record* rec_out, cudaStream_t st) {
    size_t sm = (sizeof(SelScratch) + sizeof(float)) * c->L;
    select_leaf_kernel<<<n, 32, sm, st>>>(*c, cam, leaf_out, rec_out);
    LG_LAUNCH_CHECK();
    lg_mark(c, LG_M_SELECT, st);
    return LG_OK;
}

extern "C" int lg_edt_squared(lg_context* c, const uint8_t* mask, int n, uint32_t* d2, int32_t* argmax, void* stream) {
    if (!c || !mask || n < 1) return LG_E_ARG;
    if (n > c->B) return LG_E_CAPACITY;
    cudaStream_t st = (cudaStream_t)stream;
    LG_CUDA(cudaMemsetAsync(c->edt_best, 0, sizeof(unsigned long long) * n, st));
    if (d2) {       // the full field
        const SrcMaskZero src{mask, c->P, c->W};
        edt_vcol_kernel<SrcMaskZero><<<dim3((c->W + VC_NT - 1) / VC_NT, n), VC_NT, 0, st>>>(*c, src);
        LG_LAUNCH_CHECK();
        int per_frame = (148 * 8 * 2 + n - 1) / n;
        per_frame = per_frame < 1 ? 1 : (per_frame > c->H ? c->H : per_frame);
        edt_row_kernel<<<dim3(n, per_frame), EDT_NT, c->W * sizeof(unsigned), st>>>(*c, d2, c->edt_best);
        LG_LAUNCH_CHECK();
    } else {        // the arg-max alone: the search the leaf selection uses, on the caller's mask
        const int blocks = ((c->W + 7) / 8) * c->n_bands;
        mask_bits_kernel<<<dim3((blocks + 127) / 128, n), 128, 0, st>>>(mask, c->P, c->W, c->H, c->ubits, c->ub_stride, c->ub_pitch,
                                                                         c->cellocc, c->n_bands);
        LG_LAUNCH_CHECK();
        int rc = run_edt_argmax(c, n, st);
        if (rc) return rc;
    }
    if (argmax) {
        edt_argmax_out_kernel<<<(n + 63) / 64, 64, 0, st>>>(c->edt_best, argmax, n);
        LG_LAUNCH_CHECK();
    }
    return LG_OK;
}
