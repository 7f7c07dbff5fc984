// Stage 1 - optimal-leaf selection (reference scripts/utils/leaf_scorer.py:25-203, 277-306).
//
// Kernels (all batched over frames, no host synchronisation):
//   leaf_rows_kernel     ONE pass over labels + depth, 128-bit loads, 8 pixels per thread: per-label pixel count, coordinate
//                        sums, depth sum, sum of ray lengths, bounding box, border contact, depth key range, first leaf
//                        pixel of the frame - and, from the same registers, the depth values of the leaf pixels grouped
//                        by label inside every image row (what the median needs) and the bit mask of the leaf union
//                        (what the distance transform needs).  Nothing re-reads the inputs.
//   leaf_median_kernel   exact np.median per label by radix selection over the label's per-row sub-blocks
//   edt_vcol_kernel      column pass of the exact squared Euclidean distance transform of the leaf union, from the bit
//                        mask: per column vertical bit words + the distance to the nearest leaf pixel above / below every
//                        32-row word (column distances are then three loads and a few bit operations, never stored), and
//                        the minima of the column distance over 32-column chunks per row and per block of 8 rows
//   edt_seed_kernel / edt_blockmax_kernel / edt_row_kernel   row pass as a pruned search for the background pixel
//                        farthest from every leaf (the only thing leaf_scorer.py:67-71 takes from its distance field)
//   select_leaf_kernel   the per-leaf scores, tall-leaf rule, Pareto front and weighted pick
//
// Integer sums are exact and order independent, so results do not depend on scheduling: coordinate
// sums are 64-bit integers, depth and ray-length sums are fixed point (2^-28 m and 2^-36).
#include <math_constants.h>

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
        c.kmin[i] = 0xFFFFFFFFu; c.kmax[i] = 0;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        c.first_leaf[i] = 0xFFFFFFFFu; c.edt_best[i] = 0ull; c.status[i] = 0; c.list_n[i] = 0;
    }
}

// ray_tab[y * W + x] = sum over columns x' <= x of round(2^36 * sqrt(((x' - cx)^2 + (y - cy)^2) / f^2 + 1)): row-wise prefix
// sums of the length of the viewing ray through a pixel per unit depth (leaf_scorer.py:104-113 with X = md (x - cx) / f,
// Y = md (y - cy) / f, Z = md).  The sum over a horizontal run of pixels is then a difference of two entries; the per-pixel
// terms are integers, so any grouping of the pixels gives the same total.  The table depends on the camera only: it is
// built once per camera and read (L2-resident) by every frame.  One warp per row: 32-pixel segments, warp scan, carry.
__global__ void ray_table_kernel(unsigned long long* __restrict__ tab, int H, int W, lg_camera cam) {
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

// ---------------------------------------------------------------------------------------------------
// the one pass over the inputs
// ---------------------------------------------------------------------------------------------------
// A CTA owns a range of image rows of one frame; one iteration handles one row ("tile"), a thread 8 consecutive pixels.
//
// Statistics.  Labels are piecewise constant (a leaf is 100-350 px wide): a thread's 8 pixels almost always carry one
// label, and the threads of a warp (256 px) one to three.  Everything that depends on the position only has a closed form
// (count, coordinate sums, bounding box, border contact) or is a difference of two ray_tab entries; per pixel only the depth
// is accumulated (2^-28 fixed point, exact and order independent) and its key range tracked.  Threads whose 8 pixels are
// uniform are reduced per label with full-warp redux operations (a short loop over the distinct labels of the warp), and
// one lane per label updates the CTA's shared-memory table; the few threads that contain a label boundary (~12 per row)
// add their runs directly.  What only depends on the row - pixel count, sum of y, vertical extent, top / bottom border -
// is added once per row and label by the warp that scans the row's counts.  The table goes to the frame's table once,
// when the CTA is done.
//
// Grouped depth values.  np.median needs the depths of every leaf's pixels.  Inside a row the leaf pixels are stored
// grouped by ascending label: seg[y * W + off[l] ...], with the row's offsets off[0..L] (u16) in tile_off.  The positions
// inside a row do not depend on any other row, so no frame-wide count or scan is needed before the values can be written,
// and the median kernel walks the rows of its label's bounding box.  Only labels >= 1 are stored: the reference drops the
// smallest id present (the background), which is 0 whenever 0 occurs at all.
//
// Union mask.  ubits[y][x / 8] bit x % 8 = (label of pixel (x, y) >= 1): one byte per thread and row.
constexpr int TS_PX = 8;

struct SmemLeaf {
    unsigned cnt, sx, sy, bx0, bx1, by0, by1, border, kmin, kmax;
    unsigned long long sdep, sdist;
};

template <bool VEC>
__global__ void __launch_bounds__(512, 2) leaf_rows_kernel(lg_context c, const int16_t* __restrict__ labels,
                                                           const float* __restrict__ depth, int rows_per_cta) {
    extern __shared__ __align__(16) unsigned char ts_smem[];
    __shared__ unsigned s_first, s_bad;
    const int L = c.L, W = c.W, H = c.H, NT = blockDim.x;
    const size_t P = c.P;
    SmemLeaf* tab = reinterpret_cast<SmemLeaf*>(ts_smem);
    unsigned* cur2 = reinterpret_cast<unsigned*>(tab + L);        // [2][L]: per-row pixel counts, then write cursors
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int l = tid; l < L; l += NT) {
        SmemLeaf z;
        z.cnt = 0; z.sx = 0; z.sy = 0; z.bx0 = 0xFFFFFFFFu; z.by0 = 0xFFFFFFFFu; z.bx1 = 0; z.by1 = 0;
        z.border = 0; z.kmin = 0xFFFFFFFFu; z.kmax = 0; z.sdep = 0; z.sdist = 0;
        tab[l] = z;
        cur2[l] = 0; cur2[L + l] = 0;
    }
    if (tid == 0) { s_first = 0xFFFFFFFFu; s_bad = 0; }
    __syncthreads();
    const int16_t* lp = labels + (size_t)b * P;
    const float* dp = depth + (size_t)b * P;
    float* seg = c.seg + (size_t)b * P;
    uint8_t* ub = c.ubits + (size_t)b * c.ub_stride;
    const int ubw = (W + 7) >> 3;
    const int row0 = blockIdx.x * rows_per_cta, row1 = min(row0 + rows_per_cta, H);
    uint16_t* toff_base = c.tile_off + (size_t)b * H * c.lstride;
    const int x0 = tid * TS_PX;
    const int npx = max(0, min(TS_PX, W - x0));
    const unsigned lanes_lt = (1u << lane) - 1u;
    const unsigned llim = (unsigned)L * 0x00010001u;      // L in both halfwords (L <= 1024)

    // 2^28 * depth is exact in float32 (power-of-two scale), so the fixed-point term equals the float64 formulation
    auto fixed28 = [](float v) -> long long { return __float2ll_rn(fminf(fmaxf(v, -2048.f), 2048.f) * 268435456.f); };

    // Items (label il >= 0 or -1 for none, pixels [ix, ix + ilen) of the current row, their depth sum and key range) are
    // reduced per distinct label of the warp with full-warp redux operations; one lane per label updates the shared-memory
    // table.  Returns the lanes that hold an item with this lane's label (0 for label <= 0: nothing to place).
    const unsigned long long* rrow = nullptr;
    unsigned* tcnt = nullptr;
    auto reduce_items = [&](int il, unsigned ix, unsigned ilen, long long isdep, unsigned ikmn, unsigned ikmx) -> unsigned {
        unsigned gmask = 0;
        unsigned todo = __ballot_sync(FULL, il >= 0);
        while (todo) {
            const int src = __ffs(todo) - 1;
            const int cl = __shfl_sync(FULL, il, src);
            const bool mine = il == cl;
            const unsigned gm = __ballot_sync(FULL, mine);
            todo &= ~gm;
            const unsigned gpx = __reduce_add_sync(FULL, mine ? ilen : 0u);
            if (cl == 0) {
                if (lane == src) atomicAdd(&tcnt[0], gpx);
                continue;
            }
            if (mine) gmask = gm;
            const unsigned gsx = __reduce_add_sync(FULL, mine ? ilen * ix + ilen * (ilen - 1u) / 2u : 0u);
            const unsigned gxa = __reduce_min_sync(FULL, mine ? ix : 0xFFFFFFFFu);
            const unsigned gxb = __reduce_max_sync(FULL, mine ? ix + ilen - 1u : 0u);
            const unsigned gkmn = __reduce_min_sync(FULL, mine ? ikmn : 0xFFFFFFFFu);
            const unsigned gkmx = __reduce_max_sync(FULL, mine ? ikmx : 0u);
            // 64-bit sums as two redux operations: value = hi * 2^24 + lo, lo in [0, 2^24)
            const unsigned glo = __reduce_add_sync(FULL, mine ? (unsigned)(isdep & 0xFFFFFFll) : 0u);
            const int ghi = __reduce_add_sync(FULL, mine ? (int)(isdep >> 24) : 0);
            unsigned long long gray = 0;
            if (gxb - gxa + 1u == gpx) {
                // the group's pixels are one horizontal run: one lane takes the difference of two table entries
                if (lane == src) gray = rrow[gxb] - (gxa > 0 ? rrow[gxa - 1] : 0ull);
            } else {
                unsigned long long d = 0;
                if (mine) d = rrow[ix + ilen - 1u] - (ix > 0 ? rrow[ix - 1] : 0ull);
                const unsigned rlo = __reduce_add_sync(FULL, (unsigned)(d & 0xFFFFFFull));
                const unsigned rhi = __reduce_add_sync(FULL, (unsigned)(d >> 24));
                gray = ((unsigned long long)rhi << 24) + rlo;
            }
            if (lane == src) {
                SmemLeaf* t = &tab[cl];
                atomicAdd(&tcnt[cl], gpx);
                atomicAdd(&t->sx, gsx);
                atomicMin(&t->bx0, gxa); atomicMax(&t->bx1, gxb);
                atomicMin(&t->kmin, gkmn); atomicMax(&t->kmax, gkmx);
                atomicAdd(&t->sdep, (unsigned long long)(((long long)ghi << 24) + (long long)glo));
                atomicAdd(&t->sdist, gray);
            }
        }
        return gmask;
    };

    // the next row's pixels are in flight while the current row is reduced
    uint4 nl = make_uint4(0, 0, 0, 0);
    float nd[TS_PX];
#pragma unroll
    for (int k = 0; k < TS_PX; ++k) nd[k] = 0.f;
    auto fetch = [&](int y) {
        const size_t p0 = (size_t)y * W + x0;
        if (VEC) {
            if (npx > 0) {
                nl = *reinterpret_cast<const uint4*>(lp + p0);
                const float4 d0 = *reinterpret_cast<const float4*>(dp + p0), d1 = *reinterpret_cast<const float4*>(dp + p0 + 4);
                nd[0] = d0.x; nd[1] = d0.y; nd[2] = d0.z; nd[3] = d0.w; nd[4] = d1.x; nd[5] = d1.y; nd[6] = d1.z; nd[7] = d1.w;
            }
        } else {
            unsigned w[4] = {0, 0, 0, 0};
#pragma unroll
            for (int k = 0; k < TS_PX; ++k) {
                const bool ok = k < npx;
                const unsigned v = ok ? (unsigned)(unsigned short)lp[p0 + k] : 0u;
                w[k >> 1] |= v << (16 * (k & 1));
                nd[k] = ok ? dp[p0 + k] : 0.f;
            }
            nl = make_uint4(w[0], w[1], w[2], w[3]);
        }
    };
    if (row0 < row1) fetch(row0);
    int buf = 0;
    for (int y = row0; y < row1; ++y, buf ^= 1) {
        tcnt = cur2 + buf * L;
        rrow = c.ray_tab + (size_t)y * W;
        const int16_t* lrow = lp + (size_t)y * W;
        const float* drow = dp + (size_t)y * W;
        const uint4 cl4 = nl;
        float val[TS_PX];
#pragma unroll
        for (int k = 0; k < TS_PX; ++k) val[k] = nd[k];
        if (y + 1 < row1) fetch(y + 1);
        // ---- this thread's 8 pixels, branch free: union byte, labels outside the table, uniformity
        const unsigned cw[4] = {cl4.x, cl4.y, cl4.z, cl4.w};
        unsigned ubyte = 0, badw = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const unsigned pos = __vcmpgts2(cw[q], 0u);                 // 0xFFFF per halfword that is >= 1
            ubyte |= ((pos & 1u) | ((pos >> 15) & 2u)) << (2 * q);
            badw |= __vcmplts2(cw[q], 0u) | __vcmpges2(cw[q], llim);
        }
        if (npx < TS_PX) {                                              // ragged end of the row (only without 128-bit loads)
            ubyte &= (1u << npx) - 1u;
            if (npx == 0) badw = 0;
        }
        const bool uniform = npx == TS_PX && cl4.x == cl4.y && cl4.y == cl4.z && cl4.z == cl4.w && (cl4.x >> 16) == (cl4.x & 0xFFFFu);
        const bool mixed = npx > 0 && !uniform;
        if (badw) s_bad = 1;
        if (npx > 0) ub[(size_t)y * ubw + tid] = (uint8_t)ubyte;
        {
            const unsigned first = ubyte ? (unsigned)((size_t)y * W) + (unsigned)x0 + (unsigned)(__ffs(ubyte) - 1) : 0xFFFFFFFFu;
            const unsigned wm = __reduce_min_sync(FULL, first);
            if (lane == 0 && wm != 0xFFFFFFFFu) atomicMin(&s_first, wm);
        }
        // left / right image border (top / bottom rows are handled per row below)
        if (x0 == 0 && npx > 0) { const int lb_ = (int)(short)(cl4.x & 0xFFFFu); if (lb_ >= 1 && lb_ < L) atomicOr(&tab[lb_].border, 1u); }
        if (npx > 0 && x0 + npx == W) { const int le = (int)lrow[W - 1]; if (le >= 1 && le < L) atomicOr(&tab[le].border, 1u); }
        // ---- uniform threads: one item of 8 pixels each
        const int l0 = (int)(short)(cl4.x & 0xFFFFu);
        const int ul = (uniform && !badw) ? l0 : -1;
        long long sdep = 0;
        unsigned kmn = 0xFFFFFFFFu, kmx = 0;
        if (ul >= 1) {
#pragma unroll
            for (int k = 0; k < TS_PX; ++k) {
                sdep += fixed28(val[k]);
                const unsigned key = f2key(val[k]);
                kmn = min(kmn, key); kmx = max(kmx, key);
            }
        }
        const unsigned gmask = reduce_items(ul, (unsigned)x0, TS_PX, sdep, kmn, kmx);
        // ---- threads with a label boundary (about a dozen per row): their pixels are dealt out one per lane, four
        //      threads at a time, and go through the same reduction as items of one pixel (values re-read through L1)
        const unsigned mixed_lanes = __ballot_sync(FULL, mixed);
        for (unsigned mm = mixed_lanes; mm; ) {
            const unsigned srcl = __fns(mm, 0, (lane >> 3) + 1);         // the (lane / 8 + 1)-th remaining mixed lane
            const int x = srcl != 0xFFFFFFFFu ? ((warp << 5) + (int)srcl) * TS_PX + (lane & 7) : W;
            int il = -1;
            long long isdep = 0;
            unsigned ikey = 0;
            if (x < W) {
                il = (int)lrow[x];
                if (il < 0 || il >= L) il = -1;
                if (il >= 1) { const float v = drow[x]; isdep = fixed28(v); ikey = f2key(v); }
            }
            reduce_items(il, (unsigned)x, 1u, isdep, ikey, ikey);
#pragma unroll
            for (int q = 0; q < 4; ++q) mm &= mm - 1u;                  // four mixed lanes done
        }
        __syncthreads();
        // ---- per row and label (warp 0; lane owns labels [la, lb)): pixel count, sum of y, vertical extent, top / bottom
        //      border; then the row's offsets = exclusive scan of the counts of the labels >= 1.  The other warps clear the
        //      other buffer for the next row meanwhile (everybody left its cursors behind before the barrier above).
        if (warp == 0) {
            const int chunk = (L + 31) >> 5;
            const int la = min(lane * chunk, L), lb = min(la + chunk, L);
            const bool edge_row = y == 0 || y == H - 1;
            unsigned mine = 0;
            for (int l = la; l < lb; ++l) {
                const unsigned n = tcnt[l];
                if (n) {
                    SmemLeaf* t = &tab[l];
                    t->cnt += n;
                    if (l >= 1) {
                        t->sy += (unsigned)y * n;
                        t->by0 = min(t->by0, (unsigned)y); t->by1 = max(t->by1, (unsigned)y);
                        if (edge_row) atomicOr(&t->border, 1u);
                        mine += n;
                    }
                }
            }
            unsigned incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned t = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += t;
            }
            unsigned run = incl - mine;
            uint16_t* toff = toff_base + (size_t)y * c.lstride;
            for (int l = la; l < lb; ++l) {
                const unsigned n = l >= 1 ? tcnt[l] : 0u;
                toff[l] = (uint16_t)run;
                tcnt[l] = run;
                run += n;
            }
            if (lane == 31) toff[L] = (uint16_t)incl;
        }
        if (warp != 0 || NT == 32) {
            unsigned* other = cur2 + (buf ^ 1) * L;
            const int first_t = NT == 32 ? 0 : 32, n_t = NT == 32 ? 32 : NT - 32;
            for (int l = tid - first_t; l < L; l += n_t) other[l] = 0;
        }
        __syncthreads();
        // ---- placement: one cursor bump per label group, the values go out in lane order
        {
            float* segr = seg + (size_t)y * W;
            unsigned base = 0;
            const int leader = gmask ? __ffs(gmask) - 1 : lane;
            if (gmask && lane == leader) base = atomicAdd(&tcnt[ul], (unsigned)(TS_PX * __popc(gmask)));
            base = __shfl_sync(FULL, base, leader);
            if (gmask) {
                float* o = segr + base + TS_PX * __popc(gmask & lanes_lt);
#pragma unroll
                for (int k = 0; k < TS_PX; ++k) o[k] = val[k];
            }
            for (unsigned mm = mixed_lanes; mm; ) {
                const unsigned srcl = __fns(mm, 0, (lane >> 3) + 1);
                const int x = srcl != 0xFFFFFFFFu ? ((warp << 5) + (int)srcl) * TS_PX + (lane & 7) : W;
                int il = -1;
                if (x < W) { il = (int)lrow[x]; if (il >= L) il = -1; }
                unsigned todo = __ballot_sync(FULL, il >= 1);
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    const int cl = __shfl_sync(FULL, il, src);
                    const bool mine = il == cl;
                    const unsigned gm = __ballot_sync(FULL, mine);
                    todo &= ~gm;
                    unsigned pbase = 0;
                    if (lane == src) pbase = atomicAdd(&tcnt[cl], (unsigned)__popc(gm));
                    pbase = __shfl_sync(FULL, pbase, src);
                    if (mine) segr[pbase + __popc(gm & lanes_lt)] = drow[x];
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) mm &= mm - 1u;
            }
        }
    }
    __syncthreads();
    for (int l = tid; l < L; l += NT) {
        const SmemLeaf t = tab[l];
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
                atomicMin(&c.kmin[o], t.kmin); atomicMax(&c.kmax[o], t.kmax);
                if (t.border) atomicOr(&c.border[o], 1u);
            }
        }
    }
    if (tid == 0) {
        if (s_first != 0xFFFFFFFFu) atomicMin(&c.first_leaf[b], s_first);
        if (s_bad) atomicOr(&c.status[b], LG_ST_LABEL_RANGE);
    }
}

// background id = smallest id present (torch.unique(mask)[1:], leaf_scorer.py:32)
__device__ __forceinline__ int background_id(const uint32_t* cnt, int L) {
    for (int l = 0; l < L; ++l)
        if (cnt[l]) return l;
    return -1;
}

template <int NT>
__device__ __forceinline__ void block_sum3(unsigned& a, unsigned& b, unsigned& c, unsigned* sm) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        a += __shfl_xor_sync(0xFFFFFFFFu, a, d);
        b += __shfl_xor_sync(0xFFFFFFFFu, b, d);
        c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
    }
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) { sm[w * 3] = a; sm[w * 3 + 1] = b; sm[w * 3 + 2] = c; }
    __syncthreads();
    a = 0; b = 0; c = 0;
#pragma unroll
    for (int k = 0; k < NT / 32; ++k) { a += sm[k * 3]; b += sm[k * 3 + 1]; c += sm[k * 3 + 2]; }
}

constexpr int MED_NT = 256;
constexpr int MED_CAP = 4096;    // candidate keys kept in shared memory once the search range is this small
// np.median(depth[labels == l]) for every label of every frame (leaf_scorer.py:41-47): radix selection on the label's
// values, which leaf_rows_kernel left grouped by label inside every image row.  A pass over the values walks the rows of
// the label's bounding box, one warp per row: two entries of the row's offset table give the sub-block.  Keys are
// ranked relative to the label's smallest key; a label that does not fit shared memory gets one streaming pass with a
// 256-bin histogram of the top 7-8 bits of its key range, a second pass compacts the <= 4096 candidates of the median's
// bin into shared memory, and two-bit radix rounds finish there.
__global__ void __launch_bounds__(MED_NT) leaf_median_kernel(lg_context c) {
    const int l = blockIdx.x, b = blockIdx.y, L = c.L;
    const uint32_t* cnt = c.cnt + (size_t)b * L;
    __shared__ unsigned sm[MED_NT / 32 * 3];
    extern __shared__ unsigned s_keys[];     // [MED_CAP]
    __shared__ unsigned s_n;
    __shared__ unsigned s_hist[256], s_sel[3];
    __shared__ int s_bg;
    if (threadIdx.x == 0) { s_bg = background_id(cnt, L); s_n = 0; }
    __syncthreads();
    const unsigned n = cnt[l];
    if (n == 0 || l == s_bg || l == 0) {
        if (threadIdx.x == 0) c.median[(size_t)b * L + l] = CUDART_NAN_F;
        return;
    }
    const size_t o = (size_t)b * L + l;
    const unsigned kmin = c.kmin[o], kmax = c.kmax[o];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // rows that hold pixels of this label
    const unsigned W = (unsigned)c.W;
    const int t_first = (int)c.by0[o], t_last = (int)c.by1[o];
    const float* seg = c.seg + (size_t)b * c.P;
    const uint16_t* toff = c.tile_off + (size_t)b * c.H * c.lstride + l;
    // f(key, valid) for every value of the label; whole warps call it together (valid = false on the padding lanes)
    // (warp w takes rows t_first + w, + 8, ...; its lanes fetch the offsets of 32 of those rows at once)
    auto for_each_key = [&](auto f) {
        constexpr int NW = MED_NT / 32;
        for (int base = t_first + warp; base <= t_last; base += NW * 32) {
            const int tr = base + NW * lane;
            unsigned a = 0, z = 0;
            if (tr <= t_last) { const uint16_t* e = toff + (size_t)tr * c.lstride; a = e[0]; z = e[1]; }
            unsigned live = __ballot_sync(FULL, z > a);
            while (live) {
                const int s = __ffs(live) - 1;
                live &= live - 1;
                const unsigned aa = __shfl_sync(FULL, a, s), zz = __shfl_sync(FULL, z, s);
                const float* v = seg + (size_t)(base + NW * s) * W;
                for (unsigned i0 = aa; i0 < zz; i0 += 32) {
                    const unsigned i = i0 + lane;
                    const bool ok = i < zz;
                    f(ok ? f2key(v[i]) - kmin : 0u, ok);
                }
            }
        }
    };
    const unsigned k_lo = (n & 1) ? n / 2 : n / 2 - 1;   // rank of the lower middle
    unsigned klo;            // key of rank k_lo
    unsigned below = 0;      // elements whose key is smaller than every current candidate
    unsigned m = n;          // current candidates
    bool in_smem = false;
    unsigned set_below = 0, set_m = 0;   // the compacted set: its size and the number of elements below it
    auto compact = [&](unsigned prefix, unsigned pmask) {   // candidates (key & pmask) == prefix -> s_keys
        for_each_key([&](unsigned key, bool ok) {
            const bool hit = ok && (key & pmask) == prefix;
            const unsigned ball = __ballot_sync(FULL, hit);
            if (ball) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(&s_n, __popc(ball));
                base = __shfl_sync(FULL, base, 0);
                if (hit) s_keys[base + __popc(ball & ((1u << lane) - 1u))] = key;
            }
        });
        __syncthreads();
    };
    if (kmin == kmax) {
        klo = 0;
        below = 0; m = n;
    } else {
        const int top = 31 - __clz(kmax - kmin);          // highest set bit of the largest relative key
        int shift = top & ~1;                              // the digit (shift+1, shift) contains it
        unsigned pmask = shift >= 30 ? 0u : ~((4u << shift) - 1u);
        unsigned prefix = 0u;
        if (n <= MED_CAP) { compact(prefix, pmask); in_smem = true; set_below = 0; set_m = n; }
        else {
            int s0 = max(top - 7, 0);
            s0 += s0 & 1;                                      // even, so that the two-bit rounds end at bit 0
            for (int i = threadIdx.x; i < 256; i += MED_NT) s_hist[i] = 0;
            __syncthreads();
            for_each_key([&](unsigned key, bool ok) { if (ok) atomicAdd(&s_hist[key >> s0], 1u); });
            __syncthreads();
            if (threadIdx.x < 32) {                            // the bin that holds rank k_lo
                unsigned mine = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) mine += s_hist[8 * lane + j];
                unsigned incl = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= d) incl += t;
                }
                const unsigned excl = incl - mine;
                if (k_lo >= excl && k_lo < incl) {
                    unsigned acc = excl;
                    for (int j = 0; j < 8; ++j) {
                        const unsigned h = s_hist[8 * lane + j];
                        if (k_lo < acc + h) { s_sel[0] = 8u * lane + j; s_sel[1] = acc; s_sel[2] = h; break; }
                        acc += h;
                    }
                }
            }
            __syncthreads();
            prefix = s_sel[0] << s0;
            below = s_sel[1];
            m = s_sel[2];
            pmask = s0 == 0 ? 0xFFFFFFFFu : ~((1u << s0) - 1u);
            shift = s0 - 2;
            if (m <= MED_CAP && s0 > 0) {
                compact(prefix, pmask);
                in_smem = true; set_below = below; set_m = m;
            }
        }
        for (; shift >= 0; shift -= 2) {
            unsigned c0 = 0, c1 = 0, c2 = 0;
            if (in_smem) {
                for (unsigned i = threadIdx.x; i < set_m; i += MED_NT) {
                    const unsigned key = s_keys[i];
                    if ((key & pmask) == prefix) {
                        const unsigned d = (key >> shift) & 3u;
                        c0 += (d == 0); c1 += (d == 1); c2 += (d == 2);
                    }
                }
            } else {
                for_each_key([&](unsigned key, bool ok) {
                    if (ok && (key & pmask) == prefix) {
                        const unsigned d = (key >> shift) & 3u;
                        c0 += (d == 0); c1 += (d == 1); c2 += (d == 2);
                    }
                });
            }
            block_sum3<MED_NT>(c0, c1, c2, sm);
            const unsigned kk = k_lo - below;
            unsigned d;
            if (kk < c0) { d = 0; m = c0; }
            else if (kk < c0 + c1) { d = 1; below += c0; m = c1; }
            else if (kk < c0 + c1 + c2) { d = 2; below += c0 + c1; m = c2; }
            else { d = 3; below += c0 + c1 + c2; m = m - (c0 + c1 + c2); }
            prefix |= d << shift;
            pmask |= 3u << shift;
            if (!in_smem && m <= MED_CAP && shift > 0) {
                compact(prefix, pmask);
                in_smem = true; set_below = below; set_m = m;
            }
        }
        klo = prefix;
    }
    float med = key2f(klo + kmin);
    if (!(n & 1)) {
        // upper middle (rank k_lo + 1): klo again when enough elements are <= klo, else the smallest larger key
        float hi;
        if (below + m > k_lo + 1) {
            hi = med;
        } else {
            unsigned mn = 0xFFFFFFFFu;
            if (in_smem && k_lo + 1 < set_below + set_m) {
                for (unsigned i = threadIdx.x; i < set_m; i += MED_NT) { const unsigned key = s_keys[i]; if (key > klo) mn = min(mn, key); }
            } else {
                for_each_key([&](unsigned key, bool ok) { if (ok && key > klo) mn = min(mn, key); });
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, d));
            __syncthreads();
            if (lane == 0) sm[threadIdx.x >> 5] = mn;
            __syncthreads();
            mn = 0xFFFFFFFFu;
            for (int w = 0; w < MED_NT / 32; ++w) mn = min(mn, sm[w]);
            hi = key2f(mn + kmin);
        }
        med = __fmul_rn(__fadd_rn(med, hi), 0.5f);   // float32 mean of the two middles
    }
    if (threadIdx.x == 0) c.median[(size_t)b * L + l] = med;
}

// ---------------------------------------------------------------------------------------------------
// exact squared Euclidean distance transform
// ---------------------------------------------------------------------------------------------------
// Column pass without a distance image.  For every column the kernel keeps the source pixels as vertical bit words
// (bit r of word yw = row 32 yw + r is a source) plus, per word, the distance from its first row to the nearest source
// strictly above (vup) and from its last row to the nearest source strictly below (vdn).  The column distance g(x, y) of
// any pixel is then three coalesced loads and a count-leading / find-first on the word (g_of), so the row pass computes it
// where it needs it instead of reading a [H][W] image the column pass would have to write (2 B/px, and the row pass reads
// about 1/32 of it).  What the arg-max search reads everywhere are the minima of g over 32-column chunks: per row (gmin)
// and per block of 8 rows (g8).
struct SrcUnionBits {       // source = leaf pixel of the union mask leaf_rows_kernel wrote (rows of ceil(W / 8) bytes)
    const uint8_t* ub;
    size_t stride;
    int ubw;
    __device__ __forceinline__ bool at(int b, int x, int y) const {
        return (ub[(size_t)b * stride + (size_t)y * ubw + (x >> 3)] >> (x & 7)) & 1u;
    }
};
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
__global__ void __launch_bounds__(VC_NT) edt_vcol_kernel(lg_context c, SRC src, int want_min) {
    const int W = c.W, H = c.H, Hw = c.Hw, b = blockIdx.y;
    const int x = blockIdx.x * VC_NT + threadIdx.x, lane = threadIdx.x & 31;
    const bool in = x < W;
    const int xc = in ? x : W - 1;             // out-of-range lanes shadow the last column and store nothing
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
            for (int r = 0; r < 32; ++r) w |= (src.at(b, xc, ybase + r) ? 1u : 0u) << r;
        } else {
            for (int r = 0; ybase + r < H; ++r) w |= (src.at(b, xc, ybase + r) ? 1u : 0u) << r;
        }
        if (in) {
            vb[(size_t)yw * W + x] = w;
            vu[(size_t)yw * W + x] = (uint16_t)min(since + 1u, 0xFFFFu);
        }
        since = w ? (unsigned)__clz(w) : min(since + 32u, 0xFFFFu);
    }
    // upward: the distance to the nearest source below each word, and the chunk minima of g
    const int chunk = x >> 5, nchunks = c.edt_nchunks;
    const bool wr = want_min && lane == 0 && chunk < nchunks;
    uint16_t* gm = c.edt_gmin + (size_t)b * H * nchunks + chunk;
    uint16_t* g8 = c.edt_g8 + (size_t)b * c.H8 * nchunks + chunk;
    unsigned below = 0xFFFFu;                  // distance from the current word's last row to the nearest source strictly below
    unsigned m8 = 0xFFFFu;
    for (int yw = Hw - 1; yw >= 0; --yw) {
        const size_t o = (size_t)yw * W + xc;
        const unsigned w = vb[o];              // written by this thread (or, for a shadow lane, by the last column's thread of
        const unsigned up0 = vu[o];            // this very warp: same value in either order)
        if (in) vd[(size_t)yw * W + x] = (uint16_t)below;
        if (want_min) {
            const int ybase = yw << 5;
#pragma unroll 4
            for (int r = 31; r >= 0; --r) {
                const int y = ybase + r;
                if (y >= H) continue;
                unsigned g = 0u;
                if (!((w >> r) & 1u)) {
                    const unsigned mu = w << (31 - r), md = w >> r;
                    const unsigned du = mu ? (unsigned)__clz(mu) : (unsigned)r + up0;
                    const unsigned dd = md ? (unsigned)(__ffs(md) - 1) : (unsigned)(31 - r) + below;
                    g = min(min(du, dd), 0xFFFFu);
                }
                const unsigned m = __reduce_min_sync(FULL, in ? g : 0xFFFFu);
                m8 = min(m8, m);
                if (wr) {
                    gm[(size_t)y * nchunks] = (uint16_t)m;
                    if ((y & 7) == 0) g8[(size_t)(y >> 3) * nchunks] = (uint16_t)m8;
                }
                if ((y & 7) == 0) m8 = 0xFFFFu;
            }
        }
        below = w ? (unsigned)__ffs(w) : min(below + 32u, 0xFFFFu);
    }
}

// exact search for pixel x of a row whose squared column distances are in srow (shared memory): d2(x) = min over x' of
// (x - x')^2 + g(x')^2, abandoned as soon as the running value drops below lb (such a pixel cannot be the maximum)
__device__ __forceinline__ unsigned edt_row_search(const unsigned* srow, int W, int x, unsigned lb) {
    unsigned bestd = srow[x];
    if (bestd < lb) return bestd;
    if (lb) {   // probes at doubling offsets: almost every pixel near a source drops below the bound here
        for (unsigned k = 1; k * k < bestd; k <<= 1) {
            const int xl = x - (int)k, xr = x + (int)k;
            if (xl < 0 && xr >= W) break;
            const unsigned kk = k * k;
            if (xl >= 0) { const unsigned s = srow[xl]; if (s != 0xFFFFFFFFu) bestd = min(bestd, s + kk); }
            if (xr < W) { const unsigned s = srow[xr]; if (s != 0xFFFFFFFFu) bestd = min(bestd, s + kk); }
        }
        if (bestd < lb) return bestd;
    }
    for (unsigned k = 1; k * k < bestd; ++k) {
        const int xl = x - (int)k, xr = x + (int)k;
        if (xl < 0 && xr >= W) break;
        const unsigned kk = k * k;
        if (xl >= 0) { const unsigned s = srow[xl]; if (s != 0xFFFFFFFFu) bestd = min(bestd, s + kk); }
        if (xr < W) { const unsigned s = srow[xr]; if (s != 0xFFFFFFFFu) bestd = min(bestd, s + kk); }
        if (bestd < lb) break;
    }
    return bestd;
}

constexpr int EDT_NT = 256;
// Row pass, full field (d2out != nullptr: lg_edt_squared) or arg-max with pruning against the frame's running maximum
// (small images, where the block search below has nothing to prune with).  One CTA per row at a time; rows are visited in a
// permuted order so that the bound comes from all over the frame early.
__global__ void __launch_bounds__(EDT_NT) edt_row_kernel(lg_context c, uint32_t* __restrict__ d2out,
                                                          unsigned long long* __restrict__ best, int row_stride) {
    extern __shared__ unsigned srow[];   // g squared, 0xFFFFFFFF = no source in that column
    __shared__ unsigned long long sbest[EDT_NT / 32];
    __shared__ unsigned s_lb;
    const int b = blockIdx.x, W = c.W, H = c.H;
    const size_t P = c.P;
    const EdtCols cols = edt_cols_of(c, b);
    const bool prune = (d2out == nullptr) && (best != nullptr);
    unsigned long long mybest = 0;       // CTA-wide best so far (identical in every thread)
    for (int yi = blockIdx.y; yi < H; yi += gridDim.y) {
        const int y = (int)(((long long)yi * row_stride) % H);
        int any = 0;
        for (int x = threadIdx.x; x < W; x += EDT_NT) {
            const unsigned v = cols.g_of(x, y);
            srow[x] = (v == 0xFFFFu) ? 0xFFFFFFFFu : v * v;
            any |= (v != 0xFFFFu);
        }
        if (threadIdx.x == 0) {
            unsigned v = 0;
            if (prune) {
                const unsigned long long gb = *reinterpret_cast<volatile unsigned long long*>(&best[b]);
                v = (unsigned)(max(gb, mybest) >> 32);
            }
            s_lb = v;
        }
        any = __syncthreads_or(any);     // srow and s_lb complete; does the row see any source at all?
        const unsigned lb = s_lb;
        unsigned long long rowbest = 0;
        if (any) {
            for (int x = threadIdx.x; x < W; x += EDT_NT) {
                const unsigned bestd = edt_row_search(srow, W, x, lb);
                if (bestd < lb) continue;
                const size_t idx = (size_t)y * W + x;
                if (d2out) d2out[(size_t)b * P + idx] = bestd;
                rowbest = max(rowbest, ((unsigned long long)bestd << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)idx));
            }
        } else if (d2out) {
            for (int x = threadIdx.x; x < W; x += EDT_NT) d2out[(size_t)b * P + (size_t)y * W + x] = 0xFFFFFFFFu;
        }
        // the reduction only runs when some thread improved on the CTA's best; the barrier also protects srow
        const int improved = __syncthreads_or(best != nullptr && rowbest > mybest);
        if (improved) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) rowbest = max(rowbest, __shfl_xor_sync(FULL, rowbest, d));
            if ((threadIdx.x & 31) == 0) sbest[threadIdx.x >> 5] = rowbest;
            __syncthreads();
            for (int w = 0; w < EDT_NT / 32; ++w) mybest = max(mybest, sbest[w]);
            if (threadIdx.x == 0 && mybest > *reinterpret_cast<volatile unsigned long long*>(&best[b])) atomicMax(&best[b], mybest);
            __syncthreads();             // sbest is reused by the next improving row
        }
    }
}

// Arg-max search.  The first maximum of the field is found by branch and bound on upper bounds that come from the chunk
// minima alone: every pixel of chunk j of row y is at most 31 + 32 dj columns away from the column of chunk j +- dj that
// holds gmin, so d2 <= gmin(j')^2 + (32 |j - j'| + 31)^2 for every j'; for a block of 8 rows the same holds with
// (g8(j') + 7) in place of gmin (the nearest source of the block's best column is at most 7 rows further from any other
// row of the block).  A chunk whose bound is below the frame's running maximum (best[b], raised with atomicMax by every
// warp that finds a larger exact distance) can neither hold the maximum nor tie with it, so the result is exactly the
// first maximum of the full field whatever the schedule.
//
// mrow: chunk minima (linear, 0xFFFF = none) in shared memory; add: 0 for a row, 7 for a block of rows.  Returns in
// alive[k] bit `lane`: chunk 32 k + lane may still hold a pixel at squared distance >= lb.
__device__ __forceinline__ bool edt_alive_chunks(const unsigned* mrow, int nchunks, unsigned lb, unsigned add, int lane,
                                                 unsigned (&alive)[4]) {
    bool any_alive = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int j = 32 * k + lane;
        bool keep = false;
        if (j < nchunks) {
            auto sq = [&](int jj) -> unsigned long long {
                const unsigned v = mrow[jj];
                if (v == 0xFFFFu) return ~0ull >> 1;
                const unsigned long long t = (unsigned long long)v + add;
                return t * t;
            };
            unsigned long long ub = sq(j) + 31ull * 31ull;
            for (int dj = 1; dj < nchunks && ub >= lb; ++dj) {
                const unsigned long long off = (unsigned long long)(32 * dj + 31) * (32 * dj + 31);
                if (off >= lb) break;
                if (j - dj >= 0) ub = min(ub, sq(j - dj) + off);
                if (j + dj < nchunks) ub = min(ub, sq(j + dj) + off);
            }
            keep = ub >= lb;
        }
        alive[k] = __ballot_sync(FULL, keep);
        any_alive |= alive[k] != 0u;
    }
    return any_alive;
}

// A lower bound to start from.  On the grid of cells (8 rows x 32 columns; a cell is occupied when it holds a source,
// i.e. its g8 is 0) the kernel runs a small two-pass distance transform in shared memory - vertical distance to the
// nearest occupied cell per cell column, then the lower envelope along the cell rows in pixel units - and every warp
// evaluates the EXACT distance at the middle pixel of the best cell of its share.  The maximum of the true field is
// within about a cell of the coarse one, so the search below starts with a bound that prunes nearly everything.
// coarse == 0 (cell grid too large for shared memory): the cells are ranked by g8 alone.  One CTA per frame.
constexpr int SEED_NT = 256;
__global__ void __launch_bounds__(SEED_NT) edt_seed_kernel(lg_context c, int coarse) {
    extern __shared__ uint16_t s_gc[];                  // [H8][nchunks] vertical distance in cells, 0xFFFF = column has no source
    const int b = blockIdx.x, W = c.W, H = c.H, nchunks = c.edt_nchunks, H8 = c.H8;
    const int tid = threadIdx.x, lane = tid & 31;
    const EdtCols cols = edt_cols_of(c, b);
    const uint16_t* g8 = c.edt_g8 + (size_t)b * H8 * nchunks;
    const int cells = H8 * nchunks;
    unsigned long long bestv = 0;
    int bestc = -1;
    if (coarse) {
        for (int j = tid; j < nchunks; j += SEED_NT) {
            unsigned d = 0xFFFFu;
            for (int i = 0; i < H8; ++i) {
                d = g8[(size_t)i * nchunks + j] == 0 ? 0u : min(d + 1u, 0xFFFFu);
                s_gc[i * nchunks + j] = (uint16_t)d;
            }
            d = 0xFFFFu;
            for (int i = H8 - 1; i >= 0; --i) {
                d = min((unsigned)s_gc[i * nchunks + j], min(d + 1u, 0xFFFFu));
                s_gc[i * nchunks + j] = (uint16_t)d;
            }
        }
        __syncthreads();
        for (int cell = tid; cell < cells; cell += SEED_NT) {
            const int i = cell / nchunks, j = cell - i * nchunks;
            const uint16_t* row = s_gc + i * nchunks;
            if (row[j] == 0) continue;                  // occupied: holds a source
            unsigned long long d2 = ~0ull;
            for (int jj = 0; jj < nchunks; ++jj) {
                const unsigned v = row[jj];
                if (v == 0xFFFFu) continue;
                const long long dy = 8ll * v, dx = 32ll * (jj - j);
                d2 = min(d2, (unsigned long long)(dy * dy + dx * dx));
            }
            if (d2 != ~0ull && d2 > bestv) { bestv = d2; bestc = cell; }
        }
    } else {
        for (int cell = tid; cell < cells; cell += SEED_NT) {
            const unsigned v = g8[cell];
            if (v != 0xFFFFu && v > bestv) { bestv = v; bestc = cell; }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long ov = __shfl_xor_sync(FULL, bestv, d);
        const int oc = __shfl_xor_sync(FULL, bestc, d);
        if (ov > bestv || (ov == bestv && oc >= 0 && (bestc < 0 || oc < bestc))) { bestv = ov; bestc = oc; }
    }
    if (bestc < 0) return;                                  // warp-uniform
    const int y = min((bestc / nchunks) * 8 + 4, H - 1), x = min((bestc % nchunks) * 32 + 16, W - 1);
    // exact d2 at (x, y): min over x' of g(x', y)^2 + (x - x')^2
    unsigned long long d2 = ~0ull;
    for (int xx = lane; xx < W; xx += 32) {
        const unsigned g = cols.g_of(xx, y);
        if (g == 0xFFFFu) continue;
        const long long dx = xx - x;
        d2 = min(d2, (unsigned long long)g * g + (unsigned long long)(dx * dx));
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) d2 = min(d2, __shfl_xor_sync(FULL, d2, d));
    if (lane == 0 && d2 != ~0ull && d2 > 0 && d2 <= 0xFFFFFFFFull) {
        const unsigned idx = (unsigned)((size_t)y * W + x);
        atomicMax(&c.edt_best[b], (d2 << 32) | (unsigned long long)(0xFFFFFFFFu - idx));
    }
}

// Main part: ONE WARP PER BLOCK OF 8 ROWS.  The block is judged by its g8 row, each of its rows by its gmin row, and only
// rows with a surviving chunk compute their column distances (into the warp's shared-memory row) and run the exact
// search, for the surviving chunks only.
constexpr int EDTW_NT = 256;
__global__ void __launch_bounds__(EDTW_NT) edt_blockmax_kernel(lg_context c, int blk_stride) {
    extern __shared__ unsigned sm_rows[];            // per warp: W squared column distances, then 128 chunk minima
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = c.W, H = c.H, nchunks = c.edt_nchunks, H8 = c.H8;
    unsigned* srow = sm_rows + (size_t)warp * (W + 128);
    unsigned* mrow = srow + W;
    const EdtCols cols = edt_cols_of(c, b);
    const uint16_t* gmf = c.edt_gmin + (size_t)b * H * nchunks;
    const uint16_t* g8f = c.edt_g8 + (size_t)b * H8 * nchunks;
    unsigned long long* best = c.edt_best;
    unsigned long long mybest = 0;                   // this warp's best (identical in all lanes)
    const int per_pass = gridDim.y * (EDTW_NT / 32);
    auto bound = [&]() -> unsigned {
        unsigned long long gb = 0;
        if (lane == 0) gb = *reinterpret_cast<volatile unsigned long long*>(&best[b]);
        gb = __shfl_sync(FULL, gb, 0);
        return (unsigned)(max(gb, mybest) >> 32);
    };
    for (int bi = blockIdx.y * (EDTW_NT / 32) + warp; bi < H8; bi += per_pass) {
        const int blk = (int)(((long long)bi * blk_stride) % H8);
        unsigned lb = bound();
        unsigned alive[4];
        if (lb) {                                    // with no bound yet nothing can be dropped
            __syncwarp();
            bool has_source = false;
            for (int j = lane; j < nchunks; j += 32) { const unsigned v = g8f[(size_t)blk * nchunks + j]; mrow[j] = v; has_source |= v != 0xFFFFu; }
            if (!__any_sync(FULL, has_source)) continue;
            __syncwarp();
            if (!edt_alive_chunks(mrow, nchunks, lb, 7u, lane, alive)) continue;
        }
        const int y_end = min(blk * 8 + 8, H);
        for (int y = blk * 8; y < y_end; ++y) {
            lb = bound();
            __syncwarp();
            bool has_source = false;
            for (int j = lane; j < nchunks; j += 32) { const unsigned v = gmf[(size_t)y * nchunks + j]; mrow[j] = v; has_source |= v != 0xFFFFu; }
            if (!__any_sync(FULL, has_source)) continue;      // no source in any column of this row: nothing to rank
            __syncwarp();
            if (!edt_alive_chunks(mrow, nchunks, lb, 0u, lane, alive)) continue;
            for (int x = lane; x < W; x += 32) {
                const unsigned v = cols.g_of(x, y);
                srow[x] = (v == 0xFFFFu) ? 0xFFFFFFFFu : v * v;
            }
            __syncwarp();
            unsigned long long rowbest = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                unsigned rem = alive[k];
                while (rem) {
                    const int j = 32 * k + (__ffs(rem) - 1);
                    rem &= rem - 1;
                    const int x = (j << 5) + lane;
                    if (x >= W) continue;
                    const unsigned bestd = edt_row_search(srow, W, x, lb);
                    if (bestd < lb) continue;
                    const unsigned idx = (unsigned)((size_t)y * W + x);
                    rowbest = max(rowbest, ((unsigned long long)bestd << 32) | (unsigned long long)(0xFFFFFFFFu - idx));
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) rowbest = max(rowbest, __shfl_xor_sync(FULL, rowbest, d));
            if (rowbest > mybest) {
                mybest = rowbest;
                if (lane == 0 && mybest > *reinterpret_cast<volatile unsigned long long*>(&best[b])) atomicMax(&best[b], mybest);
            }
        }
    }
}

// items are visited in the order (i * stride) mod n: stride ~ 0.618 n, coprime with n
static int golden_stride(int n) {
    auto gcd = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
    int s = (int)(n * 0.6180339887);
    if (s < 1) s = 1;
    while (gcd(s, n) != 1) ++s;
    return s % n ? s % n : 1;
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

// column pass + arg-max search of the union distance transform for n frames, on `st`
template <class SRC>
static int run_edt_argmax(lg_context* c, SRC src, int n, cudaStream_t st, bool mark) {
    const bool small = c->H < 64;      // too few 8-row blocks to prune with: plain pruned row search
    edt_vcol_kernel<SRC><<<dim3((c->W + VC_NT - 1) / VC_NT, n), VC_NT, 0, st>>>(*c, src, small ? 0 : 1);
    LG_LAUNCH_CHECK();
    if (mark) lg_mark(c, LG_M_EDT_COL, st);
    if (small) {
        int per_frame = (148 * 8 * 2 + n - 1) / n;
        per_frame = per_frame < 1 ? 1 : (per_frame > c->H ? c->H : per_frame);
        edt_row_kernel<<<dim3(n, per_frame), EDT_NT, c->W * sizeof(unsigned), st>>>(*c, nullptr, c->edt_best, golden_stride(c->H));
        LG_LAUNCH_CHECK();
        return LG_OK;
    }
    {
        const size_t sm_seed = (size_t)c->H8 * c->edt_nchunks * sizeof(uint16_t);
        const int coarse = sm_seed <= 200 * 1024;
        if (coarse) LG_ENSURE_SMEM(edt_seed_kernel, sm_seed);
        edt_seed_kernel<<<n, SEED_NT, coarse ? sm_seed : 0, st>>>(*c, coarse);
        LG_LAUNCH_CHECK();
    }
    const size_t smw = (size_t)(EDTW_NT / 32) * (c->W + 128) * sizeof(unsigned);
    LG_ENSURE_SMEM(edt_blockmax_kernel, smw);
    int per_frame = (148 * 4 * 2 + n - 1) / n;       // about two waves of CTAs over the batch
    const int max_ctas = (c->H8 + EDTW_NT / 32 - 1) / (EDTW_NT / 32);
    per_frame = per_frame < 1 ? 1 : (per_frame > max_ctas ? max_ctas : per_frame);
    edt_blockmax_kernel<<<dim3(n, per_frame), EDTW_NT, smw, st>>>(*c, golden_stride(c->H8));
    LG_LAUNCH_CHECK();
    return LG_OK;
}

int lg_run_stage1(lg_context* c, const int16_t* labels, const float* depth, int n, lg_camera cam, cudaStream_t st) {
    clear_tables_kernel<<<64, 256, 0, st>>>(*c, n);
    LG_LAUNCH_CHECK();
    if (!c->ray_valid || c->ray_cam.f != cam.f || c->ray_cam.cx != cam.cx || c->ray_cam.cy != cam.cy) {
        ray_table_kernel<<<(c->H + 7) / 8, 256, 0, st>>>(c->ray_tab, c->H, c->W, cam);
        LG_LAUNCH_CHECK();
        c->ray_cam = cam;
        c->ray_valid = 1;
    }
    // the one pass over labels + depth: per-leaf statistics, grouped depth values, union bit mask
    {
        const size_t smem = c->L * (sizeof(SmemLeaf) + 2 * sizeof(unsigned));
        long long rpc = ((long long)c->H * n + 8191) / 8192;                // several thousand CTAs over the batch
        rpc = rpc < 1 ? 1 : (rpc > 16 ? 16 : rpc);
        const dim3 grid((unsigned)((c->H + rpc - 1) / rpc), n);
        const int nt = (((c->W + 7) / 8 + 31) / 32) * 32;                    // one thread per 8 pixels of a row
        const bool vec = (c->W % 8 == 0) && ((reinterpret_cast<uintptr_t>(labels) | reinterpret_cast<uintptr_t>(depth)) % 16 == 0);
        if (vec) {
            LG_ENSURE_SMEM(leaf_rows_kernel<true>, smem);
            leaf_rows_kernel<true><<<grid, nt, smem, st>>>(*c, labels, depth, (int)rpc);
        } else {
            LG_ENSURE_SMEM(leaf_rows_kernel<false>, smem);
            leaf_rows_kernel<false><<<grid, nt, smem, st>>>(*c, labels, depth, (int)rpc);
        }
        LG_LAUNCH_CHECK();
    }
    lg_mark(c, LG_M_STATS, st);
    // the distance transform of the union (arg-max only) is independent of the medians: it runs beside them
    cudaStream_t aux = lg_fork(c, 0, st);
    int rc = run_edt_argmax(c, SrcUnionBits{c->ubits, c->ub_stride, (c->W + 7) / 8}, n, aux, true);
    lg_mark(c, LG_M_EDT_ROW, aux);
    if (!rc) {
        rc = lg_ensure_smem_impl((const void*)leaf_median_kernel, MED_CAP * sizeof(unsigned));
        if (!rc) {
            leaf_median_kernel<<<dim3(c->L, n), MED_NT, MED_CAP * sizeof(unsigned), st>>>(*c);
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
    const SrcMaskZero src{mask, c->P, c->W};
    if (d2) {       // the full field: every row, no pruning
        edt_vcol_kernel<SrcMaskZero><<<dim3((c->W + VC_NT - 1) / VC_NT, n), VC_NT, 0, st>>>(*c, src, 0);
        LG_LAUNCH_CHECK();
        int per_frame = (148 * 8 * 2 + n - 1) / n;
        per_frame = per_frame < 1 ? 1 : (per_frame > c->H ? c->H : per_frame);
        edt_row_kernel<<<dim3(n, per_frame), EDT_NT, c->W * sizeof(unsigned), st>>>(*c, d2, c->edt_best, golden_stride(c->H));
        LG_LAUNCH_CHECK();
    } else {
        int rc = run_edt_argmax(c, src, n, st, false);
        if (rc) return rc;
    }
    if (argmax) {
        edt_argmax_out_kernel<<<(n + 63) / 64, 64, 0, st>>>(c->edt_best, argmax, n);
        LG_LAUNCH_CHECK();
    }
    return LG_OK;
}
