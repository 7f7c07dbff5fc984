// Stage 1 - optimal-leaf selection (reference scripts/utils/leaf_scorer.py:25-203, 277-306).
//
// Kernels (all batched over frames, no host synchronisation):
//   leaf_stats_kernel    one walk down every column of labels + depth: per-label pixel count, coordinate sums, depth
//                        sum, sum of ray lengths, bounding box, border contact, depth key range, first leaf pixel of
//                        the frame; the same walk is the column pass of the union distance transform
//   leaf_offsets_kernel  exclusive scan of the counts -> where each label's depth values go
//   leaf_scatter_kernel  groups the depth values by label (order inside a group is irrelevant)
//   leaf_median_kernel   exact np.median per label by radix selection on the grouped values
//   edt_row_kernel / edt_rowmax_kernel   row pass of the exact squared Euclidean distance transform of (labels >= 1),
//                        as a pruned search for the background pixel farthest from every leaf (the only thing
//                        leaf_scorer.py:67-71 takes from its distance field)
//   edt_col_kernel       stand-alone column pass (lg_edt_squared on a caller-supplied mask)
//   select_leaf_kernel   the per-leaf scores, tall-leaf rule, Pareto front and weighted pick
//
// Integer sums are exact and order independent, so results do not depend on scheduling: coordinate
// sums are 64-bit integers, depth and ray-length sums are fixed point (2^-28 m and 2^-36).
#include <math_constants.h>

#include "lg_internal.cuh"

namespace {

constexpr int ST_NT = 256;
constexpr int ST_PX = 8;
constexpr double DEP_SCALE = 268435456.0;       // 2^28
constexpr double DIST_SCALE = 68719476736.0;    // 2^36

struct SmemLeaf {
    unsigned cnt, sx, sy, bx0, bx1, by0, by1, border, kmin, kmax;
    unsigned long long sdep, sdist;
};

__device__ __forceinline__ unsigned f2key(float f) {   // order-preserving float -> uint32 key
    unsigned u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
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
        c.seg_cur[i] = 0;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        c.first_leaf[i] = 0xFFFFFFFFu; c.edt_best[i] = 0ull; c.status[i] = 0; c.list_n[i] = 0;
    }
}

// ray_tab[y * W + x] = sum over rows y' <= y of round(2^36 * sqrt(((x - cx)^2 + (y' - cy)^2) / f^2 + 1)): column-wise prefix
// sums of the length of the viewing ray through a pixel per unit depth (leaf_scorer.py:104-113 with X = md (x - cx) / f,
// Y = md (y - cy) / f, Z = md).  The sum over a vertical run of pixels is then a difference of two entries.  The table
// depends on the camera only: it is built once per camera and read (L2-resident) by every frame.
__global__ void ray_table_kernel(unsigned long long* __restrict__ tab, int H, int W, lg_camera cam) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const double inv_f2 = 1.0 / (cam.f * cam.f);
    const double ddx = (double)x - cam.cx;
    unsigned long long acc = 0;
    for (int y = 0; y < H; ++y) {
        const double ddy = (double)y - cam.cy;
        const double sv = sqrt((ddx * ddx + ddy * ddy) * inv_f2 + 1.0);
        acc += (unsigned long long)__double2ll_rn(sv * DIST_SCALE);
        tab[(size_t)y * W + x] = acc;
    }
}

// One pass over labels + depth.  A thread walks down one column: labels are piecewise constant along a column
// (a leaf is ~100 rows tall), so the thread carries the sums of its current run in registers and updates the CTA's
// shared-memory table only when the label changes - about ten updates per column instead of one per 8 pixels.
// Everything that depends on the pixel position only has a closed form per run (count, coordinate sums, bounding
// box, border contact) or is a difference of two ray_tab entries; per pixel only the depth is accumulated (2^-28
// fixed point, exact and order independent) and its key range tracked.  Loads run ST_U rows ahead.
constexpr int ST_U = 8;
#ifndef LG_STATS_NT
#define LG_STATS_NT 128
#endif
#ifndef LG_STATS_MINB
#define LG_STATS_MINB 8
#endif
constexpr int STC_NT = LG_STATS_NT;   // columns per CTA
constexpr int STC_BND = 40;   // leaf-run boundaries (20 runs) a column can record for the distance transform

// the vertical run [ya, yb] of label `cur` in column x ends: add its sums to the CTA's table (rare: ~10 per column)
__device__ __noinline__ void stats_flush_run(SmemLeaf* tab, int cur, int x, int ya, int yb, int W, int H, unsigned kmn, unsigned kmx,
                                             long long sdep, const unsigned long long* rt) {
    SmemLeaf* t = &tab[cur];
    const unsigned len = (unsigned)(yb - ya + 1);
    atomicAdd(&t->cnt, len);
    if (cur > 0) {
        atomicAdd(&t->sx, (unsigned)x * len);
        atomicAdd(&t->sy, (unsigned)(ya + yb) * len / 2u);
        atomicMin(&t->bx0, (unsigned)x); atomicMax(&t->bx1, (unsigned)x);
        atomicMin(&t->by0, (unsigned)ya); atomicMax(&t->by1, (unsigned)yb);
        if (x == 0 || x == W - 1 || ya == 0 || yb == H - 1) atomicOr(&t->border, 1u);
        atomicMin(&t->kmin, kmn); atomicMax(&t->kmax, kmx);
        atomicAdd(&t->sdep, (unsigned long long)sdep);
        const unsigned long long hi = rt[(size_t)yb * W], lo = ya > 0 ? rt[(size_t)(ya - 1) * W] : 0ull;
        atomicAdd(&t->sdist, hi - lo);
    }
}
// The same walk is the column pass of the union distance transform (edt_col_kernel with source = label >= 1): the
// kernel also writes the column distances c.edt_g and, on the way back up, their chunk minima c.edt_gmin.
__global__ void __launch_bounds__(STC_NT, LG_STATS_MINB) leaf_stats_kernel(lg_context c, const int16_t* __restrict__ labels,
                                                            const float* __restrict__ depth) {
    extern __shared__ SmemLeaf tab[];
    __shared__ unsigned s_first, s_bad;
    const int L = c.L, W = c.W, H = c.H;
    const size_t P = c.P;
    const int b = blockIdx.y;
    // rows where this column enters (even entries) / leaves (odd entries) the union of all leaves: [STC_BND][STC_NT]
    uint16_t* bnd = reinterpret_cast<uint16_t*>(tab + L) + threadIdx.x;
    int nb = 0;
    bool in_src = false;
    for (int l = threadIdx.x; l < L; l += STC_NT) {
        SmemLeaf z;
        z.cnt = 0; z.sx = 0; z.sy = 0; z.bx0 = 0xFFFFFFFFu; z.by0 = 0xFFFFFFFFu; z.bx1 = 0; z.by1 = 0;
        z.border = 0; z.kmin = 0xFFFFFFFFu; z.kmax = 0; z.sdep = 0; z.sdist = 0;
        tab[l] = z;
    }
    if (threadIdx.x == 0) { s_first = 0xFFFFFFFFu; s_bad = 0; }
    __syncthreads();
    const int x = blockIdx.x * STC_NT + threadIdx.x;
    if (x < W) {
        const int16_t* lp = labels + (size_t)b * P + x;
        const float* dp = depth + (size_t)b * P + x;
        const unsigned long long* rt = c.ray_tab + x;
        int cur = -1, ya = 0;
        bool seen_leaf = false;
        long long sdep = 0;
        unsigned kmn = 0xFFFFFFFFu, kmx = 0;
        auto step = [&](int y, int l, float dv) {     // one pixel of the downward walk
            if ((l >= 1) != in_src) {                 // the column enters / leaves the union of the leaves
                if (nb < STC_BND) bnd[nb * STC_NT] = (uint16_t)y;
                ++nb;
                in_src = !in_src;
            }
            if (l < 0 || l >= L) { s_bad = 1; l = -1; }
            if (l != cur) {
                if (cur >= 0) stats_flush_run(tab, cur, x, ya, y - 1, W, H, kmn, kmx, sdep, rt);
                cur = l; ya = y; sdep = 0; kmn = 0xFFFFFFFFu; kmx = 0;
                if (l >= 1 && !seen_leaf) { atomicMin(&s_first, (unsigned)((size_t)y * W + x)); seen_leaf = true; }
            }
            if (l > 0) {
                // 2^28 * depth is exact in float32 (power-of-two scale), so this equals the float64 formulation
                const float dd = fminf(fmaxf(dv, -2048.f), 2048.f);
                sdep += __float2ll_rn(dd * 268435456.f);
                const unsigned key = f2key(dv);
                kmn = min(kmn, key); kmx = max(kmx, key);
            }
        };
        int nl[ST_U];
        float nd[ST_U];
#pragma unroll
        for (int k = 0; k < ST_U; ++k) {
            nl[k] = k < H ? (int)lp[(size_t)k * W] : 0;
            nd[k] = k < H ? dp[(size_t)k * W] : 0.f;
        }
        const int16_t* lrow = lp;        // row y0 of this column
        const float* drow = dp;
        const size_t bump = (size_t)ST_U * W;
        int y0 = 0;
        for (; y0 + ST_U <= H; y0 += ST_U) {     // full batches: no bounds checks, running row pointers
            int cl[ST_U];
            float cd[ST_U];
#pragma unroll
            for (int k = 0; k < ST_U; ++k) { cl[k] = nl[k]; cd[k] = nd[k]; }
            if (y0 + 2 * ST_U <= H) {            // next batch in flight while this one is reduced
#pragma unroll
                for (int k = 0; k < ST_U; ++k) { nl[k] = (int)lrow[bump + (size_t)k * W]; nd[k] = drow[bump + (size_t)k * W]; }
            } else {
#pragma unroll
                for (int k = 0; k < ST_U; ++k) {
                    const bool ok = y0 + ST_U + k < H;
                    nl[k] = ok ? (int)lrow[bump + (size_t)k * W] : 0;
                    nd[k] = ok ? drow[bump + (size_t)k * W] : 0.f;
                }
            }
#pragma unroll
            for (int k = 0; k < ST_U; ++k) step(y0 + k, cl[k], cd[k]);
            lrow += bump; drow += bump;
        }
#pragma unroll
        for (int k = 0; k < ST_U; ++k)           // the last H % ST_U rows (already fetched)
            if (y0 + k < H) step(y0 + k, nl[k], nd[k]);
        if (cur >= 0) stats_flush_run(tab, cur, x, ya, H - 1, W, H, kmn, kmx, sdep, rt);
        if (in_src) {                             // close the last run at the bottom edge
            if (nb < STC_BND) bnd[nb * STC_NT] = (uint16_t)H;
            ++nb;
        }
    }
    __syncthreads();
    for (int l = threadIdx.x; l < L; l += STC_NT) {
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
    if (threadIdx.x == 0) {
        if (s_first != 0xFFFFFFFFu) atomicMin(&c.first_leaf[b], s_first);
        if (s_bad) atomicOr(&c.status[b], LG_ST_LABEL_RANGE);
    }
    {   // Column pass of the union distance transform, from the recorded run boundaries (no second look at the labels):
        // g(y) = 0 inside a run, else the distance to the nearest run end above / run start below; plus the minimum of
        // g over every chunk of 32 columns.  Whole warps take part: the reduction needs every lane.
        const bool in = x < W;
        uint16_t* gp = c.edt_g + (size_t)b * P + (in ? x : 0);
        uint16_t* gm = c.edt_gmin + (size_t)b * H * c.edt_nchunks;
        const int chunk = x >> 5, lane = threadIdx.x & 31, nchunks = c.edt_nchunks;
        const bool wr_min = lane == 0 && chunk < nchunks;
        const bool overflow = nb > STC_BND;
        if (in && overflow) {      // more runs than the table holds (never on real frames): plain two sweeps for this column
            const int16_t* lp = labels + (size_t)b * P + x;
            unsigned d = 0xFFFFu;
            for (int y = 0; y < H; ++y) { d = lp[(size_t)y * W] >= 1 ? 0u : min(d + 1u, 0xFFFFu); gp[(size_t)y * W] = (uint16_t)d; }
            d = 0xFFFFu;
            for (int y = H - 1; y >= 0; --y) { d = min((unsigned)gp[(size_t)y * W], min(d + 1u, 0xFFFFu)); gp[(size_t)y * W] = (uint16_t)d; }
        }
        int ri = 0;
        int next_start = (in && !overflow && nb > 0) ? (int)bnd[0] : 0x7FFFFFF;
        int cur_last = -1;           // last row of the run the walk is in, -1 when in a gap
        int last_src = -0x7FFFFFF;   // last leaf row above
        uint16_t* grow = gp;
        uint16_t* mrow = gm + chunk;
        for (int y = 0; y < H; ++y) {
            unsigned gv = 0xFFFFu;
            if (in) {
                if (overflow) {
                    gv = *grow;
                } else {
                    if (y == next_start) {
                        cur_last = (int)bnd[(ri + 1) * STC_NT] - 1;
                        ri += 2;
                        next_start = ri < nb ? (int)bnd[ri * STC_NT] : 0x7FFFFFF;
                    }
                    if (cur_last >= 0) {
                        gv = 0u;
                        if (y == cur_last) { last_src = y; cur_last = -1; }
                    } else {
                        gv = (unsigned)min(min(y - last_src, next_start - y), 0xFFFF);
                    }
                    *grow = (uint16_t)gv;
                }
            }
            const unsigned m = __reduce_min_sync(0xFFFFFFFFu, gv);
            if (wr_min) *mrow = (uint16_t)m;
            grow += W; mrow += nchunks;
        }
    }
}

// background id = smallest id present (torch.unique(mask)[1:], leaf_scorer.py:32)
__device__ __forceinline__ int background_id(const uint32_t* cnt, int L) {
    for (int l = 0; l < L; ++l)
        if (cnt[l]) return l;
    return -1;
}

__global__ void leaf_offsets_kernel(lg_context c, int n) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const uint32_t* cnt = c.cnt + (size_t)b * c.L;
    uint32_t* off = c.seg_off + (size_t)b * (c.L + 1);
    int bg = background_id(cnt, c.L);
    uint32_t o = 0;
    for (int l = 0; l < c.L; ++l) {
        off[l] = o;
        if (cnt[l] && l != bg) o += cnt[l];
    }
    off[c.L] = o;
}

// Groups the depth values by label.  Global atomics are taken once per (CTA, label): the CTA counts its
// pixels per label in shared memory, reserves one contiguous range per label in the frame's segment and
// hands out positions inside it from shared-memory cursors.
__global__ void __launch_bounds__(ST_NT) leaf_scatter_kernel(lg_context c, const int16_t* __restrict__ labels,
                                                              const float* __restrict__ depth) {
    extern __shared__ unsigned s_cur[];   // [L] count, then absolute write cursor
    const int L = c.L;
    const size_t P = c.P;
    const int b = blockIdx.y;
    const size_t base = ((size_t)blockIdx.x * ST_NT + threadIdx.x) * ST_PX;
    const uint32_t* cnt = c.cnt + (size_t)b * L;
    __shared__ int s_bg;
    if (threadIdx.x == 0) s_bg = background_id(cnt, L);
    for (int l = threadIdx.x; l < L; l += ST_NT) s_cur[l] = 0;
    __syncthreads();
    const int bg = s_bg;
    const int16_t* lp = labels + (size_t)b * P;
    const float* dp = depth + (size_t)b * P;
    int lab[ST_PX];
    float val[ST_PX];
    const int npx = base < P ? (int)min((size_t)ST_PX, P - base) : 0;
    if (npx == ST_PX && ((reinterpret_cast<uintptr_t>(lp + base) | reinterpret_cast<uintptr_t>(dp + base)) & 15) == 0) {
        const uint4 lv = *reinterpret_cast<const uint4*>(lp + base);
        const float4 d0 = *reinterpret_cast<const float4*>(dp + base), d1 = *reinterpret_cast<const float4*>(dp + base + 4);
        const unsigned w[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { lab[2 * k] = (int16_t)(w[k] & 0xFFFFu); lab[2 * k + 1] = (int16_t)(w[k] >> 16); }
        val[0] = d0.x; val[1] = d0.y; val[2] = d0.z; val[3] = d0.w; val[4] = d1.x; val[5] = d1.y; val[6] = d1.z; val[7] = d1.w;
    } else {
#pragma unroll
        for (int k = 0; k < ST_PX; ++k) {
            lab[k] = k < npx ? (int)lp[base + k] : -1;
            val[k] = k < npx ? dp[base + k] : 0.f;
        }
    }
    // run starts: bit k set when pixel k opens a run; labels that are not scattered become -1
    unsigned starts = 0;
#pragma unroll
    for (int k = 0; k < ST_PX; ++k) {
        if (lab[k] < 0 || lab[k] >= L || lab[k] == bg) lab[k] = -1;
        if (k == 0 || lab[k] != lab[k - 1]) starts |= 1u << k;
    }
#pragma unroll
    for (int k = 0; k < ST_PX; ++k) {
        if (((starts >> k) & 1u) && lab[k] >= 0) {
            const unsigned rest = starts >> (k + 1);
            const int len = rest ? __ffs(rest) : ST_PX - k;
            atomicAdd(&s_cur[lab[k]], (unsigned)len);
        }
    }
    __syncthreads();
    for (int l = threadIdx.x; l < L; l += ST_NT) {
        const unsigned n = s_cur[l];
        if (n) s_cur[l] = c.seg_off[(size_t)b * (L + 1) + l] + atomicAdd(&c.seg_cur[(size_t)b * L + l], n);
    }
    __syncthreads();
    float* seg = c.seg + (size_t)b * P;
    unsigned pos = 0;
#pragma unroll
    for (int k = 0; k < ST_PX; ++k) {
        if (lab[k] >= 0) {
            if ((starts >> k) & 1u) {
                const unsigned rest = starts >> (k + 1);
                const int len = rest ? __ffs(rest) : ST_PX - k;
                pos = atomicAdd(&s_cur[lab[k]], (unsigned)len);
            }
            seg[pos++] = val[k];
        }
    }
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
// np.median(depth[labels == l]) for every label of every frame (leaf_scorer.py:41-47): radix selection on the
// values grouped by leaf_scatter_kernel, two key bits per pass.  The passes start at the first bit in which the
// label's smallest and largest key differ (leaf_stats_kernel), and as soon as the surviving candidates fit in
// shared memory they are compacted there, so only the first two or three passes stream the whole segment.
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
    if (n == 0 || l == s_bg) {
        if (threadIdx.x == 0) c.median[(size_t)b * L + l] = CUDART_NAN_F;
        return;
    }
    const float* v = c.seg + (size_t)b * c.P + c.seg_off[(size_t)b * (L + 1) + l];
    const unsigned kmin = c.kmin[(size_t)b * L + l], kmax = c.kmax[(size_t)b * L + l];
    const unsigned k_lo = (n & 1) ? n / 2 : n / 2 - 1;   // rank of the lower middle
    unsigned klo;            // key of rank k_lo
    unsigned below = 0;      // elements whose key is smaller than every current candidate
    unsigned m = n;          // current candidates
    bool in_smem = false;
    unsigned set_below = 0, set_m = 0;   // the compacted set: its size and the number of elements below it
    const int lane = threadIdx.x & 31;
    auto compact = [&](unsigned prefix, unsigned pmask) {   // candidates (key & pmask) == prefix -> s_keys
        for (unsigned i0 = 0; i0 < n; i0 += MED_NT) {
            const unsigned i = i0 + threadIdx.x;
            unsigned key = 0;
            bool hit = false;
            if (i < n) { key = (f2key(v[i]) - kmin); hit = (key & pmask) == prefix; }
            const unsigned ball = __ballot_sync(0xFFFFFFFFu, hit);
            if (ball) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(&s_n, __popc(ball));
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (hit) s_keys[base + __popc(ball & ((1u << lane) - 1u))] = key;
            }
        }
        __syncthreads();
    };
    // keys are ranked relative to the label's smallest key: the search range is [0, kmax - kmin], so the first
    // digit already splits the values that are present instead of a power-of-two block around them
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
            // Segments that do not fit: the first pass over the segment resolves the top 7-8 bits of the range at once
            // (a 256-bin histogram in shared memory), which normally leaves few enough candidates to compact them in
            // the second pass; two-bit digits would stream the segment once per factor of four.
            int s0 = max(top - 7, 0);
            s0 += s0 & 1;                                      // even, so that the two-bit rounds end at bit 0
            for (int i = threadIdx.x; i < 256; i += MED_NT) s_hist[i] = 0;
            __syncthreads();
            for (unsigned i = threadIdx.x; i < n; i += MED_NT) atomicAdd(&s_hist[(f2key(v[i]) - kmin) >> s0], 1u);
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
                for (unsigned i = threadIdx.x; i < n; i += MED_NT) {
                    const unsigned key = (f2key(v[i]) - kmin);
                    if ((key & pmask) == prefix) {
                        const unsigned d = (key >> shift) & 3u;
                        c0 += (d == 0); c1 += (d == 1); c2 += (d == 2);
                    }
                }
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
                for (unsigned i = threadIdx.x; i < n; i += MED_NT) { const unsigned key = (f2key(v[i]) - kmin); if (key > klo) mn = min(mn, key); }
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
struct EdtSrc {
    const int16_t* labels;   // source (distance 0) where labels >= 1
    const uint8_t* mask;     // source where mask == 0
    __device__ __forceinline__ bool is_source(size_t i) const { return labels ? (labels[i] >= 1) : (mask[i] == 0); }
};

// pass 1: per column, distance (in rows) to the nearest source pixel in that column; 0xFFFF = none.
// gmin (optional): for every row the minimum of g over each chunk of 32 columns (one warp), [H][nchunks] per frame -
// the arg-max search reads these instead of the rows themselves.
__global__ void edt_col_kernel(EdtSrc src, uint16_t* __restrict__ g, uint16_t* __restrict__ gmin, int nchunks, int H, int W,
                               size_t P) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = x < W;
    const int xc = in ? x : W - 1;             // out-of-range lanes shadow the last column and store nothing
    const size_t fo = (size_t)blockIdx.y * P;
    uint16_t* gp = g + fo;
    constexpr int U = 8;    // rows whose loads are in flight together (the sweep itself is sequential)
    unsigned d = 0xFFFFu;
    for (int y0 = 0; y0 < H; y0 += U) {
        bool srcv[U];
#pragma unroll
        for (int k = 0; k < U; ++k) srcv[k] = (y0 + k < H) ? src.is_source(fo + (size_t)(y0 + k) * W + xc) : false;
#pragma unroll
        for (int k = 0; k < U; ++k) {
            if (y0 + k < H) {
                d = srcv[k] ? 0u : min(d + 1u, 0xFFFFu);
                if (in) gp[(size_t)(y0 + k) * W + x] = (uint16_t)d;
            }
        }
    }
    d = 0xFFFFu;
    const int chunk = x >> 5, lane = threadIdx.x & 31;
    uint16_t* gm = gmin ? gmin + (size_t)blockIdx.y * H * nchunks : nullptr;
    for (int y0 = H - 1; y0 >= 0; y0 -= U) {
        unsigned cur[U];
#pragma unroll
        for (int k = 0; k < U; ++k) cur[k] = (y0 - k >= 0) ? gp[(size_t)(y0 - k) * W + xc] : 0xFFFFu;
#pragma unroll
        for (int k = 0; k < U; ++k) {
            if (y0 - k >= 0) {
                d = min(cur[k], min(d + 1u, 0xFFFFu));
                if (in) gp[(size_t)(y0 - k) * W + x] = (uint16_t)d;
                if (gm) {
                    const unsigned m = __reduce_min_sync(0xFFFFFFFFu, in ? d : 0xFFFFu);
                    if (lane == 0 && chunk < nchunks) gm[(size_t)(y0 - k) * nchunks + chunk] = (uint16_t)m;
                }
            }
        }
    }
}

constexpr int EDT_NT = 256;
// pass 2: per row, d2(x) = min over x' of (x - x')^2 + g(x')^2.  The search around x stops as soon
// as the horizontal offset alone exceeds the best distance found (Meijster's lower-envelope bound).
//
// When only the arg-max of the field is wanted (d2out == nullptr, the leaf-selection use) a pixel is
// dropped as soon as its running upper bound falls below the largest exact distance found so far in
// the frame (best[b], shared by all CTAs through atomicMax): such a pixel can neither be the maximum
// nor tie with it, so the result is exactly the first maximum of the full field whatever the
// scheduling.  CTAs walk rows `stride` apart (a permutation of the rows) so that the bound comes
// from all over the frame early, and every CTA handles several rows to profit from it.
__global__ void __launch_bounds__(EDT_NT) edt_row_kernel(const uint16_t* __restrict__ g, uint32_t* __restrict__ d2out,
                                                          unsigned long long* __restrict__ best, int H, int W, size_t P,
                                                          int row_stride, int yi_begin, int yi_end) {
    extern __shared__ unsigned srow[];   // g squared, 0xFFFFFFFF = no source in that column
    __shared__ unsigned long long sbest[EDT_NT / 32];
    __shared__ unsigned schunk[128];     // minimum of srow over each 32-column chunk (rows up to 4096 wide)
    __shared__ unsigned s_lb;            // pruning bound of the current row
    constexpr int MAXV = 16;             // row values per thread held in registers: rows up to 4096 wide
    const int b = blockIdx.x;
    const bool prune = (d2out == nullptr) && (best != nullptr);
    unsigned long long mybest = 0;       // CTA-wide best so far (identical in every thread)
    const int nv = (W + EDT_NT - 1) / EDT_NT;
    // the next row's column distances are fetched while the current row is searched
    unsigned short nxt[MAXV];
    auto fetch = [&](int yi) {
        if (yi < yi_end) {
            const int y = (int)(((long long)yi * row_stride) % H);
            const uint16_t* gp = g + (size_t)b * P + (size_t)y * W;
#pragma unroll
            for (int k = 0; k < MAXV; ++k) {
                const int x = threadIdx.x + k * EDT_NT;
                nxt[k] = (k < nv && x < W) ? gp[x] : (unsigned short)0xFFFFu;
            }
        }
    };
    fetch(yi_begin + blockIdx.y);
    for (int yi = yi_begin + blockIdx.y; yi < yi_end; yi += gridDim.y) {
        const int y = (int)(((long long)yi * row_stride) % H);
        int any = 0;
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const int x = threadIdx.x + k * EDT_NT;
            if (k < nv && x < W) {
                const unsigned v = nxt[k];
                srow[x] = (v == 0xFFFFu) ? 0xFFFFFFFFu : v * v;
                any |= (v != 0xFFFFu);
            }
        }
        // one thread samples the frame's running maximum: the bound must be the same for the whole CTA (it steers
        // barriers and warp shuffles below), and other CTAs raise best[b] concurrently
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
        fetch(yi + gridDim.y);
        unsigned long long rowbest = 0;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int nchunks = (W + 31) >> 5;
        if (any && lb) {   // minimum of every 32-column chunk: lets a warp discard a whole chunk at once
            for (int j = warp; j < nchunks; j += EDT_NT / 32) {
                const int x = (j << 5) + lane;
                unsigned v = x < W ? srow[x] : 0xFFFFFFFFu;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v = min(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
                if (lane == 0) schunk[j] = v;
            }
            __syncthreads();
        }
        if (any) {
            for (int j = warp; j < nchunks; j += EDT_NT / 32) {
                if (lb) {
                    // every pixel of chunk j is at most 31 + 32*dj columns away from the best column of chunk j +- dj
                    unsigned long long ub = (unsigned long long)schunk[j] + 31ull * 31ull;
                    for (int dj = 1; dj < nchunks; ++dj) {
                        const unsigned long long off = (unsigned long long)(32 * dj + 31) * (32 * dj + 31);
                        if (off >= lb || ub < lb) break;
                        if (j - dj >= 0) ub = min(ub, (unsigned long long)schunk[j - dj] + off);
                        if (j + dj < nchunks) ub = min(ub, (unsigned long long)schunk[j + dj] + off);
                    }
                    if (ub < lb) continue;
                }
                const int x = (j << 5) + lane;
                if (x >= W) continue;
                unsigned bestd = srow[x];
                if (bestd < lb) continue;
                if (lb) {   // probes at doubling offsets: almost every pixel near a source drops below the bound here
                    for (unsigned k = 1; k * k < bestd; k <<= 1) {
                        const int xl = x - (int)k, xr = x + (int)k;
                        if (xl < 0 && xr >= W) break;
                        const unsigned kk = k * k;
                        if (xl >= 0) { unsigned s = srow[xl]; if (s != 0xFFFFFFFFu) bestd = min(bestd, s + kk); }
                        if (xr < W) { unsigned s = srow[xr]; if (s != 0xFFFFFFFFu) bestd = min(bestd, s + kk); }
                    }
                    if (bestd < lb) continue;
                }
                for (unsigned k = 1; k * k < bestd; ++k) {
                    int xl = x - (int)k, xr = x + (int)k;
                    if (xl < 0 && xr >= W) break;
                    unsigned kk = k * k;
                    if (xl >= 0) { unsigned s = srow[xl]; if (s != 0xFFFFFFFFu) bestd = min(bestd, s + kk); }
                    if (xr < W) { unsigned s = srow[xr]; if (s != 0xFFFFFFFFu) bestd = min(bestd, s + kk); }
                    if (bestd < lb) break;
                }
                if (bestd < lb) continue;
                size_t idx = (size_t)y * W + x;
                if (d2out) d2out[(size_t)b * P + idx] = bestd;
                unsigned long long key = ((unsigned long long)bestd << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)idx);
                rowbest = max(rowbest, key);
            }
        } else if (d2out) {
            for (int x = threadIdx.x; x < W; x += EDT_NT) d2out[(size_t)b * P + (size_t)y * W + x] = 0xFFFFFFFFu;
        }
        // the reduction only runs when some thread improved on the CTA's best (rare once the bound is tight);
        // the barrier also protects srow against the next row's writers
        const int improved = __syncthreads_or(best != nullptr && rowbest > mybest);
        if (improved) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) rowbest = max(rowbest, __shfl_xor_sync(0xFFFFFFFFu, rowbest, d));
            if ((threadIdx.x & 31) == 0) sbest[threadIdx.x >> 5] = rowbest;
            __syncthreads();
            for (int w = 0; w < EDT_NT / 32; ++w) mybest = max(mybest, sbest[w]);
            if (threadIdx.x == 0 && mybest > *reinterpret_cast<volatile unsigned long long*>(&best[b])) atomicMax(&best[b], mybest);
            __syncthreads();             // sbest is reused by the next improving row
        }
    }
}

// Arg-max search, main part: ONE WARP PER ROW.  A row is first judged by its chunk minima alone (gmin, 1/32 of the
// data): chunk j cannot hold the maximum if min over j' of gmin(j')^2 + (32 |j - j'| + 31)^2 is below the frame's
// running maximum.  Only rows with a surviving chunk load their column distances (into the warp's shared-memory
// row) and run the exact search, for the surviving chunks only.  Exactness argument as for edt_row_kernel: a pixel
// is only dropped when an upper bound of its distance is below an exact distance found elsewhere in the frame.
constexpr int EDTW_NT = 256;
__global__ void __launch_bounds__(EDTW_NT) edt_rowmax_kernel(const uint16_t* __restrict__ g, const uint16_t* __restrict__ gmin,
                                                              unsigned long long* __restrict__ best, int nchunks, int H, int W,
                                                              size_t P, int row_stride, int yi_begin, int yi_end) {
    extern __shared__ unsigned sm_rows[];            // per warp: W squared column distances, then 128 chunk minima
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned* srow = sm_rows + (size_t)warp * (W + 128);
    unsigned* mrow = srow + W;
    const uint16_t* gmf = gmin + (size_t)b * H * nchunks;
    unsigned long long mybest = 0;                   // this warp's best (identical in all lanes)
    const int rows_per_pass = gridDim.y * (EDTW_NT / 32);
    for (int yi = yi_begin + blockIdx.y * (EDTW_NT / 32) + warp; yi < yi_end; yi += rows_per_pass) {
        const int y = (int)(((long long)yi * row_stride) % H);
        unsigned long long gb = 0;
        if (lane == 0) gb = *reinterpret_cast<volatile unsigned long long*>(&best[b]);
        gb = __shfl_sync(0xFFFFFFFFu, gb, 0);
        const unsigned lb = (unsigned)(max(gb, mybest) >> 32);
        __syncwarp();
        bool has_source = false;
        for (int j = lane; j < nchunks; j += 32) {
            const unsigned v = gmf[(size_t)y * nchunks + j];
            mrow[j] = (v == 0xFFFFu) ? 0xFFFFFFFFu : v * v;
            has_source |= v != 0xFFFFu;
        }
        if (!__any_sync(0xFFFFFFFFu, has_source)) continue;      // no source in any column of this row: nothing to rank
        __syncwarp();
        // chunks that may still hold a pixel at distance >= lb
        unsigned alive[4] = {0u, 0u, 0u, 0u};        // bit `lane` of alive[k]: chunk 32 k + lane survives
        bool any_alive = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = 32 * k + lane;
            bool keep = false;
            if (j < nchunks) {
                unsigned long long ub = (unsigned long long)mrow[j] + 31ull * 31ull;
                for (int dj = 1; dj < nchunks && ub >= lb; ++dj) {
                    const unsigned long long off = (unsigned long long)(32 * dj + 31) * (32 * dj + 31);
                    if (off >= lb) break;
                    if (j - dj >= 0) ub = min(ub, (unsigned long long)mrow[j - dj] + off);
                    if (j + dj < nchunks) ub = min(ub, (unsigned long long)mrow[j + dj] + off);
                }
                keep = ub >= lb;
            }
            alive[k] = __ballot_sync(0xFFFFFFFFu, keep);
            any_alive |= alive[k] != 0u;
        }
        if (!any_alive) continue;
        // exact search on the surviving chunks
        const uint16_t* gp = g + (size_t)b * P + (size_t)y * W;
        for (int x = lane; x < W; x += 32) {
            const unsigned v = gp[x];
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
                unsigned bestd = srow[x];
                if (bestd < lb) continue;
                if (lb) {
                    for (unsigned kk = 1; kk * kk < bestd; kk <<= 1) {
                        const int xl = x - (int)kk, xr = x + (int)kk;
                        if (xl < 0 && xr >= W) break;
                        const unsigned q = kk * kk;
                        if (xl >= 0) { const unsigned t = srow[xl]; if (t != 0xFFFFFFFFu) bestd = min(bestd, t + q); }
                        if (xr < W) { const unsigned t = srow[xr]; if (t != 0xFFFFFFFFu) bestd = min(bestd, t + q); }
                    }
                    if (bestd < lb) continue;
                }
                for (unsigned kk = 1; kk * kk < bestd; ++kk) {
                    const int xl = x - (int)kk, xr = x + (int)kk;
                    if (xl < 0 && xr >= W) break;
                    const unsigned q = kk * kk;
                    if (xl >= 0) { const unsigned t = srow[xl]; if (t != 0xFFFFFFFFu) bestd = min(bestd, t + q); }
                    if (xr < W) { const unsigned t = srow[xr]; if (t != 0xFFFFFFFFu) bestd = min(bestd, t + q); }
                    if (bestd < lb) break;
                }
                if (bestd < lb) continue;
                const unsigned idx = (unsigned)((size_t)y * W + x);
                rowbest = max(rowbest, ((unsigned long long)bestd << 32) | (unsigned long long)(0xFFFFFFFFu - idx));
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) rowbest = max(rowbest, __shfl_xor_sync(0xFFFFFFFFu, rowbest, d));
        if (rowbest > mybest) {
            mybest = rowbest;
            if (lane == 0 && mybest > *reinterpret_cast<volatile unsigned long long*>(&best[b])) atomicMax(&best[b], mybest);
        }
    }
}

// rows are visited in the order (i * stride) mod H: stride ~ 0.618 H, coprime with H
static int edt_row_stride(int H) {
    auto gcd = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
    int s = (int)(H * 0.6180339887);
    if (s < 1) s = 1;
    while (gcd(s, H) != 1) ++s;
    return s % H ? s % H : 1;
}
static dim3 edt_row_grid(int n, int H) {
    int per_frame = (148 * 8 * 2 + n - 1) / n;     // ~two waves of 8 CTAs per SM over the whole batch
    if (per_frame > H) per_frame = H;
    if (per_frame < 1) per_frame = 1;
    return dim3(n, per_frame);
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

// row pass for n frames; d2 == nullptr: arg-max only (pruned search).  The arg-max search first runs a few well
// spread rows per frame one after the other (edt_row_kernel, one CTA per frame: the first row is searched in full,
// the following ones against the bound it leaves), then all remaining rows with one warp per row.
static int run_edt_rows(lg_context* c, int n, uint32_t* d2, cudaStream_t st) {
    const int stride = edt_row_stride(c->H);
    const size_t sm = c->W * sizeof(unsigned);
    if (d2 || c->H < 64) {
        edt_row_kernel<<<edt_row_grid(n, c->H), EDT_NT, sm, st>>>(c->edt_g, d2, c->edt_best, c->H, c->W, c->P, stride, 0, c->H);
        LG_LAUNCH_CHECK();
        return LG_OK;
    }
    const int seed = 16;
    int seed_ctas = 296 / n;                         // small batches: spread the seed rows over a few CTAs per frame
    seed_ctas = seed_ctas < 1 ? 1 : (seed_ctas > seed ? seed : seed_ctas);
    edt_row_kernel<<<dim3(n, seed_ctas), EDT_NT, sm, st>>>(c->edt_g, nullptr, c->edt_best, c->H, c->W, c->P, stride, 0, seed);
    LG_LAUNCH_CHECK();
    const size_t smw = (size_t)(EDTW_NT / 32) * (c->W + 128) * sizeof(unsigned);
    LG_ENSURE_SMEM(edt_rowmax_kernel, smw);
    int per_frame = (148 * 4 * 2 + n - 1) / n;       // about two waves of CTAs over the batch
    const int max_ctas = (c->H - seed + EDTW_NT / 32 - 1) / (EDTW_NT / 32);
    per_frame = per_frame < 1 ? 1 : (per_frame > max_ctas ? max_ctas : per_frame);
    edt_rowmax_kernel<<<dim3(n, per_frame), EDTW_NT, smw, st>>>(c->edt_g, c->edt_gmin, c->edt_best, c->edt_nchunks, c->H, c->W,
                                                                 c->P, stride, seed, c->H);
    LG_LAUNCH_CHECK();
    return LG_OK;
}

int lg_run_stage1(lg_context* c, const int16_t* labels, const float* depth, int n, lg_camera cam, cudaStream_t st) {
    clear_tables_kernel<<<64, 256, 0, st>>>(*c, n);
    LG_LAUNCH_CHECK();
    const int tiles = (int)((c->P + ST_NT * ST_PX - 1) / (ST_NT * ST_PX));
    if (!c->ray_valid || c->ray_cam.f != cam.f || c->ray_cam.cx != cam.cx || c->ray_cam.cy != cam.cy) {
        ray_table_kernel<<<(c->W + 127) / 128, 128, 0, st>>>(c->ray_tab, c->H, c->W, cam);
        LG_LAUNCH_CHECK();
        c->ray_cam = cam;
        c->ray_valid = 1;
    }
    // per-leaf statistics + column pass of the union distance transform in one walk over the columns
    LG_ENSURE_SMEM(leaf_stats_kernel, c->L * sizeof(SmemLeaf) + (size_t)STC_BND * STC_NT * sizeof(uint16_t));
    leaf_stats_kernel<<<dim3((c->W + STC_NT - 1) / STC_NT, n), STC_NT,
                        c->L * sizeof(SmemLeaf) + (size_t)STC_BND * STC_NT * sizeof(uint16_t), st>>>(*c, labels, depth);
    LG_LAUNCH_CHECK();
    lg_mark(c, LG_M_STATS, st);
    // the row pass (arg-max only) is independent of the medians: it runs beside scatter + median
    cudaStream_t aux = lg_fork(c, 0, st);
    lg_mark(c, LG_M_EDT_COL, aux);
    int rc = run_edt_rows(c, n, nullptr, aux);
    if (rc) return rc;
    lg_mark(c, LG_M_EDT_ROW, aux);
    leaf_offsets_kernel<<<(n + 63) / 64, 64, 0, st>>>(*c, n);
    LG_LAUNCH_CHECK();
    leaf_scatter_kernel<<<dim3(tiles, n), ST_NT, c->L * sizeof(unsigned), st>>>(*c, labels, depth);
    LG_LAUNCH_CHECK();
    lg_mark(c, LG_M_SCATTER, st);
    LG_ENSURE_SMEM(leaf_median_kernel, MED_CAP * sizeof(unsigned));
    leaf_median_kernel<<<dim3(c->L, n), MED_NT, MED_CAP * sizeof(unsigned), st>>>(*c);
    LG_LAUNCH_CHECK();
    lg_mark(c, LG_M_MEDIAN, st);
    return lg_join(c, 0, aux, st);
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
    EdtSrc src{nullptr, mask};
    edt_col_kernel<<<dim3((c->W + 127) / 128, n), 128, 0, st>>>(src, c->edt_g, d2 ? nullptr : c->edt_gmin, c->edt_nchunks, c->H, c->W,
                                                                c->P);
    LG_LAUNCH_CHECK();
    int rc = run_edt_rows(c, n, d2, st);
    if (rc) return rc;
    if (argmax) {
        edt_argmax_out_kernel<<<(n + 63) / 64, 64, 0, st>>>(c->edt_best, argmax, n);
        LG_LAUNCH_CHECK();
    }
    return LG_OK;
}
