// Self-supervised training samples as a by-product of grasp selection
// (reference scripts/utils/ml_grasp_optimizer/data_collector.py:83-348, 420-487).
//
// Runs after lg_process_batch / lg_select_grasp_point on the same frames: the score maps, the chamfer field, the
// leaf bitmask and the winning contour of every frame are still in the context.  One CTA per frame produces
//   slot 0      the positive sample: raw 32x32 windows of depth, mask and the seven score maps at the grasp point
//   slots 1..3  its rot90 copies with depth noise and score jitter (:250-299)
//   slots 4..6  up to three negatives drawn from the tip / stem / edge candidate sets (:301-348)
// The reference draws from the global generators of `random` and `torch`; here every draw is a pure function of
// (seed, frame index, purpose, counter) - oracle.CollectorRng restates it and the golden vectors were made by
// injecting that generator into the unmodified reference.
//
// Scratch (all free once the batch has been processed): the collector's own tip lists (allocated by its first call); chamfer forward-pass
// buffer -> the two erosion maps; hull work array -> header, stem row offsets, border turn-back points.
#include <math_constants.h>

#include "lg_internal.cuh"

namespace {

constexpr int CO_NT = 512;
constexpr int CO_WARPS = CO_NT / 32;
constexpr int CO_HDR = 16;
constexpr int PSZ = LG_CHANNELS * LG_PATCH * LG_PATCH;
constexpr int PPX = LG_PATCH * LG_PATCH;
enum { HD_NTIP = 0, HD_QTIP, HD_NSTEM, HD_NEDGE, HD_S0, HD_SX, HD_SY, HD_EDGE_CAP, HD_EDGE_OVER, HD_READY };

struct Bits {
    const uint32_t* w;
    int wpr, bw, bh;
    __device__ __forceinline__ int get(int lx, int ly) const {
        if (lx < 0 || ly < 0 || lx >= bw || ly >= bh) return 0;
        return (w[ly * wpr + (lx >> 5)] >> (lx & 31)) & 1u;
    }
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
struct Rng {
    unsigned long long base;
    __device__ unsigned long long draw(unsigned stream, unsigned a, unsigned b) const {
        return mix64(base ^ (((unsigned long long)stream << 56) | ((unsigned long long)a << 32) | b));
    }
    __device__ double u01(unsigned stream, unsigned a, unsigned b) const { return (double)(draw(stream, a, b) >> 11) * 0x1p-53; }
};
enum { RNG_NOISE_FACTOR = 1, RNG_SCORE_JITTER = 2, RNG_PICK = 3, RNG_NORMAL = 4 };

struct Frame {
    const lg_context* c;
    LgMaskSrc src;
    const float* depth;
    LgRegion r;
    Bits B;
    int ox, oy, W, H, id;
    size_t fo;
    __device__ __forceinline__ int leaf(int x, int y) const { return B.get(x - ox, y - oy); }
};

// CTA-wide sum of an int; every thread gets the result.  s_red: CO_WARPS ints.
__device__ int block_sum(int v, int* s_red) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < CO_WARPS; ++w) t += s_red[w];
    return t;
}

// Exclusive position of this thread's flag among the CTA's flags (thread order), and the CTA total.
__device__ int block_rank(bool flag, int* s_red, int* total) {
    const unsigned ball = __ballot_sync(0xFFFFFFFFu, flag);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = __popc(ball);
    __syncthreads();
    int before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < CO_WARPS; ++w) {
        const int n = s_red[w];
        if (w < warp) before += n;
        all += n;
    }
    *total = all;
    return before + __popc(ball & ((1u << lane) - 1u));
}

// The nine raw channels of pixel (x, y) exactly as the reference's maps hold them, cast to float32 (:139-156).
__device__ __forceinline__ void raw_channels(const Frame& F, int x, int y, float* v) {
    const lg_context& c = *F.c;
    const size_t p = (size_t)y * F.W + x, fo = F.fo;
    const bool M = F.src.at(fo, p, F.id);
    const bool in = x >= F.r.sx0 && x < F.r.sx1 && y >= F.r.sy0 && y < F.r.sy1;
    const double step = (0.2 - 1.0) / (double)(F.H - 1);
    v[0] = F.depth[fo + p];
    v[1] = M ? 1.f : 0.f;
    v[2] = in ? (float)c.m_sdf[fo + p] : 0.f;
    v[3] = in ? (float)c.m_app[fo + p] : 0.f;
    v[4] = in ? c.m_flat[fo + p] : 1.f;
    v[5] = M ? (float)((y == F.H - 1) ? 0.2 : ((double)y * step + 1.0)) : 0.f;
    v[6] = M ? c.di[fo + p] : 0.f;
    v[7] = in ? (float)c.m_acc[fo + p] : 0.f;
    v[8] = in ? c.m_stem[fo + p] : 0.f;
}

// _extract_patches (:91-173): the window must lie inside the image (plain slicing), depth and scores must be finite
// and the mask window must not be empty.  On success the raw patch is written to `out`.  Uniform return value.
__device__ bool extract(const Frame& F, int x, int y, float* __restrict__ out, int* s_red) {
    if (x - 16 < 0 || y - 16 < 0 || x + 16 > F.W || y + 16 > F.H) return false;
    float v[PPX / CO_NT][LG_CHANNELS];
    int bad = 0, any = 0;
#pragma unroll
    for (int q = 0; q < PPX / CO_NT; ++q) {
        const int pi = threadIdx.x + q * CO_NT;
        raw_channels(F, x - 16 + (pi & 31), y - 16 + (pi >> 5), v[q]);
#pragma unroll
        for (int ch = 0; ch < LG_CHANNELS; ++ch) bad |= !isfinite(v[q][ch]);
        any |= v[q][1] != 0.f;
    }
    const int nbad = block_sum(bad, s_red);
    const int nany = block_sum(any, s_red);
    if (nbad || !nany) return false;
#pragma unroll
    for (int q = 0; q < PPX / CO_NT; ++q)
#pragma unroll
        for (int ch = 0; ch < LG_CHANNELS; ++ch) out[ch * PPX + threadIdx.x + q * CO_NT] = v[q][ch];
    return true;
}

// _rotate_point (:402-418): double arithmetic with numpy's cos / sin of np.radians(90 k), int() truncation.
__device__ void rotate_point(int x, int y, int k, int* nx, int* ny) {
    const double COS[4] = {1.0, 6.123233995736766e-17, -1.0, -1.8369701987210297e-16};
    const double SIN[4] = {0.0, 1.0, 1.2246467991473532e-16, -1.0};
    const double xd = (double)(x - 16), yd = (double)(y - 16);
    const double rx = __dsub_rn(__dmul_rn(xd, COS[k]), __dmul_rn(yd, SIN[k]));
    const double ry = __dadd_rn(__dmul_rn(xd, SIN[k]), __dmul_rn(yd, COS[k]));
    *nx = (int)__dadd_rn(rx, 16.0);
    *ny = (int)__dadd_rn(ry, 16.0);
}

// ---- candidate sets ----------------------------------------------------------------------------------------
// tip points (:420-440): leaf pixels whose chamfer value is the maximum of its 5x5 neighbourhood, in raster order.
__device__ int build_tip_list(const Frame& F, uint32_t* tip_idx, float* tip_val, int* s_red) {
    const LgRegion& r = F.r;
    const int bw = r.x1 - r.x0, bh = r.y1 - r.y0, area = bw * bh;
    const float* di = F.c->di + F.fo;
    int n = 0;
    for (int base = 0; base < area; base += CO_NT) {
        const int i = base + threadIdx.x;
        bool is_max = false;
        int x = 0, y = 0;
        float d = 0.f;
        if (i < area) {
            y = r.y0 + i / bw; x = r.x0 + i % bw;
            if (F.leaf(x, y)) {
                d = di[(size_t)y * F.W + x];
                is_max = true;
                for (int dy = -2; dy <= 2 && is_max; ++dy)
                    for (int dx = -2; dx <= 2; ++dx) {
                        const int xx = x + dx, yy = y + dy;
                        if (xx < 0 || yy < 0 || xx >= F.W || yy >= F.H || !F.leaf(xx, yy)) continue;
                        if (di[(size_t)yy * F.W + xx] > d) { is_max = false; break; }
                    }
            }
        }
        int total;
        const int pos = block_rank(is_max, s_red, &total);
        if (is_max) { tip_idx[n + pos] = (uint32_t)(y * F.W + x); tip_val[n + pos] = d; }
        n += total;
    }
    __syncthreads();
    return n;
}

// Element of rank `rank` when the tip list is sorted by value descending, raster order among equals (list.sort is
// stable): the largest T with count(value >= T) > rank, then the (rank - count(value > T))-th entry equal to T.
__device__ uint32_t tip_select(const uint32_t* tip_idx, const float* tip_val, int n, int rank, int* s_red, int* s_out) {
    unsigned T = 0;
    for (int bit = 30; bit >= 0; --bit) {          // chamfer values are positive floats: their bit patterns order like ints
        const unsigned cand = T | (1u << bit);
        int cnt = 0;
        for (int i = threadIdx.x; i < n; i += CO_NT) cnt += __float_as_uint(tip_val[i]) >= cand;
        if (block_sum(cnt, s_red) > rank) T = cand;
    }
    int gt = 0;
    for (int i = threadIdx.x; i < n; i += CO_NT) gt += __float_as_uint(tip_val[i]) > T;
    int m = rank - block_sum(gt, s_red);
    for (int base = 0; base < n; base += CO_NT) {
        const int i = base + threadIdx.x;
        const bool eq = i < n && __float_as_uint(tip_val[i]) == T;
        int total;
        const int pos = block_rank(eq, s_red, &total);
        if (eq && pos == m) *s_out = (int)tip_idx[i];
        if (m < total) break;
        m -= total;
    }
    __syncthreads();
    return (uint32_t)*s_out;
}

// stem points (:442-459): the leaf restricted to rows >= int(0.75 H), eroded twice by the 5x5 ellipse
// (rows of the element: 1, 5, 5, 5, 1 wide).  cv2 erodes literally twice; pixels outside the image never constrain.
__device__ __forceinline__ bool se5(int dx, int dy) { return (dy == -2 || dy == 2) ? dx == 0 : true; }

__device__ int build_stem(const Frame& F, uint8_t* e1, uint8_t* e2, int32_t* row_off, int* s_red) {
    const LgRegion& r = F.r;
    const int h75 = (int)(0.75 * (double)F.H);
    const int ya = max(r.y0, h75), yb = r.y1;
    const int bw = r.x1 - r.x0;
    const int area = (yb > ya) ? (yb - ya) * bw : 0;
    for (int i = threadIdx.x; i < area; i += CO_NT) {
        const int y = ya + i / bw, x = r.x0 + i % bw;
        bool keep = true;
        for (int dy = -2; dy <= 2 && keep; ++dy)
            for (int dx = -2; dx <= 2; ++dx) {
                if (!se5(dx, dy)) continue;
                const int xx = x + dx, yy = y + dy;
                if (xx < 0 || yy < 0 || xx >= F.W || yy >= F.H) continue;
                if (yy < h75 || !F.leaf(xx, yy)) { keep = false; break; }
            }
        e1[(size_t)y * F.W + x] = keep;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < area; i += CO_NT) {
        const int y = ya + i / bw, x = r.x0 + i % bw;
        bool keep = true;
        for (int dy = -2; dy <= 2 && keep; ++dy)
            for (int dx = -2; dx <= 2; ++dx) {
                if (!se5(dx, dy)) continue;
                const int xx = x + dx, yy = y + dy;
                if (xx < 0 || yy < 0 || xx >= F.W || yy >= F.H) continue;
                // first-pass values outside the processed rectangle are 0: such a pixel is off the leaf or above h75
                const bool in = xx >= r.x0 && xx < r.x1 && yy >= ya && yy < yb;
                if (!in || !e1[(size_t)yy * F.W + xx]) { keep = false; break; }
            }
        e2[(size_t)y * F.W + x] = keep;
    }
    __syncthreads();
    // per-row counts, then offsets (row_off[y - ya] = survivors in rows before y)
    for (int y = ya + (threadIdx.x >> 5); y < yb; y += CO_WARPS) {
        int cnt = 0;
        for (int x = r.x0 + (threadIdx.x & 31); x < r.x1; x += 32) cnt += e2[(size_t)y * F.W + x];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
        if ((threadIdx.x & 31) == 0) row_off[y - ya + 1] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        row_off[0] = 0;
        for (int y = ya; y < yb; ++y) { acc += row_off[y - ya + 1]; row_off[y - ya + 1] = acc; }
        s_red[CO_WARPS] = acc;
    }
    __syncthreads();
    return (yb > ya) ? s_red[CO_WARPS] : 0;
}

// rank-th stem point in raster order
__device__ uint32_t stem_select(const Frame& F, const uint8_t* e2, const int32_t* row_off, int rank, int* s_red, int* s_out) {
    const LgRegion& r = F.r;
    const int h75 = (int)(0.75 * (double)F.H);
    const int ya = max(r.y0, h75), yb = r.y1;
    for (int y = ya + threadIdx.x; y < yb; y += CO_NT)
        if (row_off[y - ya] <= rank && rank < row_off[y - ya + 1]) s_out[1] = y;
    __syncthreads();
    const int y = s_out[1];
    int m = rank - row_off[y - ya];
    for (int base = r.x0; base < r.x1; base += CO_NT) {
        const int x = base + threadIdx.x;
        const bool on = x < r.x1 && e2[(size_t)y * F.W + x];
        int total;
        const int pos = block_rank(on, s_red, &total);
        if (on && pos == m) *s_out = y * F.W + x;
        if (m < total) break;
        m -= total;
    }
    __syncthreads();
    return (uint32_t)*s_out;
}

// edge points (:461-487): points of the winning outer contour where the border doubles back (see the oracle's
// collector_edge_points for why the reference's angle test reduces to "previous point == next point").  The Moore
// trace of lg_orient.cu runs against cv2's direction: in cv2's order the list is the start (if it qualifies)
// followed by the trace-order list reversed.  One thread; a border is a few thousand steps.
__device__ void build_edges(const Bits& B, int sx, int sy, int32_t* hdr, uint32_t* list, int cap) {
    const int DX[8] = {1, 1, 0, -1, -1, -1, 0, 1};
    const int DY[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    int cx = sx, cy = sy, db = 4, first = -1;
    int n = 0, m = 0, over = 0;
    int p1x = 0, p1y = 0;              // pts[1]
    int ax = 0, ay = 0, bx = 0, by = 0;  // pts[n-2], pts[n-1]
    auto emit = [&](int x, int y) {
        if (m < cap) list[m++] = (uint32_t)x | ((uint32_t)y << 16); else over = 1;
    };
    for (long long guard = 0; guard < (1ll << 26); ++guard) {
        int d = -1;
        for (int k = 1; k <= 8; ++k) {
            const int dd = (db + k) & 7;
            if (B.get(cx + DX[dd], cy + DY[dd])) { d = dd; break; }
        }
        if (d < 0) { n = 1; break; }                       // isolated pixel
        if (cx == sx && cy == sy && first >= 0 && d == first) break;
        if (first < 0) first = d;
        // append (cx, cy) as pts[n]
        if (n == 1) { p1x = cx; p1y = cy; }
        if (n >= 2 && ax == cx && ay == cy) emit(bx, by);  // pts[n-1] sits between two visits of the same pixel
        ax = bx; ay = by; bx = cx; by = cy;
        ++n;
        cx += DX[d]; cy += DY[d];
        db = (d + ((d & 1) ? 5 : 6)) & 7;
    }
    int s0 = 0;
    if (n == 1) s0 = 1;                                    // prev == curr == next
    else if (n == 2) { emit(bx, by); s0 = 1; }
    else if (n >= 3) {
        if (ax == sx && ay == sy) emit(bx, by);            // pts[n-1]: prev pts[n-2], next pts[0]
        s0 = (bx == p1x && by == p1y);                     // pts[0]: prev pts[n-1], next pts[1]
    }
    hdr[HD_S0] = s0; hdr[HD_NEDGE] = m + s0; hdr[HD_EDGE_OVER] = over;
}

__device__ __forceinline__ uint32_t edge_select(const int32_t* hdr, const uint32_t* list, int rank, int ox, int oy, int W) {
    const int s0 = hdr[HD_S0], m = hdr[HD_NEDGE] - s0;
    int lx, ly;
    if (s0 && rank == 0) { lx = hdr[HD_SX]; ly = hdr[HD_SY]; }
    else { const uint32_t e = list[m - 1 - (rank - s0)]; lx = e & 0xFFFF; ly = e >> 16; }
    return (uint32_t)((ly + oy) * W + (lx + ox));
}

struct Sets {
    uint32_t* tip_idx; float* tip_val; uint8_t *e1, *e2; int32_t *hdr, *row_off; uint32_t* edge_list;
};

__device__ Sets frame_sets(const lg_context& c, int b) {
    Sets S;
    S.tip_idx = c.tip_idx_buf + (size_t)b * c.P;
    S.tip_val = reinterpret_cast<float*>(c.tip_val_buf + (size_t)b * c.P);
    uint8_t* e = reinterpret_cast<uint8_t*>(c.dt_fwd);
    S.e1 = e + (size_t)b * c.P;
    S.e2 = e + ((size_t)c.B + b) * c.P;
    S.hdr = c.hull + (size_t)b * (12 * (c.H + 2));
    S.row_off = S.hdr + CO_HDR;
    S.edge_list = reinterpret_cast<uint32_t*>(S.row_off + c.H + 2);
    return S;
}

__device__ Frame frame_of(const lg_context& c, LgMaskSrc src, const float* depth, int b) {
    Frame F;
    F.c = &c; F.src = src; F.depth = depth; F.r = c.region[b];
    F.W = c.W; F.H = c.H; F.fo = (size_t)b * c.P; F.id = src.id(b);
    F.ox = F.r.x0 - 1; F.oy = F.r.y0 - 1;
    F.B.w = c.bits + (size_t)b * c.bits_stride;
    F.B.bw = F.r.x1 - F.r.x0 + 2; F.B.bh = F.r.y1 - F.r.y0 + 2; F.B.wpr = (F.B.bw + 31) >> 5;
    return F;
}

__device__ uint32_t set_select(const Frame& F, const Sets& S, int kind, int rank, int* s_red, int* s_out) {
    if (kind == 0) return tip_select(S.tip_idx, S.tip_val, S.hdr[HD_NTIP], rank, s_red, s_out);
    if (kind == 1) return stem_select(F, S.e2, S.row_off, rank, s_red, s_out);
    return edge_select(S.hdr, S.edge_list, rank, F.ox, F.oy, F.W);
}

__global__ void __launch_bounds__(CO_NT) collect_kernel(lg_context c, LgMaskSrc src, const float* __restrict__ depth,
                                                        unsigned long long seed, unsigned long long first_index,
                                                        const int32_t* __restrict__ grasp_xy, const double* __restrict__ total_in,
                                                        float* __restrict__ patches, lg_sample_meta* __restrict__ meta,
                                                        int32_t* __restrict__ set_sizes) {
    const int b = blockIdx.x, tid = threadIdx.x;
    __shared__ int s_red[CO_WARPS + 1];
    __shared__ int s_out[2];
    __shared__ double s_dred[CO_WARPS];
    lg_sample_meta* mt = meta + (size_t)b * LG_SAMPLES_PER_FRAME;
    float* outp = patches + (size_t)b * LG_SAMPLES_PER_FRAME * PSZ;
    const Sets S = frame_sets(c, b);
    if (tid < LG_SAMPLES_PER_FRAME) {
        lg_sample_meta z;
        z.valid = 0; z.label = 0; z.is_augmented = 0; z.kind = tid; z.x = 0; z.y = 0; z.total_score = 0.0;
        mt[tid] = z;
    }
    if (tid < 3) set_sizes[b * 3 + tid] = 0;
    if (tid < CO_HDR) S.hdr[tid] = 0;
    __syncthreads();
    const Frame F = frame_of(c, src, depth, b);
    if (!F.r.ok) return;
    const lg_frame_result* res = &c.results[b];
    int gx, gy;
    if (grasp_xy) { gx = grasp_xy[2 * b]; gy = grasp_xy[2 * b + 1]; }
    else {
        if (res->n_candidates <= 0) return;               // select_grasp_point returned (None, None, None)
        gx = res->grasp_x; gy = res->grasp_y;
    }
    const int W = F.W, H = F.H;
    // ---- candidate sets (kept in scratch for lg_collector_points)
    const int n_tip = build_tip_list(F, S.tip_idx, S.tip_val, s_red);
    const int q_tip = n_tip > 0 ? max(1, n_tip / 4) : 0;
    const int n_stem = build_stem(F, S.e1, S.e2, S.row_off, s_red);
    const LgOrient* orient = &c.orient[b];
    if (tid == 0) {
        S.hdr[HD_NTIP] = n_tip; S.hdr[HD_QTIP] = q_tip; S.hdr[HD_NSTEM] = n_stem;
        S.hdr[HD_SX] = orient->win_lx; S.hdr[HD_SY] = orient->win_ly;
        const int cap = 12 * (H + 2) - CO_HDR - (H + 2);
        S.hdr[HD_EDGE_CAP] = cap;
        if (orient->has_angle && orient->win_lx >= 0) build_edges(F.B, orient->win_lx, orient->win_ly, S.hdr, S.edge_list, cap);
        S.hdr[HD_READY] = 1;
        __threadfence_block();
    }
    __syncthreads();
    const int n_edge = S.hdr[HD_NEDGE];
    if (tid == 0) { set_sizes[b * 3] = q_tip; set_sizes[b * 3 + 1] = n_stem; set_sizes[b * 3 + 2] = n_edge; }

    // ---- positive sample (:175-240)
    if (gx < 0 || gy < 0 || gx >= W || gy >= H) return;
    if (gy < 16 || gy >= H - 16 || gx < 16 || gx >= W - 16) return;      // _check_boundaries
    if (!extract(F, gx, gy, outp, s_red)) return;
    double total;
    if (total_in) total = total_in[b];
    else {
        // np.max(traditional_score) over the frame (the call site, grasp_point_selector_bkp.py:146-152): the maximum over
        // the score rectangle, and 0.2 = 0.2 * flatness(1.0) wherever the frame extends beyond it (nothing else is
        // non-zero off the leaf and the flatness of empty surroundings is exactly 1)
        double mx = -CUDART_INF;
        const int rw = F.r.sx1 - F.r.sx0, rh = F.r.sy1 - F.r.sy0;
        for (int i = tid; i < rw * rh; i += CO_NT)
            mx = fmax(mx, c.m_trad[F.fo + (size_t)(F.r.sy0 + i / rw) * W + F.r.sx0 + i % rw]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mx = fmax(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, d));
        if ((tid & 31) == 0) s_dred[tid >> 5] = mx;
        __syncthreads();
        mx = s_dred[0];
        for (int w = 1; w < CO_WARPS; ++w) mx = fmax(mx, s_dred[w]);
        if (rw < W || rh < H) mx = fmax(mx, 0.2);
        total = mx;
        __syncthreads();
    }
    if (tid == 0) {
        lg_sample_meta z;
        z.valid = 1; z.label = 1; z.is_augmented = 0; z.kind = 0; z.x = gx; z.y = gy; z.total_score = total;
        mt[0] = z;
    }
    __syncthreads();     // slot 0 is read back below

    // ---- rot90 copies with depth noise and score jitter (:250-299)
    const Rng rng{mix64(seed + 0x9E3779B97F4A7C15ull * (first_index + (unsigned long long)b + 1ull))};
    double dsum = 0.0;
    for (int i = tid; i < PPX; i += CO_NT) dsum += (double)outp[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) dsum += __shfl_xor_sync(0xFFFFFFFFu, dsum, d);
    if ((tid & 31) == 0) s_dred[tid >> 5] = dsum;
    __syncthreads();
    dsum = 0.0;
    for (int w = 0; w < CO_WARPS; ++w) dsum += s_dred[w];
    const float mean = (float)(dsum / (double)PPX);
    for (int k = 1; k <= 3; ++k) {
        float* o = outp + (size_t)k * PSZ;
        const float factor = (float)(0.01 + (0.02 - 0.01) * rng.u01(RNG_NOISE_FACTOR, k, 0));
        const float scale = __fmul_rn(factor, mean);
        for (int i = tid; i < PPX; i += CO_NT) {
            const int oi = i >> 5, oj = i & 31;
            int si, sj;                                   // rot90 by k quarter turns: out[oi][oj] = in[si][sj]
            if (k == 1) { si = oj; sj = 31 - oi; }
            else if (k == 2) { si = 31 - oi; sj = 31 - oj; }
            else { si = 31 - oj; sj = oi; }
            const int s = si * 32 + sj;
            const double u1 = ((double)(rng.draw(RNG_NORMAL, k, 2 * i) >> 11) + 1.0) * 0x1p-53;
            const double u2 = (double)(rng.draw(RNG_NORMAL, k, 2 * i + 1) >> 11) * 0x1p-53;
            const float z = (float)(sqrt(-2.0 * log(u1)) * cos(2.0 * 3.141592653589793 * u2));
            o[i] = fmaxf(__fadd_rn(outp[s], __fmul_rn(z, scale)), 0.f);
            o[PPX + i] = outp[PPX + s] > 0.5f ? 1.f : 0.f;
#pragma unroll
            for (int ch = 2; ch < LG_CHANNELS; ++ch) o[ch * PPX + i] = outp[ch * PPX + s];
        }
        if (tid == 0) {
            lg_sample_meta z;
            z.valid = 1; z.label = 1; z.is_augmented = 1; z.kind = k;
            rotate_point(gx, gy, k, &z.x, &z.y);
            z.total_score = __dmul_rn(total, 0.95 + (1.0 - 0.95) * rng.u01(RNG_SCORE_JITTER, k, 0));
            mt[k] = z;
        }
    }

    // ---- negatives (:301-348): up to 3, at most 10 rounds over the non-empty sets
    const int sizes[3] = {q_tip, n_stem, n_edge};
    int collected = 0;
    for (int attempt = 0; attempt < 10 && collected < 3; ++attempt) {
        for (int kind = 0; kind < 3; ++kind) {
            if (sizes[kind] == 0 || collected >= 3) continue;
            const int rank = (int)(rng.draw(RNG_PICK, attempt, kind) % (unsigned long long)sizes[kind]);
            const uint32_t p = set_select(F, S, kind, rank, s_red, s_out);
            const int px = (int)(p % (uint32_t)W), py = (int)(p / (uint32_t)W);
            const int slot = 4 + collected;
            if (!extract(F, px, py, outp + (size_t)slot * PSZ, s_red)) continue;
            if (tid == 0) {
                lg_sample_meta z;
                z.valid = 1; z.label = 0; z.is_augmented = 0; z.kind = 4 + kind; z.x = px; z.y = py; z.total_score = 0.0;
                mt[slot] = z;
            }
            ++collected;
        }
    }
}

// points of the candidate sets by rank (parity tests; any training code that wants its own sampling)
__global__ void __launch_bounds__(CO_NT) collector_points_kernel(lg_context c, LgMaskSrc src, int kind, const uint32_t* __restrict__ ranks,
                                                                 int nq, int32_t* __restrict__ xy) {
    const int b = blockIdx.x;
    __shared__ int s_red[CO_WARPS + 1];
    __shared__ int s_out[2];
    const Sets S = frame_sets(c, b);
    const Frame F = frame_of(c, src, nullptr, b);
    const int size = !S.hdr[HD_READY] ? 0 : kind == 0 ? S.hdr[HD_QTIP] : kind == 1 ? S.hdr[HD_NSTEM] : S.hdr[HD_NEDGE];
    for (int q = 0; q < nq; ++q) {
        int x = -1, y = -1;
        if (size > 0) {
            const uint32_t p = set_select(F, S, kind, (int)(ranks[(size_t)b * nq + q] % (uint32_t)size), s_red, s_out);
            x = (int)(p % (uint32_t)c.W); y = (int)(p / (uint32_t)c.W);
        }
        if (threadIdx.x == 0) { xy[((size_t)b * nq + q) * 2] = x; xy[((size_t)b * nq + q) * 2 + 1] = y; }
        __syncthreads();
    }
}

}  // namespace

int lg_context_dev_alloc(lg_context* c, void** p, size_t bytes);
// The tip lists (12 bytes per pixel of capacity) belong to the collector alone: its first call allocates them, so that a
// context that only selects grasps does not carry them.
static int ensure_collector_scratch(lg_context* c) {
    if (c->tip_val_buf && c->tip_idx_buf) return LG_OK;
    void* p = nullptr;
    int rc = lg_context_dev_alloc(c, &p, (size_t)c->B * c->P * sizeof(double));
    if (rc) return rc;
    c->tip_val_buf = static_cast<double*>(p);
    rc = lg_context_dev_alloc(c, &p, (size_t)c->B * c->P * sizeof(uint32_t));
    if (rc) return rc;
    c->tip_idx_buf = static_cast<uint32_t*>(p);
    return LG_OK;
}

int lg_run_collect(lg_context* c, LgMaskSrc src, const float* depth, int n, unsigned long long seed, unsigned long long first_index,
                   const int32_t* grasp_xy, const double* total, float* patches, lg_sample_meta* meta, int32_t* set_sizes,
                   cudaStream_t st) {
    { const int rc = ensure_collector_scratch(c); if (rc) return rc; }
    collect_kernel<<<n, CO_NT, 0, st>>>(*c, src, depth, seed, first_index, grasp_xy, total, patches, meta, set_sizes);
    LG_LAUNCH_CHECK();
    return LG_OK;
}

int lg_run_collector_points(lg_context* c, LgMaskSrc src, int n, int kind, const uint32_t* ranks, int nq, int32_t* xy, cudaStream_t st) {
    { const int rc = ensure_collector_scratch(c); if (rc) return rc; }
    collector_points_kernel<<<n, CO_NT, 0, st>>>(*c, src, kind, ranks, nq, xy);
    LG_LAUNCH_CHECK();
    return LG_OK;
}
