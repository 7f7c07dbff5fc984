// Stage 2 - per-pixel grasp scoring, candidates, patches and the CV/ML fusion
// (reference scripts/utils/grasp_point_selector.py).  Compile with -fmad=false: the reference combines
// float64 / float32 terms with separately rounded products and sums, and candidate indices only stay
// identical if this file does the same.
//
//   score_kernel   one pass over the leaf rectangle: masked depth tile -> 5x5 Gaussian -> Sobel ->
//                  flatness; closed-form approach / accessibility; sdf_score from the chamfer field;
//                  stem penalty from the leaf bitmask; traditional score; valid mask     (:256-288, 502-701)
//   nms_tiles_kernel / nms_kernel   top-20 greedy pick with the +-10 px mark: best alive key per 32 x 8 tile,
//                  20 rounds of (arg-max over the tiles, refresh of the tiles under the pick)   (:447-482)
//   gather_kernel  9 x 32 x 32 patch tensor written in the CNN's input layout    (:59-127, 392-445)
//   fuse_kernel    ML rescale + confidence weighting + 3-D / pre-grasp points    (:133-136, 205-249, 152-180, 754-826)
#include <math_constants.h>

#include <cuda_bf16.h>

#include "lg_internal.cuh"

namespace {

constexpr int SC_TW = 32, SC_TH = 8;

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

// any leaf pixel on local bitmask row ly within local columns [xa, xb) ?
__device__ __forceinline__ bool bits_any(const uint32_t* bits, int wpr, int bw, int bh, int ly, int xa, int xb) {
    if (ly < 0 || ly >= bh) return false;
    xa = max(xa, 0); xb = min(xb, bw);
    if (xa >= xb) return false;
    const uint32_t* row = bits + ly * wpr;
    int wa = xa >> 5, wb = (xb - 1) >> 5;
    for (int wi = wa; wi <= wb; ++wi) {
        uint32_t m = row[wi];
        if (wi == wa) m &= 0xFFFFFFFFu << (xa & 31);
        if (wi == wb) { int hi = xb - (wb << 5); if (hi < 32) m &= (1u << hi) - 1u; }
        if (m) return true;
    }
    return false;
}

__device__ __forceinline__ float exp32(float x) { return (float)exp((double)x); }

// grid = (CTAs per frame, frames): a CTA walks the 32 x 8 tiles of its frame's score rectangle (the rectangle is
// only known on the device, so the grid cannot be sized to it; a grid over the whole frame would launch ~10x more
// CTAs than there is work).
__global__ void __launch_bounds__(SC_TW * SC_TH) score_kernel(lg_context c, LgMaskSrc src, const float* __restrict__ depth,
                                                               lg_camera cam, int full, double* iso_out) {
    const int b = blockIdx.y;
    const LgRegion r = c.region[b];
    if (!r.ok && !full) return;
    const int W = c.W, H = c.H;
    const int tiles_x = (r.sx1 - r.sx0 + SC_TW - 1) / SC_TW, tiles_y = (r.sy1 - r.sy0 + SC_TH - 1) / SC_TH;
    __shared__ float zt[SC_TH + 6][SC_TW + 6];
    __shared__ float st[SC_TH + 2][SC_TW + 2];
    const int tid = threadIdx.x;
    const size_t fo = (size_t)b * c.P;
    const int id = src.id(b);
    for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
    const int tx0 = r.sx0 + (tile % tiles_x) * SC_TW, ty0 = r.sy0 + (tile / tiles_x) * SC_TH;
    __syncthreads();     // the previous tile's readers of zt / st are done
    // masked depth tile with reflect-101 borders (image_processor.py:60-61); at most three turns per thread, unrolled so
    // that their label and depth loads are all in flight together
#pragma unroll
    for (int i = tid; i < (SC_TH + 6) * (SC_TW + 6); i += SC_TW * SC_TH) {
        const int ly = i / (SC_TW + 6), lx = i - ly * (SC_TW + 6);
        const int y = reflect101(ty0 - 3 + ly, H), x = reflect101(tx0 - 3 + lx, W);
        const size_t p = (size_t)y * W + x;
        const float m = src.at(fo, p, id) ? 1.f : 0.f;
        zt[ly][lx] = __fmul_rn(depth[fo + p], m);
    }
    __syncthreads();
    // Gaussian 5x5 at in-image positions of the (tile + 1) ring
    for (int i = tid; i < (SC_TH + 2) * (SC_TW + 2); i += SC_TW * SC_TH) {
        const int ly = i / (SC_TW + 2), lx = i - ly * (SC_TW + 2);
        const int y = ty0 - 1 + ly, x = tx0 - 1 + lx;
        if (y >= 0 && y < H && x >= 0 && x < W) {
            float acc = __fmul_rn(c.gauss[0], zt[ly][lx]);
#pragma unroll
            for (int k = 1; k < 25; ++k) acc = __fadd_rn(acc, __fmul_rn(c.gauss[k], zt[ly + k / 5][lx + k % 5]));
            st[ly][lx] = acc;
        }
    }
    __syncthreads();
    // reflect-101 padding of the smoothed image (grasp_point_selector.py:648)
    for (int i = tid; i < (SC_TH + 2) * (SC_TW + 2); i += SC_TW * SC_TH) {
        const int ly = i / (SC_TW + 2), lx = i - ly * (SC_TW + 2);
        const int y = ty0 - 1 + ly, x = tx0 - 1 + lx;
        if (!(y >= 0 && y < H && x >= 0 && x < W)) {
            const int ry = reflect101(y, H) - (ty0 - 1), rx = reflect101(x, W) - (tx0 - 1);
            float v = 0.f;
            if (ry >= 0 && ry < SC_TH + 2 && rx >= 0 && rx < SC_TW + 2) v = st[ry][rx];
            st[ly][lx] = v;
        }
    }
    __syncthreads();
    const int lx = tid % SC_TW, ly = tid / SC_TW;
    const int x = tx0 + lx, y = ty0 + ly;
    double trad = 0.0;
    if (x < r.sx1 && y < r.sy1) {
        const size_t p = (size_t)y * W + x;
        const bool M = src.at(fo, p, id);
        // flatness (:650-655)
        const float s00 = st[ly][lx], s01 = st[ly][lx + 1], s02 = st[ly][lx + 2];
        const float s10 = st[ly + 1][lx], s12 = st[ly + 1][lx + 2];
        const float s20 = st[ly + 2][lx], s21 = st[ly + 2][lx + 1], s22 = st[ly + 2][lx + 2];
        float gx = -s00;
        gx = __fadd_rn(gx, s02);
        gx = __fadd_rn(gx, __fmul_rn(-2.f, s10));
        gx = __fadd_rn(gx, __fmul_rn(2.f, s12));
        gx = __fadd_rn(gx, -s20);
        gx = __fadd_rn(gx, s22);
        float gy = -s00;
        gy = __fadd_rn(gy, __fmul_rn(-2.f, s01));
        gy = __fadd_rn(gy, -s02);
        gy = __fadd_rn(gy, s20);
        gy = __fadd_rn(gy, __fmul_rn(2.f, s21));
        gy = __fadd_rn(gy, s22);
        const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
        const float flat = exp32(__fmul_rn(-mag, 5.f));
        double sdf_score = 0.0, approach = 0.0, access = 0.0;
        float stem = 0.f;
        float di = 0.f;
        if (M) {
            di = c.di[fo + p];
            const uint32_t* mx = c.dt_max + (size_t)b * 2;
            const float sdf_max = __fmul_rn((float)max(mx[0], mx[1]), 1.0f / 65536.0f);
            const LgOrient o = c.orient[b];
            // sdf_score (:526-567)
            const float dm = __fsub_rn(di, 20.f);
            const float interior = exp32(__fdiv_rn(-__fmul_rn(dm, dm), 800.f));
            const float sdf = __fdiv_rn(di, sdf_max);
            const double dx = (double)x - cam.cx, dy = (double)y - cam.cy;
            const double r2 = dx * dx + dy * dy;
            double nrm = sqrt(r2);
            const double rr = nrm;
            if (nrm == 0.0) nrm = 1.0;
            double align = 1.0;
            if (o.has_angle) align = fabs((dx / nrm) * o.sin_a - (dy / nrm) * o.cos_a);
            sdf_score = ((double)__fmul_rn(0.4f, interior) + 0.4 * align) + (double)__fmul_rn(0.2f, sdf);
            // approach (:569-593)
            approach = fabs(cam.f / sqrt(r2 + cam.f * cam.f));
            // accessibility (:502-524)
            const double diag = sqrt((double)((long long)W * W + (long long)H * H));
            const double fwd = (rr == 0.0) ? 1.0 : dx / rr;
            access = 0.7 * (1.0 - rr / diag) + 0.3 * fwd;
            // stem penalty (:688-701): dilation of (leaf AND bottom third) by the 30x30 ellipse, AND leaf
            const int h3 = H - H / 3;
            if (y + (LG_SE_STEM - 1 - LG_SE_STEM / 2) >= h3) {
                const uint32_t* bits = c.bits + (size_t)b * c.bits_stride;
                const int ox = r.x0 - 1, oy = r.y0 - 1;
                const int bw = r.x1 - r.x0 + 2, bh = r.y1 - r.y0 + 2, wpr = (bw + 31) >> 5;
                for (int j = 0; j < LG_SE_STEM; ++j) {
                    const int yy = y + j - LG_SE_STEM / 2;
                    if (yy < h3 || yy >= H) continue;
                    if (bits_any(bits, wpr, bw, bh, yy - oy, x + c.se30_a[j] - LG_SE_STEM / 2 - ox,
                                 x + c.se30_b[j] - LG_SE_STEM / 2 - ox)) { stem = 1.f; break; }
                }
            }
        }
        // combination (:272-277); evaluated for every pixel like the reference (flatness is not masked)
        // (weights: 0.4 / 0.3 / 0.2 / 0.1 unless lg_set_score_weights changed them; the flatness term is a float32 product
        // in the reference, float32 map times Python scalar)
        trad = (((c.w_trad[0] * approach + c.w_trad[1] * sdf_score) + (double)__fmul_rn((float)c.w_trad[2], flat)) + c.w_trad[3] * access) *
               (double)__fsub_rn(1.f, stem);
        const bool valid = (di > 20.f) && M && (stem < 0.8f);
        c.m_sdf[fo + p] = sdf_score; c.m_app[fo + p] = approach; c.m_acc[fo + p] = access; c.m_trad[fo + p] = trad;
        c.m_flat[fo + p] = flat; c.m_stem[fo + p] = stem; c.m_valid[fo + p] = valid ? 1 : 0;
        if (iso_out) {
            double iso = 0.0;
            if (M) iso = (y == H - 1) ? 0.2 : ((double)y * ((0.2 - 1.0) / (double)(H - 1)) + 1.0);
            iso_out[fo + p] = iso;
        }
    }
    }   // tiles
}

// fill the per-frame maps with their analytic values outside the score rectangle (full mode only needs
// nothing: the rectangle is the frame).  Not needed in region mode: consumers special-case the outside.

// ---------------------------------------------------------------------------------------------------
// candidates
// ---------------------------------------------------------------------------------------------------
// _get_candidate_points (grasp_point_selector.py:447-482): 20 rounds of (arg-max of the remaining keys, suppress
// everything within +-20 px of the pick).  Whether a pixel is still in the running depends on the picks alone - it is alive
// iff its key is positive and no earlier pick lies within LG_NMS_REACH of it (Chebyshev) - so no list of keys is edited or
// rebuilt.  The score rectangle is cut into tiles of 32 x 8 pixels; a table holds every tile's best alive key with its
// pixel (tile_key / tile_id).  A round is then an arg-max over the table (a few hundred entries in shared memory) and a
// fresh look at the at most 18 tiles the new pick's 41 x 41 window touches, one warp per tile: 20 rounds cost 40 barriers
// and 20 trips to the L2, whatever the number of positive keys.  Keys are compared as the bit patterns of positive doubles
// (same order); among equal keys the larger flat index wins, as in the sequential search over the index-ordered list.
constexpr int NMS_NT = 640;               // 20 warps: one per tile a pick can touch (3 x 6) and a little spare
constexpr int NMS_NW = NMS_NT / 32;
constexpr int NMS_TW = 32, NMS_TH = 8;
constexpr int NMS_TCACHE = 3072;          // table entries kept in shared memory (a 1440 x 1080 frame has 6075 tiles, a leaf ~300)
constexpr int NMS_INIT_NT = 256;

struct NmsGeom {
    int rx0, ry0, rx1, ry1;               // rectangle that carries keys (exclusive upper bounds)
    int ax0, ay0;                         // origin of the tile grid: the rectangle's corner rounded down to the tile size
    int tx, ty;                           // tiles per row / column
};
__device__ __forceinline__ NmsGeom nms_geom(const LgRegion& r, bool whole_frame, int W, int H) {
    NmsGeom g;
    g.rx0 = whole_frame ? 0 : r.sx0; g.ry0 = whole_frame ? 0 : r.sy0;
    g.rx1 = whole_frame ? W : r.sx1; g.ry1 = whole_frame ? H : r.sy1;
    g.ax0 = g.rx0 & ~(NMS_TW - 1); g.ay0 = g.ry0 & ~(NMS_TH - 1);
    g.tx = max(0, (g.rx1 - g.ax0 + NMS_TW - 1) / NMS_TW); g.ty = max(0, (g.ry1 - g.ay0 + NMS_TH - 1) / NMS_TH);
    return g;
}

// lexicographic maximum of (key, index) over the warp; every lane gets the result
__device__ __forceinline__ void nms_warp_max(unsigned long long& k, unsigned& i) {
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mh = __reduce_max_sync(0xFFFFFFFFu, hi);
    const unsigned ml = __reduce_max_sync(0xFFFFFFFFu, hi == mh ? lo : 0u);
    const unsigned mi = __reduce_max_sync(0xFFFFFFFFu, (hi == mh && lo == ml) ? i : 0u);
    k = ((unsigned long long)mh << 32) | ml;
    i = mi;
}

// Best alive key of one tile (all lanes call; lane = column, 8 rows per lane).  The picks live in the lanes: lane q < npick
// holds pick q in (pqx, pqy).  key == nullptr never happens; valid marks the pixels that may be picked at all.
__device__ __forceinline__ void nms_tile_best(const double* __restrict__ key, const uint8_t* __restrict__ valid, int W,
                                              const NmsGeom& g, int tile, int pqx, int pqy, int npick, int lane,
                                              unsigned long long& bk, unsigned& bi) {
    const int tcy = tile / g.tx, tcx = tile - tcy * g.tx;
    const int X0 = g.ax0 + tcx * NMS_TW, Y0 = g.ay0 + tcy * NMS_TH;
    const int x = X0 + lane;
    // the picks whose window reaches this tile
    const bool reach = lane < npick && pqx + LG_NMS_REACH >= X0 && pqx - LG_NMS_REACH <= X0 + NMS_TW - 1 &&
                       pqy + LG_NMS_REACH >= Y0 && pqy - LG_NMS_REACH <= Y0 + NMS_TH - 1;
    unsigned rel = __ballot_sync(0xFFFFFFFFu, reach);
    const bool xin = x >= g.rx0 && x < g.rx1;
    double kv[NMS_TH];
    unsigned ok = 0;                      // bit j: row j of this column may be picked
#pragma unroll
    for (int j = 0; j < NMS_TH; ++j) {    // all loads of the tile are in flight together
        const int y = Y0 + j;
        const bool in = xin && y >= g.ry0 && y < g.ry1;
        const size_t p = (size_t)y * W + x;
        kv[j] = in ? key[p] : 0.0;
        if (in && valid[p]) ok |= 1u << j;
    }
    while (rel) {
        const int q = __ffs(rel) - 1;
        rel &= rel - 1;
        const int qx = __shfl_sync(0xFFFFFFFFu, pqx, q), qy = __shfl_sync(0xFFFFFFFFu, pqy, q);
        const int jlo = max(qy - LG_NMS_REACH - Y0, 0), jhi = min(qy + LG_NMS_REACH - Y0, NMS_TH - 1);
        if (abs(x - qx) <= LG_NMS_REACH && jlo <= jhi) ok &= ~(((1u << (jhi - jlo + 1)) - 1u) << jlo);
    }
    bk = 0ull;
    int bj = 0;
#pragma unroll
    for (int j = 0; j < NMS_TH; ++j) {
        const bool pos = kv[j] > 0.0;     // false for NaN
        const unsigned long long kb = (unsigned long long)__double_as_longlong(kv[j]);
        if (((ok >> j) & 1u) && pos && kb >= bk) { bk = kb; bj = j; }      // a later row is a larger index: it wins a tie
    }
    bi = bk ? (unsigned)((size_t)(Y0 + bj) * W + x) : 0u;
    nms_warp_max(bk, bi);
}

// the table of every frame before the first pick: a warp per tile.  ext_score / ext_valid: caller-supplied full-frame
// maps (lg_candidate_points); otherwise the traditional score and the valid mask on the frame's score rectangle.
__global__ void __launch_bounds__(NMS_INIT_NT) nms_tiles_kernel(lg_context c, const double* ext_score, const uint8_t* ext_valid) {
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const LgRegion r = c.region[b];
    if (!ext_score && !r.ok) return;
    const NmsGeom g = nms_geom(r, ext_score != nullptr, c.W, c.H);
    const int T = g.tx * g.ty;
    const size_t fo = (size_t)b * c.P;
    const double* key = ext_score ? ext_score + fo : c.m_trad + fo;
    const uint8_t* valid = ext_valid ? ext_valid + fo : c.m_valid + fo;
    unsigned long long* tk = c.tile_key + (size_t)b * c.tile_cap;
    uint32_t* ti = c.tile_id + (size_t)b * c.tile_cap;
    for (int tile = blockIdx.x * (NMS_INIT_NT / 32) + warp; tile < T; tile += gridDim.x * (NMS_INIT_NT / 32)) {
        unsigned long long bk;
        unsigned bi;
        nms_tile_best(key, valid, c.W, g, tile, 0, 0, 0, lane, bk, bi);
        if (lane == 0) { tk[tile] = bk; ti[tile] = bi; }
    }
}

__global__ void __launch_bounds__(NMS_NT, 2) nms_kernel(lg_context c, const double* ext_score, const uint8_t* ext_valid,
                                                      int32_t* ext_xy, int32_t* ext_count) {
    __shared__ unsigned long long s_tk[NMS_TCACHE];
    __shared__ unsigned s_ti[NMS_TCACHE];
    __shared__ unsigned long long wk[2][NMS_NW];
    __shared__ unsigned wi[2][NMS_NW];
    __shared__ int px[LG_TOP_K], py[LG_TOP_K];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = c.W;
    const size_t fo = (size_t)b * c.P;
    lg_frame_result* res = &c.results[b];
    const LgRegion r = c.region[b];
    if (!ext_score && !r.ok) {
        if (tid == 0) { res->n_candidates = 0; res->n_positive = 0; }
        return;
    }
    const NmsGeom g = nms_geom(r, ext_score != nullptr, W, c.H);
    const int T = g.tx * g.ty;
    const double* key = ext_score ? ext_score + fo : c.m_trad + fo;
    const uint8_t* valid = ext_valid ? ext_valid + fo : c.m_valid + fo;
    unsigned long long* tk = c.tile_key + (size_t)b * c.tile_cap;
    unsigned* ti = c.tile_id + (size_t)b * c.tile_cap;
    if (T <= NMS_TCACHE) {                 // the usual case: the table lives in shared memory from here on
        for (int t = tid; t < T; t += NMS_NT) { s_tk[t] = tk[t]; s_ti[t] = ti[t]; }
        tk = s_tk; ti = s_ti;
    }
    __syncthreads();
    int cnt = 0;
    int pqx = 0, pqy = 0;                  // lane q of every warp holds pick q
    for (int it = 0; it < LG_TOP_K; ++it) {
        unsigned long long bk = 0ull;
        unsigned bi = 0u;
        for (int t = tid; t < T; t += NMS_NT) {
            const unsigned long long k = tk[t];
            const unsigned i = ti[t];
            if (k > bk || (k == bk && i > bi)) { bk = k; bi = i; }
        }
        nms_warp_max(bk, bi);
        const int par = it & 1;
        if (lane == 0) { wk[par][warp] = bk; wi[par][warp] = bi; }
        __syncthreads();
        bk = lane < NMS_NW ? wk[par][lane] : 0ull;       // every warp reduces the warps' winners: no second broadcast
        bi = lane < NMS_NW ? wi[par][lane] : 0u;
        nms_warp_max(bk, bi);
        if (bk == 0ull) break;                           // no alive key left (uniform)
        const int qx = (int)(bi % (unsigned)W), qy = (int)(bi / (unsigned)W);
        if (lane == it) { pqx = qx; pqy = qy; }
        if (tid == 0) {
            px[it] = qx; py[it] = qy;
            res->cand_x[it] = qx; res->cand_y[it] = qy; res->trad[it] = __longlong_as_double((long long)bk);
        }
        ++cnt;
        if (it == LG_TOP_K - 1) break;
        // the tiles the pick's window touches get their best alive key again
        const int c0 = (max(qx - LG_NMS_REACH, g.rx0) - g.ax0) / NMS_TW, c1 = (min(qx + LG_NMS_REACH, g.rx1 - 1) - g.ax0) / NMS_TW;
        const int r0 = (max(qy - LG_NMS_REACH, g.ry0) - g.ay0) / NMS_TH, r1 = (min(qy + LG_NMS_REACH, g.ry1 - 1) - g.ay0) / NMS_TH;
        const int ncx = c1 - c0 + 1, na = ncx * (r1 - r0 + 1);
        for (int a = warp; a < na; a += NMS_NW) {
            const int tile = (r0 + a / ncx) * g.tx + c0 + a % ncx;
            // A tile keeps its entry when it has no alive key anyway or when its best pixel lies outside the new window:
            // a pick only removes pixels, so a maximum that survives stays the maximum.
            const unsigned long long ok_ = tk[tile];
            const unsigned oi = ti[tile];
            if (ok_ == 0ull || abs((int)(oi % (unsigned)W) - qx) > LG_NMS_REACH || abs((int)(oi / (unsigned)W) - qy) > LG_NMS_REACH) continue;
            unsigned long long nk;
            unsigned ni;
            nms_tile_best(key, valid, W, g, tile, pqx, pqy, cnt, lane, nk, ni);
            if (lane == 0) { tk[tile] = nk; ti[tile] = ni; }
        }
        __syncthreads();
    }
    __syncthreads();
    if (warp == 0) {
        const int n_pos = cnt;
        // zero-key fill: every positive key is picked or suppressed by now, so the remaining picks are the non-suppressed
        // pixels in descending flat index (oracle: candidate_points).  Lane q tests pick q; the first pick that hits
        // decides where the walk continues, as in the sequential loop.
        long long i = (long long)c.P - 1;
        while (cnt < LG_TOP_K && i >= 0) {
            const int x = (int)(i % W), y = (int)(i / W);
            const unsigned hits = __ballot_sync(0xFFFFFFFFu, lane < cnt && abs(x - pqx) <= LG_NMS_REACH && abs(y - pqy) <= LG_NMS_REACH);
            if (hits) {
                const int nx = __shfl_sync(0xFFFFFFFFu, pqx, __ffs(hits) - 1) - LG_NMS_REACH - 1;
                i = (nx >= 0) ? (long long)y * W + nx : (long long)y * W - 1;
                continue;
            }
            if (lane == cnt) { pqx = x; pqy = y; }
            if (lane == 0) {
                px[cnt] = x; py[cnt] = y;
                res->cand_x[cnt] = x; res->cand_y[cnt] = y;
                double t;
                if (ext_score) t = ext_score[fo + i];
                else if (x >= r.sx0 && x < r.sx1 && y >= r.sy0 && y < r.sy1) t = c.m_trad[fo + i];
                else t = (double)0.2f;   // outside the leaf rectangle only the (unmasked) flatness term is left, and it is 1
                res->trad[cnt] = t;
            }
            ++cnt;
            --i;
        }
        if (lane == 0) {
            res->n_candidates = cnt;
            res->n_positive = n_pos;
            if (cnt == 0) atomicOr(&c.status[b], LG_ST_NO_CANDIDATE);
            if (ext_xy) {
                for (int k = 0; k < LG_TOP_K; ++k) {
                    ext_xy[(b * LG_TOP_K + k) * 2] = k < cnt ? px[k] : -1;
                    ext_xy[(b * LG_TOP_K + k) * 2 + 1] = k < cnt ? py[k] : -1;
                }
                ext_count[b] = cnt;
            }
        }
    }
}

__global__ void clear_status_kernel(lg_context c, int n) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n) c.status[b] = 0;
}

// ---------------------------------------------------------------------------------------------------
// patches
// ---------------------------------------------------------------------------------------------------
// Which of the n x 20 candidate slots get an ML score (the reference only scores windows that need no padding:
// bool replicate-pad raises, grasp_point_selector.py:425-437), and where their patch goes: valid slots are numbered
// consecutively so that the CNN runs on exactly that many patches (c.slot_map, c.cnn_count).  One CTA, block scan.
constexpr int CS_NT = 1024;
__global__ void __launch_bounds__(CS_NT) compact_slots_kernel(lg_context c, int n) {
    __shared__ int s_warp[CS_NT / 32];
    __shared__ int s_carry;
    const int total = n * LG_TOP_K, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = c.W, H = c.H;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < total; base += CS_NT) {
        const int i = base + tid;
        int ok = 0;
        if (i < total) {
            const int b = i / LG_TOP_K, k = i - b * LG_TOP_K;
            lg_frame_result* res = &c.results[b];
            const int ncand = c.region[b].ok ? res->n_candidates : 0;
            if (k < ncand) {
                const int cx = res->cand_x[k], cy = res->cand_y[k];
                ok = (cx - 16 >= 0 && cy - 16 >= 0 && cx + 16 <= W && cy + 16 <= H) ? 1 : 0;
            }
            res->ml_valid[k] = ok;
        }
        int incl = ok;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += s_warp[w];
        const int carry = s_carry;
        if (i < total) c.slot_map[i] = ok ? carry + woff + incl - 1 : -1;
        __syncthreads();
        if (tid == CS_NT - 1) s_carry = carry + woff + incl;
        __syncthreads();
    }
    if (tid == 0) *c.cnn_count = s_carry;
}

constexpr int GA_NT = 256;
// PACKED = false: float32 [slot][9][32][32] into c.patches (the fp32 CNN, lg_patches).  PACKED = true: straight into the
// input layout of the tensor-core convolutions (lg_cnn_bf16.cu: two planes of 16-byte units - channels 0-7 | channel 8 and
// zeros - at position lead + slot * 33 * 33 + (y + 1) * 33 + (x + 1), row 0 / column 0 of every patch zero), so that no
// fp32 patch tensor is written and re-read on the way to the CNN.
template <bool PACKED>
__global__ void __launch_bounds__(GA_NT) gather_kernel(lg_context c, LgMaskSrc src, const float* __restrict__ depth,
                                                        uint4* __restrict__ packed, long long plane_rows, int lead) {
    const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    lg_frame_result* res = &c.results[b];
    const int slot = c.slot_map[b * LG_TOP_K + k];
    if (slot < 0) return;                      // no ML score for this candidate: nothing to build
    float* out = c.patches + (size_t)slot * (LG_CHANNELS * LG_PATCH * LG_PATCH);
    constexpr int PITCH = LG_PATCH + 1, PP = PITCH * PITCH;
    uint4* pk = PACKED ? packed + lead + (long long)slot * PP : nullptr;
    if (PACKED) {
        // the zero halo of this patch (row 0, column 0), the PITCH + 1 positions behind it (halo of the next patch, or the
        // tail the convolutions read past the last one), and - by the first patch - the lead rows
        const uint4 z = make_uint4(0, 0, 0, 0);
        long long pos = -1;
        if (tid < PITCH) pos = tid;
        else if (tid < PITCH + LG_PATCH) pos = (long long)(tid - PITCH + 1) * PITCH;
        else if (tid < 2 * PITCH + LG_PATCH + 1) pos = PP + (tid - PITCH - LG_PATCH);
        if (pos >= 0) { pk[pos] = z; pk[plane_rows + pos] = z; }
        if (slot == 0)
            for (int i = tid; i < lead; i += GA_NT) { packed[i] = z; packed[plane_rows + i] = z; }
    }
    const LgRegion r = c.region[b];
    const int W = c.W, H = c.H;
    __shared__ float smn[GA_NT / 32][LG_CHANNELS], smx[GA_NT / 32][LG_CHANNELS];
    __shared__ float fmn[LG_CHANNELS], fmx[LG_CHANNELS];
    const int cx = res->cand_x[k], cy = res->cand_y[k];
    const size_t fo = (size_t)b * c.P;
    const int id = src.id(b);
    float v[LG_CHANNELS][4];
    float mn[LG_CHANNELS], mx[LG_CHANNELS];
#pragma unroll
    for (int ch = 0; ch < LG_CHANNELS; ++ch) { mn[ch] = CUDART_INF_F; mx[ch] = -CUDART_INF_F; }
    const double step = (0.2 - 1.0) / (double)(H - 1);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int pi = tid + q * GA_NT;
        const int y = min(max(cy - 16 + (pi >> 5), 0), H - 1), x = min(max(cx - 16 + (pi & 31), 0), W - 1);
        const size_t p = (size_t)y * W + x;
        const bool M = src.at(fo, p, id);
        const bool in = x >= r.sx0 && x < r.sx1 && y >= r.sy0 && y < r.sy1;
        v[0][q] = depth[fo + p];
        v[1][q] = M ? 1.f : 0.f;
        v[2][q] = in ? (float)c.m_sdf[fo + p] : 0.f;
        v[3][q] = in ? (float)c.m_app[fo + p] : 0.f;
        v[4][q] = in ? c.m_flat[fo + p] : 1.f;
        v[5][q] = M ? (float)((y == H - 1) ? 0.2 : ((double)y * step + 1.0)) : 0.f;
        v[6][q] = M ? c.di[fo + p] : 0.f;
        v[7][q] = in ? (float)c.m_acc[fo + p] : 0.f;
        v[8][q] = in ? c.m_stem[fo + p] : 0.f;
#pragma unroll
        for (int ch = 0; ch < LG_CHANNELS; ++ch) { mn[ch] = fminf(mn[ch], v[ch][q]); mx[ch] = fmaxf(mx[ch], v[ch][q]); }
    }
#pragma unroll
    for (int ch = 0; ch < LG_CHANNELS; ++ch) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mn[ch] = fminf(mn[ch], __shfl_xor_sync(0xFFFFFFFFu, mn[ch], d));
            mx[ch] = fmaxf(mx[ch], __shfl_xor_sync(0xFFFFFFFFu, mx[ch], d));
        }
        if ((tid & 31) == 0) { smn[tid >> 5][ch] = mn[ch]; smx[tid >> 5][ch] = mx[ch]; }
    }
    __syncthreads();
    if (tid < LG_CHANNELS) {
        float a = smn[0][tid], z = smx[0][tid];
        for (int w = 1; w < GA_NT / 32; ++w) { a = fminf(a, smn[w][tid]); z = fmaxf(z, smx[w][tid]); }
        fmn[tid] = a; fmx[tid] = z;
    }
    __syncthreads();
#pragma unroll
    for (int ch = 0; ch < LG_CHANNELS; ++ch) {
        const float lo = fmn[ch], hi = fmx[ch];
        const bool norm = (ch != 1) && (hi > lo);
        const float den = __fsub_rn(hi, lo);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float val = v[ch][q];
            if (norm) val = __fdiv_rn(__fsub_rn(val, lo), den);
            if (PACKED) v[ch][q] = val;
            else out[ch * (LG_PATCH * LG_PATCH) + tid + q * GA_NT] = val;
        }
    }
    if (PACKED) {
        auto pack2 = [](float a, float b2) -> unsigned { __nv_bfloat162 h = __floats2bfloat162_rn(a, b2); return *reinterpret_cast<unsigned*>(&h); };
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int pi = tid + q * GA_NT;
            const long long pos = (long long)((pi >> 5) + 1) * PITCH + (pi & 31) + 1;
            pk[pos] = make_uint4(pack2(v[0][q], v[1][q]), pack2(v[2][q], v[3][q]), pack2(v[4][q], v[5][q]), pack2(v[6][q], v[7][q]));
            pk[plane_rows + pos] = make_uint4(pack2(v[8][q], 0.f), 0u, 0u, 0u);
        }
    }
}

__global__ void export_patches_kernel(lg_context c, float* __restrict__ out) {
    const int k = blockIdx.x, b = blockIdx.y;
    const int slot = c.slot_map[b * LG_TOP_K + k];
    const size_t sz = LG_CHANNELS * LG_PATCH * LG_PATCH;
    float* o = out + ((size_t)b * LG_TOP_K + k) * sz;
    const float* src = slot >= 0 ? c.patches + (size_t)slot * sz : nullptr;
    for (int i = threadIdx.x; i < (int)sz; i += blockDim.x) o[i] = src ? src[i] : 0.f;
}

// ---------------------------------------------------------------------------------------------------
// fusion + 3-D points
// ---------------------------------------------------------------------------------------------------
// One warp per frame: lane k owns candidate k (rescale + fusion term), the pick is a warp arg-max with the serial
// loop's rule (a later candidate replaces the current best only if strictly larger), and the 31 rows of the pre-grasp
// dilation element are tested by 31 lanes at once.
// rec (optional): the frame's candidate records for the multi-GPU aggregation, float32 [20][4] = (x, y, traditional score,
// ML score or 0), (-1, -1, 0, 0) in unused slots - written here so that no host or framework code reshuffles results.
__global__ void __launch_bounds__(32) fuse_kernel(lg_context c, LgMaskSrc src, const float* __restrict__ depth, lg_camera cam, int have_ml, int n,
                                                  float4* __restrict__ rec) {
    const int b = blockIdx.x, lane = threadIdx.x;
    if (b >= n) return;
    lg_frame_result* res = &c.results[b];
    const LgRegion r = c.region[b];
    const int W = c.W, H = c.H;
    const size_t fo = (size_t)b * c.P;
    const int nc = r.ok ? res->n_candidates : 0;
    if (lane == 0) {
        res->status = c.status[b];
        res->leaf_id = c.leaf_id[b];
        res->region[0] = r.x0; res->region[1] = r.y0; res->region[2] = r.x1; res->region[3] = r.y1;
        res->angle = c.orient[b].angle;
        const uint32_t* mxq = c.dt_max + (size_t)b * 2;
        res->sdf_max = __fmul_rn((float)max(mxq[0], mxq[1]), 1.0f / 65536.0f);
        if (!r.ok) { res->n_candidates = 0; res->n_positive = 0; }
    }
    // per candidate: ML score and the fused score (:205-237)
    double comb = -CUDART_INF, trad_k = 0.0;
    bool has_comb = false;
    if (lane < LG_TOP_K) {
        const int k = lane;
        const bool mlk = have_ml && nc > 1 && k < nc && res->ml_valid[k];
        if (k < nc) trad_k = res->trad[k];
        if (mlk) {
            const float lg = c.logits[c.slot_map[b * LG_TOP_K + k]];
            res->logit[k] = lg;
            const float s = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-lg)));
            const double ml = tanh((double)s * 3.0) * 0.5 + 0.5;
            res->ml[k] = ml;
            const double conf = 1.0 - fabs(ml - 0.5) * 2.0;
            const double w = fmin(0.3, conf * 0.6);
            comb = (1.0 - w) * trad_k + w * ml;
            has_comb = true;
        } else {
            res->logit[k] = CUDART_NAN_F; res->ml[k] = CUDART_NAN; res->ml_valid[k] = 0;
        }
        if (rec) {
            float4 r4 = make_float4(-1.f, -1.f, 0.f, 0.f);
            if (k < nc) r4 = make_float4((float)res->cand_x[k], (float)res->cand_y[k], (float)trad_k, mlk ? (float)res->ml[k] : 0.f);
            rec[(size_t)b * LG_TOP_K + k] = r4;
        }
    }
    if (nc == 0) {
        if (lane == 0) {
            res->best_index = -1; res->ml_used = 0; res->best_score = 0.0;
            res->grasp_x = -1; res->grasp_y = -1;
            for (int k = 0; k < 3; ++k) { res->grasp_3d[k] = CUDART_NAN; res->pre_grasp[k] = CUDART_NAN; }
        }
        return;
    }
    // the serial loop starts from (trad[0], candidate 0) and takes candidate k when its fused score is strictly larger:
    // the winner is the largest fused score (first one among equals) if that beats trad[0], else candidate 0
    double best_score = has_comb ? comb : -CUDART_INF;
    int best = has_comb ? lane : 0x7FFFFFFF;
    if (!(best_score == best_score)) { best_score = -CUDART_INF; best = 0x7FFFFFFF; }     // NaN never wins a '>' test
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const double os = __shfl_xor_sync(0xFFFFFFFFu, best_score, d);
        const int ok = __shfl_xor_sync(0xFFFFFFFFu, best, d);
        if (os > best_score || (os == best_score && ok < best)) { best_score = os; best = ok; }
    }
    const double trad0 = __shfl_sync(0xFFFFFFFFu, trad_k, 0);
    int ml_used = 0;
    if (best != 0x7FFFFFFF && best_score > trad0) ml_used = 1;
    else { best = 0; best_score = trad0; }
    const int u = res->cand_x[best], v = res->cand_y[best];
    const double z = (double)depth[fo + (size_t)v * W + u];
    const double X = (z * ((double)u - cam.cx)) / cam.f, Y = (z * ((double)v - cam.cy)) / cam.f;
    // pre-grasp (:754-819)
    const double nrm = sqrt(X * X + Y * Y + z * z);
    const double d0 = X / nrm, d1 = Y / nrm;
    const uint32_t* bits = c.bits + (size_t)b * c.bits_stride;
    const int ox = r.x0 - 1, oy = r.y0 - 1, bw = r.x1 - r.x0 + 2, bh = r.y1 - r.y0 + 2, wpr = (bw + 31) >> 5;
    bool found = false;
    double p0 = 0.0, p1 = 0.0;
    for (int i = 0; i < 5 && !found; ++i) {
        const double dist = 0.05 + (double)i * 0.01;
        const double t0 = X - d0 * dist, t1 = Y - d1 * dist;
        const int pu = (int)((t0 * cam.f / z) + cam.cx), pv = (int)((t1 * cam.f / z) + cam.cy);
        if (!(pu >= 0 && pu < W && pv >= 0 && pv < H)) continue;
        bool hit = false;
        if (lane < LG_SE_PRE)
            hit = bits_any(bits, wpr, bw, bh, pv + lane - LG_SE_PRE / 2 - oy, pu + c.se31_a[lane] - LG_SE_PRE / 2 - ox,
                           pu + c.se31_b[lane] - LG_SE_PRE / 2 - ox);
        const bool blocked = __any_sync(0xFFFFFFFFu, hit);
        if (!blocked) {
            const double e0 = t0 - X, e1 = t1 - Y;
            if (sqrt(e0 * e0 + e1 * e1 + 0.0) >= 0.05) { p0 = t0; p1 = t1; found = true; }
        }
    }
    if (!found) { p0 = X - d0 * 0.10; p1 = Y - d1 * 0.10; }
    if (lane == 0) {
        res->best_index = best; res->best_score = best_score; res->ml_used = ml_used;
        res->grasp_x = u; res->grasp_y = v;
        res->grasp_3d[0] = X; res->grasp_3d[1] = Y; res->grasp_3d[2] = z;
        res->pre_grasp[0] = p0; res->pre_grasp[1] = p1; res->pre_grasp[2] = z;
    }
}

// copy the internal maps out in the reference's dtypes (standalone score-map API)
__global__ void export_maps_kernel(lg_context c, int n, double* sdf, double* app, float* flat, float* dist, double* acc,
                                   float* stem, double* trad, uint8_t* valid, double* angle) {
    const size_t total = (size_t)n * c.P;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        if (sdf) sdf[i] = c.m_sdf[i];
        if (app) app[i] = c.m_app[i];
        if (flat) flat[i] = c.m_flat[i];
        if (dist) dist[i] = c.di[i];
        if (acc) acc[i] = c.m_acc[i];
        if (stem) stem[i] = c.m_stem[i];
        if (trad) trad[i] = c.m_trad[i];
        if (valid) valid[i] = c.m_valid[i];
    }
    if (angle)
        for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x) angle[b] = c.orient[b].angle;
}

__global__ void export_orient_kernel(lg_context c, int n, double* out5) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const LgOrient o = c.orient[b];
    out5[b * 5] = o.has_angle ? o.angle : CUDART_NAN;
    out5[b * 5 + 1] = o.major; out5[b * 5 + 2] = o.minor; out5[b * 5 + 3] = o.cx; out5[b * 5 + 4] = o.cy;
}

// per-patch min-max normalisation of caller-built raw patches (get_ml_score, :83-123): every channel but the mask
__global__ void __launch_bounds__(256) normalize_patches_kernel(const float* __restrict__ raw, float* __restrict__ out) {
    const int ch = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
    const float* src = raw + ((size_t)n * LG_CHANNELS + ch) * (LG_PATCH * LG_PATCH);
    float* dst = out + ((size_t)n * LG_CHANNELS + ch) * (LG_PATCH * LG_PATCH);
    __shared__ float smn[8], smx[8];
    float v[4], mn = CUDART_INF_F, mx = -CUDART_INF_F;
#pragma unroll
    for (int q = 0; q < 4; ++q) { v[q] = src[tid + q * 256]; mn = fminf(mn, v[q]); mx = fmaxf(mx, v[q]); }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, d));
        mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, d));
    }
    if ((tid & 31) == 0) { smn[tid >> 5] = mn; smx[tid >> 5] = mx; }
    __syncthreads();
    mn = smn[0]; mx = smx[0];
    for (int w = 1; w < 8; ++w) { mn = fminf(mn, smn[w]); mx = fmaxf(mx, smx[w]); }
    const bool norm = ch != 1 && mx > mn;
    const float den = __fsub_rn(mx, mn);
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[tid + q * 256] = norm ? __fdiv_rn(__fsub_rn(v[q], mn), den) : v[q];
}

// ImageProcessor.smooth_depth (image_processor.py:56-64): reflect padding by 2 and the 5x5 Gaussian, taps added in
// row-major order like the fused flatness path above; any image size >= 3 x 3
struct GaussTaps { float g[25]; };
__global__ void __launch_bounds__(256) smooth_depth_kernel(const float* __restrict__ in, float* __restrict__ out, int h, int w,
                                                            GaussTaps taps) {
    __shared__ float t[8 + 4][32 + 4];
    const size_t fo = (size_t)blockIdx.z * h * w;
    const int tx0 = blockIdx.x * 32, ty0 = blockIdx.y * 8;
    for (int i = threadIdx.x; i < 12 * 36; i += 256) {
        const int ly = i / 36, lx = i - ly * 36;
        const int y = reflect101(ty0 - 2 + ly, h), x = reflect101(tx0 - 2 + lx, w);
        t[ly][lx] = in[fo + (size_t)min(max(y, 0), h - 1) * w + min(max(x, 0), w - 1)];
    }
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int x = tx0 + lx, y = ty0 + ly;
    if (x >= w || y >= h) return;
    float acc = __fmul_rn(taps.g[0], t[ly][lx]);
#pragma unroll
    for (int k = 1; k < 25; ++k) acc = __fadd_rn(acc, __fmul_rn(taps.g[k], t[ly + k / 5][lx + k % 5]));
    out[fo + (size_t)y * w + x] = acc;
}

}  // namespace

int lg_run_smooth_depth(const float* in, int n, int h, int w, const float* gauss25, float* out, cudaStream_t st) {
    GaussTaps taps;
    for (int k = 0; k < 25; ++k) taps.g[k] = gauss25[k];
    smooth_depth_kernel<<<dim3((w + 31) / 32, (h + 7) / 8, n), 256, 0, st>>>(in, out, h, w, taps);
    LG_LAUNCH_CHECK();
    return LG_OK;
}

#define TRY_SMEM(kernel, bytes) LG_ENSURE_SMEM(kernel, bytes)

int lg_run_scores(lg_context* c, LgMaskSrc src, const float* depth, int n, lg_camera cam, int full, double* iso_out,
                  cudaStream_t st) {
    const int all_tiles = ((c->W + SC_TW - 1) / SC_TW + 1) * ((c->H + SC_TH - 1) / SC_TH + 1);
    int per_frame = (148 * 8 * 2 + n - 1) / n;          // about two waves of CTAs over the batch, at least 64 per frame
    per_frame = per_frame < 64 ? 64 : (per_frame > all_tiles ? all_tiles : per_frame);
    score_kernel<<<dim3(per_frame, n), SC_TW * SC_TH, 0, st>>>(*c, src, depth, cam, full, iso_out);
    LG_LAUNCH_CHECK();
    return LG_OK;
}

// CTAs per frame of the table kernel: a warp per tile and pass, enough passes in flight to fill the GPU
static int nms_init_ctas(const lg_context* c, int n) {
    const int per_cta = NMS_INIT_NT / 32;
    int g = (LG_NUM_SM_HINT * 8 * 4 + n - 1) / n;
    const int most = (c->tile_cap + per_cta - 1) / per_cta;
    return g < 8 ? 8 : (g > most ? most : g);
}

int lg_run_nms(lg_context* c, int n, cudaStream_t st) {
    nms_tiles_kernel<<<dim3(nms_init_ctas(c, n), n), NMS_INIT_NT, 0, st>>>(*c, nullptr, nullptr);
    LG_LAUNCH_CHECK();
    nms_kernel<<<n, NMS_NT, 0, st>>>(*c, nullptr, nullptr, nullptr, nullptr);
    LG_LAUNCH_CHECK();
    return LG_OK;
}

long long lg_cnn_input_plane_rows(long long n_patches);
int lg_cnn_input_lead();

// packed != 0: write the tensor-core CNN's input (c->cnn_act0) instead of the float32 patch tensor
int lg_run_gather(lg_context* c, LgMaskSrc src, const float* depth, int n, lg_camera cam, int packed, cudaStream_t st) {
    (void)cam;
    compact_slots_kernel<<<1, CS_NT, 0, st>>>(*c, n);
    LG_LAUNCH_CHECK();
    if (packed)
        gather_kernel<true><<<dim3(LG_TOP_K, n), GA_NT, 0, st>>>(*c, src, depth, reinterpret_cast<uint4*>(c->cnn_act0),
                                                                 lg_cnn_input_plane_rows((long long)n * LG_TOP_K), lg_cnn_input_lead());
    else
        gather_kernel<false><<<dim3(LG_TOP_K, n), GA_NT, 0, st>>>(*c, src, depth, nullptr, 0, 0);
    LG_LAUNCH_CHECK();
    c->patches_valid = packed ? 0 : 1;
    return LG_OK;
}

// the patch tensor in the dense [frames][20][9][32][32] order of the API (zeros where no patch was built)
int lg_run_export_patches(lg_context* c, float* out, int n, cudaStream_t st) {
    export_patches_kernel<<<dim3(LG_TOP_K, n), 256, 0, st>>>(*c, out);
    LG_LAUNCH_CHECK();
    return LG_OK;
}

int lg_run_fuse(lg_context* c, LgMaskSrc src, const float* depth, int n, lg_camera cam, int have_ml, lg_frame_result* out,
                float* rec_out, cudaStream_t st) {
    fuse_kernel<<<n, 32, 0, st>>>(*c, src, depth, cam, have_ml, n, reinterpret_cast<float4*>(rec_out));
    LG_LAUNCH_CHECK();
    if (out && out != c->results)
        LG_CUDA(cudaMemcpyAsync(out, c->results, sizeof(lg_frame_result) * n, cudaMemcpyDeviceToDevice, st));
    return LG_OK;
}

int lg_run_export_maps(lg_context* c, int n, double* sdf, double* app, float* flat, float* dist, double* acc, float* stem,
                       double* trad, uint8_t* valid, double* angle, cudaStream_t st) {
    export_maps_kernel<<<LG_NUM_SM_HINT * 8, 256, 0, st>>>(*c, n, sdf, app, flat, dist, acc, stem, trad, valid, angle);
    LG_LAUNCH_CHECK();
    return LG_OK;
}

int lg_run_candidates_from_maps(lg_context* c, const double* score, const uint8_t* valid, int n, int32_t* xy, int32_t* count,
                                cudaStream_t st) {
    clear_status_kernel<<<(n + 63) / 64, 64, 0, st>>>(*c, n);
    LG_LAUNCH_CHECK();
    nms_tiles_kernel<<<dim3(nms_init_ctas(c, n), n), NMS_INIT_NT, 0, st>>>(*c, score, valid);
    LG_LAUNCH_CHECK();
    nms_kernel<<<n, NMS_NT, 0, st>>>(*c, score, valid, xy, count);
    LG_LAUNCH_CHECK();
    return LG_OK;
}

int lg_run_export_orient(lg_context* c, int n, double* out5, cudaStream_t st) {
    export_orient_kernel<<<(n + 63) / 64, 64, 0, st>>>(*c, n, out5);
    LG_LAUNCH_CHECK();
    return LG_OK;
}

int lg_run_normalize_patches(const float* raw, int n, float* out, cudaStream_t st) {
    normalize_patches_kernel<<<dim3(LG_CHANNELS, n), 256, 0, st>>>(raw, out);
    LG_LAUNCH_CHECK();
    return LG_OK;
}
