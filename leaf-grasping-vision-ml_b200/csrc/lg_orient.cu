// Leaf orientation: largest outer contour -> minimum-area rectangle -> angle
// (reference scripts/utils/grasp_point_selector.py:718-752: cv2.findContours(RETR_EXTERNAL) +
//  max(cv2.contourArea) + cv2.minAreaRect, then angle += 90 when width < height).
//
// One CTA per frame.  The chosen leaf is turned into a bitmask over its bounding box (kept in global
// scratch, L1-resident; the stem-penalty and pre-grasp dilations reuse it), then:
//   rows -> runs (parallel)                                   -> 8-connected components by union-find on runs
//   outer border of the contending components by Moore tracing -> shoelace area == cv2.contourArea
//   winner's per-row extents (parallel)                        -> convex hull (two monotone chains)
//   rotating calipers in float32, restated from OpenCV's rotcalipers.cpp
// The serial parts run on one thread: they touch a few hundred runs / hull points per frame and the
// batch supplies the parallelism (one frame per CTA, many CTAs per SM).
#include <math_constants.h>

#include "lg_internal.cuh"

namespace {

constexpr int OR_NT = 128;
// The serial sections (union-find over runs, border trace, hull, calipers) are chains of dependent loads; on global
// scratch every link costs an L2 round trip (the data was written by other threads of the CTA, so L1 does not hold it).
// Working sets that fit are therefore staged in shared memory: the bitmask, the run tables, and - aliased on the run
// tables, which are dead by then - the row extents, hull points and caliper work arrays.  Larger leaves use the global
// scratch as before.
constexpr int OR_S_BITS = 8192;     // bitmask words (32 KB)
constexpr int OR_S_RUNS = 1536;     // runs: 3 x u16 + 6 x i32 per run (45 KB)
constexpr int OR_S_ROWS = 960;      // rows of the winning component: 12 ints per row (45 KB), aliases the run tables
constexpr int OR_S_ROWF = 1104;     // row -> first run table (rows of the bounding box + 2)
constexpr size_t OR_SMEM = (size_t)OR_S_BITS * 4 + (size_t)OR_S_ROWF * 4 + (size_t)OR_S_ROWS * 8 + (size_t)OR_S_RUNS * 30;
static_assert((size_t)OR_S_ROWS * 40 <= (size_t)OR_S_RUNS * 30, "hull points and caliper arrays alias the run tables");

struct Bits {
    const uint32_t* w;
    int wpr, bw, bh;
};

// Lock-free union-find on run indices (shared or global memory).  Roots only ever decrease.
__device__ __forceinline__ int uf_root(const int32_t* parent, int i) {
    int p = ((const volatile int32_t*)parent)[i];
    while (p != i) { i = p; p = ((const volatile int32_t*)parent)[i]; }
    return i;
}
__device__ void uf_union(int32_t* parent, int a, int b) {
    while (true) {
        a = uf_root(parent, a);
        b = uf_root(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }     // a > b: hang a under b
        const int old = atomicMin(&parent[a], b);
        if (old == a) return;                              // a was still a root
        a = old;                                           // somebody re-parented a meanwhile: join that tree with b's
    }
}

// bits (lx-1, lx, lx+1) of row ly as bits 0..2; pixels outside the bitmask read as 0
__device__ __forceinline__ unsigned bits3(const Bits& B, int lx, int ly) {
    if (ly < 0 || ly >= B.bh) return 0u;
    const uint32_t* row = B.w + ly * B.wpr;
    const int x0 = lx - 1, wi = x0 >> 5, o = x0 & 31;                 // x0 = -1 -> wi = -1, o = 31
    const uint32_t lo = (wi >= 0 && wi < B.wpr) ? row[wi] : 0u;
    const uint32_t hi = (o > 29 && wi + 1 >= 0 && wi + 1 < B.wpr) ? row[wi + 1] : 0u;
    return __funnelshift_r(lo, hi, o) & 7u;
}

// |shoelace| * 2 of the outer border that starts at the raster-first pixel (sx, sy) of a component.
// Moore tracing: from the current pixel the eight neighbours are examined clockwise, starting behind the direction
// the trace arrived from.  The three rows around the pixel are fetched at once and folded into an 8-bit neighbour
// mask (bit d = neighbour in direction d, 0 = east, clockwise in image coordinates), so that a step costs one round
// of loads instead of up to eight dependent probes.
__device__ long long trace_twice_area(const Bits& B, int sx, int sy) {
    const int DX[8] = {1, 1, 0, -1, -1, -1, 0, 1};
    const int DY[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    int cx = sx, cy = sy, db = 4, first = -1;
    long long a00 = 0;
    for (long long guard = 0; guard < (1ll << 26); ++guard) {
        const unsigned wm = bits3(B, cx, cy - 1), w0 = bits3(B, cx, cy), wp = bits3(B, cx, cy + 1);
        const unsigned nb = ((w0 >> 2) & 1u) | (((wp >> 2) & 1u) << 1) | (((wp >> 1) & 1u) << 2) | ((wp & 1u) << 3) |
                            ((w0 & 1u) << 4) | ((wm & 1u) << 5) | (((wm >> 1) & 1u) << 6) | (((wm >> 2) & 1u) << 7);
        if (nb == 0) break;                                   // isolated pixel
        const int r = (db + 1) & 7;
        const unsigned rot = ((nb >> r) | (nb << (8 - r))) & 0xFFu;
        const int d = (r + __ffs(rot) - 1) & 7;               // first set neighbour in the order db+1, db+2, ...
        if (cx == sx && cy == sy && first >= 0 && d == first) break;
        if (first < 0) first = d;
        const int nx = cx + DX[d], ny = cy + DY[d];
        a00 += (long long)cx * ny - (long long)nx * cy;
        cx = nx; cy = ny;
        db = (d + ((d & 1) ? 5 : 6)) & 7;
    }
    return a00 < 0 ? -a00 : a00;
}

struct RectOut {
    float cx, cy, w, h, angle_deg;
};

// OpenCV rotatingCalipers(CALIPERS_MINAREARECT) + the tail of cv::minAreaRect, float32, no FMA.
// edge i = hull[i] -> hull[i + 1] (the whole CTA takes part)
__device__ void rect_edges(const int32_t* hull, int n, float* vx, float* vy, float* inv) {
    for (int i = threadIdx.x; i < n; i += OR_NT) {
        const int j = (i + 1 < n) ? i + 1 : 0;
        const double dx = (double)(float)hull[2 * j] - (double)(float)hull[2 * i];
        const double dy = (double)(float)hull[2 * j + 1] - (double)(float)hull[2 * i + 1];
        vx[i] = (float)dx; vy[i] = (float)dy;
        inv[i] = (float)(1.0 / sqrt(dx * dx + dy * dy));
    }
}

__device__ RectOut min_area_rect(const int32_t* hull, int n, const float* vx, const float* vy, const float* inv) {
    RectOut R;
    R.cx = R.cy = R.w = R.h = R.angle_deg = 0.f;
    if (n == 1) { R.cx = (float)hull[0]; R.cy = (float)hull[1]; return R; }
    if (n == 2) {
        float x0 = (float)hull[0], y0 = (float)hull[1], x1 = (float)hull[2], y1 = (float)hull[3];
        R.cx = __fmul_rn(__fadd_rn(x0, x1), 0.5f);
        R.cy = __fmul_rn(__fadd_rn(y0, y1), 0.5f);
        double dx = (double)x1 - x0, dy = (double)y1 - y0;
        R.w = (float)sqrt(dx * dx + dy * dy);
        R.h = 0.f;
        R.angle_deg = (float)(atan2(dy, dx) * 180.0 / 3.14159265358979323846);
        return R;
    }
    int left = 0, bottom = 0, right = 0, top = 0;
    float px0 = (float)hull[0], py0 = (float)hull[1];
    float left_x = px0, right_x = px0, top_y = py0, bottom_y = py0;
    for (int i = 0; i < n; ++i) {          // vx, vy, inv: edge vectors and inverse lengths, filled by rect_edges
        px0 = (float)hull[2 * i]; py0 = (float)hull[2 * i + 1];
        if (px0 < left_x) { left_x = px0; left = i; }
        if (px0 > right_x) { right_x = px0; right = i; }
        if (py0 > top_y) { top_y = py0; top = i; }
        if (py0 < bottom_y) { bottom_y = py0; bottom = i; }
    }
    float orientation = 0.f;
    {
        double ax = vx[n - 1], ay = vy[n - 1];
        for (int i = 0; i < n; ++i) {
            double bx = vx[i], by = vy[i];
            double conv = ax * by - ay * bx;
            if (conv != 0) { orientation = conv > 0 ? 1.f : -1.f; break; }
            ax = bx; ay = by;
        }
    }
    float base_a = orientation, base_b = 0.f;
    int seq[4] = {bottom, right, top, left};
    float minarea = 3.402823466e+38f;
    int b_left = 0, b_bottom = 0;
    float b_a = 1.f, b_w = 0.f, b_b = 0.f, b_h = 0.f;
    auto PX = [&](int i) { return (float)hull[2 * i]; };
    auto PY = [&](int i) { return (float)hull[2 * i + 1]; };
    for (int k = 0; k < n; ++k) {
        float dp0 = __fadd_rn(__fmul_rn(base_a, vx[seq[0]]), __fmul_rn(base_b, vy[seq[0]]));
        float dp1 = __fadd_rn(__fmul_rn(-base_b, vx[seq[1]]), __fmul_rn(base_a, vy[seq[1]]));
        float dp2 = __fsub_rn(__fmul_rn(-base_a, vx[seq[2]]), __fmul_rn(base_b, vy[seq[2]]));
        float dp3 = __fsub_rn(__fmul_rn(base_b, vx[seq[3]]), __fmul_rn(base_a, vy[seq[3]]));
        float maxcos = __fmul_rn(dp0, inv[seq[0]]);
        int main_el = 0;
        float c1 = __fmul_rn(dp1, inv[seq[1]]);
        if (c1 > maxcos) { main_el = 1; maxcos = c1; }
        float c2 = __fmul_rn(dp2, inv[seq[2]]);
        if (c2 > maxcos) { main_el = 2; maxcos = c2; }
        float c3 = __fmul_rn(dp3, inv[seq[3]]);
        if (c3 > maxcos) { main_el = 3; maxcos = c3; }
        int pi = seq[main_el];
        float lx = __fmul_rn(vx[pi], inv[pi]), ly = __fmul_rn(vy[pi], inv[pi]);
        if (main_el == 0) { base_a = lx; base_b = ly; }
        else if (main_el == 1) { base_a = ly; base_b = -lx; }
        else if (main_el == 2) { base_a = -lx; base_b = -ly; }
        else { base_a = -ly; base_b = lx; }
        seq[main_el] = (seq[main_el] + 1 == n) ? 0 : seq[main_el] + 1;
        float dx = __fsub_rn(PX(seq[1]), PX(seq[3])), dy = __fsub_rn(PY(seq[1]), PY(seq[3]));
        float width = __fadd_rn(__fmul_rn(dx, base_a), __fmul_rn(dy, base_b));
        dx = __fsub_rn(PX(seq[2]), PX(seq[0])); dy = __fsub_rn(PY(seq[2]), PY(seq[0]));
        float height = __fadd_rn(__fmul_rn(-dx, base_b), __fmul_rn(dy, base_a));
        float area = __fmul_rn(width, height);
        if (area <= minarea) {
            minarea = area; b_left = seq[3]; b_a = base_a; b_w = width; b_b = base_b; b_h = height; b_bottom = seq[0];
        }
    }
    float A1 = b_a, B1 = b_b, A2 = -b_b, B2 = b_a;
    float C1 = __fadd_rn(__fmul_rn(A1, PX(b_left)), __fmul_rn(PY(b_left), B1));
    float C2 = __fadd_rn(__fmul_rn(A2, PX(b_bottom)), __fmul_rn(PY(b_bottom), B2));
    float idet = __fdiv_rn(1.f, __fsub_rn(__fmul_rn(A1, B2), __fmul_rn(A2, B1)));
    float ox = __fmul_rn(__fsub_rn(__fmul_rn(C1, B2), __fmul_rn(C2, B1)), idet);
    float oy = __fmul_rn(__fsub_rn(__fmul_rn(A1, C2), __fmul_rn(A2, C1)), idet);
    float o2 = __fmul_rn(A1, b_w), o3 = __fmul_rn(B1, b_w), o4 = __fmul_rn(A2, b_h), o5 = __fmul_rn(B2, b_h);
    R.cx = __fadd_rn(ox, __fmul_rn(__fadd_rn(o2, o4), 0.5f));
    R.cy = __fadd_rn(oy, __fmul_rn(__fadd_rn(o3, o5), 0.5f));
    R.w = (float)sqrt((double)o2 * o2 + (double)o3 * o3);
    R.h = (float)sqrt((double)o4 * o4 + (double)o5 * o5);
    R.angle_deg = (float)(atan2((double)o3, (double)o2) * 180.0 / 3.14159265358979323846);
    return R;
}

__global__ void __launch_bounds__(OR_NT) orient_kernel(lg_context c, LgMaskSrc src, int n) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const LgRegion r = c.region[b];
    LgOrient* out = &c.orient[b];
    __shared__ int s_total, s_win, s_ytop, s_ybot, s_fail, s_wlx, s_wly;
    __shared__ int s_redp[OR_NT / 32], s_redi[OR_NT / 32], s_nborder, s_big;
    if (!r.ok) {
        if (tid == 0) {
            LgOrient o;
            o.angle = CUDART_NAN; o.cos_a = 1; o.sin_a = 0; o.major = o.minor = o.cx = o.cy = 0;
            o.has_angle = 0; o.n_hull = 0; o.win_lx = o.win_ly = -1; o.status = 0; o.pad = 0;
            *out = o;
        }
        return;
    }
    const int W = c.W, H = c.H;
    const int ox = r.x0 - 1, oy = r.y0 - 1;             // origin of the local frame (1 px empty ring)
    const int bw = r.x1 - r.x0 + 2, bh = r.y1 - r.y0 + 2;
    const int wpr = (bw + 31) >> 5;
    uint32_t* bits = c.bits + (size_t)b * c.bits_stride;
    extern __shared__ __align__(16) unsigned char or_smem[];
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(or_smem);
    int32_t* s_rowf = reinterpret_cast<int32_t*>(or_smem + (size_t)OR_S_BITS * 4);
    int32_t* s_ext = s_rowf + OR_S_ROWF;
    unsigned char* s_work = reinterpret_cast<unsigned char*>(s_ext + 2 * OR_S_ROWS);
    const bool bits_in_smem = bh * wpr <= OR_S_BITS;
    const size_t fo = (size_t)b * c.P;
    const int id = src.id(b);
    // 1. bitmask (global copy: the stem-penalty and pre-grasp dilations and the sample collector read it later)
    for (int i = tid; i < bh * wpr; i += OR_NT) {
        const int ly = i / wpr, wi = i - ly * wpr;
        const int y = oy + ly;
        uint32_t word = 0;
        if (y >= r.y0 && y < r.y1) {
            const size_t rowo = (size_t)y * W;
#pragma unroll
            for (int k = 0; k < 32; ++k) {     // loads from clamped addresses, so that all 32 are in flight together
                const int x = ox + wi * 32 + k;
                const bool ok = x >= r.x0 && x < r.x1;
                word |= (uint32_t)(src.at(fo, rowo + (ok ? x : r.x0), id) && ok) << k;
            }
        }
        bits[i] = word;
        if (bits_in_smem) s_bits[i] = word;
    }
    __syncthreads();
    if (bits_in_smem) bits = s_bits;
    Bits B{bits, wpr, bw, bh};
    // Number of leaf pixels with a background pixel among their 8 neighbours (all components together).  Every pixel of
    // a component that is not such a border pixel lies strictly inside the component's outer contour, so
    // 2 * (pixels of the component - border pixels) bounds twice the contour area from below (Pick's theorem), which
    // usually decides the largest-contour question without tracing anything (step 6).
    {
        int nbp = 0;
        for (int i = tid; i < bh * wpr; i += OR_NT) {
            const int ly = i / wpr, wi = i - ly * wpr;
            const uint32_t m = bits[i];
            if (!m) continue;
            uint32_t inner = 0xFFFFFFFFu;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = ly + dy;
                uint32_t c0 = 0, l0 = 0, r0 = 0;
                if (yy >= 0 && yy < bh) {
                    const uint32_t* row = bits + yy * wpr;
                    c0 = row[wi];
                    l0 = wi > 0 ? row[wi - 1] : 0u;
                    r0 = wi + 1 < wpr ? row[wi + 1] : 0u;
                }
                inner &= c0 & ((c0 << 1) | (l0 >> 31)) & ((c0 >> 1) | (r0 << 31));
            }
            nbp += __popc(m & ~inner);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) nbp += __shfl_xor_sync(0xFFFFFFFFu, nbp, d);
        if ((tid & 31) == 0) s_redp[tid >> 5] = nbp;
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < OR_NT / 32; ++w) t += s_redp[w]; s_nborder = t; }
    }
    int32_t* row_first = (bh + 1 <= OR_S_ROWF) ? s_rowf : c.row_first + (size_t)b * (H + 3);
    // 2. runs per row
    for (int ly = tid; ly < bh; ly += OR_NT) {
        int cnt = 0;
        uint32_t carry = 0;
        for (int wi = 0; wi < wpr; ++wi) {
            uint32_t w = bits[ly * wpr + wi];
            cnt += __popc(w & ~((w << 1) | carry));
            carry = w >> 31;
        }
        row_first[ly + 1] = cnt;
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        row_first[0] = 0;
        for (int ly = 0; ly < bh; ++ly) { acc += row_first[ly + 1]; row_first[ly + 1] = acc; }
        s_total = acc;
        s_fail = acc > c.run_cap;
    }
    __syncthreads();
    const int total = s_total;
    if (s_fail || total == 0) {
        if (tid == 0) {
            LgOrient o;
            o.angle = CUDART_NAN; o.cos_a = 1; o.sin_a = 0; o.major = o.minor = o.cx = o.cy = 0;
            o.has_angle = 0; o.n_hull = 0; o.win_lx = o.win_ly = -1; o.status = s_fail ? LG_ST_RUNS_OVERFLOW : 0; o.pad = 0;
            *out = o;
            if (s_fail) atomicOr(&c.status[b], LG_ST_RUNS_OVERFLOW);
        }
        return;
    }
    const bool runs_in_smem = total <= OR_S_RUNS;
    const int rcap = runs_in_smem ? OR_S_RUNS : c.run_cap;
    int32_t* parent = runs_in_smem ? reinterpret_cast<int32_t*>(s_work) : c.run_parent + (size_t)b * c.run_cap * 6;
    int32_t* cpix = parent + rcap;
    int32_t* cx0 = cpix + rcap;
    int32_t* cx1 = cx0 + rcap;
    int32_t* cy0 = cx1 + rcap;
    int32_t* cy1 = cy0 + rcap;
    uint16_t* rx0 = runs_in_smem ? reinterpret_cast<uint16_t*>(cy1 + rcap) : c.run_x0 + (size_t)b * c.run_cap;
    uint16_t* rx1 = runs_in_smem ? rx0 + rcap : c.run_x1 + (size_t)b * c.run_cap;
    uint16_t* ry = runs_in_smem ? rx1 + rcap : c.run_y + (size_t)b * c.run_cap;
    // 3. extract runs
    for (int ly = tid; ly < bh; ly += OR_NT) {
        int o = row_first[ly];
        bool open = false;
        int start = 0;
        for (int wi = 0; wi < wpr; ++wi) {
            uint32_t m = bits[ly * wpr + wi];
            const int xb = wi * 32;
            if (open) {
                if (m == 0xFFFFFFFFu) continue;
                int e = __ffs(~m) - 1;
                rx0[o] = (uint16_t)start; rx1[o] = (uint16_t)(xb + e - 1); ry[o] = (uint16_t)ly; ++o;
                open = false;
                m &= ~((1u << e) - 1u);
            }
            while (m) {
                int s = __ffs(m) - 1;
                uint32_t t = m >> s;
                int e = (t == 0xFFFFFFFFu) ? 32 : (__ffs(~t) - 1);
                if (s + e >= 32) { open = true; start = xb + s; break; }
                rx0[o] = (uint16_t)(xb + s); rx1[o] = (uint16_t)(xb + s + e - 1); ry[o] = (uint16_t)ly; ++o;
                m &= ~(((1u << e) - 1u) << s);
            }
        }
        if (open) { rx0[o] = (uint16_t)start; rx1[o] = (uint16_t)(bw - 1); ry[o] = (uint16_t)ly; ++o; }
    }
    for (int i = tid; i < total; i += OR_NT) {
        parent[i] = i; cpix[i] = 0; cx0[i] = 0x7FFFFFFF; cx1[i] = -1; cy0[i] = 0x7FFFFFFF; cy1[i] = -1;
    }
    __syncthreads();
    // 4. components: one thread per row links the runs of its row with the touching runs of the row above (lock-free
    //    union-find, the smaller run index becomes the root: a component's root is its raster-first run)
    for (int ly = 1 + tid; ly < bh; ly += OR_NT) {
        int i = row_first[ly], iend = row_first[ly + 1];
        int j = row_first[ly - 1], jend = row_first[ly];
        while (i < iend && j < jend) {
            if ((int)rx1[j] + 1 < (int)rx0[i]) ++j;
            else if ((int)rx1[i] + 1 < (int)rx0[j]) ++i;
            else {
                uf_union(parent, i, j);
                if (rx1[j] < rx1[i]) ++j; else ++i;
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < total; i += OR_NT) parent[i] = uf_root(parent, i);     // a write only ever shortens a path
    __syncthreads();
    // 5. per-component pixel count and bounding box (integer atomics: order independent)
    for (int i = tid; i < total; i += OR_NT) {
        const int root = parent[i];
        atomicAdd(&cpix[root], (int)rx1[i] - (int)rx0[i] + 1);
        atomicMin(&cx0[root], (int)rx0[i]); atomicMax(&cx1[root], (int)rx1[i]);
        atomicMin(&cy0[root], (int)ry[i]); atomicMax(&cy1[root], (int)ry[i]);
    }
    __syncthreads();
    // the component with most pixels, the first one among equals
    {
        int bp = -1, bi = 0x7FFFFFFF;
        for (int i = tid; i < total; i += OR_NT)
            if (parent[i] == i && cpix[i] > bp) { bp = cpix[i]; bi = i; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const int op = __shfl_xor_sync(0xFFFFFFFFu, bp, d), oi = __shfl_xor_sync(0xFFFFFFFFu, bi, d);
            if (op > bp || (op == bp && oi < bi)) { bp = op; bi = oi; }
        }
        if ((tid & 31) == 0) { s_redp[tid >> 5] = bp; s_redi[tid >> 5] = bi; }
    }
    __syncthreads();
    if (tid == 0) {
        int bp = s_redp[0], big = s_redi[0];
        for (int w = 1; w < OR_NT / 32; ++w)
            if (s_redp[w] > bp || (s_redp[w] == bp && s_redi[w] < big)) { bp = s_redp[w]; big = s_redi[w]; }
        s_big = big;
    }
    __syncthreads();
    // 6. winner = outer contour of largest polygon area.  A component whose bounding box cannot hold the lower bound
    //    of the biggest component's area is out; only if some component survives that test are contours traced (serial).
    const int big = s_big;
    const long long area_lb = 2ll * ((long long)cpix[big] - (long long)s_nborder);
    int contender = 0;
    for (int i = tid; i < total; i += OR_NT)
        if (parent[i] == i && i != big && 2ll * (cx1[i] - cx0[i]) * (cy1[i] - cy0[i]) >= area_lb) contender = 1;
    contender = __syncthreads_or(contender);
    if (tid == 0) {
        int win = big;
        if (contender) {
            long long best_a = trace_twice_area(B, rx0[big], ry[big]);
            for (int i = 0; i < total; ++i) {
                if (parent[i] != i || i == big) continue;
                long long bound = 2ll * (cx1[i] - cx0[i]) * (cy1[i] - cy0[i]);   // twice the bbox polygon area
                if (bound < best_a) continue;
                long long a = trace_twice_area(B, rx0[i], ry[i]);
                // cv2 lists contours last-found first and max() keeps the first maximum: ties go to the later start
                if (a > best_a || (a == best_a && i > win)) { best_a = a; win = i; }
            }
        }
        s_win = win; s_ytop = cy0[win]; s_ybot = cy1[win]; s_wlx = rx0[win]; s_wly = ry[win];
    }
    __syncthreads();
    const int win = s_win, ytop = s_ytop, ybot = s_ybot;
    // per-frame scratch of 12*(H+2) ints: hull points | per-row (minx, maxx) | float work arrays
    const int nrows = ybot - ytop + 1;
    const bool hull_in_smem = runs_in_smem && nrows <= OR_S_ROWS;   // (the global run tables do not alias the hull scratch)
    int32_t* hull = hull_in_smem ? reinterpret_cast<int32_t*>(s_work) : c.hull + (size_t)b * (12 * (H + 2));
    int32_t* ext = hull_in_smem ? s_ext : hull + 4 * (H + 2);
    const int wlx = s_wlx, wly = s_wly;
    // 7. winner's row extents
    for (int ly = ytop + tid; ly <= ybot; ly += OR_NT) {
        int mn = 1 << 30, mx = -1;
        for (int i = row_first[ly]; i < row_first[ly + 1]; ++i)
            if (parent[i] == win) { mn = min(mn, (int)rx0[i]); mx = max(mx, (int)rx1[i]); }
        ext[2 * (ly - ytop)] = mn; ext[2 * (ly - ytop) + 1] = mx;
    }
    __syncthreads();
    // 8. hull + calipers
    if (tid == 0) {
        // hull points are written in image coordinates; chain 1 walks down the left side, chain 2 up the right
        int nh = 0;
        auto crossz = [&](int ax, int ay, int bx, int by, int px, int py) -> long long {
            return (long long)(bx - ax) * (py - by) - (long long)(by - ay) * (px - bx);
        };
        const int rows = ybot - ytop + 1;
        for (int k = 0; k < rows; ++k) {
            int px = ext[2 * k] + ox, py = ytop + k + oy;
            while (nh >= 2 && crossz(hull[2 * (nh - 2)], hull[2 * (nh - 2) + 1], hull[2 * (nh - 1)], hull[2 * (nh - 1) + 1], px, py) >= 0) --nh;
            hull[2 * nh] = px; hull[2 * nh + 1] = py; ++nh;
        }
        const int base = nh;   // chain 2 may not pop below the bottom-left vertex
        for (int k = rows - 1; k >= 0; --k) {
            int px = ext[2 * k + 1] + ox, py = ytop + k + oy;
            if (nh > 0 && hull[2 * (nh - 1)] == px && hull[2 * (nh - 1) + 1] == py) continue;
            while (nh >= base + 1 && nh >= 2 &&
                   crossz(hull[2 * (nh - 2)], hull[2 * (nh - 2) + 1], hull[2 * (nh - 1)], hull[2 * (nh - 1) + 1], px, py) >= 0) --nh;
            hull[2 * nh] = px; hull[2 * nh + 1] = py; ++nh;
        }
        if (nh > 1 && hull[2 * (nh - 1)] == hull[0] && hull[2 * (nh - 1) + 1] == hull[1]) --nh;
        // the closing turn at the start vertex can still be flat: drop a last point collinear with it
        while (nh >= 3 && crossz(hull[2 * (nh - 2)], hull[2 * (nh - 2) + 1], hull[2 * (nh - 1)], hull[2 * (nh - 1) + 1], hull[0], hull[1]) >= 0) --nh;
        s_total = nh;
    }
    __syncthreads();
    const int nh = s_total;
    float* fs = hull_in_smem ? reinterpret_cast<float*>(hull + 4 * OR_S_ROWS) : reinterpret_cast<float*>(hull + 6 * (H + 2));
    if (nh >= 3) rect_edges(hull, nh, fs, fs + nh, fs + 2 * nh);
    __syncthreads();
    if (tid == 0) {
        RectOut R = min_area_rect(hull, nh, fs, fs + nh, fs + 2 * nh);
        double ang = (double)R.angle_deg;
        if (R.w < R.h) ang = ang + 90.0;
        LgOrient o;
        o.angle = ang * (3.14159265358979323846 / 180.0);
        o.cos_a = cos(o.angle); o.sin_a = sin(o.angle);
        o.major = fmaxf(R.w, R.h); o.minor = fminf(R.w, R.h); o.cx = R.cx; o.cy = R.cy;
        o.has_angle = 1; o.n_hull = nh; o.status = 0; o.pad = 0;
        o.win_lx = wlx; o.win_ly = wly;
        *out = o;
    }
}

// bounding box of a u8 mask per frame (entry for the standalone score-map API)
__global__ void mask_bbox_kernel(lg_context c, const uint8_t* mask, int full) {
    const int b = blockIdx.y;
    const int W = c.W;
    const size_t P = c.P;
    __shared__ unsigned sx0, sx1, sy0, sy1;
    if (threadIdx.x == 0) { sx0 = 0xFFFFFFFFu; sy0 = 0xFFFFFFFFu; sx1 = 0; sy1 = 0; }
    __syncthreads();
    unsigned x0 = 0xFFFFFFFFu, y0 = 0xFFFFFFFFu, x1 = 0, y1 = 0;
    bool any = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (size_t)gridDim.x * blockDim.x) {
        if (mask[(size_t)b * P + i]) {
            unsigned x = (unsigned)(i % W), y = (unsigned)(i / W);
            x0 = min(x0, x); x1 = max(x1, x); y0 = min(y0, y); y1 = max(y1, y);
            any = true;
        }
    }
    if (any) { atomicMin(&sx0, x0); atomicMax(&sx1, x1); atomicMin(&sy0, y0); atomicMax(&sy1, y1); }
    __syncthreads();
    if (threadIdx.x == 0 && sx0 != 0xFFFFFFFFu) {
        // bx0.. of label slot 1 double as the scratch for this reduction
        size_t o = (size_t)b * c.L + 1;
        atomicMin(&c.bx0[o], sx0); atomicMax(&c.bx1[o], sx1); atomicMin(&c.by0[o], sy0); atomicMax(&c.by1[o], sy1);
    }
}

__global__ void mask_region_kernel(lg_context c, int n, int full) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    size_t o = (size_t)b * c.L + 1;
    LgRegion r;
    r.ok = c.bx0[o] != 0xFFFFFFFFu;
    r.x0 = r.ok ? (int)c.bx0[o] : 0; r.x1 = r.ok ? (int)c.bx1[o] + 1 : 0;
    r.y0 = r.ok ? (int)c.by0[o] : 0; r.y1 = r.ok ? (int)c.by1[o] + 1 : 0;
    if (full) { r.sx0 = 0; r.sy0 = 0; r.sx1 = c.W; r.sy1 = c.H; }
    else {
        r.sx0 = max(0, r.x0 - LG_REGION_PAD); r.sy0 = max(0, r.y0 - LG_REGION_PAD);
        r.sx1 = min(c.W, r.x1 + LG_REGION_PAD); r.sy1 = min(c.H, r.y1 + LG_REGION_PAD);
    }
    c.region[b] = r;
    c.leaf_id[b] = r.ok ? 1 : -1;
}

__global__ void mask_clear_kernel(lg_context c, int n) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    size_t o = (size_t)b * c.L + 1;
    c.bx0[o] = 0xFFFFFFFFu; c.by0[o] = 0xFFFFFFFFu; c.bx1[o] = 0; c.by1[o] = 0;
    c.status[b] = 0;
}

}  // namespace

int lg_run_orientation(lg_context* c, LgMaskSrc src, int n, cudaStream_t st) {
    LG_ENSURE_SMEM(orient_kernel, OR_SMEM);
    LG_PREFER_LARGE_SMEM(orient_kernel);
    orient_kernel<<<n, OR_NT, OR_SMEM, st>>>(*c, src, n);
    LG_LAUNCH_CHECK();
    return LG_OK;
}

int lg_run_mask_regions(lg_context* c, const uint8_t* mask, int n, int full, cudaStream_t st) {
    mask_clear_kernel<<<(n + 63) / 64, 64, 0, st>>>(*c, n);
    LG_LAUNCH_CHECK();
    mask_bbox_kernel<<<dim3(64, n), 256, 0, st>>>(*c, mask, full);
    LG_LAUNCH_CHECK();
    mask_region_kernel<<<(n + 63) / 64, 64, 0, st>>>(*c, n, full);
    LG_LAUNCH_CHECK();
    return LG_OK;
}
