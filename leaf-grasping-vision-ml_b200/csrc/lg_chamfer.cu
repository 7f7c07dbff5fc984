// 5x5 chamfer distance transform, bit-exact with cv2.distanceTransform(mask, DIST_L2, 5) (IPP off)
// as the reference calls it at scripts/utils/grasp_point_selector.py:266,529-530.
//
// OpenCV's algorithm is two raster sweeps over an integer (Q16) field.  Within one row the only
// sequential tap is t[x] = min(t[x], t[x-1] + a), which is a running minimum of (t[x] - a*x); the
// other seven taps read the two previous rows.  So a row is: 7-tap stencil (parallel) + one prefix-min
// scan (parallel), and rows follow each other.  One CTA owns one transform and walks its rows with the
// last three rows of the field in shared memory; many transforms (frames x {inside, outside}) run
// side by side, which is where the HBM bandwidth comes from.
//
// Traffic per transform of h x w pixels: source read once (1 or 2 B/px), forward field written and
// read back (4+4 B/px, L2-resident when the batch is small), result written once (4 B/px) when asked.
#include <stdlib.h>

#include "lg_internal.cuh"

namespace {

constexpr int CH_NT = 512;
constexpr int CH_MAXI = 8;   // pixels per thread per row -> rows up to 4096 wide

struct ChamferArgs {
    LgMaskSrc src;
    const LgRegion* region;   // used when rect_mode (variant 0 only)
    int n, H, W;
    int rect_mode;            // 1: variant 0 runs on region bbox grown by 1 px, 0: full frame
    int invert_base;          // variant v transforms (mask != 0) ^ invert_base ^ v
    int32_t* fwd;             // [nvar][n][P]
    float* out_f32;           // variant 0 only, [n][P] (may be null)
    uint32_t* out_q16;        // variant 0 only
    uint32_t* out_max;        // [n][nvar] (may be null)
    const uint32_t* need_full; // [n] (may be null): variant 1 only runs for frames whose entry is non-zero
    size_t P;
    int warp_path;            // 1: rectangles up to CHW_MAXW wide (variant 0, rect_mode) are left to chamfer_warp_kernel
    int nvar;                 // variants per frame (layout of out_max)
    int var_base;             // first variant of this launch (variant = blockIdx.y + var_base)
};

// LPT = source / forward values each thread fetches per row (row width <= LPT * CH_NT), D = rows fetched ahead.
// The fetches of row ly + 1 + D are issued D iterations before their values are stored to shared memory, so the
// global-memory latency of the row stream is hidden behind D row steps instead of being paid once per row.
template <int LPT, int D>
__global__ void __launch_bounds__(CH_NT, LPT <= 3 ? 2 : 1) chamfer_kernel(ChamferArgs A) {
    extern __shared__ int smem[];
    __shared__ int wtot[2][CH_NT / 32];
    __shared__ unsigned smax[CH_NT / 32];
    const int b = blockIdx.x, var = blockIdx.y + A.var_base, nvar = A.nvar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = A.W, H = A.H;
    if (var == 1 && A.need_full && A.need_full[b] == 0) return;   // outside_max_kernel already produced this frame's maximum
    int rx0 = 0, ry0 = 0, rx1 = W, ry1 = H;
    int id = 1;
    if (A.src.labels) {
        id = A.src.leaf_id[b];
        if (id < 0) {
            if (A.out_max && tid == 0) A.out_max[b * nvar + var] = 0;
            return;
        }
    }
    if (A.rect_mode && var == 0) {
        LgRegion r = A.region[b];
        if (!r.ok) {
            if (A.out_max && tid == 0) A.out_max[b * nvar + var] = 0;
            return;
        }
        rx0 = max(0, r.x0 - 1); ry0 = max(0, r.y0 - 1); rx1 = min(W, r.x1 + 1); ry1 = min(H, r.y1 + 1);
    }
    const int w = rx1 - rx0, h = ry1 - ry0;
    const int wpad = w + 4;
    const int ipt = (w + CH_NT - 1) / CH_NT;
    const bool invert = ((A.invert_base ^ var) & 1) != 0;
    // Rows above the first source row stay at the sentinel in the forward sweep; when the leaf's bounding box is
    // known (pipeline mode) the outside transform starts its forward sweep there and the backward sweep reads
    // the sentinel instead of the (unwritten) forward rows.
    int fwd_first = 0;
    if (A.rect_mode && var != 0 && invert) {
        const LgRegion r = A.region[b];
        if (r.ok) fwd_first = min(max(r.y0, 0), h - 1);
    }
    int* ring = smem;              // 3 rows of wpad ints
    int* sbuf = smem + 3 * wpad;   // 2 rows of w ints: sources (pass 0) or forward values (pass 1)
    const size_t fo = (size_t)b * A.P;
    int32_t* fwd = A.fwd + ((size_t)var * A.n + b) * A.P;
    float* of = (var == 0 && A.out_f32) ? A.out_f32 + fo : nullptr;
    uint32_t* oq = (var == 0 && A.out_q16) ? A.out_q16 + fo : nullptr;
    unsigned my_max = 0;

    for (int pass = 0; pass < 2; ++pass) {
        const bool flip = pass == 1;
        const int row_first = flip ? 0 : fwd_first;             // first logical row of this sweep
        const int fwd_rows_valid = h - fwd_first;               // backward sweep: logical rows < this were written
        int xcol[LPT];                                          // physical column of this thread's i-th element
#pragma unroll
        for (int i = 0; i < LPT; ++i) xcol[i] = flip ? (rx1 - 1 - (tid + i * CH_NT)) : (rx0 + tid + i * CH_NT);
        auto rowoff = [&](int ly) -> size_t { return (size_t)(flip ? (ry1 - 1 - ly) : (ry0 + ly)) * W; };
        // fetch only issues the loads (raw label / mask / forward value); nothing may consume them here, or the
        // thread would wait for the memory round trip in every row.  commit() interprets them D rows later.
        auto fetch = [&](int ly, int* dst) {
            const bool row_ok = ly < h && !(flip && ly >= fwd_rows_valid);
            const size_t ro = row_ok ? rowoff(ly) : 0;
#pragma unroll
            for (int i = 0; i < LPT; ++i) {
                int v = 0;
                if (row_ok && tid + i * CH_NT < w) {
                    const size_t p = fo + ro + xcol[i];
                    if (pass == 0) v = A.src.labels ? (int)A.src.labels[p] : (int)A.src.mask[p];
                    else v = fwd[p - fo];
                }
                dst[i] = v;
            }
        };
        auto commit = [&](int ly, const int* srcv, int* dst) {   // registers -> shared row
            const bool unwritten = flip && ly >= fwd_rows_valid;
#pragma unroll
            for (int i = 0; i < LPT; ++i) {
                const int lx = tid + i * CH_NT;
                if (lx < w) {
                    int v = srcv[i];
                    if (pass == 0) v = ((A.src.labels ? (v == id) : (v != 0)) != invert) ? 1 : 0;   // 0 = source pixel
                    else if (unwritten) v = LG_CH_INF;
                    dst[lx] = v;
                }
            }
        };
        auto store_row = [&](int ly, const int* row) {   // row has the +2 border offset
            const size_t ro = rowoff(ly);
#pragma unroll
            for (int i = 0; i < LPT; ++i) {
                const int lx = tid + i * CH_NT;
                if (lx < w) {
                    const int t = row[lx + 2];
                    const size_t p = ro + xcol[i];
                    if (pass == 0) {
                        fwd[p] = t;
                    } else {
                        const unsigned q = (t >= LG_CH_INF) ? LG_CH_DIST_MAX : (unsigned)t;
                        my_max = max(my_max, q);
                        if (oq) oq[p] = q;
                        if (of) of[p] = __fmul_rn((float)q, 1.0f / 65536.0f);
                    }
                }
            }
        };
        int pre[D][LPT];
        for (int i = tid; i < 3 * wpad; i += CH_NT) ring[i] = LG_CH_INF;
        {
            int first[LPT];
            fetch(row_first, first);
#pragma unroll
            for (int d = 1; d <= D; ++d) fetch(row_first + d, pre[d % D]);
            commit(row_first, first, sbuf + (row_first & 1) * w);
        }
        __syncthreads();
        const int xb = tid * ipt;
        for (int ly0 = row_first; ly0 < h; ly0 += D) {
#pragma unroll
            for (int dd = 0; dd < D; ++dd) {
                const int ly = ly0 + dd;
                if (ly >= h) break;
                int* cur = ring + (ly % 3) * wpad;
                const int* p1 = ring + ((ly + 2) % 3) * wpad;
                const int* p2 = ring + ((ly + 1) % 3) * wpad;
                const int* sb = sbuf + (ly & 1) * w;
                if (ly + 1 < h) commit(ly + 1, pre[(dd + 1) % D], sbuf + ((ly + 1) & 1) * w);
                fetch(ly + 1 + D, pre[(dd + 1) % D]);
                if (ly > row_first) store_row(ly - 1, p1);
                // the thread's ipt consecutive pixels share their 5x5 neighbourhood loads
                int a1[LPT + 4], a2[LPT + 2];
#pragma unroll
                for (int j = 0; j < LPT + 4; ++j) { const int xx = xb + j; a1[j] = xx < wpad ? p1[xx] : LG_CH_INF; }
#pragma unroll
                for (int j = 0; j < LPT + 2; ++j) { const int xx = xb + 1 + j; a2[j] = xx < wpad ? p2[xx] : LG_CH_INF; }
                int pv[LPT];
                int run = 0x7FFFFFFF;
#pragma unroll
                for (int i = 0; i < LPT; ++i) {
                    const int x = xb + i;
                    if (i < ipt && x < w) {
                        const int s = sb[x];
                        int m = min(a2[i], a2[i + 2]) + LG_CH_C;
                        m = min(m, min(a1[i], a1[i + 4]) + LG_CH_C);
                        m = min(m, min(a1[i + 1], a1[i + 3]) + LG_CH_B);
                        m = min(m, a1[i + 2] + LG_CH_A);
                        int u;
                        if (pass == 0) u = s ? min(m, LG_CH_INF) : 0;
                        else u = min(s, m);
                        run = min(run, u - LG_CH_A * x);
                    }
                    pv[i] = run;
                }
                int incl = run;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= d) incl = min(incl, t);
                }
                int excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
                if (lane == 0) excl = 0x7FFFFFFF;
                if (lane == 31) wtot[ly & 1][warp] = incl;
                __syncthreads();
                {   // minimum over the preceding warps: one shuffle reduction instead of a serial chain
                    int wv = (lane < warp && lane < CH_NT / 32) ? wtot[ly & 1][lane] : 0x7FFFFFFF;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) wv = min(wv, __shfl_xor_sync(0xFFFFFFFFu, wv, d));
                    excl = min(excl, wv);
                }
#pragma unroll
                for (int i = 0; i < LPT; ++i) {
                    const int x = xb + i;
                    if (i < ipt && x < w) {
                        int t = min(pv[i], excl);
                        // t can only stay at the sentinel when nothing finite precedes x
                        t = (t > LG_CH_INF) ? LG_CH_INF : min(t + LG_CH_A * x, LG_CH_INF);
                        cur[x + 2] = t;
                    }
                }
                __syncthreads();
            }
        }
        store_row(h - 1, ring + ((h - 1) % 3) * wpad);
        __syncthreads();
    }
    if (A.out_max) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) my_max = max(my_max, __shfl_xor_sync(0xFFFFFFFFu, my_max, d));
        if (lane == 0) smax[warp] = my_max;
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < CH_NT / 32; ++k) my_max = max(my_max, smax[k]);
            A.out_max[b * nvar + var] = my_max;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Fast path (row width and rectangle multiples of 8, 16-byte aligned frames): every thread owns 8 consecutive
// pixels of the row.  The thread that fetches a pixel's source flag / forward value is the one that consumes it,
// so the row stream stays in registers (fetched CH8_D rows ahead); the two previous rows are read from the
// shared-memory ring with three 128-bit loads each, the new row is written with four 64-bit stores and goes to
// global memory straight from registers.  ~3x fewer instructions per pixel than the generic kernel above.
// ---------------------------------------------------------------------------------------------------------
constexpr int CH8_D = 4;

template <bool LABELS>
__global__ void __launch_bounds__(512) chamfer8_kernel(ChamferArgs A) {
    extern __shared__ __align__(16) int smem[];
    __shared__ int wtot[2][16];
    __shared__ unsigned smax[16];
    const int b = blockIdx.x, var = blockIdx.y + A.var_base, nvar = A.nvar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int W = A.W, H = A.H;
    if (var == 1 && A.need_full && A.need_full[b] == 0) return;   // outside_max_kernel already produced this frame's maximum
    int rx0 = 0, ry0 = 0, rx1 = W, ry1 = H;
    int id = 1;
    if (LABELS) {
        id = A.src.leaf_id[b];
        if (id < 0) {
            if (A.out_max && tid == 0) A.out_max[b * nvar + var] = 0;
            return;
        }
    }
    const bool invert = ((A.invert_base ^ var) & 1) != 0;
    int fwd_first = 0;
    if (A.rect_mode) {
        const LgRegion r = A.region[b];
        if (var == 0) {
            if (!r.ok) {
                if (A.out_max && tid == 0) A.out_max[b * nvar + var] = 0;
                return;
            }
            rx0 = max(0, r.x0 - 1) & ~7; rx1 = min(W, (r.x1 + 1 + 7) & ~7);
            ry0 = max(0, r.y0 - 1); ry1 = min(H, r.y1 + 1);
        } else if (invert && r.ok) {
            fwd_first = min(max(r.y0, 0), H - 1);     // rows above the first source row stay at the sentinel
        }
    }
    const int w = rx1 - rx0, h = ry1 - ry0;
    if (A.warp_path && A.rect_mode && var == 0 && w <= 512) return;    // chamfer_warp_kernel's
    const int wpad = w + 8;                            // stored position = lx + 2; borders at 0,1 and w+2,w+3
    int* ring = smem;                                  // 3 rows
    const size_t fo = (size_t)b * A.P;
    int32_t* fwd = A.fwd + ((size_t)var * A.n + b) * A.P;
    float* of = (var == 0 && A.out_f32) ? A.out_f32 + fo : nullptr;
    uint32_t* oq = (var == 0 && A.out_q16) ? A.out_q16 + fo : nullptr;
    unsigned my_max = 0;
    const int lx0 = tid * 8;
    const bool active = lx0 < w;

    for (int pass = 0; pass < 2; ++pass) {
        const bool flip = pass == 1;
        const int row_first = flip ? 0 : fwd_first;
        const int fwd_rows_valid = h - fwd_first;
        // physical start of this thread's 8 pixels (ascending addresses; logical order is reversed when flipped)
        const int xphys = flip ? (rx1 - 8 - lx0) : (rx0 + lx0);
        auto rowoff = [&](int ly) -> size_t { return (size_t)(flip ? (ry1 - 1 - ly) : (ry0 + ly)) * W + xphys; };
        uint4 pre[CH8_D][2];
        auto fetch = [&](int ly, uint4* dst) {             // raw loads only; interpreted when the row is computed
            dst[0] = make_uint4(0, 0, 0, 0); dst[1] = dst[0];
            if (!active || ly >= h || (flip && ly >= fwd_rows_valid)) return;
            const size_t p = rowoff(ly);
            if (pass == 0) {
                if (LABELS) dst[0] = *reinterpret_cast<const uint4*>(A.src.labels + fo + p);
                else { const uint2 m = *reinterpret_cast<const uint2*>(A.src.mask + fo + p); dst[0].x = m.x; dst[0].y = m.y; }
            } else {
                dst[0] = *reinterpret_cast<const uint4*>(fwd + p);
                dst[1] = *reinterpret_cast<const uint4*>(fwd + p + 4);
            }
        };
        for (int i = tid; i < 3 * wpad; i += blockDim.x) ring[i] = LG_CH_INF;
#pragma unroll
        for (int d = 0; d < CH8_D; ++d) fetch(row_first + d, pre[d]);
        __syncthreads();
        for (int ly0 = row_first; ly0 < h; ly0 += CH8_D) {
#pragma unroll
            for (int dd = 0; dd < CH8_D; ++dd) {
                const int ly = ly0 + dd;
                if (ly >= h) break;
                int* cur = ring + (ly % 3) * wpad;
                const int* p1 = ring + ((ly + 2) % 3) * wpad;
                const int* p2 = ring + ((ly + 1) % 3) * wpad;
                // this row's source flags (pass 0) or forward values (pass 1), in logical order
                int sv[8];
                {
                    const uint4 r0 = pre[dd][0], r1 = pre[dd][1];
                    if (pass == 0) {
                        int raw[8];
                        if (LABELS) {
                            const unsigned q[4] = {r0.x, r0.y, r0.z, r0.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) { raw[2 * k] = (int16_t)(q[k] & 0xFFFFu); raw[2 * k + 1] = (int16_t)(q[k] >> 16); }
#pragma unroll
                            for (int k = 0; k < 8; ++k) sv[k] = ((raw[k] == id) != invert) ? 1 : 0;
                        } else {
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const unsigned byte = ((k < 4 ? r0.x : r0.y) >> (8 * (k & 3))) & 0xFFu;
                                sv[k] = ((byte != 0) != invert) ? 1 : 0;
                            }
                        }
                    } else {
                        const bool unwritten = ly >= fwd_rows_valid;
                        const unsigned q[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                        for (int k = 0; k < 8; ++k) sv[k] = unwritten ? LG_CH_INF : (int)q[7 - k];   // reversed: logical order
                    }
                }
                fetch(ly + CH8_D, pre[dd]);
                int out[8];
                int run = 0x7FFFFFFF;
                int pv[8];
                if (active) {
                    int a1[12], a2[12];
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const int4 v1 = *reinterpret_cast<const int4*>(p1 + lx0 + 4 * j);
                        const int4 v2 = *reinterpret_cast<const int4*>(p2 + lx0 + 4 * j);
                        a1[4 * j] = v1.x; a1[4 * j + 1] = v1.y; a1[4 * j + 2] = v1.z; a1[4 * j + 3] = v1.w;
                        a2[4 * j] = v2.x; a2[4 * j + 1] = v2.y; a2[4 * j + 2] = v2.z; a2[4 * j + 3] = v2.w;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        // stored index of pixel lx0+i is lx0+i+2; taps at columns -2..+2 are a[i..i+4]
                        int m = min(min(a2[i + 1], a2[i + 3]), min(a1[i], a1[i + 4])) + LG_CH_C;
                        m = min(m, min(a1[i + 1], a1[i + 3]) + LG_CH_B);
                        m = min(m, a1[i + 2] + LG_CH_A);
                        int u;
                        if (pass == 0) u = sv[i] ? min(m, LG_CH_INF) : 0;
                        else u = min(sv[i], m);
                        run = min(run, u - LG_CH_A * (lx0 + i));
                        pv[i] = run;
                    }
                }
                int incl = run;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= d) incl = min(incl, t);
                }
                int excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
                if (lane == 0) excl = 0x7FFFFFFF;
                if (lane == 31) wtot[ly & 1][warp] = incl;
                __syncthreads();
                {
                    int wv = (lane < warp && lane < nwarps) ? wtot[ly & 1][lane] : 0x7FFFFFFF;
#pragma unroll
                    for (int d = 8; d > 0; d >>= 1) wv = min(wv, __shfl_xor_sync(0xFFFFFFFFu, wv, d));
                    wv = min(wv, __shfl_xor_sync(0xFFFFFFFFu, wv, 16));
                    excl = min(excl, wv);
                }
                if (active) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        int t = min(pv[i], excl);
                        t = (t > LG_CH_INF) ? LG_CH_INF : min(t + LG_CH_A * (lx0 + i), LG_CH_INF);
                        out[i] = t;
                    }
                    int2* cw = reinterpret_cast<int2*>(cur + lx0 + 2);
                    cw[0] = make_int2(out[0], out[1]); cw[1] = make_int2(out[2], out[3]);
                    cw[2] = make_int2(out[4], out[5]); cw[3] = make_int2(out[6], out[7]);
                    const size_t p = rowoff(ly);
                    if (pass == 0) {
                        *reinterpret_cast<int4*>(fwd + p) = make_int4(out[0], out[1], out[2], out[3]);
                        *reinterpret_cast<int4*>(fwd + p + 4) = make_int4(out[4], out[5], out[6], out[7]);
                    } else {
                        unsigned q[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            q[i] = (out[7 - i] >= LG_CH_INF) ? LG_CH_DIST_MAX : (unsigned)out[7 - i];   // ascending addresses
                            my_max = max(my_max, q[i]);
                        }
                        if (oq) {
                            *reinterpret_cast<uint4*>(oq + p) = make_uint4(q[0], q[1], q[2], q[3]);
                            *reinterpret_cast<uint4*>(oq + p + 4) = make_uint4(q[4], q[5], q[6], q[7]);
                        }
                        if (of) {
                            const float sc = 1.0f / 65536.0f;
                            *reinterpret_cast<float4*>(of + p) = make_float4(__fmul_rn((float)q[0], sc), __fmul_rn((float)q[1], sc),
                                                                            __fmul_rn((float)q[2], sc), __fmul_rn((float)q[3], sc));
                            *reinterpret_cast<float4*>(of + p + 4) = make_float4(__fmul_rn((float)q[4], sc), __fmul_rn((float)q[5], sc),
                                                                                __fmul_rn((float)q[6], sc), __fmul_rn((float)q[7], sc));
                        }
                    }
                }
                __syncthreads();
            }
        }
        __syncthreads();
    }
    if (A.out_max) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) my_max = max(my_max, __shfl_xor_sync(0xFFFFFFFFu, my_max, d));
        if (lane == 0) smax[warp] = my_max;
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < nwarps; ++k) my_max = max(my_max, smax[k]);
            A.out_max[b * nvar + var] = my_max;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// One warp per leaf rectangle (up to 512 pixels wide: every leaf of a 1440 x 1080 frame).  The sweeps are a chain of
// row steps, and with a CTA per rectangle every step paid two block barriers and a trip through shared memory for a
// row that two warps could hold.  Here a lane owns 16 consecutive pixels and keeps the two previous rows in REGISTERS
// (plus two halo pixels on either side, fetched with four shuffles when the row was produced); the running minimum along
// the row is 16 dependent steps in the lane, a five-step warp scan of the lane totals, and no barrier at all.  Source
// flags / forward values are fetched two rows ahead.  Same arithmetic, same order of minima as chamfer8_kernel: the
// result is bit-identical.
// ---------------------------------------------------------------------------------------------------------
constexpr int CHW_PX = 16;
constexpr int CHW_MAXW = 32 * CHW_PX;
constexpr int CHW_D = 3;            // rows fetched ahead = rows per turn of the unrolled loop

template <bool LABELS>
__global__ void __launch_bounds__(32) chamfer_warp_kernel(ChamferArgs A) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const int W = A.W, H = A.H;
    int id = 1;
    if (LABELS) {
        id = A.src.leaf_id[b];
        if (id < 0) {
            if (A.out_max && lane == 0) A.out_max[b * A.nvar] = 0;
            return;
        }
    }
    const LgRegion r = A.region[b];
    if (!r.ok) {
        if (A.out_max && lane == 0) A.out_max[b * A.nvar] = 0;
        return;
    }
    const bool invert = (A.invert_base & 1) != 0;
    const int rx0 = max(0, r.x0 - 1) & ~7, rx1 = min(W, (r.x1 + 1 + 7) & ~7);
    const int ry0 = max(0, r.y0 - 1), ry1 = min(H, r.y1 + 1);
    const int w = rx1 - rx0, h = ry1 - ry0;
    if (w > CHW_MAXW) return;                           // chamfer8_kernel's
    const size_t fo = (size_t)b * A.P;
    int32_t* fwd = A.fwd + (size_t)b * A.P;             // variant 0
    float* of = A.out_f32 ? A.out_f32 + fo : nullptr;
    uint32_t* oq = A.out_q16 ? A.out_q16 + fo : nullptr;
    unsigned my_max = 0;
    const int lx0 = lane * CHW_PX;
    const int nv = max(0, min(CHW_PX, w - lx0));        // valid pixels of this lane: 16, 8 or 0 (w is a multiple of 8)

    for (int pass = 0; pass < 2; ++pass) {
        const bool flip = pass == 1;
        // physical start of this lane's 16 pixels (ascending addresses; logical order is reversed when flipped)
        const int xphys = flip ? (rx1 - CHW_PX - lx0) : (rx0 + lx0);
        auto rowoff = [&](int ly) -> long long { return (long long)(flip ? (ry1 - 1 - ly) : (ry0 + ly)) * W + xphys; };
        // a flipped lane with 8 valid pixels owns the UPPER half of its 16 physical positions
        uint4 pre[CHW_D][4];
        auto fetch = [&](int ly, uint4* dst) {
#pragma unroll
            for (int k = 0; k < 4; ++k) dst[k] = make_uint4(0, 0, 0, 0);
            if (nv == 0 || ly >= h) return;
            const long long p = rowoff(ly);
            if (pass == 0) {
                if (LABELS) {
                    dst[0] = *reinterpret_cast<const uint4*>(A.src.labels + fo + p);
                    if (nv > 8) dst[1] = *reinterpret_cast<const uint4*>(A.src.labels + fo + p + 8);
                } else {
                    const uint2 m0 = *reinterpret_cast<const uint2*>(A.src.mask + fo + p);
                    dst[0].x = m0.x; dst[0].y = m0.y;
                    if (nv > 8) { const uint2 m1 = *reinterpret_cast<const uint2*>(A.src.mask + fo + p + 8); dst[0].z = m1.x; dst[0].w = m1.y; }
                }
            } else {
                if (nv > 8) {
                    dst[0] = *reinterpret_cast<const uint4*>(fwd + p);
                    dst[1] = *reinterpret_cast<const uint4*>(fwd + p + 4);
                }
                dst[2] = *reinterpret_cast<const uint4*>(fwd + p + 8);
                dst[3] = *reinterpret_cast<const uint4*>(fwd + p + 12);
            }
        };
        // Three row buffers with halos (index i + 2 = logical pixel lx0 + i, i in [-2, 18)) rotate through the roles
        // "row before last", "last row", "this row"; the row loop is unrolled by three so that the rotation costs nothing.
        int ra[CHW_PX + 4], rb[CHW_PX + 4], rc[CHW_PX + 4];
#pragma unroll
        for (int i = 0; i < CHW_PX + 4; ++i) { ra[i] = LG_CH_INF; rb[i] = LG_CH_INF; rc[i] = LG_CH_INF; }
#pragma unroll
        for (int d = 0; d < CHW_D; ++d) fetch(d, pre[d]);
        auto row_step = [&](int ly, uint4* mine, const int (&p1)[CHW_PX + 4], const int (&p2)[CHW_PX + 4], int (&cur)[CHW_PX + 4]) {
            // this row's source flags (pass 0) or forward values (pass 1), in logical order
            int sv[CHW_PX];
            {
                const unsigned q[16] = {mine[0].x, mine[0].y, mine[0].z, mine[0].w, mine[1].x, mine[1].y, mine[1].z, mine[1].w,
                                        mine[2].x, mine[2].y, mine[2].z, mine[2].w, mine[3].x, mine[3].y, mine[3].z, mine[3].w};
                if (pass == 0) {
#pragma unroll
                    for (int k = 0; k < CHW_PX; ++k) {
                        bool leaf;
                        if (LABELS) leaf = (int)(int16_t)((q[k >> 1] >> (16 * (k & 1))) & 0xFFFFu) == id;
                        else leaf = ((q[k >> 2] >> (8 * (k & 3))) & 0xFFu) != 0;
                        sv[k] = (leaf != invert) ? 1 : 0;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < CHW_PX; ++k) sv[k] = (int)q[15 - k];     // reversed: logical order
                }
            }
            fetch(ly + CHW_D, mine);
            int pv[CHW_PX];
            int run = 0x7FFFFFFF;
#pragma unroll
            for (int i = 0; i < CHW_PX; ++i) {
                // taps at columns -2..+2 of pixel i are p[i..i+4]
                int m = min(min(p2[i + 1], p2[i + 3]), min(p1[i], p1[i + 4])) + LG_CH_C;
                m = min(m, min(p1[i + 1], p1[i + 3]) + LG_CH_B);
                m = min(m, p1[i + 2] + LG_CH_A);
                int u;
                if (pass == 0) u = sv[i] ? min(m, LG_CH_INF) : 0;
                else u = min(sv[i], m);
                if (i >= nv) u = LG_CH_INF;                                      // beyond the rectangle
                run = min(run, u - LG_CH_A * (lx0 + i));
                pv[i] = run;
            }
            int incl = run;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl = min(incl, t);
            }
            int excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
            if (lane == 0) excl = 0x7FFFFFFF;
#pragma unroll
            for (int i = 0; i < CHW_PX; ++i) {
                int t = min(pv[i], excl);
                t = (t > LG_CH_INF) ? LG_CH_INF : min(t + LG_CH_A * (lx0 + i), LG_CH_INF);
                cur[i + 2] = i < nv ? t : LG_CH_INF;
            }
            {   // the new row's halos come from the neighbouring lanes
                const int l14 = __shfl_up_sync(0xFFFFFFFFu, cur[CHW_PX], 1), l15 = __shfl_up_sync(0xFFFFFFFFu, cur[CHW_PX + 1], 1);
                const int r0 = __shfl_down_sync(0xFFFFFFFFu, cur[2], 1), r1 = __shfl_down_sync(0xFFFFFFFFu, cur[3], 1);
                cur[0] = lane == 0 ? LG_CH_INF : l14; cur[1] = lane == 0 ? LG_CH_INF : l15;
                cur[CHW_PX + 2] = lane == 31 ? LG_CH_INF : r0; cur[CHW_PX + 3] = lane == 31 ? LG_CH_INF : r1;
            }
            if (nv > 0) {
                const long long p = rowoff(ly);
                if (pass == 0) {
                    *reinterpret_cast<int4*>(fwd + p) = make_int4(cur[2], cur[3], cur[4], cur[5]);
                    *reinterpret_cast<int4*>(fwd + p + 4) = make_int4(cur[6], cur[7], cur[8], cur[9]);
                    if (nv > 8) {
                        *reinterpret_cast<int4*>(fwd + p + 8) = make_int4(cur[10], cur[11], cur[12], cur[13]);
                        *reinterpret_cast<int4*>(fwd + p + 12) = make_int4(cur[14], cur[15], cur[16], cur[17]);
                    }
                } else {
                    unsigned q[CHW_PX];
#pragma unroll
                    for (int i = 0; i < CHW_PX; ++i) {                          // ascending addresses
                        const int o = cur[2 + 15 - i];
                        q[i] = (o >= LG_CH_INF) ? LG_CH_DIST_MAX : (unsigned)o;
                        if (15 - i < nv) my_max = max(my_max, q[i]);
                    }
                    const float sc = 1.0f / 65536.0f;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (nv <= 8 && g < 2) continue;                         // the lower half lies outside the rectangle
                        if (oq) *reinterpret_cast<uint4*>(oq + p + 4 * g) = make_uint4(q[4 * g], q[4 * g + 1], q[4 * g + 2], q[4 * g + 3]);
                        if (of) *reinterpret_cast<float4*>(of + p + 4 * g) =
                            make_float4(__fmul_rn((float)q[4 * g], sc), __fmul_rn((float)q[4 * g + 1], sc),
                                        __fmul_rn((float)q[4 * g + 2], sc), __fmul_rn((float)q[4 * g + 3], sc));
                    }
                }
            }
        };
        for (int ly0 = 0; ly0 < h; ly0 += 3) {
            row_step(ly0, pre[0], rc, rb, ra);                  // (last row, row before last) -> this row
            if (ly0 + 1 < h) row_step(ly0 + 1, pre[1], ra, rc, rb);
            if (ly0 + 2 < h) row_step(ly0 + 2, pre[2], rb, ra, rc);
        }
        __syncwarp();
        __threadfence_block();          // pass 1 reads the forward field other lanes wrote
    }
    if (A.out_max) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) my_max = max(my_max, __shfl_xor_sync(0xFFFFFFFFu, my_max, d));
        if (lane == 0) A.out_max[b * A.nvar] = my_max;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Maximum of the OUTSIDE transform without running it.
//
// The two-sweep 5x5 chamfer transform computes, for every pixel p, min over source pixels q of the chamfer norm
// N(p - q) (cost of the cheapest path of a/b/c moves; tests/test_oracle_golden.py holds the integer identity on
// adversarial masks), and N satisfies the triangle inequality in integers.  The stage only needs max over p of that
// distance to the leaf (the sdf normalisation, grasp_point_selector.py:531-532), so it is found by branch and bound:
// the distance at the centre c of a cell is exact (minimum over the leaf's boundary pixels), every pixel of the
// cell is at most d(c) + N(w/2, h/2) away, and cells whose bound cannot beat the best exact value seen so far are
// dropped; survivors are split 4 x 4 until they are single pixels.  Exact by construction; ~10^6 norm evaluations
// per frame instead of two sweeps over the whole frame with 2 H dependent row steps.  Frames whose boundary or
// cell lists overflow fall back to the sweeps (need_full).
// ---------------------------------------------------------------------------------------------------------
constexpr int BND_NT = 256;
__global__ void __launch_bounds__(BND_NT) leaf_boundary_kernel(lg_context c, LgMaskSrc src) {
    const int b = blockIdx.y;
    const LgRegion r = c.region[b];
    if (!r.ok) return;
    const int W = c.W, H = c.H, id = src.id(b);
    const size_t fo = (size_t)b * c.P;
    const int lane = threadIdx.x & 31;
    uint32_t* list = c.bnd_list + (size_t)b * LG_BND_CAP;
    for (int y = r.y0 + blockIdx.x; y < r.y1; y += gridDim.x) {
        for (int xb = r.x0; xb < r.x1; xb += BND_NT) {
            const int x = xb + threadIdx.x;
            bool hit = false;
            if (x < r.x1 && src.at(fo, (size_t)y * W + x, id)) {
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int xx = x + dx, yy = y + dy;
                        if ((dx | dy) != 0 && xx >= 0 && xx < W && yy >= 0 && yy < H && !src.at(fo, (size_t)yy * W + xx, id)) hit = true;
                    }
            }
            const unsigned ball = __ballot_sync(0xFFFFFFFFu, hit);
            if (ball) {
                unsigned base = 0;
                if (lane == (__ffs(ball) - 1)) base = atomicAdd(&c.bnd_count[b], __popc(ball));
                base = __shfl_sync(0xFFFFFFFFu, base, __ffs(ball) - 1);
                if (hit) {
                    const unsigned pos = base + __popc(ball & ((1u << lane) - 1u));
                    if (pos < LG_BND_CAP) list[pos] = (unsigned)x | ((unsigned)y << 16);
                }
            }
        }
    }
}

__device__ __forceinline__ int chamfer_norm_q16(int dx, int dy) {
    dx = abs(dx); dy = abs(dy);
    const int mx = max(dx, dy), mn = min(dx, dy);
    // the two linear pieces of the norm; it is convex, so it is their maximum
    return max(LG_CH_A * mx + (LG_CH_C - 2 * LG_CH_A) * mn, (LG_CH_C - LG_CH_B) * mx + (2 * LG_CH_B - LG_CH_C) * mn);
}

constexpr int OM_NT = 512;
constexpr int OM_CELLS = 1024;
constexpr int OM_STAGE = 6144;
__global__ void __launch_bounds__(OM_NT) outside_max_kernel(lg_context c) {
    __shared__ unsigned s_bnd[OM_STAGE];
    __shared__ unsigned s_cell[2][OM_CELLS][2];      // (x0 | y0 << 16, w | h << 16)
    __shared__ int s_val[OM_CELLS];
    __shared__ int s_cnt[2], s_lb, s_over;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const LgRegion r = c.region[b];
    const unsigned Q = r.ok ? c.bnd_count[b] : 0u;
    if (!r.ok || Q == 0u) {                          // no leaf, or the leaf covers the frame: the transform is zero
        if (tid == 0) { c.dt_max[b * 2 + 1] = 0; c.need_full[b] = 0; }
        return;
    }
    if (Q > (unsigned)OM_STAGE || Q > (unsigned)LG_BND_CAP) {
        if (tid == 0) c.need_full[b] = 1;
        return;
    }
    const int W = c.W, H = c.H;
    const uint32_t* list = c.bnd_list + (size_t)b * LG_BND_CAP;
    for (unsigned i = tid; i < Q; i += OM_NT) s_bnd[i] = list[i];
    int G = 64;
    while (((W + G - 1) / G) * ((H + G - 1) / G) > OM_CELLS / 2) G *= 2;
    const int ncx = (W + G - 1) / G, ncy = (H + G - 1) / G;
    for (int i = tid; i < ncx * ncy; i += OM_NT) {
        const int x0 = (i % ncx) * G, y0 = (i / ncx) * G;
        s_cell[0][i][0] = (unsigned)x0 | ((unsigned)y0 << 16);
        s_cell[0][i][1] = (unsigned)min(G, W - x0) | ((unsigned)min(G, H - y0) << 16);
    }
    if (tid == 0) { s_cnt[0] = ncx * ncy; s_cnt[1] = 0; s_lb = 0; s_over = 0; }
    __syncthreads();
    int cur = 0;
    while (true) {
        const int ncell = s_cnt[cur];
        // exact distance at every cell centre: one warp per cell, lanes over the boundary pixels
        for (int i = warp; i < ncell; i += OM_NT / 32) {
            const unsigned c0 = s_cell[cur][i][0], c1 = s_cell[cur][i][1];
            const int cx = (int)(c0 & 0xFFFFu) + (int)(c1 & 0xFFFFu) / 2, cy = (int)(c0 >> 16) + (int)(c1 >> 16) / 2;
            int dmin = 0x7FFFFFFF;
            for (unsigned q = lane; q < Q; q += 32) {
                const unsigned v = s_bnd[q];
                dmin = min(dmin, chamfer_norm_q16(cx - (int)(v & 0xFFFFu), cy - (int)(v >> 16)));
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) dmin = min(dmin, __shfl_xor_sync(0xFFFFFFFFu, dmin, d));
            if (lane == 0) { s_val[i] = dmin; atomicMax(&s_lb, dmin); }
        }
        __syncthreads();
        const int lb = s_lb;
        const int nxt = cur ^ 1;
        for (int i = tid; i < ncell; i += OM_NT) {
            const unsigned c0 = s_cell[cur][i][0], c1 = s_cell[cur][i][1];
            const int x0 = (int)(c0 & 0xFFFFu), y0 = (int)(c0 >> 16), w = (int)(c1 & 0xFFFFu), h = (int)(c1 >> 16);
            if (w == 1 && h == 1) continue;                                  // exact already
            if (s_val[i] + chamfer_norm_q16(w / 2, h / 2) <= lb) continue;   // nothing in the cell can exceed the best
            const int kx = min(4, w), ky = min(4, h);
            const int pos = atomicAdd(&s_cnt[nxt], kx * ky);
            if (pos + kx * ky > OM_CELLS) { s_over = 1; continue; }
            for (int j = 0; j < ky; ++j)
                for (int k = 0; k < kx; ++k) {
                    const int xa = x0 + (w * k) / kx, xb = x0 + (w * (k + 1)) / kx;
                    const int ya = y0 + (h * j) / ky, yb = y0 + (h * (j + 1)) / ky;
                    s_cell[nxt][pos + j * kx + k][0] = (unsigned)xa | ((unsigned)ya << 16);
                    s_cell[nxt][pos + j * kx + k][1] = (unsigned)(xb - xa) | ((unsigned)(yb - ya) << 16);
                }
        }
        __syncthreads();
        if (s_over) {
            if (tid == 0) c.need_full[b] = 1;
            return;
        }
        if (s_cnt[nxt] == 0) break;
        __syncthreads();
        if (tid == 0) s_cnt[cur] = 0;
        cur = nxt;
        __syncthreads();
    }
    if (tid == 0) { c.dt_max[b * 2 + 1] = (uint32_t)s_lb; c.need_full[b] = 0; }
}

}  // namespace

int lg_run_chamfer(lg_context* c, LgMaskSrc src, int n, int rect_mode, int invert_base, int nvar, int var_first, int var_count,
                   float* out0, uint32_t* q0, uint32_t* out_max, const uint32_t* need_full, cudaStream_t st) {
    if (c->W > CH_NT * CH_MAXI) {
        lg_set_error("chamfer transform supports widths up to %d", CH_NT * CH_MAXI);
        return LG_E_ARG;
    }
    ChamferArgs A;
    A.src = src; A.region = c->region; A.n = n; A.H = c->H; A.W = c->W; A.rect_mode = rect_mode;
    A.invert_base = invert_base; A.fwd = c->dt_fwd; A.out_f32 = out0; A.out_q16 = q0; A.out_max = out_max; A.P = c->P;
    A.need_full = need_full; A.warp_path = 0; A.nvar = nvar; A.var_base = var_first;
    // fast path: 8 pixels per thread, vector loads/stores -> needs 16-byte aligned rows
    const bool aligned = (c->W % 8 == 0) &&
                         ((reinterpret_cast<uintptr_t>(src.labels) | reinterpret_cast<uintptr_t>(src.mask) |
                           reinterpret_cast<uintptr_t>(out0) | reinterpret_cast<uintptr_t>(q0)) % 16 == 0) &&
                         (src.labels || c->P % 8 == 0);
    static const bool force_generic = getenv("LG_CHAMFER_GENERIC") != nullptr;
    if (aligned && !force_generic) {
        static const bool no_warp = getenv("LG_CHAMFER_NO_WARP") != nullptr;    // A/B switch
        // Frames wider than 2048 px (4K) have many leaves wider than one warp's 512 px; those would run in chamfer8_kernel
        // AFTER the warp kernel on the same stream (measured at 3840 x 2160: 1.5 -> 2.1 ms per 16 frames), so there every
        // rectangle stays with chamfer8_kernel.
        A.warp_path = (rect_mode && !no_warp && c->W <= 2048) ? 1 : 0;
        if (A.warp_path && var_first == 0) {      // the leaf rectangles that one warp can hold (all of them at 1440 x 1080)
            // The kernel uses no shared memory, but it runs beside kernels that do (orientation, outside maximum): ask for
            // the large carve-out so that an SM it occupies does not have to drain before it can take their CTAs.
            LG_PREFER_LARGE_SMEM(chamfer_warp_kernel<true>);
            LG_PREFER_LARGE_SMEM(chamfer_warp_kernel<false>);
            if (src.labels) chamfer_warp_kernel<true><<<n, 32, 0, st>>>(A);
            else chamfer_warp_kernel<false><<<n, 32, 0, st>>>(A);
            LG_LAUNCH_CHECK();
        }
        const int threads = ((c->W / 8 + 31) / 32) * 32;
        const size_t sm8 = (size_t)3 * (c->W + 8) * sizeof(int);
        LG_ENSURE_SMEM(chamfer8_kernel<true>, sm8);
        LG_ENSURE_SMEM(chamfer8_kernel<false>, sm8);
        if (src.labels) chamfer8_kernel<true><<<dim3(n, var_count), threads, sm8, st>>>(A);
        else chamfer8_kernel<false><<<dim3(n, var_count), threads, sm8, st>>>(A);
        LG_LAUNCH_CHECK();
        return LG_OK;
    }
    size_t sm = (size_t)(3 * (c->W + 4) + 2 * c->W) * sizeof(int);
    LG_ENSURE_SMEM((chamfer_kernel<3, 4>), sm);
    LG_ENSURE_SMEM((chamfer_kernel<CH_MAXI, 2>), sm);
    if (c->W <= 3 * CH_NT) chamfer_kernel<3, 4><<<dim3(n, var_count), CH_NT, sm, st>>>(A);
    else chamfer_kernel<CH_MAXI, 2><<<dim3(n, var_count), CH_NT, sm, st>>>(A);
    LG_LAUNCH_CHECK();
    return LG_OK;
}

// maximum of the outside transform by branch and bound (frames that overflow its lists are flagged in need_full
// and handled by variant 1 of the sweep kernels)
int lg_run_outside_max(lg_context* c, LgMaskSrc src, int n, cudaStream_t st) {
    LG_CUDA(cudaMemsetAsync(c->bnd_count, 0, sizeof(uint32_t) * n, st));
    LG_PREFER_LARGE_SMEM(leaf_boundary_kernel);
    LG_PREFER_LARGE_SMEM(outside_max_kernel);
    leaf_boundary_kernel<<<dim3(16, n), BND_NT, 0, st>>>(*c, src);
    LG_LAUNCH_CHECK();
    outside_max_kernel<<<n, OM_NT, 0, st>>>(*c);
    LG_LAUNCH_CHECK();
    return LG_OK;
}

extern "C" int lg_chamfer_transform(lg_context* c, const uint8_t* mask, int n, int invert, float* dist,
                                    uint32_t* q16, uint32_t* max_q16, void* stream) {
    if (!c || !mask || n < 1) return LG_E_ARG;
    if (n > c->B) return LG_E_CAPACITY;
    LgMaskSrc src{nullptr, mask, nullptr};
    return lg_run_chamfer(c, src, n, 0, invert ? 1 : 0, 1, 0, 1, dist, q16, max_q16, nullptr, (cudaStream_t)stream);
}
