// GraspPointCNN forward, eval mode (reference scripts/utils/ml_grasp_optimizer/model.py:102-128):
//   3 x [conv3x3-BN-ReLU, conv3x3-BN-ReLU, maxpool2]  ->  x * sigmoid(conv1x1)  ->  global average
//   ->  Linear-BN-ReLU x3  ->  Linear(64, 1).      Dropout is the identity in eval mode; BatchNorm is
//   folded into the preceding conv / linear on the host (cnn.py:fold_batchnorm).
//
// fp32 path (this file, part 1): direct convolution as a register-tiled implicit GEMM on CUDA cores,
// NHWC activations, weights [ky][kx][Cin][Cout]; max-pool fused into the second conv of each block.
// It is the parity anchor (1e-5 against torch fp32 on the CPU).
//
// Weight blob (float32), in order:
//   for layer l in 0..5:  w[3][3][Cin_l][Cout_l], b[Cout_l]       Cin/Cout = 9/64, 64/64, 64/128, 128/128, 128/256, 256/256
//   attention: w[256], b[1]
//   fc0: w[256][256] (in-major: w[i][o]), b[256];  fc1: w[256][128], b[128];  fc2: w[128][64], b[64];  fc3: w[64], b[1]
#include "lg_internal.cuh"

namespace {

constexpr int CV_NT = 128;      // 16 pixel quads x 8 channel groups
constexpr int CV_KC = 16;       // input channels per shared-memory stage
constexpr int CV_CO = 64;       // output channels per CTA

struct ConvArgs {
    const float* in;
    long long sn, sy, sx, sc;   // input strides in elements (n, y, x, channel)
    int H, W, Cin, Cout;
    const float* wt;            // [9][Cin][Cout]
    const float* bias;
    float* out;                 // NHWC, pooled when POOL
};

template <bool POOL>
__global__ void __launch_bounds__(CV_NT) conv3x3_relu_kernel(ConvArgs A) {
    __shared__ float s_in[CV_KC][10][12];            // 8x8 tile + halo, row padded to 12
    __shared__ __align__(16) float s_w[9][CV_KC][CV_CO];
    const int tid = threadIdx.x;
    const int tiles_x = (A.W + 7) / 8;
    const int ty0 = (blockIdx.x / tiles_x) * 8, tx0 = (blockIdx.x % tiles_x) * 8;
    const int co0 = blockIdx.y * CV_CO;
    const int n = blockIdx.z;
    const int quad = tid & 15, cg = tid >> 4;        // quad: 4x4 grid of 2x2 pixel quads; cg: 8 channels
    const int qy = (quad >> 2) * 2, qx = (quad & 3) * 2;
    float acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[p][k] = 0.f;
    const float* inb = A.in + (long long)n * A.sn;
    for (int c0 = 0; c0 < A.Cin; c0 += CV_KC) {
        const int kc = min(CV_KC, A.Cin - c0);
        __syncthreads();
        for (int i = tid; i < CV_KC * 100; i += CV_NT) {
            const int ci = i / 100, rem = i - ci * 100, ly = rem / 10, lx = rem - ly * 10;
            const int y = ty0 - 1 + ly, x = tx0 - 1 + lx;
            float v = 0.f;
            if (ci < kc && y >= 0 && y < A.H && x >= 0 && x < A.W) v = inb[(long long)y * A.sy + (long long)x * A.sx + (long long)(c0 + ci) * A.sc];
            s_in[ci][ly][lx] = v;
        }
        for (int i = tid; i < 9 * CV_KC * CV_CO; i += CV_NT) {
            const int t = i / (CV_KC * CV_CO), rem = i - t * (CV_KC * CV_CO), ci = rem / CV_CO, co = rem - ci * CV_CO;
            float v = 0.f;
            if (ci < kc && co0 + co < A.Cout) v = A.wt[((long long)t * A.Cin + (c0 + ci)) * A.Cout + co0 + co];
            s_w[t][ci][co] = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int t = 0; t < 9; ++t) {
            const int ky = t / 3, kx = t - ky * 3;
#pragma unroll 4
            for (int ci = 0; ci < CV_KC; ++ci) {
                const float i00 = s_in[ci][qy + ky][qx + kx], i01 = s_in[ci][qy + ky][qx + kx + 1];
                const float i10 = s_in[ci][qy + ky + 1][qx + kx], i11 = s_in[ci][qy + ky + 1][qx + kx + 1];
                const float4 w0 = *reinterpret_cast<const float4*>(&s_w[t][ci][cg * 8]);
                const float4 w1 = *reinterpret_cast<const float4*>(&s_w[t][ci][cg * 8 + 4]);
                const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    acc[0][k] = fmaf(i00, w[k], acc[0][k]);
                    acc[1][k] = fmaf(i01, w[k], acc[1][k]);
                    acc[2][k] = fmaf(i10, w[k], acc[2][k]);
                    acc[3][k] = fmaf(i11, w[k], acc[3][k]);
                }
            }
        }
    }
    const int co = co0 + cg * 8;
    if (co >= A.Cout || ty0 + qy >= A.H || tx0 + qx >= A.W) return;     // partial channel tile / tile beyond a small image
    if (POOL) {
        const int Ho = A.H / 2, Wo = A.W / 2;
        const int oy = (ty0 + qy) / 2, ox = (tx0 + qx) / 2;
        float* o = A.out + (((long long)n * Ho + oy) * Wo + ox) * A.Cout + co;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float bsv = A.bias[co + k];
            float m = fmaxf(fmaxf(acc[0][k], acc[1][k]), fmaxf(acc[2][k], acc[3][k]));
            o[k] = fmaxf(m + bsv, 0.f);
        }
    } else {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int y = ty0 + qy + (p >> 1), x = tx0 + qx + (p & 1);
            float* o = A.out + (((long long)n * A.H + y) * A.W + x) * A.Cout + co;
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = fmaxf(acc[p][k] + A.bias[co + k], 0.f);
        }
    }
}

// attention + global average + MLP.  in: NHWC [n][4][4][256].  One CTA handles TL_PB patches so that every MLP weight
// fetched from L2 is used TL_PB times: the attention-pooled vectors of the patches are formed one after the other,
// then the four linear layers run for all of them together.
constexpr int TL_NT = 256;
constexpr int TL_PB = 8;
// act != nullptr: the features come straight from the last tensor-core convolution instead - bf16, plane-major with
// halo (8x8 pixels per patch on a 9x9 grid, `lead` zero rows in front, act_rows rows per plane of 8 channels) - and the
// final 2x2 max-pool is taken while they are loaded.
__global__ void __launch_bounds__(TL_NT) cnn_tail_kernel(const float* __restrict__ feat, const uint4* __restrict__ act,
                                                          long long act_rows, int lead, const float* __restrict__ blob_tail,
                                                          float* __restrict__ logits, int n_host, const int32_t* __restrict__ n_dev) {
    const int n = n_dev ? min(*n_dev, n_host) : n_host;
    if ((int)blockIdx.x * TL_PB >= n) return;
    __shared__ float s_f[16][256];
    __shared__ float s_att[16];
    __shared__ float s_a[TL_PB][256], s_b[TL_PB][256];
    const int n0 = blockIdx.x * TL_PB, tid = threadIdx.x;
    const int np = min(TL_PB, n - n0);
    const float* aw = blob_tail;              // 256
    const float* ab = aw + 256;               // 1
    const float* w0 = ab + 1;                 // 256x256
    const float* b0 = w0 + 256 * 256;
    const float* w1 = b0 + 256;               // 256x128
    const float* b1 = w1 + 256 * 128;
    const float* w2 = b1 + 128;               // 128x64
    const float* b2 = w2 + 128 * 64;
    const float* w3 = b2 + 64;                // 64
    const float* b3 = w3 + 64;
    for (int p = 0; p < TL_PB; ++p) {
        if (p >= np) { s_a[p][tid] = 0.f; continue; }
        __syncthreads();
        if (act) {
            for (int i = tid; i < 16 * 32; i += TL_NT) {            // (pooled pixel, plane of 8 channels)
                const int px = i >> 5, pl = i & 31;
                const uint4* src = act + (long long)pl * act_rows + lead + (long long)(n0 + p) * 81 + (2 * (px >> 2) + 1) * 9 + 2 * (px & 3) + 1;
                const uint4 q[4] = {src[0], src[1], src[9], src[10]};
                float m[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) m[e] = 0.f;              // post-ReLU values: 0 is the identity of max
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t w[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        m[2 * e] = fmaxf(m[2 * e], __uint_as_float(w[e] << 16));
                        m[2 * e + 1] = fmaxf(m[2 * e + 1], __uint_as_float(w[e] & 0xFFFF0000u));
                    }
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) s_f[px][pl * 8 + e] = m[e];
            }
        } else {
            const float* f = feat + (size_t)(n0 + p) * 16 * 256;
            for (int i = tid; i < 16 * 256; i += TL_NT) s_f[i >> 8][i & 255] = f[i];
        }
        __syncthreads();
        {   // attention logit per pixel: 16 pixels x 256 channels, 16 threads per pixel
            const int px = tid >> 4, l = tid & 15;
            float s = 0.f;
            for (int ch = l; ch < 256; ch += 16) s = fmaf(s_f[px][ch], aw[ch], s);
#pragma unroll
            for (int d = 8; d > 0; d >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
            if (l == 0) s_att[px] = 1.f / (1.f + expf(-(s + ab[0])));
        }
        __syncthreads();
        float s = 0.f;
#pragma unroll
        for (int px = 0; px < 16; ++px) s = fmaf(s_f[px][tid], s_att[px], s);
        s_a[p][tid] = s * (1.f / 16.f);
    }
    __syncthreads();
    {
        float acc[TL_PB];
#pragma unroll
        for (int p = 0; p < TL_PB; ++p) acc[p] = b0[tid];
#pragma unroll 16      // sixteen weight rows in flight: the loop is a chain of L2 round trips otherwise
        for (int i = 0; i < 256; ++i) {
            const float w = w0[i * 256 + tid];
#pragma unroll
            for (int p = 0; p < TL_PB; ++p) acc[p] = fmaf(s_a[p][i], w, acc[p]);
        }
#pragma unroll
        for (int p = 0; p < TL_PB; ++p) s_b[p][tid] = fmaxf(acc[p], 0.f);
    }
    __syncthreads();
    if (tid < 128) {
        float acc[TL_PB];
#pragma unroll
        for (int p = 0; p < TL_PB; ++p) acc[p] = b1[tid];
#pragma unroll 16      // sixteen weight rows in flight: the loop is a chain of L2 round trips otherwise
        for (int i = 0; i < 256; ++i) {
            const float w = w1[i * 128 + tid];
#pragma unroll
            for (int p = 0; p < TL_PB; ++p) acc[p] = fmaf(s_b[p][i], w, acc[p]);
        }
#pragma unroll
        for (int p = 0; p < TL_PB; ++p) s_a[p][tid] = fmaxf(acc[p], 0.f);
    }
    __syncthreads();
    if (tid < 64) {
        float acc[TL_PB];
#pragma unroll
        for (int p = 0; p < TL_PB; ++p) acc[p] = b2[tid];
#pragma unroll 16      // sixteen weight rows in flight: the loop is a chain of L2 round trips otherwise
        for (int i = 0; i < 128; ++i) {
            const float w = w2[i * 64 + tid];
#pragma unroll
            for (int p = 0; p < TL_PB; ++p) acc[p] = fmaf(s_a[p][i], w, acc[p]);
        }
#pragma unroll
        for (int p = 0; p < TL_PB; ++p) s_b[p][tid] = fmaxf(acc[p], 0.f);
    }
    __syncthreads();
    {   // final 64 -> 1: warp p handles patch p
        const int p = tid >> 5, lane = tid & 31;
        if (p < np) {
            float s = s_b[p][lane] * w3[lane] + s_b[p][lane + 32] * w3[lane + 32];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
            if (lane == 0) logits[n0 + p] = s + b3[0];
        }
    }
}

// Attention + global average + classifier for any architecture GraspPointCNN can take (model.py:30-86): C final
// channels on S2 = s x s positions, attention none / spatial / channel / hybrid.  One CTA per patch, features in
// shared memory.  Blob: [spatial: w[C], b] [channel: w1[C][C/16], b1[C/16], w2[C/16][C], b2[C]] fc0 w[C][C], b; fc1
// w[C][C/2], b; fc2 w[C/2][C/4], b; fc3 w[C/4], b (BatchNorm1d folded, weights in-major).
__global__ void __launch_bounds__(256) cnn_tail_generic_kernel(const float* __restrict__ feat, const float* __restrict__ blob,
                                                                float* __restrict__ logits, int C, int S2, int attention) {
    extern __shared__ float sm[];
    float* f = sm;                    // [S2][C]
    float* att_s = f + S2 * C;        // [S2]
    float* att_c = att_s + S2;        // [C]
    float* va = att_c + C;            // [C]
    float* vb = va + C;               // [C]
    const int n = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const float* src = feat + (size_t)n * S2 * C;
    for (int i = tid; i < S2 * C; i += NT) f[i] = src[i];
    for (int i = tid; i < S2; i += NT) att_s[i] = 1.f;
    for (int i = tid; i < C; i += NT) att_c[i] = 1.f;
    __syncthreads();
    const float* w = blob;
    const bool spatial = attention == 1 || attention == 3, channel = attention == 2 || attention == 3;
    if (spatial) {
        const int lane = tid & 31, warp = tid >> 5;
        for (int p = warp; p < S2; p += NT / 32) {
            float a = 0.f;
            for (int ch = lane; ch < C; ch += 32) a = fmaf(f[p * C + ch], w[ch], a);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xFFFFFFFFu, a, d);
            if (lane == 0) att_s[p] = 1.f / (1.f + expf(-(a + w[C])));
        }
        w += C + 1;
    }
    if (channel) {
        const int R = C / 16;
        const float *w1 = w, *b1 = w1 + C * R, *w2 = b1 + R, *b2 = w2 + R * C;
        for (int ch = tid; ch < C; ch += NT) {
            float a = 0.f;
            for (int p = 0; p < S2; ++p) a += f[p * C + ch];
            va[ch] = a / (float)S2;                 // AdaptiveAvgPool2d(1) of the un-attended features
        }
        __syncthreads();
        for (int j = tid; j < R; j += NT) {
            float a = b1[j];
            for (int ch = 0; ch < C; ++ch) a = fmaf(va[ch], w1[ch * R + j], a);
            vb[j] = fmaxf(a, 0.f);
        }
        __syncthreads();
        for (int ch = tid; ch < C; ch += NT) {
            float a = b2[ch];
            for (int j = 0; j < R; ++j) a = fmaf(vb[j], w2[j * C + ch], a);
            att_c[ch] = 1.f / (1.f + expf(-a));
        }
        w = b2 + C;
    }
    __syncthreads();
    for (int ch = tid; ch < C; ch += NT) {          // x * attention, then global average
        float a = 0.f;
        for (int p = 0; p < S2; ++p) a = fmaf(f[p * C + ch] * att_c[ch], att_s[p], a);
        va[ch] = a / (float)S2;
    }
    __syncthreads();
    int din = C;
    float *cur = va, *nxt = vb;
    for (int layer = 0; layer < 3; ++layer) {
        const int dout = layer == 0 ? C : (layer == 1 ? C / 2 : C / 4);
        const float* bias = w + (size_t)din * dout;
        for (int o = tid; o < dout; o += NT) {
            float a = bias[o];
            for (int i = 0; i < din; ++i) a = fmaf(cur[i], w[(size_t)i * dout + o], a);
            nxt[o] = fmaxf(a, 0.f);
        }
        __syncthreads();
        w = bias + dout;
        din = dout;
        float* t = cur; cur = nxt; nxt = t;
    }
    if (tid < 32) {
        float a = 0.f;
        for (int i = tid; i < din; i += 32) a = fmaf(cur[i], w[i], a);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xFFFFFFFFu, a, d);
        if (tid == 0) logits[n] = a + w[din];
    }
}

const int kCin[6] = {9, 64, 64, 128, 128, 256};
const int kCout[6] = {64, 64, 128, 128, 256, 256};
const int kHW[6] = {32, 32, 16, 16, 8, 8};

}  // namespace

uint64_t lg_cnn_blob_floats() {
    uint64_t n = 0;
    for (int l = 0; l < 6; ++l) n += 9ull * kCin[l] * kCout[l] + kCout[l];
    n += 256 + 1 + 256 * 256 + 256 + 256 * 128 + 128 + 128 * 64 + 64 + 64 + 1;
    return n;
}

bool lg_cnn_config_ok(const lg_cnn_config* g) {
    if (!g || g->n_blocks < 1 || g->n_blocks > 4 || g->attention < 0 || g->attention > 3) return false;
    for (int b = 0; b < g->n_blocks; ++b)
        if (g->filters[b] < 8 || g->filters[b] > 1024 || g->filters[b] % 8) return false;
    const int C = g->filters[g->n_blocks - 1];
    return C % 16 == 0 && (LG_PATCH >> g->n_blocks) >= 1;
}

bool lg_cnn_config_is_default(const lg_cnn_config* g) {
    return g->n_blocks == 3 && g->filters[0] == 64 && g->filters[1] == 128 && g->filters[2] == 256 && g->attention == 1;
}

uint64_t lg_cnn_config_floats(const lg_cnn_config* g) {
    uint64_t n = 0;
    int cin = LG_CHANNELS;
    for (int b = 0; b < g->n_blocks; ++b) {
        const uint64_t f = g->filters[b];
        n += 9ull * cin * f + f + 9ull * f * f + f;
        cin = (int)f;
    }
    const uint64_t C = cin;
    if (g->attention == 1 || g->attention == 3) n += C + 1;
    if (g->attention == 2 || g->attention == 3) n += C * (C / 16) + C / 16 + (C / 16) * C + C;
    n += C * C + C + C * (C / 2) + C / 2 + (C / 2) * (C / 4) + C / 4 + C / 4 + 1;
    return n;
}

// fp32 forward of a non-default architecture: the same direct-convolution kernel on every layer, generic tail
static int run_cnn_generic(lg_context* c, const float* patches, int n, float* logits, cudaStream_t st) {
    const lg_cnn_config& g = c->cnn.cfg;
    size_t per_patch = 0;                                  // floats of the largest activation of one patch
    for (int b = 0, s = LG_PATCH; b < g.n_blocks; ++b, s >>= 1) per_patch = max(per_patch, (size_t)s * s * g.filters[b]);
    const int chunk = (int)min(min((size_t)n, (size_t)32768), c->cnn_act_bytes / (per_patch * sizeof(float)));   // gridDim.z <= 65535
    if (chunk < 1) { lg_set_error("CNN activation scratch too small for this architecture"); return LG_E_CAPACITY; }
    const int C = g.filters[g.n_blocks - 1], s_final = LG_PATCH >> g.n_blocks, S2 = s_final * s_final;
    const size_t tail_smem = ((size_t)S2 * C + S2 + 3 * (size_t)C) * sizeof(float);
    LG_ENSURE_SMEM(cnn_tail_generic_kernel, tail_smem);
    for (int done = 0; done < n; done += chunk) {
        const int m = min(chunk, n - done);
        float* buf[2] = {(float*)c->cnn_act0, (float*)c->cnn_act1};
        int cur = -1;                                       // -1: the input is the patch tensor (NCHW)
        const float* w = c->cnn.blob;
        int cin = LG_CHANNELS, hw = LG_PATCH;
        for (int l = 0; l < 2 * g.n_blocks; ++l) {
            const int cout = g.filters[l / 2];
            ConvArgs A;
            if (cur < 0) { A.in = patches + (size_t)done * LG_CHANNELS * LG_PATCH * LG_PATCH; A.sn = 9 * 32 * 32; A.sc = 32 * 32; A.sy = 32; A.sx = 1; }
            else { A.in = buf[cur]; A.sn = (long long)hw * hw * cin; A.sy = (long long)hw * cin; A.sx = cin; A.sc = 1; }
            const int dst = cur < 0 ? 0 : cur ^ 1;
            A.H = hw; A.W = hw; A.Cin = cin; A.Cout = cout; A.wt = w; A.bias = w + 9ull * cin * cout; A.out = buf[dst];
            dim3 grid(((hw + 7) / 8) * ((hw + 7) / 8), (cout + CV_CO - 1) / CV_CO, m);
            if (l & 1) conv3x3_relu_kernel<true><<<grid, CV_NT, 0, st>>>(A);
            else conv3x3_relu_kernel<false><<<grid, CV_NT, 0, st>>>(A);
            LG_LAUNCH_CHECK();
            w += 9ull * cin * cout + cout;
            cin = cout;
            if (l & 1) hw >>= 1;
            cur = dst;
        }
        cnn_tail_generic_kernel<<<m, 256, tail_smem, st>>>(buf[cur], w, logits + done, C, S2, g.attention);
        LG_LAUNCH_CHECK();
    }
    return LG_OK;
}

int lg_run_cnn_bf16(lg_context* c, const float* patches, int n, const int32_t* n_dev, float* logits, cudaStream_t st);
int lg_run_cnn_bf16_variant(lg_context* c, const float* patches, int n, float* logits, cudaStream_t st);

// the generic tail (any attention type) on fp32 NHWC features [m][S2][C]; w = the blob behind the convolution weights
int lg_launch_cnn_tail_generic(const float* feat, const float* w, float* logits, int C, int S2, int attention, int m, cudaStream_t st) {
    const size_t tail_smem = ((size_t)S2 * C + S2 + 3 * (size_t)C) * sizeof(float);
    LG_ENSURE_SMEM(cnn_tail_generic_kernel, tail_smem);
    cnn_tail_generic_kernel<<<m, 256, tail_smem, st>>>(feat, w, logits, C, S2, attention);
    LG_LAUNCH_CHECK();
    return LG_OK;
}

// attention + average + MLP on fp32 NHWC [n][4][4][256] features (shared by the fp32 and the bf16 conv paths)
int lg_launch_cnn_tail(const float* feat, const float* blob_tail, float* logits, int n, const int32_t* n_dev, cudaStream_t st) {
    cnn_tail_kernel<<<(n + TL_PB - 1) / TL_PB, TL_NT, 0, st>>>(feat, nullptr, 0, 0, blob_tail, logits, n, n_dev);
    LG_LAUNCH_CHECK();
    return LG_OK;
}

// the same on the bf16 output of the last tensor-core convolution (plane-major, 9x9 halo layout): pools while loading
int lg_launch_cnn_tail_bf16(const void* act, long long act_rows, int lead, const float* blob_tail, float* logits, int n,
                            const int32_t* n_dev, cudaStream_t st) {
    cnn_tail_kernel<<<(n + TL_PB - 1) / TL_PB, TL_NT, 0, st>>>(nullptr, reinterpret_cast<const uint4*>(act), act_rows, lead,
                                                                blob_tail, logits, n, n_dev);
    LG_LAUNCH_CHECK();
    return LG_OK;
}

int lg_run_cnn(lg_context* c, const float* patches, int n, const int32_t* n_dev, float* logits, int use_bf16, cudaStream_t st) {
    if (!c->cnn.loaded) {
        lg_set_error("lg_cnn_forward: no weights loaded (lg_set_cnn_weights)");
        return LG_E_ARG;
    }
    if (!c->cnn.is_default) {
        if (use_bf16) {
            // the attention variants of the default encoder: tensor-core convolutions, generic fp32 tail
            if (c->cnn.bf16_convs && patches) return lg_run_cnn_bf16_variant(c, patches, n, logits, st);
            lg_set_error("the bf16 tensor-core path covers the encoder [64, 128, 256] only; use the fp32 path");
            return LG_E_ARG;
        }
        return run_cnn_generic(c, patches, n, logits, st);
    }
    if (use_bf16) return lg_run_cnn_bf16(c, patches, n, n_dev, logits, st);
    if (!patches) { lg_set_error("the fp32 CNN needs the float32 patch tensor"); return LG_E_ARG; }
    // the fp32 anchor path sizes its grids on the host: with a device-side count it simply runs all n slots
    const float* blob = c->cnn.blob;
    const int cap = min(c->cnn_cap, 32768);               // patches ride on gridDim.z (<= 65535)
    for (int done = 0; done < n; done += cap) {
        const int m = min(cap, n - done);
        const float* in = patches + (size_t)done * LG_CHANNELS * LG_PATCH * LG_PATCH;
        float* a0 = (float*)c->cnn_act0;
        float* a1 = (float*)c->cnn_act1;
        const float* w = blob;
        for (int l = 0; l < 6; ++l) {
            ConvArgs A;
            const int hw = kHW[l];
            if (l == 0) { A.in = in; A.sn = 9 * 32 * 32; A.sc = 32 * 32; A.sy = 32; A.sx = 1; }
            else { A.in = (l & 1) ? a0 : a1; A.sn = (long long)hw * hw * kCin[l]; A.sy = (long long)hw * kCin[l]; A.sx = kCin[l]; A.sc = 1; }
            A.H = hw; A.W = hw; A.Cin = kCin[l]; A.Cout = kCout[l]; A.wt = w; A.bias = w + 9ull * kCin[l] * kCout[l];
            A.out = (l & 1) ? a1 : a0;
            dim3 grid((hw / 8) * (hw / 8), kCout[l] / CV_CO, m);
            if (l & 1) conv3x3_relu_kernel<true><<<grid, CV_NT, 0, st>>>(A);
            else conv3x3_relu_kernel<false><<<grid, CV_NT, 0, st>>>(A);
            LG_LAUNCH_CHECK();
            w += 9ull * kCin[l] * kCout[l] + kCout[l];
        }
        int rc = lg_launch_cnn_tail(a1, w, logits + done, m, nullptr, st);
        if (rc) return rc;
    }
    return LG_OK;
}
