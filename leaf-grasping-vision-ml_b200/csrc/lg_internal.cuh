// Internal declarations shared by the translation units of liblgb200.so (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <stdint.h>
#include <stdio.h>

#include "../../include/leafgrasp.h"

#define LG_NUM_SM_HINT 148

// OpenCV DIST_L2 5x5 chamfer weights in Q16 and its saturation value (oracle: chamfer5_q16)
#define LG_CH_A 65536
#define LG_CH_B 91750
#define LG_CH_C 143976
#define LG_CH_DIST_MAX (0xFFFFFFFFu - 143976u)
#define LG_CH_INF (1 << 30)

#define LG_MIN_LEAF_AREA 10000u
#define LG_NMS_REACH 20  // two +-10 marks overlap  <=>  Chebyshev distance <= 20
#define LG_REGION_PAD 16 // score maps are produced on the leaf bbox grown by half a patch
#define LG_SE_STEM 30
#define LG_SE_PRE 31
#define LG_BAND 8            // rows per band of the stage-1 pass (8 x 4096 pixels still fit a u16 offset)
#define LG_PROF_MARKS 20
#define LG_PROF_RING 32
#define LG_BND_CAP 16384
#define LG_MAX_HOST_CHUNKS 64
#define LG_HOST_CHUNK_FRAMES 16
enum LgMark { LG_M_START = 0, LG_M_STATS, LG_M_SCATTER, LG_M_MEDIAN, LG_M_EDT_COL, LG_M_EDT_ROW, LG_M_SELECT, LG_M_CHAMFER,
              LG_M_ORIENT, LG_M_SCORE, LG_M_NMS, LG_M_GATHER, LG_M_CNN, LG_M_FUSE, LG_M_COUNT,
              // fork / join points of the auxiliary stream (not stages): see lg_stage_times
              LG_M_FORK1 = LG_M_COUNT, LG_M_JOIN1, LG_M_FORK2, LG_M_JOIN2 };

struct LgRegion {
    int x0, y0, x1, y1;  // bounding box of the chosen leaf, exclusive upper bounds
    int ok;              // 0: frame has no chosen leaf -> every stage-2 kernel skips it
    int sx0, sy0, sx1, sy1;  // rectangle the score maps are written on (bbox +- LG_REGION_PAD, or full frame)
};

// Where "is this pixel part of the leaf" comes from: the label image + the chosen id, or a u8 mask.
struct LgMaskSrc {
    const int16_t* labels;
    const uint8_t* mask;
    const int32_t* leaf_id;  // [frames], used with labels
    __device__ __forceinline__ int id(int b) const { return labels ? leaf_id[b] : 1; }
    __device__ __forceinline__ bool at(size_t frame_off, size_t idx, int id_) const {
        return labels ? (labels[frame_off + idx] == (int16_t)id_) : (mask[frame_off + idx] != 0);
    }
};

struct LgOrient {
    double angle;      // rad, after the +90 rule; NaN when there is no contour
    double cos_a, sin_a;
    double major, minor, cx, cy;
    int has_angle;
    int n_hull;
    unsigned status;
    unsigned pad;
    int win_lx, win_ly;  // raster-first pixel of the winning contour's component, bitmask coordinates; -1 = none
};

// Folded CNN weights on the device (see cnn.py:pack_weights for the blob layout)
struct LgCnn {
    float* blob;        // fp32 blob
    uint64_t n_floats;
    void* bf16_blob;    // bf16 copy of conv weights, tensor-core layout (default architecture only)
    int loaded;
    lg_cnn_config cfg;  // architecture the blob belongs to
    int is_default;     // 3 blocks [64,128,256], spatial attention: the one the live node builds
    int bf16_convs;     // encoder [64,128,256] with ANY attention: the six convolutions can run on the tensor cores
};

// Per-stage timing (lg_set_profiling): a ring of event sets, so that the stage times of up to LG_PROF_RING consecutive calls
// can be read back afterwards and a benchmark does not have to synchronise after every step of its timed region.
struct LgProf {
    int on;
    int slot;                                // ring slot of the call in progress / last call
    int calls;                               // lg_process_batch calls since lg_set_profiling(ctx, 1)
    cudaEvent_t ev[LG_PROF_RING][LG_PROF_MARKS];
    int seen[LG_PROF_RING][LG_PROF_MARKS];
};

struct lg_context {
    int B, H, W, L;
    size_t P;
    uint64_t bytes;
    void* allocs;                  // std::vector<void*>* of everything dev_alloc handed out (freed by lg_destroy)
    // ---- stage 1 tables, [B][L] unless noted
    uint32_t* cnt;
    unsigned long long *sx, *sy, *sdep, *sdist;
    uint32_t *bx0, *bx1, *by0, *by1, *border;
    uint32_t* krange;              // [B][2] bounds of the depth keys of the frame's leaf pixels (median search range)
    unsigned long long* ray_tab;   // [P] summed-area table of the viewing-ray length per unit depth, 2^36 fixed point, for ray_cam
    lg_camera ray_cam;
    int ray_valid;
    uint32_t* first_leaf;          // [B] flat index of the first pixel with label >= 1
    float* seg;                    // [B][seg_stride] depth values of the leaf pixels, grouped by label inside every band of LG_BAND rows
    size_t seg_stride;             // P rounded up to 8 values: every band's part starts 32-byte aligned
    uint16_t* band_off;            // [B][n_bands][2][lstride] per band: offsets of every label's block part (in blocks of 64
                                   // values) and pixel part (in values), entry L = total
    int n_bands, lstride;
    uint8_t* ubits;                // [B][H][ub_pitch] (frame stride ub_stride) bit x%8 of byte x/8: pixel belongs to a leaf
    uint8_t* cellocc;              // [B][n_bands][ub_pitch] 1: the 8 x 8 block holds a leaf pixel
    size_t ub_stride;
    int ub_pitch;                  // bytes per row of ubits, a multiple of 4, padding bits zero
    int am_cs;                     // cell size of the distance-transform search (lg_edt_cell_size)
    float* median;                 // [B][L]
    // column pass of the exact distance transform of a caller's mask (lg_edt_squared): per column vertical bit
    // words of the sources and the distance to the nearest source above / below every word
    uint32_t* vbits;               // [B][Hw][W]
    uint16_t *vup, *vdn;           // [B][Hw][W]
    int Hw;                        // ceil(H / 32)
    unsigned long long* edt_best;  // [B] packed (d2 << 32 | ~index)
    int32_t* leaf_id;              // [B]
    lg_leaf_record* records;       // [B][L]
    uint32_t* status;              // [B]
    LgRegion* region;              // [B]
    // ---- stage 2
    int32_t* dt_fwd;               // [2][B][P] forward-pass scratch of the two chamfer transforms
    float* di;                     // [B][P] distance inside (distance_map)
    uint32_t* dt_max;              // [B][2] max Q16 of (inside, outside)
    uint32_t* bnd_list;            // [B][LG_BND_CAP] boundary pixels of the chosen leaf, x | y << 16
    uint32_t* bnd_count;           // [B]
    uint32_t* need_full;           // [B] 1: the outside maximum of this frame needs the full sweeps
    uint32_t* bits;                // [B][bits_stride] leaf bitmask on bbox+1 ring
    size_t bits_stride;
    int run_cap;
    uint16_t *run_x0, *run_x1, *run_y;   // [B][run_cap]
    int32_t* run_parent;                 // [B][run_cap]
    int32_t* row_first;                  // [B][H+3] first run index of each region row
    int32_t* hull;                       // [B][2*(H+2)][2]
    LgOrient* orient;                    // [B]
    double *m_sdf, *m_app, *m_acc, *m_trad;  // [B][P]
    float *m_flat, *m_stem;                  // [B][P]
    uint8_t* m_valid;                        // [B][P]
    double* tip_val_buf;                     // [B][P] scratch of the sample collector (tip values, used as float), allocated by its first call
    uint32_t* tip_idx_buf;                   // [B][P] scratch of the sample collector (tip indices), likewise
    unsigned long long* tile_key;            // [B][tile_cap] candidate search: best alive key (bits of the double) of every 32 x 8 tile
    uint32_t* tile_id;                       // [B][tile_cap] and its flat pixel index
    int tile_cap;                            // ceil(W / 32) * ceil(H / 8)
    float* patches;                          // [B*20][9][32][32]
    float* logits;                           // [B*20], indexed by compact slot
    int32_t* slot_map;                       // [B*20] compact patch index of every (frame, candidate), -1 = no ML score
    int32_t* cnn_count;                      // [1] number of valid slots of the current batch
    int patch_export;                        // 1 (default): the float32 patch tensor is kept for lg_patches; 0: throughput mode
    int patches_valid;                       // the last gather wrote c->patches
    lg_frame_result* results;                // [B]
    float* rec_out;                          // caller-owned [frames][20][4] candidate records (lg_set_record_output), or null
    // CNN scratch
    void* cnn_act0;
    void* cnn_act1;
    size_t cnn_act_bytes;
    int cnn_cap;                             // patches the activation scratch holds
    LgCnn cnn;
    // staging for the *_host entry point: inputs are copied chunk by chunk on copy_stream while the
    // previous chunk is processed on the caller's stream
    int16_t* in_labels;
    float* in_depth;
    lg_frame_result* results_all;            // [B] results of all chunks of one host call
    void* host_pipe;                         // lg_host.cu: host thread pool + run-length staging of the host entry point
    int host_rle;                            // 1 (default): labels cross the link run-length encoded
    uint64_t last_h2d_bytes, last_d2h_bytes; // what the last lg_process_batch_host call moved
    cudaStream_t copy_stream;
    cudaEvent_t copy_ev[LG_MAX_HOST_CHUNKS];
    cudaEvent_t copy_gate;                   // the caller's stream reached the start of this host call
    // Independent branches run side by side: the distance transform of the leaf union next to the per-leaf
    // statistics, the orientation next to the chamfer transforms.  aux_stream forks from and joins the caller's stream.
    cudaStream_t aux_stream, aux2_stream;
    cudaEvent_t ev_fork[3], ev_join[3];
    int overlap;
    // optional per-stage timing (lg_set_profiling): events recorded on the stream the stage runs on
    LgProf* prof;                            // host-side state, kept out of this struct: kernels take the struct by value
    // weights of the traditional score: approach, sdf_score, flatness, accessibility (grasp_point_selector.py:272-277)
    double w_trad[4];
    // constants
    float gauss[25];
    int se30_a[LG_SE_STEM], se30_b[LG_SE_STEM];   // per structuring-element row: first / last+1 column
    int se31_a[LG_SE_PRE], se31_b[LG_SE_PRE];
};

void lg_set_error(const char* fmt, ...);
void lg_host_pipe_destroy(lg_context* c);
#define LG_CUDA(expr)                                                                       \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            lg_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            return LG_E_CUDA;                                                               \
        }                                                                                   \
    } while (0)
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device, per-kernel setting: the sizes already granted are kept
// per (device, kernel) behind a mutex, so that several devices or host threads in one process each get theirs.
int lg_ensure_smem_impl(const void* kernel, size_t bytes);
#define LG_ENSURE_SMEM(kernel, bytes)                                        \
    do {                                                                     \
        int rc__ = lg_ensure_smem_impl((const void*)(kernel), (size_t)(bytes)); \
        if (rc__) return rc__;                                               \
    } while (0)
int lg_prefer_large_smem_impl(const void* kernel);
#define LG_PREFER_LARGE_SMEM(kernel)                                 \
    do {                                                             \
        int rc__ = lg_prefer_large_smem_impl((const void*)(kernel)); \
        if (rc__) return rc__;                                       \
    } while (0)
extern std::atomic<unsigned long long> g_lg_launches;    // host threads may each drive a context
#define LG_LAUNCH_CHECK()            \
    do {                             \
        ++g_lg_launches;             \
        LG_CUDA(cudaGetLastError()); \
    } while (0)
static inline void lg_mark(lg_context* c, int id, cudaStream_t st) {
    LgProf* p = c->prof;
    if (p && p->on) { cudaEventRecord(p->ev[p->slot][id], st); p->seen[p->slot][id] = 1; }
}

// stage launchers (defined across the .cu files); all asynchronous on `st`
int lg_run_stage1(lg_context* c, const int16_t* labels, const float* depth, int n, lg_camera cam, cudaStream_t st);
int lg_edt_cell_size(int H, int W);
int lg_run_select(lg_context* c, int n, lg_camera cam, int32_t* leaf_out, lg_leaf_record* rec_out, cudaStream_t st);
// aux = lg_fork(c, k, st): stream for the side branch (st itself when overlap is off); lg_join makes st wait for it
cudaStream_t lg_fork(lg_context* c, int k, cudaStream_t st);
int lg_join(lg_context* c, int k, cudaStream_t aux, cudaStream_t st);
// variants var_first .. var_first + var_count - 1 of the nvar a frame has (0: the mask as given, 1: its complement)
int lg_run_chamfer(lg_context* c, LgMaskSrc src, int n, int rect_mode, int invert_base, int nvar, int var_first, int var_count,
                   float* out0, uint32_t* q0, uint32_t* out_max, const uint32_t* need_full, cudaStream_t st);
int lg_run_outside_max(lg_context* c, LgMaskSrc src, int n, cudaStream_t st);
int lg_run_orientation(lg_context* c, LgMaskSrc src, int n, cudaStream_t st);
int lg_run_scores(lg_context* c, LgMaskSrc src, const float* depth, int n, lg_camera cam, int full,
                  double* iso_out, cudaStream_t st);
int lg_run_nms(lg_context* c, int n, cudaStream_t st);
// packed != 0: the patches go straight into the tensor-core CNN's input layout (no float32 patch tensor: lg_patches has nothing to export)
int lg_run_gather(lg_context* c, LgMaskSrc src, const float* depth, int n, lg_camera cam, int packed, cudaStream_t st);
// n patches; n_dev (device int, may be null) = the number actually present (<= n): the kernels read it on the device
int lg_run_cnn(lg_context* c, const float* patches, int n, const int32_t* n_dev, float* logits, int use_bf16, cudaStream_t st);
int lg_run_export_patches(lg_context* c, float* out, int n, cudaStream_t st);
bool lg_cnn_config_ok(const lg_cnn_config* g);
bool lg_cnn_config_is_default(const lg_cnn_config* g);
uint64_t lg_cnn_config_floats(const lg_cnn_config* g);
int lg_run_fuse(lg_context* c, LgMaskSrc src, const float* depth, int n, lg_camera cam, int have_ml,
                lg_frame_result* out, float* rec_out, cudaStream_t st);
int lg_run_smooth_depth(const float* in, int n, int h, int w, const float* gauss25, float* out, cudaStream_t st);
int lg_run_collect(lg_context* c, LgMaskSrc src, const float* depth, int n, unsigned long long seed, unsigned long long first_index,
                   const int32_t* grasp_xy, const double* total, float* patches, lg_sample_meta* meta, int32_t* set_sizes,
                   cudaStream_t st);
int lg_run_collector_points(lg_context* c, LgMaskSrc src, int n, int kind, const uint32_t* ranks, int nq, int32_t* xy,
                            cudaStream_t st);
int lg_run_mask_regions(lg_context* c, const uint8_t* mask, int n, int full, cudaStream_t st);
