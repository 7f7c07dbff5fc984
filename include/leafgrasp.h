/*
 * leafgrasp.h - C-ABI of the B200-native grasp-selection hot path.
 *
 * The reference (Srecharan/Leaf-Grasping-Vision-ML) is pure Python and has no FFI of its own; the
 * boundary it offers is three Python classes called from scripts/leaf_grasp_node_v3.py:114-119.
 * Each entry point below names the reference routine it replaces (paths relative to the reference
 * checkout).  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; plain C types only;
 *   - images are row-major [frames][H][W]; pixel (x, y) = (column, row); flat index y*W + x;
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream); no entry point synchronises
 *     the host unless its comment says so;
 *   - return value: 0 = ok, negative = LG_E_* below.  Data-dependent problems found on the device
 *     (label out of range, region too large ...) are reported per frame in lg_frame_result.status.
 *   - there is no CPU implementation behind any of these: without a CUDA device they fail.
 */
#ifndef LEAFGRASP_H_
#define LEAFGRASP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LG_OK 0
#define LG_E_ARG (-1)      /* bad argument (null pointer, size out of range) */
#define LG_E_CUDA (-2)     /* a CUDA runtime call failed; see lg_last_error() */
#define LG_E_CAPACITY (-3) /* batch / image larger than the context was created for */

#define LG_TOP_K 20        /* grasp_point_selector.py:197 */
#define LG_PATCH 32        /* grasp_point_selector.py:66  */
#define LG_CHANNELS 9      /* grasp_point_selector.py:127 */

/* per-frame status bits (lg_frame_result.status) */
#define LG_ST_NO_LEAF 1u        /* select_optimal_leaf would return None (leaf_scorer.py:49-50,144-146) */
#define LG_ST_LABEL_RANGE 2u    /* a label id outside [0, max_labels) was seen; frame result undefined */
#define LG_ST_RUNS_OVERFLOW 4u  /* chosen leaf has more row runs than the orientation scratch holds */
#define LG_ST_NO_CANDIDATE 8u   /* _get_candidate_points returned [] */

typedef struct lg_context lg_context; /* opaque: device scratch sized for (max_frames, H, W) */

/* Camera constants, as set by GraspPointSelector.set_camera_params / OptimalLeafSelector.set_camera_params
 * (grasp_point_selector.py:145-150, leaf_scorer.py:19-23): f = P[0,0], cx = P[0,2], cy = P[1,2]. */
typedef struct lg_camera {
    double f, cx, cy;
} lg_camera;

/* One leaf's stage-1 record (leaf_scorer.py:74-138). */
typedef struct lg_leaf_record {
    int32_t leaf_id;
    uint32_t area;         /* pixel count */
    float median_depth;    /* np.median(depth[mask]) */
    float mean_depth;      /* np.mean(depth[mask]) */
    double centroid_x, centroid_y;
    double clutter, distance, visibility; /* the three scores */
    double mean_distance;  /* raw_scores['distance'] */
    int32_t is_tall;       /* median < mean of medians */
    int32_t is_candidate;  /* area >= 10000 */
} lg_leaf_record;

/* Everything select_optimal_leaf + select_grasp_point return for one frame
 * (leaf_grasp_node_v3.py:114-119, grasp_point_selector.py:184-253). */
typedef struct lg_frame_result {
    uint32_t status;             /* LG_ST_* bits */
    int32_t leaf_id;             /* -1 when no leaf */
    int32_t n_candidates;        /* <= LG_TOP_K */
    int32_t n_positive;          /* candidates whose key (traditional score x valid) is > 0 */
    int32_t cand_x[LG_TOP_K], cand_y[LG_TOP_K];
    double trad[LG_TOP_K];       /* traditional_score at each candidate */
    float logit[LG_TOP_K];       /* CNN output; NaN where no ML score was produced */
    double ml[LG_TOP_K];         /* tanh(3*sigmoid(logit))/2 + 1/2 */
    int32_t ml_valid[LG_TOP_K];  /* 1 where the reference's get_ml_score returns a value */
    int32_t best_index;          /* index into cand_* of the fused pick */
    int32_t ml_used;             /* a fused score beat candidate 0's traditional score */
    double best_score;
    int32_t grasp_x, grasp_y;    /* grasp_point_2d */
    double grasp_3d[3];          /* get_3d_grasp_point */
    double pre_grasp[3];         /* calculate_pre_grasp_point */
    double angle;                /* estimate_leaf_orientation angle (rad), NaN if no contour */
    float sdf_max;               /* max |dist_inside - dist_outside| */
    int32_t region[4];           /* x0, y0, x1, y1 (exclusive) of the chosen leaf's bounding box */
} lg_frame_result;

/* ---- lifetime ------------------------------------------------------------------------------ */

/* Allocate device scratch for batches of up to max_frames images of height x width whose label ids
 * lie in [0, max_labels).  Uses the current CUDA device. */
int lg_create(lg_context** out, int max_frames, int height, int width, int max_labels);
void lg_destroy(lg_context* ctx);
const char* lg_last_error(void);
/* Bytes of device memory the context holds. */
uint64_t lg_context_bytes(const lg_context* ctx);

/* Folded GraspPointCNN weights (model.py:5-100 with eval-mode BatchNorm folded into the preceding
 * conv / linear).  `blob_host` is a HOST float array laid out as described in cnn.py:pack_weights;
 * it is copied to the device (synchronous).  Passing NULL clears the model (CV-only path,
 * grasp_point_selector.py:52-57). */
int lg_set_cnn_weights(lg_context* ctx, const float* blob_host, uint64_t n_floats);

/* Any architecture the reference's constructor accepts (model.py:6-86; the hyper-parameter sweep of
 * train_model_mlflow.py:173-182 produces them): n_blocks encoder blocks with filters[] channels, attention
 * 0 = none, 1 = spatial, 2 = channel, 3 = hybrid.  Blob layout: cnn.py:pack_weights.  The default architecture
 * (3 blocks 64/128/256, spatial) runs on the tensor cores (use_bf16) or the fp32 path; the others on the fp32 path. */
typedef struct lg_cnn_config {
    int32_t n_blocks;
    int32_t filters[4];
    int32_t attention;
} lg_cnn_config;
int lg_set_cnn_model(lg_context* ctx, const lg_cnn_config* cfg, const float* blob_host, uint64_t n_floats);
uint64_t lg_cnn_model_floats(const lg_cnn_config* cfg);   /* 0 for an unsupported architecture */

/* ---- whole path ----------------------------------------------------------------------------- */

/* The per-frame path of leaf_grasp_node_v3.py:110-119 for `frames` independent frames:
 * OptimalLeafSelector.select_optimal_leaf (leaf_scorer.py:25) then GraspPointSelector.select_grasp_point
 * (grasp_point_selector.py:184) on the chosen leaf, including the batched GraspPointCNN forward and the
 * CV/ML fusion.  labels int16 [frames][H][W], depth float32 [frames][H][W]; results (device)
 * lg_frame_result[frames].  use_bf16_cnn != 0 runs the tensor-core CNN, 0 the fp32 one. */
int lg_process_batch(lg_context* ctx, const int16_t* labels, const float* depth, int frames,
                     const lg_camera* cam_host, lg_frame_result* results, int use_bf16_cnn, void* stream);

/* Same, but labels/depth/results are HOST buffers: copies in (16-frame chunks on a copy stream, each chunk processed
 * as soon as it has landed), runs, copies the results back and synchronises the stream.  This is the call the
 * end-to-end benchmark makes.  The inputs should be PINNED (cudaHostAlloc / cudaHostRegister / torch pin_memory):
 * pageable memory is accepted, but CUDA then stages every copy through its own pinned buffer and the copies no longer
 * overlap the kernels; lg_host_memory_is_pinned tells which kind a pointer is. */
int lg_process_batch_host(lg_context* ctx, const int16_t* labels_host, const float* depth_host, int frames,
                          const lg_camera* cam_host, lg_frame_result* results_host, int use_bf16_cnn,
                          void* stream);

/* Weights of the traditional score, traditional = (approach * a + sdf_score * s + flatness * f + accessibility * c) *
 * (1 - stem_penalty) (grasp_point_selector.py:272-277).  The default is what the reference's CODE uses, 0.4 / 0.3 / 0.2 /
 * 0.1, and every parity test runs with it; the reference's README.md:83-87 advertises another set (approach 0.40, edge =
 * sdf 0.20, flatness 0.25, accessibility 0.15), which SURVEY.md 8(a') asks to be selectable: pass it here. */
int lg_set_score_weights(lg_context* ctx, double approach, double sdf, double flatness, double accessibility);

/* The label image is piecewise constant, so lg_process_batch_host run-length encodes it with host threads inside the call
 * (lossless; a frame with more than H*W/16 runs is copied as it is) and a kernel expands it on the device: ~0.1 MB instead
 * of 3.1 MB per 1440 x 1080 frame cross the host-to-device link, which bounds the call.  on = 0 copies the labels raw.
 * lg_host_call_bytes: bytes the last lg_process_batch_host call copied to / from the device. */
int lg_set_host_label_rle(lg_context* ctx, int on);
int lg_host_call_bytes(const lg_context* ctx, uint64_t* h2d_bytes, uint64_t* d2h_bytes);

/* The encoder those host threads run (pure host code, needs no device): the runs of one label image, one 32-bit word
 * `first column | label << 16` per run - a run starts at column 0 of every row and wherever a label differs from its left
 * neighbour - with rowoff[y] = index of row y's first run and rowoff[height] = the number of runs (rowoff holds height + 1
 * words).  Returns the number of runs, or 0xFFFFFFFF when they do not fit run_cap (or an argument is bad; width <= 65535).
 * isa: -1 = the fastest version the CPU supports (lg_rle_host_isa: 2 AVX-512BW, 32 labels per step; 1 AVX2, 16; 0 portable),
 * 0..2 = that version if the CPU has it; every version produces the same words.  Replaces nothing in the reference: it is
 * the price of not sending leaf_grasp_node_v3.py:110's mask tensor over the link as it is. */
uint32_t lg_rle_encode_labels(const int16_t* labels, int height, int width, uint32_t* runs, uint32_t run_cap,
                              uint32_t* rowoff, int isa);
int lg_rle_host_isa(void);

/* 1: page-locked host memory known to CUDA, 0: pageable (or not host memory), < 0: error. */
int lg_host_memory_is_pinned(const void* host_ptr);

/* Candidate records for the multi-GPU aggregation (the only exchange between ranks is one all-gather of
 * fixed-size records): when `records` (DEVICE float32 [frames][LG_TOP_K][4], caller-owned, NULL = off) is set, every
 * following lg_process_batch / lg_process_batch_host call also writes, per frame and candidate slot, (x, y, traditional
 * score, ML score or 0) - (-1, -1, 0, 0) in unused slots - straight from the fusion kernel; frame i of a call goes to
 * records + i * 80 floats.  The buffer is the send buffer of the NCCL all-gather; nothing is reshuffled on the host. */
int lg_set_record_output(lg_context* ctx, float* records);

/* ---- stage 1: OptimalLeafSelector.select_optimal_leaf (leaf_scorer.py:25-203) ---------------- */

/* leaf_id int32 [frames] (-1 = None); records lg_leaf_record [frames][max_labels] (entry i describes
 * label id i; area 0 = absent), either may be NULL.  Also fills the context's per-frame regions used
 * by stage 2. */
int lg_select_leaf(lg_context* ctx, const int16_t* labels, const float* depth, int frames,
                   const lg_camera* cam_host, int32_t* leaf_id, lg_leaf_record* records, void* stream);

/* ---- distance transforms -------------------------------------------------------------------- */

/* cv2.distanceTransform(mask, DIST_L2, 5) with IPP off (grasp_point_selector.py:266,529-530), batched:
 * mask uint8 [n][H][W] (non-zero = inside, zero pixels are the sources).  dist float32 [n][H][W] and/or
 * q16 uint32 [n][H][W] (OpenCV's integer field before the float conversion); either may be NULL.
 * max_q16 uint32 [n] may be NULL. */
int lg_chamfer_transform(lg_context* ctx, const uint8_t* mask, int n, int invert, float* dist,
                         uint32_t* q16, uint32_t* max_q16, void* stream);

/* Exact squared Euclidean distance transform (two-pass, Meijster lower envelope): for every non-zero
 * pixel of mask the squared distance to the nearest zero pixel; uint32 [n][H][W].  The leaf-selection
 * stage uses the same kernels on (labels < 1) to find the background pixel farthest from any leaf
 * (leaf_scorer.py:67-71, scikit-fmm substitute).  argmax int32 [n] (flat index of the first maximum)
 * may be NULL; d2 may be NULL. */
int lg_edt_squared(lg_context* ctx, const uint8_t* mask, int n, uint32_t* d2, int32_t* argmax, void* stream);

/* ---- stage 2: GraspPointSelector._calculate_all_scores (grasp_point_selector.py:256-288) ----- */

/* Full-frame score maps of ONE leaf mask per frame, as the reference returns them:
 * mask uint8 [frames][H][W]; outputs (each may be NULL) sdf_score/approach/accessibility/isolation/
 * traditional float64, flatness/distance/stem float32, valid uint8, all [frames][H][W].
 * angle_out double [frames] (estimate_leaf_orientation, :718-752) may be NULL. */
int lg_score_maps(lg_context* ctx, const uint8_t* mask, const float* depth, int frames,
                  const lg_camera* cam_host, double* sdf_score, double* approach, float* flatness,
                  double* isolation, float* distance, double* accessibility, float* stem, double* traditional,
                  uint8_t* valid, double* angle_out, void* stream);

/* GraspPointSelector._get_candidate_points (grasp_point_selector.py:447-482): score float64 and valid
 * uint8 [frames][H][W] -> xy int32 [frames][LG_TOP_K][2], count int32 [frames]. */
int lg_candidate_points(lg_context* ctx, const double* score, const uint8_t* valid, int frames,
                        int32_t* xy, int32_t* count, void* stream);

/* ---- stage 3: GraspPointCNN.forward (ml_grasp_optimizer/model.py:102-128) ------------------- */

/* patches float32 [n][9][32][32] -> logits float32 [n].  Needs lg_set_cnn_weights first. */
int lg_cnn_forward(lg_context* ctx, const float* patches, int n, float* logits, int use_bf16, void* stream);

/* The bf16 tensor-core path stopped after conv layer `layer` (0..5; the 2x2 max-pool of model.py:25-28 applied
 * after layers 1, 3 and 5): features float32 [n][C][S][S] with (C, S) = (64,32) (64,16) (128,16) (128,8) (256,8)
 * for layers 0..4 and float32 [n][4][4][256] (channels last) for layer 5.  n must fit one activation chunk
 * (>= 2048 patches).  Used by the per-layer parity tests. */
int lg_cnn_bf16_features(lg_context* ctx, const float* patches, int n, int layer, float* features, void* stream);

/* ---- training samples ---------------------------------------------------- */

/* EnhancedGraspDataCollector.collect_sample (ml_grasp_optimizer/data_collector.py:175-348) for every frame of the
 * batch the context has just processed: call it right after lg_process_batch (pass the same `labels`, mask NULL) or
 * lg_select_grasp_point (pass the same `mask`, labels NULL) with the same depth, on the same stream.
 * Per frame LG_SAMPLES_PER_FRAME slots: 0 = the positive sample at the grasp point, 1..3 = its rot90 copies with
 * depth noise and score jitter (:250-299), 4..6 = up to three negatives from the tip / stem / edge candidate sets
 * (:301-348, 420-487).  patches float32 [frames][7][9][32][32]: RAW (un-normalised) windows, channels depth, mask,
 * sdf_score, approach_score, flatness_map, isolation_map, distance_map, accessibility_map, stem_penalty (:139-143);
 * meta [frames][7]; set_sizes int32 [frames][3] = sizes of the tip (largest quarter), stem and edge sets.
 * A slot with valid == 0 holds no sample (the reference added none: window leaves the image, non-finite depth, ...).
 * grasp_xy int32 [frames][2] (device) overrides the grasp points, NULL = the ones just selected; total_score double
 * [frames] (device) is the `total_score` argument, NULL = max of traditional_score over the frame (the reference's
 * call site, grasp_point_selector_bkp.py:146-152).
 * Randomness: the reference uses the global generators of `random` and `torch`; here every draw is a pure function
 * of (seed, first_frame_index + frame, purpose, counter) - see oracle CollectorRng for the definition.
 * The call reuses scratch of the candidate search (NMS lists) and of the chamfer transform; lg_frame_result is kept. */
#define LG_SAMPLES_PER_FRAME 7
typedef struct lg_sample_meta {
    int32_t valid;          /* 1: the slot holds a sample */
    int32_t label;          /* 1 positive, 0 negative */
    int32_t is_augmented;
    int32_t kind;           /* 0 positive, 1..3 rot90 x k, 4 tip, 5 stem, 6 edge */
    int32_t x, y;           /* 'grasp_point' as the reference stores it (rotated copies: _rotate_point, :402-418) */
    double total_score;
} lg_sample_meta;
int lg_collect_samples(lg_context* ctx, const int16_t* labels, const uint8_t* mask, const float* depth, int frames,
                       uint64_t seed, uint64_t first_frame_index, const int32_t* grasp_xy, const double* total_score,
                       float* patches, lg_sample_meta* meta, int32_t* set_sizes, void* stream);
/* Points of a candidate set in the reference's list order (_get_tip_points / _get_stem_points / _get_edge_points,
 * :420-487), after lg_collect_samples on the same batch: kind 0 tip, 1 stem, 2 edge; ranks uint32 [frames][nq]
 * (device) are taken modulo the set size; xy int32 [frames][nq][2], (-1, -1) for an empty set. */
int lg_collector_points(lg_context* ctx, const int16_t* labels, const uint8_t* mask, int frames, int kind,
                        const uint32_t* ranks, int nq, int32_t* xy, void* stream);
uint64_t lg_sizeof_sample_meta(void);

/* ---- stage 2 on a caller-supplied leaf mask --------------------------------------------------- */

/* GraspPointSelector.select_grasp_point (grasp_point_selector.py:184-253) for one binary leaf mask per
 * frame: mask uint8 [frames][H][W] (non-zero = leaf), depth float32; results (device) as lg_process_batch
 * (leaf_id is 1 for a non-empty mask). */
int lg_select_grasp_point(lg_context* ctx, const uint8_t* mask, const float* depth, int frames,
                          const lg_camera* cam_host, lg_frame_result* results, int use_bf16_cnn, void* stream);

/* GraspPointSelector.estimate_leaf_orientation (grasp_point_selector.py:718-752): out5 double [frames][5] =
 * angle (rad; NaN when there is no contour), major axis, minor axis, centre x, centre y. */
int lg_leaf_orientation(lg_context* ctx, const uint8_t* mask, int frames, double* out5, void* stream);

/* The [frames][LG_TOP_K][9][32][32] float32 patch tensor the last lg_process_batch / lg_select_grasp_point
 * call fed to the CNN (rows of candidates without an ML score are zero). */
int lg_patches(lg_context* ctx, float* patches_out, int frames, void* stream);

/* Drop-in mode (on = 1, the default) keeps that float32 patch tensor - what get_ml_score (grasp_point_selector.py:59-127)
 * builds per candidate - for lg_patches.  Throughput mode (on = 0): with the bf16 tensor-core CNN the gather kernel writes
 * the CNN's input layout directly and no float32 patch tensor exists (lg_patches then fails); results are identical. */
int lg_set_patch_export(lg_context* ctx, int on);

/* ImageProcessor.smooth_depth (image_processor.py:56-64): reflect padding by 2 + the 5x5 Gaussian (sigma = 5/6, float32
 * taps) on n float32 images of height x width (>= 3 x 3), any size, no context needed.  out float32 [n][height][width]. */
int lg_smooth_depth(const float* depth, int n, int height, int width, float* out, void* stream);

/* Per-patch min-max normalisation of get_ml_score (grasp_point_selector.py:83-123) on caller-built raw
 * patches float32 [n][9][32][32] (channel 1, the mask, is passed through). */
int lg_normalize_patches(lg_context* ctx, const float* raw, int n, float* out, void* stream);

/* ---- measurement hooks (bench.py) ------------------------------------------------------------------ */

/* Record a CUDA event on the launch stream after every stage of lg_process_batch. */
int lg_set_profiling(lg_context* ctx, int on);
/* ms[14]: device milliseconds of each stage of the last lg_process_batch call, in the order
 * [unused, leaf_stats, scatter, median, edt_columns, edt_rows, select, chamfer, orientation, score_maps,
 *  candidates, patches, cnn, fuse] (edt_columns is ~0 in lg_process_batch: the column pass of the union distance
 *  transform is fused into leaf_stats).  A stage's time runs from the previous mark on the stream it ran on; stages on the
 * internal stream overlap the others, so the sum can exceed the step time.  Synchronises on the recorded events. */
int lg_stage_times(lg_context* ctx, float* ms, int n);
/* The same, averaged over the lg_process_batch calls made since lg_set_profiling(ctx, 1) (the last 32 at most; the
 * host-buffer entry point makes one call per chunk): a benchmark can time every stage of its whole timed region
 * without synchronising between steps.  calls_out (may be NULL) = number of calls averaged. */
int lg_stage_times_mean(lg_context* ctx, float* ms, int n, int* calls_out);
/* Independent stages (union distance transform | per-leaf statistics, orientation | chamfer transforms) run side by
 * side on an internal stream that forks from and joins `stream`; on = 0 serialises them on `stream` (clean per-stage
 * times).  Default: on, unless the environment has LG_NO_OVERLAP=1.  Results are identical either way. */
int lg_set_overlap(lg_context* ctx, int on);
/* Number of kernels this library has launched in this process so far. */
uint64_t lg_launch_count(void);

/* ---- ABI self-description (the Python binding checks its struct layouts against these) ---------- */
uint64_t lg_sizeof_frame_result(void);
uint64_t lg_sizeof_leaf_record(void);
uint64_t lg_cnn_weight_floats(void);

#ifdef __cplusplus
}
#endif
#endif /* LEAFGRASP_H_ */
