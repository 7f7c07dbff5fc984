// C-ABI entry points (include/leafgrasp.h): context lifetime, the whole-path pipeline and the
// per-stage calls the drop-in Python classes and the parity tests use.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

#include "lg_internal.cuh"

int lg_run_export_maps(lg_context* c, int n, double* sdf, double* app, float* flat, float* dist, double* acc, float* stem,
                       double* trad, uint8_t* valid, double* angle, cudaStream_t st);
int lg_run_candidates_from_maps(lg_context* c, const double* score, const uint8_t* valid, int n, int32_t* xy, int32_t* count,
                                cudaStream_t st);
uint64_t lg_cnn_blob_floats();
int lg_cnn_prepare_bf16(lg_context* c);
int lg_run_export_orient(lg_context* c, int n, double* out5, cudaStream_t st);
int lg_run_normalize_patches(const float* raw, int n, float* out, cudaStream_t st);

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_lg_launches{0};

void lg_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* lg_last_error(void) { return g_err; }

int lg_ensure_smem_impl(const void* kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return LG_OK;                 // within the default limit: nothing to opt in to
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> granted;
    int dev = 0;
    LG_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    size_t& have = granted[std::make_pair(dev, kernel)];
    if (bytes > have) {
        LG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        have = bytes;
    }
    return LG_OK;
}

// Kernels that run side by side on the two streams should agree on the shared-memory carve-out of an SM: an SM configured
// for a small carve-out has to drain before it can take a CTA that needs a large one, which serialises the "concurrent"
// kernels (measured: orientation beside the chamfer sweeps 0.78 -> 0.54 ms with the hint).  Asked for once per (device, kernel).
int lg_prefer_large_smem_impl(const void* kernel) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, bool> done;
    int dev = 0;
    LG_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    bool& d = done[std::make_pair(dev, kernel)];
    if (!d) {
        LG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        d = true;
    }
    return LG_OK;
}

namespace {

// 5x5 Gaussian, sigma = 5/6, normalised in float64 and stored as float32 exactly as
// ImageProcessor._create_gaussian_kernel does (image_processor.py:25-32); bit patterns from NumPy.
const uint32_t kGaussBits[25] = {
    0x3a3de028u, 0x3bcdce01u, 0x3c5367f0u, 0x3bcdce01u, 0x3a3de028u, 0x3bcdce01u, 0x3d5f11f3u, 0x3de5242eu, 0x3d5f11f3u,
    0x3bcdce01u, 0x3c5367f0u, 0x3de5242eu, 0x3e6b60b5u, 0x3de5242eu, 0x3c5367f0u, 0x3bcdce01u, 0x3d5f11f3u, 0x3de5242eu,
    0x3d5f11f3u, 0x3bcdce01u, 0x3a3de028u, 0x3bcdce01u, 0x3c5367f0u, 0x3bcdce01u, 0x3a3de028u};

// cv2.getStructuringElement(MORPH_ELLIPSE, (n, n)) restated: row i spans columns [a, b)
void ellipse_rows(int n, int* a, int* b) {
    const int r = n / 2, c = n / 2;
    const double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
    for (int i = 0; i < n; ++i) {
        const int dy = i - r;
        int j1 = 0, j2 = 0;
        if (abs(dy) <= r) {
            const int dx = (int)lrint(c * sqrt((r * r - dy * dy) * inv_r2));
            j1 = c - dx > 0 ? c - dx : 0;
            j2 = c + dx + 1 < n ? c + dx + 1 : n;
        }
        a[i] = j1; b[i] = j2;
    }
}

template <typename T>
int dev_alloc(lg_context* c, T** p, size_t count) {
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = sizeof(T);
    LG_CUDA(cudaMalloc((void**)p, bytes));
    c->bytes += bytes;
    static_cast<std::vector<void*>*>(c->allocs)->push_back((void*)*p);
    return LG_OK;
}

}  // namespace

#define TRY(x)                \
    do {                      \
        int rc__ = (x);       \
        if (rc__) return rc__; \
    } while (0)

extern "C" int lg_create(lg_context** out, int max_frames, int height, int width, int max_labels) {
    if (!out || max_frames < 1 || height < 8 || width < 8 || max_labels < 2 || max_labels > 1024 || width > 4096 ||
        height > 16384) {
        lg_set_error("lg_create: bad arguments (frames=%d H=%d W=%d labels=%d)", max_frames, height, width, max_labels);
        return LG_E_ARG;
    }
    lg_context* c = new (std::nothrow) lg_context();
    if (!c) return LG_E_ARG;
    memset(c, 0, sizeof(*c));
    c->B = max_frames; c->H = height; c->W = width; c->L = max_labels;
    c->patch_export = 1;
    c->host_rle = 1;
    c->w_trad[0] = 0.4; c->w_trad[1] = 0.3; c->w_trad[2] = 0.2; c->w_trad[3] = 0.1;
    c->P = (size_t)height * width;
    c->allocs = new (std::nothrow) std::vector<void*>();
    if (!c->allocs) { delete c; return LG_E_ARG; }
    const size_t B = max_frames, L = max_labels, P = c->P;
    int rc = LG_OK;
    auto A = [&](auto pp, size_t count) { if (!rc) rc = dev_alloc(c, pp, count); };
    A(&c->cnt, B * L); A(&c->sx, B * L); A(&c->sy, B * L); A(&c->sdep, B * L); A(&c->sdist, B * L);
    A(&c->bx0, B * L); A(&c->bx1, B * L); A(&c->by0, B * L); A(&c->by1, B * L); A(&c->border, B * L);
    A(&c->krange, B * 2); A(&c->ray_tab, P);
    c->seg_stride = (P + 7) & ~(size_t)7;
    A(&c->first_leaf, B); A(&c->seg, B * c->seg_stride); A(&c->median, B * L);
    c->lstride = (int)((L + 1 + 7) & ~(size_t)7);
    c->n_bands = (height + LG_BAND - 1) / LG_BAND;
    A(&c->band_off, B * (size_t)c->n_bands * 2 * c->lstride);
    c->ub_pitch = (((width + 7) / 8) + 3) & ~3;
    c->ub_stride = ((size_t)height * c->ub_pitch + 15) & ~(size_t)15;
    A(&c->ubits, B * c->ub_stride); A(&c->cellocc, B * (size_t)c->n_bands * c->ub_pitch);
    c->am_cs = lg_edt_cell_size(height, width);
    c->Hw = (height + 31) / 32;
    A(&c->vbits, B * (size_t)c->Hw * width); A(&c->vup, B * (size_t)c->Hw * width); A(&c->vdn, B * (size_t)c->Hw * width);
    A(&c->edt_best, B); A(&c->leaf_id, B); A(&c->records, B * L); A(&c->status, B); A(&c->region, B);
    A(&c->dt_fwd, 2 * B * P); A(&c->di, B * P); A(&c->dt_max, B * 2);
    A(&c->bnd_list, B * (size_t)LG_BND_CAP); A(&c->bnd_count, B); A(&c->need_full, B);
    c->bits_stride = (size_t)((width + 2 + 31) / 32) * (height + 2);
    A(&c->bits, B * c->bits_stride);
    c->run_cap = 8 * (height + 2);
    A(&c->run_x0, B * (size_t)c->run_cap); A(&c->run_x1, B * (size_t)c->run_cap); A(&c->run_y, B * (size_t)c->run_cap);
    A(&c->run_parent, B * (size_t)c->run_cap * 6); A(&c->row_first, B * (size_t)(height + 3));
    A(&c->hull, B * (size_t)(12 * (height + 2))); A(&c->orient, B);
    A(&c->m_sdf, B * P); A(&c->m_app, B * P); A(&c->m_acc, B * P); A(&c->m_trad, B * P);
    A(&c->m_flat, B * P); A(&c->m_stem, B * P); A(&c->m_valid, B * P);
    // (the sample collector's tip lists, 12 bytes per pixel of capacity, are allocated by its first call: lg_collect.cu)
    c->tile_cap = ((width + 31) / 32) * ((height + 7) / 8);
    A(&c->tile_key, B * (size_t)c->tile_cap); A(&c->tile_id, B * (size_t)c->tile_cap);
    A(&c->patches, B * LG_TOP_K * (size_t)(LG_CHANNELS * LG_PATCH * LG_PATCH)); A(&c->logits, B * LG_TOP_K);
    A(&c->slot_map, B * LG_TOP_K); A(&c->cnn_count, 1);
    A(&c->results, B);
    // the two CNN activation buffers are allocated when a model is loaded (ensure_cnn_scratch): a context that never
    // runs the CNN (OptimalLeafSelector's) does not pay for them
    c->cnn_cap = (int)(B * LG_TOP_K < 2048 ? 2048 : B * LG_TOP_K);
    c->cnn_act_bytes = (size_t)c->cnn_cap * 32 * 32 * 64 * sizeof(float);
    // (the staging buffers of the host entry point - 9.3 MB per 1440 x 1080 frame - are allocated by its first call)
    if (rc) { lg_destroy(c); return rc; }
    {
        cudaError_t e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
        for (int i = 0; i < LG_MAX_HOST_CHUNKS && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&c->copy_ev[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->copy_gate, cudaEventDisableTiming);
        // The side stream carries the latency chains (distance-transform search, orientation): its few CTAs should not queue
        // behind the thousands of CTAs of the kernel they run beside, so the stream gets the highest priority.
        int prio_lo = 0, prio_hi = 0;
        if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, prio_hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->aux2_stream, cudaStreamNonBlocking, prio_hi);
        for (int i = 0; i < 3 && e == cudaSuccess; ++i) {
            e = cudaEventCreateWithFlags(&c->ev_fork[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming);
        }
        const char* no = getenv("LG_NO_OVERLAP");
        c->overlap = !(no && no[0] == '1');
        if (e != cudaSuccess) { lg_set_error("lg_create: %s", cudaGetErrorString(e)); lg_destroy(c); return LG_E_CUDA; }
    }
    memcpy(c->gauss, kGaussBits, sizeof(kGaussBits));
    ellipse_rows(LG_SE_STEM, c->se30_a, c->se30_b);
    ellipse_rows(LG_SE_PRE, c->se31_a, c->se31_b);
    cudaError_t e = cudaMemset(c->results, 0, sizeof(lg_frame_result) * B);
    if (e == cudaSuccess) e = cudaMemset(c->dt_max, 0, sizeof(uint32_t) * 2 * B);
    if (e == cudaSuccess) e = cudaMemset(c->orient, 0, sizeof(LgOrient) * B);
    if (e == cudaSuccess) e = cudaMemset(c->ubits, 0, B * c->ub_stride);     // the padding bits of every row stay zero
    if (e != cudaSuccess) { lg_set_error("lg_create: %s", cudaGetErrorString(e)); lg_destroy(c); return LG_E_CUDA; }
    *out = c;
    return LG_OK;
}

extern "C" void lg_destroy(lg_context* c) {
    if (!c) return;
    lg_host_pipe_destroy(c);
    if (c->allocs) {
        std::vector<void*>* v = static_cast<std::vector<void*>*>(c->allocs);
        for (void* p : *v)
            if (p) cudaFree(p);
        delete v;
    }
    if (c->cnn.blob) cudaFree(c->cnn.blob);
    if (c->cnn.bf16_blob) cudaFree(c->cnn.bf16_blob);
    for (int i = 0; i < LG_MAX_HOST_CHUNKS; ++i)
        if (c->copy_ev[i]) cudaEventDestroy(c->copy_ev[i]);
    if (c->copy_gate) cudaEventDestroy(c->copy_gate);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (int i = 0; i < 3; ++i) {
        if (c->ev_fork[i]) cudaEventDestroy(c->ev_fork[i]);
        if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
    }
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->aux2_stream) cudaStreamDestroy(c->aux2_stream);
    if (c->prof) {
        for (int r = 0; r < LG_PROF_RING; ++r)
            for (int i = 0; i < LG_PROF_MARKS; ++i)
                if (c->prof->ev[r][i]) cudaEventDestroy(c->prof->ev[r][i]);
        delete c->prof;
    }
    delete c;
}

extern "C" uint64_t lg_context_bytes(const lg_context* c) { return c ? c->bytes : 0; }

static int ensure_cnn_scratch(lg_context* c) {
    if (!c->cnn_act0) TRY(dev_alloc(c, (unsigned char**)&c->cnn_act0, c->cnn_act_bytes));
    if (!c->cnn_act1) TRY(dev_alloc(c, (unsigned char**)&c->cnn_act1, c->cnn_act_bytes));
    return LG_OK;
}

extern "C" int lg_set_cnn_model(lg_context* c, const lg_cnn_config* cfg, const float* blob_host, uint64_t n_floats) {
    if (!c) return LG_E_ARG;
    if (!blob_host) { c->cnn.loaded = 0; return LG_OK; }
    if (!lg_cnn_config_ok(cfg)) { lg_set_error("lg_set_cnn_model: unsupported architecture"); return LG_E_ARG; }
    const uint64_t want = lg_cnn_config_floats(cfg);
    if (n_floats != want) {
        lg_set_error("lg_set_cnn_model: blob has %llu floats, expected %llu", (unsigned long long)n_floats, (unsigned long long)want);
        return LG_E_ARG;
    }
    if (c->cnn.blob && c->cnn.n_floats != n_floats) { cudaFree(c->cnn.blob); c->cnn.blob = nullptr; }
    if (!c->cnn.blob) LG_CUDA(cudaMalloc((void**)&c->cnn.blob, n_floats * sizeof(float)));
    LG_CUDA(cudaMemcpy(c->cnn.blob, blob_host, n_floats * sizeof(float), cudaMemcpyHostToDevice));
    c->cnn.n_floats = n_floats;
    c->cnn.cfg = *cfg;
    c->cnn.is_default = lg_cnn_config_is_default(cfg) ? 1 : 0;
    c->cnn.bf16_convs = (cfg->n_blocks == 3 && cfg->filters[0] == 64 && cfg->filters[1] == 128 && cfg->filters[2] == 256) ? 1 : 0;
    c->cnn.loaded = 1;
    TRY(ensure_cnn_scratch(c));
    return c->cnn.bf16_convs ? lg_cnn_prepare_bf16(c) : LG_OK;
}

extern "C" int lg_set_cnn_weights(lg_context* c, const float* blob_host, uint64_t n_floats) {
    lg_cnn_config def;
    def.n_blocks = 3; def.filters[0] = 64; def.filters[1] = 128; def.filters[2] = 256; def.filters[3] = 0; def.attention = 1;
    return lg_set_cnn_model(c, &def, blob_host, n_floats);
}

extern "C" uint64_t lg_cnn_model_floats(const lg_cnn_config* cfg) { return lg_cnn_config_ok(cfg) ? lg_cnn_config_floats(cfg) : 0; }

// k = 0, 1: the side stream (stage 1, stage 2); k = 2: the second side stream of stage 2 (no timing marks of its own)
cudaStream_t lg_fork(lg_context* c, int k, cudaStream_t st) {
    if (!c->overlap) return st;
    cudaStream_t side = k == 2 ? c->aux2_stream : c->aux_stream;
    if (cudaEventRecord(c->ev_fork[k], st) != cudaSuccess || cudaStreamWaitEvent(side, c->ev_fork[k], 0) != cudaSuccess)
        return st;
    if (k < 2) lg_mark(c, k == 0 ? LG_M_FORK1 : LG_M_FORK2, side);
    return side;
}

int lg_join(lg_context* c, int k, cudaStream_t aux, cudaStream_t st) {
    if (aux != st) {
        LG_CUDA(cudaEventRecord(c->ev_join[k], aux));
        LG_CUDA(cudaStreamWaitEvent(st, c->ev_join[k], 0));
    }
    if (k < 2) lg_mark(c, k == 0 ? LG_M_JOIN1 : LG_M_JOIN2, st);
    return LG_OK;
}

static int check_batch(lg_context* c, const void* a, const void* b, int frames) {
    if (!c || !a || !b || frames < 1) { lg_set_error("null pointer or empty batch"); return LG_E_ARG; }
    if (frames > c->B) { lg_set_error("batch of %d frames exceeds context capacity %d", frames, c->B); return LG_E_CAPACITY; }
    return LG_OK;
}

extern "C" int lg_select_leaf(lg_context* c, const int16_t* labels, const float* depth, int frames, const lg_camera* cam,
                              int32_t* leaf_id, lg_leaf_record* records, void* stream) {
    TRY(check_batch(c, labels, depth, frames));
    if (!cam) return LG_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    TRY(lg_run_stage1(c, labels, depth, frames, *cam, st));
    TRY(lg_run_select(c, frames, *cam, leaf_id, records, st));
    return LG_OK;
}

static int run_stage2(lg_context* c, LgMaskSrc src, const float* depth, int n, lg_camera cam, int full, double* iso_out,
                      cudaStream_t st) {
    // inside transform on the leaf rectangle (+ its distance map), outside transform on the whole frame (max only)
    // Three independent pieces: the inside transform (a chain of row steps), the orientation, and the maximum of the
    // outside transform (sdf normalisation: branch and bound, the sweeps only as a fallback for the frames it flags).
    // They share no scratch, so each of the two short ones gets a side stream of its own beside the inside transform.
    cudaStream_t aux = lg_fork(c, 1, st);
    cudaStream_t aux2 = lg_fork(c, 2, st);
    int rc = lg_run_orientation(c, src, n, aux);
    if (!rc) rc = lg_run_outside_max(c, src, n, aux2);
    lg_mark(c, LG_M_ORIENT, aux);                   // serialised (overlap off): after both; overlapped: the orientation's end
    if (!rc) rc = lg_run_chamfer(c, src, n, full ? 0 : 1, 0, 2, 0, 1, c->di, nullptr, c->dt_max, c->need_full, st);
    int rcj = lg_join(c, 1, aux, st);               // also on an error path: the side streams must not stay unordered
    { const int rcj2 = lg_join(c, 2, aux2, st); if (!rcj) rcj = rcj2; }
    if (!rc && !rcj) rc = lg_run_chamfer(c, src, n, full ? 0 : 1, 0, 2, 1, 1, c->di, nullptr, c->dt_max, c->need_full, st);
    lg_mark(c, LG_M_CHAMFER, st);
    if (rc) return rc;
    TRY(rcj);
    TRY(lg_run_scores(c, src, depth, n, cam, full, iso_out, st));
    lg_mark(c, LG_M_SCORE, st);
    return LG_OK;
}

static int process_batch_impl(lg_context* c, const int16_t* labels, const float* depth, int frames, const lg_camera* cam,
                              lg_frame_result* results, float* rec_out, int use_bf16_cnn, void* stream) {
    TRY(check_batch(c, labels, depth, frames));
    if (!cam) return LG_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (c->prof && c->prof->on) {
        LgProf* p = c->prof;
        p->slot = p->calls % LG_PROF_RING;
        ++p->calls;
        memset(p->seen[p->slot], 0, sizeof(p->seen[0]));
    }
    lg_mark(c, LG_M_START, st);
    TRY(lg_run_stage1(c, labels, depth, frames, *cam, st));
    TRY(lg_run_select(c, frames, *cam, nullptr, c->records, st));
    LgMaskSrc src{labels, nullptr, c->leaf_id};
    TRY(run_stage2(c, src, depth, frames, *cam, 0, nullptr, st));
    TRY(lg_run_nms(c, frames, st));
    lg_mark(c, LG_M_NMS, st);
    const int have_ml = c->cnn.loaded;
    if (have_ml) {
        // throughput mode (lg_set_patch_export(ctx, 0)): the gather writes the tensor-core CNN's input directly
        const int packed = use_bf16_cnn && !c->patch_export && c->cnn.is_default && frames * LG_TOP_K <= c->cnn_cap;
        TRY(lg_run_gather(c, src, depth, frames, *cam, packed, st));
        lg_mark(c, LG_M_GATHER, st);
        TRY(lg_run_cnn(c, packed ? nullptr : c->patches, frames * LG_TOP_K, c->cnn_count, c->logits, use_bf16_cnn, st));
        lg_mark(c, LG_M_CNN, st);
    }
    TRY(lg_run_fuse(c, src, depth, frames, *cam, have_ml, results, rec_out, st));
    lg_mark(c, LG_M_FUSE, st);
    return LG_OK;
}

extern "C" int lg_process_batch(lg_context* c, const int16_t* labels, const float* depth, int frames, const lg_camera* cam,
                                lg_frame_result* results, int use_bf16_cnn, void* stream) {
    return process_batch_impl(c, labels, depth, frames, cam, results, c ? c->rec_out : nullptr, use_bf16_cnn, stream);
}

extern "C" int lg_set_record_output(lg_context* c, float* records) {
    if (!c) return LG_E_ARG;
    c->rec_out = records;
    return LG_OK;
}

// (lg_process_batch_host: lg_host.cu)
int lg_process_batch_device_impl(lg_context* c, const int16_t* labels, const float* depth, int frames, const lg_camera* cam,
                                 lg_frame_result* results, float* rec_out, int use_bf16_cnn, void* stream) {
    return process_batch_impl(c, labels, depth, frames, cam, results, rec_out, use_bf16_cnn, stream);
}
int lg_context_dev_alloc(lg_context* c, void** p, size_t bytes) {
    unsigned char* q = nullptr;
    TRY(dev_alloc(c, &q, bytes));
    *p = q;
    return LG_OK;
}

extern "C" int lg_set_score_weights(lg_context* c, double approach, double sdf, double flatness, double accessibility) {
    if (!c || !(approach >= 0) || !(sdf >= 0) || !(flatness >= 0) || !(accessibility >= 0)) {
        lg_set_error("lg_set_score_weights: weights must be non-negative numbers");
        return LG_E_ARG;
    }
    c->w_trad[0] = approach; c->w_trad[1] = sdf; c->w_trad[2] = flatness; c->w_trad[3] = accessibility;
    return LG_OK;
}

extern "C" int lg_host_memory_is_pinned(const void* host_ptr) {
    if (!host_ptr) return LG_E_ARG;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host_ptr) != cudaSuccess) { cudaGetLastError(); return 0; }
    return a.type == cudaMemoryTypeHost ? 1 : 0;
}

extern "C" int lg_score_maps(lg_context* c, const uint8_t* mask, const float* depth, int frames, const lg_camera* cam,
                             double* sdf_score, double* approach, float* flatness, double* isolation, float* distance,
                             double* accessibility, float* stem, double* traditional, uint8_t* valid, double* angle_out,
                             void* stream) {
    TRY(check_batch(c, mask, depth, frames));
    if (!cam) return LG_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    TRY(lg_run_mask_regions(c, mask, frames, 1, st));
    LgMaskSrc src{nullptr, mask, nullptr};
    TRY(run_stage2(c, src, depth, frames, *cam, 1, isolation, st));
    TRY(lg_run_export_maps(c, frames, sdf_score, approach, flatness, distance, accessibility, stem, traditional, valid,
                           angle_out, st));
    return LG_OK;
}

extern "C" int lg_candidate_points(lg_context* c, const double* score, const uint8_t* valid, int frames, int32_t* xy,
                                   int32_t* count, void* stream) {
    TRY(check_batch(c, score, valid, frames));
    if (!xy || !count) return LG_E_ARG;
    return lg_run_candidates_from_maps(c, score, valid, frames, xy, count, (cudaStream_t)stream);
}

extern "C" int lg_cnn_forward(lg_context* c, const float* patches, int n, float* logits, int use_bf16, void* stream) {
    if (!c || !patches || !logits || n < 1) return LG_E_ARG;
    return lg_run_cnn(c, patches, n, nullptr, logits, use_bf16, (cudaStream_t)stream);
}

extern "C" int lg_select_grasp_point(lg_context* c, const uint8_t* mask, const float* depth, int frames,
                                     const lg_camera* cam, lg_frame_result* results, int use_bf16_cnn, void* stream) {
    TRY(check_batch(c, mask, depth, frames));
    if (!cam) return LG_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    TRY(lg_run_mask_regions(c, mask, frames, 0, st));
    LgMaskSrc src{nullptr, mask, nullptr};
    TRY(run_stage2(c, src, depth, frames, *cam, 0, nullptr, st));
    TRY(lg_run_nms(c, frames, st));
    const int have_ml = c->cnn.loaded;
    if (have_ml) {
        TRY(lg_run_gather(c, src, depth, frames, *cam, 0, st));
        TRY(lg_run_cnn(c, c->patches, frames * LG_TOP_K, c->cnn_count, c->logits, use_bf16_cnn, st));
    }
    TRY(lg_run_fuse(c, src, depth, frames, *cam, have_ml, results, nullptr, st));
    return LG_OK;
}

extern "C" int lg_leaf_orientation(lg_context* c, const uint8_t* mask, int frames, double* out5, void* stream) {
    if (!c || !mask || !out5 || frames < 1) return LG_E_ARG;
    if (frames > c->B) return LG_E_CAPACITY;
    cudaStream_t st = (cudaStream_t)stream;
    TRY(lg_run_mask_regions(c, mask, frames, 0, st));
    LgMaskSrc src{nullptr, mask, nullptr};
    TRY(lg_run_orientation(c, src, frames, st));
    return lg_run_export_orient(c, frames, out5, st);
}

extern "C" int lg_set_patch_export(lg_context* c, int on) {
    if (!c) return LG_E_ARG;
    c->patch_export = on ? 1 : 0;
    return LG_OK;
}

extern "C" int lg_patches(lg_context* c, float* patches_out, int frames, void* stream) {
    if (!c || !patches_out || frames < 1 || frames > c->B) return LG_E_ARG;
    if (!c->patches_valid) {
        lg_set_error("lg_patches: the last call ran in throughput mode (lg_set_patch_export(ctx, 0)): no float32 patch tensor was written");
        return LG_E_ARG;
    }
    return lg_run_export_patches(c, patches_out, frames, (cudaStream_t)stream);
}

extern "C" int lg_smooth_depth(const float* depth, int n, int height, int width, float* out, void* stream) {
    if (!depth || !out || n < 1 || height < 3 || width < 3) {
        lg_set_error("lg_smooth_depth: bad arguments (reflect padding by 2 needs an image of at least 3 x 3)");
        return LG_E_ARG;
    }
    float g[25];
    memcpy(g, kGaussBits, sizeof(g));
    return lg_run_smooth_depth(depth, n, height, width, g, out, (cudaStream_t)stream);
}

extern "C" int lg_normalize_patches(lg_context* c, const float* raw, int n, float* out, void* stream) {
    if (!c || !raw || !out || n < 1) return LG_E_ARG;
    return lg_run_normalize_patches(raw, n, out, (cudaStream_t)stream);
}

extern "C" int lg_collect_samples(lg_context* c, const int16_t* labels, const uint8_t* mask, const float* depth, int frames,
                                  uint64_t seed, uint64_t first_frame_index, const int32_t* grasp_xy, const double* total_score,
                                  float* patches, lg_sample_meta* meta, int32_t* set_sizes, void* stream) {
    if (!c || !depth || !patches || !meta || !set_sizes || frames < 1 || (labels == nullptr) == (mask == nullptr)) {
        lg_set_error("lg_collect_samples: bad arguments (exactly one of labels / mask)");
        return LG_E_ARG;
    }
    if (frames > c->B) return LG_E_CAPACITY;
    LgMaskSrc src{labels, mask, labels ? c->leaf_id : nullptr};
    return lg_run_collect(c, src, depth, frames, seed, first_frame_index, grasp_xy, total_score, patches, meta, set_sizes,
                          (cudaStream_t)stream);
}

extern "C" int lg_collector_points(lg_context* c, const int16_t* labels, const uint8_t* mask, int frames, int kind,
                                   const uint32_t* ranks, int nq, int32_t* xy, void* stream) {
    if (!c || !ranks || !xy || frames < 1 || nq < 1 || kind < 0 || kind > 2 || (labels == nullptr) == (mask == nullptr)) {
        lg_set_error("lg_collector_points: bad arguments");
        return LG_E_ARG;
    }
    if (frames > c->B) return LG_E_CAPACITY;
    LgMaskSrc src{labels, mask, labels ? c->leaf_id : nullptr};
    return lg_run_collector_points(c, src, frames, kind, ranks, nq, xy, (cudaStream_t)stream);
}

extern "C" uint64_t lg_sizeof_sample_meta(void) { return sizeof(lg_sample_meta); }
extern "C" uint64_t lg_sizeof_frame_result(void) { return sizeof(lg_frame_result); }
extern "C" uint64_t lg_sizeof_leaf_record(void) { return sizeof(lg_leaf_record); }
extern "C" uint64_t lg_cnn_weight_floats(void) { return lg_cnn_blob_floats(); }

extern "C" int lg_set_profiling(lg_context* c, int on) {
    if (!c) return LG_E_ARG;
    if (!c->prof) {
        if (!on) return LG_OK;
        c->prof = new (std::nothrow) LgProf();
        if (!c->prof) return LG_E_ARG;
        memset(c->prof, 0, sizeof(LgProf));
        for (int r = 0; r < LG_PROF_RING; ++r)
            for (int i = 0; i < LG_PROF_MARKS; ++i) LG_CUDA(cudaEventCreate(&c->prof->ev[r][i]));
    }
    c->prof->on = on ? 1 : 0;
    c->prof->slot = 0;
    c->prof->calls = 0;
    memset(c->prof->seen, 0, sizeof(c->prof->seen));
    return LG_OK;
}

/* ms[i] = device time of stage i of one recorded lg_process_batch call = time between the mark recorded after it and
 * the mark that precedes it on the stream it ran on (stages on the auxiliary stream start at their fork mark, stages
 * that follow a join start at the join mark); 0 for stages that did not run.  Stages on different streams overlap, so
 * the sum can exceed the step time.  Synchronises on the recorded events. */
static int stage_times_of_slot(lg_context* c, int slot, float* ms) {
    const int* seen = c->prof->seen[slot];
    cudaEvent_t* ev = c->prof->ev[slot];
    for (int i = 0; i < LG_M_COUNT; ++i) ms[i] = 0.f;
    if (!seen[LG_M_START]) return LG_OK;
    int pred[LG_PROF_MARKS];
    for (int i = 0; i < LG_PROF_MARKS; ++i) pred[i] = i - 1;
    pred[LG_M_START] = -1;
    pred[LG_M_FORK1] = LG_M_STATS;   pred[LG_M_EDT_COL] = LG_M_FORK1;
    pred[LG_M_JOIN1] = LG_M_EDT_ROW; pred[LG_M_SELECT] = LG_M_JOIN1;
    pred[LG_M_FORK2] = LG_M_SELECT;  pred[LG_M_ORIENT] = LG_M_FORK2; pred[LG_M_CHAMFER] = LG_M_SELECT;
    pred[LG_M_JOIN2] = LG_M_ORIENT;  pred[LG_M_SCORE] = LG_M_CHAMFER;   // the chamfer mark is the last one before the scores
    if (!seen[LG_M_FORK1]) {   // overlap off: stats (+ column pass), row pass, scatter, median on the caller's stream
        pred[LG_M_EDT_COL] = LG_M_STATS; pred[LG_M_SCATTER] = LG_M_EDT_ROW; pred[LG_M_JOIN1] = LG_M_MEDIAN;
    }
    if (!seen[LG_M_FORK2]) { pred[LG_M_ORIENT] = LG_M_SELECT; pred[LG_M_CHAMFER] = LG_M_ORIENT; pred[LG_M_JOIN2] = LG_M_ORIENT; }
    for (int i = 1; i < LG_M_COUNT; ++i) {
        if (!seen[i]) continue;
        int p = pred[i];
        while (p >= 0 && !seen[p]) p = pred[p];
        if (p < 0) continue;
        LG_CUDA(cudaEventSynchronize(ev[i]));
        LG_CUDA(cudaEventSynchronize(ev[p]));
        LG_CUDA(cudaEventElapsedTime(&ms[i], ev[p], ev[i]));
    }
    return LG_OK;
}

extern "C" int lg_stage_times(lg_context* c, float* ms, int n) {
    if (!c || !ms || n < LG_M_COUNT) return LG_E_ARG;
    for (int i = 0; i < n; ++i) ms[i] = 0.f;
    if (!c->prof || !c->prof->on || c->prof->calls == 0) return LG_OK;
    return stage_times_of_slot(c, c->prof->slot, ms);
}

extern "C" int lg_stage_times_mean(lg_context* c, float* ms, int n, int* calls_out) {
    if (!c || !ms || n < LG_M_COUNT) return LG_E_ARG;
    for (int i = 0; i < n; ++i) ms[i] = 0.f;
    const bool on = c->prof && c->prof->on;
    const int calls = !on ? 0 : (c->prof->calls < LG_PROF_RING ? c->prof->calls : LG_PROF_RING);
    if (calls_out) *calls_out = calls;
    if (calls == 0) return LG_OK;
    for (int r = 0; r < calls; ++r) {
        float one[LG_PROF_MARKS];
        TRY(stage_times_of_slot(c, r, one));
        for (int i = 0; i < LG_M_COUNT; ++i) ms[i] += one[i];
    }
    for (int i = 0; i < LG_M_COUNT; ++i) ms[i] /= (float)calls;
    return LG_OK;
}

extern "C" uint64_t lg_launch_count(void) { return g_lg_launches.load(); }

extern "C" int lg_set_overlap(lg_context* c, int on) {
    if (!c) return LG_E_ARG;
    c->overlap = on ? 1 : 0;
    return LG_OK;
}
