// The host-buffer entry point (include/leafgrasp.h: lg_process_batch_host): the step in front of the path, reference
// scripts/leaf_grasp_node_v3.py:110-111 (`torch.from_numpy(...).to(device)` of the mask and depth messages).
//
// End to end the path is bound by the host-to-device link: 9.33 MB per 1440 x 1080 frame (int16 labels 3.11 MB + float32
// depth 6.22 MB) against ~4 ms of kernels per 256 frames.  Two things keep the link busy and its load small:
//
//   * chunks of 16 frames are copied on a private stream and processed on the caller's stream as soon as their event
//     fires, so the kernels hide under the copies;
//   * the LABEL image does not cross the link as it is.  An instance-label image is piecewise constant (a few dozen runs
//     per row), so host threads run-length encode it inside the call - one 32-bit word (first column | label << 16) per
//     run, per-row offsets - while the previous chunk's depth is in flight, ~90 KB cross the link instead of 3.11 MB, and
//     a small kernel expands the runs into the int16 image the path reads.  Lossless: the device sees the same labels
//     bit for bit (`host_batch_identical_to_device_batch` in bench.py, tests/test_gpu_round2.py).  A frame whose runs do
//     not fit the staging (more than P / 16, i.e. label noise) is copied raw.  lg_set_host_label_rle(ctx, 0) switches
//     the encoding off; lg_host_call_bytes reports what the last call really moved.
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <sched.h>
#include <stdlib.h>
#include <string.h>

#include "lg_internal.cuh"

int lg_process_batch_device_impl(lg_context* c, const int16_t* labels, const float* depth, int frames, const lg_camera* cam,
                                 lg_frame_result* results, float* rec_out, int use_bf16_cnn, void* stream);
int lg_context_dev_alloc(lg_context* c, void** p, size_t bytes);
// lg_rle_host.cpp (host compiler, AVX-512BW / AVX2 / portable by what the CPU has): the runs of one frame; returns their
// number, or 0xFFFFFFFF when they do not fit `run_cap`
extern "C" uint32_t lg_rle_encode_labels(const int16_t* labels, int height, int width, uint32_t* runs, uint32_t run_cap,
                                         uint32_t* rowoff, int isa);

namespace {

constexpr int RLE_SLOTS = 3;            // chunk staging buffers in flight

// A small blocking fork-join pool: run(n, fn) calls fn(0..n-1) on the workers and the calling thread.
class Pool {
 public:
    explicit Pool(int workers) {
        for (int i = 0; i < workers; ++i) th_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> l(mu_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    void run(int n, const std::function<void(int)>& fn) {
        {
            std::lock_guard<std::mutex> l(mu_);
            fn_ = &fn; n_ = n; next_.store(0); left_ = n; ++gen_;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> l(mu_);
        done_.wait(l, [this] { return left_ == 0 && active_ == 0; });     // no worker is still looking at this job
        fn_ = nullptr;
    }

 private:
    void work() {
        for (;;) {
            const int i = next_.fetch_add(1);
            if (i >= n_) return;
            (*fn_)(i);
            std::lock_guard<std::mutex> l(mu_);
            if (--left_ == 0) done_.notify_all();
        }
    }
    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                ++active_;
            }
            work();
            std::lock_guard<std::mutex> l(mu_);
            if (--active_ == 0) done_.notify_all();
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* fn_ = nullptr;
    int n_ = 0, left_ = 0, active_ = 0;
    std::atomic<int> next_{0};
    unsigned long long gen_ = 0;
    bool stop_ = false;
};

struct HostPipe {
    Pool* pool = nullptr;
    int threads = 1;
    int chunk_cap = 0;                  // frames per staging slot
    uint32_t run_cap = 0;               // runs per frame the staging holds
    uint32_t* h_runs[RLE_SLOTS] = {};   // pinned [chunk_cap][run_cap]
    uint32_t* h_pack[RLE_SLOTS] = {};   // pinned [chunk_cap * run_cap]: the chunk's runs packed back to back (what is copied)
    uint32_t* h_rowoff[RLE_SLOTS] = {}; // pinned [chunk_cap][H + 2]: row offsets, [H + 1] = the frame's start in the packed runs; [0] = 0xFFFFFFFF: frame copied raw
    uint32_t* d_runs[RLE_SLOTS] = {};
    uint32_t* d_rowoff[RLE_SLOTS] = {};
    cudaEvent_t copied[RLE_SLOTS] = {};     // the slot's host buffers have been read
    cudaEvent_t expanded[RLE_SLOTS] = {};   // the slot's device buffers have been read
    bool used[RLE_SLOTS] = {};
};

// one warp per image row: the row's runs (first column | label << 16) back into int16 labels
__global__ void __launch_bounds__(256) expand_labels_kernel(const uint32_t* __restrict__ runs_all,
                                                            const uint32_t* __restrict__ rowoff_all, int16_t* __restrict__ out,
                                                            size_t P, int H, int W) {
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const int y = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (y >= H) return;
    const uint32_t* ro = rowoff_all + (size_t)f * (H + 2);
    if (ro[0] == 0xFFFFFFFFu) return;            // this frame's labels were copied as they are
    const uint32_t a = ro[y], b = ro[y + 1];
    const uint32_t* runs = runs_all + ro[H + 1];      // the chunk's runs are packed: this frame's start at ro[H + 1]
    int16_t* orow = out + (size_t)f * P + (size_t)y * W;
    for (uint32_t j0 = a; j0 < b; j0 += 32) {
        const uint32_t mine = j0 + lane < b ? runs[j0 + lane] : 0u;
        const uint32_t next = j0 + lane + 1 < b ? runs[j0 + lane + 1] : (uint32_t)W;     // only its column is used
        const int cnt = (int)min(32u, b - j0);
        for (int k = 0; k < cnt; ++k) {
            const uint32_t r = __shfl_sync(0xFFFFFFFFu, mine, k);
            const int x0 = (int)(r & 0xFFFFu), x1 = (int)(__shfl_sync(0xFFFFFFFFu, next, k) & 0xFFFFu);
            const int16_t v = (int16_t)(r >> 16);
            for (int x = x0 + lane; x < x1; x += 32) orow[x] = v;
        }
    }
}

int allowed_cpus() {
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        const int n = CPU_COUNT(&set);
        if (n > 0) return n;
    }
    const unsigned hc = std::thread::hardware_concurrency();
    return hc ? (int)hc : 1;
}

int ensure_pipe(lg_context* c, int chunk) {
    HostPipe* hp = static_cast<HostPipe*>(c->host_pipe);
    if (hp && hp->chunk_cap >= chunk) return LG_OK;
    if (!hp) {
        hp = new (std::nothrow) HostPipe();
        if (!hp) return LG_E_ARG;
        c->host_pipe = hp;
        // host threads of the label encoder: the CPUs this process may use, shared with the other ranks of the node when a
        // launcher says how many there are (torchrun: LOCAL_WORLD_SIZE); LG_HOST_THREADS overrides
        int t = allowed_cpus();
        const char* lw = getenv("LOCAL_WORLD_SIZE");
        if (lw && atoi(lw) > 1) t /= atoi(lw);
        if (t > 16) t = 16;
        if (t < 2) t = 2;
        const char* e = getenv("LG_HOST_THREADS");
        if (e && atoi(e) > 0) t = atoi(e);
        hp->threads = t;
        hp->pool = new (std::nothrow) Pool(t - 1);
        if (!hp->pool) return LG_E_ARG;
        for (int s = 0; s < RLE_SLOTS; ++s) {
            LG_CUDA(cudaEventCreateWithFlags(&hp->copied[s], cudaEventDisableTiming));
            LG_CUDA(cudaEventCreateWithFlags(&hp->expanded[s], cudaEventDisableTiming));
        }
    }
    // (re)size the staging: only grows, and only while nothing is in flight (the caller synchronises every call)
    hp->run_cap = (uint32_t)(c->P / 16 < 64 ? 64 : c->P / 16);
    for (int s = 0; s < RLE_SLOTS; ++s) {
        if (hp->h_runs[s]) cudaFreeHost(hp->h_runs[s]);
        if (hp->h_rowoff[s]) cudaFreeHost(hp->h_rowoff[s]);
        if (hp->h_pack[s]) cudaFreeHost(hp->h_pack[s]);
        hp->h_runs[s] = nullptr; hp->h_rowoff[s] = nullptr; hp->h_pack[s] = nullptr;
        LG_CUDA(cudaHostAlloc((void**)&hp->h_runs[s], (size_t)chunk * hp->run_cap * sizeof(uint32_t), cudaHostAllocDefault));
        LG_CUDA(cudaHostAlloc((void**)&hp->h_pack[s], (size_t)chunk * hp->run_cap * sizeof(uint32_t), cudaHostAllocDefault));
        LG_CUDA(cudaHostAlloc((void**)&hp->h_rowoff[s], (size_t)chunk * (c->H + 2) * sizeof(uint32_t), cudaHostAllocDefault));
        void* p = nullptr;
        int rc = lg_context_dev_alloc(c, &p, (size_t)chunk * hp->run_cap * sizeof(uint32_t));
        if (rc) return rc;
        hp->d_runs[s] = static_cast<uint32_t*>(p);
        rc = lg_context_dev_alloc(c, &p, (size_t)chunk * (c->H + 2) * sizeof(uint32_t));
        if (rc) return rc;
        hp->d_rowoff[s] = static_cast<uint32_t*>(p);
        hp->used[s] = false;
    }
    hp->chunk_cap = chunk;
    return LG_OK;
}

}  // namespace

void lg_host_pipe_destroy(lg_context* c) {
    HostPipe* hp = static_cast<HostPipe*>(c->host_pipe);
    if (!hp) return;
    delete hp->pool;
    for (int s = 0; s < RLE_SLOTS; ++s) {
        if (hp->h_runs[s]) cudaFreeHost(hp->h_runs[s]);
        if (hp->h_rowoff[s]) cudaFreeHost(hp->h_rowoff[s]);
        if (hp->h_pack[s]) cudaFreeHost(hp->h_pack[s]);
        if (hp->copied[s]) cudaEventDestroy(hp->copied[s]);
        if (hp->expanded[s]) cudaEventDestroy(hp->expanded[s]);
    }
    delete hp;
    c->host_pipe = nullptr;
}

extern "C" int lg_set_host_label_rle(lg_context* c, int on) {
    if (!c) return LG_E_ARG;
    c->host_rle = on ? 1 : 0;
    return LG_OK;
}

extern "C" int lg_host_call_bytes(const lg_context* c, uint64_t* h2d, uint64_t* d2h) {
    if (!c) return LG_E_ARG;
    if (h2d) *h2d = c->last_h2d_bytes;
    if (d2h) *d2h = c->last_d2h_bytes;
    return LG_OK;
}

extern "C" int lg_process_batch_host(lg_context* c, const int16_t* labels_host, const float* depth_host, int frames,
                                     const lg_camera* cam, lg_frame_result* results_host, int use_bf16_cnn, void* stream) {
    if (!c || !labels_host || !depth_host || frames < 1) { lg_set_error("null pointer or empty batch"); return LG_E_ARG; }
    if (frames > c->B) { lg_set_error("batch of %d frames exceeds context capacity %d", frames, c->B); return LG_E_CAPACITY; }
    if (!results_host || !cam) return LG_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (!c->in_labels) {     // first host call of this context: device staging for a full batch
        const size_t Bc = (size_t)c->B;
        void* p = nullptr;
        int rc = lg_context_dev_alloc(c, &p, Bc * c->P * sizeof(int16_t));
        if (rc) return rc;
        c->in_labels = static_cast<int16_t*>(p);
        rc = lg_context_dev_alloc(c, &p, Bc * c->P * sizeof(float));
        if (rc) return rc;
        c->in_depth = static_cast<float*>(p);
        rc = lg_context_dev_alloc(c, &p, Bc * sizeof(lg_frame_result));
        if (rc) return rc;
        c->results_all = static_cast<lg_frame_result*>(p);
    }
    static const int chunk0 = [] { const char* e = getenv("LG_HOST_CHUNK"); const int v = e ? atoi(e) : 0; return v >= 1 ? v : LG_HOST_CHUNK_FRAMES; }();
    int chunk = chunk0;
    while ((frames + chunk - 1) / chunk > LG_MAX_HOST_CHUNKS) chunk *= 2;
    const int n_chunks = (frames + chunk - 1) / chunk;
    static const bool env_off = [] { const char* e = getenv("LG_HOST_RLE"); return e && e[0] == '0'; }();
    const bool rle = c->host_rle && !env_off && c->W <= 0xFFFF;
    HostPipe* hp = nullptr;
    if (rle) {
        int rc = ensure_pipe(c, chunk);
        if (rc) return rc;
        hp = static_cast<HostPipe*>(c->host_pipe);
    }
    const int H = c->H, W = c->W;
    const size_t P = c->P;
    uint64_t h2d = 0;
    LG_CUDA(cudaEventRecord(c->copy_gate, st));                 // staging is free once earlier work on st is done
    LG_CUDA(cudaStreamWaitEvent(c->copy_stream, c->copy_gate, 0));
    // Chunk by chunk: encode the labels (host threads), queue the chunk's copies, queue its kernels.  Everything on the
    // device is asynchronous, so while the host encodes chunk k + 1 the link carries chunk k and the GPU computes k - 1.
    for (int k = 0; k < n_chunks; ++k) {
        const size_t off = (size_t)k * chunk * P;
        const int m = frames - k * chunk < chunk ? frames - k * chunk : chunk;
        const int slot = k % RLE_SLOTS;
        if (rle) {
            // depth needs no preparation: its copy is queued first and crosses the link while the host threads encode the labels
            LG_CUDA(cudaMemcpyAsync(c->in_depth + off, depth_host + off, (size_t)m * P * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
            h2d += (uint64_t)m * P * sizeof(float);
            if (hp->used[slot]) {
                LG_CUDA(cudaEventSynchronize(hp->copied[slot]));                       // the slot's last copies have left the host buffers
                LG_CUDA(cudaStreamWaitEvent(c->copy_stream, hp->expanded[slot], 0));   // and its device buffers have been expanded
            }
            uint32_t* runs = hp->h_runs[slot];
            uint32_t* rowoff = hp->h_rowoff[slot];
            const uint32_t cap = hp->run_cap;
            std::vector<uint32_t> n_runs((size_t)m), base((size_t)m + 1);
            const std::function<void(int)> job = [&](int f) {
                uint32_t* ro = rowoff + (size_t)f * (H + 2);
                n_runs[(size_t)f] = lg_rle_encode_labels(labels_host + off + (size_t)f * P, H, W, runs + (size_t)f * cap, cap, ro, -1);
                if (n_runs[(size_t)f] == 0xFFFFFFFFu) ro[0] = 0xFFFFFFFFu;
            };
            hp->pool->run(m, job);
            // pack the frames' runs back to back: ONE copy per chunk instead of one per frame
            base[0] = 0;
            for (int f = 0; f < m; ++f) base[(size_t)f + 1] = base[(size_t)f] + (n_runs[(size_t)f] == 0xFFFFFFFFu ? 0u : n_runs[(size_t)f]);
            uint32_t* pack = hp->h_pack[slot];
            const std::function<void(int)> move = [&](int f) {
                rowoff[(size_t)f * (H + 2) + H + 1] = base[(size_t)f];
                if (n_runs[(size_t)f] != 0xFFFFFFFFu)
                    memcpy(pack + base[(size_t)f], runs + (size_t)f * cap, (size_t)n_runs[(size_t)f] * sizeof(uint32_t));
            };
            hp->pool->run(m, move);
            LG_CUDA(cudaMemcpyAsync(hp->d_rowoff[slot], rowoff, (size_t)m * (H + 2) * sizeof(uint32_t), cudaMemcpyHostToDevice, c->copy_stream));
            h2d += (uint64_t)m * (H + 2) * sizeof(uint32_t);
            if (base[(size_t)m]) {
                LG_CUDA(cudaMemcpyAsync(hp->d_runs[slot], pack, (size_t)base[(size_t)m] * sizeof(uint32_t), cudaMemcpyHostToDevice, c->copy_stream));
                h2d += (uint64_t)base[(size_t)m] * sizeof(uint32_t);
            }
            for (int f = 0; f < m; ++f) {
                if (n_runs[(size_t)f] == 0xFFFFFFFFu) {          // too many runs: this frame's labels go as they are
                    LG_CUDA(cudaMemcpyAsync(c->in_labels + off + (size_t)f * P, labels_host + off + (size_t)f * P, P * sizeof(int16_t),
                                            cudaMemcpyHostToDevice, c->copy_stream));
                    h2d += P * sizeof(int16_t);
                }
            }
            LG_CUDA(cudaEventRecord(c->copy_ev[k], c->copy_stream));
            LG_CUDA(cudaEventRecord(hp->copied[slot], c->copy_stream));
            hp->used[slot] = true;
            LG_CUDA(cudaStreamWaitEvent(st, c->copy_ev[k], 0));
            expand_labels_kernel<<<dim3((H + 7) / 8, m), 256, 0, st>>>(hp->d_runs[slot], hp->d_rowoff[slot],
                                                                        c->in_labels + off, P, H, W);
            LG_LAUNCH_CHECK();
            LG_CUDA(cudaEventRecord(hp->expanded[slot], st));
        } else {
            LG_CUDA(cudaMemcpyAsync(c->in_labels + off, labels_host + off, (size_t)m * P * sizeof(int16_t), cudaMemcpyHostToDevice, c->copy_stream));
            LG_CUDA(cudaMemcpyAsync(c->in_depth + off, depth_host + off, (size_t)m * P * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
            h2d += (uint64_t)m * P * (sizeof(int16_t) + sizeof(float));
            LG_CUDA(cudaEventRecord(c->copy_ev[k], c->copy_stream));
            LG_CUDA(cudaStreamWaitEvent(st, c->copy_ev[k], 0));
        }
        int rc = lg_process_batch_device_impl(c, c->in_labels + off, c->in_depth + off, m, cam, c->results_all + (size_t)k * chunk,
                                              c->rec_out ? c->rec_out + (size_t)k * chunk * LG_TOP_K * 4 : nullptr, use_bf16_cnn, stream);
        if (rc) return rc;
    }
    LG_CUDA(cudaMemcpyAsync(results_host, c->results_all, sizeof(lg_frame_result) * frames, cudaMemcpyDeviceToHost, st));
    LG_CUDA(cudaStreamSynchronize(st));
    c->last_h2d_bytes = h2d;
    c->last_d2h_bytes = sizeof(lg_frame_result) * (uint64_t)frames;
    return LG_OK;
}
