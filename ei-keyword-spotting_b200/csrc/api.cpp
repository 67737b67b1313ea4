// eikws-b200: C ABI (include/eikws_b200.h).  Thin host layer: model container -> plan -> kernel launches.
// No CPU compute path exists here; every classify/features call launches the CUDA kernel or fails.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "eikws_b200.h"
#include "kernels.h"
#include "model_graph.h"
#include "plan.h"

namespace eikws {
bool capture_compiled_model(const eikws_compiled_model_t *cm, ModelGraph &g, std::string &err);
}

using namespace eikws;

struct eikws_handle {
    int device = 0;
    ModelGraph graph;
    HostPlan host;
    DevicePlan dev;
    int sm_count = 148;
    int ctas_per_sm = 4;    // clip groups (160 threads each) resident per SM
    int clips_per_cta = 2;  // clip groups per CTA: 2 CTAs x 2 groups measured best (profiles/r1_ab_clip_groups.txt)
    int skew_ns = 14000;  // start offset between the CTAs that share an SM (see kernels.cu)
    int pipelined = 0;      // software-pipelined classify kernel (kernels.cu eikws_pipelined_kernel)
    int split = 1;          // two-kernel classify path for int16 clips (kernels.cu eikws_logmel_kernel -> eikws_cepstral_kernel)
    size_t split_chunk_clips = 65536;  // clips per kernel pair: bounds the hand-over scratch (6,480 B per clip) at 425 MB
    // eikws_set_kernel_timing: CUDA events around the two kernels of every split launch (a ring of triples, harvested when it is full
    // or when eikws_split_kernel_ms asks), accumulated per kernel
    static constexpr int kEvRing = 32;
    cudaEvent_t split_ev[kEvRing][3] = {};
    int ev_used = 0;
    double ev_ms[2] = {0.0, 0.0};
    uint64_t ev_launches = 0;
    int kernel_timing = 0;
    cudaMemPool_t pool = nullptr;      // stream-ordered allocations of that scratch: no state shared between callers' streams
    int work_claiming = 1;  // work-claiming schedule of the shortcut kernel (kDyn in kernels.cu)
    int cmvn_shortcut = 1;  // certified CMVN shortcut of the tensor-core variant (kernels.cu cmvn_certified; exact fallback inside the kernel)
    int tensor_core = 1;  // block 1 of the fused classifier as a tcgen05 UMMA (when the plan allows it; +2.4 %, profiles/r1_ab_tensor_core_block1.txt)
    std::atomic<uint64_t> launches{0};
    std::mutex mu;  // serialises the host-buffer and single-clip paths (they share staging buffers) from the first byte staged to the last result copied
    // staging for the host-buffer entry points: a ring of kHostRing chunk slots per buffer (chunk i uses slot i % kHostRing on stream
    // i % 2, so a slot is only ever reused by a later chunk of the SAME stream: stream order is the only synchronisation needed)
    void *d_in = nullptr;
    size_t d_in_bytes = 0;
    float *d_probs = nullptr;
    size_t d_probs_bytes = 0;
    float *d_feat = nullptr;
    size_t d_feat_bytes = 0;
    int8_t *d_qfeat = nullptr;
    size_t d_qfeat_bytes = 0;
    float *h_pinned = nullptr;  // one clip of floats for eikws_run_classifier_signal
    cudaStream_t stream = nullptr, stream2 = nullptr;
    size_t host_chunk_clips = 8192;  // host-buffer path: clips per pipelined chunk (262 MB of int16 PCM)
    static constexpr size_t kHostRing = 4;  // chunk slots in flight (even: see above)
};

namespace {
thread_local std::string t_err;
int fail(int code, const std::string &msg) {
    t_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char *what) {
    t_err = std::string(what) + ": " + cudaGetErrorString(e);
    return EIKWS_ERR_CUDA;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

int ensure(void **p, size_t *have, size_t need) {
    if (*have >= need) return EIKWS_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *have = 0;
    cudaError_t e = cudaMalloc(p, need);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(staging)");
    *have = need;
    return EIKWS_OK;
}

int grid_for(const eikws_handle *h, size_t n_clips) {
    size_t g = static_cast<size_t>(h->sm_count) * h->ctas_per_sm;  // persistent: one wave of resident CTAs
    if (n_clips < g) g = n_clips;
    return static_cast<int>(g ? g : 1);
}

// add the recorded event triples to the per-kernel sums (waits for the launches they belong to)
cudaError_t harvest_split_events(eikws_handle *h) {
    for (int i = 0; i < h->ev_used; i++) {
        cudaError_t e = cudaEventSynchronize(h->split_ev[i][2]);
        float a = 0.0f, b = 0.0f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&a, h->split_ev[i][0], h->split_ev[i][1]);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&b, h->split_ev[i][1], h->split_ev[i][2]);
        if (e != cudaSuccess) {
            h->ev_used = 0;
            return e;
        }
        h->ev_ms[0] += a;
        h->ev_ms[1] += b;
        h->ev_launches++;
    }
    h->ev_used = 0;
    return cudaSuccess;
}

int launch(eikws_handle *h, const void *clips, bool f32, const float *features_in, size_t n, bool run_nn, float *probs,
           float *feat, int8_t *qfeat, cudaStream_t st, float *dbg = nullptr) {
    if (n == 0) return EIKWS_OK;
    if (clips && (reinterpret_cast<uintptr_t>(clips) & 15)) return fail(EIKWS_ERR_BAD_ARG, "clip buffer must be 16-byte aligned (TMA bulk copy)");
    LaunchArgs a;
    a.plan = h->dev.d_plan;
    a.clips = clips;
    a.input_is_f32 = f32;
    a.features_in = features_in;
    a.n_clips = n;
    a.run_nn = run_nn;
    a.nn_fused = h->host.dev.nn.fused.enabled != 0;
    a.nn_float = h->host.dev.nn.float_mode != 0;
    a.nn_tc = h->tensor_core != 0 && h->host.dev.nn.fused.tc_enabled != 0 && dbg == nullptr;
    a.cmvn_certified = h->cmvn_shortcut != 0;
    a.work_claiming = h->work_claiming != 0;
    a.pipelined = h->pipelined != 0;
    if (a.nn_float && qfeat) return fail(EIKWS_ERR_BAD_ARG, "a float32 model has no quantised input tensor");
    a.probs = probs;
    a.features_out = feat;
    a.qfeatures_out = qfeat;
    a.debug_taps = dbg;
    a.grid = grid_for(h, n);
    a.clips_per_cta = h->clips_per_cta;
    a.sm_count = h->sm_count;
    a.skew_ns = (n >= static_cast<size_t>(h->sm_count) * h->ctas_per_sm * 8) ? h->skew_ns : 0;  // only worth it for long launches
    a.pre_cof = h->host.dev.mfcc.pre_cof;
    a.nn_smem_bytes = h->dev.nn_smem_bytes;
    a.stream = st;
    cudaError_t e;
    if (h->split && !h->pipelined && clips && !features_in && run_nn && !feat && !dbg &&
        (a.nn_float || (a.nn_fused && h->tensor_core != 0 && h->host.dev.nn.fused.tc_enabled != 0 && a.cmvn_certified && h->clips_per_cta == 2))) {
        a.nn_tc = !a.nn_float;
        // two kernels per chunk, the log-mel records in between in stream-ordered scratch
        const size_t L = h->graph.labels.size();
        for (size_t off = 0; off < n; off += h->split_chunk_clips) {
            const size_t m = n - off < h->split_chunk_clips ? n - off : h->split_chunk_clips;
            void *scratch = nullptr;
            if ((e = cudaMallocFromPoolAsync(&scratch, split_scratch_bytes(m), h->pool, st)) != cudaSuccess) return cuda_fail(e, "cudaMallocFromPoolAsync(log-mel scratch)");
            a.clips = static_cast<const char *>(clips) + off * static_cast<size_t>(kSamples) * (f32 ? 4 : 2);
            a.n_clips = m;
            a.probs = probs + off * L;
            a.qfeatures_out = qfeat ? qfeat + off * static_cast<size_t>(kFeatures) : nullptr;
            a.split = true;
            a.logmel = static_cast<float *>(scratch);
            a.split_events = nullptr;
            if (h->kernel_timing) {
                if (h->ev_used == eikws_handle::kEvRing && (e = harvest_split_events(h)) != cudaSuccess) return cuda_fail(e, "split kernel timing");
                a.split_events = h->split_ev[h->ev_used++];
            }
            e = launch_run_classifier(a);
            cudaError_t e2 = cudaFreeAsync(scratch, st);
            if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
            if (e2 != cudaSuccess) return cuda_fail(e2, "cudaFreeAsync(log-mel scratch)");
            h->launches += 2;
        }
        return EIKWS_OK;
    }
    e = launch_run_classifier(a);
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    h->launches++;
    return EIKWS_OK;
}
}  // namespace

extern "C" {

const char *eikws_last_error(void) { return t_err.c_str(); }

int eikws_model_from_compiled(const eikws_compiled_model_t *cm, void **blob, size_t *bytes) {
    if (!blob || !bytes) return fail(EIKWS_ERR_BAD_ARG, "null output argument");
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    ModelGraph g;
    std::string err;
    if (!capture_compiled_model(cm, g, err)) return fail(EIKWS_ERR_TFLITE, err);
    std::vector<uint8_t> out;
    serialize_model(g, out);
    void *p = std::malloc(out.size());
    if (!p) return fail(EIKWS_ERR_ALLOC_FAILED, "malloc failed");
    std::memcpy(p, out.data(), out.size());
    *blob = p;
    *bytes = out.size();
    return EIKWS_OK;
}

void eikws_free(void *p) { std::free(p); }

int eikws_create(const void *model_blob, size_t bytes, int device, eikws_handle **out) {
    if (!out) return fail(EIKWS_ERR_BAD_ARG, "null output argument");
    *out = nullptr;
    eikws_handle *h = new (std::nothrow) eikws_handle();
    if (!h) return fail(EIKWS_ERR_ALLOC_FAILED, "out of memory");
    std::string err;
    if (!parse_model(model_blob, bytes, h->graph, err)) {
        delete h;
        return fail(EIKWS_ERR_BAD_ARG, err);
    }
    int rc = build_host_plan(h->graph, h->host, err);
    if (rc != EIKWS_OK) {
        delete h;
        return fail(rc, err);
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        delete h;
        return fail(EIKWS_ERR_CUDA, std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    }
    if (device < 0 || device >= ndev) {
        delete h;
        return fail(EIKWS_ERR_BAD_ARG, "device index out of range");
    }
    h->device = device;
    DeviceGuard guard(device);
    if (!guard.ok) {
        delete h;
        return fail(EIKWS_ERR_CUDA, "cudaSetDevice failed");
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        delete h;
        return cuda_fail(e, "cudaGetDeviceProperties");
    }
    if (prop.major < 10) {
        delete h;
        return fail(EIKWS_ERR_CUDA, "this library contains sm_100a code only; device compute capability is " + std::to_string(prop.major) + "." +
                                        std::to_string(prop.minor));
    }
    h->sm_count = prop.multiProcessorCount;
    if ((e = upload_plan(h->host, h->dev)) != cudaSuccess) {
        free_plan(h->dev);
        delete h;
        return cuda_fail(e, "plan upload");
    }
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        free_plan(h->dev);
        delete h;
        return cuda_fail(e, "cudaStreamCreate");
    }
    {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        uint64_t keep = UINT64_MAX;  // freed scratch stays in the pool: the next launch's allocation costs microseconds
        if ((e = cudaMemPoolCreate(&h->pool, &props)) != cudaSuccess || (e = cudaMemPoolSetAttribute(h->pool, cudaMemPoolAttrReleaseThreshold, &keep)) != cudaSuccess) {
            // no stream-ordered allocator on this device / driver configuration: the handle still works, on the single fused kernel (which needs no scratch)
            (void)cudaGetLastError();
            if (h->pool) cudaMemPoolDestroy(h->pool);
            h->pool = nullptr;
            h->split = 0;
        }
    }
    *out = h;
    return EIKWS_OK;
}

void eikws_destroy(eikws_handle *h) {
    if (!h) return;
    DeviceGuard guard(h->device);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->d_in) cudaFree(h->d_in);
    if (h->d_probs) cudaFree(h->d_probs);
    if (h->d_feat) cudaFree(h->d_feat);
    if (h->d_qfeat) cudaFree(h->d_qfeat);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    for (auto &triple : h->split_ev)
        for (cudaEvent_t ev : triple)
            if (ev) cudaEventDestroy(ev);
    if (h->pool) {
        cudaDeviceSynchronize();  // stream-ordered frees of the scratch may still be pending on callers' streams
        cudaMemPoolDestroy(h->pool);
    }
    free_plan(h->dev);
    delete h;
}

int eikws_label_count(const eikws_handle *h) { return h ? static_cast<int>(h->graph.labels.size()) : 0; }
int eikws_feature_count(const eikws_handle *h) { return h ? static_cast<int>(h->graph.nn_input_frame_size) : 0; }
int eikws_raw_sample_count(const eikws_handle *h) { return h ? static_cast<int>(h->graph.raw_sample_count) : 0; }
int eikws_device(const eikws_handle *h) { return h ? h->device : -1; }
const char *eikws_label(const eikws_handle *h, int i) {
    if (!h || i < 0 || i >= static_cast<int>(h->graph.labels.size())) return nullptr;
    return h->graph.labels[i].c_str();
}
uint64_t eikws_launch_count(const eikws_handle *h) { return h ? h->launches.load() : 0; }

int eikws_set_ctas_per_sm(eikws_handle *h, int n) {  // tuning knob (not in the public header)
    if (!h || n < 1 || n > 8) return EIKWS_ERR_BAD_ARG;
    h->ctas_per_sm = n;
    return EIKWS_OK;
}
int eikws_set_clips_per_cta(eikws_handle *h, int n) {  // tuning knob (not in the public header)
    if (!h || (n != 1 && n != 2 && n != 4)) return EIKWS_ERR_BAD_ARG;
    h->clips_per_cta = n;
    return EIKWS_OK;
}
int eikws_set_tensor_core(eikws_handle *h, int on) {  // tuning knob (not in the public header)
    if (!h || (on != 0 && on != 1)) return EIKWS_ERR_BAD_ARG;
    h->tensor_core = on;
    return EIKWS_OK;
}
int eikws_set_cmvn_shortcut(eikws_handle *h, int on) {  // tuning knob (not in the public header): 0 = every CMVN chain with the reference's operation sequence
    if (!h || (on != 0 && on != 1)) return EIKWS_ERR_BAD_ARG;
    h->cmvn_shortcut = on;
    return EIKWS_OK;
}
int eikws_set_work_claiming(eikws_handle *h, int on) {  // tuning knob (not in the public header)
    if (!h || (on != 0 && on != 1)) return EIKWS_ERR_BAD_ARG;
    h->work_claiming = on;
    return EIKWS_OK;
}
int eikws_set_pipelined(eikws_handle *h, int on) {  // tuning knob: the software-pipelined classify kernel
    if (!h || (on != 0 && on != 1)) return EIKWS_ERR_BAD_ARG;
    h->pipelined = on;
    return EIKWS_OK;
}
int eikws_set_split(eikws_handle *h, int on) {  // tuning knob: the two-kernel classify path (spectral kernel + cepstral / classifier kernel)
    if (!h || (on != 0 && on != 1)) return EIKWS_ERR_BAD_ARG;
    if (on && !h->pool) return fail(EIKWS_ERR_UNSUPPORTED, "the two-kernel path needs the stream-ordered allocator (cudaMemPoolCreate failed at eikws_create)");
    h->split = on;
    return EIKWS_OK;
}
// measurement aid (bench.py's roofline of the dominant kernel): with timing on, every split launch records CUDA events around its two
// kernels on the launch stream.  eikws_split_kernel_ms waits for the recorded launches and returns the AVERAGE milliseconds per launch
// of {spectral kernel, cepstral / classifier kernel} since timing was switched on (or since the last call), and how many launches that was.
int eikws_set_kernel_timing(eikws_handle *h, int on) {
    if (!h || (on != 0 && on != 1)) return EIKWS_ERR_BAD_ARG;
    DeviceGuard guard(h->device);
    std::lock_guard<std::mutex> lock(h->mu);
    if (on)
        for (auto &triple : h->split_ev)
            for (cudaEvent_t &ev : triple)
                if (!ev) {
                    cudaError_t e = cudaEventCreate(&ev);
                    if (e != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
                }
    h->kernel_timing = on;
    h->ev_used = 0;
    h->ev_ms[0] = h->ev_ms[1] = 0.0;
    h->ev_launches = 0;
    return EIKWS_OK;
}
int eikws_split_kernel_ms(eikws_handle *h, float *ms2, uint64_t *launches) {
    if (!h || !ms2) return EIKWS_ERR_BAD_ARG;
    DeviceGuard guard(h->device);
    std::lock_guard<std::mutex> lock(h->mu);
    cudaError_t e = harvest_split_events(h);
    if (e != cudaSuccess) return cuda_fail(e, "split kernel timing");
    if (h->ev_launches == 0) return fail(EIKWS_ERR_BAD_ARG, "no split launch has been timed (kernel timing off, or the launches took another path)");
    ms2[0] = static_cast<float>(h->ev_ms[0] / static_cast<double>(h->ev_launches));
    ms2[1] = static_cast<float>(h->ev_ms[1] / static_cast<double>(h->ev_launches));
    if (launches) *launches = h->ev_launches;
    h->ev_ms[0] = h->ev_ms[1] = 0.0;
    h->ev_launches = 0;
    return EIKWS_OK;
}
int eikws_set_skew_ns(eikws_handle *h, int ns) {  // tuning knob (not in the public header)
    if (!h || ns < 0 || ns > 1000000) return EIKWS_ERR_BAD_ARG;
    h->skew_ns = ns;
    return EIKWS_OK;
}

// ---- device-buffer entry points -------------------------------------------------------------------------
int eikws_classify_i16_device(eikws_handle *h, const int16_t *d_pcm, size_t n, float *d_probs, void *stream) {
    if (!h || !d_pcm || !d_probs) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    DeviceGuard guard(h->device);
    return launch(h, d_pcm, false, nullptr, n, true, d_probs, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}
int eikws_classify_f32_device(eikws_handle *h, const float *d_samples, size_t n, float *d_probs, void *stream) {
    if (!h || !d_samples || !d_probs) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    DeviceGuard guard(h->device);
    return launch(h, d_samples, true, nullptr, n, true, d_probs, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}
int eikws_features_i16_device(eikws_handle *h, const int16_t *d_pcm, size_t n, float *d_features, int8_t *d_q, void *stream) {
    if (!h || !d_pcm || (!d_features && !d_q)) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    DeviceGuard guard(h->device);
    return launch(h, d_pcm, false, nullptr, n, false, nullptr, d_features, d_q, static_cast<cudaStream_t>(stream));
}
int eikws_features_f32_device(eikws_handle *h, const float *d_samples, size_t n, float *d_features, int8_t *d_q, void *stream) {
    if (!h || !d_samples || (!d_features && !d_q)) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    DeviceGuard guard(h->device);
    return launch(h, d_samples, true, nullptr, n, false, nullptr, d_features, d_q, static_cast<cudaStream_t>(stream));
}
int eikws_infer_device(eikws_handle *h, const float *d_features, size_t n, float *d_probs, void *stream) {
    if (!h || !d_features || !d_probs) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    DeviceGuard guard(h->device);
    return launch(h, nullptr, false, d_features, n, true, d_probs, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}
// debug/parity tap: classify and also return float + int8 features
int eikws_classify_taps_i16_device(eikws_handle *h, const int16_t *d_pcm, size_t n, float *d_probs, float *d_features, int8_t *d_q,
                                   void *stream) {
    if (!h || !d_pcm || !d_probs) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    DeviceGuard guard(h->device);
    return launch(h, d_pcm, false, nullptr, n, true, d_probs, d_features, d_q, static_cast<cudaStream_t>(stream));
}

// Core/Src/main.cpp:507-521: pcm[i] = (int16_t)(i2s[skip * i] >> shift)   (firmware: skip 4, shift 8)
int eikws_decimate_i2s_device(eikws_handle *h, const int32_t *d_i2s, size_t n_out, int skip, int shift, int16_t *d_pcm, void *stream) {
    if (!h || !d_i2s || !d_pcm || skip < 1 || shift < 0 || shift > 31) return fail(EIKWS_ERR_BAD_ARG, "bad argument");
    if (reinterpret_cast<uintptr_t>(d_pcm) & 15) return fail(EIKWS_ERR_BAD_ARG, "output buffer must be 16-byte aligned");
    DeviceGuard guard(h->device);
    cudaError_t e = launch_decimate_i2s(d_i2s, n_out, skip, shift, d_pcm, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "decimate launch");
    h->launches++;
    return EIKWS_OK;
}

// mix_audio (dataset-curation.py:93-137) for already-16-kHz float32 material, then the PCM_16 conversion of sf.write (:190-206):
// d_pcm[c][i] = PCM16(0.5 * word_vol * word[c][i] + 0.5 * bg_vol * bg[bg_start[c] + i]).  d_words may be NULL (the script's
// background-only clips, word_path=None); word c holds word_len[c] valid samples (zero-padded / truncated to one second).
int eikws_mix_audio_device(eikws_handle *h, const float *d_words, const uint32_t *d_word_len, size_t word_stride, const float *d_bg, size_t bg_len,
                           const uint32_t *d_bg_start, size_t max_bg_start, double word_vol, double bg_vol, size_t n_clips, int16_t *d_pcm,
                           void *stream) {
    if (!h || !d_bg || !d_bg_start || !d_pcm || (d_words && !d_word_len)) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    if (bg_len < static_cast<size_t>(kSamples) || max_bg_start > bg_len - kSamples)
        return fail(EIKWS_ERR_BAD_ARG, "background track shorter than one clip past the largest start offset");
    if (d_words && ((word_stride & 1) || (reinterpret_cast<uintptr_t>(d_words) & 7))) return fail(EIKWS_ERR_BAD_ARG, "word rows must be 8-byte aligned with an even stride");
    if (reinterpret_cast<uintptr_t>(d_pcm) & 3) return fail(EIKWS_ERR_BAD_ARG, "output buffer must be 4-byte aligned");
    if (n_clips == 0) return EIKWS_OK;
    DeviceGuard guard(h->device);
    // 0.5 * word_vol is a Python float product (double); 0.5 * bg_vol multiplies a float32 array, i.e. is rounded to float32 first
    const double half_word = 0.5 * word_vol;
    const float half_bg = static_cast<float>(0.5 * bg_vol);
    cudaError_t e = launch_mix_audio(d_words, d_word_len, word_stride, d_bg, d_bg_start, half_word, half_bg, n_clips, d_pcm, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "mix_audio launch");
    h->launches++;
    return EIKWS_OK;
}

int eikws_synth_i16_device(eikws_handle *h, int16_t *d_pcm, size_t n, uint64_t first_clip, uint64_t seed, void *stream) {
    if (!h || !d_pcm) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    DeviceGuard guard(h->device);
    cudaError_t e = launch_synth(d_pcm, n, first_clip, seed, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "synth launch");
    return EIKWS_OK;
}

// ---- host-buffer entry points (synchronous) ---------------------------------------------------------------
// Host-buffer path: the batch is cut into chunks that alternate between two streams, so the H2D copy of chunk i+1
// (PCIe) overlaps the kernel and the D2H of chunk i.  Device staging is a ring of kHostRing chunk slots (not the whole
// batch): memory stays bounded for any n, results land in the caller's buffers in place.
// The caller holds h->mu and has selected the device.
static int host_run_locked(eikws_handle *h, const void *in, size_t in_bytes_per_clip, bool f32, const float *features_in, size_t n, bool run_nn,
                           float *probs, float *features, int8_t *qfeatures) {
    if (n == 0) return EIKWS_OK;
    const size_t L = h->graph.labels.size(), F = h->graph.nn_input_frame_size;
    int rc;
    cudaError_t e;
    if (!h->stream2 && (e = cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking)) != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
    const size_t chunk = h->host_chunk_clips;
    const size_t slots = (n + chunk - 1) / chunk < eikws_handle::kHostRing ? (n + chunk - 1) / chunk : eikws_handle::kHostRing;
    const size_t cap = (n < chunk ? n : chunk) * slots;  // clips the ring holds
    if (features_in) {
        if ((rc = ensure(reinterpret_cast<void **>(&h->d_feat), &h->d_feat_bytes, cap * F * 4))) return rc;
    } else {
        if ((rc = ensure(&h->d_in, &h->d_in_bytes, cap * in_bytes_per_clip))) return rc;
        if (features && (rc = ensure(reinterpret_cast<void **>(&h->d_feat), &h->d_feat_bytes, cap * F * 4))) return rc;
    }
    if (probs && (rc = ensure(reinterpret_cast<void **>(&h->d_probs), &h->d_probs_bytes, cap * L * 4))) return rc;
    if (qfeatures && (rc = ensure(reinterpret_cast<void **>(&h->d_qfeat), &h->d_qfeat_bytes, cap * F))) return rc;
    rc = EIKWS_OK;
    size_t ci = 0;
    for (size_t off = 0; off < n && rc == EIKWS_OK; off += chunk, ci++) {
        const size_t m = n - off < chunk ? n - off : chunk;
        const size_t so = (ci % slots) * chunk;  // first clip of this chunk's ring slot
        cudaStream_t st = (ci & 1) ? h->stream2 : h->stream;
        const void *d_src = nullptr;
        if (features_in) {
            if ((e = cudaMemcpyAsync(h->d_feat + so * F, features_in + off * F, m * F * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) {
                rc = cuda_fail(e, "H2D features");
                break;
            }
        } else {
            uint8_t *dst = static_cast<uint8_t *>(h->d_in) + so * in_bytes_per_clip;
            if ((e = cudaMemcpyAsync(dst, static_cast<const uint8_t *>(in) + off * in_bytes_per_clip, m * in_bytes_per_clip, cudaMemcpyHostToDevice,
                                     st)) != cudaSuccess) {
                rc = cuda_fail(e, "H2D clips");
                break;
            }
            d_src = dst;
        }
        rc = launch(h, d_src, f32, features_in ? h->d_feat + so * F : nullptr, m, run_nn, probs ? h->d_probs + so * L : nullptr,
                    (features && !features_in) ? h->d_feat + so * F : nullptr, qfeatures ? h->d_qfeat + so * F : nullptr, st);
        if (rc) break;
        if (probs && (e = cudaMemcpyAsync(probs + off * L, h->d_probs + so * L, m * L * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
            rc = cuda_fail(e, "D2H probs");
        else if (features && !features_in &&
                 (e = cudaMemcpyAsync(features + off * F, h->d_feat + so * F, m * F * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
            rc = cuda_fail(e, "D2H features");
        else if (qfeatures && (e = cudaMemcpyAsync(qfeatures + off * F, h->d_qfeat + so * F, m * F, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
            rc = cuda_fail(e, "D2H qfeatures");
    }
    // always drain both streams, also on an error path: queued D2H copies must not write into the caller's buffers after the return
    const cudaError_t e1 = cudaStreamSynchronize(h->stream), e2 = cudaStreamSynchronize(h->stream2);
    if (rc != EIKWS_OK) return rc;
    if (e1 != cudaSuccess) return cuda_fail(e1, "kernel execution");
    if (e2 != cudaSuccess) return cuda_fail(e2, "kernel execution");
    return EIKWS_OK;
}
static int host_run(eikws_handle *h, const void *in, size_t in_bytes_per_clip, bool f32, const float *features_in, size_t n, bool run_nn,
                    float *probs, float *features, int8_t *qfeatures) {
    if (n == 0) return EIKWS_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    if (!guard.ok) return fail(EIKWS_ERR_CUDA, "cudaSetDevice failed");
    return host_run_locked(h, in, in_bytes_per_clip, f32, features_in, n, run_nn, probs, features, qfeatures);
}
// one clip pulled through the signal callback into the handle's pinned buffer, then classified / transformed -- the lock is held from
// the first byte the callback writes to the last result byte copied back, so concurrent callers cannot see each other's clip
}  // extern "C"
template <class Run>
static int signal_run(eikws_handle *h, eikws_get_data_fn get_data, size_t total_length, Run run) {
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    if (!guard.ok) return fail(EIKWS_ERR_CUDA, "cudaSetDevice failed");
    if (!h->h_pinned) {
        cudaError_t e = cudaMallocHost(reinterpret_cast<void **>(&h->h_pinned), sizeof(float) * kSamples);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMallocHost");
    }
    if (get_data(0, total_length, h->h_pinned) != 0) return fail(EIKWS_ERR_DSP, "signal get_data callback failed");
    return run(h->h_pinned);
}
extern "C" {

int eikws_classify_i16_host(eikws_handle *h, const int16_t *pcm, size_t n, float *probs) {
    if (!h || !pcm || !probs) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    return host_run(h, pcm, static_cast<size_t>(kSamples) * 2, false, nullptr, n, true, probs, nullptr, nullptr);
}
int eikws_classify_f32_host(eikws_handle *h, const float *samples, size_t n, float *probs) {
    if (!h || !samples || !probs) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    return host_run(h, samples, static_cast<size_t>(kSamples) * 4, true, nullptr, n, true, probs, nullptr, nullptr);
}
int eikws_features_i16_host(eikws_handle *h, const int16_t *pcm, size_t n, float *features, int8_t *qfeatures) {
    if (!h || !pcm || (!features && !qfeatures)) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    return host_run(h, pcm, static_cast<size_t>(kSamples) * 2, false, nullptr, n, false, nullptr, features, qfeatures);
}
int eikws_features_f32_host(eikws_handle *h, const float *samples, size_t n, float *features, int8_t *qfeatures) {
    if (!h || !samples || (!features && !qfeatures)) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    return host_run(h, samples, static_cast<size_t>(kSamples) * 4, true, nullptr, n, false, nullptr, features, qfeatures);
}
int eikws_infer_host(eikws_handle *h, const float *features, size_t n, float *probs) {
    if (!h || !features || !probs) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    return host_run(h, nullptr, 0, false, features, n, true, probs, nullptr, nullptr);
}
int eikws_classify_taps_i16_host(eikws_handle *h, const int16_t *pcm, size_t n, float *probs, float *features, int8_t *qfeatures) {
    if (!h || !pcm || !probs) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    return host_run(h, pcm, static_cast<size_t>(kSamples) * 2, false, nullptr, n, true, probs, features, qfeatures);
}

// ---- the sibling MFE DSP block (extract_mfe_features of the reference's newer SDK copy, L432 ei_run_dsp.h:369-418) ---------
// Geometry = the impulse's MFCC block (frame length/stride, filters, FFT, band, window); features [n][49 * 32].
int eikws_mfe_feature_count(const eikws_handle *h) { return h ? kFrames * kFilters : 0; }

static int launch_mfe_on(eikws_handle *h, const void *d_clips, bool f32, size_t n, float *d_out, cudaStream_t st) {
    if (n == 0) return EIKWS_OK;
    if (reinterpret_cast<uintptr_t>(d_clips) & 15) return fail(EIKWS_ERR_BAD_ARG, "clip buffer must be 16-byte aligned (TMA bulk copy)");
    MfeArgs a;
    a.plan = h->dev.d_plan;
    a.clips = d_clips;
    a.input_is_f32 = f32;
    a.n_clips = n;
    a.out = d_out;
    a.grid = grid_for(h, n);
    a.stream = st;
    cudaError_t e = launch_mfe(a);
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    h->launches++;
    return EIKWS_OK;
}
int eikws_mfe_i16_device(eikws_handle *h, const int16_t *d_pcm, size_t n, float *d_features, void *stream) {
    if (!h || !d_pcm || !d_features) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    DeviceGuard guard(h->device);
    return launch_mfe_on(h, d_pcm, false, n, d_features, static_cast<cudaStream_t>(stream));
}
int eikws_mfe_f32_device(eikws_handle *h, const float *d_samples, size_t n, float *d_features, void *stream) {
    if (!h || !d_samples || !d_features) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    DeviceGuard guard(h->device);
    return launch_mfe_on(h, d_samples, true, n, d_features, static_cast<cudaStream_t>(stream));
}
static int mfe_host_locked(eikws_handle *h, const void *in, size_t bytes_per_clip, bool f32, size_t n, float *features) {
    if (n == 0) return EIKWS_OK;
    const size_t F = static_cast<size_t>(kFrames) * kFilters;
    int rc;
    cudaError_t e;
    if ((rc = ensure(&h->d_in, &h->d_in_bytes, n * bytes_per_clip))) return rc;
    if ((rc = ensure(reinterpret_cast<void **>(&h->d_feat), &h->d_feat_bytes, n * F * 4))) return rc;
    if ((e = cudaMemcpyAsync(h->d_in, in, n * bytes_per_clip, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess) return cuda_fail(e, "H2D clips");
    if ((rc = launch_mfe_on(h, h->d_in, f32, n, h->d_feat, h->stream))) return rc;
    if ((e = cudaMemcpyAsync(features, h->d_feat, n * F * 4, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess) return cuda_fail(e, "D2H features");
    if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return cuda_fail(e, "kernel execution");
    return EIKWS_OK;
}
static int mfe_host(eikws_handle *h, const void *in, size_t bytes_per_clip, bool f32, size_t n, float *features) {
    if (n == 0) return EIKWS_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    if (!guard.ok) return fail(EIKWS_ERR_CUDA, "cudaSetDevice failed");
    return mfe_host_locked(h, in, bytes_per_clip, f32, n, features);
}
int eikws_mfe_i16_host(eikws_handle *h, const int16_t *pcm, size_t n, float *features) {
    if (!h || !pcm || !features) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    return mfe_host(h, pcm, static_cast<size_t>(kSamples) * 2, false, n, features);
}
int eikws_mfe_f32_host(eikws_handle *h, const float *samples, size_t n, float *features) {
    if (!h || !samples || !features) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    return mfe_host(h, samples, static_cast<size_t>(kSamples) * 4, true, n, features);
}
// extract_mfe_features(signal_t*, matrix_t*, void *config) through the pull callback; cfg mirrors ei_dsp_config_mfe_t
// (L432 model-parameters/model_metadata.h:103-112) and must describe the geometry the kernels are specialised for
int eikws_extract_mfe_signal(eikws_handle *h, eikws_get_data_fn get_data, size_t total_length, const eikws_mfe_config *cfg, float *features,
                             size_t capacity) {
    if (!h || !get_data || !features || !cfg) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    const MfccConfig &m = h->graph.mfcc;
    if (cfg->axes != 1) return fail(EIKWS_ERR_DSP, "MFE block: axes must be 1 (ei_run_dsp.h:372-374)");
    if (cfg->frame_length != m.frame_length || cfg->frame_stride != m.frame_stride || cfg->num_filters != m.num_filters ||
        cfg->fft_length != m.fft_length || cfg->low_frequency != m.low_frequency || cfg->high_frequency != m.high_frequency ||
        cfg->win_size != m.win_size)
        return fail(EIKWS_ERR_UNSUPPORTED, "MFE block: only the geometry of the impulse's MFCC block is implemented");
    if (total_length != h->graph.raw_sample_count) return fail(EIKWS_ERR_DSP, "signal length does not match EI_CLASSIFIER_RAW_SAMPLE_COUNT");
    if (capacity < static_cast<size_t>(kFrames) * kFilters) return fail(EIKWS_ERR_DSP, "MFE block: output matrix too small (ei_run_dsp.h:384-388)");
    return signal_run(h, get_data, total_length,
                      [&](const float *clip) { return mfe_host_locked(h, clip, static_cast<size_t>(kSamples) * 4, true, 1, features); });
}

// ---- single clip through the reference's pull callback -----------------------------------------------------
// run_classifier (ei_run_classifier.h:650-714).  The reference pulls the signal in ~98 pieces
// ((off,320) frames and (off-1,1) history samples); any call pattern is legal, so the whole clip is pulled once.
int eikws_run_classifier_signal(eikws_handle *h, eikws_get_data_fn get_data, size_t total_length, float *values, int *t_dsp_ms,
                                int *t_cls_ms) {
    if (!h || !get_data || !values) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    // extract_mfcc_features fails with EIDSP_MATRIX_SIZE_MISMATCH -> EI_IMPULSE_DSP_ERROR when the signal would
    // produce more features than the block owns (ei_run_dsp.h:279-283); the kernel is specialised to exactly
    // EI_CLASSIFIER_RAW_SAMPLE_COUNT samples.
    if (total_length != h->graph.raw_sample_count) return fail(EIKWS_ERR_DSP, "signal length does not match EI_CLASSIFIER_RAW_SAMPLE_COUNT");
    auto t0 = std::chrono::steady_clock::now();
    int rc = signal_run(h, get_data, total_length, [&](const float *clip) {
        return host_run_locked(h, clip, static_cast<size_t>(kSamples) * 4, true, nullptr, 1, true, values, nullptr, nullptr);
    });
    auto t1 = std::chrono::steady_clock::now();
    // the fused kernel does DSP and classification in one launch; the whole latency is reported as dsp
    if (t_dsp_ms) *t_dsp_ms = static_cast<int>(std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count());
    if (t_cls_ms) *t_cls_ms = 0;
    return rc;
}

int eikws_extract_mfcc_signal(eikws_handle *h, eikws_get_data_fn get_data, size_t total_length, float *features) {
    if (!h || !get_data || !features) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    if (total_length != h->graph.raw_sample_count) return fail(EIKWS_ERR_DSP, "signal length does not match EI_CLASSIFIER_RAW_SAMPLE_COUNT");
    return signal_run(h, get_data, total_length, [&](const float *clip) {
        return host_run_locked(h, clip, static_cast<size_t>(kSamples) * 4, true, nullptr, 1, false, nullptr, features, nullptr);
    });
}

// ---- continuous mode: many audio streams advancing one slice per call ---------------------------------------------
// run_classifier_continuous (ei_run_classifier.h:184-282).  The reference serves ONE stream with static state; here
// n_streams streams advance in lock step, so the bookkeeping the reference keeps in statics (slice_offset,
// feature_buffer_full :116-121, extract_mfcc_per_slice_features' first_run ei_run_dsp.h:313, the MAF index) is shared.
struct eikws_streams {
    eikws_handle *h = nullptr;
    size_t n = 0;
    int slices = 4, slice_size = 0, maf_len = 2;
    bool first_run = false, window_full = false;
    size_t slice_offset = 0;
    int maf_idx = 0;
    float *d_features = nullptr, *d_maf_buf = nullptr, *d_maf_sum = nullptr;
    int16_t *d_slices = nullptr;  // staging for the host entry point
    float *d_probs = nullptr;
};

void eikws_streams_destroy(eikws_streams *s) {
    if (!s) return;
    DeviceGuard guard(s->h->device);
    if (s->d_features) cudaFree(s->d_features);
    if (s->d_maf_buf) cudaFree(s->d_maf_buf);
    if (s->d_maf_sum) cudaFree(s->d_maf_sum);
    if (s->d_slices) cudaFree(s->d_slices);
    if (s->d_probs) cudaFree(s->d_probs);
    delete s;
}

// power-up state: run_classifier_init (:164-172) plus what the reference can only reset by restarting
int eikws_streams_reset(eikws_streams *s) {
    if (!s) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    DeviceGuard guard(s->h->device);
    const size_t L = s->h->graph.labels.size();
    cudaError_t e;
    // Stream-ordering contract: the state is cleared on the handle's own stream and the call returns only after the clears have
    // completed, so a following eikws_streams_push_*_device on ANY stream sees the power-up state.  (A plain cudaMemset runs on
    // the legacy default stream, which the non-blocking compute streams do not synchronise with.)
    cudaStream_t st = s->h->stream;
    if ((e = cudaMemsetAsync(s->d_features, 0, s->n * kFeatures * 4, st)) != cudaSuccess) return cuda_fail(e, "cudaMemset");
    if ((e = cudaMemsetAsync(s->d_maf_buf, 0, s->n * L * s->maf_len * 4, st)) != cudaSuccess) return cuda_fail(e, "cudaMemset");
    if ((e = cudaMemsetAsync(s->d_maf_sum, 0, s->n * L * 4, st)) != cudaSuccess) return cuda_fail(e, "cudaMemset");
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return cuda_fail(e, "cudaMemset");
    s->first_run = false;
    s->window_full = false;
    s->slice_offset = 0;
    s->maf_idx = 0;
    return EIKWS_OK;
}

int eikws_streams_create(eikws_handle *h, size_t n_streams, int slices_per_window, eikws_streams **out) {
    if (!h || !out || n_streams == 0 || slices_per_window < 2 || kSamples % slices_per_window) return fail(EIKWS_ERR_BAD_ARG, "bad argument");
    *out = nullptr;
    if (!h->host.dev.nn.fused.enabled) return fail(EIKWS_ERR_UNSUPPORTED, "continuous mode needs the fused int8 classifier plan");
    const int slice = kSamples / slices_per_window;
    if ((slice * 2) % 16 || slice < 2 * kFrameLen) return fail(EIKWS_ERR_UNSUPPORTED, "unsupported slice size");
    eikws_streams *s = new (std::nothrow) eikws_streams();
    if (!s) return fail(EIKWS_ERR_ALLOC_FAILED, "out of memory");
    s->h = h;
    s->n = n_streams;
    s->slices = slices_per_window;
    s->slice_size = slice;
    s->maf_len = slices_per_window >> 1;
    DeviceGuard guard(h->device);
    const size_t L = h->graph.labels.size();
    cudaError_t e;
    if ((e = cudaMalloc(reinterpret_cast<void **>(&s->d_features), n_streams * kFeatures * 4)) != cudaSuccess ||
        (e = cudaMalloc(reinterpret_cast<void **>(&s->d_maf_buf), n_streams * L * s->maf_len * 4)) != cudaSuccess ||
        (e = cudaMalloc(reinterpret_cast<void **>(&s->d_maf_sum), n_streams * L * 4)) != cudaSuccess) {
        eikws_streams_destroy(s);
        return cuda_fail(e, "cudaMalloc(stream state)");
    }
    int rc = eikws_streams_reset(s);
    if (rc) {
        eikws_streams_destroy(s);
        return rc;
    }
    *out = s;
    return EIKWS_OK;
}

int eikws_streams_slice_size(const eikws_streams *s) { return s ? s->slice_size : 0; }

// One slice for every stream.  d_slices [n_streams][slice_size] int16, d_probs [n_streams][labels] (written only when
// *has_result becomes 1: the window is full and the values are the moving-average-filtered probabilities).
// `beyond` = the float the application's callback returns for indices past the slice (the reference reads one such
// sample per slice, see oracle/ref_harness.cpp); a zero-padded buffer gives 0.
static int streams_push_device(eikws_streams *s, const void *d_slices, bool f32, float beyond, float *d_probs, int *has_result, void *stream) {
    if (!s || !d_slices || !d_probs || !has_result) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    if (reinterpret_cast<uintptr_t>(d_slices) & 15) return fail(EIKWS_ERR_BAD_ARG, "slice buffer must be 16-byte aligned");
    eikws_handle *h = s->h;
    DeviceGuard guard(h->device);
    const MfccConfig &c = h->graph.mfcc;
    // the stream's new bookkeeping is computed into locals and committed only after the launch has succeeded, so a refused or
    // failed call leaves host state and device window in step
    int total_length = s->slice_size;
    if (s->first_run) total_length += static_cast<int>(c.frame_length * static_cast<float>(c.sample_rate));  // ei_run_dsp.h:322-324
    const int n_frames = (total_length - kFrameLen) / kFrameStride;  // calculate_no_of_stack_frames (processing.hpp:260-284)
    const size_t feature_size = static_cast<size_t>(n_frames) * kCepstra;
    if (s->slice_offset + feature_size > kFeatures) return fail(EIKWS_ERR_DSP, "Would write outside feature buffer");
    const size_t offset_now = s->slice_offset;
    size_t new_offset = s->slice_offset;
    bool new_full = s->window_full;
    if (!new_full) {  // ei_run_classifier.h:230-238
        new_offset += feature_size;
        if (new_offset > kFeatures - feature_size) {
            new_full = true;
            new_offset -= feature_size;
        }
    }
    ContinuousArgs a;
    a.plan = h->dev.d_plan;
    a.slices = d_slices;
    a.input_is_f32 = f32;
    a.slice_size = s->slice_size;
    a.n_frames = n_frames;
    a.total_length = total_length;
    a.beyond = beyond;
    a.n_streams = s->n;
    a.state_features = s->d_features;
    a.maf_buf = s->d_maf_buf;
    a.maf_sum = s->d_maf_sum;
    a.slice_offset = static_cast<int>(offset_now);
    a.window_full = new_full ? 1 : 0;
    a.maf_idx = s->maf_idx;
    a.maf_len = s->maf_len;
    a.cmvn_certified = h->cmvn_shortcut != 0;
    a.probs = d_probs;
    a.sm_count = h->sm_count;
    size_t g = static_cast<size_t>(h->sm_count) * 4;
    a.grid = static_cast<int>(s->n < g ? s->n : g);
    a.stream = static_cast<cudaStream_t>(stream);
    cudaError_t e = launch_continuous(a);
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    h->launches++;
    s->first_run = true;
    s->slice_offset = new_offset;
    s->window_full = new_full;
    *has_result = new_full ? 1 : 0;
    if (new_full && ++s->maf_idx >= s->maf_len) s->maf_idx = 0;
    return EIKWS_OK;
}

int eikws_streams_push_i16_device(eikws_streams *s, const int16_t *d_slices, float beyond, float *d_probs, int *has_result, void *stream) {
    return streams_push_device(s, d_slices, false, beyond, d_probs, has_result, stream);
}
int eikws_streams_push_f32_device(eikws_streams *s, const float *d_slices, float beyond, float *d_probs, int *has_result, void *stream) {
    return streams_push_device(s, d_slices, true, beyond, d_probs, has_result, stream);
}

static int streams_push_host(eikws_streams *s, const void *slices, bool f32, float beyond, float *probs, int *has_result) {
    if (!s || !slices || !probs || !has_result) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    eikws_handle *h = s->h;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    const size_t L = h->graph.labels.size(), bytes = s->n * static_cast<size_t>(s->slice_size) * (f32 ? 4 : 2);
    cudaError_t e;
    if (!s->d_slices && (e = cudaMalloc(reinterpret_cast<void **>(&s->d_slices), s->n * static_cast<size_t>(s->slice_size) * 4)) != cudaSuccess)
        return cuda_fail(e, "cudaMalloc");
    if (!s->d_probs && (e = cudaMalloc(reinterpret_cast<void **>(&s->d_probs), s->n * L * 4)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMemcpyAsync(s->d_slices, slices, bytes, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess) return cuda_fail(e, "H2D slices");
    int rc = streams_push_device(s, s->d_slices, f32, beyond, s->d_probs, has_result, h->stream);
    if (rc) return rc;
    if (*has_result && (e = cudaMemcpyAsync(probs, s->d_probs, s->n * L * 4, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess)
        return cuda_fail(e, "D2H probs");
    if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return cuda_fail(e, "kernel execution");
    return EIKWS_OK;
}

int eikws_streams_push_i16_host(eikws_streams *s, const int16_t *slices, float beyond, float *probs, int *has_result) {
    return streams_push_host(s, slices, false, beyond, probs, has_result);
}
int eikws_streams_push_f32_host(eikws_streams *s, const float *slices, float beyond, float *probs, int *has_result) {
    return streams_push_host(s, slices, true, beyond, probs, has_result);
}

// stage taps of the fused kernel for parity debugging (tests only): per clip P[129][49] (transposed power spectra),
// log-mel [49][33], pre-CMVN cepstra [49][13]; returns the record length through *floats_per_clip when taps == NULL
int eikws_debug_stage_taps_i16_host(eikws_handle *h, const int16_t *pcm, size_t n, float *taps, int *floats_per_clip) {
    if (floats_per_clip) *floats_per_clip = debug_tap_floats();
    if (!taps) return EIKWS_OK;
    if (!h || !pcm) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    int rc;
    cudaError_t e;
    const size_t L = h->graph.labels.size(), rec = static_cast<size_t>(debug_tap_floats());
    if ((rc = ensure(&h->d_in, &h->d_in_bytes, n * kSamples * 2))) return rc;
    if ((rc = ensure(reinterpret_cast<void **>(&h->d_probs), &h->d_probs_bytes, n * L * 4))) return rc;
    float *d_taps = nullptr;
    if ((e = cudaMalloc(reinterpret_cast<void **>(&d_taps), n * rec * 4)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(taps)");
    if ((e = cudaMemcpyAsync(h->d_in, pcm, n * kSamples * 2, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess) return cuda_fail(e, "H2D");
    rc = launch(h, h->d_in, false, nullptr, n, true, h->d_probs, nullptr, nullptr, h->stream, d_taps);
    if (!rc && (e = cudaMemcpyAsync(taps, d_taps, n * rec * 4, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess) rc = cuda_fail(e, "D2H");
    if (!rc && (e = cudaStreamSynchronize(h->stream)) != cudaSuccess) rc = cuda_fail(e, "kernel execution");
    cudaFree(d_taps);
    return rc;
}

// tests only: CMVN + input quantisation of caller-supplied pre-CMVN cepstra [n][49][13] (host) -> int8 features [n][637] (host);
// shortcut != 0 runs the certified path of the default classify kernel (cmvn_certified / cmvn_resolve), 0 every chain exactly
int eikws_debug_cmvn_quantise_host(eikws_handle *h, const float *cepstra, size_t n, int shortcut, int8_t *q) {
    if (!h || !cepstra || !q) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    if (!h->host.dev.mfcc.input_is_int8) return fail(EIKWS_ERR_UNSUPPORTED, "the model has no quantised input tensor");
    if (n == 0) return EIKWS_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    int rc;
    cudaError_t e;
    if ((rc = ensure(reinterpret_cast<void **>(&h->d_feat), &h->d_feat_bytes, n * kFeatures * 4))) return rc;
    if ((rc = ensure(reinterpret_cast<void **>(&h->d_qfeat), &h->d_qfeat_bytes, n * kFeatures))) return rc;
    if ((e = cudaMemcpyAsync(h->d_feat, cepstra, n * kFeatures * 4, cudaMemcpyHostToDevice, h->stream)) != cudaSuccess) return cuda_fail(e, "H2D");
    if ((e = launch_debug_cmvn_quantise(h->dev.d_plan, h->d_feat, n, shortcut, h->d_qfeat, h->stream)) != cudaSuccess) return cuda_fail(e, "kernel launch");
    h->launches++;
    if ((e = cudaMemcpyAsync(q, h->d_qfeat, n * kFeatures, cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess) return cuda_fail(e, "D2H");
    if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return cuda_fail(e, "kernel execution");
    return EIKWS_OK;
}

// ---- several GPUs of one box behind one call ----------------------------------------------------------------------------
// The application's loop (nucleo-l476-keyword-spotting/Core/Src/main.cpp:190-194) classifies one window after the other on one
// core; a batch host shards its clips over the box instead.  Clips are independent, so device d of D gets the contiguous range
// [d*n/D, (d+1)*n/D) (sizes differ by at most one clip), one host thread and one pair of streams per device, no exchange
// between devices; results land in place.
struct eikws_multi {
    std::vector<eikws_handle *> hs;
};

static void shard_range(size_t n, size_t parts, size_t i, size_t *lo, size_t *hi) {
    const size_t base = n / parts, rem = n % parts;
    *lo = i * base + (i < rem ? i : rem);
    *hi = *lo + base + (i < rem ? 1 : 0);
}

void eikws_multi_destroy(eikws_multi *m) {
    if (!m) return;
    for (eikws_handle *h : m->hs) eikws_destroy(h);
    delete m;
}

int eikws_multi_create(const void *model_blob, size_t bytes, const int *devices, int n_devices, eikws_multi **out) {
    if (!out) return fail(EIKWS_ERR_BAD_ARG, "null output argument");
    *out = nullptr;
    std::vector<int> devs;
    if (devices) {
        if (n_devices < 1) return fail(EIKWS_ERR_BAD_ARG, "empty device list");
        devs.assign(devices, devices + n_devices);
    } else {  // every visible device (n_devices > 0 caps the count)
        int nd = 0;
        cudaError_t e = cudaGetDeviceCount(&nd);
        if (e != cudaSuccess || nd <= 0) return fail(EIKWS_ERR_CUDA, std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
        if (n_devices > 0 && n_devices < nd) nd = n_devices;
        for (int d = 0; d < nd; d++) devs.push_back(d);
    }
    for (size_t i = 0; i < devs.size(); i++)
        for (size_t j = 0; j < i; j++)
            if (devs[i] == devs[j]) return fail(EIKWS_ERR_BAD_ARG, "device listed twice");
    eikws_multi *m = new (std::nothrow) eikws_multi();
    if (!m) return fail(EIKWS_ERR_ALLOC_FAILED, "out of memory");
    for (int d : devs) {
        eikws_handle *h = nullptr;
        int rc = eikws_create(model_blob, bytes, d, &h);
        if (rc != EIKWS_OK) {
            eikws_multi_destroy(m);
            return rc;  // eikws_create has set the message
        }
        m->hs.push_back(h);
    }
    *out = m;
    return EIKWS_OK;
}

int eikws_multi_device_count(const eikws_multi *m) { return m ? static_cast<int>(m->hs.size()) : 0; }
eikws_handle *eikws_multi_handle(eikws_multi *m, int i) { return (m && i >= 0 && i < static_cast<int>(m->hs.size())) ? m->hs[i] : nullptr; }
void eikws_multi_shard(const eikws_multi *m, size_t n_clips, int i, size_t *first, size_t *count) {
    size_t lo = 0, hi = 0;
    if (m && i >= 0 && i < static_cast<int>(m->hs.size())) shard_range(n_clips, m->hs.size(), static_cast<size_t>(i), &lo, &hi);
    if (first) *first = lo;
    if (count) *count = hi - lo;
}

}  // extern "C"
template <class Fn>
static int multi_run(eikws_multi *m, size_t n, Fn per_device) {
    const size_t D = m->hs.size();
    std::vector<int> rcs(D, EIKWS_OK);
    std::vector<std::string> errs(D);
    std::vector<std::thread> th;
    th.reserve(D);
    for (size_t d = 0; d < D; d++) {
        th.emplace_back([&, d]() {
            size_t lo, hi;
            shard_range(n, D, d, &lo, &hi);
            if (hi == lo) return;
            rcs[d] = per_device(m->hs[d], lo, hi - lo);
            if (rcs[d] != EIKWS_OK) errs[d] = t_err;  // the message lives in the worker's thread-local slot
        });
    }
    for (std::thread &t : th) t.join();
    for (size_t d = 0; d < D; d++)
        if (rcs[d] != EIKWS_OK) return fail(rcs[d], "device " + std::to_string(m->hs[d]->device) + ": " + errs[d]);
    return EIKWS_OK;
}
extern "C" {

int eikws_multi_classify_i16_host(eikws_multi *m, const int16_t *pcm, size_t n, float *probs) {
    if (!m || m->hs.empty() || !pcm || !probs) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    const size_t L = m->hs[0]->graph.labels.size();
    return multi_run(m, n, [&](eikws_handle *h, size_t first, size_t count) {
        return eikws_classify_i16_host(h, pcm + first * kSamples, count, probs + first * L);
    });
}
int eikws_multi_classify_f32_host(eikws_multi *m, const float *samples, size_t n, float *probs) {
    if (!m || m->hs.empty() || !samples || !probs) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    const size_t L = m->hs[0]->graph.labels.size();
    return multi_run(m, n, [&](eikws_handle *h, size_t first, size_t count) {
        return eikws_classify_f32_host(h, samples + first * kSamples, count, probs + first * L);
    });
}
// Device-resident shards: d_pcm[i] / d_probs[i] live on device i of the set and hold n_clips[i] clips; every launch is
// asynchronous on streams[i] (NULL array or entry = that device's default stream).  One host thread issues all of them.
int eikws_multi_classify_i16_device(eikws_multi *m, const int16_t *const *d_pcm, const size_t *n_clips, float *const *d_probs, void *const *streams) {
    if (!m || m->hs.empty() || !d_pcm || !n_clips || !d_probs) return fail(EIKWS_ERR_BAD_ARG, "null argument");
    for (size_t d = 0; d < m->hs.size(); d++) {
        if (n_clips[d] == 0) continue;
        int rc = eikws_classify_i16_device(m->hs[d], d_pcm[d], n_clips[d], d_probs[d], streams ? streams[d] : nullptr);
        if (rc != EIKWS_OK) return rc;
    }
    return EIKWS_OK;
}

// Page-locked host memory for a pure-C host (no CUDA headers needed): buffers handed to the *_host entry points are copied at
// PCIe speed and asynchronously only when they are pinned; cudaHostAllocPortable makes them so for every device of the box.
void *eikws_host_alloc(size_t bytes) {
    void *p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaHostAlloc");
        return nullptr;
    }
    return p;
}
void eikws_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

// ---- parity taps of host-side derived data (tests only; no GPU needed) ------------------------------------------
int eikws_debug_host_plan(const void *model_blob, size_t bytes, float *filterbank_129x32, int32_t *conv_mult, int32_t *conv_shift,
                          int max_channels, int *n_channels) {
    ModelGraph g;
    HostPlan hp;
    std::string err;
    if (!parse_model(model_blob, bytes, g, err)) return fail(EIKWS_ERR_BAD_ARG, err);
    int rc = build_host_plan(g, hp, err);
    if (rc) return fail(rc, err);
    if (filterbank_129x32) std::memcpy(filterbank_129x32, hp.filterbank.data(), hp.filterbank.size() * sizeof(float));
    int n = 0;
    for (int o = 0; o < hp.dev.nn.n_ops; o++) {
        const NnOpDev &op = hp.dev.nn.ops[o];
        if (op.kind != kNnConv1d) continue;
        // pointers are still null on the host image; find this op's tables through the fixups
        size_t moff = 0, soff = 0;
        for (const auto &f : hp.fixups) {
            const uint8_t *base = reinterpret_cast<const uint8_t *>(&hp.dev);
            if (base + f.first == reinterpret_cast<const uint8_t *>(&op.mult)) moff = f.second;
            if (base + f.first == reinterpret_cast<const uint8_t *>(&op.shift)) soff = f.second;
        }
        for (int c = 0; c < op.out_c && n < max_channels; c++, n++) {
            if (conv_mult) std::memcpy(&conv_mult[n], hp.blob.data() + moff + 4 * c, 4);
            if (conv_shift) std::memcpy(&conv_shift[n], hp.blob.data() + soff + 4 * c, 4);
        }
    }
    if (n_channels) *n_channels = n;
    return EIKWS_OK;
}

}  // extern "C"
