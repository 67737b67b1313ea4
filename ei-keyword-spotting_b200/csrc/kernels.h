// eikws-b200: host-callable launchers of kernels.cu
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "dev_plan.h"

namespace eikws {

struct LaunchArgs {
    const DevPlan *plan = nullptr;       // device pointer
    const void *clips = nullptr;         // device: [n_clips][16000] int16 or float
    bool input_is_f32 = false;
    const float *features_in = nullptr;  // device: run_inference only (clips ignored)
    size_t n_clips = 0;
    bool run_nn = true;
    bool nn_fused = false;               // the plan has a fused classifier (NnFusedDev.enabled)
    bool nn_float = false;               // float32 graph (NnDev.float_mode)
    bool nn_tc = false;                  // block 1 of the fused classifier on the tensor core (NnFusedDev.tc_enabled, 2 clip groups per CTA)
    bool cmvn_certified = false;         // certified CMVN shortcut (tensor-core variant, no float feature output): see cmvn_certified()
    bool pipelined = false;              // with cmvn_certified + nn_tc: the software-pipelined kernel (eikws_pipelined_kernel: FFT of clip s interleaved with the post-FFT slices of clip s-1)
    bool split = false;                  // with cmvn_certified + nn_tc, int16 clips: the two-kernel path (eikws_logmel_kernel -> eikws_cepstral_kernel); needs logmel
    float *logmel = nullptr;             // device scratch of split_scratch_bytes(n_clips): the log-mel records handed from the first kernel to the second
    cudaEvent_t *split_events = nullptr; // optional, 3 events: recorded before the first kernel, between the two, after the second
    bool work_claiming = false;          // with cmvn_certified: frame pairs and the UMMA issue are claimed dynamically (kDyn in kernels.cu)
    float *probs = nullptr;              // device: [n_clips][labels]
    float *features_out = nullptr;       // device, optional: [n_clips][637]
    int8_t *qfeatures_out = nullptr;     // device, optional: [n_clips][637]
    float *debug_taps = nullptr;         // device, tests only: [n_clips][debug_tap_floats()] (int16 classify path)
    int grid = 0;                        // CTAs if each holds one clip group (launch divides by clips_per_cta)
    int clips_per_cta = 1;               // 160-thread clip groups per CTA (1, 2 or 4): int16 + fused classifier path only
    int sm_count = 0;                    // SMs of the device (CTA j of an SM = blockIdx / sm_count)
    int skew_ns = 0;                     // start offset between co-resident CTAs (0 = none)
    float pre_cof = 0.0f;                // MfccDev::pre_cof again, as a kernel argument (a constant-bank operand of the pre-emphasis multiply)
    int nn_smem_bytes = 0;               // activation arena + conv row scratch
    cudaStream_t stream = nullptr;
};

cudaError_t launch_run_classifier(const LaunchArgs &a);
size_t split_scratch_bytes(size_t n_clips);  // LaunchArgs::logmel for a split launch of n_clips (at most 131,072 clips per launch)

// the sibling MFE DSP block: [n_clips][49 * 32] features
struct MfeArgs {
    const DevPlan *plan = nullptr;
    const void *clips = nullptr;  // device: [n_clips][16000] int16 or float32
    bool input_is_f32 = false;
    size_t n_clips = 0;
    float *out = nullptr;         // device: [n_clips][49 * 32]
    int grid = 0;
    cudaStream_t stream = nullptr;
};
cudaError_t launch_mfe(const MfeArgs &a);

// one slice for every stream (run_classifier_continuous); requires the fused int8 classifier plan
struct ContinuousArgs {
    const DevPlan *plan = nullptr;
    const void *slices = nullptr;     // device: [n_streams][slice_size] int16 or float
    bool input_is_f32 = false;
    int slice_size = 0, n_frames = 0, total_length = 0;
    float beyond = 0.0f;
    size_t n_streams = 0;
    float *state_features = nullptr;  // device: [n_streams][637]
    float *maf_buf = nullptr;         // device: [n_streams][labels][maf_len]
    float *maf_sum = nullptr;         // device: [n_streams][labels]
    int slice_offset = 0, window_full = 0, maf_idx = 0, maf_len = 1;
    bool cmvn_certified = false;      // certified CMVN shortcut for the window classification (see cmvn_certified in kernels.cu)
    float *probs = nullptr;           // device: [n_streams][labels]
    int grid = 0, sm_count = 0;
    cudaStream_t stream = nullptr;
};
cudaError_t launch_continuous(const ContinuousArgs &a);
cudaError_t launch_decimate_i2s(const int32_t *i2s, size_t n_out, int skip, int shift, int16_t *pcm, cudaStream_t st);
// tests only: CMVN + input quantisation of caller-supplied pre-CMVN cepstra [n][49][13] -> int8 [n][637]; shortcut = certified path
cudaError_t launch_debug_cmvn_quantise(const DevPlan *plan, const float *cepstra, size_t n, int shortcut, int8_t *q_out, cudaStream_t st);
// mix_audio arithmetic (dataset-curation.py:93-137) + PCM_16 conversion; words may be null (background-only clips)
cudaError_t launch_mix_audio(const float *words, const uint32_t *word_len, size_t word_stride, const float *bg, const uint32_t *bg_start,
                             double half_word_vol, float half_bg_vol, size_t n_clips, int16_t *out, cudaStream_t st);
cudaError_t launch_synth(int16_t *pcm, size_t n_clips, uint64_t first_clip, uint64_t seed, cudaStream_t st);
int kernel_threads();
int debug_tap_floats();  // P[129][49] + logmel[49][33] + cepstra[49][13]
// bytes of shared memory available to the classifier arena inside the fused kernel's overlay
int nn_smem_capacity(bool input_is_f32);
// float32 graphs may also use the (dead after CMVN) GT area: up to the mbarrier
int nn_smem_capacity_float_graph();

}  // namespace eikws
