/* eikws_b200.h -- C ABI of libeikws_b200.so
 *
 * B200-native (sm_100a CUDA) implementation of the hot path of ShawnHymel/ei-keyword-spotting:
 *     run_classifier(signal_t*, ei_impulse_result_t*, bool)
 *         edge-impulse-sdk/classifier/ei_run_classifier.h:650-714
 *   = extract_mfcc_features   edge-impulse-sdk/classifier/ei_run_dsp.h:256-308
 *   + run_inference           edge-impulse-sdk/classifier/ei_run_classifier.h:293-641
 * Plain pointers and sizes only; no C++/torch types.  Error returns use the reference's
 * EI_IMPULSE_ERROR values (edge-impulse-sdk/porting/ei_classifier_porting.h:34-43) plus the
 * EIKWS_* extensions below.  There is no CPU compute fallback: every classify/features call
 * runs CUDA kernels or fails.
 *
 * The reference processes one clip per call through a pull callback; one callback per clip
 * cannot feed a GPU, so the batch entry points take contiguous clips
 * ([n_clips][raw_sample_count], int16 PCM or float) -- same arithmetic per clip.
 */
#ifndef EIKWS_B200_H
#define EIKWS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* EI_IMPULSE_ERROR values (ei_classifier_porting.h:34-43) */
#define EIKWS_OK 0
#define EIKWS_ERR_SHAPES_DONT_MATCH (-1)
#define EIKWS_ERR_CANCELED (-2)
#define EIKWS_ERR_TFLITE (-3)
#define EIKWS_ERR_DSP (-5)
#define EIKWS_ERR_TFLITE_ARENA_ALLOC_FAILED (-6)
#define EIKWS_ERR_ALLOC_FAILED (-8)
/* extensions */
#define EIKWS_ERR_UNSUPPORTED (-100) /* model/DSP configuration outside what the kernels implement */
#define EIKWS_ERR_CUDA (-101)        /* CUDA runtime failure (no device, launch error, ...)          */
#define EIKWS_ERR_BAD_ARG (-102)

typedef struct eikws_handle eikws_handle;

/* What a generated Edge Impulse export provides; filled by the drop-in header
 * include/edge-impulse-sdk/classifier/ei_run_classifier.h from the UNMODIFIED
 * model-parameters/model_metadata.h + tflite-model/trained_model_compiled.{h,cpp}. */
typedef struct {
    int (*init)(void *(*alloc_fnc)(size_t, size_t)); /* trained_model_init   (trained_model_compiled.cpp:380) */
    void *(*input)(int index);                       /* trained_model_input  (:446) -> TfLiteTensor*          */
    void *(*output)(int index);                      /* trained_model_output (:453) -> TfLiteTensor*          */
    int (*reset)(void (*free_fnc)(void *));          /* trained_model_reset  (:467)                           */
    uint32_t raw_sample_count;                       /* EI_CLASSIFIER_RAW_SAMPLE_COUNT                        */
    uint32_t nn_input_frame_size;                    /* EI_CLASSIFIER_NN_INPUT_FRAME_SIZE                     */
    uint32_t label_count;                            /* EI_CLASSIFIER_LABEL_COUNT                             */
    int32_t frequency;                               /* EI_CLASSIFIER_FREQUENCY                               */
    const char *const *labels;                       /* ei_classifier_inferencing_categories                  */
    /* ei_dsp_config_mfcc_t (model_metadata.h:92-104) */
    int32_t mfcc_num_cepstral;
    float mfcc_frame_length;
    float mfcc_frame_stride;
    int32_t mfcc_num_filters;
    int32_t mfcc_fft_length;
    int32_t mfcc_win_size;
    int32_t mfcc_low_frequency;
    int32_t mfcc_high_frequency;
    float mfcc_pre_cof;
    int32_t mfcc_pre_shift;
} eikws_compiled_model_t;

/* ---- model ingestion (host only, no GPU needed) ------------------------------------------- */
/* Runs the generated model's init against this library's Register_* operators
 * (replaces TFL/micro/kernels/{conv,add,pooling,fully_connected,softmax,reshape}.cc registrations,
 * micro_ops.h:33-75) and serialises the recorded graph ("EIKWSMDL" container). *blob is malloc'ed;
 * release with eikws_free(). */
int eikws_model_from_compiled(const eikws_compiled_model_t *cm, void **blob, size_t *bytes);
void eikws_free(void *p);

/* ---- lifecycle ----------------------------------------------------------------------------- */
/* Lowers the model to a device plan on CUDA device `device`.  Fails with EIKWS_ERR_CUDA if no
 * usable GPU, EIKWS_ERR_UNSUPPORTED if the graph/DSP config is outside the implemented family. */
int eikws_create(const void *model_blob, size_t bytes, int device, eikws_handle **out);
void eikws_destroy(eikws_handle *h);

int eikws_label_count(const eikws_handle *h);
int eikws_feature_count(const eikws_handle *h);    /* EI_CLASSIFIER_NN_INPUT_FRAME_SIZE */
int eikws_raw_sample_count(const eikws_handle *h); /* EI_CLASSIFIER_RAW_SAMPLE_COUNT    */
int eikws_device(const eikws_handle *h);
const char *eikws_label(const eikws_handle *h, int i);

/* ---- batch hot path, DEVICE buffers (on the handle's device) ------------------------------- */
/* run_classifier over n_clips clips of int16 PCM (the demos' signal: int16 -> x/32768,
 * numpy.hpp:1289-1298).  d_probs: [n_clips][label_count] float = result.classification[i].value.
 * stream: a cudaStream_t passed as void* (NULL = default stream).  Asynchronous.
 * Exactness: for an int8 model the outputs are byte-identical to the reference CPU code built with
 * -ffp-contract=off; the classify kernels may decide the int8 rounding of a CMVN output from double-precision window
 * statistics and a proven error bound instead of running the reference's float chain, and run that chain whenever the
 * bound cannot decide.  Float features (eikws_features_*) always come from the reference's exact
 * operation sequence and are bit-identical. */
int eikws_classify_i16_device(eikws_handle *h, const int16_t *d_pcm, size_t n_clips, float *d_probs, void *stream);
/* same, float samples as a signal_t callback would deliver them */
int eikws_classify_f32_device(eikws_handle *h, const float *d_samples, size_t n_clips, float *d_probs, void *stream);
/* extract_mfcc_features only: d_features [n_clips][feature_count] float (may be NULL),
 * d_qfeatures [n_clips][feature_count] int8 = the quantised NN input (ei_run_classifier.h:436-444; may be NULL) */
int eikws_features_i16_device(eikws_handle *h, const int16_t *d_pcm, size_t n_clips, float *d_features,
                              int8_t *d_qfeatures, void *stream);
int eikws_features_f32_device(eikws_handle *h, const float *d_samples, size_t n_clips, float *d_features,
                              int8_t *d_qfeatures, void *stream);
/* run_inference only (ei_run_classifier.h:293): float features in, probabilities out */
int eikws_infer_device(eikws_handle *h, const float *d_features, size_t n_clips, float *d_probs, void *stream);

/* ---- batch hot path, HOST buffers (H2D + kernels + D2H inside the call, synchronous) ------- */
int eikws_classify_i16_host(eikws_handle *h, const int16_t *pcm, size_t n_clips, float *probs);
int eikws_classify_f32_host(eikws_handle *h, const float *samples, size_t n_clips, float *probs);
int eikws_features_i16_host(eikws_handle *h, const int16_t *pcm, size_t n_clips, float *features, int8_t *qfeatures);
int eikws_features_f32_host(eikws_handle *h, const float *samples, size_t n_clips, float *features, int8_t *qfeatures);
int eikws_infer_host(eikws_handle *h, const float *features, size_t n_clips, float *probs);

/* Thread safety: the *_host entry points and the single-clip calls below serialise on the handle (they share its staging
 * buffers; the lock is held from the first byte staged to the last result copied); the *_device entry points only launch and
 * may be called concurrently on different streams.  Host buffers are copied at full PCIe speed, asynchronously, only when
 * they are page-locked: eikws_host_alloc / eikws_host_free give a pure-C host such memory without CUDA headers. */
void *eikws_host_alloc(size_t bytes); /* NULL on failure (eikws_last_error) */
void eikws_host_free(void *p);

/* ---- several GPUs of one box behind one call ----------------------------------------------------------------
 * The reference application classifies window after window on one core (nucleo-l476-keyword-spotting/Core/Src/main.cpp:190-194);
 * a batch host shards its clips over the box instead.  Clips are independent: device i of D gets the contiguous range
 * eikws_multi_shard(m, n, i, &first, &count) (sizes differ by at most one clip), one host thread and one stream pair per
 * device, no exchange between devices; results land in place and are byte-identical to a single-device run.
 * devices == NULL selects the first n_devices visible devices (n_devices <= 0: all of them). */
typedef struct eikws_multi eikws_multi;
int eikws_multi_create(const void *model_blob, size_t bytes, const int *devices, int n_devices, eikws_multi **out);
void eikws_multi_destroy(eikws_multi *m);
int eikws_multi_device_count(const eikws_multi *m);
eikws_handle *eikws_multi_handle(eikws_multi *m, int i); /* the i-th device's handle (owned by m) */
void eikws_multi_shard(const eikws_multi *m, size_t n_clips, int i, size_t *first, size_t *count);
int eikws_multi_classify_i16_host(eikws_multi *m, const int16_t *pcm, size_t n_clips, float *probs);
int eikws_multi_classify_f32_host(eikws_multi *m, const float *samples, size_t n_clips, float *probs);
/* device-resident shards: d_pcm[i] / d_probs[i] live on device i of the set and hold n_clips[i] clips; asynchronous on
 * streams[i] (streams == NULL or streams[i] == NULL: that device's default stream) */
int eikws_multi_classify_i16_device(eikws_multi *m, const int16_t *const *d_pcm, const size_t *n_clips, float *const *d_probs, void *const *streams);

/* ---- single clip through the reference's pull callback (signal_t::get_data with
 * EIDSP_SIGNAL_C_FN_POINTER=1, numpy_types.h:242-249).  Used by the drop-in run_classifier(). --- */
typedef int (*eikws_get_data_fn)(size_t offset, size_t length, float *out_ptr);
int eikws_run_classifier_signal(eikws_handle *h, eikws_get_data_fn get_data, size_t total_length, float *values,
                                int *timing_dsp_ms, int *timing_classification_ms);

/* extract_mfcc_features (ei_run_dsp.h:256-308) for one clip pulled through the same callback: features[feature_count] */
int eikws_extract_mfcc_signal(eikws_handle *h, eikws_get_data_fn get_data, size_t total_length, float *features);

/* ---- the sibling MFE DSP block: extract_mfe_features of the reference's newer SDK copy
 * (nucleo-l432-keyword-spotting/keyword-spotting-02-v3/edge-impulse-sdk/classifier/ei_run_dsp.h:369-418):
 * mel filterbank energies of the raw frames (no pre-emphasis, no log), sliding-window mean subtraction, min/max scaling of
 * the whole matrix.  The block uses the geometry of the impulse's MFCC block; features are [n][eikws_mfe_feature_count()]
 * = [n][frames * num_filters], row-major [frame][filter] like the reference's output matrix. --- */
int eikws_mfe_feature_count(const eikws_handle *h);
int eikws_mfe_i16_device(eikws_handle *h, const int16_t *d_pcm, size_t n_clips, float *d_features, void *stream);
int eikws_mfe_f32_device(eikws_handle *h, const float *d_samples, size_t n_clips, float *d_features, void *stream);
int eikws_mfe_i16_host(eikws_handle *h, const int16_t *pcm, size_t n_clips, float *features);
int eikws_mfe_f32_host(eikws_handle *h, const float *samples, size_t n_clips, float *features);
/* mirrors ei_dsp_config_mfe_t (same directory, model-parameters/model_metadata.h:103-112) */
typedef struct {
    int axes;
    float frame_length;
    float frame_stride;
    int num_filters;
    int fft_length;
    int low_frequency;
    int high_frequency;
    int win_size;
} eikws_mfe_config;
/* one clip pulled through the signal callback; EIKWS_ERR_UNSUPPORTED when cfg is not the MFCC block's geometry */
int eikws_extract_mfe_signal(eikws_handle *h, eikws_get_data_fn get_data, size_t total_length, const eikws_mfe_config *cfg,
                             float *features, size_t capacity);

/* ---- continuous mode: run_classifier_continuous (ei_run_classifier.h:184-282) over many streams --------------
 * Every call feeds ONE slice (raw_sample_count / slices_per_window samples) of every stream; all streams advance in
 * lock step.  Per stream the semantics are those of one reference process from power-up: per-slice MFCC without CMVN
 * (extract_mfcc_per_slice_features, ei_run_dsp.h:310-366), a 637-feature window, CMVN + classifier over the whole
 * window once it is full, a moving average over slices_per_window/2 results (run_moving_average_filter :134-145).
 * `beyond`: from the second slice on the reference asks the signal callback for ONE sample past the end of the slice
 * (index slice_size + 319); pass the float that callback returns there (a zero-padded buffer: 0). */
typedef struct eikws_streams eikws_streams;
int eikws_streams_create(eikws_handle *h, size_t n_streams, int slices_per_window, eikws_streams **out);
void eikws_streams_destroy(eikws_streams *s);
int eikws_streams_reset(eikws_streams *s); /* run_classifier_init (:164-172) + power-up */
int eikws_streams_slice_size(const eikws_streams *s);
/* d_slices [n_streams][slice_size] int16 (device, 16-byte aligned); d_probs [n_streams][label_count] is written and
 * *has_result set to 1 only once the window is full */
int eikws_streams_push_i16_device(eikws_streams *s, const int16_t *d_slices, float beyond, float *d_probs, int *has_result, void *stream);
int eikws_streams_push_i16_host(eikws_streams *s, const int16_t *slices, float beyond, float *probs, int *has_result);
/* same with float samples as the signal callback delivers them */
int eikws_streams_push_f32_device(eikws_streams *s, const float *d_slices, float beyond, float *d_probs, int *has_result, void *stream);
int eikws_streams_push_f32_host(eikws_streams *s, const float *slices, float beyond, float *probs, int *has_result);

/* ---- caller-side ingest ----------------------------------------------------------------------------------------
 * The firmware's microphone path (nucleo-l476-keyword-spotting/Core/Src/main.cpp:507-521): SAI words (32 kHz stereo,
 * 24 valid bits in 32) -> 16 kHz mono int16: d_pcm[i] = (int16_t)(d_i2s[skip * i] >> shift); firmware: skip 4, shift 8.
 * d_pcm must be 16-byte aligned; d_i2s holds at least skip * n_out words. */
int eikws_decimate_i2s_device(eikws_handle *h, const int32_t *d_i2s, size_t n_out, int skip, int shift, int16_t *d_pcm, void *stream);

/* The arithmetic of the dataset tooling's mix_audio (dataset-curation.py:93-137) and of its PCM_16 write (:190-206) for material that
 * is already 16 kHz float32 (the resampling librosa.load does there is NOT part of this call):
 *   d_pcm[c][i] = PCM16(0.5 * word_vol * word[c][i] + 0.5 * bg_vol * bg[bg_start[c] + i]),  i < raw_sample_count,
 * word c = d_words + c * word_stride with d_word_len[c] valid samples (shorter: zero-padded, longer: truncated; d_words == NULL:
 * the script's background-only clips), bg_start[c] <= max_bg_start <= bg_len - raw_sample_count.  PCM16(x) = lrint(x * 32767),
 * low 16 bits (libsndfile's unclipped double -> short conversion).  The volumes are doubles because the script's are Python floats.  One streaming kernel: 8 bytes read, 2 written per sample. */
int eikws_mix_audio_device(eikws_handle *h, const float *d_words, const uint32_t *d_word_len, size_t word_stride, const float *d_bg,
                           size_t bg_len, const uint32_t *d_bg_start, size_t max_bg_start, double word_vol, double bg_vol, size_t n_clips,
                           int16_t *d_pcm, void *stream);

/* ---- diagnostics --------------------------------------------------------------------------- */
const char *eikws_last_error(void); /* thread-local text of the last failure */
/* number of kernel launches issued by this handle since creation (bench.py's gpu_launches claim) */
uint64_t eikws_launch_count(const eikws_handle *h);
/* deterministic synthetic int16 clips written on the device (bench/test input generator);
 * clip c of the stream is generated for index first_clip + c. */
int eikws_synth_i16_device(eikws_handle *h, int16_t *d_pcm, size_t n_clips, uint64_t first_clip, uint64_t seed, void *stream);


/* ---- tuning knobs (A/B measurement of kernel variants; defaults are the measured best, results never change) ---- */
int eikws_set_ctas_per_sm(eikws_handle *h, int n);   /* clip groups resident per SM (1..8)                              */
int eikws_set_clips_per_cta(eikws_handle *h, int n); /* clip groups per CTA: 1, 2 or 4                                   */
int eikws_set_tensor_core(eikws_handle *h, int on);  /* block 1 of the fused classifier as a tcgen05 UMMA                */
int eikws_set_cmvn_shortcut(eikws_handle *h, int on);/* certified CMVN shortcut (0: every chain with the exact sequence) */
int eikws_set_work_claiming(eikws_handle *h, int on);/* work-claiming schedule of the shortcut kernel                    */
int eikws_set_pipelined(eikws_handle *h, int on);     /* software-pipelined classify kernel (two clips in different stages per CTA) */
int eikws_set_split(eikws_handle *h, int on);         /* two-kernel classify path (default on; int16 or float32 clips, tensor-core int8 lowering or float32 graph): spectral kernel, then cepstral / classifier kernel; 0 = the single fused kernel */
int eikws_set_kernel_timing(eikws_handle *h, int on); /* measurement aid: record CUDA events around the two kernels of every split launch (device entry points; not for concurrent callers) */
int eikws_split_kernel_ms(eikws_handle *h, float *ms2, uint64_t *launches); /* waits for the timed launches: average ms per launch of ms2[0] spectral kernel, ms2[1] cepstral / classifier kernel; resets the sums */
int eikws_set_skew_ns(eikws_handle *h, int ns);      /* start offset between the CTAs that share an SM                   */

/* ---- parity taps (tests only) --------------------------------------------------------------- */
/* classify and also return the float features and the int8 classifier input */
int eikws_classify_taps_i16_device(eikws_handle *h, const int16_t *d_pcm, size_t n_clips, float *d_probs, float *d_features,
                                   int8_t *d_qfeatures, void *stream);
int eikws_classify_taps_i16_host(eikws_handle *h, const int16_t *pcm, size_t n_clips, float *probs, float *features, int8_t *qfeatures);
/* per clip: P[129][49] power spectra, log-mel [49][33], pre-CMVN cepstra [49][13]; taps == NULL only reports the record length */
int eikws_debug_stage_taps_i16_host(eikws_handle *h, const int16_t *pcm, size_t n_clips, float *taps, int *floats_per_clip);
/* the CMVN + input-quantisation stage alone on caller-supplied pre-CMVN cepstra [n][49][13] -> int8 features [n][637];
 * shortcut != 0: the certified path of the default classify kernel, 0: every chain with the reference's sequence */
int eikws_debug_cmvn_quantise_host(eikws_handle *h, const float *cepstra, size_t n, int shortcut, int8_t *qfeatures);
/* host-side derived tables, no GPU needed: dense mel filterbank [129][32], conv/FC requantisation multipliers and shifts */
int eikws_debug_host_plan(const void *model_blob, size_t bytes, float *filterbank_129x32, int32_t *conv_mult, int32_t *conv_shift,
                          int max_channels, int *n_channels);

#ifdef __cplusplus
}
#endif
#endif /* EIKWS_B200_H */
