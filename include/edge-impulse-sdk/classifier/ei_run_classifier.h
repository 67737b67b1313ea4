/* eikws-b200 drop-in for edge-impulse-sdk/classifier/ei_run_classifier.h
 *
 * Same include path, same entry points, same types as the reference header
 *   run_classifier(signal_t*, ei_impulse_result_t*, bool debug = false)      reference :650-714
 *   run_inference(ei::matrix_t*, ei_impulse_result_t*, bool debug = false)   reference :293-641
 *   run_classifier_init()                                                    reference :164-172
 * but the work is done by libeikws_b200.so on a B200: the MFCC block and the int8 classifier run as one fused
 * sm_100a kernel.  The application keeps its UNMODIFIED generated files
 *   model-parameters/model_metadata.h, model-parameters/dsp_blocks.h, tflite-model/trained_model_compiled.{h,cpp}
 * and replaces the `edge-impulse-sdk/` directory by this repo's include/edge-impulse-sdk.
 * On the first call the generated trained_model_init() is executed once against the library's recording
 * Register_*() operators (tensorflow/lite/micro/kernels/micro_ops.h) and the captured graph is lowered to a
 * device plan; nothing is recomputed per call (the reference re-inits the model on every call, :352/:487).
 *
 * Batch extension (one callback per clip cannot feed a GPU): run_classifier_batch_i16 / _f32 below, and the
 * C ABI in eikws_b200.h.  Environment: EIKWS_DEVICE selects the CUDA device (default 0).
 */
#ifndef _EDGE_IMPULSE_RUN_CLASSIFIER_H_
#define _EDGE_IMPULSE_RUN_CLASSIFIER_H_

#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "model-parameters/model_metadata.h"

#include "../dsp/numpy_types.h"
#include "../porting/ei_classifier_porting.h"
#include "ei_classifier_types.h"
#include "eikws_b200.h"

#if EI_CLASSIFIER_INFERENCING_ENGINE != EI_CLASSIFIER_TFLITE || EI_CLASSIFIER_COMPILED != 1
#error "eikws-b200 accelerates EON-compiled TFLite impulses (EI_CLASSIFIER_INFERENCING_ENGINE == EI_CLASSIFIER_TFLITE, EI_CLASSIFIER_COMPILED == 1)"
#endif

#ifndef __cplusplus
/* ---- C translation units ---------------------------------------------------------------------------------------------
 * The reference declares run_classifier `extern "C"` but defines it inside a C++ header (anonymous namespace, default
 * argument: reference :106-108, :650-653), and its README has users delete the SDK's C wrapper ei_run_classifier_c.*
 * (README.md:185), so a pure-C firmware-style host cannot call it as shipped.  Here a .c file includes this header (exactly
 * ONE translation unit of the program may, like the reference: model_metadata.h defines globals), gets the declarations
 * below, and the program links
 *     edge-impulse-sdk/classifier/ei_run_classifier_c.cpp   (C++ shim, defines these symbols with C linkage)
 *     tflite-model/trained_model_compiled.cpp               (the unmodified generated model)
 *     -leikws_b200
 * `debug` is explicit in C (no default arguments); signal_t::get_data is the plain function pointer flavour
 * (EIDSP_SIGNAL_C_FN_POINTER == 1, reference numpy_types.h:242-249).  See examples/static_buffer.c. */
EI_IMPULSE_ERROR run_classifier(signal_t *signal, ei_impulse_result_t *result, bool debug);
EI_IMPULSE_ERROR run_classifier_continuous(signal_t *signal, ei_impulse_result_t *result, bool debug);
void run_classifier_init(void);
/* batch extension: n_clips contiguous clips of EI_CLASSIFIER_RAW_SAMPLE_COUNT samples in host memory (pin it with
 * eikws_host_alloc for full PCIe speed); sharded over the devices given to ei_b200_init */
EI_IMPULSE_ERROR run_classifier_batch_i16(const int16_t *pcm, size_t n_clips, ei_impulse_result_t *results);
EI_IMPULSE_ERROR run_classifier_batch_f32(const float *samples, size_t n_clips, ei_impulse_result_t *results);
/* optional: choose the GPUs (devices == NULL: the first n_devices visible ones, n_devices <= 0: all of them).  Without it the
 * first call lowers the impulse on device $EIKWS_DEVICE (default 0).  ei_b200_shutdown releases every device resource. */
EI_IMPULSE_ERROR ei_b200_init(const int *devices, int n_devices);
void ei_b200_shutdown(void);
#else /* __cplusplus */

#include "ei_model_types.h"
#include "ei_run_dsp.h"
#include "model-parameters/dsp_blocks.h"
#include "tflite-model/trained_model_compiled.h"

using ei::matrix_t;
using ei::signal_t;

namespace {

static int eikws_dropin_init_thunk(void *(*a)(size_t, size_t)) { return (int)trained_model_init(a); }
static void *eikws_dropin_input_thunk(int i) { return trained_model_input(i); }
static void *eikws_dropin_output_thunk(int i) { return trained_model_output(i); }
static int eikws_dropin_reset_thunk(void (*f)(void *)) { return (int)trained_model_reset(f); }

/* Runs the generated trained_model_init() against the library's recording operators and serialises the captured impulse. */
static int eikws_dropin_capture(void **blob, size_t *bytes) {
    if (ei_dsp_blocks_size != 1) {
        ei_printf("ERR: eikws-b200 supports impulses with exactly one (MFCC) DSP block\n");
        return EIKWS_ERR_UNSUPPORTED;
    }
    const ei_dsp_config_mfcc_t *c = (const ei_dsp_config_mfcc_t *)ei_dsp_blocks[0].config;
    eikws_compiled_model_t cm;
    cm.init = eikws_dropin_init_thunk;
    cm.input = eikws_dropin_input_thunk;
    cm.output = eikws_dropin_output_thunk;
    cm.reset = eikws_dropin_reset_thunk;
    cm.raw_sample_count = EI_CLASSIFIER_RAW_SAMPLE_COUNT;
    cm.nn_input_frame_size = EI_CLASSIFIER_NN_INPUT_FRAME_SIZE;
    cm.label_count = EI_CLASSIFIER_LABEL_COUNT;
    cm.frequency = EI_CLASSIFIER_FREQUENCY;
    cm.labels = ei_classifier_inferencing_categories;
    cm.mfcc_num_cepstral = c->num_cepstral;
    cm.mfcc_frame_length = c->frame_length;
    cm.mfcc_frame_stride = c->frame_stride;
    cm.mfcc_num_filters = c->num_filters;
    cm.mfcc_fft_length = c->fft_length;
    cm.mfcc_win_size = c->win_size;
    cm.mfcc_low_frequency = c->low_frequency;
    cm.mfcc_high_frequency = c->high_frequency;
    cm.mfcc_pre_cof = c->pre_cof;
    cm.mfcc_pre_shift = c->pre_shift;
    int rc = eikws_model_from_compiled(&cm, blob, bytes);
    if (rc != EIKWS_OK) ei_printf("ERR: eikws-b200 could not capture the compiled model: %s\n", eikws_last_error());
    return rc;
}

/* process-wide state, like the reference's statics: the device set of ei_b200_init (or the single lazily created handle) */
static eikws_multi *eikws_dropin_devices = NULL;
static eikws_handle *eikws_dropin_single = NULL;
static bool eikws_dropin_tried = false;

/* Lowers the impulse on first use; returns NULL (after printing the reason) when that fails. */
static eikws_handle *eikws_dropin_handle() {
    if (eikws_dropin_devices) return eikws_multi_handle(eikws_dropin_devices, 0);
    if (eikws_dropin_tried) return eikws_dropin_single;
    eikws_dropin_tried = true;
    void *blob = NULL;
    size_t bytes = 0;
    if (eikws_dropin_capture(&blob, &bytes) != EIKWS_OK) return NULL;
    const char *dev = getenv("EIKWS_DEVICE");
    int rc = eikws_create(blob, bytes, dev ? atoi(dev) : 0, &eikws_dropin_single);
    eikws_free(blob);
    if (rc != EIKWS_OK) {
        ei_printf("ERR: eikws-b200 could not create the device plan (%d): %s\n", rc, eikws_last_error());
        eikws_dropin_single = NULL;
    }
    return eikws_dropin_single;
}

/* signal_t::get_data may be a std::function (EIDSP_SIGNAL_C_FN_POINTER == 0, the SDK default); the C ABI takes a
 * plain function pointer, so the current signal is parked here for the duration of the call. run_classifier is
 * serialised per process exactly like the reference (which keeps global state, ei_run_dsp.h:251). */
static signal_t *eikws_dropin_current_signal = NULL;
static int eikws_dropin_get_data(size_t offset, size_t length, float *out) {
    return eikws_dropin_current_signal->get_data(offset, length, out);
}

static EI_IMPULSE_ERROR eikws_dropin_error(int rc) {
    switch (rc) {
        case EIKWS_OK: return EI_IMPULSE_OK;
        case EIKWS_ERR_DSP: return EI_IMPULSE_DSP_ERROR;
        case EIKWS_ERR_SHAPES_DONT_MATCH: return EI_IMPULSE_ERROR_SHAPES_DONT_MATCH;
        case EIKWS_ERR_CANCELED: return EI_IMPULSE_CANCELED;
        case EIKWS_ERR_ALLOC_FAILED: return EI_IMPULSE_ALLOC_FAILED;
        case EIKWS_ERR_CUDA: return EI_IMPULSE_ALLOC_FAILED;  /* no device / device failure */
        default: return EI_IMPULSE_TFLITE_ERROR;
    }
}

static void eikws_dropin_fill_result(ei_impulse_result_t *result, const float *values, bool debug) {
    for (uint32_t ix = 0; ix < EI_CLASSIFIER_LABEL_COUNT; ix++) {
        result->classification[ix].label = ei_classifier_inferencing_categories[ix];
        result->classification[ix].value = values[ix];
        if (debug) {
            ei_printf("%s:\t", ei_classifier_inferencing_categories[ix]);
            ei_printf_float(values[ix]);
            ei_printf("\n");
        }
    }
    result->anomaly = 0.0f;
}

}  // namespace

static int eikws_dropin_extract_mfcc(ei::signal_t *signal, ei::matrix_t *output_matrix) {
    eikws_handle *h = eikws_dropin_handle();
    if (!h) return -1004; /* EIDSP_OUT_OF_MEM class of failure */
    if (output_matrix->rows * output_matrix->cols < (uint32_t)EI_CLASSIFIER_NN_INPUT_FRAME_SIZE) return EIDSP_MATRIX_SIZE_MISMATCH;
    eikws_dropin_current_signal = signal;
    int rc = eikws_extract_mfcc_signal(h, &eikws_dropin_get_data, signal->total_length, output_matrix->buffer);
    if (rc != EIKWS_OK) return EIDSP_MATRIX_SIZE_MISMATCH;
    output_matrix->cols = output_matrix->rows * output_matrix->cols;
    output_matrix->rows = 1;
    return EIDSP_OK;
}

/* reference (L432 copy) ei_run_dsp.h:369-418; the output matrix becomes [1][frames * num_filters] (:414-415) */
__attribute__((unused)) static int eikws_dropin_extract_mfe(ei::signal_t *signal, ei::matrix_t *output_matrix, void *config_ptr) {
    eikws_handle *h = eikws_dropin_handle();
    if (!h) return -1004;
    if (!config_ptr) return EIDSP_MATRIX_SIZE_MISMATCH;
    eikws_dropin_current_signal = signal;
    int rc = eikws_extract_mfe_signal(h, &eikws_dropin_get_data, signal->total_length, (const eikws_mfe_config *)config_ptr,
                                      output_matrix->buffer, (size_t)output_matrix->rows * output_matrix->cols);
    if (rc != EIKWS_OK) return EIDSP_MATRIX_SIZE_MISMATCH;
    output_matrix->cols = (uint32_t)eikws_mfe_feature_count(h);
    output_matrix->rows = 1;
    return EIDSP_OK;
}

namespace {

/* continuous mode state of the (single) stream this process serves, like the reference's statics (:116-121, :187) */
static eikws_streams *eikws_dropin_stream = NULL;
static bool eikws_dropin_stream_started = false;

/* reference :164-172.  Note: the reference's init leaves extract_mfcc_per_slice_features' `first_run` and the feature
 * window untouched (they can only be reset by a restart); here init returns the stream to its power-up state. */
extern "C" void run_classifier_init(void) {
    if (eikws_dropin_stream) eikws_streams_reset(eikws_dropin_stream);
    eikws_dropin_stream_started = false;
}

/* reference :184-282: one slice (EI_CLASSIFIER_SLICE_SIZE samples) per call; result is filled once the 1-second window
 * is full (values are the moving-average-filtered probabilities), otherwise left untouched like the reference does. */
extern "C" EI_IMPULSE_ERROR run_classifier_continuous(signal_t *signal, ei_impulse_result_t *result, bool debug = false) {
    eikws_handle *h = eikws_dropin_handle();
    if (!h) return EI_IMPULSE_TFLITE_ARENA_ALLOC_FAILED;
    if (!eikws_dropin_stream && eikws_streams_create(h, 1, EI_CLASSIFIER_SLICES_PER_MODEL_WINDOW, &eikws_dropin_stream) != EIKWS_OK) {
        ei_printf("ERR: eikws-b200 continuous mode: %s\n", eikws_last_error());
        return EI_IMPULSE_ALLOC_FAILED;
    }
    if (signal->total_length != (size_t)EI_CLASSIFIER_SLICE_SIZE) return EI_IMPULSE_DSP_ERROR;
    static float slice[EI_CLASSIFIER_SLICE_SIZE];
    uint64_t t0 = ei_read_timer_ms();
    if (signal->get_data(0, EI_CLASSIFIER_SLICE_SIZE, slice) != 0) return EI_IMPULSE_DSP_ERROR;
    float beyond = 0.0f;
    if (eikws_dropin_stream_started) {
        /* the reference pretends the slice is one frame longer (ei_run_dsp.h:322-324, it mutates the caller's signal_t)
         * and its pre-emphasis then asks the callback for the LAST sample of that longer signal (processing.hpp:68) */
        const ei_dsp_config_mfcc_t *c = (const ei_dsp_config_mfcc_t *)ei_dsp_blocks[0].config;
        signal->total_length += (size_t)(c->frame_length * (float)EI_CLASSIFIER_FREQUENCY);
        if (signal->get_data(signal->total_length - 1, 1, &beyond) != 0) return EI_IMPULSE_DSP_ERROR;
    }
    eikws_dropin_stream_started = true;
    float values[EI_CLASSIFIER_LABEL_COUNT];
    int has_result = 0;
    int rc = eikws_streams_push_f32_host(eikws_dropin_stream, slice, beyond, values, &has_result);
    if (rc != EIKWS_OK) return eikws_dropin_error(rc);
    result->timing.dsp = (int)(ei_read_timer_ms() - t0);
    result->timing.classification = 0;
    if (has_result) eikws_dropin_fill_result(result, values, debug);
    if (ei_run_impulse_check_canceled() == EI_IMPULSE_CANCELED) return EI_IMPULSE_CANCELED;
    return EI_IMPULSE_OK;
}

/* reference :293-641: classify an already-extracted feature matrix */
extern "C" EI_IMPULSE_ERROR run_inference(ei::matrix_t *fmatrix, ei_impulse_result_t *result, bool debug = false) {
    eikws_handle *h = eikws_dropin_handle();
    if (!h) return EI_IMPULSE_TFLITE_ARENA_ALLOC_FAILED;
    if (fmatrix->rows * fmatrix->cols != (uint32_t)EI_CLASSIFIER_NN_INPUT_FRAME_SIZE) return EI_IMPULSE_ERROR_SHAPES_DONT_MATCH;
    float values[EI_CLASSIFIER_LABEL_COUNT];
    uint64_t t0 = ei_read_timer_ms();
    int rc = eikws_infer_host(h, fmatrix->buffer, 1, values);
    result->timing.classification = (int)(ei_read_timer_ms() - t0);
    if (rc != EIKWS_OK) return eikws_dropin_error(rc);
    if (debug) ei_printf("Predictions (time: %d ms.):\n", result->timing.classification);
    eikws_dropin_fill_result(result, values, debug);
    if (ei_run_impulse_check_canceled() == EI_IMPULSE_CANCELED) return EI_IMPULSE_CANCELED;
    return EI_IMPULSE_OK;
}

/* reference :650-714 */
extern "C" EI_IMPULSE_ERROR run_classifier(signal_t *signal, ei_impulse_result_t *result, bool debug = false) {
    eikws_handle *h = eikws_dropin_handle();
    if (!h) return EI_IMPULSE_TFLITE_ARENA_ALLOC_FAILED;
    float values[EI_CLASSIFIER_LABEL_COUNT];
    int t_dsp = 0, t_cls = 0;
    eikws_dropin_current_signal = signal;
    int rc = eikws_run_classifier_signal(h, &eikws_dropin_get_data, signal->total_length, values, &t_dsp, &t_cls);
    if (rc != EIKWS_OK) {
        if (rc == EIKWS_ERR_DSP) ei_printf("ERR: Failed to run DSP process (%d)\n", rc);
        return eikws_dropin_error(rc);
    }
    result->timing.sampling = 0;
    result->timing.dsp = t_dsp;
    result->timing.classification = t_cls;
    result->timing.anomaly = 0;
    if (debug) ei_printf("Predictions (DSP+NN fused, time: %d ms.):\n", t_dsp);
    eikws_dropin_fill_result(result, values, debug);
    if (ei_run_impulse_check_canceled() == EI_IMPULSE_CANCELED) return EI_IMPULSE_CANCELED;
    return EI_IMPULSE_OK;
}

/* ---- batch extension: n_clips contiguous clips of EI_CLASSIFIER_RAW_SAMPLE_COUNT samples (host memory), sharded over the
 * devices of ei_b200_init when it was called ---- */
static EI_IMPULSE_ERROR eikws_dropin_batch(const void *clips, bool f32, size_t n_clips, ei_impulse_result_t *results) {
    eikws_handle *h = eikws_dropin_handle();
    if (!h) return EI_IMPULSE_TFLITE_ARENA_ALLOC_FAILED;
    float *values = (float *)malloc(sizeof(float) * EI_CLASSIFIER_LABEL_COUNT * (n_clips ? n_clips : 1));
    if (!values) return EI_IMPULSE_ALLOC_FAILED;
    int rc;
    if (eikws_dropin_devices)
        rc = f32 ? eikws_multi_classify_f32_host(eikws_dropin_devices, (const float *)clips, n_clips, values)
                 : eikws_multi_classify_i16_host(eikws_dropin_devices, (const int16_t *)clips, n_clips, values);
    else
        rc = f32 ? eikws_classify_f32_host(h, (const float *)clips, n_clips, values) : eikws_classify_i16_host(h, (const int16_t *)clips, n_clips, values);
    for (size_t i = 0; rc == EIKWS_OK && i < n_clips; i++) {
        memset(&results[i].timing, 0, sizeof(results[i].timing));
        eikws_dropin_fill_result(&results[i], values + i * EI_CLASSIFIER_LABEL_COUNT, false);
    }
    free(values);
    return eikws_dropin_error(rc);
}
extern "C" EI_IMPULSE_ERROR run_classifier_batch_i16(const int16_t *pcm, size_t n_clips, ei_impulse_result_t *results) {
    return eikws_dropin_batch(pcm, false, n_clips, results);
}
extern "C" EI_IMPULSE_ERROR run_classifier_batch_f32(const float *samples, size_t n_clips, ei_impulse_result_t *results) {
    return eikws_dropin_batch(samples, true, n_clips, results);
}

/* ---- device set (the library owns its device buffers behind an init / teardown pair; the two-argument call
 * sequence keeps working without it).  devices == NULL: the first n_devices visible GPUs (n_devices <= 0: all). ---- */
extern "C" void ei_b200_shutdown(void) {
    if (eikws_dropin_stream) eikws_streams_destroy(eikws_dropin_stream);
    eikws_dropin_stream = NULL;
    eikws_dropin_stream_started = false;
    if (eikws_dropin_devices) eikws_multi_destroy(eikws_dropin_devices);
    eikws_dropin_devices = NULL;
    if (eikws_dropin_single) eikws_destroy(eikws_dropin_single);
    eikws_dropin_single = NULL;
    eikws_dropin_tried = false;
}
extern "C" EI_IMPULSE_ERROR ei_b200_init(const int *devices, int n_devices) {
    ei_b200_shutdown();
    void *blob = NULL;
    size_t bytes = 0;
    int rc = eikws_dropin_capture(&blob, &bytes);
    if (rc != EIKWS_OK) return eikws_dropin_error(rc);
    rc = eikws_multi_create(blob, bytes, devices, n_devices, &eikws_dropin_devices);
    eikws_free(blob);
    if (rc != EIKWS_OK) {
        ei_printf("ERR: eikws-b200 could not create the device plans (%d): %s\n", rc, eikws_last_error());
        eikws_dropin_devices = NULL;
    }
    return eikws_dropin_error(rc);
}

}  // namespace
#endif /* __cplusplus */
#endif /* _EDGE_IMPULSE_RUN_CLASSIFIER_H_ */
