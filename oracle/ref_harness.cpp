// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or shipped with the product.
//
// Thin C-ABI harness around the UNMODIFIED reference (Edge Impulse SDK + generated
// model, compiled in place from /root/reference by oracle/Makefile into
// oracle/_ref/liboracle_ref_<model>.so).  It adds no arithmetic of its own: every
// number it returns is produced by the reference's own functions
//   run_classifier            edge-impulse-sdk/classifier/ei_run_classifier.h:650-714
//   run_inference             edge-impulse-sdk/classifier/ei_run_classifier.h:293-641
//   extract_mfcc_features     edge-impulse-sdk/classifier/ei_run_dsp.h:256-308
//   speechpy::feature::*      edge-impulse-sdk/dsp/speechpy/feature.hpp
// Stage taps (filterbank, mfe, pre-CMVN mfcc, every TFLite tensor) are read by
// calling the reference's internal functions directly / by intercepting
// trained_model_reset() so the arena can be copied out before it is freed.
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

// Intercept the arena teardown (ei_run_classifier.h:487) so intermediate tensors
// can be dumped.  Both the declaration (trained_model_compiled.h) and the call
// site are renamed by this macro; the real symbol is re-declared below.
#define trained_model_reset ref_hook_trained_model_reset
#include "edge-impulse-sdk/classifier/ei_run_classifier.h"
#undef trained_model_reset
TfLiteStatus trained_model_reset(void (*free_fnc)(void *ptr));

// ---- porting layer the reference expects the application to supply
// (edge-impulse-sdk/porting/ei_classifier_porting.h:48-76) -------------------
EI_IMPULSE_ERROR ei_sleep(int32_t) { return EI_IMPULSE_OK; }
EI_IMPULSE_ERROR ei_run_impulse_check_canceled() { return EI_IMPULSE_OK; }
uint64_t ei_read_timer_us() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (uint64_t)ts.tv_sec * 1000000ull + ts.tv_nsec / 1000;
}
uint64_t ei_read_timer_ms() { return ei_read_timer_us() / 1000; }
static int g_quiet = 1;
void ei_printf(const char *format, ...) {
    if (g_quiet) return;
    va_list ap;
    va_start(ap, format);
    vprintf(format, ap);
    va_end(ap);
}
void ei_printf_float(float f) { ei_printf("%f", f); }
void DebugLog(const char *s) { ei_printf("%s", s); }

// ---- signal sources ---------------------------------------------------------
static const int16_t *g_pcm = nullptr;
static const float *g_f32 = nullptr;
// The demos' callback: numpy::int16_to_float (numpy.hpp:1289-1298), see
// nucleo-l476-keyword-spotting/Core/Src/main.cpp:526-531.
static int get_data_i16(size_t offset, size_t length, float *out) {
    return ei::numpy::int16_to_float(g_pcm + offset, out, length);
}
static int get_data_f32(size_t offset, size_t length, float *out) {
    memcpy(out, g_f32 + offset, length * sizeof(float));
    return 0;
}

// ---- tensor dump taken just before the arena is freed ------------------------
#define REF_MAX_TENSORS 64
#define REF_MAX_TENSOR_BYTES 8192
static int g_dump_enabled = 0;
static int g_dump_count = 0;
static int g_dump_bytes[REF_MAX_TENSORS];
static uint8_t g_dump[REF_MAX_TENSORS][REF_MAX_TENSOR_BYTES];

TfLiteStatus ref_hook_trained_model_reset(void (*free_fnc)(void *ptr)) {
    if (g_dump_enabled) {
        // tensor 0 is the model input, so trained_model_input(0) is the table base
        // (tflite-model/trained_model_compiled.cpp:443-448).
        TfLiteTensor *t = trained_model_input(0);
        TfLiteTensor *out = trained_model_output(0);
        int n = (int)(out - t) + 1;
        if (n > REF_MAX_TENSORS) n = REF_MAX_TENSORS;
        g_dump_count = n;
        for (int i = 0; i < n; i++) {
            int b = (int)t[i].bytes;
            if (b > REF_MAX_TENSOR_BYTES) b = REF_MAX_TENSOR_BYTES;
            g_dump_bytes[i] = b;
            if (t[i].data.data) memcpy(g_dump[i], t[i].data.data, b);
        }
    }
    return trained_model_reset(free_fnc);
}

extern "C" {

void ref_set_quiet(int q) { g_quiet = q; }
int ref_label_count() { return EI_CLASSIFIER_LABEL_COUNT; }
int ref_feature_count() { return EI_CLASSIFIER_NN_INPUT_FRAME_SIZE; }
int ref_raw_sample_count() { return EI_CLASSIFIER_RAW_SAMPLE_COUNT; }
const char *ref_label(int i) { return ei_classifier_inferencing_categories[i]; }

// Full path, int16 PCM in (what the firmware feeds): probs[L], optional features[637].
int ref_run_classifier_i16(const int16_t *pcm, int n, float *probs) {
    g_pcm = pcm;
    signal_t sig;
    sig.total_length = (size_t)n;
    sig.get_data = &get_data_i16;
    ei_impulse_result_t res;
    memset(&res, 0, sizeof(res));
    EI_IMPULSE_ERROR e = run_classifier(&sig, &res, false);
    for (int i = 0; i < EI_CLASSIFIER_LABEL_COUNT; i++) probs[i] = res.classification[i].value;
    return (int)e;
}

int ref_run_classifier_f32(const float *x, int n, float *probs) {
    g_f32 = x;
    signal_t sig;
    sig.total_length = (size_t)n;
    sig.get_data = &get_data_f32;
    ei_impulse_result_t res;
    memset(&res, 0, sizeof(res));
    EI_IMPULSE_ERROR e = run_classifier(&sig, &res, false);
    for (int i = 0; i < EI_CLASSIFIER_LABEL_COUNT; i++) probs[i] = res.classification[i].value;
    return (int)e;
}

// DSP block only: features[EI_CLASSIFIER_NN_INPUT_FRAME_SIZE].
int ref_mfcc_i16(const int16_t *pcm, int n, float *features) {
    g_pcm = pcm;
    signal_t sig;
    sig.total_length = (size_t)n;
    sig.get_data = &get_data_i16;
    ei::matrix_t fm(1, EI_CLASSIFIER_NN_INPUT_FRAME_SIZE, features);
    return ei_dsp_blocks[0].extract_fn(&sig, &fm, ei_dsp_blocks[0].config);
}

int ref_mfcc_f32(const float *x, int n, float *features) {
    g_f32 = x;
    signal_t sig;
    sig.total_length = (size_t)n;
    sig.get_data = &get_data_f32;
    ei::matrix_t fm(1, EI_CLASSIFIER_NN_INPUT_FRAME_SIZE, features);
    return ei_dsp_blocks[0].extract_fn(&sig, &fm, ei_dsp_blocks[0].config);
}

// NN only: probs[L]; if tensors_out != NULL every TFLite tensor is copied out as
// consecutive (int32 bytes, payload) records; returns number of tensors via *n_tensors.
int ref_run_inference(const float *features, float *probs, uint8_t *tensors_out, int tensors_cap,
                      int *n_tensors) {
    ei::matrix_t fm(1, EI_CLASSIFIER_NN_INPUT_FRAME_SIZE, (float *)features);
    ei_impulse_result_t res;
    memset(&res, 0, sizeof(res));
    g_dump_enabled = tensors_out != nullptr;
    EI_IMPULSE_ERROR e = run_inference(&fm, &res, false);
    g_dump_enabled = 0;
    for (int i = 0; i < EI_CLASSIFIER_LABEL_COUNT; i++) probs[i] = res.classification[i].value;
    if (tensors_out) {
        int off = 0;
        int n = 0;
        for (int i = 0; i < g_dump_count; i++) {
            if (off + 4 + g_dump_bytes[i] > tensors_cap) break;
            int32_t b = g_dump_bytes[i];
            memcpy(tensors_out + off, &b, 4);
            memcpy(tensors_out + off + 4, g_dump[i], b);
            off += 4 + b;
            n++;
        }
        if (n_tensors) *n_tensors = n;
    }
    return (int)e;
}

// ---- stage taps of the DSP block (reference functions called directly) --------
// Mel filterbank exactly as mfe() builds it (feature.hpp:243-253): out[coefficients x num_filters], transposed.
int ref_filterbank(float *out, int *rows, int *cols) {
    ei_dsp_config_mfcc_t c = *(ei_dsp_config_mfcc_t *)ei_dsp_blocks[0].config;
    uint16_t coefficients = c.fft_length / 2 + 1;
    uint32_t high = c.high_frequency == 0 ? EI_CLASSIFIER_FREQUENCY / 2 : c.high_frequency;
    ei::matrix_t fb(c.num_filters, coefficients);
    int r = ei::speechpy::feature::filterbanks(&fb, c.num_filters, coefficients, EI_CLASSIFIER_FREQUENCY,
                                               c.low_frequency, high, true);
    *rows = fb.rows;
    *cols = fb.cols;
    memcpy(out, fb.buffer, sizeof(float) * fb.rows * fb.cols);
    return r;
}

static class ei::speechpy::processing::preemphasis *g_pre = nullptr;
static int pre_get_data(size_t offset, size_t length, float *out) { return g_pre->get_data(offset, length, out); }

// mfe taps: mel[frames x num_filters] (after zero handling), energy[frames]   (feature.hpp:193-318)
int ref_mfe_i16(const int16_t *pcm, int n, float *mel, float *energy, int *frames) {
    ei_dsp_config_mfcc_t c = *(ei_dsp_config_mfcc_t *)ei_dsp_blocks[0].config;
    g_pcm = pcm;
    signal_t sig;
    sig.total_length = (size_t)n;
    sig.get_data = &get_data_i16;
    class ei::speechpy::processing::preemphasis pre(&sig, c.pre_shift, c.pre_cof);
    g_pre = &pre;
    signal_t psig;
    psig.total_length = (size_t)n;
    psig.get_data = &pre_get_data;
    ei::matrix_size_t sz =
        ei::speechpy::feature::calculate_mfe_buffer_size(n, EI_CLASSIFIER_FREQUENCY, c.frame_length, c.frame_stride, c.num_filters);
    ei::matrix_t fm(sz.rows, sz.cols, mel);
    ei::matrix_t em(sz.rows, 1, energy);
    *frames = sz.rows;
    return ei::speechpy::feature::mfe(&fm, &em, &psig, EI_CLASSIFIER_FREQUENCY, c.frame_length, c.frame_stride,
                                      c.num_filters, c.fft_length, c.low_frequency, c.high_frequency);
}

// mfcc before CMVN: out[frames x num_cepstral]   (feature.hpp:370-439)
int ref_mfcc_nocmvn_i16(const int16_t *pcm, int n, float *out, int *frames) {
    ei_dsp_config_mfcc_t c = *(ei_dsp_config_mfcc_t *)ei_dsp_blocks[0].config;
    g_pcm = pcm;
    signal_t sig;
    sig.total_length = (size_t)n;
    sig.get_data = &get_data_i16;
    class ei::speechpy::processing::preemphasis pre(&sig, c.pre_shift, c.pre_cof);
    g_pre = &pre;
    signal_t psig;
    psig.total_length = (size_t)n;
    psig.get_data = &pre_get_data;
    ei::matrix_size_t sz =
        ei::speechpy::feature::calculate_mfcc_buffer_size(n, EI_CLASSIFIER_FREQUENCY, c.frame_length, c.frame_stride, c.num_cepstral);
    ei::matrix_t fm(sz.rows, sz.cols, out);
    *frames = sz.rows;
    return ei::speechpy::feature::mfcc(&fm, &psig, EI_CLASSIFIER_FREQUENCY, c.frame_length, c.frame_stride, c.num_cepstral,
                                       c.num_filters, c.fft_length, c.low_frequency, c.high_frequency);
}

// The reference's own sliding-window CMVN (processing.hpp:326-389, called by extract_mfcc_features at ei_run_dsp.h:297-302)
// on a caller-supplied [frames x num_cepstral] matrix of pre-CMVN cepstra, in place: lets the tests drive the stage with
// matrices no audio clip produces.
int ref_cmvnw_f32(float *m, int frames) {
    const ei_dsp_config_mfcc_t *c = (const ei_dsp_config_mfcc_t *)ei_dsp_blocks[0].config;
    ei::matrix_t fm(frames, c->num_cepstral, m);
    return ei::speechpy::processing::cmvnw(&fm, c->win_size, true);
}

// One frame's magnitude spectrum via numpy::rfft (numpy.hpp:1091-1156): out[n_fft/2+1]
#ifdef REF_HAS_MFE_BLOCK
// The sibling DSP block of the newer SDK copy (L432): extract_mfe_features (ei_run_dsp.h:369-418) = mel filterbank energies
// (no pre-emphasis, no log), sliding-window mean subtraction, min/max scaling of the whole matrix.  The block's config
// takes the geometry of the model's MFCC block (no shipped impulse carries an MFE block of its own).
int ref_mfe_block_i16(const int16_t *pcm, int n, float *features, int capacity) {
    const ei_dsp_config_mfcc_t *mc = (const ei_dsp_config_mfcc_t *)ei_dsp_blocks[0].config;
    ei_dsp_config_mfe_t cfg;
    cfg.axes = 1;
    cfg.frame_length = mc->frame_length;
    cfg.frame_stride = mc->frame_stride;
    cfg.num_filters = mc->num_filters;
    cfg.fft_length = mc->fft_length;
    cfg.low_frequency = mc->low_frequency;
    cfg.high_frequency = mc->high_frequency;
    cfg.win_size = mc->win_size;
    g_pcm = pcm;
    signal_t sig;
    sig.total_length = (size_t)n;
    sig.get_data = &get_data_i16;
    ei::matrix_t fm(1, capacity, features);
    int r = extract_mfe_features(&sig, &fm, &cfg);
    return r == 0 ? (int)(fm.rows * fm.cols) : r;
}
#endif

int ref_rfft_mag(const float *frame, int frame_len, int n_fft, float *out) {
    return ei::numpy::rfft(frame, frame_len, out, n_fft / 2 + 1, n_fft);
}

// ---- continuous mode (run_classifier_continuous, ei_run_classifier.h:184-282) -------------------------------
// One call = one slice of EI_CLASSIFIER_SLICE_SIZE samples.  The reference keeps its state in statics that
// run_classifier_init() only partly resets (extract_mfcc_per_slice_features' `first_run`, ei_run_dsp.h:313, and the
// feature matrix are never reset), so one loaded copy of this library models exactly ONE stream from power-up.
// From the second slice on the reference asks the callback for sample total_length-1 = SLICE_SIZE+319, i.e. beyond
// the slice (ei_run_dsp.h:322-324 + processing.hpp:68); the firmware then reads past its buffer.  Here the slice is
// staged in a buffer whose tail [SLICE_SIZE, SLICE_SIZE+320) holds `beyond` (as int16), which makes that read defined.
static int16_t g_slice_buf[EI_CLASSIFIER_SLICE_SIZE + 1024];
int ref_slice_size() { return EI_CLASSIFIER_SLICE_SIZE; }
int ref_run_classifier_continuous_i16(const int16_t *slice, int16_t beyond, float *probs, int *has_result) {
    memcpy(g_slice_buf, slice, sizeof(int16_t) * EI_CLASSIFIER_SLICE_SIZE);
    for (int i = EI_CLASSIFIER_SLICE_SIZE; i < EI_CLASSIFIER_SLICE_SIZE + 1024; i++) g_slice_buf[i] = beyond;
    g_pcm = g_slice_buf;
    signal_t sig;
    sig.total_length = EI_CLASSIFIER_SLICE_SIZE;
    sig.get_data = &get_data_i16;
    ei_impulse_result_t res;
    memset(&res, 0, sizeof(res));
    for (int i = 0; i < EI_CLASSIFIER_LABEL_COUNT; i++) res.classification[i].value = -1.0f;  // untouched until the window is full
    EI_IMPULSE_ERROR e = run_classifier_continuous(&sig, &res, false);
    *has_result = res.classification[0].value >= 0.0f;
    for (int i = 0; i < EI_CLASSIFIER_LABEL_COUNT; i++) probs[i] = res.classification[i].value;
    return (int)e;
}

// CPU timing loop used by bench.py's reference arm: runs run_classifier over
// `count` clips laid out back to back, returns seconds elapsed.
double ref_time_run_classifier_i16(const int16_t *pcm, int n, int count, float *probs_last) {
    uint64_t t0 = ei_read_timer_us();
    for (int i = 0; i < count; i++) ref_run_classifier_i16(pcm + (size_t)i * n, n, probs_last);
    return (double)(ei_read_timer_us() - t0) * 1e-6;
}

// Same loops, but every clip's probabilities are kept ([count][labels]): bench.py compares them with the GPU's outputs
// for the same clips, so the timed CPU baseline doubles as a parity check against the unmodified reference.
double ref_time_run_classifier_i16_all(const int16_t *pcm, int n, int count, float *probs_all) {
    uint64_t t0 = ei_read_timer_us();
    for (int i = 0; i < count; i++) ref_run_classifier_i16(pcm + (size_t)i * n, n, probs_all + (size_t)i * EI_CLASSIFIER_LABEL_COUNT);
    return (double)(ei_read_timer_us() - t0) * 1e-6;
}
double ref_time_run_classifier_f32_all(const float *x, int n, int count, float *probs_all) {
    uint64_t t0 = ei_read_timer_us();
    for (int i = 0; i < count; i++) ref_run_classifier_f32(x + (size_t)i * n, n, probs_all + (size_t)i * EI_CLASSIFIER_LABEL_COUNT);
    return (double)(ei_read_timer_us() - t0) * 1e-6;
}

}  // extern "C"
