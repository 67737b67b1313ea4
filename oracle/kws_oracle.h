/* TEST INFRASTRUCTURE ONLY -- see kws_oracle.c.  Not part of the product; the product
 * (ei-keyword-spotting_b200/) must never include, link or dlopen anything from oracle/. */
#ifndef KWS_ORACLE_H
#define KWS_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors ei_dsp_config_mfcc_t (model-parameters/model_metadata.h:92-104) plus the
 * sample rate (EI_CLASSIFIER_FREQUENCY, model_metadata.h:48). */
typedef struct {
    int num_cepstral;
    float frame_length;
    float frame_stride;
    int num_filters;
    int fft_length;
    int win_size;
    int low_frequency;
    int high_frequency; /* 0 => sample_rate/2 (feature.hpp:203-205) */
    float pre_cof;
    int pre_shift;
    int sample_rate;
} kws_mfcc_cfg;

/* optional stage taps; any pointer may be NULL */
typedef struct {
    float *filterbank; /* [(fft/2+1) x num_filters] transposed, as mfe() holds it          */
    float *power;      /* [frames x (fft/2+1)]  power spectrum                            */
    float *energy;     /* [frames]                                                        */
    float *mel;        /* [frames x num_filters] after zero handling, before log          */
    float *mfcc;       /* [frames x num_cepstral] before CMVN                             */
} kws_mfcc_taps;

int kws_oracle_num_frames(const kws_mfcc_cfg *c, int n_samples);
/* x: float samples as the signal_t callback would return them */
int kws_oracle_mfcc_f32(const kws_mfcc_cfg *c, const float *x, int n, float *features, kws_mfcc_taps *taps);
/* pcm: int16, converted exactly like numpy::int16_to_float (numpy.hpp:1289-1298) */
int kws_oracle_mfcc_i16(const kws_mfcc_cfg *c, const int16_t *pcm, int n, float *features, kws_mfcc_taps *taps);

/* the sibling MFE DSP block (extract_mfe_features of the reference's newer SDK copy) with the geometry of the MFCC block:
 * features[frames * num_filters]; returns the feature count or a negative error */
int kws_oracle_mfe_block_f32(const kws_mfcc_cfg *c, const float *x, int n, float *features);
int kws_oracle_mfe_block_i16(const kws_mfcc_cfg *c, const int16_t *pcm, int n, float *features);

/* ---- classifier (model container = the raw graph written by tools/ingest) ---- */
typedef struct kws_model kws_model;
kws_model *kws_model_load(const void *blob, size_t bytes); /* NULL on parse error */
void kws_model_free(kws_model *m);
int kws_model_num_labels(const kws_model *m);
int kws_model_num_features(const kws_model *m);
int kws_model_num_tensors(const kws_model *m);
const kws_mfcc_cfg *kws_model_mfcc_cfg(const kws_model *m);
const char *kws_model_label(const kws_model *m, int i);
/* run_inference (ei_run_classifier.h:341-493): features -> probs; optional copy of every
 * tensor's bytes after invoke (tensor_out[i] may be NULL; sized by kws_model_tensor_bytes). */
int kws_oracle_run_inference(const kws_model *m, const float *features, float *probs, void **tensor_out);
int kws_model_tensor_bytes(const kws_model *m, int i);
/* run_classifier (ei_run_classifier.h:650-714) on int16 PCM */
int kws_oracle_run_classifier_i16(const kws_model *m, const int16_t *pcm, int n, float *probs, float *features_out);
int kws_oracle_run_classifier_f32(const kws_model *m, const float *x, int n, float *probs, float *features_out);

/* the CMVN stage (processing.hpp:326-389) + input quantisation (ei_run_classifier.h:436-444) on caller-supplied pre-CMVN
 * cepstra [frames x num_cepstral]; features_out (optional) = the float CMVN output */
int kws_oracle_cmvn_quantise(const kws_model *m, const float *cepstra, int8_t *q, float *features_out);

/* ---- continuous mode: run_classifier_continuous (ei_run_classifier.h:184-282) for ONE stream from power-up ---- */
typedef struct kws_stream kws_stream;
kws_stream *kws_stream_new(const kws_model *m, int slices_per_window); /* EI_CLASSIFIER_SLICES_PER_MODEL_WINDOW (4) */
void kws_stream_free(kws_stream *s);
int kws_stream_slice_size(const kws_stream *s);
/* one slice of slice_size samples; `beyond` = what the signal callback returns for indices past the slice (see
 * oracle/ref_harness.cpp); *has_result = 1 once the feature window is full (probs are the MAF-filtered values) */
int kws_stream_push_i16(kws_stream *s, const int16_t *slice, int16_t beyond, float *probs, int *has_result);

#ifdef __cplusplus
}
#endif
#endif
