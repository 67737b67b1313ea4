// eikws-b200 model ingestion: compile this file TOGETHER WITH an UNMODIFIED Edge Impulse export
//   <export>/model-parameters/model_metadata.h
//   <export>/tflite-model/trained_model_compiled.cpp
// against this repo's include/ (which supplies edge-impulse-sdk/tensorflow/lite/c/*.h and micro_ops.h at the
// paths the generated code includes) and link libeikws_b200.so.  Running the binary executes the generated
// trained_model_init() against the library's recording Register_*() operators and writes the captured impulse
// as an "EIKWSMDL" container.  See tools/ingest_model.py for the driver and INTEGRATION.md.
//
// -DEIKWS_MFCC_CFG=<name of the ei_dsp_config_mfcc_t instance in model_metadata.h> is supplied by the driver
// (the instance is named after the DSP block id, e.g. ei_dsp_config_28 at model_metadata.h:120).
#include <stdio.h>
#include <stdlib.h>

#include "edge-impulse-sdk/tensorflow/lite/c/common.h"
#include "eikws_b200.h"
#include "model-parameters/model_metadata.h"

// the generated runtime (tflite-model/trained_model_compiled.cpp:380-475)
TfLiteStatus trained_model_init(void *(*alloc_fnc)(size_t, size_t));
TfLiteTensor *trained_model_input(int index);
TfLiteTensor *trained_model_output(int index);
TfLiteStatus trained_model_reset(void (*free_fnc)(void *ptr));

static int init_thunk(void *(*a)(size_t, size_t)) { return (int)trained_model_init(a); }
static void *input_thunk(int i) { return trained_model_input(i); }
static void *output_thunk(int i) { return trained_model_output(i); }
static int reset_thunk(void (*f)(void *)) { return (int)trained_model_reset(f); }

int main(int argc, char **argv) {
    if (argc != 2) {
        fprintf(stderr, "usage: %s <out.eikwsmdl>\n", argv[0]);
        return 2;
    }
    eikws_compiled_model_t cm;
    cm.init = init_thunk;
    cm.input = input_thunk;
    cm.output = output_thunk;
    cm.reset = reset_thunk;
    cm.raw_sample_count = EI_CLASSIFIER_RAW_SAMPLE_COUNT;
    cm.nn_input_frame_size = EI_CLASSIFIER_NN_INPUT_FRAME_SIZE;
    cm.label_count = EI_CLASSIFIER_LABEL_COUNT;
    cm.frequency = EI_CLASSIFIER_FREQUENCY;
    cm.labels = ei_classifier_inferencing_categories;
    cm.mfcc_num_cepstral = EIKWS_MFCC_CFG.num_cepstral;
    cm.mfcc_frame_length = EIKWS_MFCC_CFG.frame_length;
    cm.mfcc_frame_stride = EIKWS_MFCC_CFG.frame_stride;
    cm.mfcc_num_filters = EIKWS_MFCC_CFG.num_filters;
    cm.mfcc_fft_length = EIKWS_MFCC_CFG.fft_length;
    cm.mfcc_win_size = EIKWS_MFCC_CFG.win_size;
    cm.mfcc_low_frequency = EIKWS_MFCC_CFG.low_frequency;
    cm.mfcc_high_frequency = EIKWS_MFCC_CFG.high_frequency;
    cm.mfcc_pre_cof = EIKWS_MFCC_CFG.pre_cof;
    cm.mfcc_pre_shift = EIKWS_MFCC_CFG.pre_shift;
    void *blob = NULL;
    size_t bytes = 0;
    int rc = eikws_model_from_compiled(&cm, &blob, &bytes);
    if (rc != 0) {
        fprintf(stderr, "ingest failed (%d): %s\n", rc, eikws_last_error());
        return 1;
    }
    FILE *f = fopen(argv[1], "wb");
    if (!f || fwrite(blob, 1, bytes, f) != bytes) {
        fprintf(stderr, "cannot write %s\n", argv[1]);
        return 1;
    }
    fclose(f);
    eikws_free(blob);
    printf("%s: %zu bytes, %d labels\n", argv[1], bytes, (int)EI_CLASSIFIER_LABEL_COUNT);
    return 0;
}
