// Host twin of the reference's known-answer harness (Arduino example static_buffer.ino:72-81 inside
// embedded-demos/arduino/.../ei-keyword-spotting-03-arduino-1.0.2.zip): classify one statically provided clip through
// the UNCHANGED reference call sequence -- signal_t + run_classifier -- against the drop-in header.
//
//   g++ -std=gnu++14 -I<repo>/include -I<export root> examples/static_buffer.cpp <export root>/tflite-model/trained_model_compiled.cpp \
//       -L<repo>/ei-keyword-spotting_b200 -leikws_b200 -Wl,-rpath,<repo>/ei-keyword-spotting_b200 -o static_buffer
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "edge-impulse-sdk/classifier/ei_run_classifier.h"

static int16_t clip[EI_CLASSIFIER_RAW_SAMPLE_COUNT];

// the firmware's callback (nucleo-l476-keyword-spotting/Core/Src/main.cpp:526-531): int16 -> float in [-1, 1)
static int get_signal_data(size_t offset, size_t length, float *out_ptr) {
    for (size_t i = 0; i < length; i++) out_ptr[i] = (float)clip[offset + i] / 32768.0f;
    return 0;
}

int main(int argc, char **argv) {
    // deterministic test tone + noise; a raw 16 kHz mono PCM_16 file may be given instead
    if (argc > 1) {
        FILE *f = fopen(argv[1], "rb");
        if (!f || fread(clip, sizeof(int16_t), EI_CLASSIFIER_RAW_SAMPLE_COUNT, f) != EI_CLASSIFIER_RAW_SAMPLE_COUNT) {
            fprintf(stderr, "cannot read %d samples from %s\n", EI_CLASSIFIER_RAW_SAMPLE_COUNT, argv[1]);
            return 2;
        }
        fclose(f);
    } else {
        uint32_t s = 12345;
        for (int i = 0; i < EI_CLASSIFIER_RAW_SAMPLE_COUNT; i++) {
            s = s * 1664525u + 1013904223u;
            clip[i] = (int16_t)(((i / 8) % 2 ? 4000 : -4000) + (int)((s >> 20) & 1023) - 512);
        }
    }
    signal_t signal;
    signal.total_length = EI_CLASSIFIER_RAW_SAMPLE_COUNT;
    signal.get_data = &get_signal_data;
    if (argc > 2 && !strcmp(argv[2], "mfe")) {
        // the sibling DSP block of the newer SDK copy, called the way generated dsp_blocks.h entries are called:
        // extract_fn(signal, matrix, config).  No shipped impulse carries an MFE block, so its config (the layout of
        // ei_dsp_config_mfe_t) is filled from the MFCC block's geometry.
        const ei_dsp_config_mfcc_t *mc = (const ei_dsp_config_mfcc_t *)ei_dsp_blocks[0].config;
        eikws_mfe_config cfg = {1, mc->frame_length, mc->frame_stride, (int)mc->num_filters, (int)mc->fft_length, (int)mc->low_frequency,
                                (int)mc->high_frequency, (int)mc->win_size};
        static float feats[4096];
        ei::matrix_t fm(1, 4096, feats);
        int rc = extract_mfe_features(&signal, &fm, &cfg);
        if (rc != 0) {
            printf("extract_mfe_features returned %d\n", rc);
            return 1;
        }
        printf("MFE features: %u x %u\n", (unsigned)fm.rows, (unsigned)fm.cols);
        for (uint32_t i = 0; i < fm.cols; i++) printf("mfe %u %.9g\n", (unsigned)i, feats[i]);
        return 0;
    }
    ei_impulse_result_t result;
    EI_IMPULSE_ERROR r = run_classifier(&signal, &result, false);
    if (r != EI_IMPULSE_OK) {
        printf("run_classifier returned %d\n", (int)r);
        return 1;
    }
    printf("Predictions (DSP: %d ms., Classification: %d ms.):\n", result.timing.dsp, result.timing.classification);
    for (size_t ix = 0; ix < EI_CLASSIFIER_LABEL_COUNT; ix++) printf("    %s: %.5f\n", result.classification[ix].label, result.classification[ix].value);
    return 0;
}
