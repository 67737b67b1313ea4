/* Pure-C host twin of the reference's known-answer harness (Arduino example static_buffer.ino:72-81 inside
 * embedded-demos/arduino/.../ei-keyword-spotting-03-arduino-1.0.2.zip): one statically provided clip through the reference's
 * call sequence -- signal_t + run_classifier -- from a C11 translation unit.  The C wrapper the reference's README has
 * users delete (README.md:185) is replaced by edge-impulse-sdk/classifier/ei_run_classifier_c.cpp.
 *
 *   gcc -std=c11 -c examples/static_buffer.c -I<repo>/include -I<export>
 *   g++ -std=gnu++14 -c <repo>/include/edge-impulse-sdk/classifier/ei_run_classifier_c.cpp <export>/tflite-model/trained_model_compiled.cpp \
 *       -I<repo>/include -I<export>
 *   g++ static_buffer.o ei_run_classifier_c.o trained_model_compiled.o -L<repo>/ei-keyword-spotting_b200 -leikws_b200 \
 *       -Wl,-rpath,<repo>/ei-keyword-spotting_b200 -o static_buffer_c
 */
#include <stdint.h>
#include <stdio.h>

#include "edge-impulse-sdk/classifier/ei_run_classifier.h"

static int16_t clip[EI_CLASSIFIER_RAW_SAMPLE_COUNT];

/* the firmware's callback (nucleo-l476-keyword-spotting/Core/Src/main.cpp:526-531): int16 -> float in [-1, 1) */
static int get_signal_data(size_t offset, size_t length, float *out_ptr) {
    for (size_t i = 0; i < length; i++) out_ptr[i] = (float)clip[offset + i] / 32768.0f;
    return 0;
}

int main(int argc, char **argv) {
    if (argc > 1) { /* a raw 16 kHz mono PCM_16 file */
        FILE *f = fopen(argv[1], "rb");
        if (!f || fread(clip, sizeof(int16_t), EI_CLASSIFIER_RAW_SAMPLE_COUNT, f) != EI_CLASSIFIER_RAW_SAMPLE_COUNT) {
            fprintf(stderr, "cannot read %d samples from %s\n", EI_CLASSIFIER_RAW_SAMPLE_COUNT, argv[1]);
            return 2;
        }
        fclose(f);
    } else { /* the deterministic test tone + noise of examples/static_buffer.cpp */
        uint32_t s = 12345;
        for (int i = 0; i < EI_CLASSIFIER_RAW_SAMPLE_COUNT; i++) {
            s = s * 1664525u + 1013904223u;
            clip[i] = (int16_t)(((i / 8) % 2 ? 4000 : -4000) + (int)((s >> 20) & 1023) - 512);
        }
    }
    signal_t signal;
    signal.total_length = EI_CLASSIFIER_RAW_SAMPLE_COUNT;
    signal.get_data = &get_signal_data;
    ei_impulse_result_t result;
    EI_IMPULSE_ERROR r = run_classifier(&signal, &result, false);
    if (r != EI_IMPULSE_OK) {
        printf("run_classifier returned %d\n", (int)r);
        return 1;
    }
    printf("Predictions (DSP: %d ms., Classification: %d ms.):\n", result.timing.dsp, result.timing.classification);
    for (size_t ix = 0; ix < EI_CLASSIFIER_LABEL_COUNT; ix++) printf("    %s: %.5f\n", result.classification[ix].label, result.classification[ix].value);
    ei_b200_shutdown();
    return 0;
}
