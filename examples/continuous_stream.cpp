// Host twin of the firmware main loop (nucleo-l476-keyword-spotting/Core/Src/main.cpp:178-232): feed the audio one
// 250 ms slice at a time through the UNCHANGED reference call sequence -- signal_t + run_classifier_continuous.
// usage: continuous_stream <raw 16 kHz mono PCM_16 file>   (prints one line per slice once the window is full)
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "edge-impulse-sdk/classifier/ei_run_classifier.h"

static int16_t *audio = NULL;
static size_t n_audio = 0, slice_start = 0;

// like the firmware's callback (main.cpp:526-531); indices past the current slice return 0 (the firmware reads past
// its buffer there -- the reference asks for one such sample per slice)
static int get_audio_signal_data(size_t offset, size_t length, float *out_ptr) {
    for (size_t i = 0; i < length; i++) {
        size_t k = offset + i;
        out_ptr[i] = k < EI_CLASSIFIER_SLICE_SIZE ? (float)audio[slice_start + k] / 32768.0f : 0.0f;
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc != 2) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    fseek(f, 0, SEEK_END);
    n_audio = (size_t)ftell(f) / 2;
    fseek(f, 0, SEEK_SET);
    audio = (int16_t *)malloc(n_audio * 2);
    if (fread(audio, 2, n_audio, f) != n_audio) return 2;
    fclose(f);
    run_classifier_init();
    for (slice_start = 0; slice_start + EI_CLASSIFIER_SLICE_SIZE <= n_audio; slice_start += EI_CLASSIFIER_SLICE_SIZE) {
        signal_t signal;
        signal.total_length = EI_CLASSIFIER_SLICE_SIZE;
        signal.get_data = &get_audio_signal_data;
        ei_impulse_result_t result;
        for (size_t ix = 0; ix < EI_CLASSIFIER_LABEL_COUNT; ix++) result.classification[ix].value = -1.0f;
        EI_IMPULSE_ERROR r = run_classifier_continuous(&signal, &result, false);
        if (r != EI_IMPULSE_OK) {
            printf("run_classifier_continuous returned %d\n", (int)r);
            return 1;
        }
        if (result.classification[0].value < 0.0f) continue;  // window not full yet
        printf("slice %zu:", slice_start / EI_CLASSIFIER_SLICE_SIZE);
        for (size_t ix = 0; ix < EI_CLASSIFIER_LABEL_COUNT; ix++) printf(" %.8f", result.classification[ix].value);
        printf("\n");
    }
    return 0;
}
