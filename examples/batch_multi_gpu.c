/* Pure-C batch host: every GPU of the box behind the reference's data types.  The reference's application loop
 * (nucleo-l476-keyword-spotting/Core/Src/main.cpp:190-194) classifies one window after the other on one core; here a batch
 * of clips is sharded contiguously over the devices given to ei_b200_init (no exchange between devices) and the results
 * come back as the reference's ei_impulse_result_t records.
 *
 *   batch_multi_gpu_c [n_clips] [n_devices, 0 = all visible]
 * prints a checksum of the probabilities of the whole batch computed on n_devices GPUs and again on one GPU: they must agree
 * byte for byte (shard invariance).  Build: see examples/static_buffer.c.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "edge-impulse-sdk/classifier/ei_run_classifier.h"

/* splitmix64-based deterministic noise, one stream per clip */
static uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static uint64_t checksum(const ei_impulse_result_t *r, size_t n) {
    uint64_t h = 1469598103934665603ull; /* FNV-1a over the float bit patterns */
    for (size_t i = 0; i < n; i++)
        for (size_t l = 0; l < EI_CLASSIFIER_LABEL_COUNT; l++) {
            uint32_t bits;
            memcpy(&bits, &r[i].classification[l].value, 4);
            h = (h ^ bits) * 1099511628211ull;
        }
    return h;
}

int main(int argc, char **argv) {
    const size_t n = argc > 1 ? (size_t)strtoull(argv[1], NULL, 10) : 4096;
    const int n_dev = argc > 2 ? atoi(argv[2]) : 0;
    int16_t *pcm = (int16_t *)eikws_host_alloc(n * EI_CLASSIFIER_RAW_SAMPLE_COUNT * sizeof(int16_t)); /* page-locked: full PCIe speed */
    ei_impulse_result_t *res = (ei_impulse_result_t *)malloc(n * sizeof(ei_impulse_result_t));
    if (!pcm || !res) {
        fprintf(stderr, "allocation failed: %s\n", eikws_last_error());
        return 2;
    }
    for (size_t c = 0; c < n; c++) {
        const int amp = 200 + (int)(mix(c) % 6000);
        for (size_t i = 0; i < EI_CLASSIFIER_RAW_SAMPLE_COUNT; i += 4) {
            uint64_t r = mix((c << 20) + i);
            for (int k = 0; k < 4; k++) pcm[c * EI_CLASSIFIER_RAW_SAMPLE_COUNT + i + k] = (int16_t)(((int)((r >> (16 * k)) & 0xffff) - 32768) * amp / 32768);
        }
    }
    uint64_t sums[2] = {0, 0};
    for (int pass = 0; pass < 2; pass++) {
        EI_IMPULSE_ERROR e = ei_b200_init(NULL, pass == 0 ? n_dev : 1);
        if (e != EI_IMPULSE_OK) {
            printf("ei_b200_init returned %d: %s\n", (int)e, eikws_last_error());
            return 1;
        }
        e = run_classifier_batch_i16(pcm, n, res);
        if (e != EI_IMPULSE_OK) {
            printf("run_classifier_batch_i16 returned %d: %s\n", (int)e, eikws_last_error());
            return 1;
        }
        sums[pass] = checksum(res, n);
        printf("pass %d (%s): %zu clips, checksum %016llx, clip 0:", pass, pass == 0 ? "requested devices" : "one device", n, (unsigned long long)sums[pass]);
        for (size_t l = 0; l < EI_CLASSIFIER_LABEL_COUNT; l++) printf(" %s=%.5f", res[0].classification[l].label, res[0].classification[l].value);
        printf("\n");
    }
    ei_b200_shutdown();
    eikws_host_free(pcm);
    free(res);
    if (sums[0] != sums[1]) {
        printf("MISMATCH between the sharded and the single-device run\n");
        return 1;
    }
    printf("sharded == single device\n");
    return 0;
}
