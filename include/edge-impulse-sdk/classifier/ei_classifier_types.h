/* eikws-b200 drop-in for edge-impulse-sdk/classifier/ei_classifier_types.h (reference :30-53): the
 * caller-allocated result POD.  `label` points at the static strings of model_metadata.h. */
#ifndef EIKWS_EI_CLASSIFIER_TYPES_H_
#define EIKWS_EI_CLASSIFIER_TYPES_H_

#include <stdint.h>

#include "model-parameters/model_metadata.h"

typedef struct {
    const char *label;
    float value;
} ei_impulse_result_classification_t;

typedef struct {
    int sampling;
    int dsp;
    int classification;
    int anomaly;
} ei_impulse_result_timing_t;

typedef struct {
    ei_impulse_result_classification_t classification[EI_CLASSIFIER_LABEL_COUNT];
    float anomaly;
    ei_impulse_result_timing_t timing;
} ei_impulse_result_t;

typedef struct {
    uint32_t buf_idx;
    float running_sum;
    float maf_buffer[EI_CLASSIFIER_SLICES_PER_MODEL_WINDOW >> 1];
} ei_impulse_maf;

#endif /* EIKWS_EI_CLASSIFIER_TYPES_H_ */
