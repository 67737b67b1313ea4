/* eikws-b200 drop-in for edge-impulse-sdk/classifier/ei_model_types.h (reference :30-34): the DSP block
 * descriptor the generated model-parameters/dsp_blocks.h instantiates. */
#ifndef EIKWS_EI_MODEL_TYPES_H_
#define EIKWS_EI_MODEL_TYPES_H_

#include <stddef.h>
#include <stdint.h>

#include "../dsp/numpy_types.h"

typedef struct {
    size_t n_output_features;
    int (*extract_fn)(ei::signal_t *signal, ei::matrix_t *output_matrix, void *config);
    void *config;
} ei_model_dsp_t;

#endif /* EIKWS_EI_MODEL_TYPES_H_ */
