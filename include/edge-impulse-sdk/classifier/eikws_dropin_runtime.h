/* eikws-b200: forward declarations shared by the drop-in headers (defined in ei_run_classifier.h). */
#ifndef EIKWS_DROPIN_RUNTIME_H_
#define EIKWS_DROPIN_RUNTIME_H_
#include "../dsp/numpy_types.h"
static int eikws_dropin_extract_mfcc(ei::signal_t *signal, ei::matrix_t *output_matrix);
static int eikws_dropin_extract_mfe(ei::signal_t *signal, ei::matrix_t *output_matrix, void *config_ptr);
#endif
