/* eikws-b200 drop-in for edge-impulse-sdk/classifier/ei_run_dsp.h: the DSP block entry point whose address the
 * generated dsp_blocks.h stores (`&extract_mfcc_features`, dsp_blocks.h:33).  Reference: ei_run_dsp.h:256-308.
 * Here it is a thin call into libeikws_b200.so (the fused MFCC kernel run in features-only mode). */
#ifndef EIKWS_EI_RUN_DSP_H_
#define EIKWS_EI_RUN_DSP_H_

#include "../dsp/numpy_types.h"
#include "eikws_dropin_runtime.h"

#ifndef EIDSP_OK
#define EIDSP_OK 0
#define EIDSP_MATRIX_SIZE_MISMATCH (-1002)
#endif

__attribute__((unused)) static int extract_mfcc_features(ei::signal_t *signal, ei::matrix_t *output_matrix, void *config_ptr) {
    (void)config_ptr;  // the configuration was captured when the impulse was lowered (eikws_dropin_handle)
    return eikws_dropin_extract_mfcc(signal, output_matrix);
}

/* The sibling MFE block of the newer SDK copy (nucleo-l432-keyword-spotting/.../classifier/ei_run_dsp.h:369-418).
 * config_ptr points at the application's ei_dsp_config_mfe_t (model-parameters/model_metadata.h:103-112 of that export);
 * it is read through a layout twin so that exports without that type (L476) still compile. */
__attribute__((unused)) static int extract_mfe_features(ei::signal_t *signal, ei::matrix_t *output_matrix, void *config_ptr) {
    return eikws_dropin_extract_mfe(signal, output_matrix, config_ptr);
}

#endif /* EIKWS_EI_RUN_DSP_H_ */
