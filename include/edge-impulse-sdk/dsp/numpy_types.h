/* eikws-b200 drop-in for edge-impulse-sdk/dsp/numpy_types.h: the two types that cross the run_classifier
 * boundary -- ei::signal_t (reference :234-253) and ei::matrix_t (reference :55-127). */
#ifndef EIKWS_EIDSP_NUMPY_TYPES_H_
#define EIKWS_EIDSP_NUMPY_TYPES_H_

#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef __cplusplus
#if !defined(EIDSP_SIGNAL_C_FN_POINTER) || EIDSP_SIGNAL_C_FN_POINTER == 0
#include <functional>
#endif
namespace ei {
#endif

/* A matrix that owns its (calloc'ed) buffer unless one is handed in -- same contract as the reference. */
typedef struct ei_matrix {
    float *buffer;
    uint32_t rows;
    uint32_t cols;
    bool buffer_managed_by_me;
#ifdef __cplusplus
    ei_matrix(uint32_t n_rows, uint32_t n_cols, float *a_buffer = NULL) {
        if (a_buffer) {
            buffer = a_buffer;
            buffer_managed_by_me = false;
        } else {
            buffer = (float *)calloc((size_t)n_rows * n_cols * sizeof(float), 1);
            buffer_managed_by_me = true;
        }
        rows = n_rows;
        cols = n_cols;
    }
    ~ei_matrix() {
        if (buffer && buffer_managed_by_me) free(buffer);
    }
#endif
} matrix_t;

/* Sensor signal: a pull callback plus the total length.  The callee pulls; no sample outside
 * [0, total_length) is requested; a non-zero return aborts the run with EI_IMPULSE_DSP_ERROR. */
typedef struct ei_signal_t {
#if defined(EIDSP_SIGNAL_C_FN_POINTER) && EIDSP_SIGNAL_C_FN_POINTER == 1 || !defined(__cplusplus)
    int (*get_data)(size_t, size_t, float *);
#else
    std::function<int(size_t offset, size_t length, float *out_ptr)> get_data;
#endif
    size_t total_length;
} signal_t;

#ifdef __cplusplus
}  // namespace ei
#endif
#endif /* EIKWS_EIDSP_NUMPY_TYPES_H_ */
