/* eikws-b200: C-linkage shim for pure-C hosts of the drop-in edge-impulse-sdk/classifier/ei_run_classifier.h.
 *
 * The reference defines run_classifier & co. inside a C++ header (anonymous namespace => internal linkage, reference
 * ei_run_classifier.h:106-108, :650-653) and tells users to delete the SDK's C wrapper (README.md:185).  A .c file that
 * includes the drop-in header only gets declarations; this translation unit, compiled with g++ and linked into the same
 * program, defines them with C linkage by forwarding to the header's implementation.
 *
 *   gcc -std=c11 -c app.c -I<repo>/include -I<export>
 *   g++ -std=gnu++14 -c <repo>/include/edge-impulse-sdk/classifier/ei_run_classifier_c.cpp -I<repo>/include -I<export>
 *   g++ -std=gnu++14 -c <export>/tflite-model/trained_model_compiled.cpp -I<repo>/include -I<export>
 *   g++ app.o ei_run_classifier_c.o trained_model_compiled.o -L<repo>/ei-keyword-spotting_b200 -leikws_b200
 */
#define EIDSP_SIGNAL_C_FN_POINTER 1 /* the C signal_t carries a plain function pointer (reference numpy_types.h:242-249) */
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

/* The generated model_metadata.h DEFINES globals (ei_classifier_inferencing_categories, ei_dsp_config_<id>), so exactly one
 * translation unit of a program may own them -- the application's .c file.  This unit's copy gets internal linkage. */
namespace {
#include "model-parameters/model_metadata.h"
}

#define run_classifier eikws_cxx_run_classifier
#define run_classifier_continuous eikws_cxx_run_classifier_continuous
#define run_classifier_init eikws_cxx_run_classifier_init
#define run_classifier_batch_i16 eikws_cxx_run_classifier_batch_i16
#define run_classifier_batch_f32 eikws_cxx_run_classifier_batch_f32
#define ei_b200_init eikws_cxx_b200_init
#define ei_b200_shutdown eikws_cxx_b200_shutdown
#include "edge-impulse-sdk/classifier/ei_run_classifier.h"
#undef run_classifier
#undef run_classifier_continuous
#undef run_classifier_init
#undef run_classifier_batch_i16
#undef run_classifier_batch_f32
#undef ei_b200_init
#undef ei_b200_shutdown

extern "C" {
EI_IMPULSE_ERROR run_classifier(ei::signal_t *signal, ei_impulse_result_t *result, bool debug) { return eikws_cxx_run_classifier(signal, result, debug); }
EI_IMPULSE_ERROR run_classifier_continuous(ei::signal_t *signal, ei_impulse_result_t *result, bool debug) {
    return eikws_cxx_run_classifier_continuous(signal, result, debug);
}
void run_classifier_init(void) { eikws_cxx_run_classifier_init(); }
EI_IMPULSE_ERROR run_classifier_batch_i16(const int16_t *pcm, size_t n_clips, ei_impulse_result_t *results) {
    return eikws_cxx_run_classifier_batch_i16(pcm, n_clips, results);
}
EI_IMPULSE_ERROR run_classifier_batch_f32(const float *samples, size_t n_clips, ei_impulse_result_t *results) {
    return eikws_cxx_run_classifier_batch_f32(samples, n_clips, results);
}
EI_IMPULSE_ERROR ei_b200_init(const int *devices, int n_devices) { return eikws_cxx_b200_init(devices, n_devices); }
void ei_b200_shutdown(void) { eikws_cxx_b200_shutdown(); }
}
