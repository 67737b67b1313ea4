/* eikws-b200 drop-in for edge-impulse-sdk/porting/ei_classifier_porting.h (reference :34-76):
 * the EI_IMPULSE_ERROR codes and the porting hooks an application may override.  Unlike the reference,
 * weak POSIX defaults are provided here, so a host application links without supplying them. */
#ifndef EIKWS_EI_CLASSIFIER_PORTING_H_
#define EIKWS_EI_CLASSIFIER_PORTING_H_

#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <time.h>

#if defined(__cplusplus) && EI_C_LINKAGE == 1
extern "C" {
#endif

typedef enum {
    EI_IMPULSE_OK = 0,
    EI_IMPULSE_ERROR_SHAPES_DONT_MATCH = -1,
    EI_IMPULSE_CANCELED = -2,
    EI_IMPULSE_TFLITE_ERROR = -3,
    EI_IMPULSE_DSP_ERROR = -5,
    EI_IMPULSE_TFLITE_ARENA_ALLOC_FAILED = -6,
    EI_IMPULSE_CUBEAI_ERROR = -7,
    EI_IMPULSE_ALLOC_FAILED = -8
} EI_IMPULSE_ERROR;

EI_IMPULSE_ERROR ei_sleep(int32_t time_ms);
EI_IMPULSE_ERROR ei_run_impulse_check_canceled();
uint64_t ei_read_timer_ms();
uint64_t ei_read_timer_us();
void ei_printf(const char *format, ...);
void ei_printf_float(float f);

#if !defined(EIKWS_NO_DEFAULT_PORTING)
/* weak defaults: an application definition (as the firmware's main.cpp:536-554 provides) wins at link time */
#if defined(__cplusplus) || defined(_POSIX_C_SOURCE) || defined(_GNU_SOURCE)
__attribute__((weak)) EI_IMPULSE_ERROR ei_sleep(int32_t time_ms) {
    struct timespec ts = {time_ms / 1000, (long)(time_ms % 1000) * 1000000L};
    nanosleep(&ts, NULL);
    return EI_IMPULSE_OK;
}
__attribute__((weak)) uint64_t ei_read_timer_us() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (uint64_t)ts.tv_sec * 1000000ull + (uint64_t)ts.tv_nsec / 1000;
}
#else /* a strict ISO C11 translation unit (gcc -std=c11): the POSIX clock calls are hidden, C11's own are used */
__attribute__((weak)) EI_IMPULSE_ERROR ei_sleep(int32_t time_ms) {
    struct timespec t0, t1;
    timespec_get(&t0, TIME_UTC);
    do timespec_get(&t1, TIME_UTC);
    while ((int64_t)(t1.tv_sec - t0.tv_sec) * 1000 + (t1.tv_nsec - t0.tv_nsec) / 1000000 < time_ms);
    return EI_IMPULSE_OK;
}
__attribute__((weak)) uint64_t ei_read_timer_us() {
    struct timespec ts;
    timespec_get(&ts, TIME_UTC);
    return (uint64_t)ts.tv_sec * 1000000ull + (uint64_t)ts.tv_nsec / 1000;
}
#endif
__attribute__((weak)) EI_IMPULSE_ERROR ei_run_impulse_check_canceled() { return EI_IMPULSE_OK; }
__attribute__((weak)) uint64_t ei_read_timer_ms() { return ei_read_timer_us() / 1000; }
__attribute__((weak)) void ei_printf(const char *format, ...) {
    va_list ap;
    va_start(ap, format);
    vprintf(format, ap);
    va_end(ap);
}
__attribute__((weak)) void ei_printf_float(float f) { ei_printf("%f", (double)f); }
#endif

#if defined(__cplusplus) && EI_C_LINKAGE == 1
}
#endif
#endif /* EIKWS_EI_CLASSIFIER_PORTING_H_ */
