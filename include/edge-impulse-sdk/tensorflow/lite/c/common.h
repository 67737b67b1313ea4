/* eikws-b200: the slice of the TensorFlow-Lite C operator API that an Edge Impulse
 * EON-compiled model (tflite-model/trained_model_compiled.cpp) is written against.
 *
 * Replaces, at the same include path, the reference's
 *   edge-impulse-sdk/tensorflow/lite/c/common.h            (TfLiteContext :519-617,
 *   TfLiteTensor :379-487, TfLiteNode :620-700, TfLiteRegistration :703-760)
 * so that the UNMODIFIED generated model file compiles against this library and its
 * Register_*() operators (micro_ops.h).  Only the members the generated code and the
 * capture layer touch are declared; this library never interprets the graph on the
 * CPU -- `prepare` records it and the batch runs as CUDA kernels (csrc/).
 */
#ifndef EIKWS_TFLITE_C_COMMON_H_
#define EIKWS_TFLITE_C_COMMON_H_

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum TfLiteStatus { kTfLiteOk = 0, kTfLiteError = 1, kTfLiteDelegateError = 2 } TfLiteStatus;

/* element types; numeric values are part of the .tflite schema (model_metadata.h:34-36 relies on 1 and 9) */
typedef enum {
    kTfLiteNoType = 0,
    kTfLiteFloat32 = 1,
    kTfLiteInt32 = 2,
    kTfLiteUInt8 = 3,
    kTfLiteInt64 = 4,
    kTfLiteString = 5,
    kTfLiteBool = 6,
    kTfLiteInt16 = 7,
    kTfLiteComplex64 = 8,
    kTfLiteInt8 = 9,
    kTfLiteFloat16 = 10,
    kTfLiteFloat64 = 11
} TfLiteType;

typedef struct TfLiteIntArray {
    int size;
    int data[];
} TfLiteIntArray;

typedef struct TfLiteFloatArray {
    int size;
    float data[];
} TfLiteFloatArray;

typedef struct TfLiteQuantizationParams {
    float scale;
    int32_t zero_point;
} TfLiteQuantizationParams;

typedef enum TfLiteQuantizationType { kTfLiteNoQuantization = 0, kTfLiteAffineQuantization = 1 } TfLiteQuantizationType;

typedef struct TfLiteQuantization {
    TfLiteQuantizationType type;
    void *params; /* TfLiteAffineQuantization* when type == kTfLiteAffineQuantization */
} TfLiteQuantization;

typedef struct TfLiteAffineQuantization {
    TfLiteFloatArray *scale;
    TfLiteIntArray *zero_point;
    int32_t quantized_dimension;
} TfLiteAffineQuantization;

typedef union TfLitePtrUnion {
    int32_t *i32;
    int64_t *i64;
    float *f;
    char *raw;
    const char *raw_const;
    uint8_t *uint8;
    bool *b;
    int16_t *i16;
    int8_t *int8;
    void *data;
} TfLitePtrUnion;

typedef enum TfLiteAllocationType {
    kTfLiteMemNone = 0,
    kTfLiteMmapRo,
    kTfLiteArenaRw,
    kTfLiteArenaRwPersistent,
    kTfLiteDynamic,
    kTfLitePersistentRo
} TfLiteAllocationType;

typedef struct TfLiteTensor {
    TfLiteQuantization quantization;
    TfLiteQuantizationParams params;
    TfLitePtrUnion data;
    TfLiteIntArray *dims;
    size_t bytes;
    TfLiteType type;
    TfLiteAllocationType allocation_type;
    bool is_variable;
} TfLiteTensor;

typedef struct TfLiteNode {
    TfLiteIntArray *inputs;
    TfLiteIntArray *outputs;
    TfLiteIntArray *intermediates;
    TfLiteIntArray *temporaries;
    void *user_data;
    void *builtin_data;
    const void *custom_initial_data;
    int custom_initial_data_size;
} TfLiteNode;

typedef struct TfLiteContext {
    size_t tensors_size;
    TfLiteTensor *tensors;
    void *impl_;
    TfLiteStatus (*AllocatePersistentBuffer)(struct TfLiteContext *ctx, size_t bytes, void **ptr);
    TfLiteStatus (*RequestScratchBufferInArena)(struct TfLiteContext *ctx, size_t bytes, int *buffer_idx);
    void *(*GetScratchBuffer)(struct TfLiteContext *ctx, int buffer_idx);
    void (*ReportError)(struct TfLiteContext *, const char *msg, ...);
} TfLiteContext;

typedef struct TfLiteRegistration {
    void *(*init)(TfLiteContext *context, const char *buffer, size_t length);
    void (*free)(TfLiteContext *context, void *buffer);
    TfLiteStatus (*prepare)(TfLiteContext *context, TfLiteNode *node);
    TfLiteStatus (*invoke)(TfLiteContext *context, TfLiteNode *node);
    const char *(*profiling_string)(const TfLiteContext *context, const TfLiteNode *node);
    int32_t builtin_code; /* BuiltinOperator value of the op (schema numbering) */
    const char *custom_name;
    int version;
} TfLiteRegistration;

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* EIKWS_TFLITE_C_COMMON_H_ */
