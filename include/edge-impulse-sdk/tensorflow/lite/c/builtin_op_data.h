/* eikws-b200: builtin-operator parameter blocks that EON-compiled Edge Impulse models
 * aggregate-initialise (`const TfLiteConvParams opdata1 = { kTfLitePaddingSame, 1,1, ... }`,
 * tflite-model/trained_model_compiled.cpp:235-279), so member ORDER is the contract.
 * Replaces the reference's edge-impulse-sdk/tensorflow/lite/c/builtin_op_data.h at the
 * same include path; only the operators this library accelerates are declared.
 */
#ifndef EIKWS_TFLITE_C_BUILTIN_OP_DATA_H_
#define EIKWS_TFLITE_C_BUILTIN_OP_DATA_H_

#include <stdbool.h>
#include <stdint.h>

#include "common.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { kTfLitePaddingUnknown = 0, kTfLitePaddingSame, kTfLitePaddingValid } TfLitePadding;

typedef struct {
    int width;
    int height;
    int width_offset;
    int height_offset;
} TfLitePaddingValues;

typedef enum {
    kTfLiteActNone = 0,
    kTfLiteActRelu,
    kTfLiteActRelu1, /* min(max(-1, x), 1) */
    kTfLiteActRelu6,
    kTfLiteActTanh,
    kTfLiteActSignBit,
    kTfLiteActSigmoid
} TfLiteFusedActivation;
#define kTfLiteActReluN1To1 kTfLiteActRelu1

typedef struct {
    TfLitePadding padding;
    int stride_width;
    int stride_height;
    TfLiteFusedActivation activation;
    int dilation_width_factor;
    int dilation_height_factor;
} TfLiteConvParams;

typedef struct {
    TfLitePadding padding;
    int stride_width;
    int stride_height;
    int depth_multiplier;
    TfLiteFusedActivation activation;
    int dilation_width_factor;
    int dilation_height_factor;
} TfLiteDepthwiseConvParams;

typedef struct {
    TfLitePadding padding;
    int stride_width;
    int stride_height;
    int filter_width;
    int filter_height;
    TfLiteFusedActivation activation;
    struct {
        TfLitePaddingValues padding;
    } computed;
} TfLitePoolParams;

typedef enum {
    kTfLiteFullyConnectedWeightsFormatDefault = 0,
    kTfLiteFullyConnectedWeightsFormatShuffled4x16Int8 = 1
} TfLiteFullyConnectedWeightsFormat;

typedef struct {
    TfLiteFusedActivation activation;
    TfLiteFullyConnectedWeightsFormat weights_format;
    bool keep_num_dims;
    bool asymmetric_quantize_inputs;
} TfLiteFullyConnectedParams;

typedef struct {
    float beta;
} TfLiteSoftmaxParams;

typedef struct {
    TfLiteFusedActivation activation;
} TfLiteAddParams;

#define TFLITE_RESHAPE_PARAMS_MAX_DIMENSION_COUNT 8
typedef struct {
    int shape[TFLITE_RESHAPE_PARAMS_MAX_DIMENSION_COUNT];
    int num_dimensions;
} TfLiteReshapeParams;

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* EIKWS_TFLITE_C_BUILTIN_OP_DATA_H_ */
