// eikws-b200: host-side description of one Edge Impulse impulse (MFCC DSP block config +
// the quantised TFLite graph), as captured from the generated model files
//   model-parameters/model_metadata.h        (EI_CLASSIFIER_* macros :38-68, MFCC config :120-132)
//   tflite-model/trained_model_compiled.cpp  (tensor table :280-312, node table :312-328)
// and its on-disk container ("EIKWSMDL" v1, layout documented in include/eikws_model_format.md).
// The container holds only RAW model data (tensor bytes, scales, zero points, op parameters);
// every derived quantity (fixed-point multipliers, tables) is computed by plan.cpp.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace eikws {

// TFLite BuiltinOperator numbering (schema) for the ops this library lowers.
enum OpCode : uint32_t {
    kOpAdd = 0,
    kOpAveragePool2D = 1,
    kOpConv2D = 3,
    kOpDepthwiseConv2D = 4,
    kOpFullyConnected = 9,
    kOpMaxPool2D = 17,
    kOpReshape = 22,
    kOpSoftmax = 25,
};

// TfLiteType numbering (edge-impulse-sdk/tensorflow/lite/c/common.h)
enum ElemType : uint32_t { kF32 = 1, kI32 = 2, kU8 = 3, kI8 = 9 };

struct MfccConfig {  // ei_dsp_config_mfcc_t (model_metadata.h:92-104) + EI_CLASSIFIER_FREQUENCY (:48)
    int32_t sample_rate = 16000;
    int32_t num_cepstral = 13;
    float frame_length = 0.02f;
    float frame_stride = 0.02f;
    int32_t num_filters = 32;
    int32_t fft_length = 256;
    int32_t win_size = 101;
    int32_t low_frequency = 300;
    int32_t high_frequency = 0;
    float pre_cof = 0.98f;
    int32_t pre_shift = 1;
};

struct TensorDesc {
    uint32_t type = 0;
    bool is_const = false;
    std::vector<int32_t> dims;
    uint32_t bytes = 0;
    std::vector<float> scales;  // per-tensor (1) or per-channel (n) affine quantisation; empty = none
    std::vector<int32_t> zero_points;
    int32_t quantized_dimension = 0;
    std::vector<uint8_t> data;  // constants only

    float scale() const { return scales.empty() ? 0.f : scales[0]; }
    int32_t zero_point() const { return zero_points.empty() ? 0 : zero_points[0]; }
};

// params, in order, per op:
//   CONV_2D            padding, stride_w, stride_h, activation, dilation_w, dilation_h
//   DEPTHWISE_CONV_2D  padding, stride_w, stride_h, depth_multiplier, activation, dilation_w, dilation_h
//   ADD                activation
//   MAX/AVERAGE_POOL   padding, stride_w, stride_h, filter_w, filter_h, activation
//   FULLY_CONNECTED    activation
//   SOFTMAX            beta (IEEE-754 bits)
//   RESHAPE            (none)
struct NodeDesc {
    uint32_t op = 0;
    std::vector<int32_t> inputs, outputs, params;
};

struct ModelGraph {
    std::vector<TensorDesc> tensors;
    std::vector<NodeDesc> nodes;
    uint32_t input = 0, output = 0;
    std::vector<std::string> labels;
    uint32_t raw_sample_count = 16000;  // EI_CLASSIFIER_RAW_SAMPLE_COUNT
    uint32_t nn_input_frame_size = 0;   // EI_CLASSIFIER_NN_INPUT_FRAME_SIZE
    MfccConfig mfcc;
};

void serialize_model(const ModelGraph &g, std::vector<uint8_t> &out);
// returns false and fills err on malformed input
bool parse_model(const void *blob, size_t bytes, ModelGraph &g, std::string &err);
// structural checks on an untrusted graph (run by parse_model): indices, shapes vs byte sizes, per-operator arity, divisors
bool validate_model(const ModelGraph &g, std::string &err);

}  // namespace eikws
