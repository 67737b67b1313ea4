// eikws-b200: operator registrations a generated model binds with
//   registrations[OP_x] = *tflite::ops::micro::Register_x();
// (tflite-model/trained_model_compiled.cpp:415-420).  Replaces the reference's
// edge-impulse-sdk/tensorflow/lite/micro/kernels/micro_ops.h:33-75 at the same path.
// The functions live in libeikws_b200.so (csrc/tflm_capture.cpp): `prepare` records the
// node into the graph that is lowered to a CUDA plan; `invoke` is not a compute path.
#ifndef EIKWS_TFLITE_MICRO_KERNELS_MICRO_OPS_H_
#define EIKWS_TFLITE_MICRO_KERNELS_MICRO_OPS_H_

#include "../../c/common.h"

namespace tflite {
namespace ops {
namespace micro {

TfLiteRegistration *Register_RESHAPE();
TfLiteRegistration *Register_CONV_2D();
TfLiteRegistration *Register_DEPTHWISE_CONV_2D();
TfLiteRegistration *Register_ADD();
TfLiteRegistration *Register_MAX_POOL_2D();
TfLiteRegistration *Register_AVERAGE_POOL_2D();
TfLiteRegistration *Register_FULLY_CONNECTED();
TfLiteRegistration *Register_SOFTMAX();

}  // namespace micro
}  // namespace ops
}  // namespace tflite

#endif  // EIKWS_TFLITE_MICRO_KERNELS_MICRO_OPS_H_
