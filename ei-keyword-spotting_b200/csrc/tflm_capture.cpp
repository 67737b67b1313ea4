// eikws-b200: the operator-plugin boundary of the reference.
//
// An EON-compiled Edge Impulse model (tflite-model/trained_model_compiled.cpp) binds its ops with
//   registrations[OP_x] = *tflite::ops::micro::Register_x();            (:415-420)
// and drives init(ctx, builtin_data, 0) -> prepare(ctx, node) -> invoke(ctx, node)  (:428-438, :457-465)
// through TfLiteRegistration (edge-impulse-sdk/tensorflow/lite/c/common.h:703-760).  The reference's
// Register_* functions (TFL/micro/kernels/{conv,add,pooling,fully_connected,softmax,reshape}.cc)
// compute on the CPU in `invoke`.  Ours do not: `prepare` RECORDS the node, and after
// trained_model_init() returns the recorded graph is turned into a ModelGraph that plan.cpp
// lowers to CUDA kernels.  `invoke` reports an error: there is no CPU compute path in this library.
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "edge-impulse-sdk/tensorflow/lite/c/builtin_op_data.h"
#include "edge-impulse-sdk/tensorflow/lite/c/common.h"
#include "edge-impulse-sdk/tensorflow/lite/micro/kernels/micro_ops.h"
#include "eikws_b200.h"
#include "model_graph.h"

namespace eikws {
namespace {

struct Recorded {
    uint32_t op;
    TfLiteNode *node;
};
std::mutex g_mu;
std::vector<Recorded> g_nodes;
TfLiteContext *g_ctx = nullptr;

template <uint32_t OP>
void *op_init(TfLiteContext *, const char *buffer, size_t) {
    // user_data is not needed (the builtin params travel in node->builtin_data); hand back the
    // params pointer so generated code that checks for non-null user_data is satisfied.
    return const_cast<char *>(buffer);
}

template <uint32_t OP>
TfLiteStatus op_prepare(TfLiteContext *ctx, TfLiteNode *node) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_ctx = ctx;
    g_nodes.push_back({OP, node});
    return kTfLiteOk;
}

TfLiteStatus op_invoke_unsupported(TfLiteContext *, TfLiteNode *) {
    std::fprintf(stderr,
                 "eikws-b200: trained_model_invoke() has no CPU compute path; call run_classifier()/"
                 "eikws_classify_* (the graph runs as CUDA kernels)\n");
    return kTfLiteError;
}

template <uint32_t OP>
TfLiteRegistration *make_registration() {
    static TfLiteRegistration r = {/*init=*/op_init<OP>,
                                   /*free=*/nullptr,
                                   /*prepare=*/op_prepare<OP>,
                                   /*invoke=*/op_invoke_unsupported,
                                   /*profiling_string=*/nullptr,
                                   /*builtin_code=*/static_cast<int32_t>(OP),
                                   /*custom_name=*/nullptr,
                                   /*version=*/1};
    return &r;
}

void push_params(NodeDesc &n, std::initializer_list<int32_t> v) { n.params.assign(v); }

bool build_graph(ModelGraph &g, std::string &err) {
    if (!g_ctx || g_nodes.empty()) {
        err = "no graph recorded: trained_model_init() was not run against this library's Register_* ops";
        return false;
    }
    TfLiteContext *ctx = g_ctx;
    g.tensors.clear();
    g.nodes.clear();
    for (size_t i = 0; i < ctx->tensors_size; i++) {
        const TfLiteTensor &t = ctx->tensors[i];
        TensorDesc d;
        d.type = static_cast<uint32_t>(t.type);
        d.is_const = t.allocation_type == kTfLiteMmapRo;
        if (t.dims)
            for (int k = 0; k < t.dims->size; k++) d.dims.push_back(t.dims->data[k]);
        d.bytes = static_cast<uint32_t>(t.bytes);
        if (t.quantization.type == kTfLiteAffineQuantization && t.quantization.params) {
            const TfLiteAffineQuantization *q = static_cast<const TfLiteAffineQuantization *>(t.quantization.params);
            int n = q->scale ? q->scale->size : 0;
            for (int k = 0; k < n; k++) d.scales.push_back(q->scale->data[k]);
            for (int k = 0; k < n; k++)
                d.zero_points.push_back(q->zero_point && k < q->zero_point->size ? q->zero_point->data[k] : 0);
            d.quantized_dimension = q->quantized_dimension;
        }
        if (d.is_const) {
            if (!t.data.data) {
                err = "constant tensor without data";
                return false;
            }
            d.data.assign(static_cast<const uint8_t *>(t.data.data), static_cast<const uint8_t *>(t.data.data) + t.bytes);
        }
        g.tensors.push_back(std::move(d));
    }
    for (const Recorded &r : g_nodes) {
        NodeDesc n;
        n.op = r.op;
        for (int k = 0; k < r.node->inputs->size; k++) n.inputs.push_back(r.node->inputs->data[k]);
        for (int k = 0; k < r.node->outputs->size; k++) n.outputs.push_back(r.node->outputs->data[k]);
        const void *bd = r.node->builtin_data;
        switch (r.op) {
            case kOpConv2D: {
                const TfLiteConvParams *p = static_cast<const TfLiteConvParams *>(bd);
                push_params(n, {p->padding, p->stride_width, p->stride_height, p->activation, p->dilation_width_factor,
                                p->dilation_height_factor});
                break;
            }
            case kOpDepthwiseConv2D: {
                const TfLiteDepthwiseConvParams *p = static_cast<const TfLiteDepthwiseConvParams *>(bd);
                push_params(n, {p->padding, p->stride_width, p->stride_height, p->depth_multiplier, p->activation,
                                p->dilation_width_factor, p->dilation_height_factor});
                break;
            }
            case kOpAdd: {
                const TfLiteAddParams *p = static_cast<const TfLiteAddParams *>(bd);
                push_params(n, {p->activation});
                break;
            }
            case kOpMaxPool2D:
            case kOpAveragePool2D: {
                const TfLitePoolParams *p = static_cast<const TfLitePoolParams *>(bd);
                push_params(n, {p->padding, p->stride_width, p->stride_height, p->filter_width, p->filter_height, p->activation});
                break;
            }
            case kOpFullyConnected: {
                const TfLiteFullyConnectedParams *p = static_cast<const TfLiteFullyConnectedParams *>(bd);
                push_params(n, {p->activation});
                break;
            }
            case kOpSoftmax: {
                const TfLiteSoftmaxParams *p = static_cast<const TfLiteSoftmaxParams *>(bd);
                int32_t bits;
                std::memcpy(&bits, &p->beta, 4);
                push_params(n, {bits});
                break;
            }
            default: break;
        }
        g.nodes.push_back(std::move(n));
    }
    return true;
}

void *aligned_arena_alloc(size_t align, size_t size) {
    void *p = nullptr;
    if (align < sizeof(void *)) align = sizeof(void *);
    if (posix_memalign(&p, align, size ? size : align)) return nullptr;
    return p;
}

}  // namespace

// Run the generated model's own init against our Register_* ops and lift the result into a graph.
bool capture_compiled_model(const eikws_compiled_model_t *cm, ModelGraph &g, std::string &err) {
    if (!cm || !cm->init || !cm->input || !cm->output) {
        err = "eikws_compiled_model_t is incomplete";
        return false;
    }
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_nodes.clear();
        g_ctx = nullptr;
    }
    if (cm->init(aligned_arena_alloc) != 0) {
        err = "trained_model_init() failed";
        return false;
    }
    bool ok;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        ok = build_graph(g, err);
        if (ok) {
            TfLiteTensor *base = g_ctx->tensors;
            g.input = static_cast<uint32_t>(static_cast<TfLiteTensor *>(cm->input(0)) - base);
            g.output = static_cast<uint32_t>(static_cast<TfLiteTensor *>(cm->output(0)) - base);
        }
        g_nodes.clear();
        g_ctx = nullptr;
    }
    if (cm->reset) cm->reset(free);
    if (!ok) return false;
    g.raw_sample_count = cm->raw_sample_count;
    g.nn_input_frame_size = cm->nn_input_frame_size;
    g.labels.clear();
    for (uint32_t i = 0; i < cm->label_count; i++) g.labels.emplace_back(cm->labels[i] ? cm->labels[i] : "");
    g.mfcc.sample_rate = cm->frequency;
    g.mfcc.num_cepstral = cm->mfcc_num_cepstral;
    g.mfcc.frame_length = cm->mfcc_frame_length;
    g.mfcc.frame_stride = cm->mfcc_frame_stride;
    g.mfcc.num_filters = cm->mfcc_num_filters;
    g.mfcc.fft_length = cm->mfcc_fft_length;
    g.mfcc.win_size = cm->mfcc_win_size;
    g.mfcc.low_frequency = cm->mfcc_low_frequency;
    g.mfcc.high_frequency = cm->mfcc_high_frequency;
    g.mfcc.pre_cof = cm->mfcc_pre_cof;
    g.mfcc.pre_shift = cm->mfcc_pre_shift;
    return true;
}

}  // namespace eikws

namespace tflite {
namespace ops {
namespace micro {
TfLiteRegistration *Register_RESHAPE() { return eikws::make_registration<eikws::kOpReshape>(); }
TfLiteRegistration *Register_CONV_2D() { return eikws::make_registration<eikws::kOpConv2D>(); }
TfLiteRegistration *Register_DEPTHWISE_CONV_2D() { return eikws::make_registration<eikws::kOpDepthwiseConv2D>(); }
TfLiteRegistration *Register_ADD() { return eikws::make_registration<eikws::kOpAdd>(); }
TfLiteRegistration *Register_MAX_POOL_2D() { return eikws::make_registration<eikws::kOpMaxPool2D>(); }
TfLiteRegistration *Register_AVERAGE_POOL_2D() { return eikws::make_registration<eikws::kOpAveragePool2D>(); }
TfLiteRegistration *Register_FULLY_CONNECTED() { return eikws::make_registration<eikws::kOpFullyConnected>(); }
TfLiteRegistration *Register_SOFTMAX() { return eikws::make_registration<eikws::kOpSoftmax>(); }
}  // namespace micro
}  // namespace ops
}  // namespace tflite
