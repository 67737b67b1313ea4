// eikws-b200: POD description of one impulse as the CUDA kernels consume it.  Built on the host by
// plan.cpp from a ModelGraph, uploaded once at eikws_create(); all pointers are device pointers.
#pragma once
#include <vector_types.h>  // float2

#include <cstdint>

namespace eikws {

// ---- fixed DSP geometry the fused MFCC kernel is specialised for --------------------------------
// Every model shipped with the reference uses this geometry (model_metadata.h:120-132 in the L476,
// L432 and Arduino exports): 16 kHz, 20 ms frames / 20 ms stride, 256-point FFT, 32 mel filters,
// 13 cepstra, CMVN window 101.  plan.cpp rejects anything else with EIKWS_ERR_UNSUPPORTED.
constexpr int kSamples = 16000;                     // EI_CLASSIFIER_RAW_SAMPLE_COUNT
constexpr int kFrameLen = 320;                      // round(16000 * 0.02)
constexpr int kFrameStride = 320;
constexpr int kFrames = 49;                         // floor((16000 - 320) / 320)
constexpr int kNfft = 256;
constexpr int kNcfft = 128;                         // complex FFT length inside kiss_fftr
constexpr int kBins = 129;
constexpr int kFilters = 32;
constexpr int kCepstra = 13;
constexpr int kFeatures = kFrames * kCepstra;       // 637
constexpr int kWin = 101;
constexpr int kPad = 50;
constexpr int kPadRows = kFrames + 2 * kPad;        // 149
constexpr int kFbMaxTaps = 8;                       // max strictly-positive weights of one mel filter

struct MfccDev {
    float pre_cof;
    float q_scale;  // NN input quantisation (tensor 0): q = (int8)(round(f / q_scale) + q_zp)
    int32_t q_zp;
    float q_inv_scale;  // (float)(1 / q_scale): used only by the certified CMVN shortcut, whose bound covers its rounding
    int32_t input_is_int8;
    const float2 *tw;      // [128]  kiss_fft twiddles  (float)cos/sin(-2*pi*i/128)
    const float2 *stw;     // [64]   kiss_fftr super twiddles
    const float2 *dtw;     // [16]   twiddles of the 16-point FFT inside the 32-point DCT
    const float2 *dstw;    // [8]    its super twiddles
    const float2 *dcs;     // [17]   (cosf, sinf)((float)(i*pi/64))
    const int32_t *fb_first;  // [32] first bin with a strictly positive weight
    const int32_t *fb_count;  // [32] number of consecutive bins with strictly positive weight
    const float *fb_w;        // [32][kFbMaxTaps]
    int32_t fb_max_taps;      // widest filter of this model
    const uint8_t *pad_src;   // [149] source frame of every row of the symmetric-padded matrix
};

// ---- classifier ------------------------------------------------------------------------------------
enum NnOpKind : int32_t {
    kNnConv1d = 1, kNnAddLut = 2, kNnMaxPool = 3, kNnSoftmax = 4,           // int8 graph
    kNnConv1dF32 = 11, kNnAddF32 = 12, kNnMaxPoolF32 = 13, kNnSoftmaxF32 = 14  // float32 twin (BASELINE config 5)
};

struct NnOpDev {
    int32_t kind;
    int32_t in_off, out_off;  // byte offsets of the activation tensors inside the CTA's arena
    // conv1d / fully connected:  in [in_w][in_c] -> out [out_w][out_c], window kw, stride, left pad
    // maxpool:                   in [in_h][in_w][in_c] -> out [out_h][out_w][in_c]
    int32_t in_h, in_w, in_c, out_h, out_w, out_c;
    int32_t kh, kw, stride_h, stride_w, pad_h, pad_w;
    int32_t k_words;        // conv: ceil(kw*in_c / 4)
    int32_t in_zp, out_zp;  // zero points
    int32_t act_min, act_max;
    int32_t n_elems;        // add: number of output elements; softmax: depth; conv: bytes of the padded input row
    int32_t n_const;        // add: number of elements of the broadcast constant operand
    const int32_t *weights;   // conv: [out_c][k_words] packed int8 (zero padded)
    const int32_t *bias;      // conv: [out_c] bias + in_offset * sum(weights)
    const int32_t *mult;      // conv: [out_c] quantised multiplier
    const int32_t *shift;     // conv: [out_c] shift (positive = left)
    const uint8_t *lut;       // add: [n_const][256] output byte for input byte q (index q+128)
    const int32_t *exp_lut;   // softmax: [256] exp_on_negative_values for diff = -i, or -1 if below diff_min
    // float32 ops
    const float *wf;          // conv/fc: [kw*in_c][out_c] (transposed so that lanes over out_c read consecutively)
    const float *bf;          // conv/fc: [out_c] bias or null; add: [n_const] constant operand
    float fmin, fmax;         // fused activation range; softmax: fmin = beta
};

constexpr int kMaxNnOps = 16;

// ---- fused classifier plan (the topology of every model the reference ships) --------------------------------
// [CONV_2D 1xk, stride 1] -> [ADD constant + ReLU] -> [MAX_POOL over the conv width] , twice, then
// FULLY_CONNECTED -> SOFTMAX.  Activations live in shared memory as channel-padded rows [pos][cp] (cp multiple of
// 16 bytes, halo rows and padding lanes pre-filled with the zero point) so that every operand fetch is an aligned
// 128-bit load and the pooled result is written straight into the next stage's padded input.
struct NnFusedStage {
    int32_t in_w, in_c, cp, kw, pad_w, out_c;   // conv geometry; cp = padded channel count (bytes per row)
    int32_t pool, pool_out;                     // pool window (== stride) along the conv width, pooled positions
    int32_t in_zp, conv_out_zp, conv_act_min, conv_act_max, pool_act_min, pool_act_max;
    int32_t in_off, in_rows;                    // padded input buffer: byte offset in the arena, rows = in_w + kw - 1
    int32_t out_off, out_cp, out_row0, out_rows, out_fill;  // consumer layout: row stride, first interior row, total rows, halo byte
    const int32_t *weights;  // [out_c][kw][cp/4] packed int8, zero in the padding lanes
    const int32_t *bias;     // [out_c] bias + in_offset * sum(weights)
    const int32_t *mult;     // [out_c]
    const int32_t *shift;    // [out_c]
    const uint8_t *lut;      // [out_c][256] ADD+activation table over the conv output byte
};

struct NnFusedDev {
    int32_t enabled;
    int32_t shape;  // 0: conv 1x7 + pool 7, second block pooled by the tail (STM32 exports); 1: conv 1x3 + pool 2 (SAME), twice (Arduino zip)
    NnFusedStage st[2];
    // tail: FULLY_CONNECTED [fc_d] -> [fc_o], SOFTMAX over fc_o
    int32_t fc_in_off, fc_d, fc_o, fc_in_zp, fc_out_zp, fc_act_min, fc_act_max, fc_mult, fc_shift;
    int32_t tail_pool, tail_pool_act_min, tail_pool_act_max;  // stage 1 leaves its max-pool (over tail_pool positions) to the tail
    int32_t tail_off;        // scratch for the fc output / softmax output bytes
    // tensor-core lowering of block 1 (tcgen05.mma.kind::i8, see kernels.cu): the stage-0 filter as the A operand of a
    // 128 x N x 32 UMMA, K-major without swizzle: [8 K-chunks of 16 B = one tap][64 rows][16 B]; rows 0..out_c-1 and
    // 32..32+out_c-1 both hold the output channels (two TMEM sub-partitions can then read every channel), tap 7 is zero
    int32_t tc_enabled;
    const int8_t *tc_w;      // [8][64][16]
    const int8_t *fc_w;      // [fc_o][fc_d]
    const int32_t *fc_bias;  // [fc_o] bias + in_offset * sum(weights)
    const int32_t *exp_lut;  // [256]
};

struct NnDev {
    int32_t n_ops;
    int32_t in_off;       // arena offset of the quantised input tensor
    int32_t out_off;      // arena offset of the output tensor
    int32_t n_in;         // == feature count
    int32_t n_out;        // == label count
    int32_t arena_bytes;  // activation arena (two ping-pong buffers)
    int32_t row_bytes;    // scratch for the zero-point padded conv input row
    float out_scale;
    int32_t out_zp;
    int32_t float_mode;   // 1: every activation tensor is float32 (offsets are still bytes)
    NnOpDev ops[kMaxNnOps];
    NnFusedDev fused;
};

struct DevPlan {
    MfccDev mfcc;
    NnDev nn;
};

}  // namespace eikws
