// eikws-b200: host-side planning.  Everything the reference recomputes on EVERY run_classifier call
// (mel filterbank feature.hpp:243-253, FFT twiddles kiss_fft.cpp:351-357 / kiss_fftr.cpp:51-57, DCT
// cos/sin fast-dct-fft.cpp:71-74, per-channel requantisation multipliers conv.cc:588-637) is computed
// once here with the same host arithmetic (same libm calls, same float/double mix) and uploaded.
#include "plan.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <deque>

#include "eikws_b200.h"
#include "kernels.h"
#include "quant_math.h"

namespace eikws {
namespace {

// libm entry points reached through volatile function pointers.  The tables below must hold exactly what the
// reference computes AT RUN TIME with the C library (kiss_fft_alloc, fast-dct-fft.cpp:71-74, functions.hpp:52-54);
// with compile-time-constant geometry GCC would otherwise fold cosf/sinf through MPFR (correctly rounded), which
// differs from glibc's results by one ulp for some arguments (seen: cosf/sinf(i*pi/64), i = 6 and 11).
float (*volatile rt_cosf)(float) = cosf;
float (*volatile rt_sinf)(float) = sinf;
float (*volatile rt_expf)(float) = expf;
double (*volatile rt_cos)(double) = cos;
double (*volatile rt_sin)(double) = sin;

struct Fixup {
    size_t field_off;  // byte offset of the pointer member inside DevPlan
    size_t blob_off;
};
struct Builder {
    HostPlan &hp;
    std::vector<Fixup> fixups;
    size_t push(const void *data, size_t bytes) {
        size_t off = (hp.blob.size() + 15) & ~size_t(15);
        hp.blob.resize(off + bytes);
        if (bytes) std::memcpy(hp.blob.data() + off, data, bytes);
        return off;
    }
    template <typename P>
    void bind(P *&field, size_t blob_off) {
        fixups.push_back({static_cast<size_t>(reinterpret_cast<uint8_t *>(&field) - reinterpret_cast<uint8_t *>(&hp.dev)), blob_off});
        field = nullptr;
    }
};

// ---- DSP tables ----------------------------------------------------------------------------------------
// numpy::log (numpy.hpp:1350-1371), host copy used only for the mel scale (functions.hpp:42-44)
float fastlog_host(float a) {
    int32_t g;
    std::memcpy(&g, &a, 4);
    int32_t e = static_cast<int32_t>((static_cast<uint32_t>(g) - 0x3f2aaaabu) & 0xff800000u);
    g = static_cast<int32_t>(static_cast<uint32_t>(g) - static_cast<uint32_t>(e));
    float m;
    std::memcpy(&m, &g, 4);
    float i = static_cast<float>(e) * 1.19209290e-7f;
    float f = m - 1.0f;
    float s = f * f;
    float r = std::fmaf(0.230836749f, f, -0.279208571f);
    float t = std::fmaf(0.331826031f, f, -0.498910338f);
    r = std::fmaf(r, s, t);
    r = std::fmaf(r, s, f);
    r = std::fmaf(i, 0.693147182f, r);
    return r;
}

void linspace(float start, float stop, uint32_t number, float *out) {  // numpy.hpp:1257-1280
    if (number == 1) {
        out[0] = start;
        return;
    }
    float step = (stop - start) / static_cast<float>(number - 1);
    for (uint32_t ix = 0; ix < number - 1; ix++) out[ix] = start + static_cast<float>(ix) * step;
    out[number - 1] = stop;
}

// feature::filterbanks (feature.hpp:54-171) -> dense [bins][filters]
void mel_filterbank(std::vector<float> &fb, int filters, int bins, uint32_t fs, uint32_t low, uint32_t high) {
    const int n = filters + 2;
    std::vector<float> mels(n), hz(n);
    std::vector<int> bin(n);
    fb.assign(static_cast<size_t>(bins) * filters, 0.0f);
    auto to_mel = [](float f) { return static_cast<float>(1127.0 * static_cast<double>(fastlog_host(1 + f / 700.0f))); };
    linspace(to_mel(static_cast<float>(low)), to_mel(static_cast<float>(high)), n, mels.data());
    for (int i = 0; i < n; i++) {
        hz[i] = 700.0f * (rt_expf(mels[i] / 1127.0f) - 1.0f);
        if (hz[i] < static_cast<float>(low)) hz[i] = static_cast<float>(low);
        if (hz[i] > static_cast<float>(high)) hz[i] = static_cast<float>(high);
        if (i == n - 1) hz[i] = static_cast<float>(static_cast<double>(hz[i]) - 0.001);
    }
    for (int i = 0; i < n; i++) bin[i] = static_cast<int>(std::floor(static_cast<float>(bins + 1) * hz[i] / static_cast<float>(fs)));
    for (int i = 0; i < filters; i++) {
        const int left = bin[i], middle = bin[i + 1], right = bin[i + 2];
        const int zn = right - left + 1;
        if (zn <= 0) continue;
        std::vector<float> z(zn);
        linspace(static_cast<float>(left), static_cast<float>(right), zn, z.data());
        for (int zx = 0; zx < zn; zx++) {
            float x = z[zx], o = 0.0f;  // functions::triangle (functions.hpp:90-104)
            if (x > left && x <= middle) o = (x - left) / (middle - left);
            if (x < right && middle <= x) o = (right - x) / (right - middle);
            if (left + zx >= 0 && left + zx < bins) fb[static_cast<size_t>(left + zx) * filters + i] = o;  // (always true for a validated band)
        }
    }
}

void fft_twiddles(int nfft, std::vector<float2> &tw) {  // kiss_fft_alloc (kiss_fft.cpp:351-357)
    tw.resize(nfft);
    for (int i = 0; i < nfft; i++) {
        const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
        double phase = -2 * pi * i / nfft;
        tw[i] = make_float2(static_cast<float>(rt_cos(phase)), static_cast<float>(rt_sin(phase)));
    }
}
void super_twiddles(int ncfft, std::vector<float2> &stw) {  // kiss_fftr_alloc (kiss_fftr.cpp:51-57)
    stw.resize(ncfft / 2);
    for (int i = 0; i < ncfft / 2; i++) {
        double phase = -3.14159265358979323846264338327 * (static_cast<double>(i + 1) / ncfft + .5);
        stw[i] = make_float2(static_cast<float>(rt_cos(phase)), static_cast<float>(rt_sin(phase)));
    }
}

void pad_rows(int rows, int before, int after, std::vector<uint8_t> &src) {  // numpy::pad_1d_symmetric (numpy.hpp:479-541)
    src.assign(rows + before + after, 0);
    int idx = 0;
    bool up = true;
    for (int ix = before - 1; ix >= 0; ix--) {
        src[ix] = static_cast<uint8_t>(idx);
        if (idx == 0 && !up) up = true;
        else if (idx == rows - 1 && up) up = false;
        else if (up) idx++;
        else idx--;
    }
    for (int ix = 0; ix < rows; ix++) src[before + ix] = static_cast<uint8_t>(ix);
    idx = rows - 1;
    up = false;
    for (int ix = 0; ix < after; ix++) {
        src[ix + before + rows] = static_cast<uint8_t>(idx);
        if (idx == 0 && !up) up = true;
        else if (idx == rows - 1 && up) up = false;
        else if (up) idx++;
        else idx--;
    }
}

// ---- classifier lowering ------------------------------------------------------------------------------
void dims4(const TensorDesc &t, int d[4]) {  // RuntimeShape::ExtendedShape(4, ...)
    for (int i = 0; i < 4; i++) d[i] = 1;
    int n = static_cast<int>(t.dims.size());
    for (int i = 0; i < n && i < 4; i++) d[4 - n + i] = t.dims[i];
}

// CalculateActivationRangeQuantized (kernel_util_lite.cc:174-220), int8 only
void activation_range_i8(int act, const TensorDesc &out, int32_t *lo, int32_t *hi) {
    const float scale = out.scale();
    const int32_t zp = out.zero_point();
    auto quantize = [&](float f) { return zp + static_cast<int32_t>(std::round(f / scale)); };
    *lo = -128;
    *hi = 127;
    if (act == 1) {  // kTfLiteActRelu
        *lo = std::max(-128, quantize(0.0f));
    } else if (act == 3) {  // kTfLiteActRelu6
        *lo = std::max(-128, quantize(0.0f));
        *hi = std::min(127, quantize(6.0f));
    } else if (act == 2) {  // kTfLiteActRelu1
        *lo = std::max(-128, quantize(-1.0f));
        *hi = std::min(127, quantize(1.0f));
    }
}

// acc = sum w*(x + in_offset) = sum w*x + in_offset*sum w with in_offset = -in_zp: the constant part is folded into the bias.
// int32 accumulators wrap like the reference's on x86; done in 64 bits here so that a hostile bias cannot trigger signed overflow
int32_t fold_bias(int32_t bias, int32_t in_zp, int32_t wsum) {
    return static_cast<int32_t>(static_cast<uint32_t>(static_cast<int64_t>(bias) - static_cast<int64_t>(in_zp) * static_cast<int64_t>(wsum)));
}

int same_or_valid_pad(int padding, int stride, int dilation, int in, int filt) {  // padding.h:32-57
    const int eff = (filt - 1) * dilation + 1;
    int out = 0;
    if (padding == 1) out = (in + stride - 1) / stride;
    else if (padding == 2) out = (in + stride - eff) / stride;
    int total = (out - 1) * stride + eff - in;
    if (total < 0) total = 0;
    return total / 2;
}


// ---- float32 graph (BASELINE config 5): same operator set, TFLite float reference semantics ----------------------
// CalculateActivationRange (kernel_util.h): None [lowest,max], Relu [0,max], Relu6 [0,6], ReluN1To1 [-1,1]
void activation_range_f32(int act, float *lo, float *hi) {
    *lo = -FLT_MAX;
    *hi = FLT_MAX;
    if (act == 1) *lo = 0.0f;
    else if (act == 3) { *lo = 0.0f; *hi = 6.0f; }
    else if (act == 2) { *lo = -1.0f; *hi = 1.0f; }
}

int build_float_nn(const ModelGraph &g, HostPlan &hp, Builder &b, std::string &err) {
    NnDev &nn = hp.dev.nn;
    MfccDev &mf = hp.dev.mfcc;
    const TensorDesc &tin = g.tensors[g.input], &tout = g.tensors[g.output];
    if (tin.bytes != kFeatures * 4 || static_cast<size_t>(tout.bytes) != g.labels.size() * 4) {
        err = "input/output tensor size does not match feature/label count";
        return EIKWS_ERR_SHAPES_DONT_MATCH;
    }
    mf.q_scale = 1.0f;
    mf.q_inv_scale = 1.0f;
    mf.q_zp = 0;
    mf.input_is_int8 = 0;
    uint32_t max_bytes = 0;
    for (const TensorDesc &t : g.tensors)
        if (!t.is_const && t.bytes > max_bytes) max_bytes = t.bytes;
    const int buf_bytes = static_cast<int>((max_bytes + 15) & ~15u);
    std::vector<int> off(g.tensors.size(), -1);
    off[g.input] = 0;
    nn.n_ops = 0;
    nn.float_mode = 1;
    auto other = [&](int o) { return o == 0 ? buf_bytes : 0; };
    for (const NodeDesc &n : g.nodes) {
        if (n.op == kOpReshape) {
            if (n.inputs.empty() || off[n.inputs[0]] < 0) {
                err = "reshape of an unplaced tensor";
                return EIKWS_ERR_UNSUPPORTED;
            }
            off[n.outputs[0]] = off[n.inputs[0]];
            continue;
        }
        if (nn.n_ops >= kMaxNnOps) {
            err = "graph has more compute nodes than kMaxNnOps";
            return EIKWS_ERR_UNSUPPORTED;
        }
        NnOpDev &op = nn.ops[nn.n_ops];
        std::memset(&op, 0, sizeof(op));
        const TensorDesc &out = g.tensors[n.outputs[0]];
        if (out.type != kF32) {
            err = "mixed-type graph";
            return EIKWS_ERR_UNSUPPORTED;
        }
        if (n.op == kOpConv2D || n.op == kOpFullyConnected) {
            const TensorDesc &in = g.tensors[n.inputs[0]], &flt = g.tensors[n.inputs[1]];
            const bool has_bias = n.inputs.size() > 2 && n.inputs[2] >= 0;
            if (off[n.inputs[0]] < 0 || !flt.is_const || flt.type != kF32 || in.type != kF32 ||
                (has_bias && (!g.tensors[n.inputs[2]].is_const || g.tensors[n.inputs[2]].type != kF32))) {
                err = "conv/fc: unsupported operand types";
                return EIKWS_ERR_UNSUPPORTED;
            }
            op.kind = kNnConv1dF32;
            if (n.op == kOpConv2D) {
                int di[4], df[4], dq[4];
                dims4(in, di);
                dims4(flt, df);
                dims4(out, dq);
                if (di[0] != 1 || di[1] != 1 || df[1] != 1 || dq[1] != 1 || n.params[4] != 1 || df[3] != di[3] || dq[3] != df[0]) {
                    err = "conv: only 1xk convolutions over [1,1,W,C] inputs with dilation 1 are implemented";
                    return EIKWS_ERR_UNSUPPORTED;
                }
                op.in_w = di[2];
                op.in_c = di[3];
                op.out_w = dq[2];
                op.out_c = dq[3];
                op.kw = df[2];
                op.stride_w = n.params[1];
                op.pad_w = same_or_valid_pad(n.params[0], n.params[1], 1, di[2], df[2]);
                activation_range_f32(n.params[3], &op.fmin, &op.fmax);
            } else {
                const int depth = flt.dims.back(), outc = out.dims.back();
                if (static_cast<int>(in.bytes) != depth * 4 || static_cast<int>(flt.bytes) != depth * outc * 4 || static_cast<int>(out.bytes) != outc * 4) {
                    err = "fully_connected: only batch 1 with a dense [out,in] filter is implemented";
                    return EIKWS_ERR_UNSUPPORTED;
                }
                op.in_w = 1;
                op.in_c = depth;
                op.out_w = 1;
                op.out_c = outc;
                op.kw = 1;
                op.stride_w = 1;
                activation_range_f32(n.params[0], &op.fmin, &op.fmax);
            }
            const int K = op.kw * op.in_c;
            const float *w = reinterpret_cast<const float *>(flt.data.data());
            std::vector<float> wt(static_cast<size_t>(K) * op.out_c);
            for (int oc = 0; oc < op.out_c; oc++)
                for (int k = 0; k < K; k++) wt[static_cast<size_t>(k) * op.out_c + oc] = w[oc * K + k];
            b.bind(op.wf, b.push(wt.data(), wt.size() * 4));
            if (has_bias) b.bind(op.bf, b.push(g.tensors[n.inputs[2]].data.data(), g.tensors[n.inputs[2]].data.size()));
            op.in_off = off[n.inputs[0]];
        } else if (n.op == kOpAdd) {
            int ia = n.inputs[0], ic = n.inputs[1];
            if (g.tensors[ia].is_const && !g.tensors[ic].is_const) std::swap(ia, ic);  // float addition commutes exactly
            const TensorDesc &a = g.tensors[ia], &cst = g.tensors[ic];
            const size_t nc = cst.dims.size(), na = a.dims.size();
            bool ok = !a.is_const && cst.is_const && off[ia] >= 0 && a.type == kF32 && cst.type == kF32 && nc <= na && out.bytes == a.bytes;
            for (size_t i = 0; ok && i < nc; i++) ok = cst.dims[nc - 1 - i] == a.dims[na - 1 - i];
            if (!ok || cst.bytes == 0 || a.bytes % cst.bytes) {
                err = "add: only activation + trailing-dims constant is implemented";
                return EIKWS_ERR_UNSUPPORTED;
            }
            op.kind = kNnAddF32;
            op.n_elems = static_cast<int32_t>(out.bytes / 4);
            op.n_const = static_cast<int32_t>(cst.bytes / 4);
            activation_range_f32(n.params[0], &op.fmin, &op.fmax);
            b.bind(op.bf, b.push(cst.data.data(), cst.data.size()));
            op.in_off = off[ia];
        } else if (n.op == kOpMaxPool2D) {
            const TensorDesc &in = g.tensors[n.inputs[0]];
            int di[4], dq[4];
            dims4(in, di);
            dims4(out, dq);
            if (off[n.inputs[0]] < 0 || in.type != kF32 || di[0] != 1 || di[3] != dq[3]) {
                err = "max_pool: unsupported operand";
                return EIKWS_ERR_UNSUPPORTED;
            }
            op.kind = kNnMaxPoolF32;
            op.in_h = di[1];
            op.in_w = di[2];
            op.in_c = di[3];
            op.out_h = dq[1];
            op.out_w = dq[2];
            op.out_c = dq[3];
            op.stride_w = n.params[1];
            op.stride_h = n.params[2];
            op.kw = n.params[3];
            op.kh = n.params[4];
            op.pad_h = same_or_valid_pad(n.params[0], op.stride_h, 1, op.in_h, op.kh);
            op.pad_w = same_or_valid_pad(n.params[0], op.stride_w, 1, op.in_w, op.kw);
            activation_range_f32(n.params[5], &op.fmin, &op.fmax);
            op.in_off = off[n.inputs[0]];
        } else if (n.op == kOpSoftmax) {
            const TensorDesc &in = g.tensors[n.inputs[0]];
            int outer = 1;
            for (size_t i = 0; i + 1 < in.dims.size(); i++) outer *= in.dims[i];
            if (off[n.inputs[0]] < 0 || in.type != kF32 || outer != 1) {
                err = "softmax: only a single float row is implemented";
                return EIKWS_ERR_UNSUPPORTED;
            }
            op.kind = kNnSoftmaxF32;
            op.n_elems = in.dims.back();
            std::memcpy(&op.fmin, &n.params[0], 4);  // beta
            op.in_off = off[n.inputs[0]];
        } else {
            err = "operator " + std::to_string(n.op) + " is not implemented for float32 graphs";
            return EIKWS_ERR_UNSUPPORTED;
        }
        op.out_off = other(op.in_off);
        off[n.outputs[0]] = op.out_off;
        nn.n_ops++;
    }
    if (off[g.output] < 0) {
        err = "output tensor is never produced";
        return EIKWS_ERR_UNSUPPORTED;
    }
    nn.in_off = 0;
    nn.out_off = off[g.output];
    nn.n_in = kFeatures;
    nn.n_out = static_cast<int32_t>(g.labels.size());
    nn.arena_bytes = 2 * buf_bytes;
    nn.out_scale = 1.0f;
    nn.out_zp = 0;
    hp.nn_smem_bytes = nn.arena_bytes;
    if (hp.nn_smem_bytes > nn_smem_capacity_float_graph()) {
        err = "float activations do not fit the fused kernel's shared-memory overlay";
        return EIKWS_ERR_UNSUPPORTED;
    }
    return EIKWS_OK;
}

}  // namespace

// QuantizeMultiplier (quantization_util.cc:53-91)
void quantize_multiplier(double dm, int32_t *qm_out, int *shift) {
    if (dm == 0.) {
        *qm_out = 0;
        *shift = 0;
        return;
    }
    const double q = std::frexp(dm, shift);
    int64_t q_fixed = static_cast<int64_t>(std::round(q * static_cast<double>(1ll << 31)));
    if (q_fixed == (1ll << 31)) {
        q_fixed /= 2;
        ++*shift;
    }
    if (*shift < -31) {
        *shift = 0;
        q_fixed = 0;
    }
    *qm_out = static_cast<int32_t>(q_fixed);
}

int build_host_plan(const ModelGraph &g, HostPlan &hp, std::string &err) {
    hp = HostPlan();
    Builder b{hp, {}};
    const MfccConfig &c = g.mfcc;

    // ---------- DSP geometry: only the family the fused kernel is specialised for ----------
    const int frame_len = static_cast<int>(std::round(static_cast<float>(c.sample_rate) * c.frame_length));
    const float stride_f = std::round(static_cast<float>(c.sample_rate) * c.frame_stride);
    const int frames = static_cast<int>(std::floor(static_cast<float>(static_cast<int>(g.raw_sample_count) - frame_len) / stride_f));
    if (g.raw_sample_count != kSamples || frame_len != kFrameLen || static_cast<int>(stride_f) != kFrameStride || frames != kFrames ||
        c.fft_length != kNfft || c.num_filters != kFilters || c.num_cepstral != kCepstra || c.win_size != kWin || c.pre_shift != 1 ||
        g.nn_input_frame_size != kFeatures) {
        err = "unsupported DSP geometry (need 16000 samples, 20 ms/20 ms frames at 16 kHz, fft 256, 32 filters, 13 cepstra, win 101, shift 1)";
        return EIKWS_ERR_UNSUPPORTED;
    }
    const uint32_t high = c.high_frequency == 0 ? static_cast<uint32_t>(c.sample_rate) / 2 : static_cast<uint32_t>(c.high_frequency);
    mel_filterbank(hp.filterbank, kFilters, kBins, static_cast<uint32_t>(c.sample_rate), static_cast<uint32_t>(c.low_frequency), high);
    std::vector<int32_t> fb_first(kFilters, 0), fb_count(kFilters, 0);
    std::vector<float> fb_w(kFilters * kFbMaxTaps, 0.0f);
    for (int j = 0; j < kFilters; j++) {
        int first = -1, last = -1;
        for (int k = 0; k < kBins; k++) {
            float w = hp.filterbank[static_cast<size_t>(k) * kFilters + j];
            if (w < 0.0f || w != w) {
                err = "mel filterbank produced a negative/NaN weight";
                return EIKWS_ERR_UNSUPPORTED;
            }
            if (w > 0.0f) {
                if (first < 0) first = k;
                last = k;
            }
        }
        if (first >= 0) {
            // zero weights inside [first,last] are kept as explicit taps (adding p*0 is exact), but a triangle has none
            if (last - first + 1 > kFbMaxTaps) {
                err = "mel filter wider than kFbMaxTaps";
                return EIKWS_ERR_UNSUPPORTED;
            }
            fb_first[j] = first;
            fb_count[j] = last - first + 1;
            for (int k = first; k <= last; k++) fb_w[j * kFbMaxTaps + (k - first)] = hp.filterbank[static_cast<size_t>(k) * kFilters + j];
        }
    }
    std::vector<float2> tw, stw, dtw, dstw, dcs(kFilters / 2 + 1);
    fft_twiddles(kNcfft, tw);
    super_twiddles(kNcfft, stw);
    fft_twiddles(kFilters / 2, dtw);
    super_twiddles(kFilters / 2, dstw);
    for (int i = 0; i < kFilters / 2 + 1; i++) {
        // fast-dct-fft.cpp:71-74: float temp = i * M_PI / (len * 2); cos(temp)/sin(temp) resolve to the float overloads
        float temp = static_cast<float>(static_cast<double>(i) * 3.14159265358979323846264338327950288 / static_cast<double>(kFilters * 2));
        dcs[i] = make_float2(rt_cosf(temp), rt_sinf(temp));
    }
    std::vector<uint8_t> psrc;
    pad_rows(kFrames, kPad, kPad, psrc);

    MfccDev &mf = hp.dev.mfcc;
    mf.pre_cof = c.pre_cof;
    mf.fb_max_taps = *std::max_element(fb_count.begin(), fb_count.end());
    b.bind(mf.tw, b.push(tw.data(), tw.size() * sizeof(float2)));
    b.bind(mf.stw, b.push(stw.data(), stw.size() * sizeof(float2)));
    b.bind(mf.dtw, b.push(dtw.data(), dtw.size() * sizeof(float2)));
    b.bind(mf.dstw, b.push(dstw.data(), dstw.size() * sizeof(float2)));
    b.bind(mf.dcs, b.push(dcs.data(), dcs.size() * sizeof(float2)));
    b.bind(mf.fb_first, b.push(fb_first.data(), fb_first.size() * 4));
    b.bind(mf.fb_count, b.push(fb_count.data(), fb_count.size() * 4));
    b.bind(mf.fb_w, b.push(fb_w.data(), fb_w.size() * 4));
    b.bind(mf.pad_src, b.push(psrc.data(), psrc.size()));

    // ---------- classifier ----------
    NnDev &nn = hp.dev.nn;
    const TensorDesc &tin = g.tensors[g.input], &tout = g.tensors[g.output];
    if (tin.type == kF32 && tout.type == kF32) {
        int rc = build_float_nn(g, hp, b, err);
        if (rc != EIKWS_OK) return rc;
        for (const Fixup &f : b.fixups) hp.fixups.emplace_back(f.field_off, f.blob_off);
        return EIKWS_OK;
    }
    if (tin.type != kI8 || tout.type != kI8) {
        err = "unsupported input/output tensor type (int8 and float32 graphs are implemented)";
        return EIKWS_ERR_UNSUPPORTED;
    }
    if (tin.bytes != kFeatures || static_cast<size_t>(tout.bytes) != g.labels.size()) {
        err = "input/output tensor size does not match feature/label count";
        return EIKWS_ERR_SHAPES_DONT_MATCH;
    }
    mf.q_scale = tin.scale();
    mf.q_inv_scale = static_cast<float>(1.0 / static_cast<double>(tin.scale()));
    mf.q_zp = tin.zero_point();
    mf.input_is_int8 = 1;

    // raw material kept aside for the fused plan (see build_fused below)
    struct ConvSrc {
        int op_index;
        const int8_t *w;
        std::vector<int32_t> raw_bias, mult, shift;
    };
    std::vector<ConvSrc> conv_src;
    std::deque<std::vector<int8_t>> dense_filters;  // depthwise filters expanded to dense [out_c][kw][in_c] (stable addresses)
    std::vector<std::pair<int, std::vector<uint8_t>>> lut_src;
    std::vector<int32_t> exp_lut_src;

    uint32_t max_bytes = 0;
    for (const TensorDesc &t : g.tensors)
        if (!t.is_const && t.bytes > max_bytes) max_bytes = t.bytes;
    const int buf_bytes = static_cast<int>((max_bytes + 15) & ~15u);
    std::vector<int> off(g.tensors.size(), -1);
    off[g.input] = 0;
    int max_row = 0;
    nn.n_ops = 0;
    auto other = [&](int o) { return o == 0 ? buf_bytes : 0; };

    for (const NodeDesc &n : g.nodes) {
        if (n.op == kOpReshape) {
            if (n.inputs.empty() || off[n.inputs[0]] < 0) {
                err = "reshape of an unplaced tensor";
                return EIKWS_ERR_UNSUPPORTED;
            }
            off[n.outputs[0]] = off[n.inputs[0]];
            continue;
        }
        if (nn.n_ops >= kMaxNnOps) {
            err = "graph has more compute nodes than kMaxNnOps";
            return EIKWS_ERR_UNSUPPORTED;
        }
        NnOpDev &op = nn.ops[nn.n_ops];
        std::memset(&op, 0, sizeof(op));
        const TensorDesc &out = g.tensors[n.outputs[0]];
        if (out.type != kI8) {
            err = "non-int8 activation tensor";
            return EIKWS_ERR_UNSUPPORTED;
        }
        if (n.op == kOpConv2D || n.op == kOpDepthwiseConv2D || n.op == kOpFullyConnected) {
            const TensorDesc &in = g.tensors[n.inputs[0]], &flt = g.tensors[n.inputs[1]];
            const bool has_bias = n.inputs.size() > 2 && n.inputs[2] >= 0;
            if (off[n.inputs[0]] < 0 || !flt.is_const || flt.type != kI8 || in.type != kI8) {
                err = "conv/fc: unsupported operand types";
                return EIKWS_ERR_UNSUPPORTED;
            }
            for (int32_t z : flt.zero_points)
                if (z != 0) {
                    err = "conv/fc: non-zero filter zero point";
                    return EIKWS_ERR_UNSUPPORTED;
                }
            int32_t lo, hi;
            op.kind = kNnConv1d;
            std::vector<double> eff;  // real multiplier per output channel
            const int8_t *w = reinterpret_cast<const int8_t *>(flt.data.data());
            if (n.op == kOpDepthwiseConv2D) {
                // DepthwiseConvPerChannel (integer_ops/depthwise_conv.h:22-121): output channel oc = m + ic * depth_multiplier
                // sums filter[0][0][fx][oc] * (x[.][ic] + in_offset) over fx only.  In integer arithmetic that is exactly the
                // dense 1xk convolution whose filter is zero for every other input channel, so the filter is expanded once
                // here and the op runs on the conv kernels (a zero weight contributes exactly 0 to acc and to the bias fold).
                int di[4], df[4], dq[4];
                dims4(in, di);
                dims4(flt, df);
                dims4(out, dq);
                const int padding = n.params[0], sw = n.params[1], dm = n.params[3], act = n.params[4], dw = n.params[5];
                if (di[0] != 1 || di[1] != 1 || df[0] != 1 || df[1] != 1 || dq[1] != 1 || dw != 1 || dm < 1 || df[3] != di[3] * dm || dq[3] != df[3]) {
                    err = "depthwise conv: only 1xk filters over [1,1,W,C] inputs with dilation 1 are implemented";
                    return EIKWS_ERR_UNSUPPORTED;
                }
                op.in_w = di[2];
                op.in_c = di[3];
                op.out_w = dq[2];
                op.out_c = dq[3];
                op.kw = df[2];
                op.stride_w = sw;
                op.pad_w = same_or_valid_pad(padding, sw, dw, di[2], df[2]);
                activation_range_i8(act, out, &lo, &hi);
                dense_filters.emplace_back(static_cast<size_t>(op.out_c) * op.kw * op.in_c, static_cast<int8_t>(0));
                std::vector<int8_t> &dense = dense_filters.back();
                for (int oc = 0; oc < op.out_c; oc++) {
                    for (int fx = 0; fx < op.kw; fx++) dense[(static_cast<size_t>(oc) * op.kw + fx) * op.in_c + oc / dm] = w[fx * op.out_c + oc];
                    const float fs = flt.scales.size() > 1 ? flt.scales[oc] : flt.scales[0];  // per-channel along filter dimension 3
                    eff.push_back(static_cast<double>(in.scale()) * static_cast<double>(fs) / static_cast<double>(out.scale()));
                }
                w = dense.data();
            } else if (n.op == kOpConv2D) {
                int di[4], df[4], dq[4];
                dims4(in, di);
                dims4(flt, df);
                dims4(out, dq);
                const int padding = n.params[0], sw = n.params[1], act = n.params[3], dw = n.params[4];
                if (di[0] != 1 || di[1] != 1 || df[1] != 1 || dq[1] != 1 || dw != 1 || df[3] != di[3] || dq[3] != df[0]) {
                    err = "conv: only 1xk convolutions over [1,1,W,C] inputs with dilation 1 are implemented";
                    return EIKWS_ERR_UNSUPPORTED;
                }
                op.in_w = di[2];
                op.in_c = di[3];
                op.out_w = dq[2];
                op.out_c = dq[3];
                op.kw = df[2];
                op.stride_w = sw;
                op.pad_w = same_or_valid_pad(padding, sw, dw, di[2], df[2]);
                activation_range_i8(act, out, &lo, &hi);
                for (int oc = 0; oc < op.out_c; oc++) {
                    // PopulateConvolutionQuantizationParams (kernel_util_lite.cc:88-101)
                    const float fs = flt.scales.size() > 1 ? flt.scales[oc] : flt.scales[0];
                    eff.push_back(static_cast<double>(in.scale()) * static_cast<double>(fs) / static_cast<double>(out.scale()));
                }
            } else {
                const int depth = flt.dims.back();
                const int outc = out.dims.back();
                int batches = 1;
                for (size_t i = 0; i + 1 < out.dims.size(); i++) batches *= out.dims[i];
                if (batches != 1 || static_cast<int>(in.bytes) != depth || static_cast<int>(flt.bytes) != depth * outc) {
                    err = "fully_connected: only batch 1 with a dense [out,in] filter is implemented";
                    return EIKWS_ERR_UNSUPPORTED;
                }
                op.in_w = 1;
                op.in_c = depth;
                op.out_w = 1;
                op.out_c = outc;
                op.kw = 1;
                op.stride_w = 1;
                op.pad_w = 0;
                activation_range_i8(n.params[0], out, &lo, &hi);
                // GetQuantizedConvolutionMultipler (kernel_util_lite.cc:158-171): FLOAT product of the two scales
                const double real = static_cast<double>(in.scale() * flt.scale()) / static_cast<double>(out.scale());
                eff.assign(outc, real);
            }
            op.in_h = op.out_h = op.kh = op.stride_h = 1;
            const int K = op.kw * op.in_c;
            op.k_words = (K + 3) / 4;
            op.in_zp = in.zero_point();
            op.out_zp = out.zero_point();
            op.act_min = lo;
            op.act_max = hi;
            std::vector<int32_t> packed(static_cast<size_t>(op.out_c) * op.k_words, 0), bias(op.out_c), mult(op.out_c), shift(op.out_c);
            const int32_t *bsrc = has_bias ? reinterpret_cast<const int32_t *>(g.tensors[n.inputs[2]].data.data()) : nullptr;
            if (has_bias && (!g.tensors[n.inputs[2]].is_const || g.tensors[n.inputs[2]].type != kI32)) {
                err = "conv/fc: bias must be a constant int32 tensor";
                return EIKWS_ERR_UNSUPPORTED;
            }
            for (int oc = 0; oc < op.out_c; oc++) {
                int32_t wsum = 0;
                uint8_t *dst = reinterpret_cast<uint8_t *>(&packed[static_cast<size_t>(oc) * op.k_words]);
                for (int k = 0; k < K; k++) {
                    dst[k] = static_cast<uint8_t>(w[oc * K + k]);
                    wsum += w[oc * K + k];
                }
                // acc = sum w*(x + in_offset) = sum w*x + in_offset*sum w, with in_offset = -in_zp
                bias[oc] = fold_bias(bsrc ? bsrc[oc] : 0, op.in_zp, wsum);
                int sh;
                quantize_multiplier(eff[oc], &mult[oc], &sh);
                if (!(eff[oc] >= 0.0) || sh < -31 || sh > 30) {  // also refuses NaN: the shifts below would be undefined
                    err = "conv/fc: requantisation multiplier out of range";
                    return EIKWS_ERR_UNSUPPORTED;
                }
                shift[oc] = sh;
            }
            {
                ConvSrc cs;
                cs.op_index = nn.n_ops;
                cs.w = w;
                cs.mult = mult;
                cs.shift = shift;
                for (int oc = 0; oc < op.out_c; oc++) cs.raw_bias.push_back(bsrc ? bsrc[oc] : 0);
                conv_src.push_back(std::move(cs));
            }
            const int lead = op.pad_w * op.in_c, data_bytes = op.in_w * op.in_c;
            int need = (op.out_w - 1) * op.stride_w * op.in_c + 4 * op.k_words + 8;
            if (need < lead + data_bytes) need = lead + data_bytes;
            op.n_elems = (need + 3) & ~3;  // row_bytes
            if (op.n_elems > max_row) max_row = op.n_elems;
            b.bind(op.weights, b.push(packed.data(), packed.size() * 4));
            b.bind(op.bias, b.push(bias.data(), bias.size() * 4));
            b.bind(op.mult, b.push(mult.data(), mult.size() * 4));
            b.bind(op.shift, b.push(shift.data(), shift.size() * 4));
            op.in_off = off[n.inputs[0]];
        } else if (n.op == kOpAdd) {
            // add::CalculateOpData (add.cc:271-309) + AddElementwise (integer_ops/add.h:27-57), tabulated over the
            // 256 values of the activation operand for every element of the constant operand
            int ia = n.inputs[0], ic = n.inputs[1];
            bool const_is_second = true;
            if (g.tensors[ia].is_const && !g.tensors[ic].is_const) {
                std::swap(ia, ic);
                const_is_second = false;
            }
            const TensorDesc &a = g.tensors[ia], &cst = g.tensors[ic];
            if (a.is_const || !cst.is_const || off[ia] < 0 || a.type != kI8 || cst.type != kI8) {
                err = "add: only activation + constant int8 operands are implemented";
                return EIKWS_ERR_UNSUPPORTED;
            }
            // broadcast must be over leading dimensions only (constant dims == trailing dims of the activation)
            const size_t nc = cst.dims.size(), na = a.dims.size();
            bool ok = nc <= na && out.bytes == a.bytes;
            for (size_t i = 0; ok && i < nc; i++) ok = cst.dims[nc - 1 - i] == a.dims[na - 1 - i];
            if (!ok || cst.bytes == 0 || a.bytes % cst.bytes) {
                err = "add: unsupported broadcast pattern";
                return EIKWS_ERR_UNSUPPORTED;
            }
            const TensorDesc &in1 = const_is_second ? a : cst, &in2 = const_is_second ? cst : a;
            const int left_shift = 20;
            const double twice_max = 2 * static_cast<double>(std::max(in1.scale(), in2.scale()));
            int32_t m1, m2, mo;
            int s1, s2, so;
            quantize_multiplier(static_cast<double>(in1.scale()) / twice_max, &m1, &s1);
            quantize_multiplier(static_cast<double>(in2.scale()) / twice_max, &m2, &s2);
            quantize_multiplier(twice_max / ((1 << left_shift) * static_cast<double>(out.scale())), &mo, &so);
            if (s1 > 0 || s2 > 0 || so > 0 || s1 < -31 || s2 < -31 || so < -31) {  // MultiplyByQuantizedMultiplierSmallerThanOneExp needs exponents <= 0
                err = "add: operand / output scales out of the range the fixed-point rescaling supports";
                return EIKWS_ERR_UNSUPPORTED;
            }
            int32_t lo, hi;
            activation_range_i8(n.params[0], out, &lo, &hi);
            std::vector<uint8_t> lut(static_cast<size_t>(cst.bytes) * 256);
            const int8_t *cv = reinterpret_cast<const int8_t *>(cst.data.data());
            for (uint32_t ci = 0; ci < cst.bytes; ci++)
                for (int q = -128; q <= 127; q++) {
                    const int32_t v1 = const_is_second ? q : cv[ci], v2 = const_is_second ? cv[ci] : q;
                    const int32_t x1 = (-in1.zero_point() + v1) * (1 << left_shift), x2 = (-in2.zero_point() + v2) * (1 << left_shift);
                    const int32_t sum = qm::mul_smaller_than_one(x1, m1, s1) + qm::mul_smaller_than_one(x2, m2, s2);
                    int32_t o = qm::mul_smaller_than_one(sum, mo, so) + out.zero_point();
                    o = std::min(hi, std::max(lo, o));
                    lut[static_cast<size_t>(ci) * 256 + static_cast<size_t>(q + 128)] = static_cast<uint8_t>(static_cast<int8_t>(o));
                }
            op.kind = kNnAddLut;
            op.n_elems = static_cast<int32_t>(out.bytes);
            op.n_const = static_cast<int32_t>(cst.bytes);
            b.bind(op.lut, b.push(lut.data(), lut.size()));
            lut_src.emplace_back(nn.n_ops, lut);
            op.in_off = off[ia];
        } else if (n.op == kOpMaxPool2D) {
            const TensorDesc &in = g.tensors[n.inputs[0]];
            int di[4], dq[4];
            dims4(in, di);
            dims4(out, dq);
            if (off[n.inputs[0]] < 0 || in.type != kI8 || di[0] != 1 || di[3] != dq[3]) {
                err = "max_pool: unsupported operand";
                return EIKWS_ERR_UNSUPPORTED;
            }
            op.kind = kNnMaxPool;
            op.in_h = di[1];
            op.in_w = di[2];
            op.in_c = di[3];
            op.out_h = dq[1];
            op.out_w = dq[2];
            op.out_c = dq[3];
            op.stride_w = n.params[1];
            op.stride_h = n.params[2];
            op.kw = n.params[3];
            op.kh = n.params[4];
            op.pad_h = same_or_valid_pad(n.params[0], op.stride_h, 1, op.in_h, op.kh);
            op.pad_w = same_or_valid_pad(n.params[0], op.stride_w, 1, op.in_w, op.kw);
            activation_range_i8(n.params[5], out, &op.act_min, &op.act_max);
            op.in_off = off[n.inputs[0]];
        } else if (n.op == kOpSoftmax) {
            // CalculateSoftmaxParams (softmax.cc:187-226), PreprocessSoftmaxScaling / CalculateInputRadius
            // (quantization_util.cc:269-334)
            const TensorDesc &in = g.tensors[n.inputs[0]];
            int outer = 1;
            for (size_t i = 0; i + 1 < in.dims.size(); i++) outer *= in.dims[i];
            if (off[n.inputs[0]] < 0 || in.type != kI8 || outer != 1 || out.zero_point() != -128 || out.scale() != 1.f / 256) {
                err = "softmax: only a single int8 row with output scale 1/256, zero point -128 is implemented";
                return EIKWS_ERR_UNSUPPORTED;
            }
            float beta;
            std::memcpy(&beta, &n.params[0], 4);
            const int kBits = 5;
            double rm = static_cast<double>(beta) * static_cast<double>(in.scale()) * static_cast<double>(1 << (31 - kBits));
            rm = std::min(rm, (1ll << 31) - 1.0);
            int32_t mult;
            int left_shift;
            quantize_multiplier(rm, &mult, &left_shift);
            if (!(rm > 0.0) || left_shift < 0 || left_shift > 30) {
                err = "softmax: beta * input scale out of range";
                return EIKWS_ERR_UNSUPPORTED;
            }
            const double max_rescaled = 1.0 * ((1 << kBits) - 1) * static_cast<double>(1ll << (31 - kBits)) / static_cast<double>(1ll << left_shift);
            const int diff_min = -1 * static_cast<int>(std::floor(max_rescaled));
            std::vector<int32_t> elut(256);
            for (int i = 0; i < 256; i++) {
                const int diff = -i;
                elut[i] = diff >= diff_min ? qm::exp_on_negative_q5_26(qm::mul_greater_than_one(diff, mult, left_shift)) : -1;
            }
            op.kind = kNnSoftmax;
            op.n_elems = in.dims.back();
            b.bind(op.exp_lut, b.push(elut.data(), elut.size() * 4));
            exp_lut_src = elut;
            op.in_off = off[n.inputs[0]];
        } else {
            err = "operator " + std::to_string(n.op) + " is not implemented (supported: RESHAPE, CONV_2D 1xk, DEPTHWISE_CONV_2D 1xk, ADD const, MAX_POOL_2D, FULLY_CONNECTED, SOFTMAX)";
            return EIKWS_ERR_UNSUPPORTED;
        }
        op.out_off = other(op.in_off);
        off[n.outputs[0]] = op.out_off;
        nn.n_ops++;
    }
    if (off[g.output] < 0) {
        err = "output tensor is never produced";
        return EIKWS_ERR_UNSUPPORTED;
    }
    nn.in_off = 0;
    nn.out_off = off[g.output];
    nn.n_in = kFeatures;
    nn.n_out = static_cast<int32_t>(g.labels.size());
    nn.arena_bytes = 2 * buf_bytes;
    nn.row_bytes = max_row;
    nn.out_scale = tout.scale();
    nn.out_zp = tout.zero_point();
    hp.nn_smem_bytes = nn.arena_bytes + ((max_row + 15) & ~15);

    // ---------- fused plan: conv -> add(const)+act -> maxpool, twice, then fully connected -> softmax ----------
    {
        NnFusedDev &fu = nn.fused;
        std::memset(&fu, 0, sizeof(fu));
        auto find_conv = [&](int idx) -> const ConvSrc * {
            for (const ConvSrc &c : conv_src)
                if (c.op_index == idx) return &c;
            return nullptr;
        };
        auto find_lut = [&](int idx) -> const std::vector<uint8_t> * {
            for (const auto &l : lut_src)
                if (l.first == idx) return &l.second;
            return nullptr;
        };
        bool ok = nn.n_ops == 8 && nn.ops[0].kind == kNnConv1d && nn.ops[1].kind == kNnAddLut && nn.ops[2].kind == kNnMaxPool &&
                  nn.ops[3].kind == kNnConv1d && nn.ops[4].kind == kNnAddLut && nn.ops[5].kind == kNnMaxPool && nn.ops[6].kind == kNnConv1d &&
                  nn.ops[7].kind == kNnSoftmax && !exp_lut_src.empty();
        int arena = 0;
        auto alloc = [&](int bytes) {
            int o = arena;
            arena += (bytes + 15) & ~15;
            return o;
        };
        // two stage shapes are implemented (kernels.cu, fused_stage0 / fused_stage1): the shipped STM32 topology -- conv 1x7, pool 7
        // (fu.shape 0; its tiny second block leaves the pool to the tail warp) -- and the Arduino-zip topology -- conv 1x3, pool 2
        // with SAME padding, i.e. a partial last window (fu.shape 1; FC over pool_out x out_c inputs)
        int shape = -1;
        for (int s2 = 0; ok && s2 < 2; s2++) {
            const NnOpDev &cv = nn.ops[3 * s2], &ad = nn.ops[3 * s2 + 1], &pl = nn.ops[3 * s2 + 2];
            const ConvSrc *cs = find_conv(3 * s2);
            const std::vector<uint8_t> *lut = find_lut(3 * s2 + 1);
            NnFusedStage &st = fu.st[s2];
            const int cp = (cv.in_c + 15) & ~15;
            // the pool must run along the conv width (tensor reshaped to [1, W, 1, C]) with window == stride and no leading padding
            ok = cs && lut && cv.stride_w == 1 && cv.out_w == cv.in_w && ad.n_const == cv.out_c && ad.n_elems == cv.out_w * cv.out_c &&
                 pl.in_w == 1 && pl.kw == 1 && pl.in_h == cv.out_w && pl.in_c == cv.out_c && pl.pad_h == 0 && pl.out_w == 1 && pl.kh == pl.stride_h;
            if (!ok) break;
            const bool shape_a = cv.kw == 7 && pl.kh == 7 && pl.out_h * 7 == pl.in_h && cp == (s2 == 0 ? 16 : 32);
            const bool shape_b = cv.kw == 3 && pl.kh == 2 && pl.out_h == (pl.in_h + 1) / 2 && cp == 16;
            const int this_shape = shape_a ? 0 : (shape_b ? 1 : -1);
            if (s2 == 0) shape = this_shape;
            ok = this_shape >= 0 && this_shape == shape;
            if (s2 == 0) ok = ok && cv.in_w == kFrames && cv.in_c == kCepstra;
            if (!ok) break;
            st.in_w = cv.in_w;
            st.in_c = cv.in_c;
            st.cp = cp;
            st.kw = cv.kw;
            st.pad_w = cv.pad_w;
            st.out_c = cv.out_c;
            st.pool = pl.kh;
            st.pool_out = pl.out_h;
            if (s2 == 1 && shape == 0) {
                // the second block is tiny (7 x 10 outputs): one thread per conv output, the max-pool moves to the tail warp
                ok = pl.out_h == 1;
                st.pool = 1;
                st.pool_out = cv.out_w;
                fu.tail_pool = pl.kh;
                fu.tail_pool_act_min = pl.act_min;
                fu.tail_pool_act_max = pl.act_max;
            } else if (s2 == 1) {
                fu.tail_pool = 1;  // pooled inside the stage
                fu.tail_pool_act_min = -128;
                fu.tail_pool_act_max = 127;
            }
            st.in_zp = cv.in_zp;
            st.conv_out_zp = cv.out_zp;
            st.conv_act_min = cv.act_min;
            st.conv_act_max = cv.act_max;
            const bool pool_in_tail = s2 == 1 && shape == 0;
            st.pool_act_min = pool_in_tail ? -128 : pl.act_min;
            st.pool_act_max = pool_in_tail ? 127 : pl.act_max;
            // padded input rows: the image plus the conv halo, and every row the LAST pool group's window touches (a partial window
            // still reads its full POOL + KW - 1 rows; the outputs beyond the image are dropped, not their reads)
            st.in_rows = std::max(cv.in_w + cv.kw - 1, st.pool_out * st.pool + cv.kw - 1);
            // the fused stage pools the ACCUMULATORS (max commutes with monotone steps): needs non-negative multipliers
            // and a non-decreasing ADD+activation table for every channel
            for (int oc = 0; ok && oc < cv.out_c; oc++) {
                ok = cs->mult[oc] >= 0;
                for (int i = 1; ok && i < 256; i++)
                    ok = static_cast<int8_t>((*lut)[static_cast<size_t>(oc) * 256 + i]) >= static_cast<int8_t>((*lut)[static_cast<size_t>(oc) * 256 + i - 1]);
            }
            if (!ok) break;
            std::vector<int32_t> packed(static_cast<size_t>(cv.out_c) * cv.kw * (cp / 4), 0), bias(cv.out_c);
            for (int oc = 0; oc < cv.out_c; oc++) {
                int32_t wsum = 0;
                for (int kx = 0; kx < cv.kw; kx++)
                    for (int c2 = 0; c2 < cv.in_c; c2++) {
                        const int8_t wv = cs->w[(oc * cv.kw + kx) * cv.in_c + c2];
                        reinterpret_cast<uint8_t *>(packed.data())[(static_cast<size_t>(oc) * cv.kw + kx) * cp + c2] = static_cast<uint8_t>(wv);
                        wsum += wv;
                    }
                bias[oc] = fold_bias(cs->raw_bias[oc], cv.in_zp, wsum);
            }
            b.bind(st.weights, b.push(packed.data(), packed.size() * 4));
            b.bind(st.bias, b.push(bias.data(), bias.size() * 4));
            b.bind(st.mult, b.push(cs->mult.data(), cs->mult.size() * 4));
            b.bind(st.shift, b.push(cs->shift.data(), cs->shift.size() * 4));
            b.bind(st.lut, b.push(lut->data(), lut->size()));
        }
        if (ok) {
            const NnOpDev &fc = nn.ops[6], &sm = nn.ops[7];
            const ConvSrc *cs = find_conv(6);
            // FC input = block 2's output: [tail_pool positions][out_c] still to be pooled by the tail (shape 0) or the pooled
            // [pool_out][out_c] matrix (shape 1)
            const int fc_in = shape == 0 ? fu.st[1].out_c : fu.st[1].pool_out * fu.st[1].out_c;
            const int blk2_bytes = shape == 0 ? fc.in_c * fu.tail_pool : fc_in;
            ok = cs && fc.in_w == 1 && fc.kw == 1 && fc.in_c == fc_in && (shape != 0 || fu.st[1].pool_out == fu.tail_pool) &&
                 fu.st[1].in_w == fu.st[0].pool_out && fu.st[1].in_c == fu.st[0].out_c && fc.out_c <= 32 && (shape != 0 || fc.in_c <= 32) &&
                 fu.st[0].in_rows * fu.st[0].cp <= 1024 && fu.st[1].in_rows * fu.st[1].cp <= 512 && blk2_bytes <= 256 &&  // region S of the kernel's shared memory map
                 fu.st[0].out_c <= 32 && fu.st[1].out_c <= 32 && sm.n_elems == fc.out_c && fc.out_c == static_cast<int>(g.labels.size());
            if (ok) {
                NnFusedStage &s0 = fu.st[0], &s1 = fu.st[1];
                s0.in_off = alloc(s0.in_rows * s0.cp);
                s1.in_off = alloc(s1.in_rows * s1.cp);
                fu.fc_in_off = alloc(blk2_bytes);  // block 2's output
                fu.tail_off = alloc(64);
                s0.out_off = s1.in_off;  // stage 0 writes stage 1's padded input
                s0.out_cp = s1.cp;
                s0.out_row0 = s1.pad_w;
                s0.out_rows = s1.in_rows;
                s0.out_fill = s1.in_zp;
                s1.out_off = fu.fc_in_off;  // stage 1 writes the dense FC input
                s1.out_cp = s1.out_c;
                s1.out_row0 = 0;
                s1.out_rows = s1.pool_out;
                s1.out_fill = 0;
                fu.fc_d = fc.in_c;
                fu.fc_o = fc.out_c;
                fu.fc_in_zp = fc.in_zp;
                fu.fc_out_zp = fc.out_zp;
                fu.fc_act_min = fc.act_min;
                fu.fc_act_max = fc.act_max;
                fu.fc_mult = cs->mult[0];
                fu.fc_shift = cs->shift[0];
                std::vector<int32_t> fbias(fc.out_c);
                for (int oc = 0; oc < fc.out_c; oc++) {
                    int32_t wsum = 0;
                    for (int d = 0; d < fc.in_c; d++) wsum += cs->w[oc * fc.in_c + d];
                    fbias[oc] = fold_bias(cs->raw_bias[oc], fc.in_zp, wsum);
                }
                b.bind(fu.fc_w, b.push(cs->w, static_cast<size_t>(fc.out_c) * fc.in_c));
                b.bind(fu.fc_bias, b.push(fbias.data(), fbias.size() * 4));
                b.bind(fu.exp_lut, b.push(exp_lut_src.data(), exp_lut_src.size() * 4));
                fu.enabled = 1;
                fu.shape = shape;
                // A operand of the tensor-core block 1 (dev_plan.h): needs the shipped stage-0 shape (7 taps of one 16-byte row)
                if (s0.kw == 7 && s0.cp == 16 && s0.out_c <= 30 && s0.in_w == kFrames && s0.pool == 7 && s0.pool_out == 7 && s0.pad_w == 3) {
                    const ConvSrc *c0 = find_conv(0);
                    std::vector<int8_t> tcw(8 * 64 * 16, 0);
                    for (int oc = 0; oc < s0.out_c; oc++)
                        for (int kx = 0; kx < s0.kw; kx++)
                            for (int c2 = 0; c2 < s0.in_c; c2++) {
                                const int8_t wv = c0->w[(oc * s0.kw + kx) * s0.in_c + c2];
                                tcw[(static_cast<size_t>(kx) * 64 + oc) * 16 + c2] = wv;
                                tcw[(static_cast<size_t>(kx) * 64 + 32 + oc) * 16 + c2] = wv;
                            }
                    b.bind(fu.tc_w, b.push(tcw.data(), tcw.size()));
                    fu.tc_enabled = 1;
                }
                if (arena > hp.nn_smem_bytes) hp.nn_smem_bytes = arena;
            }
        }
        if (!ok) std::memset(&fu, 0, sizeof(fu));
    }
    if (hp.nn_smem_bytes > nn_smem_capacity(false)) {
        err = "classifier activations do not fit the fused kernel's shared-memory overlay";
        return EIKWS_ERR_UNSUPPORTED;
    }
    for (const Fixup &f : b.fixups) hp.fixups.emplace_back(f.field_off, f.blob_off);
    return EIKWS_OK;
}

cudaError_t upload_plan(const HostPlan &hp, DevicePlan &dp) {
    dp = DevicePlan();
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&dp.d_blob), hp.blob.size() ? hp.blob.size() : 16);
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(dp.d_blob, hp.blob.data(), hp.blob.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    DevPlan img = hp.dev;
    for (const auto &f : hp.fixups) {  // rebase table pointers onto the device allocation
        uint8_t *p = dp.d_blob + f.second;
        std::memcpy(reinterpret_cast<uint8_t *>(&img) + f.first, &p, sizeof(p));
    }
    e = cudaMalloc(reinterpret_cast<void **>(&dp.d_plan), sizeof(DevPlan));
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(dp.d_plan, &img, sizeof(DevPlan), cudaMemcpyHostToDevice);
    dp.nn_smem_bytes = hp.nn_smem_bytes;
    if (e != cudaSuccess) return e;
    // pageable cudaMemcpy may return before the DMA has landed, and the kernels run on non-blocking streams that do not
    // synchronise with the default stream: make the plan globally visible before the handle exists
    return cudaDeviceSynchronize();
}

void free_plan(DevicePlan &dp) {
    if (dp.d_plan) cudaFree(dp.d_plan);
    if (dp.d_blob) cudaFree(dp.d_blob);
    dp = DevicePlan();
}

}  // namespace eikws
