// "EIKWSMDL" v1 container: writer and reader (see include/eikws_model_format.md).
#include "model_graph.h"

#include <cstring>

namespace eikws {
namespace {

struct Writer {
    std::vector<uint8_t> &o;
    void u32(uint32_t v) {
        uint8_t b[4];
        std::memcpy(b, &v, 4);
        o.insert(o.end(), b, b + 4);
    }
    void i32(int32_t v) { u32(static_cast<uint32_t>(v)); }
    void f32(float f) {
        uint32_t v;
        std::memcpy(&v, &f, 4);
        u32(v);
    }
    void bytes(const void *p, size_t n) {
        const uint8_t *b = static_cast<const uint8_t *>(p);
        o.insert(o.end(), b, b + n);
        while (o.size() & 3) o.push_back(0);
    }
};

struct Reader {
    const uint8_t *p, *end;
    bool ok = true;
    uint32_t u32() {
        if (end - p < 4) {
            ok = false;
            return 0;
        }
        uint32_t v;
        std::memcpy(&v, p, 4);
        p += 4;
        return v;
    }
    int32_t i32() { return static_cast<int32_t>(u32()); }
    float f32() {
        uint32_t v = u32();
        float f;
        std::memcpy(&f, &v, 4);
        return f;
    }
    const uint8_t *bytes(size_t n) {
        size_t padded = (n + 3) & ~size_t(3);
        if (static_cast<size_t>(end - p) < padded) {
            ok = false;
            return nullptr;
        }
        const uint8_t *q = p;
        p += padded;
        return q;
    }
};

}  // namespace

void serialize_model(const ModelGraph &g, std::vector<uint8_t> &out) {
    out.clear();
    Writer w{out};
    out.insert(out.end(), {'E', 'I', 'K', 'W', 'S', 'M', 'D', 'L'});
    w.u32(1);
    w.u32(static_cast<uint32_t>(g.tensors.size()));
    w.u32(static_cast<uint32_t>(g.nodes.size()));
    w.u32(g.input);
    w.u32(g.output);
    w.u32(static_cast<uint32_t>(g.labels.size()));
    w.u32(g.raw_sample_count);
    w.u32(g.nn_input_frame_size);
    w.i32(g.mfcc.sample_rate);
    w.i32(g.mfcc.num_cepstral);
    w.f32(g.mfcc.frame_length);
    w.f32(g.mfcc.frame_stride);
    w.i32(g.mfcc.num_filters);
    w.i32(g.mfcc.fft_length);
    w.i32(g.mfcc.win_size);
    w.i32(g.mfcc.low_frequency);
    w.i32(g.mfcc.high_frequency);
    w.f32(g.mfcc.pre_cof);
    w.i32(g.mfcc.pre_shift);
    for (const std::string &s : g.labels) {
        w.u32(static_cast<uint32_t>(s.size()));
        w.bytes(s.data(), s.size());
    }
    for (const TensorDesc &t : g.tensors) {
        w.u32(t.type);
        w.u32(t.is_const ? 1u : 0u);
        w.u32(static_cast<uint32_t>(t.dims.size()));
        for (int32_t d : t.dims) w.i32(d);
        w.u32(t.bytes);
        w.u32(static_cast<uint32_t>(t.scales.size()));
        for (float s : t.scales) w.f32(s);
        for (int32_t z : t.zero_points) w.i32(z);
        w.i32(t.quantized_dimension);
        if (t.is_const) w.bytes(t.data.data(), t.data.size());
    }
    for (const NodeDesc &n : g.nodes) {
        w.u32(n.op);
        w.u32(static_cast<uint32_t>(n.inputs.size()));
        for (int32_t v : n.inputs) w.i32(v);
        w.u32(static_cast<uint32_t>(n.outputs.size()));
        for (int32_t v : n.outputs) w.i32(v);
        w.u32(static_cast<uint32_t>(n.params.size()));
        for (int32_t v : n.params) w.i32(v);
    }
}

bool parse_model(const void *blob, size_t bytes, ModelGraph &g, std::string &err) {
    g = ModelGraph();
    if (!blob || bytes < 16 || std::memcmp(blob, "EIKWSMDL", 8) != 0) {
        err = "not an EIKWSMDL container";
        return false;
    }
    Reader r{static_cast<const uint8_t *>(blob) + 8, static_cast<const uint8_t *>(blob) + bytes};
    if (r.u32() != 1) {
        err = "unsupported EIKWSMDL version";
        return false;
    }
    uint32_t nt = r.u32(), nn = r.u32();
    g.input = r.u32();
    g.output = r.u32();
    uint32_t nl = r.u32();
    g.raw_sample_count = r.u32();
    g.nn_input_frame_size = r.u32();
    g.mfcc.sample_rate = r.i32();
    g.mfcc.num_cepstral = r.i32();
    g.mfcc.frame_length = r.f32();
    g.mfcc.frame_stride = r.f32();
    g.mfcc.num_filters = r.i32();
    g.mfcc.fft_length = r.i32();
    g.mfcc.win_size = r.i32();
    g.mfcc.low_frequency = r.i32();
    g.mfcc.high_frequency = r.i32();
    g.mfcc.pre_cof = r.f32();
    g.mfcc.pre_shift = r.i32();
    if (!r.ok || nt > 4096 || nn > 4096 || nl > 4096) {
        err = "truncated or implausible header";
        return false;
    }
    for (uint32_t i = 0; i < nl; i++) {
        uint32_t len = r.u32();
        const uint8_t *s = r.bytes(len);
        if (!r.ok) break;
        g.labels.emplace_back(reinterpret_cast<const char *>(s), len);
    }
    for (uint32_t i = 0; i < nt && r.ok; i++) {
        TensorDesc t;
        t.type = r.u32();
        t.is_const = r.u32() != 0;
        uint32_t nd = r.u32();
        if (nd > 8) r.ok = false;
        for (uint32_t d = 0; d < nd && r.ok; d++) t.dims.push_back(r.i32());
        t.bytes = r.u32();
        uint32_t nq = r.u32();
        if (nq > 65536) r.ok = false;
        for (uint32_t q = 0; q < nq && r.ok; q++) t.scales.push_back(r.f32());
        for (uint32_t q = 0; q < nq && r.ok; q++) t.zero_points.push_back(r.i32());
        t.quantized_dimension = r.i32();
        if (t.is_const && r.ok) {
            const uint8_t *d = r.bytes(t.bytes);
            if (r.ok) t.data.assign(d, d + t.bytes);
        }
        g.tensors.push_back(std::move(t));
    }
    for (uint32_t i = 0; i < nn && r.ok; i++) {
        NodeDesc n;
        n.op = r.u32();
        uint32_t a = r.u32();
        if (a > 16) r.ok = false;
        for (uint32_t k = 0; k < a && r.ok; k++) n.inputs.push_back(r.i32());
        a = r.u32();
        if (a > 16) r.ok = false;
        for (uint32_t k = 0; k < a && r.ok; k++) n.outputs.push_back(r.i32());
        a = r.u32();
        if (a > 32) r.ok = false;
        for (uint32_t k = 0; k < a && r.ok; k++) n.params.push_back(r.i32());
        g.nodes.push_back(std::move(n));
    }
    if (!r.ok) {
        err = "truncated EIKWSMDL container";
        return false;
    }
    if (g.input >= g.tensors.size() || g.output >= g.tensors.size()) {
        err = "input/output tensor index out of range";
        return false;
    }
    for (const NodeDesc &n : g.nodes) {
        for (int32_t v : n.inputs)
            if (v >= static_cast<int32_t>(g.tensors.size())) {
                err = "node input index out of range";
                return false;
            }
        for (int32_t v : n.outputs)
            if (v < 0 || v >= static_cast<int32_t>(g.tensors.size())) {
                err = "node output index out of range";
                return false;
            }
    }
    return validate_model(g, err);
}

// Structural validation of an untrusted container, done once so that the lowerings in plan.cpp can index freely: every
// index they dereference exists, every constant holds as many bytes as its shape says, every divisor is non-zero.
// (Whether the graph is one the kernels IMPLEMENT is still decided by plan.cpp -> EIKWS_ERR_UNSUPPORTED.)
bool validate_model(const ModelGraph &g, std::string &err) {
    auto bad = [&](const std::string &m) {
        err = "malformed model: " + m;
        return false;
    };
    const MfccConfig &c = g.mfcc;
    if (g.raw_sample_count == 0 || g.raw_sample_count > (1u << 24) || g.nn_input_frame_size == 0 || g.nn_input_frame_size > (1u << 24))
        return bad("raw sample count / feature count out of range");
    if (c.sample_rate <= 0 || c.sample_rate > 1000000) return bad("sample rate out of range");
    if (!(c.frame_length > 0.0f) || !(c.frame_length < 10.0f) || !(c.frame_stride > 0.0f) || !(c.frame_stride < 10.0f)) return bad("frame length/stride out of range");
    if (c.num_filters < 1 || c.num_filters > 1024 || c.num_cepstral < 1 || c.num_cepstral > 1024 || c.fft_length < 2 || c.fft_length > 65536 ||
        c.win_size < 1 || c.win_size > 65535)
        return bad("MFCC block parameter out of range");
    if (c.low_frequency < 0 || c.high_frequency < 0 || c.high_frequency > c.sample_rate / 2 ||
        c.low_frequency >= (c.high_frequency == 0 ? c.sample_rate / 2 : c.high_frequency))
        return bad("mel band must satisfy 0 <= low_frequency < high_frequency <= sample_rate / 2");
    if (!(c.pre_cof == c.pre_cof)) return bad("pre-emphasis coefficient is NaN");
    if (g.labels.empty() || g.labels.size() > 1024) return bad("label count out of range");
    const int nt = static_cast<int>(g.tensors.size());
    for (int i = 0; i < nt; i++) {
        const TensorDesc &t = g.tensors[i];
        size_t elem = 0;
        switch (t.type) {
            case kF32: case kI32: elem = 4; break;
            case kU8: case kI8: elem = 1; break;
            default: return bad("tensor " + std::to_string(i) + " has an unsupported element type");
        }
        uint64_t count = 1;
        for (int32_t d : t.dims) {
            if (d < 1 || d > (1 << 24)) return bad("tensor " + std::to_string(i) + " has a non-positive or huge dimension");
            count *= static_cast<uint64_t>(d);
            if (count > (1u << 26)) return bad("tensor " + std::to_string(i) + " is implausibly large");
        }
        if (static_cast<uint64_t>(t.bytes) != count * elem) return bad("tensor " + std::to_string(i) + ": byte size does not match its shape");
        if (t.is_const && t.data.size() != t.bytes) return bad("tensor " + std::to_string(i) + ": constant data size mismatch");
        if (t.scales.size() != t.zero_points.size()) return bad("tensor " + std::to_string(i) + ": scale / zero-point count mismatch");
        for (float sc : t.scales)
            if (!(sc > 0.0f) || !(sc < 3.0e38f)) return bad("tensor " + std::to_string(i) + ": quantisation scale must be finite and positive");
        if (t.type == kI8 || t.type == kU8)
            for (int32_t z : t.zero_points)
                if (z < -128 || z > 255) return bad("tensor " + std::to_string(i) + ": zero point outside the 8-bit range");
        if (t.scales.size() > 1) {
            if (t.quantized_dimension < 0 || t.quantized_dimension >= static_cast<int32_t>(t.dims.size()) ||
                static_cast<size_t>(t.dims[t.quantized_dimension]) != t.scales.size())
                return bad("tensor " + std::to_string(i) + ": per-channel quantisation does not match its shape");
        }
        if (t.type == kI8 && !t.is_const && t.scales.empty()) return bad("tensor " + std::to_string(i) + ": int8 activation without quantisation parameters");
    }
    auto need = [&](const NodeDesc &n, size_t in, size_t out, size_t par) { return n.inputs.size() >= in && n.outputs.size() >= out && n.params.size() >= par; };
    for (size_t k = 0; k < g.nodes.size(); k++) {
        const NodeDesc &n = g.nodes[k];
        const std::string at = "node " + std::to_string(k);
        bool ok = true;
        size_t required_inputs = 1;
        switch (n.op) {
            case kOpConv2D: ok = need(n, 2, 1, 6); required_inputs = 2; break;
            case kOpDepthwiseConv2D: ok = need(n, 2, 1, 7); required_inputs = 2; break;
            case kOpFullyConnected: ok = need(n, 2, 1, 1); required_inputs = 2; break;
            case kOpAdd: ok = need(n, 2, 1, 1); required_inputs = 2; break;
            case kOpMaxPool2D: case kOpAveragePool2D: ok = need(n, 1, 1, 6); break;
            case kOpSoftmax: ok = need(n, 1, 1, 1); break;
            case kOpReshape: ok = need(n, 1, 1, 0); break;
            default: ok = need(n, 1, 1, 0); break;  // unknown operators are refused by the lowering; keep their indices sane
        }
        if (!ok) return bad(at + ": too few inputs, outputs or parameters for its operator");
        for (size_t i = 0; i < n.inputs.size(); i++) {
            if (n.inputs[i] < -1 || n.inputs[i] >= nt) return bad(at + ": input index out of range");
            if (i < required_inputs && n.inputs[i] < 0) return bad(at + ": a required input is missing");
        }
        for (int32_t v : n.outputs)
            if (v < 0 || v >= nt) return bad(at + ": output index out of range");
        for (size_t i = 0; i < required_inputs; i++)
            if (g.tensors[n.inputs[i]].dims.empty()) return bad(at + ": scalar operand");
        const TensorDesc &out = g.tensors[n.outputs[0]];
        if (out.dims.empty()) return bad(at + ": scalar output");
        if (out.is_const) return bad(at + ": writes a constant tensor");
        if (n.op == kOpConv2D || n.op == kOpDepthwiseConv2D || n.op == kOpFullyConnected) {
            const TensorDesc &flt = g.tensors[n.inputs[1]];
            const size_t out_c = static_cast<size_t>(out.dims.back());
            if (flt.type == kI8 && flt.scales.size() != 1 && flt.scales.size() != out_c) return bad(at + ": filter scales must be per-tensor or one per output channel");
            if (n.inputs.size() > 2 && n.inputs[2] >= 0 && static_cast<size_t>(g.tensors[n.inputs[2]].bytes) != 4 * out_c)
                return bad(at + ": bias size does not match the output channels");
            if (n.op != kOpFullyConnected && (n.params[1] < 1 || n.params[2] < 1 || n.params[1] > 1024 || n.params[2] > 1024)) return bad(at + ": stride out of range");
            if (n.op == kOpDepthwiseConv2D && (n.params[3] < 1 || n.params[3] > 1024)) return bad(at + ": depth multiplier out of range");
        }
        if (n.op == kOpMaxPool2D || n.op == kOpAveragePool2D) {
            for (int i = 1; i <= 4; i++)
                if (n.params[i] < 1 || n.params[i] > 65536) return bad(at + ": pool stride / window out of range");
        }
    }
    return true;
}

}  // namespace eikws
