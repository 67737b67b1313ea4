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
    return true;
}

}  // namespace eikws
