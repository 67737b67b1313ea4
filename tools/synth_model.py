#!/usr/bin/env python
"""Synthesise the 12-label (10 keywords + _noise/_unknown) int8 model of BASELINE.json config 4.

The reference ships no such model, so one is derived from its L432 export (the copy whose generated files use
include-path-relative includes) IN THE SAME GENERATED-FILE FORMAT, so that
(1) the unmodified reference SDK can be compiled against it as the oracle, and (2) the product ingests it through the
very same Register_*/trained_model_init boundary as a real Edge Impulse export.

    python tools/synth_model.py <L432 export root> <out dir>
    python tools/synth_model.py --float <L432 export root (template)> <L476 export root (weights)> <out dir>
      -> BASELINE config 5: float32 twin of the L476 model (weights dequantised, all tensors kTfLiteFloat32)
    python tools/synth_model.py --depthwise <L432 export root> <out dir>
      -> the L432 graph with its second convolution replaced by DEPTHWISE_CONV_2D (depth multiplier 1, 30 channels, k7):
         the operator the north star names but no shipped graph uses

Nothing from the reference is stored in this repo: its generated files are read as TEMPLATES at generation time and
only the data tables are substituted (regex), the result is written under the (git-ignored) output directory:
    <out>/model-parameters/{model_metadata.h,dsp_blocks.h}   <out>/tflite-model/trained_model_compiled.{h,cpp}
New tables (seeded numpy PRNG, seed 0x12AB): every conv / fully-connected weight ~ clip(round(N(0,35))), per-channel
weight scales = shipped scale x log-uniform[0.7,1.4], ADD constants uniform[-127,127], FC becomes [12,10] with
bias ~ U[-400,400]; activation scales/zero points are kept so the graph stays numerically sane.
"""
import os
import re
import sys

import numpy as np

LABELS = ["_noise", "_unknown", "down", "go", "left", "no", "off", "on", "right", "stop", "up", "yes"]
SEED = 0x12AB


def fmt_int_array(a, per_line=16):
    a = [int(v) for v in a]
    lines = [", ".join(str(v) for v in a[i:i + per_line]) + ", " for i in range(0, len(a), per_line)]
    return "{ \n  " + "\n  ".join(lines) + "\n}"


def fmt_float_array(a):
    return ", ".join(repr(float(np.float32(v).astype(np.float64))) if False else f"{float(v):.20g}" for v in a) + ", "


def sub_one(pattern, repl, text, flags=re.S):
    new, n = re.subn(pattern, lambda m: repl, text, count=1, flags=flags)
    if n != 1:
        raise RuntimeError("template pattern not found: " + pattern[:60])
    return new


def set_data(text, idx, ctype, dim_expr, values):
    pat = r"const ALIGN\(8\) \w+ tensor_data%d\[[^\]]*\] = \{.*?\};" % idx
    return sub_one(pat, f"const ALIGN(8) {ctype} tensor_data{idx}[{dim_expr}] = {fmt_int_array(values)};", text)


def set_dims(text, idx, dims):
    pat = r"const TfArray<\d+, int> tensor_dimension%d = \{[^;]*\};" % idx
    return sub_one(pat, f"const TfArray<{len(dims)}, int> tensor_dimension{idx} = {{ {len(dims)}, {{ {','.join(str(d) for d in dims)} }} }};", text)


def get_scales(text, idx):
    m = re.search(r"const TfArray<(\d+), float> quant%d_scale = \{ \d+, \{ ([^}]*)\} \};" % idx, text)
    return np.array([float(v) for v in m.group(2).split(",") if v.strip()], np.float64)


def set_scales(text, idx, scales):
    n = len(scales)
    text = sub_one(r"const TfArray<\d+, float> quant%d_scale = \{[^;]*\};" % idx,
                   f"const TfArray<{n}, float> quant{idx}_scale = {{ {n}, {{ {fmt_float_array(scales)}}} }};", text)
    return sub_one(r"const TfArray<\d+, int> quant%d_zero = \{[^;]*\};" % idx,
                   f"const TfArray<{n}, int> quant{idx}_zero = {{ {n}, {{ {','.join(['0'] * n)} }} }};", text)


def set_tensor_bytes(text, data_name, nbytes):
    # tensorData[] row of a constant tensor: { kTfLiteMmapRo, <type>, (void*)tensor_dataN, (TfLiteIntArray*)&tensor_dimensionN, <bytes>, ...
    pat = r"(\(void\*\)%s, \(TfLiteIntArray\*\)&tensor_dimension\d+, )\d+(,)" % data_name
    new, n = re.subn(pat, lambda m: m.group(1) + str(nbytes) + m.group(2), text, count=1)
    if n != 1:
        raise RuntimeError("tensorData row not found for " + data_name)
    return new


def set_arena_tensor_bytes(text, dim_idx, nbytes):
    pat = r"(tensor_arena \+ \d+, \(TfLiteIntArray\*\)&tensor_dimension%d, )\d+(,)" % dim_idx
    new, n = re.subn(pat, lambda m: m.group(1) + str(nbytes) + m.group(2), text, count=1)
    if n != 1:
        raise RuntimeError("arena tensor row not found for dimension %d" % dim_idx)
    return new


def get_int_array(text, idx):
    m = re.search(r"const ALIGN\(8\) \w+ tensor_data%d\[[^\]]*\] = \{(.*?)\};" % idx, text, flags=re.S)
    body = re.sub(r"/\*.*?\*/", " ", m.group(1), flags=re.S)  # generated files annotate rows with /* [i][j][][] */
    return np.array([int(v) for v in body.replace("\n", " ").split(",") if v.strip()], np.int64)


def fmt_f32_array(a, per_line=8):
    a = [np.float32(v) for v in a]
    lines = [", ".join("%.9g" % float(v) for v in a[i:i + per_line]) + ", " for i in range(0, len(a), per_line)]
    return "{ \n  " + "\n  ".join(lines) + "\n}"


def main_float(template_root, weights_root, out_root):
    """BASELINE config 5: float32 twin of the int8 model under <weights_root> (same topology): w = scale_c * q,
    biases b = scale * q, all tensors kTfLiteFloat32, no quantisation, arena re-planned without aliasing."""
    cpp = open(os.path.join(template_root, "tflite-model", "trained_model_compiled.cpp")).read()
    src = open(os.path.join(weights_root, "tflite-model", "trained_model_compiled.cpp")).read()
    hdr = open(os.path.join(template_root, "tflite-model", "trained_model_compiled.h")).read()
    meta = open(os.path.join(template_root, "model-parameters", "model_metadata.h")).read()
    src_meta = open(os.path.join(weights_root, "model-parameters", "model_metadata.h")).read()
    blocks = open(os.path.join(template_root, "model-parameters", "dsp_blocks.h")).read()
    n_labels = int(re.search(r"#define EI_CLASSIFIER_LABEL_COUNT\s+(\d+)", src_meta).group(1))
    labels = re.search(r"ei_classifier_inferencing_categories\[\] = \{([^}]*)\}", src_meta).group(1)
    dims = {2: "30", 3: "10", 4: str(n_labels), 5: f"{n_labels}*10", 6: "30", 7: "30*1*7*13", 8: "10", 9: "10*1*7*30"}
    per_channel = {7: 7 * 13, 9: 7 * 30, 6: 1, 8: 1}  # elements per output channel for per-channel scaled tensors
    for idx in (2, 3, 4, 5, 6, 7, 8, 9):
        q = get_int_array(src, idx).astype(np.float32)
        sc = get_scales(src, idx).astype(np.float32)
        if len(sc) > 1:
            sc = np.repeat(sc, per_channel[idx])
        vals = (sc * q).astype(np.float32)  # float32 product, the value a float model would hold
        pat = r"const ALIGN\(8\) \w+ tensor_data%d\[[^\]]*\] = \{.*?\};" % idx
        cpp = sub_one(pat, f"const ALIGN(8) float tensor_data{idx}[{dims[idx]}] = {fmt_f32_array(vals)};", cpp)
    cpp = set_dims(cpp, 4, [n_labels])
    cpp = set_dims(cpp, 5, [n_labels, 10])
    cpp = set_dims(cpp, 29, [1, n_labels])
    cpp = set_dims(cpp, 30, [1, n_labels])
    # tensor table: every int8 / quantised int32 tensor becomes float32 without quantisation; arena offsets are
    # re-planned sequentially (no aliasing), 16-byte aligned
    elems = {}
    for m in re.finditer(r"const TfArray<(\d+), int> tensor_dimension(\d+) = \{ \d+, \{ ([^}]*)\} \};", cpp):
        elems[int(m.group(2))] = int(np.prod([int(v) for v in m.group(3).split(",")]))
    off = 0

    def row(m):
        nonlocal off
        alloc, typ, place, dim, nbytes, quant = m.group(1), m.group(2), m.group(3), int(m.group(4)), int(m.group(5)), m.group(6)
        is_float = typ == "kTfLiteInt8" or (typ == "kTfLiteInt32" and "Affine" in quant)
        if is_float:
            typ, nbytes, quant = "kTfLiteFloat32", 4 * elems[dim], "{kTfLiteNoQuantization, nullptr}"
        if alloc == "kTfLiteArenaRw":
            place = f"tensor_arena + {off}"
            off += (nbytes + 15) // 16 * 16
        return f"{{ {alloc}, {typ}, {place}, (TfLiteIntArray*)&tensor_dimension{dim}, {nbytes}, {quant}, }}"

    cpp, n = re.subn(r"\{ (kTfLite\w+), (kTfLite\w+), ([^,]+), \(TfLiteIntArray\*\)&tensor_dimension(\d+), (\d+), (\{kTfLite\w+, [^}]*\}), \}", row, cpp)
    assert n == 31, n
    cpp = sub_one(r"constexpr int kTensorArenaSize = \d+;", f"constexpr int kTensorArenaSize = {off + 8192};", cpp)
    meta = sub_one(r"#define EI_CLASSIFIER_LABEL_COUNT\s+\d+", f"#define EI_CLASSIFIER_LABEL_COUNT                {n_labels}", meta)
    meta = sub_one(r"const char\* ei_classifier_inferencing_categories\[\] = \{[^}]*\};",
                   "const char* ei_classifier_inferencing_categories[] = {" + labels + "};", meta)
    for key, val in (("EI_CLASSIFIER_TFLITE_INPUT_DATATYPE", "EI_CLASSIFIER_DATATYPE_FLOAT32"), ("EI_CLASSIFIER_TFLITE_INPUT_QUANTIZED", "0"),
                     ("EI_CLASSIFIER_TFLITE_OUTPUT_DATATYPE", "EI_CLASSIFIER_DATATYPE_FLOAT32"), ("EI_CLASSIFIER_TFLITE_OUTPUT_QUANTIZED", "0")):
        meta = sub_one(r"#define %s\s+\S+" % key, f"#define {key}      {val}", meta)
    # keep the DSP block of the weights' model (low/high frequency etc.)
    cfg = re.search(r"ei_dsp_config_mfcc_t \w+ = \{(.*?)\};", src_meta, flags=re.S).group(1)
    meta = sub_one(r"(ei_dsp_config_mfcc_t \w+ = \{).*?(\};)", re.search(r"ei_dsp_config_mfcc_t \w+ = \{", meta).group(0) + cfg + "};", meta)
    for sub, name, text in (("tflite-model", "trained_model_compiled.cpp", cpp), ("tflite-model", "trained_model_compiled.h", hdr),
                            ("model-parameters", "model_metadata.h", meta), ("model-parameters", "dsp_blocks.h", blocks)):
        os.makedirs(os.path.join(out_root, sub), exist_ok=True)
        with open(os.path.join(out_root, sub, name), "w") as f:
            f.write(text)
    print("synthesised float32 twin under", out_root, "arena", off + 8192)


def main_depthwise(src_root, out_root):
    """conv k7 13->30, ADD+ReLU, pool 7, DEPTHWISE conv k7 (30 channels, multiplier 1), ADD+ReLU, pool 7, FC 30->3, softmax"""
    rng = np.random.default_rng(SEED + 1)
    cpp = open(os.path.join(src_root, "tflite-model", "trained_model_compiled.cpp")).read()
    hdr = open(os.path.join(src_root, "tflite-model", "trained_model_compiled.h")).read()
    meta = open(os.path.join(src_root, "model-parameters", "model_metadata.h")).read()
    blocks = open(os.path.join(src_root, "model-parameters", "dsp_blocks.h")).read()
    n_labels = int(re.search(r"#define EI_CLASSIFIER_LABEL_COUNT\s+(\d+)", meta).group(1))
    C2 = 30  # channels through the second block

    def weights(count):
        return np.clip(np.round(rng.normal(0, 35, count)), -127, 127).astype(np.int64)

    # operator table: one more registration, node 7 switches to it
    cpp = sub_one(r"OP_SOFTMAX,\s+OP_LAST", "OP_SOFTMAX, OP_DEPTHWISE_CONV_2D,  OP_LAST", cpp)
    cpp = sub_one(r"(registrations\[OP_CONV_2D\] = \*tflite::ops::micro::Register_CONV_2D\(\);)",
                  "registrations[OP_CONV_2D] = *tflite::ops::micro::Register_CONV_2D();\n"
                  "  registrations[OP_DEPTHWISE_CONV_2D] = *tflite::ops::micro::Register_DEPTHWISE_CONV_2D();", cpp)
    cpp = sub_one(r"const TfLiteConvParams opdata7 = \{[^;]*\};", "const TfLiteDepthwiseConvParams opdata7 = { kTfLitePaddingSame, 1,1, 1, kTfLiteActNone, 1,1 };", cpp)
    cpp = sub_one(r"(&opdata7\)\), )OP_CONV_2D", "&opdata7)), OP_DEPTHWISE_CONV_2D", cpp)
    # depthwise filter [1, 1, 7, C2], per-channel scales along dimension 3; bias int32[C2]
    cpp = set_data(cpp, 9, "int8_t", f"1*1*7*{C2}", weights(7 * C2))
    cpp = set_dims(cpp, 9, [1, 1, 7, C2])
    w_s = (np.exp(rng.uniform(np.log(0.01), np.log(0.03), C2))).astype(np.float32).astype(np.float64)
    cpp = set_scales(cpp, 9, w_s)
    cpp = sub_one(r"(const TfLiteAffineQuantization quant9 = \{ \(TfLiteFloatArray\*\)&quant9_scale, \(TfLiteIntArray\*\)&quant9_zero, )0( \};)",
                  "const TfLiteAffineQuantization quant9 = { (TfLiteFloatArray*)&quant9_scale, (TfLiteIntArray*)&quant9_zero, 3 };", cpp)
    cpp = set_data(cpp, 8, "int32_t", str(C2), rng.integers(-300, 301, C2))
    cpp = set_dims(cpp, 8, [C2])
    cpp = set_scales(cpp, 8, (w_s * get_scales(cpp, 22)[0]).astype(np.float32).astype(np.float64))
    cpp = set_tensor_bytes(cpp, "tensor_data8", 4 * C2)
    cpp = set_tensor_bytes(cpp, "tensor_data9", 7 * C2)
    # ADD constant of block 2 and the fully connected layer follow the channel count
    cpp = set_data(cpp, 3, "int8_t", str(C2), rng.integers(-127, 128, C2))
    cpp = set_dims(cpp, 3, [C2])
    cpp = set_tensor_bytes(cpp, "tensor_data3", C2)
    cpp = set_data(cpp, 5, "int8_t", f"{n_labels}*{C2}", weights(n_labels * C2))
    cpp = set_dims(cpp, 5, [n_labels, C2])
    cpp = set_tensor_bytes(cpp, "tensor_data5", n_labels * C2)
    cpp = sub_one(r"const ALIGN\(8\) int32_t tensor_data14\[3\] = \{[^;]*\};", f"const ALIGN(8) int32_t tensor_data14[3] = {{ 1, 7, {C2}, }};", cpp)
    cpp = sub_one(r"const ALIGN\(8\) int32_t tensor_data15\[4\] = \{[^;]*\};", f"const ALIGN(8) int32_t tensor_data15[4] = {{ 1, 7, 1, {C2}, }};", cpp)
    for idx, dims in ((23, [1, 1, 7, C2]), (24, [1, 7, C2]), (25, [1, 7, C2]), (26, [1, 7, 1, C2]), (27, [1, 1, 1, C2]), (28, [1, C2])):
        cpp = set_dims(cpp, idx, dims)
    # arena: re-plan every activation tensor sequentially (no aliasing), 16-byte aligned, sizes from the dimensions
    elems = {}
    for m in re.finditer(r"const TfArray<(\d+), int> tensor_dimension(\d+) = \{ \d+, \{ ([^}]*)\} \};", cpp):
        elems[int(m.group(2))] = int(np.prod([int(v) for v in m.group(3).split(",")]))
    off = 0

    def row(m):
        nonlocal off
        dim = int(m.group(2))
        place = f"tensor_arena + {off}"
        off += (elems[dim] + 15) // 16 * 16
        return f"{{ kTfLiteArenaRw, kTfLiteInt8, {place}, (TfLiteIntArray*)&tensor_dimension{dim}, {elems[dim]},"

    cpp, n = re.subn(r"\{ kTfLiteArenaRw, kTfLiteInt8, (tensor_arena \+ \d+), \(TfLiteIntArray\*\)&tensor_dimension(\d+), \d+,", row, cpp)
    assert n == 16, n
    cpp = sub_one(r"constexpr int kTensorArenaSize = \d+;", f"constexpr int kTensorArenaSize = {off + 8192};", cpp)
    for sub, name, text in (("tflite-model", "trained_model_compiled.cpp", cpp), ("tflite-model", "trained_model_compiled.h", hdr),
                            ("model-parameters", "model_metadata.h", meta), ("model-parameters", "dsp_blocks.h", blocks)):
        os.makedirs(os.path.join(out_root, sub), exist_ok=True)
        with open(os.path.join(out_root, sub, name), "w") as f:
            f.write(text)
    print("synthesised depthwise variant under", out_root, "arena", off + 8192)


def main(src_root, out_root):
    rng = np.random.default_rng(SEED)
    n = len(LABELS)
    cpp = open(os.path.join(src_root, "tflite-model", "trained_model_compiled.cpp")).read()
    hdr = open(os.path.join(src_root, "tflite-model", "trained_model_compiled.h")).read()
    meta = open(os.path.join(src_root, "model-parameters", "model_metadata.h")).read()
    blocks = open(os.path.join(src_root, "model-parameters", "dsp_blocks.h")).read()

    def weights(count):
        return np.clip(np.round(rng.normal(0, 35, count)), -127, 127).astype(np.int64)

    # conv1 (tensor 7, per-channel scales 7, bias-scales 6), conv2 (9 / 8), ADD constants 2 and 3
    cpp = set_data(cpp, 7, "int8_t", "30*1*7*13", weights(30 * 7 * 13))
    cpp = set_data(cpp, 9, "int8_t", "10*1*7*30", weights(10 * 7 * 30))
    cpp = set_data(cpp, 2, "int8_t", "30", rng.integers(-127, 128, 30))
    cpp = set_data(cpp, 3, "int8_t", "10", rng.integers(-127, 128, 10))
    in_scale = {7: get_scales(cpp, 16)[0], 9: get_scales(cpp, 22)[0]}
    for w_idx, b_idx in ((7, 6), (9, 8)):
        s = get_scales(cpp, w_idx) * np.exp(rng.uniform(np.log(0.7), np.log(1.4), len(get_scales(cpp, w_idx))))
        s = s.astype(np.float32).astype(np.float64)
        cpp = set_scales(cpp, w_idx, s)
        cpp = set_scales(cpp, b_idx, (s * in_scale[w_idx]).astype(np.float32).astype(np.float64))
    # fully connected: weights [12,10] (tensor 5), bias int32[12] (tensor 4), outputs 29/30 become [1,12]
    cpp = set_data(cpp, 5, "int8_t", f"{n}*10", weights(n * 10))
    cpp = set_dims(cpp, 5, [n, 10])
    cpp = set_data(cpp, 4, "int32_t", str(n), rng.integers(-400, 401, n))
    cpp = set_dims(cpp, 4, [n])
    cpp = set_dims(cpp, 29, [1, n])
    cpp = set_dims(cpp, 30, [1, n])
    cpp = set_tensor_bytes(cpp, "tensor_data4", 4 * n)
    cpp = set_tensor_bytes(cpp, "tensor_data5", 10 * n)
    cpp = set_arena_tensor_bytes(cpp, 29, n)
    cpp = set_arena_tensor_bytes(cpp, 30, n)
    meta = sub_one(r"#define EI_CLASSIFIER_LABEL_COUNT\s+\d+", f"#define EI_CLASSIFIER_LABEL_COUNT                {n}", meta)
    meta = sub_one(r"const char\* ei_classifier_inferencing_categories\[\] = \{[^}]*\};",
                   "const char* ei_classifier_inferencing_categories[] = { " + ", ".join(f'"{l}"' for l in LABELS) + " };", meta)
    for sub, name, text in (("tflite-model", "trained_model_compiled.cpp", cpp), ("tflite-model", "trained_model_compiled.h", hdr),
                            ("model-parameters", "model_metadata.h", meta), ("model-parameters", "dsp_blocks.h", blocks)):
        os.makedirs(os.path.join(out_root, sub), exist_ok=True)
        with open(os.path.join(out_root, sub, name), "w") as f:
            f.write(text)
    print("synthesised", n, "label model under", out_root)


if __name__ == "__main__":
    if len(sys.argv) == 5 and sys.argv[1] == "--float":  # --float <template export (L432 style)> <weights export> <out>
        main_float(sys.argv[2], sys.argv[3], sys.argv[4])
    elif len(sys.argv) == 4 and sys.argv[1] == "--depthwise":
        main_depthwise(sys.argv[2], sys.argv[3])
    elif len(sys.argv) == 3:
        main(sys.argv[1], sys.argv[2])
    else:
        sys.exit(__doc__)
