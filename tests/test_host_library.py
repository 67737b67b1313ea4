"""CPU tests of the product's host side (no GPU, no compute calls): the C-ABI library loads and exports every symbol
include/eikws_b200.h declares, the model container parses, host-side derived tables match the reference-generated
goldens, and failure modes are loud."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol(eikws):
    lib = eikws.load_library()
    header = open(os.path.join(ROOT, "include", "eikws_b200.h")).read()
    names = set(re.findall(r"\b(eikws_[a-z0-9_]+)\s*\(", header))
    names -= {"eikws_get_data_fn"}
    assert len(names) >= 20
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in eikws_b200.h but not exported"


@pytest.mark.parametrize("name", ["l476", "l432", "gsc12"])
def test_host_plan_filterbank_matches_reference(eikws, name):
    fb, mult, shift = eikws.debug_host_plan(name)
    g = np.load(os.path.join(GOLDEN, f"golden_{name}.npz"))
    assert np.array_equal(fb, g["filterbank"])
    n_labels = len(g["labels"])
    assert len(mult) == 30 + 10 + n_labels  # conv1, conv2, fully connected channels
    assert np.all((mult >= (1 << 30)) & (mult < (1 << 31))) and np.all(shift <= 0)


def test_third_topology_is_lowered_by_the_generic_plan(eikws):
    """the Arduino-zip model (conv k3 -> pool 2 (SAME, padded) -> conv k3 -> pool 2 -> FC 208 -> 6): not the fused shape, so
    it exercises the generic op plan built from the captured Register_* graph"""
    fb, mult, shift = eikws.debug_host_plan("zip6")
    g = np.load(os.path.join(GOLDEN, "golden_zip6.npz"))
    assert np.array_equal(fb, g["filterbank"])  # high_frequency 0 -> 8000 Hz, like L432
    assert len(mult) == 8 + 16 + 6 and np.all(shift <= 0)


def test_depthwise_conv_is_lowered_as_a_sparse_dense_conv(eikws):
    """DEPTHWISE_CONV_2D: synthesised variant of the L432 graph; 30 + 30 + 3 requantisation channels"""
    fb, mult, shift = eikws.debug_host_plan("dw3")
    assert len(mult) == 30 + 30 + 3 and np.all((mult >= (1 << 30)) & (mult < (1 << 31)))


def test_float_graph_is_lowered(eikws):
    fb, mult, _ = eikws.debug_host_plan("l476f32")  # BASELINE config 5: float32 twin, no requantisation tables
    assert fb.shape == (129, 32) and len(mult) == 0


def test_malformed_model_is_rejected(eikws):
    lib = eikws.load_library()
    n = C.c_int(0)
    for blob in (b"", b"garbage!" * 8, eikws.model_blob("l476")[:100]):
        rc = lib.eikws_debug_host_plan(blob, len(blob), None, None, None, 0, C.byref(n))
        assert rc == -102 and lib.eikws_last_error()


def test_create_without_gpu_fails_loudly(eikws):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(eikws.EikwsError) as ei:
        eikws.Impulse("l476")
    assert ei.value.code == -101  # EIKWS_ERR_CUDA: no fallback path


def test_unsupported_geometry_is_rejected(eikws):
    """a model whose DSP block is outside the specialised geometry must be refused, not silently mis-computed"""
    import struct
    blob = bytearray(eikws.model_blob("l476"))
    # header: magic(8) version(4) n_tensors n_nodes input output n_labels raw_samples n_features sample_rate num_cepstral ...
    off_fft = 8 + 4 * (1 + 7 + 1 + 4)  # fft_length field
    assert struct.unpack_from("<i", blob, off_fft)[0] == 256
    struct.pack_into("<i", blob, off_fft, 512)
    lib = eikws.load_library()
    n = C.c_int(0)
    rc = lib.eikws_debug_host_plan(bytes(blob), len(blob), None, None, None, 0, C.byref(n))
    assert rc == -100


def test_synth_is_deterministic_and_mixed(synth):
    a, b = synth.synth_clips(40), synth.synth_clips(40)
    assert np.array_equal(a, b) and a.dtype == np.int16 and a.shape == (40, 16000)
    assert np.array_equal(synth.synth_clips(5, first_clip=10), a[10:15])
    big = synth.synth_clips(200, first_clip=100)
    assert (big == 0).all(axis=1).any()                    # silent clips exist in the mixture
    assert (np.abs(big.astype(np.int32)) >= 32767).any()   # saturating clips exist in the mixture


def test_wav_ingest_roundtrip(eikws, synth, tmp_path):
    """PCM_16 / 16 kHz / mono WAV files (what dataset-curation.py writes) -> [n,16000] int16 batches, pad/truncate to 1 s"""
    import eikws_b200.ingest as ingest
    clip = synth.synth_clips(1)[0]
    paths = []
    for i, n in enumerate((16000, 12000, 20000)):
        p = str(tmp_path / f"c{i}.wav")
        ingest.write_wav(p, np.resize(clip, n))
        paths.append(p)
    batch = ingest.read_wav_clips(paths)
    assert batch.shape == (3, 16000) and batch.dtype == np.int16
    assert np.array_equal(batch[0], clip)
    assert np.array_equal(batch[1][:12000], clip[:12000]) and np.all(batch[1][12000:] == 0)
    assert np.array_equal(batch[2], np.resize(clip, 20000)[:16000])
    import wave
    bad = str(tmp_path / "bad.wav")
    with wave.open(bad, "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(44100)
        w.writeframes(b"\0" * 400)
    with pytest.raises(ValueError):
        ingest.read_wav(bad)


def test_cmvn_rounding_trick_equals_the_float_cast(tmp_path):
    """the kernel's two-FMA rounding of the CMVN variance sum (csrc/kernels.cu, cmvn_chains) against the reference's
    (float)(double) cast (numpy.hpp:819-825): tools/check_round_trick.c, millions of operands incl. exact ties and denormals"""
    import subprocess
    exe = str(tmp_path / "check_round_trick")
    subprocess.run(["gcc", "-O2", "-march=native", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tools", "check_round_trick.c"), "-lm"], check=True)
    out = subprocess.run([exe, "30000000"], check=True, capture_output=True, text=True).stdout
    m = re.search(r"checked=(\d+) mismatches=(\d+) exact_ties=(\d+) float_denormal_sums=(\d+)", out)
    assert m and int(m.group(2)) == 0 and int(m.group(1)) > 10_000_000 and int(m.group(3)) > 100_000 and int(m.group(4)) > 10_000, out


def test_tuning_knobs_are_exported_and_validate_their_arguments(eikws):
    """the schedule / lowering knobs tools/ab.py and the parity tests flip (not part of the public header): present, and a null
    handle or an out-of-range value is refused instead of being stored"""
    lib = eikws.load_library()
    for name in ("eikws_set_cmvn_shortcut", "eikws_set_work_claiming", "eikws_set_tensor_core", "eikws_set_clips_per_cta",
                 "eikws_set_ctas_per_sm", "eikws_set_skew_ns", "eikws_classify_taps_i16_device"):
        assert hasattr(lib, name), name
    for name in ("eikws_set_cmvn_shortcut", "eikws_set_work_claiming", "eikws_set_tensor_core", "eikws_set_clips_per_cta", "eikws_set_split",
                 "eikws_set_pipelined", "eikws_set_kernel_timing"):
        fn = getattr(lib, name)
        fn.argtypes = [C.c_void_p, C.c_int]
        fn.restype = C.c_int
        assert fn(None, 1) != 0
    # the kernel-timing query of the two-kernel path: null handle / null output refused
    lib.eikws_split_kernel_ms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.eikws_split_kernel_ms.restype = C.c_int
    ms = (C.c_float * 2)()
    assert lib.eikws_split_kernel_ms(None, ms, None) != 0


def _mutations(blob: bytes):
    """(description, mutated container) pairs: every one must be refused by parse/validate or by the lowering -- never crash.
    Layout (include/eikws_model_format.md): magic(8) version n_tensors n_nodes input output n_labels raw n_features, then the
    11 MFCC fields, labels, tensors, nodes."""
    import struct
    out = []

    def put(off, fmt, v, what):
        b = bytearray(blob)
        struct.pack_into(fmt, b, off, v)
        out.append((what, bytes(b)))

    hdr = 8 + 4 * 8
    put(hdr + 4 * 8, "<i", 40000, "high_frequency above Nyquist (mel_filterbank used to write past its 129-row table)")
    put(hdr + 4 * 8, "<i", 100, "high_frequency below low_frequency")
    put(hdr + 4 * 7, "<i", -5, "negative low_frequency")
    put(hdr + 4 * 0, "<i", 0, "sample rate 0")
    put(8 + 4 * 3, "<I", 4000, "input tensor index out of range")
    put(8 + 4 * 1, "<I", 5000, "implausible tensor count")
    # walk the container to reach the tensors and nodes
    off = hdr + 4 * 11
    n_t, n_n, _, _, n_l = struct.unpack_from("<5I", blob, 12)
    for _ in range(n_l):
        (ln,) = struct.unpack_from("<I", blob, off)
        off += 4 + ((ln + 3) & ~3)
    tens = []
    for _ in range(n_t):
        t0 = off
        ttype, is_const, nd = struct.unpack_from("<3I", blob, off)
        off += 12 + 4 * nd
        (nbytes,) = struct.unpack_from("<I", blob, off)
        bytes_off = off
        (nq,) = struct.unpack_from("<I", blob, off + 4)
        scales_off = off + 8
        off += 8 + 8 * nq + 4
        if is_const:
            off += (nbytes + 3) & ~3
        tens.append(dict(start=t0, type=ttype, is_const=is_const, nd=nd, bytes_off=bytes_off, nq=nq, scales_off=scales_off))
    nodes = []
    for _ in range(n_n):
        n0 = off
        op, n_in = struct.unpack_from("<2I", blob, off)
        in_off = off + 8
        off = in_off + 4 * n_in
        (n_out,) = struct.unpack_from("<I", blob, off)
        out_off = off + 4
        off = out_off + 4 * n_out
        (n_par,) = struct.unpack_from("<I", blob, off)
        par_cnt_off = off
        off += 4 + 4 * n_par
        nodes.append(dict(start=n0, op=op, n_in=n_in, in_off=in_off, n_out=n_out, out_off=out_off, n_par=n_par, par_cnt_off=par_cnt_off))
    assert off == len(blob), "container walked to its end"
    conv = next(n for n in nodes if n["op"] == 3)
    # a CONV_2D with no parameters: the parameter words are cut out of the container
    b = bytearray(blob)
    struct.pack_into("<I", b, conv["par_cnt_off"], 0)
    del b[conv["par_cnt_off"] + 4: conv["par_cnt_off"] + 4 + 4 * conv["n_par"]]
    out.append(("CONV_2D with n_params = 0 (the lowering used to index params[4])", bytes(b)))
    put(conv["in_off"], "<i", -1, "CONV_2D whose input tensor is -1")
    put(conv["in_off"] + 4, "<i", -1, "CONV_2D whose filter tensor is -1")
    put(conv["out_off"], "<i", n_t + 3, "node output index out of range")
    put(conv["par_cnt_off"] + 4 + 4 * 1, "<i", 0, "CONV_2D stride 0")
    pool = next(n for n in nodes if n["op"] == 17)
    put(pool["par_cnt_off"] + 4 + 4 * 2, "<i", 0, "MAX_POOL_2D stride 0")
    act = next(t for t in tens if not t["is_const"] and t["nd"] > 0)
    put(act["start"] + 12, "<i", 0, "activation tensor with a zero dimension")
    put(act["bytes_off"], "<I", 7, "activation tensor whose byte size contradicts its shape")
    qt = next((t for t in tens if t["nq"] >= 1), None)  # (the float32 graph carries no quantisation parameters)
    if qt:
        put(qt["scales_off"], "<f", 0.0, "zero quantisation scale")
        put(qt["scales_off"], "<f", float("nan"), "NaN quantisation scale")
    cst = next(t for t in tens if t["is_const"] and t["type"] in (9, 1) and t["nd"] == 4)
    put(cst["start"] + 12, "<i", 60, "filter whose shape disagrees with its data size")
    put(cst["start"], "<I", 77, "unknown element type")
    out.append(("truncated in the middle of the node table", blob[: nodes[len(nodes) // 2]["start"] + 6]))
    return out


@pytest.mark.parametrize("name", ["l476", "l476f32", "dw3", "zip6"])
def test_mutated_containers_are_refused_not_crashed_on(eikws, name):
    """the container is untrusted input of two public entry points (eikws_create, eikws_debug_host_plan): structural damage must
    come back as an error code from parse_model / validate_model / the lowering, never as an out-of-bounds access (this test runs
    the lowering on every mutation; under ASan/valgrind it is the regression test of the findings in ADVICE.md)"""
    lib = eikws.load_library()
    blob = eikws.model_blob(name)
    n = C.c_int(0)
    assert lib.eikws_debug_host_plan(blob, len(blob), None, None, None, 0, C.byref(n)) == 0
    muts = _mutations(blob)
    assert len(muts) >= 17
    for what, bad in muts:
        rc = lib.eikws_debug_host_plan(bad, len(bad), None, None, None, 0, C.byref(n))
        assert rc in (-102, -100, -1), f"{name}: {what}: rc {rc}"
        assert lib.eikws_last_error(), what
    # random single-word damage anywhere behind the header: any outcome but a crash is acceptable
    rng = np.random.default_rng(1)
    for _ in range(300):
        b = bytearray(blob)
        off = int(rng.integers(12, len(b) - 4))
        b[off:off + 4] = int(rng.choice([0, 0xFFFFFFFF, 0x7FFFFFFF, 0x80000000, 1 << 20, int(rng.integers(0, 1 << 32))])).to_bytes(4, "little")
        lib.eikws_debug_host_plan(bytes(b), len(b), None, None, None, 0, C.byref(n))


def test_multi_device_api_without_gpu_fails_loudly(eikws):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(eikws.EikwsError) as ei:
        eikws.MultiImpulse("l476")
    assert ei.value.code == -101
    lib = eikws.load_library()
    assert lib.eikws_multi_device_count(None) == 0 and lib.eikws_multi_handle(None, 0) is None
    assert lib.eikws_host_alloc(64) is None and b"cudaHostAlloc" in lib.eikws_last_error()


def test_mix_audio_restatement_known_answers():
    """oracle/mix_audio_oracle.py (numpy restatement of dataset-curation.py:93-137 + the PCM_16 write; PARITY UNPINNED, see the
    module) on values worked out by hand: padding, truncation, the float32 background product, round-half-even, 16-bit wrap"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mix_audio_oracle", os.path.join(ROOT, "oracle", "mix_audio_oracle.py"))
    mo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mo)
    bg = np.zeros(40000, np.float32)
    bg[100:16100] = np.float32(0.5)
    x = mo.mix_audio(np.array([1.0, -1.0, 0.25], np.float32), bg, 100, word_vol=1.0, bg_vol=0.1)
    assert x.dtype == np.float64 and x.shape == (16000,)
    b = float(np.float32(0.05) * np.float32(0.5))  # the background term is a float32 product
    assert x[0] == 0.5 + b and x[1] == -0.5 + b and x[2] == 0.125 + b and x[3] == b and x[15999] == b  # zero padding behind the word
    long_word = np.arange(20000, dtype=np.float32) / np.float32(20000)
    assert np.array_equal(mo.mix_audio(long_word, np.zeros(16000, np.float32), 0, 2.0, 1.0), long_word[:16000].astype(np.float64))  # truncation
    assert np.array_equal(mo.mix_audio(None, bg, 100, 3.0, 2.0), np.full(16000, 0.5))  # background-only clip
    pcm = mo.to_pcm16(np.array([0.0, 1.0, -1.0, 0.5 / 32767, 1.5 / 32767, 2.5 / 32767, 2.0, -1.0001]))
    assert pcm.tolist() == [0, 32767, -32767, 0, 2, 2, -2, 32766]  # ties to even; 65534 and -32770 wrap through 16 bits


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the unmodified reference on the host cores, no GPU, nothing of the product loaded): ONE JSON line with
    the base contract's keys, impl = reference, a cpu_baseline describing the run and an e2e object repeating the value"""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-clips-per-core", "128"],
                         check=True, capture_output=True, text=True, timeout=600).stdout.strip().splitlines()
    assert len(out) == 1, out
    d = json.loads(out[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "clips_per_sec" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["gpu_launches"] == 0 and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
