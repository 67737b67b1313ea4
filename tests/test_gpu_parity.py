"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI (libeikws_b200.so), against
(a) the golden vectors generated from the unmodified reference and (b) the plain-C oracle on fresh seeded inputs.
Bar: bit-exact int8 classifier outputs AND bit-identical float MFCC features (north star asks 1e-5; we hold 0)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle_lib import PortOracle

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = ("l476", "l432", "gsc12", "l476f32", "zip6", "dw3")
FLOAT_MODELS = ("l476f32",)
PROB_TOL_F32 = 1e-5  # north-star tolerance for the float32 path (GPU expf vs glibc expf in the softmax)
FEATURE_TOL = 1e-5  # north-star tolerance on the float MFCC coefficients (we additionally assert exact equality)


def golden(name):
    return np.load(os.path.join(GOLDEN, f"golden_{name}.npz"))


def golden_clips(synth, g):
    sp = synth.special_clips()
    return np.concatenate([synth.synth_clips(int(g["n_synth"]), 0, int(g["seed"])), np.stack([sp[str(k)] for k in g["special_names"]]),
                           synth.speechlike_clips(g["speech_params"])])


def same_floats(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


@pytest.fixture(scope="module")
def impulses(eikws):
    d = {m: eikws.Impulse(m, device=0) for m in MODELS}
    yield d
    for v in d.values():
        v.close()


@pytest.mark.parametrize("name", MODELS)
def test_run_classifier_matches_reference_golden(name, impulses, synth):
    g = golden(name)
    imp = impulses[name]
    assert imp.labels == [str(s) for s in g["labels"]]
    clips = golden_clips(synth, g)
    if name in FLOAT_MODELS:
        probs, feats = imp.run_classifier(clips), imp.extract_mfcc_features(clips)
    else:
        probs, feats, q = imp.run_classifier_taps(clips)
    assert np.nanmax(np.abs(feats - g["features"])) <= FEATURE_TOL
    bad = np.where(~((feats == g["features"]) | (np.isnan(feats) & np.isnan(g["features"]))).all(axis=1))[0]
    assert bad.size == 0, f"clips with non-identical features: {bad[:10]}"
    if name in FLOAT_MODELS:
        assert np.max(np.abs(probs - g["probs"])) <= PROB_TOL_F32
    else:
        assert np.array_equal(probs, g["probs"])


@pytest.mark.parametrize("name", MODELS)
def test_features_only_and_float_input(name, impulses, synth):
    g = golden(name)
    imp = impulses[name]
    clips = golden_clips(synth, g)
    feats = imp.extract_mfcc_features(clips)
    assert same_floats(feats, g["features"])
    x = clips[:8].astype(np.float32) / np.float32(32768)
    assert same_floats(imp.extract_mfcc_features(x), g["features_f32in"])
    if name in FLOAT_MODELS:
        assert np.max(np.abs(imp.run_classifier(x) - g["probs"][:8])) <= PROB_TOL_F32
    else:
        assert np.array_equal(imp.run_classifier(x), g["probs"][:8])


@pytest.mark.parametrize("name", MODELS)
def test_int8_classifier_matches_reference_golden(name, impulses):
    """run_inference alone on crafted feature vectors (saturating, wrapping the float->int8 cast, inf/nan)"""
    g = golden(name)
    if name in FLOAT_MODELS:  # finite rows only: NaN/inf propagate differently through max() on CPU and GPU
        rows = np.r_[0:64, 72:96]
        probs = impulses[name].run_inference(g["nn_features"][rows])
        assert np.max(np.abs(probs - g["nn_probs"][rows])) <= PROB_TOL_F32
        return
    probs = impulses[name].run_inference(g["nn_features"])
    assert np.array_equal(probs, g["nn_probs"])


@pytest.mark.parametrize("name", MODELS)
def test_quantised_input_matches_oracle(name, impulses, synth):
    if name in FLOAT_MODELS:
        pytest.skip("a float32 graph has no quantised input")
    g = golden(name)
    imp = impulses[name]
    clips = golden_clips(synth, g)
    _, q = imp.extract_mfcc_features(clips, quantized=True)
    port = PortOracle(name)
    _, tens = port.run_inference(g["features"], want_tensors=True)
    want = np.stack([t[0] for t in tens]).view(np.int8)  # tensor 0 = quantised NN input
    assert np.array_equal(q, want)


@pytest.mark.parametrize("name", MODELS)
def test_fresh_clips_against_plain_c_oracle(name, impulses, synth):
    imp = impulses[name]
    port = PortOracle(name)
    clips = synth.synth_clips(256, first_clip=5000, seed=0xC0FFEE)
    if name in FLOAT_MODELS:
        probs, feats = imp.run_classifier(clips), imp.extract_mfcc_features(clips)
    else:
        probs, feats, _ = imp.run_classifier_taps(clips)
    want_p, want_f = port.run_classifier_i16(clips, want_features=True)
    assert same_floats(feats, want_f)
    if name in FLOAT_MODELS:
        assert np.max(np.abs(probs - want_p)) <= PROB_TOL_F32
    else:
        assert np.array_equal(probs, want_p)


def test_device_path_equals_host_path_and_synth_twin(impulses, synth):
    import torch
    imp = impulses["l476"]
    n = 1000  # not a multiple of the grid: exercises the ragged tail of the persistent loop
    d_clips = imp.synth_clips_device(n, first_clip=123, seed=0xE1D5)
    torch.cuda.synchronize()
    h_clips = d_clips.cpu().numpy()
    assert np.array_equal(h_clips[:64], synth.synth_clips(64, first_clip=123, seed=0xE1D5))
    d_probs = imp.run_classifier_device(d_clips)
    torch.cuda.synchronize()
    assert np.array_equal(d_probs.cpu().numpy(), imp.run_classifier(h_clips))
    feats = imp.extract_mfcc_features_device(d_clips)
    d_probs2 = imp.run_inference_device(feats)
    torch.cuda.synchronize()
    assert torch.equal(d_probs, d_probs2)  # fused == MFCC kernel + inference kernel


def test_mfe_block_matches_reference(impulses, synth):
    """the sibling MFE DSP block (extract_mfe_features, newer SDK copy): bit-identical to the unmodified reference (golden,
    L432 band) and to the C oracle on fresh clips for both bands (L476: 300-4000 Hz, L432: 300-8000 Hz); int16, float32 and
    device entry points; NaN rows of degenerate clips (max == min) included"""
    import torch
    g = golden("l432")
    clips = golden_clips(synth, g)
    got = impulses["l432"].extract_mfe_features(clips)
    assert same_floats(got, g["mfe_features"])
    fresh = synth.synth_clips(300, first_clip=9000, seed=0xFEED)
    for name in ("l476", "l432"):
        imp = impulses[name]
        want = PortOracle(name).mfe_block_i16(fresh)
        assert same_floats(imp.extract_mfe_features(fresh), want)
        xf = fresh.astype(np.float32) / np.float32(32768)
        assert same_floats(imp.extract_mfe_features(xf), want)
        d = torch.from_numpy(fresh).to("cuda:0")
        assert same_floats(imp.extract_mfe_features_device(d).cpu().numpy(), want)


@pytest.mark.parametrize("name", ["l476", "zip6", "l476f32"])
def test_ragged_batch_sizes(name, impulses):
    """batches that do not fill the persistent grid or the clip groups of a CTA: 1, 2, 3 clips (one group idle), odd sizes
    around one full wave (148 SMs x 4 groups = 592), sizes that leave the last CTA half empty"""
    import torch
    imp = impulses[name]
    d = imp.synth_clips_device(1300, first_clip=77, seed=0xE1D5)
    full = imp.run_classifier_device(d).clone()
    torch.cuda.synchronize()
    port = PortOracle(name)
    idx = [0, 1, 2, 590, 591, 592, 593, 1183, 1184, 1185, 1299]
    want = port.run_classifier_i16(d[idx].cpu().numpy())
    got = full[idx].cpu().numpy()
    assert np.array_equal(got, want) if name not in FLOAT_MODELS else np.max(np.abs(got - want)) <= PROB_TOL_F32
    for n in (1, 2, 3, 5, 591, 592, 593, 1183, 1185):
        out = imp.run_classifier_device(d[:n].contiguous())
        assert torch.equal(out, full[:n]), f"n={n}"
        off = 1300 - n  # the same clips at another position of another batch
        out = imp.run_classifier_device(d[off:].contiguous())
        assert torch.equal(out, full[off:]), f"tail n={n}"
    if name == "l476":  # host-buffer entry point: 8192-clip chunks on two streams, the last chunk odd-sized
        big = imp.synth_clips_device(8192 + 593, first_clip=77, seed=0xE1D5)
        want = imp.run_classifier_device(big).cpu().numpy()
        assert np.array_equal(imp.run_classifier(big.cpu().numpy()), want)
        assert np.array_equal(want[:1300], full.cpu().numpy())


def test_batch_properties_at_full_size(impulses, synth):
    """BASELINE config 2 size (65,536 clips): size-independent properties -- permutation/sharding invariance
    (a clip's result does not depend on its position or on its neighbours), duplicates agree, probabilities are
    valid int8-softmax outputs, and a subsample agrees with the oracle."""
    import torch
    imp = impulses["l476"]
    n = 65536
    d = imp.synth_clips_device(n, first_clip=0, seed=0xE1D5)
    p = imp.run_classifier_device(d)
    torch.cuda.synchronize()
    # sharded 8 ways (what 8 GPUs would each see) == unsharded
    shard = n // 8
    for s in (0, 3, 7):
        ps = imp.run_classifier_device(d[s * shard:(s + 1) * shard].contiguous())
        assert torch.equal(ps, p[s * shard:(s + 1) * shard])
    perm = torch.randperm(n, device=d.device, generator=torch.Generator(device=d.device).manual_seed(1))
    pp = imp.run_classifier_device(d[perm].contiguous())
    assert torch.equal(pp, p[perm])
    ph = p.cpu().numpy()
    assert np.all((ph >= 0) & (ph <= 255 / 256)) and np.all(ph * 256 == np.round(ph * 256))
    assert np.all(np.abs(ph.sum(axis=1) - 1.0) <= 4 / 256 + 1e-6)
    idx = np.arange(0, n, 1024)
    port = PortOracle("l476")
    assert np.array_equal(ph[idx], port.run_classifier_i16(d[torch.from_numpy(idx).to(d.device)].cpu().numpy()))


def test_error_behaviour(eikws, impulses):
    import torch
    imp = impulses["l476"]
    lib = eikws.load_library()
    d = torch.zeros(16000 * 2 + 8, dtype=torch.int16, device="cuda:0")
    out = torch.zeros(4, dtype=torch.float32, device="cuda:0")
    # misaligned clip pointer: refused (TMA bulk copies need 16-byte alignment), nothing launched
    rc = lib.eikws_classify_i16_device(imp._h, C.c_void_p(d.data_ptr() + 2), 1, C.c_void_p(out.data_ptr()), None)
    assert rc == -102
    # empty batch is a no-op
    assert lib.eikws_classify_i16_device(imp._h, C.c_void_p(d.data_ptr()), 0, C.c_void_p(out.data_ptr()), None) == 0
    # single-clip pull API: wrong signal length -> EI_IMPULSE_DSP_ERROR like ei_run_dsp.h:279-283
    CB = C.CFUNCTYPE(C.c_int, C.c_size_t, C.c_size_t, C.POINTER(C.c_float))

    def get_data(off, length, outp):
        for i in range(length):
            outp[i] = 0.0
        return 0

    vals = (C.c_float * 4)()
    lib.eikws_run_classifier_signal.argtypes = [C.c_void_p, CB, C.c_size_t, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    assert lib.eikws_run_classifier_signal(imp._h, CB(get_data), 16320, vals, None, None) == -5
    assert lib.eikws_run_classifier_signal(imp._h, CB(get_data), 16000, vals, None, None) == 0
    g = golden("l476")
    sil = [str(k) for k in g["special_names"]].index("silence") + int(g["n_synth"])
    assert np.array_equal(np.array(vals[:], np.float32), g["probs"][sil])


@pytest.mark.parametrize("name", ["l476", "l432", "gsc12", "zip6"])
def test_continuous_mode_matches_oracle(name, eikws, impulses, synth):
    """run_classifier_continuous for 37 concurrent streams x 14 slices against the plain-C oracle, stream by stream"""
    from oracle_lib import PortStream
    imp = impulses[name]
    n_streams, n_slices = 37, 14
    streams = eikws.Streams(imp, n_streams)
    assert streams.slice_size == 4000
    audio = synth.synth_clips(n_streams * 4, first_clip=900).reshape(n_streams, -1)[:, : n_slices * 4000]
    port = PortOracle(name)
    oracles = [PortStream(port) for _ in range(n_streams)]
    for i in range(n_slices):
        sl = audio[:, i * 4000:(i + 1) * 4000]
        got = streams.push(sl)
        want = [o.push(sl[k]) for k, o in enumerate(oracles)]
        assert (got is None) == (want[0] is None)
        if got is not None:
            assert np.array_equal(got, np.stack(want)), f"slice {i}"
    # reset == power-up: the same audio gives the same answers again
    streams.reset()
    again = [streams.push(audio[:, i * 4000:(i + 1) * 4000]) for i in range(5)]
    oracles = [PortStream(port) for _ in range(n_streams)]
    want5 = [[o.push(audio[k, i * 4000:(i + 1) * 4000]) for k, o in enumerate(oracles)] for i in range(5)]
    assert again[2] is None and np.array_equal(again[4], np.stack(want5[4]))
    streams.close()


def test_i2s_decimation_matches_firmware_isr(impulses):
    """Core/Src/main.cpp:507-521: every 4th 32-bit SAI word, top 16 of 24 bits"""
    import torch
    imp = impulses["l476"]
    g = torch.Generator(device="cuda:0").manual_seed(5)
    for n_out in (16000 * 33 + 5, 7, 8):
        i2s = torch.randint(-(1 << 23), 1 << 23, (4 * n_out + 3,), dtype=torch.int32, device="cuda:0", generator=g)
        got = imp.decimate_i2s_device(i2s, n_out)
        want = (i2s[: 4 * n_out : 4] >> 8).to(torch.int16)
        torch.cuda.synchronize()
        assert torch.equal(got, want)


@pytest.mark.parametrize("name", ["l476", "l432", "gsc12", "dw3"])
def test_tensor_core_block1_is_bit_identical(name, impulses, synth):
    """block 1 of the fused int8 classifier as a tcgen05.mma.kind::i8 over the in-place sliding windows of the quantised
    feature matrix (kernels.cu, use_tc): integer sums, so the outputs must equal the dp4a path's and the C oracle's byte for
    byte -- goldens (incl. the saturating / wrapping edge-case clips), fresh clips, ragged batch sizes (a clip group idle)"""
    import torch
    imp = impulses[name]
    g = golden(name)
    clips = golden_clips(synth, g)
    d = imp.synth_clips_device(1300, first_clip=4242, seed=0xBEEF)
    try:
        imp.set_tensor_core(False)
        base = imp.run_classifier_device(d).clone()
        imp.set_tensor_core(True)
        before = imp.launch_count
        assert np.array_equal(imp.run_classifier(clips), g["probs"])
        tc = imp.run_classifier_device(d).clone()
        assert torch.equal(tc, base)
        idx = [0, 1, 2, 591, 592, 593, 1299]
        assert np.array_equal(tc[idx].cpu().numpy(), PortOracle(name).run_classifier_i16(d[idx].cpu().numpy()))
        for n in (1, 2, 3, 5, 591, 593, 1185):
            assert torch.equal(imp.run_classifier_device(d[:n].contiguous()), base[:n]), f"n={n}"
            assert torch.equal(imp.run_classifier_device(d[1300 - n:].contiguous()), base[1300 - n:]), f"tail n={n}"
        probs, feats, qfeats = imp.run_classifier_taps(clips)  # features / quantised features come out of the same launch
        assert same_floats(feats, g["features"]) and np.array_equal(probs, g["probs"])
        assert imp.launch_count > before
    finally:
        imp.set_tensor_core(True)  # the default


@pytest.mark.parametrize("name", ["l476", "l432", "gsc12", "dw3", "zip6"])
def test_certified_cmvn_shortcut_is_bit_identical(name, impulses, synth):
    """The default classify kernel decides round(f / scale) of most CMVN outputs from double-precision window statistics plus
    a rigorous error bound and runs the reference's operation sequence only for the chains the bound cannot certify
    (kernels.cu, cmvn_certified; derivation in DESIGN.md).  The int8 classifier input and the outputs must equal, byte for
    byte, those of the kernel that runs every chain with the reference's sequence: goldens (incl. silence / DC / impulse
    clips, where every chain takes the exact path), the oracle's input tensor, 65,536 fresh clips, ragged batch sizes."""
    import torch
    imp = impulses[name]
    g = golden(name)
    clips = golden_clips(synth, g)
    port = PortOracle(name)
    _, tens = port.run_inference(g["features"], want_tensors=True)
    want_q = np.stack([t[0] for t in tens]).view(np.int8)
    try:
        imp.set_cmvn_shortcut(True)
        d_clips = torch.from_numpy(clips).to("cuda:0")
        probs, q = imp.run_classifier_taps_device(d_clips)
        torch.cuda.synchronize()
        assert np.array_equal(q.cpu().numpy(), want_q)
        assert np.array_equal(probs.cpu().numpy(), g["probs"])
        assert np.array_equal(imp.run_classifier(clips), g["probs"])
        n = 65536
        d = imp.synth_clips_device(n, first_clip=777000, seed=0x5EED)
        p1, q1 = imp.run_classifier_taps_device(d)
        p1b = imp.run_classifier_device(d)
        imp.set_cmvn_shortcut(False)
        p0, q0 = imp.run_classifier_taps_device(d)
        torch.cuda.synchronize()
        bad = (q1 != q0).any(dim=1).nonzero().flatten()
        assert bad.numel() == 0, f"clips whose quantised features differ: {bad[:10].tolist()}"
        assert torch.equal(p1, p0) and torch.equal(p1b, p0)
        imp.set_cmvn_shortcut(True)
        for m in (1, 2, 3, 5, 591, 593, 1185):
            assert torch.equal(imp.run_classifier_device(d[:m].contiguous()), p0[:m]), f"n={m}"
            assert torch.equal(imp.run_classifier_device(d[n - m:].contiguous()), p0[n - m:]), f"tail n={m}"
        idx = [0, 1, 2, 3, 40000, 65535]
        assert np.array_equal(p1[idx].cpu().numpy(), port.run_classifier_i16(d[idx].cpu().numpy()))
        # the same shortcut on the dp4a lowering (tensor core off) and on float-input clips (the samples a demo callback delivers)
        imp.set_tensor_core(False)
        assert torch.equal(imp.run_classifier_device(d), p0)
        imp.set_clips_per_cta(1)
        assert torch.equal(imp.run_classifier_device(d[:4099].contiguous()), p0[:4099])
        imp.set_clips_per_cta(2)
        imp.set_tensor_core(True)
        x = (d[:8192].to(torch.float32) / 32768.0).contiguous()
        assert torch.equal(imp.run_classifier_device(x), p0[:8192])
        assert np.array_equal(imp.run_classifier(clips.astype(np.float32) / np.float32(32768)), g["probs"])
        # float samples no int16 source produces (inf, NaN, huge, denormal-small): the bound must refuse, never mis-certify
        bad = x[:64].clone()
        for i, v in enumerate([float("inf"), float("-inf"), float("nan"), 1e30, -3e38, 1e-30, 1e-42, 65504.0]):
            bad[i, 100 + 37 * i] = v
            bad[8 + i, 320 * (3 + 5 * i):320 * (4 + 5 * i)] = v
            bad[16 + i] *= v if np.isfinite(v) else 1.0
        with_shortcut = imp.run_classifier_device(bad).clone()
        imp.set_cmvn_shortcut(False)
        without = imp.run_classifier_device(bad).clone()
        imp.set_cmvn_shortcut(True)
        torch.cuda.synchronize()
        assert torch.equal(with_shortcut, without)
    finally:
        imp.set_cmvn_shortcut(True)
        imp.set_tensor_core(True)
        imp.set_clips_per_cta(2)


@pytest.mark.parametrize("name", ["l476", "gsc12"])
def test_device_cmvn_shortcut_on_adversarial_cepstra(name, impulses):
    """The DEVICE implementation of the certified CMVN shortcut (cmvn_certified / cmvn_resolve, MUFU sqrt.approx / rcp.approx
    inside) fed directly with pre-CMVN cepstra through eikws_debug_cmvn_quantise_host -- matrices no audio clip produces:
    the 72 hand-shaped adversarial matrices of test_cmvn_bound.py, denormal / huge / inf / NaN columns, and > 10^6 random
    matrices in which one frame per column was moved ONTO a rounding boundary k + 1/2 of the int8 quantisation (to within the
    reference's own rounding errors: the only place where a nearly-right shortcut flips a bit).  Bar: byte equality with the
    every-chain kernel on all of them, and with the CPU oracle's CMVN (pinned to the unmodified reference's cmvnw on these very
    matrix families, test_cmvn_bound.py) on every hand-shaped matrix and a 12,288-matrix subset.
    reference: processing.hpp:326-389, numpy.hpp:767-825, ei_run_classifier.h:436-444"""
    import cmvn_cases
    imp = impulses[name]
    port = PortOracle(name)
    scale = {"l476": 0.046360891312360764}.get(name)
    if scale is None:  # the input scale of the model, from its container (tensor 0 of the int8 graph is the input)
        _, t0 = port.run_inference(np.zeros((1, 637), np.float32), want_tensors=True)
        zp = int(t0[0][0].view(np.int8)[0])
        probe = np.zeros((1, 637), np.float32)
        lo, hi = 0.0, 64.0
        for _ in range(60):  # f with round(f / scale) == 40 |-> bisect the 39.5 boundary
            mid = 0.5 * (lo + hi)
            probe[0, 0] = mid
            _, t = port.run_inference(probe, want_tensors=True)
            if int(t[0][0].view(np.int8)[0]) - zp >= 40:
                hi = mid
            else:
                lo = mid
        scale = hi / 39.5
    hand = np.stack(cmvn_cases.adversarial_matrices(scale) + cmvn_cases.special_value_matrices())
    with np.errstate(all="ignore"):
        want = port.cmvn_quantise(hand)
    q_short = imp.debug_cmvn_quantise(hand, shortcut=True)
    q_exact = imp.debug_cmvn_quantise(hand, shortcut=False)
    for i in range(len(hand)):
        assert np.array_equal(q_exact[i], want[i]), f"hand-shaped matrix {i}: every-chain kernel differs from the oracle"
        assert np.array_equal(q_short[i], want[i]), f"hand-shaped matrix {i}: certified shortcut differs from the oracle"
    total = checked_cpu = near = 0
    chunk = 65536
    for c in range(16 if name == "l476" else 2):  # 1,048,576 matrices (13.6 M targeted chains) for the headline model
        F, r = cmvn_cases.near_boundary_matrices(chunk, scale, seed=1000 + c)
        qs = imp.debug_cmvn_quantise(F, shortcut=True)
        qe = imp.debug_cmvn_quantise(F, shortcut=False)
        bad = np.nonzero((qs != qe).any(axis=1))[0]
        assert bad.size == 0, f"chunk {c}: shortcut and every-chain kernels disagree on matrices {bad[:8].tolist()}"
        sub = slice(0, 768)
        with np.errstate(all="ignore"):
            w, f = port.cmvn_quantise(F[sub], want_features=True)
        assert np.array_equal(qs[sub], w), f"chunk {c}: shortcut differs from the CPU oracle"
        t = (f.reshape(-1, 49, 13) / np.float32(scale))[np.arange(768)[:, None], r[sub], np.arange(13)[None, :]].astype(np.float64)
        near += int((np.abs(t - np.floor(t) - 0.5) < 1e-4).sum())
        checked_cpu += 768
        total += chunk
    assert total >= (1 << 20 if name == "l476" else 1 << 17) and checked_cpu >= (12288 if name == "l476" else 1536)
    assert near > 0.3 * checked_cpu * 13  # the targeted chains really sit at the boundaries


def test_multi_device_host_api_equals_one_device(eikws, impulses, synth):
    """eikws_multi_*: contiguous shards over every visible GPU, one host thread + stream pair per device, results in place ==
    the single-device result, byte for byte (ragged batch sizes included).  With one visible GPU the set has one device and the
    test still covers the thread/shard plumbing; the 2-, 4-, 8-way split runs where the box has the GPUs."""
    import torch
    multi = eikws.MultiImpulse("l476")
    try:
        assert multi.device_count == torch.cuda.device_count()
        for n in (1, 7, 4097, 20000):
            clips = synth.synth_clips(min(n, 512), first_clip=555)
            clips = np.resize(clips, (n, 16000)) if n > 512 else clips[:n]
            want = impulses["l476"].run_classifier(clips)
            got = multi.run_classifier(clips)
            assert np.array_equal(got, want), f"n={n}"
            covered = sum(multi.shard(n, i)[1] for i in range(multi.device_count))
            assert covered == n and multi.shard(n, 0)[0] == 0
        # the same through float samples
        x = synth.synth_clips(33, first_clip=9).astype(np.float32) / np.float32(32768)
        assert np.array_equal(multi.run_classifier(x), impulses["l476"].run_classifier(x))
    finally:
        multi.close()
    if torch.cuda.device_count() >= 2:  # an explicit device list, in reverse order
        m2 = eikws.MultiImpulse("l476", devices=[1, 0])
        try:
            clips = synth.synth_clips(301, first_clip=1234)
            assert np.array_equal(m2.run_classifier(clips), impulses["l476"].run_classifier(clips))
        finally:
            m2.close()
    lib = eikws.load_library()
    h = C.c_void_p()
    blob = eikws.model_blob("l476")
    arr = (C.c_int * 2)(0, 0)
    assert lib.eikws_multi_create(blob, len(blob), arr, 2, C.byref(h)) == -102  # a device listed twice
    arr = (C.c_int * 1)(99)
    assert lib.eikws_multi_create(blob, len(blob), arr, 1, C.byref(h)) == -102  # no such device


def test_single_clip_calls_from_many_threads_do_not_mix_clips(eikws, impulses, synth):
    """eikws_run_classifier_signal holds the handle's lock from the callback's first write into the shared pinned staging buffer
    to the last result byte: 8 threads x 40 calls, each thread with its own clip, every result must be its own clip's"""
    import threading
    lib = eikws.load_library()
    imp = impulses["l476"]
    clips = synth.synth_clips(8, first_clip=31)
    want = imp.run_classifier(clips)
    CB = C.CFUNCTYPE(C.c_int, C.c_size_t, C.c_size_t, C.POINTER(C.c_float))
    lib.eikws_run_classifier_signal.argtypes = [C.c_void_p, CB, C.c_size_t, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    errors = []

    def worker(k):
        x = clips[k].astype(np.float32) / np.float32(32768)

        def get_data(off, length, out):
            C.memmove(out, x[off:off + length].ctypes.data, 4 * length)
            return 0

        cb = CB(get_data)
        vals = (C.c_float * imp.label_count)()
        for _ in range(40):
            rc = lib.eikws_run_classifier_signal(imp._h, cb, 16000, vals, None, None)
            if rc != 0 or not np.array_equal(np.frombuffer(vals, np.float32), want[k]):
                errors.append((k, rc))
                return

    ts = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


def test_mix_audio_matches_the_numpy_restatement(impulses):
    """the dataset tooling's mix_audio + PCM_16 write (dataset-curation.py:93-137, 190-206) as one streaming kernel, against
    oracle/mix_audio_oracle.py (PARITY UNPINNED: librosa / soundfile are absent, see that module): words shorter and longer than a
    second, odd lengths, background-only clips, volumes that overflow 16 bits (libsndfile wraps), exact .5 ties of the rounding"""
    import importlib.util
    import torch
    spec = importlib.util.spec_from_file_location("mix_audio_oracle", os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "mix_audio_oracle.py"))
    mo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mo)
    imp = impulses["l476"]
    rng = np.random.default_rng(77)
    n, stride = 257, 20000
    words = (rng.standard_normal((n, stride)) * 0.3).astype(np.float32)
    lens = rng.integers(0, stride + 1, n).astype(np.uint32)
    lens[:6] = [0, 1, 15999, 16000, 16001, 20000]
    words[6, :16000] = (np.arange(16000) - 8000).astype(np.float32) / np.float32(32767.0)  # x * 32767 lands on integers and .5 ties
    lens[6] = 16000
    bg = (rng.standard_normal(16000 * 40) * 0.2).astype(np.float32)
    bg[:16000] = 0.0
    starts = rng.integers(0, len(bg) - 16000 + 1, n).astype(np.uint32)
    starts[6] = 0
    starts[7] = len(bg) - 16000
    d_words, d_lens = torch.from_numpy(words).cuda(), torch.from_numpy(lens.view(np.int32)).cuda()
    d_bg, d_starts = torch.from_numpy(bg).cuda(), torch.from_numpy(starts.view(np.int32)).cuda()
    for wv, bv in ((1.0, 0.1), (0.7, 1.0), (9.0, 3.0), (2.0, 0.0)):
        got = imp.mix_audio_device(d_words, d_lens, d_bg, d_starts, wv, bv).cpu().numpy()
        for c in range(n):
            want = mo.to_pcm16(mo.mix_audio(words[c, :lens[c]], bg, int(starts[c]), wv, bv))
            assert np.array_equal(got[c], want), f"clip {c} word_vol {wv} bg_vol {bv}"
    got = imp.mix_audio_device(None, None, d_bg, d_starts, 1.0, 0.25).cpu().numpy()  # the script's _noise clips
    for c in (0, 7, 100):
        assert np.array_equal(got[c], mo.to_pcm16(mo.mix_audio(None, bg, int(starts[c]), 1.0, 0.25)))
    # the mixed clips feed the classifier like any other int16 batch
    assert imp.run_classifier_device(imp.mix_audio_device(d_words, d_lens, d_bg, d_starts, 1.0, 0.1)).shape == (n, imp.label_count)
    # throughput of the streaming kernel (8 B read + 2 B written per sample), for DESIGN.md
    big = 8192
    w2 = torch.randn((big, 16000), device="cuda") * 0.3
    l2 = torch.full((big,), 16000, dtype=torch.int32, device="cuda")
    s2 = torch.randint(0, len(bg) - 16000, (big,), dtype=torch.int32, device="cuda")
    imp.mix_audio_device(w2, l2, d_bg, s2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        imp.mix_audio_device(w2, l2, d_bg, s2)
    e1.record()
    torch.cuda.synchronize()
    gbs = 10 * big * 16000 * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    print(f"mix_audio: {gbs:.0f} GB/s algorithmic ({10 * big / (e0.elapsed_time(e1) * 1e-3) / 1e6:.1f} M clips/s)")
    assert gbs > 500


@pytest.mark.parametrize("name", ["l476", "l432", "gsc12", "dw3"])
def test_pipelined_kernel_is_bit_identical(name, impulses, synth):
    """eikws_pipelined_kernel (every warp interleaves the FFT of clip s with the post-FFT slices of clip s-1; mbarrier-linked slices,
    one CTA barrier per clip) against the goldens of the unmodified reference, the oracle's int8 input tensor, the phase-by-phase
    kernel on 65,536 fresh clips, and ragged batch sizes from one clip up (short CTAs: 1, 2, 3 clips per CTA and uneven tails)"""
    import torch
    imp = impulses[name]
    g = golden(name)
    clips = golden_clips(synth, g)
    port = PortOracle(name)
    _, tens = port.run_inference(g["features"], want_tensors=True)
    want_q = np.stack([t[0] for t in tens]).view(np.int8)
    n = 65536
    d = imp.synth_clips_device(n, first_clip=424242, seed=0xBEEF)
    p_ref, q_ref = imp.run_classifier_taps_device(d)
    try:
        imp.set_pipelined(True)
        before = imp.launch_count
        probs, q = imp.run_classifier_taps_device(torch.from_numpy(clips).to("cuda:0"))
        torch.cuda.synchronize()
        assert np.array_equal(q.cpu().numpy(), want_q)
        assert np.array_equal(probs.cpu().numpy(), g["probs"])
        assert np.array_equal(imp.run_classifier(clips), g["probs"])
        p1, q1 = imp.run_classifier_taps_device(d)
        p1b = imp.run_classifier_device(d)
        torch.cuda.synchronize()
        bad = (q1 != q_ref).any(dim=1).nonzero().flatten()
        assert bad.numel() == 0, f"clips whose quantised features differ: {bad[:10].tolist()}"
        assert torch.equal(p1, p_ref) and torch.equal(p1b, p_ref)
        for m in (1, 2, 3, 5, 295, 296, 297, 591, 593, 887, 889, 1185, 4099):
            assert torch.equal(imp.run_classifier_device(d[:m].contiguous()), p_ref[:m]), f"n={m}"
            assert torch.equal(imp.run_classifier_device(d[n - m:].contiguous()), p_ref[n - m:]), f"tail n={m}"
        assert imp.launch_count > before
    finally:
        imp.set_pipelined(False)


@pytest.mark.parametrize("name", ["l476", "l432", "gsc12", "dw3"])
def test_split_kernels_are_bit_identical(name, impulses, synth):
    """the two-kernel classify path (eikws_logmel_kernel: a warp per eight frames of the batch's frame sequence, no CTA barrier;
    eikws_cepstral_kernel: DCT / certified CMVN / int8 CNN of one clip per CTA) against the goldens of the unmodified reference, the
    oracle's int8 input tensor, the fused phase-by-phase kernel on 65,536 fresh clips, and ragged batch sizes from one clip up (work
    units that straddle clips, units beyond the batch, CTAs with 0 / 1 / 2 clips, a chunk boundary of the hand-over scratch)"""
    import torch
    imp = impulses[name]
    g = golden(name)
    clips = golden_clips(synth, g)
    port = PortOracle(name)
    _, tens = port.run_inference(g["features"], want_tensors=True)
    want_q = np.stack([t[0] for t in tens]).view(np.int8)
    n = 65536 + 4099  # more than one chunk of the scratch
    d = imp.synth_clips_device(n, first_clip=515151, seed=0xFACE)
    try:
        imp.set_split(False)
        p_ref, q_ref = imp.run_classifier_taps_device(d)
        imp.set_split(True)
        before = imp.launch_count
        probs, q = imp.run_classifier_taps_device(torch.from_numpy(clips).to("cuda:0"))
        torch.cuda.synchronize()
        assert imp.launch_count == before + 2, "the split path launches two kernels per chunk"
        assert np.array_equal(q.cpu().numpy(), want_q)
        assert np.array_equal(probs.cpu().numpy(), g["probs"])
        assert np.array_equal(imp.run_classifier(clips), g["probs"])
        p1, q1 = imp.run_classifier_taps_device(d)
        p1b = imp.run_classifier_device(d)
        torch.cuda.synchronize()
        bad = (q1 != q_ref).any(dim=1).nonzero().flatten()
        assert bad.numel() == 0, f"clips whose quantised features differ: {bad[:10].tolist()}"
        assert torch.equal(p1, p_ref) and torch.equal(p1b, p_ref)
        for m in (1, 2, 3, 5, 7, 8, 9, 163, 739, 740, 741, 1479, 1481, 2961, 4099):
            assert torch.equal(imp.run_classifier_device(d[:m].contiguous()), p_ref[:m]), f"n={m}"
            assert torch.equal(imp.run_classifier_device(d[n - m:].contiguous()), p_ref[n - m:]), f"tail n={m}"
    finally:
        imp.set_split(True)


@pytest.mark.parametrize("name,f32_input", [("l476", True), ("l476f32", False), ("l476f32", True), ("gsc12", True)])
def test_split_kernels_float_clips_and_float_graph(name, f32_input, impulses, synth):
    """the two-kernel path with float32 clips (spectral kernel on four-frame units, 65 chunks per frame) and / or the float32 graph
    (eikws_cepstral_f32_kernel: the reference's CMVN chains + the float op plan): goldens of the unmodified reference, then the
    fused kernel on fresh clips -- the two paths run the same device functions in the same order, so even the float probabilities
    must be bit-identical -- and ragged batch sizes"""
    import torch
    imp = impulses[name]
    g = golden(name)
    clips = golden_clips(synth, g)
    gin = (clips.astype(np.float32) / np.float32(32768.0)) if f32_input else clips
    n = 8192 + 37
    d = imp.synth_clips_device(n, first_clip=717171, seed=0xF10A7)
    if f32_input:
        d = (d.to(torch.float32) / 32768.0).contiguous()
    try:
        imp.set_split(False)
        p_ref = imp.run_classifier_device(d).clone()
        imp.set_split(True)
        before = imp.launch_count
        got = imp.run_classifier(gin)
        assert imp.launch_count == before + 2, "the split path launches two kernels per chunk"
        if name == "l476f32":
            assert np.abs(got - g["probs"]).max() <= 1e-5  # float softmax: GPU expf vs glibc (north_star's tolerance)
        else:
            assert np.array_equal(got, g["probs"])
        p1 = imp.run_classifier_device(d)
        torch.cuda.synchronize()
        bad = (p1 != p_ref).any(dim=1).nonzero().flatten()
        assert bad.numel() == 0, f"clips whose probabilities differ from the fused kernel's: {bad[:10].tolist()}"
        for m in (1, 2, 3, 4, 5, 163, 739, 741, 1481):
            assert torch.equal(imp.run_classifier_device(d[:m].contiguous()), p_ref[:m]), f"n={m}"
            assert torch.equal(imp.run_classifier_device(d[n - m:].contiguous()), p_ref[n - m:]), f"tail n={m}"
    finally:
        imp.set_split(True)
