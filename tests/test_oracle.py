"""CPU tests (no GPU): the plain-C oracle (oracle/kws_oracle.c) against the committed golden vectors generated from
the unmodified reference, and -- when the reference build is present (build container) -- against the reference live."""
import os

import numpy as np
import pytest

from oracle_lib import PortOracle, RefOracle, have_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = ("l476", "l432", "gsc12", "l476f32", "zip6", "dw3")
INTACT = {19: (210, 1470), 21: (0, 210), 23: (0, 70), 25: (10, 70), 27: (0, 10)}


def golden(name):
    return np.load(os.path.join(GOLDEN, f"golden_{name}.npz"))


def golden_clips(synth, g):
    sp = synth.special_clips()
    return np.concatenate([synth.synth_clips(int(g["n_synth"]), 0, int(g["seed"])), np.stack([sp[str(k)] for k in g["special_names"]]),
                           synth.speechlike_clips(g["speech_params"])])


def same_floats(a, b):
    """bit-for-bit equality of float arrays, except that +0.0 and -0.0 are allowed to differ and NaNs must coincide"""
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


@pytest.mark.parametrize("name", MODELS)
def test_port_matches_golden_mfcc(name, synth):
    g = golden(name)
    port = PortOracle(name)
    clips = golden_clips(synth, g)
    feats, taps = port.mfcc_i16(clips, taps=True)
    assert np.array_equal(taps[0]["filterbank"], g["filterbank"])
    assert same_floats(taps[0]["mel"], g["mel0"])
    assert same_floats(taps[0]["energy"], g["energy0"])
    assert same_floats(taps[0]["mfcc"], g["mfcc0"])
    assert same_floats(feats, g["features"])
    # 1e-5 is the north-star tolerance; we are at 0
    assert np.nanmax(np.abs(feats - g["features"])) == 0.0


@pytest.mark.parametrize("name", MODELS)
def test_port_matches_golden_float_input(name, synth):
    g = golden(name)
    port = PortOracle(name)
    clips = golden_clips(synth, g)[:8]
    x = clips.astype(np.float32) / np.float32(32768)
    assert same_floats(port.mfcc_f32(x), g["features_f32in"])
    # int16 -> x/32768 is what the demo callback does, so both entry points must agree
    assert same_floats(g["features_f32in"], g["features"][:8])


@pytest.mark.parametrize("name", MODELS)
def test_port_matches_golden_run_classifier(name, synth):
    g = golden(name)
    port = PortOracle(name)
    assert port.labels == [str(s) for s in g["labels"]]
    probs = port.run_classifier_i16(golden_clips(synth, g))
    assert np.array_equal(probs, g["probs"])


@pytest.mark.parametrize("name", MODELS)
def test_port_matches_golden_int8_classifier(name):
    g = golden(name)
    port = PortOracle(name)
    probs, tens = port.run_inference(g["nn_features"], want_tensors=True)
    assert same_floats(probs, g["nn_probs"])
    nt = port.n_tensors
    intact = {int(k): (int(lo), int(hi)) for k, lo, hi in g["intact"]} if "intact" in g.files else INTACT
    for k, (lo, hi) in intact.items():
        got = np.stack([t[k][lo:hi] for t in tens])
        assert np.array_equal(got, g[f"nn_t{k}"]), f"tensor {k}"
    assert np.array_equal(np.stack([t[nt - 2] for t in tens]), g["nn_t_fc"])
    assert np.array_equal(np.stack([t[nt - 1] for t in tens]), g["nn_t_out"])


def test_mfe_block_port_matches_golden(synth):
    """extract_mfe_features of the newer SDK copy (L432 ei_run_dsp.h:369-418): the golden comes from the unmodified reference;
    degenerate clips (silence, DC) have max == min and turn into NaN rows in both (0 * inf)"""
    g = golden("l432")
    got = PortOracle("l432").mfe_block_i16(golden_clips(synth, g))
    want = g["mfe_features"]
    assert got.shape == want.shape == (len(want), 49 * 32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    finite = np.isfinite(want).all(axis=1)
    assert finite.sum() >= 40 and np.all(want[finite].min(axis=1) == 0.0) and np.all(np.abs(want[finite].max(axis=1) - 1.0) < 1e-6)


def test_silence_hits_epsilon_paths(synth):
    """all-zero clip: every power bin is 0 -> energy and mel are replaced by FLT_EPSILON (feature.hpp:295-297,
    functions.hpp:63-69).  All frames are identical, yet c0's window mean (sum of 101 equal floats / 101) is not exactly
    c0, so CMVN yields d/(|d|+eps) = -0.888... -- the kind of degenerate sensitivity only a bit-exact pipeline reproduces."""
    port = PortOracle("l476")
    feats, taps = port.mfcc_i16(synth.special_clips()["silence"], taps=True)
    eps = np.float32(np.finfo(np.float32).eps)
    assert np.all(taps[0]["energy"] == eps) and np.all(taps[0]["mel"] == eps)
    f = feats.reshape(49, 13)
    assert np.all(f[:, 1:] == 0) and np.all(f[:, 0] == f[0, 0]) and -1.0 < f[0, 0] < -0.5


def test_frame_geometry():
    port = PortOracle("l476")
    c = port.cfg.contents
    assert port.lib.kws_oracle_num_frames(port.cfg, 16000) == 49
    assert port.lib.kws_oracle_num_frames(port.cfg, 16320) == 50  # would overflow the 637-feature block
    assert (c.num_cepstral, c.num_filters, c.fft_length, c.win_size) == (13, 32, 256, 101)


def test_oversized_signal_is_a_dsp_error(synth):
    """ei_run_dsp.h:279-283: more frames than the block owns -> EIDSP_MATRIX_SIZE_MISMATCH -> EI_IMPULSE_DSP_ERROR (-5)"""
    import ctypes as C
    port = PortOracle("l476")
    x = np.zeros(16320, np.int16)
    probs = np.zeros(4, np.float32)
    rc = port.lib.kws_oracle_run_classifier_i16(port.m, x.ctypes.data_as(C.POINTER(C.c_int16)), 16320,
                                               probs.ctypes.data_as(C.POINTER(C.c_float)), None)
    assert rc == -5


@pytest.mark.skipif(not (have_ref("l476") and have_ref("l432") and have_ref("gsc12") and have_ref("l476f32") and have_ref("zip6") and have_ref("dw3")), reason="reference build (oracle/_ref) not present")
@pytest.mark.parametrize("name", MODELS)
def test_port_matches_reference_live(name, synth):
    ref, port = RefOracle(name), PortOracle(name)
    clips = synth.synth_clips(48, first_clip=1000, seed=0xBEEF)
    fr, fp = ref.mfcc_i16(clips), port.mfcc_i16(clips)
    assert same_floats(fr, fp)
    assert np.array_equal(ref.run_classifier_i16(clips), port.run_classifier_i16(clips))
    rng = np.random.default_rng(7)
    F = rng.normal(0, 2.0, (64, 637)).astype(np.float32)
    assert same_floats(ref.run_inference(F), port.run_inference(F))


@pytest.mark.skipif(not (have_ref("l476") and have_ref("l432")), reason="reference build (oracle/_ref) not present")
@pytest.mark.parametrize("name", ["l476", "l432"])
def test_continuous_mode_port_matches_reference_live(name, synth):
    """run_classifier_continuous over 16 slices of one stream: window fill (11+12+12+12 frames), CMVN over the whole
    window incl. its two never-written zero rows, moving average over 2 results, window shift"""
    from oracle_lib import PortStream, RefStream
    ref, port = RefStream(name), PortStream(PortOracle(name))
    audio = synth.synth_clips(4, first_clip=3).reshape(-1)
    n_results = 0
    for i in range(16):
        sl = audio[i * 4000:(i + 1) * 4000]
        a, b = ref.push(sl), port.push(sl)
        assert (a is None) == (b is None)
        assert (a is None) == (i < 3)  # the 637-feature window is full after the 4th slice
        if a is not None:
            assert np.array_equal(a, b)
            n_results += 1
    assert n_results == 13
