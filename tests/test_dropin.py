"""The drop-in header path: examples/static_buffer.cpp is the reference's known-answer call sequence (signal_t +
run_classifier) compiled UNCHANGED against include/edge-impulse-sdk and the UNMODIFIED generated model files of both
shipped exports (built by __graft_entry__.build() into examples/_build/)."""
import os
import re
import subprocess

import numpy as np
import pytest

from oracle_lib import PortOracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "examples", "_build")


def _binary(tag, example="static_buffer"):
    p = os.path.join(BUILD, f"{example}_{tag}")
    if not os.path.isfile(p):
        pytest.skip("examples/_build not present (run __graft_entry__.build() where /root/reference exists)")
    return p


@pytest.mark.parametrize("tag", ["l476", "l432"])
def test_dropin_fails_loudly_without_gpu(tag):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([_binary(tag)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1
    assert "run_classifier returned -6" in r.stdout and "no usable CUDA device" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["l476", "l432"])
def test_dropin_run_classifier_matches_oracle(tag, tmp_path, synth):
    clip = synth.synth_clips(1, first_clip=77)[0]
    f = tmp_path / "clip.pcm"
    clip.tofile(f)
    r = subprocess.run([_binary(tag), str(f)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = [float(v) for v in re.findall(r":\s+([0-9.]+)\s*$", r.stdout, flags=re.M)]
    port = PortOracle(tag)
    want = port.run_classifier_i16(clip)[0]
    labels = re.findall(r"^\s+(\S+): [0-9.]+\s*$", r.stdout, flags=re.M)
    assert labels == port.labels
    assert np.allclose(got, want, atol=5e-6)  # printed with 5 decimals; values are multiples of 1/256


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["l476", "l432"])
def test_dropin_extract_mfe_features_matches_oracle(tag, tmp_path, synth):
    """the sibling MFE DSP block through the drop-in ei_run_dsp.h entry point (extract_fn(signal, matrix, config))"""
    clip = synth.synth_clips(1, first_clip=78)[0]
    f = tmp_path / "clip.pcm"
    clip.tofile(f)
    r = subprocess.run([_binary(tag), str(f), "mfe"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "MFE features: 1 x 1568" in r.stdout
    got = np.array([float(v) for v in re.findall(r"^mfe \d+ (\S+)$", r.stdout, flags=re.M)], np.float32)
    want = PortOracle(tag).mfe_block_i16(clip)[0]
    assert np.array_equal(got, want)  # %.9g round-trips a float32


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["l476", "l432"])
def test_dropin_run_classifier_continuous_matches_oracle(tag, tmp_path, synth):
    """the firmware main loop (signal_t + run_classifier_continuous per 250 ms slice) against the drop-in header"""
    from oracle_lib import PortStream
    audio = synth.synth_clips(3, first_clip=41).reshape(-1)  # 12 slices
    f = tmp_path / "audio.pcm"
    audio.tofile(f)
    r = subprocess.run([_binary(tag, "continuous_stream"), str(f)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = {int(m.group(1)): [float(v) for v in m.group(2).split()] for m in re.finditer(r"^slice (\d+):(.*)$", r.stdout, flags=re.M)}
    stream = PortStream(PortOracle(tag))
    for i in range(12):
        want = stream.push(audio[i * 4000:(i + 1) * 4000])
        assert (want is None) == (i not in got)
        if want is not None:
            assert np.allclose(got[i], want, atol=1e-8)
