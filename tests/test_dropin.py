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


def test_c_hosts_compile_as_c11(tmp_path):
    """a pure-C translation unit (gcc -std=c11, -Wall -Werror) gets run_classifier & co. from the drop-in header; the C++ shim
    ei_run_classifier_c.cpp defines them with C linkage (the reference's README has users delete its C wrapper, README.md:185)"""
    ref = "/root/reference/embedded-demos/stm32cubeide/nucleo-l432-keyword-spotting/keyword-spotting-02-v3"
    if not os.path.isdir(ref):
        pytest.skip("needs the reference export to compile against")
    inc = ["-I" + os.path.join(ROOT, "include"), "-I" + ref]
    for ex in ("static_buffer", "batch_multi_gpu"):
        subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", *inc, "-c", os.path.join(ROOT, "examples", ex + ".c"), "-o", str(tmp_path / (ex + ".o"))], check=True)
    shim = os.path.join(ROOT, "include", "edge-impulse-sdk", "classifier", "ei_run_classifier_c.cpp")
    subprocess.run(["g++", "-std=gnu++14", "-w", *inc, "-c", shim, "-o", str(tmp_path / "shim.o")], check=True)
    syms = subprocess.run(["nm", "--defined-only", str(tmp_path / "shim.o")], capture_output=True, text=True, check=True).stdout
    for s in ("run_classifier", "run_classifier_continuous", "run_classifier_init", "run_classifier_batch_i16", "run_classifier_batch_f32",
              "ei_b200_init", "ei_b200_shutdown"):
        assert re.search(rf" T {s}$", syms, flags=re.M), f"{s} is not defined with C linkage by the shim"
    # the shim's copy of the generated globals is internal: the application's .c file owns the external ones
    assert not re.search(r" [DB] ei_classifier_inferencing_categories$", syms, flags=re.M)


@pytest.mark.parametrize("tag", ["l476", "l432"])
def test_c_host_fails_loudly_without_gpu(tag):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([_binary(tag, "static_buffer_c")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "run_classifier returned -6" in r.stdout and "no usable CUDA device" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["l476", "l432"])
def test_c_host_run_classifier_matches_oracle(tag, tmp_path, synth):
    clip = synth.synth_clips(1, first_clip=79)[0]
    f = tmp_path / "clip.pcm"
    clip.tofile(f)
    r = subprocess.run([_binary(tag, "static_buffer_c"), str(f)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = [float(v) for v in re.findall(r":\s+([0-9.]+)\s*$", r.stdout, flags=re.M)]
    port = PortOracle(tag)
    assert re.findall(r"^\s+(\S+): [0-9.]+\s*$", r.stdout, flags=re.M) == port.labels
    assert np.allclose(got, port.run_classifier_i16(clip)[0], atol=5e-6)


@pytest.mark.gpu
def test_c_batch_host_shards_over_every_gpu():
    """examples/batch_multi_gpu.c: ei_b200_init(all devices) + run_classifier_batch_i16 == the same batch on one device"""
    r = subprocess.run([_binary("l476", "batch_multi_gpu_c"), "3001", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "sharded == single device" in r.stdout, r.stdout + r.stderr
