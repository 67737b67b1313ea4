"""Caller-side ingest: the on-disk format the reference's tooling produces and consumes.

dataset-curation.py writes 1-second, 16 kHz, mono, PCM_16 WAV files (dataset-curation.py:190-206, 348-351, via
soundfile); `read_wav_clips` turns such files into the [n, 16000] int16 batches the classifier takes, applying the
same pad/truncate-to-one-second rule as mix_audio (dataset-curation.py:93-137: shorter clips are zero-padded at the
end, longer ones truncated).  Only the stdlib `wave` module is used.
"""
import wave

import numpy as np

N_SAMPLES = 16000


def read_wav(path: str) -> np.ndarray:
    """one PCM_16 / 16 kHz / mono WAV file -> int16 samples (raises on any other format: no silent resampling)"""
    with wave.open(path, "rb") as w:
        if w.getnchannels() != 1 or w.getsampwidth() != 2 or w.getframerate() != 16000 or w.getcomptype() != "NONE":
            raise ValueError(f"{path}: need 16 kHz mono PCM_16 (got {w.getnchannels()} ch, {8 * w.getsampwidth()} bit, {w.getframerate()} Hz)")
        return np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").copy()


def to_clip(samples: np.ndarray) -> np.ndarray:
    """pad with zeros / truncate to exactly one second"""
    out = np.zeros(N_SAMPLES, np.int16)
    n = min(len(samples), N_SAMPLES)
    out[:n] = samples[:n]
    return out


def read_wav_clips(paths) -> np.ndarray:
    return np.stack([to_clip(read_wav(p)) for p in paths]) if len(paths) else np.zeros((0, N_SAMPLES), np.int16)


def write_wav(path: str, samples: np.ndarray) -> None:
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(16000)
        w.writeframes(np.ascontiguousarray(samples, dtype="<i2").tobytes())
