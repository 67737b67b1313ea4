"""Deterministic synthetic 1-second 16 kHz int16 clips (host/numpy twin of eikws_synth_kernel in csrc/kernels.cu).

Integer arithmetic only, so the host and the GPU generate bit-identical clips from (seed, clip index):
a counter-based splitmix64 stream, Irwin-Hall(4) noise, and a per-clip kind drawn from the clip hash:
  70 % noise sigma~3000 | 10 % sigma~300 | 10 % sigma~12000 (saturates) | 5 % silence | 5 % square wave + noise
(the mixture SURVEY.md §8(d) asks for).  `special_clips()` adds the edge cases used by the parity tests.
"""
import numpy as np

N_SAMPLES = 16000
DEFAULT_SEED = 0xE1D5

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return x ^ (x >> np.uint64(31))


def synth_clips(n_clips: int, first_clip: int = 0, seed: int = DEFAULT_SEED) -> np.ndarray:
    """[n_clips, 16000] int16, identical to eikws_synth_i16_device(first_clip, seed)."""
    out = np.empty((n_clips, N_SAMPLES), np.int16)
    i = np.arange(N_SAMPLES, dtype=np.uint64)
    for c in range(n_clips):
        clip = np.uint64(first_clip + c)
        with np.errstate(over="ignore"):
            hc = _splitmix64(np.array([np.uint64(seed) ^ ((clip * np.uint64(0xD1B54A32D192ED03)) & _M64)], np.uint64))[0]
            r = _splitmix64((hc + i) & _M64)
        kind = int(hc % np.uint64(20))
        s = ((r & np.uint64(0xFFFF)) + ((r >> np.uint64(16)) & np.uint64(0xFFFF)) + ((r >> np.uint64(32)) & np.uint64(0xFFFF)) +
             (r >> np.uint64(48))).astype(np.int64) - 131070
        if kind < 14:
            v = (s * 81) >> 10
        elif kind < 16:
            v = (s * 8) >> 10
        elif kind < 18:
            v = (s * 325) >> 10
        elif kind == 18:
            v = np.zeros(N_SAMPLES, np.int64)
        else:
            period = 16 + int((hc >> np.uint64(8)) % np.uint64(240))
            sq = np.where(((i.astype(np.int64) // (period // 2)) & 1) == 1, -8000, 8000)
            v = sq + ((s * 8) >> 10)
        out[c] = np.clip(v, -32768, 32767).astype(np.int16)
    return out


def special_clips() -> dict:
    """Hand-built edge cases (name -> [16000] int16) that drive the reference's corner paths:
    FLT_EPSILON substitutions (silence, feature.hpp:295-297 / functions.hpp:63-69), int16 saturation,
    a single impulse (CMVN outliers and the float->int8 cast overflow of ei_run_classifier.h:440),
    DC, Nyquist-rate alternation, a chirp, a ramp, and a one-sample-from-silence clip."""
    n = np.arange(N_SAMPLES)
    d = {}
    d["silence"] = np.zeros(N_SAMPLES, np.int16)
    d["dc_pos"] = np.full(N_SAMPLES, 12345, np.int16)
    d["dc_min"] = np.full(N_SAMPLES, -32768, np.int16)
    d["full_scale_square"] = np.where((n // 20) % 2 == 0, 32767, -32768).astype(np.int16)
    d["nyquist"] = np.where(n % 2 == 0, 20000, -20000).astype(np.int16)
    imp = np.zeros(N_SAMPLES, np.int16)
    imp[8000] = 32767
    d["impulse_mid"] = imp
    imp0 = np.zeros(N_SAMPLES, np.int16)
    imp0[0] = -32768
    d["impulse_first"] = imp0
    impl = np.zeros(N_SAMPLES, np.int16)
    impl[15999] = 30000  # only reachable through the pre-emphasis wrap-around (processing.hpp:68)
    d["impulse_last"] = impl
    one = np.zeros(N_SAMPLES, np.int16)
    one[321] = 1
    d["one_lsb"] = one
    ph = 2 * np.pi * (200.0 * n / 16000.0 + 0.5 * 3800.0 * (n / 16000.0) ** 2)
    d["chirp"] = np.round(12000 * np.sin(ph)).astype(np.int16)
    d["ramp"] = ((n * 4) % 65536 - 32768).astype(np.int16)
    d["tone_1k"] = np.round(8000 * np.sin(2 * np.pi * 1000.0 * n / 16000.0)).astype(np.int16)
    burst = np.zeros(N_SAMPLES, np.int64)
    burst[4000:4400] = synth_clips(1, first_clip=7)[0][:400].astype(np.int64) * 3
    d["burst"] = np.clip(burst, -32768, 32767).astype(np.int16)
    return d
