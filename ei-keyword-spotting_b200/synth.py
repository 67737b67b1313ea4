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


# ---- speech-like clips (integer arithmetic only: reproducible bit for bit on any machine) ------------------------------
def _isin(phase: np.ndarray) -> np.ndarray:
    """integer sine: phase in 1/65536 turns -> [-32767, 32767] (parabola + one refinement step, all int64)"""
    x = phase.astype(np.int64) & 0xFFFF
    x = np.where(x >= 32768, x - 65536, x)
    y = (x * (32768 - np.abs(x))) >> 13
    y = y + ((((y * np.abs(y)) >> 15) - y) * 7373 >> 15)
    return np.clip(y, -32767, 32767)


def speechlike_clip(params) -> np.ndarray:
    """One voiced-word-like clip from 10 small integers: a harmonic stack on a gliding pitch, weighted by three formant peaks that
    move from one vowel to another, under a syllable envelope, with a noise burst (a 'consonant') in front and light background
    noise -- closer to what the classifiers were trained on than Gaussian noise, so every label gets to win (tests/golden).
    params = (f0_hz, glide_hz, vowel_a, vowel_b, onset_ms, length_ms, level, burst_ms, burst_level, seed)"""
    f0, glide, va, vb, onset_ms, length_ms, level, burst_ms, burst_level, seed = (int(v) for v in params)
    vowels = ((730, 1090, 2440), (270, 2290, 3010), (300, 870, 2240), (530, 1840, 2480), (660, 1720, 2410), (440, 1020, 2240),
              (490, 1350, 1690), (640, 1190, 2390))  # Peterson-Barney style (F1, F2, F3) in Hz
    n = np.arange(N_SAMPLES, dtype=np.int64)
    on, ln = onset_ms * 16, max(1, length_ms * 16)
    t = np.clip(n - on, 0, ln)  # position inside the syllable
    # pitch glides linearly over the syllable; phase = cumulative sum of the per-sample increment (1/65536 turns)
    f_inst = f0 * ln + glide * t  # Hz * ln
    inc = (f_inst * 65536) // (16000 * ln)
    phase = np.cumsum(inc)
    fa, fb = vowels[va % 8], vowels[vb % 8]
    voiced = np.zeros(N_SAMPLES, np.int64)
    for h in range(1, 24):
        fh = h * (f0 + glide // 2)
        if fh > 3800:
            break
        w = 0
        for k in range(3):  # formant k moves from vowel a to vowel b; triangular resonance of half-width 220 Hz, later ones weaker
            # weight evaluated at the syllable midpoint (constant per harmonic keeps everything integer and cheap)
            fc = (fa[k] + fb[k]) // 2
            w += max(0, 220 - abs(fh - fc)) * (4 - k)
        w = max(w, 12)
        voiced += (w * _isin(phase * h)) >> 7
    env = (_isin((t * 32768) // ln) * level) >> 15  # half-sine syllable envelope, 0 outside
    env = np.where((n >= on) & (n < on + ln), env, 0)
    x = (voiced * env) >> 18
    with np.errstate(over="ignore"):
        r = _splitmix64((np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + n.astype(np.uint64)) & _M64)
    noise = ((r & np.uint64(0xFFFF)).astype(np.int64) + ((r >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.int64)) - 65535
    b0, b1 = max(0, on - burst_ms * 16), on
    x = x + np.where((n >= b0) & (n < b1), (noise * burst_level) >> 12, 0) + ((noise * 3) >> 10)
    return np.clip(x, -32768, 32767).astype(np.int16)


def speechlike_clips(param_rows) -> np.ndarray:
    rows = np.asarray(param_rows).reshape(-1, 10)
    return np.stack([speechlike_clip(r) for r in rows]) if len(rows) else np.zeros((0, N_SAMPLES), np.int16)
