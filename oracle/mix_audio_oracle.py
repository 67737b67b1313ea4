"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the arithmetic of the reference's dataset tooling for one output clip:

    mix_audio                 /root/reference/dataset-curation.py:93-137   (pad / truncate, 0.5 * (word_vol * w + bg_vol * bg[window]))
    sf.write(.., 'PCM_16')    /root/reference/dataset-curation.py:190-206  (float -> 16-bit PCM)

PARITY UNPINNED: the script's own I/O lives in librosa.load (resampling, float32 decoding) and soundfile / libsndfile (the PCM_16
conversion), neither of which is installed in the build image, so this restatement cannot be checked against the reference
itself.  It restates (a) the script's expression with the dtypes its operands have under the NumPy 1.x promotion rules the script
was written for -- `0.5 * word_vol * i` is a Python float times a sample: double; `0.5 * bg_vol * bg[window]` is a Python float
times a float32 array: float32; list + array: float64 -- and (b) libsndfile's published double -> short conversion for normalised
data without clipping: lrint(x * 0x7FFF), truncated to 16 bits.  Resampling is out of scope: inputs are 16 kHz float32.
Only tests/ may import this module.
"""
import numpy as np

N = 16000


def mix_audio(word, bg, start, word_vol=1.0, bg_vol=1.0):
    """word: float32 [len] at 16 kHz or None (dataset-curation.py:104-106); bg: float32 [bg_len]; returns float64 [16000]"""
    if word is None:
        waveform = np.zeros(N, np.float64)                       # [0] * int(sample_time * sample_rate)
    else:
        waveform = np.asarray(word, np.float32).astype(np.float64)  # 0.5 * word_vol * i promotes every sample to double anyway
        if len(waveform) < N:
            waveform = np.append(waveform, np.zeros(N - len(waveform)))   # :114-116
        waveform = waveform[:N]                                            # :119
    word_term = (0.5 * float(word_vol)) * waveform                         # [0.5 * word_vol * i for i in waveform], :131
    bg_term = (np.float32(0.5 * float(bg_vol)) * np.asarray(bg, np.float32)[start:start + N]).astype(np.float32)  # :132, float32 array
    return word_term + bg_term.astype(np.float64)


def to_pcm16(x):
    """libsndfile d2s_array, normalised doubles, clipping off: lrint(x * 32767) (round half to even), low 16 bits"""
    q = np.rint(np.asarray(x, np.float64) * 32767.0).astype(np.int64)
    return (q & 0xFFFF).astype(np.uint16).view(np.int16)
