"""ctypes bindings of the two CPU checkers under oracle/ (TEST INFRASTRUCTURE; never imported by the product).

  RefOracle  : oracle/_ref/liboracle_ref_<model>.so -- the unmodified reference compiled in place
  PortOracle : oracle/liboracle_port.so             -- our plain-C restatement (kws_oracle.c)
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
MODELS_DIR = os.path.join(ROOT, "ei-keyword-spotting_b200", "models")
MODEL_FILES = {"l476": "l476_yes_no.eikwsmdl", "l432": "l432_trick_or_treat.eikwsmdl", "gsc12": "gsc12_synth.eikwsmdl", "l476f32": "l476_f32_twin.eikwsmdl", "zip6": "zip6_arduino.eikwsmdl", "dw3": "dw3_depthwise_synth.eikwsmdl"}

N_SAMPLES = 16000
N_FEATURES = 637


def model_blob(name: str) -> bytes:
    with open(os.path.join(MODELS_DIR, MODEL_FILES[name]), "rb") as f:
        return f.read()


def ref_path(name: str) -> str:
    return os.path.join(ORACLE_DIR, "_ref", f"liboracle_ref_{name}.so")


def have_ref(name: str) -> bool:
    return os.path.isfile(ref_path(name))


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class RefOracle:
    def __init__(self, name: str):
        self.lib = C.CDLL(ref_path(name))
        L = self.lib
        L.ref_label.restype = C.c_char_p
        L.ref_time_run_classifier_i16.restype = C.c_double
        self.n_labels = L.ref_label_count()
        self.n_features = L.ref_feature_count()
        self.labels = [L.ref_label(i).decode() for i in range(self.n_labels)]

    def run_classifier_i16(self, pcm: np.ndarray) -> np.ndarray:
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1, N_SAMPLES)
        out = np.zeros((pcm.shape[0], self.n_labels), np.float32)
        for i in range(pcm.shape[0]):
            rc = self.lib.ref_run_classifier_i16(_p(pcm[i], C.c_int16), N_SAMPLES, _p(out[i], C.c_float))
            assert rc == 0, rc
        return out

    def run_classifier_f32(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, N_SAMPLES)
        out = np.zeros((x.shape[0], self.n_labels), np.float32)
        for i in range(x.shape[0]):
            rc = self.lib.ref_run_classifier_f32(_p(x[i], C.c_float), N_SAMPLES, _p(out[i], C.c_float))
            assert rc == 0, rc
        return out

    def mfcc_i16(self, pcm: np.ndarray) -> np.ndarray:
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1, N_SAMPLES)
        out = np.zeros((pcm.shape[0], self.n_features), np.float32)
        for i in range(pcm.shape[0]):
            rc = self.lib.ref_mfcc_i16(_p(pcm[i], C.c_int16), N_SAMPLES, _p(out[i], C.c_float))
            assert rc == 0, rc
        return out

    @property
    def has_mfe_block(self) -> bool:
        return hasattr(self.lib, "ref_mfe_block_i16")  # only the newer SDK copy (L432) ships extract_mfe_features

    def mfe_block_i16(self, pcm: np.ndarray) -> np.ndarray:
        """extract_mfe_features with the geometry of the model's MFCC block: [n][49 * 32]"""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1, N_SAMPLES)
        out = np.zeros((pcm.shape[0], 49 * 32), np.float32)
        for i in range(pcm.shape[0]):
            rc = self.lib.ref_mfe_block_i16(_p(pcm[i], C.c_int16), N_SAMPLES, _p(out[i], C.c_float), out.shape[1])
            assert rc == out.shape[1], rc
        return out

    def mfcc_f32(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, N_SAMPLES)
        out = np.zeros((x.shape[0], self.n_features), np.float32)
        for i in range(x.shape[0]):
            rc = self.lib.ref_mfcc_f32(_p(x[i], C.c_float), N_SAMPLES, _p(out[i], C.c_float))
            assert rc == 0, rc
        return out

    def run_inference(self, features: np.ndarray, want_tensors=False):
        features = np.ascontiguousarray(features, dtype=np.float32).reshape(-1, self.n_features)
        probs = np.zeros((features.shape[0], self.n_labels), np.float32)
        tensors = []
        cap = 1 << 18
        buf = np.zeros(cap, np.uint8)
        n = C.c_int(0)
        for i in range(features.shape[0]):
            rc = self.lib.ref_run_inference(_p(features[i], C.c_float), _p(probs[i], C.c_float),
                                            _p(buf, C.c_uint8) if want_tensors else None, cap, C.byref(n))
            assert rc == 0, rc
            if want_tensors:
                off, ts = 0, []
                for _ in range(n.value):
                    b = int(np.frombuffer(buf[off:off + 4].tobytes(), np.int32)[0])
                    ts.append(buf[off + 4:off + 4 + b].copy())
                    off += 4 + b
                tensors.append(ts)
        return (probs, tensors) if want_tensors else probs

    def filterbank(self) -> np.ndarray:
        out = np.zeros(129 * 64, np.float32)
        r, c = C.c_int(0), C.c_int(0)
        assert self.lib.ref_filterbank(_p(out, C.c_float), C.byref(r), C.byref(c)) == 0
        return out[: r.value * c.value].reshape(r.value, c.value)

    def mfe_i16(self, pcm: np.ndarray):
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(N_SAMPLES)
        mel = np.zeros(49 * 32, np.float32)
        en = np.zeros(49, np.float32)
        fr = C.c_int(0)
        assert self.lib.ref_mfe_i16(_p(pcm, C.c_int16), N_SAMPLES, _p(mel, C.c_float), _p(en, C.c_float), C.byref(fr)) == 0
        return mel.reshape(49, 32), en

    def mfcc_nocmvn_i16(self, pcm: np.ndarray) -> np.ndarray:
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(N_SAMPLES)
        out = np.zeros(49 * 13, np.float32)
        fr = C.c_int(0)
        assert self.lib.ref_mfcc_nocmvn_i16(_p(pcm, C.c_int16), N_SAMPLES, _p(out, C.c_float), C.byref(fr)) == 0
        return out.reshape(49, 13)

    def cmvnw(self, cepstra: np.ndarray) -> np.ndarray:
        """the reference's own processing::cmvnw on pre-CMVN cepstra [n][49][13] -> float features [n][637]"""
        out = np.ascontiguousarray(cepstra, dtype=np.float32).reshape(-1, N_FEATURES).copy()
        for i in range(out.shape[0]):
            rc = self.lib.ref_cmvnw_f32(_p(out[i], C.c_float), 49)
            assert rc == 0, rc
        return out

    def time_run_classifier_all(self, clips: np.ndarray):
        """(seconds, probs [n][labels]) of the reference's run_classifier over int16 or float32 clips"""
        f32 = clips.dtype == np.float32
        clips = np.ascontiguousarray(clips, dtype=np.float32 if f32 else np.int16).reshape(-1, N_SAMPLES)
        out = np.zeros((clips.shape[0], self.n_labels), np.float32)
        fn = self.lib.ref_time_run_classifier_f32_all if f32 else self.lib.ref_time_run_classifier_i16_all
        fn.restype = C.c_double
        t = float(fn(_p(clips, C.c_float if f32 else C.c_int16), N_SAMPLES, clips.shape[0], _p(out, C.c_float)))
        return t, out

    def time_run_classifier_i16(self, pcm: np.ndarray) -> float:
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1, N_SAMPLES)
        last = np.zeros(self.n_labels, np.float32)
        return float(self.lib.ref_time_run_classifier_i16(_p(pcm, C.c_int16), N_SAMPLES, pcm.shape[0], _p(last, C.c_float)))


class _Cfg(C.Structure):
    _fields_ = [("num_cepstral", C.c_int), ("frame_length", C.c_float), ("frame_stride", C.c_float),
                ("num_filters", C.c_int), ("fft_length", C.c_int), ("win_size", C.c_int), ("low_frequency", C.c_int),
                ("high_frequency", C.c_int), ("pre_cof", C.c_float), ("pre_shift", C.c_int), ("sample_rate", C.c_int)]


class _Taps(C.Structure):
    _fields_ = [("filterbank", C.POINTER(C.c_float)), ("power", C.POINTER(C.c_float)), ("energy", C.POINTER(C.c_float)),
                ("mel", C.POINTER(C.c_float)), ("mfcc", C.POINTER(C.c_float))]


def build_port() -> str:
    path = os.path.join(ORACLE_DIR, "liboracle_port.so")
    if not os.path.isfile(path):
        import subprocess
        subprocess.run(["make", "-C", ORACLE_DIR, "port"], check=True, stdout=subprocess.DEVNULL)
    return path


class PortOracle:
    def __init__(self, name: str):
        self.lib = C.CDLL(build_port())
        L = self.lib
        L.kws_model_load.restype = C.c_void_p
        L.kws_model_load.argtypes = [C.c_char_p, C.c_size_t]
        L.kws_model_mfcc_cfg.restype = C.POINTER(_Cfg)
        L.kws_model_mfcc_cfg.argtypes = [C.c_void_p]
        L.kws_model_label.restype = C.c_char_p
        L.kws_model_label.argtypes = [C.c_void_p, C.c_int]
        for f in ("kws_model_num_labels", "kws_model_num_features", "kws_model_num_tensors"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.kws_model_tensor_bytes.argtypes = [C.c_void_p, C.c_int]
        L.kws_oracle_run_inference.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_void_p)]
        L.kws_oracle_run_classifier_i16.argtypes = [C.c_void_p, C.POINTER(C.c_int16), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.kws_oracle_run_classifier_f32.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.kws_oracle_mfcc_i16.argtypes = [C.POINTER(_Cfg), C.POINTER(C.c_int16), C.c_int, C.POINTER(C.c_float), C.POINTER(_Taps)]
        L.kws_oracle_mfcc_f32.argtypes = [C.POINTER(_Cfg), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.POINTER(_Taps)]
        self._blob = model_blob(name)
        self.m = L.kws_model_load(self._blob, len(self._blob))
        assert self.m, "kws_model_load failed"
        self.n_labels = L.kws_model_num_labels(self.m)
        self.n_features = L.kws_model_num_features(self.m)
        self.n_tensors = L.kws_model_num_tensors(self.m)
        self.cfg = L.kws_model_mfcc_cfg(self.m)
        self.labels = [L.kws_model_label(self.m, i).decode() for i in range(self.n_labels)]

    def mfe_block_i16(self, pcm: np.ndarray) -> np.ndarray:
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1, N_SAMPLES)
        out = np.zeros((pcm.shape[0], 49 * 32), np.float32)
        for i in range(pcm.shape[0]):
            rc = self.lib.kws_oracle_mfe_block_i16(self.cfg, _p(pcm[i], C.c_int16), N_SAMPLES, _p(out[i], C.c_float))
            assert rc == out.shape[1], rc
        return out

    def mfcc_i16(self, pcm: np.ndarray, taps=False):
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1, N_SAMPLES)
        out = np.zeros((pcm.shape[0], self.n_features), np.float32)
        tp = None
        res = []
        for i in range(pcm.shape[0]):
            if taps:
                d = dict(filterbank=np.zeros((129, 32), np.float32), power=np.zeros((49, 129), np.float32),
                         energy=np.zeros(49, np.float32), mel=np.zeros((49, 32), np.float32), mfcc=np.zeros((49, 13), np.float32))
                tp = _Taps(*[_p(d[k], C.c_float) for k in ("filterbank", "power", "energy", "mel", "mfcc")])
                res.append(d)
            rc = self.lib.kws_oracle_mfcc_i16(self.cfg, _p(pcm[i], C.c_int16), N_SAMPLES, _p(out[i], C.c_float), C.byref(tp) if taps else None)
            assert rc == 0, rc
        return (out, res) if taps else out

    def mfcc_f32(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, N_SAMPLES)
        out = np.zeros((x.shape[0], self.n_features), np.float32)
        for i in range(x.shape[0]):
            assert self.lib.kws_oracle_mfcc_f32(self.cfg, _p(x[i], C.c_float), N_SAMPLES, _p(out[i], C.c_float), None) == 0
        return out

    def run_inference(self, features: np.ndarray, want_tensors=False):
        features = np.ascontiguousarray(features, dtype=np.float32).reshape(-1, self.n_features)
        probs = np.zeros((features.shape[0], self.n_labels), np.float32)
        all_t = []
        for i in range(features.shape[0]):
            if want_tensors:
                ts = [np.zeros(max(1, self.lib.kws_model_tensor_bytes(self.m, t)), np.uint8) for t in range(self.n_tensors)]
                arr = (C.c_void_p * self.n_tensors)(*[t.ctypes.data for t in ts])
                rc = self.lib.kws_oracle_run_inference(self.m, _p(features[i], C.c_float), _p(probs[i], C.c_float), arr)
                all_t.append([t[: self.lib.kws_model_tensor_bytes(self.m, k)] for k, t in enumerate(ts)])
            else:
                rc = self.lib.kws_oracle_run_inference(self.m, _p(features[i], C.c_float), _p(probs[i], C.c_float), None)
            assert rc == 0, rc
        return (probs, all_t) if want_tensors else probs

    def cmvn_quantise(self, cepstra: np.ndarray, want_features=False):
        """CMVN + int8 input quantisation of pre-CMVN cepstra [n][49][13] -> int8 [n][637] (and the float features)"""
        cep = np.ascontiguousarray(cepstra, dtype=np.float32).reshape(-1, self.n_features)
        q = np.zeros(cep.shape, np.int8)
        f = np.zeros(cep.shape, np.float32)
        self.lib.kws_oracle_cmvn_quantise.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int8), C.POINTER(C.c_float)]
        for i in range(cep.shape[0]):
            rc = self.lib.kws_oracle_cmvn_quantise(self.m, _p(cep[i], C.c_float), _p(q[i], C.c_int8), _p(f[i], C.c_float))
            assert rc == 0, rc
        return (q, f) if want_features else q

    def run_classifier_i16(self, pcm: np.ndarray, want_features=False):
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1, N_SAMPLES)
        probs = np.zeros((pcm.shape[0], self.n_labels), np.float32)
        feats = np.zeros((pcm.shape[0], self.n_features), np.float32)
        for i in range(pcm.shape[0]):
            rc = self.lib.kws_oracle_run_classifier_i16(self.m, _p(pcm[i], C.c_int16), N_SAMPLES, _p(probs[i], C.c_float), _p(feats[i], C.c_float))
            assert rc == 0, rc
        return (probs, feats) if want_features else probs

    def run_classifier_f32(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, N_SAMPLES)
        probs = np.zeros((x.shape[0], self.n_labels), np.float32)
        for i in range(x.shape[0]):
            assert self.lib.kws_oracle_run_classifier_f32(self.m, _p(x[i], C.c_float), N_SAMPLES, _p(probs[i], C.c_float), None) == 0
        return probs


class PortStream:
    """one reference-semantics audio stream (run_classifier_continuous) on the plain-C oracle"""

    def __init__(self, port: PortOracle, slices_per_window: int = 4):
        self.port = port
        L = port.lib
        L.kws_stream_new.restype = C.c_void_p
        L.kws_stream_new.argtypes = [C.c_void_p, C.c_int]
        L.kws_stream_free.argtypes = [C.c_void_p]
        L.kws_stream_push_i16.argtypes = [C.c_void_p, C.POINTER(C.c_int16), C.c_int16, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        self.s = L.kws_stream_new(port.m, slices_per_window)
        self.slice_size = N_SAMPLES // slices_per_window

    def push(self, slice_i16: np.ndarray, beyond: int = 0):
        sl = np.ascontiguousarray(slice_i16, dtype=np.int16).reshape(self.slice_size)
        probs = np.zeros(self.port.n_labels, np.float32)
        has = C.c_int(0)
        rc = self.port.lib.kws_stream_push_i16(self.s, _p(sl, C.c_int16), C.c_int16(beyond), _p(probs, C.c_float), C.byref(has))
        assert rc == 0, rc
        return probs if has.value else None

    def __del__(self):
        try:
            self.port.lib.kws_stream_free(self.s)
        except Exception:
            pass


class RefStream:
    """one stream on the unmodified reference: its continuous-mode state lives in statics that cannot be reset, so every
    stream gets its own private copy of the shared library"""

    def __init__(self, name: str):
        import shutil
        import tempfile
        self._dir = tempfile.mkdtemp(prefix="eikws_refstream_")
        path = os.path.join(self._dir, "ref.so")
        shutil.copy(ref_path(name), path)
        self.lib = C.CDLL(path)
        self.n_labels = self.lib.ref_label_count()
        self.slice_size = self.lib.ref_slice_size()

    def push(self, slice_i16: np.ndarray, beyond: int = 0):
        sl = np.ascontiguousarray(slice_i16, dtype=np.int16).reshape(self.slice_size)
        probs = np.zeros(self.n_labels, np.float32)
        has = C.c_int(0)
        rc = self.lib.ref_run_classifier_continuous_i16(_p(sl, C.c_int16), C.c_int16(beyond), _p(probs, C.c_float), C.byref(has))
        assert rc == 0, rc
        return probs if has.value else None
