"""eikws-b200: Python mirror of the reference's classifier interface over libeikws_b200.so (ctypes).

The product is the C-ABI shared library (include/eikws_b200.h); this module is plumbing for tests, bench.py and
Python callers: it only marshals pointers.  It deliberately has NO fallback: if the CUDA library is missing or no
B200 is present, construction raises.

Names follow the reference (edge-impulse-sdk/classifier/ei_run_classifier.h): `run_classifier` (:650),
`run_inference` (:293), `extract_mfcc_features` (ei_run_dsp.h:256); batch arguments replace signal_t.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EIKWS_B200_LIB") or os.path.join(_HERE, "libeikws_b200.so")  # the override is for A/B timing of kernel builds
MODELS_DIR = os.path.join(_HERE, "models")
MODELS = {"l476": "l476_yes_no.eikwsmdl", "l432": "l432_trick_or_treat.eikwsmdl", "gsc12": "gsc12_synth.eikwsmdl", "l476f32": "l476_f32_twin.eikwsmdl", "zip6": "zip6_arduino.eikwsmdl", "dw3": "dw3_depthwise_synth.eikwsmdl"}

EI_IMPULSE_OK = 0
EI_IMPULSE_DSP_ERROR = -5
N_SAMPLES = 16000


class EikwsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"eikws error {code}: {msg}")
        self.code = code


_lib = None


def load_library() -> C.CDLL:
    """dlopen the in-tree CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, sz, i32, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
    L.eikws_last_error.restype = C.c_char_p
    L.eikws_create.argtypes = [C.c_char_p, sz, i32, C.POINTER(vp)]
    L.eikws_destroy.argtypes = [vp]
    L.eikws_label.restype = C.c_char_p
    L.eikws_label.argtypes = [vp, i32]
    for f in ("eikws_label_count", "eikws_feature_count", "eikws_raw_sample_count", "eikws_device"):
        getattr(L, f).argtypes = [vp]
    L.eikws_launch_count.restype = u64
    L.eikws_launch_count.argtypes = [vp]
    L.eikws_set_ctas_per_sm.argtypes = [vp, i32]
    L.eikws_set_clips_per_cta.argtypes = [vp, i32]
    L.eikws_set_skew_ns.argtypes = [vp, i32]
    L.eikws_set_tensor_core.argtypes = [vp, i32]
    L.eikws_set_cmvn_shortcut.argtypes = [vp, i32]
    L.eikws_set_work_claiming.argtypes = [vp, i32]
    L.eikws_set_pipelined.argtypes = [vp, i32]
    L.eikws_set_split.argtypes = [vp, i32]
    L.eikws_set_kernel_timing.argtypes = [vp, i32]
    L.eikws_split_kernel_ms.argtypes = [vp, vp, vp]
    L.eikws_classify_i16_device.argtypes = [vp, vp, sz, vp, vp]
    L.eikws_classify_f32_device.argtypes = [vp, vp, sz, vp, vp]
    L.eikws_features_i16_device.argtypes = [vp, vp, sz, vp, vp, vp]
    L.eikws_features_f32_device.argtypes = [vp, vp, sz, vp, vp, vp]
    L.eikws_infer_device.argtypes = [vp, vp, sz, vp, vp]
    L.eikws_classify_taps_i16_device.argtypes = [vp, vp, sz, vp, vp, vp, vp]
    L.eikws_synth_i16_device.argtypes = [vp, vp, sz, u64, u64, vp]
    L.eikws_decimate_i2s_device.argtypes = [vp, vp, sz, i32, i32, vp, vp]
    L.eikws_classify_i16_host.argtypes = [vp, vp, sz, vp]
    L.eikws_classify_f32_host.argtypes = [vp, vp, sz, vp]
    L.eikws_features_i16_host.argtypes = [vp, vp, sz, vp, vp]
    for _fn in (L.eikws_mfe_i16_device, L.eikws_mfe_f32_device):
        _fn.argtypes = [vp, vp, sz, vp, vp]
    for _fn in (L.eikws_mfe_i16_host, L.eikws_mfe_f32_host):
        _fn.argtypes = [vp, vp, sz, vp]
    L.eikws_mfe_feature_count.argtypes = [vp]
    L.eikws_features_f32_host.argtypes = [vp, vp, sz, vp, vp]
    L.eikws_infer_host.argtypes = [vp, vp, sz, vp]
    L.eikws_classify_taps_i16_host.argtypes = [vp, vp, sz, vp, vp, vp]
    L.eikws_streams_create.argtypes = [vp, sz, i32, C.POINTER(vp)]
    L.eikws_streams_destroy.argtypes = [vp]
    L.eikws_streams_reset.argtypes = [vp]
    L.eikws_streams_slice_size.argtypes = [vp]
    L.eikws_streams_push_i16_host.argtypes = [vp, vp, C.c_float, vp, C.POINTER(i32)]
    L.eikws_streams_push_i16_device.argtypes = [vp, vp, C.c_float, vp, C.POINTER(i32), vp]
    L.eikws_debug_host_plan.argtypes = [C.c_char_p, sz, vp, vp, vp, i32, C.POINTER(i32)]
    L.eikws_debug_cmvn_quantise_host.argtypes = [vp, vp, sz, i32, vp]
    L.eikws_mix_audio_device.argtypes = [vp, vp, vp, sz, vp, sz, vp, sz, C.c_double, C.c_double, sz, vp, vp]
    L.eikws_multi_create.argtypes = [C.c_char_p, sz, C.POINTER(i32), i32, C.POINTER(vp)]
    L.eikws_multi_destroy.argtypes = [vp]
    L.eikws_multi_device_count.argtypes = [vp]
    L.eikws_multi_handle.restype = vp
    L.eikws_multi_handle.argtypes = [vp, i32]
    L.eikws_multi_shard.argtypes = [vp, sz, i32, C.POINTER(sz), C.POINTER(sz)]
    L.eikws_multi_classify_i16_host.argtypes = [vp, vp, sz, vp]
    L.eikws_multi_classify_f32_host.argtypes = [vp, vp, sz, vp]
    L.eikws_host_alloc.restype = vp
    L.eikws_host_alloc.argtypes = [sz]
    L.eikws_host_free.argtypes = [vp]
    _lib = L
    return L


def model_blob(name_or_path: str) -> bytes:
    path = os.path.join(MODELS_DIR, MODELS[name_or_path]) if name_or_path in MODELS else name_or_path
    with open(path, "rb") as f:
        return f.read()


def _check(rc):
    if rc != 0:
        raise EikwsError(rc, load_library().eikws_last_error().decode(errors="replace"))


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data)


class Impulse:
    """One Edge Impulse impulse (MFCC block + int8 classifier) resident on one B200."""

    def __init__(self, model="l476", device=0):
        self._lib = load_library()
        self._blob = model_blob(model) if isinstance(model, str) else bytes(model)
        h = C.c_void_p()
        _check(self._lib.eikws_create(self._blob, len(self._blob), int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        self.label_count = self._lib.eikws_label_count(h)
        self.feature_count = self._lib.eikws_feature_count(h)
        self.raw_sample_count = self._lib.eikws_raw_sample_count(h)
        self.labels = [self._lib.eikws_label(h, i).decode() for i in range(self.label_count)]

    def close(self):
        if getattr(self, "_h", None):
            self._lib.eikws_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self._lib.eikws_launch_count(self._h))

    def set_clips_per_cta(self, n: int):
        _check(self._lib.eikws_set_clips_per_cta(self._h, n))

    def set_ctas_per_sm(self, n: int):
        _check(self._lib.eikws_set_ctas_per_sm(self._h, n))

    def set_tensor_core(self, on: bool):
        """block 1 of the fused int8 classifier as a tcgen05 UMMA (int16 clips, two clip groups per CTA)"""
        _check(self._lib.eikws_set_tensor_core(self._h, 1 if on else 0))

    def set_cmvn_shortcut(self, on: bool):
        """certified CMVN shortcut (default on): window statistics in one double-precision pass, rounding decision certified by
        a rigorous error bound, uncertified chains recomputed with the reference's operation sequence -- same int8 features"""
        _check(self._lib.eikws_set_cmvn_shortcut(self._h, 1 if on else 0))

    def set_work_claiming(self, on: bool):
        """work-claiming schedule of the shortcut kernel (frame pairs and the UMMA issue claimed from shared counters)"""
        _check(self._lib.eikws_set_work_claiming(self._h, 1 if on else 0))

    def set_pipelined(self, on: bool):
        """the software-pipelined classify kernel: FFT of clip s interleaved, warp by warp, with the post-FFT slices of clip s-1"""
        _check(self._lib.eikws_set_pipelined(self._h, 1 if on else 0))

    def set_split(self, on: bool):
        """the two-kernel classify path (int16 clips): barrier-free spectral kernel, then the cepstral / classifier kernel"""
        _check(self._lib.eikws_set_split(self._h, 1 if on else 0))

    def set_kernel_timing(self, on: bool):
        """record CUDA events around the two kernels of every split launch (see split_kernel_ms)"""
        _check(self._lib.eikws_set_kernel_timing(self._h, 1 if on else 0))

    def split_kernel_ms(self):
        """(spectral kernel ms, cepstral / classifier kernel ms, launches): averages per launch over the split launches since timing was
        switched on or since the last call; waits for them"""
        ms = (C.c_float * 2)()
        cnt = C.c_uint64(0)
        _check(self._lib.eikws_split_kernel_ms(self._h, ms, C.byref(cnt)))
        return float(ms[0]), float(ms[1]), int(cnt.value)

    def set_skew_ns(self, ns: int):
        _check(self._lib.eikws_set_skew_ns(self._h, ns))

    # ---- host (numpy) batch API: H2D + kernel + D2H inside the call -----------------------------------
    def run_classifier(self, clips: np.ndarray) -> np.ndarray:
        """clips: [n, 16000] int16 PCM or float32 samples -> [n, label_count] float32 (classification[i].value)."""
        clips = np.ascontiguousarray(clips).reshape(-1, self.raw_sample_count)
        out = np.empty((clips.shape[0], self.label_count), np.float32)
        if clips.dtype == np.int16:
            _check(self._lib.eikws_classify_i16_host(self._h, _np_ptr(clips), clips.shape[0], _np_ptr(out)))
        elif clips.dtype == np.float32:
            _check(self._lib.eikws_classify_f32_host(self._h, _np_ptr(clips), clips.shape[0], _np_ptr(out)))
        else:
            raise TypeError("clips must be int16 or float32")
        return out

    def run_classifier_taps(self, clips: np.ndarray):
        """int16 clips -> (probs, float features [n,637], int8 quantised NN input [n,637])."""
        clips = np.ascontiguousarray(clips, dtype=np.int16).reshape(-1, self.raw_sample_count)
        n = clips.shape[0]
        probs = np.empty((n, self.label_count), np.float32)
        feat = np.empty((n, self.feature_count), np.float32)
        q = np.empty((n, self.feature_count), np.int8)
        _check(self._lib.eikws_classify_taps_i16_host(self._h, _np_ptr(clips), n, _np_ptr(probs), _np_ptr(feat), _np_ptr(q)))
        return probs, feat, q

    def extract_mfcc_features(self, clips: np.ndarray, quantized=False):
        clips = np.ascontiguousarray(clips).reshape(-1, self.raw_sample_count)
        n = clips.shape[0]
        feat = np.empty((n, self.feature_count), np.float32)
        q = np.empty((n, self.feature_count), np.int8) if quantized else None
        fn = self._lib.eikws_features_i16_host if clips.dtype == np.int16 else self._lib.eikws_features_f32_host
        if clips.dtype not in (np.int16, np.float32):
            raise TypeError("clips must be int16 or float32")
        _check(fn(self._h, _np_ptr(clips), n, _np_ptr(feat), _np_ptr(q) if quantized else None))
        return (feat, q) if quantized else feat

    def extract_mfe_features(self, clips: np.ndarray) -> np.ndarray:
        """the sibling MFE DSP block (extract_mfe_features, L432 SDK copy) with the MFCC block's geometry: [n][49 * 32]"""
        clips = np.ascontiguousarray(clips).reshape(-1, self.raw_sample_count)
        if clips.dtype not in (np.int16, np.float32):
            raise TypeError("clips must be int16 or float32")
        feat = np.empty((clips.shape[0], self._lib.eikws_mfe_feature_count(self._h)), np.float32)
        fn = self._lib.eikws_mfe_i16_host if clips.dtype == np.int16 else self._lib.eikws_mfe_f32_host
        _check(fn(self._h, _np_ptr(clips), clips.shape[0], _np_ptr(feat)))
        return feat

    def run_inference(self, features: np.ndarray) -> np.ndarray:
        features = np.ascontiguousarray(features, dtype=np.float32).reshape(-1, self.feature_count)
        out = np.empty((features.shape[0], self.label_count), np.float32)
        _check(self._lib.eikws_infer_host(self._h, _np_ptr(features), features.shape[0], _np_ptr(out)))
        return out

    def debug_cmvn_quantise(self, cepstra: np.ndarray, shortcut: bool) -> np.ndarray:
        """tests only: CMVN + int8 input quantisation of pre-CMVN cepstra [n,49,13] on the device -> int8 [n,637];
        shortcut=True runs the certified path of the default classify kernel, False every chain with the reference's sequence"""
        cep = np.ascontiguousarray(cepstra, dtype=np.float32).reshape(-1, self.feature_count)
        q = np.empty(cep.shape, np.int8)
        _check(self._lib.eikws_debug_cmvn_quantise_host(self._h, _np_ptr(cep), cep.shape[0], 1 if shortcut else 0, _np_ptr(q)))
        return q

    # ---- device (torch) batch API: tensors already resident in HBM, asynchronous on the current stream -----
    def run_classifier_device(self, clips, out=None):
        import torch
        assert clips.is_cuda and clips.is_contiguous() and clips.device.index == self.device
        n = clips.numel() // self.raw_sample_count
        if out is None:
            out = torch.empty((n, self.label_count), dtype=torch.float32, device=clips.device)
        stream = C.c_void_p(torch.cuda.current_stream(clips.device).cuda_stream)
        if clips.dtype == torch.int16:
            _check(self._lib.eikws_classify_i16_device(self._h, C.c_void_p(clips.data_ptr()), n, C.c_void_p(out.data_ptr()), stream))
        elif clips.dtype == torch.float32:
            _check(self._lib.eikws_classify_f32_device(self._h, C.c_void_p(clips.data_ptr()), n, C.c_void_p(out.data_ptr()), stream))
        else:
            raise TypeError("clips must be int16 or float32")
        return out

    def run_classifier_taps_device(self, clips, want_features=False):
        """int16 clips on the device -> (probs, int8 quantised NN input [n,637][, float features]) from the classify kernel itself;
        without the float features the kernel is the one run_classifier_device launches"""
        import torch
        assert clips.is_cuda and clips.is_contiguous() and clips.dtype == torch.int16 and clips.device.index == self.device
        n = clips.numel() // self.raw_sample_count
        probs = torch.empty((n, self.label_count), dtype=torch.float32, device=clips.device)
        q = torch.empty((n, self.feature_count), dtype=torch.int8, device=clips.device)
        feat = torch.empty((n, self.feature_count), dtype=torch.float32, device=clips.device) if want_features else None
        stream = C.c_void_p(torch.cuda.current_stream(clips.device).cuda_stream)
        _check(self._lib.eikws_classify_taps_i16_device(self._h, C.c_void_p(clips.data_ptr()), n, C.c_void_p(probs.data_ptr()),
                                                        C.c_void_p(feat.data_ptr()) if want_features else None, C.c_void_p(q.data_ptr()), stream))
        return (probs, q, feat) if want_features else (probs, q)

    def extract_mfcc_features_device(self, clips, features=None, qfeatures=None):
        import torch
        assert clips.is_cuda and clips.is_contiguous() and clips.device.index == self.device
        n = clips.numel() // self.raw_sample_count
        if features is None and qfeatures is None:
            features = torch.empty((n, self.feature_count), dtype=torch.float32, device=clips.device)
        stream = C.c_void_p(torch.cuda.current_stream(clips.device).cuda_stream)
        fn = self._lib.eikws_features_i16_device if clips.dtype == torch.int16 else self._lib.eikws_features_f32_device
        _check(fn(self._h, C.c_void_p(clips.data_ptr()), n, C.c_void_p(features.data_ptr()) if features is not None else None,
                  C.c_void_p(qfeatures.data_ptr()) if qfeatures is not None else None, stream))
        return features if qfeatures is None else (features, qfeatures)

    def extract_mfe_features_device(self, clips, features=None):
        import torch
        assert clips.is_cuda and clips.is_contiguous() and clips.device.index == self.device
        n = clips.numel() // self.raw_sample_count
        if features is None:
            features = torch.empty((n, self._lib.eikws_mfe_feature_count(self._h)), dtype=torch.float32, device=clips.device)
        stream = C.c_void_p(torch.cuda.current_stream(clips.device).cuda_stream)
        fn = self._lib.eikws_mfe_i16_device if clips.dtype == torch.int16 else self._lib.eikws_mfe_f32_device
        _check(fn(self._h, C.c_void_p(clips.data_ptr()), n, C.c_void_p(features.data_ptr()), stream))
        return features

    def run_inference_device(self, features, out=None):
        import torch
        assert features.is_cuda and features.is_contiguous() and features.dtype == torch.float32
        n = features.numel() // self.feature_count
        if out is None:
            out = torch.empty((n, self.label_count), dtype=torch.float32, device=features.device)
        stream = C.c_void_p(torch.cuda.current_stream(features.device).cuda_stream)
        _check(self._lib.eikws_infer_device(self._h, C.c_void_p(features.data_ptr()), n, C.c_void_p(out.data_ptr()), stream))
        return out

    def decimate_i2s_device(self, i2s, n_out, skip=4, shift=8):
        """firmware microphone path (main.cpp:507-521): int32 SAI words -> int16 16 kHz mono, pcm[i] = i2s[skip*i] >> shift"""
        import torch
        assert i2s.is_cuda and i2s.dtype == torch.int32 and i2s.is_contiguous() and i2s.numel() >= skip * n_out
        out = torch.empty(n_out, dtype=torch.int16, device=i2s.device)
        stream = C.c_void_p(torch.cuda.current_stream(i2s.device).cuda_stream)
        _check(self._lib.eikws_decimate_i2s_device(self._h, C.c_void_p(i2s.data_ptr()), n_out, skip, shift, C.c_void_p(out.data_ptr()), stream))
        return out

    def mix_audio_device(self, words, word_len, bg, bg_start, word_vol=1.0, bg_vol=0.1):
        """mix_audio of the dataset tooling (dataset-curation.py:93-137) + PCM_16 conversion for 16 kHz float32 tensors on the device:
        words [n, stride] float32 (or None: background-only clips) with word_len [n] uint32 valid samples, bg [bg_len] float32,
        bg_start [n] uint32 -> [n, 16000] int16"""
        import torch
        n = int(bg_start.numel())
        assert bg.is_cuda and bg.dtype == torch.float32 and bg_start.dtype in (torch.int32, torch.uint32) and bg.is_contiguous()
        assert words is None or (words.dtype == torch.float32 and words.stride(1) == 1 and word_len.dtype in (torch.int32, torch.uint32))
        out = torch.empty((n, self.raw_sample_count), dtype=torch.int16, device=bg.device)
        stream = C.c_void_p(torch.cuda.current_stream(bg.device).cuda_stream)
        max_start = int(bg_start.max().item()) if n else 0
        _check(self._lib.eikws_mix_audio_device(self._h, C.c_void_p(words.data_ptr()) if words is not None else None,
                                                C.c_void_p(word_len.data_ptr()) if words is not None else None,
                                                int(words.stride(0)) if words is not None else 0, C.c_void_p(bg.data_ptr()), int(bg.numel()),
                                                C.c_void_p(bg_start.data_ptr()), max_start, C.c_double(word_vol), C.c_double(bg_vol), n,
                                                C.c_void_p(out.data_ptr()), stream))
        return out

    def synth_clips_device(self, n_clips, first_clip=0, seed=0xE1D5):
        import torch
        out = torch.empty((n_clips, self.raw_sample_count), dtype=torch.int16, device=f"cuda:{self.device}")
        stream = C.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)
        _check(self._lib.eikws_synth_i16_device(self._h, C.c_void_p(out.data_ptr()), n_clips, first_clip, seed, stream))
        return out


class MultiImpulse:
    """One impulse replicated on several GPUs of the box (eikws_multi_*): host batches are sharded contiguously, one host
    thread and stream pair per device, no exchange between devices."""

    def __init__(self, model="l476", devices=None, n_devices=0):
        self._lib = load_library()
        self._blob = model_blob(model) if isinstance(model, str) else bytes(model)
        h = C.c_void_p()
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            _check(self._lib.eikws_multi_create(self._blob, len(self._blob), arr, len(devices), C.byref(h)))
        else:
            _check(self._lib.eikws_multi_create(self._blob, len(self._blob), None, int(n_devices), C.byref(h)))
        self._m = h
        self.device_count = self._lib.eikws_multi_device_count(h)
        self.label_count = self._lib.eikws_label_count(self._lib.eikws_multi_handle(h, 0))

    def close(self):
        if getattr(self, "_m", None):
            self._lib.eikws_multi_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shard(self, n_clips: int, i: int):
        first, count = C.c_size_t(0), C.c_size_t(0)
        self._lib.eikws_multi_shard(self._m, n_clips, i, C.byref(first), C.byref(count))
        return first.value, count.value

    def run_classifier(self, clips: np.ndarray) -> np.ndarray:
        clips = np.ascontiguousarray(clips).reshape(-1, N_SAMPLES)
        out = np.empty((clips.shape[0], self.label_count), np.float32)
        if clips.dtype == np.int16:
            _check(self._lib.eikws_multi_classify_i16_host(self._m, _np_ptr(clips), clips.shape[0], _np_ptr(out)))
        elif clips.dtype == np.float32:
            _check(self._lib.eikws_multi_classify_f32_host(self._m, _np_ptr(clips), clips.shape[0], _np_ptr(out)))
        else:
            raise TypeError("clips must be int16 or float32")
        return out


class Streams:
    """run_classifier_continuous over `n_streams` concurrent audio streams (one slice per stream per push)."""

    def __init__(self, impulse: "Impulse", n_streams: int, slices_per_window: int = 4):
        self._lib = load_library()
        self.impulse = impulse
        self.n_streams = n_streams
        h = C.c_void_p()
        _check(self._lib.eikws_streams_create(impulse._h, n_streams, slices_per_window, C.byref(h)))
        self._s = h
        self.slice_size = self._lib.eikws_streams_slice_size(h)

    def close(self):
        if getattr(self, "_s", None):
            self._lib.eikws_streams_destroy(self._s)
            self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        _check(self._lib.eikws_streams_reset(self._s))

    def push(self, slices: np.ndarray, beyond: float = 0.0):
        """slices [n_streams, slice_size] int16 -> probabilities [n_streams, labels] (moving-average filtered) or None
        while the feature window is still filling"""
        slices = np.ascontiguousarray(slices, dtype=np.int16).reshape(self.n_streams, self.slice_size)
        probs = np.zeros((self.n_streams, self.impulse.label_count), np.float32)
        has = C.c_int(0)
        _check(self._lib.eikws_streams_push_i16_host(self._s, _np_ptr(slices), C.c_float(beyond), _np_ptr(probs), C.byref(has)))
        return probs if has.value else None

    def push_device(self, slices, probs, beyond: float = 0.0) -> bool:
        import torch
        assert slices.is_cuda and slices.dtype == torch.int16 and slices.is_contiguous()
        has = C.c_int(0)
        stream = C.c_void_p(torch.cuda.current_stream(slices.device).cuda_stream)
        _check(self._lib.eikws_streams_push_i16_device(self._s, C.c_void_p(slices.data_ptr()), C.c_float(beyond), C.c_void_p(probs.data_ptr()),
                                                       C.byref(has), stream))
        return bool(has.value)


def debug_host_plan(model="l476"):
    """Host-side derived tables (no GPU needed): dense mel filterbank [129,32] and the conv/FC requantisation
    multipliers/shifts, for parity tests against the oracle."""
    L = load_library()
    blob = model_blob(model) if isinstance(model, str) else bytes(model)
    fb = np.zeros((129, 32), np.float32)
    mult = np.zeros(256, np.int32)
    shift = np.zeros(256, np.int32)
    n = C.c_int(0)
    _check(L.eikws_debug_host_plan(blob, len(blob), _np_ptr(fb), _np_ptr(mult), _np_ptr(shift), 256, C.byref(n)))
    return fb, mult[: n.value].copy(), shift[: n.value].copy()
