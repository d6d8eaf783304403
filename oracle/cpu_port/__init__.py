"""ORACLE — TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Loader of the best-effort C statement of the gated path
(oracle/cpu_port/ewk_cpu.c): compiled with the host's gcc into a temporary directory at first use (-march=native: a
binary built in one container must not travel to another CPU), tables taken from oracle/librosa_restated.py.
Only tests/ and bench.py's cpu_baseline leg import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ewk_cpu.c")
_lib = None


class Params(C.Structure):
    _fields_ = [("similarity_threshold", C.c_double), ("pre_speech_silence", C.c_double), ("speech_duration_min", C.c_double),
                ("speech_duration_max", C.c_double), ("post_speech_silence", C.c_double), ("timeout", C.c_double),
                ("ring_samples", C.c_int)]


EVENT_DTYPE = np.dtype([("tick", "<i4"), ("seg_len", "<i4"), ("matched", "<i4"), ("score", "<f4")])


def compile_to(out_dir: str, native: bool = True) -> str:
    """gcc -O3 [-march=native] -fopenmp -shared -fPIC ewk_cpu.c -> out_dir/libewk_cpu.so"""
    out = os.path.join(out_dir, "libewk_cpu.so")
    cmd = ["gcc", "-O3", "-std=gnu11", "-fopenmp", "-shared", "-fPIC", "-o", out, SRC, "-lm"]
    if native:
        cmd.insert(2, "-march=native")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed on oracle/cpu_port/ewk_cpu.c:\n" + r.stderr)
    return out


def load():
    global _lib
    if _lib is not None:
        return _lib
    import scipy.fft
    from .. import librosa_restated as L
    lib = C.CDLL(compile_to(tempfile.mkdtemp(prefix="ewk_cpu_")))
    f32p, i16p = C.POINTER(C.c_float), C.POINTER(C.c_int16)
    lib.ewk_cpu_init.argtypes = [f32p, f32p, C.POINTER(C.c_double)]
    lib.ewk_cpu_features.argtypes = [i16p, C.c_int, f32p, f32p]
    lib.ewk_cpu_score.argtypes = [f32p, f32p, f32p, f32p]
    lib.ewk_cpu_score.restype = C.c_float
    lib.ewk_cpu_detect_batch.argtypes = [i16p, C.c_int, C.c_int64, C.c_int64, f32p, f32p, C.POINTER(Params), C.c_void_p, C.c_int,
                                         C.POINTER(C.c_int32), C.c_int]
    mel = np.ascontiguousarray(L.mel_filterbank(), dtype=np.float32)
    dct = np.ascontiguousarray(scipy.fft.dct(np.eye(L.N_MELS), axis=0, type=2, norm="ortho")[:L.N_MFCC], dtype=np.float32)
    hann = np.ascontiguousarray(L.hann_window(), dtype=np.float64)
    rc = lib.ewk_cpu_init(mel.ctypes.data_as(f32p), dct.ctypes.data_as(f32p), hann.ctypes.data_as(C.POINTER(C.c_double)))
    if rc != 0:
        raise RuntimeError("ewk_cpu_init failed")
    _lib = lib
    return lib


def features(pcm_i16):
    """WordMatcher.extract_mfcc on int16 PCM -> (mean[20], std[20]) float32."""
    lib = load()
    q = np.ascontiguousarray(pcm_i16, dtype=np.int16)
    mean, std = np.empty(20, np.float32), np.empty(20, np.float32)
    f32p = C.POINTER(C.c_float)
    rc = lib.ewk_cpu_features(q.ctypes.data_as(C.POINTER(C.c_int16)), len(q), mean.ctypes.data_as(f32p), std.ctypes.data_as(f32p))
    if rc < 0:
        raise ValueError(f"ewk_cpu_features: {rc}")
    return mean, std


def detect_batch(pcm_i16, template_i16, *, ring_seconds=10, threads=None, cap=256, similarity_threshold=75.0,
                 pre_speech_silence=0.8, speech_duration_min=0.3, speech_duration_max=2.0, post_speech_silence=0.4, timeout=30.0):
    """pcm_i16 [n_streams, n] -> list of per-stream structured event arrays (tick, seg_len, matched, score)."""
    lib = load()
    q = np.ascontiguousarray(pcm_i16, dtype=np.int16)
    ns, n = q.shape
    rm, rs = features(template_i16)
    prm = Params(similarity_threshold, pre_speech_silence, speech_duration_min, speech_duration_max, post_speech_silence,
                 timeout, ring_seconds * 16000)
    ev = np.zeros((ns, cap), dtype=EVENT_DTYPE)
    counts = np.zeros(ns, np.int32)
    f32p = C.POINTER(C.c_float)
    rc = lib.ewk_cpu_detect_batch(q.ctypes.data_as(C.POINTER(C.c_int16)), ns, n, n, rm.ctypes.data_as(f32p), rs.ctypes.data_as(f32p),
                                  C.byref(prm), ev.ctypes.data, cap, counts.ctypes.data_as(C.POINTER(C.c_int32)),
                                  threads or (os.cpu_count() or 1))
    if rc != 0:
        raise RuntimeError("ewk_cpu_detect_batch failed")
    return [ev[s, :min(int(counts[s]), cap)] for s in range(ns)]
