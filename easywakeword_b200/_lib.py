"""ctypes binding of libewk.so (include/ewk.h).  Fails loudly when the CUDA library is missing —
there is no CPU fallback in this package."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libewk.so")

EWK_OK, EWK_ERR_ARG, EWK_ERR_NO_TEMPLATE, EWK_ERR_CUDA, EWK_ERR_STATE, EWK_ERR_NOMEM = 0, -1, -2, -3, -4, -5
PCM_F32, PCM_I16 = 0, 1
HOST, DEVICE = 0, 1


class EwkError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("n_streams", C.c_int32), ("ring_samples", C.c_int32), ("slack_samples", C.c_int32),
                ("pcm_format", C.c_int32), ("max_templates", C.c_int32), ("max_events", C.c_int32)]


_lib = None


def _p(t):
    return C.POINTER(t)


def load():
    """dlopen libewk.so (built in-tree by easywakeword_b200.build) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EwkError(f"{LIB_PATH} is missing: build it with `python -m easywakeword_b200.build` "
                       "(nvcc, sm_100a). easywakeword_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32p = C.c_void_p, C.c_int, C.c_int64, _p(C.c_float)
    protos = {
        "ewk_abi_version": (C.c_int, []),
        "ewk_last_error": (C.c_char_p, [vp]),
        "ewk_host_table": (C.c_int, [i32, f32p, i32]),
        "ewk_create": (C.c_int, [i32, _p(Config), _p(vp)]),
        "ewk_destroy": (C.c_int, [vp]),
        "ewk_set_cuda_stream": (C.c_int, [vp, vp]),
        "ewk_synchronize": (C.c_int, [vp]),
        "ewk_extract_mfcc": (C.c_int, [vp, vp, i64, i32, f32p, f32p, f32p, i64]),
        "ewk_set_template": (C.c_int, [vp, i32, f32p, i64]),
        "ewk_set_template_features": (C.c_int, [vp, i32, f32p, f32p, i64]),
        "ewk_get_template": (C.c_int, [vp, i32, f32p, f32p, _p(i64)]),
        "ewk_clear_template": (C.c_int, [vp, i32]),
        "ewk_similarity_batch": (C.c_int, [vp, i32, vp, i32, i32, _p(i64), _p(i64), i32, C.c_float, f32p,
                                           _p(C.c_uint8), f32p]),
    }
    for name, (res, args) in protos.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._protos = protos
    _lib = lib
    return lib


def f32ptr(a):
    return a.ctypes.data_as(_p(C.c_float))


def i64ptr(a):
    return a.ctypes.data_as(_p(C.c_int64))


def u8ptr(a):
    return a.ctypes.data_as(_p(C.c_uint8))


def host_table(which):
    lib = load()
    n = lib.ewk_host_table(which, None, 0)
    out = np.empty(n, dtype=np.float32)
    lib.ewk_host_table(which, f32ptr(out), n)
    return out


def check(lib, ctx, rc):
    """Map ewk_status to the reference's exception types (SURVEY §8(b) 'Errors')."""
    if rc == EWK_OK:
        return
    msg = lib.ewk_last_error(ctx).decode("utf-8", "replace")
    if rc in (EWK_ERR_ARG, EWK_ERR_NO_TEMPLATE):
        raise ValueError(msg)
    if rc == EWK_ERR_NOMEM:
        raise MemoryError(msg)
    raise EwkError(f"libewk error {rc}: {msg}")


class Context:
    """Owns one ewk_ctx (one GPU)."""

    def __init__(self, device=0, n_streams=0, ring_samples=160000, slack_samples=16000, pcm_format=PCM_I16,
                 max_templates=4, max_events=0):
        self.lib = load()
        self.cfg = Config(n_streams, ring_samples, slack_samples, pcm_format, max_templates,
                          max_events or max(1024, 4 * n_streams))
        h = C.c_void_p()
        rc = self.lib.ewk_create(device, C.byref(self.cfg), C.byref(h))
        if rc != EWK_OK:
            msg = self.lib.ewk_last_error(None).decode("utf-8", "replace")
            if rc == EWK_ERR_ARG:
                raise ValueError(msg)
            raise EwkError(f"libewk error {rc}: {msg}")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.ewk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        check(self.lib, self.h, rc)

    # ---- level 2
    def extract_mfcc(self, audio, want_frames=False):
        a = np.ascontiguousarray(audio, dtype=np.float32).reshape(-1)
        mean = np.empty(20, np.float32)
        std = np.empty(20, np.float32)
        frames = None
        if want_frames:
            nf = 1 + len(a) // 160
            frames = np.empty((nf, 20), np.float32)
        self._ck(self.lib.ewk_extract_mfcc(self.h, a.ctypes.data, len(a), HOST, f32ptr(mean), f32ptr(std),
                                           f32ptr(frames) if want_frames else None, 0 if frames is None else len(frames)))
        return (mean, std, frames) if want_frames else (mean, std)

    def set_template(self, slot, audio):
        a = np.ascontiguousarray(audio, dtype=np.float32).reshape(-1)
        self._ck(self.lib.ewk_set_template(self.h, slot, f32ptr(a), len(a)))

    def set_template_features(self, slot, mean, std, n_samples):
        m = np.ascontiguousarray(mean, dtype=np.float32)
        s = np.ascontiguousarray(std, dtype=np.float32)
        self._ck(self.lib.ewk_set_template_features(self.h, slot, f32ptr(m), f32ptr(s), int(n_samples)))

    def get_template(self, slot):
        m = np.empty(20, np.float32)
        s = np.empty(20, np.float32)
        n = C.c_int64()
        self._ck(self.lib.ewk_get_template(self.h, slot, f32ptr(m), f32ptr(s), C.byref(n)))
        return m, s, n.value

    def clear_template(self, slot):
        self._ck(self.lib.ewk_clear_template(self.h, slot))

    def similarity_batch(self, slot, pcm, offsets, lens, threshold=75.0, want_features=False):
        pcm = np.ascontiguousarray(pcm).reshape(-1)
        if pcm.dtype == np.int16:
            fmt = PCM_I16
        else:
            pcm = pcm.astype(np.float32, copy=False)
            fmt = PCM_F32
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        ln = np.ascontiguousarray(lens, dtype=np.int64)
        n = len(off)
        scores = np.empty(n, np.float32)
        matched = np.empty(n, np.uint8)
        feats = np.empty((n, 40), np.float32) if want_features else None
        self._ck(self.lib.ewk_similarity_batch(self.h, slot, pcm.ctypes.data, fmt, HOST, i64ptr(off), i64ptr(ln), n,
                                               float(threshold), f32ptr(scores), u8ptr(matched),
                                               f32ptr(feats) if want_features else None))
        return (scores, matched.astype(bool), feats) if want_features else (scores, matched.astype(bool))

    def synchronize(self):
        self._ck(self.lib.ewk_synchronize(self.h))
