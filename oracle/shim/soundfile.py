"""ORACLE SHIM — `soundfile` subset (PCM16 WAV through stdlib `wave`).  Test infrastructure only.
libsndfile semantics: float -> int16 on write is lrint(x*32767); int16 -> float on read is /32768."""
import numpy as np

from oracle import librosa_restated as _r


def write(file, data, samplerate, subtype=None, **kw):
    _r.wav_write_pcm16(file, data, samplerate)


def read(file, dtype="float64", always_2d=False, **kw):
    pcm, sr = _r.wav_read_pcm16(file)
    dt = np.dtype(dtype)
    out = pcm.astype(dt) if dt.kind == "i" else pcm.astype(dt) / dt.type(32768.0)
    if always_2d and out.ndim == 1:
        out = out[:, None]
    return out, sr
