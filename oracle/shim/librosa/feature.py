"""ORACLE SHIM — librosa.feature subset used by the reference (wakeword.py:561-563, 878)."""
from oracle import librosa_restated as _r


def mfcc(y=None, sr=22050, n_mfcc=20, n_fft=2048, hop_length=512, **kw):
    if kw:
        raise TypeError(f"oracle shim: unsupported mfcc kwargs {sorted(kw)}")
    return _r.mfcc(y, sr=sr, n_mfcc=n_mfcc, n_fft=n_fft, hop_length=hop_length)


def rms(y=None, frame_length=2048, hop_length=512, **kw):
    if kw:
        raise TypeError(f"oracle shim: unsupported rms kwargs {sorted(kw)}")
    return _r.rms(y, frame_length=frame_length, hop_length=hop_length)
