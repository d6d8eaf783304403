"""ORACLE SHIM — test infrastructure only.

A stand-in for the un-vendored `librosa` wheel (0.11.0, /root/reference/uv.lock:324-325)
so that /root/reference/easywakeword/wakeword.py can be imported UNMODIFIED in this
container.  Everything forwards to oracle/librosa_restated.py.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from oracle import librosa_restated as _r  # noqa: E402

from . import feature  # noqa: E402,F401

__version__ = "0.11.0+restated"


def load(path, sr=22050, mono=True, **kw):
    return _r.load(path, sr=sr)


def resample(y, orig_sr, target_sr, **kw):
    if orig_sr == target_sr:
        return y
    raise NotImplementedError("oracle shim: resampling is not restated")
