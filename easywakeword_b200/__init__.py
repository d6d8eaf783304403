"""easywakeword_b200 — B200-native hot path of EasyWakeWord behind the reference's own surface.

    from easywakeword_b200 import WakeWord            # drop-in for easywakeword.WakeWord
    from easywakeword_b200 import WakeWordBank        # N streams on one GPU (the multiroom case)

Importing the package needs neither a GPU nor the built library; constructing a detector does
(there is no CPU fallback).  Build the library with `python -m easywakeword_b200.build`.
"""
__version__ = "0.1.0"
__all__ = ["WakeWord", "WordMatcher", "SoundBuffer", "WakeWordBank"]


def __getattr__(name):
    if name in ("WakeWord", "WordMatcher", "SoundBuffer"):
        from . import wakeword
        return getattr(wakeword, name)
    if name == "WakeWordBank":
        from .bank import WakeWordBank
        return WakeWordBank
    raise AttributeError(name)
