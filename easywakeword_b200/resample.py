"""Sample-rate conversion for templates / feeds that are not 16 kHz.

The reference delegates this to librosa.load / librosa.resample (soxr_hq; wakeword.py:588, 866-870).
It is outside the round-1 hot path; until the device resampler lands this fails loudly rather than
silently producing features at the wrong rate."""


def resample_to_16k(y, sr_native, sr_target=16000):
    raise NotImplementedError(
        f"audio at {sr_native} Hz must be resampled to {sr_target} Hz before use; "
        "easywakeword_b200 does not resample yet (SURVEY §8(f) N3)")
