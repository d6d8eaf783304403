"""Synthetic 16 kHz PCM streams for tests and bench (no dataset exists offline).

Shapes follow SURVEY.md §8(d): stationary Gaussian background well below / around the
reference's silence floor (SoundBuffer.MIN_THRESHOLD = 0.005, wakeword.py:409) with the
wake-word template mixed in at random sample offsets and gains.  Every stream is a pure
function of (seed, stream id), so any rank can regenerate any stream.

The reference's own tests use the same kind of signals: sines, a 4-partial "speech-like"
tone and seed-42 Gaussian noise (tests/test_wakeword_simulated.py:47-69, 166-167).
"""
from __future__ import annotations

import numpy as np

SR = 16000


def sine(freq=440.0, duration=1.0, amp=0.5, sr=SR):
    """generate_wav of the reference's tests (tests/test_wakeword_simulated.py:47-52)."""
    t = np.linspace(0, duration, int(sr * duration), endpoint=False)
    return (amp * np.sin(2 * np.pi * freq * t)).astype(np.float32)


def speech_like(duration=1.0, sr=SR):
    """generate_speech_like_audio of the reference's tests (tests/test_wakeword_simulated.py:55-69)."""
    t = np.linspace(0, duration, int(sr * duration), endpoint=False)
    audio = (0.3 * np.sin(2 * np.pi * 150 * t) + 0.2 * np.sin(2 * np.pi * 500 * t)
             + 0.15 * np.sin(2 * np.pi * 1500 * t) + 0.1 * np.sin(2 * np.pi * 2500 * t))
    envelope = np.sin(np.pi * t / duration) ** 0.5
    return (audio * envelope).astype(np.float32)


def synthetic_word(seed=0, duration=0.97, sr=SR):
    """A deterministic word-shaped template for boxes without the bundled WAV: three voiced
    'syllables' (harmonic stacks with moving formant weights) under raised-cosine envelopes."""
    rng = np.random.default_rng(seed)
    n = int(duration * sr)
    t = np.arange(n) / sr
    y = np.zeros(n)
    edges = np.linspace(0.04, duration - 0.04, 4)
    for i in range(3):
        a, b = edges[i], edges[i + 1] - 0.03
        env = np.where((t >= a) & (t <= b), 0.5 - 0.5 * np.cos(2 * np.pi * (t - a) / (b - a)), 0.0)
        f0 = rng.uniform(110, 180)
        formants = rng.uniform([300, 900, 2200], [800, 1800, 3200])
        syl = np.zeros(n)
        for h in range(1, 40):
            f = h * f0
            if f > 7000:
                break
            w = sum(np.exp(-0.5 * ((f - fc) / 120.0) ** 2) for fc in formants) + 0.02
            syl += w * np.sin(2 * np.pi * f * t + rng.uniform(0, 2 * np.pi))
        y += env * syl
    y *= 0.11 / np.max(np.abs(y))
    return y.astype(np.float32)


def to_int16(x):
    """float PCM in [-1, 1) -> int16, the device ring's compact format (x == q / 32768 exactly)."""
    return np.clip(np.rint(np.asarray(x, dtype=np.float64) * 32768.0), -32768, 32767).astype(np.int16)


def from_int16(q):
    return q.astype(np.float32) / np.float32(32768.0)


def _distractor(rng, n, sr=SR):
    """A non-word burst of word length: an enveloped two-tone or a noise burst (same RMS scale as the word)."""
    t = np.arange(n) / sr
    env = np.sin(np.pi * np.arange(n) / n) ** 0.5
    if rng.random() < 0.5:
        f = rng.uniform(200, 3000)
        y = 0.03 * (np.sin(2 * np.pi * f * t) + 0.5 * np.sin(2 * np.pi * 2.7 * f * t)) * env
    else:
        y = rng.standard_normal(n) * 0.02 * env
    return y.astype(np.float32)


def stream(seed, seconds, word, *, noise_sigma=0.002, inserts_per_10s=(1, 3), gain=(1.0, 4.0),
           zero_gaps=0, distractor_prob=0.0, sr=SR):
    """One float32 stream: N(0, noise_sigma^2) background + `word` inserted 1..3 times per 10 s at
    random sample offsets with gain U(gain).  zero_gaps > 0 blanks that many 0.3 s stretches to exact
    zeros (exercises the top_db floor / amin path).  With distractor_prob > 0 an insertion is replaced
    by a non-word burst with that probability (offset recorded with gain < 0).
    Returns (pcm_f32, inserts[(offset, gain)])."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    x = rng.standard_normal(n).astype(np.float32) * np.float32(noise_sigma)
    inserts = []
    lw = len(word)
    for w0 in range(0, n, 10 * sr):
        w1 = min(n, w0 + 10 * sr)
        if w1 - w0 < lw + 2:
            break
        k = int(rng.integers(inserts_per_10s[0], inserts_per_10s[1] + 1))
        placed = []
        for _ in range(k):
            for _try in range(20):
                off = int(rng.integers(w0, w1 - lw))
                if all(abs(off - p) > lw for p in placed):
                    placed.append(off)
                    break
        for off in sorted(placed):
            g = float(rng.uniform(*gain))
            if distractor_prob > 0 and rng.random() < distractor_prob:
                x[off:off + lw] += np.float32(g) * _distractor(rng, lw, sr)
                inserts.append((off, -g))
            else:
                x[off:off + lw] += np.float32(g) * word
                inserts.append((off, g))
    for _ in range(zero_gaps):
        off = int(rng.integers(0, max(1, n - int(0.3 * sr))))
        x[off:off + int(0.3 * sr)] = 0.0
    return x, inserts


def stream_batch(seed0, n_streams, seconds, word, *, as_int16=True, **kw):
    """[n_streams, seconds*sr] batch; stream s uses seed0 + s."""
    n = int(round(seconds * SR))
    out = np.empty((n_streams, n), dtype=np.int16 if as_int16 else np.float32)
    for s in range(n_streams):
        x, _ = stream(seed0 + s, seconds, word, **kw)
        out[s] = to_int16(x) if as_int16 else x
    return out
