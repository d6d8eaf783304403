"""ORACLE — TEST INFRASTRUCTURE ONLY.  Runs the REFERENCE'S OWN CODE, unmodified.

Imports /root/reference/easywakeword/wakeword.py as it lies (read-only) on top of the
three shims in oracle/shim (librosa restatement, soundfile via stdlib `wave`, inert
sounddevice) and drives its classes deterministically:

  * WordMatcher / SoundBuffer are the reference's bytes (wakeword.py:405-639);
  * WakeWord._detect_word (wakeword.py:1036-1159) is driven by a fake clock that
    replaces the module global ``time``: ``time() = k * 0.1`` (float64, integer k)
    and ``sleep(dt)`` advances k by one tick and feeds that tick's audio through the
    real ``SoundBuffer._add_sound_to_buffer`` in ``block``-sample callbacks.
  * level 3 (``_transcribe_audio``) is stubbed to return None (out of scope).

/root/reference does not exist on the GPU box.  HERE this module is used by oracle/gen_golden.py
(to write tests/golden/*) and by the CPU-side tests that pin oracle/ewk_oracle.py against the
reference.  On the GPU box the only consumer is bench.py's CPU arm (`--impl reference`,
`cpu_baseline`): when oracle/stage_ref.py has staged the reference's unmodified module under
oracle/_ref/ (git-ignored, travels with gpurun), the arm times THESE classes (kind "reference")
instead of the oracle port.  Nothing under -m gpu or smoke() imports it.
"""
from __future__ import annotations

import importlib
import os
import sys
import threading

import numpy as np

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _find_root():
    """EWK_REFERENCE_ROOT, else the read-only tree, else the copy staged by oracle/stage_ref.py."""
    env = os.environ.get("EWK_REFERENCE_ROOT")
    for root in ([env] if env else ["/root/reference", _STAGED]):
        if os.path.isfile(os.path.join(root, "easywakeword", "wakeword.py")):
            return root
    return env or "/root/reference"


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "easywakeword", "wakeword.py"))


def import_reference():
    """-> the reference's `easywakeword.wakeword` module (unmodified source)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (_REPO, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    # the shims must win over any real wheel, the reference package over ours
    if sys.path[0] != _SHIM:
        sys.path.remove(_SHIM)
        sys.path.insert(0, _SHIM)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(1, REFERENCE_ROOT)
    for name in ("librosa", "soundfile", "sounddevice"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(_SHIM):
            del sys.modules[name]
    mod = importlib.import_module("easywakeword.wakeword")
    assert mod.__file__.startswith(REFERENCE_ROOT), mod.__file__
    return mod


class FakeClock:
    """Stands in for the `time` module inside the reference (wakeword.py:15)."""

    def __init__(self, stream, block, tick_samples=1600, tick_seconds=0.1):
        self.k = 0
        self.stream = np.asarray(stream, dtype=np.float32)
        self.block = int(block)
        self.tick_samples = int(tick_samples)
        self.tick_seconds = tick_seconds
        self.fed = 0
        self.buffer = None  # reference SoundBuffer
        self.exhausted = False

    def time(self):
        return self.k * self.tick_seconds

    def sleep(self, dt):
        target = ((self.k + 1) * self.tick_samples // self.block) * self.block
        if target > len(self.stream):
            self.exhausted = True
            raise StopIteration("stream exhausted")
        self.k += 1
        while self.fed < target:
            blk = self.stream[self.fed : self.fed + self.block].reshape(-1, 1)
            self.buffer._add_sound_to_buffer(blk, self.block, None, None)
            self.fed += self.block


def make_wakeword(mod, wavword, *, textword="computer", numberofwords=1, timeout=30,
                  similarity_threshold=75.0, pre_speech_silence=0.8, speech_duration_min=0.69,
                  speech_duration_max=1.38, post_speech_silence=0.4, buffer_seconds=10):
    """A reference WakeWord without __init__ (HEAD's __init__ calls a method that does not
    exist, wakeword.py:786) — the same bypass the reference's tests use
    (tests/test_helpers.py:65-126)."""
    ww = object.__new__(mod.WakeWord)
    ww.textword = textword
    ww.wavword = str(wavword)
    ww.numberofwords = numberofwords
    ww.timeout = timeout
    ww.callback = None
    ww.device = None
    ww.similarity_threshold = similarity_threshold
    ww.buffer_seconds = buffer_seconds
    ww.verbose = False
    ww.retry_count = 3
    ww.retry_backoff = 0.5
    ww.pre_speech_silence = pre_speech_silence
    ww.post_speech_silence = post_speech_silence
    ww.speech_duration_min = speech_duration_min
    ww.speech_duration_max = speech_duration_max
    ww._sound_buffer = None
    ww._matcher = None
    ww._listening = False
    ww._listen_thread = None
    ww._stop_event = threading.Event()
    ww._transcribe_audio = lambda audio: None  # level 3 stubbed
    return ww


def run_reference_stream(stream, template, *, block=512, restart_on_timeout=True, timing=None, keep_audio=True, **params):
    """Drive the reference's SoundBuffer + WordMatcher + _detect_word over `stream`.

    template: float32 samples (set_reference) or a WAV path (load_reference_from_file).
    Returns a dict with the per-tick is_silent / silence_threshold trace and the list of
    level-2 events (tick, seg_len, score, matched) plus timeouts.
    """
    mod = import_reference()
    clock = FakeClock(stream, block)
    saved_time = mod.time
    mod.time = clock
    try:
        ww = make_wakeword(mod, template if isinstance(template, (str, os.PathLike)) else "<mem>", **params)
        buf = mod.SoundBuffer(seconds=ww.buffer_seconds, device=None)
        clock.buffer = buf
        ww._sound_buffer = buf
        matcher = mod.WordMatcher(sample_rate=mod.SoundBuffer.FREQUENCY)
        if isinstance(template, (str, os.PathLike)):
            matcher.load_reference_from_file(str(template), ww.textword)
        else:
            matcher.set_reference(np.asarray(template, dtype=np.float32), ww.textword)
        ww._matcher = matcher

        trace = {"tick": [], "silent": [], "thr": []}
        events = []
        timeouts = []

        real_is_silent = buf.is_silent

        def is_silent_logged():
            r = bool(real_is_silent())
            trace["tick"].append(clock.k)
            trace["silent"].append(r)
            trace["thr"].append(float(buf.silence_threshold))
            return r

        buf.is_silent = is_silent_logged
        real_matches = matcher.matches

        def matches_logged(audio, threshold=75.0):
            ok, sim = real_matches(audio, threshold=threshold)
            ev = {"tick": clock.k, "seg_len": int(len(audio)), "score": float(sim), "matched": bool(ok)}
            if keep_audio:
                ev["audio"] = np.array(audio, dtype=np.float64)
            events.append(ev)
            return ok, sim

        matcher.matches = matches_logged

        full_tick = None
        try:
            ww._wait_for_buffer()
            full_tick = clock.k
            if timing is not None:                     # bench.py: steady state only (the ring fill is untimed)
                import time as _time
                timing["t_full"] = _time.perf_counter()
            while True:
                try:
                    ww._detect_word()
                except TimeoutError:
                    timeouts.append(clock.k)
                    if not restart_on_timeout:
                        break
        except StopIteration:
            pass
        if timing is not None:
            import time as _time
            timing["t_end"] = _time.perf_counter()
            timing["steady_ticks"] = clock.k - (full_tick if full_tick is not None else clock.k)
        return {
            "full_tick": full_tick,
            "trace_tick": np.asarray(trace["tick"], dtype=np.int64),
            "trace_silent": np.asarray(trace["silent"], dtype=np.bool_),
            "trace_thr": np.asarray(trace["thr"], dtype=np.float64),
            "events": events,
            "timeouts": timeouts,
            "frame_size": int(buf.frame_size),
            "ticks_run": clock.k,
        }
    finally:
        mod.time = saved_time
