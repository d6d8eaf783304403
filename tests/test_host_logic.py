"""CPU: host-side mirror of the reference interface — constructor validation and messages, the
auto-calculated speech durations, the timing state machine against the reference-generated goldens,
timeout behaviour with duck-typed buffers (the reference's own seam, tests/test_wakeword_simulated.py:
301-328), WAV loading, sharding arithmetic."""
import os
import threading
import time

import numpy as np
import pytest

from easywakeword_b200 import synth
from easywakeword_b200.wakeword import TimingMachine, WakeWord, analyze_reference_audio_duration, load_wav_16k
from easywakeword_b200.dist import owner_of, padded_shard, shard_range


@pytest.fixture()
def wav(tmp_path, word_i16):
    import wave
    p = tmp_path / "word.wav"
    with wave.open(str(p), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(word_i16.tobytes())
    return str(p)


# ---- constructor validation: same messages as wakeword.py:743-763 (reference tests :520-661)
@pytest.mark.parametrize("kw,msg", [
    (dict(numberofwords=0), "numberofwords must be at least 1"),
    (dict(buffer_seconds=0), "buffer_seconds must be positive"),
    (dict(retry_count=-1), "retry_count must be non-negative"),
    (dict(retry_backoff=-0.1), "retry_backoff must be non-negative"),
    (dict(pre_speech_silence=0), "pre_speech_silence must be positive"),
    (dict(speech_duration_min=-1), "speech_duration_min must be positive"),
    (dict(speech_duration_max=0), "speech_duration_max must be positive"),
    (dict(speech_duration_min=2.0, speech_duration_max=1.0), "speech_duration_min must be <= speech_duration_max"),
    (dict(post_speech_silence=0), "post_speech_silence must be positive"),
])
def test_constructor_validation(wav, kw, msg):
    with pytest.raises(ValueError, match=msg):
        WakeWord("hello", wav, **kw)


def test_defaults_and_auto_durations(wav):
    ww = WakeWord("OK Computer ", wav)
    assert ww.textword == "ok computer" and ww.numberofwords == 2 and ww.timeout == 30
    assert ww.similarity_threshold == 75.0 and ww.pre_speech_silence == 0.8 and ww.post_speech_silence == 0.4
    assert ww.buffer_seconds == 10 and ww.retry_count == 3 and ww.retry_backoff == 0.5
    # reference_word.wav: VAD duration 0.69 s, max = 2 x min
    assert ww.speech_duration_min == pytest.approx(0.69) and ww.speech_duration_max == pytest.approx(1.38)
    assert not ww.is_listening()
    ww2 = WakeWord("hello", wav, speech_duration_min=0.5)
    assert (ww2.speech_duration_min, ww2.speech_duration_max) == (0.5, 1.0)
    ww3 = WakeWord("hello", wav, speech_duration_min=0.4, speech_duration_max=1.5)
    assert (ww3.speech_duration_min, ww3.speech_duration_max) == (0.4, 1.5)
    ww4 = WakeWord("hello", "/nonexistent.wav")
    assert (ww4.speech_duration_min, ww4.speech_duration_max) == (0.3, 2.0)


def test_vad_duration_equals_oracle(word):
    from oracle import ewk_oracle as O
    for a in (word, synth.speech_like(0.8), synth.speech_like(0.5), synth.sine(440)):
        assert analyze_reference_audio_duration(a) == O.analyze_reference_audio_duration(a)
    assert 0.2 <= analyze_reference_audio_duration(synth.speech_like(1.0)) <= 1.5   # tests/test_wakeword_simulated.py:238-249


def test_start_requires_callback(wav):
    with pytest.raises(ValueError, match="Callback must be set"):
        WakeWord("hello", wav).start()


def test_load_wav(wav, word):
    assert np.array_equal(load_wav_16k(wav), word)


def test_read_wav_formats_host_side(tmp_path, word_i16):
    """The RIFF reader behind load_16k (librosa.load's decoding step: libsndfile float conversion) on the encodings
    libsndfile reads from WAV: PCM 8/16/32, IEEE float, and G.711 mu-law / A-law (format tags 7 / 6), whose expansion
    tables are pinned against CPython's audioop where it still exists (< 3.13)."""
    import struct
    import warnings
    from easywakeword_b200.resample import ALAW_TABLE, ULAW_TABLE, read_wav

    def write(name, tag, bits, ch, sr, payload):
        block = ch * bits // 8
        hdr = struct.pack("<4sI4s4sIHHIIHH4sI", b"RIFF", 36 + len(payload), b"WAVE", b"fmt ", 16, tag, ch, sr, sr * block, block,
                          bits, b"data", len(payload))
        p = tmp_path / name
        p.write_bytes(hdr + payload + (b"\0" if len(payload) & 1 else b""))
        return p

    q = word_i16[:4001]
    y, sr = read_wav(write("p16.wav", 1, 16, 1, 16000, q.astype("<i2").tobytes()))
    assert sr == 16000 and np.array_equal(y, q.astype(np.float32) / np.float32(32768))
    y, _ = read_wav(write("p8.wav", 1, 8, 1, 8000, bytes(range(256))))
    assert np.array_equal(y, (np.arange(256, dtype=np.float32) - 128) / 128)
    y, _ = read_wav(write("p32.wav", 1, 32, 2, 44100, (q[:4000].astype(np.int32) << 16).astype("<i4").tobytes()))
    assert y.shape == (2000, 2) and np.array_equal(y.reshape(-1), q[:4000].astype(np.float32) / np.float32(32768))
    y, _ = read_wav(write("f32.wav", 3, 32, 1, 48000, (q.astype(np.float32) / 32768).astype("<f4").tobytes()))
    assert np.array_equal(y, q.astype(np.float32) / np.float32(32768))
    codes = np.arange(256, dtype=np.uint8)
    for tag, table in ((7, ULAW_TABLE), (6, ALAW_TABLE)):
        y, sr = read_wav(write(f"g711_{tag}.wav", tag, 8, 1, 8000, codes.tobytes()))
        assert sr == 8000 and np.array_equal(y, table.astype(np.float32) / np.float32(32768))
    assert ULAW_TABLE[0] == -32124 and ULAW_TABLE[0x80] == 32124 and ULAW_TABLE[0xFF] == 0 and ULAW_TABLE[0x7F] == 0
    assert ALAW_TABLE[0x2A] == -32256 and ALAW_TABLE[0xAA] == 32256 and ALAW_TABLE[0xD5] == 8 and ALAW_TABLE[0x55] == -8
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import audioop
    except ImportError:
        audioop = None
    if audioop is not None:
        assert np.array_equal(np.frombuffer(audioop.ulaw2lin(codes.tobytes(), 2), "<i2"), ULAW_TABLE)
        assert np.array_equal(np.frombuffer(audioop.alaw2lin(codes.tobytes(), 2), "<i2"), ALAW_TABLE)
    with pytest.raises(ValueError):
        read_wav(write("adpcm.wav", 2, 4, 1, 8000, b"\0" * 64))


# ---- timing state machine vs goldens from the reference's _detect_word
def test_timing_machine_reproduces_reference_events(golden_detect):
    g, cases = golden_detect
    checked = 0
    for c in cases:
        n = c["name"]
        p = dict(pre_speech_silence=0.8, speech_duration_min=0.3, speech_duration_max=2.0, post_speech_silence=0.4,
                 timeout=30.0)
        p.update(c["params"])
        ticks, silent = g[f"{n}_trace_tick"], g[f"{n}_trace_silent"]
        m = TimingMachine(p["pre_speech_silence"], p["speech_duration_min"], p["speech_duration_max"], p["post_speech_silence"])
        events, timeouts = [], []
        i = 0
        start = ticks[0] * 0.1
        m.enter(bool(silent[0]), ticks[0] * 0.1)
        i = 1
        while i < len(ticks):
            k = int(ticks[i])
            if ticks[i] == ticks[i - 1]:                 # the reference re-sampled after a timeout
                timeouts.append(k)
                m = TimingMachine(p["pre_speech_silence"], p["speech_duration_min"], p["speech_duration_max"], p["post_speech_silence"])
                m.enter(bool(silent[i]), k * 0.1)
                i += 1
                continue
            cut = m.step(bool(silent[i]), k * 0.1)
            if cut is not None:
                back, n_drop = cut
                seg_len = int(back * 16000) - n_drop
                if seg_len / 16000 <= 3.0:
                    events.append((k, seg_len))
            i += 1
        assert [e[0] for e in events] == list(g[f"{n}_ev_tick"]), n
        assert [e[1] for e in events] == list(g[f"{n}_ev_len"]), n
        assert timeouts == list(g[f"{n}_timeouts"]), n
        checked += len(events)
    assert checked >= 60


# ---- the reference's own seam: duck-typed buffer / matcher objects on a WakeWord
class _SilentBuffer:
    def is_buffer_full(self): return True
    def is_silent(self): return True
    def stop(self): pass
    def return_last_n_seconds(self, n): return np.zeros(int(n * 16000), dtype=np.float32)


def test_detect_word_times_out_on_silent_buffer(wav):
    ww = WakeWord("hello", wav, numberofwords=1, timeout=1, pre_speech_silence=0.5, speech_duration_min=0.3,
                  speech_duration_max=1.5, post_speech_silence=0.3)
    ww._sound_buffer = _SilentBuffer()
    ww._matcher = object()
    t0 = time.time()
    with pytest.raises(TimeoutError):
        ww._detect_word()
    assert 0.9 < time.time() - t0 < 2.5


def test_detect_word_runs_levels_2_and_3_on_scripted_audio(wav, monkeypatch):
    """A scripted buffer (silence 1.0 s, sound 0.9 s, silence) under a fake clock: level 2 is called once
    with the cut the reference would make; level 3 confirms."""
    import easywakeword_b200.wakeword as W

    class Clock:
        def __init__(self): self.k = 0
        def time(self): return self.k * 0.1
        def sleep(self, dt): self.k += 1
    clk = Clock()
    monkeypatch.setattr(W, "time", clk)

    class Buf(_SilentBuffer):
        def is_silent(self): return not (10 < clk.k <= 19)
        def return_last_n_seconds(self, n):
            self.asked = n
            return np.arange(int(n * 16000), dtype=np.float64)

    class Matcher:
        def matches(self, audio, threshold=75.0):
            self.audio = audio
            return True, 99.0

    class Stt:
        def transcribe(self, audio): return "Hello."

    ww = WakeWord("hello", wav, numberofwords=1, timeout=30, speech_duration_min=0.69, speech_duration_max=1.38,
                  transcriber=Stt())
    ww._sound_buffer, ww._matcher = Buf(), Matcher()
    assert ww._detect_word() == "Hello."
    # sound seen at ticks 11..19, end at tick 20, post-silence satisfied at tick 24
    assert clk.k == 24
    from oracle.ewk_oracle import segment_bounds
    n_back, n_drop = segment_bounds(11 * 0.1, 20 * 0.1, 24 * 0.1)
    assert len(ww._matcher.audio) == n_back - n_drop
    # word-count mismatch -> keep listening -> timeout
    clk.k = 0
    ww2 = WakeWord("hello there", wav, numberofwords=2, timeout=5, speech_duration_min=0.69, speech_duration_max=1.38,
                   transcriber=Stt())
    ww2._sound_buffer, ww2._matcher = Buf(), Matcher()
    with pytest.raises(TimeoutError):
        ww2._detect_word()


def test_level3_preprocessing_matches_reference_formula():
    from oracle import ewk_oracle as O
    x = np.random.default_rng(0).standard_normal(5000) * 0.1 + 0.02
    assert np.array_equal(WakeWord.prepare_for_transcription(x), O.prepare_for_level3(x))


def test_stop_is_safe_on_half_built_objects(wav):
    ww = object.__new__(WakeWord)
    ww.stop()
    WakeWord("hello", wav).stop()


# ---- sharding arithmetic
@pytest.mark.parametrize("n,world", [(4096, 1), (4096, 8), (65536, 8), (10, 4), (7, 8)])
def test_shard_ranges_partition_streams(n, world):
    seen = []
    for r in range(world):
        a, b = shard_range(n, world, r)
        assert 0 <= a <= b <= n and b - a <= padded_shard(n, world)
        seen.extend(range(a, b))
        for s in range(a, b):
            assert owner_of(s, n, world) == (r, s - a)
    assert seen == list(range(n))


# ---- level-3 hand-off of the drop-in WakeWord (round-1 advisor finding: a facade that silently never confirms)
def test_missing_level3_backend_is_announced(wav):
    import importlib.util
    if importlib.util.find_spec("whisper") is not None:
        pytest.skip("openai-whisper is installed: a backend exists")
    with pytest.warns(RuntimeWarning, match="no level-3 speech-to-text backend"):
        WakeWord("hello", wav)
    with pytest.warns(RuntimeWarning, match="ignores external_whisper_url, stt_backend"):
        WakeWord("hello", wav, external_whisper_url="http://localhost:1", stt_backend="external", transcriber=object())


def test_initial_prompt_reaches_a_backend_that_accepts_it(wav):
    """wakeword.py:1029 passes initial_prompt=f"Wake word: {textword}" to the level-3 call."""
    seen = {}

    class WithPrompt:
        def transcribe(self, audio, initial_prompt=None):
            seen["prompt"] = initial_prompt
            return "hello"

    class Plain:
        def transcribe(self, audio):
            seen["plain"] = True
            return "hello"

    x = np.zeros(1600, np.float32)
    assert WakeWord("hello", wav, numberofwords=1, transcriber=WithPrompt())._confirm(x) == "hello"
    assert seen["prompt"] == "Wake word: hello"
    assert WakeWord("hello", wav, numberofwords=1, transcriber=Plain())._confirm(x) == "hello" and seen["plain"]
