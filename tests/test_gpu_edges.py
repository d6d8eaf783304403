"""GPU edge cases of the level-1 data plane and the C-ABI: odd callback sizes (generic gate path, chunks that
straddle ticks, a tail of the ring that belongs to no chunk), ragged per-stream pushes, event-queue overflow,
argument errors, empty inputs."""
import numpy as np
import pytest

from easywakeword_b200 import synth

pytestmark = pytest.mark.gpu


def _trace_vs_oracle(word, block, fmt, seconds=24, seed=4321, noise=0.007, push=None):
    from easywakeword_b200 import _lib
    from oracle import ewk_oracle as O
    x, _ = synth.stream(seed, seconds, word, noise_sigma=noise, gain=(2.0, 4.0), inserts_per_10s=(2, 3))
    q = synth.to_int16(x)
    xf = synth.from_int16(q)
    P = dict(speech_duration_min=0.5, speech_duration_max=1.6, timeout=6.0)
    o = O.detect_stream(xf, word, block=block, fast=True, **P)
    ctx = _lib.Context(device=0, n_streams=1, ring_samples=160000, slack_samples=40000,
                       pcm_format=_lib.PCM_I16 if fmt == "i16" else _lib.PCM_F32, max_events=1024)
    ctx.set_template(0, word)
    ctx.set_stream_params(0, frame_size=block, **P)
    data = q if fmt == "i16" else xf
    push = push or block
    silent, thr, events = [], [], []
    pos, ticks = 0, 0
    n = (len(data) // block) * block
    while pos < n:
        m = min(push, n - pos)
        ctx.push(data[pos:pos + m].reshape(1, -1))
        pos += m
        due = pos // 1600 - ticks
        if due > 0:
            tr = ctx.tick(due, trace=True)
            silent.append(tr["silent"][0]); thr.append(tr["thr"][0])
            ticks += due
            events.append(ctx.poll().copy())
    ctx.close()
    silent = np.concatenate(silent); thr = np.concatenate(thr); ev = np.concatenate(events)
    gt = o["trace_tick"]
    _, first = np.unique(gt, return_index=True)
    gt, gs, gthr = gt[first], o["trace_silent"][first], o["trace_thr"][first]
    keep = gt <= ticks
    assert np.array_equal(silent[gt[keep] - 1], gs[keep])
    if fmt == "i16":
        assert np.array_equal(thr[gt[keep] - 1], gthr[keep])
    else:
        np.testing.assert_allclose(thr[gt[keep] - 1], gthr[keep], rtol=1e-12)
    l2 = ev[ev["kind"] == 2]
    want = [e for e in o["events"] if e["tick"] <= ticks]
    assert list(l2["tick"]) == [e["tick"] for e in want] and list(l2["seg_len"]) == [e["seg_len"] for e in want]
    assert list(ev[ev["kind"] == 1]["tick"]) == [t for t in o["timeouts"] if t <= ticks]
    assert len(want) >= 2 and gthr.max() > 0.0051
    return len(want)


@pytest.mark.parametrize("block,fmt", [(333, "i16"), (1000, "f32"), (4800, "i16"), (160, "i16")])
def test_odd_callback_sizes_follow_the_reference(word, block, fmt):
    """frame_size 333: 480 chunks + a 160-sample tail in no chunk; 1000: chunks straddle ticks; 4800: one callback
    spans three ticks; 160: the smallest supported frame (1000 chunks)."""
    _trace_vs_oracle(word, block, fmt)


def test_large_unaligned_pushes(word):
    """pushes of 7 callbacks (2331 samples: neither tick- nor 8-sample aligned) ahead of the ticks"""
    _trace_vs_oracle(word, 333, "i16", push=333 * 7)


def test_event_queue_overflow_is_counted(word):
    from easywakeword_b200 import _lib
    n = 64
    ctx = _lib.Context(device=0, n_streams=n, ring_samples=16000, slack_samples=8000, max_events=16)
    ctx.set_template(0, word)
    ctx.set_stream_params(-1, frame_size=1600, timeout=0.25)       # a timeout event every 3 ticks per stream
    z = np.zeros((n, 1600), np.int16)
    for _ in range(30):
        ctx.push(z)
        ctx.tick(1)
    ev = ctx.poll()
    assert len(ev) == 16 and ctx.dropped > 0 and (ev["kind"] == 1).all()
    ctx.push(z); ctx.tick(1)
    assert len(ctx.poll()) <= 16
    ctx.close()


def test_argument_errors(word):
    from easywakeword_b200 import _lib
    ctx = _lib.Context(device=0, n_streams=2, ring_samples=16000, slack_samples=1600)
    with pytest.raises(ValueError):
        ctx.extract_mfcc(np.zeros(0, np.float32))                  # librosa rejects empty audio too
    with pytest.raises(ValueError, match="must be positive"):
        ctx.set_stream_params(0, pre_speech_silence=0.0)
    with pytest.raises(ValueError, match="speech_duration_min must be <= speech_duration_max"):
        ctx.set_stream_params(0, speech_duration_min=2.0, speech_duration_max=1.0)
    with pytest.raises(ValueError):
        ctx.set_stream_params(5)                                   # no such stream
    with pytest.raises(TypeError):
        ctx.push(np.zeros((2, 160), np.float32))                   # ring is int16
    with pytest.raises(ValueError):
        ctx.set_template(9, word)                                  # slot out of range
    ctx.set_stream_params(-1, frame_size=1600)
    ctx.push(np.zeros((2, 1600), np.int16))
    ctx.tick(1)
    for _ in range(12):
        ctx.push(np.zeros((2, 1600), np.int16)); ctx.tick(1)
    with pytest.raises(_lib.EwkError, match="un-gated"):
        for _ in range(4):
            ctx.push(np.zeros((2, 1600), np.int16))                # running ahead of the ticks beyond the slack
    with pytest.raises(ValueError):
        _lib.Context(device=0, n_streams=1, ring_samples=10)       # ring shorter than one tick
    ctx.close()


@pytest.mark.parametrize("fmt", ["i16", "f32"])
def test_bulk_staged_gate_path(word, fmt):
    """frame_size 1600 with half-tick pushes: K1 cannot produce per-block sums (unaligned pushes), every tick is
    still the aligned single-chunk case, so K2 stages the ticks with cp.async.bulk + mbarrier."""
    _trace_vs_oracle(word, 1600, fmt, push=800)


def test_presummed_and_staged_paths_agree(word):
    """the same stream through K1 block sums (1 s pushes) and through the staged path (0.05 s pushes): identical"""
    a = _trace_vs_oracle(word, 1600, "i16", push=16000, seed=99)
    b = _trace_vs_oracle(word, 1600, "i16", push=800, seed=99)
    assert a == b
