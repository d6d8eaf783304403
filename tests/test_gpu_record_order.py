"""The 8-byte per-stream result record (what multi-GPU runs gather) must carry the stream's LATEST level-2 evaluation
even when one ewk_tick call queues several candidates of one stream and different K3 CTAs score them in arbitrary
order (round-1 advisor finding: read-modify-write race on results[]).  pre / min / post of 0.1 s force an event every
~0.7 s, so one 32-tick call queues 4-5 events per stream."""
import numpy as np
import pytest

from easywakeword_b200 import synth

pytestmark = pytest.mark.gpu


def burst_stream(seed, seconds):
    """quiet / loud / quiet pattern on the 0.1 s grid: 2 quiet ticks, 2 loud ticks (a different tone every burst, so
    consecutive events of a stream score differently), 3 quiet ticks."""
    rng = np.random.default_rng(seed)
    n = int(seconds * 16000)
    x = (rng.standard_normal(n) * 0.0004).astype(np.float32)
    t = np.arange(3200) / 16000.0
    k = 0
    pos = int(rng.integers(0, 7)) * 1600
    while pos + 7 * 1600 <= n:
        f = 200.0 + 137.0 * ((seed * 7 + k * 13) % 23)
        x[pos + 3200:pos + 6400] += (0.05 * np.sin(2 * np.pi * f * t) * np.hanning(3200)).astype(np.float32) + \
            (rng.standard_normal(3200) * 0.01).astype(np.float32)
        pos += 7 * 1600
        k += 1
    return x


def test_result_record_is_the_latest_event(word):
    from easywakeword_b200 import _lib
    n, seconds = 384, 22.4
    P = dict(frame_size=1600, pre_speech_silence=0.1, speech_duration_min=0.1, speech_duration_max=0.35,
             post_speech_silence=0.1, timeout=0.0, similarity_threshold=60.0)
    xs = np.stack([synth.to_int16(burst_stream(900 + s, seconds)) for s in range(n)])
    ctx = _lib.Context(device=0, n_streams=n, ring_samples=16000, slack_samples=3200 * 17, pcm_format=_lib.PCM_I16,
                       max_templates=1, max_events=1 << 16)
    try:
        ctx.set_template(0, word)
        ctx.set_stream_params(-1, **P)
        checked = multi = 0
        last_score = np.full(n, np.nan, np.float32)
        last_match = np.zeros(n, bool)
        ever = np.zeros(n, bool)
        all_ev = []
        for rep, p in enumerate(range(0, xs.shape[1] - 32 * 1600 + 1, 32 * 1600)):
            ctx.push(np.ascontiguousarray(xs[:, p:p + 32 * 1600]))
            ctx.tick(32)                       # two gate launches, ONE K3 launch over all their candidates
            res = ctx.results()
            ev = ctx.poll()
            ev = ev[ev["kind"] == 2]
            ev = ev[np.lexsort((ev["tick"], ev["stream"]))]
            all_ev.append(ev)
            per_stream = np.bincount(ev["stream"], minlength=n)
            multi += int((per_stream >= 2).sum())
            idx = np.cumsum(per_stream) - 1    # the last event (largest tick) of every stream that had one in this call
            has = per_stream > 0
            last_score[has] = ev["score"][idx[has]]
            last_match[has] = ev["matched"][idx[has]].astype(bool)
            ever |= has
            assert np.array_equal(res["score"][ever], last_score[ever], equal_nan=True), rep
            assert np.array_equal((res["flags"][ever] & 1).astype(bool), last_match[ever]), rep
            assert np.array_equal(((res["flags"] >> 4) & 1).astype(bool), has), rep
            checked += int(has.sum())
        assert multi >= 1000, multi            # stream-calls with two or more events scored by one K3 launch
        print(f"record == latest event in {checked} stream-calls, {multi} of them with >= 2 events in one launch")
        # the events themselves are the oracle's (0.3 s segments, 0.1 s timing, 1 s ring): spot check
        from oracle import ewk_oracle as O
        ev = np.concatenate(all_ev)
        po = {k: v for k, v in P.items() if k != "frame_size"}
        po["timeout"] = 1e12
        for s in (0, 191, n - 1):
            o = O.detect_stream(synth.from_int16(xs[s]), word, block=1600, buffer_seconds=1, fast=True, **po)
            mine = ev[ev["stream"] == s]
            ref = [e for e in o["events"] if e["tick"] <= int(mine["tick"].max())]
            assert list(mine["tick"]) == [e["tick"] for e in ref] and len(ref) >= 20
            assert list(mine["seg_len"]) == [e["seg_len"] for e in ref]
            assert np.abs(mine["score"] - np.array([e["score"] for e in ref])).max() <= 0.01
    finally:
        ctx.close()


def test_candidate_overflow_keeps_the_queue_consistent(word):
    """More candidates than event slots: the ones that found a slot are all scored (none left pending, K3's length-class
    lists hold exactly them), the rest are counted as dropped, and the next call works on a clean queue."""
    from easywakeword_b200 import _lib
    n, cap = 256, 64
    P = dict(frame_size=1600, pre_speech_silence=0.1, speech_duration_min=0.1, speech_duration_max=0.35,
             post_speech_silence=0.1, timeout=0.0, similarity_threshold=60.0)
    xs = np.stack([synth.to_int16(burst_stream(1900 + s, 9.6)) for s in range(n)])
    ctx = _lib.Context(device=0, n_streams=n, ring_samples=16000, slack_samples=3200 * 17, pcm_format=_lib.PCM_I16,
                       max_templates=1, max_events=cap)
    try:
        ctx.set_template(0, word)
        ctx.set_stream_params(-1, **P)
        seen_drop = 0
        for p in range(0, xs.shape[1] - 32 * 1600 + 1, 32 * 1600):
            ctx.push(np.ascontiguousarray(xs[:, p:p + 32 * 1600]))
            ctx.tick(32)
            ev = ctx.poll()
            assert len(ev) <= cap
            assert (ev["kind"] == 2).all(), np.unique(ev["kind"])          # nothing left pending, no timeouts configured
            assert np.isfinite(ev["score"]).all() and (ev["seg_len"] > 0).all()
            seen_drop += ctx.dropped
        assert seen_drop > 0
        # with room for everything the same audio gives the same scores for the events that fit before
        ctx2 = _lib.Context(device=0, n_streams=n, ring_samples=16000, slack_samples=3200 * 17, pcm_format=_lib.PCM_I16,
                            max_templates=1, max_events=1 << 15)
        try:
            ctx2.set_template(0, word)
            ctx2.set_stream_params(-1, **P)
            ctx.close()
            ctx = _lib.Context(device=0, n_streams=n, ring_samples=16000, slack_samples=3200 * 17, pcm_format=_lib.PCM_I16,
                               max_templates=1, max_events=cap)
            ctx.set_template(0, word)
            ctx.set_stream_params(-1, **P)
            blk = np.ascontiguousarray(xs[:, :32 * 1600])
            ctx.push(blk); ctx.tick(32)
            ctx2.push(blk); ctx2.tick(32)
            few, full = ctx.poll(), ctx2.poll()
            assert len(full) > cap and len(few) == cap
            key = {(int(e["stream"]), int(e["tick"])): float(e["score"]) for e in full}
            for e in few:
                assert key[(int(e["stream"]), int(e["tick"]))] == float(e["score"])
        finally:
            ctx2.close()
    finally:
        ctx.close()
