"""CPU: the best-effort C statement of the gated path (oracle/cpu_port, bench.py's `cpu_baseline.best_effort_c`) against
the numpy oracle: features within 1e-4, identical level-2 events (tick, segment length, decision), scores within 0.01."""
import numpy as np
import pytest

from easywakeword_b200 import synth


@pytest.fixture(scope="module")
def cpu_port():
    import shutil
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from oracle import cpu_port
    cpu_port.load()
    return cpu_port


def test_features_match_oracle(cpu_port, word, word_i16):
    from oracle import ewk_oracle as O
    rng = np.random.default_rng(11)
    cases = [word_i16, word_i16[:4000], (rng.standard_normal(20000) * 800).astype(np.int16),
             np.concatenate([word_i16[:7000], np.zeros(3000, np.int16), word_i16[7000:]]), word_i16[:159]]
    for q in cases:
        mean, std = cpu_port.features(q)
        rm, rs = O.extract_mfcc(synth.from_int16(q))
        assert np.linalg.norm(mean - rm) <= 1e-4 * np.linalg.norm(rm)
        assert np.linalg.norm(std - rs) <= 1e-4 * max(np.linalg.norm(rs), 1.0)


def test_detect_matches_oracle(cpu_port, word, word_i16):
    from oracle import ewk_oracle as O
    P = dict(similarity_threshold=75.0, speech_duration_min=0.69, speech_duration_max=1.38, timeout=8.0)
    streams = [synth.to_int16(synth.stream(3100 + i, 30.0, word, noise_sigma=0.002, gain=(1.0, 4.0),
                                           distractor_prob=0.3 if i % 2 else 0.0)[0]) for i in range(6)]
    q = np.stack(streams)
    got = cpu_port.detect_batch(q, word_i16, threads=3, **P)
    n = 0
    for i, ev in enumerate(got):
        o = O.detect_stream(synth.from_int16(q[i]), word, block=1600, fast=True, **P)
        assert list(ev["tick"]) == [e["tick"] for e in o["events"]], i
        assert list(ev["seg_len"]) == [e["seg_len"] for e in o["events"]], i
        for a, b in zip(ev, o["events"]):
            assert abs(float(a["score"]) - b["score"]) <= 0.01
            if abs(b["score"] - 75.0) > 0.01:
                assert bool(a["matched"]) == b["matched"]
            n += 1
    assert n >= 8
