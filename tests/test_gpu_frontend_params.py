"""GPU parity for the front-end parameters ewk_config exposes (ABI 2): `preemphasis` and `n_mfcc`.

The reference hard-codes both (no pre-emphasis, n_mfcc=20: wakeword.py:561-563) and lists exposing them as wanted
(LEARNINGS.md:87; BASELINE.json north_star names pre-emphasis).  At the reference's values every result must be
bit-identical to a context that never heard of the parameters; at other values the device path is compared with the
oracle (librosa.effects.preemphasis / librosa.feature.mfcc(n_mfcc=...) restated) — MFCC <= 1e-4 per-frame relative L2,
scores within 0.01."""
import numpy as np
import pytest

from easywakeword_b200 import synth
from helpers import mfcc_rel_l2

pytestmark = pytest.mark.gpu
MFCC_RTOL = 1e-4
SCORE_ATOL = 0.01


def _signals(word):
    rng = np.random.default_rng(11)
    x, _ = synth.stream(77, 3.0, word, gain=(2.0, 3.0), inserts_per_10s=(1, 1))
    return {
        "word": word,
        "stream": synth.from_int16(synth.to_int16(x)),
        "noise_odd_len": (rng.standard_normal(12799) * 0.01).astype(np.float32),
        "sine440": synth.sine(440.0, 1.0),
        "short": (rng.standard_normal(700) * 0.02).astype(np.float32),
        "gap": np.concatenate([word, np.zeros(4000, np.float32), 0.5 * word]).astype(np.float32),
    }


def test_default_parameters_are_bit_identical(word):
    from easywakeword_b200 import _lib
    a = _lib.Context(device=0, n_streams=0, max_templates=1)
    b = _lib.Context(device=0, n_streams=0, max_templates=1, preemphasis=0.0, n_mfcc=20)
    try:
        for name, x in _signals(word).items():
            ma, sa, fa = a.extract_mfcc(x, want_frames=True)
            mb, sb, fb = b.extract_mfcc(x, want_frames=True)
            assert np.array_equal(fa, fb) and np.array_equal(ma, mb) and np.array_equal(sa, sb), name
    finally:
        a.close()
        b.close()


@pytest.mark.parametrize("coef", [0.97, 0.5])
def test_preemphasis_matches_oracle(word, coef):
    from easywakeword_b200 import _lib
    from oracle import ewk_oracle as O
    ctx = _lib.Context(device=0, n_streams=0, max_templates=2, preemphasis=coef)
    try:
        ctx.set_template(0, word)
        m = O.WordMatcherOracle(preemphasis=coef)
        m.set_reference(word)
        worst_f = worst_s = 0.0
        sig = _signals(word)
        for name, x in sig.items():
            mean, std, frames = ctx.extract_mfcc(x, want_frames=True)
            ref = O.mfcc_frames(x, preemphasis=coef)
            assert frames.shape == (ref.shape[1], 20), name
            err = mfcc_rel_l2(frames.T, ref)
            worst_f = max(worst_f, float(err.max()))
            assert err.max() <= MFCC_RTOL, (name, float(err.max()))
            rm, rs = O.extract_mfcc(x, preemphasis=coef)
            assert np.linalg.norm(mean - rm) <= 1e-4 * np.linalg.norm(rm), name
        # scores of segments cut at odd / even offsets out of one buffer (pair alignment of the loader)
        x = sig["stream"]
        offs = np.array([0, 1, 4001, 16000, 20480])
        lens = np.array([len(x), 17601, 15503, 12799, 9000])
        scores, _ = ctx.similarity_batch(0, x, offs, lens, threshold=75.0)
        q16 = synth.to_int16(x)
        scores16, _ = ctx.similarity_batch(0, q16, offs, lens, threshold=75.0)
        for i, (o, l) in enumerate(zip(offs, lens)):
            ref = float(m.calculate_similarity(x[o:o + l]))
            worst_s = max(worst_s, abs(float(scores[i]) - ref))
            assert abs(float(scores[i]) - ref) <= SCORE_ATOL, (i, float(scores[i]), ref)
            assert scores16[i] == scores[i]            # the stream is int16-representable: both loaders see the same samples
        print(f"[a={coef}] worst per-frame rel L2 {worst_f:.2e}, worst |score - oracle| {worst_s:.2e}")
    finally:
        ctx.close()


def test_preemphasis_detect_and_dense_match_oracle(word):
    """The gated path (K1/K2/K3 from the rings) and the dense path (K4, window-local filter state) with a = 0.97."""
    from easywakeword_b200.bank import WakeWordBank
    from oracle import ewk_oracle as O
    coef = 0.97
    P = dict(speech_duration_min=0.69, speech_duration_max=1.38, timeout=30.0)
    xs = [synth.from_int16(synth.to_int16(synth.stream(5200 + i, 24.0, word, gain=(1.5, 4.0), zero_gaps=i)[0])) for i in range(2)]
    q = np.stack([synth.to_int16(x) for x in xs])
    bank = WakeWordBank(2, [word], frame_size=1600, preemphasis=coef, **P)
    try:
        evs = []
        for p in range(0, q.shape[1], 16000):
            bank.step(np.ascontiguousarray(q[:, p:p + 16000]))
            evs.append(bank.poll().copy())
        ev = np.concatenate(evs)
        n = 0
        m = O.WordMatcherOracle(preemphasis=coef)
        m.set_reference(word)
        for i in range(2):
            o = O.detect_stream(xs[i], word, block=1600, fast=True, matcher=m, **P)
            mine = ev[(ev["stream"] == i) & (ev["kind"] == 2)]
            assert list(mine["tick"]) == [e["tick"] for e in o["events"]]
            for a, b in zip(mine, o["events"]):
                assert abs(float(a["score"]) - b["score"]) <= SCORE_ATOL
                n += 1
        assert n >= 3
        # dense: the last 8 s are still in the 10 s rings
        hop_end = q.shape[1] // 160
        sc = bank.dense_scores(hop_end - 500, 500)
        hops = np.arange(hop_end - 500, hop_end, 13)
        for i in range(2):
            ref = O.dense_scores(xs[i], [word], hops, preemphasis=coef)[:, 0]
            got = sc[i, hops - (hop_end - 500), 0]
            assert np.abs(got - ref).max() <= SCORE_ATOL, (i, float(np.abs(got - ref).max()))
    finally:
        bank.close()


@pytest.mark.parametrize("n_mfcc", [13, 1])
def test_n_mfcc_matches_oracle(word, n_mfcc):
    from easywakeword_b200 import _lib
    from oracle import ewk_oracle as O
    ctx = _lib.Context(device=0, n_streams=0, max_templates=2, n_mfcc=n_mfcc)
    try:
        ctx.set_template(0, word)
        m = O.WordMatcherOracle(n_mfcc=n_mfcc)
        m.set_reference(word)
        for name, x in _signals(word).items():
            mean, std = ctx.extract_mfcc(x)
            rm, rs = O.extract_mfcc(x, n_mfcc=n_mfcc)
            assert np.all(mean[n_mfcc:] == 0) and np.all(std[n_mfcc:] == 0), name
            assert np.linalg.norm(mean[:n_mfcc] - rm) <= 1e-4 * np.linalg.norm(rm), name
            assert np.abs(std[:n_mfcc] - rs).max() <= 1e-4 * max(1.0, float(np.abs(rs).max())), name
            sc, _ = ctx.similarity_batch(0, x, [0], [len(x)])
            with np.errstate(all="ignore"):
                ref = float(m.calculate_similarity(x))
            if np.isnan(ref):
                assert np.isnan(sc[0]), name
            else:
                assert abs(float(sc[0]) - ref) <= SCORE_ATOL, (name, float(sc[0]), ref)
    finally:
        ctx.close()


def test_bad_front_end_parameters_are_refused():
    from easywakeword_b200 import _lib
    for kw in (dict(preemphasis=1.0), dict(preemphasis=-0.1), dict(n_mfcc=21), dict(n_mfcc=-1)):
        with pytest.raises(ValueError):
            _lib.Context(device=0, n_streams=0, max_templates=1, **kw)
