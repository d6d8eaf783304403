"""GPU parity: K3 (segment MFCC + match) through the C-ABI vs the reference-generated goldens and
the oracle.  Tolerances are BASELINE.json's: MFCC <= 1e-4 relative (per-frame L2),
scores within 0.01 on the 0-100 scale, decisions identical."""
import numpy as np
import pytest

from helpers import mfcc_rel_l2

pytestmark = pytest.mark.gpu

MFCC_RTOL = 1e-4     # per-frame relative L2
SCORE_ATOL = 0.01


@pytest.fixture(scope="module")
def ctx():
    from easywakeword_b200 import _lib
    c = _lib.Context(device=0, n_streams=0, max_templates=4)
    yield c
    c.close()


def test_mfcc_frames_match_reference_goldens(ctx, golden_matcher):
    g = golden_matcher
    worst = 0.0
    for name in g["names"]:
        a = g[f"in_{name}"]
        mean, std, frames = ctx.extract_mfcc(a, want_frames=True)
        ref = g[f"mfcc_{name}"]                       # [20, F] from the reference's code path
        assert frames.shape == (ref.shape[1], 20), name
        err = mfcc_rel_l2(frames.T, ref)
        worst = max(worst, float(err.max()))
        assert err.max() <= MFCC_RTOL, (name, float(err.max()))
        atol = 1e-4 * float(np.abs(ref).max())
        assert np.abs(frames.T - ref).max() <= atol, name
        scale = float(np.linalg.norm(g[f"mean_{name}"]))
        assert np.linalg.norm(mean - g[f"mean_{name}"]) <= 1e-4 * scale, name
        assert np.abs(std - g[f"std_{name}"]).max() <= 1e-4 * max(1.0, float(np.abs(g[f"std_{name}"]).max())), name
    print("worst per-frame rel L2:", worst)


def test_scores_and_decisions_match_reference_goldens(ctx, golden_matcher, word):
    g = golden_matcher
    tpl = {"word": word, "sine440": g["in_sine440"], "speech_like": g["in_speech_like"]}
    names = [str(n) for n in g["names"]]
    pcm = np.concatenate([g[f"in_{n}"] for n in names])
    lens = np.array([len(g[f"in_{n}"]) for n in names])
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    worst = 0.0
    for slot, (k, a) in enumerate(tpl.items()):
        ctx.set_template(slot, a)
        scores, matched = ctx.similarity_batch(slot, pcm, offs, lens, threshold=75.0)
        for i, n in enumerate(names):
            ref = float(g[f"score_{k}_{n}"])
            if np.isnan(ref):
                assert np.isnan(scores[i]) and not matched[i], (k, n)
                continue
            worst = max(worst, abs(float(scores[i]) - ref))
            assert abs(float(scores[i]) - ref) <= SCORE_ATOL, (k, n, float(scores[i]), ref)
            if abs(ref - 75.0) > SCORE_ATOL:
                assert bool(matched[i]) == bool(g[f"match_{k}_{n}"]), (k, n)
    print("worst |score - ref|:", worst)


@pytest.mark.parametrize("sig", ["sine440", "speech_like", "word"])
def test_self_similarity_is_exactly_100(ctx, golden_matcher, sig):
    """tests/test_wakeword_simulated.py:107-118, 194-205; tests/test_cross_platform.py:94-109."""
    a = golden_matcher[f"in_{sig}"]
    ctx.set_template(0, a)
    s1, m1 = ctx.similarity_batch(0, a, [0], [len(a)])
    s2, m2 = ctx.similarity_batch(0, a, [0], [len(a)])
    assert s1[0] == 100.0 and m1[0] and s1[0] == s2[0]


def test_int16_and_f32_inputs_agree(ctx, golden_matcher, word):
    from easywakeword_b200 import synth
    ctx.set_template(0, word)
    x = golden_matcher["in_stream_3s_i16"]
    q = synth.to_int16(x)
    assert np.array_equal(synth.from_int16(q), x)
    sf, _ = ctx.similarity_batch(0, x, [0, 1000], [len(x), 20000])
    si, _ = ctx.similarity_batch(0, q, [0, 1000], [len(q), 20000])
    assert np.array_equal(sf, si)


def test_long_input_spills_to_workspace(ctx, word):
    """set_reference on audio longer than the 3.0 s cap (global workspace path) vs the oracle."""
    from oracle import ewk_oracle as O
    rng = np.random.default_rng(11)
    x = (rng.standard_normal(5 * 16000) * 0.01).astype(np.float32)
    x[20000:20000 + len(word)] += 2 * word
    mean, std, frames = ctx.extract_mfcc(x, want_frames=True)
    ref = O.mfcc_frames(x)
    assert frames.shape == (501, 20)
    assert mfcc_rel_l2(frames.T, ref).max() <= MFCC_RTOL


def test_no_template_raises(ctx):
    ctx.clear_template(1)
    with pytest.raises(ValueError, match="No reference word set"):
        ctx.similarity_batch(1, np.zeros(16000, np.float32), [0], [16000])


def test_batch_of_many_segments_vs_oracle(ctx, word):
    from oracle import ewk_oracle as O
    from easywakeword_b200 import synth
    x, _ = synth.stream(901, 30.0, word, gain=(1.0, 4.0))
    q = synth.to_int16(x)
    xf = synth.from_int16(q)
    rng = np.random.default_rng(3)
    offs = rng.integers(0, len(q) - 48000, size=64)
    lens = rng.integers(160, 48001, size=64)
    ctx.set_template(0, word)
    scores, matched, feats = ctx.similarity_batch(0, q, offs, lens, threshold=90.0, want_features=True)
    m = O.WordMatcherOracle()
    m.set_reference(word)
    for i in range(0, 64, 5):
        seg = xf[offs[i]:offs[i] + lens[i]]
        ok, sim = m.matches(seg, threshold=90.0)
        assert abs(float(scores[i]) - float(sim)) <= SCORE_ATOL
        if abs(float(sim) - 90.0) > SCORE_ATOL:
            assert bool(matched[i]) == bool(ok)
        mean, std = O.extract_mfcc(seg)
        assert np.linalg.norm(feats[i, :20] - mean) <= 1e-4 * np.linalg.norm(mean)


def test_floor_from_stored_logmel_is_bit_identical_to_recompute(ctx, golden_matcher, word, monkeypatch):
    """K3 floors a frame by re-running the DCT on its stored log-mel row; with EWK_SEG_LM=0 it recomputes the
    frame from the PCM.  Both must give the same bits (frames, features, scores, queue path included)."""
    from easywakeword_b200 import _lib
    from easywakeword_b200.bank import WakeWordBank
    from easywakeword_b200.synth import stream_batch
    g = golden_matcher
    names = [str(n) for n in g["names"]]
    monkeypatch.setenv("EWK_SEG_LM", "0")
    plain = _lib.Context(device=0, n_streams=0, max_templates=1)
    monkeypatch.delenv("EWK_SEG_LM")
    try:
        plain.set_template(0, word)
        ctx.set_template(0, word)
        for n in names:
            a = g[f"in_{n}"]
            m0, s0, f0 = plain.extract_mfcc(a, want_frames=True)
            m1, s1, f1 = ctx.extract_mfcc(a, want_frames=True)
            assert np.array_equal(f0, f1) and np.array_equal(m0, m1) and np.array_equal(s0, s1), n
    finally:
        plain.close()
    # queue form (K2 -> K3 on the device rings)
    pcm16 = stream_batch(900, 8, 20.0, word, zero_gaps=2)
    events = []
    for flag in ("0", "1"):
        monkeypatch.setenv("EWK_SEG_LM", flag)
        bank = WakeWordBank(8, [word], device=0)
        try:
            evs = []
            for b in range(0, pcm16.shape[1], 16000):
                bank.step(np.ascontiguousarray(pcm16[:, b:b + 16000]))
                evs.append(bank.poll())
            events.append(np.concatenate(evs))
        finally:
            bank.close()
    monkeypatch.delenv("EWK_SEG_LM")
    assert len(events[0]) == len(events[1]) and len(events[0]) > 0
    for f in events[0].dtype.names:
        assert np.array_equal(events[0][f], events[1][f], equal_nan=events[0][f].dtype.kind == "f"), f
