"""GPU parity: K4 dense per-hop scoring through the C-ABI vs goldens produced by the reference's
WordMatcher.calculate_similarity on every dense window (strided), and vs the oracle on extra hops.
Tolerance: 0.01 on the 0-100 scale (BASELINE.json)."""
import numpy as np
import pytest

from easywakeword_b200 import synth
from helpers import sha

pytestmark = pytest.mark.gpu
SCORE_ATOL = 0.01


def make_ctx(n_streams, fmt, templates):
    from easywakeword_b200 import _lib
    ctx = _lib.Context(device=0, n_streams=n_streams, ring_samples=200000, slack_samples=16000,
                       pcm_format=_lib.PCM_I16 if fmt == "i16" else _lib.PCM_F32, max_templates=4)
    for i, t in enumerate(templates):
        ctx.set_template(i, t)
    ctx.set_stream_params(-1, frame_size=1600, template_count=len(templates))
    return ctx


@pytest.mark.parametrize("fmt", ["i16", "f32"])
def test_dense_scores_match_reference_goldens(fmt, golden_dense, word):
    g = golden_dense
    tpls = [word, g["tpl2"]]
    xs = []
    for si in range(2):
        x, _ = synth.stream(int(g[f"s{si}_seed"]), 12.0, word, gain=(1.0, 4.0), zero_gaps=int(g[f"s{si}_zero_gaps"]))
        x = synth.from_int16(synth.to_int16(x))
        assert sha(x) == str(g[f"s{si}_sha"])
        xs.append(x)
    ctx = make_ctx(2, fmt, tpls)
    pcm = np.stack([synth.to_int16(x) if fmt == "i16" else x for x in xs])
    ctx.push(pcm)
    sc = ctx.dense_scores(200, 1000, 0, 2)          # hops 200 .. 1199
    assert sc.shape == (2, 1000, 2)
    worst = 0.0
    for si in range(2):
        hops = g[f"s{si}_hops"]
        ref = g[f"s{si}_scores"]
        got = sc[si, hops - 200, :].astype(np.float64)
        assert not np.isnan(got).any()
        worst = max(worst, float(np.abs(got - ref).max()))
        assert np.abs(got - ref).max() <= SCORE_ATOL, (si, float(np.abs(got - ref).max()))
    print(f"[{fmt}] worst |dense score - reference| = {worst:.2e} over {2 * len(hops) * 2} windows")
    # chunking of the request must not matter (sub-chunk boundaries, history recomputation)
    a = ctx.dense_scores(200, 37, 0, 2)
    b = ctx.dense_scores(237, 100, 0, 2)
    assert np.array_equal(a, sc[:, :37]) and np.array_equal(b, sc[:, 37:137])
    one = ctx.dense_scores(300, 64, 1, 1)
    assert np.array_equal(one[:, :, 0], sc[:, 100:164, 1])
    ctx.close()


def test_dense_vs_oracle_extra_hops_and_early_windows(word):
    """Hops the goldens do not cover (incl. digital-silence stretches: the floored-frame path) and windows that
    would start before the stream (NaN)."""
    from oracle import ewk_oracle as O
    x, _ = synth.stream(3100, 6.0, word, gain=(2.0, 4.0), zero_gaps=5, noise_sigma=0.0005)
    x = synth.from_int16(synth.to_int16(x))
    ctx = make_ctx(1, "i16", [word])
    ctx.push(synth.to_int16(x).reshape(1, -1))
    sc = ctx.dense_scores(0, 600, 0, 1)[0, :, 0]
    n_back, _ = O.dense_window(len(word))
    assert np.isnan(sc[:n_back]).all() and not np.isnan(sc[n_back:]).any()
    hops = np.arange(n_back, 600, 11)
    ref = O.dense_scores(x, [word], hops)[:, 0]
    assert np.abs(sc[hops] - ref).max() <= SCORE_ATOL, float(np.abs(sc[hops] - ref).max())
    ctx.close()


def test_dense_self_window_is_exactly_100(word):
    """A window that is exactly the template scores 100.0 (the reference's self-match tests)."""
    x = np.concatenate([np.zeros(160 * 50, np.float32), word, np.zeros(160 * 20 + 97, np.float32)])
    ctx = make_ctx(1, "f32", [word])
    ctx.push(x.reshape(1, -1))
    from oracle import ewk_oracle as O
    n_back, _ = O.dense_window(len(word))
    sc = ctx.dense_scores(50 + n_back, 1, 0, 1)
    assert sc[0, 0, 0] == 100.0
    ctx.close()


def test_dense_argument_errors(word):
    ctx = make_ctx(1, "i16", [word])
    ctx.push(np.zeros((1, 16000), np.int16))
    with pytest.raises(Exception):
        ctx.dense_scores(0, 200, 0, 1)         # hop 199 not pushed yet
    ctx.clear_template(0)
    with pytest.raises(ValueError, match="No reference word set"):
        ctx.dense_scores(0, 50, 0, 1)
    ctx.close()


def test_dense_sweep_multi_template_two_word_phrase(word):
    """BASELINE config 5 in small: several templates of different lengths (incl. a 2-word phrase of 2.1 s) scored at
    every hop of a chunked stream; sampled hops against the oracle."""
    from easywakeword_b200.bank import WakeWordBank
    from oracle import ewk_oracle as O
    phrase = np.concatenate([word, np.zeros(2400, np.float32), word[::-1]]).astype(np.float32)   # word + 0.15 s + reversed word
    short = synth.synthetic_word(seed=9, duration=0.5)
    tpls = [word, phrase, short]
    xs = []
    for i in range(3):
        x, _ = synth.stream(8100 + i, 9.0, phrase if i == 1 else word, gain=(1.5, 3.0), inserts_per_10s=(2, 2),
                            zero_gaps=2 if i == 2 else 0)
        xs.append(synth.from_int16(synth.to_int16(x)))
    q = np.stack([synth.to_int16(x) for x in xs])
    bank = WakeWordBank(3, tpls, buffer_seconds=6, max_push_seconds=2.0)
    got = {}
    for hop0, sc in bank.dense_sweep(np.ascontiguousarray(q[:, p:p + 24000]) for p in range(0, q.shape[1], 24000)):
        assert sc.shape == (3, 150, 3)
        for h in range(sc.shape[1]):
            got[hop0 + h] = sc[:, h, :]
    bank.close()
    assert sorted(got) == list(range(1, 901))
    rng = np.random.default_rng(2)
    worst = 0.0
    for h in sorted(rng.choice(np.arange(215, 900), size=24, replace=False)):
        for s in range(3):
            ref = O.dense_scores(xs[s], tpls, [int(h)])[0]
            worst = max(worst, float(np.nanmax(np.abs(got[int(h)][s] - ref))))
            assert np.allclose(got[int(h)][s], ref, atol=SCORE_ATOL, equal_nan=True), (h, s, got[int(h)][s], ref)
    # before a template's first complete window the score is NaN, from then on a number
    for k, t in enumerate(tpls):
        nb, _ = O.dense_window(len(t))
        assert np.isnan(got[nb - 1][0, k]) and not np.isnan(got[nb][0, k])
    print(f"dense sweep, 3 templates (0.5 s / 0.97 s / 2.1 s): worst |score - oracle| = {worst:.2e}")


def test_dense_config2_64_streams_strided_vs_oracle(word):
    """BASELINE configs[1] / SURVEY §8(d) config 2: 64 concurrent streams (seeds 1000 + s, sigma = 0.002 background, 1-3
    insertions per 10 s at arbitrary sample offsets, gain U(1, 4); 4 of them with exact-zero gaps so that the top_db
    floor path runs), 10 s rolling buffers, every hop scored against 2 templates; every 7th hop of every stream is
    compared with oracle.dense_scores (= WordMatcher.calculate_similarity on that window)."""
    from oracle import ewk_oracle as O
    n, seconds = 64, 10.0
    short = synth.synthetic_word(seed=9, duration=0.5)
    tpls = [word, short]
    xs = []
    for s in range(n):
        x, _ = synth.stream(1000 + s, seconds, word, noise_sigma=0.002, gain=(1.0, 4.0), zero_gaps=3 if s % 16 == 5 else 0)
        xs.append(synth.from_int16(synth.to_int16(x)))
    q = np.stack([synth.to_int16(x) for x in xs])
    ctx = make_ctx(n, "i16", tpls)
    ctx.push(q)
    sc = ctx.dense_scores(0, 1000, 0, 2)
    # the same request in streaming form (100 hops per call, the carried rows path) must give the same bits
    ctx2 = make_ctx(n, "i16", tpls)
    parts = []
    for p in range(0, q.shape[1], 16000):
        ctx2.push(np.ascontiguousarray(q[:, p:p + 16000]))
        parts.append(ctx2.dense_scores(p // 160 + 1, 100, 0, 2))
    ctx2.close()
    stream_form = np.concatenate(parts, axis=1)
    assert np.array_equal(stream_form[:, :-1], sc[:, 1:], equal_nan=True)
    hops = np.arange(3, 1000, 7)
    worst, n_cmp, n_nan = 0.0, 0, 0
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=8) as ex:                      # numpy's FFT / einsum release the GIL
        refs = list(ex.map(lambda x: O.dense_scores(x, tpls, hops), xs))
    for s in range(n):
        ref = refs[s]
        got = sc[s, hops, :]
        assert np.array_equal(np.isnan(got), np.isnan(ref)), s
        ok = ~np.isnan(ref)
        n_nan += int((~ok).sum())
        d = np.abs(got[ok].astype(np.float64) - ref[ok])
        worst = max(worst, float(d.max()))
        n_cmp += int(ok.sum())
        assert d.max() <= SCORE_ATOL, (s, float(d.max()))
    ctx.close()
    print(f"config 2: {n_cmp} windows of 64 streams x 2 templates within {worst:.2e} of the oracle ({n_nan} before the first window)")


def test_dense_five_templates_of_assorted_lengths(word):
    """More templates than fit the two-CTAs-per-SM geometry (the kernel then runs 32-warp CTAs): 0.04 s-granular lengths
    from the 640-sample minimum to the 3.0 s cap, scored in two calls; sampled hops against the oracle."""
    from oracle import ewk_oracle as O
    rng = np.random.default_rng(5)
    long3 = np.concatenate([word, np.zeros(1600, np.float32), word[::-1], np.zeros(900, np.float32), 0.5 * word])[:48000].astype(np.float32)
    tpls = [word, synth.synthetic_word(seed=3, duration=0.5)[:640],  # the shortest supported template
            synth.sine(440.0, 1.0, 0.3), long3, (rng.standard_normal(7777) * 0.05).astype(np.float32)]
    xs = []
    for i in range(2):
        x, _ = synth.stream(8800 + i, 12.0, word, gain=(1.5, 3.0), inserts_per_10s=(2, 3), zero_gaps=i)
        xs.append(synth.from_int16(synth.to_int16(x)))
    q = np.stack([synth.to_int16(x) for x in xs])
    from easywakeword_b200 import _lib
    ctx = _lib.Context(device=0, n_streams=2, ring_samples=200000, slack_samples=16000, pcm_format=_lib.PCM_I16, max_templates=8)
    for i, t in enumerate(tpls):
        ctx.set_template(i, t)
    ctx.set_stream_params(-1, frame_size=1600, template_count=len(tpls))
    ctx.push(q)
    a = ctx.dense_scores(0, 700, 0, 5)
    b = ctx.dense_scores(700, 500, 0, 5)
    sc = np.concatenate([a, b], axis=1)
    one = ctx.dense_scores(350, 300, 3, 1)                               # a single long template on its own
    assert np.array_equal(one[:, :, 0], sc[:, 350:650, 3], equal_nan=True)
    ctx.close()
    hops = np.sort(rng.choice(np.arange(5, 1200), size=40, replace=False))
    worst = 0.0
    for s in range(2):
        ref = O.dense_scores(xs[s], tpls, hops)
        got = sc[s, hops, :]
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        ok = ~np.isnan(ref)
        worst = max(worst, float(np.abs(got[ok] - ref[ok]).max()))
        assert np.abs(got[ok] - ref[ok]).max() <= SCORE_ATOL, (s, worst)
    print(f"5 templates (640 .. 48000 samples): worst |score - oracle| = {worst:.2e}")
