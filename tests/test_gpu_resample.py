"""GPU: K7 resample_kernel (ewk_resample) against oracle/resample_restated.py — the same filter
specification evaluated in float64 — within 2e-6 of full scale; chunked streaming equals the one-shot result
exactly; a 44.1 kHz / 48 kHz copy of the bundled template scores like the 16 kHz original.  Parity with the
reference's soxr is unpinned (no soxr here, no golden in the reference): see the oracle's header."""
import struct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ATOL = 2e-6


@pytest.fixture(scope="module")
def ctx():
    from easywakeword_b200 import _lib
    c = _lib.Context(device=0, n_streams=0, max_templates=2)
    yield c
    c.close()


@pytest.mark.parametrize("sr", [8000, 11025, 22050, 32000, 44100, 48000, 96000])
def test_one_shot_matches_oracle(ctx, sr):
    from oracle import resample_restated as R
    rng = np.random.default_rng(sr)
    n = int(0.4 * sr) + 17
    t = np.arange(n) / sr
    x = (0.3 * np.sin(2 * np.pi * 523.0 * t) + 0.05 * rng.standard_normal(n)).astype(np.float32)
    got = ctx.resample(x, sr)
    ref = R.resample(x, sr)
    assert got.shape == ref.shape == (int(np.ceil(n * 16000 / sr)),)
    assert np.abs(got - ref).max() <= ATOL, float(np.abs(got - ref).max())


def test_int16_rows_and_edges(ctx):
    from oracle import resample_restated as R
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((5, 4410)) * 3000).astype(np.int16)
    x[2] = 0
    got = ctx.resample(x, 44100)
    assert got.shape == (5, 1600) and not got[2].any()
    for r in range(5):
        ref = R.resample(x[r].astype(np.float32) / np.float32(32768.0), 44100)
        assert np.abs(got[r] - ref).max() <= ATOL
    # inputs shorter than the filter, length 1
    for n in (1, 2, 100):
        y = rng.standard_normal(n).astype(np.float32)
        assert np.abs(ctx.resample(y, 48000) - R.resample(y, 48000)).max() <= ATOL


def test_unsupported_rate_fails_loudly(ctx):
    from easywakeword_b200 import _lib
    with pytest.raises(ValueError, match="not supported"):
        ctx.resample(np.zeros(100, np.float32), 44101)          # 16000 phases
    with pytest.raises(ValueError):
        _lib.resample_info(44101)


@pytest.mark.parametrize("sr,dtype", [(44100, np.int16), (48000, np.float32), (8000, np.int16)])
def test_streaming_equals_one_shot(ctx, sr, dtype):
    from easywakeword_b200.resample import StreamResampler
    rng = np.random.default_rng(sr + 1)
    n = 3 * sr // 2 + 5
    x = rng.standard_normal((3, n)) * 0.1
    x = (x * 32767).astype(np.int16) if dtype == np.int16 else x.astype(np.float32)
    whole = ctx.resample(x, sr)
    rs = StreamResampler(sr, 3, ctx=ctx, dtype=dtype)
    parts, pos = [], 0
    for size in [1, 7, 333, 4096, 1000, 12345, 1]:
        parts.append(rs.push(x[:, pos:pos + size]))
        pos += size
    while pos < n:
        parts.append(rs.push(x[:, pos:pos + 5000]))
        pos += 5000
    parts.append(rs.flush())
    got = np.concatenate(parts, axis=1)
    assert got.shape == whole.shape
    assert np.array_equal(got, whole)
    assert rs.hist.shape[1] <= 2 * rs.W + 5000                   # history stays bounded


def _write_wav(path, y, sr, bits=16, ch=1, fmt_tag=1):
    if fmt_tag == 3:
        raw = y.astype("<f4").tobytes()
        bits = 32
    elif bits == 16:
        raw = np.clip(np.rint(y * 32767.0), -32768, 32767).astype("<i2").tobytes()
    elif bits == 24:
        v = np.clip(np.rint(y * 8388607.0), -8388608, 8388607).astype(np.int32)
        raw = b"".join(struct.pack("<i", int(s))[:3] for s in v.reshape(-1))
    elif bits == 8:
        raw = (np.clip(np.rint(y * 127.0), -128, 127) + 128).astype(np.uint8).tobytes()
    hdr = struct.pack("<4sI4s4sIHHIIHH4sI", b"RIFF", 36 + len(raw), b"WAVE", b"fmt ", 16, fmt_tag, ch, sr,
                      sr * ch * bits // 8, ch * bits // 8, bits, b"data", len(raw))
    with open(path, "wb") as f:
        f.write(hdr + raw)


def test_template_at_other_rates_scores_like_the_original(ctx, word, tmp_path):
    """The bundled word, band-limited-interpolated to 44.1 / 48 kHz (float64 oracle), written as WAV, loaded through
    the package loader (decode + device conversion back to 16 kHz) and matched against the 16 kHz original."""
    from easywakeword_b200.resample import load_16k, read_wav
    from oracle import resample_restated as R
    ctx.set_template(0, word)
    for sr, kw in ((44100, dict(bits=16)), (48000, dict(fmt_tag=3)), (48000, dict(bits=24)), (22050, dict(bits=16, ch=2))):
        d = R.design(16000, sr)                                     # 16 k -> sr with the same specification
        up = R.resample(word, 16000, sr)
        assert len(up) == int(np.ceil(len(word) * sr / 16000)) and d["L"] >= 1
        y = np.stack([up, up], axis=1).reshape(-1) if kw.get("ch") == 2 else up
        p = tmp_path / f"w_{sr}_{kw.get('bits', 32)}.wav"
        _write_wav(p, y, sr, **kw)
        raw, sr2 = read_wav(p)
        assert sr2 == sr and raw.shape[0] == len(up)
        back = load_16k(p, ctx=ctx)
        assert abs(len(back) - len(word)) <= 1
        m = min(len(back), len(word))
        mid = slice(400, m - 400)
        assert np.abs(back[mid] - word[mid]).max() < 2e-3           # content above 7.3 kHz is filtered twice
        scores, matched = ctx.similarity_batch(0, back, [0], [len(back)], threshold=75.0)
        assert matched[0] and scores[0] > 97.0, (sr, kw, float(scores[0]))
