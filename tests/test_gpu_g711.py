"""GPU: G.711 ingest (ewk_push_g711) — 8-bit mu-law / A-law codes expanded on the device and pushed like PCM16.  Pushing
the codes must give the same ring contents, events and result records, bit for bit, as pushing the expanded samples
(expansion by CPython's audioop where it exists, else by the package's host tables, which the CPU suite pins against
audioop)."""
import warnings

import numpy as np
import pytest

from easywakeword_b200 import synth
from easywakeword_b200.resample import ALAW_TABLE, ULAW_TABLE

pytestmark = pytest.mark.gpu


def _encode(q, law):
    """16-bit samples -> nearest G.711 code (table search; ties and the two zero codes do not matter here)."""
    table = (ULAW_TABLE if law == "ulaw" else ALAW_TABLE).astype(np.int32)
    order = np.argsort(table, kind="stable")
    srt = table[order]
    i = np.clip(np.searchsorted(srt, q.astype(np.int32)), 1, 255)
    pick = np.where(np.abs(srt[i] - q) < np.abs(srt[i - 1] - q), i, i - 1)
    return order[pick].astype(np.uint8)


def _expand(codes, law):
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import audioop
        f = audioop.ulaw2lin if law == "ulaw" else audioop.alaw2lin
        return np.frombuffer(f(codes.tobytes(), 2), "<i2").reshape(codes.shape)
    except ImportError:
        return (ULAW_TABLE if law == "ulaw" else ALAW_TABLE)[codes]


@pytest.mark.parametrize("law", ["ulaw", "alaw"])
def test_g711_push_equals_pcm_push(word, law):
    from easywakeword_b200.bank import WakeWordBank
    n, secs = 12, 24
    pcm = synth.stream_batch(8300, n, float(secs), word, gain=(2.0, 4.0))
    codes = _encode(pcm, law)
    lin = np.ascontiguousarray(_expand(codes, law))
    assert np.abs(lin.astype(np.int32) - pcm).max() < 1100           # companding error, largest at full scale
    out = []
    for feed in ("codes", "pcm"):
        bank = WakeWordBank(n, [word], device=0, buffer_seconds=5, speech_duration_min=0.5, speech_duration_max=1.6)
        try:
            evs = []
            for b in range(0, secs * 16000, 16000):
                if feed == "codes":
                    bank.push_g711(np.ascontiguousarray(codes[:, b:b + 16000]), law=law)
                else:
                    bank.push(np.ascontiguousarray(lin[:, b:b + 16000]))
                bank.tick(10)
                evs.append(bank.poll().copy())
            ring = np.stack([bank.ctx.read_last(s, 5 * 16000) for s in range(n)])
            out.append((np.concatenate(evs), bank.results(), ring))
        finally:
            bank.close()
    (e0, r0, g0), (e1, r1, g1) = out
    assert np.array_equal(g0, g1) and np.array_equal(g0, lin[:, -5 * 16000:].astype(np.float32) / np.float32(32768))
    assert len(e0) == len(e1) and (e0["kind"] == 2).sum() >= 6
    for f in e0.dtype.names:
        assert np.array_equal(e0[f], e1[f], equal_nan=e0[f].dtype.kind == "f"), f
    for f in r0.dtype.names:
        assert np.array_equal(r0[f], r1[f], equal_nan=r0[f].dtype.kind == "f"), f


def test_g711_unaligned_rows_and_errors(word):
    from easywakeword_b200 import _lib
    rng = np.random.default_rng(4)
    ctx = _lib.Context(device=0, n_streams=3, ring_samples=16000, slack_samples=16000, pcm_format=_lib.PCM_I16)
    ctx.set_stream_params(-1, live=1)
    big = rng.integers(0, 256, size=(3, 5003), dtype=np.uint8)
    codes = big[:, 1:4998]                                           # odd length, rows not 16-byte aligned, strided
    ctx.push_g711(codes, law="alaw")
    ctx.tick(1)
    for s in range(3):
        want = ALAW_TABLE[codes[s]].astype(np.float32) / np.float32(32768)
        assert np.array_equal(ctx.read_last(s, codes.shape[1]), want)
    with pytest.raises(Exception):
        ctx.push_g711(codes.astype(np.int16))
    with pytest.raises(Exception):
        ctx.push_g711(codes, law="g722")
    ctx.close()
    ctx = _lib.Context(device=0, n_streams=1, ring_samples=16000, slack_samples=16000, pcm_format=_lib.PCM_F32)
    with pytest.raises(Exception):
        ctx.push_g711(codes[:1])                                     # float rings take no G.711 feed
    ctx.close()
