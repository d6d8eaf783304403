"""GPU, BASELINE.json's full single-GPU size (4096 streams): size-independent properties of the hot path
plus oracle spot checks.  The bank holds 64 distinct synthetic streams replicated 64 times in a shuffled
order, so every replica must produce bit-identical events / scores, and a sample of the distinct streams
must agree with the oracle decision for decision."""
import numpy as np
import pytest

from easywakeword_b200 import synth

pytestmark = pytest.mark.gpu
N, UNIQUE, SECONDS = 4096, 64, 18


@pytest.fixture(scope="module")
def big(word):
    uniq = np.stack([synth.to_int16(synth.stream(7000 + i, SECONDS, word, gain=(1.0, 4.0),
                                                 distractor_prob=0.3 if i % 3 == 0 else 0.0)[0]) for i in range(UNIQUE)])
    perm = np.random.default_rng(1).permutation(N)
    which = perm % UNIQUE                     # stream s carries distinct stream which[s]
    return uniq, which


def test_full_size_gated_path(big, word):
    from easywakeword_b200.bank import WakeWordBank
    from oracle import ewk_oracle as O
    uniq, which = big
    bank = WakeWordBank(N, [word], frame_size=1600, similarity_threshold=95.0, speech_duration_min=0.69,
                        speech_duration_max=1.38, timeout=4.0, max_events=1 << 16)
    events = []
    block = np.empty((N, 16000), np.int16)
    for p in range(0, SECONDS * 16000, 16000):
        block[:] = uniq[which, p:p + 16000]
        bank.step(block)
        events.append(bank.poll().copy())
        assert bank.ctx.dropped == 0
    ev = np.concatenate(events)
    res = bank.results()
    bank.close()
    assert len(ev) > N                                       # timeouts (4 s) and level-2 evaluations everywhere
    # replication invariance: every replica of a distinct stream reports exactly the same event list
    by_stream = {}
    order = np.lexsort((ev["kind"], ev["tick"], ev["stream"]))
    ev = ev[order]
    starts = np.searchsorted(ev["stream"], np.arange(N + 1))
    fields = ("kind", "tick", "seg_start", "seg_len", "template_slot", "score", "matched")
    canon = {}
    n_l2 = 0
    for s in range(N):
        mine = ev[starts[s]:starts[s + 1]]
        key = tuple(mine[f].tobytes() for f in fields)
        u = int(which[s])
        if u in canon:
            assert key == canon[u][0], f"stream {s} (copy of {u}) differs from stream {canon[u][1]}"
        else:
            canon[u] = (key, s)
            n_l2 += int((mine["kind"] == 2).sum())
    assert n_l2 >= 40
    # the dense per-stream result record agrees with the last level-2 event of the stream
    for s in range(0, N, 97):
        mine = ev[starts[s]:starts[s + 1]]
        l2 = mine[mine["kind"] == 2]
        if len(l2):
            assert res["score"][s] == l2["score"][-1] and (res["flags"][s] & 1) == l2["matched"][-1]
            assert (res["flags"][s] >> 8) == len(l2)
        else:
            assert np.isnan(res["score"][s])
    # oracle spot check on a sample of the distinct streams: identical decisions, scores within 0.01
    n_dec = n_no = 0
    for u in range(0, UNIQUE, 3):
        s = canon[u][1]
        mine = ev[starts[s]:starts[s + 1]]
        o = O.detect_stream(synth.from_int16(uniq[u]), word, block=1600, fast=True, similarity_threshold=95.0,
                            speech_duration_min=0.69, speech_duration_max=1.38, timeout=4.0)
        l2 = mine[mine["kind"] == 2]
        assert list(l2["tick"]) == [e["tick"] for e in o["events"]]
        assert list(l2["seg_len"]) == [e["seg_len"] for e in o["events"]]
        assert list(mine[mine["kind"] == 1]["tick"]) == o["timeouts"]
        for a, b in zip(l2, o["events"]):
            assert abs(float(a["score"]) - b["score"]) <= 0.01
            if abs(b["score"] - 95.0) > 0.01:
                assert bool(a["matched"]) == b["matched"]
                n_dec += 1
                n_no += not b["matched"]
    assert n_dec >= 8
    print(f"full size: {len(ev)} events over {N} streams, {n_l2} distinct level-2 evaluations, "
          f"{n_dec} decisions checked against the oracle ({n_no} of them no-match)")


def test_full_size_dense_properties(big, word):
    from easywakeword_b200 import _lib
    uniq, which = big
    secs = 4
    ctx = _lib.Context(device=0, n_streams=N, ring_samples=secs * 16000, slack_samples=3200, pcm_format=_lib.PCM_I16)
    ctx.set_template(0, word)
    ctx.set_stream_params(-1, live=1)
    ctx.push(np.ascontiguousarray(uniq[which, :secs * 16000]))
    hop0, nh = 100, 300
    sc = ctx.dense_scores(hop0, nh, 0, 1)[:, :, 0]
    ctx.close()
    assert sc.shape == (N, nh) and not np.isnan(sc).any()
    assert ((sc >= 0) & (sc <= 100.0001)).all()
    # replicas agree bit for bit; distinct streams do not
    first = {}
    for s in range(N):
        u = int(which[s])
        if u in first:
            assert np.array_equal(sc[s], sc[first[u]])
        else:
            first[u] = s
    assert not np.array_equal(sc[first[0]], sc[first[1]])
    # oracle on a handful of (stream, hop) pairs
    from oracle import ewk_oracle as O
    rng = np.random.default_rng(5)
    for _ in range(6):
        u = int(rng.integers(0, UNIQUE)); h = int(rng.integers(hop0, hop0 + nh))
        ref = O.dense_scores(synth.from_int16(uniq[u, :secs * 16000]), [word], [h])[0, 0]
        assert abs(float(sc[first[u], h - hop0]) - float(ref)) <= 0.01


def test_config4_shard_size_overlap_equals_sequential(big, word):
    """8192 streams on one GPU — the per-GPU shard of BASELINE configs[3] (65 536 streams / 8 GPUs) — pushed from device
    memory in overlap mode (K3 beside the next bulk push) and in sequential order: identical event lists and result
    records, and every replica of a distinct stream reports the same events."""
    import torch
    from easywakeword_b200 import _lib
    from easywakeword_b200.bank import WakeWordBank
    uniq, _ = big
    n8, secs = 8192, 14
    which8 = np.random.default_rng(2).permutation(n8) % UNIQUE
    dev_pcm = torch.from_numpy(np.ascontiguousarray(
        uniq[which8, :secs * 16000].reshape(n8, secs, 16000).transpose(1, 0, 2))).to("cuda:0")       # [secs][n8][16000]
    out = []
    for overlap in (False, True):
        bank = WakeWordBank(n8, [word], frame_size=1600, similarity_threshold=95.0, speech_duration_min=0.69,
                            speech_duration_max=1.38, timeout=4.0, max_events=1 << 17, max_push_seconds=2.0)
        bank.ctx.set_overlap(overlap)
        try:
            evs = []
            for j in range(secs):
                bank.push((dev_pcm[j].data_ptr(), n8, 16000, 16000), where=_lib.DEVICE)
                bank.tick(10)
                if j % 4 == 3:
                    evs.append(bank.poll().copy())
                    assert bank.ctx.dropped == 0
            evs.append(bank.poll().copy())
            out.append((np.concatenate(evs), bank.results()))
        finally:
            bank.close()
    (e0, r0), (e1, r1) = out
    assert len(e0) == len(e1) and (e0["kind"] == 2).sum() > 2 * UNIQUE
    for f in e0.dtype.names:
        assert np.array_equal(e0[f], e1[f], equal_nan=e0[f].dtype.kind == "f"), f
    for f in r0.dtype.names:
        assert np.array_equal(r0[f], r1[f], equal_nan=r0[f].dtype.kind == "f"), f
    order = np.lexsort((e1["kind"], e1["tick"], e1["stream"]))
    ev = e1[order]
    starts = np.searchsorted(ev["stream"], np.arange(n8 + 1))
    canon = {}
    for s in range(n8):
        mine = ev[starts[s]:starts[s + 1]]
        key = tuple(mine[f].tobytes() for f in ("kind", "tick", "seg_start", "seg_len", "score", "matched"))
        u = int(which8[s])
        assert canon.setdefault(u, key) == key, f"stream {s} (copy of {u}) differs"
