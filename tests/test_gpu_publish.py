"""GPU: peer publication of the per-stream result records (ewk_set_results_peers) — the multi-GPU exchange done by
K2 / K3 themselves.  On one GPU the "peers" are two local buffers; `test_two_gpu_exchange` runs the real thing over
NVLink under torchrun when the box has two devices (the driver's one-GPU box skips it)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("overlap", [False, True])
def test_published_copies_equal_results(word, overlap):
    import torch
    from easywakeword_b200.bank import WakeWordBank
    from easywakeword_b200.synth import stream_batch
    n, stride, off = 24, 64, 17
    pcm16 = stream_batch(5100, n, 20.0, word, distractor_prob=0.3)
    dests = [torch.full((2, stride, 2), -7, dtype=torch.int32, device="cuda:0") for _ in range(2)]
    sigs = [torch.zeros(2, 16, dtype=torch.int64, device="cuda:0") for _ in range(2)]
    bank = WakeWordBank(n, [word], device=0, buffer_seconds=5, speech_duration_min=0.5, speech_duration_max=1.6)
    ctx = bank.ctx
    try:
        ctx.set_overlap(overlap)
        assert ctx.publish_parity() == -1
        ctx.set_results_peers([d.data_ptr() for d in dests], stride_records=stride, offset_records=off,
                              signals=[g.data_ptr() for g in sigs], slot=0)
        assert ctx.publish_parity() == -1 and ctx.publish_seq() == 0
        prev, n_scored = None, 0
        for i, b in enumerate(range(0, pcm16.shape[1], 16000)):
            bank.step(np.ascontiguousarray(pcm16[:, b:b + 16000]))
            par = ctx.publish_parity()
            assert par == i % 2 and ctx.publish_seq() == i + 1
            ctx.wait_published(1, i + 1)                       # device-side gate on the own slot: K3's completion signal
            seqs, timed_out = ctx.published_seq(par, 2)
            assert not timed_out and seqs.tolist() == [i + 1, 0]
            for g in sigs:                                     # both destinations got the signal, in this call's row only
                assert g.cpu().numpy()[par].tolist() == [i + 1] + [0] * 15
            res = ctx.results()                                # joins K3 and synchronises
            raw = res.view(np.int32).reshape(n, 2)
            for d in dests:
                h = d.cpu().numpy()
                np.testing.assert_array_equal(h[par, off:off + n], raw)
                assert (h[par, :off] == -7).all() and (h[par, off + n:] == -7).all()
                if prev is not None:                           # the other parity still holds the previous call
                    np.testing.assert_array_equal(h[1 - par, off:off + n], prev)
            prev = raw.copy()
            n_scored += int((res["flags"] & 16 != 0).sum())
        assert n_scored > 5                                    # K3 published scored records, not only K2 flags
        # nobody ever writes slot 1: a wait that includes it gives up after its timeout instead of hanging the device
        ctx.wait_published(2, ctx.publish_seq(), timeout_ms=30)
        _, timed_out = ctx.published_seq(ctx.publish_parity(), 2)
        assert timed_out
        _, timed_out = ctx.published_seq(ctx.publish_parity(), 2)
        assert not timed_out                                   # the flag is cleared by the read
        # switching it off stops the stores
        ctx.set_results_peers([])
        for d in dests:
            d.fill_(-7)
        bank.step(np.ascontiguousarray(pcm16[:, :16000]))
        ctx.results()
        assert all(bool((d == -7).all()) for d in dests)
    finally:
        bank.close()


def test_publish_argument_errors(word):
    import torch
    from easywakeword_b200 import _lib
    ctx = _lib.Context(device=0, n_streams=4, ring_samples=16000, slack_samples=16000)
    buf = torch.zeros(2, 8, 2, dtype=torch.int32, device="cuda:0")
    with pytest.raises(Exception):
        ctx.set_results_peers([buf.data_ptr()], stride_records=3, offset_records=0)      # cannot hold 4 records
    with pytest.raises(Exception):
        ctx.set_results_peers([buf.data_ptr()], stride_records=8, offset_records=5)      # 5 + 4 > 8
    with pytest.raises(Exception):
        ctx.set_results_peers([buf.data_ptr() + 4], stride_records=8, offset_records=0)  # misaligned
    with pytest.raises(Exception):
        ctx.set_results_peers([buf.data_ptr()] * 17, stride_records=8, offset_records=0)
    with pytest.raises(Exception):
        ctx.set_results_peers([buf.data_ptr()], stride_records=8, offset_records=0, signals=[buf.data_ptr()], slot=1)
    ctx.set_results_peers([buf.data_ptr()], stride_records=8, offset_records=4)
    with pytest.raises(Exception):
        ctx.wait_published(1, 1)                               # no signal rows installed
    ctx.close()


def test_two_gpu_exchange():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs of one NVLink domain")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(REPO, "tests", "mgpu_peer_exchange.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "peer exchange ok" in r.stdout
