"""torchrun script (world_size >= 2, one rank per GPU): the peer-published global record array of every rank must equal
the NCCL all-gather of the ranks' local records after every step, in sequential and in overlap mode, for unequal
shards.  Run by tests/test_gpu_publish.py::test_two_gpu_exchange and by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_peer_exchange.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    from easywakeword_b200.bank import WakeWordBank
    from easywakeword_b200.dist import PeerResultExchange, ResultGather
    from easywakeword_b200.synth import stream_batch
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    word = np.load(os.path.join(REPO, "tests", "golden", "reference_word.npz"))["pcm_i16"].astype(np.float32) / np.float32(32768)
    n_total = 37                                              # unequal shards
    pcm16 = stream_batch(6100, n_total, 16.0, word, distractor_prob=0.3)
    ex = PeerResultExchange(n_total, world, rank, dev)
    ga = ResultGather(n_total, world, rank, device=dev)
    mine = pcm16[ex.first:ex.last]
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    checked = 0
    for overlap, lockstep in ((False, False), (True, False), (True, True)):
        bank = WakeWordBank(ex.last - ex.first, [word], device=local, buffer_seconds=5, speech_duration_min=0.5,
                            speech_duration_max=1.6, cuda_stream=stream.cuda_stream)
        ctx = bank.ctx
        ctx.set_results_buffer(ga.local.data_ptr())
        ctx.set_overlap(overlap)
        ex.install(ctx)
        for b in range(0, mine.shape[1], 16000):
            bank.step(np.ascontiguousarray(mine[:, b:b + 16000]))
            if lockstep:                                      # symmetric-memory barrier behind K3
                ex.barrier(ctx)
                ex.finish(stream)
            else:                                             # put-with-signal: wait for every rank's sequence number
                ex.wait(ctx)
                seqs, timed_out = ex.published(ctx)
                assert not timed_out and (seqs == ctx.publish_seq()).all(), (seqs, ctx.publish_seq())
            ctx.join()
            ref = ga.gather()                                 # NCCL all-gather of the local records
            got = ex.records(ctx.publish_parity())
            torch.cuda.synchronize(dev)
            assert torch.equal(got, ref), f"rank {rank} overlap {overlap} lockstep {lockstep} step {b // 16000}"
            checked += 1
        scored = int(((ref[:, 1] >> 8) > 0).sum())
        assert scored > 5, scored
        ctx.set_results_peers([])
        bank.close()
    # the packaged form: ShardedBank with either exchange returns the same global record array
    from easywakeword_b200.dist import ShardedBank
    outs = {}
    for mode in ("nccl", "peer"):
        sb = ShardedBank(n_total, [word], world=world, rank=rank, device=local, exchange=mode, buffer_seconds=5,
                         speech_duration_min=0.5, speech_duration_max=1.6, cuda_stream=stream.cuda_stream, overlap=True)
        assert (sb.peer is not None) == (mode == "peer")
        got = []
        for b in range(0, mine.shape[1], 16000):
            sb.step(np.ascontiguousarray(mine[:, b:b + 16000]))
            got.append(sb.gather().clone())
            dist.barrier()                                    # a consumer that lags must not be overrun by two calls
        outs[mode] = torch.stack(got)
        sb.close()
    assert torch.equal(outs["nccl"], outs["peer"])
    dist.barrier()
    if rank == 0:
        print(f"peer exchange ok: {checked} steps x {world} ranks identical to the all-gather")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
