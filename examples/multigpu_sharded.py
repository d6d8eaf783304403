"""65 536 rooms over the GPUs of one box (BASELINE configs[3]): one process per GPU, contiguous shards, and the 8-byte
per-room result records reaching every GPU by peer stores from the producing kernels over NVLink (no collective).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/multigpu_sharded.py \
        [rooms=65536] [seconds=30]

Each rank feeds its own rooms (here: synthetic audio); `gather()` returns the records of ALL rooms in global order on
every rank once every rank's signal for the step has arrived.  The reference's shape for this is one WakeWord object
and one thread per room (/root/reference/examples/multiroom_async.py:14-35)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easywakeword_b200 import synth                                   # noqa: E402
from easywakeword_b200.dist import ResultGather, ShardedBank, bind_host_near_gpu   # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bind_host_near_gpu(local)
    rooms = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    seconds = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    word = synth.synthetic_word()
    bank = ShardedBank(rooms, [word], world=world, rank=rank, device=local, exchange="auto", frame_size=1600,
                       speech_duration_min=0.5, speech_duration_max=1.6, overlap=True)
    n_local = bank.last - bank.first
    # 64 distinct synthetic rooms per rank, replicated (a real feed would arrive from the network)
    distinct = synth.stream_batch(9000 + 64 * rank, 64, float(seconds), word, gain=(1.5, 4.0))
    which = np.arange(n_local) % 64
    block = np.empty((n_local, 16000), np.int16)
    for t in range(seconds):
        block[:] = distinct[which, t * 16000:(t + 1) * 16000]
        bank.step(block)                                              # K1 push + 10 ticks of K2 / K3 for this rank's rooms
        records = ResultGather.decode(bank.gather())                  # every room of every rank, on every rank
        for ev in bank.poll_global():                                 # this rank's level-2 evaluations (global room ids)
            if ev["kind"] == 2 and ev["matched"] and rank == 0 and ev["stream"] < 4:
                print(f"t={t + 1:2d}s room {int(ev['stream'])}: score {float(ev['score']):.1f}")
        if rank == 0 and t % 10 == 9:
            heard = int(((records["flags"] >> 8) > 0).sum())
            print(f"t={t + 1}s: {heard} of {rooms} rooms have had a level-2 evaluation "
                  f"({'peer stores' if bank.peer is not None else 'NCCL all-gather'})")
    bank.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
