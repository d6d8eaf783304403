"""Plain host->device copy bandwidth with 1..N ranks copying at the same time (one process per GPU, torchrun).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/tools/h2d_probe.py [--mb 131]

Every rank owns a pinned host buffer of the bench's step size (131 MB = 4096 streams x 1.0 s int16) and copies it to
its GPU with one cudaMemcpyAsync (torch copy_ non_blocking) per iteration; for k in 1, 2, 4, ..., N the first k ranks
copy together (barrier before, CUDA events around) while the others wait.  Prints one JSON line: per k the GB/s of every
active rank, their min and their sum — the platform's ceiling for the end-to-end leg of bench.py at that rank count
(VERDICT r1 #3a: is 23 GB/s per GPU at 8 ranks the host's limit or the bench's?)."""
import argparse
import json
import os

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=131.072)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = int(args.mb * 1e6) // 2
    host = torch.empty(n, dtype=torch.int16).pin_memory()
    host.fill_(rank + 1)
    dst = torch.empty(n, dtype=torch.int16, device=dev)
    back = torch.empty(n, dtype=torch.int16).pin_memory()
    out = {}

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for direction in ("h2d", "d2h"):
        k = 1
        while k <= world:
            active = rank < k
            for _ in range(3):
                if active:
                    (dst.copy_(host, non_blocking=True) if direction == "h2d" else back.copy_(dst, non_blocking=True))
            sync()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if active:
                for _ in range(args.reps):
                    (dst.copy_(host, non_blocking=True) if direction == "h2d" else back.copy_(dst, non_blocking=True))
            b.record()
            sync()
            gbs = (n * 2 * args.reps / (a.elapsed_time(b) * 1e-3) / 1e9) if active else 0.0
            if world > 1:
                t = torch.zeros(world, device=dev)
                t[rank] = gbs
                dist.all_reduce(t)
                vals = [round(float(v), 2) for v in t[:k]]
            else:
                vals = [round(gbs, 2)]
            out[f"{direction}_{k}"] = {"per_rank_gbs": vals, "min": min(vals), "sum": round(sum(vals), 2)}
            k *= 2
    if rank == 0:
        print(json.dumps({"probe": "pinned host <-> device copy, concurrent ranks", "mb_per_copy": args.mb, "world": world, **out}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
