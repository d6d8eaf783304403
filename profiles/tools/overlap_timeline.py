"""Per-kernel elapsed times (CUDA events on each kernel's own stream) with K3 beside the next push, i.e. including
the interference the kernels cause each other, next to the same loop in sequential order.
    python profiles/tools/overlap_timeline.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from easywakeword_b200 import _lib, synth                      # noqa: E402
from easywakeword_b200.bank import WakeWordBank                # noqa: E402

n, steps = 4096, 20
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
word = np.load(os.path.join(REPO, "tests", "golden", "reference_word.npz"))["pcm_i16"].astype(np.float32) / np.float32(32768.0)
pool = np.stack([synth.stream_batch(100 + 64 * (s // 64), 64, 10.0, word) for s in range(0, n, 64)]).reshape(n, -1)
dev = torch.device("cuda", 0)
pool_dev = torch.from_numpy(np.ascontiguousarray(pool.reshape(n, 10, 16000).transpose(1, 0, 2))).to(dev)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
bank = WakeWordBank(n, [word], device=0, buffer_seconds=10, max_push_seconds=2.0, cuda_stream=stream.cuda_stream,
                    max_events=1 << 17, speech_duration_min=0.69, speech_duration_max=1.38)
ctx = bank.ctx


def step(i):
    t = pool_dev[i % 10]
    bank.push((t.data_ptr(), n, 16000, 16000), where=_lib.DEVICE)
    bank.tick(10)


for i in range(12):
    step(i)
bank.poll()
for overlap in (False, True):
    ctx.set_overlap(overlap)
    for i in range(5):
        step(i)
    bank.poll()
    ctx.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        step(i)
    ctx.join()
    e1.record(stream)
    torch.cuda.synchronize()
    prof = ctx.profile_read()
    ctx.profile(False)
    bank.poll()
    print("overlap" if overlap else "sequential", "step us", round(e0.elapsed_time(e1) / steps * 1e3, 1),
          {k: round(v["ms"] / max(1, v["launches"]) * 1e3, 1) for k, v in prof.items() if v["launches"]})
