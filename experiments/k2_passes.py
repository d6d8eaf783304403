"""Why does K2 (tick_gate) take 30 us in bench.py's first profile pass and 25 us in a later one?  Repeats the profile loop
under different conditions on one bank (bench workload) and prints the per-kernel times."""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
import bench
from bench import *   # noqa
from easywakeword_b200 import _lib
from easywakeword_b200.bank import WakeWordBank

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
n = N_STREAMS
word, _ = load_word()
pool = np.empty((POOL_SECONDS, n, STEP_SAMPLES), np.int16)
make_pool(0, n, word, pool)
pool_dev = torch.from_numpy(pool).to(dev)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
bank = WakeWordBank(n, [word], device=0, buffer_seconds=RING_SECONDS, pcm_dtype=np.int16, max_push_seconds=2 * STEP_SECONDS,
                    cuda_stream=stream.cuda_stream, max_events=1 << 17, **PARAMS)
ctx = bank.ctx
j = [0]


def step(poll):
    bank.push((pool_dev.data_ptr() + (j[0] % POOL_SECONDS) * n * STEP_SAMPLES * 2, n, STEP_SAMPLES, STEP_SAMPLES), where=_lib.DEVICE)
    j[0] += 1
    bank.tick(TICKS_PER_STEP)
    return len(bank.poll()) if poll else 0


for _ in range(RING_SECONDS + 3):
    step(True)


def run(tag, K=20, poll=False, overlap=False):
    ctx.set_overlap(overlap)
    ctx.profile(True)
    for _ in range(K):
        step(poll)
    p = ctx.profile_read()
    ctx.profile(False)
    print(tag, {k: round(1e3 * v["ms"] / max(1, v["launches"]), 2) for k, v in p.items() if v["launches"]}, "j", j[0], flush=True)


run("A no poll, sequential   ")
results = torch.zeros(n, 2, dtype=torch.int32, device=dev)
ctx.set_results_buffer(results.data_ptr())
run("A2 external result array")
pin = _lib.PinnedArray((POOL_SECONDS, n, STEP_SAMPLES), np.int16)
pin.array[...] = pool
host = torch.from_numpy(pin.array)
ctx.set_overlap(True)
for _ in range(6):                       # host pushes issued one step ahead, as bench.py's e2e leg does
    bank.push((host.data_ptr() + (j[0] % POOL_SECONDS) * n * STEP_SAMPLES * 2, n, STEP_SAMPLES, STEP_SAMPLES), where=_lib.HOST)
    j[0] += 1
    bank.tick(TICKS_PER_STEP)
    bank.poll()
run("A3 after host pushes    ")
run("A4 again                ")
run("B no poll again         ")
bank.poll()
run("C after poll            ")
run("D polling every step    ", poll=True)
run("E 10 steps              ", K=10)
run("F 10 steps              ", K=10)
run("G 5 steps               ", K=5)
run("H 5 steps               ", K=5)
run("I overlap mode, no poll ", overlap=True)
run("J sequential            ")
bank.close()
