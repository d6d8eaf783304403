"""Per-CTA timeline of one K3 (segment_queue) launch, from a measuring build of the library:
    nvcc ... -DEWK_K3_TRACE -o experiments/libewk_trace.so easywakeword_b200/csrc/ewk_api.cu
    (on the GPU box) cp experiments/libewk_trace.so easywakeword_b200/libewk.so
    EWK_K3_TRACE=gpurun_out/k3_trace.bin python bench.py --no-cpu --no-extra --steps 8
    python profiles/tools/k3_timeline.py gpurun_out/k3_trace.bin
The library dumps the last launch's records when the bank closes: per CTA the SM id, start, tables-loaded and exit
times (globaltimer, ns) and per segment start / end / samples / event index."""
import sys

import numpy as np

SEGS, WORDS = 14, 8 + 4 * 14


def main(path):
    a = np.fromfile(path, dtype=np.int64).reshape(-1, WORDS)
    a = a[a[:, 1] > 0]
    t0 = a[:, 1].min()
    start, tab, end, nseg = (a[:, 1] - t0) / 1e3, (a[:, 2] - t0) / 1e3, (a[:, 3] - t0) / 1e3, a[:, 4]
    print(f"CTAs {len(a)}  SMs {len(set(a[:, 0]))}  segments {nseg.sum()}  span {end.max():.1f} us")
    q = lambda x: "min %.1f  p10 %.1f  median %.1f  p90 %.1f  max %.1f" % tuple(np.percentile(x, [0, 10, 50, 90, 100]))
    print("CTA start (us)        :", q(start))
    print("tables loaded - start :", q(tab - start))
    print("CTA exit (us)         :", q(end))
    print("segments per CTA      :", np.bincount(nseg))
    seg = a[:, 8:].reshape(len(a), SEGS, 4)
    dur, frames = [], []
    for c in range(len(a)):
        for k in range(min(nseg[c], SEGS)):
            dur.append((seg[c, k, 1] - seg[c, k, 0]) / 1e3)
            frames.append(1 + seg[c, k, 2] // 160)
    dur, frames = np.array(dur), np.array(frames)
    print("segment frames        :", q(frames))
    print("segment time (us)     :", q(dur))
    rounds = -(-frames // 14)
    for r in sorted(set(rounds)):
        m = rounds == r
        print(f"  {r:2d} rounds of 14 frames: n {m.sum():4d}  time median {np.median(dur[m]):.1f} us  min {dur[m].min():.1f}  max {dur[m].max():.1f}")
    # first vs later segments of a CTA (SM shared by 2 CTAs all the time vs. a neighbour that has already left)
    first = np.array([(seg[c, 0, 1] - seg[c, 0, 0]) / 1e3 / max(1, -(-(1 + seg[c, 0, 2] // 160) // 14)) for c in range(len(a)) if nseg[c] > 0])
    last = np.array([(seg[c, nseg[c] - 1, 1] - seg[c, nseg[c] - 1, 0]) / 1e3 / max(1, -(-(1 + seg[c, nseg[c] - 1, 2] // 160) // 14))
                     for c in range(len(a)) if 1 < nseg[c] <= SEGS])
    print("us per round, first segment of a CTA:", q(first))
    if len(last):
        print("us per round, last segment of a CTA :", q(last))
    # SM-level: when does an SM run out of work
    sm_end = {}
    for c in range(len(a)):
        sm_end[a[c, 0]] = max(sm_end.get(a[c, 0], 0.0), end[c])
    e = np.array(list(sm_end.values()))
    print("SM idle from (us)     :", q(e), f" mean busy share {e.mean() / end.max():.3f}")
    busy = sum(dur) / (len(a) * end.max())
    print(f"CTA-time inside segments / (CTAs x span): {busy:.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
