"""Threshold tuning with per-hop scores — what the reference's examples/tune_threshold.py does by calling
WordMatcher.calculate_similarity on consecutive chunks, at 10 ms resolution and for whole recordings at once.

    python examples/tune_threshold_dense.py word.wav recording1.wav recording2.wav ...

Prints, per recording, the distribution of the per-hop similarity and the hops above a few thresholds.
"""
import sys

import numpy as np

from easywakeword_b200 import WakeWordBank
from easywakeword_b200 import synth
from easywakeword_b200.wakeword import load_wav_16k


def main():
    if len(sys.argv) >= 3:
        word = load_wav_16k(sys.argv[1])
        recs = [load_wav_16k(p) for p in sys.argv[2:]]
    else:
        word = synth.synthetic_word()
        recs = [synth.stream(i, 20.0, word, gain=(1.0, 3.0))[0] for i in range(4)]
    n = min(len(r) for r in recs) // 160 * 160
    pcm = np.stack([synth.to_int16(r[:n]) for r in recs])
    bank = WakeWordBank(len(recs), [word], buffer_seconds=int(np.ceil(n / 16000)) + 1, max_push_seconds=n / 16000)
    (hop0, scores), = bank.dense_sweep([pcm])
    bank.close()
    s = scores[:, :, 0]
    for i in range(len(recs)):
        v = s[i][~np.isnan(s[i])]
        print(f"recording {i}: per-hop similarity min {v.min():5.1f}  median {np.median(v):5.1f}  max {v.max():5.1f}  |  "
              + "  ".join(f">={t}: {int((v >= t).sum()):5d} hops" for t in (75, 85, 95)))


if __name__ == "__main__":
    main()
