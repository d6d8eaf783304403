"""Many rooms on one GPU — the batched form of the reference's examples/multiroom_async.py (one WakeWord
object and one thread per room there; one WakeWordBank here).

    python examples/multiroom_bank.py computer.wav --rooms 64

Audio comes from `audio_source()`: replace it with whatever delivers [rooms, n] int16 blocks at 16 kHz
(here: synthetic rooms with the wake word mixed in).  Level 3 (speech-to-text) is any object with
`transcribe(audio) -> str`, e.g. the reference's WhisperTranscriber; without one, level-2 matches are reported.
"""
import argparse
import sys

import numpy as np

from easywakeword_b200 import WakeWordBank
from easywakeword_b200 import synth
from easywakeword_b200.wakeword import load_wav_16k


def audio_source(rooms, word, seconds=30, block=16000):
    streams = synth.stream_batch(1000, rooms, seconds, word, gain=(1.5, 4.0))
    for p in range(0, streams.shape[1], block):
        yield np.ascontiguousarray(streams[:, p:p + block])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("wav", nargs="?", help="16 kHz PCM16 WAV of the wake word (default: a synthetic word)")
    ap.add_argument("--rooms", type=int, default=64)
    ap.add_argument("--threshold", type=float, default=75.0)
    args = ap.parse_args()
    word = load_wav_16k(args.wav) if args.wav else synth.synthetic_word()
    bank = WakeWordBank(args.rooms, [word], similarity_threshold=args.threshold)

    def on_match(room, tick, score, text):
        print(f"room {room:4d}  t={tick / 10:6.1f} s  MFCC similarity {score:5.1f} %  {text or ''}")

    log = bank.run(audio_source(args.rooms, word), on_match=on_match)
    n2 = sum(1 for e in log if e["kind"] == 2)
    print(f"{n2} candidate segments scored, {sum(1 for e in log if e['kind'] == 2 and e['matched'])} matches", file=sys.stderr)
    bank.close()


if __name__ == "__main__":
    main()
