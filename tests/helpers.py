"""Shared test helpers (stream regeneration identical to oracle/gen_golden.py)."""
import hashlib

import numpy as np

from easywakeword_b200 import synth


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def detect_stream_for(case, word):
    if case.get("special") == "config1":
        rng = np.random.default_rng(7)
        s = (rng.standard_normal(400000) * 0.002).astype(np.float32)
        s[224000:224000 + len(word)] += (3.0 * word).astype(np.float32)
        return s
    x, _ = synth.stream(case["seed"], case["seconds"], word, noise_sigma=case["noise"], gain=tuple(case["gain"]),
                        zero_gaps=case.get("zero_gaps", 0), distractor_prob=case.get("distractor_prob", 0.0),
                        inserts_per_10s=tuple(case.get("inserts", (1, 3))))
    return synth.from_int16(synth.to_int16(x))


def mfcc_rel_l2(a, b):
    """Per-frame relative L2 error of MFCC matrices [20, F] (SURVEY §7 'hard parts': the parity metric)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    num = np.linalg.norm(a - b, axis=0)
    den = np.linalg.norm(b, axis=0)
    return num / np.maximum(den, 1e-30)
