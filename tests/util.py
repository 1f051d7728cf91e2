"""Shared helpers for the parity tests."""
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def golden():
    with open(os.path.join(HERE, "golden", "ref_vectors.json")) as f:
        return json.load(f)


def seeded_records(name):
    from debwt_b200 import synth
    return {
        "c1_like_200k": lambda: synth.config1(200_000),
        "c2_like_1m": lambda: synth.config2(1_000_000, 10_000, 5, 0.05),
        "c4_like_5x100k": lambda: synth.config4(100_000, 5, 0.001),
        "c3_like_600k_3rec": lambda: synth.config3(600_000, 3),
    }[name]()


def as_bytes_records(recs):
    return [r.encode() if isinstance(r, str) else bytes(np.asarray(r, dtype=np.uint8)) for r in recs]


def sha(b):
    return hashlib.sha256(b).hexdigest()
