"""Generates tests/golden/ref_vectors.json by running the compiled reference (oracle/_ref/deBWT,
built by `make -C oracle ref` from /root/reference/src) at -t 1 (the canonical setting).

Run in the build container only (needs oracle/_ref).  Small cases store the reference's three
output files verbatim (hex); larger cases store their SHA-256 and are regenerated from seeds.
"""
import hashlib
import json
import os
import random
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.refrun import run_reference  # noqa: E402
from debwt_b200 import synth  # noqa: E402


def rnd(rng, n, alpha="ACGT"):
    return "".join(rng.choice(alpha) for _ in range(n))


def small_cases():
    rng = random.Random(20261017)
    cases = {}
    cases["survey_golden"] = ["ACGTACGTTGCATGCAAGCTTAGCTAGGATCCATGCAAGCTA",
                              "TTGCATGCAAGCTTAGCTAGGATCCATGCAAGCTACGTACGG",
                              "GGATCCATGCAAGCTACGTACGTTGCATGCAAGCTTAGCTAG"]
    cases["single_min_record"] = [rnd(rng, 33)]
    cases["single_random_1k"] = [rnd(rng, 1000)]
    r = rnd(rng, 60)
    cases["pathological"] = ["A" * 70, "T" * 45, "AC" * 60, "ACG" * 30, r, r, rnd(rng, 33), "T" * 33, "A" * 70]
    base = rnd(rng, 1500)
    recs = []
    for _ in range(6):
        x = list(base)
        for _ in range(4):
            x[rng.randrange(len(x))] = rng.choice("ACGT")
        recs.append("".join(x))
    cases["haplotypes_6x1500"] = recs
    el = rnd(rng, 200)
    cases["planted_repeats"] = [rnd(rng, 300) + el + rnd(rng, 100) + el + rnd(rng, 50) + el + rnd(rng, 40), el + rnd(rng, 77) + el]
    cases["two_letter"] = [rnd(rng, 400, "AC"), rnd(rng, 300, "GT"), rnd(rng, 200, "AC")]
    cases["lowercase_mixed"] = [rnd(rng, 120).lower(), rnd(rng, 90)]
    return cases


def seeded_cases():
    """(name, generator call) -> records as numpy uint8 arrays"""
    return {
        "c1_like_200k": lambda: synth.config1(200_000),
        "c2_like_1m": lambda: synth.config2(1_000_000, 10_000, 5, 0.05),
        "c4_like_5x100k": lambda: synth.config4(100_000, 5, 0.001),
        "c3_like_600k_3rec": lambda: synth.config3(600_000, 3),
    }


def run(records):
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "in.fa")
        with open(fa, "w") as f:
            for i, r in enumerate(records):
                s = r if isinstance(r, str) else bytes(r).decode()
                f.write(f">r{i}\n")
                for j in range(0, len(s), 70):
                    f.write(s[j:j + 70] + "\n")
        return run_reference(fa, threads=1)


def main():
    out = {"_how": "oracle/_ref/deBWT -t 1 -k 32 (reference compiled from /root/reference/src with gcc -O2 -fcommon, "
                   "Jellyfish replaced by oracle/jellyfish_standin.c); see tests/golden/make_golden.py",
           "small": {}, "seeded": {}}
    for name, recs in small_cases().items():
        try:
            r = run(recs)
        except RuntimeError as e:      # the reference itself crashes on some tiny inputs; record that
            out.setdefault("reference_crashes", {})[name] = {"records": recs, "tail": str(e)[-120:]}
            print(name, "REFERENCE CRASHED")
            continue
        out["small"][name] = {"records": recs, "bwt": r.bwt.hex(), "sharp": r.sharp.hex(), "dollar": r.dollar.hex()}
        print(name, len(r.bwt))
    for name, gen in seeded_cases().items():
        recs = gen()
        r = run(recs)
        out["seeded"][name] = {"n_bases": int(sum(len(x) for x in recs)), "n_records": len(recs),
                               "bwt_sha256": hashlib.sha256(r.bwt).hexdigest(),
                               "sharp": r.sharp.hex(), "dollar": r.dollar.hex()}
        print(name, len(r.bwt))
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
