"""CPU tests: the oracle (oracle/bwt_oracle.c, oracle/stages.py) against the reference's golden
vectors (tests/golden/ref_vectors.json, produced by oracle/_ref/deBWT -t 1) -- SURVEY.md §8c."""
import random

import numpy as np
import pytest

from oracle import coracle, stages as st
from tests.util import as_bytes_records, golden, seeded_records, sha

G = golden()


@pytest.mark.parametrize("name", sorted(G["small"]))
def test_c_oracle_matches_reference_small(name):
    case = G["small"][name]
    sym, _ = st.text_from_records(case["records"])
    w, s, d = coracle.bwt(sym)
    assert w.tobytes().hex() == case["bwt"]
    assert s.tobytes().hex() == case["sharp"]
    assert d.tobytes().hex() == case["dollar"]


@pytest.mark.parametrize("name", sorted(G["small"]))
def test_stage_restatement_matches_reference_small(name):
    case = G["small"][name]
    w, s, d = st.out_bytes(*st.build_bwt(case["records"]))
    assert w.hex() == case["bwt"] and s.hex() == case["sharp"] and d.hex() == case["dollar"]


@pytest.mark.parametrize("name", sorted(G["seeded"]))
def test_c_oracle_matches_reference_seeded(name):
    case = G["seeded"][name]
    recs = as_bytes_records(seeded_records(name))
    assert sum(len(r) for r in recs) == case["n_bases"]
    sym, _ = st.text_from_records(recs)
    w, s, d = coracle.bwt(sym)
    assert sha(w.tobytes()) == case["bwt_sha256"]
    assert s.tobytes().hex() == case["sharp"]
    assert d.tobytes().hex() == case["dollar"]


def test_reference_crash_case_still_defined():
    # the reference binary crashes on a lone 33-bp record; the definition is still unambiguous
    for name, case in G.get("reference_crashes", {}).items():
        a = st.suffix_sort_bwt(case["records"])
        b = st.build_bwt(case["records"])
        assert all((x == y).all() for x, y in zip(a, b)), name


def test_stage_restatement_equals_suffix_sort_random():
    rng = random.Random(3)

    def rnd(n, alpha="ACGT"):
        return "".join(rng.choice(alpha) for _ in range(n))
    for it in range(40):
        mode = it % 4
        if mode == 0:
            recs = [rnd(rng.randint(33, 150)) for _ in range(rng.randint(1, 4))]
        elif mode == 1:
            base = rnd(rng.randint(40, 100))
            recs = []
            for _ in range(rng.randint(2, 4)):
                r = list(base)
                for _ in range(rng.randint(0, 3)):
                    r[rng.randrange(len(r))] = rng.choice("ACGT")
                recs.append("".join(r))
        elif mode == 2:
            recs = [rnd(rng.randint(33, 120), "AC") for _ in range(rng.randint(1, 3))]
        else:
            r = rnd(50)
            recs = [r, r, rnd(35) + r]
        a = st.suffix_sort_bwt(recs)
        b = st.build_bwt(recs)
        sym, _ = st.text_from_records(recs)
        c = coracle.bwt(sym)
        assert all((x == y).all() for x, y in zip(a, b)), recs
        assert all((x == y).all() for x, y in zip(a, c)), recs


def test_c_stage_helpers_match_numpy():
    recs = as_bytes_records(seeded_records("c4_like_5x100k"))[:2]
    sym, seps = st.text_from_records(recs)
    assert (coracle.pack_text(sym) == st.pack_text(sym)).all()
    keys = coracle.extract_keys(sym)
    assert (keys == st.extract_keys(sym, seps)).all()
    sk = coracle.sort_keys(keys)
    assert (sk == np.sort(keys)).all()
    km, ct = coracle.rle(sk)
    km2, ct2 = st.rle(sk)
    assert (km == km2).all() and (ct == ct2).all() and int(ct.sum()) == keys.size


@pytest.mark.parametrize("name", ["c4_like_5x100k", "c3_like_600k_3rec"])
def test_digit_histograms_from_text_equal_key_histograms(name):
    # the identity the CUDA path uses to skip the histogram sweep over the keys (text_hist_kernel)
    recs = as_bytes_records(seeded_records(name))
    recs = recs + [recs[0][:33], recs[0][:40]]             # shortest legal records: the two corrections nearly touch
    sym, seps = st.text_from_records(recs)
    want = st.digit_histograms(st.extract_keys(sym, seps))
    got = st.digit_histograms_from_text(sym, seps)
    assert (want == got).all()
    assert int(got[0].sum()) == sym.size - 32 * len(recs)


def test_lf_inversion_roundtrip():
    recs = as_bytes_records(seeded_records("c3_like_600k_3rec"))
    sym, _ = st.text_from_records(recs)
    b = coracle.bwt_symbols(sym)
    ok, t = coracle.invert_bwt(b)
    assert ok and (t == sym).all()
    b2 = b.copy()
    i, j = 1000, 77777
    if b2[i] != b2[j]:
        b2[i], b2[j] = b2[j], b2[i]
        ok2, t2 = coracle.invert_bwt(b2)
        assert (not ok2) or not (t2 == sym).all()


def test_rejects_short_record():
    with pytest.raises(ValueError):
        st.text_from_records(["ACGT" * 8])
