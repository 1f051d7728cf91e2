"""K9 (branch codes + blue entries) and K10 (segmented sort) through their own C-ABI entry points, against the stage
restatements of oracle/stages.py (reference src/generateSP.c:534-683, src/sortBlue.c:76-280)."""
import numpy as np
import pytest

from debwt_b200 import api
from oracle import stages as st
from tests.util import as_bytes_records, golden, seeded_records

pytestmark = pytest.mark.gpu
G = golden()


def _oracle_k9(recs):
    sym, seps = st.text_from_records(recs)
    sk = st.sort_keys(st.extract_keys(sym, seps))
    gmask, _ = st.group_masks(sk, sym, seps)
    specials = st.special_suffixes(sym, seps, sk)
    return st.branch_codes(sym, seps, sk, gmask, specials)


def _cases():
    out = {n: [r.upper().encode() for r in c["records"]] for n, c in G["small"].items()}
    out["c4_like_5x100k"] = as_bytes_records(seeded_records("c4_like_5x100k"))
    out["c3_like_600k_3rec"] = as_bytes_records(seeded_records("c3_like_600k_3rec"))
    return out


@pytest.mark.parametrize("name", sorted(_cases()))
def test_k9_codes_and_blue_entries_match_restatement(name):
    recs = _cases()[name]
    sp, blue = _oracle_k9(recs)
    text, seps = api.join_records(recs)
    codes, head, spi, prev = api.k_codes(text, seps)
    assert codes.size == sp.size and (codes == sp).all()
    got = sorted(zip(head.tolist(), spi.tolist(), prev.tolist()))
    assert got == sorted(blue)


def _check_sorted(codes, offs, spi, prev, got_spi, got_prev):
    spb = bytes(codes.tolist())
    for s in range(offs.size - 1):
        lo, hi = int(offs[s]), int(offs[s + 1])
        want = sorted(zip(spi[lo:hi].tolist(), prev[lo:hi].tolist()), key=lambda e: spb[e[0]:])
        assert sorted(got_spi[lo:hi].tolist()) == sorted(spi[lo:hi].tolist())            # a permutation inside the segment
        assert got_prev[lo:hi].tolist() == [p for _, p in want], s                         # what the BWT sees


@pytest.mark.parametrize("name", ["haplotypes_6x1500", "planted_repeats", "c4_like_5x100k", "c3_like_600k_3rec"])
def test_k10_sorts_the_restatements_segments(name):
    recs = _cases()[name]
    sp, blue = _oracle_k9(recs)
    blue.sort(key=lambda e: e[0])
    heads = np.array([e[0] for e in blue], dtype=np.int64)
    offs = np.concatenate(([0], np.flatnonzero(np.diff(heads)) + 1, [heads.size])).astype(np.uint64)
    spi = np.array([e[1] for e in blue], dtype=np.uint64)
    prev = np.array([e[2] for e in blue], dtype=np.uint8)
    got_spi, got_prev = api.k_sort_blue(sp, offs, spi, prev)
    _check_sorted(sp, offs, spi, prev, got_spi, got_prev)


@pytest.mark.parametrize("kind", ["random", "long_common_prefixes", "separators", "one_huge_segment"])
def test_k10_synthetic_code_strings(kind):
    rng = np.random.default_rng(11)
    if kind == "random":
        codes = rng.integers(0, 4, size=50_000).astype(np.uint8)
        sizes = rng.integers(2, 300, size=300)
    elif kind == "long_common_prefixes":                       # a 700-code unit repeated: comparisons run hundreds of codes deep
        unit = rng.integers(0, 4, size=700).astype(np.uint8)
        codes = np.tile(unit, 60)
        codes[rng.integers(0, codes.size, size=40)] ^= 1
        sizes = rng.integers(2, 2000, size=40)
    elif kind == "separators":                                 # '#' codes inside the strings: '#' > T, equal '#' compared through
        codes = rng.integers(0, 4, size=30_000).astype(np.uint8)
        codes[rng.integers(0, codes.size, size=600)] = 4
        codes[::97] = 4
        sizes = rng.integers(2, 600, size=120)
    else:
        unit = rng.integers(0, 4, size=90).astype(np.uint8)
        codes = np.tile(unit, 400)
        codes[rng.integers(0, codes.size, size=200)] ^= 2
        sizes = np.array([20_000])
    codes = np.concatenate([codes, [5]]).astype(np.uint8)      # the one '$' code ends the sequence
    offs = np.concatenate(([0], np.cumsum(sizes))).astype(np.uint64)
    m = int(offs[-1])
    spi = np.concatenate([rng.choice(codes.size - 1, size=int(s), replace=False) for s in sizes]).astype(np.uint64)
    prev = rng.integers(0, 4, size=m).astype(np.uint8)
    prev[rng.integers(0, m, size=5)] = 4
    got_spi, got_prev = api.k_sort_blue(codes, offs, spi, prev)
    _check_sorted(codes, offs, spi, prev, got_spi, got_prev)
