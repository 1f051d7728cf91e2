"""f2: FM-index tables (occ checkpoints, C-array), backward-search counts and the LF-inversion verifier, through the C
ABI, against numpy / the oracle.  Also the on-device workload generators against debwt_b200/synth.py."""
import numpy as np
import pytest

from debwt_b200 import api, synth
from oracle import coracle, stages as st
from tests.util import as_bytes_records, seeded_records

pytestmark = pytest.mark.gpu


def _decode(words, n):
    sh = (2 * (31 - np.arange(32))).astype(np.uint64)
    return ((words[:, None] >> sh[None, :]) & np.uint64(3)).astype(np.uint8).reshape(-1)[:n]


def _occ_numpy(words, sharp, dollar, n):
    codes = _decode(words, n).astype(np.int64)
    codes[sharp.astype(np.int64)] = 4
    codes[int(dollar[0])] = 4
    rows = (n >> 5) + 1
    occ = np.zeros((rows, 4), dtype=np.uint64)
    for c in range(4):
        cum = np.concatenate(([0], np.cumsum(codes == c)))
        occ[:, c] = cum[np.minimum(np.arange(rows) * 32, n)]
    tot = [(codes == c).sum() for c in range(4)]
    carr = np.array([0, tot[0], tot[0] + tot[1], tot[0] + tot[1] + tot[2], sum(tot), n - 1], dtype=np.uint64)
    return occ, carr


@pytest.mark.parametrize("name", ["c4_like_5x100k", "c3_like_600k_3rec", "c2_like_1m"])
def test_index_tables_counts_and_verifier(name):
    recs = as_bytes_records(seeded_records(name))
    text, seps = api.join_records(recs)
    n = text.size
    with api.BwtBuilder() as b:
        b.set_text(text, seps)
        b.build()
        words, sharp, dollar = b.result()
        occ, carr = b.index()
        want_occ, want_c = _occ_numpy(words, sharp, dollar, n)
        assert (carr == want_c).all()
        assert occ.shape == want_occ.shape and (occ == want_occ).all()
        # backward search against a plain count on the text (patterns: substrings, mutated substrings, absent)
        rng = np.random.default_rng(3)
        raw = text.tobytes()
        pats = []
        for L in (1, 2, 5, 12, 31, 32, 33, 64, 200):
            for _ in range(6):
                o = int(rng.integers(0, n - L - 1))
                p = raw[o:o + L]
                if b"#" in p or b"$" in p:
                    continue
                pats.append(p)
                q = bytearray(p)
                q[int(rng.integers(0, L))] = b"ACGT"[int(rng.integers(0, 4))]
                pats.append(bytes(q))
        pats.append(b"ACGT" * 40)
        got = b.count(pats)

        def plain_count(p):
            c, i = 0, raw.find(p)
            while i >= 0:
                c += 1
                i = raw.find(p, i + 1)
            return c
        assert [int(x) for x in got] == [plain_count(p) for p in pats]
        # verifier: the build inverts to T ...
        bad, _ = b.verify(text)
        assert bad == 0
        # ... and not to a text that differs in one base or in one separator position
        t2 = text.copy()
        t2[n // 2] = ord("A") if t2[n // 2] != ord("A") else ord("C")
        assert b.verify(t2)[0] > 0
    # an independent check of the same BWT: the oracle's sequential LF walk
    ok, inv = coracle.invert_bwt(coracle.unpack_bwt(words, n, sharp, dollar))
    assert ok and (inv == st.text_from_records(recs)[0]).all()


def test_verifier_on_small_golden_cases():
    from tests.util import golden
    for name, case in golden()["small"].items():
        recs = [r.upper().encode() for r in case["records"]]
        text, seps = api.join_records(recs)
        with api.BwtBuilder() as b:
            b.set_text(text, seps)
            b.build()
            assert b.verify(text)[0] == 0, name


def test_device_generators_match_numpy():
    torch = pytest.importorskip("torch")
    from debwt_b200 import synth_gpu
    assert (synth_gpu.random_bases(7, 100_003).cpu().numpy() == synth.random_bases(7, 100_003)).all()
    # genome_like at a scale where all three families have copies and copies overlap
    n = 4_000_000
    want = synth.genome_like(n, 5, scale=40.0)
    got = synth_gpu.genome_like(n, 5, scale=40.0).cpu().numpy()
    assert (got == want).all()
    base = synth.random_bases(11, 300_000)
    assert (synth_gpu.mutate(torch.from_numpy(base).cuda(), 13, 0.001).cpu().numpy() == synth._mutate(base, 13, 0.001)).all()
    # whole configs, as the bench uses them
    t3, s3 = synth_gpu.config3(3_000_000, 4)
    w3, ws3 = api.join_records(synth.config3(3_000_000, 4))
    assert (t3.cpu().numpy() == w3).all() and (s3 == ws3).all()
    t4, s4 = synth_gpu.config4(500_000, 4)
    w4, ws4 = api.join_records(synth.config4(500_000, 4))
    assert (t4.cpu().numpy() == w4).all() and (s4 == ws4).all()


@pytest.mark.parametrize("kind", ["poly_a", "tandem", "two_copies"])
def test_degenerate_repeats_invert(kind):
    # one (k+1)-mer repeated more than 65 536 times in a row in the sorted keys: the target-tiled in-edge join leaves the tail of
    # that run on its overflow list (mark_edges_over_kernel); long tandem repeats / whole-record copies: tie runs that K10 walks
    # for thousands of codes.  Beyond what a CPU suffix sort does in seconds, so the check is the LF inversion on the GPU.
    rng = np.random.default_rng(3)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    if kind == "poly_a":
        recs = [np.concatenate([np.full(90_000, ord("A"), np.uint8), acgt[rng.integers(0, 4, 3000)]]),
                np.concatenate([acgt[rng.integers(0, 4, 2000)], np.full(70_000, ord("A"), np.uint8), acgt[rng.integers(0, 4, 50)]])]
    elif kind == "tandem":
        unit = acgt[rng.integers(0, 4, 37)]
        recs = [np.concatenate([np.tile(unit, 4000), acgt[rng.integers(0, 4, 500)]]), np.tile(unit, 1500)]
    else:
        a = acgt[rng.integers(0, 4, 120_000)]
        recs = [a, a.copy(), a[:60_000].copy()]
    text, seps = api.join_records(recs)
    with api.BwtBuilder() as b:
        b.set_records(recs)
        b.build()
        assert b.verify(text)[0] == 0
