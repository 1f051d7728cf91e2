"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the
oracle on the same inputs, against the reference's golden vectors, and -- at full size -- through
size-independent properties (LF-walk inversion, sortedness, multiset preservation)."""
import random

import numpy as np
import pytest

from debwt_b200 import api, binding, synth
from debwt_b200.binding import DebwtError
from oracle import coracle, stages as st
from tests.util import as_bytes_records, golden, seeded_records, sha

pytestmark = pytest.mark.gpu
G = golden()


def rnd(rng, n, alpha="ACGT"):
    return "".join(rng.choice(alpha) for _ in range(n))


# ---- K1 / K2 ---------------------------------------------------------------------------------
@pytest.mark.parametrize("lens", [[33], [64], [95, 33, 40], [1000, 37, 64, 129], [100_003]])
def test_pack_and_extract_match_oracle(lens):
    rng = random.Random(sum(lens))
    recs = [rnd(rng, n) for n in lens]
    if len(recs) > 1:
        recs[1] = recs[1].lower()
    text, seps = api.join_records(recs)
    sym, oseps = st.text_from_records(recs)
    assert (api.k_pack(text) == coracle.pack_text(sym)).all()
    assert (api.k_extract(text, seps) == coracle.extract_keys(sym)).all()


def test_pack_rejects_non_acgt():
    text, seps = api.join_records(["ACGTN" * 10])
    with pytest.raises(DebwtError):
        api.k_pack(text)


# ---- K3 --------------------------------------------------------------------------------------
TMA_CFGS = [32, 33]       # radix_sort_tma.cu: bulk-store write-out (parity chain), two tile shapes


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9] + TMA_CFGS)
@pytest.mark.parametrize("n", [0, 1, 2, 255, 3839, 3840, 3841, 4095, 4096, 4097, 100_000, 1_000_003])
def test_radix_sort_random(cfg, n):
    rng = np.random.default_rng(n + cfg)
    keys = rng.integers(0, 2**64, size=n, dtype=np.uint64)
    out, _ = api.k_radix_sort(keys, cfg=cfg)
    assert (out == np.sort(keys)).all()


@pytest.mark.parametrize("cfg", [0, 8, 32, 33])
@pytest.mark.parametrize("kind", ["all_equal", "few_distinct", "sorted", "reversed", "low_bits_only", "high_bits_only", "dup_heavy"])
def test_radix_sort_structured(kind, cfg):
    n = 300_001
    rng = np.random.default_rng(7)
    if kind == "all_equal":
        keys = np.full(n, 0xDEADBEEF12345678, dtype=np.uint64)
    elif kind == "few_distinct":
        keys = rng.integers(0, 2**64, size=5, dtype=np.uint64)[rng.integers(0, 5, size=n)]
    elif kind == "sorted":
        keys = np.sort(rng.integers(0, 2**64, size=n, dtype=np.uint64))
    elif kind == "reversed":
        keys = np.sort(rng.integers(0, 2**64, size=n, dtype=np.uint64))[::-1].copy()
    elif kind == "low_bits_only":
        keys = rng.integers(0, 2**16, size=n, dtype=np.uint64)
    elif kind == "high_bits_only":
        keys = rng.integers(0, 2**16, size=n, dtype=np.uint64) << np.uint64(48)
    else:
        keys = rng.integers(0, 2**64, size=n // 100, dtype=np.uint64)[rng.integers(0, n // 100, size=n)]
    out, _ = api.k_radix_sort(keys, cfg=cfg)
    assert (out == np.sort(keys)).all()


# ---- K4 / K5-K7 ------------------------------------------------------------------------------
def test_rle_matches_oracle():
    recs = as_bytes_records(seeded_records("c4_like_5x100k"))
    sym, _ = st.text_from_records(recs)
    sk = coracle.sort_keys(coracle.extract_keys(sym))
    km, ct = api.k_rle(sk)
    okm, oct_ = coracle.rle(sk)
    assert (km == okm).all() and (ct == oct_).all()
    km1, ct1 = api.k_rle(sk[:1])
    assert km1.size == 1 and ct1[0] == 1


@pytest.mark.parametrize("name", ["haplotypes_6x1500", "pathological", "planted_repeats", "survey_golden"])
def test_group_masks_match_stage_restatement(name):
    recs = G["small"][name]["records"]
    text, seps = api.join_records(recs)
    sym, oseps = st.text_from_records(recs)
    sk = st.sort_keys(st.extract_keys(sym, oseps))
    want, _ = st.group_masks(sk, sym, oseps)
    got = api.k_group_masks(text, seps)
    assert (got == want).all()


# ---- whole path ---------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(G["small"]))
def test_bwt_matches_reference_golden_small(name):
    case = G["small"][name]
    w, s, d = api.build_bwt(case["records"])
    assert w.tobytes().hex() == case["bwt"]
    assert s.tobytes().hex() == case["sharp"]
    assert d.tobytes().hex() == case["dollar"]


@pytest.mark.parametrize("name", sorted(G["seeded"]))
def test_bwt_matches_reference_golden_seeded(name):
    case = G["seeded"][name]
    w, s, d = api.build_bwt(seeded_records(name))
    assert sha(w.tobytes()) == case["bwt_sha256"]
    assert s.tobytes().hex() == case["sharp"]
    assert d.tobytes().hex() == case["dollar"]


def test_bwt_reference_crash_case():
    for name, case in G.get("reference_crashes", {}).items():
        sym, _ = st.text_from_records(case["records"])
        want = coracle.bwt(sym)
        got = api.build_bwt(case["records"])
        assert all((a == b).all() for a, b in zip(want, got)), name


def test_bwt_random_small_inputs_vs_oracle():
    rng = random.Random(11)
    for it in range(60):
        mode = it % 5
        if mode == 0:
            recs = [rnd(rng, rng.randint(33, 300)) for _ in range(rng.randint(1, 5))]
        elif mode == 1:
            base = rnd(rng, rng.randint(40, 400))
            recs = []
            for _ in range(rng.randint(2, 6)):
                r = list(base)
                for _ in range(rng.randint(0, 4)):
                    r[rng.randrange(len(r))] = rng.choice("ACGT")
                recs.append("".join(r))
        elif mode == 2:
            recs = [rnd(rng, rng.randint(33, 500), "AC") for _ in range(rng.randint(1, 3))]
        elif mode == 3:
            r = rnd(rng, rng.randint(33, 90))
            recs = [r, r, rnd(rng, 35) + r, r]
        else:
            recs = ["A" * rng.randint(33, 200), "T" * rng.randint(33, 80), "AC" * rng.randint(17, 90),
                    "ACG" * rng.randint(11, 50), "A" * rng.randint(33, 200)]
        sym, _ = st.text_from_records(recs)
        want = coracle.bwt(sym)
        got = api.build_bwt(recs)
        assert all((a == b).all() for a, b in zip(want, got)), (it, recs)


def test_bwt_large_segments_vs_oracle():
    # one element copied many times with point mutations: blue segments well beyond one warp / one
    # shared-memory tile, deep branch-code comparisons
    rng = np.random.default_rng(5)
    master = synth.random_bases(77, 3000)
    recs = []
    for r in range(3):
        parts = []
        for c in range(900):
            el = master.copy()
            idx = rng.integers(0, el.size, size=30)
            el[idx] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=30)]
            parts.append(el)
        recs.append(np.concatenate(parts))
    sym, _ = st.text_from_records(as_bytes_records(recs))
    want = coracle.bwt(sym)
    with api.BwtBuilder() as b:
        b.set_records(recs)
        b.build()
        got = b.result()
        stats = b.stats()
    assert stats["n_blue"] > 100_000
    assert all((a == b_).all() for a, b_ in zip(want, got))


@pytest.mark.parametrize("n_rec", [40, 513, 700])
def test_bwt_many_short_records(n_rec):
    # 32 R sentinel-window suffixes: device all-pairs ranking up to 16384, host sort beyond
    rng = random.Random(n_rec)
    base = rnd(rng, 60)
    recs = []
    for i in range(n_rec):
        if i % 3 == 0:
            recs.append(rnd(rng, rng.randint(33, 70)))
        else:
            r = list(base[:rng.randint(33, 60)])
            r[rng.randrange(len(r))] = rng.choice("ACGT")
            recs.append("".join(r))
    sym, _ = st.text_from_records(recs)
    want = coracle.bwt(sym)
    got = api.build_bwt(recs)
    assert all((a == b).all() for a, b in zip(want, got))


@pytest.mark.parametrize("copies,el_len,muts", [(6000, 150, 4), (20000, 90, 3), (30000, 64, 1), (9000, 400, 2)])
def test_bwt_huge_segments_vs_oracle(copies, el_len, muts):
    # segments beyond one shared-memory block (4096 entries): sample-sort split (tie-heavy with 1-2 mutations per
    # copy) and, for items with a separator code inside a word, the chunked network over HBM
    rng = np.random.default_rng(copies)
    master = synth.random_bases(99, el_len)
    parts = []
    for c in range(copies):
        el = master.copy()
        idx = rng.integers(0, el.size, size=muts)
        el[idx] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=muts)]
        parts.append(el)
    recs = [np.concatenate(parts[:copies // 2]), np.concatenate(parts[copies // 2:])]
    sym, _ = st.text_from_records(as_bytes_records(recs))
    want = coracle.bwt(sym)
    with api.BwtBuilder() as b:
        b.set_records(recs)
        b.build()
        got = b.result()
    assert all((a == b_).all() for a, b_ in zip(want, got))


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("name", ["golden_small", "large_segments", "many_records", "c4_like"])
def test_blue_grouping_modes(name, mode):
    # K9's two ways of bringing the blue entries into their segments (per-segment cursors / dense append + radix sort on
    # the branch id; debwt_set_blue_grouping) give the oracle's BWT
    if name == "golden_small":
        cases = [G["small"][k]["records"] for k in sorted(G["small"])]
    elif name == "large_segments":
        rng = np.random.default_rng(15)
        master = synth.random_bases(78, 2500)
        parts = []
        for c in range(700):
            el = master.copy()
            idx = rng.integers(0, el.size, size=20)
            el[idx] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=20)]
            parts.append(el)
        cases = [[np.concatenate(parts[:300]), np.concatenate(parts[300:])]]
    elif name == "many_records":
        rng = random.Random(99)
        base = rnd(rng, 80)
        cases = [[base[:rng.randint(33, 80)] if i % 2 else rnd(rng, rng.randint(33, 90)) for i in range(300)]]
    else:
        base = synth.config2(200_000, 4_000, 3, 0.1)[0]
        cases = [[base] + [synth._mutate(base, 5 + i, 0.001) for i in range(4)]]
    for recs in cases:
        sym, _ = st.text_from_records(as_bytes_records(recs) if not isinstance(recs[0], str) else recs)
        want = coracle.bwt(sym)
        with api.BwtBuilder(blue_grouping=mode) as b:
            b.set_records(recs)
            b.build()
            got = b.result()
        assert all((a == b_).all() for a, b_ in zip(want, got))


def test_bwt_errors():
    with pytest.raises(DebwtError):
        api.build_bwt(["ACGT" * 8])                 # 32 bp: "Length <= 32!" (src/collect#$.c:41-45)
    with pytest.raises(DebwtError):
        api.build_bwt(["ACGTN" * 20])
    with pytest.raises(DebwtError):
        api.build_bwt(["ACGT" * 10 + "#" + "ACGT" * 10])      # separators are reserved
    with pytest.raises(DebwtError):
        api.build_bwt([])
    with pytest.raises(DebwtError):
        with api.BwtBuilder() as b:
            b.set_records(["ACGT" * 20])
            b.build(k=11)


def test_k_independent_and_builder_reuse():
    recs = G["small"]["haplotypes_6x1500"]["records"]
    with api.BwtBuilder() as b:
        outs = []
        for k in (32, 16, 12):
            b.set_records(recs)
            b.build(k)
            outs.append(b.result())
    for o in outs[1:]:
        assert all((x == y).all() for x, y in zip(outs[0], o))


def test_config1_full_size_vs_oracle():
    recs = synth.config1()                                  # 4.6 Mbp, BASELINE.json configs[0]
    sym, _ = st.text_from_records(as_bytes_records(recs))
    want = coracle.bwt(sym)
    got = api.build_bwt(recs)
    assert all((a == b).all() for a, b in zip(want, got))


@pytest.mark.slow
def test_config2_full_size_lf_inversion():
    recs = synth.config2()                                  # 100 Mbp planted repeats, configs[1]
    w, s, d = api.build_bwt(recs)
    n = recs[0].size + 1
    shifts = (2 * (31 - np.arange(32))).astype(np.uint64)
    codes = ((w[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.uint8).reshape(-1)[:n]
    codes[s.astype(np.int64)] = 4
    codes[int(d[0])] = 5
    ok, text = coracle.invert_bwt(codes)
    assert ok
    sym, _ = st.text_from_records(as_bytes_records(recs))
    assert (text == sym).all()


# ---- f1: streaming ingest (debwt_ingest_*) -------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c4_like_5x100k", "c3_like_600k_3rec"])
def test_streaming_ingest_equals_set_text(name):
    recs = as_bytes_records(seeded_records(name))
    text, seps = api.join_records(recs)
    with api.BwtBuilder() as b:
        b.set_text(text, seps)
        b.build()
        want = b.result()
        for step in (1_000_003, 4097, 33):
            if step < 1000 and text.size > 700_000:
                continue
            b.ingest((text[i:i + step] for i in range(0, text.size, step)), seps, n_hint=0 if step == 4097 else text.size)
            b.build()
            got = b.result()
            assert all((x == y).all() for x, y in zip(want, got)), step
            b.build()                                       # the packed input survives a build
            assert all((x == y).all() for x, y in zip(want, b.result()))


def test_streaming_ingest_large_and_errors():
    # larger than one 32 MB staging window, hint far too small (the packed text is moved to a larger block)
    recs = [synth.random_bases(21, 41_000_000), synth.random_bases(22, 30_000_001)]
    text, seps = api.join_records(recs)
    with api.BwtBuilder() as b:
        b.ingest([text[:40_000_000], text[40_000_000:]], seps, n_hint=1_000_000)
        b.build()
        assert b.verify(text)[0] == 0
        bad = text.copy()
        bad[12345] = ord("N")
        b.ingest([bad], seps, n_hint=text.size)
        with pytest.raises(binding.DebwtError):
            b.build()
        with pytest.raises(binding.DebwtError):
            b.ingest([text[:100]], seps)                    # separators do not match what was streamed


# ---- f3: seeded IUPAC policy (debwt_set_ambiguity_policy) ---------------------------------------------------------------
def test_ambiguity_policy_matches_restatement():
    rng = np.random.default_rng(9)
    recs = [synth.random_bases(31, 60_000), synth.random_bases(32, 45_001)]
    text, seps = api.join_records(recs)
    codes = np.frombuffer(b"NVDBHWSKMYRnvdbhwskmyr", dtype=np.uint8)
    hits = rng.choice(np.setdiff1d(np.arange(text.size), seps.astype(np.int64)), size=4000, replace=False)
    amb = text.copy()
    amb[hits] = codes[rng.integers(0, codes.size, size=hits.size)]
    with api.BwtBuilder() as b:
        b.set_text(amb, seps)
        with pytest.raises(DebwtError):                      # default policy: reject, like the reference asks of its input
            b.build()
        for seed in (0, 12345):
            resolved = st.resolve_ambiguity(amb, seed)
            assert not (resolved == amb)[hits].any() and (resolved != amb).sum() == hits.size
            b.set_ambiguity_policy(True, seed)
            b.set_text(amb, seps)
            b.build()
            got = b.result()
            assert b.verify(resolved)[0] == 0                # the BWT is the BWT of the resolved text ...
            b.ingest([amb[:33_333], amb[33_333:]], seps)     # ... also when the text is streamed in (positions carry over)
            b.build()
            assert all((x == y).all() for x, y in zip(got, b.result()))
            b.set_ambiguity_policy(False)
            b.set_text(resolved, seps)
            b.build()
            assert all((x == y).all() for x, y in zip(got, b.result()))
        b.set_ambiguity_policy(True, 1)
        junk = amb.copy()
        junk[777] = ord("!")
        b.set_text(junk, seps)
        with pytest.raises(DebwtError):
            b.build()


# ---- f4: many short records (the sentinel-window suffixes go through the bitonic network beyond 512 records) ----------
@pytest.mark.parametrize("n_rec", [600, 3000])
def test_many_records_gpu_sentinel_sort(n_rec):
    rng = np.random.default_rng(n_rec)
    base = synth.random_bases(41, 400)
    recs = []
    for i in range(n_rec):
        r = base[: int(rng.integers(40, 400))].copy() if i % 3 else synth.random_bases(1000 + i, int(rng.integers(33, 200)))
        if i % 7 == 0 and r.size > 50:
            r[int(rng.integers(0, r.size))] = ord("A")
        recs.append(r)
    text, seps = api.join_records(recs)
    with api.BwtBuilder() as b:
        b.set_text(text, seps)
        b.build()
        words, sharp, dollar = b.result()
        assert b.verify(text)[0] == 0
    sym, _ = st.text_from_records([bytes(r) for r in recs])
    ow, os_, od = coracle.bwt(sym)
    assert (words == ow).all() and (sharp == os_).all() and (dollar == od).all()
