"""CPU model of the K10 work-list refinement (debwt_b200/csrc/bluesort.cu: split_kernel, refine_kernel).

The device sorts the blue entries of one multi-in k-mer by the branch-code string that starts at their spIndex
(reference: src/sortBlue.c:76-280, cmpSP :109-173).  It never compares strings pairwise on large segments: it refines
runs of entries one code word at a time -- sample-sort split of everything above the tiny class with equality buckets (a
window with a separator code is bucketed by an order-preserving word, order_word()), a three-way peel around a dominant
word, a packed network on the first 27 codes of plain words, direct comparisons only for short runs, a comparator fallback
when a small item holds a separator code.  This file restates that control flow in plain Python with scaled-down
thresholds and checks it against `sorted()` on the code strings, so the host-visible logic (what becomes an item, at
which depth, when a run is finished) is pinned on CPU; the CUDA kernels are checked against the oracle in
tests/test_gpu_parity.py.
"""
import random

import pytest

SPLIT_ABOVE = 16    # device: 512    items above this are cut by the split first
SHORT = 4           # device: 32     runs up to this size are ranked by direct comparisons
ADV = 27            # device: 27     codes the packed network of refine_kernel<.., 512> compares per step
SPLIT_TARGET = 8    # device: 256
MAX_BUCKETS = 8     # device: 256
OVERSAMPLE = 2      # device: 8
MAX_PEEL = 3        # device: 64
W = 32              # codes per word


class Codes:
    """code string with separator codes: symbols 0..3, 4 = '#', 5 = '$' (unique, last)"""

    def __init__(self, syms):
        self.syms = bytes(syms)
        self.n = len(syms)

    def word(self, s, depth):
        """(integer value of the 32 two-bit codes at s + depth, plain flag) like text_window32 + fetch_sep"""
        p = s + depth
        chunk = self.syms[p:p + W]
        plain = len(chunk) == W and all(c < 4 for c in chunk)
        val = 0
        for i in range(W):
            c = chunk[i] if i < len(chunk) else 3
            val = (val << 2) | min(c, 3)
        return val, plain

    def order_word(self, s, depth):
        """(word with ones from the first separator code on, has-a-separator flag): debwt::order_word"""
        p = s + depth
        chunk = self.syms[p:p + W]
        val, np_ = 0, False
        for i in range(W):
            c = chunk[i] if i < len(chunk) else 5
            if c >= 4:
                np_ = True
            val = (val << 2) | (3 if np_ else c)
        return val, np_

    def less(self, a, b):
        return self.syms[a:] < self.syms[b:]

    def key(self, e):
        return self.syms[e >> 4:]


def refine_segment(entries, codes, stats=None):
    """entries: list of (spIndex << 4) | prev of one segment; returns them in the order the device leaves them in"""
    ent = list(entries)
    stats = stats if stats is not None else {}
    items = [(0, len(ent), 0)]                  # (offset, length, depth): entries agree on their first `depth` codes
    rounds = 0
    while items:
        rounds += 1
        nxt = []
        cur = []
        for off, ln, depth in items:            # split_kernel: huge items are cut before the round's refinement
            if ln > SPLIT_ABOVE:
                _split(ent, codes, off, ln, depth, cur, nxt, stats)
            else:
                cur.append((off, ln, depth))
        for off, ln, depth in cur:
            _refine_item(ent, codes, off, ln, depth, nxt, stats)
        items = nxt
        assert rounds < 10000
    stats["rounds"] = rounds
    return ent


def _mixed(ent, off, ln):
    return len({e & 15 for e in ent[off:off + ln]}) > 1


def _full_sort(ent, codes, off, ln):
    ent[off:off + ln] = sorted(ent[off:off + ln], key=codes.key)


def _split(ent, codes, off, ln, depth, cur, nxt, stats):
    seg = ent[off:off + ln]
    words = [codes.order_word(e >> 4, depth) for e in seg]
    if not _mixed(ent, off, ln):
        return
    nb = max(2, min(MAX_BUCKETS, -(-ln // SPLIT_TARGET)))
    m = nb * OVERSAMPLE
    sample = sorted(words[t * ln // m][0] for t in range(m))
    split = [sample[t * OVERSAMPLE + OVERSAMPLE - 1] for t in range(nb - 1)]
    buckets = [[] for _ in range(2 * nb - 1)]
    for e, (w, np_) in zip(seg, words):
        lo = sum(1 for s in split if s < w)      # first splitter >= w
        # equal to a splitter: plain windows tie on 32 codes (equality bucket); one with a separator code sorts after them
        b = 2 * lo + ((2 if np_ else 1) if lo < len(split) and split[lo] == w else 0)
        buckets[b].append(e)
    if any(len(members) == ln and not b & 1 for b, members in enumerate(buckets)):
        stats["fallback"] = stats.get("fallback", 0) + 1      # no progress: separator windows that all tie -> comparator
        _full_sort(ent, codes, off, ln)
        return
    pos = off
    stats["splits"] = stats.get("splits", 0) + 1
    for b, members in enumerate(buckets):
        ent[pos:pos + len(members)] = members
        if len(members) >= 2:
            item = (pos, len(members), depth + (W if b & 1 else 0))
            (nxt if len(members) > SPLIT_ABOVE else cur).append(item)      # still long: cut again next round
        pos += len(members)


def _rank_short(ent, codes, off, ln):
    ent[off:off + ln] = sorted(ent[off:off + ln], key=codes.key)


def _refine_item(ent, codes, off, ln, depth, nxt, stats):
    lo, vlen, peels = off, ln, 0
    while True:
        if peels == MAX_PEEL:
            nxt.append((lo, vlen, depth))
            return
        if not _mixed(ent, lo, vlen):
            return                              # one prev symbol: any order gives the same BWT (src/sortBlue.c:192-219)
        words = [codes.word(e >> 4, depth) for e in ent[lo:lo + vlen]]
        if not all(p for _, p in words):
            stats["fallback"] = stats.get("fallback", 0) + 1
            _full_sort(ent, codes, lo, vlen)
            return
        vals = [w for w, _ in words]
        if vlen > SHORT:
            pivot = vals[vlen >> 1]
            lt = [e for e, w in zip(ent[lo:lo + vlen], vals) if w < pivot]
            eq = [e for e, w in zip(ent[lo:lo + vlen], vals) if w == pivot]
            gt = [e for e, w in zip(ent[lo:lo + vlen], vals) if w > pivot]
            if 2 * len(eq) >= vlen:             # dominant word: three-way peel, the equal part goes 32 codes deeper
                stats["peels"] = stats.get("peels", 0) + 1
                ent[lo:lo + vlen] = lt + eq + gt
                for o, part in ((lo, lt), (lo + len(lt) + len(eq), gt)):
                    if len(part) > SHORT:
                        nxt.append((o, len(part), depth))
                    elif len(part) >= 2:
                        _rank_short(ent, codes, o, len(part))
                lo, vlen, depth, peels = lo + len(lt), len(eq), depth + W, peels + 1
                continue
        vals = [v >> (2 * (W - ADV)) for v in vals]                  # the packed network: the first ADV codes | index
        order = sorted(range(vlen), key=lambda t: (vals[t], t))
        seg = [ent[lo + t] for t in order]
        vals = [vals[t] for t in order]
        ent[lo:lo + vlen] = seg
        h = 0
        while h < vlen:                          # runs of equal words
            t = h
            while t + 1 < vlen and vals[t + 1] == vals[h]:
                t += 1
            size = t + 1 - h
            if size >= 2 and _mixed(ent, lo + h, size):
                if size > SHORT:
                    nxt.append((lo + h, size, depth + ADV))
                else:
                    _rank_short(ent, codes, lo + h, size)
            h = t + 1
        return


def make_case(rng, n_codes, n_entries, family=None, seps=0):
    """code string + one segment of entries; `family`: (copies, length, mutations) of a planted repeat"""
    syms = [rng.randrange(4) for _ in range(n_codes)]
    starts = []
    if family:
        copies, length, muts = family
        master = [rng.randrange(4) for _ in range(length)]
        for c in range(copies):
            p = rng.randrange(0, n_codes - length - 1000)
            el = list(master)
            for _ in range(muts):
                el[rng.randrange(length)] = rng.randrange(4)
            syms[p:p + length] = el
            starts.append(p + (rng.randrange(0, 3) if muts else 1))
    for _ in range(seps):
        syms[rng.randrange(n_codes - 1)] = 4
    syms[-1] = 5
    while len(starts) < n_entries:
        starts.append(rng.randrange(n_codes - 1000))      # away from the '$' code: words near it are not plain
    starts = sorted(set(starts))
    return Codes(syms), [(s << 4) | rng.randrange(4) for s in starts]


@pytest.mark.parametrize("seed,n_codes,n_entries,family,seps", [
    (1, 4000, 40, None, 0),                       # below every threshold but the short-run one
    (2, 4000, 300, None, 0),                      # split, distinct words
    (3, 20000, 200, (180, 400, 1), 0),            # near-identical copies: dominant-word peels, deep ties
    (4, 20000, 500, (450, 150, 3), 0),            # split with heavy equality buckets
    (5, 20000, 300, (250, 300, 2), 40),           # separator codes: comparator fallback
    (6, 200000, 120, (100, 900, 0), 0),           # exact copies: ties end only at the end of the element (many peels)
])
def test_refinement_model_orders_like_string_sort(seed, n_codes, n_entries, family, seps):
    rng = random.Random(seed)
    codes, entries = make_case(rng, n_codes, n_entries, family, seps)
    rng.shuffle(entries)
    stats = {}
    got = refine_segment(entries, codes, stats)
    want = sorted(entries, key=codes.key)
    # entries of a finished run may stay in any order when their prev symbols agree: the BWT is what must match
    assert [e & 15 for e in got] == [e & 15 for e in want]
    assert sorted(got) == sorted(entries)
    if family and family[2] <= 1 and not seps:
        assert stats.get("peels", 0) > 0 or stats["rounds"] > 5      # deep ties: peeled, or cut again 32 codes deeper per round
    if n_entries > SPLIT_ABOVE:
        assert stats.get("splits", 0) > 0
    if seps:
        assert stats.get("fallback", 0) > 0


def test_order_word_is_consistent_with_the_comparator():
    """split_kernel buckets by (order_word, has-separator): whenever that pair differs, it must order two windows like cmpSP"""
    rng = random.Random(9)
    for trial in range(300):
        n = 200
        syms = [rng.randrange(2) * 3 for _ in range(n)]            # only A and T: long common prefixes, all-T tails
        for _ in range(rng.randrange(1, 12)):
            syms[rng.randrange(n - 1)] = 4
        syms[-1] = 5
        codes = Codes(syms)
        starts = rng.sample(range(n - 1), 40)
        keyed = [(codes.order_word(s, 0), s) for s in starts]
        for (ka, a) in keyed:
            for (kb, b) in keyed:
                if ka < kb:
                    assert codes.less(a, b), (trial, a, b)
