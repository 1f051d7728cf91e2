"""Deterministic synthetic genomes for the BASELINE.json configs.

splitmix64 stream, 32 bases per 64-bit output taken from the top bits down, upper case.
Bit-reproducible: numpy here, trivially restatable in C.  Not part of the device path.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
M64 = (1 << 64) - 1


def splitmix64(seed: int, count: int) -> np.ndarray:
    """`count` successive splitmix64 outputs for `seed` (vectorised: state_i = seed + (i+1)*gamma)."""
    with np.errstate(over="ignore"):
        z = (np.arange(1, count + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) + np.uint64(seed & M64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def random_codes(seed: int, n: int) -> np.ndarray:
    """n iid-uniform base codes (0..3); generated in chunks so that multi-Gbp sequences fit in memory."""
    nw = (n + 31) // 32
    out = np.empty(nw * 32, dtype=np.uint8)
    shifts = (2 * (31 - np.arange(32))).astype(np.uint64)
    gamma = np.uint64(0x9E3779B97F4A7C15)
    chunk = 1 << 22
    with np.errstate(over="ignore"):
        for lo in range(0, nw, chunk):
            hi = min(nw, lo + chunk)
            z = (np.arange(lo + 1, hi + 1, dtype=np.uint64) * gamma) + np.uint64(seed & M64)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            out[lo * 32:hi * 32] = ((z[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.uint8).reshape(-1)
    return out[:n]


def random_bases(seed: int, n: int) -> np.ndarray:
    """n iid-uniform bases as ASCII bytes."""
    return _ACGT[random_codes(seed, n)]


def _rand_ints(seed: int, count: int, bound: int) -> np.ndarray:
    return (splitmix64(seed, count) % np.uint64(bound)).astype(np.int64)


def config1(n: int = 4_600_000):
    """C1: one record, iid uniform, seed 1."""
    return [random_bases(1, n)]


def config2(n: int = 100_000_000, elem_len: int = 10_000, n_elems: int = 50, frac: float = 0.05):
    """C2: iid background (seed 2) + exact copies of `n_elems` elements (seed 3) over distinct
    elem_len-aligned slots (seed 4) covering `frac` of the length."""
    seq = random_bases(2, n)
    elems = random_bases(3, elem_len * n_elems).reshape(n_elems, elem_len)
    n_slots = n // elem_len
    n_copies = int(n * frac) // elem_len
    if n_slots == 0 or n_copies == 0:
        return [seq]
    order = np.argsort(splitmix64(4, n_slots), kind="stable")[:n_copies]      # distinct slots
    which = _rand_ints(4 + (1 << 32), n_copies, n_elems)
    for slot, e in zip(order.tolist(), which.tolist()):
        seq[slot * elem_len:(slot + 1) * elem_len] = elems[e]
    return [seq]


def _mutate(seq: np.ndarray, seed: int, rate: float) -> np.ndarray:
    """iid substitutions at `rate`, never to the same base."""
    n = seq.size
    r = splitmix64(seed, n)
    thr = np.uint64(int(rate * (1 << 53)))
    hit = (r >> np.uint64(11)) < thr
    out = seq.copy()
    idx = np.flatnonzero(hit)
    code = np.searchsorted(_ACGT, out[idx])          # ACGT is sorted in ASCII
    delta = ((r[idx] & np.uint64(0x7FF)) % np.uint64(3)).astype(np.int64) + 1
    out[idx] = _ACGT[(code + delta) & 3]
    return out


def _mutate_many(master: np.ndarray, seeds: np.ndarray, rate: float) -> np.ndarray:
    """rows i = _mutate(master, seeds[i], rate), computed for all rows at once"""
    L = master.size
    with np.errstate(over="ignore"):
        z = (np.arange(1, L + 1, dtype=np.uint64)[None, :] * np.uint64(0x9E3779B97F4A7C15)) + seeds.astype(np.uint64)[:, None]
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        r = z ^ (z >> np.uint64(31))
    thr = np.uint64(int(rate * (1 << 53)))
    hit = (r >> np.uint64(11)) < thr
    code = np.searchsorted(_ACGT, master)
    delta = ((r & np.uint64(0x7FF)) % np.uint64(3)).astype(np.int64) + 1
    out = np.where(hit, _ACGT[(code[None, :] + delta) & 3], master[None, :])
    return out.astype(np.uint8)


def _insert_family(seq: np.ndarray, seed: int, copies: int, length: int, sub_rate: float):
    if copies <= 0 or seq.size <= length:
        return
    master = random_bases(seed, length)
    offs = _rand_ints(seed + (1 << 33), copies, seq.size - length).tolist()
    batch = max(1, (1 << 24) // length)
    for lo in range(0, copies, batch):
        hi = min(copies, lo + batch)
        if sub_rate == 0.0:
            els = None
        else:
            els = _mutate_many(master, np.arange(lo, hi, dtype=np.uint64) + np.uint64(seed + (1 << 34)), sub_rate)
        for i in range(lo, hi):                      # later copies overwrite earlier ones, in order
            o = offs[i]
            seq[o:o + length] = master if els is None else els[i - lo]


def genome_like(n: int, seed0: int, scale: float = 1.0) -> np.ndarray:
    """C3-style sequence: iid background (seed0) + three repeat families (seed0+1..3):
    1e6 x 300 bp at 10 %, 1e5 x 6 kbp at 3 %, 2e3 x 10 kbp exact, counts scaled by n/3.1e9*scale."""
    seq = random_bases(seed0, n)
    f = n / 3.1e9 * scale
    _insert_family(seq, seed0 + 1, int(1_000_000 * f), 300, 0.10)
    _insert_family(seq, seed0 + 2, int(100_000 * f), 6000, 0.03)
    _insert_family(seq, seed0 + 3, int(2_000 * f), 10_000, 0.0)
    return seq


def genome_like_prefix(n: int, seed0: int, m: int, scale: float = 1.0) -> np.ndarray:
    """the first m bases of genome_like(n, seed0, scale) without generating the other n - m: the background is indexed by
    position, and only the copies of each family that reach into [0, m) are applied (in copy order, as _insert_family does).
    Used by the CPU arms of bench.py, which run the reference on a bounded prefix of a multi-Gbp workload."""
    m = min(m, n)
    seq = random_bases(seed0, m)
    f = n / 3.1e9 * scale
    for seed, copies, length, rate in ((seed0 + 1, int(1_000_000 * f), 300, 0.10), (seed0 + 2, int(100_000 * f), 6000, 0.03),
                                       (seed0 + 3, int(2_000 * f), 10_000, 0.0)):
        if copies <= 0 or n <= length:
            continue
        master = random_bases(seed, length)
        offs = _rand_ints(seed + (1 << 33), copies, n - length)
        for i in np.flatnonzero(offs < m).tolist():              # ascending copy index: later copies overwrite earlier ones
            o = int(offs[i])
            el = master if rate == 0.0 else _mutate_many(master, np.array([i + seed + (1 << 34)], dtype=np.uint64), rate)[0]
            k = min(length, m - o)
            seq[o:o + k] = el[:k]
    return seq


def config3(n: int = 3_100_000_000, n_records: int = 24):
    """C3: human-sized genome split into `n_records` records (each < 2^31 bp)."""
    seq = genome_like(n, 5)
    per = -(-n // n_records)
    return [seq[i:i + per] for i in range(0, n, per)]


def config4(base_len: int = 300_000_000, n_genomes: int = 10, rate: float = 0.001):
    """C4: base genome (seed 9) + n_genomes-1 copies with iid substitutions at `rate` (seeds 10..)."""
    base = genome_like(base_len, 9)
    return [base] + [_mutate(base, 10 + i, rate) for i in range(n_genomes - 1)]


def write_fasta(records, path: str, width: int = 80):
    with open(path, "wb") as f:
        for i, r in enumerate(records):
            f.write(b">seq%d\n" % i)
            r = np.asarray(r, dtype=np.uint8)
            nfull = r.size // width
            if nfull:
                body = np.empty((nfull, width + 1), dtype=np.uint8)
                body[:, :width] = r[:nfull * width].reshape(nfull, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if r.size % width:
                f.write(r[nfull * width:].tobytes() + b"\n")
