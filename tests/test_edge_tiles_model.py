"""CPU model of the target-tiled in-edge join (debwt_b200/csrc/stages.cu: edge_bounds_kernel, mark_edges_kernel).

Source (k+1)-mer W = c.X marks the group of the k-mer X = the last 31 bases of W; its query is q = W << 2 (X followed by A) and
its target is lower_bound(keys, q).  The device serves every tile of ME_TILE consecutive target keys with four blocks, one per
first base c, and finds the run of the c block whose targets fall into tile t from the two keys at the tile's edges:

    bounds(c, t) = cbeg(c)                                              if t == 0
                   cend(c)                                              if t * ME_TILE >= n
                   lower_bound(keys, c << 62 | ((keys[a - 1] >> 2) + 1))   with a = t * ME_TILE   (cend(c) if the +1 overflows 62 bits)

This file checks on numpy that these runs partition every c block and that every source's target lies in its tile, for random
keys, duplicate-heavy keys and the all-T corner (the reference's in-edge lists: src/getKmer.c:62-109, src/INandOut.c:282-343)."""
import numpy as np
import pytest

U = np.uint64
TILE = 64          # device: 8192
MASK62 = U((1 << 62) - 1)


def bounds(keys, c, t):
    n = keys.size
    cbeg = int(np.searchsorted(keys, U(c) << U(62), side="left"))
    cend = n if c == 3 else int(np.searchsorted(keys, U(c + 1) << U(62), side="left"))
    a = t * TILE
    if a == 0:
        return cbeg
    if a >= n:
        return cend
    low = (int(keys[a - 1]) >> 2) + 1
    if low >> 62:
        return cend
    return int(np.searchsorted(keys, U((c << 62) | low), side="left"))


@pytest.mark.parametrize("kind", ["random", "duplicates", "low_entropy", "poly_t"])
def test_tile_runs_partition_the_sources_and_hold_their_targets(kind):
    rng = np.random.default_rng(11)
    n = 5000
    if kind == "random":
        keys = rng.integers(0, 1 << 64, size=n, dtype=np.uint64)
    elif kind == "duplicates":
        keys = rng.integers(0, 1 << 64, size=40, dtype=np.uint64)[rng.integers(0, 40, size=n)]
    elif kind == "low_entropy":                      # A / T only: long shared prefixes, queries that tie with keys
        bits = rng.integers(0, 2, size=(n, 32), dtype=np.uint64) * U(3)
        keys = np.zeros(n, dtype=np.uint64)
        for j in range(32):
            keys = (keys << U(2)) | bits[:, j]
        keys[: n // 10] = keys[0]
    else:
        keys = np.full(n, ~U(0), dtype=np.uint64)
        keys[: n // 2] = rng.integers(0, 1 << 64, size=n // 2, dtype=np.uint64)
    keys = np.sort(keys)
    tiles = -(-n // TILE)
    for c in range(4):
        b = [bounds(keys, c, t) for t in range(tiles + 1)]
        cbeg = int(np.searchsorted(keys, U(c) << U(62), side="left"))
        cend = n if c == 3 else int(np.searchsorted(keys, U(c + 1) << U(62), side="left"))
        assert b[0] == cbeg and b[-1] == cend
        assert all(x <= y for x, y in zip(b, b[1:]))                        # runs are disjoint, in order, and cover the block
        for t in range(tiles):
            src = keys[b[t]:b[t + 1]]
            if src.size == 0:
                continue
            q = (src & MASK62) << U(2)                                      # W << 2
            target = np.searchsorted(keys, q, side="left")
            lo, hi = t * TILE, min(n, (t + 1) * TILE)
            last = t == tiles - 1
            assert (target >= lo).all()
            assert (target < hi).all() or (last and (target <= n).all())    # a query beyond every key belongs to the last tile
