"""CPU tests of the sharded (multi-GPU) path's host logic: the orchestration of debwt_b200.dist runs on
CPU tensors with the numpy restatement of the stage kernels (tests/numpy_ops.py), on a world_size-2 and
a world_size-3 gloo group, and must reproduce the oracle / the reference's golden vectors."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from debwt_b200 import api, dist as D
from oracle import coracle, stages as st
from tests.util import golden

G = golden()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cases():
    rng = random.Random(5)

    def rnd(n, alpha="ACGT"):
        return "".join(rng.choice(alpha) for _ in range(n))
    base = rnd(400)
    hap = []
    for _ in range(4):
        x = list(base)
        for _ in range(3):
            x[rng.randrange(len(x))] = rng.choice("ACGT")
        hap.append("".join(x))
    r = rnd(70)
    return {
        "survey_golden": G["small"]["survey_golden"]["records"],
        "pathological": G["small"]["pathological"]["records"],
        "haplotypes": hap,
        "dups": [r, r, rnd(40) + r, "T" * 50, "A" * 64],
        "single": [rnd(700)],
    }


def _worker(rank, world, port, names, out):
    import torch.distributed as dist
    from tests.numpy_ops import NumpyOps
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = D.Comm()
        cases = _cases()
        for name in names:
            text, seps = api.join_records(cases[name])
            stats = {}
            w, s, d = D.build_sharded(text, seps, comm, NumpyOps(), stats)
            if rank == 0:
                sym, _ = st.text_from_records(cases[name])
                ow, os_, od = coracle.bwt(sym)
                ok = (w == ow).all() and (s == os_).all() and (d == od).all()
                out.put((name, bool(ok), stats.get("keys_local")))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_orchestration_gloo(world):
    names = list(_cases())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, names, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    got = {}
    while not q.empty():
        name, ok, kl = q.get()
        got[name] = (ok, kl)
    assert set(got) == set(names)
    for name, (ok, kl) in got.items():
        assert ok, name
        assert len(kl) == world


def test_sharded_single_rank_matches_oracle():
    from tests.numpy_ops import NumpyOps
    comm = D.Comm()
    assert comm.size == 1
    for name, recs in _cases().items():
        text, seps = api.join_records(recs)
        w, s, d = D.build_sharded(text, seps, comm, NumpyOps())
        sym, _ = st.text_from_records(recs)
        ow, os_, od = coracle.bwt(sym)
        assert (w == ow).all() and (s == os_).all() and (d == od).all(), name


def test_geometry_helpers():
    seps = np.array([40, 90, 200], dtype=np.uint64)
    # records [0,40) [41,90) [91,200): valid window starts 0..8, 41..58, 91..168
    assert D.valid_windows_before(0, seps) == 0
    assert D.valid_windows_before(9, seps) == 9
    assert D.valid_windows_before(41, seps) == 9
    assert D.valid_windows_before(60, seps) == 9 + 18
    assert D.valid_windows_before(201, seps) == 9 + 18 + 78 == 201 - 32 * 3
    wp, tot = D.slice_geometry(201, 4)
    assert tot == 4 * wp and tot >= (201 + 63) // 32 + 1
    sp = D.choose_splitters(np.arange(0, 1000, dtype=np.uint64) * np.uint64(7), 4)
    assert sp.size == 3 and (sp & np.uint64(3) == 0).all() and (np.diff(sp.astype(np.int64)) >= 0).all()
