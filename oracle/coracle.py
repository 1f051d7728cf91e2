"""ORACLE / TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/liboracle.so (oracle/bwt_oracle.c).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "bwt_oracle.c")
    if force or not os.path.isfile(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-g", "-Wall", "-shared", "-fPIC", "-o", so, src])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        vp, u64 = ctypes.c_void_p, ctypes.c_uint64
        L.oracle_bwt.argtypes = [vp, u64, vp]; L.oracle_bwt.restype = ctypes.c_int
        L.oracle_pack_bwt.argtypes = [vp, u64, vp, vp, vp]; L.oracle_pack_bwt.restype = u64
        L.oracle_pack_text.argtypes = [vp, u64, vp]; L.oracle_pack_text.restype = None
        L.oracle_extract_keys.argtypes = [vp, u64, vp]; L.oracle_extract_keys.restype = u64
        L.oracle_sort_keys.argtypes = [vp, u64]; L.oracle_sort_keys.restype = None
        L.oracle_rle.argtypes = [vp, u64, vp, vp]; L.oracle_rle.restype = u64
        L.oracle_invert_bwt.argtypes = [vp, u64, vp]; L.oracle_invert_bwt.restype = ctypes.c_int
        L.oracle_unpack_bwt.argtypes = [vp, u64, vp, u64, u64, vp]; L.oracle_unpack_bwt.restype = None
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def bwt_symbols(sym: np.ndarray) -> np.ndarray:
    sym = np.ascontiguousarray(sym, dtype=np.uint8)
    out = np.empty_like(sym)
    if lib().oracle_bwt(_p(sym), sym.size, _p(out)) != 0:
        raise MemoryError
    return out


def pack_bwt(bwt: np.ndarray):
    bwt = np.ascontiguousarray(bwt, dtype=np.uint8)
    n = bwt.size
    words = np.zeros((n + 31) // 32, dtype=np.uint64)
    sharp = np.zeros(int((bwt == 4).sum()) + 1, dtype=np.uint64)
    dollar = np.zeros(1, dtype=np.uint64)
    ns = lib().oracle_pack_bwt(_p(bwt), n, _p(words), _p(sharp), _p(dollar))
    return words, sharp[:ns].copy(), dollar


def bwt(sym: np.ndarray):
    """(words, sharp_rows, dollar_row) exactly as the reference's three output files hold them."""
    return pack_bwt(bwt_symbols(sym))


def pack_text(sym: np.ndarray) -> np.ndarray:
    sym = np.ascontiguousarray(sym, dtype=np.uint8)
    words = np.zeros((sym.size + 32 + 31) // 32, dtype=np.uint64)
    lib().oracle_pack_text(_p(sym), sym.size, _p(words))
    return words


def extract_keys(sym: np.ndarray) -> np.ndarray:
    sym = np.ascontiguousarray(sym, dtype=np.uint8)
    keys = np.empty(sym.size, dtype=np.uint64)
    m = lib().oracle_extract_keys(_p(sym), sym.size, _p(keys))
    return keys[:m].copy()


def sort_keys(keys: np.ndarray) -> np.ndarray:
    k = np.ascontiguousarray(keys, dtype=np.uint64).copy()
    lib().oracle_sort_keys(_p(k), k.size)
    return k


def rle(sorted_keys: np.ndarray):
    s = np.ascontiguousarray(sorted_keys, dtype=np.uint64)
    km = np.empty(s.size, dtype=np.uint64)
    ct = np.empty(s.size, dtype=np.uint64)
    d = lib().oracle_rle(_p(s), s.size, _p(km), _p(ct))
    return km[:d].copy(), ct[:d].copy()


def unpack_bwt(words: np.ndarray, n: int, sharp: np.ndarray, dollar: np.ndarray) -> np.ndarray:
    words = np.ascontiguousarray(words, dtype=np.uint64)
    sharp = np.ascontiguousarray(sharp, dtype=np.uint64)
    out = np.empty(n, dtype=np.uint8)
    lib().oracle_unpack_bwt(_p(words), n, _p(sharp), sharp.size, int(dollar[0]), _p(out))
    return out


def invert_bwt(bwt_sym: np.ndarray):
    b = np.ascontiguousarray(bwt_sym, dtype=np.uint8)
    t = np.zeros_like(b)
    rc = lib().oracle_invert_bwt(_p(b), b.size, _p(t))
    return rc == 0, t
