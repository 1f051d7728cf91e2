"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes

import numpy as np
import pytest

from debwt_b200 import api, binding


def test_library_exports_every_declared_symbol():
    L = binding.lib()
    names = binding.declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), n
    assert set(binding._SIGS) == set(names)


def test_library_exports_every_device_stage_symbol():
    L = binding.lib()
    names = binding.declared_symbols(binding.DEV_HEADER)
    assert len(names) >= 26
    for n in names:
        assert hasattr(L, n), n


def test_stats_struct_matches_header_layout(tmp_path):
    """the ctypes mirror has the size and field offsets a C compiler gives include/debwt_b200.h"""
    import os
    import subprocess
    src = tmp_path / "sz.c"
    fields = [n for n, _ in binding.Stats._fields_]
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "debwt_b200.h"\nint main(void){printf("%zu", sizeof(debwt_stats));'
                   + "".join('printf(" %%zu", offsetof(debwt_stats, %s));' % f for f in fields) + "return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.dirname(binding.HEADER), "-o", str(exe), str(src)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out[0] == ctypes.sizeof(binding.Stats)
    assert out[1:] == [getattr(binding.Stats, f).offset for f in fields]


def test_no_cpu_fallback_without_device():
    if binding.lib().debwt_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(binding.DebwtError) as e:
        api.build_bwt(["ACGT" * 20])
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(binding.DebwtError):
        api.k_radix_sort(np.arange(10, dtype=np.uint64))


def test_join_records_layout():
    text, seps = api.join_records(["ACGT" * 10, b"TTTT" * 9])
    assert bytes(text[40:41]) == b"#" and bytes(text[-1:]) == b"$"
    assert list(seps) == [40, 77] and text.size == 78
