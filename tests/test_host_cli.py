"""Host side of the drop-in boundary: the FASTA/FASTQ reader (host/fastx.h) on CPU, and -- on a GPU --
the deBWT executable with the reference's command line writing the reference's three files."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from tests.util import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = golden()


def fnv(b):
    h = 1469598103934665603
    for c in b:
        h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.fixture(scope="module")
def dumper(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("fx") / "fastx_dump")
    subprocess.check_call(["gcc", "-O2", "-o", exe, os.path.join(ROOT, "host", "fastx_dump.c"), "-lz"])
    return exe


def run_dump(exe, path):
    """records as read by fastx_read; the streaming reader (fastx_stream, what host/deBWT uses) must agree"""
    out = subprocess.run([exe, path], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    out_s = subprocess.run([exe, "-s", path], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    assert out == out_s
    recs = [tuple(int(x) for x in l.split()) for l in out[:-1]]
    return recs, out[-1]


def test_fastx_reads_fasta_fastq_gzip_crlf(dumper, tmp_path):
    seqs = [b"ACGT" * 30, b"T" * 75, b"GATTACA" * 11, b"ac" * 40]
    want = [(len(s), fnv(s)) for s in seqs]
    fa = tmp_path / "a.fa"
    with open(fa, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">r%d some description\n" % i)
            for j in range(0, len(s), 60):
                f.write(s[j:j + 60] + b"\n")
    assert run_dump(dumper, str(fa))[0] == want
    crlf = tmp_path / "b.fa"
    with open(crlf, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">r%d\r\n" % i + s[:50] + b"\r\n" + s[50:] + b"\r\n")
    assert run_dump(dumper, str(crlf))[0] == want
    gz = tmp_path / "c.fa.gz"
    with gzip.open(gz, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">r%d\n" % i + s + b"\n")
    assert run_dump(dumper, str(gz))[0] == want
    fq = tmp_path / "d.fq"
    with open(fq, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b"@r%d\n" % i + s + b"\n+\n" + b"@" * len(s) + b"\n")     # '@' quality values must not start a record
    assert run_dump(dumper, str(fq))[0] == want
    nonl = tmp_path / "e.fa"
    with open(nonl, "wb") as f:
        f.write(b">only\n" + seqs[0])                                          # no trailing newline
    assert run_dump(dumper, str(nonl))[0] == want[:1]
    big = tmp_path / "f.fa"
    s = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[np.random.default_rng(1).integers(0, 4, 3_000_000)])
    with open(big, "wb") as f:
        f.write(b">big\n")
        for j in range(0, len(s), 80):
            f.write(s[j:j + 80] + b"\n")
    assert run_dump(dumper, str(big))[0] == [(len(s), fnv(s))]


def test_cli_usage_and_errors(tmp_path):
    exe = os.path.join(ROOT, "host", "deBWT")
    if not os.path.isfile(exe):
        pytest.skip("host/deBWT not built")
    assert subprocess.run([exe]).returncode == 1
    fa = tmp_path / "s.fa"
    fa.write_text(">a\n" + "ACGT" * 8 + "\n")
    r = subprocess.run([exe, "-o", str(tmp_path / "o"), str(fa)], capture_output=True, text=True)
    assert r.returncode == 1 and "Length <= 32!" in r.stderr                    # src/collect#$.c:41-45
    r = subprocess.run([exe, "-o", str(tmp_path / "o"), "-k", "40", str(fa)], capture_output=True, text=True)
    assert r.returncode == 1 and "12 to 32" in r.stderr                         # src/main.c:45-46
    r = subprocess.run([exe, "-o", "/nonexistent_dir/x", str(fa)], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot create" in r.stderr                    # src/main.c:55-58


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["survey_golden", "pathological", "haplotypes_6x1500"])
def test_cli_writes_reference_files(name, tmp_path):
    exe = os.path.join(ROOT, "host", "deBWT")
    case = G["small"][name]
    fa = tmp_path / "in.fa"
    with open(fa, "w") as f:
        for i, r in enumerate(case["records"]):
            f.write(f">r{i}\n")
            for j in range(0, len(r), 70):
                f.write(r[j:j + 70] + "\n")
    out = tmp_path / "out.bwt"
    r = subprocess.run([exe, "-o", str(out), "-t", "8", "-k", "32", "-j", "/ignored", str(fa)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert out.read_bytes().hex() == case["bwt"]
    assert (tmp_path / "out.bwt.#").read_bytes().hex() == case["sharp"]
    assert (tmp_path / "out.bwt.$").read_bytes().hex() == case["dollar"]


@pytest.mark.gpu
def test_cli_multi_gpu_list(tmp_path):
    """deBWT -g a,b,...: the sharded build behind the C ABI (debwt_build_multi); two ranks share cuda:0 when the box has one GPU"""
    import torch
    exe = os.path.join(ROOT, "host", "deBWT")
    case = G["small"]["haplotypes_6x1500"]
    fa = tmp_path / "in.fa"
    with open(fa, "w") as f:
        for i, r in enumerate(case["records"]):
            f.write(f">r{i}\n{r}\n")
    out1, out2 = tmp_path / "one.bwt", tmp_path / "multi.bwt"
    assert subprocess.run([exe, "-o", str(out1), str(fa)], capture_output=True, text=True).returncode == 0
    devs = ",".join(str(d) for d in range(min(torch.cuda.device_count(), 4))) if torch.cuda.device_count() > 1 else "0,0,0"
    r = subprocess.run([exe, "-o", str(out2), "-g", devs, str(fa)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for ext in ("", ".#", ".$"):
        assert open(str(out1) + ext, "rb").read() == open(str(out2) + ext, "rb").read(), ext
    assert open(out1, "rb").read().hex() == case["bwt"]
