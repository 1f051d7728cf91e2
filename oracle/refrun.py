"""ORACLE / TEST INFRASTRUCTURE ONLY -- runs the compiled reference (oracle/_ref/deBWT).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It never reads /root/reference: it only executes the binary that
`make -C oracle ref` produced from the reference sources (reference CLI: src/main.c:25-53).

The reference writes its temp files into dirname(argv[0]) and is not re-entrant
(SURVEY.md appendix C), so every run gets a private scratch copy of the *built* artefacts
(binary + stand-in script), never of reference sources.
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import tempfile
import time
from dataclasses import dataclass, field

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return (os.path.isfile(os.path.join(REF_DIR, "deBWT"))
            and os.path.isfile(os.path.join(REF_DIR, "jf", "bin", "jellyfish"))
            and os.path.isfile(os.path.join(REF_DIR, "src", "kmercounting.sh")))


@dataclass
class RefResult:
    bwt: bytes
    sharp: bytes
    dollar: bytes
    wall_s: float                  # whole process, stand-in included
    standin_s: float               # Jellyfish stand-in (count + dump)
    proper_s: float                # "deBWT proper": wall_s - standin_s
    threads: int
    log: str = field(repr=False, default="")


def run_reference(fasta_path: str, threads: int = 1, k: int = 32, timeout: float | None = None,
                  scratch_root: str | None = None) -> RefResult:
    """Run the reference on `fasta_path`; -t 1 is the canonical (bug-free) setting, SURVEY.md §0."""
    if not available():
        raise RuntimeError("oracle/_ref is not built (run `make -C oracle ref` where /root/reference exists)")
    scratch = tempfile.mkdtemp(prefix="debwt_ref_", dir=scratch_root)
    try:
        os.makedirs(os.path.join(scratch, "src"))
        shutil.copy2(os.path.join(REF_DIR, "deBWT"), os.path.join(scratch, "deBWT"))
        shutil.copy2(os.path.join(REF_DIR, "src", "kmercounting.sh"), os.path.join(scratch, "src", "kmercounting.sh"))
        out = os.path.join(scratch, "result.bwt")
        cmd = [os.path.join(scratch, "deBWT"), "-o", out, "-t", str(threads), "-k", str(k),
               "-j", os.path.join(REF_DIR, "jf"), os.path.abspath(fasta_path)]
        t0 = time.perf_counter()
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout, cwd=scratch)
        wall = time.perf_counter() - t0
        log = p.stdout.decode(errors="replace")
        if p.returncode != 0 or not os.path.isfile(out):
            raise RuntimeError(f"reference failed (exit {p.returncode}):\n{log[-2000:]}")
        m = re.search(r"kmercounting-standin t0=([0-9.]+) t1=([0-9.]+)", log)
        standin = (float(m.group(2)) - float(m.group(1))) if m else 0.0
        with open(out, "rb") as f:
            bwt = f.read()
        with open(out + ".#", "rb") as f:
            sharp = f.read()
        with open(out + ".$", "rb") as f:
            dollar = f.read()
        return RefResult(bwt, sharp, dollar, wall, standin, wall - standin, threads, log)
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
