#!/bin/bash
# File-to-file timing of the host CLI (SURVEY 8d "also report file-to-file"): tools/file_to_file.sh
# writes the C2 workload as FASTA (plain and gzip) and runs host/deBWT on it twice (second run: warm page cache).
set -e
python - <<'PY'
from debwt_b200 import synth
synth.write_fasta(synth.config2(), "/tmp/c2.fa")
PY
gzip -1 -k -f /tmp/c2.fa
for f in /tmp/c2.fa /tmp/c2.fa /tmp/c2.fa.gz; do
  t0=$(date +%s%N); ./host/deBWT -o /tmp/c2.out -k 32 $f; t1=$(date +%s%N)
  echo "wall $(( (t1 - t0) / 1000000 )) ms ($f)"
done
sha256sum /tmp/c2.out | cut -c1-16
