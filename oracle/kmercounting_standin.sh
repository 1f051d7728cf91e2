#!/bin/sh
# ORACLE / TEST INFRASTRUCTURE ONLY.
# Installed by oracle/Makefile as oracle/_ref/src/kmercounting.sh, the path the reference binary
# runs through system() (reference src/main.c:70).  Same five positional arguments as the
# reference's script (SOURCE THREADS BIN JROOT K); it drives the Jellyfish stand-in
# (oracle/jellyfish_standin.c) through the two-command count/dump sequence the reference expects
# and leaves the text dump in $BIN/out for mySort (reference src/mySort.c:48).
set -e
SOURCE=$1
THREADS=$2
BIN=$3
JROOT=$4
K=$5
T0=$(date +%s.%N)
"$JROOT/bin/jellyfish" count -m "$K" -o "$BIN/output" -c 40 -s 4G -t "$THREADS" "$SOURCE"
"$JROOT/bin/jellyfish" dump -c -t -o "$BIN/out" "$BIN/output"
T1=$(date +%s.%N)
echo "kmercounting-standin t0=$T0 t1=$T1"
rm -f "$BIN/output"
