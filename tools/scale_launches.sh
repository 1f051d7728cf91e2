#!/bin/bash
# ncu launch list (durations only) of every non-sort kernel of one full-size build: tools/scale_launches.sh c3|c4
cfg=${1:-c3}
if [ "$cfg" = c4 ]; then export DEBWT_CFG=c4; args="3e9 10"; else args="3.1e9 24"; fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:"^(?!.*(onesweep|radix_))" --csv --log-file gpurun_out/launches_full_$cfg.csv \
  python tools/c3_validate.py $args > gpurun_out/launches_full_$cfg.log 2>&1
tail -c 400 gpurun_out/launches_full_$cfg.log
