#!/bin/bash
# ncu of k2_cells on C4 (one launch of 148 chains x 2000 steps): raw metrics + per-source-line stall samples as CSV (the .ncu-rep stays on the box)
set -u
OUT=gpurun_out/c4
mkdir -p $OUT
CMD="python bench.py --config c4 --steps-per-set 2000 --no-cpu --no-configs --no-literal --steps 1 --warmup 3"
$CMD 2> $OUT/bench.err | grep '^{' | tail -1 > $OUT/bench_line.json; echo "bench rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k2_cells --launch-skip 3 -c 1 -f -o /tmp/k2_cells $CMD > $OUT/ncu.log 2>&1
ncu -i /tmp/k2_cells.ncu-rep --page raw --csv > $OUT/k2_cells_raw.csv 2> /dev/null
ncu -i /tmp/k2_cells.ncu-rep --page source --csv > $OUT/k2_cells_source.csv 2> /dev/null
ls -la $OUT
