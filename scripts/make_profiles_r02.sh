#!/bin/bash
# Regenerates the measured material of profiles/r02_* on a GPU box:  gpurun --timeout 3000 -- 'bash scripts/make_profiles_r02.sh'
# Only text summaries come back (the .ncu-rep files stay on the box: gpurun_out/ is limited to 64 MiB).
set -u
OUT=gpurun_out/r02
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 5 --warmup 3 2> $OUT/bench.err | grep '^{' | tail -1 > $OUT/bench_line.json; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 3 2> $OUT/bench_ref.err | grep '^{' | tail -1 > $OUT/bench_reference_arm.json
CMD="python bench.py --no-cpu --no-configs --steps 2 --warmup 3"
# ncu: only after the same command exited 0 above
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_chains --launch-skip 4 -c 1 -f -o /tmp/k2_chains $CMD > $OUT/ncu_k2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_spec --launch-skip 4 -c 1 -f -o /tmp/k2_spec $CMD > $OUT/ncu_spec.log 2>&1
python scripts/ncu_summary.py full /tmp/k2_chains.ncu-rep $OUT/k2_chains_full.json 51200000 > /dev/null 2>&1
python scripts/ncu_summary.py full /tmp/k2_spec.ncu-rep $OUT/k2_spec_full.json 400000 > /dev/null 2>&1
python scripts/ncu_summary.py launches $OUT/launches.csv $OUT/launches_summary.md > /dev/null 2>&1
ls -la $OUT
