#!/bin/bash
# Regenerates the measured material of profiles/ on a GPU box:  gpurun --timeout 3000 -- 'bash scripts/make_profiles.sh'
# (then summarise here with scripts/ncu_summary.py / scripts/ncu_lines.py and copy what is to be judged into profiles/).
set -u
OUT=gpurun_out/r01
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"
for c in c2 c1 c5 c3 c4; do
  python bench.py --config $c --steps 5 --warmup 3 2> $OUT/bench_$c.err | tail -1 > $OUT/bench_$c.json; echo "bench $c rc=$?"
done
python bench.py --replicas 1 --steps 5 --warmup 3 --no-cpu 2> $OUT/bench_c2_r1.err | tail -1 > $OUT/bench_c2_r1.json
python bench.py --replicas 8 --steps 5 --warmup 3 --no-cpu 2> $OUT/bench_c2_r8.err | tail -1 > $OUT/bench_c2_r8.json
python bench.py --impl reference --steps 2 --warmup 1 2> $OUT/bench_ref.err | tail -1 > $OUT/bench_ref.json
python scripts/latency_table.py 40 > $OUT/latency_40.txt 2>&1
# ncu: only after the same commands exited 0 above
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --no-cpu --steps 2 --warmup 3 > $OUT/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_chains --launch-skip 4 -c 1 -f -o $OUT/k2_chains python bench.py --no-cpu --steps 2 --warmup 3 > $OUT/ncu_k2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_spec --launch-skip 4 -c 1 -f -o $OUT/k2_spec python bench.py --no-cpu --replicas 1 --steps 2 --warmup 3 > $OUT/ncu_spec.log 2>&1
ls -la $OUT
