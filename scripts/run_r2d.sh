mkdir -p gpurun_out/r2d
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k2_pool -c 1 -o /tmp/pool16 python bench.py --config c1 --steps 3 --warmup 3 --no-cpu --pool 16 > gpurun_out/r2d/ncu_pool.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k2_chains -c 1 -o /tmp/classic python bench.py --config c1 --steps 3 --warmup 3 --no-cpu --pool 0 > gpurun_out/r2d/ncu_classic.log 2>&1
for K in pool16 classic; do
  ncu -i /tmp/$K.ncu-rep --page raw --csv > gpurun_out/r2d/$K.raw.csv 2>/dev/null
  ncu -i /tmp/$K.ncu-rep --page source --csv --print-source sass > gpurun_out/r2d/$K.sass.csv 2>/dev/null
done
ls -la gpurun_out/r2d
