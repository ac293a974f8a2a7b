mkdir -p gpurun_out/r2e
for P in 12 16 0; do timeout 300 python bench.py --config c1 --steps 3 --warmup 3 --no-cpu --pool $P > gpurun_out/r2e/bench_c1_pool$P.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2e/bench_c1_pool$P.json')); print('c1 pool $P', d['value'], d['roofline']['kernel'], d['roofline']['frac'])"; done
for P in 12 16 0; do timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-configs --no-literal --pool $P > gpurun_out/r2e/bench_pool$P.json 2> gpurun_out/r2e/bench_pool$P.err; python -c "
import json; d=json.load(open('gpurun_out/r2e/bench_pool$P.json')); print('c2 pool $P', d['value'], d['roofline']['kernel'], d['roofline']['frac'], d['e2e']['value'])"; done
