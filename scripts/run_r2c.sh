mkdir -p gpurun_out/r2c
timeout 900 python -m pytest tests/test_gpu_chains.py tests/test_gpu_baseline_configs.py tests/test_gpu_api_state.py -m gpu -q --maxfail=25 > gpurun_out/r2c/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c/pytest.log; tail -30 gpurun_out/r2c/pytest.log
for P in 16 20 0; do timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-configs --no-literal --pool $P > gpurun_out/r2c/bench_pool$P.json 2> gpurun_out/r2c/bench_pool$P.err; echo "pool $P rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2c/bench_pool$P.json')); print('pool $P', d['value'], d['roofline']['kernel'], d['roofline']['frac'], d['e2e']['value'])"; done
for P in 16 20 0; do timeout 300 python bench.py --config c1 --steps 3 --warmup 3 --no-cpu --pool $P > gpurun_out/r2c/bench_c1_pool$P.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2c/bench_c1_pool$P.json')); print('c1 pool $P', d['value'], d['roofline']['kernel'], d['roofline']['frac'])"; done
