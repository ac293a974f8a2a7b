mkdir -p gpurun_out/r2j
for C in c2 c1 c5 c3; do timeout 300 python bench.py --config $C --steps 5 --warmup 3 --no-cpu --no-configs --no-literal 2>/dev/null | grep '^{' | tail -1 > gpurun_out/r2j/bench_$C.json; python -c "
import json; d=json.load(open('gpurun_out/r2j/bench_$C.json')); print('$C', d['value'], d['roofline']['kernel'], round(d['roofline']['frac'],4), round(d['roofline']['frac_executed'],4), d['e2e']['value'])"; done
timeout 900 python -m pytest tests/test_gpu_chains.py tests/test_gpu_baseline_configs.py tests/test_gpu_api_state.py tests/test_gpu_f4.py -m gpu -q > gpurun_out/r2j/pytest.log 2>&1; tail -4 gpurun_out/r2j/pytest.log
