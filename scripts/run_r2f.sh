mkdir -p gpurun_out/r2f
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --durations=12 > gpurun_out/r2f/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f/pytest.log; tail -45 gpurun_out/r2f/pytest.log
