mkdir -p gpurun_out/r2g
BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2g/bench_n2.json 2> gpurun_out/r2g/bench_n2.err; echo "bench n2 rc=$?"; tail -12 gpurun_out/r2g/bench_n2.err; ls -la gpurun_out/r2g/bench_n2.json; cut -c1-300 gpurun_out/r2g/bench_n2.json
