mkdir -p gpurun_out/r2h
timeout 900 python -m pytest tests/test_gpu_chains.py tests/test_gpu_host_driver.py -m gpu -q -k "rdf or widom" > gpurun_out/r2h/pytest_rdf.log 2>&1; tail -3 gpurun_out/r2h/pytest_rdf.log
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py tests/test_gpu_bound_reference.py -m gpu -q -s -k "million or isotherm or golden_traces" > gpurun_out/r2h/pytest_prints.log 2>&1; grep -E "max \|ep|k= |EnergyOfParticle calls|passed|failed" gpurun_out/r2h/pytest_prints.log
# RDF cost: NVT run with and without SampleRDF (N = 400 bulk), 2000 production steps
python - <<'PY' > gpurun_out/r2h/rdf_timing.txt 2>&1
import sys, json, numpy as np
sys.path.insert(0, "mc-porousmaterials_b200"); sys.path.insert(0, ".")
import mcpm_b200 as m
from oracle import oracle as O
g = np.load("tests/golden/reference_vectors.npz"); meta = json.load(open("tests/golden/reference_vectors.json"))
s = O.make_system(**meta["bulk"]["system"]); s.n_equil_sets = 0
for rdf in (0, 1):
    with m.Engine(148, 1024) as e:
        e.set_system(s)
        for c in range(148): e.set_positions(c, g["bulk_x"], g["bulk_y"], g["bulk_z"])
        e.run(1, 1, 200, seed=1, sample_rdf=rdf)
        e.run(2, 1, 2000, seed=1, sample_rdf=rdf)
        print("sample_rdf", rdf, "N", len(g["bulk_x"]), "148 chains x 2000 production steps:", e.last_kernel_ms(), "ms", e.last_kernel_name())
PY
cat gpurun_out/r2h/rdf_timing.txt
# ncu: C5 (k2_pair) and K3 launch
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k2_spec -c 1 -o /tmp/c5pair python bench.py --config c5 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2h/ncu_c5.log 2>&1
ncu -i /tmp/c5pair.ncu-rep --page raw --csv > gpurun_out/r2h/c5_k2_pair.raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:k3_box_energy -c 4 -o /tmp/k3 python bench.py --steps 3 --warmup 3 --no-cpu --no-literal > gpurun_out/r2h/ncu_k3.log 2>&1
ncu -i /tmp/k3.ncu-rep --page raw --csv > gpurun_out/r2h/k3.raw.csv 2>/dev/null
ls -la gpurun_out/r2h
