mkdir -p gpurun_out/r2i
timeout 900 python -m pytest tests/test_gpu_chains.py tests/test_gpu_host_driver.py -m gpu -q -k "rdf or widom" > gpurun_out/r2i/pytest_rdf.log 2>&1; tail -3 gpurun_out/r2i/pytest_rdf.log
python - <<'PY' > gpurun_out/r2i/rdf_timing.txt 2>&1
import sys, json, numpy as np
sys.path.insert(0, "mc-porousmaterials_b200"); sys.path.insert(0, ".")
import mcpm_b200 as m
from oracle import oracle as O
g = np.load("tests/golden/reference_vectors.npz"); meta = json.load(open("tests/golden/reference_vectors.json"))
for case, thr in (("bulk", 0), ("c1", 0), ("c1", 256)):
    s = O.make_system(**meta[case]["system"]); s.n_equil_sets = 0; s.n_swap = 0
    for rdf in (0, 1):
        with m.Engine(148, 1024) as e:
            e.set_system(s)
            for c in range(148): e.set_positions(c, g[case + "_x"], g[case + "_y"], g[case + "_z"])
            e.run(1, 1, 200, seed=1, sample_rdf=rdf, threads=thr)
            e.run(2, 1, 2000, seed=1, sample_rdf=rdf, threads=thr)
            n = len(g[case + "_x"])
            ms = e.last_kernel_ms()
            print(case, "threads", thr or "auto", "sample_rdf", rdf, "N", n, "148 chains x 2000 production steps:", round(ms, 2), "ms;", "pairs binned/s %.3g" % (148 * 2000 * n * (n - 1) / 2 / (ms * 1e-3)) if rdf else "")
PY
cat gpurun_out/r2i/rdf_timing.txt
