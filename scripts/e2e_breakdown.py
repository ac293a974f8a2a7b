"""Where does the end-to-end step go?  (GPU box)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-porousmaterials_b200")); sys.path.insert(0, ROOT)
import bench, mcpm_b200 as m
w = bench.build_workload("c2", 0, 0)
pts, R, S = w["points"], w["replicas"], w["steps_per_set"]
P = len(pts); n_chains = P * R
nmax = max(len(p[1]) for p in pts); cap = int(-(-int(nmax * 1.15 + 40) // 32) * 32)
eng = m.Engine(n_chains, cap); cap = eng.capacity
hx = torch.zeros((n_chains, cap), dtype=torch.float64).pin_memory(); hy = torch.zeros_like(hx).pin_memory(); hz = torch.zeros_like(hx).pin_memory()
X, Y, Z = hx.numpy(), hy.numpy(), hz.numpy(); n0 = np.zeros(n_chains, dtype=np.int32)
for c in range(n_chains):
    sysd, x, y, z, dr = pts[c % P]; eng.set_system(sysd, c); eng.set_dr(c, dr); n0[c] = len(x)
    X[c, :len(x)] = x; Y[c, :len(x)] = y; Z[c, :len(x)] = z
eng.set_positions_all(n0, X, Y, Z)
for k in range(3): st = eng.run(k + 1, 1, S, seed=1)
nb, _, _, _ = eng.get_positions_all(X, Y, Z)
T = {"h2d": 0, "energy": 0, "run": 0, "d2h": 0, "py": 0}
for it in range(5):
    t0 = time.perf_counter(); eng.set_positions_all(nb, X, Y, Z)
    t1 = time.perf_counter(); ep = np.array([s.pair for s in st]); ew = np.array([s.wall for s in st])
    t2 = time.perf_counter(); eng.set_running_energy(ep, ew)
    t3 = time.perf_counter(); st = eng.run(10 + it, 1, S, seed=1)
    t4 = time.perf_counter(); nb, _, _, _ = eng.get_positions_all(X, Y, Z)
    t5 = time.perf_counter()
    T["h2d"] += t1 - t0; T["py"] += t2 - t1; T["energy"] += t3 - t2; T["run"] += t4 - t3; T["d2h"] += t5 - t4
print({k: round(1e3 * v / 5, 2) for k, v in T.items()}, "kernel ms", eng.last_kernel_ms())
