"""Experiment: C2 isotherm split into capacity classes run as concurrent contexts (separate streams) vs one context."""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-porousmaterials_b200"))
import mcpm_b200 as m
from mcpm_b200 import workloads

g = np.load(os.path.join(ROOT, "tests", "golden", "isotherm_c2_start.npz"))
R, S = 64, 10000
pts = [(float(g["mu"][k]), g[f"x{k}"], g[f"y{k}"], g[f"z{k}"], float(g["dr"][k])) for k in range(40)]
classes = [(128, [p for p in pts if len(p[1]) < 90]), (320, [p for p in pts if 90 <= len(p[1]) < 280]), (704, [p for p in pts if len(p[1]) >= 280])]
if len(sys.argv) > 1 and sys.argv[1] == "two":
    classes = [(320, [p for p in pts if len(p[1]) < 280]), (704, [p for p in pts if len(p[1]) >= 280])]
if len(sys.argv) > 1 and sys.argv[1] == "single":
    classes = [(768, pts)]
engines = []
for cap, cp in classes:
    if not cp: continue
    e = m.Engine(len(cp) * R, cap)
    for c in range(len(cp) * R):
        mu, x, y, z, dr = cp[c % len(cp)]
        e.set_system(workloads.argon_slit(mu=mu, n_equil_sets=0), c)
        e.set_dr(c, dr)
        e.set_positions(c, x, y, z)
    engines.append(e)
    print("class cap", cap, "points", len(cp), "chains", e.n_chains)
TB = int(sys.argv[2]) if len(sys.argv) > 2 else 32
def run(e, k): e.run(k, 1, S, seed=5, threads=(TB if e.capacity >= 700 and len(engines) > 1 else 32))
for it in range(5):
    t0 = time.perf_counter()
    th = [threading.Thread(target=run, args=(e, it + 1)) for e in engines]
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    print("iter", it, "wall %.2f ms" % (dt * 1e3), "kernel ms", [round(e.last_kernel_ms(), 2) for e in engines], "moves/s %.4g" % (40 * R * S / dt))
