"""One K2 configuration for ncu: python scripts/profile_case.py <point> <threads> <chains> <steps>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-porousmaterials_b200"))
import mcpm_b200 as m
from mcpm_b200 import workloads

k, T, nchains, S = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
cap = int(sys.argv[5]) if len(sys.argv) > 5 else 768
g = np.load(os.path.join(ROOT, "tests", "golden", "isotherm_c2_start.npz"))
x, y, z = g[f"x{k}"], g[f"y{k}"], g[f"z{k}"]
with m.Engine(nchains, cap) as e:
    e.set_system(workloads.argon_slit(mu=float(g["mu"][k]), n_equil_sets=0))
    for c in range(nchains):
        e.set_positions(c, x, y, z)
        e.set_dr(c, float(g["dr"][k]))
    e.run(1, 1, 200, seed=1, threads=T)
    st = e.run(2, 1, S, seed=1, threads=T)
    ms = e.last_kernel_ms()
    print("point %d N=%d T=%d chains=%d cap=%d: %.3f ms, %.3f us/step/chain, %.4g moves/s" % (k, len(x), T, nchains, cap, ms, 1e3 * ms / S, nchains * S / (ms * 1e-3)))
