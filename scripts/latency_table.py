"""Per-step latency of K2 as a function of loading N and threads per chain T (GPU box only).

    python scripts/latency_table.py [chains_per_case]
Prints us/step for one state point at a time, every chain the same system, so the kernel time is the
chain latency (all chains resident, no tail effect)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-porousmaterials_b200"))
import mcpm_b200 as m
from mcpm_b200 import workloads

g = np.load(os.path.join(ROOT, "tests", "golden", "isotherm_c2_start.npz"))
nchains = int(sys.argv[1]) if len(sys.argv) > 1 else 148
S = 4000
print("chains", nchains, "steps", S)
for k in (0, 5, 12, 27, 39):
    x, y, z = g[f"x{k}"], g[f"y{k}"], g[f"z{k}"]
    row = []
    for T in (32, 128, 256, "spec"):
        with m.Engine(nchains, 768) as e:
            label = T
            e.use_speculation(1 if isinstance(T, str) else 0)
            if isinstance(T, str):
                T = 0
            e.set_system(workloads.argon_slit(mu=float(g["mu"][k]), n_equil_sets=0))
            for c in range(nchains):
                e.set_positions(c, x, y, z)
                e.set_dr(c, float(g["dr"][k]))
            e.run(1, 1, 500, seed=1, threads=T)
            st = e.run(2, 1, S, seed=1, threads=T)
            ms = e.last_kernel_ms()
            evals = sum(int(s.pair_evals) for s in st)
            row.append("T=%s: %.2f us/step (%.3g evals/s)" % (label, 1e3 * ms / S, evals / (ms * 1e-3) if False else (evals - 0) / (ms * 1e-3)))
    print("point", k, "N=%d" % len(x), " | ".join(row))
