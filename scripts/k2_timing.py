"""Phase timing of a translation step in k2_chains (one warp per chain) under the C2 bench load
(needs: make -C mc-porousmaterials_b200/csrc EXTRA=-DMCPM_K2_TIMING)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-porousmaterials_b200"))
sys.path.insert(0, ROOT)
import mcpm_b200 as m
import bench

w = bench.build_workload("c2", 128, 0)
pts = w["points"]
P, R = len(pts), 128
nch = P * R
with m.Engine(nch, 704) as e:
    for c in range(nch):
        d, x, y, z, dr = pts[c % P]
        e.set_system(d, c)
        e.set_positions(c, x, y, z)
        e.set_dr(c, dr)
    e.run(1, 1, 2000, seed=1)
    e.reset_accumulators()
    e.run(2, 1, 10000, seed=1)
    ms = e.last_kernel_ms()
    print("chains", nch, "kernel", e.last_kernel_name(), "%.2f ms, %.4g moves/s" % (ms, nch * 10000 / (ms * 1e-3)))
    print("point N | prelude (uniforms, selection, new position) | scan | wall + reductions | dE + Metropolis | commit + bookkeeping | total   [clk per translation step]")
    for k in (0, 5, 12, 27, 39):
        h = e.get_rdf(k).reshape(-1)[:8]
        r = max(h[5], 1)
        print("%2d %4d | %6.0f | %6.0f | %6.0f | %6.0f | %6.0f | %6.0f" % ((k, len(pts[k][1])) + tuple(h[:5] / r) + (h[:5].sum() / r,)))
