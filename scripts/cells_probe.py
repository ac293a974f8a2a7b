"""K2 through the cell list on config C4: time per step with and without Widom sampling, cell list vs all-pairs."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-porousmaterials_b200"))
sys.path.insert(0, ROOT)
import mcpm_b200 as m
from mcpm_b200 import workloads
import bench

w = bench.build_workload("c4", 0, 0)
d, x, y, z, dr = w["points"][0]
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 148
S = 500
for label, n_equil, cells in (("cells, no sampling", 10, True), ("cells, Widom", 0, True), ("all pairs, no sampling", 10, False), ("all pairs, Widom", 0, False)):
    sysd = workloads.argon_bulk(n_equil_sets=n_equil)
    with m.Engine(nch, 23040) as e:
        e.use_cell_list(cells)
        e.set_system(sysd)
        for c in range(nch):
            e.set_positions(c, x, y, z)
            e.set_dr(c, dr)
        e.run(1, 1, 100, seed=1)
        st = e.run(2, 1, S, seed=1)
        ms = e.last_kernel_ms()
        print("%-24s %s: %.3f ms for %d steps = %.2f us/step, %.3g moves/s, acceptance %.2f" % (label, e.last_kernel_name(), ms, S, 1e3 * ms / S, nch * S / (ms * 1e-3), st[0].tot_disp_acc / max(1, st[0].tot_disp)))
