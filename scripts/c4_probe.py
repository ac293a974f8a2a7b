"""Where a step of k2_cells goes on C4 (148 chains of N = 20 000): with / without Widom sampling, two amplitudes."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-porousmaterials_b200"))
import mcpm_b200 as M
from mcpm_b200 import workloads

x, y, z = workloads.c4_lattice()
nch = int(os.environ.get("CHAINS", "148"))
for dr in (0.5,):
    for nos in (0, 1):
        with M.Engine(nch, 20032) as e:
            d = workloads.argon_bulk(n_equil_sets=0); d.dr = dr
            e.set_system(d)
            for c in range(nch):
                e.set_positions(c, x, y, z)
            e.run(1, 1, 500, seed=1, no_sampling=nos)
            st = e.run(2, 2, 2000, seed=1, no_sampling=nos)
            ms = e.last_kernel_ms()
            acc = sum(s.tot_disp_acc for s in st) / max(1, sum(s.tot_disp for s in st))
            print(f"dr={dr} no_sampling={nos} kernel={e.last_kernel_name()} {ms:.2f} ms -> {ms * 1e3 / 4000:.2f} us/step/chain, {nch * 4000 / ms / 1e3:.3g} moves/s, acceptance {acc:.2f}", flush=True)
