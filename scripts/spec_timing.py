"""Phase timing of the speculative kernel (needs libmcpm.so built with -DMCPM_SPEC_TIMING):
    make -C mc-porousmaterials_b200/csrc EXTRA=-DMCPM_SPEC_TIMING ; python scripts/spec_timing.py [point]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-porousmaterials_b200"))
import mcpm_b200 as m
from mcpm_b200 import workloads

g = np.load(os.path.join(ROOT, "tests", "golden", "isotherm_c2_start.npz"))
S = 4000
for k in [int(a) for a in sys.argv[1:]] or [0, 27, 39]:
    x, y, z = g[f"x{k}"], g[f"y{k}"], g[f"z{k}"]
    with m.Engine(40, 768) as e:
        e.use_speculation(1)
        e.set_system(workloads.argon_slit(mu=float(g["mu"][k]), n_equil_sets=0))
        for c in range(40):
            e.set_positions(c, x, y, z)
            e.set_dr(c, float(g["dr"][k]))
        e.run(1, 1, 500, seed=1)
        e.run(2, 1, S, seed=1)
        ms = e.last_kernel_ms()
        full = e.get_rdf(0).reshape(-1)
        h = full[:80].reshape(10, 8)
        sub = full[80:88]
    print("point", k, "N", len(x), "%.3f us/step" % (1e3 * ms / S), "rounds", h[0, 4], "commits", h[0, 5], "kept/round %.2f" % (h[0, 6] / h[0, 4]))
    print("  warp: eval  wait  commit  book   (clk per round)")
    for w in range(10):
        r = h[w, 4]
        print("  %d: %7.0f %7.0f %7.0f %7.0f  total %7.0f" % (w, h[w, 0] / r, h[w, 1] / r, h[w, 2] / r, h[w, 3] / r, h[w, :4].sum() / r))
    print("  translation candidate (warp 1): draws->r %.0f | ->xn %.0f | scan %.0f | wall %.0f | allsum %.0f | metropolis %.0f | publish %.0f  (n=%d)" % tuple(list(sub[:7] / max(sub[7], 1)) + [sub[7]]))
