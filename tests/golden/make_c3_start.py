"""Equilibrated start configuration for BASELINE config C3 (CH4 in a slit, 300 K, mu = -3050 K, Rcut 92 => L = 184 A,
N ~ 5k).  Run in the build container:  python tests/golden/make_c3_start.py
Start: 3 layers of a 41x41 square grid (spacing 4.4 A, 5043 sites, first 5000 occupied), then STEPS GCMC steps
(disp:swap = 1:3) with the CPU oracle (trace-identical to the compiled reference), mt19937_64 seed 3003."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

STEPS = 150_000


def main():
    s = O.methane_slit(rcut=92.0)
    g = (np.arange(41) + 0.5) * 4.4
    xs, ys, zs = [], [], []
    for zl in (3.6, 10.0, 16.4):
        for a in g:
            for b in g:
                xs.append(a); ys.append(b); zs.append(zl)
    x, y, z = np.array(xs[:5000]), np.array(ys[:5000]), np.array(zs[:5000])
    c = O.COracle(s, 8192)
    c.set_positions(x, y, z)
    c.seed(3003)
    print("start energy/particle", c.box_energy()[0][3] / 5000)
    hist = []
    for k in range(STEPS // 10000):
        c.run_sets(1, 1, 10000, 0)
        hist.append(c.state()["n"])
    x, y, z = c.get_positions()
    print("n history", hist, "dr", c.state()["dr"], "energy/particle", c.box_energy()[0][3] / len(x))
    np.savez_compressed(os.path.join(HERE, "c3_start.npz"), x=x, y=y, z=z, dr=c.state()["dr"], n_history=np.array(hist))


if __name__ == "__main__":
    main()
