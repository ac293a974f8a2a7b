"""Equilibrated start configurations for BASELINE config C2 (40-point Ar isotherm).

Run in the build container:  python tests/golden/make_isotherm_start.py
For every chemical potential mu_k = -1800 + k*666/39 K the CPU oracle (trace-identical to the compiled
reference, see tests/test_oracle_vs_reference.py) starts from N0 = 50 random particles with mt19937_64
seed 1000+k and runs STEPS GCMC steps (disp:swap = 1:1, dr adapted every 10 000 steps during the first
half).  Both bench arms (GPU and reference CPU) and the ensemble tests start from these states.
"""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

STEPS = 800_000
SET = 10_000


def mu_grid():
    return [-1800.0 + k * (666.0 / 39.0) for k in range(40)]


def one(k):
    mu = mu_grid()[k]
    nsets = STEPS // SET
    s = O.argon_slit(mu=mu, n_equil_sets=nsets // 2)
    c = O.COracle(s, 1024)
    c.seed(1000 + k)
    c.initial_config(50)
    c.box_energy()
    hist = []
    for st in range(1, nsets + 1):
        c.run_sets(st, 1, SET, 0)
        hist.append(c.state()["n"])
    x, y, z = c.get_positions()
    return k, mu, x, y, z, c.state()["dr"], hist


def main():
    with mp.Pool(min(8, os.cpu_count())) as pool:
        res = pool.map(one, list(range(39, -1, -1)))
    res.sort()
    out = {"mu": np.array([r[1] for r in res]), "dr": np.array([r[5] for r in res]),
           "n": np.array([len(r[2]) for r in res], dtype=np.int32),
           "n_history": np.array([r[6] for r in res], dtype=np.int32)}
    for k, mu, x, y, z, dr, hist in res:
        out[f"x{k}"] = x; out[f"y{k}"] = y; out[f"z{k}"] = z
    np.savez_compressed(os.path.join(HERE, "isotherm_c2_start.npz"), **out)
    for k, mu, x, y, z, dr, hist in res:
        print(k, round(mu, 1), len(x), round(dr, 3), hist[::10])


if __name__ == "__main__":
    main()
