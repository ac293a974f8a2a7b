"""Small K2/K1/K3 run for compute-sanitizer (racecheck / memcheck): multi-warp chains, swaps, Widom+RDF kernels."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mc-porousmaterials_b200"))
import mcpm_b200 as m
from mcpm_b200 import workloads
g = np.load(os.path.join(ROOT, "tests", "golden", "isotherm_c2_start.npz"))
for T in (32, 128):
    with m.Engine(6, 704) as e:
        for c in range(6):
            k = (5, 20, 39)[c % 3]
            e.set_system(workloads.argon_slit(mu=float(g["mu"][k]), n_equil_sets=1), c)
            e.set_positions(c, g[f"x{k}"], g[f"y{k}"], g[f"z{k}"])
        st = e.run(1, 2, 150, seed=3, threads=T, check_energy=1)
        print("T", T, [s.n for s in st], e.box_energy(0)[0][3], e.particle_energy(1, 3, [1.0, 2.0, 9.0])[1][3])
with m.Engine(2, 512) as e:       # NVT: Widom + RDF sampling kernel
    d = workloads.argon_bulk(size=30.0, rcut=12.0, n_equil_sets=0)
    rng = np.random.default_rng(1)
    for c in range(2):
        e.set_system(d, c)
        e.set_positions(c, rng.random(300) * 30, rng.random(300) * 30, rng.random(300) * 30)
    st = e.run(1, 1, 60, seed=4, threads=64, sample_rdf=1)
    print("nvt", st[0].widom_insertions, e.get_rdf(0).sum())
print("done")
