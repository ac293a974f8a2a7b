"""Golden vectors for the cylindrical pores of row f4 (CylindricalLJ10_4, src/potentials.cpp:412-444, and CylindricalSteele10_4_3, :445-469),
made by the COMPILED, UNMODIFIED reference (whose 2F1 is the vendored Numerical-Recipes hypgeo).  Run in the build container:

    python tests/golden/make_golden_cyl.py        -> tests/golden/reference_cyl.npz + reference_cyl.json

Per wall: energies of an equilibrated configuration (BoxEnergy, EnergyOfParticle for stored / moved / ghost positions, incl. positions
outside the pore and on its axis) and a GCMC trace (insertions through InsertParticle's cylinder branch, src/moves.cpp:37-45; PBC along x only).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_golden_f4 import energies, sysdict, trace  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    ref = O.RefOracle()
    out, meta = {}, {}
    for name, wk, layers in (("cyl3", 3, 3), ("cyl4", 4, 1)):
        s = O.make_system(geometry=2, wall_kind=wk, n_layers=layers, n_disp=1, n_vol=0, n_swap=1, n_equil_sets=2, pair_kind=0, width=(30.0, 30.0, 30.0),
                          temperature=87.0, eps_ff=119.8, sig_ff=3.405, rcut=15.0, molar_mass=39.948, mu=-1450.0, eps_sf=57.92, sig_sf=3.4025,
                          rho_s=0.382, delta=3.35, dr=3.0, pressure=-1.0, dv=1e-3)
        ref.setup(s)
        ref.seed(50 + wk)
        ref.initial_config(100)
        ref.box_energy()
        ref.run(60000, 0)
        special = np.array([[15.0, 15.0, 30.5], [3.0, 31.0, 15.0], [7.0, 15.0, 15.0], [22.0, 15.0, 28.3], [1.0, 24.0, 24.0]])   # outside x2, axis, near wall x2
        energies(ref, name, out, special)
        meta[name] = {"system": sysdict(s), "n": len(out[name + "_x"])}
        meta[name + "_trace"] = trace(ref, name + "t", s, out[name + "_x"], out[name + "_y"], out[name + "_z"], 600 + wk, 3, 1500, 0, out)
        tr = out[name + "t_trace"]
        print(name, "n", meta[name]["n"], "box", out[name + "_box"], "accepted", int((tr & 1).sum()), "kinds", np.bincount(tr >> 1), "final n", out[name + "t_per_set"][-1][0])
    np.savez_compressed(os.path.join(HERE, "reference_cyl.npz"), **out)
    with open(os.path.join(HERE, "reference_cyl.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
