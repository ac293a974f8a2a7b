"""Golden vectors for the two-box (Gibbs ensemble) branches, made by the COMPILED, UNMODIFIED reference through oracle/ref_harness.cpp.
Run in the build container (needs /root/reference):

    python tests/golden/make_golden_gibbs.py     -> tests/golden/reference_gibbs.npz + reference_gibbs.json

What runs: the reference's own step-loop body (src/main.cpp:67-71) with thermoSys.nBoxes == 2 — MoveParticle's box selection
(src/moves.cpp:127-137), SwapParticle's Gibbs branch (src/moves.cpp:342-400), CorrectEnergy, ResetStats / AdjustMCMoves on both
boxes — under a seeded mt19937_64.  NVolumeAttempts = 0: ChangeVolume aborts on entry in the stock build (quirk Q19).

Cases:
  vle    LJ argon at T* ~ 1: a liquid-like bulk box (L = 24 A) and a vapour-like bulk box (L = 40 A)
  pore   a graphite slit pore (H = 20 A, SlitLJ_Pot) exchanging particles with a bulk reservoir box (L = 50 A)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def lattice(n, w, rng):
    m = int(np.ceil(n ** (1.0 / 3.0)))
    site = (np.arange(m) + 0.5) / m
    g = np.array(np.meshgrid(site, site, site, indexing="ij")).reshape(3, -1)[:, :n]
    return (g * np.array(w)[:, None] * np.array([1.0, 1.0, 0.6])[:, None] + np.array([0.0, 0.0, 0.2 * w[2]])[:, None]
            + (rng.random((3, n)) - 0.5) * 0.2)


def case(name, a, b, n0, seed, equil_steps, sets, sps, out):
    rng = np.random.default_rng(seed)
    ref = O.RefOracle(a); ref.add_box(b)
    ref.set_positions_box(0, *lattice(n0[0], a.width, rng)); ref.set_positions_box(1, *lattice(n0[1], b.width, rng))
    ref.box_energy_box(0); ref.box_energy_box(1)
    ref.seed(seed)
    ref.reset_stats()
    ref.gibbs_run(equil_steps, 1)
    for ib in (0, 1):
        x, y, z = ref.get_positions_box(ib)
        out[f"{name}_x{ib}"] = x; out[f"{name}_y{ib}"] = y; out[f"{name}_z{ib}"] = z
        out[f"{name}_box{ib}"] = ref.box_energy_box(ib)
    ref.seed(seed + 1)
    tr, per_set = [], []
    for s in range(1, sets + 1):
        ref.reset_stats()
        tr.append(ref.gibbs_run(sps, 1))
        row = []
        st = [ref.state_box(0), ref.state_box(1)]
        wd = [ref.widom_box(0), ref.widom_box(1)]
        ref.adjust(s)
        for ib in (0, 1):
            row += [st[ib]["n"], st[ib]["acceptance"], st[ib]["rejection"], st[ib]["n_displacements"], ref.state_box(ib)["dr"],
                    wd[ib][0][0], wd[ib][0][1], wd[ib][1][0], wd[ib][1][1]]
        row += [st[0]["accept_swap"], st[0]["reject_swap"], st[0]["n_swaps"]]
        per_set.append(row)
    out[name + "_trace"] = np.concatenate(tr); out[name + "_per_set"] = np.array(per_set, dtype=np.float64)
    for ib in (0, 1):
        x, y, z = ref.get_positions_box(ib)
        out[f"{name}_fx{ib}"] = x; out[f"{name}_fy{ib}"] = y; out[f"{name}_fz{ib}"] = z
        out[f"{name}_fbox{ib}"] = ref.box_energy_box(ib)
    ref.close()
    return {"box0": a.as_dict(), "box1": b.as_dict(), "seed": seed + 1, "sets": sets, "steps_per_set": sps, "correct": 1}


def main():
    out, meta = {}, {}
    kw = dict(geometry=0, wall_kind=0, n_layers=0, n_disp=4, n_vol=0, n_swap=1, n_equil_sets=2, pair_kind=0, temperature=120.0,
              eps_ff=119.8, sig_ff=3.405, rcut=10.0, molar_mass=39.948, mu=0.0, eps_sf=0.0, sig_sf=0.0, rho_s=0.0, delta=0.0, dr=0.6,
              pressure=-1.0, dv=1e-3)
    a = O.make_system(width=(24.0, 24.0, 24.0), **kw)
    b = O.make_system(width=(40.0, 40.0, 40.0), **{**kw, "dr": 6.0})
    meta["vle"] = case("vle", a, b, (230, 60), 7001, 120000, 3, 2500, out)
    pore = O.make_system(**{**kw, "geometry": 1, "wall_kind": 1, "n_layers": 1, "width": (30.0, 30.0, 20.0), "temperature": 87.0, "eps_sf": 57.92,
                            "sig_sf": 3.4025, "rho_s": 0.382, "delta": 0.0, "n_disp": 2, "dr": 1.0, "rcut": 15.0})
    bulk = O.make_system(**{**kw, "width": (50.0, 50.0, 50.0), "temperature": 87.0, "n_disp": 2, "dr": 10.0, "rcut": 15.0})
    meta["pore"] = case("pore", pore, bulk, (150, 100), 7101, 120000, 3, 2500, out)
    np.savez_compressed(os.path.join(HERE, "reference_gibbs.npz"), **out)
    with open(os.path.join(HERE, "reference_gibbs.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    for k in ("vle", "pore"):
        tr = out[k + "_trace"]
        print(k, "steps", len(tr), "codes", dict(zip(*np.unique(tr, return_counts=True))), "last set", out[k + "_per_set"][-1])


if __name__ == "__main__":
    main()
