"""Golden vectors for row f4 of SURVEY.md §8 (spherical pore, hard spheres, NPT volume moves), made by the COMPILED, UNMODIFIED
reference through oracle/ref_harness.cpp.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden_f4.py        -> tests/golden/reference_f4.npz + reference_f4.json

Cases (every trace: the reference's own MoveParticle / SwapParticle / ChangeVolume / AdjustMCMoves under a seeded mt19937_64):
  sphere   Ar in a spherical cavity (Geometry sphere, Size 30 A, "lj" wall = SphericalLJ_Pot, src/potentials.cpp:474-504; no PBC):
           energies of an equilibrated configuration (BoxEnergy, EnergyOfParticle for stored / moved / ghost positions, a few
           positions outside the cavity) and a GCMC trace (insertions through InsertParticle's sphere branch, src/moves.cpp:25-36)
  hs       hard spheres in a periodic box (HardSphere_Pot, src/potentials.cpp:40-54; INT_MAX per overlap, self counted): energies of
           a jittered lattice and of a configuration with overlaps, an NVT trace and a GCMC trace (deletions always accepted: Q17)
  npt      LJ argon, bulk, NPT (NVolumeAttempts > 0, ExternalPressure; src/moves.cpp:191-231) AS SHIPPED, i.e. with CorrectEnergy after
           every step (src/main.cpp:71): ChangeVolume's old energy is the running box energy
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def sysdict(s):
    return s.as_dict()


def energies(ref, name, out, extra_trials=None, seed=5):
    x, y, z = ref.get_positions()
    n = len(x)
    box, pp, pw = ref.box_energy()
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, n, size=8)
    w = ref.system.width
    trials = np.column_stack([rng.random(8) * w[0], rng.random(8) * w[1], rng.random(8) * w[2]])
    if extra_trials is not None:
        trials = np.vstack([trials, extra_trials]); idx = np.concatenate([idx, rng.integers(0, n, size=len(extra_trials))])
    out[name + "_x"] = x; out[name + "_y"] = y; out[name + "_z"] = z
    out[name + "_box"] = box; out[name + "_pp"] = pp; out[name + "_pw"] = pw
    out[name + "_idx"] = idx; out[name + "_trials"] = trials
    out[name + "_moved"] = np.array([ref.moved_energy(int(i), *t) for i, t in zip(idx, trials)])
    out[name + "_ghost"] = np.array([ref.trial_energy(*t) for t in trials])
    out[name + "_cur"] = np.array([ref.particle_energy(int(i)) for i in idx])


def trace(ref, name, system, x, y, z, seed, sets, sps, correct, out):
    ref.setup(system)
    ref.set_positions(x, y, z)
    ref.seed(seed)
    ref.box_energy()
    tr, per_set, vol = [], [], []
    for s in range(1, sets + 1):
        ref.reset_stats()
        _, t = ref.run(sps, correct, True)
        tr.append(t)
        st = ref.state(); vs = ref.volume_state()
        ref.adjust(s)
        vs2 = ref.volume_state()
        per_set.append([st["n"], st["acceptance"], st["rejection"], st["n_displacements"], st["accept_swap"], st["reject_swap"], st["n_swaps"], ref.state()["dr"]])
        vol.append([vs["volume"], vs["width"][2], vs2["dv"], vs["acceptance_vol"], vs["rejection_vol"], vs["n_vol_changes"], vs["as_shipped_energy"]])
    fx, fy, fz = ref.get_positions()
    out[name + "_trace"] = np.concatenate(tr); out[name + "_per_set"] = np.array(per_set, dtype=np.float64); out[name + "_vol"] = np.array(vol)
    out[name + "_fx"] = fx; out[name + "_fy"] = fy; out[name + "_fz"] = fz
    out[name + "_fbox"] = ref.box_energy()[0]
    return {"system": sysdict(system), "seed": seed, "sets": sets, "steps_per_set": sps, "n0": len(x), "correct": correct}


def main():
    ref = O.RefOracle()
    out, meta = {}, {}
    # ---------------------------------------------------------------- sphere
    sph = O.make_system(geometry=3, wall_kind=2, n_layers=1, n_disp=1, n_vol=0, n_swap=1, n_equil_sets=2, pair_kind=0, width=(30.0, 30.0, 30.0),
                        temperature=87.0, eps_ff=119.8, sig_ff=3.405, rcut=15.0, molar_mass=39.948, mu=-1400.0, eps_sf=57.92, sig_sf=3.4025,
                        rho_s=0.382, delta=0.0, dr=3.0, pressure=-1.0, dv=1e-3)
    ref.setup(sph)
    ref.seed(31)
    ref.initial_config(120)
    ref.box_energy()
    ref.run(60000, 0)
    outside = np.array([[15.0, 15.0, 30.5], [32.0, 15.0, 15.0], [15.0, 15.0, 15.0], [15.0, 15.0, 28.2]])   # two outside the cavity, centre, near wall
    energies(ref, "sph", out, outside)
    meta["sph"] = {"system": sysdict(sph), "n": len(out["sph_x"])}
    meta["sph_trace"] = trace(ref, "spht", sph, out["sph_x"], out["sph_y"], out["sph_z"], 404, 3, 1500, 0, out)
    # ---------------------------------------------------------------- hard spheres
    hs = O.make_system(geometry=0, wall_kind=0, n_layers=0, n_disp=1, n_vol=0, n_swap=0, n_equil_sets=2, pair_kind=1, width=(30.0, 30.0, 30.0),
                       temperature=120.0, eps_ff=0.0, sig_ff=3.4, rcut=12.0, molar_mass=39.948, mu=-900.0, eps_sf=0.0, sig_sf=0.0, rho_s=0.0,
                       delta=0.0, dr=1.0, pressure=-1.0, dv=1e-3)
    rng = np.random.default_rng(9)
    site = (np.arange(6) + 0.5) * 5.0
    gx, gy, gz = np.meshgrid(site, site, site, indexing="ij")
    p = np.column_stack([gx.ravel(), gy.ravel(), gz.ravel()]) + (rng.random((216, 3)) - 0.5) * 0.8
    ref.setup(hs); ref.set_positions(p[:, 0], p[:, 1], p[:, 2])
    energies(ref, "hs", out)
    meta["hs"] = {"system": sysdict(hs), "n": 216}
    q = p.copy(); q[:40] = q[40:80] + 1.0                                   # 40 overlapping pairs
    ref.setup(hs); ref.set_positions(q[:, 0], q[:, 1], q[:, 2])
    energies(ref, "hsov", out)
    meta["hsov"] = {"system": sysdict(hs), "n": 216}
    meta["hs_trace"] = trace(ref, "hst", hs, p[:, 0], p[:, 1], p[:, 2], 77, 3, 1500, 0, out)
    hsg = O.make_system(**{**sysdict(hs), "n_swap": 1, "mu": -700.0})
    meta["hsg_trace"] = trace(ref, "hsgt", hsg, p[:, 0], p[:, 1], p[:, 2], 78, 2, 1500, 0, out)
    # ---------------------------------------------------------------- NPT (as shipped: CorrectEnergy on)
    # The stock build ABORTS in MC::ChangeVolume (`Eigen::Vector3d olddv(2);`, src/moves.cpp:198, trips Eigen's size assertion: src/Makefile
    # does not define NDEBUG).  The NPT vectors therefore come from the same sources compiled with -DNDEBUG (make -C oracle ref_ndebug).
    ref.close()
    ref = O.RefOracle(ndebug=True)
    g = np.load(os.path.join(HERE, "reference_vectors.npz"))
    npt = O.make_system(geometry=0, wall_kind=0, n_layers=0, n_disp=5, n_vol=1, n_swap=0, n_equil_sets=4, pair_kind=0, width=(30.0, 30.0, 30.0),
                        temperature=120.0, eps_ff=119.8, sig_ff=3.405, rcut=12.0, molar_mass=39.948, mu=0.0, eps_sf=0.0, sig_sf=0.0, rho_s=0.0,
                        delta=0.0, dr=0.5, pressure=3.0e7, dv=0.02)
    meta["npt_trace"] = trace(ref, "nptt", npt, g["bulk_x"], g["bulk_y"], g["bulk_z"], 2025, 4, 1500, 1, out)
    np.savez_compressed(os.path.join(HERE, "reference_f4.npz"), **out)
    with open(os.path.join(HERE, "reference_f4.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    for k in ("spht", "hst", "hsgt", "nptt"):
        tr = out[k + "_trace"]
        print(k, "steps", len(tr), "accepted", int((tr & 1).sum()), "kinds", np.bincount(tr >> 1), "per_set", out[k + "_per_set"][-1], "vol", out[k + "_vol"][-1])


if __name__ == "__main__":
    main()
