"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference compiled into oracle/_ref/libmcref.so by
`make -C oracle ref`):   python tests/golden/make_golden.py

The reference ships no tests or fixtures of its own, so these vectors — produced by
driving the compiled reference's own EnergyOfParticle / BoxEnergy / MoveParticle / SwapParticle /
AdjustMCMoves / ReadInputFile through oracle/ref_harness.cpp — are what pins the C oracle and the CUDA
path.  Everything is stored at full double precision (.npz).
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


def energies_case(ref, name, system, n0, seed, equil_steps, out):
    """Config equilibrated by the reference itself, then its per-particle and box energies."""
    ref.setup(system)
    ref.seed(seed)
    ref.initial_config(n0)
    ref.reset_stats()
    ref.run(equil_steps, 0)
    x, y, z = ref.get_positions()
    box, pp, pw = ref.box_energy()
    n = len(x)
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, n, size=8)
    trials = np.column_stack([rng.random(8) * system.width[0], rng.random(8) * system.width[1],
                              rng.random(8) * system.width[2]])
    moved = np.array([ref.moved_energy(int(i), *t) for i, t in zip(idx, trials)])
    ghost = np.array([ref.trial_energy(*t) for t in trials])
    cur = np.array([ref.particle_energy(int(i)) for i in idx])
    out[name + "_x"] = x; out[name + "_y"] = y; out[name + "_z"] = z
    out[name + "_box"] = box; out[name + "_pp"] = pp; out[name + "_pw"] = pw
    out[name + "_idx"] = idx; out[name + "_trials"] = trials
    out[name + "_moved"] = moved; out[name + "_ghost"] = ghost; out[name + "_cur"] = cur
    return {"system": sysdict(system), "n": n, "seed": seed, "equil_steps": equil_steps}


def trace_case(ref, name, system, x, y, z, seed, sets, steps_per_set, out):
    """Accept/reject trace of the reference's step loop (main.cpp:63-84 without CorrectEnergy's
    recompute, which does not touch the chain) from a given configuration and mt19937_64 seed."""
    ref.setup(system)
    ref.set_positions(x, y, z)
    ref.seed(seed)
    ref.box_energy()
    trace = []
    per_set = []
    for s in range(1, sets + 1):
        ref.reset_stats()
        _, tr = ref.run(steps_per_set, 0, True)
        trace.append(tr)
        st = ref.state()
        ref.adjust(s)
        per_set.append([st["n"], st["acceptance"], st["rejection"], st["n_displacements"], st["accept_swap"],
                        st["reject_swap"], st["n_swaps"], ref.state()["dr"]])
    fx, fy, fz = ref.get_positions()
    box, _, _ = ref.box_energy()
    out[name + "_trace"] = np.concatenate(trace)
    out[name + "_per_set"] = np.array(per_set, dtype=np.float64)
    out[name + "_fx"] = fx; out[name + "_fy"] = fy; out[name + "_fz"] = fz
    out[name + "_fbox"] = box
    return {"system": sysdict(system), "seed": seed, "sets": sets, "steps_per_set": steps_per_set, "n0": len(x)}


def main():
    assert O.have_reference_lib(), "build oracle/_ref first: make -C oracle ref"
    ref = O.RefOracle()
    out, meta = {}, {}

    # known answers, evaluated by the reference itself
    s = O.argon_slit()
    ref.setup(s)
    r0 = 2.0 ** (1.0 / 6.0) * 3.405
    ref.set_positions([10.0, 10.0 + r0], [10.0, 10.0], [10.0, 10.0])
    ka = {"two_ar_pair": float(ref.particle_energy(0)[0])}
    ref.set_positions([10.0], [10.0], [3.4025])
    ka["wall_z_sigma"] = float(ref.particle_energy(0)[2])
    ref.set_positions([10.0], [10.0], [10.0])
    ka["wall_z_center"] = float(ref.particle_energy(0)[2])
    ref.set_positions([10.0], [10.0], [20.5])
    ka["wall_outside"] = float(ref.particle_energy(0)[2])
    ka["zz_c1"] = ref.swap_zz(); ka["volume_c1"] = ref.volume()
    meta["known_answers"] = ka

    # energy vectors
    meta["c1"] = energies_case(ref, "c1", O.argon_slit(), 450, 12345, 120000, out)
    meta["c1raw"] = energies_case(ref, "c1raw", O.argon_slit(), 300, 999, 0, out)          # overlaps, ~1e26 K
    meta["ch4"] = energies_case(ref, "ch4", O.methane_slit(), 200, 4242, 60000, out)
    bulk = O.argon_bulk(size=30.0, rcut=12.0, T=120.0)
    meta["bulk"] = energies_case(ref, "bulk", bulk, 400, 77, 40000, out)

    # traces: GCMC from the equilibrated C1 config, and NVT translations in bulk
    meta["c1_trace"] = trace_case(ref, "c1t", O.argon_slit(n_equil_sets=2), out["c1_x"], out["c1_y"], out["c1_z"],
                                  2024, 4, 2500, out)
    lowmu = O.argon_slit(mu=-1900.0, n_equil_sets=1)   # box nearly empty: exercises Q12 (moves on N=0)
    meta["low_trace"] = trace_case(ref, "lowt", lowmu, out["c1_x"][:3], out["c1_y"][:3], out["c1_z"][:3], 7, 2, 4000, out)
    meta["bulk_trace"] = trace_case(ref, "bulkt", bulk, out["bulk_x"], out["bulk_y"], out["bulk_z"], 31337, 2, 2000, out)

    # the reference's Random() stream for one seed (quirk Q1)
    ref.seed(12345)
    out["rng_12345"] = np.array([ref.random() for _ in range(4096)])

    # parser: BASELINE.md's config-1 input file through the reference's ReadInputFile
    path = os.path.join(HERE, "ar_cfg1.in")
    ref.close()
    ref = O.RefOracle()   # ReadInputFile accumulates into thermoSys counters: it needs a fresh MC object
    rc, ps, n0, proj, fname, bname, sets = ref.read_input(path)
    assert rc == 0
    meta["parser_c1"] = {"system": sysdict(ps), "n0": n0, "project": proj, "fluid": fname, "box": bname,
                         "sets": [float(v) for v in sets]}

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "reference_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()
