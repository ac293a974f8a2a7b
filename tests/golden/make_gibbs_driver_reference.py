"""Loadings of the two boxes of tests/golden/gibbs_ar.in as the COMPILED, UNMODIFIED reference program gives them (oracle/_ref/simulate, built by
oracle/Makefile from /root/reference/src): RUNS independent runs (the reference seeds itself from random_device) of 3 equilibration + SETS
production sets of 20 000 steps; per production set the particle numbers and energies per particle PrintStats shows.

    python tests/golden/make_gibbs_driver_reference.py     -> tests/golden/gibbs_driver_reference.json (+ the PrintParams echo)
"""
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
RUNS, SETS = 6, 15


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "simulate")
    text = open(os.path.join(HERE, "gibbs_ar.in")).read()
    text = re.sub(r"ProductionSets \d+", f"ProductionSets {SETS}", text)
    procs, dirs = [], []
    for r in range(RUNS):
        d = tempfile.mkdtemp(prefix="gibbs_ref_")
        open(os.path.join(d, "gibbs.in"), "w").write(text)
        procs.append(subprocess.Popen([exe, "gibbs.in"], cwd=d, stdout=open(os.path.join(d, "run.log"), "w"), stderr=subprocess.STDOUT))
        dirs.append(d)
    for p in procs:
        p.wait()
    n = {"liquid": [], "vapour": []}
    e = {"liquid": [], "vapour": []}
    for d in dirs:
        log = open(os.path.join(d, "run.log")).read()
        prod = log.split("\nSet: ")[1:]                                   # production sets only ("Equilibrium set:" blocks come first)
        for b in n:
            rows = [re.search(r"Box " + b + r":.*?NumParticles; BoxSize; Energy/Particle:\s+(\d+); [\d.]+; (-?[\d.]+)", blk, re.S) for blk in prod]
            n[b].append([int(m.group(1)) for m in rows]); e[b].append([float(m.group(2)) for m in rows])
    out = {"runs": RUNS, "production_sets": len(n["liquid"][0]), "steps_per_set": 20000}
    for b in n:
        means = np.array([np.mean(r) for r in n[b]]); emeans = np.array([np.mean(r) for r in e[b]])
        out[b] = {"mean_n": float(means.mean()), "sem_n": float(means.std(ddof=1) / np.sqrt(RUNS)), "run_means_n": means.tolist(),
                  "mean_e_per_particle": float(emeans.mean()), "sem_e_per_particle": float(emeans.std(ddof=1) / np.sqrt(RUNS))}
    params = open(os.path.join(dirs[0], "run.log")).read()
    params = params[params.index("Simulation parameters:"):params.index("Initial configuration set.")]
    params = re.sub(r"Production sets: \d+", "Production sets: 6", params)
    out["print_params"] = params
    json.dump(out, open(os.path.join(HERE, "gibbs_driver_reference.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "print_params"}, indent=1))


if __name__ == "__main__":
    main()
