"""The reference's OWN step loop (MoveParticle / SwapParticle / CorrectEnergy as shipped) on C1: stock build (CPU energies) vs the build bound to
libmcpm.so at the energy seam (oracle/energy_bound.cpp: K1 as a resident service, K3 for BoxEnergy).  Run on a GPU box."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O

g = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
meta = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))
m = meta["c1_trace"]
s = O.make_system(**m["system"])
n0 = m["n0"]
x, y, z = g["c1_x"][:n0], g["c1_y"][:n0], g["c1_z"][:n0]
for correct in (0, 1):
    for bound in (False, True):
        r = O.RefOracle(bound=bound)
        r.setup(s); r.set_positions(x, y, z); r.seed(m["seed"]); r.box_energy(); r.reset_stats()
        r.run(200, correct)                       # warm-up (CUDA context, first launches)
        steps = 4000
        t0 = time.perf_counter(); r.run(steps, correct); dt = time.perf_counter() - t0
        st = r.bound_stats()
        extra = "" if not st else f"  [{st['particle_calls']} EnergyOfParticle {1e6 * st['particle_seconds'] / max(1, st['particle_calls']):.1f} us each, {st['box_calls']} BoxEnergy {1e6 * st['box_seconds'] / max(1, st['box_calls']):.0f} us each]"
        print(f"CorrectEnergy {'as shipped' if correct else 'off'}, {'bound to libmcpm.so' if bound else 'stock reference (CPU)'}: {1e6 * dt / steps:.1f} us per step, N = {r.state()['n']}{extra}", flush=True)
        r.close()
