"""GPU: the COMPILED drop-in at the energy seam.

oracle/_ref/libmcref_bound.so is the unmodified reference (MC.cpp moves.cpp read.cpp print.cpp sample.cpp potentials.cpp) with
src/energy.cpp replaced by oracle/energy_bound.cpp, whose MC::EnergyOfParticle / MC::BoxEnergy call libmcpm.so (K1 / K3) and mirror the
reference's accepted moves with mcpm_apply_move / mcpm_append / mcpm_remove_swap_last.  The reference's OWN MoveParticle /
SwapParticle / AdjustMCMoves, its own mt19937_64 stream and bookkeeping then have to reproduce the golden traces of the stock build.
"""
import time

import numpy as np
import pytest

from conftest import O, system_from_meta

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bound():
    if not O.have_bound_reference_lib():
        pytest.skip("oracle/_ref/libmcref_bound.so not built (make -C oracle ref_bound; needs /root/reference)")
    import psutil
    if psutil.virtual_memory().available < 8 * 2 ** 30:
        pytest.skip("not enough free memory for the reference's 5.6 GB arrays")
    r = O.RefOracle(bound=True)
    assert r.bound_stats() is not None
    yield r
    r.close()


def drive(ref, system, x, y, z, seed, sets, steps_per_set, correct=0):
    """The recipe of tests/golden/make_golden.py::trace_case, call for call."""
    ref.setup(system)
    ref.set_positions(x, y, z)
    ref.seed(seed)
    ref.box_energy()
    trace, per_set = [], []
    for s in range(1, sets + 1):
        ref.reset_stats()
        _, tr = ref.run(steps_per_set, correct, True)
        trace.append(tr)
        st = ref.state()
        ref.adjust(s)
        per_set.append([st["n"], st["acceptance"], st["rejection"], st["n_displacements"], st["accept_swap"],
                        st["reject_swap"], st["n_swaps"], ref.state()["dr"]])
    return np.concatenate(trace), np.array(per_set, dtype=np.float64)


@pytest.mark.parametrize("name,mkey,start", [("c1t", "c1_trace", "c1"), ("lowt", "low_trace", "c1"), ("bulkt", "bulk_trace", "bulk")])
def test_reference_moves_on_gpu_energies_reproduce_the_golden_traces(bound, golden, name, mkey, start):
    arrays, meta = golden
    m = meta[mkey]
    s = system_from_meta(m["system"])
    n0 = m["n0"]
    x, y, z = arrays[start + "_x"][:n0], arrays[start + "_y"][:n0], arrays[start + "_z"][:n0]
    t0 = time.perf_counter()
    trace, per_set = drive(bound, s, x, y, z, m["seed"], m["sets"], m["steps_per_set"])
    dt = time.perf_counter() - t0
    d = np.nonzero(trace != arrays[name + "_trace"])[0]
    assert len(d) == 0, f"first divergence at step {d[0]}"
    assert np.array_equal(per_set, arrays[name + "_per_set"])
    fx, fy, fz = bound.get_positions()
    assert np.array_equal(fx, arrays[name + "_fx"]) and np.array_equal(fy, arrays[name + "_fy"]) and np.array_equal(fz, arrays[name + "_fz"])
    box, _, _ = bound.box_energy()                      # K3 on the device copy == the stock build's BoxEnergy of the same configuration
    assert np.allclose(box, arrays[name + "_fbox"], rtol=1e-10, atol=0)
    st = bound.bound_stats()
    steps = m["sets"] * m["steps_per_set"]
    assert st["particle_calls"] >= steps and st["uploads"] <= 2          # state followed by mirroring, not by re-uploads
    print(f"{name}: {steps} steps of the reference's own moves in {dt:.2f} s; {st['particle_calls']} EnergyOfParticle calls through "
          f"mcpm_particle_energy at {1e6 * st['particle_seconds'] / st['particle_calls']:.1f} us each (N0 = {n0}), "
          f"{st['mirrored']} slots mirrored, {st['uploads']} uploads")


def test_bound_build_with_correctenergy_as_shipped(bound, golden):
    """The as-shipped loop: CorrectEnergy's ratio test fires MC::BoxEnergy (K3 through the binding) every few steps and rewrites
    every particle record; the chain must not notice."""
    arrays, meta = golden
    m = meta["c1_trace"]
    s = system_from_meta(m["system"])
    trace, per_set = drive(bound, s, arrays["c1_x"], arrays["c1_y"], arrays["c1_z"], m["seed"], 1, 1500, correct=1)
    assert np.array_equal(trace, arrays["c1t_trace"][:1500])
    assert bound.bound_stats()["box_calls"] > 10


def test_bound_build_energies_equal_the_stock_build(bound, golden):
    arrays, meta = golden
    for case in ("c1", "ch4", "bulk"):
        s = system_from_meta(meta[case]["system"])
        bound.setup(s)
        bound.set_positions(arrays[case + "_x"], arrays[case + "_y"], arrays[case + "_z"])
        box, pp, pw = bound.box_energy()
        assert np.allclose(box, arrays[case + "_box"], rtol=1e-10, atol=0)
        assert np.allclose(pp, arrays[case + "_pp"], rtol=1e-10, atol=0) and np.allclose(pw, arrays[case + "_pw"], rtol=1e-10, atol=0)
        for i in (0, 7, len(pp) - 1):
            assert np.allclose(bound.particle_energy(i), [pp[i], 0.0, pw[i], pp[i] + pw[i]], rtol=1e-10, atol=0)
