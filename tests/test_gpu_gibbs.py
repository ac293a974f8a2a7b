"""GPU: two linked boxes (Gibbs ensemble; row f4 of SURVEY.md §8) through the C ABI — MoveParticle's box selection and SwapParticle's
Gibbs branch (src/moves.cpp:127-137, 342-400) in k2_gibbs — against golden vectors of the compiled reference
(tests/golden/reference_gibbs.npz; the C oracle is pinned to the same vectors in tests/test_oracle_gibbs.py) and, on the device's own
Philox stream, against the C oracle."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gibbs():
    return dict(np.load(os.path.join(GOLDEN, "reference_gibbs.npz"))), json.load(open(os.path.join(GOLDEN, "reference_gibbs.json")))


def code_mask(trace):
    """The direction of a transfer attempt whose donor box is empty (kind 6) cannot be observed in the reference: no box bit there."""
    return np.where(((trace >> 1) & 7) == 6, 0x0F, 0xFF).astype(np.uint8)


def linked_engine(mcpm, arrays, m, name, cap=1024):
    e = mcpm.Engine(2, cap)
    e.set_system(O.make_system(**m["box0"]), 0)
    e.set_system(O.make_system(**m["box1"]), 1)
    for ib in (0, 1):
        e.set_positions(ib, arrays[f"{name}_x{ib}"], arrays[f"{name}_y{ib}"], arrays[f"{name}_z{ib}"])
    e.link_boxes(0, 1)
    return e


@pytest.mark.parametrize("name", ["vle", "pore"])
def test_replay_traces_match_the_reference(mcpm, gibbs, name):
    """The reference's own mt19937_64 stream replayed on the GPU: the same boxes chosen, the same transfers attempted and accepted, the same
    per-box counters, dr adaptation, Widom/BAR sums and final configurations of both boxes."""
    arrays, meta = gibbs
    m = meta[name]
    sets, sps = m["sets"], m["steps_per_set"]
    u = O.mt_uniforms(m["seed"], 8 * sets * sps)
    with linked_engine(mcpm, arrays, m, name) as e:
        for ib in (0, 1):
            box, _, _ = e.box_energy(ib)
            assert np.allclose(box, arrays[f"{name}_box{ib}"], rtol=1e-10, atol=0)
        traces, per_set = [], []
        pos = 0
        for k in range(1, sets + 1):
            rep = np.zeros((2, len(u) - pos)); rep[0] = u[pos:]
            stats, tr = e.run(k, 1, sps, replay=rep, trace=True, check_energy=1)
            pos += int(stats[0].replay_consumed)
            traces.append(tr[0])
            row = []
            for ib in (0, 1):
                st = stats[ib]
                assert st.energy == pytest.approx(st.energy_check, rel=1e-9, abs=1e-6)
                row += [st.n, st.acceptance, st.rejection, st.n_displacements, st.dr, st.widom_insert, st.widom_delete, st.widom_insertions, st.widom_deletions]
            row += [stats[0].accept_swap, stats[0].reject_swap, stats[0].n_swaps]
            per_set.append(row)
        assert e.last_kernel_name() == "k2_gibbs"
        got, ref = np.concatenate(traces), arrays[name + "_trace"]
        mask = code_mask(ref)
        d = np.nonzero((got & mask) != (ref & mask))[0]
        assert len(d) == 0, f"first divergence at step {d[0]}: {got[d[0]]} vs {ref[d[0]]}"
        got_sets, want = np.array(per_set, dtype=np.float64), arrays[name + "_per_set"]
        exact = [0, 1, 2, 3, 4, 7, 8, 9, 10, 11, 12, 13, 16, 17, 18, 19, 20]          # counters and dr: bit for bit
        assert np.array_equal(got_sets[:, exact], want[:, exact])
        sums = [5, 6, 14, 15]                                                          # Widom sums: device exp / cosh
        assert np.allclose(got_sets[:, sums], want[:, sums], rtol=1e-10, atol=0, equal_nan=True)
        for ib in (0, 1):
            fx, fy, fz = e.get_positions(ib)
            assert np.array_equal(fx, arrays[f"{name}_fx{ib}"]) and np.array_equal(fy, arrays[f"{name}_fy{ib}"]) and np.array_equal(fz, arrays[f"{name}_fz{ib}"])
            box, _, _ = e.box_energy(ib)
            assert np.allclose(box, arrays[f"{name}_fbox{ib}"], rtol=1e-10, atol=1e-9)


def test_philox_run_follows_the_oracle(mcpm, gibbs):
    """Several sets in ONE launch on the device's Philox stream (dr adaptation of both boxes on the device) vs the C oracle on the same stream;
    mu of both boxes as MC::ComputeChemicalPotential forms it from the transfer's Widom sums."""
    arrays, meta = gibbs
    m = meta["vle"]
    a, b = O.make_system(**m["box0"]), O.make_system(**m["box1"])
    oa, ob = O.COracle(a, 1024), O.COracle(b, 1024)
    oa.link(ob)
    for ib, o in ((0, oa), (1, ob)):
        o.set_positions(arrays[f"vle_x{ib}"], arrays[f"vle_y{ib}"], arrays[f"vle_z{ib}"]); o.box_energy()
    oa.philox(2026, 0, 0)
    otr = oa.gibbs_run_sets(1, 4, 3000, 0, True)
    with linked_engine(mcpm, arrays, m, "vle") as e:
        stats, tr = e.run(1, 4, 3000, seed=2026, trace=True, check_energy=1)
        d = np.nonzero(tr[0] != otr)[0]
        assert len(d) == 0, f"first divergence at step {d[0]}"
        n_total = 0
        for ib, o in ((0, oa), (1, ob)):
            so, st = o.state(), stats[ib]
            n_total += st.n
            assert (st.n, st.acceptance, st.n_displacements) == (so["n"], so["acceptance"], so["n_displacements"])
            assert st.dr == so["dr"]
            assert st.energy == pytest.approx(so["exact_energy"], rel=1e-9) and st.energy == pytest.approx(st.energy_check, rel=1e-9)
            assert st.samples == 2 * 3000 and st.sum_n == pytest.approx(so["sum_n"], rel=1e-12)
            w, wi = o.get_widom()
            assert (st.widom_insertions, st.widom_deletions) == (wi[0], wi[1])
            assert st.widom_insert == pytest.approx(w[0], rel=1e-10) and st.widom_delete == pytest.approx(w[1], rel=1e-10)
            mu_ex, mu = e.chemical_potential(ib)
            omu_ex, omu = o.chem_pot()
            assert np.allclose([mu_ex, mu], [omu_ex, omu], rtol=1e-9, atol=1e-9, equal_nan=True)
        assert n_total == len(arrays["vle_x0"]) + len(arrays["vle_x1"])                # transfers conserve the particle count
        assert (stats[0].accept_swap, stats[0].n_swaps) == (oa.state()["accept_swap"], oa.state()["n_swaps"])
        assert stats[1].accept_swap == stats[0].accept_swap


def test_many_pairs_are_independent_systems(mcpm, gibbs):
    """64 linked pairs in one launch: pair k draws from box 0's stream id; pair 0 equals the single-pair run, and the pairs differ."""
    arrays, meta = gibbs
    m = meta["pore"]
    with linked_engine(mcpm, arrays, m, "pore") as e:
        ref_stats, ref_tr = e.run(1, 2, 1500, seed=5, trace=True)
        ref_n = (ref_stats[0].n, ref_stats[1].n)
    npairs = 64
    with mcpm.Engine(2 * npairs, 1024) as e:
        for k in range(npairs):
            for ib in (0, 1):
                e.set_system(O.make_system(**m[f"box{ib}"]), 2 * k + ib)
                e.set_positions(2 * k + ib, arrays[f"pore_x{ib}"], arrays[f"pore_y{ib}"], arrays[f"pore_z{ib}"])
            e.link_boxes(2 * k, 2 * k + 1)
        ids = np.arange(2 * npairs, dtype=np.uint32); ids[0] = 0; ids[2::2] = 1000 + np.arange(npairs - 1)
        e.set_stream_ids(ids)
        stats, tr = e.run(1, 2, 1500, seed=5, trace=True, check_energy=1)
        assert np.array_equal(tr[0], ref_tr[0]) and (stats[0].n, stats[1].n) == ref_n
        assert len({tr[2 * k].tobytes() for k in range(npairs)}) == npairs
        for k in range(npairs):
            assert stats[2 * k].n + stats[2 * k + 1].n == len(arrays["pore_x0"]) + len(arrays["pore_x1"])
            for ib in (0, 1):
                st = stats[2 * k + ib]
                assert st.energy == pytest.approx(st.energy_check, rel=1e-9, abs=1e-6)


def test_link_rules(mcpm, gibbs):
    arrays, meta = gibbs
    m = meta["vle"]
    a, b = O.make_system(**m["box0"]), O.make_system(**m["box1"])
    with mcpm.Engine(3, 256) as e:
        e.set_system(a, 0); e.set_system(b, 1); e.set_system(a, 2)
        with pytest.raises(mcpm.McpmError):
            e.link_boxes(0, 0)
        hot = O.make_system(**{**m["box1"], "temperature": 150.0})
        e.set_system(hot, 1)
        with pytest.raises(mcpm.McpmError):
            e.link_boxes(0, 1)                                   # one ThermodynamicSystem: one temperature
        e.set_system(b, 1)
        e.link_boxes(0, 1)
        with pytest.raises(mcpm.McpmError):
            e.link_boxes(1, 2)                                   # linked already
        with pytest.raises(mcpm.McpmError) as ei:
            e.run(1, 1, 10, seed=1)                              # chain 2 is a single box: not in the same context
        assert ei.value.code == 4


@pytest.mark.parametrize("pore", ["sphere", "cylinder"])
def test_pore_geometries_exchange_with_a_bulk_box(mcpm, pore):
    """Box 0 a spherical cavity (SphericalLJ_Pot, no periodic axis) or a cylindrical pore (CylindricalLJ10_4, periodic along x), box 1 a bulk
    reservoir: transfers insert through InsertParticle's sphere / cylinder branch of the receiving box.  Device Philox vs the C oracle on the
    same stream (the oracle's single-box restatement of these geometries is pinned to the compiled reference in tests/test_oracle_f4.py)."""
    kw = dict(n_layers=1, n_disp=3, n_vol=0, n_swap=1, n_equil_sets=1, pair_kind=0, temperature=87.0, eps_ff=119.8, sig_ff=3.405, rcut=12.0,
              molar_mass=39.948, mu=0.0, eps_sf=57.92, sig_sf=3.4025, rho_s=0.382, delta=0.0, pressure=-1.0, dv=1e-3)
    if pore == "sphere":
        a = O.make_system(geometry=3, wall_kind=2, width=(30.0, 30.0, 30.0), dr=2.0, **kw)
    else:
        a = O.make_system(geometry=2, wall_kind=3, width=(24.0, 26.0, 26.0), dr=2.0, **kw)
    b = O.make_system(geometry=0, wall_kind=0, width=(40.0, 40.0, 40.0), dr=8.0, **{**kw, "eps_sf": 0.0, "sig_sf": 0.0, "rho_s": 0.0})
    rng = np.random.default_rng(17)
    # a loose start inside the pore (within 7 A of the axis / centre) and a dilute reservoir
    pa = np.array(a.width)[:, None] * 0.5 + (rng.random((3, 40)) - 0.5) * 9.0
    if pore == "cylinder":
        pa[0] = rng.random(40) * 24.0
    pb = rng.random((3, 120)) * 40.0
    oa, ob = O.COracle(a, 1024), O.COracle(b, 1024)
    oa.link(ob)
    oa.set_positions(*pa); ob.set_positions(*pb); oa.box_energy(); ob.box_energy()
    oa.philox(77, 0, 0)
    otr = oa.gibbs_run_sets(1, 2, 2500, 0, True)
    with mcpm.Engine(2, 1024) as e:
        e.set_system(a, 0); e.set_system(b, 1)
        e.set_positions(0, *pa); e.set_positions(1, *pb)
        e.link_boxes(0, 1)
        stats, tr = e.run(1, 2, 2500, seed=77, trace=True, check_energy=1)
        d = np.nonzero(tr[0] != otr)[0]
        assert len(d) == 0, f"first divergence at step {d[0]}: {tr[0][d[0]]} vs {otr[d[0]]}"
        assert ((tr[0] >> 1) & 7 == 5).sum() > 500 and (tr[0] & 0x1F == ((5 << 1) | 1)).sum() > 5      # transfers INTO the pore were accepted
        for ib, o in ((0, oa), (1, ob)):
            so, st = o.state(), stats[ib]
            assert (st.n, st.acceptance, st.n_displacements) == (so["n"], so["acceptance"], so["n_displacements"]) and st.dr == so["dr"]
            assert st.energy == pytest.approx(st.energy_check, rel=1e-9, abs=1e-6) and st.energy == pytest.approx(so["exact_energy"], rel=1e-9, abs=1e-6)
            gx, gy, gz = e.get_positions(ib)
            ox, oy, oz = o.get_positions()
            assert np.allclose(gx, ox, rtol=1e-13, atol=1e-12) and np.allclose(gy, oy, rtol=1e-13, atol=1e-12) and np.allclose(gz, oz, rtol=1e-13, atol=1e-12)
