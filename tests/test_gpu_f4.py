"""GPU: row f4 of SURVEY.md §8 — spherical pore (SphericalLJ_Pot, sphere insertion and out-of-pore guard), hard spheres (HardSphere_Pot) and
NPT volume moves (MC::ChangeVolume) — through the C ABI against golden vectors of the compiled reference (tests/golden/reference_f4.npz;
the C oracle is pinned to the same vectors in tests/test_oracle_f4.py).  These systems run through the GENERAL kernels."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, O

pytestmark = pytest.mark.gpu
RTOL = 1e-10


@pytest.fixture(scope="module")
def f4():
    """reference_f4 (sphere, hard spheres, NPT) and reference_cyl (cylindrical walls) merged."""
    arrays = dict(np.load(os.path.join(GOLDEN, "reference_f4.npz")))
    arrays.update(dict(np.load(os.path.join(GOLDEN, "reference_cyl.npz"))))
    meta = json.load(open(os.path.join(GOLDEN, "reference_f4.json")))
    meta.update(json.load(open(os.path.join(GOLDEN, "reference_cyl.json"))))
    return arrays, meta


def close(a, b, rtol=RTOL):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return bool(np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(b), np.mean(np.abs(b)) + 1e-300)))


@pytest.mark.parametrize("name", ["sph", "hs", "hsov", "cyl3", "cyl4"])
def test_k1_k3_energies(mcpm, f4, name):
    arrays, meta = f4
    s = O.make_system(**meta[name]["system"])
    x, y, z = arrays[name + "_x"], arrays[name + "_y"], arrays[name + "_z"]
    with mcpm.Engine(1, len(x) + 8) as e:
        e.set_system(s)
        e.set_positions(0, x, y, z)
        box, pp, pw = e.box_energy(0)
        assert close(box, arrays[name + "_box"]) and close(pp, arrays[name + "_pp"]) and close(pw, arrays[name + "_pw"])
        for k, (i, t) in enumerate(zip(arrays[name + "_idx"], arrays[name + "_trials"])):
            cur, moved = e.particle_energy(0, int(i), t)
            assert close(cur, arrays[name + "_cur"][k]) and close(moved, arrays[name + "_moved"][k]), k
            _, ghost = e.particle_energy(0, -1, t)
            assert close(ghost, arrays[name + "_ghost"][k]), k
        cur_b, moved_b = e.particle_energy_batch(0, arrays[name + "_idx"].astype(np.int32), arrays[name + "_trials"])
        assert close(cur_b, arrays[name + "_cur"]) and close(moved_b, arrays[name + "_moved"])


def test_unsupported_systems_are_refused(mcpm):
    with mcpm.Engine(1, 64) as e:
        for change in ({"n_vol": 1}, {"n_vol": 1, "pressure": 1e7, "geometry": 1}, {"n_vol": 1, "pressure": 1e7, "n_swap": 1}, {"wall_kind": 5},
                       {"pair_kind": 2}, {"geometry": 4}):
            s = O.argon_bulk(size=30.0, rcut=12.0)
            s.n_swap = 0
            for k, v in change.items():
                setattr(s, k, v)
            with pytest.raises(mcpm.McpmError) as ei:
                e.set_system(s)
            assert ei.value.code in (1, 4), change


@pytest.mark.parametrize("threads", [0, 64, 256])
@pytest.mark.parametrize("name,mkey,start", [("spht", "sph_trace", "sph"), ("hst", "hs_trace", "hs"), ("hsgt", "hsg_trace", "hs"), ("nptt", "npt_trace", None),
                                              ("cyl3t", "cyl3_trace", "cyl3"), ("cyl4t", "cyl4_trace", "cyl4")])
def test_replay_traces_match_the_reference(mcpm, f4, golden, name, mkey, start, threads):
    """The reference's own mt19937_64 stream replayed on the GPU: identical accept/reject sequence (translations, insertions through the
    sphere branch of InsertParticle, hard-sphere deletions that are always accepted, volume moves against the as-shipped running energy),
    per-set counters, dr and dv adaptation, box and final configuration."""
    arrays, meta = f4
    m = meta[mkey]
    s = O.make_system(**m["system"])
    if start is None:
        g, _ = golden
        x, y, z = g["bulk_x"], g["bulk_y"], g["bulk_z"]
    else:
        x, y, z = arrays[start + "_x"], arrays[start + "_y"], arrays[start + "_z"]
    sets, sps = m["sets"], m["steps_per_set"]
    u = O.mt_uniforms(m["seed"], 8 * sets * sps)
    with mcpm.Engine(1, 1024) as e:
        e.set_system(s)
        e.set_positions(0, x, y, z)
        traces, per_set, vol = [], [], []
        pos = 0
        for k in range(1, sets + 1):
            stats, tr = e.run(k, 1, sps, replay=u[pos:], trace=True, threads=threads, check_energy=1, no_sampling=1)
            st = stats[0]
            pos += int(st.replay_consumed)
            traces.append(tr[0])
            per_set.append([st.n, st.acceptance, st.rejection, st.n_displacements, st.accept_swap, st.reject_swap, st.n_swaps, st.dr])
            vol.append([st.volume, st.width, st.dv, st.accept_vol, st.reject_vol, st.n_vol_changes, st.energy_as_shipped])
            assert st.energy == pytest.approx(st.energy_check, rel=RTOL)
        assert e.last_kernel_name() == "k2_chains"
        got = np.concatenate(traces)
        d = np.nonzero(got != arrays[name + "_trace"])[0]
        assert len(d) == 0, f"first divergence at step {d[0]}"
        assert np.array_equal(np.array(per_set, dtype=np.float64), arrays[name + "_per_set"])
        if name == "nptt":
            v, rv = np.array(vol), arrays[name + "_vol"]
            assert np.array_equal(v[:, 3:6], rv[:, 3:6])                                   # acceptanceVol, rejectionVol, nVolChanges
            assert np.allclose(v[:, :3], rv[:, :3], rtol=1e-13, atol=0)                    # volume, width, dv (device exp/log/cbrt: last-ulp differences)
            assert np.allclose(v[:, 6], rv[:, 6], rtol=1e-9, atol=0)                       # the reference's running box energy, as shipped
            assert e.get_system(0).width[2] == pytest.approx(rv[-1, 1], rel=1e-13)         # the chain's system follows its box
        fx, fy, fz = e.get_positions(0)
        tol = dict(rtol=1e-13, atol=1e-12) if name in ("spht", "nptt", "cyl3t", "cyl4t") else dict(rtol=0, atol=0)   # device sin / cos (sphere insertions), exp / log / cbrt (box width)
        assert np.allclose(fx, arrays[name + "_fx"], **tol) and np.allclose(fy, arrays[name + "_fy"], **tol) and np.allclose(fz, arrays[name + "_fz"], **tol)
        box, _, _ = e.box_energy(0)
        assert close(box, arrays[name + "_fbox"])


def test_npt_philox_follows_the_oracle(mcpm, golden):
    """Device Philox + NPT in ONE launch of several sets (dv adaptation on the device) vs the C oracle on the same stream."""
    g, _ = golden
    s = O.make_system(geometry=0, wall_kind=0, n_layers=0, n_disp=4, n_vol=1, n_swap=0, n_equil_sets=10, pair_kind=0, width=(30.0, 30.0, 30.0),
                      temperature=120.0, eps_ff=119.8, sig_ff=3.405, rcut=12.0, molar_mass=39.948, mu=0.0, eps_sf=0.0, sig_sf=0.0, rho_s=0.0,
                      delta=0.0, dr=0.5, pressure=3.0e7, dv=0.02)
    x, y, z = g["bulk_x"], g["bulk_y"], g["bulk_z"]
    o = O.COracle(s, 1024)
    o.set_positions(x, y, z); o.box_energy(); o.philox(11, 0, 0)
    otr = o.run_sets(1, 3, 1200, 1, True)                                  # as shipped: CorrectEnergy after every step
    with mcpm.Engine(1, 1024) as e:
        e.set_system(s)
        e.set_positions(0, x, y, z)
        st, tr = e.run(1, 3, 1200, seed=11, trace=True, check_energy=1)
    d = np.nonzero(tr[0] != otr)[0]
    assert len(d) == 0, f"first divergence at step {d[0]}"
    vs = o.volume_state()
    assert (st[0].accept_vol, st[0].reject_vol, st[0].n_vol_changes) == (vs["acceptance_vol"], vs["rejection_vol"], vs["n_vol_changes"])
    assert st[0].volume == pytest.approx(vs["volume"], rel=1e-13) and st[0].dv == pytest.approx(vs["dv"], rel=1e-13)
    assert st[0].energy == pytest.approx(st[0].energy_check, rel=RTOL)
    assert st[0].energy == pytest.approx(o.state()["exact_energy"], rel=1e-9)
    assert (tr[0] >> 1 == 4).sum() > 400 and st[0].accept_vol > 20
