"""GPU: K1 (EnergyOfParticle) and K3 (BoxEnergy) through the C ABI against the CPU oracle and the
golden vectors of the compiled reference.  Tolerance: 1e-10 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import O, rel_err, system_from_meta

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def close(a, b, rtol=RTOL):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= rtol * np.abs(b) + 1e-300)


@pytest.mark.parametrize("name", ["c1", "c1raw", "ch4", "bulk"])
def test_k3_box_energy_golden(mcpm, golden, name):
    arrays, meta = golden
    s = system_from_meta(meta[name]["system"])
    x, y, z = arrays[name + "_x"], arrays[name + "_y"], arrays[name + "_z"]
    with mcpm.Engine(1, len(x) + 8) as e:
        e.set_system(s)
        e.set_positions(0, x, y, z)
        box, pp, pw = e.box_energy(0)
    assert close(box, arrays[name + "_box"])
    assert close(pp, arrays[name + "_pp"])
    assert close(pw, arrays[name + "_pw"])


@pytest.mark.parametrize("name", ["c1", "c1raw", "ch4", "bulk"])
def test_k1_particle_energy_golden(mcpm, golden, name):
    arrays, meta = golden
    s = system_from_meta(meta[name]["system"])
    x, y, z = arrays[name + "_x"], arrays[name + "_y"], arrays[name + "_z"]
    with mcpm.Engine(1, len(x) + 8) as e:
        e.set_system(s)
        e.set_positions(0, x, y, z)
        for k, (i, t) in enumerate(zip(arrays[name + "_idx"], arrays[name + "_trials"])):
            cur, moved = e.particle_energy(0, int(i), t)
            assert close(cur, arrays[name + "_cur"][k])
            assert close(moved, arrays[name + "_moved"][k])
            _, ghost = e.particle_energy(0, -1, t)
            assert close(ghost, arrays[name + "_ghost"][k])
            cur2, _ = e.particle_energy(0, int(i))
            assert np.array_equal(cur2, cur)


def test_known_answers(mcpm, golden):
    _, meta = golden
    ka = meta["known_answers"]
    r0 = 2.0 ** (1.0 / 6.0) * 3.405
    with mcpm.Engine(1, 32) as e:
        e.set_system(O.argon_slit())
        e.set_positions(0, [10.0, 10.0 + r0], [10.0, 10.0], [10.0, 10.0])
        assert e.particle_energy(0, 0)[0][0] == pytest.approx(-119.8, rel=1e-12)
        for zv, key in ((3.4025, "wall_z_sigma"), (10.0, "wall_z_center"), (20.5, "wall_outside")):
            e.set_positions(0, [10.0], [10.0], [zv])
            assert e.particle_energy(0, 0)[0][2] == pytest.approx(ka[key], rel=RTOL)
            assert e.box_energy(0)[0][2] == pytest.approx(ka[key], rel=RTOL)


def random_config(system, n, seed, spill=0.0):
    rng = np.random.default_rng(seed)
    w = np.array(system.width)
    p = rng.random((n, 3)) * w
    if spill:
        p[:, :2] += (rng.random((n, 2)) - 0.5) * spill * w[:2]   # lateral coordinates outside [0,L): general minimum image
    return p[:, 0].copy(), p[:, 1].copy(), p[:, 2].copy()


@pytest.mark.parametrize("system,n,seed,spill", [
    (O.argon_slit(), 1, 1, 0.0),
    (O.argon_slit(), 2, 2, 0.0),
    (O.argon_slit(), 33, 3, 0.0),
    (O.argon_slit(), 700, 4, 0.0),
    (O.argon_slit(), 500, 5, 6.0),                       # positions up to 3 box lengths away
    (O.argon_slit(H=9.0, rcut=14.0, T=77.0), 257, 6, 0.0),
    (O.methane_slit(rcut=46.0), 2500, 7, 0.0),           # multi-tile K3, multi-CTA K1
    (O.argon_bulk(size=40.0, rcut=12.0), 1300, 8, 0.0),
    (O.argon_bulk(size=40.0, rcut=12.0), 600, 9, 4.0),
    (O.make_system(geometry=1, wall_kind=0, n_layers=0, n_disp=1, n_vol=0, n_swap=0, n_equil_sets=0,
                   width=(30.0, 30.0, 15.0), temperature=100.0, eps_ff=100.0, sig_ff=3.0, rcut=15.0,
                   molar_mass=40.0, mu=0.0, eps_sf=0.0, sig_sf=0.0, rho_s=0.0, delta=0.0, dr=1.0), 300, 10, 0.0),
    (O.make_system(geometry=1, wall_kind=1, n_layers=7, n_disp=1, n_vol=0, n_swap=1, n_equil_sets=0,
                   width=(30.0, 30.0, 25.0), temperature=100.0, eps_ff=100.0, sig_ff=3.0, rcut=15.0,
                   molar_mass=40.0, mu=-1000.0, eps_sf=50.0, sig_sf=3.2, rho_s=0.38, delta=3.0, dr=1.0), 300, 11, 0.0),
])
def test_k1_k3_vs_oracle_random(mcpm, system, n, seed, spill):
    x, y, z = random_config(system, n, seed, spill)
    c = O.COracle(system, n + 8)
    c.set_positions(x, y, z)
    obox, opp, opw = c.box_energy()
    rng = np.random.default_rng(seed + 100)
    with mcpm.Engine(2, n + 8) as e:
        e.set_system(system)
        e.set_positions(1, x, y, z)          # chain 1: exercises the chain offset
        e.set_positions(0, x[:1], y[:1], z[:1])
        box, pp, pw = e.box_energy(1)
        assert close(box, obox) and close(pp, opp) and close(pw, opw)
        allbox = e.box_energy_all()
        assert close(allbox[1], obox)
        for _ in range(6):
            i = int(rng.integers(0, n))
            t = rng.random(3) * np.array(system.width)
            if spill:
                t[:2] += 2.5 * np.array(system.width)[:2]
            cur, moved = e.particle_energy(1, i, t)
            assert close(cur, c.particle_energy(i))
            assert close(moved, c.moved_energy(i, *t))
            assert close(e.particle_energy(1, -1, t)[1], c.trial_energy(*t))


def test_empty_chain_and_bad_arguments(mcpm):
    s = O.argon_slit()
    with mcpm.Engine(1, 64) as e:
        with pytest.raises(mcpm.McpmError):
            e.set_positions(0, [1.0], [1.0], [1.0])           # no system yet
        e.set_system(s)
        e.set_positions(0, [], [], [])
        box, pp, pw = e.box_energy(0)
        assert np.all(box == 0.0) and len(pp) == 0
        _, ghost = e.particle_energy(0, -1, [5.0, 5.0, 10.0])
        assert ghost[0] == 0.0 and ghost[2] == pytest.approx(O.COracle(s, 8).slit_wall(10.0), rel=RTOL)
        with pytest.raises(mcpm.McpmError):
            e.set_positions(0, np.zeros(100), np.zeros(100), np.zeros(100))   # over capacity
        with pytest.raises(mcpm.McpmError):
            e.particle_energy(0, 5)
        bad = O.argon_slit(); bad.n_vol = 1
        with pytest.raises(mcpm.McpmError):
            e.set_system(bad)                                   # volume moves need a bulk box and a pressure
        bad = O.argon_slit(); bad.geometry = 4
        with pytest.raises(mcpm.McpmError):
            e.set_system(bad)


def test_state_mutation_matches_reference_semantics(mcpm):
    """append / remove_swap_last / apply_move mirror src/moves.cpp:302-311, :329-332, :163."""
    s = O.argon_slit()
    x, y, z = random_config(s, 50, 21)
    c = O.COracle(s, 128)
    c.set_positions(x, y, z)
    with mcpm.Engine(1, 128) as e:
        e.set_system(s)
        e.set_positions(0, x, y, z)
        e.append(0, [1.0, 2.0, 3.0])
        e.remove_swap_last(0, 7)
        e.apply_move(0, 3, [4.0, 5.0, 6.0])
        gx, gy, gz = e.get_positions(0)
    ex = np.append(x, 1.0); ey = np.append(y, 2.0); ez = np.append(z, 3.0)
    ex[7], ey[7], ez[7] = ex[-1], ey[-1], ez[-1]
    ex, ey, ez = ex[:-1], ey[:-1], ez[:-1]
    ex[3], ey[3], ez[3] = 4.0, 5.0, 6.0
    assert np.array_equal(gx, ex) and np.array_equal(gy, ey) and np.array_equal(gz, ez)


@pytest.mark.parametrize("system,n,seed", [
    (O.argon_bulk(size=40.0, rcut=12.0), 9000, 31),                       # 3x3x3 cells: every neighbour cell distinct
    (O.argon_bulk(size=100.0, rcut=12.0), 20000, 32),                     # config C4: 8x8x8 cells
    (O.make_system(geometry=1, wall_kind=1, n_layers=3, n_disp=1, n_vol=0, n_swap=1, n_equil_sets=0,
                   width=(50.0, 62.0, 30.0), temperature=120.0, eps_ff=119.8, sig_ff=3.405, rcut=12.0,
                   molar_mass=39.948, mu=-1200.0, eps_sf=57.92, sig_sf=3.4025, rho_s=0.382, delta=3.35, dr=1.0), 8500, 33),
])
def test_k3_cell_list_matches_all_pairs_and_oracle(mcpm, system, n, seed):
    """K3 through the cell list == K3 all-pairs == MC::BoxEnergy (oracle), per particle and in total."""
    x, y, z = random_config(system, n, seed)
    with mcpm.Engine(1, n + 32) as e:
        e.set_system(system)
        e.set_positions(0, x, y, z)
        e.use_cell_list(True)
        box_c, pp_c, pw_c = e.box_energy(0)
        ms_cells = e.last_kernel_ms()
        e.use_cell_list(False)
        box_a, pp_a, pw_a = e.box_energy(0)
        ms_all = e.last_kernel_ms()
    assert close(box_c, box_a) and close(pp_c, pp_a) and close(pw_c, pw_a)
    c = O.COracle(system, n + 8)           # N = 20 000: 4e8 evaluations, ~5 s of CPU
    c.set_positions(x, y, z)
    obox, opp, opw = c.box_energy()
    assert close(box_c, obox) and close(pp_c, opp) and close(pw_c, opw)
    print(f"n={n}: cell-list K3 {ms_cells:.3f} ms, all-pairs K3 {ms_all:.3f} ms")
    if n >= 20000:
        assert ms_cells < ms_all


@pytest.mark.parametrize("n,outside", [(1, False), (333, False), (1024, False), (1025, False), (2600, False), (5400, False), (700, True)])
def test_k1_batch_matches_oracle_and_single_calls(mcpm, n, outside):
    """mcpm_particle_energy_batch: a mixed batch (stored + trial, stored only, ghost) in one launch — one CTA slice of 1024 staged
    particles up to six of them — against the oracle, and bit-identical to the one-request launches (same per-thread partner
    order, same reduction order).  `outside` puts coordinates and trials outside the box: the general minimum image."""
    s = O.methane_slit(rcut=40.0, n_equil_sets=0)
    rng = np.random.default_rng(n)
    x, y, z = rng.random(n) * 80.0, rng.random(n) * 80.0, 3.0 + rng.random(n) * 14.0
    if outside:
        x += 80.0 * rng.integers(-2, 3, size=n)
    o = O.COracle(s, n + 8)
    o.set_positions(x, y, z)
    m = 37
    idx = rng.integers(-1, n, size=m).astype(np.int32)
    trial = np.column_stack([rng.random(m) * 80.0 + (160.0 if outside else 0.0), rng.random(m) * 80.0, 2.0 + rng.random(m) * 16.0])
    with mcpm.Engine(1, n + 40) as e:
        e.set_system(s)
        e.set_positions(0, x, y, z)
        cur, moved = e.particle_energy_batch(0, idx, trial)
        for r in range(m):
            i = int(idx[r])
            want_moved = o.moved_energy(i, *trial[r]) if i >= 0 else o.trial_energy(*trial[r])
            assert close(moved[r], want_moved), (r, i)
            if i >= 0:
                assert close(cur[r], o.particle_energy(i)), (r, i)
            c1, m1 = e.particle_energy(0, i, trial[r])
            assert np.array_equal(m1, moved[r]) and (i < 0 or np.array_equal(c1, cur[r]))
        stored = np.flatnonzero(idx >= 0)
        if len(stored):
            cur2, _ = e.particle_energy_batch(0, idx[stored])           # no trial positions: the stored particles where they are
            assert np.array_equal(cur2, cur[stored])
        big = np.arange(2500) % max(n, 1)                              # more requests than one launch takes (1024)
        cur3, _ = e.particle_energy_batch(0, big.astype(np.int32))
        assert np.array_equal(cur3[:n][: min(n, 50)], cur3[n:2 * n][: min(n, 50)]) if 2 * n <= 2500 else True
        assert close(cur3[0], o.particle_energy(0))


@pytest.mark.parametrize("n,outside,cap", [(1, False, 64), (333, False, 512), (2600, False, 4096), (9000, False, 12000), (700, True, 1024)])
def test_k1_resident_service_matches_oracle_and_launches(mcpm, n, outside, cap):
    """mcpm_k1_resident: the persistent one-CTA kernel (positions staged once in shared memory — or served from global memory when the
    capacity does not fit one SM: cap = 12000 — requests through a mailbox in mapped host memory) gives what the per-call launches and the
    oracle give, follows a host-driven move loop (apply_move / append / remove_swap_last through the same mailbox, written through to
    global memory), survives its idle time-out, and steps aside for every other entry point."""
    import time
    s = O.methane_slit(rcut=40.0, n_equil_sets=0)
    rng = np.random.default_rng(100 + n)
    x, y, z = rng.random(n) * 80.0, rng.random(n) * 80.0, 3.0 + rng.random(n) * 14.0
    if outside:
        x += 80.0 * rng.integers(-2, 3, size=n)
    o = O.COracle(s, cap)
    o.set_positions(x, y, z)
    m = 25
    idx = rng.integers(-1, n, size=m)
    trial = np.column_stack([rng.random(m) * 80.0 + (160.0 if outside else 0.0), rng.random(m) * 80.0, 2.0 + rng.random(m) * 16.0])
    with mcpm.Engine(1, cap) as e:
        e.set_system(s)
        e.set_positions(0, x, y, z)
        launched = [e.particle_energy(0, int(i), t) for i, t in zip(idx, trial)]
        l0 = e.launch_count()
        e.k1_resident(0, True)
        for r in range(m):
            i = int(idx[r])
            cur, moved = e.particle_energy(0, i, trial[r])
            assert close(moved, launched[r][1], 1e-11) and (i < 0 or close(cur, launched[r][0], 1e-11))
            assert close(moved, o.moved_energy(i, *trial[r]) if i >= 0 else o.trial_energy(*trial[r]))
        starts, requests = e.k1_resident_stats()
        assert requests == m and 1 <= starts <= 2 and e.launch_count() == l0 + starts     # ONE launch for all the calls (two if the host was descheduled past the idle time-out)
        # a host-driven move loop: accepted translations, insertions and deletions mirrored through the mailbox
        px, py, pz = x.copy(), y.copy(), z.copy()
        for k in range(30):
            kind = k % 3 if len(px) > 1 else 1
            if kind == 0:
                i = int(rng.integers(0, len(px))); t = np.array([rng.random() * 80.0, rng.random() * 80.0, 3.0 + rng.random() * 14.0])
                e.apply_move(0, i, t); px[i], py[i], pz[i] = t
            elif kind == 1:
                t = np.array([rng.random() * 80.0, rng.random() * 80.0, 3.0 + rng.random() * 14.0])
                e.append(0, t); px = np.append(px, t[0]); py = np.append(py, t[1]); pz = np.append(pz, t[2])
            else:
                i = int(rng.integers(0, len(px)))
                e.remove_swap_last(0, i)
                px[i], py[i], pz[i] = px[-1], py[-1], pz[-1]; px, py, pz = px[:-1], py[:-1], pz[:-1]
            o.set_positions(px, py, pz)
            j = int(rng.integers(0, len(px))); t = np.array([rng.random() * 80.0, rng.random() * 80.0, 2.0 + rng.random() * 16.0])
            cur, moved = e.particle_energy(0, j, t)
            assert close(cur, o.particle_energy(j)) and close(moved, o.moved_energy(j, *t)), k
        if outside:
            assert e.k1_resident_stats()[0] <= 3                                          # the general service took everything: no fall-back launches
        time.sleep(0.05)                                                                  # longer than the idle time-out: the kernel has left
        starts = e.k1_resident_stats()[0]
        cur, _ = e.particle_energy(0, 0)
        assert close(cur, o.particle_energy(0)) and e.k1_resident_stats()[0] == starts + 1     # restarted on demand
        # other entry points see the current configuration (write-through) and quiesce the service
        gx, gy, gz = e.get_positions(0)
        assert np.array_equal(gx, px) and np.array_equal(gy, py) and np.array_equal(gz, pz)
        e.particle_energy(0, 0)                                                           # (get_positions has quiesced the service: back on)
        l1 = e.launch_count()
        box, pp, pw = e.box_energy(0)                                                     # up to 1536 particles the service answers BoxEnergy itself
        ob, opp, opw = o.box_energy()
        assert close(box, ob) and close(pp, opp) and close(pw, opw)
        assert (e.launch_count() == l1) == (len(px) <= 1536)
        cur, _ = e.particle_energy(0, 0)                                                  # and the service is there (or comes back) afterwards
        assert close(cur, o.particle_energy(0))
        if not outside and n > 1:
            # a trial position outside the box cannot go to a service built on the fast minimum image: that request is launched instead
            far = np.array([250.0, 40.0, 10.0])
            cur, moved = e.particle_energy(0, 0, far)
            assert close(moved, o.moved_energy(0, *far))
        st = e.run(1, 1, 200, seed=3, check_energy=1)                                     # K2 on the same context
        assert st[0].energy == pytest.approx(st[0].energy_check, rel=1e-9)
        e.k1_resident(0, False)
