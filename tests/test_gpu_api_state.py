"""GPU: state handling of the C ABI — Philox stream ids across contexts, invalidation in mcpm_set_system, moves-only runs."""
import numpy as np
import pytest

from conftest import O, system_from_meta

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def test_stream_ids_make_chains_independent_of_the_sharding(mcpm, golden):
    """Four chains in ONE context == the same four chains sharded 2 + 2 over two contexts when each context is told the
    chains' global ids; without the ids the second context would repeat the first one's uniforms (round-1 advisor finding)."""
    arrays, meta = golden
    s = system_from_meta(meta["c1"]["system"])
    x, y, z = arrays["c1_x"][:300], arrays["c1_y"][:300], arrays["c1_z"][:300]

    def run(ids, set_ids=True):
        with mcpm.Engine(len(ids), 1024) as e:
            e.set_system(s)
            for c in range(len(ids)):
                e.set_positions(c, x, y, z)
            if set_ids:
                e.set_stream_ids(ids)
            st, tr = e.run(1, 2, 1500, seed=99, trace=True, threads=32)
            return tr.copy(), [e.get_positions(c) for c in range(len(ids))]

    tr_all, pos_all = run([0, 1, 2, 3])
    tr_a, pos_a = run([0, 2])
    tr_b, pos_b = run([1, 3])
    for local, g in enumerate([0, 2]):
        assert np.array_equal(tr_a[local], tr_all[g])
        assert all(np.array_equal(p, q) for p, q in zip(pos_a[local], pos_all[g]))
    for local, g in enumerate([1, 3]):
        assert np.array_equal(tr_b[local], tr_all[g])
        assert all(np.array_equal(p, q) for p, q in zip(pos_b[local], pos_all[g]))
    assert len({tr_all[c].tobytes() for c in range(4)}) == 4               # four different chains
    tr_dup, _ = run([1, 3], set_ids=False)                                 # default ids = local indices 0, 1
    assert np.array_equal(tr_dup[0], tr_all[0]) and np.array_equal(tr_dup[1], tr_all[1])


def test_stream_ids_follow_the_oracle(mcpm, golden):
    arrays, meta = golden
    s = system_from_meta(meta["c1"]["system"])
    x, y, z = arrays["c1_x"], arrays["c1_y"], arrays["c1_z"]
    with mcpm.Engine(1, 1024) as e:
        e.set_system(s)
        e.set_positions(0, x, y, z)
        e.set_stream_ids([4711])
        st, tr = e.run(1, 1, 2000, seed=5, trace=True)
    o = O.COracle(s, 1024)
    o.set_positions(x, y, z); o.box_energy(); o.philox(5, 4711, 0)
    assert np.array_equal(tr[0], o.run_sets(1, 1, 2000, 0, True))


@pytest.mark.parametrize("what", ["sigma", "rcut", "box", "mu"])
def test_set_system_invalidates_stale_energies_and_cache(mcpm, golden, what):
    """Changing an energy-relevant parameter of a chain that holds a configuration must drop the running energies and the
    per-particle energy cache (the configuration checksum cannot see a parameter change); a mu-only update keeps both.
    After the change the chain must follow the oracle started from the same configuration with the new system."""
    arrays, meta = golden
    d = dict(meta["c1"]["system"])
    s1 = system_from_meta(d)
    if what == "sigma":
        d["sig_ff"] = 3.3
    elif what == "rcut":
        d["rcut"] = 15.0
    elif what == "box":
        d["width"] = [36.0, 36.0, 20.0]          # smaller lateral box: some coordinates now lie outside [0, L]
    else:
        d["mu"] = -1250.0
    s2 = system_from_meta(d)
    x, y, z = arrays["c1_x"], arrays["c1_y"], arrays["c1_z"]
    with mcpm.Engine(1, 1024) as e:
        e.set_system(s1)
        e.set_positions(0, x, y, z)
        e.run(1, 1, 3000, seed=1, threads=32)
        e.set_dr(0, 1.7)
        px, py, pz = e.get_positions(0)
        before = e.stats()[0]
        e.set_system(s2)
        assert e.stats()[0].dr == 1.7                                        # same desc.dr: the adapted amplitude stays
        # (the smaller box starts from overlaps across the new boundary, ~1e18 K: running sums that relax from there keep a
        # rounding residue of ~1e3 K in the oracle too, so that case re-anchors once, as the hard-overlap tests do)
        st, tr = e.run(2, 1, 2500, seed=2, trace=True, threads=32, check_energy=2 if what == "box" else 1)
        st2, tr2 = e.run(3, 1, 500, seed=2, trace=True, threads=32, check_energy=1)
        gx, gy, gz = e.get_positions(0)
    o = O.COracle(s2, 1024)
    o.set_positions(px, py, pz); o.box_energy(); o.set_dr(1.7); o.philox(2, 0, 3000)
    otr = np.concatenate([o.run_sets(2, 1, 2500, 0, True), o.run_sets(3, 1, 500, 0, True)])
    d0 = np.nonzero(np.concatenate([tr[0], tr2[0]]) != otr)[0]
    assert len(d0) == 0, f"first divergence at step {d0[0]}"
    ox, oy, oz = o.get_positions()
    assert np.array_equal(gx, ox) and np.array_equal(gy, oy) and np.array_equal(gz, oz)
    assert st2[0].energy == pytest.approx(st2[0].energy_check, rel=RTOL)
    if what != "box":
        assert st[0].energy == pytest.approx(st[0].energy_check, rel=RTOL)
        assert st2[0].energy == pytest.approx(o.state()["exact_energy"], rel=RTOL)
    assert st2[0].energy_check == pytest.approx(o.box_energy()[0][3], rel=RTOL)
    if what == "mu":
        assert before.pair_evals + 10 * len(px) ** 2 > st[0].pair_evals       # no N^2 rebuild was needed


def test_moves_only_run_skips_widom(mcpm, golden):
    """no_sampling: an NVT run (n_swap = 0) in a production set runs ONLY MoveParticle, as MC::MinimizeEnergy does — same chain
    as with sampling (Widom uses its own stream), no Widom sums, fewer executed pair evaluations, and the hot kernel."""
    arrays, meta = golden
    s = system_from_meta(meta["bulk"]["system"])
    s.n_equil_sets = 0
    x, y, z = arrays["bulk_x"], arrays["bulk_y"], arrays["bulk_z"]
    out = {}
    for mode in (0, 1):
        with mcpm.Engine(1, 1024) as e:
            e.set_system(s)
            e.set_positions(0, x, y, z)
            st, tr = e.run(1, 1, 3000, seed=8, trace=True, threads=32, check_energy=1, no_sampling=mode)
            out[mode] = (tr[0].copy(), e.get_positions(0), st[0].widom_insertions + st[0].widom_deletions, st[0].pair_evals_algorithmic)
            assert st[0].energy == pytest.approx(st[0].energy_check, rel=RTOL)
    assert np.array_equal(out[0][0], out[1][0])
    assert all(np.array_equal(a, b) for a, b in zip(out[0][1], out[1][1]))
    assert out[0][2] == 3000 + 2 and out[1][2] == 2                          # counts start at one (ResetStats)
    assert out[1][3] < out[0][3]


def test_bulk_transfers_of_ragged_chains(mcpm):
    """mcpm_set_positions_all / mcpm_get_positions_all copy runs of neighbouring chains with similar loadings, each run as wide as its largest
    chain: every live particle (and the free slot) of every chain must arrive and come back whatever the mix, and the byte counters must
    show the saving over one copy as wide as the largest chain."""
    import numpy as np
    from conftest import O
    s = O.argon_slit(n_equil_sets=0)
    rng = np.random.default_rng(3)
    cap = 512
    loadings = np.concatenate([np.full(40, 30), np.full(40, 450), [0, 511, 1, 300], rng.integers(0, 500, size=36)]).astype(np.int32)
    nch = len(loadings)
    with mcpm.Engine(nch, cap) as e:
        e.set_system(s)
        X = np.zeros((nch, e.capacity)); Y = np.zeros_like(X); Z = np.zeros_like(X)
        for k, n in enumerate(loadings):
            X[k, :n + 1] = rng.random(n + 1) * 40.0; Y[k, :n + 1] = rng.random(n + 1) * 40.0; Z[k, :n + 1] = 3.0 + rng.random(n + 1) * 14.0
        b0 = e.transfer_bytes()
        e.set_positions_all(loadings, X, Y, Z)
        b1 = e.transfer_bytes()
        for k in (0, 39, 40, 79, 80, 81, 82, 83, 100, nch - 1):
            x, y, z = e.get_positions(k)
            n = int(loadings[k])
            assert np.array_equal(x, X[k, :n]) and np.array_equal(y, Y[k, :n]) and np.array_equal(z, Z[k, :n]), k
        n2, X2, Y2, Z2 = e.get_positions_all()
        assert np.array_equal(n2, loadings)
        for k, n in enumerate(loadings):
            assert np.array_equal(X2[k, :n + 1], X[k, :n + 1]) and np.array_equal(Z2[k, :n + 1], Z[k, :n + 1]), k   # incl. the free slot
        one_wide_copy = 3 * nch * 512 * 8
        assert 0 < b1[0] - b0[0] < 0.75 * one_wide_copy and e.transfer_bytes()[1] - b1[1] == b1[0] - b0[0]
        box = e.box_energy_all()
        o = O.COracle(s, cap + 8)
        for k in (0, 40, 83, nch - 1):
            n = int(loadings[k])
            o.set_positions(X[k, :n], Y[k, :n], Z[k, :n])
            assert np.allclose(np.asarray(box[k]), o.box_energy()[0], rtol=1e-10, atol=1e-9)
