"""GPU: K2 (batched persistent chains) through the C ABI.

* replay mode: the reference's own mt19937_64 uniform stream -> identical accept/reject sequence,
  counters, dr adaptation and final configuration as the compiled reference (golden) and the C oracle;
* Philox mode: the device stream is bit-identical to its host restatement, and a Philox-driven chain
  follows the oracle fed with the same stream step for step;
* bookkeeping: running energies equal a full recompute to 1e-10.
"""
import numpy as np
import pytest

from conftest import O, system_from_meta

pytestmark = pytest.mark.gpu
RTOL = 1e-10

TRACE_CASES = {"c1t": ("c1_trace", "c1"), "lowt": ("low_trace", "c1"), "bulkt": ("bulk_trace", "bulk")}


def first_divergence(a, b):
    d = np.nonzero(a != b)[0]
    return int(d[0]) if len(d) else -1


@pytest.mark.parametrize("cache", [True, False, "subwarp", "spec", "dual", "pair", "classic"], ids=["cache", "nocache", "subwarp", "spec", "dual", "pair", "classic"])
@pytest.mark.parametrize("threads", [32, 64, 128, 256, 512])
@pytest.mark.parametrize("name", ["c1t", "lowt", "bulkt"])
def test_replay_trace_matches_reference_golden(mcpm, golden, name, threads, cache):
    """'cache' at 32 threads = the pooled persistent kernel k2_pool, 'classic' = one CTA per chain (k2_chains) for one-warp chains too;
    'dual' = the opt-in dual-candidate kernel (k2_dual, one warp per chain, two steps per round)."""
    if cache == "classic" and (threads != 32 or not mcpm.has_experiments()):
        pytest.skip("'classic' only differs from 'cache' in the experiments build, where the pooled kernel replaces the one-warp variant")
    if cache in ("subwarp", "dual") and not mcpm.has_experiments():
        pytest.skip("experimental kernels: build with `make EXPERIMENTS=1` and run with MCPM_EXPERIMENTS=1")
    if cache == "subwarp" and threads != 32:
        pytest.skip("the sub-warp kernel runs one warp per chain")
    if cache == "spec" and threads != 32:
        pytest.skip("the speculative kernel has one CTA shape")
    if cache in ("dual", "pair") and threads != 32:
        pytest.skip("the dual-candidate and two-warp kernels only replace the one-warp variant")
    arrays, meta = golden
    mkey, start = TRACE_CASES[name]
    m = meta[mkey]
    s = system_from_meta(m["system"])
    n0, sets, sps = m["n0"], m["sets"], m["steps_per_set"]
    x, y, z = arrays[start + "_x"][:n0], arrays[start + "_y"][:n0], arrays[start + "_z"][:n0]
    u = O.mt_uniforms(m["seed"], 8 * sets * sps)
    with mcpm.Engine(1, 1024) as e:
        e.use_energy_cache(bool(cache))
        e.use_subwarp(cache == "subwarp")
        e.use_speculation({"spec": 1, "pair": 2}.get(cache, 0))
        e.use_dual_candidates(cache == "dual")
        e.use_pool(16 if (cache is True and mcpm.has_experiments()) else 0)
        e.set_system(s)
        e.set_positions(0, x, y, z)
        per_set, traces = [], []
        pos = 0
        for k in range(1, sets + 1):       # one call per set so that per-set counters can be compared
            stats, tr = e.run(k, 1, sps, replay=u[pos:], trace=True, threads=threads, check_energy=1)
            st = stats[0]
            # how many uniforms did this set consume?  1 + 7/6/4/2 per step
            codes = tr[0]
            mv = codes >> 1
            acc = codes & 1
            pos += int(np.sum(np.where(mv == 0, 8, np.where(mv == 1, 7, np.where(mv == 2, 5, 3)))))
            traces.append(codes)
            per_set.append([st.n, st.acceptance, st.rejection, st.n_displacements, st.accept_swap, st.reject_swap,
                            st.n_swaps, st.dr])
            assert st.energy == pytest.approx(st.energy_check, rel=RTOL)
            assert st.pair == pytest.approx(st.pair_check, rel=RTOL, abs=1e-9)
        if cache in (True, "classic") and threads == 32:
            assert e.last_kernel_name() == ("k2_pool" if (cache is True and name != "bulkt" and mcpm.has_experiments()) else "k2_chains")   # (NVT samples Widom: never pooled)
        got = np.concatenate(traces)
        div = first_divergence(got, arrays[name + "_trace"])
        assert div == -1, f"first divergence at step {div}"
        assert np.array_equal(np.array(per_set, dtype=np.float64), arrays[name + "_per_set"])
        fx, fy, fz = e.get_positions(0)
        # positions are produced by the same fp64 operations in the same order: bit-identical
        assert np.array_equal(fx, arrays[name + "_fx"]) and np.array_equal(fy, arrays[name + "_fy"]) and np.array_equal(fz, arrays[name + "_fz"])
        box, _, _ = e.box_energy(0)
        assert np.allclose(box, arrays[name + "_fbox"], rtol=RTOL, atol=0)


def test_replay_multi_set_single_launch(mcpm, golden):
    """All sets in ONE launch (dr adaptation on the device) == the per-set golden sequence."""
    arrays, meta = golden
    m = meta["c1_trace"]
    s = system_from_meta(m["system"])
    sets, sps = m["sets"], m["steps_per_set"]
    u = O.mt_uniforms(m["seed"], 8 * sets * sps)
    with mcpm.Engine(1, 1024) as e:
        e.set_system(s)
        e.set_positions(0, arrays["c1_x"], arrays["c1_y"], arrays["c1_z"])
        stats, tr = e.run(1, sets, sps, replay=u, trace=True, threads=128)
        assert np.array_equal(tr[0], arrays["c1t_trace"])
        assert stats[0].dr == arrays["c1t_per_set"][-1][7]
        assert stats[0].n == int(arrays["c1t_per_set"][-1][0])


def test_device_philox_stream_is_the_documented_one(mcpm):
    with mcpm.Engine(1, 32) as e:
        e.set_system(O.argon_slit())
        for seed, chain, first in [(0, 0, 0), (12345, 7, 3), (2 ** 40 + 17, 1023, 2 ** 33 + 5)]:
            got = e.philox_uniforms(seed, chain, first, 40)
            want = np.array([[O.philox_uniform(seed, chain, first + s, k) for k in range(8)] for s in range(40)])
            assert np.array_equal(got, want)


@pytest.mark.parametrize("cache", [True, False, "spec", "dual", "pair"], ids=["cache", "nocache", "spec", "dual", "pair"])
@pytest.mark.parametrize("threads", [32, 128])
def test_philox_chain_follows_oracle(mcpm, golden, threads, cache):
    """Device Philox + Metropolis on the GPU == the C oracle fed with the same counter-based stream."""
    if cache == "dual" and not mcpm.has_experiments():
        pytest.skip("experimental kernel: build with `make EXPERIMENTS=1` and run with MCPM_EXPERIMENTS=1")
    if cache in ("spec", "dual", "pair") and threads != 32:
        pytest.skip("one CTA shape")
    arrays, meta = golden
    s = system_from_meta(meta["c1"]["system"])
    s.n_equil_sets = 1
    x, y, z = arrays["c1_x"], arrays["c1_y"], arrays["c1_z"]
    seed, nchains, sets, sps = 987654321, 3, 2, 1500
    with mcpm.Engine(nchains, 1024) as e:
        e.use_energy_cache(bool(cache))
        e.use_speculation({"spec": 1, "pair": 2}.get(cache, 0))
        e.use_dual_candidates(cache == "dual")
        pooled = cache is True and threads == 32 and mcpm.has_experiments()
        e.use_pool(16 if pooled else 0)                                   # the pooled one-warp kernel: experiments build only
        e.set_system(s)
        for c in range(nchains):
            e.set_positions(c, x, y, z)
        stats, tr = e.run(1, sets, sps, seed=seed, trace=True, threads=threads, check_energy=1)
        assert e.last_kernel_name() == {True: "k2_pool" if pooled else "k2_chains", False: "k2_chains", "spec": "k2_spec", "dual": "k2_dual", "pair": "k2_pair"}[cache]
        for c in range(nchains):
            o = O.COracle(s, 1024)
            o.set_positions(x, y, z)
            o.box_energy()
            evals0 = o.state()["pair_evals"]          # the initial BoxEnergy's n*n evaluations
            o.philox(seed, c, 0)
            otr = o.run_sets(1, sets, sps, 0, True)
            div = first_divergence(tr[c], otr)
            assert div == -1, f"chain {c}: first divergence at step {div}"
            st, ost = stats[c], o.state()
            assert st.n == ost["n"] and st.dr == ost["dr"]
            assert st.tot_disp == ost["tot_disp"] and st.tot_disp_acc == ost["tot_disp_acc"]
            assert st.tot_ins + st.tot_del == ost["tot_swap"]
            assert st.pair_evals_algorithmic == ost["pair_evals"] - evals0    # what the reference's moves evaluate
            if not cache:      # with the energy cache fewer partner scans are needed (and the rebuild adds n*n)
                assert st.pair_evals == ost["pair_evals"] - evals0
            elif cache in ("spec", "dual", "pair"):   # discarded candidates are executed evaluations too
                assert st.pair_evals > 0
            else:
                assert st.pair_evals < ost["pair_evals"] - evals0 + len(x) ** 2
            assert st.samples == ost["samples"] == sps
            assert st.sum_n == ost["sum_n"] and st.sum_n2 == ost["sum_n2"]
            assert st.sum_e == pytest.approx(ost["sum_e"], rel=RTOL)
            assert st.energy == pytest.approx(ost["exact_energy"], rel=RTOL)
            assert st.energy == pytest.approx(st.energy_check, rel=RTOL)
            gx, gy, gz = e.get_positions(c)
            ox, oy, oz = o.get_positions()
            assert np.array_equal(gx, ox) and np.array_equal(gy, oy) and np.array_equal(gz, oz)
        # chains with different ids decorrelate
        assert not np.array_equal(tr[0], tr[1])


def test_philox_stream_independent_of_launch_partitioning(mcpm, golden):
    arrays, meta = golden
    s = system_from_meta(meta["c1"]["system"])
    x, y, z = arrays["c1_x"], arrays["c1_y"], arrays["c1_z"]
    out = []
    for split in (1, 3):
        with mcpm.Engine(1, 1024) as e:
            e.set_system(s)
            e.set_positions(0, x, y, z)
            trs = []
            for k in range(split):
                _, tr = e.run(20, 1, 1203 // split, seed=5, trace=True, threads=64)
                trs.append(tr[0])
            out.append((np.concatenate(trs), e.get_positions(0)))
    assert np.array_equal(out[0][0], out[1][0])
    for a, b in zip(out[0][1], out[1][1]):
        assert np.array_equal(a, b)


def test_capacity_overflow_is_reported(mcpm):
    s = O.argon_slit(mu=-900.0)
    rng = np.random.default_rng(1)
    with mcpm.Engine(1, 32) as e:
        e.set_system(s)
        e.set_positions(0, rng.random(30) * 40, rng.random(30) * 40, 3.0 + rng.random(30) * 14)
        with pytest.raises(mcpm.McpmError) as ei:
            e.run(1, 1, 20000, seed=1)
        assert ei.value.code == 3
        assert e.stats()[0].overflow == 1


def _replay_vs_oracle(mcpm, system, x, y, z, capacity, threads, nsteps, seed, expect_kernel=None, draws_per_step=8):
    u = O.mt_uniforms(seed, draws_per_step * nsteps)
    o = O.COracle(system, capacity)
    o.set_positions(x, y, z)
    o.box_energy()
    o.replay(u)
    otr = o.run_sets(1, 1, nsteps, 0, True)
    with mcpm.Engine(1, capacity) as e:
        e.set_system(system)
        e.set_positions(0, x, y, z)
        stats, tr = e.run(1, 1, nsteps, replay=u, trace=True, threads=threads, check_energy=1)
        if expect_kernel:
            assert e.last_kernel_name() == expect_kernel
        gx, gy, gz = e.get_positions(0)
    div = first_divergence(tr[0], otr)
    assert div == -1, f"first divergence at step {div}"
    ox, oy, oz = o.get_positions()
    assert np.array_equal(gx, ox) and np.array_equal(gy, oy) and np.array_equal(gz, oz)
    st, ost = stats[0], o.state()
    assert st.n == ost["n"] and st.replay_consumed == o.draws()
    assert st.energy == pytest.approx(st.energy_check, rel=RTOL)
    assert st.energy == pytest.approx(ost["exact_energy"], rel=RTOL)


@pytest.mark.parametrize("threads", [32, 64, 256, 512])
def test_general_minimum_image_chain(mcpm, golden, threads):
    """Coordinates outside the box (a restart file may hold them): the chain must use the reference's general
    d -= round(d/L)*L image, and moved particles are wrapped exactly as MC::PBC does."""
    arrays, meta = golden
    s = system_from_meta(meta["c1"]["system"])
    x, y, z = arrays["c1_x"].copy(), arrays["c1_y"].copy(), arrays["c1_z"].copy()
    rng = np.random.default_rng(5)
    x += 40.0 * rng.integers(-2, 3, size=len(x)); y += 40.0 * rng.integers(-2, 3, size=len(y))
    _replay_vs_oracle(mcpm, s, x, y, z, 1024, threads, 3000, 77)


@pytest.mark.parametrize("capacity,threads", [(6016, 512), (6016, 1024), (12000, 512), (12000, 1024), (12000, 0)])
def test_large_chains_shared_and_global_memory_paths(mcpm, capacity, threads):
    """N = 3000 CH4-like slit: capacity 6016 is staged in shared memory (144 KB), capacity 12000 exceeds one CTA's
    shared memory and is advanced in place in global memory (the path config C4, N = 20 000, takes)."""
    s = O.methane_slit(rcut=60.0, mu=-3050.0)
    rng = np.random.default_rng(11)
    n = 3000
    x, y, z = rng.random(n) * 120.0, rng.random(n) * 120.0, 3.5 + rng.random(n) * 13.0
    _replay_vs_oracle(mcpm, s, x, y, z, capacity, threads, 400, 123)


def test_bulk_global_path_translations_only(mcpm):
    s = O.argon_bulk(size=70.0, rcut=12.0)
    rng = np.random.default_rng(12)
    n = 5000
    x, y, z = rng.random(n) * 70.0, rng.random(n) * 70.0, rng.random(n) * 70.0
    _replay_vs_oracle(mcpm, s, x, y, z, 10240, 1024, 300, 9)


def _jittered_lattice(m, size, seed):
    rng = np.random.default_rng(seed)
    site = np.arange(m) * (size / m)
    gx, gy, gz = np.meshgrid(site, site, site, indexing="ij")
    p = np.column_stack([gx.ravel(), gy.ravel(), gz.ravel()]) + (rng.random((m ** 3, 3)) - 0.5) * 0.5 + 0.4
    return p[:, 0].copy(), p[:, 1].copy(), p[:, 2].copy()


@pytest.mark.parametrize("ensemble", ["nvt", "nvt_production", "gcmc"])
def test_cell_list_chain_follows_oracle(mcpm, ensemble):
    """K2 through a cell list (k2_cells: bulk box with 5 cut-off lengths per axis, N = 5000, positions in global memory) against
    the all-pairs C oracle under the reference's mt19937_64 stream: same trace, bit-identical positions, energies to 1e-10.
    'nvt_production' also runs ComputeWidom every step (13 draws per step), 'gcmc' exercises the list updates of
    insertions and deletions (swap-with-last renames an entry)."""
    s = O.argon_bulk(size=70.0, rcut=12.0)
    x, y, z = _jittered_lattice(17, 70.0, 12)       # 4913 particles, no overlaps
    if ensemble == "gcmc":
        s.n_swap, s.mu = 2, -900.0
    if ensemble == "nvt_production":
        s.n_equil_sets = 0
    _replay_vs_oracle(mcpm, s, x, y, z, 10240, 0, 400, 9, expect_kernel="k2_cells", draws_per_step=13)


def test_cell_list_widom_and_accumulators(mcpm):
    """Philox mode through the cell list: Widom/BAR sums, production accumulators and counters equal the oracle's."""
    s = O.argon_bulk(size=70.0, rcut=12.0)
    s.n_equil_sets = 1
    x, y, z = _jittered_lattice(17, 70.0, 13)
    seed, sets, sps = 77, 2, 300
    o = O.COracle(s, 8192)
    o.set_positions(x, y, z); o.box_energy(); o.philox(seed, 0, 0)
    otr = o.run_sets(1, sets, sps, 0, True)
    with mcpm.Engine(1, 10240) as e:
        e.set_system(s)
        e.set_positions(0, x, y, z)
        stats, tr = e.run(1, sets, sps, seed=seed, trace=True, check_energy=1)
        assert e.last_kernel_name() == "k2_cells"
        gx_, gy_, gz_ = e.get_positions(0)
    assert first_divergence(tr[0], otr) == -1
    ox, oy, oz = o.get_positions()
    assert np.array_equal(gx_, ox) and np.array_equal(gy_, oy) and np.array_equal(gz_, oz)
    st, ost = stats[0], o.state()
    (wi, wd), (ni, nd) = o.get_widom()
    assert (st.widom_insertions, st.widom_deletions) == (ni, nd)
    assert st.widom_insert == pytest.approx(wi, rel=1e-9) and st.widom_delete == pytest.approx(wd, rel=1e-9)
    assert st.samples == ost["samples"] and st.sum_n == ost["sum_n"]
    assert st.sum_e == pytest.approx(ost["sum_e"], rel=RTOL)
    assert st.tot_disp_acc == ost["tot_disp_acc"] and st.dr == ost["dr"]
    assert st.energy == pytest.approx(st.energy_check, rel=RTOL)
    assert st.pair_evals < st.pair_evals_algorithmic / 3          # ~27 cells of 125 instead of all N partners


@pytest.mark.parametrize("threads", [32, 128])
def test_nvt_widom_and_rdf_follow_oracle(mcpm, golden, threads):
    """NVT (n_swap = 0): production steps run ComputeWidom and ComputeRDF after the move, as src/main.cpp:72-75 does.
    Device Philox (move stream + sampling stream) vs the C oracle fed with the same streams."""
    arrays, meta = golden
    s = system_from_meta(meta["bulk"]["system"])
    s.n_equil_sets = 1
    x, y, z = arrays["bulk_x"], arrays["bulk_y"], arrays["bulk_z"]
    seed, sets, sps = 4321, 3, 400
    o = O.COracle(s, 1024)
    o.set_positions(x, y, z)
    o.box_energy()
    o.philox(seed, 0, 0)
    o.set_sample_rdf(1)
    otr = o.run_sets(1, sets, sps, 0, True)
    with mcpm.Engine(1, 1024) as e:
        e.set_system(s)
        e.set_positions(0, x, y, z)
        stats, tr = e.run(1, sets, sps, seed=seed, trace=True, threads=threads, sample_rdf=1, check_energy=1)
        hist = e.get_rdf(0)
        mu_ex, mu = e.chemical_potential(0)
        gx, gy, gz = e.get_positions(0)
    assert first_divergence(tr[0], otr) == -1
    ox, oy, oz = o.get_positions()
    assert np.array_equal(gx, ox) and np.array_equal(gy, oy) and np.array_equal(gz, oz)
    st = stats[0]
    (wi, wd), (ni, nd) = o.get_widom()
    assert (st.widom_insertions, st.widom_deletions) == (ni, nd) and ni + nd == sps + 2      # last set only; counts start at 1
    assert st.widom_insert == pytest.approx(wi, rel=1e-10) and st.widom_delete == pytest.approx(wd, rel=1e-10)
    omu = o.chem_pot()
    assert mu_ex == pytest.approx(omu[0], rel=1e-9) and mu == pytest.approx(omu[1], rel=1e-9)
    ohist = o.get_rdf()
    assert hist.sum() == ohist.sum() > 0
    assert np.array_equal(hist, ohist)                    # same binning as NeighDistance/deltaR, every production step
    assert st.energy == pytest.approx(st.energy_check, rel=RTOL)


def test_nvt_widom_replay_matches_reference_stream(mcpm, golden):
    """Replay mode: the Widom draws continue the reference's sequential mt19937_64 stream (8 + 5 or 3 per step)."""
    arrays, meta = golden
    s = system_from_meta(meta["bulk"]["system"])
    s.n_equil_sets = 0
    x, y, z = arrays["bulk_x"], arrays["bulk_y"], arrays["bulk_z"]
    u = O.mt_uniforms(99, 13 * 500)
    o = O.COracle(s, 1024)
    o.set_positions(x, y, z); o.box_energy(); o.replay(u)
    otr = o.run_sets(1, 1, 500, 0, True)
    with mcpm.Engine(1, 1024) as e:
        e.set_system(s)
        e.set_positions(0, x, y, z)
        stats, tr = e.run(1, 1, 500, replay=u, trace=True, threads=64)
    assert first_divergence(tr[0], otr) == -1
    (wi, wd), (ni, nd) = o.get_widom()
    st = stats[0]
    assert st.replay_consumed == o.draws()
    assert (st.widom_insertions, st.widom_deletions) == (ni, nd)
    assert st.widom_insert == pytest.approx(wi, rel=1e-10) and st.widom_delete == pytest.approx(wd, rel=1e-10)


def test_results_do_not_depend_on_threads_per_chain(mcpm, golden):
    """Race detector by construction: 60 000 steps of 4 chains at 32/64/256/512 threads per chain must give bit-identical
    configurations, counters and accumulators (every thread decides redundantly from the same partial sums, and
    the owner-writes / register-forwarding scheme has one barrier per step; compute-sanitizer is closed on this pool)."""
    arrays, meta = golden
    s = system_from_meta(meta["c1"]["system"])
    s.n_equil_sets = 2
    x, y, z = arrays["c1_x"], arrays["c1_y"], arrays["c1_z"]
    ref = None
    for threads in (32, 64, 256, 512):
        with mcpm.Engine(4, 1024) as e:
            e.set_system(s)
            for c in range(4):
                e.set_positions(c, x, y, z)
            stats = e.run(1, 6, 10000, seed=2025, threads=threads)
            pos = [e.get_positions(c) for c in range(4)]
        sig = [(st.n, st.dr, st.tot_disp_acc, st.tot_ins_acc, st.tot_del_acc, st.sum_n, st.pair_evals_algorithmic) for st in stats]
        if ref is None:
            ref = (sig, pos)
            continue
        assert sig == ref[0], f"threads={threads}"
        for a, b in zip(pos, ref[1]):
            for u, v in zip(a, b):
                assert np.array_equal(u, v), f"threads={threads}"


def test_energy_cache_survives_launches_and_uploads(mcpm, golden):
    """The per-particle energy cache is kept across launches (checksum of the configuration), is rebuilt after the host
    changes a configuration, and never changes the chain: split and unsplit runs, with and without re-uploading the
    positions in between, give bit-identical results."""
    arrays, meta = golden
    s = system_from_meta(meta["c1"]["system"])
    x, y, z = arrays["c1_x"], arrays["c1_y"], arrays["c1_z"]
    results = []
    for mode in ("one", "split", "reupload", "nocache"):
        with mcpm.Engine(2, 1024) as e:
            e.use_energy_cache(mode != "nocache")
            e.set_system(s)
            for c in range(2):
                e.set_positions(c, x, y, z)
            if mode in ("one", "nocache"):
                st = e.run(50, 1, 6000, seed=77, threads=32, check_energy=1)
            else:
                e.run(50, 1, 2500, seed=77, threads=32)
                if mode == "reupload":
                    n, X, Y, Z = e.get_positions_all()
                    st0 = e.stats()
                    e.set_positions_all(n, X, Y, Z)                    # unchanged configuration: the cache is reused
                    e.set_running_energy([q.pair for q in st0], [q.wall for q in st0])
                e.run(50, 1, 3500, seed=77, threads=32)
                st = e.run(50, 1, 0, seed=77, threads=32, check_energy=1)
            results.append(([(q.n, q.tot_disp_acc, q.tot_ins_acc, q.tot_del_acc) for q in st], [e.get_positions(c) for c in range(2)],
                            [(q.energy, q.energy_check) for q in st]))
    for sig, pos, en in results[1:]:
        assert sig == results[0][0]
        for a, b in zip(pos, results[0][1]):
            for u, v in zip(a, b):
                assert np.array_equal(u, v)
    for sig, pos, en in results:
        for running, check in en:
            assert running == pytest.approx(check, rel=RTOL)


def test_energy_cache_with_hard_overlaps(mcpm):
    """Random start (energies ~1e20 K): accepted moves push huge terms through the cache, which must be rebuilt on the
    spot; the chain stays identical to the cache-free one and the running energy stays exact after relaxation."""
    s = O.argon_slit(mu=-1300.0, rcut=12.0, n_equil_sets=0)
    rng = np.random.default_rng(8)
    n0 = 150
    x, y, z = rng.random(n0) * 24, rng.random(n0) * 24, rng.random(n0) * 20
    out = []
    for cache in (True, False, "spec"):
        with mcpm.Engine(1, 512) as e:
            e.use_energy_cache(bool(cache))
            e.use_speculation(1 if cache == "spec" else 0)
            e.set_system(s)
            e.set_positions(0, x, y, z)
            st, tr = e.run(1, 1, 20000, seed=5, threads=64, trace=True, check_energy=2)
            st2 = e.run(2, 1, 5000, seed=5, threads=64, check_energy=1)
            out.append((tr[0].copy(), e.get_positions(0), st2[0].energy, st2[0].energy_check))
    for other in out[1:]:
        assert first_divergence(out[0][0], other[0]) == -1
        for a, b in zip(out[0][1], other[1]):
            assert np.array_equal(a, b)
    for _, _, running, check in out:
        assert running == pytest.approx(check, rel=1e-9)


@pytest.mark.parametrize("start", ["dense", "empty", "bulk"])
def test_speculative_moves_are_the_sequential_chain(mcpm, golden, start):
    """The speculative kernel (8 candidate steps per round, first accepted one committed, later ones re-evaluated) and the
    dual-candidate kernel (2 per round inside one warp) must
    reproduce the one-step-at-a-time kernel bit for bit: trace, configuration, counters, dr adaptation, accumulators.
    'empty' starts from an empty pore (stale-slot moves, quirk Q12, force serial rounds) and fills it."""
    arrays, meta = golden
    s = system_from_meta(meta["c1"]["system"])
    s.n_equil_sets = 2
    if start == "bulk":          # periodic in z, GCMC
        s = system_from_meta(meta["bulk_trace"]["system"])
        s.n_equil_sets = 2
        s.n_swap, s.mu = 1, -1100.0
        x, y, z = arrays["bulk_x"], arrays["bulk_y"], arrays["bulk_z"]
    elif start == "dense":
        x, y, z = arrays["c1_x"], arrays["c1_y"], arrays["c1_z"]
    else:
        x = y = z = np.zeros(0)
    nch, sets, sps = 3, 4, 5003
    out = []
    kernels = [("k2_chains", 0, 64), ("k2_spec", 1, 0), ("k2_pair", 2, 32), ("k2_quad", 4, 0)]
    if mcpm.has_experiments():
        kernels.append(("k2_dual", 0, 32))
    for kernel, spec, threads in kernels:
        with mcpm.Engine(nch, 1024) as e:
            e.use_speculation(spec)
            e.use_dual_candidates(kernel == "k2_dual")
            e.set_system(s)
            for c in range(nch):
                e.set_positions(c, x, y, z)
            stats, tr = e.run(1, sets, sps, seed=424242, trace=True, threads=threads, check_energy=1)
            assert e.last_kernel_name() == kernel
            out.append((tr.copy(), [e.get_positions(c) for c in range(nch)], stats))
    for c, other in [(c, other) for c in range(nch) for other in range(1, len(out))]:
        assert first_divergence(out[0][0][c], out[other][0][c]) == -1
        for a, b in zip(out[0][1][c], out[other][1][c]):
            assert np.array_equal(a, b)
        p, q = out[0][2][c], out[other][2][c]
        for f in ("n", "dr", "acceptance", "rejection", "n_displacements", "accept_swap", "reject_swap", "n_swaps", "tot_disp", "tot_disp_acc",
                  "tot_ins", "tot_ins_acc", "tot_del", "tot_del_acc", "tot_empty", "steps_done", "samples", "sum_n", "sum_n2", "pair_evals_algorithmic"):
            assert getattr(p, f) == getattr(q, f), f
        assert q.sum_e == pytest.approx(p.sum_e, rel=RTOL) and q.sum_e2 == pytest.approx(p.sum_e2, rel=RTOL)
        assert q.energy == pytest.approx(q.energy_check, rel=RTOL, abs=1e-9)
        assert q.energy == pytest.approx(p.energy, rel=RTOL, abs=1e-9)
    if start == "empty":
        assert out[1][2][0].tot_empty > 0 and out[1][2][0].n > 0


def test_speculation_is_chosen_for_few_chains_only(mcpm, golden):
    """Automatic mode: a context with no more chains than SMs runs the speculative kernel (its executed evaluation count
    includes discarded candidates, so it differs from the standard kernel's), a caller-fixed thread count does not."""
    arrays, meta = golden
    s = system_from_meta(meta["c1"]["system"])
    x, y, z = arrays["c1_x"], arrays["c1_y"], arrays["c1_z"]
    evals = {}
    for mode, threads in (("auto", 0), ("off", 0), ("fixed", 128)):
        with mcpm.Engine(2, 1024) as e:
            if mode == "off":
                e.use_speculation(0)
            e.set_system(s)
            for c in range(2):
                e.set_positions(c, x, y, z)
            st = e.run(50, 1, 4000, seed=3, threads=threads)
            evals[mode] = (st[0].pair_evals, st[0].pair_evals_algorithmic, st[0].n)
    assert evals["auto"][1:] == evals["off"][1:] == evals["fixed"][1:]
    assert evals["off"][0] == evals["fixed"][0]
    assert evals["auto"][0] > evals["off"][0]


@pytest.mark.parametrize("packed", [True, False], ids=["subwarp", "onewarp"])
def test_subwarp_packing_mixed_loadings(mcpm, golden, packed):
    if packed and not mcpm.has_experiments():
        pytest.skip("experimental kernel: build with `make EXPERIMENTS=1` and run with MCPM_EXPERIMENTS=1")
    """The throughput kernel packs 4 / 2 / 1 chains per warp by loading.  12 chains of very different size (3 ... 504
    particles, different mu) against the C oracle fed with the same Philox streams: identical traces and configurations."""
    arrays, meta = golden
    base = meta["c1"]["system"]
    x, y, z = arrays["c1_x"], arrays["c1_y"], arrays["c1_z"]
    sizes = [504, 3, 40, 150, 10, 260, 90, 504, 33, 200, 5, 120]
    mus = [-1134.0, -1900.0, -1700.0, -1500.0, -1800.0, -1350.0, -1600.0, -1134.0, -1750.0, -1400.0, -1850.0, -1550.0]
    seed, sets, sps = 31337, 2, 1500
    with mcpm.Engine(len(sizes), 704) as e:
        e.use_subwarp(packed)
        systems = []
        for c, (n0, mu) in enumerate(zip(sizes, mus)):
            d = dict(base); d["mu"] = mu; d["n_equil_sets"] = 1
            s = system_from_meta(d)
            systems.append(s)
            e.set_system(s, c)
            e.set_positions(c, x[:n0], y[:n0], z[:n0])
        stats, tr = e.run(1, sets, sps, seed=seed, trace=True, threads=32, check_energy=1)
        for c, (n0, s) in enumerate(zip(sizes, systems)):
            o = O.COracle(s, 1024)
            o.set_positions(x[:n0], y[:n0], z[:n0])
            o.box_energy()
            o.philox(seed, c, 0)
            otr = o.run_sets(1, sets, sps, 0, True)
            div = first_divergence(tr[c], otr)
            assert div == -1, f"chain {c} (n0={n0}): first divergence at step {div}"
            st, ost = stats[c], o.state()
            assert st.n == ost["n"] and st.dr == ost["dr"]
            assert (st.acceptance, st.rejection, st.n_displacements, st.accept_swap, st.reject_swap, st.n_swaps) == \
                (ost["acceptance"], ost["rejection"], ost["n_displacements"], ost["accept_swap"], ost["reject_swap"], ost["n_swaps"])
            assert st.samples == ost["samples"] and st.sum_n == ost["sum_n"]
            assert st.sum_e == pytest.approx(ost["sum_e"], rel=RTOL)
            assert st.energy == pytest.approx(st.energy_check, rel=RTOL, abs=1e-9)
            gx, gy, gz = e.get_positions(c)
            ox, oy, oz = o.get_positions()
            assert np.array_equal(gx, ox) and np.array_equal(gy, oy) and np.array_equal(gz, oz)


def test_subwarp_chain_outgrows_its_slot(mcpm, golden):
    """A chain that starts small (packed 4 per warp) and fills up during the launch is paused when its slot is full and
    finishes the call alone in a warp; the result is the chain the oracle produces."""
    if not mcpm.has_experiments():
        pytest.skip("experimental kernel: build with `make EXPERIMENTS=1` and run with MCPM_EXPERIMENTS=1")
    arrays, meta = golden
    d = dict(meta["c1"]["system"]); d["n_equil_sets"] = 0          # mu = -1134 K: N grows towards ~550
    s = system_from_meta(d)
    x, y, z = arrays["c1_x"][:20], arrays["c1_y"][:20], arrays["c1_z"][:20]
    seed, sps = 99, 60000
    with mcpm.Engine(8, 704) as e:
        e.use_subwarp(True)
        e.set_system(s)
        for c in range(8):
            e.set_positions(c, x, y, z)
        stats, tr = e.run(1, 1, sps, seed=seed, trace=True, threads=32, check_energy=1)
        launches = e.launch_count()
        o = O.COracle(s, 1024)
        o.set_positions(x, y, z); o.box_energy(); o.philox(seed, 5, 0)
        otr = o.run_sets(1, 1, sps, 0, True)
        assert first_divergence(tr[5], otr) == -1
        assert stats[5].n == o.state()["n"] > 176                  # it did outgrow the 4-per-warp slot (704/4 = 176)
        assert stats[5].samples == sps and stats[5].sum_n == o.state()["sum_n"]
        assert stats[5].steps_done == sps
        gx, gy, gz = e.get_positions(5)
        ox, oy, oz = o.get_positions()
        assert np.array_equal(gx, ox) and np.array_equal(gy, oy) and np.array_equal(gz, oz)
        assert launches >= 3                                        # K3 initialisation is 2 launches, then >= 2 K2 launches


@pytest.mark.parametrize("kind", [0, 1, 2], ids=["translation", "insertion", "deletion"])
def test_metropolis_screen_never_changes_a_decision(mcpm, kind):
    """The device decides u < f with bounds and a single-precision screen before it evaluates the reference's fp64
    expression (src/moves.cpp:170,306,326).  Stress the screen: uniforms placed at f (1 + d) for |d| from 1e-12 to 1e-2 on
    both sides, over the whole range of arguments, plus random pairs; every decision must equal the fp64 one (the two
    libm's exp may differ in the last ulp, hence |d| >= 1e-12)."""
    rng = np.random.default_rng(100 + kind)
    m = 20000
    zzV = 37.5
    n = rng.integers(1, 700, size=m).astype(np.int32)
    if kind == 0:
        x = np.concatenate([rng.random(m) * 40.0, -rng.random(100)])           # dE/T, some downhill
        nn = np.concatenate([n, n[:100]])
        f = np.exp(-x)
    elif kind == 1:
        x = rng.random(m) * 70.0 - 45.0                                        # -E/T
        nn = n
        f = zzV * np.exp(x) / (nn + 1.0)
    else:
        x = rng.random(m) * 70.0 - 60.0                                        # E/T
        nn = n
        f = nn * np.exp(x) / zzV
    d = rng.choice([1e-12, 1e-9, 1e-7, 3e-6, 1e-5, 5e-5, 9e-5, 1.1e-4, 3e-4, 1e-3, 1e-2], size=len(x)) * rng.choice([-1.0, 1.0], size=len(x))
    u_min = 0.9999 * 2.0 ** -53        # the smallest non-zero uniform of the 53-bit generators (quirk Q1 scaling included)
    u_edge = np.clip(f * (1.0 + d), u_min, 0.9999)
    u_rand = np.maximum(rng.random(len(x)) * 0.9999, u_min)
    with mcpm.Engine(1, 32) as e:
        for u in (u_edge, u_rand):
            got = e.metropolis_probe(kind, u, x, nn, zzV)
            want = u < f
            sure = np.abs(u - f) > 1e-13 * f                                   # away from the last-ulp ambiguity of exp itself
            bad = np.nonzero(sure & (got != want))[0]
            assert len(bad) == 0, f"{len(bad)} wrong decisions, first: u={u[bad[0]]!r} x={x[bad[0]]!r} n={nn[bad[0]]} f={f[bad[0]]!r}"
            assert sure.mean() > 0.5


@pytest.mark.parametrize("nch,kernel", [(300, "k2_quad"), (1000, "k2_pair"), (1500, "k2_chains")])
def test_lean_speculation_is_chosen_by_chains_per_sm(mcpm, golden, nch, kernel):
    """Automatic mode between 'fewer chains than SMs' (k2_spec) and 'more than 8 chains per SM' (one warp per chain): four or two
    candidate warps per chain.  Same chains as the one-step kernel."""
    arrays, meta = golden
    s = system_from_meta(meta["c1"]["system"])
    x, y, z = arrays["c1_x"][:200], arrays["c1_y"][:200], arrays["c1_z"][:200]
    out = []
    for mode in (-1, 0):
        with mcpm.Engine(nch, 256) as e:
            e.use_speculation(mode)
            e.set_system(s)
            for c in range(nch):
                e.set_positions(c, x, y, z)
            st = e.run(50, 1, 700, seed=11, check_energy=1)
            if mode == -1:
                assert e.last_kernel_name() == kernel          # 148 SMs: 3, 7 and 11 chains per SM
            out.append([(q.n, q.tot_disp_acc, q.tot_ins_acc, q.tot_del_acc, q.sum_n) for q in st[:5]] + [e.get_positions(3)])
            for q in st[:5]:
                assert q.energy == pytest.approx(q.energy_check, rel=RTOL)
    assert out[0][:5] == out[1][:5]
    for a, b in zip(out[0][5], out[1][5]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("warps", [16, 20])
def test_pooled_kernel_mixed_loadings_and_slot_classes(mcpm, golden, warps):
    if not mcpm.has_experiments():
        pytest.skip("experimental kernel: build with `make EXPERIMENTS=1` and run with MCPM_EXPERIMENTS=1")
    """k2_pool: 40 chains of very different size (3 ... 504 particles: three slot classes in one CTA mix) against the C oracle on
    the same Philox streams — traces, configurations, counters, accumulators — and bit-identical to the one-CTA-per-chain kernel."""
    arrays, meta = golden
    base = meta["c1"]["system"]
    x, y, z = arrays["c1_x"], arrays["c1_y"], arrays["c1_z"]
    sizes = [504, 3, 40, 150, 10, 260, 90, 504, 33, 200, 5, 120] * 3 + [300, 20, 450, 64]
    mus = [-1134.0, -1900.0, -1700.0, -1500.0, -1800.0, -1350.0, -1600.0, -1134.0, -1750.0, -1400.0, -1850.0, -1550.0] * 3 + [-1300.0, -1800.0, -1200.0, -1650.0]
    seed, sets, sps = 31337, 2, 1500
    results = {}
    for pool in (warps, 0):
        with mcpm.Engine(len(sizes), 704) as e:
            e.use_pool(pool)
            systems = []
            for c, (n0, mu) in enumerate(zip(sizes, mus)):
                d = dict(base); d["mu"] = mu; d["n_equil_sets"] = 1
                s = system_from_meta(d)
                systems.append(s)
                e.set_system(s, c)
                e.set_positions(c, x[:n0], y[:n0], z[:n0])
            stats, tr = e.run(1, sets, sps, seed=seed, trace=True, threads=32, check_energy=1)
            assert e.last_kernel_name() == ("k2_pool" if pool else "k2_chains")
            results[pool] = (tr.copy(), [e.get_positions(c) for c in range(len(sizes))], [(q.n, q.dr, q.sum_n, q.acceptance, q.n_swaps, q.steps_done, q.pair_evals) for q in stats],
                             np.array([(q.sum_e, q.energy) for q in stats]))
            if pool:
                for c in (0, 1, 5, 8, 36, 38):
                    n0, s = sizes[c], systems[c]
                    o = O.COracle(s, 1024)
                    o.set_positions(x[:n0], y[:n0], z[:n0]); o.box_energy(); o.philox(seed, c, 0)
                    otr = o.run_sets(1, sets, sps, 0, True)
                    assert first_divergence(tr[c], otr) == -1, c
                    st, ost = stats[c], o.state()
                    assert (st.n, st.dr, st.samples, st.sum_n) == (ost["n"], ost["dr"], ost["samples"], ost["sum_n"])
                    assert (st.acceptance, st.rejection, st.n_displacements, st.accept_swap, st.reject_swap, st.n_swaps) == \
                        (ost["acceptance"], ost["rejection"], ost["n_displacements"], ost["accept_swap"], ost["reject_swap"], ost["n_swaps"])
                    assert st.energy == pytest.approx(st.energy_check, rel=RTOL, abs=1e-9)
                    for a, b in zip(e.get_positions(c), o.get_positions()):
                        assert np.array_equal(a, b)
    assert np.array_equal(results[warps][0], results[0][0])
    assert results[warps][2] == results[0][2]
    assert np.allclose(results[warps][3], results[0][3], rtol=1e-12, atol=0)      # running sums: same terms, last-bit differences only
    for a, b in zip(results[warps][1], results[0][1]):
        assert all(np.array_equal(u, v) for u, v in zip(a, b))


@pytest.mark.parametrize("mode", ["philox", "replay"])
def test_pooled_chain_outgrows_its_slot(mcpm, golden, mode):
    if not mcpm.has_experiments():
        pytest.skip("experimental kernel: build with `make EXPERIMENTS=1` and run with MCPM_EXPERIMENTS=1")
    """Chains that start with 20 particles get a small slot; at mu = -1134 K they fill up to ~550 during the launch, pause when the slot
    is full and are finished by a follow-up launch in full-size slots.  The result is the chain the oracle produces (same trace, counters
    of the interrupted set, accumulators, dr adaptation, replay cursor)."""
    arrays, meta = golden
    d = dict(meta["c1"]["system"]); d["n_equil_sets"] = 2
    s = system_from_meta(d)
    x, y, z = arrays["c1_x"][:20], arrays["c1_y"][:20], arrays["c1_z"][:20]
    seed, sets, sps = 99, 4, 15000
    u = O.mt_uniforms(4242, 8 * sets * sps) if mode == "replay" else None
    nch = 1 if mode == "replay" else 6
    with mcpm.Engine(nch, 704) as e:
        e.use_pool(16)
        e.set_system(s)
        for c in range(nch):
            e.set_positions(c, x, y, z)
        launches0 = e.launch_count()
        stats, tr = e.run(1, sets, sps, seed=seed, replay=u, trace=True, threads=32, check_energy=1)
        assert e.last_kernel_name() == "k2_pool"
        assert e.launch_count() - launches0 >= 3 + 1                  # K3 initialisation (2 launches) + at least two K2 launches
        c = nch - 1
        o = O.COracle(s, 1024)
        o.set_positions(x, y, z); o.box_energy()
        if mode == "replay":
            o.replay(u)
        else:
            o.philox(seed, c, 0)
        otr = o.run_sets(1, sets, sps, 0, True)
        assert first_divergence(tr[c], otr) == -1
        st, ost = stats[c], o.state()
        assert st.n == ost["n"] > 200                                 # it did outgrow the slot of a 20-particle chain
        assert (st.dr, st.samples, st.sum_n, st.steps_done) == (ost["dr"], ost["samples"], ost["sum_n"], sets * sps)
        assert (st.acceptance, st.rejection, st.n_displacements, st.accept_swap, st.reject_swap, st.n_swaps) == \
            (ost["acceptance"], ost["rejection"], ost["n_displacements"], ost["accept_swap"], ost["reject_swap"], ost["n_swaps"])
        assert st.tot_disp == ost["tot_disp"] and st.tot_ins + st.tot_del == ost["tot_swap"]
        if mode == "replay":
            assert st.replay_consumed == o.draws()
        assert st.energy == pytest.approx(st.energy_check, rel=RTOL)
        for a, b in zip(e.get_positions(c), o.get_positions()):
            assert np.array_equal(a, b)
