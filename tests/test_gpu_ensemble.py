"""GPU: ensemble parity — loadings and energies from the device-Philox engine agree with the CPU oracle
(trace-identical to the compiled reference) within statistical error (north_star: within 2 sigma; the test
uses 3 sigma of the combined standard error to keep the false-failure rate of the suite low)."""
import numpy as np
import pytest

from conftest import O

pytestmark = pytest.mark.gpu


EQUIL, PROD = 200_000, 200_000


@pytest.mark.parametrize("mu,H", [(-1400.0, 20.0), (-1250.0, 12.0)])
def test_loading_matches_oracle_within_error(mcpm, mu, H):
    """Same protocol on both sides (random start, EQUIL steps, PROD sampled steps); the error bars come from
    INDEPENDENT chains on each side (12 CPU chains with different mt19937_64 seeds, 64 GPU replicas)."""
    s = O.argon_slit(mu=mu, H=H, rcut=12.0, n_equil_sets=1)
    n0 = 40
    n_cpu, e_cpu = [], []
    start = None
    for k in range(12):
        c = O.COracle(s, 2048)
        c.seed(1000 + k)
        c.initial_config(n0)
        if start is None:
            start = c.get_positions()
        c.box_energy()
        c.run_sets(1, 1, EQUIL, 0)                     # set 1 <= n_equil_sets: not sampled
        c.run_sets(2, 1, PROD, 0)                      # production: accumulators on
        st = c.state()
        assert st["samples"] == PROD
        n_cpu.append(st["sum_n"] / PROD); e_cpu.append(st["sum_e"] / PROD)
        c.close()
    R = 64
    with mcpm.Engine(R, 1024) as e:
        e.set_system(s)
        for r in range(R):
            e.set_positions(r, *start)
        e.run(1, 1, EQUIL, seed=7, check_energy=2)     # random start: energies ~1e20 K; re-anchor after equilibration
        stats = e.run(2, 1, PROD, seed=7, check_energy=1)
    for st in stats:
        assert st.samples == PROD
        assert abs(st.energy - st.energy_check) <= 1e-9 * abs(st.energy_check)
    n_gpu = np.array([st.sum_n / st.samples for st in stats]); e_gpu = np.array([st.sum_e / st.samples for st in stats])
    def mean_sem(v):
        v = np.asarray(v); return v.mean(), v.std(ddof=1) / np.sqrt(len(v))
    (ng, sg), (nc, scpu) = mean_sem(n_gpu), mean_sem(n_cpu)
    (eg, seg), (ec, sec) = mean_sem(e_gpu), mean_sem(e_cpu)
    assert abs(ng - nc) <= 3.0 * np.hypot(sg, scpu), ("<N>", ng, sg, nc, scpu)
    assert abs(eg - ec) <= 3.0 * np.hypot(seg, sec), ("<E>", eg, seg, ec, sec)
    # and the spread over chains is comparable (same ensemble, not just the same mean)
    assert 0.4 < n_gpu.std(ddof=1) / np.std(n_cpu, ddof=1) < 2.5


def test_replicas_are_statistically_independent(mcpm):
    s = O.argon_slit(mu=-1500.0, rcut=12.0, n_equil_sets=0)
    rng = np.random.default_rng(3)
    n0 = 30
    x, y, z = rng.random(n0) * 24, rng.random(n0) * 24, 3.0 + rng.random(n0) * 14
    with mcpm.Engine(32, 512) as e:
        e.set_system(s)
        for r in range(32):
            e.set_positions(r, x, y, z)
        stats, tr = e.run(1, 1, 4000, seed=11, trace=True)
    sig = [tuple(t[:200]) for t in tr]
    assert len(set(sig)) == 32                         # no two chains share an accept/reject sequence


def test_ideal_gas_known_answer(mcpm):
    """Closed-form check that needs no oracle: with eps = 0 and no wall, GCMC samples a Poisson distribution of N with
    mean zz*V (zz = exp(mu/T)/Lambda^3, src/moves.cpp:296-298), so <N> = Var(N) = zz*V.  (The reference's [0,0.9999)
    uniforms bias insertion/deletion acceptance identically, the ratio and hence the mean are unaffected to ~1e-4.)"""
    import math
    T, L = 100.0, 30.0
    lam = 6.62607015e-34 / math.sqrt(2.0 * math.pi * (39.948 / 6.02214076e23 * 1e-3) * 1.380649e-23 * T) * 1e10
    target = 60.0
    mu = T * math.log(target * lam ** 3 / L ** 3)
    s = O.make_system(geometry=0, wall_kind=0, n_layers=0, n_disp=1, n_vol=0, n_swap=1, n_equil_sets=0,
                      width=(L, L, L), temperature=T, eps_ff=0.0, sig_ff=3.4, rcut=10.0, molar_mass=39.948, mu=mu,
                      eps_sf=0.0, sig_sf=0.0, rho_s=0.0, delta=0.0, dr=3.0)
    rng = np.random.default_rng(0)
    R = 128
    with mcpm.Engine(R, 256) as e:
        e.set_system(s)
        for r in range(R):
            e.set_positions(r, rng.random(60) * L, rng.random(60) * L, rng.random(60) * L)
        e.run(1, 1, 20000, seed=3)
        e.reset_accumulators()
        stats = e.run(2, 1, 60000, seed=3)
    mean = np.array([st.sum_n / st.samples for st in stats])
    var = np.array([st.sum_n2 / st.samples - (st.sum_n / st.samples) ** 2 for st in stats])
    sem = mean.std(ddof=1) / np.sqrt(R)
    assert abs(mean.mean() - target) < 4.0 * sem + 0.01 * target * 1e-2, (mean.mean(), sem)
    assert var.mean() == pytest.approx(target, rel=0.03)
    assert all(st.energy == 0.0 for st in stats)
