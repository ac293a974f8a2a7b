"""The BASELINE.json workloads as system descriptions and synthetic start configurations.  No numerics here."""
from __future__ import annotations

import math

import numpy as np

from .abi import GEOM_BULK, GEOM_SLIT, WALL_NONE, WALL_SLIT_LJ, SystemDesc

KB, NA, PLANCK = 1.380649e-23, 6.02214076e23, 6.62607015e-34  # reference src/MC.h:16-18


def _desc(**kw) -> SystemDesc:
    d = SystemDesc()
    for k, v in kw.items():
        if k == "width":
            d.width[0], d.width[1], d.width[2] = v
        else:
            setattr(d, k, v)
    return d


def thermal_wavelength(molar_mass: float, T: float) -> float:
    """Lambda in angstrom, as the reference computes it (src/moves.cpp:296-297)."""
    m = molar_mass / NA * 1e-3
    return PLANCK / math.sqrt(2.0 * math.pi * m * KB * T) * 1e10


def argon_slit(mu=-1134.0, H=20.0, rcut=20.0, T=87.0, n_disp=1, n_swap=1, n_equil_sets=10) -> SystemDesc:
    """C1 / one point of C2: Ar in a graphite slit (3 layers per wall); lateral size 2*rcut (src/read.cpp:134)."""
    L = 2.0 * rcut
    return _desc(geometry=GEOM_SLIT, wall_kind=WALL_SLIT_LJ, n_layers=3, n_disp=n_disp, n_vol=0, n_swap=n_swap,
                 n_equil_sets=n_equil_sets, width=(L, L, H), temperature=T, eps_ff=119.8, sig_ff=3.405, rcut=rcut,
                 molar_mass=39.948, mu=mu, eps_sf=57.92, sig_sf=3.4025, rho_s=0.382, delta=3.35, dr=0.1 * H)


def isotherm_mu_grid(n_points: int = 40):
    """C2: mu_k = -1800 + k*(666/39) K, k = 0..39."""
    return [-1800.0 + k * (666.0 / 39.0) for k in range(n_points)]


def methane_slit(mu=-3050.0, H=20.0, rcut=92.0, T=300.0, n_equil_sets=10) -> SystemDesc:
    """C3: TraPPE-UA CH4, 300 K, high loading; Rcut 92 => L = 184 A, N ~ 5k; disp:swap = 1:3."""
    L = 2.0 * rcut
    return _desc(geometry=GEOM_SLIT, wall_kind=WALL_SLIT_LJ, n_layers=3, n_disp=1, n_vol=0, n_swap=3,
                 n_equil_sets=n_equil_sets, width=(L, L, H), temperature=T, eps_ff=148.0, sig_ff=3.73, rcut=rcut,
                 molar_mass=16.043, mu=mu, eps_sf=64.37, sig_sf=3.565, rho_s=0.382, delta=3.35, dr=0.1 * H)


def argon_bulk(size=100.0, rcut=12.0, T=120.0, n_equil_sets=10) -> SystemDesc:
    """C4: dense NVT LJ argon, cubic periodic box, translations only."""
    return _desc(geometry=GEOM_BULK, wall_kind=WALL_NONE, n_layers=0, n_disp=1, n_vol=0, n_swap=0,
                 n_equil_sets=n_equil_sets, width=(size, size, size), temperature=T, eps_ff=119.8, sig_ff=3.405,
                 rcut=rcut, molar_mass=39.948, mu=0.0, eps_sf=0.0, sig_sf=0.0, rho_s=0.0, delta=0.0, dr=0.1 * size)


def replica_sweep(n_h: int = 32, n_t: int = 32):
    """C5: pore width x temperature sweep, H in 8..39 A, T in 77..139 K, Rcut 17 (L = 34 A),
    mu(T) = T*ln(8.44e-5 * Lambda(T)^3)."""
    out = []
    for ih in range(n_h):
        for it in range(n_t):
            H = 8.0 + ih; T = 77.0 + 2.0 * it
            mu = T * math.log(8.44e-5 * thermal_wavelength(39.948, T) ** 3)
            out.append(argon_slit(mu=mu, H=H, rcut=17.0, T=T))
    return out


def c4_lattice(n: int = 20000, size: float = 100.0, m: int = 28, seed: int = 2024):
    """C4 start: simple-cubic lattice of m^3 sites (spacing size/m), first n sites occupied, uniform
    jitter +-0.2 A per axis (numpy default_rng(seed)), shifted by 0.3 A so that every coordinate lies inside the box."""
    rng = np.random.default_rng(seed)
    site = np.arange(m) * (size / m)
    gx, gy, gz = np.meshgrid(site, site, site, indexing="ij")
    p = np.column_stack([gx.ravel(), gy.ravel(), gz.ravel()])[:n] + (rng.random((n, 3)) - 0.5) * 0.4 + 0.3
    return p[:, 0].copy(), p[:, 1].copy(), p[:, 2].copy()


def c5_start(desc: SystemDesc, rng, n0: int = 100):
    """C5 start: n0 random particles, at least 3 A away from either wall."""
    return rng.random(n0) * desc.width[0], rng.random(n0) * desc.width[1], 3.0 + rng.random(n0) * (desc.width[2] - 6.0)


def bench_capacity(nmax: int) -> int:
    """Particle slots per chain used by bench.py: 15 % + 40 headroom over the largest start loading, rounded up to 32."""
    return int(-(-int(nmax * 1.15 + 40) // 32) * 32)
