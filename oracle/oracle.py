"""TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.

ctypes bindings for the two CPU checkers:

* ``COracle``   -> oracle/liboracle.so, the plain-C restatement (oracle/mcpm_oracle.c);
* ``RefOracle`` -> oracle/_ref/libmcref.so, the UNMODIFIED reference compiled from /root/reference
  plus the subclass harness oracle/ref_harness.cpp.  Each instance touches ~5.6 GB (reference
  constructor, src/MC.cpp:41-55) and takes ~6 s to construct: create ONE per process.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import
this module.  The product (mc-porousmaterials_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBORACLE = os.path.join(HERE, "liboracle.so")
LIBMCREF = os.path.join(HERE, "_ref", "libmcref.so")
LIBMCREF_NDEBUG = os.path.join(HERE, "_ref", "libmcref_ndebug.so")   # same sources, -DNDEBUG (MC::ChangeVolume aborts on an Eigen assertion otherwise)
LIBMCREF_BOUND = os.path.join(HERE, "_ref", "libmcref_bound.so")   # the reference with its energy seam bound to libmcpm.so
SIMULATE = os.path.join(HERE, "_ref", "simulate")
REFERENCE_ROOT = "/root/reference"

_dp = C.POINTER(C.c_double)


class System(C.Structure):
    """Mirror of orc_system / mcref_system (single box, single species)."""

    _fields_ = [
        ("geometry", C.c_int), ("wall_kind", C.c_int), ("n_layers", C.c_int),
        ("n_disp", C.c_int), ("n_vol", C.c_int), ("n_swap", C.c_int),
        ("n_equil_sets", C.c_int), ("pair_kind", C.c_int),
        ("width", C.c_double * 3), ("temperature", C.c_double),
        ("eps_ff", C.c_double), ("sig_ff", C.c_double), ("rcut", C.c_double),
        ("molar_mass", C.c_double), ("mu", C.c_double),
        ("eps_sf", C.c_double), ("sig_sf", C.c_double), ("rho_s", C.c_double), ("delta", C.c_double),
        ("dr", C.c_double), ("pressure", C.c_double), ("dv", C.c_double),
    ]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if name == "width" else v
        return d


def make_system(**kw) -> System:
    s = System()
    for k, v in kw.items():
        if k == "width":
            s.width = (C.c_double * 3)(*v)
        else:
            setattr(s, k, v)
    return s


def build(ref: bool = True, quiet: bool = True) -> None:
    """Compile the checkers (building the checker is not using it)."""
    out = subprocess.DEVNULL if quiet else None
    subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=out)
    if ref and os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        subprocess.check_call(["make", "-C", HERE, "-j8", "ref"], stdout=out)


def _arr(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def decode_state(d, i):
    return {
        "pair": d[0], "many": d[1], "wall": d[2], "energy": d[3], "dr": d[4], "volume": d[5],
        "old_energy": d[6], "temperature": d[7],
        "n": int(i[0]), "acceptance": int(i[1]), "rejection": int(i[2]), "n_displacements": int(i[3]),
        "accept_swap": int(i[4]), "reject_swap": int(i[5]), "n_swaps": int(i[6]),
    }


class COracle:
    """The C restatement.  One instance = one chain."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(LIBORACLE):
                build(ref=False)
            L = C.CDLL(LIBORACLE)
            L.orc_create.restype = C.c_void_p
            L.orc_create.argtypes = [C.POINTER(System), C.c_int]
            L.orc_destroy.argtypes = [C.c_void_p]
            L.orc_set_positions.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
            L.orc_get_positions.argtypes = [C.c_void_p, _dp, _dp, _dp]
            L.orc_get_positions.restype = C.c_int
            L.orc_particle_energy.argtypes = [C.c_void_p, C.c_int, _dp]
            L.orc_trial_energy.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, _dp]
            L.orc_moved_energy.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, _dp]
            L.orc_box_energy.argtypes = [C.c_void_p, _dp, _dp, _dp]
            L.orc_slit_wall.argtypes = [C.POINTER(System), C.c_double]
            L.orc_slit_wall.restype = C.c_double
            L.orc_volume.argtypes = [C.POINTER(System)]
            L.orc_volume.restype = C.c_double
            L.orc_swap_zz.argtypes = [C.POINTER(System)]
            L.orc_swap_zz.restype = C.c_double
            L.orc_rng_mt.argtypes = [C.c_void_p, C.c_uint64]
            L.orc_rng_replay.argtypes = [C.c_void_p, _dp, C.c_size_t]
            L.orc_rng_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64]
            L.orc_random.argtypes = [C.c_void_p]
            L.orc_random.restype = C.c_double
            L.orc_draws.argtypes = [C.c_void_p]
            L.orc_draws.restype = C.c_size_t
            L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
            L.orc_philox_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_int]
            L.orc_philox_uniform.restype = C.c_double
            L.orc_mt_fill.argtypes = [C.c_uint64, _dp, C.c_size_t]
            L.orc_initial_config.argtypes = [C.c_void_p, C.c_int]
            L.orc_step.argtypes = [C.c_void_p, C.c_int]
            L.orc_step.restype = C.c_int
            L.orc_run.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]
            L.orc_run.restype = C.c_longlong
            L.orc_reset_stats.argtypes = [C.c_void_p]
            L.orc_set_dr.argtypes = [C.c_void_p, C.c_double]
            L.orc_get_volume_state.argtypes = [C.c_void_p, _dp, C.POINTER(C.c_longlong)]
            L.orc_wall_energy.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
            L.orc_wall_energy.restype = C.c_double
            L.orc_hyp2f1.argtypes = [C.c_int, C.c_double]
            L.orc_hyp2f1.restype = C.c_double
            L.orc_adjust.argtypes = [C.c_void_p, C.c_int]
            L.orc_run_sets.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_void_p]
            L.orc_get_state.argtypes = [C.c_void_p, _dp, C.POINTER(C.c_longlong)]
            L.orc_widom.argtypes = [C.c_void_p, _dp, C.POINTER(C.c_longlong)]
            L.orc_chem_pot.argtypes = [C.c_void_p, _dp]
            L.orc_rdf.argtypes = [C.c_void_p, _dp]
            L.orc_rdf_reset.argtypes = [C.c_void_p]
            L.orc_set_sample_rdf.argtypes = [C.c_void_p, C.c_int]
            L.orc_get_rdf.argtypes = [C.c_void_p, _dp]
            L.orc_get_widom.argtypes = [C.c_void_p, _dp, C.POINTER(C.c_longlong)]
            L.orc_link_boxes.argtypes = [C.c_void_p, C.c_void_p]
            L.orc_gibbs_step.argtypes = [C.c_void_p, C.c_int]
            L.orc_gibbs_step.restype = C.c_int
            L.orc_gibbs_run_sets.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_void_p]
            cls._lib = L
        return cls._lib

    def __init__(self, system: System, capacity: int = 4096):
        self.L = self.lib()
        self.system = system
        self.capacity = capacity
        self.h = self.L.orc_create(C.byref(system), capacity)
        self._replay = None

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state
    def set_positions(self, x, y, z):
        x, px = _arr(x); y, py = _arr(y); z, pz = _arr(z)
        self.L.orc_set_positions(self.h, len(x), px, py, pz)

    def get_positions(self):
        x = np.zeros(self.capacity); y = np.zeros(self.capacity); z = np.zeros(self.capacity)
        n = self.L.orc_get_positions(self.h, x.ctypes.data_as(_dp), y.ctypes.data_as(_dp), z.ctypes.data_as(_dp))
        return x[:n].copy(), y[:n].copy(), z[:n].copy()

    # -- energies: {pair, many, wall, total}
    def particle_energy(self, index0):
        out = np.zeros(4); self.L.orc_particle_energy(self.h, index0, out.ctypes.data_as(_dp)); return out

    def trial_energy(self, x, y, z):
        out = np.zeros(4); self.L.orc_trial_energy(self.h, x, y, z, out.ctypes.data_as(_dp)); return out

    def moved_energy(self, index0, x, y, z):
        out = np.zeros(4); self.L.orc_moved_energy(self.h, index0, x, y, z, out.ctypes.data_as(_dp)); return out

    def box_energy(self):
        n = self.state()["n"]
        out = np.zeros(4); pp = np.zeros(max(n, 1)); pw = np.zeros(max(n, 1))
        self.L.orc_box_energy(self.h, out.ctypes.data_as(_dp), pp.ctypes.data_as(_dp), pw.ctypes.data_as(_dp))
        return out, pp[:n], pw[:n]

    def slit_wall(self, z):
        return self.L.orc_slit_wall(C.byref(self.system), z)

    def volume(self):
        return self.L.orc_volume(C.byref(self.system))

    def swap_zz(self):
        return self.L.orc_swap_zz(C.byref(self.system))

    # -- rng
    def seed(self, s):
        self.L.orc_rng_mt(self.h, s)

    def replay(self, u):
        self._replay = np.ascontiguousarray(u, dtype=np.float64)
        self.L.orc_rng_replay(self.h, self._replay.ctypes.data_as(_dp), len(self._replay))

    def philox(self, seed, chain_id, first_step=0):
        self.L.orc_rng_philox(self.h, seed, chain_id, first_step)

    def random(self):
        return self.L.orc_random(self.h)

    def draws(self):
        return self.L.orc_draws(self.h)

    # -- moves
    def initial_config(self, n):
        self.L.orc_initial_config(self.h, n)

    def step(self, correct_energy=0):
        return self.L.orc_step(self.h, correct_energy)

    def run(self, nsteps, correct_energy=0, trace=False):
        buf = np.zeros(nsteps if trace else 1, dtype=np.uint8)
        acc = self.L.orc_run(self.h, nsteps, correct_energy, buf.ctypes.data if trace else None)
        return (acc, buf) if trace else acc

    def reset_stats(self):
        self.L.orc_reset_stats(self.h)

    def set_dr(self, dr):
        self.L.orc_set_dr(self.h, dr)

    def volume_state(self):
        d = np.zeros(6); i = np.zeros(3, dtype=np.int64)
        self.L.orc_get_volume_state(self.h, d.ctypes.data_as(_dp), i.ctypes.data_as(C.POINTER(C.c_longlong)))
        return {"volume": d[0], "width": [d[1], d[2], d[3]], "dv": d[4], "as_shipped_energy": d[5],
                "acceptance_vol": int(i[0]), "rejection_vol": int(i[1]), "n_vol_changes": int(i[2])}

    def wall_energy(self, x, y, z):
        return self.L.orc_wall_energy(self.h, x, y, z)

    def adjust(self, set_index):
        self.L.orc_adjust(self.h, set_index)

    def run_sets(self, first_set, n_sets, steps_per_set, correct_energy=0, trace=False):
        buf = np.zeros(n_sets * steps_per_set if trace else 1, dtype=np.uint8)
        self.L.orc_run_sets(self.h, first_set, n_sets, steps_per_set, correct_energy, buf.ctypes.data if trace else None)
        return buf if trace else None

    # -- two boxes (Gibbs ensemble): `other` becomes box 1 of this oracle's system and draws from this oracle's uniform source
    def link(self, other):
        self.peer = other
        self.L.orc_link_boxes(self.h, other.h)

    def gibbs_step(self, correct_energy=0):
        return self.L.orc_gibbs_step(self.h, correct_energy)

    def gibbs_run(self, nsteps, correct_energy=0):
        return np.array([self.L.orc_gibbs_step(self.h, correct_energy) for _ in range(nsteps)], dtype=np.uint8)

    def gibbs_run_sets(self, first_set, n_sets, steps_per_set, correct_energy=0, trace=False):
        buf = np.zeros(n_sets * steps_per_set if trace else 1, dtype=np.uint8)
        self.L.orc_gibbs_run_sets(self.h, first_set, n_sets, steps_per_set, correct_energy, buf.ctypes.data if trace else None)
        return buf if trace else None

    def state(self):
        d = np.zeros(16); i = np.zeros(16, dtype=np.int64)
        self.L.orc_get_state(self.h, d.ctypes.data_as(_dp), i.ctypes.data_as(C.POINTER(C.c_longlong)))
        s = decode_state(d, i)
        s.update({
            "exact_pair": d[8], "exact_wall": d[9], "exact_energy": d[10],
            "sum_n": d[11], "sum_n2": d[12], "sum_e": d[13], "sum_e2": d[14], "samples": int(d[15]),
            "tot_disp": int(i[8]), "tot_disp_acc": int(i[9]), "tot_swap": int(i[10]), "tot_swap_acc": int(i[11]),
            "recomputes": int(i[12]), "pair_evals": int(i[13]),
        })
        return s

    # -- sampling
    def widom(self):
        out = np.zeros(2); i = np.zeros(2, dtype=np.int64)
        self.L.orc_widom(self.h, out.ctypes.data_as(_dp), i.ctypes.data_as(C.POINTER(C.c_longlong)))
        return out, i

    def chem_pot(self):
        out = np.zeros(2); self.L.orc_chem_pot(self.h, out.ctypes.data_as(_dp)); return out

    def rdf(self):
        h = np.zeros(100); self.L.orc_rdf(self.h, h.ctypes.data_as(_dp)); return h

    def rdf_reset(self):
        self.L.orc_rdf_reset(self.h)

    def set_sample_rdf(self, on=1):
        self.L.orc_set_sample_rdf(self.h, on)

    def get_rdf(self):
        h = np.zeros(100); self.L.orc_get_rdf(self.h, h.ctypes.data_as(_dp)); return h

    def get_widom(self):
        out = np.zeros(2); i = np.zeros(2, dtype=np.int64)
        self.L.orc_get_widom(self.h, out.ctypes.data_as(_dp), i.ctypes.data_as(C.POINTER(C.c_longlong)))
        return out, i


def philox4x32_10(ctr, key):
    L = COracle.lib()
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    L.orc_philox4x32_10(c, k, o)
    return [int(v) for v in o]


def philox_uniform(seed, chain_id, step, slot):
    return COracle.lib().orc_philox_uniform(seed, chain_id, step, slot)


def mt_uniforms(seed, n):
    """n values of the reference's Random() for a given mt19937_64 seed (quirk Q1 mapping)."""
    u = np.zeros(n)
    COracle.lib().orc_mt_fill(seed, u.ctypes.data_as(_dp), n)
    return u


def have_reference_lib() -> bool:
    return os.path.exists(LIBMCREF)


def have_bound_reference_lib() -> bool:
    return os.path.exists(LIBMCREF_BOUND)


class RefOracle:
    """The compiled, unmodified reference behind the subclass harness.  ~5.6 GB per instance.
    bound=True loads oracle/_ref/libmcref_bound.so instead: the same reference objects with src/energy.cpp replaced by
    oracle/energy_bound.cpp, i.e. every energy comes from libmcpm.so (needs a GPU)."""

    _libs = {}

    @classmethod
    def lib(cls, bound=False, ndebug=False):
        path = LIBMCREF_BOUND if bound else (LIBMCREF_NDEBUG if ndebug else LIBMCREF)
        if path not in cls._libs:
            L = C.CDLL(path)
            L.mcref_create.restype = C.c_void_p
            L.mcref_destroy.argtypes = [C.c_void_p]
            L.mcref_setup.argtypes = [C.c_void_p, C.POINTER(System)]
            L.mcref_set_positions.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
            L.mcref_get_positions.argtypes = [C.c_void_p, _dp, _dp, _dp]
            L.mcref_get_positions.restype = C.c_int
            L.mcref_seed.argtypes = [C.c_void_p, C.c_uint64]
            L.mcref_random.argtypes = [C.c_void_p]
            L.mcref_random.restype = C.c_double
            L.mcref_particle_energy.argtypes = [C.c_void_p, C.c_int, _dp]
            L.mcref_trial_energy.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, _dp]
            L.mcref_moved_energy.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, _dp]
            L.mcref_box_energy.argtypes = [C.c_void_p, _dp, _dp, _dp]
            L.mcref_get_state.argtypes = [C.c_void_p, _dp, C.POINTER(C.c_int)]
            L.mcref_reset_stats.argtypes = [C.c_void_p]
            L.mcref_volume_state.argtypes = [C.c_void_p, _dp, C.POINTER(C.c_int)]
            L.mcref_adjust.argtypes = [C.c_void_p, C.c_int]
            L.mcref_initial_config.argtypes = [C.c_void_p, C.c_int]
            L.mcref_step.argtypes = [C.c_void_p, C.c_int]
            L.mcref_step.restype = C.c_int
            L.mcref_run.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]
            L.mcref_run.restype = C.c_longlong
            L.mcref_run_timed.argtypes = [C.c_void_p, C.c_longlong, C.c_int]
            L.mcref_run_timed.restype = C.c_double
            L.mcref_box_energy_timed.argtypes = [C.c_void_p, C.c_int]
            L.mcref_box_energy_timed.restype = C.c_double
            L.mcref_widom.argtypes = [C.c_void_p, _dp, C.POINTER(C.c_int)]
            L.mcref_chem_pot.argtypes = [C.c_void_p, _dp]
            L.mcref_rdf.argtypes = [C.c_void_p, C.c_int, _dp]
            L.mcref_rdf_reset.argtypes = [C.c_void_p]
            L.mcref_read_input.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(System), C.POINTER(C.c_int),
                                           C.c_char_p, C.c_char_p, C.c_char_p, _dp]
            L.mcref_read_input.restype = C.c_int
            L.mcref_volume.argtypes = [C.c_void_p]
            L.mcref_volume.restype = C.c_double
            L.mcref_swap_zz.argtypes = [C.c_void_p]
            L.mcref_swap_zz.restype = C.c_double
            L.mcref_add_box.argtypes = [C.c_void_p, C.POINTER(System)]
            L.mcref_set_positions_box.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp]
            L.mcref_get_positions_box.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
            L.mcref_get_positions_box.restype = C.c_int
            L.mcref_box_energy_box.argtypes = [C.c_void_p, C.c_int, _dp]
            L.mcref_get_state_box.argtypes = [C.c_void_p, C.c_int, _dp, C.POINTER(C.c_int)]
            L.mcref_widom_box.argtypes = [C.c_void_p, C.c_int, _dp, C.POINTER(C.c_int)]
            L.mcref_gibbs_step.argtypes = [C.c_void_p, C.c_int]
            L.mcref_gibbs_step.restype = C.c_int
            L.mcref_gibbs_run.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]
            L.mcref_gibbs_run.restype = C.c_longlong
            L.mcref_is_bound.argtypes = [_dp]
            L.mcref_is_bound.restype = C.c_int
            cls._libs[path] = L
        return cls._libs[path]

    def __init__(self, system: System | None = None, bound: bool = False, ndebug: bool = False):
        self.L = self.lib(bound, ndebug)
        self.h = self.L.mcref_create()
        self.capacity = 100000
        if system is not None:
            self.setup(system)

    def close(self):
        if self.h:
            self.L.mcref_destroy(self.h)
            self.h = None

    def setup(self, system: System):
        self.system = system
        self.L.mcref_setup(self.h, C.byref(system))

    def set_positions(self, x, y, z):
        x, px = _arr(x); y, py = _arr(y); z, pz = _arr(z)
        self.L.mcref_set_positions(self.h, len(x), px, py, pz)

    def get_positions(self):
        n = self.state()["n"]
        x = np.zeros(n + 1); y = np.zeros(n + 1); z = np.zeros(n + 1)
        self.L.mcref_get_positions(self.h, x.ctypes.data_as(_dp), y.ctypes.data_as(_dp), z.ctypes.data_as(_dp))
        return x[:n].copy(), y[:n].copy(), z[:n].copy()

    def particle_energy(self, index0):
        out = np.zeros(4); self.L.mcref_particle_energy(self.h, index0, out.ctypes.data_as(_dp)); return out

    def trial_energy(self, x, y, z):
        out = np.zeros(4); self.L.mcref_trial_energy(self.h, x, y, z, out.ctypes.data_as(_dp)); return out

    def moved_energy(self, index0, x, y, z):
        out = np.zeros(4); self.L.mcref_moved_energy(self.h, index0, x, y, z, out.ctypes.data_as(_dp)); return out

    def box_energy(self):
        n = self.state()["n"]
        out = np.zeros(4); pp = np.zeros(max(n, 1)); pw = np.zeros(max(n, 1))
        self.L.mcref_box_energy(self.h, out.ctypes.data_as(_dp), pp.ctypes.data_as(_dp), pw.ctypes.data_as(_dp))
        return out, pp[:n], pw[:n]

    def volume(self):
        return self.L.mcref_volume(self.h)

    def swap_zz(self):
        return self.L.mcref_swap_zz(self.h)

    def seed(self, s):
        self.L.mcref_seed(self.h, s)

    def random(self):
        return self.L.mcref_random(self.h)

    def initial_config(self, n):
        self.L.mcref_initial_config(self.h, n)

    def step(self, correct_energy=0):
        return self.L.mcref_step(self.h, correct_energy)

    def run(self, nsteps, correct_energy=0, trace=False):
        buf = np.zeros(nsteps if trace else 1, dtype=np.uint8)
        acc = self.L.mcref_run(self.h, nsteps, correct_energy, buf.ctypes.data if trace else None)
        return (acc, buf) if trace else acc

    def run_timed(self, nsteps, correct_energy=1):
        return self.L.mcref_run_timed(self.h, nsteps, correct_energy)

    def box_energy_timed(self, reps):
        return self.L.mcref_box_energy_timed(self.h, reps)

    def reset_stats(self):
        self.L.mcref_reset_stats(self.h)

    def adjust(self, set_index):
        self.L.mcref_adjust(self.h, set_index)

    def state(self):
        d = np.zeros(8); i = np.zeros(8, dtype=np.int32)
        self.L.mcref_get_state(self.h, d.ctypes.data_as(_dp), i.ctypes.data_as(C.POINTER(C.c_int)))
        return decode_state(d, i)

    # -- two boxes (Gibbs ensemble)
    def add_box(self, system: System):
        self.system1 = system
        self.L.mcref_add_box(self.h, C.byref(system))

    def set_positions_box(self, ib, x, y, z):
        x, px = _arr(x); y, py = _arr(y); z, pz = _arr(z)
        self.L.mcref_set_positions_box(self.h, ib, len(x), px, py, pz)

    def get_positions_box(self, ib):
        n = self.state_box(ib)["n"]
        x = np.zeros(n + 1); y = np.zeros(n + 1); z = np.zeros(n + 1)
        self.L.mcref_get_positions_box(self.h, ib, x.ctypes.data_as(_dp), y.ctypes.data_as(_dp), z.ctypes.data_as(_dp))
        return x[:n].copy(), y[:n].copy(), z[:n].copy()

    def box_energy_box(self, ib):
        out = np.zeros(4); self.L.mcref_box_energy_box(self.h, ib, out.ctypes.data_as(_dp)); return out

    def state_box(self, ib):
        d = np.zeros(8); i = np.zeros(8, dtype=np.int32)
        self.L.mcref_get_state_box(self.h, ib, d.ctypes.data_as(_dp), i.ctypes.data_as(C.POINTER(C.c_int)))
        return decode_state(d, i)

    def widom_box(self, ib):
        out = np.zeros(2); i = np.zeros(2, dtype=np.int32)
        self.L.mcref_widom_box(self.h, ib, out.ctypes.data_as(_dp), i.ctypes.data_as(C.POINTER(C.c_int)))
        return out, i

    def gibbs_run(self, nsteps, correct_energy=0):
        buf = np.zeros(nsteps, dtype=np.uint8)
        self.L.mcref_gibbs_run(self.h, nsteps, correct_energy, buf.ctypes.data)
        return buf

    def widom(self):
        out = np.zeros(2); i = np.zeros(2, dtype=np.int32)
        self.L.mcref_widom(self.h, out.ctypes.data_as(_dp), i.ctypes.data_as(C.POINTER(C.c_int)))
        return out, i

    def chem_pot(self):
        out = np.zeros(2); self.L.mcref_chem_pot(self.h, out.ctypes.data_as(_dp)); return out

    def rdf(self, enable=1):
        h = np.zeros(100); self.L.mcref_rdf(self.h, enable, h.ctypes.data_as(_dp)); return h

    def rdf_reset(self):
        self.L.mcref_rdf_reset(self.h)

    def volume_state(self):
        d = np.zeros(6); i = np.zeros(3, dtype=np.int32)
        self.L.mcref_volume_state(self.h, d.ctypes.data_as(_dp), i.ctypes.data_as(C.POINTER(C.c_int)))
        return {"volume": d[0], "width": [d[1], d[2], d[3]], "dv": d[4], "as_shipped_energy": d[5],
                "acceptance_vol": int(i[0]), "rejection_vol": int(i[1]), "n_vol_changes": int(i[2])}

    def bound_stats(self):
        """None for the stock build; for the bound build the binding's call counters and seconds."""
        out = np.zeros(6)
        if not self.L.mcref_is_bound(out.ctypes.data_as(_dp)):
            return None
        return {"particle_calls": int(out[0]), "particle_seconds": out[1], "box_calls": int(out[2]), "box_seconds": out[3],
                "mirrored": int(out[4]), "uploads": int(out[5])}

    def read_input(self, path):
        s = System(); n0 = C.c_int(0)
        proj = C.create_string_buffer(64); fname = C.create_string_buffer(64); bname = C.create_string_buffer(64)
        sets = np.zeros(4)
        rc = self.L.mcref_read_input(self.h, path.encode(), C.byref(s), C.byref(n0), proj, fname, bname,
                                     sets.ctypes.data_as(_dp))
        return rc, s, n0.value, proj.value.decode(), fname.value.decode(), bname.value.decode(), sets


# ----------------------------------------------------------------------------- canned systems

def argon_slit(mu=-1134.0, H=20.0, rcut=20.0, T=87.0, n_disp=1, n_swap=1, n_equil_sets=10, dr=None) -> System:
    """BASELINE config C1/C2: Ar in a graphite slit, 3 layers per wall."""
    L = 2.0 * rcut  # read.cpp:134 — lateral size is 2*max(rcut)
    return make_system(geometry=1, wall_kind=1, n_layers=3, n_disp=n_disp, n_vol=0, n_swap=n_swap,
                       n_equil_sets=n_equil_sets, width=(L, L, H), temperature=T,
                       eps_ff=119.8, sig_ff=3.405, rcut=rcut, molar_mass=39.948, mu=mu,
                       eps_sf=57.92, sig_sf=3.4025, rho_s=0.382, delta=3.35,
                       dr=(0.1 * H if dr is None else dr))


def methane_slit(mu=-3050.0, H=20.0, rcut=20.0, T=300.0, n_disp=1, n_swap=3, n_equil_sets=10) -> System:
    """BASELINE config C3: TraPPE-UA CH4 in a graphite slit."""
    L = 2.0 * rcut
    return make_system(geometry=1, wall_kind=1, n_layers=3, n_disp=n_disp, n_vol=0, n_swap=n_swap,
                       n_equil_sets=n_equil_sets, width=(L, L, H), temperature=T,
                       eps_ff=148.0, sig_ff=3.73, rcut=rcut, molar_mass=16.043, mu=mu,
                       eps_sf=64.37, sig_sf=3.565, rho_s=0.382, delta=3.35, dr=0.1 * H)


def argon_bulk(size=100.0, rcut=12.0, T=120.0, n_equil_sets=10) -> System:
    """BASELINE config C4: dense NVT LJ argon in a cubic periodic box."""
    return make_system(geometry=0, wall_kind=0, n_layers=0, n_disp=1, n_vol=0, n_swap=0,
                       n_equil_sets=n_equil_sets, width=(size, size, size), temperature=T,
                       eps_ff=119.8, sig_ff=3.405, rcut=rcut, molar_mass=39.948, mu=0.0,
                       eps_sf=0.0, sig_sf=0.0, rho_s=0.0, delta=0.0, dr=0.1 * size)
