"""ctypes declarations of include/mcpm.h (one to one)."""
from __future__ import annotations

import ctypes as C
import os

RNG_PHILOX, RNG_REPLAY = 0, 1
GEOM_BULK, GEOM_SLIT, GEOM_CYLINDER, GEOM_SPHERE = 0, 1, 2, 3
WALL_NONE, WALL_SLIT_LJ, WALL_SPHERE_LJ, WALL_CYL_LJ104, WALL_CYL_STEELE = 0, 1, 2, 3, 4
PAIR_LJ, PAIR_HS = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class McpmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mcpm error {code}: {msg}")
        self.code = code


class SystemDesc(C.Structure):
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

    @classmethod
    def from_any(cls, other) -> "SystemDesc":
        """Copy from any ctypes structure with the same field layout (e.g. the oracle's System)."""
        if isinstance(other, cls):
            return other
        assert C.sizeof(other) == C.sizeof(cls)
        out = cls()
        C.memmove(C.byref(out), C.byref(other), C.sizeof(cls))
        return out


class ParticleEnergy(C.Structure):
    _fields_ = [("pair", C.c_double), ("many_body", C.c_double), ("wall", C.c_double), ("energy", C.c_double)]


class BoxEnergy(C.Structure):
    _fields_ = [("pair", C.c_double), ("many_body", C.c_double), ("wall", C.c_double), ("energy", C.c_double)]


class RunParams(C.Structure):
    _fields_ = [
        ("first_set", C.c_int), ("n_sets", C.c_int), ("steps_per_set", C.c_longlong),
        ("rng_mode", C.c_int), ("seed", C.c_ulonglong),
        ("replay_u", _dp), ("replay_stride", C.c_size_t),
        ("trace", C.POINTER(C.c_ubyte)),
        ("threads_per_chain", C.c_int), ("sample_rdf", C.c_int), ("check_energy", C.c_int), ("no_sampling", C.c_int),
    ]


class ChainStats(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("overflow", C.c_int), ("dr", C.c_double),
        ("pair", C.c_double), ("wall", C.c_double), ("energy", C.c_double),
        ("pair_check", C.c_double), ("wall_check", C.c_double), ("energy_check", C.c_double),
        ("acceptance", C.c_longlong), ("rejection", C.c_longlong), ("n_displacements", C.c_longlong),
        ("accept_swap", C.c_longlong), ("reject_swap", C.c_longlong), ("n_swaps", C.c_longlong),
        ("tot_disp", C.c_longlong), ("tot_disp_acc", C.c_longlong), ("tot_ins", C.c_longlong),
        ("tot_ins_acc", C.c_longlong), ("tot_del", C.c_longlong), ("tot_del_acc", C.c_longlong),
        ("tot_empty", C.c_longlong),
        ("pair_evals", C.c_ulonglong), ("pair_evals_algorithmic", C.c_ulonglong), ("steps_done", C.c_ulonglong),
        ("sum_n", C.c_double), ("sum_n2", C.c_double), ("sum_e", C.c_double), ("sum_e2", C.c_double),
        ("samples", C.c_longlong), ("replay_consumed", C.c_ulonglong),
        ("widom_insert", C.c_double), ("widom_delete", C.c_double),
        ("widom_insertions", C.c_longlong), ("widom_deletions", C.c_longlong),
        ("accept_vol", C.c_longlong), ("reject_vol", C.c_longlong), ("n_vol_changes", C.c_longlong),
        ("volume", C.c_double), ("width", C.c_double), ("dv", C.c_double), ("energy_as_shipped", C.c_double),
        ("dr_used", C.c_double),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class SimConfig(C.Structure):
    _fields_ = [
        ("project", C.c_char * 64), ("fluid", C.c_char * 64), ("box", C.c_char * 64),
        ("n_sets", C.c_double), ("n_equil_sets", C.c_double), ("steps_per_set", C.c_double), ("print_every", C.c_double),
        ("pressure", C.c_double),
        ("n_particles", C.c_int), ("sample_rdf", C.c_int), ("print_trajectory", C.c_int), ("continue_after_crash", C.c_int),
        ("n_boxes", C.c_int), ("n_species", C.c_int),
        ("has_seed", C.c_int), ("seed", C.c_ulonglong),
        ("replicas", C.c_int), ("sweep_points", C.c_int), ("sweep_mu0", C.c_double), ("sweep_mu1", C.c_double),
        ("capacity", C.c_int), ("devices", C.c_int),
        ("system", SystemDesc),
        ("box2", C.c_char * 64), ("n_particles2", C.c_int), ("fix_volume", C.c_int * 2), ("system2", SystemDesc),
    ]


def lib_path() -> str:
    """libmcpm.so, or — only when MCPM_EXPERIMENTS=1 is set and it was built (make EXPERIMENTS=1) — libmcpm_exp.so,
    the same library plus the experimental kernels of csrc/experiments/."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if os.environ.get("MCPM_LIB"):              # kernel experiments: a variant build of the same library (scripts/variants.sh)
        return os.path.abspath(os.environ["MCPM_LIB"])
    exp = os.path.join(here, "libmcpm_exp.so")
    if os.environ.get("MCPM_EXPERIMENTS") == "1" and os.path.exists(exp):
        return exp
    return os.path.join(here, "libmcpm.so")


_LIB = None

# name -> (restype, argtypes); every symbol include/mcpm.h declares
PROTOTYPES = {
    "mcpm_version": (C.c_int, []),
    "mcpm_last_error": (C.c_char_p, []),
    "mcpm_device_count": (C.c_int, [_ip]),
    "mcpm_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "mcpm_destroy": (C.c_int, [C.c_void_p]),
    "mcpm_n_chains": (C.c_int, [C.c_void_p, _ip, _ip]),
    "mcpm_set_system": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(SystemDesc)]),
    "mcpm_get_system": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(SystemDesc)]),
    "mcpm_set_positions": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp]),
    "mcpm_get_positions": (C.c_int, [C.c_void_p, C.c_int, _ip, _dp, _dp, _dp]),
    "mcpm_set_positions_all": (C.c_int, [C.c_void_p, _ip, _dp, _dp, _dp]),
    "mcpm_get_positions_all": (C.c_int, [C.c_void_p, _ip, _dp, _dp, _dp]),
    "mcpm_particle_energy": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, C.POINTER(ParticleEnergy), C.POINTER(ParticleEnergy)]),
    "mcpm_particle_energy_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _ip, _dp, C.POINTER(ParticleEnergy), C.POINTER(ParticleEnergy)]),
    "mcpm_box_energy": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(BoxEnergy), _dp, _dp]),
    "mcpm_box_energy_all": (C.c_int, [C.c_void_p, C.POINTER(BoxEnergy)]),
    "mcpm_use_speculation": (C.c_int, [C.c_void_p, C.c_int]),
    "mcpm_use_energy_cache": (C.c_int, [C.c_void_p, C.c_int]),
    "mcpm_get_energy_cache": (C.c_int, [C.c_void_p, C.c_int, _dp, _ip]),
    "mcpm_use_cell_list": (C.c_int, [C.c_void_p, C.c_int]),
    "mcpm_apply_move": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp]),
    "mcpm_append": (C.c_int, [C.c_void_p, C.c_int, _dp]),
    "mcpm_remove_swap_last": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mcpm_run": (C.c_int, [C.c_void_p, C.POINTER(RunParams), C.POINTER(ChainStats)]),
    "mcpm_get_stats": (C.c_int, [C.c_void_p, C.POINTER(ChainStats)]),
    "mcpm_reset_accumulators": (C.c_int, [C.c_void_p]),
    "mcpm_get_rdf": (C.c_int, [C.c_void_p, C.c_int, _dp]),
    "mcpm_chemical_potential": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp]),
    "mcpm_set_running_energy": (C.c_int, [C.c_void_p, _dp, _dp]),
    "mcpm_set_stream_ids": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint)]),
    "mcpm_link_boxes": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mcpm_k1_resident": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mcpm_transfer_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
    "mcpm_k1_latency": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, _dp]),
    "mcpm_k1_resident_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
    "mcpm_set_dr": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "mcpm_set_step_counter": (C.c_int, [C.c_void_p, C.c_int, C.c_ulonglong]),
    "mcpm_last_kernel_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "mcpm_last_kernel_name": (C.c_char_p, [C.c_void_p]),
    "mcpm_metropolis_probe": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]),
    "mcpm_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_ulonglong)]),
    "mcpm_timer_start": (C.c_int, [C.c_void_p]),
    "mcpm_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "mcpm_fp64_peak": (C.c_int, [C.c_int, _dp, _dp]),
    "mcpm_read_input_file": (C.c_int, [C.c_char_p, C.POINTER(SimConfig)]),
    "mcpm_simulate": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int]),
    "mcpm_philox_uniforms": (C.c_int, [C.c_void_p, C.c_ulonglong, C.c_uint, C.c_ulonglong, C.c_int, _dp]),
}


EXPERIMENTAL = {
    "mcpm_use_subwarp": (C.c_int, [C.c_void_p, C.c_int]),
    "mcpm_use_dual_candidates": (C.c_int, [C.c_void_p, C.c_int]),
    "mcpm_use_pool": (C.c_int, [C.c_void_p, C.c_int]),
}


def lib():
    """Load libmcpm.so (built in-tree by __graft_entry__.build()).  Fails loudly if it is missing."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise McpmError(-1, f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                "— there is no CPU fallback")
        L = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        for name, (res, args) in EXPERIMENTAL.items():      # csrc/experiments/mcpm_experiments.h, libmcpm_exp.so only
            if hasattr(L, name):
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
        _LIB = L
    return _LIB


def has_experiments() -> bool:
    return all(hasattr(lib(), name) for name in EXPERIMENTAL)


def check(rc):
    if rc != 0:
        raise McpmError(rc, lib().mcpm_last_error().decode(errors="replace"))
