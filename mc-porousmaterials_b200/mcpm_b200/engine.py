"""Engine: a context of independent chains (state points / replicas) on one GPU, over the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .abi import (BoxEnergy, ChainStats, ParticleEnergy, RunParams, SystemDesc, check, lib, RNG_PHILOX, RNG_REPLAY)

_dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_dp)


def device_count() -> int:
    n = C.c_int(0)
    check(lib().mcpm_device_count(C.byref(n)))
    return n.value


def fp64_peak_tflops(device: int = 0):
    t = C.c_double(0); ms = C.c_double(0)
    check(lib().mcpm_fp64_peak(device, C.byref(t), C.byref(ms)))
    return t.value, ms.value


class Engine:
    def __init__(self, n_chains: int, capacity: int, device: int = 0):
        self.L = lib()
        h = C.c_void_p()
        check(self.L.mcpm_create(device, n_chains, capacity, C.byref(h)))
        self.h = h
        nc = C.c_int(0); cap = C.c_int(0)
        check(self.L.mcpm_n_chains(self.h, C.byref(nc), C.byref(cap)))
        self.n_chains, self.capacity, self.device = nc.value, cap.value, device

    def close(self):
        if getattr(self, "h", None):
            self.L.mcpm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- systems and positions
    def set_system(self, desc, chain: int = -1):
        d = SystemDesc.from_any(desc)
        check(self.L.mcpm_set_system(self.h, chain, C.byref(d)))

    def get_system(self, chain: int) -> SystemDesc:
        d = SystemDesc()
        check(self.L.mcpm_get_system(self.h, chain, C.byref(d)))
        return d

    def set_positions(self, chain, x, y, z):
        x = np.ascontiguousarray(x, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
        z = np.ascontiguousarray(z, dtype=np.float64)
        check(self.L.mcpm_set_positions(self.h, chain, len(x), _p(x), _p(y), _p(z)))

    def get_positions(self, chain):
        n = C.c_int(0)
        x = np.zeros(self.capacity); y = np.zeros(self.capacity); z = np.zeros(self.capacity)
        check(self.L.mcpm_get_positions(self.h, chain, C.byref(n), _p(x), _p(y), _p(z)))
        return x[:n.value].copy(), y[:n.value].copy(), z[:n.value].copy()

    def set_positions_all(self, n, x, y, z):
        """n: int32[n_chains]; x,y,z: float64[n_chains, capacity] (may be pinned)."""
        n = np.ascontiguousarray(n, dtype=np.int32)
        assert x.shape == (self.n_chains, self.capacity) and x.dtype == np.float64 and x.flags.c_contiguous
        check(self.L.mcpm_set_positions_all(self.h, n.ctypes.data_as(C.POINTER(C.c_int)), _p(x), _p(y), _p(z)))

    def get_positions_all(self, x=None, y=None, z=None):
        if x is None:
            x = np.zeros((self.n_chains, self.capacity)); y = np.zeros_like(x); z = np.zeros_like(x)
        n = np.zeros(self.n_chains, dtype=np.int32)
        check(self.L.mcpm_get_positions_all(self.h, n.ctypes.data_as(C.POINTER(C.c_int)), _p(x), _p(y), _p(z)))
        return n, x, y, z

    # ---- K1 / K3
    def particle_energy(self, chain, index, trial=None):
        """-> (cur, moved) arrays {pair, many_body, wall, energy}; see mcpm_particle_energy."""
        cur = ParticleEnergy(); mv = ParticleEnergy()
        t = None
        if trial is not None:
            t = np.ascontiguousarray(trial, dtype=np.float64)
        check(self.L.mcpm_particle_energy(self.h, chain, index, _p(t) if t is not None else None, C.byref(cur), C.byref(mv)))
        f = lambda e: np.array([e.pair, e.many_body, e.wall, e.energy])
        return f(cur), f(mv)

    def particle_energy_batch(self, chain, index, trial=None):
        """-> (cur, moved), each [count, 4] = {pair, many_body, wall, energy}; see mcpm_particle_energy_batch."""
        index = np.ascontiguousarray(index, dtype=np.int32)
        m = len(index)
        t = np.ascontiguousarray(trial, dtype=np.float64).reshape(m, 3) if trial is not None else None
        cur = (ParticleEnergy * max(m, 1))(); mv = (ParticleEnergy * max(m, 1))()
        check(self.L.mcpm_particle_energy_batch(self.h, chain, m, index.ctypes.data_as(C.POINTER(C.c_int)), _p(t) if t is not None else None, cur, mv))
        f = lambda arr: np.array([[e.pair, e.many_body, e.wall, e.energy] for e in arr[:m]]).reshape(m, 4)
        return f(cur), f(mv)

    def box_energy(self, chain, per_particle=True):
        out = BoxEnergy()
        pp = np.zeros(self.capacity); pw = np.zeros(self.capacity)
        check(self.L.mcpm_box_energy(self.h, chain, C.byref(out), _p(pp) if per_particle else None, _p(pw) if per_particle else None))
        n = self.stats()[chain].n
        return np.array([out.pair, out.many_body, out.wall, out.energy]), pp[:n].copy(), pw[:n].copy()

    def box_energy_all(self):
        arr = (BoxEnergy * self.n_chains)()
        check(self.L.mcpm_box_energy_all(self.h, arr))
        return np.array([[e.pair, e.many_body, e.wall, e.energy] for e in arr])

    def use_speculation(self, mode=-1):
        """-1 automatic (few chains), 0 never, 1 whenever the launch qualifies, 2 two-warp variant for one-warp chains (mcpm_use_speculation)."""
        check(self.L.mcpm_use_speculation(self.h, int(mode)))

    def use_dual_candidates(self, enable=True):
        """Experimental kernel (libmcpm_exp.so only); disabling it is a no-op on the product library."""
        if enable or abi.has_experiments():
            check(self.L.mcpm_use_dual_candidates(self.h, 1 if enable else 0))

    def use_subwarp(self, enable=True):
        """Experimental kernel (libmcpm_exp.so only); disabling it is a no-op on the product library."""
        if enable or abi.has_experiments():
            check(self.L.mcpm_use_subwarp(self.h, 1 if enable else 0))

    def set_stream_ids(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        assert len(ids) == self.n_chains
        check(self.L.mcpm_set_stream_ids(self.h, ids.ctypes.data_as(C.POINTER(C.c_uint))))

    def k1_resident(self, chain, enable=True):
        """Serve mcpm_particle_energy / apply_move / append / remove_swap_last of `chain` from a persistent kernel (mcpm_k1_resident)."""
        check(self.L.mcpm_k1_resident(self.h, chain, 1 if enable else 0))

    def transfer_bytes(self):
        """(host -> device, device -> host) position bytes moved so far by set_positions_all / get_positions_all."""
        a, b = C.c_ulonglong(0), C.c_ulonglong(0)
        check(self.L.mcpm_transfer_bytes(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def k1_latency(self, chain, reps, trial=None):
        """Microseconds per mcpm_particle_energy call measured inside the library (no interpreter in the loop)."""
        us = C.c_double(0.0)
        t = None if trial is None else np.ascontiguousarray(trial, dtype=np.float64)
        check(self.L.mcpm_k1_latency(self.h, chain, reps, None if t is None else _p(t), C.byref(us)))
        return us.value

    def k1_resident_stats(self):
        a, b = C.c_ulonglong(0), C.c_ulonglong(0)
        check(self.L.mcpm_k1_resident_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def link_boxes(self, box0, box1):
        """Chains box0 / box1 become the two boxes of one Gibbs-ensemble system (mcpm_link_boxes)."""
        check(self.L.mcpm_link_boxes(self.h, box0, box1))

    def use_pool(self, warps=16):
        """Experimental pooled kernel (libmcpm_exp.so only); switching it off is a no-op on the product library."""
        if warps or abi.has_experiments():
            check(self.L.mcpm_use_pool(self.h, int(warps)))

    def use_energy_cache(self, enable=True):
        check(self.L.mcpm_use_energy_cache(self.h, 1 if enable else 0))

    def energy_cache(self, chain):
        """Per-particle pair energies (K) carried by the K2 energy cache, or None if no valid cache exists."""
        n = self.stats()[chain].n
        out = np.zeros(max(n, 1)); valid = C.c_int(0)
        check(self.L.mcpm_get_energy_cache(self.h, chain, _p(out), C.byref(valid)))
        return out[:n].copy() if valid.value else None

    def use_cell_list(self, enable=True):
        check(self.L.mcpm_use_cell_list(self.h, 1 if enable else 0))

    # ---- state mutation
    def apply_move(self, chain, index, xyz):
        t = np.ascontiguousarray(xyz, dtype=np.float64)
        check(self.L.mcpm_apply_move(self.h, chain, index, _p(t)))

    def append(self, chain, xyz):
        t = np.ascontiguousarray(xyz, dtype=np.float64)
        check(self.L.mcpm_append(self.h, chain, _p(t)))

    def remove_swap_last(self, chain, index):
        check(self.L.mcpm_remove_swap_last(self.h, chain, index))

    # ---- K2
    def run(self, first_set, n_sets, steps_per_set, seed=0, replay=None, trace=False, threads=0, check_energy=0, sample_rdf=0, no_sampling=0, stats_out=None):
        """Advance every chain n_sets*steps_per_set trial moves.  replay: float64[n_chains, stride] or None.
        stats_out: a (ChainStats * n_chains) array to fill instead of a fresh one (a caller in a timing loop)."""
        p = RunParams()
        p.first_set, p.n_sets, p.steps_per_set = first_set, n_sets, steps_per_set
        p.seed = seed
        p.threads_per_chain = threads
        p.check_energy = check_energy
        p.sample_rdf = sample_rdf
        p.no_sampling = no_sampling
        keep = []
        if replay is not None:
            r = np.ascontiguousarray(replay, dtype=np.float64)
            if r.ndim == 1:
                r = r.reshape(1, -1)
            assert r.shape[0] == self.n_chains
            p.rng_mode = RNG_REPLAY
            p.replay_u = _p(r); p.replay_stride = r.shape[1]
            keep.append(r)
        else:
            p.rng_mode = RNG_PHILOX
        tr = None
        if trace:
            tr = np.zeros((self.n_chains, n_sets * steps_per_set), dtype=np.uint8)
            p.trace = tr.ctypes.data_as(C.POINTER(C.c_ubyte))
        stats = stats_out if stats_out is not None else (ChainStats * self.n_chains)()
        check(self.L.mcpm_run(self.h, C.byref(p), stats))
        return (stats, tr) if trace else stats

    def stats(self):
        stats = (ChainStats * self.n_chains)()
        check(self.L.mcpm_get_stats(self.h, stats))
        return stats

    @staticmethod
    def stats_view(stats):
        """Structured numpy view (no copy) of an array of ChainStats: stats_view(st)['pair'] etc."""
        dt = np.dtype([(name, np.dtype(ct)) for name, ct in ChainStats._fields_], align=True)
        assert dt.itemsize == C.sizeof(ChainStats)
        return np.frombuffer(stats, dtype=dt)

    def reset_accumulators(self):
        check(self.L.mcpm_reset_accumulators(self.h))

    def get_rdf(self, chain):
        h = np.zeros(100)
        check(self.L.mcpm_get_rdf(self.h, chain, _p(h)))
        return h

    def chemical_potential(self, chain):
        a = C.c_double(0); b = C.c_double(0)
        check(self.L.mcpm_chemical_potential(self.h, chain, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_running_energy(self, pair, wall):
        pair = np.ascontiguousarray(pair, dtype=np.float64); wall = np.ascontiguousarray(wall, dtype=np.float64)
        assert len(pair) == len(wall) == self.n_chains
        check(self.L.mcpm_set_running_energy(self.h, _p(pair), _p(wall)))

    def set_dr(self, chain, dr):
        check(self.L.mcpm_set_dr(self.h, chain, dr))

    def set_step_counter(self, chain, step):
        check(self.L.mcpm_set_step_counter(self.h, chain, step))

    # ---- measurement
    def metropolis_probe(self, kind, u, x, n=None, zzV=1.0):
        """Device Metropolis decisions for arrays of uniforms u and arguments x (see mcpm_metropolis_probe)."""
        u = np.ascontiguousarray(u, dtype=np.float64); x = np.ascontiguousarray(x, dtype=np.float64)
        nn = np.ascontiguousarray(n if n is not None else np.zeros(len(u)), dtype=np.int32)
        out = np.zeros(len(u), dtype=np.int32)
        check(self.L.mcpm_metropolis_probe(self.h, int(kind), len(u), u.ctypes.data, x.ctypes.data, nn.ctypes.data, float(zzV), out.ctypes.data))
        return out.astype(bool)

    def last_kernel_name(self) -> str:
        return self.L.mcpm_last_kernel_name(self.h).decode()

    def last_kernel_ms(self) -> float:
        ms = C.c_float(0)
        check(self.L.mcpm_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        n = C.c_ulonglong(0)
        check(self.L.mcpm_launch_count(self.h, C.byref(n)))
        return n.value

    def timer_start(self):
        check(self.L.mcpm_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        check(self.L.mcpm_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def philox_uniforms(self, seed, chain, first_step, n_steps):
        out = np.zeros((n_steps, 8))
        check(self.L.mcpm_philox_uniforms(self.h, seed, chain, first_step, n_steps, _p(out)))
        return out
