"""TEST / MEASUREMENT INFRASTRUCTURE — NOT PART OF THE PRODUCT.

Times the reference's own CPU implementation of the hot path on the host cores: the body of the step
loop src/main.cpp:66-71 (MoveParticle / SwapParticle + CorrectEnergy, "as shipped") for a set of state
points, one single-threaded reference process per worker (the reference has no threading; an isotherm
is run as independent processes).  Uses oracle/_ref/libmcref.so (kind "reference") when it was built,
else the C port oracle/liboracle.so (kind "port").  Only bench.py may call this.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

from . import oracle as O


_REF = None
_KIND = None


def _init(kind):
    """Pool initializer: ONE reference object per worker process (5.6 GB, ~6 s to construct)."""
    global _REF, _KIND
    _KIND = kind
    if kind == "reference":
        _REF = O.RefOracle()


def _point(args):
    (k, sysd, x, y, z, dr, seed), steps, correct = args
    s = O.make_system(**sysd)
    s.dr = dr
    if _KIND == "reference":
        ref = _REF
        ref.setup(s)
        ref.set_positions(x, y, z)
        ref.seed(seed)
        ref.box_energy()          # as main.cpp does before the loop (MinimizeEnergy -> BoxEnergy)
        ref.reset_stats()
        dt = ref.run_timed(steps, correct)
        n_end = ref.state()["n"]
    else:
        c = O.COracle(s, 4096)
        c.set_positions(x, y, z)
        c.seed(seed)
        c.box_energy()
        c.reset_stats()
        t0 = time.perf_counter()
        c.run(steps, correct)
        dt = time.perf_counter() - t0
        n_end = c.state()["n"]
        c.close()
    return k, dt, len(x), n_end, os.getpid()


def available_kind() -> str:
    return "reference" if O.have_reference_lib() else "port"


def max_workers(kind: str) -> int:
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    if kind == "reference":
        import psutil
        by_mem = int(psutil.virtual_memory().available // (7 * 2 ** 30))   # 5.6 GB touched per MC object
        cores = max(1, min(cores, by_mem))
    return cores


class CpuPool:
    """Persistent worker processes, one single-threaded reference instance each."""

    def __init__(self, workers: int | None = None, kind: str | None = None):
        self.kind = kind or available_kind()
        self.workers = workers or max_workers(self.kind)
        self.pool = mp.get_context("spawn").Pool(self.workers, initializer=_init, initargs=(self.kind,))

    def close(self):
        self.pool.close()
        self.pool.join()

    def time_points(self, points, steps: int, correct: int = 1):
        """points: list of (key, system_dict, x, y, z, dr, seed).  Every point runs `steps` passes of the
        step-loop body on whichever worker is free (longest expected first).  Returns per-point seconds
        and aggregate rates."""
        order = sorted(points, key=lambda p: len(p[2]), reverse=True)
        t0 = time.perf_counter()
        res = self.pool.map(_point, [(p, steps, correct) for p in order], chunksize=1)
        wall = time.perf_counter() - t0
        per_point, busy = {}, {}
        for k, dt, n0, n1, pid in res:
            per_point[k] = {"seconds": dt, "n_start": n0, "n_end": n1, "moves_per_s": steps / dt if dt > 0 else float("inf")}
            busy[pid] = busy.get(pid, 0.0) + dt
        total_steps = steps * len(points)
        cpu_seconds = sum(busy.values())
        return {
            "kind": self.kind, "cores": self.workers, "steps_per_point": steps, "points": len(points),
            "total_steps": total_steps, "cpu_seconds": cpu_seconds, "makespan_s": max(busy.values()), "wall_s": wall,
            # every core busy on independent state points, perfectly balanced (generous to the CPU)
            "moves_per_s": total_steps * self.workers / cpu_seconds,
            "moves_per_s_makespan": total_steps / max(busy.values()),
            "moves_per_s_per_core": total_steps / cpu_seconds,
            "per_point": per_point,
        }
