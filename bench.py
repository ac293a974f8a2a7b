#!/usr/bin/env python
"""bench.py — GCMC trial moves/s on BASELINE config C2 (40-point Ar adsorption isotherm).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's B200 engine
    python bench.py --impl reference [--gpus N] ...              # the reference's CPU path, host cores

A "step" is one SET of the reference's step loop (src/main.cpp:63-84): --steps-per-set trial moves
(translation / insertion / deletion, 1:1 disp:swap) for EVERY chain.  The headline workload per GPU is the
40-point isotherm (mu_k = -1800 + k*666/39 K, C1 system otherwise) with --replicas independent replica
chains per state point, started from the oracle-equilibrated configurations in
tests/golden/isotherm_c2_start.npz.  Chains shard over ranks with no collective on the data path (weak
scaling: every rank runs its own full set of chains; Philox streams are keyed by the global chain id).

Prints ONE JSON line (rank 0).  `value` = trial moves/s with the chain state resident in HBM, timed on
the device with CUDA events on the engine's stream (max over ranks); `e2e` = the same through the
C ABI with HOST buffers (pinned): upload all positions, run the set, download positions + statistics,
every step.  `roofline` is against the FP64 FMA pipe (measured live with a DFMA microbenchmark —
MEASURED_PEAKS.json carries no fp64 figure); `cpu_baseline` times the compiled reference on the host
cores on a bounded sample of the same workload.  Further sub-records of the same line:

  literal_isotherm   the workload north_star names literally: ONE chain per state point, 40 chains in total, sharded
                     chain -> rank (chain mod N) over the ranks present — STRONG scaling, emitted at every --gpus N;
  configs            (N = 1) short runs of the other BASELINE configs C1, C3, C4, C5: value, e2e, roofline fraction, kernel;
  k3, k1             (N = 1) full-energy kernel pair evaluations/s at N = 500 / 5 000 / 20 000 and the latency of one
                     EnergyOfParticle call through the C ABI.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "mc-porousmaterials_b200"))
sys.path.insert(0, ROOT)

FLOP_PER_PAIR_SLIT = 22      # algorithmic fp64 flop per in-cutoff pair evaluation, 2 periodic axes
BYTES_PER_PAIR = 24          # x,y,z fp64 of partner j
METRIC = "gcmc_trial_moves_per_s"
UNIT = "moves/s"
L2_NOTE = "GPU arm: 256 MiB device write between timed steps (L2 flush); CPU arm: not applicable"
MIN_WARMUP = 3               # timing rule of the bench contract, applied to both arms and reported as run


def load_start():
    g = np.load(os.path.join(ROOT, "tests", "golden", "isotherm_c2_start.npz"))
    mu = g["mu"]
    pts = [(float(mu[k]), g[f"x{k}"], g[f"y{k}"], g[f"z{k}"], float(g["dr"][k])) for k in range(len(mu))]
    return pts


def build_workload(config, replicas, steps_per_set):
    """BASELINE.json configs -> dict(points=[(SystemDesc, x, y, z, dr)], replicas, steps_per_set, ...).
    Chains are laid out point-major: chain = point * replicas + replica (see expand())."""
    from mcpm_b200 import workloads
    w = {"config": config, "flop_per_pair": FLOP_PER_PAIR_SLIT, "steps_per_set": steps_per_set, "equil_steps": 0}
    if config in ("c1", "c2"):
        pts = load_start()
        if config == "c1":
            pts = pts[-1:]                       # mu = -1134 K
            w["workload"] = "C1 Ar in graphite slit, single mu=-1134 K (H=20A, rcut=20A, T=87K, N~560, disp:swap=1:1)"
            w["replicas"] = replicas or 1776       # identical chains: one per resident slot (12 chains x 148 SMs); 1920 would run a second, nearly empty wave
        else:
            w["workload"] = "C2 40-point Ar isotherm (mu_k=-1800+k*666/39 K, H=20A, rcut=20A, T=87K, disp:swap=1:1)"
            w["replicas"] = replicas or 128
        w["points"] = [(workloads.argon_slit(mu=mu, n_equil_sets=0), x, y, z, dr) for (mu, x, y, z, dr) in pts]
        w["start"] = "tests/golden/isotherm_c2_start.npz"
    elif config == "c3":
        g = np.load(os.path.join(ROOT, "tests", "golden", "c3_start.npz"))
        w["workload"] = "C3 CH4 in slit, 300 K, mu=-3050 K, Rcut 92 (L=184A), N~5.4k, disp:swap=1:3"
        w["points"] = [(workloads.methane_slit(n_equil_sets=0), g["x"], g["y"], g["z"], float(g["dr"]))]
        w["replicas"] = replicas or 148
        w["steps_per_set"] = steps_per_set or 10000      # as C1 / C2 / C5 (round 1 used 2000 here only to keep its run short)
        w["start"] = "tests/golden/c3_start.npz"
    elif config == "c4":
        cx, cy, cz = workloads.c4_lattice()
        w["workload"] = "C4 dense NVT LJ argon, bulk L=100A, rcut=12A, T=120K, N=20000, translations + Widom sampling, one sweep (20000 trial moves) per set (K2 through a cell list: 27 cells of ~39 particles per scan instead of the reference's N)"
        w["points"] = [(workloads.argon_bulk(n_equil_sets=0), cx, cy, cz, 0.5)]
        w["replicas"] = replicas or 148
        w["steps_per_set"] = steps_per_set or 20000      # one sweep: the end-to-end copies (71 MB each way) are paid once per set, as the reference prints once per set
        w["flop_per_pair"] = 25
        w["start"] = "simple-cubic 28^3 lattice, first 20000 sites, +-0.2A jitter (numpy default_rng(2024))"
    elif config == "c5":
        rng = np.random.default_rng(12345)
        pts = []
        for d in workloads.replica_sweep():
            pts.append((d, *workloads.c5_start(d, rng), 0.1 * d.width[2]))
        w["workload"] = "C5 1024 replica chains per GPU: H in 8..39A x T in 77..139K, Rcut 17 (L=34A), mu(T)=T ln(8.44e-5 Lambda^3), N0=100"
        w["points"] = pts
        w["replicas"] = replicas or 1
        w["equil_steps"] = 200000              # untimed: random start -> equilibrated loadings
        w["start"] = "100 random particles per chain (numpy default_rng(12345)), 200k untimed equilibration steps on the device"
    else:
        raise SystemExit(f"unknown config {config}")
    w["steps_per_set"] = w["steps_per_set"] or 10000
    return w


def config_record(w):
    """The `config` object — identical in both arms (what differs between them goes into `arm`)."""
    return {"workload": w["workload"], "state_points": len(w["points"]), "start": w["start"], "l2": L2_NOTE}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([v.strip() for v in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        deadline = time.time() + 1.5            # a timed region shorter than nvidia-smi's start-up: wait for its first row (GPU still busy
        while not self.rows and time.time() < deadline and self.proc.poll() is None:   # or just released: clocks do not drop within ms)
            time.sleep(0.02)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist_mod.init_process_group(backend="nccl")
        dist = dist_mod
    return world, rank, local, dist


def allmax(dist, local, v):
    if dist is None:
        return v
    import torch
    t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allsum(dist, local, v):
    if dist is None:
        return v
    import torch
    t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier(dist, local):
    import torch
    if dist is not None:
        dist.barrier(device_ids=[local])
    torch.cuda.synchronize(local)


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_points(w, max_points=40):
    """A bounded sample of the workload's state points for the CPU arm (all of them for C1-C4)."""
    pts = w["points"]
    idx = list(range(len(pts))) if len(pts) <= max_points else [int(round(k * (len(pts) - 1) / (max_points - 1))) for k in range(max_points)]
    out = []
    for k in idx:
        s, x, y, z, dr = pts[k]
        d = {name: (list(getattr(s, name)) if name == "width" else getattr(s, name)) for name, _ in s._fields_ if name != "reserved"}
        out.append((k, d, x, y, z, dr, 5000 + k))
    return out


def run_cpu_sample(pool, w, steps, correct=1):
    return pool.time_points(cpu_points(w), steps, correct=correct)


def cpu_steps_for(w, args):
    """Steps per state point of the CPU sample, sized for ~10-30 s of CPU work (cost per step ~ N, N^2 recomputes)."""
    if args.cpu_steps:
        return args.cpu_steps
    nmax = max(len(p[1]) for p in w["points"])
    return 4000 if nmax < 1000 else (150 if nmax < 10000 else 20)


def reference_arm(args, world, rank):
    if rank != 0:
        return
    from oracle import cpu_baseline
    w = build_workload(args.config, args.replicas, args.steps_per_set)
    steps = cpu_steps_for(w, args)
    pool = cpu_baseline.CpuPool()
    t_all = []
    res = None
    for it in range(args.warmup + args.steps):
        res = run_cpu_sample(pool, w, steps, correct=1)
        if it >= args.warmup:
            t_all.append(res)
    pool.close()
    total_steps = sum(r["total_steps"] for r in t_all)
    cpu_seconds = sum(r["cpu_seconds"] for r in t_all)
    cores = t_all[0]["cores"]
    value = total_steps * cores / cpu_seconds
    sample = f"{res['points']} state points x {steps} steps of the as-shipped loop (MoveParticle/SwapParticle + CorrectEnergy) per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * cpu_seconds / cores / len(t_all), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_record(w),
        "arm": {"what": "the reference's own CPU path (oracle/_ref/libmcref.so when built, else the C port), one single-threaded process per host core",
                "state_points_sampled": res["points"], "steps_per_point_per_step": steps},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": res["kind"], "sample": sample,
                         "per_core": total_steps / cpu_seconds, "makespan_rate": res["moves_per_s_makespan"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
class Job:
    """One context of chains on one GPU + pinned host mirrors of their positions."""

    def __init__(self, m, torch, chains, steps_per_set, device, cap, stream_ids, args, equil_steps=0, seed=20240611, speculation=None):
        self.m, self.torch, self.device, self.S, self.seed, self.threads = m, torch, device, steps_per_set, seed, args.threads
        n_chains = len(chains)
        self.n_chains = n_chains
        self.eng = m.Engine(n_chains, cap, device=device)
        if speculation is not None:
            self.eng.use_speculation(speculation)
        if getattr(args, "pool", None) is not None:
            self.eng.use_pool(args.pool)
        self.cap = cap = self.eng.capacity
        self.n0 = np.zeros(n_chains, dtype=np.int32)
        hx = torch.zeros((n_chains, cap), dtype=torch.float64).pin_memory()
        hy = torch.zeros_like(hx).pin_memory(); hz = torch.zeros_like(hx).pin_memory()
        self._pins = (hx, hy, hz)
        self.X, self.Y, self.Z = hx.numpy(), hy.numpy(), hz.numpy()
        for c, (sysd, x, y, z, dr) in enumerate(chains):
            self.eng.set_system(sysd, c)
            self.eng.set_dr(c, dr)
            self.n0[c] = len(x)
            self.X[c, :len(x)] = x; self.Y[c, :len(x)] = y; self.Z[c, :len(x)] = z
        self.eng.set_stream_ids(stream_ids)
        self.eng.set_positions_all(self.n0, self.X, self.Y, self.Z)
        self.set_index = 1
        if equil_steps:
            self.eng.run(1, 1, equil_steps, seed=seed + 1, threads=self.threads, check_energy=2)
            self.eng.reset_accumulators()

    def close(self):
        self.eng.close()

    def one_step(self):
        if getattr(self, "_stats", None) is None:
            self._stats = (self.m.ChainStats * self.eng.n_chains)()      # one record buffer for the whole job (2 MB for 5120 chains)
        st = self.eng.run(self.set_index, 1, self.S, seed=self.seed, threads=self.threads, check_energy=0, no_sampling=getattr(self, "no_sampling", 0),
                          stats_out=self._stats)
        self.set_index += 1
        return st

    def time_resident(self, steps, warmup, flush, dist):
        """W untimed + K timed steps, chain state resident in HBM; CUDA events on the engine's stream."""
        eng, torch, local = self.eng, self.torch, self.device
        for _ in range(warmup):
            self.one_step()
        eng.reset_accumulators()
        launches0 = eng.launch_count()
        barrier(dist, local)
        dev_ms, kern_ms, evals, evals_x = [], [], [], []
        prev_evals = prev_x = 0
        t_wall0 = time.perf_counter()
        for _ in range(steps):
            flush.zero_()                       # L2 flush between timed iterations (256 MiB write > 126 MB L2)
            torch.cuda.synchronize(local)
            eng.timer_start()
            st = self.one_step()
            dev_ms.append(eng.timer_stop())
            kern_ms.append(eng.last_kernel_ms())
            sv = eng.stats_view(st)
            tot = int(sv["pair_evals_algorithmic"].sum()); totx = int(sv["pair_evals"].sum())
            evals.append(tot - prev_evals); prev_evals = tot
            evals_x.append(totx - prev_x); prev_x = totx
        barrier(dist, local)
        stats = eng.stats()
        sv = eng.stats_view(stats)
        mean_n = float(np.mean(sv["sum_n"] / np.maximum(sv["samples"], 1))) if stats[0].samples else float(np.mean(sv["n"]))
        return {"dev_s": sum(dev_ms) * 1e-3, "kern_ms": float(np.mean(kern_ms)), "evals": float(np.mean(evals)), "evals_x": float(np.mean(evals_x)),
                "evals_sum": float(sum(evals)), "evals_x_sum": float(sum(evals_x)), "launches": eng.launch_count() - launches0,
                "wall_s": time.perf_counter() - t_wall0, "mean_n": mean_n, "kernel": eng.last_kernel_name(), "stats": stats}

    def time_e2e(self, steps, dist):
        """The same steps through the C ABI with HOST buffers: H2D of every chain's positions and running energies, K2, D2H of
        positions and statistics — every step, inside the timed region."""
        eng, torch, local, m = self.eng, self.torch, self.device, self.m
        nbuf = np.zeros(self.n_chains, dtype=np.int32)
        n_host, _, _, _ = eng.get_positions_all(self.X, self.Y, self.Z)
        nbuf[:] = n_host
        sv = eng.stats_view(eng.stats())
        e_pair = sv["pair"].copy(); e_wall = sv["wall"].copy()   # the host owns the state
        barrier(dist, local)
        e2e_s = 0.0
        b0 = eng.transfer_bytes()
        for _ in range(steps):
            torch.cuda.synchronize(local)
            t0 = time.perf_counter()
            eng.set_positions_all(nbuf, self.X, self.Y, self.Z)       # H2D: every chain's positions, pinned host memory
            eng.set_running_energy(e_pair, e_wall)                    # ... and their running energies (else K3 re-evaluates them)
            st = self.one_step()                                      # K2
            n_host, _, _, _ = eng.get_positions_all(self.X, self.Y, self.Z)   # D2H: positions; statistics came back with run()
            nbuf[:] = n_host
            sv = eng.stats_view(st)
            e_pair = sv["pair"].copy(); e_wall = sv["wall"].copy()
            e2e_s += time.perf_counter() - t0
        barrier(dist, local)
        b1 = eng.transfer_bytes()                                         # position bytes as the library counted them (runs of similar chains)
        stat_bytes = ctypes.sizeof(m.ChainStats)
        h2d = (b1[0] - b0[0]) // max(1, steps) + self.n_chains * (4 + 16 + stat_bytes)
        d2h = (b1[1] - b0[1]) // max(1, steps) + self.n_chains * stat_bytes
        return e2e_s, h2d, d2h


def expand(w):
    """Point-major chain list of a workload (the replicas of a state point are neighbours, as in the host driver: chain = point * R + replica):
    neighbouring chains then have similar loadings, which is what the bulk position transfers exploit."""
    pts, R = w["points"], w["replicas"]
    return [pts[c // R] for c in range(len(pts) * R)]


def roofline_record(w, r, peak_tflops, traffic=None, traffic_source=None):
    k_ms, evals, evals_x, kernel = r["kern_ms"], r["evals"], r["evals_x"], r["kernel"]
    # The roofline counts the reference's (algorithmic) pair evaluations, except through the cell list: there the engine evaluates
    # ~1/19 of them (a different algorithm), and a utilisation figure must count what the FP64 pipe actually executed.
    roof_evals = evals_x if kernel == "k2_cells" else evals
    achieved = w["flop_per_pair"] * roof_evals / (k_ms * 1e-3) * 1e-12
    return {"bound": "fp64", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
            "frac_executed": w["flop_per_pair"] * evals_x / (k_ms * 1e-3) * 1e-12 / peak_tflops,
            "traffic": traffic, "traffic_source": traffic_source, "kernel": kernel, "kernel_ms": k_ms,
            "algorithmic": (f"{w['flop_per_pair']} flop x {evals:.4g} pair evaluations per launch = what the reference's moves evaluate for "
                            f"the same steps (2N per translation, N per insertion/deletion, SURVEY.md 8d); the engine executed "
                            f"{evals_x:.4g} of them (per-particle energy cache)") if kernel != "k2_cells" else
                           (f"{w['flop_per_pair']} flop x {evals_x:.4g} EXECUTED pair evaluations per launch (cell list: 27 cells around the "
                            f"position); the reference's moves evaluate {evals:.4g} for the same steps (2N per translation, SURVEY.md 8d)"),
            "peak_source": "mcpm_fp64_peak DFMA microbenchmark, measured live (MEASURED_PEAKS.json has no fp64 figure)",
            "pair_evals_per_s": evals / (k_ms * 1e-3),
            "smem_bytes_per_s": BYTES_PER_PAIR * evals / (k_ms * 1e-3)}


def literal_isotherm(m, torch, args, world, rank, local, dist, flush):
    """North_star's literal target: the 40-point isotherm with ONE chain per state point, points sharded chain -> rank
    (chain mod N).  Strong scaling: the total work is fixed, each rank runs 40/N chains, and since every chain
    already has an SM (or a whole CTA of candidate warps) to itself the time is bounded by the slowest chain's step latency."""
    from mcpm_b200 import workloads
    w = build_workload("c2", 1, args.steps_per_set)
    pts = w["points"]
    mine = [c for c in range(len(pts)) if c % world == rank]
    nmax = max(len(p[1]) for p in pts)
    job = Job(m, torch, [pts[c] for c in mine], w["steps_per_set"], local, workloads.bench_capacity(nmax), mine, args)
    r = job.time_resident(args.steps, args.warmup, flush, dist)
    e2e_s, h2d, d2h = job.time_e2e(args.steps, dist)
    job.close()
    t_dev = allmax(dist, local, r["dev_s"])
    e2e_s = allmax(dist, local, e2e_s)
    moves = float(len(pts)) * w["steps_per_set"] * args.steps
    slowest = max((st.n, c) for st, c in zip(r["stats"], mine))
    return {"what": "40 chains in total (one per state point), chain c on rank c mod N, strong scaling", "chains_total": len(pts),
            "chains_this_rank": len(mine), "value": moves / t_dev, "unit": UNIT, "scaling": "strong", "ms_per_step": 1e3 * t_dev / args.steps,
            "us_per_trial_move_per_chain": 1e6 * t_dev / (args.steps * w["steps_per_set"]),
            "e2e": {"value": moves / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "kernel": r["kernel"], "gpu_launches": int(r["launches"]),
            "bound": f"step latency of the largest chain on rank 0 (N = {slowest[0]} particles, state point {slowest[1]}): a Markov chain is "
                     "sequential, so more GPUs only remove neighbours that never competed for its SM"}


def k3_k1_records(m, torch, local):
    """Full-energy kernel (MC::BoxEnergy, N(N-1) pair evaluations as the reference counts them) at N = 500 / 5 000 / 20 000, and the
    latency of one MC::EnergyOfParticle call (K1) through the C ABI — the reference does ~8e7 pair evaluations/s per core."""
    from mcpm_b200 import workloads
    out = {"k3": [], "k1": []}
    pts = load_start()
    g3 = np.load(os.path.join(ROOT, "tests", "golden", "c3_start.npz"))
    cx, cy, cz = workloads.c4_lattice()
    cases = [("C1 N~560", workloads.argon_slit(n_equil_sets=0), pts[-1][1], pts[-1][2], pts[-1][3], 1024, True),
             ("C3 N~5.4k", workloads.methane_slit(n_equil_sets=0), g3["x"], g3["y"], g3["z"], 148, True),
             ("C4 N=20000 all pairs", workloads.argon_bulk(n_equil_sets=0), cx, cy, cz, 16, False),
             ("C4 N=20000 cell list", workloads.argon_bulk(n_equil_sets=0), cx, cy, cz, 16, True)]
    for name, sysd, x, y, z, chains, cells in cases:
        n = len(x)
        with m.Engine(chains, n + 32, device=local) as e:
            e.use_cell_list(cells)
            e.set_system(sysd)
            for c in range(chains):
                e.set_positions(c, x, y, z)
            ms = []
            for it in range(6):
                e.box_energy_all()
                if it >= 2:
                    ms.append(e.last_kernel_ms())
            t = float(np.mean(ms)) * 1e-3
            out["k3"].append({"case": name, "n": n, "chains": chains, "kernel_ms": t * 1e3, "pair_evals_algorithmic_per_s": chains * n * (n - 1.0) / t,
                              "kernel": "k3c_energy (cell list)" if (cells and n >= 8192) else "k3_box_energy"})
            if chains >= 148:
                rng = np.random.default_rng(1)
                idx = rng.integers(0, n, size=200)
                trial = np.column_stack([rng.random(200) * sysd.width[0], rng.random(200) * sysd.width[1], 3.0 + rng.random(200) * (sysd.width[2] - 6.0)])
                for i in range(20):
                    e.particle_energy(0, int(idx[i]), trial[i])
                e.timer_start()
                t0 = time.perf_counter()
                for i in range(200):
                    e.particle_energy(0, int(idx[i]), trial[i])
                dt = time.perf_counter() - t0
                dev_ms = e.timer_stop()
                t0 = time.perf_counter()
                e.particle_energy_batch(0, idx.astype(np.int32), trial)
                dtb = time.perf_counter() - t0
                e.k1_resident(0, True)                        # the persistent one-CTA service: no launch per call
                for i in range(20):
                    e.particle_energy(0, int(idx[i]), trial[i])
                t0 = time.perf_counter()
                for i in range(200):
                    e.particle_energy(0, int(idx[i]), trial[i])
                dtr = time.perf_counter() - t0
                native_resident = e.k1_latency(0, 2000, trial[0])
                e.k1_resident(0, False)
                native_launch = e.k1_latency(0, 500, trial[0])
                out["k1"].append({"case": name, "n": n, "us_per_call_through_abi": 1e6 * dt / 200, "us_device_per_call": 1e3 * dev_ms / 200,
                                  "us_per_call_resident_service": 1e6 * dtr / 200,
                                  "us_per_call_native_caller": {"launch_per_call": native_launch, "resident_service": native_resident},
                                  "us_per_request_in_a_batch_of_200": 1e6 * dtb / 200, "pair_evals_per_call": 2 * n,
                                  "what": "mcpm_particle_energy: stored + displaced position in one launch (positions staged by TMA bulk copies, "
                                          "results through mapped host memory); ctypes call overhead included"})
    return out


def gibbs_record(m):
    """Row f4: two linked boxes exchanging particles (the Gibbs ensemble at fixed volumes, k2_gibbs), 148 systems = 296 chains from the
    liquid + vapour start of tests/golden/reference_gibbs.npz; one step = 2500 trial moves (translations in either box + transfers) per system."""
    import json
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_gibbs.npz"))
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_gibbs.json")))["vle"]
    pairs, sps = 148, 2500
    with m.Engine(2 * pairs, 1024) as e:
        for k in range(pairs):
            for b in (0, 1):
                d = m.SystemDesc()
                for key, val in meta[f"box{b}"].items():
                    if key == "width":
                        for a in range(3):
                            d.width[a] = val[a]
                    else:
                        setattr(d, key, val)
                e.set_system(d, 2 * k + b)
                e.set_positions(2 * k + b, g[f"vle_x{b}"], g[f"vle_y{b}"], g[f"vle_z{b}"])
            e.link_boxes(2 * k, 2 * k + 1)
        ids = np.arange(2 * pairs, dtype=np.uint32)
        e.set_stream_ids(ids)
        e.run(1, 1, sps, seed=99)
        ms = []
        for it in range(3):
            e.run(2 + it, 1, sps, seed=99)
            ms.append(e.last_kernel_ms())
        st = e.stats()
        t = float(np.mean(ms)) * 1e-3
        return {"what": "two linked boxes per system (Gibbs ensemble at fixed volumes: translations in either box + particle transfers), LJ argon liquid + vapour boxes, "
                        "148 systems = 296 chains, 2500 trial moves per system per step", "systems": pairs, "mean_particles_box0": float(np.mean([st[2 * k].n for k in range(pairs)])),
                "mean_particles_box1": float(np.mean([st[2 * k + 1].n for k in range(pairs)])), "value": pairs * sps / t, "unit": UNIT, "kernel": e.last_kernel_name(),
                "kernel_ms": t * 1e3}


def dbg(*a):
    if os.environ.get("BENCH_DEBUG"):
        print("[bench]", *a, file=sys.stderr, flush=True)


def gpu_arm(args, world, rank, local, dist):
    import torch
    import mcpm_b200 as m
    from mcpm_b200 import workloads

    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=f"cuda:{local}")
    w = build_workload(args.config, args.replicas, args.steps_per_set)
    chains = expand(w)
    P, R, S = len(w["points"]), w["replicas"], w["steps_per_set"]
    n_chains = len(chains)
    nmax = max(len(p[1]) for p in w["points"])
    cap = args.capacity or (1024 if args.config == "c5" else workloads.bench_capacity(nmax))   # headroom for fluctuations of N; overflow is an error
    ids = np.arange(n_chains, dtype=np.uint32) + np.uint32(rank * n_chains)      # every rank: its own replicas of the workload
    job = Job(m, torch, chains, S, local, cap, ids, args, equil_steps=w["equil_steps"], speculation=args.speculation)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    r = job.time_resident(args.steps, args.warmup, flush, dist)
    e2e_s, h2d, d2h = job.time_e2e(args.steps, dist)
    clocks = sampler.stop() if rank == 0 else None     # sampled over both timed regions (device-resident and end-to-end)
    cap = job.cap
    job.close()
    t_dev = allmax(dist, local, r["dev_s"])
    e2e_s = allmax(dist, local, e2e_s)
    moves_per_rank = float(n_chains) * S * args.steps
    value = world * moves_per_rank / t_dev
    evals_all = allsum(dist, local, r["evals_sum"])
    evals_x_all = allsum(dist, local, r["evals_x_sum"])

    dbg(rank, "main workload timed", value)
    literal = literal_isotherm(m, torch, args, world, rank, local, dist, flush) if args.config == "c2" and not args.no_literal else None
    dbg(rank, "literal isotherm timed")
    if rank != 0:
        return
    peak_tflops, _ = m.fp64_peak_tflops(local)
    traffic = traffic_source = None
    summ = os.path.join(ROOT, "profiles", "k2_ncu_summary.json")
    if os.path.exists(summ) and r["kernel"] == "k2_chains" and args.config == "c2" and not args.replicas:   # the capture is of the default workload
        try:
            traffic = json.load(open(summ)).get("dram_bytes_per_launch")
            traffic_source = "profiles/k2_ncu_summary.json: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this command, NOT measured in this run"
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_record(w),
        "arm": {"what": "B200 engine (libmcpm.so), device Philox4x32-10", "replicas_per_point_per_gpu": R, "chains_per_gpu": n_chains,
                "steps_per_set": S, "capacity": cap, "mean_particles": r["mean_n"], "threads_per_chain": args.threads or "auto"},
        "pair_evals_per_s": evals_all / t_dev, "pair_evals_executed_per_s": evals_x_all / t_dev,
        "wall_s_timed_region": r["wall_s"],
        "e2e": {"value": world * moves_per_rank / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(r["launches"]),
        "clocks": clocks,
        "roofline": roofline_record(w, r, peak_tflops, traffic, traffic_source),
    }
    if literal is not None:
        line["literal_isotherm"] = literal
    if world == 1 and args.config == "c2" and not args.no_configs:
        recs = []
        for cfg in ("c1", "c3", "c4", "c5"):
            wc = build_workload(cfg, 0, 0)
            nm = max(len(p[1]) for p in wc["points"])
            chains_c = expand(wc)
            jc = Job(m, torch, chains_c, wc["steps_per_set"], local, 1024 if cfg == "c5" else workloads.bench_capacity(nm),
                     np.arange(len(chains_c), dtype=np.uint32), args, equil_steps=wc["equil_steps"])
            k = min(args.steps, 3)
            rc = jc.time_resident(k, MIN_WARMUP, flush, None)
            es, hb, db = jc.time_e2e(k, None)
            moves_only = None
            if cfg == "c4":      # the same sets as EQUILIBRATION sets (set <= nEquilSets: no ComputeWidom after the move, src/main.cpp:72-75)
                jc.no_sampling = 1
                moves_only = float(len(chains_c)) * wc["steps_per_set"] * k / jc.time_resident(k, 1, flush, None)["dev_s"]
            jc.close()
            mv = float(len(chains_c)) * wc["steps_per_set"] * k
            rf = roofline_record(wc, rc, peak_tflops)
            recs.append({"config": cfg, "workload": wc["workload"], "chains": len(chains_c), "steps_per_set": wc["steps_per_set"], "steps": k,
                         "mean_particles": rc["mean_n"], "value": mv / rc["dev_s"], "unit": UNIT,
                         "e2e": {"value": mv / es, "unit": UNIT, "h2d_bytes_per_step": hb, "d2h_bytes_per_step": db},
                         "kernel": rc["kernel"], "roofline_frac": rf["frac"], "roofline_frac_executed": rf["frac_executed"],
                         "pair_evals_per_s": rf["pair_evals_per_s"]})
            if moves_only is not None:
                recs[-1]["value_equilibration_sets"] = moves_only
        line["configs"] = recs
        line.update(k3_k1_records(m, torch, local))
        line["f4_two_boxes"] = gibbs_record(m)
    if world == 1 and not args.no_cpu:
        from oracle import cpu_baseline
        pool = cpu_baseline.CpuPool()
        cs = cpu_steps_for(w, args)
        res = run_cpu_sample(pool, w, cs, correct=1)
        fair = run_cpu_sample(pool, w, cs, correct=0)
        pool.close()
        line["cpu_baseline"] = {
            "value": res["moves_per_s"], "unit": UNIT, "cores": res["cores"], "kind": res["kind"],
            "sample": f"{res['points']} state points x {cs} steps each of the as-shipped loop (MoveParticle/SwapParticle + CorrectEnergy), "
                      f"one reference process per core, perfectly balanced aggregate",
            "per_core": res["moves_per_s_per_core"], "makespan_rate": res["moves_per_s_makespan"],
            "without_correctenergy_recompute": fair["moves_per_s"],
        }
    dbg(rank, "printing")
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--replicas", type=int, default=0, help="replica chains per state point per GPU")
    ap.add_argument("--steps-per-set", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0, help="threads per chain (0 = engine heuristic)")
    ap.add_argument("--cpu-steps", type=int, default=0, help="steps per state point of the CPU sample (0 = sized from N)")
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5"], help="BASELINE.json config (default: the isotherm)")
    ap.add_argument("--capacity", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short C1/C3/C4/C5 runs and the K3/K1 records")
    ap.add_argument("--no-literal", action="store_true", help="skip the literal 40-chain isotherm sub-record")
    ap.add_argument("--speculation", type=int, default=None, help="mcpm_use_speculation mode (experiments)")
    ap.add_argument("--pool", type=int, default=None, help="mcpm_use_pool (experiments build, MCPM_EXPERIMENTS=1): warps per persistent CTA, 0 = one CTA per chain")
    args = ap.parse_args()
    if args.warmup < MIN_WARMUP:
        print(f"bench.py: --warmup {args.warmup} raised to {MIN_WARMUP} (timing rule: at least 3 warm-up steps; both arms)", file=sys.stderr)
        args.warmup = MIN_WARMUP
    if args.impl == "reference":
        world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
        reference_arm(args, world, rank)
        return
    world, rank, local, dist = dist_setup(args.gpus)
    try:
        gpu_arm(args, world, rank, local, dist)
    finally:
        if dist is not None:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
