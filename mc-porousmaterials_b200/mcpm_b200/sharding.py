"""Multi-GPU plumbing: chains shard over ranks with no data-path collective; only a final gather of the
per-chain accumulators (a few doubles per chain) crosses ranks.  Works with any
torch.distributed backend (nccl on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_chains(n_chains: int, rank: int, world: int):
    """Chain c is owned by rank c mod world."""
    return list(range(rank, n_chains, world))


def pack_accumulators(stats) -> np.ndarray:
    """[n_local, 6] = sum_n, sum_n2, sum_e, sum_e2, samples, n  from an array of ChainStats."""
    return np.array([[s.sum_n, s.sum_n2, s.sum_e, s.sum_e2, float(s.samples), float(s.n)] for s in stats], dtype=np.float64).reshape(-1, 6)


def gather_accumulators(dist, local: np.ndarray, n_chains: int, rank: int, world: int, device=None) -> np.ndarray | None:
    """All ranks call; rank 0 returns [n_chains, 6] in global chain order (others None)."""
    if dist is None or world == 1:
        return local.copy()
    import torch
    n_max = (n_chains + world - 1) // world
    buf = torch.zeros((n_max, 6), dtype=torch.float64, device=device)
    if len(local):
        buf[:len(local)] = torch.from_numpy(local).to(buf.device)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    if rank != 0:
        return None
    full = np.zeros((n_chains, 6))
    for r in range(world):
        own = shard_chains(n_chains, r, world)
        full[own] = out[r][:len(own)].cpu().numpy()
    return full


def isotherm_from_accumulators(acc: np.ndarray, n_points: int):
    """Chains are laid out point-major within a replica (chain = replica*n_points + point).
    Returns per point: mean N, standard error of the mean over replicas, mean energy."""
    n_chains = acc.shape[0]
    reps = n_chains // n_points
    a = acc[:reps * n_points].reshape(reps, n_points, 6)
    samples = np.maximum(a[:, :, 4], 1.0)
    mean_rep = a[:, :, 0] / samples                      # per replica <N>
    mean = mean_rep.mean(axis=0)
    sem = mean_rep.std(axis=0, ddof=1) / np.sqrt(reps) if reps > 1 else np.zeros(n_points)
    energy = (a[:, :, 2] / samples).mean(axis=0)
    return mean, sem, energy
