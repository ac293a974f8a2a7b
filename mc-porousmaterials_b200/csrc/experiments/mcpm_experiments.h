/* Switches of the EXPERIMENTAL K2 kernels (csrc/experiments/, built only by `make EXPERIMENTS=1` into libmcpm_exp.so).
 * Not part of the ABI a maintainer binds (include/mcpm.h): both kernels reproduce the golden traces of the compiled reference
 * but were measured SLOWER than the default kernels on B200, so they can never be the default.
 *
 *   mcpm_use_subwarp          2 or 4 small chains packed into one warp (16 / 8 lanes each), translation / insertion / deletion
 *                             as one branch-free path; chains that outgrow their slot finish the call in a warp of their own.
 *                             1.7-2x slower than one warp per chain (round 1, config C2).
 *   mcpm_use_dual_candidates  one warp per chain evaluates TWO consecutive steps against the same configuration in one
 *                             interleaved instruction stream; the second is used only if the first is rejected.  25 % slower
 *                             (168 registers cost a resident chain per SM; selection and decision stay serial).
 *   mcpm_use_pool             (round 2) k2_pool: one persistent CTA per SM whose warps own shared-memory slots of up to three capacity
 *                             classes and take chains from work queues, longest first; warps = 12 / 14 / 16 / 20 per CTA, 0 = off.  Equal to
 *                             k2_chains at 12 warps, slower at 16 (register pressure): profiles/r02_pool.md. */
#ifndef MCPM_EXPERIMENTS_H
#define MCPM_EXPERIMENTS_H
#include "../../../include/mcpm.h"
#ifdef __cplusplus
extern "C" {
#endif
int mcpm_use_subwarp(mcpm_ctx* ctx, int enable);
int mcpm_use_dual_candidates(mcpm_ctx* ctx, int enable);
int mcpm_use_pool(mcpm_ctx* ctx, int warps);
#ifdef __cplusplus
}
#endif
#endif
