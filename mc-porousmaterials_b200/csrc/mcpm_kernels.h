// Host-callable launchers of the kernels in mcpm_kernels.cu.
#pragma once
#include <cuda_runtime.h>

#include "mcpm_types.h"

namespace mcpm {

void set_last_error(const char* msg);   // mcpm_api.cu: the per-thread message behind mcpm_last_error()

cudaError_t launch_k2(int rng_mode, int n_chains, int threads, bool fast, bool sample, bool cache, const ChainParams* params, ChainState* states, const int* order,
                      double* X, double* Y, double* Z, const RunArgs& a, int max_smem_optin, cudaStream_t stream);
size_t k2_smem_limit_cap(int max_smem_optin);
// k2_pool (EXPERIMENTS build only): persistent CTAs, one warp per chain, shared-memory slots of several capacity classes
size_t pool_state_stride_doubles();
cudaError_t launch_k2_pool(int rng_mode, bool pbz, int warps, int n_ctas, size_t smem, const ChainParams* params, ChainState* states, const int* order,
                           double* X, double* Y, double* Z, const RunArgs& a, const PoolLayout& L, cudaStream_t stream);
int k2_threads_supported(int threads, bool global, bool fast, bool sample);

// mcpm_k1.cu: MC::EnergyOfParticle for a batch of requests; positions staged by TMA bulk copies, results optionally written
// straight into mapped host memory followed by a sequence flag
cudaError_t launch_k1(bool fast, int n, int count, const ChainParams* params, const ChainState* states, const double* X, const double* Y,
                      const double* Z, int chain, int cap, const K1Request* reqs, const K1Request& single, double* scratch,
                      unsigned int* counters, double* out, double* host_out, unsigned long long* host_flag, unsigned long long seq,
                      unsigned int* done_groups, cudaStream_t stream);
int k1_slices(int n);
// resident K1: one persistent CTA serving one chain through a mailbox in mapped host memory
size_t k1_server_smem_bytes(int cap, int max_smem_optin);
cudaError_t launch_k1_server(bool fast, bool pbz, const ChainParams* params, const ChainState* states, double* X, double* Y, double* Z, int chain,
                             int cap, size_t smem, K1Mail* mail, K1Reply* reply, double* host_box, unsigned long long seq0, long long idle_cycles,
                             cudaStream_t stream);

cudaError_t launch_k3(int tiles, int n_chains, int chain0, const ChainParams* params, ChainState* states, const double* X, const double* Y,
                      const double* Z, int cap, double* per_pair, double* per_wall, double* scratch, unsigned int* counters, double* out,
                      cudaStream_t stream);

size_t k2_sub_smem_bytes(int cap_cta);
cudaError_t launch_k2_sub(int rng_mode, bool pbz, int n_ctas, const ChainParams* params, ChainState* states, const int* order, const int2* ctas,
                          double* X, double* Y, double* Z, const RunArgs& a, int cap_cta, cudaStream_t stream);

// mcpm_spec.cu: speculative-moves kernel, one CTA of 8 candidate warps + 2 helper warps per chain
size_t k2_spec_smem_bytes(int cap);
size_t k2_pair_smem_bytes(int cap);
cudaError_t launch_k2_spec(int rng_mode, bool pbz, int lean, int n_chains, const ChainParams* params, ChainState* states, const int* order,
                           double* X, double* Y, double* Z, const RunArgs& a, cudaStream_t stream);

// mcpm_dual.cu: one warp per chain, two candidate steps per round
size_t k2_dual_smem_bytes(int cap);
cudaError_t launch_k2_dual(int rng_mode, bool pbz, int n_chains, const ChainParams* params, ChainState* states, const int* order, double* X,
                           double* Y, double* Z, const RunArgs& a, cudaStream_t stream);

// mcpm_cells.cu: K2 through a cell list (large bulk-like chains), 256 threads per chain, positions in global memory
cudaError_t launch_k2_cells(int rng_mode, bool sample, int n_chains, const ChainParams* params, ChainState* states, const int* order, double* X,
                            double* Y, double* Z, const RunArgs& a, int* items, int* cell_of, int* slot_of, int cells_stride, int ccap,
                            size_t smem_items, cudaStream_t stream);
size_t k2_cells_smem_items_bytes(int cap, int cells_stride, int ccap, int max_smem_optin);

// mcpm_gibbs.cu: two linked boxes (Gibbs ensemble) per CTA; pairs[k] = (box 0's chain, box 1's chain)
cudaError_t launch_k2_gibbs(int rng_mode, int n_pairs, const ChainParams* params, ChainState* states, const int2* pairs, double* X, double* Y,
                            double* Z, const RunArgs& a, cudaStream_t stream);

cudaError_t launch_in_box(int n_chains, const ChainParams* params, const int* n_of, const double* X, const double* Y, const double* Z, int cap,
                          int* flags, cudaStream_t stream);
size_t k3c_cells_per_chain();
cudaError_t launch_k3_cells(int maxn, int n_chains, int chain0, const ChainParams* params, ChainState* states, const double* X,
                            const double* Y, const double* Z, int cap, int* cell_of, int* counts, int* starts, double* sx, double* sy,
                            double* sz, int* orig, double* per_pair, double* per_wall, double* scratch, unsigned int* counters, double* out,
                            cudaStream_t stream);

cudaError_t launch_fp64_peak(double* out, int blocks, int threads, int iters, cudaStream_t stream);
cudaError_t launch_metropolis_probe(int kind, int count, const double* u, const double* x, const int* n, double ln_zzV, double zzV, int* out,
                                    cudaStream_t stream);
cudaError_t launch_philox_fill(unsigned long long seed, unsigned int chain, unsigned long long first_step, int n_steps, double* out,
                               cudaStream_t stream);

}  // namespace mcpm
