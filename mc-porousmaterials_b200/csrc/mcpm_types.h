// Device-resident per-chain records shared by the host API (mcpm_api.cu) and the kernels.
#pragma once
#include "../../include/mcpm.h"

// Immutable during a run: what the reference keeps in Box / Fluid / ThermodynamicSystem
// (src/MC.h:62-85) for one box and one species, pre-digested for the kernels.
struct ChainParams {
    double L[3];          // Box::width
    double invL[3];
    double halfH;         // 0.5*width[2]           (sizeR, src/potentials.cpp:384)
    double sqr_pore_r;    // width[2]^2*0.25         (src/energy.cpp:53)
    double T;             // ThermodynamicSystem::temp
    double sig2;          // sigma_ff^2
    double eps4;          // 4*epsilon_ff
    double rc2;           // rcut^2
    double zzV;           // zz*volume, zz = exp(mu/T)/Lambda^3 (src/moves.cpp:296-298)
    double ln_zzV, beta;  // log(zzV), 1/T: the Metropolis tests run in the log domain
    double wsum, thr_disp, thr_vol;  // move-type thresholds of src/main.cpp:67-70 as doubles
    double wall_pref;     // 4*pi*eps_sf*rho_s*sig_sf^2 (src/potentials.cpp:400)
    double sig_sf, delta;
    int geometry, wall_kind, n_layers;
    int pbc[3];
    int n_disp, n_vol, n_swap, n_equil;
    // "exotic" systems (row f4 of SURVEY.md §8: cylinder / sphere geometry, spherical wall, hard spheres, NPT volume moves) run through
    // the GENERAL kernels (the instantiations that also serve coordinates outside the box), never through the hot ones
    int pair_kind, exotic;
    double sig_ff, eps_sf, rho_s;
    double press_k;       // ExternalPressure / kB * 1e-30, K/AA^3 (src/moves.cpp:198)
    double volume;        // Box::volume (MC::ComputeVolume, src/read.cpp:26-35): the Gibbs transfer's acceptance and Widom sums
};

// Where a chain that outgrew its sub-warp slot (mcpm_subwarp.cu) stopped, and the counters of the interrupted set.
struct ChainResume {
    int paused, set, step;
    int n_disp, acc_disp, n_ins, acc_ins, n_del, acc_del, n_empty;
    unsigned long long evals, alg;
};

// Mutable chain state; `st` is exactly what mcpm_run reports.
struct ChainState {
    mcpm_chain_stats st;
    int fast_image;                 // all positions lie in [0,L] on periodic axes
    int energy_valid;               // running box energies were initialised by a full K3 evaluation
    unsigned long long replay_pos;  // cursor into the replay stream (REPLAY mode)
    unsigned long long cache_sum;   // checksum of the configuration the energy cache (RunArgs::ep) belongs to
    double b_energy, b_old;         // NPT only: the reference's RUNNING box energy and Box::oldEnergy (as shipped), see chain_loop
    int b_valid, pad3;
    int cache_valid;                // the cache was left consistent by the last K2 launch
    unsigned int stream_id;         // Philox stream of this chain: its GLOBAL id (default: index inside the context; mcpm_set_stream_ids)
    ChainResume res;
};

struct RunArgs {
    int first_set, n_sets;
    long long steps_per_set;
    unsigned long long seed;
    const double* replay_u;      // device pointer
    unsigned long long replay_stride;
    unsigned char* trace;        // device pointer or null
    unsigned long long trace_stride;
    int check_energy;
    int cap;
    int sample_rdf;
    int pool_threads;            // k2_pool: threads per chain (32), passed at run time so that the chain loop compiles like the one-CTA-per-chain kernel's
    int no_sampling;             // moves only: the sampling kernels skip ComputeWidom / ComputeRDF (mcpm_run_params::no_sampling)
    double* ep;                  // device, [n_chains][cap]: per-particle pair sums (unscaled) of the energy cache, or null
    double* rdf;                 // device, [n_chains][MCPM_RDF_BINS] accumulated histograms
};

// One request of K1 (mcpm_k1.cu): stored particle `index` (or -1) and/or a trial position.
struct K1Request {
    int index, has_trial;
    double t[3];
};
#define MCPM_K1_MAX_BATCH 1024   // requests per launch

// Resident K1 (k1_server in mcpm_k1.cu): a one-CTA persistent kernel that keeps ONE chain's positions staged in shared memory and is
// driven through a mailbox in mapped pinned host memory — no kernel launch per MC::EnergyOfParticle call.  The host fills the
// payload and then publishes `seq`; the kernel answers in K1Reply (records first, then `seq`).
enum { K1_CMD_ENERGY = 1, K1_CMD_APPLY = 2, K1_CMD_APPEND = 3, K1_CMD_REMOVE = 4, K1_CMD_STOP = 5, K1_CMD_BOX = 6 };
#define MCPM_K1_SERVER_BOX_MAX 1536   // MC::BoxEnergy is answered by the one-CTA service up to this many particles (beyond: K3 on the whole GPU)
struct __align__(64) K1Mail {   // ONE 64-byte line: the kernel's warp 0 fetches it with a single coalesced read per poll (one PCIe round trip)
    unsigned long long seq;      // written last by the host (release); the kernel waits for its next expected value
    unsigned long long what;     // cmd | has_trial << 8 | (unsigned)index << 32
    double t[3];
    unsigned long long pad[2];
    unsigned long long seq_tail; // == seq: both ends of the line must agree before the payload is believed
};
struct __align__(64) K1Reply {  // ONE 64-byte line written by ONE warp-wide store; each 32-byte sector carries its own copy of the sequence
    unsigned long long seq;      // number, so the host believes the numbers only when both copies have arrived
    double r[6];                 // {pair, wall, E} of the stored particle, then of the trial position (the many-body terms are 0)
    unsigned long long seq_tail;
    unsigned long long alive;    // (next line) 1 while the kernel serves; 0 once it has decided to exit (idle time-out or STOP) — it takes no request after that
};

// k2_pool (mcpm_kernels.cu): one persistent CTA per SM whose warps each own a shared-memory SLOT and take chains from per-class
// work queues.  Slots come in up to three capacity classes (largest first) so that small chains do not reserve the shared memory of
// the largest state point; a slot of class k serves the queues of classes >= k.
#define MCPM_POOL_MAX_SLOTS 24
#define MCPM_POOL_MAX_CLASSES 3
struct PoolLayout {
    int n_slots, n_classes;
    int slot_cap[MCPM_POOL_MAX_SLOTS];   // particles
    int slot_off[MCPM_POOL_MAX_SLOTS];   // offset of the slot in the CTA's dynamic shared memory, in doubles
    int slot_cls[MCPM_POOL_MAX_SLOTS];
    int q_begin[MCPM_POOL_MAX_CLASSES], q_end[MCPM_POOL_MAX_CLASSES];   // class k's chains: order[q_begin[k] .. q_end[k]), longest first
    int state_off;                       // offset of the per-warp ChainParams / ChainState copies, in doubles
    int state_stride;                    // doubles per warp
    int args_off;                        // offset of the CTA's copy of RunArgs, in doubles
    int* q_next;                         // device, [n_classes]: next unclaimed entry of every queue (zeroed before the launch)
};

#define MCPM_RDF_BINS 100        // NBINS, src/MC.h:11
#define MCPM_RDF_SMEM_DOUBLES (104 + 16 * 128 / 2)   // fp64 histogram + 16 warps x 128 unsigned counts (sampling kernels)
