/* mcpm.h — C ABI of the B200-native Monte Carlo adsorption engine (libmcpm.so).
 *
 * Drop-in boundary for ONE hot path of salflorom/MC-PorousMaterials: the per-trial-move energy
 * evaluation and the Metropolis translation / insertion / deletion loop that drives it.
 * The reference has no plugin or FFI layer; its seam is a pair of C++ member functions that
 * communicate through side effects on the monolithic `MC` object (reference src/MC.h:111-112).
 * Each entry point below names the reference interface it replaces (file:line under
 * /root/reference).  INTEGRATION.md shows the reference-side binding a maintainer would add.
 *
 * Conventions: plain C types only; opaque context handle; every call returns an int status
 * (MCPM_OK == 0) and never throws across the boundary; mcpm_last_error() gives the message of the
 * last failing call on the calling thread.  One context owns one GPU and one CUDA stream; calls on
 * one context are not thread-safe, different contexts may be used from different host threads.
 * Particle indices are 0-based and order-preserving (the reference's slots 1..n, src/MC.h:12).
 * Energies are in kelvin (E/kB), lengths in angstrom, exactly as in the reference.
 * All arithmetic is fp64.  There is NO CPU fallback: without a CUDA device every compute entry
 * point fails with MCPM_ERR_CUDA.
 */
#ifndef MCPM_H
#define MCPM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCPM_VERSION 200

enum {
    MCPM_OK = 0,
    MCPM_ERR_ARG = 1,       /* bad argument (null pointer, index out of range, unsupported option) */
    MCPM_ERR_CUDA = 2,      /* CUDA runtime error or no device */
    MCPM_ERR_CAPACITY = 3,  /* a chain outgrew its particle capacity */
    MCPM_ERR_UNSUPPORTED = 4,
    MCPM_ERR_IO = 5
};

/* Box::geometry, src/read.cpp:74; derived widths :128-135, PBC flags :146-162 (bulk xyz, slit xy, cylinder x, sphere none) */
enum { MCPM_GEOM_BULK = 0, MCPM_GEOM_SLIT = 1, MCPM_GEOM_CYLINDER = 2, MCPM_GEOM_SPHERE = 3 };
/* wall dispatch, src/energy.cpp:69-77: "lj" + slit -> SlitLJ_Pot, "lj" + sphere -> SphericalLJ_Pot (src/potentials.cpp:474-504),
 * "lj10-4" + cylinder -> CylindricalLJ10_4 (:412-444), "steele10-4-3" + cylinder -> CylindricalSteele10_4_3 (:445-469) */
enum { MCPM_WALL_NONE = 0, MCPM_WALL_SLIT_LJ = 1, MCPM_WALL_SPHERE_LJ = 2, MCPM_WALL_CYL_LJ104 = 3, MCPM_WALL_CYL_STEELE = 4 };
/* pair dispatch, src/energy.cpp:79-83: "lj" -> LJ126_Pot, "hs" -> HardSphere_Pot (src/potentials.cpp:40-54) */
enum { MCPM_PAIR_LJ = 0, MCPM_PAIR_HS = 1 };
enum { MCPM_RNG_PHILOX = 0, MCPM_RNG_REPLAY = 1 };

typedef struct mcpm_ctx mcpm_ctx; /* opaque */

/* One thermodynamic state point = one chain's system (single box, single species).
 * Field meaning = the reference's parsed input (src/read.cpp:50-121) after its post-processing
 * (src/read.cpp:122-168): width[0..1] are the derived lateral sizes, dr the translation amplitude. */
typedef struct mcpm_system_desc {
    int geometry;        /* MCPM_GEOM_*;  PBC flags follow the geometry as in src/read.cpp:146-162 */
    int wall_kind;       /* MCPM_WALL_*;  SLIT_LJ = SlitLJ_Pot, src/potentials.cpp:381-402 */
    int n_layers;        /* Box::nLayersPerWall */
    int n_disp, n_vol, n_swap; /* move weights, src/main.cpp:67-70; n_vol > 0 = NPT volume moves (MC::ChangeVolume, src/moves.cpp:191-231):
                                  bulk geometry and pressure > 0 only (the two-box Gibbs branches are out of scope) */
    int n_equil_sets;    /* sets during which dr adapts and nothing is sampled, src/moves.cpp:74, src/main.cpp:72 */
    int pair_kind;       /* MCPM_PAIR_*: fluid-fluid potential */
    double width[3];     /* Box::width, angstrom; width[2] is the pore size H */
    double temperature;  /* K */
    double eps_ff, sig_ff, rcut;    /* Fluid::{epsilon,sigma,rcut}(0), src/potentials.cpp:62-64 */
    double molar_mass, mu;          /* g/mol; full chemical potential mu/kB in K, src/moves.cpp:296-298 */
    double eps_sf, sig_sf, rho_s, delta; /* Box::fluid(0).{epsilon,sigma}(0), solidDens, deltaLayers, src/potentials.cpp:384-389 */
    double dr;           /* Simulation::dr, src/read.cpp:168 */
    double pressure;     /* ExternalPressure in Pa (src/read.cpp:64); <= 0: none */
    double dv;           /* Simulation::dv, amplitude of the log-volume step (src/MC.cpp:33); <= 0: the reference's 1e-3 */
} mcpm_system_desc;

/* Energies of one particle, as MC::EnergyOfParticle leaves them in the particle record
 * (src/energy.cpp:49-51,102-104): {pairPotE, manyBodyE (always 0 here), boxE, energy}. */
typedef struct mcpm_particle_energy_t {
    double pair, many_body, wall, energy;
} mcpm_particle_energy_t;

/* Box totals, as MC::BoxEnergy leaves them (src/energy.cpp:119-122): pair is ALREADY halved. */
typedef struct mcpm_box_energy_t {
    double pair, many_body, wall, energy;
} mcpm_box_energy_t;

typedef struct mcpm_run_params {
    int first_set;             /* 1-based index of the first set of this call (src/main.cpp:63) */
    int n_sets;
    long long steps_per_set;   /* Simulation::nStepsPerSet */
    int rng_mode;              /* MCPM_RNG_PHILOX: device Philox4x32-10 keyed by (seed, chain id, step);
                                  MCPM_RNG_REPLAY: consume replay_u exactly as the reference consumes
                                  MC::Random() (1 + 7/6/4/2 draws per step) */
    unsigned long long seed;
    const double* replay_u;    /* HOST pointer, n_chains * replay_stride uniforms (REPLAY only) */
    size_t replay_stride;
    unsigned char* trace;      /* HOST pointer or NULL: n_chains * n_sets*steps_per_set bytes,
                                  (move<<1)|accepted; move 0 translation, 1 insertion, 2 deletion,
                                  3 deletion branch on an empty box */
    int threads_per_chain;     /* CTA size, multiple of 32 up to 1024; 0 = choose from the loadings */
    int sample_rdf;            /* !=0: MC::ComputeRDF (src/sample.cpp:14-44) after every production step: O(N^2) 100-bin
                                  histogram out to rcut, accumulated per chain; read it with mcpm_get_rdf */
    int check_energy;          /* 1: recompute the full box energy on the device at the end (drift check, reported in
                                  *_check); 2: also adopt it as the running energy, as MC::BoxEnergy does */
    int no_sampling;           /* !=0: moves only — no ComputeWidom / ComputeRDF after the steps even where the reference's
                                  step loop would sample (src/main.cpp:72-75).  MC::MinimizeEnergy (src/energy.cpp:13-34) is
                                  such a loop: it calls MoveParticle and nothing else. */
} mcpm_run_params;

/* Per-chain result of mcpm_run.  Per-set counters follow MC::Stats after the LAST set of the call
 * (src/MC.h:35-44; ResetStats starts n_displacements and n_swaps at 1, src/MC.cpp:81-98). */
typedef struct mcpm_chain_stats {
    int n;                     /* current number of particles */
    int overflow;              /* !=0: the chain hit its capacity and stopped */
    double dr;                 /* current translation amplitude (after AdjustMCMoves) */
    double pair, wall, energy; /* running box energies by exact bookkeeping (pair halved) */
    double pair_check, wall_check, energy_check; /* full recompute at the end if check_energy */
    long long acceptance, rejection, n_displacements;
    long long accept_swap, reject_swap, n_swaps;
    long long tot_disp, tot_disp_acc, tot_ins, tot_ins_acc, tot_del, tot_del_acc, tot_empty;
    unsigned long long pair_evals;  /* partner-loop iterations (src/potentials.cpp:65-71) the engine EXECUTED, cache rebuilds included */
    unsigned long long pair_evals_algorithmic; /* what the reference's moves evaluate for the same steps: 2N per translation,
                                                  N per insertion / deletion / Widom step (src/moves.cpp:153,165,304,324,421,429) */
    unsigned long long steps_done;  /* chain's global step counter (= Philox step index) */
    /* production accumulators (set > n_equil_sets), one sample per step */
    double sum_n, sum_n2, sum_e, sum_e2;
    long long samples;
    unsigned long long replay_consumed; /* REPLAY mode: uniforms consumed by the last mcpm_run */
    /* Widom / BAR sampling of the LAST set (MC::ComputeWidom, src/moves.cpp:410-449; active like in the reference when
     * n_swap == 0, in production sets; ResetStats restarts the counts at ONE and the sums at 0 every set) */
    double widom_insert, widom_delete;
    long long widom_insertions, widom_deletions;
    /* NPT (n_vol > 0): MC::Stats volume counters of the last set (nVolChanges starts at ONE), box after the run */
    long long accept_vol, reject_vol, n_vol_changes;
    double volume, width;      /* Box::volume, Box::width(2) (== width(0) == width(1) after a volume move, src/moves.cpp:217-218) */
    double dv;                 /* Simulation::dv after AdjustMCMoves (src/moves.cpp:79-84) */
    double energy_as_shipped;  /* the reference's RUNNING box energy (half pair updates, CorrectEnergy's ratio test: quirks Q3, Q4, Q16), which
                                  MC::ChangeVolume uses as the old energy (src/moves.cpp:220); tracked only when n_vol > 0 */
    double dr_used;            /* translation amplitude in effect DURING the last set (the reference prints its statistics
                                  before AdjustMCMoves changes it, src/main.cpp:77-83); `dr` above is the value after it */
} mcpm_chain_stats;

/* ---- library ---------------------------------------------------------------------------- */
int mcpm_version(void);
const char* mcpm_last_error(void);
int mcpm_device_count(int* count);

/* ---- context: n_chains independent systems, each with room for `capacity` particles ------- */
/* Replaces the MC constructor's fixed-capacity arrays (src/MC.cpp:13-80, src/MC.h:68,79). */
int mcpm_create(int device, int n_chains, int capacity, mcpm_ctx** out);
int mcpm_destroy(mcpm_ctx* ctx);
int mcpm_n_chains(const mcpm_ctx* ctx, int* n_chains, int* capacity);

/* chain = -1 applies the description to every chain.  Replaces ReadInputFile's state (src/read.cpp:36-172). */
int mcpm_set_system(mcpm_ctx* ctx, int chain, const mcpm_system_desc* desc);
int mcpm_get_system(const mcpm_ctx* ctx, int chain, mcpm_system_desc* desc);

/* SoA fp64 positions of one chain (host pointers). Replaces box(b).fluid(s).particle(1..n).{x,y,z}. */
int mcpm_set_positions(mcpm_ctx* ctx, int chain, int n, const double* x, const double* y, const double* z);
int mcpm_get_positions(mcpm_ctx* ctx, int chain, int* n, double* x, double* y, double* z);
/* All chains at once: n[n_chains]; x,y,z are [n_chains][capacity] row-major host arrays. */
int mcpm_set_positions_all(mcpm_ctx* ctx, const int* n, const double* x, const double* y, const double* z);
int mcpm_get_positions_all(mcpm_ctx* ctx, int* n, double* x, double* y, double* z);

/* ---- K1: single-particle energy == MC::EnergyOfParticle (src/energy.cpp:43-105) ------------ *
 * index >= 0, trial == NULL : energy of stored particle `index`                 -> cur
 * index >= 0, trial != NULL : also the energy of that particle displaced to trial[3]
 *                             (what MoveParticle evaluates, src/moves.cpp:153,165) -> cur and moved
 * index == -1, trial != NULL: energy of a ghost / inserted particle at trial[3]
 *                             (src/moves.cpp:304, :421)                          -> moved
 * Either output pointer may be NULL. One launch evaluates both positions in one pass over j; the results come back through
 * mapped host memory (no device->host copy, no stream synchronisation per call). */
int mcpm_particle_energy(mcpm_ctx* ctx, int chain, int index, const double* trial,
                         mcpm_particle_energy_t* cur, mcpm_particle_energy_t* moved);
/* K1 as a RESIDENT service for one chain: after mcpm_k1_resident(ctx, chain, 1), mcpm_particle_energy / mcpm_apply_move / mcpm_append /
 * mcpm_remove_swap_last on that chain no longer launch anything — a persistent one-CTA kernel that keeps the chain's positions staged in
 * shared memory answers through a mailbox in mapped host memory (a few microseconds per MC::EnergyOfParticle call, below the reference's
 * own ~7 us at N = 500).  The kernel starts with the first request, leaves after ~5 ms without one and whenever any other entry point
 * touches the context (it is restarted on demand), and writes accepted moves through to global memory, so results never depend on it.
 * This is the mode a host-driven move loop runs in.  enable = 0 disarms. */
int mcpm_k1_resident(mcpm_ctx* ctx, int chain, int enable);
int mcpm_k1_resident_stats(const mcpm_ctx* ctx, unsigned long long* starts, unsigned long long* requests);
/* Measurement support: wall-clock microseconds per mcpm_particle_energy call for a native caller (`reps` calls, stored particle r mod n
 * displaced to `trial`), in whatever mode the chain is in. */
int mcpm_k1_latency(mcpm_ctx* ctx, int chain, int reps, const double* trial /* 3 or NULL */, double* us_per_call);
/* The same for `count` requests against the chain's PRESENT configuration in one launch (a host loop that speculates on several
 * trial moves, Widom ghost batches, ...): index[count]; trial = NULL or [count][3]; cur / moved = NULL or [count].  The launch
 * latency that bounds a single call is paid once per batch. */
int mcpm_particle_energy_batch(mcpm_ctx* ctx, int chain, int count, const int* index, const double* trial,
                               mcpm_particle_energy_t* cur, mcpm_particle_energy_t* moved);

/* ---- K3: full box energy == MC::BoxEnergy (src/energy.cpp:106-124) ------------------------- *
 * per_pair / per_wall (nullable, n doubles each) receive every particle's pairPotE / boxE, as
 * BoxEnergy leaves them in the particle records.  Also resets the chain's running energies. */
int mcpm_box_energy(mcpm_ctx* ctx, int chain, mcpm_box_energy_t* out, double* per_pair, double* per_wall);
int mcpm_box_energy_all(mcpm_ctx* ctx, mcpm_box_energy_t* out /* [n_chains] */);
/* K3 (chains of >= 8192 particles) and K2 / mcpm_run (largest chain >= 4096 particles, threads_per_chain 0, no in-kernel RDF)
 * switch to a cell list — 27 neighbour cells instead of all N partners per energy evaluation — for bulk-like chains whose box
 * holds >= 3 cut-off lengths per periodic axis and whose coordinates lie inside the box (config C4).  Same sums up to the
 * order of the additions.  A K2 chain whose local density more than triples during a run stops with MCPM_ERR_CAPACITY (a
 * cell ran out of slots).  enable = 0 forces the all-pairs kernels (default: 1). */
int mcpm_use_cell_list(mcpm_ctx* ctx, int enable);

/* ---- state mutation for a host-driven loop (src/moves.cpp:163,184 / :302-311 / :329-332) --- */
int mcpm_apply_move(mcpm_ctx* ctx, int chain, int index, const double xyz[3]);
int mcpm_append(mcpm_ctx* ctx, int chain, const double xyz[3]);
int mcpm_remove_swap_last(mcpm_ctx* ctx, int chain, int index);

/* ---- K2: batched persistent chains == the step loop src/main.cpp:63-84 with MoveParticle
 *      (src/moves.cpp:118-186), SwapParticle single-box branch (:284-341), AdjustMCMoves (:68-78)
 *      and ResetStats (src/MC.cpp:81-98), one CTA per chain, no host round trip per move. ------- */
int mcpm_run(mcpm_ctx* ctx, const mcpm_run_params* params, mcpm_chain_stats* stats /* [n_chains], nullable */);
int mcpm_get_stats(mcpm_ctx* ctx, mcpm_chain_stats* stats /* [n_chains] */);
/* K2 keeps a per-particle pair-energy cache (default on): the old energy of a translated or deleted particle is read from
 * it instead of being re-summed, accepted moves update every partner's entry, the cache is rebuilt at each launch.
 * Same Markov chain (decisions can differ only when a uniform falls within ~1e-13 of its threshold); ~0.95 N instead of
 * 1.5 N pair evaluations per step at 20 % translation acceptance.  enable = 0 re-sums both energies every move, exactly
 * like MC::MoveParticle / MC::SwapParticle. */
int mcpm_use_energy_cache(mcpm_ctx* ctx, int enable);
/* The cache itself, for drift checks: pair[j] = the pairPotE (K) the engine currently carries for particle j of `chain`
 * (what MC::EnergyOfParticle would store, src/energy.cpp:102-104); *valid = 0 if no K2 launch has left a consistent cache
 * for the chain's present configuration (then pair[] is not written).  Compare with mcpm_box_energy's per_pair. */
int mcpm_get_energy_cache(mcpm_ctx* ctx, int chain, double* pair /* n */, int* valid);
/* Speculative moves (the latency kernel for FEW chains, e.g. one chain per isotherm point): a CTA of 8 warps evaluates
 * the next 8 steps of its chain concurrently against the same configuration, commits the first accepted one and
 * re-evaluates the steps after it; rejected steps before it stand.  The chain is exactly the sequential one (same
 * uniforms per step index, same decisions, same trace).  mode -1 (default): chosen automatically when the context has
 * no more chains than the device has SMs, none larger than 2048 particles, threads_per_chain is 0 and the energy cache is in use; 0: never; 1: whenever
 * the launch qualifies (energy cache enabled, coordinates inside the box, no in-kernel sampling, shared-memory
 * positions); 2 / 4: the LEAN variants whenever the launch qualifies — two / four candidate warps and nothing else per chain,
 * energy cache in global memory, 8 / 4 chains per SM.  Chosen automatically (mode -1) when more chains than SMs but at most
 * 8 (two warps) or 4 (four warps) share an SM (+12 % on config C5); with more chains per SM they cost residency and are 30 %
 * slower than one warp per chain (config C2).
 * Replaces nothing in the reference: it is a schedule of src/main.cpp:66-71. */
int mcpm_use_speculation(mcpm_ctx* ctx, int mode);
int mcpm_reset_accumulators(mcpm_ctx* ctx);
/* Accumulated pair-distance histogram of one chain: hist[0..99] = MC::Stats::rdf(0..99) (bin 0 is never filled,
 * each pair adds 2, src/sample.cpp:27-28).  mcpm_reset_accumulators clears it. */
int mcpm_get_rdf(mcpm_ctx* ctx, int chain, double* hist /* 100 */);
/* mu_ex and mu (K) from the Widom/BAR sums of the last set == MC::ComputeChemicalPotential, src/moves.cpp:453-465. */
int mcpm_chemical_potential(mcpm_ctx* ctx, int chain, double* mu_ex, double* mu);
/* Hand the running box energies (pair already halved, wall) of every chain back to the engine after an upload of
 * positions whose energies the caller already knows (e.g. from the previous mcpm_run's statistics): skips the
 * full K3 evaluation that mcpm_run would otherwise do to initialise its exact bookkeeping. */
int mcpm_set_running_energy(mcpm_ctx* ctx, const double* pair /* [n_chains] */, const double* wall /* [n_chains] */);
/* Philox stream of every chain: the device stream is keyed by (seed, stream id, step).  Default: the chain's index inside
 * the context.  A caller that shards one set of chains over several contexts / GPUs passes each chain's GLOBAL id here,
 * so that no two chains of the job share a stream whatever the sharding (and results do not depend on it). */
int mcpm_set_stream_ids(mcpm_ctx* ctx, const unsigned int* ids /* [n_chains] */);
/* Two-box Gibbs ensemble (thermoSys.nBoxes == 2).  Chains `box0` and `box1` — both with their systems set: same fluid, temperature,
 * move weights and n_equil_sets, n_vol == 0, any geometry / wall each (e.g. a slit pore and a bulk reservoir) — become box 0 and box 1
 * of ONE system, advanced together by mcpm_run:
 *   n_disp  MC::MoveParticle with its box selection (src/moves.cpp:127-137: rand = int(u*N_total), the first box with rand <= cumulative count)
 *   n_swap  MC::SwapParticle's Gibbs branch (src/moves.cpp:342-400): a particle transfer between the boxes, accepted with
 *           exp(-(E_in - E_out)/T) * V_in N_out / (V_out (N_in + 1)); the Widom/BAR sums of both boxes are accumulated on the way
 *           (mcpm_chemical_potential then gives each box's mu as MC::ComputeChemicalPotential does)
 * The system draws from ONE uniform stream — box 0's (stream id, replay row); its trace is box 0's row:
 *   translation: accepted | box<<4;  transfer attempt: (5<<1) | accepted | inBox<<4;  transfer with an empty donor box: (6<<1) | inBox<<4.
 * Per-box statistics come back in each chain's mcpm_chain_stats (accept_swap / reject_swap / n_swaps are the system's, tot_ins / tot_del the
 * transfers into / out of the box).  A context holds either single-box chains or linked pairs, not both.  The Gibbs branches of
 * MC::ChangeVolume (src/moves.cpp:232-283) are out of scope — the stock reference aborts on entering ChangeVolume — MCPM_ERR_UNSUPPORTED. */
int mcpm_link_boxes(mcpm_ctx* ctx, int box0, int box1);
/* Override a chain's translation amplitude / global step counter (restart, tests). */
int mcpm_set_dr(mcpm_ctx* ctx, int chain, double dr);
int mcpm_set_step_counter(mcpm_ctx* ctx, int chain, unsigned long long step);

/* ---- measurement support ------------------------------------------------------------------ */
/* Device time (CUDA events on the context's stream) and launch count of the last compute call. */
int mcpm_last_kernel_ms(const mcpm_ctx* ctx, float* ms);
/* Which K2 kernel the last mcpm_run launched: "k2_chains", "k2_spec", "k2_pair", "k2_quad" or "k2_cells" ("" before the first run). */
const char* mcpm_last_kernel_name(const mcpm_ctx* ctx);
/* Position bytes the bulk transfers (mcpm_set_positions_all / mcpm_get_positions_all) have moved so far, host -> device and back:
 * chains are copied in runs of neighbours with similar loadings, each run as wide as its largest chain. */
int mcpm_transfer_bytes(const mcpm_ctx* ctx, unsigned long long* h2d, unsigned long long* d2h);
int mcpm_launch_count(const mcpm_ctx* ctx, unsigned long long* launches);
/* CUDA-event stopwatch on the context's stream: start records an event, stop records a second one,
 * waits for it and returns the device time between the two in milliseconds. */
int mcpm_timer_start(mcpm_ctx* ctx);
int mcpm_timer_stop(mcpm_ctx* ctx, float* ms);
/* The device Metropolis tests on caller-supplied inputs, for tests: kind 0 translation u < exp(-x) with x = dE/T
 * (src/moves.cpp:170); 1 insertion u < zzV exp(x)/(n+1) with x = -E/T (src/moves.cpp:306); 2 deletion
 * u < n exp(x)/zzV with x = E/T (src/moves.cpp:326).  out[k] = 1 if accepted.  The device evaluates them with bounds
 * and a single-precision screen first; the outcome must always be that of the fp64 expression. */
int mcpm_metropolis_probe(mcpm_ctx* ctx, int kind, int count, const double* u, const double* x, const int* n, double zzV, int* out);
/* FP64 FMA microbenchmark on `device`: the roofline denominator for these kernels (TFLOP/s). */
int mcpm_fp64_peak(int device, double* tflops, double* ms);
/* The device Philox stream, for tests: uniform of (seed, chain, step, slot). */
int mcpm_philox_uniforms(mcpm_ctx* ctx, unsigned long long seed, unsigned int chain,
                         unsigned long long first_step, int n_steps, double* out /* n_steps*8 */);

/* ---- host driver (C++ in csrc/host/, re-hosting the reference's main/read/print) ------------- *
 * Parsed input file == the state MC::ReadInputFile leaves behind (src/read.cpp:36-172), single species, one box
 * or two (the two-box Gibbs ensemble at fixed volumes: `system2`, see mcpm_link_boxes).  Counts are doubles because the reference parses them with stod (src/read.cpp:56-59). */
typedef struct mcpm_sim_config {
    char project[64], fluid[64], box[64];       /* lower-cased like the reference (src/tools.h:19-30) */
    double n_sets, n_equil_sets, steps_per_set, print_every;
    double pressure;                            /* ExternalPressure, unused (NPT out of scope), -1 if absent */
    int n_particles;                            /* <Box> <Fluid> NumberOfParticles */
    int sample_rdf, print_trajectory, continue_after_crash;
    int n_boxes, n_species;                     /* what the file declared (1 or 2 boxes, 1 species) */
    /* extensions of this engine (ignored by the reference, which skips unknown lines) */
    int has_seed; unsigned long long seed;      /* Seed <n>: reproducible runs (quirk Q2) */
    int replicas;                               /* Replicas <n>: independent replica chains per state point */
    int sweep_points; double sweep_mu0, sweep_mu1; /* ChemicalPotentialSweep <mu0> <mu1> <n>: isotherm */
    int capacity;                               /* Capacity <n>: particle slots per chain (default: from N) */
    int devices;                                /* Devices <n>: GPUs to shard the chains over */
    mcpm_system_desc system;                    /* derived widths, PBC, dr as in src/read.cpp:122-168 */
    /* a second BoxName block: box 1 of a two-box (Gibbs ensemble) system — same fluid, its own geometry / size / wall */
    char box2[64];
    int n_particles2;
    int fix_volume[2];                          /* FixVolume of box 0 / box 1 (echoed by PrintParams) */
    mcpm_system_desc system2;
} mcpm_sim_config;

/* MC::ReadInputFile (src/read.cpp:36-172).  No GPU needed. */
int mcpm_read_input_file(const char* path, mcpm_sim_config* cfg);
/* The whole program src/main.cpp:16-97 on the GPU: echo (PrintParams), output tree (OutputDirectory),
 * initial configuration, "minimisation", set loop with PrintStats / PrintTrajectory / PrintLog in the
 * reference's formats.  flags: bit 0 = reproduce the log file's double-halved ffE column (quirk Q5),
 * bit 1 = quiet.  out_dir = NULL writes under the current directory like the reference. */
int mcpm_simulate(const char* input_path, const char* out_dir, int flags);

#ifdef __cplusplus
}
#endif
#endif /* MCPM_H */
