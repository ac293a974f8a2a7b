/* TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.
 *
 * CPU oracle: a plain-C restatement of the reference's hot path (salflorom/MC-PorousMaterials,
 * single box, single species).  Every function cites the reference file:line it follows.
 *
 * Pinning: the reference ships NO tests, golden vectors or fixtures ("parity unpinned" by the
 * reference's own tests).  This oracle is instead pinned against the UNMODIFIED
 * reference compiled from /root/reference (oracle/_ref/libmcref.so, built by oracle/Makefile) —
 * live in tests/test_oracle_vs_reference.py when that library is present, and through the golden
 * vectors under tests/golden/ (made by tests/golden/make_golden.py from the compiled reference).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product library (libmcpm.so) never does.
 */
#ifndef MCPM_ORACLE_H
#define MCPM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same field order as oracle/ref_harness.cpp::mcref_system. */
typedef struct {
    int geometry;   /* 0 = bulk (PBC x,y,z), 1 = slit (PBC x,y; walls at z=0 and z=H), 2 = cylinder (PBC x), 3 = sphere (no PBC)
                       read.cpp:128-162 */
    int wall_kind;  /* 0 = none, 1 = "lj" + slit -> SlitLJ_Pot, 2 = "lj" + sphere -> SphericalLJ_Pot, 3 = "lj10-4" + cylinder ->
                       CylindricalLJ10_4, 4 = "steele10-4-3" + cylinder -> CylindricalSteele10_4_3              energy.cpp:69-77 */
    int n_layers;   /* nLayersPerWall */
    int n_disp, n_vol, n_swap; /* move weights, main.cpp:67-70 */
    int n_equil_sets;
    int pair_kind;  /* 0 = "lj" -> LJ126_Pot, 1 = "hs" -> HardSphere_Pot                 energy.cpp:80-83 */
    double width[3];
    double temperature;
    double eps_ff, sig_ff, rcut;
    double molar_mass, mu;
    double eps_sf, sig_sf, rho_s, delta;
    double dr;
    double pressure; /* ExternalPressure in Pa (<= 0: none), read.cpp:64; NPT volume moves moves.cpp:191-231 */
    double dv;       /* Simulation::dv, log-volume step (MC.cpp:33: 1e-3) */
} orc_system;

typedef struct orc_chain orc_chain;

orc_chain* orc_create(const orc_system* sys, int capacity);
void orc_destroy(orc_chain* c);
void orc_set_positions(orc_chain* c, int n, const double* x, const double* y, const double* z);
int orc_get_positions(const orc_chain* c, double* x, double* y, double* z);

/* out = {pairPotE, manyBodyE(=0), boxE, energy}                              energy.cpp:43-105 */
void orc_particle_energy(orc_chain* c, int index0, double out[4]);
void orc_trial_energy(orc_chain* c, double x, double y, double z, double out[4]);
void orc_moved_energy(orc_chain* c, int index0, double x, double y, double z, double out[4]);
/* out = {pairPotE (already halved), manyBodyE, boxE, energy}                 energy.cpp:106-124 */
void orc_box_energy(orc_chain* c, double out[4], double* per_pair, double* per_wall);
double orc_slit_wall(const orc_system* s, double z);                       /* potentials.cpp:381-402 */
double orc_volume(const orc_system* s);                                    /* read.cpp:26-35 */
double orc_swap_zz(const orc_system* s);                                   /* moves.cpp:296-298 */

/* Uniform sources.  MC::Random() is uniform_real_distribution{0,0.9999}(mt19937_64), MC.h:28-33,123. */
void orc_rng_mt(orc_chain* c, uint64_t seed);                    /* own mt19937_64 + the Q1 mapping */
void orc_rng_replay(orc_chain* c, const double* u, size_t n);    /* caller-owned uniform stream */
void orc_rng_philox(orc_chain* c, uint64_t seed, uint32_t chain_id, uint64_t first_step);
double orc_random(orc_chain* c);
size_t orc_draws(const orc_chain* c);
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* The product's device stream: uniform for (seed, chain, step, slot), already scaled by 0.9999. */
double orc_philox_uniform(uint64_t seed, uint32_t chain_id, uint64_t step, int slot);
void orc_mt_fill(uint64_t seed, double* u, size_t n);            /* n reference-mapped uniforms */

void orc_initial_config(orc_chain* c, int n);                              /* moves.cpp:56-67 (slit/bulk) */
/* One pass of the step-loop body, main.cpp:67-71.  Returns (move<<1)|accepted with
 * move 0 = translation, 1 = insertion attempt, 2 = deletion attempt, 3 = deletion branch on empty box. */
int orc_step(orc_chain* c, int correct_energy);
long long orc_run(orc_chain* c, long long nsteps, int correct_energy, unsigned char* trace);
void orc_set_dr(orc_chain* c, double dr);                                  /* Simulation::dr, restart / tests */
/* dout = {volume, width[0], width[1], width[2], dv, as-shipped running energy}; iout = {acceptanceVol, rejectionVol, nVolChanges} */
void orc_get_volume_state(const orc_chain* c, double dout[6], long long iout[3]);
double orc_hyp2f1(int which, double z);                                  /* 2F1 closed forms of the cylindrical walls (test hook) */
double orc_wall_energy(orc_chain* c, double x, double y, double z);        /* boxE of EnergyOfParticle at a position */
void orc_reset_stats(orc_chain* c);                                        /* MC.cpp:81-98 */
void orc_adjust(orc_chain* c, int set);                                    /* moves.cpp:68-78 */
/* Sets first_set..first_set+n_sets-1: ResetStats; steps; sampling when set > n_equil_sets; adjust.
 * main.cpp:63-84.  Accumulates the running sums below (the product's K4 accumulators). */
void orc_run_sets(orc_chain* c, int first_set, int n_sets, long long steps_per_set, int correct_energy,
                  unsigned char* trace);

/* Two boxes (Gibbs ensemble, thermoSys.nBoxes == 2): `b` becomes box 1 of `a`'s system and draws from a's uniform source; the
 * global swap counters live in a.  moves.cpp:127-137 (box selection of MoveParticle), 342-400 (SwapParticle's Gibbs branch).
 * Codes: translation | box<<4; (5<<1)|acc | inBox<<4 transfer attempt; (6<<1) | inBox<<4 donor box empty.  n_vol must be 0. */
void orc_link_boxes(orc_chain* a, orc_chain* b);
int orc_gibbs_step(orc_chain* a, int correct_energy);
void orc_gibbs_run_sets(orc_chain* a, int first_set, int n_sets, long long steps_per_set, int correct_energy, unsigned char* trace);

/* dout = {pairPotE, manyBodyE, boxE, energy, dr, volume, oldEnergy, temperature,
 *         exact pair (halved), exact wall, exact energy, sumN, sumN2, sumE, sumE2, samples}
 * iout = {n, acceptance, rejection, nDisplacements, acceptSwap, rejectSwap, nSwaps, n,
 *         total translation attempts, total translation accepted, total swap attempts,
 *         total swap accepted, full recomputes, pair evals lo, pair evals hi, 0} */
void orc_get_state(const orc_chain* c, double dout[16], long long iout[16]);

/* Widom / BAR sampling (moves.cpp:410-449) and RDF (sample.cpp:14-44) — "next" rows of SURVEY §8f. */
void orc_widom(orc_chain* c, double out[2], long long iout[2]);
void orc_chem_pot(orc_chain* c, double out[2]);                            /* moves.cpp:453-465 */
void orc_rdf(orc_chain* c, double* hist100);
void orc_rdf_reset(orc_chain* c);
void orc_set_sample_rdf(orc_chain* c, int on);   /* orc_run_sets then calls ComputeRDF every production step */
void orc_get_rdf(const orc_chain* c, double* hist100);
void orc_get_widom(const orc_chain* c, double out[2], long long iout[2]);

#ifdef __cplusplus
}
#endif
#endif
