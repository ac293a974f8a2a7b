/* TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.  See mcpm_oracle.h for the pinning statement.
 *
 * Plain-C restatement (scalar, single-threaded, SoA storage) of the reference's per-trial energy
 * evaluation and of the Metropolis moves that call it.  It follows the reference's arithmetic
 * (sqrt distance, repeated-multiply powers, round() minimum image, position-inequality
 * self-exclusion) so that it can be pinned against the compiled reference to rounding level.
 */
#include "mcpm_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
/* MC.h:16-18 */
#define ORC_KB 1.380649e-23
#define ORC_NA 6.02214076e23
#define ORC_PLANCK 6.62607015e-34
#define ORC_NBINS 100 /* MC.h:11 */

enum { RNG_MT = 0, RNG_REPLAY = 1, RNG_PHILOX = 2 };

struct orc_chain {
    orc_system sys;
    int pbc[3];
    double volume;
    int cap, n;
    double *x, *y, *z;
    /* energies stored in the particle records by EnergyOfParticle (MC.h:53-55) */
    double *pe_pair, *pe_wall, *pe_tot;
    double sx, sy, sz, s_pair, s_wall, s_tot; /* slot 0: saved old particle, moves.cpp:152 */
    /* box running energies, as shipped (Box::energy etc., MC.h:75-76) */
    double b_pair, b_many, b_wall, b_energy, b_old;
    /* exact bookkeeping (Q3/Q4 fixed) — what the product keeps */
    double e_pair, e_wall, e_energy;
    double dr;
    /* per-set stats (MC.h:35-44, MC.cpp:81-98) */
    long long acceptance, rejection, n_displacements, accept_swap, reject_swap, n_swaps;
    long long acceptance_vol, rejection_vol, n_vol_changes;   /* MC.h:37-38 */
    double dv;                                                /* Simulation::dv */
    /* run totals */
    long long tot_disp, tot_disp_acc, tot_swap, tot_swap_acc, recomputes;
    unsigned long long pair_evals;
    double sum_n, sum_n2, sum_e, sum_e2;
    long long samples;
    /* Widom (MC.h:41-42) */
    long long widom_insertions, widom_deletions;
    double widom_insert, widom_delete, bar_c, mu_ex, mu_tot;
    double rdf[ORC_NBINS + 1];
    /* uniform source */
    int rng_mode;
    uint64_t mt[312];
    int mti;
    const double* replay;
    size_t replay_n, draws;
    uint64_t ph_seed, ph_step;
    uint32_t ph_chain;
    int ph_slot;
    int sample_rdf;
    /* two-box Gibbs system (thermoSys.nBoxes == 2): box 1 draws from box 0's uniform source; the global swap counters live in box 0 */
    orc_chain* rng_src;
    orc_chain* peer;
    int box_id;
};

/* ---------------------------------------------------------------- uniform sources ---------- */

static void mt_seed(orc_chain* c, uint64_t seed) {
    c->mt[0] = seed;
    for (int i = 1; i < 312; i++) c->mt[i] = 6364136223846793005ULL * (c->mt[i - 1] ^ (c->mt[i - 1] >> 62)) + (uint64_t)i;
    c->mti = 312;
}
static uint64_t mt_next(orc_chain* c) {
    if (c->mti >= 312) {
        for (int i = 0; i < 312; i++) {
            uint64_t y = (c->mt[i] & 0xFFFFFFFF80000000ULL) | (c->mt[(i + 1) % 312] & 0x7FFFFFFFULL);
            c->mt[i] = c->mt[(i + 156) % 312] ^ (y >> 1) ^ ((y & 1ULL) ? 0xB5026F5AA96619E9ULL : 0ULL);
        }
        c->mti = 0;
    }
    uint64_t x = c->mt[c->mti++];
    x ^= (x >> 29) & 0x5555555555555555ULL;
    x ^= (x << 17) & 0x71D67FFFEDA60000ULL;
    x ^= (x << 37) & 0xFFF7EEE000000000ULL;
    x ^= (x >> 43);
    return x;
}
/* MC::Random(), MC.h:33,123: libstdc++ generate_canonical<double,53> over one 64-bit draw, then
 * (b-a)*c + a with a=0, b=0.9999 (quirk Q1). */
static double q1_map(uint64_t v) {
    double cnl = (double)v * 5.42101086242752217e-20; /* 2^-64 */
    if (cnl >= 1.0) cnl = 1.0 - 1.1102230246251565e-16; /* nextafter(1,0) */
    return 0.9999 * cnl + 0.0;
}

void orc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* The product's device stream (csrc/mcpm_rng.cuh): call = 4*step + slot/2;
 * ctr = (call_lo, call_hi, chain, 'MCPM'), key = (seed_lo, seed_hi); 53-bit mantissas; x0.9999. */
double orc_philox_uniform(uint64_t seed, uint32_t chain_id, uint64_t step, int slot) {
    /* slots 0..7: the move's stream (tag 'MCPM'); slots 8..15: the sampling stream of the same step (tag 'MCPN'),
     * used by the Widom draws that follow the move in production steps */
    uint32_t tag = 0x4D43504Du;
    if (slot >= 8) { slot -= 8; tag = 0x4D43504Eu; }
    uint64_t call = 4ULL * step + (uint64_t)(slot >> 1);
    uint32_t ctr[4] = {(uint32_t)call, (uint32_t)(call >> 32), chain_id, tag};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t r[4];
    orc_philox4x32_10(ctr, key, r);
    uint64_t w = (slot & 1) ? (((uint64_t)r[3] << 32) | r[2]) : (((uint64_t)r[1] << 32) | r[0]);
    return 0.9999 * ((double)(w >> 11) * 1.1102230246251565e-16);
}

void orc_rng_mt(orc_chain* c, uint64_t seed) { c->rng_mode = RNG_MT; mt_seed(c, seed); c->draws = 0; }
void orc_rng_replay(orc_chain* c, const double* u, size_t n) {
    c->rng_mode = RNG_REPLAY; c->replay = u; c->replay_n = n; c->draws = 0;
}
void orc_rng_philox(orc_chain* c, uint64_t seed, uint32_t chain_id, uint64_t first_step) {
    c->rng_mode = RNG_PHILOX; c->ph_seed = seed; c->ph_chain = chain_id; c->ph_step = first_step;
    c->ph_slot = 0; c->draws = 0;
}
double orc_random(orc_chain* c) {
    if (c->rng_src) c = c->rng_src;
    c->draws++;
    if (c->rng_mode == RNG_MT) return q1_map(mt_next(c));
    if (c->rng_mode == RNG_REPLAY) return (c->draws <= c->replay_n) ? c->replay[c->draws - 1] : 0.0;
    return orc_philox_uniform(c->ph_seed, c->ph_chain, c->ph_step, c->ph_slot++);
}
size_t orc_draws(const orc_chain* c) { return c->draws; }
void orc_mt_fill(uint64_t seed, double* u, size_t n) {
    orc_chain* t = (orc_chain*)calloc(1, sizeof(orc_chain));
    mt_seed(t, seed);
    for (size_t i = 0; i < n; i++) u[i] = q1_map(mt_next(t));
    free(t);
}

/* ---------------------------------------------------------------- potentials --------------- */

/* Tools::Pow, tools.h:53-58 */
static double ipow(double v, int p) {
    double r = 1.;
    for (int i = 1; i <= p; i++) r *= v;
    return r;
}
/* MC::ComputeVolume, read.cpp:26-35 */
double orc_volume(const orc_system* s) {
    if (s->geometry == 3) return (M_PI / 6.) * ipow(s->width[2], 3);
    if (s->geometry == 2) return (M_PI / 4.) * ipow(s->width[2], 2) * s->width[0];
    if (s->geometry == 1) return s->width[0] * s->width[1] * s->width[2];
    return ipow(s->width[2], 3);
}
/* MC::ComputeBoxWidth, read.cpp:19-25 */
static double box_width(const orc_system* s, double volume) {
    if (s->geometry == 3) return 2 * cbrt(3 * volume / (4 * M_PI));
    if (s->geometry == 2) return 2 * sqrt(volume / (M_PI * s->width[0]));
    if (s->geometry == 1) return volume / (s->width[0] * s->width[1]);
    return cbrt(volume);
}
/* SwapParticle's activity, moves.cpp:296-298 */
double orc_swap_zz(const orc_system* s) {
    double particleMass = s->molar_mass / ORC_NA * 1e-3;
    double thermalWL = ORC_PLANCK / sqrt(2.0 * M_PI * particleMass * ORC_KB * s->temperature) * 1e10;
    return exp(s->mu / s->temperature) / ipow(thermalWL, 3);
}
/* MC::NeighDistance, potentials.cpp:22-33 */
static double neigh_distance(const orc_chain* c, double xi, double yi, double zi, double xj, double yj, double zj) {
    double dx = xi - xj, dy = yi - yj, dz = zi - zj;
    if (c->pbc[0]) dx -= round(dx / c->sys.width[0]) * c->sys.width[0];
    if (c->pbc[1]) dy -= round(dy / c->sys.width[1]) * c->sys.width[1];
    if (c->pbc[2]) dz -= round(dz / c->sys.width[2]) * c->sys.width[2];
    return sqrt(dx * dx + dy * dy + dz * dz);
}
/* MC::LJ126_Pot, potentials.cpp:55-73 — partner scan over the live particles 0..n-1, self skipped
 * by POSITION inequality (MC.h:56-60, quirk Q7), inclusive cutoff (Q8), no shift, no tail. */
static double lj126(orc_chain* c, double xi, double yi, double zi) {
    double eps = c->sys.eps_ff, sig = c->sys.sig_ff, rcut = c->sys.rcut, energy = 0.;
    for (int j = 0; j < c->n; j++) {
        if (xi != c->x[j] || yi != c->y[j] || zi != c->z[j]) {
            double rij = neigh_distance(c, xi, yi, zi, c->x[j], c->y[j], c->z[j]);
            if (rij <= rcut) energy += 4. * eps * (ipow(sig / rij, 12) - ipow(sig / rij, 6));
        }
    }
    c->pair_evals += (unsigned long long)c->n;
    return energy;
}
/* MC::HardSphere_Pot, potentials.cpp:40-54: the finite INT_MAX per partner within sigma; NO self-exclusion — a stored particle
 * finds its own record at distance 0 and counts itself (quirk Q17). */
static double hard_sphere(orc_chain* c, double xi, double yi, double zi) {
    double sig = c->sys.sig_ff, energy = 0.;
    for (int j = 0; j < c->n; j++) {
        double rij = neigh_distance(c, xi, yi, zi, c->x[j], c->y[j], c->z[j]);
        if (rij <= sig) energy += INT_MAX;
    }
    c->pair_evals += (unsigned long long)c->n;
    return energy;
}
/* MC::SphericalLJ_Pot, potentials.cpp:474-504 (Baksh & Yang 1991) */
static double sphere_wall(const orc_system* s, double x, double y, double z) {
    double sizeR = s->width[2] * 0.5, dens = s->rho_s, eps = s->eps_sf, sig = s->sig_sf;
    double u1 = 0, u2 = 0;
    double xPos = x - 0.5 * s->width[0], yPos = y - 0.5 * s->width[1], zPos = z - 0.5 * s->width[2];
    double r = sqrt(ipow(xPos, 2) + ipow(yPos, 2) + ipow(zPos, 2));
    dens *= sig * sig;
    sizeR /= sig;
    r /= sig;
    double xx = sizeR - r;
    for (int i = 0; i < 10; i++) {
        u1 += 1. / ipow(sizeR, i) / ipow(xx, 10 - i);
        u1 += ipow(-1.0, i) / ipow(sizeR, i) / ipow(xx - 2.0 * sizeR, 10 - i);
        if (i < 4) {
            u2 += 1. / ipow(sizeR, i) / ipow(xx, 4 - i);
            u2 += ipow(-1.0, i) / ipow(sizeR, i) / ipow(xx - 2.0 * sizeR, 4 - i);
        }
    }
    double usf = (2. / 5.) * u1 - u2;
    usf *= 2. * M_PI * eps * dens;
    return usf;
}
/* ---- Gauss hypergeometric functions 2F1(a,b;1;z) with half-integer a, b for the cylindrical walls.  The reference evaluates them with
 * the vendored Numerical-Recipes hypgeo (NumRecipes/hypgeo.h:46: power series for |z| <= 1/2, else Bulirsch-Stoer integration of the
 * hypergeometric ODE from z0 = 1/2 — or, for z > 1, i.e. OUTSIDE the pore, from z0 = i/2 through the upper half plane, returning the real
 * part).  Restated here in closed form: F(1/2,1/2) = 2K/pi, F(-1/2,1/2) = 2E/pi, F(-1/2,-1/2) = (2/pi)(2E - (1-z)K) with the complete
 * elliptic integrals K(z), E(z) (arithmetic-geometric mean; for z > 1 the real parts of their continuation from above,
 * Re K = K(1/z)/sqrt z, Re E = sqrt z E(1/z) - (z-1)/sqrt z K(1/z)), and Gauss' contiguous relation (DLMF 15.5.11)
 *   (1-a) F(a-1,b) = -(2a - 1 + (b-a) z) F(a,b) - a (z-1) F(a+1,b)
 * lowers a parameter by one at a time down to -9/2 (stable: 1e-15 against the series on 0 <= z < 1). */
static void ellip_ke01(double m, double* K, double* E) {          /* 0 <= m < 1 */
    double a = 1.0, b = sqrt(1.0 - m), c = sqrt(m), sum = 0.5 * c * c, pw = 1.0;
    for (int n = 0; n < 12; n++) {
        double an = 0.5 * (a + b), bn = sqrt(a * b);
        c = 0.5 * (a - b); a = an; b = bn;
        sum += pw * c * c; pw *= 2.0;
    }
    *K = M_PI / (2.0 * a);
    *E = *K * (1.0 - sum);
}
static void ellip_ke(double m, double* K, double* E) {
    if (m < 1.0) { ellip_ke01(m, K, E); return; }
    double k, e, s = sqrt(m);
    ellip_ke01(1.0 / m, &k, &e);
    *K = k / s;
    *E = s * e - (m - 1.0) / s * k;
}
/* T[i][j] = 2F1(1/2 - i, 1/2 - j; 1; z), i, j = 0..5 */
static void hyp_table(double z, double T[6][6]) {
    double K, E, p = 2.0 / M_PI;
    ellip_ke(z, &K, &E);
    T[0][0] = p * K; T[1][0] = T[0][1] = p * E; T[1][1] = p * (2 * E - (1 - z) * K);
    for (int j = 0; j < 6; j++) {
        double bb = 0.5 - j;
        if (j >= 2) { T[0][j] = T[j][0]; T[1][j] = T[j][1]; }        /* symmetry in (a, b) */
        for (int i = 2; i < 6; i++) {
            double aa = 0.5 - (i - 1);                               /* lower a = aa to aa - 1 */
            T[i][j] = -((2 * aa - 1 + (bb - aa) * z) * T[i - 1][j] + aa * (z - 1) * T[i - 2][j]) / (1 - aa);
        }
    }
}
static double hyp_m45_m45(double z) { double T[6][6]; hyp_table(z, T); return T[5][5]; }
static void hyp_three_halves(double z, double* f_m15_m15, double* f_m15_m05) { double T[6][6]; hyp_table(z, T); *f_m15_m15 = T[2][2]; *f_m15_m05 = T[2][1]; }
/* MC::CylindricalLJ10_4, potentials.cpp:412-444 (Tjatjopoulos et al. 1988, summed over layers as the reference does) */
static double cyl_lj104(const orc_system* s, double y, double z) {
    int nLayers = s->n_layers;
    double delta = s->delta, sizeR = s->width[2] * 0.5, dens = s->rho_s, eps = s->eps_sf, sig = s->sig_sf, u1 = 0, u2 = 0;
    double yPos = y - 0.5 * s->width[1], zPos = z - 0.5 * s->width[2];
    double rho = sqrt(ipow(yPos, 2) + ipow(zPos, 2));
    dens *= sig * sig; sizeR /= sig; delta /= sig; rho /= sig;
    double x = sizeR - rho;
    for (int i = 0; i < nLayers; i++) {
        double Ri = sizeR + i * delta, arg = ipow(1 - x / Ri, 2), T[6][6];
        hyp_table(arg, T);
        u1 += 63. / 32. * 1. / (ipow(x, 10) * ipow(2.0 - x / Ri, 10)) * T[5][5];
        u2 += 3. / (ipow(x, 4) * ipow(2.0 - x / Ri, 4)) * T[2][2];
    }
    return (u1 - u2) * M_PI * M_PI * dens * eps;
}
/* MC::CylindricalSteele10_4_3, potentials.cpp:445-469 (Siderius & Gelb 2011); alpha = 0.61 */
static double cyl_steele(const orc_system* s, double y, double z) {
    double delta = s->delta, sizeR = s->width[2] * 0.5, dens = s->rho_s, eps = s->eps_sf, sig = s->sig_sf, alpha = 0.61;
    double yPos = y - 0.5 * s->width[1], zPos = z - 0.5 * s->width[2];
    double rho = sqrt(ipow(yPos, 2) + ipow(zPos, 2));
    double q = ipow(rho / sizeR, 2), Ra = sizeR + alpha * delta, qa = ipow(rho / Ra, 2), T[6][6], Ta[6][6];
    hyp_table(q, T);
    hyp_table(qa, Ta);
    double f15 = T[2][2], g05 = Ta[2][1];
    /* 4 sqrt(pi) Gamma(5.5)/Gamma(6), 4 sqrt(pi) Gamma(2.5)/Gamma(3), (4 sqrt(pi)/3) Gamma(2.5)/Gamma(3) */
    double psi6 = 3.0925052683774523 * ipow(sig / sizeR, 10) / ipow(1 - q, 10) * T[5][5];
    double psi3 = 4.712388980384691 * ipow(sig / sizeR, 4) / ipow(1 - q, 4) * f15;
    double phi3 = 1.5707963267948966 * ipow(sig / Ra, 3) / ipow(1 - qa, 3) * g05;
    return 2 * M_PI * dens * delta * sig * sig * eps * (psi6 - psi3 - sig / delta * phi3);
}
double orc_hyp2f1(int which, double z) {   /* test hook: 0: (-9/2,-9/2;1), 1: (-3/2,-3/2;1), 2: (-3/2,-1/2;1) */
    double a, b;
    if (which == 0) return hyp_m45_m45(z);
    hyp_three_halves(z, &a, &b);
    return which == 1 ? a : b;
}
/* MC::SlitLJ_Pot, potentials.cpp:381-402 */
double orc_slit_wall(const orc_system* s, double zabs) {
    int nLayers = s->n_layers;
    double sizeR = s->width[2] * 0.5, dens = s->rho_s, eps = s->eps_sf, sig = s->sig_sf, delta = s->delta;
    double usf = 0;
    double z = zabs - sizeR;
    for (int i = 0; i < nLayers; i++) {
        double t1 = ipow(sig / (sizeR + i * delta + z), 10);
        double t2 = ipow(sig / (sizeR + i * delta - z), 10);
        double t3 = ipow(sig / (sizeR + i * delta + z), 4);
        double t4 = ipow(sig / (sizeR + i * delta - z), 4);
        usf += 0.2 * (t1 + t2) - 0.5 * (t3 + t4);
    }
    usf *= 4. * M_PI * eps * dens * sig * sig;
    return usf;
}
/* MC::EnergyOfParticle, energy.cpp:43-105 for a position (x,y,z): out-of-pore guard adds the FINITE
 * INT_MAX and the wall term is still added (Q6); pair term from LJ126_Pot. */
static double wall_at(const orc_system* s, double x, double y, double z) {
    double boxE = 0., sqrDist = 0.;
    double sqrPoreRadius = s->width[2] * s->width[2] * 0.25;
    double xPos = x - 0.5 * s->width[0], yPos = y - 0.5 * s->width[1], zPos = z - 0.5 * s->width[2];
    if (s->geometry == 3) sqrDist = xPos * xPos + yPos * yPos + zPos * zPos;     /* energy.cpp:61-63 */
    else if (s->geometry == 2) sqrDist = yPos * yPos + zPos * zPos;
    else if (s->geometry == 1) sqrDist = zPos * zPos;
    if ((sqrDist > sqrPoreRadius) || (sqrDist < 0)) boxE = INT_MAX;
    if (s->wall_kind == 2 && s->geometry == 3) boxE += sphere_wall(s, x, y, z);   /* energy.cpp:69-77 */
    else if (s->wall_kind == 3 && s->geometry == 2) boxE += cyl_lj104(s, y, z);
    else if (s->wall_kind == 4 && s->geometry == 2) boxE += cyl_steele(s, y, z);
    else if (s->wall_kind == 1 && s->geometry == 1) boxE += orc_slit_wall(s, z);
    return boxE;
}
double orc_wall_energy(orc_chain* c, double x, double y, double z) { return wall_at(&c->sys, x, y, z); }
static void energy_at(orc_chain* c, double x, double y, double z, double* pair, double* wall) {
    *wall = wall_at(&c->sys, x, y, z);
    *pair = c->sys.pair_kind == 1 ? hard_sphere(c, x, y, z) : lj126(c, x, y, z);   /* energy.cpp:79-83 */
}
/* EnergyOfParticle on a stored slot: results are written into the particle record (energy.cpp:102-104). */
static void energy_of_slot(orc_chain* c, int slot) {
    double p, w;
    energy_at(c, c->x[slot], c->y[slot], c->z[slot], &p, &w);
    c->pe_pair[slot] = p; c->pe_wall[slot] = w; c->pe_tot[slot] = p + 0. + w;
}

void orc_particle_energy(orc_chain* c, int index0, double out[4]) {
    energy_of_slot(c, index0);
    out[0] = c->pe_pair[index0]; out[1] = 0.; out[2] = c->pe_wall[index0]; out[3] = c->pe_tot[index0];
}
void orc_trial_energy(orc_chain* c, double x, double y, double z, double out[4]) {
    int slot = c->n; /* first free slot, moves.cpp:302-304 */
    c->x[slot] = x; c->y[slot] = y; c->z[slot] = z;
    energy_of_slot(c, slot);
    out[0] = c->pe_pair[slot]; out[1] = 0.; out[2] = c->pe_wall[slot]; out[3] = c->pe_tot[slot];
}
void orc_moved_energy(orc_chain* c, int i, double x, double y, double z, double out[4]) {
    double ox = c->x[i], oy = c->y[i], oz = c->z[i], op = c->pe_pair[i], ow = c->pe_wall[i], ot = c->pe_tot[i];
    c->x[i] = x; c->y[i] = y; c->z[i] = z; /* moves.cpp:163-165 */
    energy_of_slot(c, i);
    out[0] = c->pe_pair[i]; out[1] = 0.; out[2] = c->pe_wall[i]; out[3] = c->pe_tot[i];
    c->x[i] = ox; c->y[i] = oy; c->z[i] = oz; c->pe_pair[i] = op; c->pe_wall[i] = ow; c->pe_tot[i] = ot;
}
/* MC::BoxEnergy, energy.cpp:106-124 */
static void box_energy(orc_chain* c) {
    double pairPotE = 0., boxE = 0.;
    for (int j = 0; j < c->n; j++) {
        energy_of_slot(c, j);
        pairPotE += c->pe_pair[j];
        boxE += c->pe_wall[j];
    }
    c->b_many = 0.;
    c->b_pair = 0.5 * pairPotE;
    c->b_wall = boxE;
    c->b_energy = 0. + 0.5 * pairPotE + boxE;
}
void orc_box_energy(orc_chain* c, double out[4], double* per_pair, double* per_wall) {
    box_energy(c);
    out[0] = c->b_pair; out[1] = 0.; out[2] = c->b_wall; out[3] = c->b_energy;
    c->e_pair = c->b_pair; c->e_wall = c->b_wall; c->e_energy = c->b_energy;
    for (int j = 0; j < c->n; j++) {
        if (per_pair) per_pair[j] = c->pe_pair[j];
        if (per_wall) per_wall[j] = c->pe_wall[j];
    }
}
/* MC::CorrectEnergy, energy.cpp:35-42 (quirk Q4) */
static void correct_energy(orc_chain* c) {
    double rel = c->b_energy / c->b_old;
    if (rel > 1) { box_energy(c); c->recomputes++; }
}
/* MC::PBC, energy.cpp:126-130 */
static void wrap(const orc_chain* c, double* x, double* y, double* z) {
    if (c->pbc[0]) *x -= floor(*x / c->sys.width[0]) * c->sys.width[0];
    if (c->pbc[1]) *y -= floor(*y / c->sys.width[1]) * c->sys.width[1];
    if (c->pbc[2]) *z -= floor(*z / c->sys.width[2]) * c->sys.width[2];
}

/* ---------------------------------------------------------------- moves -------------------- */

/* MC::InsertParticle slit/bulk branches, moves.cpp:46-54: x,y,z = u*width, three draws in that order. */
static void insert_particle(orc_chain* c, int slot) {
    if (c->sys.geometry == 3) {          /* moves.cpp:25-36: radius, theta, phi — NOT uniform in volume, as in the reference */
        double radius = orc_random(c) * c->sys.width[2] * 0.5;
        double theta = orc_random(c) * M_PI;
        double phi = orc_random(c) * 2 * M_PI;
        c->x[slot] = radius * sin(theta) * cos(phi);
        c->y[slot] = radius * sin(theta) * sin(phi);
        c->z[slot] = radius * cos(theta);
        c->x[slot] += 0.5 * c->sys.width[0]; c->y[slot] += 0.5 * c->sys.width[1]; c->z[slot] += 0.5 * c->sys.width[2];
        return;
    }
    if (c->sys.geometry == 2) {          /* moves.cpp:37-45: rho, phi, then x */
        double rho = orc_random(c) * c->sys.width[2] * 0.5;
        double phi = orc_random(c) * 2 * M_PI;
        c->x[slot] = (orc_random(c) - 0.5) * c->sys.width[0];
        c->y[slot] = rho * sin(phi);
        c->z[slot] = rho * cos(phi);
        c->x[slot] += 0.5 * c->sys.width[0]; c->y[slot] += 0.5 * c->sys.width[1]; c->z[slot] += 0.5 * c->sys.width[2];
        return;
    }
    c->x[slot] = orc_random(c) * c->sys.width[0];
    c->y[slot] = orc_random(c) * c->sys.width[1];
    c->z[slot] = orc_random(c) * c->sys.width[2];
}
void orc_initial_config(orc_chain* c, int n) {
    c->n = n;
    for (int k = 0; k < n; k++) insert_particle(c, k);
}
/* MC::MoveParticle, moves.cpp:118-186.  Seven draws: box, species, index, dx, dy, dz, accept. */
static int move_particle_in_box(orc_chain* c);
static int move_particle(orc_chain* c) {
    (void)(int)(orc_random(c) * c->n); /* box selection, :128-137 — single box */
    return move_particle_in_box(c);
}
static int move_particle_in_box(orc_chain* c) {
    double T = c->sys.temperature;
    (void)(int)(orc_random(c) * c->n); /* species selection, :138-147 — single species */
    c->n_displacements++; c->tot_disp++;
    c->b_old = c->b_energy;
    int i = (int)(orc_random(c) * c->n); /* reference: int(u*n)+1 on 1-based slots, :150.  n==0 -> stale slot (Q12) */
    /* ithPart = particle(index); particle(0) = ithPart — the COPY precedes the energy call, :151-152 */
    c->sx = c->x[i]; c->sy = c->y[i]; c->sz = c->z[i];
    c->s_pair = c->pe_pair[i]; c->s_wall = c->pe_wall[i]; c->s_tot = c->pe_tot[i];
    energy_of_slot(c, i);
    double oldPair = c->pe_pair[i], oldWall = c->pe_wall[i], oldEnergy = c->pe_tot[i];
    double nx = c->sx, ny = c->sy, nz = c->sz;
    nx += (orc_random(c) - 0.5) * c->dr;
    ny += (orc_random(c) - 0.5) * c->dr;
    nz += (orc_random(c) - 0.5) * c->dr;
    wrap(c, &nx, &ny, &nz);
    c->x[i] = nx; c->y[i] = ny; c->z[i] = nz;
    c->pe_pair[i] = c->s_pair; c->pe_wall[i] = c->s_wall; c->pe_tot[i] = c->s_tot; /* record copied whole, :163 */
    energy_of_slot(c, i);
    double newPair = c->pe_pair[i], newWall = c->pe_wall[i], newEnergy = c->pe_tot[i];
    double arg = (newEnergy - oldEnergy) / T;
    if (orc_random(c) < exp(-arg)) {
        c->acceptance++; c->tot_disp_acc++;
        double dPair = newPair - oldPair, dWall = newWall - oldWall;
        c->b_pair += 0.5 * dPair; c->b_wall += dWall; c->b_energy += 0. + 0.5 * dPair + dWall;
        /* exact: moving i changes pair_i AND every partner's pair_j, so the halved box total moves by the
         * FULL dPair; the reference's 0.5*dPair (moves.cpp:178) is quirk Q16 */
        if (c->n > 0) { c->e_pair += dPair; c->e_wall += dWall; c->e_energy += dPair + dWall; } /* n==0: the moved slot is a ghost (Q12) */
        return 1;
    }
    c->rejection++;
    c->x[i] = c->sx; c->y[i] = c->sy; c->z[i] = c->sz;
    c->pe_pair[i] = c->s_pair; c->pe_wall[i] = c->s_wall; c->pe_tot[i] = c->s_tot;
    return 0;
}
/* MC::SwapParticle single-box branch, moves.cpp:284-341.  Returns (kind<<1)|accepted. */
static int swap_particle(orc_chain* c) {
    double T = c->sys.temperature;
    (void)(int)(orc_random(c) * c->n); /* species selection, :285-294 */
    c->b_old = c->b_energy;
    double zz = orc_swap_zz(&c->sys); /* recomputed every call in the reference, :296-298 */
    if (orc_random(c) < 0.5) {
        c->n_swaps++; c->tot_swap++;
        int in = c->n;
        insert_particle(c, in);
        energy_of_slot(c, in);
        double inEnergy = c->pe_tot[in];
        double arg = zz * c->volume * exp(-inEnergy / T) / (c->n + 1);
        if (orc_random(c) < arg) {
            c->accept_swap++; c->tot_swap_acc++;
            c->n++;
            double pairE = 0.5 * c->pe_pair[in], wallE = c->pe_wall[in];
            c->b_pair += pairE; c->b_wall += wallE; c->b_energy += 0. + pairE + wallE;
            c->e_pair += c->pe_pair[in]; c->e_wall += wallE; c->e_energy += c->pe_pair[in] + wallE; /* full pair term (Q16) */
            return (1 << 1) | 1;
        }
        c->reject_swap++;
        return (1 << 1);
    } else if (c->n > 0) {
        c->n_swaps++; c->tot_swap++;
        int out = (int)(orc_random(c) * c->n + 1) - 1; /* int(u*n+1) on 1-based slots, :323 */
        energy_of_slot(c, out);
        double outEnergy = c->pe_tot[out];
        double arg = c->n * exp(outEnergy / T) / (zz * c->volume);
        if (orc_random(c) < arg) {
            c->accept_swap++; c->tot_swap_acc++;
            double fPair = c->pe_pair[out], fWall = c->pe_wall[out]; /* fresh and un-halved, for exact bookkeeping (Q3, Q16) */
            int last = c->n - 1;
            c->x[out] = c->x[last]; c->y[out] = c->y[last]; c->z[out] = c->z[last]; /* :329 swap-with-last */
            c->pe_pair[out] = c->pe_pair[last]; c->pe_wall[out] = c->pe_wall[last]; c->pe_tot[out] = c->pe_tot[last];
            c->n--;
            /* as shipped (Q3): the record at `out` is read AFTER the overwrite, :333-339 */
            double pairE = 0.5 * c->pe_pair[out], wallE = c->pe_wall[out];
            c->b_pair -= pairE; c->b_wall -= wallE; c->b_energy -= 0. + pairE + wallE;
            c->e_pair -= fPair; c->e_wall -= fWall; c->e_energy -= fPair + fWall;
            return (2 << 1) | 1;
        }
        c->reject_swap++;
        return (2 << 1);
    }
    return (3 << 1);
}
/* MC::ChangeVolume, NPT branch (extPress > 0), single box, moves.cpp:191-231.  Three draws: box, log-volume step, accept.
 * As shipped: RescaleCenterOfMass (moves.cpp:88-117) is called BEFORE the widths change, so its ratio is 1 and the coordinates are
 * NOT rescaled (quirk Q18); all three widths are set to ComputeBoxWidth (bulk: cbrt V); the old energy is the RUNNING box energy. */
static int change_volume(orc_chain* c) {
    double T = c->sys.temperature;
    double extPress = c->sys.pressure / ORC_KB * 1e-30; /* K/AA^3 */
    c->n_vol_changes++;
    if (!(extPress > 0)) return (4 << 1);               /* Gibbs-NVT branch: two boxes, out of scope */
    (void)(int)(orc_random(c) * 1);                      /* do { ithBox = int(Random()*nBoxes) } while (fix) */
    double oldW[3] = {c->sys.width[0], c->sys.width[1], c->sys.width[2]}, oldVol = c->volume;
    double oldB[4] = {c->b_pair, c->b_many, c->b_wall, c->b_energy};
    double logNewVol = log(c->volume) + (orc_random(c) - 0.5) * c->dv;
    c->volume = exp(logNewVol);
    c->sys.width[2] = box_width(&c->sys, c->volume);
    c->sys.width[0] = c->sys.width[1] = c->sys.width[2];
    double *sp = malloc(sizeof(double) * 3 * (size_t)(c->n + 1));   /* BoxEnergy rewrites every particle record */
    for (int j = 0; j < c->n; j++) { sp[3 * j] = c->pe_pair[j]; sp[3 * j + 1] = c->pe_wall[j]; sp[3 * j + 2] = c->pe_tot[j]; }
    box_energy(c);
    double deltaE0 = c->b_energy - oldB[3];
    double deltaVol0 = c->volume - oldVol;
    double arg0 = deltaE0 + extPress * deltaVol0 - (c->n + 1) * log(c->volume / oldVol) * T;
    int acc;
    if (orc_random(c) > exp(-arg0 / T)) {
        c->rejection_vol++;
        c->sys.width[0] = oldW[0]; c->sys.width[1] = oldW[1]; c->sys.width[2] = oldW[2]; c->volume = oldVol;    /* box = oldBox0 */
        c->b_pair = oldB[0]; c->b_many = oldB[1]; c->b_wall = oldB[2]; c->b_energy = oldB[3];
        for (int j = 0; j < c->n; j++) { c->pe_pair[j] = sp[3 * j]; c->pe_wall[j] = sp[3 * j + 1]; c->pe_tot[j] = sp[3 * j + 2]; }
        acc = 0;
    } else {
        c->acceptance_vol++;
        c->e_pair = c->b_pair; c->e_wall = c->b_wall; c->e_energy = c->b_energy;   /* exact: the fresh BoxEnergy */
        acc = 1;
    }
    free(sp);
    return (4 << 1) | acc;
}
/* ---- two boxes (Gibbs ensemble, thermoSys.nBoxes == 2) ------------------------------------------------------------------
 * Scope: the branches the stock build can run — MoveParticle's box selection and SwapParticle's Gibbs branch.  ChangeVolume's
 * Gibbs branch cannot be reached in the stock build (it aborts on entry, quirk Q19) and is not restated: n_vol must be 0. */
void orc_link_boxes(orc_chain* a, orc_chain* b) {
    a->peer = b; b->peer = a; a->box_id = 0; b->box_id = 1; b->rng_src = a; a->rng_src = NULL;
}
/* MC::MoveParticle with two boxes, moves.cpp:127-137: rand = int(u*N_total); the first box with rand <= cumulative count. */
static int gibbs_move(orc_chain* a) {
    orc_chain* box[2] = {a, a->peer};
    double rnd = (int)(orc_random(a) * (box[0]->n + box[1]->n));
    int tmp = 0, ith = 0;
    for (int i = 0; i < 2; i++) {
        tmp += box[i]->n;
        if (rnd <= tmp) { ith = i; break; }
    }
    return move_particle_in_box(box[ith]) | (ith << 4);
}
static double bar_weight(const orc_chain* c, double barC, double energy);
/* MC::SwapParticle, Gibbs branch, moves.cpp:342-400.  Returns (5<<1)|accepted (6<<1 when the donor box is empty), box bit = inBox. */
static int gibbs_swap(orc_chain* a) {
    orc_chain* box[2] = {a, a->peer};
    double T = a->sys.temperature;
    a->n_swaps++; a->tot_swap++;
    int in = (orc_random(a) < 0.5) ? 0 : 1, out = 1 - in;
    orc_chain *ci = box[in], *co = box[out];
    if (co->n <= 0) return (6 << 1) | (in << 4);
    (void)(int)(orc_random(a) * ci->n);                 /* species selection, :352-360 */
    int slot = ci->n;                                   /* inIndex = nParts+1 */
    ci->b_old = ci->b_energy; co->b_old = co->b_energy;
    insert_particle(ci, slot);
    energy_of_slot(ci, slot);
    double inEnergy = ci->pe_tot[slot];
    ci->widom_insert += bar_weight(ci, ci->bar_c, inEnergy) * ci->volume * exp(-0.5 * inEnergy / T) / (ci->n + 1);
    ci->widom_insertions++;
    int o = (int)(orc_random(a) * co->n + 1) - 1;       /* :370 */
    energy_of_slot(co, o);
    double outEnergy = co->pe_tot[o];
    co->widom_delete += bar_weight(co, co->bar_c, outEnergy) * co->volume * exp(0.5 * outEnergy / T) / (co->n + 1);
    co->widom_deletions++;
    double deltaE = inEnergy - outEnergy;
    double arg = deltaE + log(co->volume * (ci->n + 1) / (ci->volume * co->n)) * T;   /* :376 */
    if (orc_random(a) < exp(-arg / T)) {
        a->accept_swap++; a->tot_swap_acc++;
        ci->n++;
        double pairE = 0.5 * ci->pe_pair[slot], wallE = ci->pe_wall[slot];
        ci->b_pair += pairE; ci->b_wall += wallE; ci->b_energy += 0. + pairE + wallE;
        ci->e_pair += ci->pe_pair[slot]; ci->e_wall += wallE; ci->e_energy += ci->pe_pair[slot] + wallE;
        double fPair = co->pe_pair[o], fWall = co->pe_wall[o];
        int last = co->n - 1;
        co->x[o] = co->x[last]; co->y[o] = co->y[last]; co->z[o] = co->z[last];
        co->pe_pair[o] = co->pe_pair[last]; co->pe_wall[o] = co->pe_wall[last]; co->pe_tot[o] = co->pe_tot[last];
        co->n--;
        pairE = 0.5 * co->pe_pair[o]; wallE = co->pe_wall[o];          /* read after the overwrite (Q3), :392-398 */
        co->b_pair -= pairE; co->b_wall -= wallE; co->b_energy -= 0. + pairE + wallE;
        co->e_pair -= fPair; co->e_wall -= fWall; co->e_energy -= fPair + fWall;
        return (5 << 1) | 1 | (in << 4);
    }
    a->reject_swap++;
    return (5 << 1) | (in << 4);
}
/* Body of the step loop for a linked pair (call on box 0), main.cpp:67-71. */
int orc_gibbs_step(orc_chain* a, int correct) {
    int nD = a->sys.n_disp, nV = a->sys.n_vol, nS = a->sys.n_swap, code = 0;
    if (a->rng_mode == RNG_PHILOX) a->ph_slot = 0;
    double r = orc_random(a) * (nD + nV + nS);
    if (r <= nD) code = gibbs_move(a);
    else if (r <= nD + nV) code = (4 << 1);             /* not restated, see above */
    else if (r <= nD + nV + nS) code = gibbs_swap(a);
    if (correct) { correct_energy(a); correct_energy(a->peer); }
    if (a->rng_mode == RNG_PHILOX) a->ph_step++;
    return code;
}
/* Sets of a linked pair: ResetStats and AdjustMCMoves act on both boxes (MC.cpp:81-98, moves.cpp:68-87); ComputeWidom is a no-op
 * with two boxes (moves.cpp:415). */
void orc_gibbs_run_sets(orc_chain* a, int first_set, int n_sets, long long steps_per_set, int correct, unsigned char* trace) {
    orc_chain* b = a->peer;
    for (int set = first_set; set < first_set + n_sets; set++) {
        orc_reset_stats(a); orc_reset_stats(b);
        for (long long s = 0; s < steps_per_set; s++) {
            int code = orc_gibbs_step(a, correct);
            if (trace) *trace++ = (unsigned char)code;
            if (set > a->sys.n_equil_sets) {
                for (int k = 0; k < 2; k++) {
                    orc_chain* c = k ? b : a;
                    double n = (double)c->n, e = c->e_energy;
                    c->sum_n += n; c->sum_n2 += n * n; c->sum_e += e; c->sum_e2 += e * e; c->samples++;
                }
            }
        }
        orc_adjust(a, set); orc_adjust(b, set);
    }
}
/* Body of the step loop, main.cpp:67-71. */
int orc_step(orc_chain* c, int correct) {
    int nD = c->sys.n_disp, nV = c->sys.n_vol, nS = c->sys.n_swap, code = 0;
    if (c->rng_mode == RNG_PHILOX) c->ph_slot = 0;
    double r = orc_random(c) * (nD + nV + nS);
    if (r <= nD) code = move_particle(c);
    else if (r <= nD + nV) code = change_volume(c);
    else if (r <= nD + nV + nS) code = swap_particle(c);
    if (correct) correct_energy(c);
    if (c->rng_mode == RNG_PHILOX) c->ph_step++;
    return code;
}
long long orc_run(orc_chain* c, long long nsteps, int correct, unsigned char* trace) {
    long long acc = 0;
    for (long long s = 0; s < nsteps; s++) {
        int code = orc_step(c, correct);
        if (trace) trace[s] = (unsigned char)code;
        acc += code & 1;
    }
    return acc;
}
/* MC::ResetStats, MC.cpp:81-98 — note nDisplacements and nSwaps restart at ONE. */
void orc_reset_stats(orc_chain* c) {
    c->accept_swap = c->reject_swap = 0;
    c->n_swaps = 1;
    c->bar_c = 1.;
    c->widom_insertions = c->widom_deletions = 1;
    c->widom_insert = c->widom_delete = 0.;
    c->acceptance = c->rejection = 0;
    c->n_displacements = 1;
    c->n_vol_changes = 1;
    c->acceptance_vol = c->rejection_vol = 0;
}
/* MC::AdjustMCMoves displacement part, moves.cpp:68-78 */
void orc_adjust(orc_chain* c, int set) {
    if (c->n > 0) {
        if (set <= c->sys.n_equil_sets) {
            double ratio = c->acceptance * 1. / (1. * c->n_displacements);
            if (ratio < 0.2) c->dr /= 1.1;
            else c->dr *= 1.1;
        }
        if (c->sys.n_vol > 0) {                         /* volume part, moves.cpp:79-84: every set, not only during equilibration */
            double ratio = c->acceptance_vol * 1. / (1. * c->n_vol_changes);
            if (ratio < 0.5 && c->dv > 1e-3) c->dv /= 1.1;
            else c->dv *= 1.1;
        }
    }
}
void orc_run_sets(orc_chain* c, int first_set, int n_sets, long long steps_per_set, int correct, unsigned char* trace) {
    for (int set = first_set; set < first_set + n_sets; set++) {
        orc_reset_stats(c);
        for (long long s = 0; s < steps_per_set; s++) {
            int code = orc_step(c, correct);
            if (trace) *trace++ = (unsigned char)code;
            if (set > c->sys.n_equil_sets) {
                /* main.cpp:72-75: ComputeWidom(); ComputeRDF(); — the Widom draws follow the move's draws */
                if (c->rng_mode == RNG_PHILOX) { c->ph_step--; c->ph_slot = 8; }
                if (c->sys.n_swap == 0) orc_widom(c, NULL, NULL);
                if (c->rng_mode == RNG_PHILOX) c->ph_step++;
                if (c->sample_rdf) orc_rdf(c, NULL);
                double n = (double)c->n, e = c->e_energy;
                c->sum_n += n; c->sum_n2 += n * n; c->sum_e += e; c->sum_e2 += e * e;
                c->samples++;
            }
        }
        orc_adjust(c, set);
    }
}

/* ---------------------------------------------------------------- sampling (next rows) ----- */

/* MC::BarWeight, moves.cpp:409 */
static double bar_weight(const orc_chain* c, double barC, double energy) {
    return 1. / cosh(0.5 * (energy - barC) / c->sys.temperature);
}
/* MC::ComputeWidom, moves.cpp:410-449, NVT branch (no volume moves). Ghost slot = outside 0..n-1.
 * Deletion picks int(u*n) on 1-based slots: slot 0 = the saved-old-particle scratch (Q14). */
void orc_widom(orc_chain* c, double out[2], long long iout[2]) {
    double T = c->sys.temperature, pair, wall, energy;
    if (c->sys.n_swap == 0) {
        (void)(int)(orc_random(c) * 1);
        if (orc_random(c) < 0.5) {
            c->widom_insertions++;
            double gx = orc_random(c) * c->sys.width[0];
            double gy = orc_random(c) * c->sys.width[1];
            double gz = orc_random(c) * c->sys.width[2];
            energy_at(c, gx, gy, gz, &pair, &wall);
            energy = pair + wall;
            c->widom_insert += bar_weight(c, c->bar_c, energy) * exp(-0.5 * energy / T);
        } else {
            c->widom_deletions++;
            int slot1 = (int)(orc_random(c) * c->n);
            if (slot1 == 0) energy_at(c, c->sx, c->sy, c->sz, &pair, &wall);
            else energy_at(c, c->x[slot1 - 1], c->y[slot1 - 1], c->z[slot1 - 1], &pair, &wall);
            energy = pair + wall;
            if (exp(0.5 * energy / T) < 1) c->widom_delete += bar_weight(c, c->bar_c, energy) * exp(0.5 * energy / T);
            else c->widom_delete += bar_weight(c, c->bar_c, energy) * exp(-0.5 * energy / T);
        }
    }
    if (out) { out[0] = c->widom_insert; out[1] = c->widom_delete; }
    if (iout) { iout[0] = c->widom_insertions; iout[1] = c->widom_deletions; }
}
/* MC::ComputeChemicalPotential, moves.cpp:453-465 */
void orc_chem_pot(orc_chain* c, double out[2]) {
    double T = c->sys.temperature;
    double particleMass = c->sys.molar_mass / ORC_NA * 1e-3;
    double thermalWL = ORC_PLANCK / sqrt(2.0 * M_PI * particleMass * ORC_KB * T) * 1e10;
    double widomParam = (c->widom_insert / c->widom_insertions) / (c->widom_delete / c->widom_deletions);
    double muIdeal = T * log(ipow(thermalWL, 3) * (c->n + 1) / c->volume);
    double muExcess = -T * log(widomParam);
    c->mu_ex = muExcess; c->mu_tot = muIdeal + muExcess;
    out[0] = c->mu_ex; out[1] = c->mu_tot;
}
/* MC::ComputeRDF same-species branch, sample.cpp:21-31 */
void orc_rdf(orc_chain* c, double* hist) {
    double deltaR = c->sys.rcut / (1. * ORC_NBINS);
    for (int i = 0; i < c->n - 1; i++)
        for (int j = i + 1; j < c->n; j++) {
            int bin = (int)(neigh_distance(c, c->x[i], c->y[i], c->z[i], c->x[j], c->y[j], c->z[j]) / deltaR) + 1;
            if (bin < ORC_NBINS) c->rdf[bin] += 2;
        }
    if (hist) memcpy(hist, c->rdf, sizeof(double) * ORC_NBINS);
}
void orc_set_dr(orc_chain* c, double dr) { c->dr = dr; }
void orc_get_volume_state(const orc_chain* c, double d[6], long long i[3]) {
    d[0] = c->volume; d[1] = c->sys.width[0]; d[2] = c->sys.width[1]; d[3] = c->sys.width[2]; d[4] = c->dv; d[5] = c->b_energy;
    i[0] = c->acceptance_vol; i[1] = c->rejection_vol; i[2] = c->n_vol_changes;
}
void orc_rdf_reset(orc_chain* c) { memset(c->rdf, 0, sizeof(c->rdf)); }
void orc_set_sample_rdf(orc_chain* c, int on) { c->sample_rdf = on; }
void orc_get_rdf(const orc_chain* c, double* hist100) { memcpy(hist100, c->rdf, sizeof(double) * ORC_NBINS); }
void orc_get_widom(const orc_chain* c, double out[2], long long iout[2]) {
    out[0] = c->widom_insert; out[1] = c->widom_delete; iout[0] = c->widom_insertions; iout[1] = c->widom_deletions;
}

/* ---------------------------------------------------------------- lifecycle ---------------- */

orc_chain* orc_create(const orc_system* sys, int capacity) {
    orc_chain* c = (orc_chain*)calloc(1, sizeof(orc_chain));
    c->sys = *sys;
    /* PBC flags from the geometry, read.cpp:146-162 */
    c->pbc[0] = (sys->geometry != 3); c->pbc[1] = (sys->geometry == 0 || sys->geometry == 1); c->pbc[2] = (sys->geometry == 0);
    c->volume = orc_volume(sys);
    c->cap = capacity + 2;
    size_t bytes = sizeof(double) * (size_t)c->cap;
    c->x = (double*)calloc(1, bytes); c->y = (double*)calloc(1, bytes); c->z = (double*)calloc(1, bytes);
    c->pe_pair = (double*)calloc(1, bytes); c->pe_wall = (double*)calloc(1, bytes); c->pe_tot = (double*)calloc(1, bytes);
    c->dr = sys->dr;
    c->dv = sys->dv > 0 ? sys->dv : 1e-3;            /* MC.cpp:33 */
    c->n_vol_changes = 0;
    c->bar_c = 1.; c->widom_insertions = c->widom_deletions = 1; /* MC.cpp:25-28 */
    orc_rng_mt(c, 5489ULL);
    return c;
}
void orc_destroy(orc_chain* c) {
    if (!c) return;
    free(c->x); free(c->y); free(c->z); free(c->pe_pair); free(c->pe_wall); free(c->pe_tot);
    free(c);
}
void orc_set_positions(orc_chain* c, int n, const double* x, const double* y, const double* z) {
    if (n > c->cap - 2) n = c->cap - 2;
    memcpy(c->x, x, sizeof(double) * (size_t)n);
    memcpy(c->y, y, sizeof(double) * (size_t)n);
    memcpy(c->z, z, sizeof(double) * (size_t)n);
    memset(c->pe_pair, 0, sizeof(double) * (size_t)c->cap);
    memset(c->pe_wall, 0, sizeof(double) * (size_t)c->cap);
    memset(c->pe_tot, 0, sizeof(double) * (size_t)c->cap);
    c->n = n;
}
int orc_get_positions(const orc_chain* c, double* x, double* y, double* z) {
    memcpy(x, c->x, sizeof(double) * (size_t)c->n);
    memcpy(y, c->y, sizeof(double) * (size_t)c->n);
    memcpy(z, c->z, sizeof(double) * (size_t)c->n);
    return c->n;
}
void orc_get_state(const orc_chain* c, double d[16], long long i[16]) {
    d[0] = c->b_pair; d[1] = c->b_many; d[2] = c->b_wall; d[3] = c->b_energy;
    d[4] = c->dr; d[5] = c->volume; d[6] = c->b_old; d[7] = c->sys.temperature;
    d[8] = c->e_pair; d[9] = c->e_wall; d[10] = c->e_energy;
    d[11] = c->sum_n; d[12] = c->sum_n2; d[13] = c->sum_e; d[14] = c->sum_e2; d[15] = (double)c->samples;
    i[0] = c->n; i[1] = c->acceptance; i[2] = c->rejection; i[3] = c->n_displacements;
    i[4] = c->accept_swap; i[5] = c->reject_swap; i[6] = c->n_swaps; i[7] = c->n;
    i[8] = c->tot_disp; i[9] = c->tot_disp_acc; i[10] = c->tot_swap; i[11] = c->tot_swap_acc;
    i[12] = c->recomputes; i[13] = (long long)(c->pair_evals & 0x7FFFFFFFFFFFFFFFULL); i[14] = 0; i[15] = 0;
}
