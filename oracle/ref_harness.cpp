// TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.
//
// Direct-call harness around the UNMODIFIED reference (salflorom/MC-PorousMaterials).
// It is compiled together with the reference's own .cpp files, *where they lie* under
// /root/reference/src (see oracle/Makefile; nothing is copied into this repo), into
// oracle/_ref/libmcref.so.  `struct Oracle : public MC` is legal because every data member
// of MC is `protected` and every method `public` (reference src/MC.h:25-162).
//
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference legs)
// may load the library built from this file.  The product (libmcpm.so) never does.
//
// What the harness drives (reference file:line):
//   EnergyOfParticle  src/energy.cpp:43      BoxEnergy       src/energy.cpp:106
//   CorrectEnergy     src/energy.cpp:35      MoveParticle    src/moves.cpp:118
//   SwapParticle      src/moves.cpp:276      AdjustMCMoves   src/moves.cpp:68
//   ResetStats        src/MC.cpp:81          ReadInputFile   src/read.cpp:36
//   InitialConfig     src/moves.cpp:56       Random          src/MC.h:123
//   ComputeWidom      src/moves.cpp:410      ComputeRDF      src/sample.cpp:14
//   step-loop body    src/main.cpp:66-76 (re-stated in mcref_step below, call for call)

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "MC.h"

extern "C" {

// Same field order as oracle/mcpm_oracle.h::orc_system (single species, single box).
struct mcref_system {
    int geometry;        // 0 = bulk, 1 = slit, 2 = cylinder, 3 = sphere
    int wall_kind;       // 0 = none, 1 = "lj" slit wall (SlitLJ_Pot), 2 = "lj" sphere wall (SphericalLJ_Pot), 3 = "lj10-4" / 4 = "steele10-4-3" cylinder walls
    int n_layers;
    int n_disp, n_vol, n_swap;  // move weights (sim.nDispAttempts ...)
    int n_equil_sets;
    int pair_kind;       // 0 = "lj", 1 = "hs"
    double width[3];
    double temperature;
    double eps_ff, sig_ff, rcut;
    double molar_mass, mu;
    double eps_sf, sig_sf, rho_s, delta;
    double dr;
    double pressure;     // Pa (thermoSys.press); <= 0: none
    double dv;           // sim.dv (<= 0: the constructor's 1e-3)
};

// Hooks of the BOUND build (oracle/energy_bound.cpp, `make ref_bound`): absent (null) in the stock build.
void mcref_bound_reset(void) __attribute__((weak));
void mcref_bound_invalidate(void) __attribute__((weak));
void mcref_bound_stats(double out[6]) __attribute__((weak));

}  // extern "C"

struct Oracle : public MC {
    void setup(const mcref_system& s) {
        if (mcref_bound_reset) mcref_bound_reset();          // a new system: the binding creates a fresh GPU context
        thermoSys.nBoxes = 1;
        thermoSys.nSpecies = 1;
        thermoSys.temp = s.temperature;
        thermoSys.nParts = 0;
        fluid(0).name = "a";
        fluid(0).vdwPot(0) = s.pair_kind == 1 ? "hs" : "lj";
        fluid(0).sigma(0) = s.sig_ff;
        fluid(0).epsilon(0) = s.eps_ff;
        fluid(0).rcut(0) = s.rcut;
        fluid(0).molarMass = s.molar_mass;
        fluid(0).mu = s.mu;
        Box& b = box(0);
        b.name = "box";
        b.geometry = s.geometry == 1 ? "slit" : (s.geometry == 2 ? "cylinder" : (s.geometry == 3 ? "sphere" : "bulk"));
        for (int k = 0; k < 3; k++) b.width(k) = s.width[k];
        // PBC flags as ReadInputFile derives them from the geometry, src/read.cpp:146-162
        b.PBC(0) = s.geometry != 3; b.PBC(1) = (s.geometry == 0 || s.geometry == 1); b.PBC(2) = s.geometry == 0;
        b.fix = false;
        thermoSys.press = s.pressure > 0 ? s.pressure : -1.;
        sim.dv(0) = s.dv > 0 ? s.dv : 1e-3;
        b.volume = ComputeVolume(b);
        thermoSys.volume = b.volume;
        b.maxRcut = s.rcut;
        b.solidDens = s.rho_s;
        b.nLayersPerWall = s.n_layers;
        b.deltaLayers = s.delta;
        b.fluid(0).vdwPot(0) = (s.wall_kind == 1 || s.wall_kind == 2) ? "lj" : (s.wall_kind == 3 ? "lj10-4" : (s.wall_kind == 4 ? "steele10-4-3" : ""));
        b.fluid(0).sigma(0) = s.sig_sf;
        b.fluid(0).epsilon(0) = s.eps_sf;
        b.fluid(0).mu = s.mu;
        b.fluid(0).nParts = 0;
        b.nParts = 0;
        b.energy = b.oldEnergy = b.pairPotE = b.boxE = b.manyBodyE = 0.;
        sim.dr(0) = s.dr;
        sim.nEquilSets = s.n_equil_sets;
        sim.nDispAttempts = s.n_disp;
        sim.nVolAttempts = s.n_vol;
        sim.nSwapAttempts = s.n_swap;
    }
    void set_positions(int n, const double* x, const double* y, const double* z) {
        if (mcref_bound_invalidate) mcref_bound_invalidate();   // the whole configuration changed behind the binding's back
        for (int k = 0; k < n; k++) {
            Particle& p = box(0).fluid(0).particle(k + 1);
            p.x = x[k]; p.y = y[k]; p.z = z[k];
            p.energy = p.manyBodyE = p.pairPotE = p.boxE = 0.;
        }
        box(0).fluid(0).nParts = n;
        box(0).nParts = n;
        thermoSys.nParts = n;
    }
    int get_positions(double* x, double* y, double* z) {
        int n = box(0).fluid(0).nParts;
        for (int k = 0; k < n; k++) {
            Particle& p = box(0).fluid(0).particle(k + 1);
            x[k] = p.x; y[k] = p.y; z[k] = p.z;
        }
        return n;
    }
    void particle_out(int slot, double out[4]) {
        Particle& p = box(0).fluid(0).particle(slot);
        out[0] = p.pairPotE; out[1] = p.manyBodyE; out[2] = p.boxE; out[3] = p.energy;
    }
    // One pass of the body of the reference step loop, src/main.cpp:67-71.
    // Returns (move<<1)|accepted; move: 0 = translation, 1 = swap-insert, 2 = swap-delete,
    // 3 = swap attempted on empty box with delete branch (no attempt counted).
    int step(int correct_energy) {
        int nD = sim.nDispAttempts, nV = sim.nVolAttempts, nS = sim.nSwapAttempts;
        int acc0 = stats.acceptance(0), sw0 = stats.acceptSwap, nsw0 = stats.nSwaps;
        int n0 = box(0).nParts;
        Particle free0 = box(0).fluid(0).particle(n0 + 1);  // slot an insertion attempt would overwrite
        int move = -1;
        double r = Random() * (nD + nV + nS);
        int vacc0 = stats.acceptanceVol(0);
        if (r <= nD) { MoveParticle(); move = 0; }
        else if (r <= nD + nV) { ChangeVolume(); move = 4; }
        else if (r <= nD + nV + nS) { SwapParticle(); move = 1; }
        if (correct_energy) CorrectEnergy();
        int accepted = 0;
        if (move == 0) accepted = stats.acceptance(0) - acc0;
        else if (move == 4) accepted = stats.acceptanceVol(0) - vacc0;
        else if (move == 1) {
            accepted = stats.acceptSwap - sw0;
            if (stats.nSwaps == nsw0) move = 3;
            else if (accepted) move = box(0).nParts > n0 ? 1 : 2;
            else move = (box(0).fluid(0).particle(n0 + 1) != free0) ? 1 : 2;  // InsertParticle rewrote the free slot
        }
        return (move << 1) | accepted;
    }
    void seed(uint64_t s) { engine.seed(s); dis.reset(); }
    double rnd() { return Random(); }
    void particle_energy(int index0, double out[4]) {
        EnergyOfParticle(0, 0, index0 + 1);
        particle_out(index0 + 1, out);
    }
    // Energy of a trial particle placed in the first free slot (as SwapParticle's insertion does,
    // src/moves.cpp:302-304): never self-counted because it lies outside 1..n.
    void trial_energy(double x, double y, double z, double out[4]) {
        int slot = box(0).fluid(0).nParts + 1;
        Particle& p = box(0).fluid(0).particle(slot);
        p.x = x; p.y = y; p.z = z;
        EnergyOfParticle(0, 0, slot);
        particle_out(slot, out);
    }
    // Energy of particle index0 displaced to (x,y,z) — what MoveParticle evaluates at
    // src/moves.cpp:163-165; the position is restored afterwards.
    void moved_energy(int index0, double x, double y, double z, double out[4]) {
        Particle& p = box(0).fluid(0).particle(index0 + 1);
        Particle save = p;
        p.x = x; p.y = y; p.z = z;
        EnergyOfParticle(0, 0, index0 + 1);
        particle_out(index0 + 1, out);
        p = save;
    }
    void box_energy(double out[4], double* per_pair, double* per_wall) {
        BoxEnergy(0);
        out[0] = box(0).pairPotE; out[1] = box(0).manyBodyE; out[2] = box(0).boxE; out[3] = box(0).energy;
        int n = box(0).fluid(0).nParts;
        for (int k = 0; k < n; k++) {
            if (per_pair) per_pair[k] = box(0).fluid(0).particle(k + 1).pairPotE;
            if (per_wall) per_wall[k] = box(0).fluid(0).particle(k + 1).boxE;
        }
    }
    void get_state(double out[8], int iout[8]) {
        out[0] = box(0).pairPotE; out[1] = box(0).manyBodyE; out[2] = box(0).boxE; out[3] = box(0).energy;
        out[4] = sim.dr(0); out[5] = box(0).volume; out[6] = box(0).oldEnergy; out[7] = thermoSys.temp;
        iout[0] = box(0).fluid(0).nParts; iout[1] = stats.acceptance(0); iout[2] = stats.rejection(0);
        iout[3] = stats.nDisplacements(0); iout[4] = stats.acceptSwap; iout[5] = stats.rejectSwap;
        iout[6] = stats.nSwaps; iout[7] = thermoSys.nParts;
    }
    void volume_state(double d[6], int i[3]) {
        d[0] = box(0).volume; d[1] = box(0).width(0); d[2] = box(0).width(1); d[3] = box(0).width(2); d[4] = sim.dv(0); d[5] = box(0).energy;
        i[0] = stats.acceptanceVol(0); i[1] = stats.rejectionVol(0); i[2] = stats.nVolChanges;
    }
    void reset_stats() { ResetStats(); }
    void adjust(int set) { AdjustMCMoves((size_t)set); }
    void initial_config(int n) {
        box(0).fluid(0).nParts = n; box(0).nParts = n; thermoSys.nParts = n;
        for (int k = 1; k <= n; k++) InsertParticle(0, 0, k);
    }
    long long run(long long nsteps, int correct_energy, unsigned char* trace) {
        long long acc = 0;
        for (long long s = 0; s < nsteps; s++) {
            int c = step(correct_energy);
            if (trace) trace[s] = (unsigned char)c;
            acc += c & 1;
        }
        return acc;
    }
    // Widom/BAR sampling step, src/moves.cpp:410-449 (only acts when nSwapAttempts == 0).
    void widom(double out[2], int iout[2]) {
        ComputeWidom();
        out[0] = stats.widomInsert(0, 0); out[1] = stats.widomDelete(0, 0);
        iout[0] = stats.widomInsertions(0, 0); iout[1] = stats.widomDeletions(0, 0);
    }
    void chem_pot(double out[2]) {
        ComputeChemicalPotential();
        out[0] = box(0).fluid(0).muEx; out[1] = box(0).fluid(0).mu;
    }
    void rdf(int enable, double* hist) {
        sim.rdf(0) = enable ? 0 : -1; sim.rdf(1) = enable ? 0 : -1;
        if (enable) ComputeRDF();
        if (hist) for (int k = 0; k < NBINS; k++) hist[k] = stats.rdf(k);
    }
    void rdf_reset() { for (int k = 0; k < NBINS; k++) stats.rdf(k) = 0.; }  // ctor resizes stats.rdf to NBINS (src/MC.cpp:29)
    // ReadInputFile + derived quantities (src/read.cpp:36-172) for parser parity.
    int read_input(const char* path, mcref_system* s, int* n0, char* proj, char* fname, char* bname,
                   double sets[4]) {
        ReadInputFile(std::string(path));
        if (thermoSys.nBoxes != 1 || thermoSys.nSpecies != 1) return -1;
        const Box& b = box(0);
        s->geometry = b.geometry == "slit" ? 1 : (b.geometry == "bulk" ? 0 : (b.geometry == "cylinder" ? 2 : (b.geometry == "sphere" ? 3 : -1)));
        const std::string& wp = box(0).fluid(0).vdwPot(0);
        s->wall_kind = (wp == "lj" && b.geometry == "slit") ? 1 : (wp == "lj" && b.geometry == "sphere") ? 2 : (wp == "lj10-4" && b.geometry == "cylinder") ? 3 :
                       (wp == "steele10-4-3" && b.geometry == "cylinder") ? 4 : 0;
        s->pair_kind = fluid(0).vdwPot(0) == "hs" ? 1 : 0;
        s->pressure = thermoSys.press; s->dv = sim.dv(0);
        s->n_layers = b.nLayersPerWall;
        s->n_disp = sim.nDispAttempts; s->n_vol = sim.nVolAttempts; s->n_swap = sim.nSwapAttempts;
        s->n_equil_sets = (int)sim.nEquilSets;
        for (int k = 0; k < 3; k++) s->width[k] = b.width(k);
        s->temperature = thermoSys.temp;
        s->eps_ff = fluid(0).epsilon(0); s->sig_ff = fluid(0).sigma(0); s->rcut = fluid(0).rcut(0);
        s->molar_mass = fluid(0).molarMass; s->mu = fluid(0).mu;
        s->eps_sf = box(0).fluid(0).epsilon(0); s->sig_sf = box(0).fluid(0).sigma(0);
        s->rho_s = b.solidDens; s->delta = b.deltaLayers;
        s->dr = sim.dr(0);
        *n0 = box(0).fluid(0).nParts;
        snprintf(proj, 64, "%s", sim.projName.c_str());
        snprintf(fname, 64, "%s", fluid(0).name.c_str());
        snprintf(bname, 64, "%s", b.name.c_str());
        sets[0] = sim.nSets; sets[1] = sim.nEquilSets; sets[2] = sim.nStepsPerSet; sets[3] = sim.printEvery;
        return 0;
    }
    // ---- two boxes (Gibbs ensemble): box 1 of the same species.  Only what the two-box branches of MoveParticle (src/moves.cpp:127-137)
    // and SwapParticle (src/moves.cpp:342-400) read is configured.
    void add_box(const mcref_system& s) {
        thermoSys.nBoxes = 2;
        Box& b = box(1);
        b.name = "box1";
        b.geometry = s.geometry == 1 ? "slit" : (s.geometry == 2 ? "cylinder" : (s.geometry == 3 ? "sphere" : "bulk"));
        for (int k = 0; k < 3; k++) b.width(k) = s.width[k];
        b.PBC(0) = s.geometry != 3; b.PBC(1) = (s.geometry == 0 || s.geometry == 1); b.PBC(2) = s.geometry == 0;
        b.fix = false;
        sim.dv(1) = s.dv > 0 ? s.dv : 1e-3;
        b.volume = ComputeVolume(b);
        thermoSys.volume = box(0).volume + b.volume;
        b.maxRcut = s.rcut;
        b.solidDens = s.rho_s; b.nLayersPerWall = s.n_layers; b.deltaLayers = s.delta;
        b.fluid(0).vdwPot(0) = (s.wall_kind == 1 || s.wall_kind == 2) ? "lj" : (s.wall_kind == 3 ? "lj10-4" : (s.wall_kind == 4 ? "steele10-4-3" : ""));
        b.fluid(0).sigma(0) = s.sig_sf; b.fluid(0).epsilon(0) = s.eps_sf; b.fluid(0).mu = s.mu;
        b.fluid(0).nParts = 0; b.nParts = 0;
        b.energy = b.oldEnergy = b.pairPotE = b.boxE = b.manyBodyE = 0.;
        sim.dr(1) = s.dr;
    }
    void set_positions_box(int ib, int n, const double* x, const double* y, const double* z) {
        for (int k = 0; k < n; k++) {
            Particle& p = box(ib).fluid(0).particle(k + 1);
            p.x = x[k]; p.y = y[k]; p.z = z[k];
            p.energy = p.manyBodyE = p.pairPotE = p.boxE = 0.;
        }
        box(ib).fluid(0).nParts = n; box(ib).nParts = n;
        thermoSys.nParts = box(0).nParts + (thermoSys.nBoxes > 1 ? box(1).nParts : 0);
    }
    int get_positions_box(int ib, double* x, double* y, double* z) {
        int n = box(ib).fluid(0).nParts;
        for (int k = 0; k < n; k++) { Particle& p = box(ib).fluid(0).particle(k + 1); x[k] = p.x; y[k] = p.y; z[k] = p.z; }
        return n;
    }
    void box_energy_box(int ib, double out[4]) {
        BoxEnergy(ib);
        out[0] = box(ib).pairPotE; out[1] = box(ib).manyBodyE; out[2] = box(ib).boxE; out[3] = box(ib).energy;
    }
    void get_state_box(int ib, double out[8], int iout[8]) {
        out[0] = box(ib).pairPotE; out[1] = box(ib).manyBodyE; out[2] = box(ib).boxE; out[3] = box(ib).energy;
        out[4] = sim.dr(ib); out[5] = box(ib).volume; out[6] = box(ib).oldEnergy; out[7] = thermoSys.temp;
        iout[0] = box(ib).fluid(0).nParts; iout[1] = stats.acceptance(ib); iout[2] = stats.rejection(ib);
        iout[3] = stats.nDisplacements(ib); iout[4] = stats.acceptSwap; iout[5] = stats.rejectSwap;
        iout[6] = stats.nSwaps; iout[7] = thermoSys.nParts;
    }
    void widom_box(int ib, double out[2], int iout[2]) {
        out[0] = stats.widomInsert(ib, 0); out[1] = stats.widomDelete(ib, 0);
        iout[0] = stats.widomInsertions(ib, 0); iout[1] = stats.widomDeletions(ib, 0);
    }
    // Step-loop body with two boxes.  Codes as oracle/mcpm_oracle.h::orc_gibbs_step: translation | box<<4; (5<<1)|acc | inBox<<4 for a
    // transfer attempt; 6<<1 when the donor box was empty (the direction is then unobservable: no box bit).
    int gibbs_step(int correct_energy) {
        int nD = sim.nDispAttempts, nV = sim.nVolAttempts, nS = sim.nSwapAttempts;
        int acc0[2] = {(int)stats.acceptance(0), (int)stats.acceptance(1)}, nd1 = stats.nDisplacements(1);
        int sw0 = stats.acceptSwap, rj0 = stats.rejectSwap;
        int n0[2] = {box(0).nParts, box(1).nParts};
        Particle free0 = box(0).fluid(0).particle(n0[0] + 1), free1 = box(1).fluid(0).particle(n0[1] + 1);
        int code = 0;
        double r = Random() * (nD + nV + nS);
        if (r <= nD) {
            MoveParticle();
            int ib = stats.nDisplacements(1) != nd1 ? 1 : 0;
            code = (stats.acceptance(ib) - acc0[ib]) | (ib << 4);
        } else if (r <= nD + nV) {
            code = 4 << 1;                       // ChangeVolume aborts on entry in the stock build (Eigen assertion): never called here
        } else if (r <= nD + nV + nS) {
            SwapParticle();
            int accepted = stats.acceptSwap - sw0;
            if (!accepted && stats.rejectSwap == rj0) code = 6 << 1;
            else {
                int in;
                if (accepted) in = box(1).nParts > n0[1] ? 1 : 0;
                else in = (box(1).fluid(0).particle(n0[1] + 1) != free1) ? 1 : 0;
                (void)free0;
                code = (5 << 1) | accepted | (in << 4);
            }
        }
        if (correct_energy) CorrectEnergy();
        return code;
    }
    double volume() { return box(0).volume; }
    double swap_zz() {  // src/moves.cpp:296-298
        double particleMass = fluid(0).molarMass / na * 1e-3;
        double thermalWL = planck / sqrt(2.0 * pi * particleMass * kb * thermoSys.temp) * 1e10;
        return exp(fluid(0).mu / thermoSys.temp) / (thermalWL * thermalWL * thermalWL);
    }
};

static Oracle* ops(void* h) { return static_cast<Oracle*>(h); }

extern "C" {

void* mcref_create(void) { return new Oracle(); }
void mcref_destroy(void* h) { delete ops(h); }
void mcref_setup(void* h, const mcref_system* s) { ops(h)->setup(*s); }
void mcref_set_positions(void* h, int n, const double* x, const double* y, const double* z) {
    ops(h)->set_positions(n, x, y, z);
}
int mcref_get_positions(void* h, double* x, double* y, double* z) { return ops(h)->get_positions(x, y, z); }
void mcref_seed(void* h, uint64_t s) { ops(h)->seed(s); }
double mcref_random(void* h) { return ops(h)->rnd(); }
void mcref_particle_energy(void* h, int index0, double out[4]) { ops(h)->particle_energy(index0, out); }
void mcref_trial_energy(void* h, double x, double y, double z, double out[4]) { ops(h)->trial_energy(x, y, z, out); }
void mcref_moved_energy(void* h, int index0, double x, double y, double z, double out[4]) {
    ops(h)->moved_energy(index0, x, y, z, out);
}
void mcref_box_energy(void* h, double out[4], double* per_pair, double* per_wall) {
    ops(h)->box_energy(out, per_pair, per_wall);
}
void mcref_get_state(void* h, double out[8], int iout[8]) { ops(h)->get_state(out, iout); }
void mcref_reset_stats(void* h) { ops(h)->reset_stats(); }
void mcref_volume_state(void* h, double d[6], int i[3]) { ops(h)->volume_state(d, i); }
void mcref_adjust(void* h, int set) { ops(h)->adjust(set); }
void mcref_initial_config(void* h, int n) { ops(h)->initial_config(n); }
int mcref_step(void* h, int correct_energy) { return ops(h)->step(correct_energy); }
long long mcref_run(void* h, long long nsteps, int correct_energy, unsigned char* trace) {
    return ops(h)->run(nsteps, correct_energy, trace);
}
// Timed run: returns seconds for nsteps passes of the step-loop body (steady_clock).
double mcref_run_timed(void* h, long long nsteps, int correct_energy) {
    auto t0 = std::chrono::steady_clock::now();
    ops(h)->run(nsteps, correct_energy, nullptr);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}
double mcref_box_energy_timed(void* h, int reps) {
    double out[4];
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; r++) ops(h)->box_energy(out, nullptr, nullptr);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}
void mcref_widom(void* h, double out[2], int iout[2]) { ops(h)->widom(out, iout); }
void mcref_chem_pot(void* h, double out[2]) { ops(h)->chem_pot(out); }
void mcref_rdf(void* h, int enable, double* hist) { ops(h)->rdf(enable, hist); }
void mcref_rdf_reset(void* h) { ops(h)->rdf_reset(); }
int mcref_read_input(void* h, const char* path, mcref_system* s, int* n0, char* proj, char* fname,
                     char* bname, double sets[4]) {
    return ops(h)->read_input(path, s, n0, proj, fname, bname, sets);
}
// 1 if this library is the bound build; out = {EnergyOfParticle calls, seconds, BoxEnergy calls, seconds, slots mirrored, uploads}
int mcref_is_bound(double out[6]) {
    if (!mcref_bound_stats) return 0;
    if (out) mcref_bound_stats(out);
    return 1;
}
void mcref_add_box(void* h, const mcref_system* s) { ops(h)->add_box(*s); }
void mcref_set_positions_box(void* h, int ib, int n, const double* x, const double* y, const double* z) { ops(h)->set_positions_box(ib, n, x, y, z); }
int mcref_get_positions_box(void* h, int ib, double* x, double* y, double* z) { return ops(h)->get_positions_box(ib, x, y, z); }
void mcref_box_energy_box(void* h, int ib, double out[4]) { ops(h)->box_energy_box(ib, out); }
void mcref_get_state_box(void* h, int ib, double out[8], int iout[8]) { ops(h)->get_state_box(ib, out, iout); }
void mcref_widom_box(void* h, int ib, double out[2], int iout[2]) { ops(h)->widom_box(ib, out, iout); }
int mcref_gibbs_step(void* h, int correct_energy) { return ops(h)->gibbs_step(correct_energy); }
long long mcref_gibbs_run(void* h, long long nsteps, int correct_energy, unsigned char* trace) {
    long long acc = 0;
    for (long long s = 0; s < nsteps; s++) {
        int c = ops(h)->gibbs_step(correct_energy);
        if (trace) trace[s] = (unsigned char)c;
        acc += c & 1;
    }
    return acc;
}
double mcref_volume(void* h) { return ops(h)->volume(); }
double mcref_swap_zz(void* h) { return ops(h)->swap_zz(); }

}  // extern "C"
