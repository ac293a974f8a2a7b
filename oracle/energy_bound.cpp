// TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.
//
// The reference-side BINDING of libmcpm.so, compiled: this translation unit takes the place of the
// reference's src/energy.cpp in a second build of the UNMODIFIED reference (oracle/Makefile target `ref_bound`:
// MC.cpp moves.cpp read.cpp print.cpp sample.cpp potentials.cpp from /root/reference/src + this file + ref_harness.cpp ->
// oracle/_ref/libmcref_bound.so, linked against mc-porousmaterials_b200/libmcpm.so).  The reference keeps its own move loop,
// RNG and bookkeeping (src/moves.cpp); every energy it asks for comes from the GPU through the C ABI:
//
//   MC::EnergyOfParticle (src/energy.cpp:43-105)  -> mcpm_particle_energy   (K1)
//   MC::BoxEnergy        (src/energy.cpp:106-124) -> mcpm_box_energy        (K3)
//
// The device copy of the configuration follows the reference's moves through the three state-mutation calls of the ABI:
// an accepted translation is mirrored by mcpm_apply_move (src/moves.cpp:163), an accepted insertion by mcpm_append
// (src/moves.cpp:302-311), an accepted deletion by mcpm_remove_swap_last (src/moves.cpp:329-332).  The reference offers no hook
// at the point of acceptance, so the binding reconciles lazily: it remembers the one slot the last call touched and the particle
// count, and brings the device up to date at the next call (at most one slot and one append/remove can be pending).
// The other small functions of src/energy.cpp (MinimizeEnergy :13-34, CorrectEnergy :35-42, PBC :126-130) are restated so
// that the bound build links; they contain no energy arithmetic.
//
// tests/test_gpu_bound_reference.py drives this library through the same harness as the stock build and checks that the
// reference's own MoveParticle / SwapParticle reproduce the golden accept/reject traces of the stock build.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "MC.h"
#include "mcpm.h"

namespace {

struct Bound {
    mcpm_ctx* ctx = nullptr;
    std::vector<double> sx, sy, sz;   // what the device holds, 0-based
    int n = 0;                        // particles on the device
    int dirty = -1;                   // 1-based slot whose record may differ from the device copy
    bool synced = false;
    int capacity = 4096;
    unsigned long long calls = 0, box_calls = 0, mirrored = 0, uploads = 0;
    double seconds = 0., box_seconds = 0.;
} g;

[[noreturn]] void die(const char* what) {   // the reference's own error behaviour on this path is exit(EXIT_FAILURE)
    std::fprintf(stderr, "energy_bound: %s failed: %s\n", what, mcpm_last_error());
    std::exit(EXIT_FAILURE);
}
#define OK(call) do { if ((call) != MCPM_OK) die(#call); } while (0)

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

extern "C" {
// Called by the harness whenever it rewrites the system or the whole configuration behind the binding's back.
void mcref_bound_reset(void) {
    if (g.ctx) mcpm_destroy(g.ctx);
    g = Bound();
    if (const char* c = std::getenv("MCPM_BOUND_CAPACITY")) g.capacity = std::max(64, std::atoi(c));
}
void mcref_bound_invalidate(void) { g.synced = false; g.dirty = -1; }
// out = {EnergyOfParticle calls, seconds in them, BoxEnergy calls, seconds in them, slots mirrored, full uploads}
void mcref_bound_stats(double out[6]) {
    out[0] = (double)g.calls; out[1] = g.seconds; out[2] = (double)g.box_calls; out[3] = g.box_seconds;
    out[4] = (double)g.mirrored; out[5] = (double)g.uploads;
}
}

void MC::EnergyOfParticle(int ithBox, int ithSpecies, int index) {
    const double t0 = now();
    Box& bx = box(ithBox);
    Fluid& fl = bx.fluid(ithSpecies);
    // ---- bind: one context, one chain, described from the state ReadInputFile / the harness left behind ----
    if (!g.ctx) {
        if (const char* c = std::getenv("MCPM_BOUND_CAPACITY")) g.capacity = std::max(64, std::atoi(c));
        mcpm_system_desc d{};
        d.geometry = bx.geometry == "slit" ? MCPM_GEOM_SLIT : MCPM_GEOM_BULK;
        d.wall_kind = (fl.vdwPot(0) == "lj" && bx.geometry == "slit") ? MCPM_WALL_SLIT_LJ : MCPM_WALL_NONE;
        d.n_layers = bx.nLayersPerWall; d.rho_s = bx.solidDens; d.delta = bx.deltaLayers;
        d.eps_sf = fl.epsilon(0); d.sig_sf = fl.sigma(0);
        d.eps_ff = fluid(ithSpecies).epsilon(0); d.sig_ff = fluid(ithSpecies).sigma(0); d.rcut = fluid(ithSpecies).rcut(0);
        d.molar_mass = fluid(ithSpecies).molarMass; d.mu = fluid(ithSpecies).mu; d.temperature = thermoSys.temp;
        for (int k = 0; k < 3; k++) d.width[k] = bx.width(k);
        d.n_disp = std::max(1, sim.nDispAttempts); d.n_vol = 0; d.n_swap = sim.nSwapAttempts; d.n_equil_sets = (int)sim.nEquilSets;
        d.dr = sim.dr(ithBox);
        OK(mcpm_create(0, 1, g.capacity, &g.ctx));
        OK(mcpm_set_system(g.ctx, 0, &d));
        // the mode a host-driven move loop runs in: K1 as a resident service (no launch per call); MCREF_BOUND_LAUNCH=1 keeps one launch per call
        if (!getenv("MCREF_BOUND_LAUNCH")) OK(mcpm_k1_resident(g.ctx, 0, 1));
    }
    // ---- reconcile the device copy with what the reference's moves did since the last call ----
    const int n = fl.nParts;
    auto upload_all = [&]() {
        g.sx.resize(n); g.sy.resize(n); g.sz.resize(n);
        for (int k = 0; k < n; k++) { const Particle& q = fl.particle(k + 1); g.sx[k] = q.x; g.sy[k] = q.y; g.sz[k] = q.z; }
        OK(mcpm_set_positions(g.ctx, 0, n, g.sx.data(), g.sy.data(), g.sz.data()));
        g.n = n; g.dirty = -1; g.synced = true; g.uploads++;
    };
    if (!g.synced || std::abs(n - g.n) > 1) upload_all();
    else {
        if (n == g.n + 1) {                                   // accepted insertion: the new particle sits in slot n (src/moves.cpp:302-311)
            const Particle& q = fl.particle(n);
            const double xyz[3] = {q.x, q.y, q.z};
            OK(mcpm_append(g.ctx, 0, xyz));
            g.sx.push_back(q.x); g.sy.push_back(q.y); g.sz.push_back(q.z);
            g.n++; g.mirrored++;
        } else if (n == g.n - 1) {                            // accepted deletion of the slot last asked about (src/moves.cpp:329-332)
            if (g.dirty < 1 || g.dirty > g.n) upload_all();
            else {
                OK(mcpm_remove_swap_last(g.ctx, 0, g.dirty - 1));
                g.sx[g.dirty - 1] = g.sx[g.n - 1]; g.sy[g.dirty - 1] = g.sy[g.n - 1]; g.sz[g.dirty - 1] = g.sz[g.n - 1];
                g.sx.pop_back(); g.sy.pop_back(); g.sz.pop_back();
                g.n--; g.mirrored++;
                if (g.dirty > g.n) g.dirty = -1;              // the deleted particle was the last one
            }
        }
        if (g.synced && g.dirty >= 1 && g.dirty <= g.n && g.dirty != index) {   // an accepted (or restored) translation, src/moves.cpp:163,184
            const Particle& q = fl.particle(g.dirty);
            const int k = g.dirty - 1;
            if (q.x != g.sx[k] || q.y != g.sy[k] || q.z != g.sz[k]) {
                const double xyz[3] = {q.x, q.y, q.z};
                OK(mcpm_apply_move(g.ctx, 0, k, xyz));
                g.sx[k] = q.x; g.sy[k] = q.y; g.sz[k] = q.z; g.mirrored++;
            }
            g.dirty = -1;
        }
    }
    if (index < 0) { g.seconds += 0.; return; }             // BoxEnergy only wanted the reconciliation (see below)
    // ---- the energy itself: K1 ----
    Particle& p = fl.particle(index);
    const double xyz[3] = {p.x, p.y, p.z};
    const bool stored = (index >= 1 && index <= n);          // else: insertion slot n+1, Widom ghost MAXPART-1, scratch slot 0
    mcpm_particle_energy_t e;
    OK(mcpm_particle_energy(g.ctx, 0, stored ? index - 1 : -1, xyz, nullptr, &e));   // at the RECORD's position, itself excluded
    p.manyBodyE = e.many_body; p.pairPotE = e.pair; p.boxE = e.wall;
    p.energy = p.pairPotE; p.energy += p.manyBodyE; p.energy += p.boxE;               // src/energy.cpp:102-104
    if (stored) g.dirty = index;
    g.calls++; g.seconds += now() - t0;
}

void MC::BoxEnergy(int ithBox) {
    const double t0 = now();
    Box& bx = box(ithBox);
    EnergyOfParticle(ithBox, 0, -1);                          // bind + reconcile only
    const int n = bx.fluid(0).nParts;
    std::vector<double> pp(std::max(1, n)), pw(std::max(1, n));
    mcpm_box_energy_t be;
    OK(mcpm_box_energy(g.ctx, 0, &be, pp.data(), pw.data()));
    for (int j = 1; j <= n; j++) {                            // BoxEnergy leaves every particle's energies behind, src/energy.cpp:113-114
        Particle& p = bx.fluid(0).particle(j);
        p.manyBodyE = 0.; p.pairPotE = pp[j - 1]; p.boxE = pw[j - 1];
        p.energy = p.pairPotE; p.energy += p.manyBodyE; p.energy += p.boxE;
    }
    bx.manyBodyE = be.many_body; bx.pairPotE = be.pair; bx.boxE = be.wall; bx.energy = be.energy;   // pair already halved, :119-122
    g.box_calls++; g.box_seconds += now() - t0;
}

// ---- the rest of src/energy.cpp, restated so that the bound build links (no energy arithmetic in here) ----
void MC::MinimizeEnergy(size_t set) {                         // src/energy.cpp:13-34
    const int sweeps = 200;
    for (int i = 0; i < thermoSys.nBoxes; i++) {
        if (box(i).nParts <= 0) continue;
        if (set != 0) { BoxEnergy(i); continue; }
        std::cout << "Minimizing initial configuration of box \"" << box(i).name << "\"..." << std::endl;
        BoxEnergy(i);
        std::cout << "\tEnergy/part. of box \"" << box(i).name << "\" before minimization: " << box(i).energy / box(i).nParts << " K" << std::endl;
        for (int sweep = 0; sweep < sweeps; sweep++)
            for (int l = 0; l < box(i).nParts; l++) MoveParticle();
        BoxEnergy(i);
        std::cout << "\tEnergy/part. of box \"" << box(i).name << "\" after minimization (" << box(i).nParts * sweeps << " moves): ";
        std::cout << box(i).energy / box(i).nParts << " K" << std::endl;
    }
}
void MC::CorrectEnergy(void) {                                // src/energy.cpp:35-42 (quirk Q4: the ratio test)
    for (int i = 0; i < thermoSys.nBoxes; i++)
        if (box(i).energy / box(i).oldEnergy > 1) BoxEnergy(i);
}
void MC::PBC(int ithBox, Particle& part) {                    // src/energy.cpp:126-130
    const Box& b = box(ithBox);
    if (b.PBC(0)) part.x -= floor(part.x / b.width(0)) * b.width(0);
    if (b.PBC(1)) part.y -= floor(part.y / b.width(1)) * b.width(1);
    if (b.PBC(2)) part.z -= floor(part.z / b.width(2)) * b.width(2);
}
