// C++ host driver above the C ABI: re-hosts the reference's main.cpp / read.cpp / print.cpp behaviour
// (single box, single species) on top of the GPU engine.  Compiled into libmcpm.so; mcpm_simulate is
// also the body of the `mcpm_simulate` executable (simulate_main.cpp).
//
//   mcpm_read_input_file  == MC::ReadInputFile        reference src/read.cpp:36-172
//   print_params          == MC::PrintParams          src/print.cpp:16-104
//   output_directory      == MC::OutputDirectory      src/print.cpp:105-124
//   print_trajectory      == MC::PrintTrajectory      src/print.cpp:163-192
//   print_log             == MC::PrintLog             src/print.cpp:193-233
//   print_stats           == MC::PrintStats           src/print.cpp:234-264
//   mcpm_simulate         == main                     src/main.cpp:16-97
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <filesystem>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/mcpm.h"

namespace mcpm { void set_last_error(const char* msg); }   // mcpm_api.cu: what mcpm_last_error() returns

namespace {

// Tools::LowerCase, src/tools.h:19-30: lower-case, strip one trailing ';'.
std::string lower(std::string s) {
    for (auto& ch : s) ch = (char)tolower((unsigned char)ch);
    if (!s.empty() && s.back() == ';') s.pop_back();
    return s;
}
// Tools::SplitString(line, ' '), src/tools.h:31-43: consecutive blanks yield empty tokens, exactly
// like getline(ss, word, ' ').
std::vector<std::string> split(const std::string& line) {
    std::vector<std::string> out;
    std::stringstream ss(line);
    std::string w;
    while (!ss.eof()) {
        std::getline(ss, w, ' ');
        out.push_back(w);
        if (out.size() >= 50) break;
    }
    while (out.size() < 4) out.push_back("");
    return out;
}
// Tools::Pow, src/tools.h:53-58
double ipow(double v, int p) { double r = 1.; for (int i = 1; i <= p; i++) r *= v; return r; }

void copy_name(char* dst, const std::string& s) { snprintf(dst, 64, "%s", s.c_str()); }

// MC::Random(), src/MC.h:28-33,123: mt19937_64 through uniform_real_distribution{0,0.9999} (quirk Q1).
struct RefRandom {
    std::mt19937_64 engine;
    std::uniform_real_distribution<double> dis{0.0, 0.9999};
    explicit RefRandom(unsigned long long seed) : engine(seed) {}
    double operator()() { return dis(engine); }
};

}  // namespace

static int read_input_file(const char* path, mcpm_sim_config* cfg);

extern "C" int mcpm_read_input_file(const char* path, mcpm_sim_config* cfg) {
    if (!path || !cfg) { mcpm::set_last_error("null argument"); return MCPM_ERR_ARG; }
    try {
        return read_input_file(path, cfg);
    } catch (const std::exception& e) {     // std::stod / stoi on a malformed or missing value: never across the C boundary
        mcpm::set_last_error((std::string("input file ") + path + ": malformed value (" + e.what() + ")").c_str());
        return MCPM_ERR_ARG;
    } catch (...) {
        mcpm::set_last_error("input file: unknown error");
        return MCPM_ERR_ARG;
    }
}

static int read_input_file(const char* path, mcpm_sim_config* cfg) {
    std::ifstream in(path);
    if (!in.is_open()) { mcpm::set_last_error((std::string("cannot open ") + path).c_str()); return MCPM_ERR_IO; }   // the reference prints "Error opening file" and exits, src/read.cpp:46-49
    memset(cfg, 0, sizeof(*cfg));
    cfg->pressure = -1.;
    cfg->replicas = 1;
    cfg->devices = 1;
    // box-independent fields are parsed into box 0's description and copied to box 1's afterwards
    mcpm_system_desc& d = cfg->system;
    d.temperature = -1.;
    std::string fluid_name, box_name[2], geometry[2] = {"bulk", "bulk"}, ff_pot, sf_pot[2];   // defaults of MC::MC(), src/MC.cpp:56
    double size[2] = {0., 0.}, rho_s[2] = {0., 0.}, delta[2] = {0., 0.}, sig_sf[2] = {0., 0.}, eps_sf[2] = {0., 0.};
    int n_layers[2] = {0, 0}, n_parts[2] = {0, 0};
    int bi = 0;                                    // the box the box-level keywords refer to: the last BoxName (src/read.cpp:72-80)
    std::string line;
    while (std::getline(in, line)) {
        auto c = split(line);
        for (int i = 0; i < 4; i++) c[i] = lower(c[i]);
        const std::string& k = c[0];
        if (k == "projectname") copy_name(cfg->project, c[1]);
        else if (k == "continueaftercrash") cfg->continue_after_crash = 1;
        else if (k == "productionsets") cfg->n_sets = std::stod(c[1]);
        else if (k == "equilibriumsets") cfg->n_equil_sets = std::stod(c[1]);
        else if (k == "stepsperset") cfg->steps_per_set = std::stod(c[1]);
        else if (k == "printeverynsets") cfg->print_every = std::stod(c[1]);
        else if (k == "ndisplacementattempts") d.n_disp = std::stoi(c[1]);
        else if (k == "nvolumeattempts") d.n_vol = std::stoi(c[1]);
        else if (k == "nswapattempts") d.n_swap = std::stoi(c[1]);
        else if (k == "externaltemperature") d.temperature = std::stod(c[1]);
        else if (k == "externalpressure") cfg->pressure = std::stod(c[1]);
        else if (k == "fluidname") { cfg->n_species++; fluid_name = c[1]; }
        else if (k == "molarmass") d.molar_mass = std::stod(c[1]);
        else if (k == "chemicalpotential") d.mu = std::stod(c[1]);
        else if (k == "boxname") { cfg->n_boxes++; bi = std::min(cfg->n_boxes - 1, 1); box_name[bi] = c[1]; }
        else if (k == "geometry") geometry[bi] = c[1];
        else if (k == "surfacedensity") rho_s[bi] = std::stod(c[1]);
        else if (k == "nlayersperwall") n_layers[bi] = std::stoi(c[1]);
        else if (k == "distancebetweenlayers") delta[bi] = std::stod(c[1]);
        else if (k == "size") size[bi] = std::stod(c[1]);
        else if (k == "fixvolume") cfg->fix_volume[bi] = 1;
        else if (c[2] == "vdwpotential" || c[2] == "sigma" || c[2] == "epsilon" || c[2] == "rcut" || c[2] == "numberofparticles") {
            // addressing rule of src/read.cpp:82-111: "<fluid> <fluid> …" = pair term, "<box> <fluid> …" = wall term
            const bool a_fluid = (c[0] == fluid_name && !fluid_name.empty()), b_fluid = (c[1] == fluid_name && !fluid_name.empty());
            for (int b = 0; b < 2; b++) {
                const bool a_box = (c[0] == box_name[b] && !box_name[b].empty());
                if (!(a_box && b_fluid)) continue;
                if (c[2] == "vdwpotential") sf_pot[b] = c[3];
                else if (c[2] == "sigma") sig_sf[b] = std::stod(c[3]);
                else if (c[2] == "epsilon") eps_sf[b] = std::stod(c[3]);
                else if (c[2] == "numberofparticles") n_parts[b] = std::stoi(c[3]);
                break;                             // Tools::FindIndex returns the first box of that name
            }
            if (a_fluid && b_fluid) {
                if (c[2] == "vdwpotential") ff_pot = c[3];
                else if (c[2] == "sigma") d.sig_ff = std::stod(c[3]);
                else if (c[2] == "epsilon") d.eps_ff = std::stod(c[3]);
                else if (c[2] == "rcut") d.rcut = std::stod(c[3]);
            }
        }
        else if (k == "samplerdf") cfg->sample_rdf = (c[1] == fluid_name && c[2] == fluid_name) ? 1 : 0;
        else if (k == "printtrajectory") cfg->print_trajectory = 1;
        // ---- extensions (the reference silently ignores unknown lines) ----
        else if (k == "seed") { cfg->has_seed = 1; cfg->seed = std::stoull(c[1]); }
        else if (k == "replicas") cfg->replicas = std::max(1, std::stoi(c[1]));
        else if (k == "chemicalpotentialsweep") { cfg->sweep_mu0 = std::stod(c[1]); cfg->sweep_mu1 = std::stod(c[2]); cfg->sweep_points = std::stoi(c[3]); }
        else if (k == "capacity") cfg->capacity = std::stoi(c[1]);
        else if (k == "devices") cfg->devices = std::max(1, std::stoi(c[1]));
    }
    copy_name(cfg->fluid, fluid_name);
    copy_name(cfg->box, box_name[0]);
    copy_name(cfg->box2, box_name[1]);
    cfg->n_particles = n_parts[0]; cfg->n_particles2 = n_parts[1];
    // pair dispatch of src/energy.cpp:79-101: "lj", "hs"; the EAM potentials are recognised and refused; anything else gives no pair energy
    d.pair_kind = ff_pot == "hs" ? MCPM_PAIR_HS : (ff_pot.rfind("eam_", 0) == 0 ? 2 : MCPM_PAIR_LJ);
    if (ff_pot != "lj" && ff_pot != "hs" && d.pair_kind == MCPM_PAIR_LJ) d.eps_ff = 0.0;      // no potential matched: an ideal gas
    d.pressure = cfg->pressure;
    d.dv = 1e-3;               // src/MC.cpp:33
    d.n_equil_sets = (int)cfg->n_equil_sets;
    cfg->system2 = d;          // the fluid, temperature and move weights are the system's
    // per-box post-processing, src/read.cpp:122-168
    const double maxRcut = d.rcut;
    for (int b = 0; b < 2; b++) {
        mcpm_system_desc& x = b ? cfg->system2 : cfg->system;
        const std::string& geo = geometry[b];
        x.rho_s = rho_s[b]; x.n_layers = n_layers[b]; x.delta = delta[b]; x.sig_sf = sig_sf[b]; x.eps_sf = eps_sf[b];
        x.width[2] = size[b];
        if (geo == "slit") { x.geometry = MCPM_GEOM_SLIT; x.width[0] = x.width[1] = 2 * maxRcut; }   // :134
        else if (geo == "bulk") { x.geometry = MCPM_GEOM_BULK; x.width[0] = x.width[1] = x.width[2]; } // :135
        else if (geo == "sphere") { x.geometry = MCPM_GEOM_SPHERE; x.width[0] = x.width[1] = x.width[2]; }   // :130
        else if (geo == "cylinder") { x.geometry = MCPM_GEOM_CYLINDER; x.width[0] = 2 * maxRcut; x.width[1] = x.width[2]; }   // :131-133
        else x.geometry = -1;
        // wall dispatch of src/energy.cpp:69-77: ("lj", slit) -> SlitLJ_Pot, ("lj", sphere) -> SphericalLJ_Pot, ("lj10-4" | "steele10-4-3", cylinder)
        // -> the two cylindrical walls; anything else matches nothing, e.g. "steele10-4-3" + slit
        x.wall_kind = (sf_pot[b] == "lj" && geo == "slit") ? MCPM_WALL_SLIT_LJ : (sf_pot[b] == "lj" && geo == "sphere") ? MCPM_WALL_SPHERE_LJ :
                      (sf_pot[b] == "lj10-4" && geo == "cylinder") ? MCPM_WALL_CYL_LJ104 : (sf_pot[b] == "steele10-4-3" && geo == "cylinder") ? MCPM_WALL_CYL_STEELE : MCPM_WALL_NONE;
        x.dr = 0.1 * x.width[2];   // :168
    }
    return MCPM_OK;
}

namespace {

struct Outputs {
    std::string root;       // "./<project>" (or <out_dir>/<project>)
    std::string box_dir;    // root/<box>
};

double box_volume(const mcpm_system_desc& d) {  // MC::ComputeVolume, src/read.cpp:26-35
    if (d.geometry == MCPM_GEOM_SPHERE) return (M_PI / 6.) * ipow(d.width[2], 3);
    if (d.geometry == MCPM_GEOM_CYLINDER) return (M_PI / 4.) * ipow(d.width[2], 2) * d.width[0];
    return d.geometry == MCPM_GEOM_SLIT ? d.width[0] * d.width[1] * d.width[2] : ipow(d.width[2], 3);
}
const char* geometry_name(int g) { return g == MCPM_GEOM_SLIT ? "slit" : g == MCPM_GEOM_SPHERE ? "sphere" : g == MCPM_GEOM_CYLINDER ? "cylinder" : "bulk"; }

// MC::InsertParticle, src/moves.cpp:21-55: sphere (radius, theta, phi), cylinder (rho, phi, x), slit / bulk (x, y, z)
template <class Rnd>
void insert_particle(const mcpm_system_desc& d, Rnd& rnd, double& x, double& y, double& z) {
    if (d.geometry == MCPM_GEOM_SPHERE) {
        const double radius = rnd() * d.width[2] * 0.5, theta = rnd() * M_PI, phi = rnd() * 2 * M_PI;
        x = radius * sin(theta) * cos(phi) + 0.5 * d.width[0]; y = radius * sin(theta) * sin(phi) + 0.5 * d.width[1]; z = radius * cos(theta) + 0.5 * d.width[2];
    } else if (d.geometry == MCPM_GEOM_CYLINDER) {
        const double rho = rnd() * d.width[2] * 0.5, phi = rnd() * 2 * M_PI;
        x = (rnd() - 0.5) * d.width[0] + 0.5 * d.width[0]; y = rho * sin(phi) + 0.5 * d.width[1]; z = rho * cos(phi) + 0.5 * d.width[2];
    } else { x = rnd() * d.width[0]; y = rnd() * d.width[1]; z = rnd() * d.width[2]; }
}

// MC::PrintParams, src/print.cpp:16-104 (same wording and order)
void print_params(const mcpm_sim_config& c, std::ostream& os) {
    const mcpm_system_desc& d = c.system;
    os << "Simulation parameters:" << std::endl;
    os << "\tEquilibration sets: " << c.n_equil_sets << std::endl;
    os << "\tProduction sets: " << c.n_sets << std::endl;
    os << "\tSteps per set: " << c.steps_per_set << std::endl;
    os << "\tPrint every n sets: " << c.print_every << std::endl;
    os << "\tNum. of displacement attempts: " << d.n_disp << std::endl;
    os << "\tNum. of change volume attempts: " << d.n_vol << std::endl;
    os << "\tNum. of swap attempts: " << d.n_swap << std::endl;
    os << std::endl;
    const int nb = c.n_boxes == 2 ? 2 : 1;
    const mcpm_system_desc* bd[2] = {&c.system, &c.system2};
    const char* bname[2] = {c.box, c.box2};
    const int bn[2] = {c.n_particles, c.n_particles2};
    double vol_total = -1.;                                          // thermoSys.volume starts at -1 and is += per box (src/MC.cpp:37, src/read.cpp:137)
    int n_total = 0;
    for (int b = 0; b < nb; b++) { vol_total += box_volume(*bd[b]); n_total += bn[b]; }
    os << "System parameters:" << std::endl;
    os << "\tNum. of particles: " << n_total << std::endl;
    os << "\tVolume (AA^3): " << vol_total << std::endl;
    os << "\tExternal temperature (K): " << d.temperature << std::endl;
    if (c.pressure >= 0) os << "\tExternal pressure (Pa): " << c.pressure << std::endl;
    os << std::endl;
    os << "Fluid parameters:" << std::endl;
    os << "Species 0: " << c.fluid << std::endl;
    os << "\tMolar mass (g/mol): " << d.molar_mass << std::endl;
    if (d.mu != 0) os << "\tChemical potential (K): " << d.mu << std::endl;
    os << std::endl;
    os << "Fluid-fluid interaction parameters:" << std::endl;
    os << c.fluid << "-" << c.fluid << std::endl;
    os << "\tVdW potential: " << (d.pair_kind == MCPM_PAIR_HS ? "hs" : "lj") << std::endl;
    os << "\tsigma (AA): " << d.sig_ff << std::endl;
    os << "\tepsilon (K): " << d.eps_ff << std::endl;
    os << "\trcut (AA): " << d.rcut << std::endl;
    os << std::endl;
    os << "Box parameters:" << std::endl;
    for (int b = 0; b < nb; b++) {
        const mcpm_system_desc& x = *bd[b];
        os << "Box " << b << ": " << bname[b] << std::endl;
        os << "\tGeometry: " << geometry_name(x.geometry) << std::endl;
        os << "\tInitial width (AA): " << x.width[2] << std::endl;
        if (x.rho_s > 0) os << "\tSurface density (AA^-2): " << x.rho_s << std::endl;
        if (x.n_layers > 1) os << "\tLayers per wall: " << x.n_layers << std::endl;
        if (x.delta > 0) os << "\tDistance between layers (AA): " << x.delta << std::endl;
        os << "\tInitial number of particles: " << bn[b] << std::endl;
        if (c.fix_volume[b]) os << "\tVolume fixed." << std::endl;
    }
    os << std::endl;
    os << "Box-fluid interaction parameters:" << std::endl;
    for (int b = 0; b < nb; b++) {
        const mcpm_system_desc& x = *bd[b];
        os << bname[b] << "-" << c.fluid << std::endl;
        if (x.wall_kind == MCPM_WALL_SLIT_LJ || x.wall_kind == MCPM_WALL_SPHERE_LJ) os << "\tVdW potential: lj" << std::endl;
        else if (x.wall_kind == MCPM_WALL_CYL_LJ104) os << "\tVdW potential: lj10-4" << std::endl;
        else if (x.wall_kind == MCPM_WALL_CYL_STEELE) os << "\tVdW potential: steele10-4-3" << std::endl;
        if (x.sig_sf > 0) os << "\tsigma (AA): " << x.sig_sf << std::endl;
        if (x.eps_sf > 0) os << "\tepsilon (K): " << x.eps_sf << std::endl;
        os << "\tInitial number of particles: " << bn[b] << std::endl;
    }
    os << std::endl;
    if (c.sample_rdf) os << "Sample RDF: " << c.fluid << "-" << c.fluid << std::endl;
    os << (c.print_trajectory ? "Print Trajectory: Yes" : "Print Trajectory: No") << std::endl;
    os << std::endl;
}

// MC::OutputDirectory, src/print.cpp:105-124
Outputs output_directory(const std::string& base, const std::string& project, const std::string& box) {
    Outputs o;
    o.root = base + "/" + project;
    o.box_dir = o.root + "/" + box;
    std::filesystem::create_directories(o.box_dir);
    return o;
}

// MC::PrintTrajectory, src/print.cpp:163-192 (default ostream formatting = 6 significant digits)
void print_trajectory(const Outputs& o, const mcpm_sim_config& c, const mcpm_system_desc& d, size_t set, double dr, int n,
                      const double* x, const double* y, const double* z) {
    std::ofstream f;
    const std::string name = o.box_dir + "/trajectory.exyz";
    if (set == 0 || !c.print_trajectory) f.open(name);
    else f.open(name, std::ios::app);
    const double tmp = 0.0, dv = 1e-3;   // sim.dv default, src/MC.cpp:33
    f << n << "\n";
    f << "Lattice=\" " << d.width[0] << " " << tmp << " " << tmp << " ";
    f << tmp << " " << d.width[1] << " " << tmp << " ";
    f << tmp << " " << tmp << " " << d.width[2] << "\" ";
    f << "Properties=species:S:1:pos:R:3 " << "Set: " << set << " Disp: " << dr << " VolDisp: " << dv << "\n";
    for (int k = 0; k < n; k++) f << c.fluid << "\t" << x[k] << "\t" << y[k] << "\t" << z[k] << "\n";
}

// MC::PrintLog, src/print.cpp:193-233.  q5 = reproduce the double-halved pair column (quirk Q5).
void print_log(const Outputs& o, const mcpm_sim_config& c, const mcpm_system_desc& d, size_t set, const mcpm_chain_stats* st, bool q5,
               bool have_mu = false, double mu_ex = 0., double mu = 0.) {
    const std::string name = o.box_dir + "/simulation_" + c.fluid + ".log";
    std::ofstream f;
    if (set == 0) {
        f.open(name);
        f << "Set\tTemp(K)\twidth(AA)\tVolume(AA^3)\t";
        f << "ffE/particle(K)\tsfE/particle(K)\tE/particle(K)\t";
        f << "NParts\t" << "Dens(g/cm^3)\t" << "muEx(K)\t" << "mu(K)\n";
        return;
    }
    const double na = 6.02214076e23;
    const bool npt = d.n_vol > 0 && st->volume > 0;         // NPT: the box the chain ended the set with
    const double volume = npt ? st->volume : box_volume(d), width = npt ? st->width : d.width[2];
    const double ff = q5 ? 0.5 * st->pair : st->pair;      // st->pair is the box pair energy (already halved once)
    const double energy = ff + st->wall;
    double density = st->n / volume;
    density *= d.molar_mass / na * 1e24;
    f.open(name, std::ios::app);
    f << std::fixed << std::setprecision(7);
    f << set << "\t" << d.temperature << "\t" << width << "\t" << volume << "\t";
    f << ff / st->n << "\t" << st->wall / st->n << "\t" << energy / st->n << "\t";
    f << st->n << "\t" << density << "\t";
    // muEx / mu: with swap moves on, Widom sampling is off (src/moves.cpp:415) and ComputeChemicalPotential divides
    // 0 by 0: the reference's log shows "nan" and "-nan" in these two columns
    if (have_mu) f << mu_ex << "\t" << mu << "\n";      // NVT: Widom/BAR values of the last set (ComputeChemicalPotential)
    else f << "nan" << "\t" << "-nan" << "\n";
}

// MC::PrintStats, src/print.cpp:234-264: a header, one block per box, a blank line
void print_stats_box(const mcpm_sim_config& c, const char* box, const mcpm_system_desc& d, size_t set, const mcpm_chain_stats* st, std::ostream& os) {
    os << "Box " << box << ":" << std::endl;
    if (d.n_vol > 0) {
        os << "\tAcceptVolRatio; RejectVolRatio:\t";
        os << st->accept_vol * 1. / st->n_vol_changes << "; " << st->reject_vol * 1. / st->n_vol_changes << std::endl;
        if (set < c.n_equil_sets) os << "\tVolume step size: " << st->dv << std::endl;
    }
    if (d.n_swap > 0) {
        os << "\tAcceptSwapRatio; RejectSwapRatio:\t";
        os << st->accept_swap * 1. / st->n_swaps << "; " << st->reject_swap * 1. / st->n_swaps << std::endl;
    }
    if (st->n_displacements > 0) {
        os << "\tAcceptDispRatio; RegectDispRatio:\t";
        os << st->acceptance * 1. / st->n_displacements << "; " << st->rejection * 1. / st->n_displacements << std::endl;
        if (set < c.n_equil_sets) os << "\tDisplacement step size: " << st->dr << std::endl;
    }
    os << "\tNumParticles; BoxSize; Energy/Particle:\t";
    os << st->n << "; " << ((d.n_vol > 0 && st->width > 0) ? st->width : d.width[2]) << "; " << st->energy / st->n << std::endl;
}
void print_stats_header(const mcpm_sim_config& c, size_t set, std::ostream& os) {
    os << std::fixed << std::setprecision(7);
    if (set < c.n_equil_sets) os << "Equilibrium set: " << set << std::endl;
    else os << "Set: " << set << std::endl;
}
void print_stats(const mcpm_sim_config& c, const mcpm_system_desc& d, size_t set, const mcpm_chain_stats* st, std::ostream& os) {
    print_stats_header(c, set, os);
    print_stats_box(c, c.box, d, set, st, os);
    os << std::endl;   // like the reference, the stream stays fixed/7 for everything printed afterwards
}

int host_fail(int code, const std::string& msg) { mcpm::set_last_error(msg.c_str()); std::cerr << "mcpm_simulate: " << msg << std::endl; return code; }

// Contexts of a run: destroyed on every way out of mcpm_simulate.
struct ContextSet {
    std::vector<mcpm_ctx*> ctx;
    ~ContextSet() { for (auto* c : ctx) if (c) mcpm_destroy(c); }
};

// One device's share of the chains: the set loop of src/main.cpp:63-84 for chains [lo, hi).
struct Shard {
    int device = 0, lo = 0, hi = 0;
    mcpm_ctx* ctx = nullptr;
    int rc = MCPM_OK;
    std::string err;
};

}  // namespace

static int simulate(const char* input_path, const char* out_dir, int flags);

extern "C" int mcpm_simulate(const char* input_path, const char* out_dir, int flags) {
    try {
        return simulate(input_path, out_dir, flags);
    } catch (const std::filesystem::filesystem_error& e) {   // unwritable output directory
        return host_fail(MCPM_ERR_IO, e.what());
    } catch (const std::exception& e) {                      // nothing may cross the C boundary
        return host_fail(MCPM_ERR_ARG, e.what());
    } catch (...) {
        return host_fail(MCPM_ERR_ARG, "unknown error");
    }
}

// Two boxes (thermoSys.nBoxes == 2, the Gibbs ensemble at fixed volumes): the same program for a PAIR of linked chains per replica
// (mcpm_link_boxes, k2_gibbs): InitialConfig and MinimizeEnergy per box, the set loop with PrintStats / PrintTrajectory / PrintLog for both
// boxes in the reference's formats (src/main.cpp:16-97, src/energy.cpp:13-34, src/print.cpp).  Replica r is the pair of chains (2r, 2r+1).
static int simulate_two_boxes(mcpm_sim_config& cfg, const char* out_dir, int flags, std::ostream& log, std::chrono::system_clock::time_point start) {
    using clock = std::chrono::system_clock;
    const bool q5 = flags & 1;
    mcpm_system_desc* sys[2] = {&cfg.system, &cfg.system2};
    const char* bname[2] = {cfg.box, cfg.box2};
    const int n0[2] = {cfg.n_particles, cfg.n_particles2};
    for (int b = 0; b < 2; b++) {
        if (sys[b]->geometry < 0) return host_fail(MCPM_ERR_UNSUPPORTED, "unknown geometry");
        if (sys[b]->pair_kind != MCPM_PAIR_LJ) return host_fail(MCPM_ERR_UNSUPPORTED, "two boxes: the Lennard-Jones fluid only");
    }
    if (cfg.system.n_vol != 0) return host_fail(MCPM_ERR_UNSUPPORTED, "volume moves between two boxes (the Gibbs branches of MC::ChangeVolume) are out of scope");
    if (cfg.sweep_points > 0 || cfg.sample_rdf) return host_fail(MCPM_ERR_UNSUPPORTED, "two boxes: no chemical-potential sweep and no RDF sampling");
    print_params(cfg, log);
    for (int b = 0; b < 2; b++)
        if (sys[b]->geometry == MCPM_GEOM_BULK && sys[b]->width[2] < 2 * sys[b]->rcut) {   // src/print.cpp:96-103
            std::cout << "Error: Box size must be larger than twice the cut-off radius." << std::endl;
            std::cout << "Box: " << bname[b] << std::endl << std::endl;
            mcpm::set_last_error("box size must be larger than twice the cut-off radius");
            return MCPM_ERR_ARG;
        }
    const int n_sets = (int)cfg.n_sets, n_equil = (int)cfg.n_equil_sets, print_every = std::max(1, (int)cfg.print_every);
    const long long steps_per_set = (long long)cfg.steps_per_set;
    const int replicas = std::max(1, cfg.replicas), n_chains = 2 * replicas;
    const int total = n0[0] + n0[1];
    const int capacity = cfg.capacity > 0 ? cfg.capacity : std::max(1024, ((total + 64 + 31) / 32) * 32);   // a box can at most hold every particle of the system
    const std::string base = out_dir ? std::string(out_dir) : std::string(".");
    std::vector<Outputs> outs(n_chains);
    for (int r = 0; r < replicas; r++) {
        std::string project = cfg.project;
        if (replicas > 1) project += "_r" + std::to_string(r);
        for (int b = 0; b < 2; b++) outs[2 * r + b] = output_directory(base, project, bname[b]);
    }
    int ndev = 0;
    if (mcpm_device_count(&ndev) != MCPM_OK || ndev < 1) return host_fail(MCPM_ERR_CUDA, std::string("no CUDA device: ") + mcpm_last_error());
    ContextSet guard;
    guard.ctx.assign(1, nullptr);
    mcpm_ctx*& ctx = guard.ctx[0];
    int rc = mcpm_create(0, n_chains, capacity, &ctx);
    if (rc) return host_fail(rc, mcpm_last_error());
    auto set_systems = [&](bool moves_only) -> int {
        for (int c = 0; c < n_chains; c++) {
            mcpm_system_desc m = *sys[c & 1];
            m.n_equil_sets = moves_only ? 0 : n_equil;
            if (moves_only) { m.n_swap = 0; m.n_vol = 0; m.n_disp = 1; }    // MinimizeEnergy: MoveParticle only, no dr adaptation
            int r2 = mcpm_set_system(ctx, c, &m);
            if (r2) return r2;
        }
        return MCPM_OK;
    };
    rc = set_systems(false);
    if (rc) return host_fail(rc, mcpm_last_error());
    std::vector<unsigned int> ids(n_chains);
    for (int c = 0; c < n_chains; c++) ids[c] = (unsigned int)(c / 2);       // one stream per system: box 0's id
    rc = mcpm_set_stream_ids(ctx, ids.data());
    for (int r = 0; r < replicas && !rc; r++) rc = mcpm_link_boxes(ctx, 2 * r, 2 * r + 1);
    if (rc) return host_fail(rc, mcpm_last_error());

    // MC::InitialConfig, src/moves.cpp:56-67: box after box, three draws per particle
    std::random_device rd;
    const unsigned long long seed = cfg.has_seed ? cfg.seed : rd();
    std::vector<std::vector<double>> X(n_chains), Y(n_chains), Z(n_chains);
    for (int r = 0; r < replicas; r++) {
        RefRandom rnd(seed + 7919ull * (unsigned long long)r);
        for (int b = 0; b < 2; b++) {
            const int c = 2 * r + b;
            X[c].resize(capacity); Y[c].resize(capacity); Z[c].resize(capacity);
            for (int k = 0; k < n0[b]; k++) insert_particle(*sys[b], rnd, X[c][k], Y[c][k], Z[c][k]);
            rc = mcpm_set_positions(ctx, c, n0[b], X[c].data(), Y[c].data(), Z[c].data());
            if (rc) return host_fail(rc, mcpm_last_error());
        }
    }
    log << "Initial configuration set." << std::endl << std::endl;
    for (int c = 0; c < n_chains; c++) print_trajectory(outs[c], cfg, *sys[c & 1], 0, sys[c & 1]->dr, n0[c & 1], X[c].data(), Y[c].data(), Z[c].data());

    std::vector<mcpm_chain_stats> stats(n_chains);
    // MC::MinimizeEnergy, src/energy.cpp:13-34: for every non-empty box, BoxEnergy, 200 * N_box calls of MoveParticle — which picks ITS box
    // among both, src/moves.cpp:127-137 — and BoxEnergy again
    log << "Starting minimization..." << std::endl;
    auto t0 = clock::now();
    rc = set_systems(true);
    if (rc) return host_fail(rc, mcpm_last_error());
    for (int b = 0; b < 2; b++) {
        if (n0[b] <= 0) continue;
        mcpm_box_energy_t be;
        rc = mcpm_box_energy(ctx, b, &be, nullptr, nullptr);
        if (rc) return host_fail(rc, mcpm_last_error());
        log << "Minimizing initial configuration of box \"" << bname[b] << "\"..." << std::endl;
        log << "\tEnergy/part. of box \"" << bname[b] << "\" before minimization: " << be.energy / n0[b] << " K" << std::endl;
        mcpm_run_params mp;
        memset(&mp, 0, sizeof(mp));
        mp.first_set = 1; mp.n_sets = 1; mp.steps_per_set = 200LL * n0[b]; mp.rng_mode = MCPM_RNG_PHILOX;
        mp.seed = seed ^ (0x6d696e696d697a65ull + (unsigned long long)b); mp.check_energy = 2; mp.no_sampling = 1;
        rc = mcpm_run(ctx, &mp, stats.data());
        if (rc) return host_fail(rc, mcpm_last_error());
        log << "\tEnergy/part. of box \"" << bname[b] << "\" after minimization (" << n0[b] * 200 << " moves): ";
        log << stats[b].energy_check / std::max(1, stats[b].n) << " K" << std::endl;
    }
    rc = set_systems(false);                                             // the real move weights; same energetics: positions and energies stay
    for (int c = 0; c < n_chains && !rc; c++) {
        rc = mcpm_set_dr(ctx, c, sys[c & 1]->dr);
        if (!rc) rc = mcpm_set_step_counter(ctx, c, 0);
    }
    if (!rc) rc = mcpm_reset_accumulators(ctx);
    if (rc) return host_fail(rc, mcpm_last_error());
    std::vector<int> nn(n_chains, 0);
    auto download = [&]() -> int {
        for (int c = 0; c < n_chains; c++) {
            int r2 = mcpm_get_positions(ctx, c, &nn[c], X[c].data(), Y[c].data(), Z[c].data());
            if (r2) return r2;
        }
        return MCPM_OK;
    };
    rc = download();
    if (rc) return host_fail(rc, mcpm_last_error());
    auto t1 = clock::now();
    log << "Minimization finished" << std::endl;
    log << "Elapsed time of minimization: " << std::chrono::duration<double>(t1 - t0).count() << " s" << std::endl << std::endl;
    for (int c = 0; c < n_chains; c++) {
        print_trajectory(outs[c], cfg, *sys[c & 1], 0, sys[c & 1]->dr, nn[c], X[c].data(), Y[c].data(), Z[c].data());
        print_log(outs[c], cfg, *sys[c & 1], 0, nullptr, q5);
    }

    // the set loop, src/main.cpp:63-84: one mcpm_run per print interval
    auto ts = clock::now();
    for (int set = 1; set <= n_sets;) {
        const int next_print = ((set + print_every - 1) / print_every) * print_every;
        const int last = std::min(n_sets, next_print);
        mcpm_run_params rp;
        memset(&rp, 0, sizeof(rp));
        rp.first_set = set; rp.n_sets = last - set + 1; rp.steps_per_set = steps_per_set; rp.rng_mode = MCPM_RNG_PHILOX; rp.seed = seed;
        rp.check_energy = 2;
        rc = mcpm_run(ctx, &rp, stats.data());
        if (rc) return host_fail(rc, mcpm_last_error());
        if (last % print_every == 0) {
            rc = download();
            if (rc) return host_fail(rc, mcpm_last_error());
            print_stats_header(cfg, (size_t)last, log);
            for (int c = 0; c < n_chains; c++) {
                mcpm_chain_stats shown = stats[c];
                shown.dr = shown.dr_used;                                   // printed BEFORE AdjustMCMoves, src/main.cpp:77-83
                if (c < 2) print_stats_box(cfg, bname[c], *sys[c], (size_t)last, &shown, log);
                print_trajectory(outs[c], cfg, *sys[c & 1], (size_t)last, shown.dr, nn[c], X[c].data(), Y[c].data(), Z[c].data());
                // ComputeChemicalPotential from the Widom/BAR sums the transfers of the last set left in both boxes (src/moves.cpp:368-374)
                double mu_ex = 0., mu = 0.;
                const bool have_mu = cfg.system.n_swap > 0;
                if (have_mu) { rc = mcpm_chemical_potential(ctx, c, &mu_ex, &mu); if (rc) return host_fail(rc, mcpm_last_error()); }
                print_log(outs[c], cfg, *sys[c & 1], (size_t)last, &stats[c], q5, have_mu, mu_ex, mu);
            }
            log << std::endl;
        }
        set = last + 1;
    }
    auto te = clock::now();
    std::time_t endTime = clock::to_time_t(te);
    log << std::endl;
    log << "Simulation finished" << std::endl;
    log << "Elapsed time of simulation: " << std::chrono::duration<double>(te - ts).count() << " s" << std::endl;
    log << "Computation finished at " << std::ctime(&endTime) << std::endl;
    log << std::endl;
    (void)start;
    return MCPM_OK;
}

static int simulate(const char* input_path, const char* out_dir, int flags) {
    using clock = std::chrono::system_clock;
    const bool q5 = flags & 1, quiet = flags & 2;
    std::ostream& os = std::cout;
    std::ostringstream sink;
    std::ostream& log = quiet ? static_cast<std::ostream&>(sink) : os;
    auto start = clock::now();
    std::time_t startTime = clock::to_time_t(start);
    log << "Starting simulation on " << std::ctime(&startTime) << std::endl;

    mcpm_sim_config cfg;
    int rc = mcpm_read_input_file(input_path, &cfg);
    if (rc == MCPM_ERR_IO) { std::cout << "Error opening file: " << (input_path ? input_path : "") << std::endl; return rc; }
    if (rc) return host_fail(rc, std::string("could not parse the input file: ") + mcpm_last_error());
    if (cfg.n_boxes < 1 || cfg.n_boxes > 2 || cfg.n_species != 1) return host_fail(MCPM_ERR_UNSUPPORTED, "one or two boxes and exactly one species are supported");
    if (cfg.continue_after_crash)   // the reference restarts from ReadTrajectory/ReadLogFile (src/read.cpp:173-281): out of scope, say so
        return host_fail(MCPM_ERR_UNSUPPORTED, "ContinueAfterCrash (restart from trajectory/log files) is not supported");
    if (cfg.n_boxes == 2) return simulate_two_boxes(cfg, out_dir, flags, log, start);
    mcpm_system_desc& sys = cfg.system;
    if (sys.geometry < 0) return host_fail(MCPM_ERR_UNSUPPORTED, "unknown geometry");
    if (sys.pair_kind > MCPM_PAIR_HS) return host_fail(MCPM_ERR_UNSUPPORTED, "EAM potentials are out of scope");
    if (sys.n_vol != 0 && !(sys.pressure > 0)) return host_fail(MCPM_ERR_UNSUPPORTED, "volume moves without ExternalPressure are the two-box Gibbs ensemble: out of scope");
    if (sys.n_vol != 0 && (sys.geometry != MCPM_GEOM_BULK || sys.n_swap != 0)) return host_fail(MCPM_ERR_UNSUPPORTED, "volume moves: bulk NPT only");
    print_params(cfg, log);
    if (sys.geometry == MCPM_GEOM_BULK && sys.width[2] < 2 * sys.rcut) {   // src/print.cpp:96-103
        std::cout << "Error: Box size must be larger than twice the cut-off radius." << std::endl;
        std::cout << "Box: " << cfg.box << std::endl << std::endl;
        mcpm::set_last_error("box size must be larger than twice the cut-off radius");
        return MCPM_ERR_ARG;
    }
    const int n_sets = (int)cfg.n_sets, n_equil = (int)cfg.n_equil_sets, print_every = std::max(1, (int)cfg.print_every);
    const long long steps_per_set = (long long)cfg.steps_per_set;
    sys.n_equil_sets = n_equil;

    // state points (isotherm sweep extension) x replicas = independent chains
    const int points = std::max(1, cfg.sweep_points), replicas = std::max(1, cfg.replicas);
    const int n_chains = points * replicas;
    std::vector<mcpm_system_desc> desc(n_chains, sys);
    std::vector<Outputs> outs(n_chains);
    const std::string base = out_dir ? std::string(out_dir) : std::string(".");
    for (int p = 0; p < points; p++)
        for (int r = 0; r < replicas; r++) {
            const int c = p * replicas + r;
            if (cfg.sweep_points > 0) desc[c].mu = points > 1 ? cfg.sweep_mu0 + p * (cfg.sweep_mu1 - cfg.sweep_mu0) / (points - 1) : cfg.sweep_mu0;
            std::string project = cfg.project;
            if (points > 1) project += "_mu" + std::to_string(p);
            if (replicas > 1) project += "_r" + std::to_string(r);
            outs[c] = output_directory(base, project, cfg.box);
        }
    int capacity = cfg.capacity > 0 ? cfg.capacity : std::max(1024, 2 * cfg.n_particles + 512);

    // shard chains over devices: chain c -> device c mod G, one context + host thread per device
    int ndev = 0;
    if (mcpm_device_count(&ndev) != MCPM_OK || ndev < 1) return host_fail(MCPM_ERR_CUDA, std::string("no CUDA device: ") + mcpm_last_error());
    const int G = std::min({std::max(1, cfg.devices), ndev, n_chains});
    std::vector<std::vector<int>> owned(G);
    for (int c = 0; c < n_chains; c++) owned[c % G].push_back(c);
    ContextSet guard;
    guard.ctx.assign(G, nullptr);
    std::vector<mcpm_ctx*>& ctx = guard.ctx;
    for (int g = 0; g < G; g++) {
        rc = mcpm_create(g, (int)owned[g].size(), capacity, &ctx[g]);
        if (rc) return host_fail(rc, mcpm_last_error());
        for (size_t k = 0; k < owned[g].size(); k++) {
            rc = mcpm_set_system(ctx[g], (int)k, &desc[owned[g][k]]);
            if (rc) return host_fail(rc, mcpm_last_error());
        }
        // the Philox stream of a chain is keyed by its GLOBAL id: no two chains of the job share uniforms, and the
        // results do not depend on how many devices the chains are sharded over
        std::vector<unsigned int> ids(owned[g].begin(), owned[g].end());
        rc = mcpm_set_stream_ids(ctx[g], ids.data());
        if (rc) return host_fail(rc, mcpm_last_error());
    }

    // MC::InitialConfig, src/moves.cpp:56-67: three draws per particle from the reference's Random()
    std::random_device rd;
    const unsigned long long seed = cfg.has_seed ? cfg.seed : rd();
    std::vector<std::vector<double>> X(n_chains), Y(n_chains), Z(n_chains);
    for (int c = 0; c < n_chains; c++) {
        RefRandom rnd(seed + 7919ull * (unsigned long long)c);
        X[c].resize(capacity); Y[c].resize(capacity); Z[c].resize(capacity);
        for (int k = 0; k < cfg.n_particles; k++) insert_particle(sys, rnd, X[c][k], Y[c][k], Z[c][k]);
    }
    auto upload = [&]() -> int {
        for (int g = 0; g < G; g++)
            for (size_t k = 0; k < owned[g].size(); k++) {
                const int c = owned[g][k];
                int r2 = mcpm_set_positions(ctx[g], (int)k, cfg.n_particles, X[c].data(), Y[c].data(), Z[c].data());
                if (r2) return r2;
            }
        return MCPM_OK;
    };
    rc = upload();
    if (rc) return host_fail(rc, mcpm_last_error());
    log << "Initial configuration set." << std::endl << std::endl;
    for (int c = 0; c < n_chains; c++) print_trajectory(outs[c], cfg, desc[c], 0, sys.dr, cfg.n_particles, X[c].data(), Y[c].data(), Z[c].data());

    std::vector<std::vector<mcpm_chain_stats>> stats(G);
    for (int g = 0; g < G; g++) stats[g].resize(owned[g].size());
    auto run_all = [&](const mcpm_run_params& rp) -> int {   // one host thread per device
        std::vector<std::thread> th;
        std::vector<int> rcs(G, MCPM_OK);
        std::vector<std::string> errs(G);
        for (int g = 0; g < G; g++)
            th.emplace_back([&, g]() { rcs[g] = mcpm_run(ctx[g], &rp, stats[g].data()); if (rcs[g]) errs[g] = mcpm_last_error(); });
        for (auto& t : th) t.join();
        for (int g = 0; g < G; g++) if (rcs[g]) return host_fail(rcs[g], errs[g]);
        return MCPM_OK;
    };
    // global chain 0 lives on device 0, local index 0 (c mod G, c div G)
    auto stats_of = [&](int c) -> const mcpm_chain_stats& { return stats[c % G][c / G]; };

    // MC::MinimizeEnergy, src/energy.cpp:13-34: BoxEnergy; 200*N MoveParticle calls at temperature T; BoxEnergy
    log << "Starting minimization..." << std::endl;
    auto t0 = clock::now();
    if (cfg.n_particles > 0) {
        std::vector<mcpm_box_energy_t> be0(owned[0].size());
        for (int g = 0; g < G; g++) {
            std::vector<mcpm_box_energy_t> be(owned[g].size());
            rc = mcpm_box_energy_all(ctx[g], be.data());
            if (rc) return host_fail(rc, mcpm_last_error());
            if (g == 0) be0 = be;
        }
        log << "Minimizing initial configuration of box \"" << cfg.box << "\"..." << std::endl;
        log << "\tEnergy/part. of box \"" << cfg.box << "\" before minimization: " << be0[0].energy / cfg.n_particles << " K" << std::endl;
        for (int g = 0; g < G; g++)
            for (size_t k = 0; k < owned[g].size(); k++) {
                mcpm_system_desc m = desc[owned[g][k]];
                m.n_swap = 0; m.n_vol = 0; m.n_disp = 1; m.n_equil_sets = 0;       // translations only, no dr adaptation (AdjustMCMoves is commented out there)
                rc = mcpm_set_system(ctx[g], (int)k, &m);
                if (rc) return host_fail(rc, mcpm_last_error());
            }
        mcpm_run_params mp;
        memset(&mp, 0, sizeof(mp));
        mp.first_set = 1; mp.n_sets = 1; mp.steps_per_set = 200LL * cfg.n_particles; mp.rng_mode = MCPM_RNG_PHILOX;
        mp.seed = seed ^ 0x6d696e696d697a65ull; mp.check_energy = 1;
        mp.no_sampling = 1;                                           // MinimizeEnergy calls MoveParticle only: no ComputeWidom after the moves
        rc = run_all(mp);
        if (rc) return rc;
        log << "\tEnergy/part. of box \"" << cfg.box << "\" after minimization (" << cfg.n_particles * 200 << " moves): ";
        log << stats_of(0).energy_check / std::max(1, stats_of(0).n) << " K" << std::endl;
        for (int g = 0; g < G; g++) {
            for (size_t k = 0; k < owned[g].size(); k++) {
                const int c = owned[g][k];
                rc = mcpm_set_system(ctx[g], (int)k, &desc[c]);       // the real move weights; same energetics, so positions and energies stay
                if (!rc) rc = mcpm_set_dr(ctx[g], (int)k, sys.dr);    // (the minimisation never adapts dr)
                if (!rc) rc = mcpm_set_step_counter(ctx[g], (int)k, 0);
                if (rc) return host_fail(rc, mcpm_last_error());
            }
            rc = mcpm_reset_accumulators(ctx[g]);
            if (rc) return host_fail(rc, mcpm_last_error());
        }
        for (int c = 0; c < n_chains; c++) {
            int n = 0;
            rc = mcpm_get_positions(ctx[c % G], c / G, &n, X[c].data(), Y[c].data(), Z[c].data());
            if (rc) return host_fail(rc, mcpm_last_error());
        }
    }
    auto t1 = clock::now();
    log << "Minimization finished" << std::endl;
    log << "Elapsed time of minimization: " << std::chrono::duration<double>(t1 - t0).count() << " s" << std::endl << std::endl;
    for (int c = 0; c < n_chains; c++) {
        print_trajectory(outs[c], cfg, desc[c], 0, sys.dr, cfg.n_particles, X[c].data(), Y[c].data(), Z[c].data());
        print_log(outs[c], cfg, desc[c], 0, nullptr, q5);
    }

    // the set loop, src/main.cpp:63-84: one mcpm_run per print interval
    auto ts = clock::now();
    for (int set = 1; set <= n_sets;) {
        const int next_print = ((set + print_every - 1) / print_every) * print_every;
        const int last = std::min(n_sets, next_print);
        mcpm_run_params rp;
        memset(&rp, 0, sizeof(rp));
        rp.first_set = set; rp.n_sets = last - set + 1; rp.steps_per_set = steps_per_set; rp.rng_mode = MCPM_RNG_PHILOX; rp.seed = seed;
        rp.sample_rdf = cfg.sample_rdf;
        rp.check_energy = 2;   // re-anchor the running energies at every print interval (src/energy.cpp:35-42 in spirit)
        rc = run_all(rp);
        if (rc) return rc;
        if (last % print_every == 0) {
            for (int c = 0; c < n_chains; c++) {
                const int g = c % G, k = c / G;
                int n = 0;
                rc = mcpm_get_positions(ctx[g], k, &n, X[c].data(), Y[c].data(), Z[c].data());
                if (rc) return host_fail(rc, mcpm_last_error());
                // the reference prints BEFORE AdjustMCMoves (src/main.cpp:77-83): report the amplitude the set ran with
                mcpm_chain_stats shown = stats[g][k];
                shown.dr = shown.dr_used;
                if (desc[c].n_vol > 0 && shown.width > 0) desc[c].width[0] = desc[c].width[1] = desc[c].width[2] = shown.width;   // NPT: the box moved
                if (c == 0) print_stats(cfg, desc[c], (size_t)last, &shown, log);
                print_trajectory(outs[c], cfg, desc[c], (size_t)last, shown.dr, n, X[c].data(), Y[c].data(), Z[c].data());
                double mu_ex = 0., mu = 0.;
                const bool have_mu = desc[c].n_swap == 0 && last > n_equil;   // Widom sampling ran (src/moves.cpp:415, src/main.cpp:72)
                if (have_mu) { rc = mcpm_chemical_potential(ctx[g], k, &mu_ex, &mu); if (rc) return host_fail(rc, mcpm_last_error()); }
                print_log(outs[c], cfg, desc[c], (size_t)last, &stats[g][k], q5, have_mu, mu_ex, mu);
            }
        }
        set = last + 1;
    }
    auto te = clock::now();

    // MC::PrintRDF, src/print.cpp:127-162 (normalisation reproduced literally, including its use of ALL sets)
    if (cfg.sample_rdf) {
        const double NB = 100.;
        for (int c = 0; c < n_chains; c++) {
            const int g = c % G, k = c / G;
            std::vector<double> hist(101, 0.0);
            rc = mcpm_get_rdf(ctx[g], k, hist.data());
            if (rc) return host_fail(rc, mcpm_last_error());
            const double deltaR = desc[c].rcut / NB;
            const int nParts = stats[g][k].n;
            const double rho = nParts / box_volume(desc[c]);
            const double constant = (4 * M_PI * rho / 3);
            std::ofstream f(outs[c].box_dir + "/rdf_" + cfg.fluid + "-" + cfg.fluid + ".dat");
            f << "nBins: " << 100 << std::endl;
            f << "r(AA)\tg(r)" << std::endl;
            for (int bin = 1; bin <= 100; bin++) {
                double gofr = hist[bin] / (1. * nParts * cfg.steps_per_set * cfg.n_sets);
                const double lowR = bin * deltaR, highR = lowR + deltaR;
                gofr /= constant * (ipow(highR, 3) - ipow(lowR, 3));
                f << (bin + 0.5) * deltaR << "\t" << gofr << std::endl;
            }
        }
    }
    // final gather: per-state-point averages over replicas and production steps (the reference computes none)
    if (n_chains > 1) {
        std::ofstream iso(base + "/" + std::string(cfg.project) + "_isotherm.dat");
        iso << "mu(K)\t<N>\tsigma_N\t<E>(K)\treplicas\tsamples\n" << std::setprecision(10);
        for (int p = 0; p < points; p++) {
            double sn = 0, sn2 = 0, se = 0; long long ns = 0;
            for (int r = 0; r < replicas; r++) {
                const mcpm_chain_stats& st = stats_of(p * replicas + r);
                sn += st.sum_n; sn2 += st.sum_n2; se += st.sum_e; ns += st.samples;
            }
            const double mean = ns ? sn / ns : 0, var = ns ? std::max(0.0, sn2 / ns - mean * mean) : 0;
            iso << desc[p * replicas].mu << "\t" << mean << "\t" << std::sqrt(var) << "\t" << (ns ? se / ns : 0) << "\t" << replicas << "\t" << ns << "\n";
        }
    }
    std::time_t endTime = clock::to_time_t(te);
    log << std::endl << "Simulation finished" << std::endl;
    log << "Elapsed time of simulation: " << std::chrono::duration<double>(te - ts).count() << " s" << std::endl;
    log << "Computation finished at " << std::ctime(&endTime) << std::endl << std::endl;
    return MCPM_OK;
}
