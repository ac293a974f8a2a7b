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
    mcpm_system_desc& d = cfg->system;
    d.temperature = -1.;
    std::string fluid_name, box_name, geometry = "bulk", ff_pot, sf_pot;   // defaults of MC::MC(), src/MC.cpp:56
    double size = 0.;
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
        else if (k == "boxname") { cfg->n_boxes++; box_name = c[1]; }
        else if (k == "geometry") geometry = c[1];
        else if (k == "surfacedensity") d.rho_s = std::stod(c[1]);
        else if (k == "nlayersperwall") d.n_layers = std::stoi(c[1]);
        else if (k == "distancebetweenlayers") d.delta = std::stod(c[1]);
        else if (k == "size") size = std::stod(c[1]);
        else if (k == "fixvolume") {}
        else if (c[2] == "vdwpotential" || c[2] == "sigma" || c[2] == "epsilon" || c[2] == "rcut" || c[2] == "numberofparticles") {
            // addressing rule of src/read.cpp:82-111: "<fluid> <fluid> …" = pair term, "<box> <fluid> …" = wall term
            const bool a_fluid = (c[0] == fluid_name && !fluid_name.empty()), b_fluid = (c[1] == fluid_name && !fluid_name.empty());
            const bool a_box = (c[0] == box_name && !box_name.empty());
            if (c[2] == "vdwpotential") { if (a_fluid && b_fluid) ff_pot = c[3]; if (a_box && b_fluid) sf_pot = c[3]; }
            else if (c[2] == "sigma") { if (a_fluid && b_fluid) d.sig_ff = std::stod(c[3]); if (a_box && b_fluid) d.sig_sf = std::stod(c[3]); }
            else if (c[2] == "epsilon") { if (a_fluid && b_fluid) d.eps_ff = std::stod(c[3]); if (a_box && b_fluid) d.eps_sf = std::stod(c[3]); }
            else if (c[2] == "rcut") { if (a_fluid && b_fluid) d.rcut = std::stod(c[3]); }
            else if (c[2] == "numberofparticles") { if (a_box && b_fluid) cfg->n_particles = std::stoi(c[3]); }
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
    copy_name(cfg->box, box_name);
    // post-processing, src/read.cpp:122-168
    const double maxRcut = d.rcut;
    d.width[2] = size;
    if (geometry == "slit") { d.geometry = MCPM_GEOM_SLIT; d.width[0] = d.width[1] = 2 * maxRcut; }   // :134
    else if (geometry == "bulk") { d.geometry = MCPM_GEOM_BULK; d.width[0] = d.width[1] = d.width[2]; } // :135
    else if (geometry == "sphere") { d.geometry = MCPM_GEOM_SPHERE; d.width[0] = d.width[1] = d.width[2]; }   // :130
    else if (geometry == "cylinder") { d.geometry = MCPM_GEOM_CYLINDER; d.width[0] = 2 * maxRcut; d.width[1] = d.width[2]; }   // :131-133
    else d.geometry = -1;
    // wall dispatch of src/energy.cpp:69-77: ("lj", slit) -> SlitLJ_Pot, ("lj", sphere) -> SphericalLJ_Pot, ("lj10-4" | "steele10-4-3", cylinder)
    // -> the two cylindrical walls; anything else matches nothing, e.g. "steele10-4-3" + slit
    d.wall_kind = (sf_pot == "lj" && geometry == "slit") ? MCPM_WALL_SLIT_LJ : (sf_pot == "lj" && geometry == "sphere") ? MCPM_WALL_SPHERE_LJ :
                  (sf_pot == "lj10-4" && geometry == "cylinder") ? MCPM_WALL_CYL_LJ104 : (sf_pot == "steele10-4-3" && geometry == "cylinder") ? MCPM_WALL_CYL_STEELE : MCPM_WALL_NONE;
    // pair dispatch of src/energy.cpp:79-101: "lj", "hs"; the EAM potentials are recognised and refused; anything else gives no pair energy
    d.pair_kind = ff_pot == "hs" ? MCPM_PAIR_HS : (ff_pot.rfind("eam_", 0) == 0 ? 2 : MCPM_PAIR_LJ);
    if (ff_pot != "lj" && ff_pot != "hs" && d.pair_kind == MCPM_PAIR_LJ) d.eps_ff = 0.0;      // no potential matched: an ideal gas
    d.pressure = cfg->pressure;
    d.dv = 1e-3;               // src/MC.cpp:33
    d.dr = 0.1 * d.width[2];   // :168
    d.n_equil_sets = (int)cfg->n_equil_sets;
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
    os << "System parameters:" << std::endl;
    os << "\tNum. of particles: " << c.n_particles << std::endl;
    os << "\tVolume (AA^3): " << box_volume(d) - 1. << std::endl;   // thermoSys.volume starts at -1 and is += (src/MC.cpp:37, src/read.cpp:137)
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
    os << "Box 0: " << c.box << std::endl;
    os << "\tGeometry: " << geometry_name(d.geometry) << std::endl;
    os << "\tInitial width (AA): " << d.width[2] << std::endl;
    if (d.rho_s > 0) os << "\tSurface density (AA^-2): " << d.rho_s << std::endl;
    if (d.n_layers > 1) os << "\tLayers per wall: " << d.n_layers << std::endl;
    if (d.delta > 0) os << "\tDistance between layers (AA): " << d.delta << std::endl;
    os << "\tInitial number of particles: " << c.n_particles << std::endl;
    os << std::endl;
    os << "Box-fluid interaction parameters:" << std::endl;
    os << c.box << "-" << c.fluid << std::endl;
    if (d.wall_kind == MCPM_WALL_SLIT_LJ || d.wall_kind == MCPM_WALL_SPHERE_LJ) os << "\tVdW potential: lj" << std::endl;
    else if (d.wall_kind == MCPM_WALL_CYL_LJ104) os << "\tVdW potential: lj10-4" << std::endl;
    else if (d.wall_kind == MCPM_WALL_CYL_STEELE) os << "\tVdW potential: steele10-4-3" << std::endl;
    if (d.sig_sf > 0) os << "\tsigma (AA): " << d.sig_sf << std::endl;
    if (d.eps_sf > 0) os << "\tepsilon (K): " << d.eps_sf << std::endl;
    os << "\tInitial number of particles: " << c.n_particles << std::endl;
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

// MC::PrintStats, src/print.cpp:234-264
void print_stats(const mcpm_sim_config& c, const mcpm_system_desc& d, size_t set, const mcpm_chain_stats* st, std::ostream& os) {
    os << std::fixed << std::setprecision(7);
    if (set < c.n_equil_sets) os << "Equilibrium set: " << set << std::endl;
    else os << "Set: " << set << std::endl;
    os << "Box " << c.box << ":" << std::endl;
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
    if (cfg.n_boxes != 1 || cfg.n_species != 1) return host_fail(MCPM_ERR_UNSUPPORTED, "exactly one box and one species are supported");
    if (cfg.continue_after_crash)   // the reference restarts from ReadTrajectory/ReadLogFile (src/read.cpp:173-281): out of scope, say so
        return host_fail(MCPM_ERR_UNSUPPORTED, "ContinueAfterCrash (restart from trajectory/log files) is not supported");
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
