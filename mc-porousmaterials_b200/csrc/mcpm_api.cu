// C-ABI implementation (include/mcpm.h): context, state transfer, launches.  No torch types, no CPU
// fallback: every compute entry point needs a CUDA device and fails with MCPM_ERR_CUDA otherwise.
#include <algorithm>
#include <chrono>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "mcpm_kernels.h"
#include "mcpm_types.h"
#ifdef MCPM_EXPERIMENTS
#include "experiments/mcpm_experiments.h"
#endif

using namespace mcpm;

static thread_local std::string g_err;

// The host driver (host/mcpm_host.cpp) reports its failures through the same per-thread message mcpm_last_error() returns.
namespace mcpm { void set_last_error(const char* msg) { g_err = msg ? msg : ""; } }

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(MCPM_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct mcpm_ctx {
    int device = 0, n_chains = 0, cap = 0, num_sms = 0, max_smem = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, tm0 = nullptr, tm1 = nullptr;
    ChainParams* d_params = nullptr;
    ChainState* d_state = nullptr;
    double *d_x = nullptr, *d_y = nullptr, *d_z = nullptr;
    int* d_order = nullptr;
    double* d_scratch = nullptr;   // partial sums of K1/K3
    size_t scratch_doubles = 0;
    unsigned int* d_counters = nullptr;
    double* d_out = nullptr;       // K1: 8 doubles; K3: 4 per chain
    double *d_pp = nullptr, *d_pw = nullptr;  // K3 per-particle outputs, one chain's worth
    double* d_rdf = nullptr;       // [n_chains][MCPM_RDF_BINS]
    // K1 (mcpm_k1.cu): request / result buffers, allocated on first use.  h_k1 is MAPPED pinned host memory: the kernel writes the
    // result records and then a sequence flag straight into it, the host spins on the flag (no copy, no stream synchronisation).
    K1Request* d_k1_reqs = nullptr;
    K1Request* h_k1_reqs = nullptr;            // pinned staging of a batch's requests
    double* d_k1_out = nullptr;
    double* d_k1_scratch = nullptr;
    unsigned int* d_k1_counters = nullptr;     // MCPM_K1_MAX_BATCH/4 group counters + 1 launch counter
    double* h_k1 = nullptr;                    // [MCPM_K1_MAX_BATCH][8] records + the flag word
    double* d_k1_alias = nullptr;              // device address of h_k1
    unsigned long long k1_seq = 0;
    // resident K1 (k1_server): mailbox in mapped pinned host memory, its own stream; armed for one chain by mcpm_k1_resident
    cudaStream_t srv_stream = nullptr;
    K1Mail* h_mail = nullptr; K1Mail* d_mail = nullptr;
    K1Reply* h_reply = nullptr; K1Reply* d_reply = nullptr;
    double* h_box = nullptr; double* d_box = nullptr;   // per-particle results of a BoxEnergy answered by the service: [cap] pair | [cap] wall
    int srv_chain = -1;                        // chain the service is armed for, or -1
    bool srv_running = false, srv_fast = false;
    unsigned long long srv_seq = 0;
    unsigned long long srv_starts = 0, srv_requests = 0;
    // cell-list K3 workspace, allocated on first use
    int *d_cell_of = nullptr, *d_cell_counts = nullptr, *d_cell_starts = nullptr, *d_orig = nullptr;
    double *d_sx = nullptr, *d_sy = nullptr, *d_sz = nullptr;
    int use_cells = 1;             // 0 disables the cell-list path (tests compare both)
    int use_cache = 1;             // per-particle energy cache in K2 (mcpm_use_energy_cache)
    const char* last_kernel = "";
    int *d_kc_items = nullptr, *d_kc_cell_of = nullptr, *d_kc_slot_of = nullptr;   // K2c cell lists (mcpm_cells.cu)
    size_t kc_items_ints = 0;
    int use_dual = 0;              // opt-in: one-warp chains evaluate two candidate steps per round (mcpm_dual.cu); measured slower in round 1
    int use_pool = 0;              // opt-in (mcpm_use_pool): one-warp chains through the persistent pooled kernel k2_pool; warps per CTA below
    int pool_warps = 12;
    int* d_pool_next = nullptr;    // k2_pool queue counters
    int use_spec = -1;             // speculative-moves kernel (mcpm_spec.cu): -1 = automatic (fewer chains than SMs), 0 = never, 1 = whenever possible
    int use_subwarp = 0;           // opt-in: the sub-warp kernel that packs 2 or 4 small chains into one warp (mcpm_use_subwarp)
    int2* d_ctas = nullptr;        // CTA descriptors of the sub-warp kernel, n_chains entries
    int* d_flags = nullptr;        // 2*n_chains ints: particle counts in, in-box flags out (bulk uploads)
    int* h_flags = nullptr;        // pinned mirror
    double* d_ep = nullptr;        // [n_chains][cap], allocated on first use
    double* d_replay = nullptr; size_t replay_doubles = 0;
    unsigned char* d_trace = nullptr; size_t trace_bytes = 0;
    double* h_pin = nullptr;       // pinned staging, 4*n_chains+8 doubles
    std::vector<ChainParams> h_params;
    std::vector<ChainState> h_state;
    bool state_pinned = false;
    std::vector<mcpm_system_desc> h_desc;
    std::vector<char> has_system;
    std::vector<int> peer;           // two-box Gibbs systems (mcpm_link_boxes): the other box of the chain's system, or -1
    std::vector<char> box_id;        // 0 / 1 inside a linked pair
    int2* d_pairs = nullptr;
    std::vector<char> cache_fresh;   // the device cache belongs to the configuration now on the device (no host write since the last K2 launch)
    float last_ms = 0.f;
    unsigned long long launches = 0;
    unsigned long long pos_bytes_h2d = 0, pos_bytes_d2h = 0;   // position bytes moved by the bulk transfers (mcpm_transfer_bytes)
};

// ---- host-side parameter digestion ---------------------------------------------------------------
// Tools::Pow, reference src/tools.h:53-58 (repeated multiply starting from 1.)
static double ipow(double v, int p) { double r = 1.; for (int i = 1; i <= p; i++) r *= v; return r; }

// MC::ComputeVolume, src/read.cpp:26-35
static double box_volume(const mcpm_system_desc& d) {
    if (d.geometry == MCPM_GEOM_SPHERE) return (M_PI / 6.) * ipow(d.width[2], 3);
    if (d.geometry == MCPM_GEOM_CYLINDER) return (M_PI / 4.) * ipow(d.width[2], 2) * d.width[0];
    if (d.geometry == MCPM_GEOM_SLIT) return d.width[0] * d.width[1] * d.width[2];
    return ipow(d.width[2], 3);
}

static int digest(const mcpm_system_desc& d, ChainParams& P) {
    if (d.geometry < MCPM_GEOM_BULK || d.geometry > MCPM_GEOM_SPHERE) return fail(MCPM_ERR_UNSUPPORTED, "geometry %d not supported", d.geometry);
    if (d.wall_kind < MCPM_WALL_NONE || d.wall_kind > MCPM_WALL_CYL_STEELE) return fail(MCPM_ERR_UNSUPPORTED, "wall_kind %d not supported", d.wall_kind);
    if (d.pair_kind != MCPM_PAIR_LJ && d.pair_kind != MCPM_PAIR_HS) return fail(MCPM_ERR_UNSUPPORTED, "pair_kind %d not supported (EAM potentials are out of scope)", d.pair_kind);
    if (d.n_vol < 0) return fail(MCPM_ERR_ARG, "move weights must be non-negative");
    if (d.n_vol > 0) {   // MC::ChangeVolume: the NPT branch of a single bulk box (src/moves.cpp:198-231); its other branches need two boxes
        if (!(d.pressure > 0)) return fail(MCPM_ERR_UNSUPPORTED, "volume moves without an external pressure are the two-box Gibbs ensemble: out of scope");
        if (d.geometry != MCPM_GEOM_BULK) return fail(MCPM_ERR_UNSUPPORTED, "volume moves are supported for the bulk geometry only");
        if (d.n_swap != 0) return fail(MCPM_ERR_UNSUPPORTED, "volume moves together with swap moves are not supported");
        if (!(d.width[0] == d.width[2] && d.width[1] == d.width[2])) return fail(MCPM_ERR_ARG, "a bulk box is cubic (src/read.cpp:135)");
    }
    if (d.n_disp < 0 || d.n_swap < 0 || d.n_disp + d.n_swap + d.n_vol <= 0) return fail(MCPM_ERR_ARG, "move weights must be non-negative and not all zero");
    if (!(d.width[0] > 0 && d.width[1] > 0 && d.width[2] > 0) || !(d.temperature > 0)) return fail(MCPM_ERR_ARG, "widths and temperature must be positive");
    if (d.n_layers < 0) return fail(MCPM_ERR_ARG, "n_layers must be >= 0");
    for (int k = 0; k < 3; k++) { P.L[k] = d.width[k]; P.invL[k] = 1.0 / d.width[k]; }
    P.halfH = d.width[2] * 0.5;
    P.sqr_pore_r = d.width[2] * d.width[2] * 0.25;               // src/energy.cpp:53
    P.T = d.temperature;
    P.sig2 = d.sig_ff * d.sig_ff;
    P.eps4 = d.pair_kind == MCPM_PAIR_HS ? 1.0 : 4. * d.eps_ff;  // hard-sphere sums are energies already (INT_MAX per overlap)
    P.rc2 = d.rcut * d.rcut;
    const double volume = box_volume(d);
    // activity, src/moves.cpp:296-298 (constants src/MC.h:16-18)
    const double kb = 1.380649e-23, na = 6.02214076e23, planck = 6.62607015e-34;
    const double particleMass = d.molar_mass / na * 1e-3;
    double zz = 0.0;
    if (d.molar_mass > 0) {
        const double thermalWL = planck / sqrt(2.0 * M_PI * particleMass * kb * d.temperature) * 1e10;
        zz = exp(d.mu / d.temperature) / ipow(thermalWL, 3);
    }
    if (d.n_swap > 0 && !(zz > 0)) return fail(MCPM_ERR_ARG, "swap moves need molar_mass > 0 and a finite activity");
    P.zzV = zz * volume;
    P.ln_zzV = P.zzV > 0 ? log(P.zzV) : -INFINITY;
    P.beta = 1.0 / d.temperature;
    P.wsum = (double)(d.n_disp + d.n_vol + d.n_swap); P.thr_disp = (double)d.n_disp; P.thr_vol = (double)(d.n_disp + d.n_vol);
    P.wall_pref = 4. * M_PI * d.eps_sf * d.rho_s * d.sig_sf * d.sig_sf;  // src/potentials.cpp:400
    P.sig_sf = d.sig_sf; P.delta = d.delta;
    P.geometry = d.geometry; P.wall_kind = d.wall_kind; P.n_layers = d.n_layers;
    // PBC flags follow the geometry, src/read.cpp:146-162: bulk xyz, slit xy, cylinder x, sphere none
    P.pbc[0] = d.geometry != MCPM_GEOM_SPHERE;
    P.pbc[1] = d.geometry == MCPM_GEOM_BULK || d.geometry == MCPM_GEOM_SLIT;
    P.pbc[2] = d.geometry == MCPM_GEOM_BULK;
    P.n_disp = d.n_disp; P.n_vol = d.n_vol; P.n_swap = d.n_swap; P.n_equil = d.n_equil_sets;
    P.pair_kind = d.pair_kind;
    P.sig_ff = d.sig_ff; P.eps_sf = d.eps_sf; P.rho_s = d.rho_s;
    P.press_k = d.pressure > 0 ? d.pressure / kb * 1e-30 : 0.0;   // src/moves.cpp:198
    P.volume = volume;
    // everything beyond "Lennard-Jones fluid, bulk or slit with the layered 10-4 wall, fixed volume" runs through the general kernels
    P.exotic = (d.geometry == MCPM_GEOM_CYLINDER || d.geometry == MCPM_GEOM_SPHERE || d.wall_kind >= MCPM_WALL_SPHERE_LJ || d.pair_kind == MCPM_PAIR_HS || d.n_vol > 0) ? 1 : 0;
    return MCPM_OK;
}

static int check_chain(const mcpm_ctx* c, int chain) {
    if (!c) return fail(MCPM_ERR_ARG, "null context");
    if (chain < 0 || chain >= c->n_chains) return fail(MCPM_ERR_ARG, "chain %d out of range [0,%d)", chain, c->n_chains);
    return MCPM_OK;
}
static bool in_box(const ChainParams& P, int n, const double* x, const double* y, const double* z) {
    if (P.exotic) return false;   // exotic systems always take the general kernels
    const double* a[3] = {x, y, z};
    for (int k = 0; k < 3; k++) {
        if (!P.pbc[k]) continue;
        for (int j = 0; j < n; j++) if (!(a[k][j] >= 0.0 && a[k][j] <= P.L[k])) return false;
    }
    return true;
}
static int push_state(mcpm_ctx* c, int chain) {
    CU(cudaMemcpyAsync(c->d_state + chain, &c->h_state[chain], sizeof(ChainState), cudaMemcpyHostToDevice, c->stream));
    return MCPM_OK;
}
static int pull_states(mcpm_ctx* c) {
    CU(cudaMemcpyAsync(c->h_state.data(), c->d_state, sizeof(ChainState) * c->n_chains, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return MCPM_OK;
}

// ---- resident K1: quiescing (the service itself lives beside K1 below) ----
// Stop the service (if it runs) and bring the device copy of the chain's record up to date: every API entry point that reads or
// writes device state outside the mailbox calls this first.
static int k1_server_stop(mcpm_ctx* c) {
    if (!c->srv_running) return MCPM_OK;
    c->srv_running = false;
    if (__atomic_load_n(&c->h_reply->alive, __ATOMIC_ACQUIRE)) {
        const unsigned long long s = ++c->srv_seq;
        c->h_mail->what = K1_CMD_STOP;
        c->h_mail->seq_tail = s;
        __atomic_store_n(&c->h_mail->seq, s, __ATOMIC_RELEASE);
    }
    CU(cudaStreamSynchronize(c->srv_stream));            // at most the idle time-out away
    CU(cudaMemcpyAsync(c->d_state + c->srv_chain, &c->h_state[c->srv_chain], sizeof(ChainState), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return MCPM_OK;
}
#define QUIESCE(c) do { if ((c) && (c)->srv_running) { int rq_ = k1_server_stop(c); if (rq_) return rq_; } } while (0)

extern "C" {

int mcpm_version(void) { return MCPM_VERSION; }
const char* mcpm_last_error(void) { return g_err.c_str(); }
int mcpm_device_count(int* count) {
    if (!count) return fail(MCPM_ERR_ARG, "null count");
    CU(cudaGetDeviceCount(count));
    return MCPM_OK;
}

int mcpm_create(int device, int n_chains, int capacity, mcpm_ctx** out) {
    if (!out) return fail(MCPM_ERR_ARG, "null out");
    *out = nullptr;
    if (n_chains <= 0 || capacity <= 0) return fail(MCPM_ERR_ARG, "n_chains and capacity must be positive");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(MCPM_ERR_CUDA, "device %d not available (%d CUDA devices)", device, ndev);
    CU(cudaSetDevice(device));
    mcpm_ctx* c = new mcpm_ctx();
    c->device = device; c->n_chains = n_chains;
    c->cap = (capacity + 31) & ~31;  // slots are owned modulo the CTA size; keep rows 256-byte aligned
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    c->max_smem = (int)prop.sharedMemPerBlockOptin;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    CU(cudaEventCreate(&c->tm0));
    CU(cudaEventCreate(&c->tm1));
    const size_t np = (size_t)n_chains * c->cap;
    CU(cudaMalloc(&c->d_params, sizeof(ChainParams) * n_chains));
    CU(cudaMalloc(&c->d_state, sizeof(ChainState) * n_chains));
    CU(cudaMalloc(&c->d_x, sizeof(double) * np));
    CU(cudaMalloc(&c->d_y, sizeof(double) * np));
    CU(cudaMalloc(&c->d_z, sizeof(double) * np));
    CU(cudaMemsetAsync(c->d_x, 0, sizeof(double) * np, c->stream));
    CU(cudaMemsetAsync(c->d_y, 0, sizeof(double) * np, c->stream));
    CU(cudaMemsetAsync(c->d_z, 0, sizeof(double) * np, c->stream));
    CU(cudaMalloc(&c->d_order, sizeof(int) * n_chains));
    const int tiles = (c->cap + 127) / 128;
    c->scratch_doubles = std::max<size_t>((size_t)2 * tiles * n_chains, 4096);
    CU(cudaMalloc(&c->d_scratch, sizeof(double) * c->scratch_doubles));
    CU(cudaMalloc(&c->d_counters, sizeof(unsigned int) * (n_chains + 1)));
    CU(cudaMemsetAsync(c->d_counters, 0, sizeof(unsigned int) * (n_chains + 1), c->stream));
    CU(cudaMalloc(&c->d_out, sizeof(double) * (4 * (size_t)n_chains + 8)));
    CU(cudaMalloc(&c->d_pp, sizeof(double) * c->cap));
    CU(cudaMalloc(&c->d_pw, sizeof(double) * c->cap));
    CU(cudaMalloc(&c->d_rdf, sizeof(double) * MCPM_RDF_BINS * (size_t)n_chains));
    CU(cudaMemsetAsync(c->d_rdf, 0, sizeof(double) * MCPM_RDF_BINS * (size_t)n_chains, c->stream));
    CU(cudaMallocHost(&c->h_pin, sizeof(double) * (4 * (size_t)n_chains + 8)));
    c->h_params.assign(n_chains, ChainParams());
    ChainState zero;
    memset(&zero, 0, sizeof(zero));
    zero.fast_image = 1;
    c->h_state.assign(n_chains, zero);
    // the records come back after every mcpm_run (2.5 MB for 5120 chains): page-locked, the copy runs at PCIe speed instead of through the
    // driver's bounce buffer (the vector is never resized)
    c->state_pinned = !getenv("MCPM_NO_PIN") && cudaHostRegister(c->h_state.data(), sizeof(ChainState) * (size_t)n_chains, cudaHostRegisterDefault) == cudaSuccess;
    if (!c->state_pinned) cudaGetLastError();
    for (int k = 0; k < n_chains; k++) c->h_state[k].stream_id = (unsigned int)k;
    c->h_desc.assign(n_chains, mcpm_system_desc());
    c->has_system.assign(n_chains, 0);
    c->cache_fresh.assign(n_chains, 0);
    c->peer.assign(n_chains, -1);
    c->box_id.assign(n_chains, 0);
    CU(cudaMemcpyAsync(c->d_state, c->h_state.data(), sizeof(ChainState) * n_chains, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *out = c;
    return MCPM_OK;
}

int mcpm_destroy(mcpm_ctx* c) {
    if (!c) return MCPM_OK;
    cudaSetDevice(c->device);
    if (c->srv_running) k1_server_stop(c);
    if (c->srv_stream) { cudaStreamSynchronize(c->srv_stream); cudaStreamDestroy(c->srv_stream); }
    if (c->h_mail) cudaFreeHost(c->h_mail);
    if (c->state_pinned) cudaHostUnregister(c->h_state.data());
    if (c->h_reply) cudaFreeHost(c->h_reply);
    if (c->h_box) cudaFreeHost(c->h_box);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_params); cudaFree(c->d_state); cudaFree(c->d_x); cudaFree(c->d_y); cudaFree(c->d_z);
    cudaFree(c->d_pool_next); cudaFree(c->d_pairs);
    cudaFree(c->d_k1_reqs); cudaFree(c->d_k1_out); cudaFree(c->d_k1_scratch); cudaFree(c->d_k1_counters);
    if (c->h_k1_reqs) cudaFreeHost(c->h_k1_reqs);
    if (c->h_k1) cudaFreeHost(c->h_k1);
    cudaFree(c->d_order); cudaFree(c->d_scratch); cudaFree(c->d_counters); cudaFree(c->d_out);
    cudaFree(c->d_pp); cudaFree(c->d_pw); cudaFree(c->d_rdf); cudaFree(c->d_ep); cudaFree(c->d_kc_items); cudaFree(c->d_kc_cell_of); cudaFree(c->d_kc_slot_of); cudaFree(c->d_ctas); cudaFree(c->d_flags);
    if (c->h_flags) cudaFreeHost(c->h_flags);
    cudaFree(c->d_cell_of); cudaFree(c->d_cell_counts); cudaFree(c->d_cell_starts); cudaFree(c->d_orig);
    cudaFree(c->d_sx); cudaFree(c->d_sy); cudaFree(c->d_sz); cudaFree(c->d_replay); cudaFree(c->d_trace);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->tm0) cudaEventDestroy(c->tm0);
    if (c->tm1) cudaEventDestroy(c->tm1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return MCPM_OK;
}

int mcpm_n_chains(const mcpm_ctx* c, int* n_chains, int* capacity) {
    if (!c) return fail(MCPM_ERR_ARG, "null context");
    if (n_chains) *n_chains = c->n_chains;
    if (capacity) *capacity = c->cap;
    return MCPM_OK;
}

// Do two digested systems give the same energies for the same configuration?  (mu, T, move weights and n_equil do not enter.)
static bool same_energetics(const ChainParams& a, const ChainParams& b) {
    for (int k = 0; k < 3; k++) if (a.L[k] != b.L[k] || a.pbc[k] != b.pbc[k]) return false;
    return a.sig2 == b.sig2 && a.eps4 == b.eps4 && a.rc2 == b.rc2 && a.wall_pref == b.wall_pref && a.sig_sf == b.sig_sf && a.delta == b.delta &&
           a.geometry == b.geometry && a.wall_kind == b.wall_kind && a.n_layers == b.n_layers && a.pair_kind == b.pair_kind &&
           a.sig_ff == b.sig_ff && a.eps_sf == b.eps_sf && a.rho_s == b.rho_s && a.exotic == b.exotic;
}

// fast_image of every chain from the positions on the device (all coordinates on periodic axes inside [0, L]).
static int refresh_fast_image(mcpm_ctx* c) {
    if (!c->d_flags) { CU(cudaMalloc(&c->d_flags, sizeof(int) * 2 * c->n_chains)); CU(cudaMallocHost(&c->h_flags, sizeof(int) * 2 * c->n_chains)); }
    for (int k = 0; k < c->n_chains; k++) c->h_flags[k] = c->h_state[k].st.n;
    CU(cudaMemcpyAsync(c->d_flags, c->h_flags, sizeof(int) * c->n_chains, cudaMemcpyHostToDevice, c->stream));
    CU(launch_in_box(c->n_chains, c->d_params, c->d_flags, c->d_x, c->d_y, c->d_z, c->cap, c->d_flags + c->n_chains, c->stream));
    c->launches++;
    CU(cudaMemcpyAsync(c->h_flags + c->n_chains, c->d_flags + c->n_chains, sizeof(int) * c->n_chains, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < c->n_chains; k++) c->h_state[k].fast_image = c->h_flags[c->n_chains + k];
    return MCPM_OK;
}

int mcpm_set_system(mcpm_ctx* c, int chain, const mcpm_system_desc* desc) {
    QUIESCE(c);
    if (!c || !desc) return fail(MCPM_ERR_ARG, "null argument");
    if (chain != -1) { int rc = check_chain(c, chain); if (rc) return rc; }
    ChainParams P;
    memset(&P, 0, sizeof(P));
    int rc = digest(*desc, P);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    const int lo = chain == -1 ? 0 : chain, hi = chain == -1 ? c->n_chains : chain + 1;
    bool recheck = false;
    for (int k = lo; k < hi; k++) {
        ChainState& S = c->h_state[k];
        // A chain that already holds a configuration keeps its running energies and energy cache only if the new system
        // gives the same energies (e.g. a mu- or T-only update in a sweep); otherwise both are stale, and a smaller box may
        // have left coordinates outside [0, L], which the fast minimum image must not see.
        if (c->has_system[k] && !same_energetics(c->h_params[k], P)) {
            S.energy_valid = 0; S.cache_valid = 0; S.cache_sum = 0ull;
            c->cache_fresh[k] = 0;
            if (S.st.n > 0) recheck = true;
        }
        // dr: taken from the description the first time and whenever the description's own dr changes; an unchanged
        // desc->dr keeps the amplitude the chain has adapted to (AdjustMCMoves), e.g. across mu updates of a sweep.
        if (P.exotic) S.fast_image = 0;                       // general kernels only
        else if (c->has_system[k] && c->h_params[k].exotic && S.st.n > 0) recheck = true;
        if (!c->has_system[k] || c->h_desc[k].dr != desc->dr) S.st.dr = desc->dr;
        if (!c->has_system[k] || c->h_desc[k].dv != desc->dv) S.st.dv = desc->dv > 0 ? desc->dv : 1e-3;     // Simulation::dv, src/MC.cpp:33
        if (P.n_vol > 0 && (!c->has_system[k] || !same_energetics(c->h_params[k], P))) S.b_valid = 0;
        c->h_params[k] = P; c->h_desc[k] = *desc; c->has_system[k] = 1;
    }
    CU(cudaMemcpyAsync(c->d_params + lo, c->h_params.data() + lo, sizeof(ChainParams) * (hi - lo), cudaMemcpyHostToDevice, c->stream));
    if (recheck) { rc = refresh_fast_image(c); if (rc) return rc; }
    CU(cudaMemcpyAsync(c->d_state + lo, c->h_state.data() + lo, sizeof(ChainState) * (hi - lo), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return MCPM_OK;
}

int mcpm_get_system(const mcpm_ctx* c, int chain, mcpm_system_desc* desc) {
    int rc = check_chain(c, chain); if (rc) return rc;
    if (!desc) return fail(MCPM_ERR_ARG, "null desc");
    if (!c->has_system[chain]) return fail(MCPM_ERR_ARG, "chain %d has no system yet", chain);
    *desc = c->h_desc[chain];
    return MCPM_OK;
}

// Two boxes of ONE system (thermoSys.nBoxes == 2): same fluid, temperature and move weights; fixed volumes (see mcpm_gibbs.cu).
static int check_pair(const mcpm_ctx* c, int a, int b) {
    const mcpm_system_desc &da = c->h_desc[a], &db = c->h_desc[b];
    if (da.temperature != db.temperature || da.eps_ff != db.eps_ff || da.sig_ff != db.sig_ff || da.rcut != db.rcut || da.molar_mass != db.molar_mass ||
        da.pair_kind != db.pair_kind)
        return fail(MCPM_ERR_ARG, "linked boxes %d and %d hold the same fluid at the same temperature (one species, one ThermodynamicSystem)", a, b);
    if (da.n_disp != db.n_disp || da.n_swap != db.n_swap || da.n_vol != db.n_vol || da.n_equil_sets != db.n_equil_sets)
        return fail(MCPM_ERR_ARG, "linked boxes %d and %d share the move weights and the number of equilibration sets (Simulation, src/MC.h)", a, b);
    if (da.n_vol != 0) return fail(MCPM_ERR_UNSUPPORTED, "volume moves between two boxes (the Gibbs branches of MC::ChangeVolume) are out of scope: the stock reference aborts in ChangeVolume");
    if (da.pair_kind != MCPM_PAIR_LJ) return fail(MCPM_ERR_UNSUPPORTED, "linked boxes support the Lennard-Jones fluid only");
    return MCPM_OK;
}
int mcpm_link_boxes(mcpm_ctx* c, int box0, int box1) {
    int rc = check_chain(c, box0); if (rc) return rc;
    rc = check_chain(c, box1); if (rc) return rc;
    if (box0 == box1) return fail(MCPM_ERR_ARG, "a box cannot be linked to itself");
    if (!c->has_system[box0] || !c->has_system[box1]) return fail(MCPM_ERR_ARG, "set the systems of both boxes first");
    if (c->peer[box0] >= 0 || c->peer[box1] >= 0) return fail(MCPM_ERR_ARG, "chain %d is linked already", c->peer[box0] >= 0 ? box0 : box1);
    rc = check_pair(c, box0, box1); if (rc) return rc;
    c->peer[box0] = box1; c->peer[box1] = box0; c->box_id[box0] = 0; c->box_id[box1] = 1;
    return MCPM_OK;
}

int mcpm_set_positions(mcpm_ctx* c, int chain, int n, const double* x, const double* y, const double* z) {
    QUIESCE(c);
    int rc = check_chain(c, chain); if (rc) return rc;
    if (n < 0 || n > c->cap) return fail(MCPM_ERR_CAPACITY, "n=%d exceeds capacity %d", n, c->cap);
    if (n > 0 && (!x || !y || !z)) return fail(MCPM_ERR_ARG, "null coordinate array");
    if (!c->has_system[chain]) return fail(MCPM_ERR_ARG, "set the system of chain %d first", chain);
    CU(cudaSetDevice(c->device));
    const size_t off = (size_t)chain * c->cap;
    if (n > 0) {
        CU(cudaMemcpyAsync(c->d_x + off, x, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->d_y + off, y, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->d_z + off, z, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    }
    ChainState& S = c->h_state[chain];
    c->cache_fresh[chain] = 0;
    S.st.n = n; S.st.overflow = 0;
    S.st.pair = S.st.wall = S.st.energy = 0.0; S.energy_valid = 0;
    S.fast_image = in_box(c->h_params[chain], n, x, y, z) ? 1 : 0;
    rc = push_state(c, chain); if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return MCPM_OK;
}

int mcpm_get_positions(mcpm_ctx* c, int chain, int* n, double* x, double* y, double* z) {
    QUIESCE(c);
    int rc = check_chain(c, chain); if (rc) return rc;
    CU(cudaSetDevice(c->device));
    const int m = c->h_state[chain].st.n;
    if (n) *n = m;
    const size_t off = (size_t)chain * c->cap;
    if (m > 0 && x && y && z) {
        CU(cudaMemcpyAsync(x, c->d_x + off, sizeof(double) * m, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(y, c->d_y + off, sizeof(double) * m, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(z, c->d_z + off, sizeof(double) * m, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return MCPM_OK;
}

// Bulk position transfers move, per chain, only the columns that can hold live particles (plus the free slot an insertion would use).
// Chains are cut into runs of neighbours with similar loadings — an isotherm's chains are ordered by state point — and every run is one
// strided copy as wide as its largest chain: about half the bytes of one copy as wide as the largest chain of the whole context (C2:
// 38 MB instead of 73 MB each way), for ~40 x 3 copy calls.  need[k] = columns of chain k.
struct RowRun { int first, count, cols; };
static std::vector<RowRun> row_runs(const std::vector<int>& need, int cap) {
    std::vector<RowRun> runs;
    const long long call_cost = 16384;                   // columns*rows a copy call is worth (~4 us at 25 GB/s, in doubles)
    RowRun cur{0, 0, 0};
    long long cur_sum = 0;                               // sum of the needs of the current run
    for (int k = 0; k < (int)need.size(); k++) {
        const int w = std::min(cap, (need[k] + 31) & ~31);
        if (cur.count == 0) { cur = RowRun{k, 1, w}; cur_sum = w; continue; }
        const int wide = std::max(cur.cols, w);
        // padding paid if chain k joins the run, against the padding of the run as it is plus a new call
        const long long joined = (long long)wide * (cur.count + 1) - (cur_sum + w);
        const long long apart = (long long)cur.cols * cur.count - cur_sum + call_cost;
        if (joined <= apart) { cur.cols = wide; cur.count++; cur_sum += w; }
        else { runs.push_back(cur); cur = RowRun{k, 1, w}; cur_sum = w; }
    }
    if (cur.count) runs.push_back(cur);
    return runs;
}
static int copy_rows(mcpm_ctx* c, const std::vector<RowRun>& runs, double* dst, const double* src, cudaMemcpyKind kind) {
    const size_t pitch = sizeof(double) * c->cap;
    for (const RowRun& r : runs) {
        const size_t off = (size_t)r.first * c->cap;
        (kind == cudaMemcpyHostToDevice ? c->pos_bytes_h2d : c->pos_bytes_d2h) += sizeof(double) * (unsigned long long)r.cols * (unsigned long long)r.count;
        CU(cudaMemcpy2DAsync(dst + off, pitch, src + off, pitch, sizeof(double) * r.cols, r.count, kind, c->stream));
    }
    return MCPM_OK;
}

int mcpm_set_positions_all(mcpm_ctx* c, const int* n, const double* x, const double* y, const double* z) {
    QUIESCE(c);
    if (!c || !n || !x || !y || !z) return fail(MCPM_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    const size_t np = (size_t)c->n_chains * c->cap;
    for (int k = 0; k < c->n_chains; k++) {
        if (n[k] < 0 || n[k] > c->cap) return fail(MCPM_ERR_CAPACITY, "chain %d: n=%d exceeds capacity %d", k, n[k], c->cap);
        if (!c->has_system[k]) return fail(MCPM_ERR_ARG, "set the system of chain %d first", k);
    }
    (void)np;
    std::vector<int> need(c->n_chains);
    for (int k = 0; k < c->n_chains; k++) need[k] = n[k] + 1;
    const std::vector<RowRun> runs = row_runs(need, c->cap);
    int rcc = copy_rows(c, runs, c->d_x, x, cudaMemcpyHostToDevice); if (rcc) return rcc;
    rcc = copy_rows(c, runs, c->d_y, y, cudaMemcpyHostToDevice); if (rcc) return rcc;
    rcc = copy_rows(c, runs, c->d_z, z, cudaMemcpyHostToDevice); if (rcc) return rcc;
    // in-box test of ~MBs of coordinates: on the device, where they just arrived (a host loop costs as much as the copy)
    if (!c->d_flags) { CU(cudaMalloc(&c->d_flags, sizeof(int) * 2 * c->n_chains)); CU(cudaMallocHost(&c->h_flags, sizeof(int) * 2 * c->n_chains)); }
    memcpy(c->h_flags, n, sizeof(int) * c->n_chains);
    CU(cudaMemcpyAsync(c->d_flags, c->h_flags, sizeof(int) * c->n_chains, cudaMemcpyHostToDevice, c->stream));
    CU(launch_in_box(c->n_chains, c->d_params, c->d_flags, c->d_x, c->d_y, c->d_z, c->cap, c->d_flags + c->n_chains, c->stream));
    c->launches++;
    CU(cudaMemcpyAsync(c->h_flags + c->n_chains, c->d_flags + c->n_chains, sizeof(int) * c->n_chains, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < c->n_chains; k++) {
        ChainState& S = c->h_state[k];
        S.st.n = n[k]; S.st.overflow = 0; S.energy_valid = 0;
        c->cache_fresh[k] = 0;
        S.fast_image = c->h_flags[c->n_chains + k];
    }
    CU(cudaMemcpyAsync(c->d_state, c->h_state.data(), sizeof(ChainState) * c->n_chains, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return MCPM_OK;
}

int mcpm_get_positions_all(mcpm_ctx* c, int* n, double* x, double* y, double* z) {
    QUIESCE(c);
    if (!c || !n || !x || !y || !z) return fail(MCPM_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    std::vector<int> need(c->n_chains);
    for (int k = 0; k < c->n_chains; k++) need[k] = c->h_state[k].st.n + 1;
    const std::vector<RowRun> runs = row_runs(need, c->cap);
    int rcc = copy_rows(c, runs, x, c->d_x, cudaMemcpyDeviceToHost); if (rcc) return rcc;
    rcc = copy_rows(c, runs, y, c->d_y, cudaMemcpyDeviceToHost); if (rcc) return rcc;
    rcc = copy_rows(c, runs, z, c->d_z, cudaMemcpyDeviceToHost); if (rcc) return rcc;
    CU(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < c->n_chains; k++) n[k] = c->h_state[k].st.n;
    return MCPM_OK;
}

// ---- K1 ---------------------------------------------------------------------------------------
static int k1_prepare(mcpm_ctx* c) {
    if (c->h_k1) return MCPM_OK;
    const size_t groups = MCPM_K1_MAX_BATCH / 4 + 1;
    CU(cudaMalloc(&c->d_k1_reqs, sizeof(K1Request) * MCPM_K1_MAX_BATCH));
    CU(cudaMallocHost(&c->h_k1_reqs, sizeof(K1Request) * MCPM_K1_MAX_BATCH));
    CU(cudaMalloc(&c->d_k1_out, sizeof(double) * 8 * MCPM_K1_MAX_BATCH));
    CU(cudaMalloc(&c->d_k1_scratch, sizeof(double) * 8 * (MCPM_K1_MAX_BATCH / 4) * (size_t)k1_slices(c->cap)));
    CU(cudaMalloc(&c->d_k1_counters, sizeof(unsigned int) * groups));
    CU(cudaMemsetAsync(c->d_k1_counters, 0, sizeof(unsigned int) * groups, c->stream));
    CU(cudaHostAlloc(&c->h_k1, sizeof(double) * (8 * MCPM_K1_MAX_BATCH + 2), cudaHostAllocMapped));
    memset(c->h_k1, 0, sizeof(double) * (8 * MCPM_K1_MAX_BATCH + 2));
    CU(cudaHostGetDevicePointer((void**)&c->d_k1_alias, c->h_k1, 0));
    CU(cudaStreamSynchronize(c->stream));
    return MCPM_OK;
}

// `count` requests (<= MCPM_K1_MAX_BATCH) in ONE launch; results are waited for on the mapped flag.
static int k1_launch(mcpm_ctx* c, int chain, int count, const K1Request* host_reqs, bool fast) {
    int rc = k1_prepare(c); if (rc) return rc;
    const int n = c->h_state[chain].st.n;
    const K1Request* dev_reqs = nullptr;
    if (count > 1) {
        memcpy(c->h_k1_reqs, host_reqs, sizeof(K1Request) * count);
        CU(cudaMemcpyAsync(c->d_k1_reqs, c->h_k1_reqs, sizeof(K1Request) * count, cudaMemcpyHostToDevice, c->stream));
        dev_reqs = c->d_k1_reqs;
    }
    const unsigned long long seq = ++c->k1_seq;
    volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(c->h_k1 + 8 * MCPM_K1_MAX_BATCH);
    CU(launch_k1(fast, n, count, c->d_params, c->d_state, c->d_x, c->d_y, c->d_z, chain, c->cap, dev_reqs, host_reqs[0], c->d_k1_scratch,
                 c->d_k1_counters, c->d_k1_out, c->d_k1_alias, reinterpret_cast<unsigned long long*>(c->d_k1_alias + 8 * MCPM_K1_MAX_BATCH), seq,
                 c->d_k1_counters + MCPM_K1_MAX_BATCH / 4, c->stream));
    c->launches++;
    // spin on the flag the kernel raises after its records are in host memory; look at the stream now and then so that a failed
    // launch is reported instead of waited for
    for (unsigned long long spins = 1; *flag != seq; spins++) {
        if ((spins & 0x3fffull) == 0) {
            cudaError_t q = cudaStreamQuery(c->stream);
            if (q == cudaSuccess) break;                       // finished: the flag write has landed by now (checked below)
            if (q != cudaErrorNotReady) return fail(MCPM_ERR_CUDA, "K1 launch failed: %s", cudaGetErrorString(q));
        }
    }
    if (*flag != seq) {
        CU(cudaStreamSynchronize(c->stream));
        if (*flag != seq) return fail(MCPM_ERR_CUDA, "K1 finished without delivering its results");
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return MCPM_OK;
}


// ---- resident K1 ------------------------------------------------------------------------------------
static int k1_server_start(mcpm_ctx* c) {
    const int chain = c->srv_chain;
    if (!c->srv_stream) {
        CU(cudaStreamCreateWithFlags(&c->srv_stream, cudaStreamNonBlocking));
        CU(cudaHostAlloc(&c->h_mail, sizeof(K1Mail), cudaHostAllocMapped));
        CU(cudaHostAlloc(&c->h_reply, sizeof(K1Reply), cudaHostAllocMapped));
        memset(c->h_mail, 0, sizeof(K1Mail)); memset(c->h_reply, 0, sizeof(K1Reply));
        CU(cudaHostGetDevicePointer((void**)&c->d_mail, c->h_mail, 0));
        CU(cudaHostGetDevicePointer((void**)&c->d_reply, c->h_reply, 0));
        CU(cudaHostAlloc(&c->h_box, sizeof(double) * 2 * (size_t)c->cap, cudaHostAllocMapped));
        CU(cudaHostGetDevicePointer((void**)&c->d_box, c->h_box, 0));
    }
    CU(cudaStreamSynchronize(c->srv_stream));            // a previous instance that timed out has left by now
    CU(cudaMemcpyAsync(c->d_state + chain, &c->h_state[chain], sizeof(ChainState), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));                // every earlier write to the chain has landed
    c->srv_fast = c->h_state[chain].fast_image != 0 && !c->h_params[chain].exotic;
    c->h_reply->alive = 1ull;
    c->h_reply->seq = c->srv_seq; c->h_reply->seq_tail = c->srv_seq;
    c->h_mail->seq = c->srv_seq; c->h_mail->seq_tail = c->srv_seq;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    const long long idle = 10000000ll;                   // ~5 ms of SM clocks without a request
    CU(launch_k1_server(c->srv_fast, c->h_params[chain].pbc[2] != 0, c->d_params, c->d_state, c->d_x, c->d_y, c->d_z, chain, c->cap,
                        k1_server_smem_bytes(c->cap, c->max_smem), c->d_mail, c->d_reply, c->d_box, c->srv_seq, idle, c->srv_stream));
    c->launches++; c->srv_starts++;
    c->srv_running = true;
    return MCPM_OK;
}

// Wait until the service has answered command `s` (both copies of the sequence number in the reply line).  false: it left before taking it.
static int k1_server_wait(mcpm_ctx* c, unsigned long long s, bool& answered) {
    answered = false;
    const volatile K1Reply* r = c->h_reply;
    for (unsigned long long spins = 1;; spins++) {
        if (__atomic_load_n(&r->seq, __ATOMIC_ACQUIRE) == s && __atomic_load_n(&r->seq_tail, __ATOMIC_ACQUIRE) == s) { answered = true; return MCPM_OK; }
        if ((spins & 0xfffull) == 0) {
            if (!__atomic_load_n(&r->alive, __ATOMIC_ACQUIRE)) {                  // it left (idle) — before or after this command?
                answered = __atomic_load_n(&r->seq, __ATOMIC_ACQUIRE) == s && __atomic_load_n(&r->seq_tail, __ATOMIC_ACQUIRE) == s;
                return MCPM_OK;
            }
            if ((spins & 0xfffffull) == 0) {
                cudaError_t q = cudaStreamQuery(c->srv_stream);
                if (q != cudaErrorNotReady && q != cudaSuccess) { c->srv_running = false; return fail(MCPM_ERR_CUDA, "resident K1 failed: %s", cudaGetErrorString(q)); }
                if (spins > (1ull << 34)) { c->srv_running = false; return fail(MCPM_ERR_CUDA, "resident K1 does not answer"); }
            }
        }
    }
}
// One command through the mailbox, waited for; restarts the service if it has timed out in the meantime.  rec: 8 doubles or null.
static int k1_server_call(mcpm_ctx* c, int cmd, int index, const double* t, double* rec) {
    for (int attempt = 0; attempt < 3; attempt++) {
        if (!c->srv_running) { int rc = k1_server_start(c); if (rc) return rc; }
        const unsigned long long s = ++c->srv_seq;
        K1Mail* m = c->h_mail;
        m->what = (unsigned long long)cmd | ((t ? 1ull : 0ull) << 8) | ((unsigned long long)(unsigned)index << 32);
        for (int k = 0; k < 3; k++) m->t[k] = t ? t[k] : 0.0;
        m->seq_tail = s;
        __atomic_store_n(&m->seq, s, __ATOMIC_RELEASE);
        c->srv_requests++;
        bool answered = false;
        int rc = k1_server_wait(c, s, answered); if (rc) return rc;
        if (answered) {
            if (rec) {
                const volatile double* r = c->h_reply->r;
                rec[0] = r[0]; rec[1] = 0.0; rec[2] = r[1]; rec[3] = r[2]; rec[4] = r[3]; rec[5] = 0.0; rec[6] = r[4]; rec[7] = r[5];
            }
            return MCPM_OK;
        }
        c->srv_running = false;                          // it left before taking the command: start again and ask again (srv_seq goes back)
        c->srv_seq = s - 1;
        c->srv_requests--;
    }
    return fail(MCPM_ERR_CUDA, "resident K1 keeps timing out");
}
static bool k1_served(const mcpm_ctx* c, int chain) { return c->srv_chain == chain && chain >= 0; }

static bool trial_in_box(const ChainParams& P, const double* t) {
    for (int k = 0; k < 3; k++) if (P.pbc[k] && !(t[k] >= 0.0 && t[k] <= P.L[k])) return false;
    return true;
}
static void k1_record(const double* r, mcpm_particle_energy_t* cur, mcpm_particle_energy_t* moved) {
    if (cur) { cur->pair = r[0]; cur->many_body = 0.0; cur->wall = r[2]; cur->energy = r[3]; }
    if (moved) { moved->pair = r[4]; moved->many_body = 0.0; moved->wall = r[6]; moved->energy = r[7]; }
}

int mcpm_particle_energy(mcpm_ctx* c, int chain, int index, const double* trial, mcpm_particle_energy_t* cur,
                         mcpm_particle_energy_t* moved) {
    int rc = check_chain(c, chain); if (rc) return rc;
    if (!c->has_system[chain]) return fail(MCPM_ERR_ARG, "set the system of chain %d first", chain);
    const int n = c->h_state[chain].st.n;
    if (index < -1 || index >= std::max(n, 1)) return fail(MCPM_ERR_ARG, "index %d out of range (n=%d)", index, n);
    if (index == -1 && !trial) return fail(MCPM_ERR_ARG, "index -1 needs a trial position");
    CU(cudaSetDevice(c->device));
    K1Request rq;
    rq.index = index; rq.has_trial = trial ? 1 : 0;
    for (int k = 0; k < 3; k++) rq.t[k] = trial ? trial[k] : 0.0;
    // the trial position must satisfy the fast-image precondition too
    const bool fast = c->h_state[chain].fast_image != 0 && (!trial || trial_in_box(c->h_params[chain], trial));
    if (k1_served(c, chain)) {
        // a service instantiated for coordinates inside the box (fast minimum image) cannot take a trial position outside it: such a
        // request falls back to a launch; a general service takes anything
        const bool chain_fast = c->h_state[chain].fast_image != 0 && !c->h_params[chain].exotic;
        const bool srv_fast = c->srv_running ? c->srv_fast : chain_fast;
        if (!srv_fast || fast) {
            double rec[8];
            rc = k1_server_call(c, K1_CMD_ENERGY, index, trial, rec); if (rc) return rc;
            k1_record(rec, cur, moved);
            return MCPM_OK;
        }
    }
    QUIESCE(c);
    rc = k1_launch(c, chain, 1, &rq, fast); if (rc) return rc;
    k1_record(c->h_k1, cur, moved);
    return MCPM_OK;
}

int mcpm_k1_resident(mcpm_ctx* c, int chain, int enable) {
    if (!c) return fail(MCPM_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    QUIESCE(c);
    if (!enable) { c->srv_chain = -1; return MCPM_OK; }
    int rc = check_chain(c, chain); if (rc) return rc;
    if (!c->has_system[chain]) return fail(MCPM_ERR_ARG, "set the system of chain %d first", chain);
    c->srv_chain = chain;                                // armed: the first mcpm_particle_energy starts the service
    return MCPM_OK;
}
int mcpm_k1_resident_stats(const mcpm_ctx* c, unsigned long long* starts, unsigned long long* requests) {
    if (!c) return fail(MCPM_ERR_ARG, "null context");
    if (starts) *starts = c->srv_starts;
    if (requests) *requests = c->srv_requests;
    return MCPM_OK;
}

int mcpm_particle_energy_batch(mcpm_ctx* c, int chain, int count, const int* index, const double* trial, mcpm_particle_energy_t* cur,
                               mcpm_particle_energy_t* moved) {
    QUIESCE(c);
    int rc = check_chain(c, chain); if (rc) return rc;
    if (!c->has_system[chain]) return fail(MCPM_ERR_ARG, "set the system of chain %d first", chain);
    if (count < 0 || (count > 0 && !index)) return fail(MCPM_ERR_ARG, "bad request list");
    const int n = c->h_state[chain].st.n;
    bool fast = c->h_state[chain].fast_image != 0;
    for (int r = 0; r < count; r++) {
        if (index[r] < -1 || index[r] >= std::max(n, 1)) return fail(MCPM_ERR_ARG, "request %d: index %d out of range (n=%d)", r, index[r], n);
        if (index[r] == -1 && !trial) return fail(MCPM_ERR_ARG, "request %d: index -1 needs a trial position", r);
        if (trial && !trial_in_box(c->h_params[chain], trial + 3 * r)) fast = false;
    }
    CU(cudaSetDevice(c->device));
    std::vector<K1Request> rq(std::min(count, MCPM_K1_MAX_BATCH));
    for (int r0 = 0; r0 < count; r0 += MCPM_K1_MAX_BATCH) {
        const int m = std::min(MCPM_K1_MAX_BATCH, count - r0);
        for (int r = 0; r < m; r++) {
            rq[r].index = index[r0 + r]; rq[r].has_trial = trial ? 1 : 0;
            for (int k = 0; k < 3; k++) rq[r].t[k] = trial ? trial[3 * (r0 + r) + k] : 0.0;
        }
        rc = k1_launch(c, chain, m, rq.data(), fast); if (rc) return rc;
        for (int r = 0; r < m; r++) k1_record(c->h_k1 + 8 * r, cur ? cur + r0 + r : nullptr, moved ? moved + r0 + r : nullptr);
    }
    return MCPM_OK;
}

// ---- K3 ---------------------------------------------------------------------------------------
// The cell-list K3 applies when the box holds >= 3 cut-off lengths along every periodic axis, all coordinates lie
// inside the box and the chain is large enough for the list to pay for itself (config C4; never for reference-format
// slit inputs, whose lateral size is 2*rcut, quirk Q10).
static bool cell_list_ok(const mcpm_ctx* c, int chain) {
    const ChainParams& P = c->h_params[chain];
    const ChainState& S = c->h_state[chain];
    if (!c->use_cells || !S.fast_image || S.st.n < 8192) return false;   // measured: 6.7x at N = 20 000, break-even near N = 4 000
    const double rc = sqrt(P.rc2);
    for (int k = 0; k < 3; k++) if (P.pbc[k] && (int)(P.L[k] / rc) < 3) return false;
    return true;
}

static int box_energy_range(mcpm_ctx* c, int chain0, int count, double* per_pair_dev, double* per_wall_dev) {
    int maxn = 1;
    bool cells = true;
    for (int k = chain0; k < chain0 + count; k++) {
        if (!c->has_system[k]) return fail(MCPM_ERR_ARG, "set the system of chain %d first", k);
        maxn = std::max(maxn, c->h_state[k].st.n);
        cells = cells && cell_list_ok(c, k);
    }
    const int tiles = (maxn + 127) / 128;
    if ((size_t)2 * tiles * count > c->scratch_doubles) return fail(MCPM_ERR_ARG, "scratch too small");
    CU(cudaEventRecord(c->ev0, c->stream));
    if (cells) {
        if (!c->d_cell_of) {
            const size_t np = (size_t)c->n_chains * c->cap, nc = k3c_cells_per_chain();
            CU(cudaMalloc(&c->d_cell_of, sizeof(int) * np)); CU(cudaMalloc(&c->d_orig, sizeof(int) * np));
            CU(cudaMalloc(&c->d_cell_counts, sizeof(int) * nc * c->n_chains)); CU(cudaMalloc(&c->d_cell_starts, sizeof(int) * (nc + 1) * c->n_chains));
            CU(cudaMalloc(&c->d_sx, sizeof(double) * np)); CU(cudaMalloc(&c->d_sy, sizeof(double) * np)); CU(cudaMalloc(&c->d_sz, sizeof(double) * np));
        }
        CU(launch_k3_cells(maxn, count, chain0, c->d_params, c->d_state, c->d_x, c->d_y, c->d_z, c->cap, c->d_cell_of, c->d_cell_counts,
                           c->d_cell_starts, c->d_sx, c->d_sy, c->d_sz, c->d_orig, per_pair_dev, per_wall_dev, c->d_scratch, c->d_counters,
                           c->d_out, c->stream));
        c->launches += 3;
    } else
    CU(launch_k3(tiles, count, chain0, c->d_params, c->d_state, c->d_x, c->d_y, c->d_z, c->cap, per_pair_dev, per_wall_dev,
                 c->d_scratch, c->d_counters, c->d_out, c->stream));
    CU(cudaEventRecord(c->ev1, c->stream));
    c->launches += 2;
    CU(cudaMemcpyAsync(c->h_pin, c->d_out, sizeof(double) * 4 * count, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
    for (int k = 0; k < count; k++) {
        mcpm_chain_stats& st = c->h_state[chain0 + k].st;
        st.pair = st.pair_check = c->h_pin[4 * k]; st.wall = st.wall_check = c->h_pin[4 * k + 2];
        st.energy = st.energy_check = c->h_pin[4 * k + 3];
        c->h_state[chain0 + k].energy_valid = 1;
    }
    return MCPM_OK;
}

int mcpm_box_energy(mcpm_ctx* c, int chain, mcpm_box_energy_t* out, double* per_pair, double* per_wall) {
    int rc = check_chain(c, chain); if (rc) return rc;
    CU(cudaSetDevice(c->device));
    if (c->srv_running && k1_served(c, chain) && c->h_state[chain].st.n > 0 && c->h_state[chain].st.n <= MCPM_K1_SERVER_BOX_MAX) {
        // a host loop that interleaves BoxEnergy with its moves (CorrectEnergy as shipped): answered by the resident service
        double rec[8];
        rc = k1_server_call(c, K1_CMD_BOX, -1, nullptr, rec); if (rc) return rc;
        const int n = c->h_state[chain].st.n;
        mcpm_chain_stats& st = c->h_state[chain].st;
        st.pair = st.pair_check = rec[0]; st.wall = st.wall_check = rec[2]; st.energy = st.energy_check = rec[3];
        c->h_state[chain].energy_valid = 1;
        if (out) { out->pair = rec[0]; out->many_body = 0.0; out->wall = rec[2]; out->energy = rec[3]; }
        if (per_pair) for (int k = 0; k < n; k++) per_pair[k] = ((volatile double*)c->h_box)[k];
        if (per_wall) for (int k = 0; k < n; k++) per_wall[k] = ((volatile double*)c->h_box)[c->cap + k];
        return MCPM_OK;
    }
    QUIESCE(c);
    rc = box_energy_range(c, chain, 1, c->d_pp, c->d_pw); if (rc) return rc;
    if (out) { out->pair = c->h_pin[0]; out->many_body = 0.0; out->wall = c->h_pin[2]; out->energy = c->h_pin[3]; }
    const int n = c->h_state[chain].st.n;
    if (n > 0 && per_pair) CU(cudaMemcpyAsync(per_pair, c->d_pp, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    if (n > 0 && per_wall) CU(cudaMemcpyAsync(per_wall, c->d_pw, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return MCPM_OK;
}

int mcpm_box_energy_all(mcpm_ctx* c, mcpm_box_energy_t* out) {
    QUIESCE(c);
    if (!c) return fail(MCPM_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    int rc = box_energy_range(c, 0, c->n_chains, nullptr, nullptr); if (rc) return rc;
    if (out) for (int k = 0; k < c->n_chains; k++) {
        out[k].pair = c->h_pin[4 * k]; out[k].many_body = 0.0; out[k].wall = c->h_pin[4 * k + 2]; out[k].energy = c->h_pin[4 * k + 3];
    }
    return MCPM_OK;
}

// ---- state mutation -----------------------------------------------------------------------------
static int write_slot(mcpm_ctx* c, int chain, int slot, const double xyz[3]) {
    const size_t off = (size_t)chain * c->cap + slot;
    CU(cudaMemcpyAsync(c->d_x + off, &xyz[0], sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_y + off, &xyz[1], sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_z + off, &xyz[2], sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const ChainParams& P = c->h_params[chain];
    for (int k = 0; k < 3; k++) if (P.pbc[k] && !(xyz[k] >= 0.0 && xyz[k] <= P.L[k])) c->h_state[chain].fast_image = 0;
    c->h_state[chain].energy_valid = 0;
    c->cache_fresh[chain] = 0;
    return MCPM_OK;
}
int mcpm_apply_move(mcpm_ctx* c, int chain, int index, const double xyz[3]) {
    int rc = check_chain(c, chain); if (rc) return rc;
    if (!xyz || index < 0 || index >= c->h_state[chain].st.n) return fail(MCPM_ERR_ARG, "bad index %d", index);
    CU(cudaSetDevice(c->device));
    if (k1_served(c, chain) && c->srv_running && (!c->srv_fast || trial_in_box(c->h_params[chain], xyz))) {
        c->h_state[chain].energy_valid = 0; c->cache_fresh[chain] = 0;
        return k1_server_call(c, K1_CMD_APPLY, index, xyz, nullptr);
    }
    QUIESCE(c);
    rc = write_slot(c, chain, index, xyz); if (rc) return rc;
    rc = push_state(c, chain); if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return MCPM_OK;
}
int mcpm_append(mcpm_ctx* c, int chain, const double xyz[3]) {
    int rc = check_chain(c, chain); if (rc) return rc;
    if (!xyz) return fail(MCPM_ERR_ARG, "null xyz");
    const int n = c->h_state[chain].st.n;
    if (n >= c->cap) return fail(MCPM_ERR_CAPACITY, "chain %d is full (%d)", chain, c->cap);
    CU(cudaSetDevice(c->device));
    if (k1_served(c, chain) && c->srv_running && (!c->srv_fast || trial_in_box(c->h_params[chain], xyz))) {
        c->h_state[chain].energy_valid = 0; c->cache_fresh[chain] = 0;
        rc = k1_server_call(c, K1_CMD_APPEND, -1, xyz, nullptr); if (rc) return rc;
        c->h_state[chain].st.n = n + 1;
        return MCPM_OK;
    }
    QUIESCE(c);
    rc = write_slot(c, chain, n, xyz); if (rc) return rc;
    c->h_state[chain].st.n = n + 1;
    rc = push_state(c, chain); if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return MCPM_OK;
}
int mcpm_remove_swap_last(mcpm_ctx* c, int chain, int index) {
    int rc = check_chain(c, chain); if (rc) return rc;
    const int n = c->h_state[chain].st.n;
    if (index < 0 || index >= n) return fail(MCPM_ERR_ARG, "bad index %d (n=%d)", index, n);
    CU(cudaSetDevice(c->device));
    if (k1_served(c, chain) && c->srv_running) {
        c->h_state[chain].energy_valid = 0; c->cache_fresh[chain] = 0;
        rc = k1_server_call(c, K1_CMD_REMOVE, index, nullptr, nullptr); if (rc) return rc;
        c->h_state[chain].st.n = n - 1;
        return MCPM_OK;
    }
    QUIESCE(c);
    const size_t off = (size_t)chain * c->cap;
    if (index != n - 1) {
        CU(cudaMemcpyAsync(c->d_x + off + index, c->d_x + off + n - 1, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpyAsync(c->d_y + off + index, c->d_y + off + n - 1, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpyAsync(c->d_z + off + index, c->d_z + off + n - 1, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    }
    c->h_state[chain].st.n = n - 1;
    c->h_state[chain].energy_valid = 0;
    c->cache_fresh[chain] = 0;
    rc = push_state(c, chain); if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return MCPM_OK;
}

// ---- K2 ---------------------------------------------------------------------------------------
static int pick_threads(const mcpm_ctx* c, int maxn) {
    // Measured on B200 (scripts/latency_table.py, bench.py --config c1..c5 --threads T): aggregate throughput is best
    // with ~8 warps resident per SM in total — one warp per chain when many chains share an SM (the per-step overhead:
    // uniforms, wall term, reductions, decision, is paid once per warp), 256 threads when shared memory or the chain
    // count leaves one chain per SM (C3: N = 5.4k, 157 KB of positions per CTA).
    const size_t smem_per_cta = (size_t)(3 * c->cap) * sizeof(double) + 1024;
    const int by_smem = std::max<int>(1, (int)((size_t)c->max_smem / smem_per_cta));
    const int by_count = (c->n_chains + c->num_sms - 1) / c->num_sms;
    const int resident = std::max(1, std::min(by_smem, by_count));
    int t = 32;
    while (t * 2 * resident <= 256) t *= 2;
    while (t > 32 && t > maxn / 2) t /= 2;       // at least two partners per thread
    return t;
}

#ifdef MCPM_EXPERIMENTS
// ---- k2_pool: capacity classes, slot mix and queues ---------------------------------------------------------------------
// Slot a chain of n particles needs for one launch: headroom for the fluctuations of N (a chain that still outgrows it pauses and
// is finished in a full-size slot).
static int pool_need(int n, int cap) { return std::min(cap, ((int)(n * 1.15) + 48 + 31) & ~31); }

// Builds the layout for the chains listed in `chains` (all of them get slots of the full capacity if `full`), reorders them by
// (class, loading descending) into `order`, and returns the dynamic shared-memory size (0: does not fit).
static size_t pool_layout(const mcpm_ctx* c, const std::vector<int>& chains, bool full, int max_warps, PoolLayout& L, std::vector<int>& order, int& n_ctas) {
    memset(&L, 0, sizeof(L));
    const int cap = c->cap;
    int top = 32;
    for (int k : chains) top = std::max(top, full ? cap : pool_need(c->h_state[k].st.n, cap));
    // geometric classes: top, top/2, top/4 (multiples of 32); a chain joins the smallest class that holds its need
    int ccap[MCPM_POOL_MAX_CLASSES] = {top, std::max(32, ((top / 2) + 31) & ~31), std::max(32, ((top / 4) + 31) & ~31)};
    std::vector<int> members[MCPM_POOL_MAX_CLASSES];
    double work[MCPM_POOL_MAX_CLASSES] = {0., 0., 0.};
    for (int k : chains) {
        const int n = c->h_state[k].st.n, need = full ? cap : pool_need(n, cap);
        int cls = 0;
        for (int q = MCPM_POOL_MAX_CLASSES - 1; q > 0; q--) if (need <= ccap[q] && ccap[q] < ccap[q - 1]) { cls = q; break; }
        members[cls].push_back(k);
        work[cls] += 1400.0 + 6.0 * n;          // cycles per step: fixed latency + scan (profiles/r01_k2_step_phases.md)
    }
    // drop empty classes (keep the order large -> small)
    int ncls = 0;
    int keep_cap[MCPM_POOL_MAX_CLASSES]; double keep_work[MCPM_POOL_MAX_CLASSES]; std::vector<int> keep_members[MCPM_POOL_MAX_CLASSES];
    for (int q = 0; q < MCPM_POOL_MAX_CLASSES; q++) if (!members[q].empty()) { keep_cap[ncls] = ccap[q]; keep_work[ncls] = work[q]; keep_members[ncls] = members[q]; ncls++; }
    if (ncls == 0) return 0;
    const int n_chains = (int)chains.size();
    int slots = std::max(1, std::min(max_warps, (n_chains + c->num_sms - 1) / c->num_sms));
    const size_t stride = pool_state_stride_doubles();
    auto slot_doubles = [](int cp) { return (size_t)3 * cp + 64; };
    int per[MCPM_POOL_MAX_CLASSES];
    for (;; slots--) {
        if (slots < 1) return 0;
        if (slots < ncls) {                        // fewer warps than classes: one class of the largest capacity
            for (int q = 1; q < ncls; q++) { keep_members[0].insert(keep_members[0].end(), keep_members[q].begin(), keep_members[q].end()); keep_work[0] += keep_work[q]; }
            ncls = 1;
        }
        double tot = 0.; for (int q = 0; q < ncls; q++) tot += keep_work[q];
        int used = 0;
        for (int q = 0; q < ncls; q++) { per[q] = std::max(1, (int)(slots * keep_work[q] / tot)); used += per[q]; }
        for (int q = 0; used < slots; q = (q + 1) % ncls) { per[q]++; used++; }          // leftovers to the larger classes first
        for (int q = ncls - 1; used > slots; q = (q + ncls - 1) % ncls) if (per[q] > 1) { per[q]--; used--; }
        size_t doubles = (size_t)slots * stride;
        for (int q = 0; q < ncls; q++) doubles += per[q] * slot_doubles(keep_cap[q]);
        if (doubles * sizeof(double) + 64 + sizeof(RunArgs) + 16 <= (size_t)c->max_smem) break;
    }
    L.n_slots = slots; L.n_classes = ncls;
    size_t off = 0;
    int w = 0;
    for (int q = 0; q < ncls; q++)
        for (int k = 0; k < per[q]; k++, w++) { L.slot_cap[w] = keep_cap[q]; L.slot_off[w] = (int)off; L.slot_cls[w] = q; off += slot_doubles(keep_cap[q]); }
    off = (off + 1) & ~(size_t)1;                  // 16-byte alignment of the record area
    L.state_off = (int)off; L.state_stride = (int)stride;
    L.args_off = (int)(off + (size_t)slots * stride);
    order.clear();
    for (int q = 0; q < ncls; q++) {
        std::stable_sort(keep_members[q].begin(), keep_members[q].end(), [&](int a, int b) { return c->h_state[a].st.n > c->h_state[b].st.n; });
        L.q_begin[q] = (int)order.size();
        order.insert(order.end(), keep_members[q].begin(), keep_members[q].end());
        L.q_end[q] = (int)order.size();
    }
    n_ctas = std::min(c->num_sms, (n_chains + slots - 1) / slots);
    return (off + (size_t)slots * stride) * sizeof(double) + ((sizeof(RunArgs) + 15) / 16) * 16;
}

#endif  // MCPM_EXPERIMENTS

int mcpm_run(mcpm_ctx* c, const mcpm_run_params* p, mcpm_chain_stats* stats) {
    QUIESCE(c);
    if (!c || !p) return fail(MCPM_ERR_ARG, "null argument");
    if (p->n_sets < 0 || p->steps_per_set < 0 || p->steps_per_set > 2000000000LL) return fail(MCPM_ERR_ARG, "set count must be >= 0 and steps_per_set in [0, 2e9]");
    if (p->rng_mode != MCPM_RNG_PHILOX && p->rng_mode != MCPM_RNG_REPLAY) return fail(MCPM_ERR_ARG, "bad rng_mode %d", p->rng_mode);
    if (p->rng_mode == MCPM_RNG_REPLAY && (!p->replay_u || p->replay_stride == 0)) return fail(MCPM_ERR_ARG, "replay mode needs replay_u");
    int threads = p->threads_per_chain;
    if (threads != 0 && (threads < 32 || threads > 1024 || (threads & (threads - 1)))) return fail(MCPM_ERR_ARG, "threads_per_chain must be a power of two in [32,1024]");
    CU(cudaSetDevice(c->device));
    int maxn = 1;
    for (int k = 0; k < c->n_chains; k++) {
        if (!c->has_system[k]) return fail(MCPM_ERR_ARG, "set the system of chain %d first", k);
        maxn = std::max(maxn, c->h_state[k].st.n);
    }
    const bool global_path = (size_t)c->cap > k2_smem_limit_cap(c->max_smem);   // positions stay in global memory / L2
    if (threads == 0) threads = global_path ? 512 : pick_threads(c, maxn);
    if (p->threads_per_chain == 0 && p->sample_rdf && !p->no_sampling) threads = std::max(threads, 256);   // ComputeRDF is O(N^2) per production step: all the threads the sampling kernel takes (2x measured)
    bool all_fast = true;
    for (int k = 0; k < c->n_chains; k++) all_fast = all_fast && c->h_state[k].fast_image && !c->h_params[k].exotic;
    bool sample = p->sample_rdf != 0;                       // in-kernel sampling: RDF on request, Widom whenever n_swap == 0
    for (int k = 0; k < c->n_chains; k++) sample = sample || c->h_params[k].n_swap == 0;
    if (p->no_sampling) sample = false;                     // moves only (MC::MinimizeEnergy): the hot kernels apply
    threads = k2_threads_supported(threads, global_path, all_fast, sample);
    bool need_init = false;
    for (int k = 0; k < c->n_chains; k++) need_init |= !c->h_state[k].energy_valid;
    if (need_init) {  // exact bookkeeping starts from a full BoxEnergy evaluation (src/energy.cpp:20)
        int rc0 = box_energy_range(c, 0, c->n_chains, nullptr, nullptr); if (rc0) return rc0;
        CU(cudaMemcpyAsync(c->d_state, c->h_state.data(), sizeof(ChainState) * c->n_chains, cudaMemcpyHostToDevice, c->stream));
    }
    bool any_linked = false, all_linked = true;
    for (int k = 0; k < c->n_chains; k++) { any_linked = any_linked || c->peer[k] >= 0; all_linked = all_linked && c->peer[k] >= 0; }
    if (any_linked) {
        // Two-box Gibbs systems (mcpm_link_boxes): one CTA per pair of boxes, mcpm_gibbs.cu
        if (!all_linked) return fail(MCPM_ERR_UNSUPPORTED, "a context holds either single-box chains or linked pairs of boxes, not both");
        if (p->sample_rdf) return fail(MCPM_ERR_UNSUPPORTED, "RDF sampling of linked boxes is not supported");
        std::vector<int2> pairs;
        for (int k = 0; k < c->n_chains; k++) {
            if (c->box_id[k] != 0) continue;
            int rcp = check_pair(c, k, c->peer[k]); if (rcp) return rcp;
            pairs.push_back(make_int2(k, c->peer[k]));
        }
        std::stable_sort(pairs.begin(), pairs.end(), [&](const int2& u, const int2& v) {
            return c->h_state[u.x].st.n + c->h_state[u.y].st.n > c->h_state[v.x].st.n + c->h_state[v.y].st.n; });
        if (!c->d_pairs) CU(cudaMalloc(&c->d_pairs, sizeof(int2) * (size_t)c->n_chains));
        CU(cudaMemcpyAsync(c->d_pairs, pairs.data(), sizeof(int2) * pairs.size(), cudaMemcpyHostToDevice, c->stream));
        RunArgs ga;
        memset(&ga, 0, sizeof(ga));
        ga.first_set = p->first_set; ga.n_sets = p->n_sets; ga.steps_per_set = p->steps_per_set; ga.seed = p->seed;
        ga.check_energy = p->check_energy; ga.cap = c->cap; ga.no_sampling = p->no_sampling ? 1 : 0;
        const unsigned long long gsteps = (unsigned long long)p->n_sets * (unsigned long long)p->steps_per_set;
        if (p->rng_mode == MCPM_RNG_REPLAY) {
            const size_t need = (size_t)c->n_chains * p->replay_stride;
            if (need > c->replay_doubles) {
                cudaFree(c->d_replay); c->d_replay = nullptr; c->replay_doubles = 0;
                CU(cudaMalloc(&c->d_replay, sizeof(double) * need));
                c->replay_doubles = need;
            }
            CU(cudaMemcpyAsync(c->d_replay, p->replay_u, sizeof(double) * need, cudaMemcpyHostToDevice, c->stream));
            ga.replay_u = c->d_replay; ga.replay_stride = p->replay_stride;
        }
        if (p->trace) {
            const size_t need = (size_t)c->n_chains * gsteps;
            if (need > c->trace_bytes) {
                cudaFree(c->d_trace); c->d_trace = nullptr; c->trace_bytes = 0;
                CU(cudaMalloc(&c->d_trace, std::max<size_t>(need, 1)));
                c->trace_bytes = need;
            }
            CU(cudaMemsetAsync(c->d_trace, 0xff, std::max<size_t>(need, 1), c->stream));
            ga.trace = c->d_trace; ga.trace_stride = gsteps;
        }
        CU(cudaEventRecord(c->ev0, c->stream));
        c->last_kernel = "k2_gibbs";
        CU(launch_k2_gibbs(p->rng_mode, (int)pairs.size(), c->d_params, c->d_state, c->d_pairs, c->d_x, c->d_y, c->d_z, ga, c->stream));
        c->launches++;
        CU(cudaEventRecord(c->ev1, c->stream));
        if (p->trace) CU(cudaMemcpyAsync(p->trace, c->d_trace, (size_t)c->n_chains * gsteps, cudaMemcpyDeviceToHost, c->stream));
        int rcg = pull_states(c); if (rcg) return rcg;
        CU(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
        int full = -1;
        for (int k = 0; k < c->n_chains; k++) {
            c->cache_fresh[k] = 0;
            if (stats) stats[k] = c->h_state[k].st;
            if (c->h_state[k].st.overflow && full < 0) full = k;
        }
        if (full >= 0) return fail(MCPM_ERR_CAPACITY, "a box of the system of chain %d reached its capacity of %d particles and the system stopped", full, c->cap);
        return MCPM_OK;
    }
    // longest chains first (cost per step ~ n)
    std::vector<int> order(c->n_chains);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return c->h_state[a].st.n > c->h_state[b].st.n; });
    CU(cudaMemcpyAsync(c->d_order, order.data(), sizeof(int) * c->n_chains, cudaMemcpyHostToDevice, c->stream));

    RunArgs a;
    memset(&a, 0, sizeof(a));
    a.first_set = p->first_set; a.n_sets = p->n_sets; a.steps_per_set = p->steps_per_set; a.seed = p->seed;
    a.check_energy = p->check_energy; a.cap = c->cap; a.pool_threads = 32;
    a.sample_rdf = p->no_sampling ? 0 : p->sample_rdf; a.rdf = c->d_rdf; a.no_sampling = p->no_sampling ? 1 : 0;
    // The energy cache pays while accepted translations (each costs a second pass over the partners) are the minority:
    // N + 2pN < 2N for p < 0.5.  Decide from the acceptance of the last set; the reference's dr adaptation aims at 0.2.
    bool cache_now = c->use_cache != 0;
    {
        double acc = 0.0, att = 0.0;
        for (int k = 0; k < c->n_chains; k++) { acc += (double)c->h_state[k].st.acceptance; att += (double)c->h_state[k].st.n_displacements - 1.0; }
        if (att >= 100.0 * c->n_chains && acc / att > 0.4) cache_now = false;
    }
    // K2 through a cell list (mcpm_cells.cu): large chains whose box holds >= 3 cut-off lengths per periodic axis (config C4).
    // ~27 cells x occupancy candidate partners per energy evaluation instead of N; measured break-even is a few thousand particles.
    bool use_k2c = c->use_cells && all_fast && !a.sample_rdf && p->threads_per_chain == 0 && maxn >= 4096;
    int kc_stride = 1, kc_ccap = 64;
    size_t kc_smem = 0;
    if (use_k2c) {
        double occ = 0.0;
        for (int k = 0; k < c->n_chains && use_k2c; k++) {
            const ChainParams& P = c->h_params[k];
            const double rc = sqrt(P.rc2);
            int dims[3];
            for (int ax = 0; ax < 3; ax++) {
                dims[ax] = std::min(16, (int)(P.L[ax] / rc));
                if (P.pbc[ax] && dims[ax] < 3) use_k2c = false;
                dims[ax] = std::max(1, dims[ax]);
            }
            const int ncell = dims[0] * dims[1] * dims[2];
            kc_stride = std::max(kc_stride, ncell);
            occ = std::max(occ, (double)c->h_state[k].st.n / ncell);
        }
        bool any_swap = false;
        for (int k = 0; k < c->n_chains; k++) any_swap = any_swap || c->h_params[k].n_swap > 0;
        // room for density fluctuations (NVT: twice the mean occupancy, > 8 sigma in a dense fluid) and, with swap moves, for growth;
        // a full cell stops the chain
        kc_ccap = any_swap ? std::max(64, (((int)(3.0 * occ) + 32 + 31) / 32) * 32) : std::max(64, (((int)(2.0 * occ) + 15) / 16) * 16);
        kc_smem = use_k2c ? k2_cells_smem_items_bytes(c->cap, kc_stride, kc_ccap, c->max_smem) : 0;
        if (use_k2c) {
            const size_t need = kc_smem ? 1 : (size_t)c->n_chains * kc_stride * kc_ccap;   // lists in shared memory need no global copy
            if (need > c->kc_items_ints) {
                cudaFree(c->d_kc_items); c->d_kc_items = nullptr; c->kc_items_ints = 0;
                CU(cudaMalloc(&c->d_kc_items, sizeof(int) * need));
                c->kc_items_ints = need;
            }
            if (!c->d_kc_cell_of) {
                CU(cudaMalloc(&c->d_kc_cell_of, sizeof(int) * (size_t)c->n_chains * c->cap));
                CU(cudaMalloc(&c->d_kc_slot_of, sizeof(int) * (size_t)c->n_chains * c->cap));
            }
        }
    }
    if (use_k2c) cache_now = false;
    // The speculative-moves kernel (mcpm_spec.cu) trades throughput for latency: it is chosen when there are too few
    // chains to fill the GPU anyway, the caller left the thread count open, and moves are mostly rejected.  It needs the
    // energy cache, coordinates inside the box, shared-memory positions, no in-kernel sampling and one geometry.
    bool same_pbz = true;
    for (int k = 1; k < c->n_chains; k++) same_pbz = same_pbz && c->h_params[k].pbc[2] == c->h_params[0].pbc[2];
    const bool spec_ok = !use_k2c && c->use_cache && all_fast && !sample && !global_path && same_pbz && k2_spec_smem_bytes(c->cap) <= (size_t)c->max_smem;
    // Automatic choice: chains that cannot fill the GPU anyway, and small enough that a candidate's scan by one warp is latency-
    // rather than FP64-throughput-bound (measured: 2.6-3x faster at N <= 600, break-even near N ~ 5000 where the discarded
    // candidates' evaluations cost as much as the latency they hide).
    // Two-warp speculative chains: when one warp per chain was chosen but at most 8 chains share an SM (one wave, e.g. config C5),
    // a second warp per chain evaluating the following step uses issue slots that would otherwise idle (+12 % measured on C5);
    // with more chains per SM it costs residency (128 registers -> 8 chains per SM) and loses.
    const int per_sm = (c->n_chains + c->num_sms - 1) / c->num_sms;
    // (2-4 chains per SM: four candidate warps per chain.)  The thread count picked for the one-step kernel does not matter here.
    const bool lean_auto = c->use_spec < 0 && p->threads_per_chain == 0 && per_sm <= 8 && c->n_chains > c->num_sms && maxn <= 2048;
    const bool pair_wanted = c->use_spec == 2 || c->use_spec == 4 || lean_auto;
    const int lean_width = c->use_spec == 4 ? 4 : (c->use_spec == 2 ? 2 : (per_sm <= 4 ? 4 : 2));
    const bool use_pair = pair_wanted && !use_k2c && k2_pair_smem_bytes(c->cap) <= (size_t)c->max_smem && (threads == 32 || lean_auto || c->use_spec == 4) && cache_now && all_fast && !sample && !global_path && same_pbz;
    const bool use_spec = !use_pair && spec_ok && (c->use_spec == 1 || (c->use_spec < 0 && cache_now && p->threads_per_chain == 0 && c->n_chains <= c->num_sms && maxn <= 2048));
    if (use_spec) cache_now = true;
    if (cache_now && !c->d_ep) CU(cudaMalloc(&c->d_ep, sizeof(double) * (size_t)c->n_chains * c->cap));
    a.ep = cache_now ? c->d_ep : nullptr;
    const unsigned long long nsteps = (unsigned long long)p->n_sets * (unsigned long long)p->steps_per_set;
    if (p->rng_mode == MCPM_RNG_REPLAY) {
        const size_t need = (size_t)c->n_chains * p->replay_stride;
        if (need > c->replay_doubles) {
            cudaFree(c->d_replay); c->d_replay = nullptr; c->replay_doubles = 0;
            CU(cudaMalloc(&c->d_replay, sizeof(double) * need));
            c->replay_doubles = need;
        }
        CU(cudaMemcpyAsync(c->d_replay, p->replay_u, sizeof(double) * need, cudaMemcpyHostToDevice, c->stream));
        a.replay_u = c->d_replay; a.replay_stride = p->replay_stride;
    }
    if (p->trace) {
        const size_t need = (size_t)c->n_chains * nsteps;
        if (need > c->trace_bytes) {
            cudaFree(c->d_trace); c->d_trace = nullptr; c->trace_bytes = 0;
            CU(cudaMalloc(&c->d_trace, std::max<size_t>(need, 1)));
            c->trace_bytes = need;
        }
        CU(cudaMemsetAsync(c->d_trace, 0xff, std::max<size_t>(need, 1), c->stream));
        a.trace = c->d_trace; a.trace_stride = nsteps;
    }
    // The sub-warp kernel (mcpm_subwarp.cu) takes over whenever one warp per chain was chosen: it needs the energy cache,
    // coordinates inside the box, shared-memory positions, no in-kernel sampling and one geometry for the whole launch.
#ifdef MCPM_EXPERIMENTS
    const bool use_sub = c->use_subwarp && threads == 32 && cache_now && all_fast && !sample && !global_path && same_pbz;
#else
    const bool use_sub = false;
#endif
    CU(cudaEventRecord(c->ev0, c->stream));
#ifdef MCPM_EXPERIMENTS
    const bool use_dual = c->use_dual && !use_k2c && !use_spec && !use_sub && threads == 32 && cache_now && all_fast && !sample && !global_path && same_pbz;
#else
    const bool use_dual = false;
#endif
    // One warp per chain on the energy cache: the persistent pooled kernel (shared-memory slots per capacity class, work queues).
#ifdef MCPM_EXPERIMENTS
    const bool use_pool = c->use_pool && !use_k2c && !use_pair && !use_spec && !use_sub && !use_dual && threads == 32 && cache_now && all_fast && !sample && !global_path && same_pbz;
#else
    const bool use_pool = false;
#endif
    c->last_kernel = use_pool ? "k2_pool" : use_k2c ? "k2_cells" : use_pair ? (lean_width == 4 ? "k2_quad" : "k2_pair") : use_spec ? "k2_spec" : (use_sub ? "k2_sub" : (use_dual ? "k2_dual" : "k2_chains"));
#ifdef MCPM_EXPERIMENTS
    if (use_pool) {
        if (!c->d_pool_next) CU(cudaMalloc(&c->d_pool_next, sizeof(int) * MCPM_POOL_MAX_CLASSES));
        for (int k = 0; k < c->n_chains; k++) c->h_state[k].res.paused = 0;
        CU(cudaMemcpyAsync(c->d_state, c->h_state.data(), sizeof(ChainState) * c->n_chains, cudaMemcpyHostToDevice, c->stream));
        std::vector<int> todo(c->n_chains);
        std::iota(todo.begin(), todo.end(), 0);
        for (int round = 0; !todo.empty(); round++) {
            if (round > 8) return fail(MCPM_ERR_CUDA, "k2_pool: chains still paused after %d launches", round);
            PoolLayout L;
            std::vector<int> pool_order;
            int n_ctas = 0;
            const size_t smem = pool_layout(c, todo, round > 0, c->pool_warps, L, pool_order, n_ctas);
            if (smem == 0) return fail(MCPM_ERR_CAPACITY, "k2_pool: a chain of capacity %d does not fit one SM's shared memory", c->cap);
            L.q_next = c->d_pool_next;
            CU(cudaMemsetAsync(c->d_pool_next, 0, sizeof(int) * MCPM_POOL_MAX_CLASSES, c->stream));
            CU(cudaMemcpyAsync(c->d_order, pool_order.data(), sizeof(int) * pool_order.size(), cudaMemcpyHostToDevice, c->stream));
            CU(launch_k2_pool(p->rng_mode, c->h_params[0].pbc[2] != 0, c->pool_warps, n_ctas, smem, c->d_params, c->d_state, c->d_order, c->d_x, c->d_y, c->d_z, a, L, c->stream));
            c->launches++;
            if (round == 0) CU(cudaEventRecord(c->ev1, c->stream));      // (follow-up launches for paused chains are rare and short)
            int rc1 = pull_states(c); if (rc1) return rc1;
            todo.clear();
            for (int k = 0; k < c->n_chains; k++) if (c->h_state[k].res.paused) todo.push_back(k);
        }
    } else
#endif
    if (use_k2c) {
        bool widom = false;
        for (int k = 0; k < c->n_chains; k++) widom = widom || (c->h_params[k].n_swap == 0 && !p->no_sampling);
        CU(launch_k2_cells(p->rng_mode, widom, c->n_chains, c->d_params, c->d_state, c->d_order, c->d_x, c->d_y, c->d_z, a, c->d_kc_items,
                           c->d_kc_cell_of, c->d_kc_slot_of, kc_stride, kc_ccap, kc_smem, c->stream));
        c->launches++;
#ifdef MCPM_EXPERIMENTS
    } else if (use_dual) {
        CU(launch_k2_dual(p->rng_mode, c->h_params[0].pbc[2] != 0, c->n_chains, c->d_params, c->d_state, c->d_order, c->d_x, c->d_y, c->d_z, a, c->stream));
        c->launches++;
#endif
    } else if (use_pair) {
        CU(launch_k2_spec(p->rng_mode, c->h_params[0].pbc[2] != 0, lean_width, c->n_chains, c->d_params, c->d_state, c->d_order, c->d_x, c->d_y, c->d_z, a, c->stream));
        c->launches++;
    } else if (use_spec) {
        CU(launch_k2_spec(p->rng_mode, c->h_params[0].pbc[2] != 0, 0, c->n_chains, c->d_params, c->d_state, c->d_order, c->d_x, c->d_y, c->d_z, a, c->stream));
        c->launches++;
#ifdef MCPM_EXPERIMENTS
    } else if (use_sub) {
        const int cap_cta = ((c->cap + 127) / 128) * 128;
        if (!c->d_ctas) CU(cudaMalloc(&c->d_ctas, sizeof(int2) * c->n_chains));
        auto chains_per_warp = [&](int n) {   // room for the fluctuations of N during the launch; outgrowing the slot pauses the chain
            const int need = n + std::max(32, n / 4) + 1;
            return need <= cap_cta / 4 ? 4 : (need <= cap_cta / 2 ? 2 : 1);
        };
        std::vector<int2> ctas;
        for (int i = 0; i < c->n_chains;) {     // `order` is sorted by loading: equal classes are contiguous
            const int k = chains_per_warp(c->h_state[order[i]].st.n);
            int cnt = 1;
            while (cnt < k && i + cnt < c->n_chains && chains_per_warp(c->h_state[order[i + cnt]].st.n) == k) cnt++;
            ctas.push_back(make_int2(i, k | (cnt << 8)));
            i += cnt;
        }
        for (int k = 0; k < c->n_chains; k++) c->h_state[k].res.paused = 0;
        CU(cudaMemcpyAsync(c->d_state, c->h_state.data(), sizeof(ChainState) * c->n_chains, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->d_ctas, ctas.data(), sizeof(int2) * ctas.size(), cudaMemcpyHostToDevice, c->stream));
        CU(launch_k2_sub(p->rng_mode, c->h_params[0].pbc[2] != 0, (int)ctas.size(), c->d_params, c->d_state, c->d_order, c->d_ctas, c->d_x,
                         c->d_y, c->d_z, a, cap_cta, c->stream));
        c->launches++;
        for (int round = 0; round < 4; round++) {   // chains that outgrew their slot continue alone in a warp with the full capacity
            int rc1 = pull_states(c); if (rc1) return rc1;
            std::vector<int> again;
            for (int k = 0; k < c->n_chains; k++) if (c->h_state[k].res.paused) again.push_back(k);
            if (again.empty()) break;
            std::vector<int2> solo;
            for (size_t i = 0; i < again.size(); i++) solo.push_back(make_int2((int)i, 1 | (1 << 8)));
            CU(cudaMemcpyAsync(c->d_order, again.data(), sizeof(int) * again.size(), cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemcpyAsync(c->d_ctas, solo.data(), sizeof(int2) * solo.size(), cudaMemcpyHostToDevice, c->stream));
            CU(launch_k2_sub(p->rng_mode, c->h_params[0].pbc[2] != 0, (int)solo.size(), c->d_params, c->d_state, c->d_order, c->d_ctas, c->d_x,
                             c->d_y, c->d_z, a, cap_cta, c->stream));
            c->launches++;
        }
#endif
    } else {
        CU(launch_k2(p->rng_mode, c->n_chains, threads, all_fast, sample, cache_now, c->d_params, c->d_state, c->d_order, c->d_x, c->d_y, c->d_z, a, c->max_smem, c->stream));
        c->launches++;
    }
    if (!use_pool) CU(cudaEventRecord(c->ev1, c->stream));
    if (p->trace) CU(cudaMemcpyAsync(p->trace, c->d_trace, (size_t)c->n_chains * nsteps, cudaMemcpyDeviceToHost, c->stream));
    int rc = pull_states(c); if (rc) return rc;
    for (int k = 0; k < c->n_chains; k++) c->cache_fresh[k] = (a.ep != nullptr && c->h_state[k].cache_valid) ? 1 : 0;
    CU(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
    if (use_k2c && p->check_energy) {   // the cell-list kernel leaves the end-of-launch energy check to K3 (itself through a cell list)
        bool stopped = false;
        for (int k = 0; k < c->n_chains; k++) stopped = stopped || c->h_state[k].st.overflow;
        if (!stopped) {
            std::vector<double> run(3 * (size_t)c->n_chains);
            for (int k = 0; k < c->n_chains; k++) { run[3 * k] = c->h_state[k].st.pair; run[3 * k + 1] = c->h_state[k].st.wall; run[3 * k + 2] = c->h_state[k].st.energy; }
            const float ms = c->last_ms;
            rc = box_energy_range(c, 0, c->n_chains, nullptr, nullptr); if (rc) return rc;     // fills *_check and overwrites the running values
            c->last_ms = ms;
            if (p->check_energy != 2)
                for (int k = 0; k < c->n_chains; k++) { c->h_state[k].st.pair = run[3 * k]; c->h_state[k].st.wall = run[3 * k + 1]; c->h_state[k].st.energy = run[3 * k + 2]; }
            CU(cudaMemcpyAsync(c->d_state, c->h_state.data(), sizeof(ChainState) * c->n_chains, cudaMemcpyHostToDevice, c->stream));
            CU(cudaStreamSynchronize(c->stream));
        }
    }
    {   // NPT: the box a chain ended with becomes its system (widths, volume-dependent constants) for K1 / K3 and the next launch
        bool any = false;
        for (int k = 0; k < c->n_chains; k++) {
            if (c->h_params[k].n_vol <= 0 || !(c->h_state[k].st.width > 0) || c->h_state[k].st.width == c->h_desc[k].width[2]) continue;
            mcpm_system_desc d = c->h_desc[k];
            d.width[0] = d.width[1] = d.width[2] = c->h_state[k].st.width;
            ChainParams P;
            memset(&P, 0, sizeof(P));
            if (digest(d, P) == MCPM_OK) { c->h_desc[k] = d; c->h_params[k] = P; any = true; }
        }
        if (any) {
            CU(cudaMemcpyAsync(c->d_params, c->h_params.data(), sizeof(ChainParams) * c->n_chains, cudaMemcpyHostToDevice, c->stream));
            CU(cudaStreamSynchronize(c->stream));
        }
    }
    int overflowed = -1;
    for (int k = 0; k < c->n_chains; k++) {
        if (stats) stats[k] = c->h_state[k].st;
        if (c->h_state[k].st.overflow && overflowed < 0) overflowed = k;
    }
    if (overflowed >= 0 && c->h_state[overflowed].st.overflow == 2)
        return fail(MCPM_ERR_CAPACITY, "chain %d: a cell of its cell list ran out of slots (the density grew by more than 3x during the run) and the chain stopped", overflowed);
    if (overflowed >= 0) return fail(MCPM_ERR_CAPACITY, "chain %d reached its capacity of %d particles and stopped", overflowed, c->cap);
    return MCPM_OK;
}

int mcpm_get_stats(mcpm_ctx* c, mcpm_chain_stats* stats) {
    if (!c || !stats) return fail(MCPM_ERR_ARG, "null argument");
    for (int k = 0; k < c->n_chains; k++) stats[k] = c->h_state[k].st;
    return MCPM_OK;
}

int mcpm_reset_accumulators(mcpm_ctx* c) {
    QUIESCE(c);
    if (!c) return fail(MCPM_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    for (int k = 0; k < c->n_chains; k++) {
        mcpm_chain_stats& st = c->h_state[k].st;
        st.sum_n = st.sum_n2 = st.sum_e = st.sum_e2 = 0.0; st.samples = 0;
        st.tot_disp = st.tot_disp_acc = st.tot_ins = st.tot_ins_acc = st.tot_del = st.tot_del_acc = st.tot_empty = 0;
        st.pair_evals = 0; st.pair_evals_algorithmic = 0;
    }
    CU(cudaMemcpyAsync(c->d_state, c->h_state.data(), sizeof(ChainState) * c->n_chains, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(c->d_rdf, 0, sizeof(double) * MCPM_RDF_BINS * (size_t)c->n_chains, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return MCPM_OK;
}

int mcpm_get_rdf(mcpm_ctx* c, int chain, double* hist) {
    QUIESCE(c);
    int rc = check_chain(c, chain); if (rc) return rc;
    if (!hist) return fail(MCPM_ERR_ARG, "null hist");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(hist, c->d_rdf + (size_t)chain * MCPM_RDF_BINS, sizeof(double) * MCPM_RDF_BINS, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return MCPM_OK;
}

// MC::ComputeChemicalPotential, src/moves.cpp:453-465
int mcpm_chemical_potential(mcpm_ctx* c, int chain, double* mu_ex, double* mu) {
    int rc = check_chain(c, chain); if (rc) return rc;
    if (!c->has_system[chain]) return fail(MCPM_ERR_ARG, "chain %d has no system yet", chain);
    const mcpm_system_desc& d = c->h_desc[chain];
    const mcpm_chain_stats& st = c->h_state[chain].st;
    const double kb = 1.380649e-23, na = 6.02214076e23, planck = 6.62607015e-34;
    const double particleMass = d.molar_mass / na * 1e-3;
    const double thermalWL = planck / sqrt(2.0 * M_PI * particleMass * kb * d.temperature) * 1e10;
    const double volume = box_volume(d);
    const double widomParam = (st.widom_insert / st.widom_insertions) / (st.widom_delete / st.widom_deletions);
    const double muIdeal = d.temperature * log(ipow(thermalWL, 3) * (st.n + 1) / volume);
    const double muExcess = -d.temperature * log(widomParam);
    if (mu_ex) *mu_ex = muExcess;
    if (mu) *mu = muIdeal + muExcess;
    return MCPM_OK;
}

int mcpm_set_running_energy(mcpm_ctx* c, const double* pair, const double* wall) {
    QUIESCE(c);
    if (!c || !pair || !wall) return fail(MCPM_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    for (int k = 0; k < c->n_chains; k++) {
        mcpm_chain_stats& st = c->h_state[k].st;
        st.pair = pair[k]; st.wall = wall[k]; st.energy = pair[k] + wall[k];
        c->h_state[k].energy_valid = 1;
    }
    CU(cudaMemcpyAsync(c->d_state, c->h_state.data(), sizeof(ChainState) * c->n_chains, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return MCPM_OK;
}

#ifdef MCPM_EXPERIMENTS
int mcpm_use_dual_candidates(mcpm_ctx* c, int enable) {
    if (!c) return fail(MCPM_ERR_ARG, "null context");
    c->use_dual = enable ? 1 : 0;
    return MCPM_OK;
}
#endif

int mcpm_use_speculation(mcpm_ctx* c, int mode) {
    if (!c) return fail(MCPM_ERR_ARG, "null context");
    if (mode < -1 || (mode > 2 && mode != 4)) return fail(MCPM_ERR_ARG, "speculation mode must be -1 (automatic), 0 (off), 1 (on), 2 or 4 (lean variants with 2 / 4 candidate warps per chain)");
    c->use_spec = mode;
    return MCPM_OK;
}

#ifdef MCPM_EXPERIMENTS
int mcpm_use_subwarp(mcpm_ctx* c, int enable) {
    if (!c) return fail(MCPM_ERR_ARG, "null context");
    c->use_subwarp = enable ? 1 : 0;
    return MCPM_OK;
}
#endif

#ifdef MCPM_EXPERIMENTS
int mcpm_use_pool(mcpm_ctx* c, int warps) {
    if (!c) return fail(MCPM_ERR_ARG, "null context");
    if (warps != 0 && warps != 12 && warps != 14 && warps != 16 && warps != 20) return fail(MCPM_ERR_ARG, "k2_pool runs 12, 16 or 20 warps per CTA (0: the one-CTA-per-chain kernel)");
    c->use_pool = warps ? 1 : 0;
    if (warps) c->pool_warps = warps;
    return MCPM_OK;
}
#endif

int mcpm_use_energy_cache(mcpm_ctx* c, int enable) {
    if (!c) return fail(MCPM_ERR_ARG, "null context");
    c->use_cache = enable ? 1 : 0;
    return MCPM_OK;
}

int mcpm_get_energy_cache(mcpm_ctx* c, int chain, double* pair, int* valid) {
    QUIESCE(c);
    int rc = check_chain(c, chain); if (rc) return rc;
    if (!pair || !valid) return fail(MCPM_ERR_ARG, "null argument");
    const ChainState& S = c->h_state[chain];
    *valid = (c->d_ep && S.cache_valid && c->cache_fresh[chain]) ? 1 : 0;
    if (!*valid) return MCPM_OK;
    CU(cudaSetDevice(c->device));
    const int n = S.st.n;
    if (n > 0) CU(cudaMemcpyAsync(pair, c->d_ep + (size_t)chain * c->cap, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const double eps4 = c->h_params[chain].eps4;          // the cache holds sums of s6(s6-1); 4*eps is applied once per sum
    for (int j = 0; j < n; j++) pair[j] *= eps4;
    return MCPM_OK;
}

int mcpm_use_cell_list(mcpm_ctx* c, int enable) {
    if (!c) return fail(MCPM_ERR_ARG, "null context");
    c->use_cells = enable ? 1 : 0;
    return MCPM_OK;
}

int mcpm_set_stream_ids(mcpm_ctx* c, const unsigned int* ids) {
    QUIESCE(c);
    if (!c || !ids) return fail(MCPM_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    for (int k = 0; k < c->n_chains; k++) c->h_state[k].stream_id = ids[k];
    CU(cudaMemcpyAsync(c->d_state, c->h_state.data(), sizeof(ChainState) * c->n_chains, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return MCPM_OK;
}

int mcpm_set_dr(mcpm_ctx* c, int chain, double dr) {
    QUIESCE(c);
    int rc = check_chain(c, chain); if (rc) return rc;
    CU(cudaSetDevice(c->device));
    c->h_state[chain].st.dr = dr;
    rc = push_state(c, chain); if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return MCPM_OK;
}
int mcpm_set_step_counter(mcpm_ctx* c, int chain, unsigned long long step) {
    QUIESCE(c);
    int rc = check_chain(c, chain); if (rc) return rc;
    CU(cudaSetDevice(c->device));
    c->h_state[chain].st.steps_done = step;
    rc = push_state(c, chain); if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return MCPM_OK;
}

// ---- measurement support ------------------------------------------------------------------------
const char* mcpm_last_kernel_name(const mcpm_ctx* c) { return c ? c->last_kernel : ""; }

int mcpm_last_kernel_ms(const mcpm_ctx* c, float* ms) {
    if (!c || !ms) return fail(MCPM_ERR_ARG, "null argument");
    *ms = c->last_ms;
    return MCPM_OK;
}
int mcpm_transfer_bytes(const mcpm_ctx* c, unsigned long long* h2d, unsigned long long* d2h) {
    if (!c) return fail(MCPM_ERR_ARG, "null context");
    if (h2d) *h2d = c->pos_bytes_h2d;
    if (d2h) *d2h = c->pos_bytes_d2h;
    return MCPM_OK;
}
int mcpm_launch_count(const mcpm_ctx* c, unsigned long long* launches) {
    if (!c || !launches) return fail(MCPM_ERR_ARG, "null argument");
    *launches = c->launches;
    return MCPM_OK;
}

int mcpm_timer_start(mcpm_ctx* c) {
    if (!c) return fail(MCPM_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaEventRecord(c->tm0, c->stream));
    return MCPM_OK;
}
int mcpm_timer_stop(mcpm_ctx* c, float* ms) {
    if (!c || !ms) return fail(MCPM_ERR_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    CU(cudaEventRecord(c->tm1, c->stream));
    CU(cudaEventSynchronize(c->tm1));
    CU(cudaEventElapsedTime(ms, c->tm0, c->tm1));
    return MCPM_OK;
}

// Host-side latency of mcpm_particle_energy as a native caller sees it (no interpreter in the loop): `reps` calls on stored particles
// 0, 1, 2, ... displaced to `trial`, wall-clock microseconds per call.  Uses whatever mode the chain is in (launch per call or resident).
int mcpm_k1_latency(mcpm_ctx* c, int chain, int reps, const double* trial, double* us_per_call) {
    int rc = check_chain(c, chain); if (rc) return rc;
    if (reps <= 0 || !us_per_call) return fail(MCPM_ERR_ARG, "bad arguments");
    const int n = std::max(1, c->h_state[chain].st.n);
    mcpm_particle_energy_t cur, moved;
    for (int r = 0; r < 16; r++) { rc = mcpm_particle_energy(c, chain, r % n, trial, &cur, &moved); if (rc) return rc; }
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; r++) { rc = mcpm_particle_energy(c, chain, r % n, trial, &cur, &moved); if (rc) return rc; }
    const auto t1 = std::chrono::steady_clock::now();
    *us_per_call = std::chrono::duration<double, std::micro>(t1 - t0).count() / reps;
    return MCPM_OK;
}

int mcpm_fp64_peak(int device, double* tflops, double* ms_out) {
    if (!tflops) return fail(MCPM_ERR_ARG, "null tflops");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(MCPM_ERR_CUDA, "device %d not available", device);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 1 << 16;
    double* d = nullptr;
    CU(cudaMalloc(&d, sizeof(double) * (size_t)threads * blocks));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {  // first repetitions warm the clocks up
        CU(cudaEventRecord(e0, 0));
        CU(launch_fp64_peak(d, blocks, threads, iters, 0));
        CU(cudaEventRecord(e1, 0));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (rep >= 2) best = std::min(best, ms);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    const double flops = 2.0 * 8.0 * (double)iters * (double)threads * (double)blocks;
    *tflops = flops / (best * 1e-3) * 1e-12;
    if (ms_out) *ms_out = best;
    return MCPM_OK;
}

int mcpm_metropolis_probe(mcpm_ctx* c, int kind, int count, const double* u, const double* x, const int* n, double zzV, int* out) {
    QUIESCE(c);
    if (!c || !u || !x || !out || count <= 0 || kind < 0 || kind > 2 || (kind > 0 && !n)) return fail(MCPM_ERR_ARG, "bad argument");
    CU(cudaSetDevice(c->device));
    double *du = nullptr, *dx = nullptr; int *dn = nullptr, *dout = nullptr;
    CU(cudaMalloc(&du, sizeof(double) * count)); CU(cudaMalloc(&dx, sizeof(double) * count));
    CU(cudaMalloc(&dn, sizeof(int) * count)); CU(cudaMalloc(&dout, sizeof(int) * count));
    CU(cudaMemcpyAsync(du, u, sizeof(double) * count, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(dx, x, sizeof(double) * count, cudaMemcpyHostToDevice, c->stream));
    if (n) CU(cudaMemcpyAsync(dn, n, sizeof(int) * count, cudaMemcpyHostToDevice, c->stream));
    else CU(cudaMemsetAsync(dn, 0, sizeof(int) * count, c->stream));
    CU(launch_metropolis_probe(kind, count, du, dx, dn, log(zzV), zzV, dout, c->stream));
    c->launches++;
    CU(cudaMemcpyAsync(out, dout, sizeof(int) * count, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(du); cudaFree(dx); cudaFree(dn); cudaFree(dout);
    return MCPM_OK;
}

int mcpm_philox_uniforms(mcpm_ctx* c, unsigned long long seed, unsigned int chain, unsigned long long first_step, int n_steps, double* out) {
    QUIESCE(c);
    if (!c || !out || n_steps <= 0) return fail(MCPM_ERR_ARG, "bad argument");
    CU(cudaSetDevice(c->device));
    double* d = nullptr;
    CU(cudaMalloc(&d, sizeof(double) * 8 * (size_t)n_steps));
    CU(launch_philox_fill(seed, chain, first_step, n_steps, d, c->stream));
    c->launches++;
    CU(cudaMemcpyAsync(out, d, sizeof(double) * 8 * (size_t)n_steps, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(d);
    return MCPM_OK;
}

}  // extern "C"
