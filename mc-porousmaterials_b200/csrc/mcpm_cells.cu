// K2c — persistent chains through a CELL LIST: large bulk-like systems whose box holds >= 3 cut-off lengths along every
// periodic axis (config C4: N = 20 000, L = 100 A, rcut = 12 A -> 8x8x8 cells of ~39 particles).  The reference scans all N
// partners per energy evaluation (src/potentials.cpp:55-73); every partner inside the cut-off lies in the 27 cells around the
// position, so ~1050 candidates are evaluated instead of 20 000 — same sums up to the order of the additions.
//
// One CTA of 8 warps per chain; positions stay in global memory (L2-resident) and are addressed through per-cell index
// lists items[cell][slot], with cell_of[j] / slot_of[j] for O(1) removal; the per-cell counts live in shared memory.
// A scan gives each warp up to four of the 27 neighbour cells and walks them in lock step (4, or 8 with old+new position,
// independent evaluations in flight per lane).  Accepted moves update the lists (thread 0) behind one extra barrier.
// The list is rebuilt — sorted by particle index inside each cell, so deterministically — at the start of every launch.
//
// Restates the same reference code as chain_loop in mcpm_kernels.cu (no energy cache): step loop src/main.cpp:66-71,
// MoveParticle src/moves.cpp:118-186, SwapParticle src/moves.cpp:284-341, ComputeWidom src/moves.cpp:410-449,
// ResetStats src/MC.cpp:81-98, AdjustMCMoves src/moves.cpp:68-78.
#include <algorithm>
#include <type_traits>

#include "mcpm_chain.cuh"
#include "mcpm_device.cuh"
#include "mcpm_kernels.h"

namespace mcpm {

namespace {

// Build parameters, measured on C4 (148 chains of N = 20 000, B200, round 2; moves/s with / without Widom sampling):
//   8 warps, 2 slots per trip, 1 CTA/SM (default)  2.94e7 / 5.31e7      12 warps 2.63e7      16 warps (128 registers) 2.29e7
//   1 slot per trip 3.03e7 / 5.26e7                __launch_bounds__(256, 2) with 296 chains: 2.2e7 / 4.2e7 (spills at 128 registers)
//   positions kept a second time in cell order (coalesced scans, no index gather): 2.69e7 / 4.36e7 — the extra stores of an accepted move
//   sit on the serial path.  The step is a chain of dependent latencies (4 L2 round trips) around an FP64-throughput-bound scan.
#ifndef MCPM_KC_WARPS
#define MCPM_KC_WARPS 8
#endif
#ifndef MCPM_KC_MINB
#define MCPM_KC_MINB 1
#endif
#ifndef MCPM_KC_H
#define MCPM_KC_H 2      // slots of each of the warp's cells per trip
#endif
constexpr int KC_WARPS = MCPM_KC_WARPS, KC_THREADS = 32 * KC_WARPS, KC_H = MCPM_KC_H;
constexpr int KC_Q = (27 + KC_WARPS - 1) / KC_WARPS;   // neighbour cells per warp
constexpr int KC_MAXCELLS = K3C_MAXC * K3C_MAXC * K3C_MAXC;

template <typename ItemT>   // unsigned short: lists in shared memory (capacity <= 65536 and they fit); int: lists in global memory
struct CellListT {
    ItemT* items;   // [ncell][ccap]: particle indices
    int* cell_of;   // global [cap]
    int* slot_of;   // global [cap]
    int* cnt;       // shared [ncell]
    int nx, ny, nz, ccap;
    double icx, icy, icz;   // cells per unit length
};

// Cell of a position, packed as cx | cy << 8 | cz << 16 (<= 16 cells per axis).  One multiply and one conversion per axis;
// the same expression bins the particles at build time, so a particle is always found in the cell its position maps to.
template <class CellList>
__device__ __forceinline__ int cell_index(const CellList& cl, const ChainParams&, double x, double y, double z) {
    const int cx = min(cl.nx - 1, max(0, (int)(x * cl.icx))), cy = min(cl.ny - 1, max(0, (int)(y * cl.icy))), cz = min(cl.nz - 1, max(0, (int)(z * cl.icz)));
    return cx | (cy << 8) | (cz << 16);
}
template <class CellList>
__device__ __forceinline__ int cell_linear(const CellList& cl, int packed) {
    return (((packed >> 16) & 0xff) * cl.ny + ((packed >> 8) & 0xff)) * cl.nx + (packed & 0xff);
}

// thread 0 only; false if the cell is full
template <class CellList>
__device__ __forceinline__ bool cell_add(const CellList& cl, int j, int c) {
    const int s = cl.cnt[c];
    if (s >= cl.ccap) return false;
    cl.cnt[c] = s + 1;
    cl.items[(size_t)c * cl.ccap + s] = j; cl.cell_of[j] = c; cl.slot_of[j] = s;
    return true;
}
template <class CellList>
__device__ __forceinline__ void cell_remove(const CellList& cl, int j) {
    const int c = cl.cell_of[j], s = cl.slot_of[j], last = cl.cnt[c] - 1;
    const int jl = cl.items[(size_t)c * cl.ccap + last];
    cl.items[(size_t)c * cl.ccap + s] = jl; cl.slot_of[jl] = s;
    cl.cnt[c] = last;
}
template <class CellList>
__device__ __forceinline__ void cell_rename(const CellList& cl, int from, int to) {
    const int c = cl.cell_of[from], s = cl.slot_of[from];
    cl.items[(size_t)c * cl.ccap + s] = to; cl.cell_of[to] = c; cl.slot_of[to] = s;
}

// Pair sums of position a (and b) over the 27 cells around cell c0.  nev counts the evaluations of this thread.
template <bool PBZ, bool TWO, class CellList>
__device__ __forceinline__ void cell_scan(const CellList& cl, const double* xs, const double* ys, const double* zs, const PairGeom& g, int c0,
                                          double xa, double ya, double za, double xb, double yb, double zb, int self, int warp, int lane,
                                          double& ea, double& eb, double& nev) {
    const int cx = c0 & 0xff, cy = (c0 >> 8) & 0xff, cz = (c0 >> 16) & 0xff;       // c0: packed cell coordinates
    int base[KC_Q], cnt[KC_Q], maxc = 0;
#pragma unroll
    for (int q = 0; q < KC_Q; q++) {
        const int k = warp + q * KC_WARPS;
        base[q] = 0; cnt[q] = 0;
        if (k < 27) {
            int z = cz + k / 9 - 1, y = cy + (k / 3) % 3 - 1, x = cx + k % 3 - 1;
            bool ok = true;
            if (PBZ) z = (z < 0) ? z + cl.nz : ((z >= cl.nz) ? z - cl.nz : z); else ok = (z >= 0 && z < cl.nz);
            y = (y < 0) ? y + cl.ny : ((y >= cl.ny) ? y - cl.ny : y);
            x = (x < 0) ? x + cl.nx : ((x >= cl.nx) ? x - cl.nx : x);
            if (ok) { const int c = (z * cl.ny + y) * cl.nx + x; cnt[q] = cl.cnt[c]; base[q] = c * cl.ccap; }
        }
        maxc = max(maxc, cnt[q]);
    }
    double a[KC_Q], b[KC_Q];
#pragma unroll
    for (int q = 0; q < KC_Q; q++) { a[q] = 0.0; b[q] = 0.0; }
    int ne = 0;
    // Two slots of each of the warp's four cells per trip: the 8 index loads and 24 position loads (L2, ~800 cycles) are all
    // in flight before the first evaluation needs them.
    for (int s = lane; s < maxc; s += 32 * KC_H) {
        int j[KC_H * KC_Q];
        bool on[KC_H * KC_Q];
        double xj[KC_H * KC_Q], yj[KC_H * KC_Q], zj[KC_H * KC_Q];
#pragma unroll
        for (int q = 0; q < KC_Q; q++)
#pragma unroll
            for (int h = 0; h < KC_H; h++) {
                on[KC_H * q + h] = (s + 32 * h) < cnt[q];
                j[KC_H * q + h] = on[KC_H * q + h] ? (int)cl.items[base[q] + s + 32 * h] : 0;
            }
#pragma unroll
        for (int t = 0; t < KC_H * KC_Q; t++) { xj[t] = xs[j[t]]; yj[t] = ys[j[t]]; zj[t] = zs[j[t]]; }
#pragma unroll
        for (int t = 0; t < KC_H * KC_Q; t++) {
            const int ex = on[t] ? self : j[t];              // inactive slots mask themselves out
            pair_acc(a[t / KC_H], dist2<true, PBZ>(g, xa, ya, za, xj[t], yj[t], zj[t]), g, j[t], ex);
            if (TWO) pair_acc(b[t / KC_H], dist2<true, PBZ>(g, xb, yb, zb, xj[t], yj[t], zj[t]), g, j[t], ex);
            ne += on[t] ? (TWO ? 2 : 1) : 0;
        }
    }
#pragma unroll
    for (int w = 1; w < KC_Q; w *= 2)
#pragma unroll
        for (int q = 0; q + w < KC_Q; q += 2 * w) { a[q] += a[q + w]; b[q] += b[q + w]; }   // (a0+a1)+(a2+a3)
    ea += a[0];
    if (TWO) eb += b[0];
    nev += (double)ne;
}

// Bulk boxes have neither walls nor an out-of-pore guard (src/energy.cpp:53-77): skip the lane-parallel wall evaluation.
__device__ __forceinline__ void wall_any(const ChainParams& P, double z_a, double z_b, double& w_a, double& w_b, int lane) {
    if (P.geometry == MCPM_GEOM_SLIT) wall_two(P, z_a, z_b, w_a, w_b, lane);
    else { w_a = 0.0; w_b = 0.0; }
}

// All-reduce of three per-thread values over the 8 warps; every thread returns the same bits.
__device__ __forceinline__ void cta_allsum3(double (&v)[3], double* partials, int parity, int warp, int lane) {
#pragma unroll
    for (int k = 0; k < 3; k++) v[k] = warp_allsum(v[k]);
    double* p = partials + parity * 64;
    if (lane == 0) { p[warp * 4] = v[0]; p[warp * 4 + 1] = v[1]; p[warp * 4 + 2] = v[2]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double s = 0.0;
        for (int w = 0; w < KC_WARPS; w++) s += p[w * 4 + k];
        v[k] = s;
    }
}

// ComputeWidom's exponentials (src/moves.cpp:425-447) are not needed by the chain itself: thread 0 only queues the sampled
// energy and warp 0 turns 32 queued samples into BAR terms at once, one per lane.
struct WidomQueue {
    double energy[32];
    int kind[32];      // 0 insertion, 1 deletion
    int count;
};
__device__ __forceinline__ void widom_flush(WidomQueue& wq, mcpm_chain_stats& st, double beta, int lane) {   // warp 0, all lanes
    const int m = wq.count;
    double ins = 0.0, del = 0.0;
    if (lane < m) {
        const double energy = wq.energy[lane];
        const double bw = bar_weight(energy, beta);
        if (wq.kind[lane] == 0) ins = bw * exp(-0.5 * energy * beta);
        else if (exp(0.5 * energy * beta) < 1) del = bw * exp(0.5 * energy * beta);
        else del = bw * exp(-0.5 * energy * beta);   // the reference's boundary for deletion statistics
    }
    ins = warp_allsum(ins); del = warp_allsum(del);
    if (lane == 0) { st.widom_insert += ins; st.widom_delete += del; wq.count = 0; }
    __syncwarp();
}

template <class Source, bool PBZ, bool SAMPLE, class CellList>
__device__ void chain_loop_cells(const ChainParams& P, ChainState& S, Source& rng, double* xs, double* ys, double* zs, double* partials,
                                 WidomQueue& wq, const CellList& cl, const RunArgs& a, unsigned char* trace, int cap, uint32_t chain) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool lead = (tid == 0);
    const PairGeom g = make_geom<true>(P);
    mcpm_chain_stats& st = S.st;   // shared memory; written by thread 0 only
    const bool widom_on = SAMPLE && (P.n_swap == 0) && !a.no_sampling;   // ComputeWidom acts only without swap moves, src/moves.cpp:415

    int n = st.n;
    double dr = st.dr, e_pair = st.pair, e_wall = st.wall;
    unsigned long long step = st.steps_done;
    int overflow = 0;
    const double wsum = P.wsum, thr_disp = P.thr_disp, thr_vol = P.thr_vol;
    const double beta = P.beta, eps4 = P.eps4, ln_zzV = P.ln_zzV, zzV = P.zzV;
    const double Lx = P.L[0], Ly = P.L[1], Lz = P.L[2];
    int parity = 0;
    SetCounters sc;
    const int steps_per_set = (int)a.steps_per_set;

    for (int set = a.first_set; set < a.first_set + a.n_sets && !overflow; set++) {
        sc.reset();                      // MC::ResetStats, src/MC.cpp:81-98
        if (lead) { st.widom_insert = st.widom_delete = 0.0; st.widom_insertions = st.widom_deletions = 1; }
        const bool production = set > P.n_equil;
        for (int s = 0; s < steps_per_set; s++) {
            rng.begin_step(a.seed, S.stream_id, step, lane);
            step++;
            int code;
            double sx = 0., sy = 0., sz = 0., s_po = 0., s_wo = 0., mx = 0., my = 0., mz = 0.;   // this step's MoveParticle, for Widom's slot 0
            bool s_acc = false;
            const double r = rng.template at<0>() * wsum;                         // src/main.cpp:67
            if (r <= thr_disp) {
                // ---------------- MoveParticle, src/moves.cpp:118-186 ----------------
                const double u_i = rng.template at<3>(), u_x = rng.template at<4>(), u_y = rng.template at<5>(),
                             u_z = rng.template at<6>(), u_a = rng.template at<7>();
                rng.end_step(8);
                const int i = (int)(u_i * (double)n);                             // n == 0: the stale slot 0 (quirk Q12)
                const double xo = xs[i], yo = ys[i], zo = zs[i];
                const double xr = __dadd_rn(xo, __dmul_rn(u_x - 0.5, dr)), yr = __dadd_rn(yo, __dmul_rn(u_y - 0.5, dr));
                const double zr = __dadd_rn(zo, __dmul_rn(u_z - 0.5, dr));
                bool odd = false;                                                 // MC::PBC: x and y are always periodic
                double xn = wrap_select(xr, Lx, odd), yn = wrap_select(yr, Ly, odd), zn = PBZ ? wrap_select(zr, Lz, odd) : zr;
                if (odd) { xn = wrap_floor(xr, Lx); yn = wrap_floor(yr, Ly); if (PBZ) zn = wrap_floor(zr, Lz); }
                const int co = cell_index(cl, P, xo, yo, zo), cn = cell_index(cl, P, xn, yn, zn);
                double v[3] = {0.0, 0.0, 0.0};
                if (co == cn) cell_scan<PBZ, true>(cl, xs, ys, zs, g, co, xo, yo, zo, xn, yn, zn, i, warp, lane, v[0], v[1], v[2]);
                else {
                    double dummy = 0.0;
                    cell_scan<PBZ, false>(cl, xs, ys, zs, g, co, xo, yo, zo, 0., 0., 0., i, warp, lane, v[0], dummy, v[2]);
                    cell_scan<PBZ, false>(cl, xs, ys, zs, g, cn, xn, yn, zn, 0., 0., 0., i, warp, lane, v[1], dummy, v[2]);
                }
                double wo, wn;
                wall_any(P, zo, zn, wo, wn, lane);
                cta_allsum3(v, partials, parity, warp, lane); parity ^= 1;
                sc.evals += (unsigned long long)v[2];
                const double po = eps4 * v[0], pn = eps4 * v[1];
                const double dE = (pn + wn) - (po + wo);                          // newEnergy - oldEnergy, src/moves.cpp:169
                const bool acc = metropolis_translation(u_a, dE * beta);          // u < exp(-dE/T), src/moves.cpp:170-172
                sc.n_disp++; sc.alg += 2ull * (unsigned long long)n;
                if (acc) {
                    sc.acc_disp++;
                    bool full = false;
                    if (n > 0) { e_pair += pn - po; e_wall += wn - wo; }          // full pair change (Q16); n == 0 moves a ghost (Q12)
                    if (lead) {
                        xs[i] = xn; ys[i] = yn; zs[i] = zn;
                        if (n > 0 && co != cn) { cell_remove(cl, i); full = !cell_add(cl, i, cell_linear(cl, cn)); }
                    }
                    if (__syncthreads_or(full ? 1 : 0)) { overflow = 2; break; }  // a cell ran out of slots
                }
                code = acc ? 1 : 0;
                if (SAMPLE) { sx = xo; sy = yo; sz = zo; s_po = po; s_wo = wo; mx = xn; my = yn; mz = zn; s_acc = acc; }
            } else if (r <= thr_vol) {
                rng.end_step(1);
                code = (4 << 1);  // volume moves are out of scope (n_vol is rejected by the host API)
            } else {
                // ---------------- SwapParticle, single box, src/moves.cpp:284-341 ----------------
                if (rng.template at<2>() < 0.5) {
                    if (n >= cap) { overflow = 1; break; }
                    const double u_x = rng.template at<3>(), u_y = rng.template at<4>(), u_z = rng.template at<5>(), u_a = rng.template at<6>();
                    rng.end_step(7);
                    const double xi = u_x * Lx, yi = u_y * Ly, zi = u_z * Lz;     // InsertParticle, src/moves.cpp:46-54
                    // the trial position is written to the first free slot even if rejected (src/moves.cpp:302-303); it is in no cell
                    if (lead) { xs[n] = xi; ys[n] = yi; zs[n] = zi; }
                    const int ci = cell_index(cl, P, xi, yi, zi);
                    double v[3] = {0.0, 0.0, 0.0}, dummy = 0.0;
                    cell_scan<PBZ, false>(cl, xs, ys, zs, g, ci, xi, yi, zi, 0., 0., 0., -1, warp, lane, v[0], dummy, v[2]);
                    double wi, wdummy;
                    wall_any(P, zi, zi, wi, wdummy, lane);
                    cta_allsum3(v, partials, parity, warp, lane); parity ^= 1;
                    sc.evals += (unsigned long long)v[2];
                    const double pi_ = eps4 * v[0];
                    const bool acc = metropolis_insertion(u_a, -((pi_ + wi) * beta), ln_zzV, zzV, n);   // src/moves.cpp:306-307
                    sc.n_ins++; sc.alg += (unsigned long long)n;
                    if (acc) {
                        bool full = false;
                        if (lead) full = !cell_add(cl, n, cell_linear(cl, ci));
                        sc.acc_ins++; n++; e_pair += pi_; e_wall += wi;
                        if (__syncthreads_or(full ? 1 : 0)) { overflow = 2; break; }
                    }
                    code = (1 << 1) | (acc ? 1 : 0);
                } else if (n > 0) {
                    const double u_o = rng.template at<3>(), u_a = rng.template at<4>();
                    rng.end_step(5);
                    const int o = (int)(u_o * (double)n + 1.0) - 1;               // int(u*n+1) on 1-based slots
                    const double xo = xs[o], yo = ys[o], zo = zs[o];
                    const int co = cell_index(cl, P, xo, yo, zo);
                    double v[3] = {0.0, 0.0, 0.0}, dummy = 0.0;
                    cell_scan<PBZ, false>(cl, xs, ys, zs, g, co, xo, yo, zo, 0., 0., 0., o, warp, lane, v[0], dummy, v[2]);
                    double wo, wdummy;
                    wall_any(P, zo, zo, wo, wdummy, lane);
                    cta_allsum3(v, partials, parity, warp, lane); parity ^= 1;
                    sc.evals += (unsigned long long)v[2];
                    const double po = eps4 * v[0];
                    const bool acc = metropolis_deletion(u_a, (po + wo) * beta, ln_zzV, zzV, n);        // src/moves.cpp:326-327
                    sc.n_del++; sc.alg += (unsigned long long)n;
                    if (acc) {
                        sc.acc_del++;
                        if (lead) {                                                // swap-with-last, src/moves.cpp:329
                            cell_remove(cl, o);
                            if (o != n - 1) { cell_rename(cl, n - 1, o); xs[o] = xs[n - 1]; ys[o] = ys[n - 1]; zs[o] = zs[n - 1]; }
                        }
                        n--;
                        e_pair -= po; e_wall -= wo;   // exact bookkeeping with the fresh energies (quirks Q3, Q16 fixed)
                        __syncthreads();
                    }
                    code = (2 << 1) | (acc ? 1 : 0);
                } else {
                    rng.end_step(3);
                    sc.n_empty++;
                    code = (3 << 1);
                }
            }
            if (SAMPLE && production && widom_on) {
                // ---------------- ComputeWidom, src/moves.cpp:410-449 (NVT branch) ----------------
                rng.load_extra(a.seed, S.stream_id, step - 1, lane);
                double energy;                                                   // draw 0 (species) is consumed but unused
                if (rng.template extra<1>() < 0.5) {                             // ghost insertion
                    const double gx = rng.template extra<2>() * Lx, gy = rng.template extra<3>() * Ly, gz = rng.template extra<4>() * Lz;
                    rng.end_step(5);
                    double v[3] = {0.0, 0.0, 0.0}, dummy = 0.0;
                    cell_scan<PBZ, false>(cl, xs, ys, zs, g, cell_index(cl, P, gx, gy, gz), gx, gy, gz, 0., 0., 0., -1, warp, lane, v[0], dummy, v[2]);
                    double wg, wdummy;
                    wall_any(P, gz, gz, wg, wdummy, lane);
                    cta_allsum3(v, partials, parity, warp, lane); parity ^= 1;
                    energy = eps4 * v[0] + wg;
                    if (lead) { st.widom_insertions++; wq.energy[wq.count] = energy; wq.kind[wq.count] = 0; wq.count++; st.pair_evals += (unsigned long long)v[2]; st.pair_evals_algorithmic += (unsigned long long)n; }
                } else {                                                         // virtual deletion
                    const int slot1 = (int)(rng.template extra<2>() * (double)n);   // 1-based slot in [0, n-1]: 0 is the scratch slot (Q14)
                    rng.end_step(3);
                    if (slot1 == 0) {
                        // slot 0 holds the old position of this step's MoveParticle (src/moves.cpp:152): see chain_loop
                        double extra = 0.0;
                        if (s_acc) pair_acc(extra, dist2<true, PBZ>(g, sx, sy, sz, mx, my, mz), g, 0, -1);
                        energy = (s_po + eps4 * extra) + s_wo;
                    } else {
                        const int k = slot1 - 1;
                        const double xk = xs[k], yk = ys[k], zk = zs[k];
                        double v[3] = {0.0, 0.0, 0.0}, dummy = 0.0;
                        cell_scan<PBZ, false>(cl, xs, ys, zs, g, cell_index(cl, P, xk, yk, zk), xk, yk, zk, 0., 0., 0., k, warp, lane, v[0], dummy, v[2]);
                        double wk, wdummy;
                        wall_any(P, zk, zk, wk, wdummy, lane);
                        cta_allsum3(v, partials, parity, warp, lane); parity ^= 1;
                        if (lead) st.pair_evals += (unsigned long long)v[2];
                        energy = eps4 * v[0] + wk;
                    }
                    if (lead) {
                        st.pair_evals_algorithmic += (unsigned long long)n;
                        st.widom_deletions++;
                        wq.energy[wq.count] = energy; wq.kind[wq.count] = 1; wq.count++;
                    }
                }
                if (warp == 0) {                     // (the queue belongs to warp 0: no barrier needed)
                    __syncwarp();
                    if (wq.count == 32) widom_flush(wq, st, beta, lane);
                }
            }
            if (lead) {
                if (trace) trace[(unsigned long long)(set - a.first_set) * (unsigned long long)a.steps_per_set + (unsigned long long)s] = (unsigned char)code;
                if (production) {
                    const double dn = (double)n, e = e_pair + e_wall;
                    st.sum_n += dn; st.sum_n2 += dn * dn; st.sum_e += e; st.sum_e2 += e * e;
                    st.samples++;
                }
            }
        }
        if (SAMPLE && warp == 0) { __syncwarp(); if (wq.count > 0) widom_flush(wq, st, beta, lane); }   // the sums are per set (ResetStats)
        if (lead) st.dr_used = dr;          // the amplitude this set ran with (statistics are printed before the adjustment, src/main.cpp:77-83)
        // MC::AdjustMCMoves displacement part, src/moves.cpp:68-78 (ResetStats starts nDisplacements at ONE)
        if (!overflow && n > 0 && set <= P.n_equil) {
            const double ratio = (double)sc.acc_disp * 1. / (1. * (double)(sc.n_disp + 1));
            if (ratio < 0.2) dr /= 1.1;
            else dr *= 1.1;
        }
        if (lead) {
            st.acceptance = sc.acc_disp; st.rejection = sc.n_disp - sc.acc_disp; st.n_displacements = 1 + sc.n_disp;
            st.accept_swap = sc.acc_ins + sc.acc_del; st.reject_swap = (sc.n_ins + sc.n_del) - (sc.acc_ins + sc.acc_del);
            st.n_swaps = 1 + sc.n_ins + sc.n_del;
            st.tot_disp += sc.n_disp; st.tot_disp_acc += sc.acc_disp; st.tot_ins += sc.n_ins; st.tot_ins_acc += sc.acc_ins;
            st.tot_del += sc.n_del; st.tot_del_acc += sc.acc_del; st.tot_empty += sc.n_empty; st.pair_evals += sc.evals; st.pair_evals_algorithmic += sc.alg;
        }
    }
    __syncthreads();
    if (lead) {
        st.n = n; st.overflow = overflow; st.dr = dr; st.steps_done = step;
        st.pair = e_pair; st.wall = e_wall; st.energy = e_pair + e_wall;
        S.replay_pos = rng.cursor();
        st.replay_consumed = rng.cursor();
        S.cache_valid = 0;             // this kernel does not maintain the energy cache
        S.cache_sum = 0ull;
    }
}

}  // namespace

// SMEM_ITEMS: the index lists (16-bit entries) live in dynamic shared memory; else 32-bit entries in global memory.
// (round 2: __launch_bounds__(KC_THREADS, 2) — two chains per SM at 128 registers — spills 450 bytes and was measured 22-37 % slower)
template <int RNG_MODE, bool SAMPLE, bool SMEM_ITEMS>
__global__ void __launch_bounds__(KC_THREADS, MCPM_KC_MINB) k2_cells(const ChainParams* __restrict__ params, ChainState* __restrict__ states,
                                                          const int* __restrict__ order, double* X, double* Y, double* Z, RunArgs a, int* items,
                                                          int* cell_of, int* slot_of, int cells_stride, int ccap) {
    typedef typename std::conditional<SMEM_ITEMS, unsigned short, int>::type ItemT;
    extern __shared__ unsigned short dyn_items[];
    __shared__ ChainParams P;
    __shared__ ChainState S;
    __shared__ int cnt[KC_MAXCELLS];
    __shared__ double partials[128];
    __shared__ WidomQueue wq;
    __shared__ int build_overflow;
    const int chain = order[blockIdx.x];
    const int cap = a.cap;
    const size_t off = (size_t)chain * (size_t)cap;
    double* xs = X + off;
    double* ys = Y + off;
    double* zs = Z + off;
    const int tid = threadIdx.x, T = blockDim.x;
    if (tid == 0) { P = params[chain]; S = states[chain]; build_overflow = 0; wq.count = 0; }
    __syncthreads();
    const CellGrid cg = cell_grid(P);
    CellListT<ItemT> cl;
    if (SMEM_ITEMS) cl.items = reinterpret_cast<ItemT*>(dyn_items);
    else cl.items = reinterpret_cast<ItemT*>(items + (size_t)chain * (size_t)cells_stride * (size_t)ccap);
    cl.cell_of = cell_of + off; cl.slot_of = slot_of + off; cl.cnt = cnt;
    cl.nx = cg.nx; cl.ny = cg.ny; cl.nz = cg.nz; cl.ccap = ccap;
    cl.icx = (double)cg.nx / P.L[0]; cl.icy = (double)cg.ny / P.L[1]; cl.icz = (double)cg.nz / P.L[2];
    const int ncell = cg.nx * cg.ny * cg.nz, n0 = S.st.n;
    // ---- build: bin with shared-memory atomics, then sort every cell by particle index (deterministic order) ----
    for (int c = tid; c < ncell; c += T) cnt[c] = 0;
    __syncthreads();
    for (int j = tid; j < n0; j += T) {
        const int c = cell_linear(cl, cell_index(cl, P, xs[j], ys[j], zs[j]));
        const int s = atomicAdd(&cnt[c], 1);
        if (s < ccap) cl.items[(size_t)c * ccap + s] = (ItemT)j; else build_overflow = 1;
    }
    __syncthreads();
    if (build_overflow) {
        if (tid == 0) { S.st.overflow = 2; states[chain] = S; }
        return;
    }
    for (int c = tid; c < ncell; c += T) {
        ItemT* it = cl.items + (size_t)c * ccap;
        const int m = cnt[c];
        for (int p = 1; p < m; p++) {          // insertion sort: ~40 entries
            const ItemT key = it[p];
            int q = p - 1;
            while (q >= 0 && it[q] > key) { it[q + 1] = it[q]; q--; }
            it[q + 1] = key;
        }
        for (int s = 0; s < m; s++) { cl.cell_of[it[s]] = c; cl.slot_of[it[s]] = s; }
    }
    __syncthreads();

    unsigned char* trace = a.trace ? a.trace + (unsigned long long)chain * a.trace_stride : nullptr;
    const bool pbz = P.pbc[2] != 0;
    if (RNG_MODE == MCPM_RNG_PHILOX) {
        PhiloxSource rng; rng.init(a.seed, S.stream_id, S.st.steps_done, tid & 31);
        if (pbz) chain_loop_cells<PhiloxSource, true, SAMPLE>(P, S, rng, xs, ys, zs, partials, wq, cl, a, trace, cap, (uint32_t)chain);
        else chain_loop_cells<PhiloxSource, false, SAMPLE>(P, S, rng, xs, ys, zs, partials, wq, cl, a, trace, cap, (uint32_t)chain);
    } else {
        ReplaySource rng; rng.init(a.replay_u + (unsigned long long)chain * a.replay_stride, 0ull, a.replay_stride);
        if (pbz) chain_loop_cells<ReplaySource, true, SAMPLE>(P, S, rng, xs, ys, zs, partials, wq, cl, a, trace, cap, (uint32_t)chain);
        else chain_loop_cells<ReplaySource, false, SAMPLE>(P, S, rng, xs, ys, zs, partials, wq, cl, a, trace, cap, (uint32_t)chain);
    }
    __syncthreads();
    if (tid == 0) states[chain] = S;
}

// Dynamic shared memory of the 16-bit in-kernel lists, or 0 if they do not fit / cannot index the capacity (-> global lists).
size_t k2_cells_smem_items_bytes(int cap, int cells_stride, int ccap, int max_smem_optin) {
    const size_t need = (size_t)cells_stride * ccap * sizeof(unsigned short);
    return (cap <= 65536 && need + 20 * 1024 <= (size_t)max_smem_optin) ? need : 0;
}

cudaError_t launch_k2_cells(int rng_mode, bool sample, int n_chains, const ChainParams* params, ChainState* states, const int* order, double* X,
                            double* Y, double* Z, const RunArgs& a, int* items, int* cell_of, int* slot_of, int cells_stride, int ccap,
                            size_t smem_items, cudaStream_t stream) {
    auto go = [&](auto kern) -> cudaError_t {
        if (smem_items) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_items);
            if (e != cudaSuccess) return e;
        }
        kern<<<n_chains, KC_THREADS, smem_items, stream>>>(params, states, order, X, Y, Z, a, items, cell_of, slot_of, cells_stride, ccap);
        return cudaGetLastError();
    };
    const bool ph = rng_mode == MCPM_RNG_PHILOX;
    if (smem_items) {
        if (ph) return sample ? go(k2_cells<MCPM_RNG_PHILOX, true, true>) : go(k2_cells<MCPM_RNG_PHILOX, false, true>);
        return sample ? go(k2_cells<MCPM_RNG_REPLAY, true, true>) : go(k2_cells<MCPM_RNG_REPLAY, false, true>);
    }
    if (ph) return sample ? go(k2_cells<MCPM_RNG_PHILOX, true, false>) : go(k2_cells<MCPM_RNG_PHILOX, false, false>);
    return sample ? go(k2_cells<MCPM_RNG_REPLAY, true, false>) : go(k2_cells<MCPM_RNG_REPLAY, false, false>);
}

}  // namespace mcpm
