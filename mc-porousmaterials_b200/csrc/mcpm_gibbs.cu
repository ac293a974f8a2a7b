// K2g — k2_gibbs: the two-box (Gibbs ensemble) branches of the reference's moves, one CTA per PAIR of boxes.
//
//   MoveParticle with thermoSys.nBoxes == 2   src/moves.cpp:118-186 (box selection :127-137)
//   SwapParticle, Gibbs branch                src/moves.cpp:342-400 (particle transfer between the boxes, Widom/BAR sums of both)
//   ResetStats / AdjustMCMoves on both boxes  src/MC.cpp:81-98, src/moves.cpp:68-78
//
// ChangeVolume's Gibbs branches are not here: the stock build of the reference aborts on entering ChangeVolume (quirk Q19), so
// two boxes at fixed volumes exchanging particles is what the reference can run; the host refuses n_vol > 0 for linked boxes.
//
// Row f4 of SURVEY.md §8 — a correctness path, not a hot one: both boxes stay in global memory (L2), every box goes through the
// GENERAL helpers (any geometry / wall the engine knows: a slit pore exchanging with a bulk reservoir is the typical use), scalars
// are replicated in every thread as in chain_loop, and a step ends with a barrier instead of register forwarding.
#include "mcpm_chain.cuh"
#include "mcpm_device.cuh"
#include "mcpm_kernels.h"

namespace mcpm {

namespace {

constexpr int GIBBS_THREADS = 128;

// MC::InsertParticle for the box's geometry, src/moves.cpp:22-54
__device__ __forceinline__ void gibbs_insert_position(const ChainParams& P, double u_x, double u_y, double u_z, double& xi, double& yi, double& zi) {
    const double Lx = P.L[0], Ly = P.L[1], Lz = P.L[2];
    const double PI = 3.14159265358979323846;
    if (P.geometry == MCPM_GEOM_SPHERE) {                // radius, theta, phi
        const double rad = u_x * Lz * 0.5, th = u_y * PI, ph = u_z * 2 * PI;
        xi = rad * sin(th) * cos(ph) + 0.5 * Lx; yi = rad * sin(th) * sin(ph) + 0.5 * Ly; zi = rad * cos(th) + 0.5 * Lz;
    } else if (P.geometry == MCPM_GEOM_CYLINDER) {       // rho, phi, then x
        const double rho = u_x * Lz * 0.5, ph = u_y * 2 * PI;
        xi = (u_z - 0.5) * Lx + 0.5 * Lx; yi = rho * sin(ph) + 0.5 * Ly; zi = rho * cos(ph) + 0.5 * Lz;
    } else { xi = u_x * Lx; yi = u_y * Ly; zi = u_z * Lz; }
}

template <class Source>
__device__ void gibbs_loop(const ChainParams* P /* [2], shared */, ChainState* S /* [2], shared */, Source& rng, double* const xs[2], double* const ys[2],
                           double* const zs[2], double* partials, const RunArgs& a, unsigned char* trace, int cap) {
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, W = T >> 5;
    const bool lead = (tid == 0);
    const uint32_t sid = S[0].stream_id;                  // one uniform stream per SYSTEM: box 0's
    PairGeom g[2] = {make_geom<false>(P[0]), make_geom<false>(P[1])};
    int n[2] = {S[0].st.n, S[1].st.n};
    double dr[2] = {S[0].st.dr, S[1].st.dr}, e_pair[2] = {S[0].st.pair, S[1].st.pair}, e_wall[2] = {S[0].st.wall, S[1].st.wall};
    const double vol[2] = {P[0].volume, P[1].volume};
    const double wsum = P[0].wsum, thr_disp = P[0].thr_disp, thr_vol = P[0].thr_vol, beta = P[0].beta, temp = P[0].T;
    unsigned long long step = S[0].st.steps_done;
    int overflow = 0, parity = 0;
    const int steps_per_set = (int)a.steps_per_set;

    for (int set = a.first_set; set < a.first_set + a.n_sets && !overflow; set++) {
        // MC::ResetStats
        int n_disp[2] = {0, 0}, acc_disp[2] = {0, 0}, n_in[2] = {0, 0}, acc_in[2] = {0, 0}, n_empty = 0;
        unsigned long long evals = 0ull;
        if (lead) for (int b = 0; b < 2; b++) { S[b].st.widom_insert = S[b].st.widom_delete = 0.0; S[b].st.widom_insertions = S[b].st.widom_deletions = 1; }
        const bool production = set > P[0].n_equil;
        for (int s = 0; s < steps_per_set; s++) {
            rng.begin_step(a.seed, sid, step, lane);
            step++;
            int code;
            const double r = rng.template at<0>() * wsum;                        // src/main.cpp:67
            if (r <= thr_disp) {
                // ---------------- MoveParticle ----------------
                const double u_b = rng.template at<1>(), u_i = rng.template at<3>(), u_x = rng.template at<4>(), u_y = rng.template at<5>(),
                             u_z = rng.template at<6>(), u_a = rng.template at<7>();   // draw 2: species
                rng.end_step(8);
                const double rnd = (double)(int)(u_b * (double)(n[0] + n[1]));   // :127: rand = int(Random()*thermoSys.nParts)
                const int b = (rnd <= (double)n[0]) ? 0 : ((rnd <= (double)(n[0] + n[1])) ? 1 : 0);   // first box with rand <= cumulative count
                const ChainParams& Pb = P[b];
                const int i = (int)(u_i * (double)n[b]);                         // n == 0: the stale first slot (Q12)
                const double xo = xs[b][i], yo = ys[b][i], zo = zs[b][i];
                const double xr = __dadd_rn(xo, __dmul_rn(u_x - 0.5, dr[b])), yr = __dadd_rn(yo, __dmul_rn(u_y - 0.5, dr[b]));
                const double zr = __dadd_rn(zo, __dmul_rn(u_z - 0.5, dr[b]));
                const double xn = g[b].px ? wrap_floor(xr, Pb.L[0]) : xr, yn = g[b].py ? wrap_floor(yr, Pb.L[1]) : yr,
                             zn = g[b].pz ? wrap_floor(zr, Pb.L[2]) : zr;        // MC::PBC: periodic axes only
                double v[2];
                pair_sum2<false, false>(xs[b], ys[b], zs[b], n[b], i, g[b], xo, yo, zo, xn, yn, zn, tid, T, v[0], v[1]);
                double wo, wn;
                wall_pair<false>(Pb, xo, yo, zo, xn, yn, zn, wo, wn, lane);
                cta_allsum<2, true>(v, partials, parity, W, warp, lane); parity ^= 1;
                evals += 2ull * (unsigned long long)n[b];
                const double po = Pb.eps4 * v[0], pn = Pb.eps4 * v[1];
                const double dE = (pn + wn) - (po + wo);
                const bool acc = metropolis_translation(u_a, dE * beta);         // :170-172
                n_disp[b]++;
                if (acc) {
                    acc_disp[b]++;
                    if (n[b] > 0) { e_pair[b] += pn - po; e_wall[b] += wn - wo; }   // exact bookkeeping (Q16 fixed); n == 0 moves a ghost
                    if (tid == (i & (T - 1))) { xs[b][i] = xn; ys[b][i] = yn; zs[b][i] = zn; }
                }
                code = (acc ? 1 : 0) | (b << 4);
            } else if (r <= thr_vol) {
                rng.end_step(1);
                code = (4 << 1);                                                 // (refused by the host: never reached)
            } else {
                // ---------------- SwapParticle, Gibbs branch: src/moves.cpp:342-400 ----------------
                const int in = (rng.template at<1>() < 0.5) ? 0 : 1, out = 1 - in;
                if (n[out] <= 0) {
                    rng.end_step(2);
                    n_empty++;
                    code = (6 << 1) | (in << 4);
                } else if (n[in] >= cap) {
                    overflow = 1; step--;
                    break;
                } else {
                    const double u_x = rng.template at<3>(), u_y = rng.template at<4>(), u_z = rng.template at<5>(), u_o = rng.template at<6>(),
                                 u_a = rng.template at<7>();                      // draw 2: species
                    rng.end_step(8);
                    double xi, yi, zi;
                    gibbs_insert_position(P[in], u_x, u_y, u_z, xi, yi, zi);
                    const int slot = n[in];
                    if (tid == (slot & (T - 1))) { xs[in][slot] = xi; ys[in][slot] = yi; zs[in][slot] = zi; }   // written even if rejected, :365
                    const int o = (int)(u_o * (double)n[out] + 1.0) - 1;          // :370
                    const double xo = xs[out][o], yo = ys[out][o], zo = zs[out][o];
                    double v[2];
                    v[0] = pair_sum1<false, false>(xs[in], ys[in], zs[in], n[in], -1, g[in], xi, yi, zi, tid, T);
                    v[1] = pair_sum1<false, false>(xs[out], ys[out], zs[out], n[out], o, g[out], xo, yo, zo, tid, T);
                    const double w_in = wall_any(P[in], xi, yi, zi), w_out = wall_any(P[out], xo, yo, zo);
                    cta_allsum<2, true>(v, partials, parity, W, warp, lane); parity ^= 1;
                    evals += (unsigned long long)(n[in] + n[out]);
                    const double p_in = P[in].eps4 * v[0], p_out = P[out].eps4 * v[1];
                    const double e_in = p_in + w_in, e_out = p_out + w_out;
                    if (lead) {   // Widom / BAR sums of the transfer, :368-374 (barC == 1: ResetStats)
                        S[in].st.widom_insert += bar_weight(e_in, beta) * vol[in] * exp(-0.5 * e_in / temp) / (double)(n[in] + 1);
                        S[in].st.widom_insertions++;
                        S[out].st.widom_delete += bar_weight(e_out, beta) * vol[out] * exp(0.5 * e_out / temp) / (double)(n[out] + 1);
                        S[out].st.widom_deletions++;
                    }
                    const double arg = (e_in - e_out) + log(vol[out] * (double)(n[in] + 1) / (vol[in] * (double)n[out])) * temp;   // :376
                    const bool acc = u_a < exp(-arg / temp);
                    n_in[in]++;
                    if (acc) {
                        acc_in[in]++;
                        const int last = n[out] - 1;
                        if (tid == (o & (T - 1))) { xs[out][o] = xs[out][last]; ys[out][o] = ys[out][last]; zs[out][o] = zs[out][last]; }   // :388
                        n[in]++; n[out]--;
                        e_pair[in] += p_in; e_wall[in] += w_in;                  // exact bookkeeping with the fresh, un-halved energies
                        e_pair[out] -= p_out; e_wall[out] -= w_out;
                    }
                    code = (5 << 1) | (acc ? 1 : 0) | (in << 4);
                }
            }
            __syncthreads();                                                     // this step's slot writes are visible to the next selection
            if (lead && trace) trace[(unsigned long long)(set - a.first_set) * (unsigned long long)a.steps_per_set + (unsigned long long)s] = (unsigned char)code;
            if (production && lead) {
                for (int b = 0; b < 2; b++) {
                    const double dn = (double)n[b], e = e_pair[b] + e_wall[b];
                    S[b].st.sum_n += dn; S[b].st.sum_n2 += dn * dn; S[b].st.sum_e += e; S[b].st.sum_e2 += e * e; S[b].st.samples++;
                }
            }
        }
        // MC::AdjustMCMoves, displacement part, per box (ResetStats starts nDisplacements at ONE)
        for (int b = 0; b < 2; b++) {
            if (lead) S[b].st.dr_used = dr[b];
            if (!overflow && n[b] > 0 && set <= P[0].n_equil) {
                const double ratio = (double)acc_disp[b] * 1. / (1. * (double)(n_disp[b] + 1));
                if (ratio < 0.2) dr[b] /= 1.1;
                else dr[b] *= 1.1;
            }
        }
        if (lead) {
            const int att = n_in[0] + n_in[1], ok = acc_in[0] + acc_in[1];
            for (int b = 0; b < 2; b++) {
                mcpm_chain_stats& st = S[b].st;
                st.acceptance = acc_disp[b]; st.rejection = n_disp[b] - acc_disp[b]; st.n_displacements = 1 + n_disp[b];
                st.accept_swap = ok; st.reject_swap = att - ok; st.n_swaps = 1 + att + n_empty;   // Stats::nSwaps etc. are global to the system
                st.tot_disp += n_disp[b]; st.tot_disp_acc += acc_disp[b];
                st.tot_ins += n_in[b]; st.tot_ins_acc += acc_in[b];            // transfers INTO this box ...
                st.tot_del += n_in[1 - b]; st.tot_del_acc += acc_in[1 - b];    // ... and out of it
                st.tot_empty += n_empty;
            }
            S[0].st.pair_evals += evals; S[0].st.pair_evals_algorithmic += evals;
        }
    }
    __syncthreads();
    double c_pair[2] = {0., 0.}, c_wall[2] = {0., 0.};
    if (a.check_energy)
        for (int b = 0; b < 2; b++) cta_box_energy<false, false, true>(xs[b], ys[b], zs[b], n[b], P[b], g[b], partials, W, warp, lane, tid, T, c_pair[b], c_wall[b]);
    if (lead) {
        for (int b = 0; b < 2; b++) {
            mcpm_chain_stats& st = S[b].st;
            st.n = n[b]; st.overflow = overflow; st.dr = dr[b]; st.steps_done = step;
            st.pair = e_pair[b]; st.wall = e_wall[b]; st.energy = e_pair[b] + e_wall[b];
            if (a.check_energy) { st.pair_check = c_pair[b]; st.wall_check = c_wall[b]; st.energy_check = c_pair[b] + c_wall[b]; }
            if (a.check_energy == 2) { st.pair = c_pair[b]; st.wall = c_wall[b]; st.energy = c_pair[b] + c_wall[b]; }
            S[b].replay_pos = rng.cursor();
            st.replay_consumed = rng.cursor();
            S[b].cache_valid = 0;                                                // this kernel does not maintain the energy cache
        }
    }
}

template <int RNG_MODE>
__global__ void __launch_bounds__(GIBBS_THREADS) k2_gibbs(const ChainParams* __restrict__ params, ChainState* __restrict__ states, const int2* __restrict__ pairs,
                                                          double* X, double* Y, double* Z, RunArgs a) {
    __shared__ ChainParams P[2];
    __shared__ ChainState S[2];
    __shared__ double partials[128];
    const int2 pr = pairs[blockIdx.x];
    const int chain[2] = {pr.x, pr.y};
    const int tid = threadIdx.x;
    if (tid < 2) { P[tid] = params[chain[tid]]; S[tid] = states[chain[tid]]; }
    __syncthreads();
    double* const xs[2] = {X + (size_t)chain[0] * a.cap, X + (size_t)chain[1] * a.cap};
    double* const ys[2] = {Y + (size_t)chain[0] * a.cap, Y + (size_t)chain[1] * a.cap};
    double* const zs[2] = {Z + (size_t)chain[0] * a.cap, Z + (size_t)chain[1] * a.cap};
    unsigned char* trace = a.trace ? a.trace + (unsigned long long)chain[0] * a.trace_stride : nullptr;   // the system's trace: box 0's row
    if (RNG_MODE == MCPM_RNG_PHILOX) {
        PhiloxSource rng; rng.init(a.seed, S[0].stream_id, S[0].st.steps_done, tid & 31);
        gibbs_loop(P, S, rng, xs, ys, zs, partials, a, trace, a.cap);
    } else {
        ReplaySource rng; rng.init(a.replay_u + (unsigned long long)chain[0] * a.replay_stride, 0ull, a.replay_stride);
        gibbs_loop(P, S, rng, xs, ys, zs, partials, a, trace, a.cap);
    }
    __syncthreads();
    if (tid < 2) states[chain[tid]] = S[tid];
}

}  // namespace

cudaError_t launch_k2_gibbs(int rng_mode, int n_pairs, const ChainParams* params, ChainState* states, const int2* pairs, double* X, double* Y,
                            double* Z, const RunArgs& a, cudaStream_t stream) {
    if (n_pairs <= 0) return cudaSuccess;
    if (rng_mode == MCPM_RNG_PHILOX) k2_gibbs<MCPM_RNG_PHILOX><<<n_pairs, GIBBS_THREADS, 0, stream>>>(params, states, pairs, X, Y, Z, a);
    else k2_gibbs<MCPM_RNG_REPLAY><<<n_pairs, GIBBS_THREADS, 0, stream>>>(params, states, pairs, X, Y, Z, a);
    return cudaGetLastError();
}

}  // namespace mcpm
