// K2d — one warp per chain, TWO candidate steps per round (the throughput kernel for many small chains).
//
// k2_chains with one warp per chain is latency-bound: a step is one dependent chain (uniforms -> selection -> partner
// scan -> warp reduction -> wall -> Metropolis -> write) and shared memory limits an SM to ~13 resident chains, i.e. ~3
// warps per scheduler to hide it (FP64 pipe 41 % busy, issue slots 54 % busy on config C2).  Like mcpm_spec.cu this
// kernel uses the fact that a chain is sequential only through its ACCEPTED moves (~10 % of the trial moves): every
// round it evaluates step s and step s+1 against the same configuration in ONE interleaved instruction stream — one
// partner loop serving both trial positions, one 32-lane wall evaluation for the four z values, one reduction of two
// values — which doubles the instruction-level parallelism of the warp.  If step s is accepted it is committed and
// step s+1 is discarded (re-evaluated next round with the same uniforms: the streams are addressed by step index);
// otherwise step s stands as rejected and step s+1 is decided as usual.  Progress is 2 - p steps per round, and the
// chain is exactly the sequential one.
//
// Restates the same reference code as chain_loop in mcpm_kernels.cu: step loop src/main.cpp:66-71, MoveParticle
// src/moves.cpp:118-186, SwapParticle src/moves.cpp:284-341, ResetStats src/MC.cpp:81-98, AdjustMCMoves src/moves.cpp:68-78.
// Needs: energy cache, coordinates inside the box (fast image), shared-memory positions, no in-kernel sampling.
#include <algorithm>

#include "../mcpm_chain.cuh"
#include "../mcpm_device.cuh"
#include "../mcpm_kernels.h"

namespace mcpm {

namespace {

enum { MV_TRANS = 0, MV_INS = 1, MV_DEL = 2, MV_EMPTY = 3, MV_VOL = 4, MV_OVERFLOW = 5 };

// Two Philox batches (16 steps) in registers: step t of the window [b, b+16) has its call k/2 in lane 4*(t & 7) + k/2 of
// batch (t - b) >> 3.  One Philox call per lane per 8 steps, as in PhiloxSource.
struct DualPhilox {
    double ua0, ub0, ua1, ub1;
    unsigned long long b;          // first step of batch 0 (multiple of 8)
    unsigned long long seed;
    uint32_t chain;
    int off[2];                    // candidate c is step b + off[c]
    __device__ __forceinline__ void init(unsigned long long seed_, uint32_t chain_, unsigned long long step, int lane) {
        seed = seed_; chain = chain_;
        b = step & ~7ull;
        philox_call(seed, chain, 4ull * b + (unsigned long long)lane, ua0, ub0);
        philox_call(seed, chain, 4ull * (b + 8ull) + (unsigned long long)lane, ua1, ub1);
    }
    __device__ __forceinline__ void begin_round(unsigned long long step, int lane) {
        if (step >= b + 8ull) {
            b += 8ull;
            ua0 = ua1; ub0 = ub1;
            philox_call(seed, chain, 4ull * (b + 8ull) + (unsigned long long)lane, ua1, ub1);
        }
        off[0] = (int)(step - b); off[1] = off[0] + 1;
    }
    template <int K>
    __device__ __forceinline__ double at(int c) const {
        const int t = off[c];
        const double v0 = (K & 1) ? ub0 : ua0, v1 = (K & 1) ? ub1 : ua1;
        return __shfl_sync(MCPM_FULL_MASK, (t & 8) ? v1 : v0, 4 * (t & 7) + (K >> 1));
    }
    __device__ __forceinline__ void second_starts_after(int) {}
    __device__ __forceinline__ void end_round(int) {}
    __device__ __forceinline__ unsigned long long cursor() const { return 0ull; }
};

// Replay: the reference's sequential stream, 1 + 7/6/4/2 draws per step; candidate 1 starts where candidate 0 ends.
struct DualReplay {
    const double* u;
    unsigned long long pos[2], n;
    __device__ __forceinline__ void init(const double* p, unsigned long long len) { u = p; pos[0] = pos[1] = 0ull; n = len; }
    __device__ __forceinline__ void begin_round(unsigned long long, int) {}
    template <int K>
    __device__ __forceinline__ double at(int c) const { return (pos[c] + K < n) ? __ldg(u + pos[c] + K) : 0.0; }
    __device__ __forceinline__ void second_starts_after(int count0) { pos[1] = pos[0] + (unsigned long long)count0; }
    __device__ __forceinline__ void end_round(int consumed) { pos[0] += (unsigned long long)consumed; }
    __device__ __forceinline__ unsigned long long cursor() const { return pos[0]; }
};

struct Prep {                      // one candidate step before its energies are known (replicated in every lane)
    int type, idx, self, count;
    bool scan;
    double x, y, z;                // position to scan (translation: new position, insertion: trial position)
    double z_old, z_new;           // wall arguments
    double ep_old;                 // cached pair sum of the selected particle (translation, deletion)
};

template <class Draws, bool PBZ>
__device__ __forceinline__ Prep prepare(const Draws& rng, int c, const ChainParams& P, const double* xs, const double* ys, const double* zs,
                                        const double* ep, int n, int cap, double dr) {
    Prep q;
    q.type = MV_EMPTY; q.idx = 0; q.self = -1; q.count = 3; q.scan = false;
    q.x = q.y = q.z = 0.0; q.z_old = q.z_new = P.halfH; q.ep_old = 0.0;
    const double r = rng.template at<0>(c) * P.wsum;                                      // src/main.cpp:67
    if (r <= P.thr_disp) {
        // draws 1,2 (box, species selection) are consumed but unused: single box, single species
        const double u_i = rng.template at<3>(c), u_x = rng.template at<4>(c), u_y = rng.template at<5>(c), u_z = rng.template at<6>(c);
        const int i = (int)(u_i * (double)n);                                             // n == 0: the stale slot 0 (quirk Q12)
        const double xo = xs[i], yo = ys[i], zo = zs[i];
        q.ep_old = (n > 0) ? ep[i] : 0.0;
        const double xr = __dadd_rn(xo, __dmul_rn(u_x - 0.5, dr)), yr = __dadd_rn(yo, __dmul_rn(u_y - 0.5, dr));
        const double zr = __dadd_rn(zo, __dmul_rn(u_z - 0.5, dr));
        bool odd = false;                                                                 // MC::PBC, src/energy.cpp:126-130
        double xn = wrap_select(xr, P.L[0], odd), yn = wrap_select(yr, P.L[1], odd), zn = PBZ ? wrap_select(zr, P.L[2], odd) : zr;
        if (odd) { xn = wrap_floor(xr, P.L[0]); yn = wrap_floor(yr, P.L[1]); if (PBZ) zn = wrap_floor(zr, P.L[2]); }
        q.type = MV_TRANS; q.idx = i; q.self = i; q.count = 8; q.scan = true;
        q.x = xn; q.y = yn; q.z = zn; q.z_old = zo; q.z_new = zn;
    } else if (r <= P.thr_vol) {
        q.type = MV_VOL; q.count = 1;
    } else if (rng.template at<2>(c) < 0.5) {                                             // draw 1 (species) unused
        if (n >= cap) { q.type = MV_OVERFLOW; q.count = 0; }
        else {
            const double u_x = rng.template at<3>(c), u_y = rng.template at<4>(c), u_z = rng.template at<5>(c);
            q.type = MV_INS; q.idx = n; q.self = -1; q.count = 7; q.scan = true;
            q.x = u_x * P.L[0]; q.y = u_y * P.L[1]; q.z = u_z * P.L[2];                   // src/moves.cpp:46-54
            q.z_old = q.z_new = q.z;
        }
    } else if (n > 0) {
        const double u_o = rng.template at<3>(c);
        const int o = (int)(u_o * (double)n + 1.0) - 1;                                   // int(u*n+1) on 1-based slots
        q.type = MV_DEL; q.idx = o; q.count = 5;
        q.z_old = q.z_new = zs[o];
        q.ep_old = ep[o];
    }
    return q;
}

// One partner loop for two trial positions with their own excluded indices: 2 positions x 2 slots = 4 chains in flight.
template <bool PBZ>
__device__ __forceinline__ void pair_sum_two(const double* xs, const double* ys, const double* zs, int n, const PairGeom& g, const Prep& a,
                                             const Prep& b, int lane, double& ea, double& eb) {
    double a0 = 0.0, b0 = 0.0, a1 = 0.0, b1 = 0.0;
    int j = lane;
    for (; j + 32 < n; j += 64) {
        const double x0 = xs[j], y0 = ys[j], z0 = zs[j];
        const double x1 = xs[j + 32], y1 = ys[j + 32], z1 = zs[j + 32];
        pair_acc(a0, dist2<true, PBZ>(g, a.x, a.y, a.z, x0, y0, z0), g, j, a.self);
        pair_acc(b0, dist2<true, PBZ>(g, b.x, b.y, b.z, x0, y0, z0), g, j, b.self);
        pair_acc(a1, dist2<true, PBZ>(g, a.x, a.y, a.z, x1, y1, z1), g, j + 32, a.self);
        pair_acc(b1, dist2<true, PBZ>(g, b.x, b.y, b.z, x1, y1, z1), g, j + 32, b.self);
    }
    if (j < n) {
        const double x0 = xs[j], y0 = ys[j], z0 = zs[j];
        pair_acc(a0, dist2<true, PBZ>(g, a.x, a.y, a.z, x0, y0, z0), g, j, a.self);
        pair_acc(b0, dist2<true, PBZ>(g, b.x, b.y, b.z, x0, y0, z0), g, j, b.self);
    }
    ea = a0 + a1; eb = b0 + b1;
}

// Wall term (see wall_two) for four z values at once: lane = which*8 + layer*2 + side.
__device__ __forceinline__ void wall_four(const ChainParams& P, double z0, double z1, double z2, double z3, double (&w)[4], int lane) {
    double acc = 0.0;
    const int which = lane >> 3;
    const double zsel = (which & 2) ? ((which & 1) ? z3 : z2) : ((which & 1) ? z1 : z0);
    if (P.wall_kind == MCPM_WALL_SLIT_LJ && P.geometry == MCPM_GEOM_SLIT) {
        double zc = zsel - P.halfH;
        if (lane & 1) zc = -zc;
        const int layer0 = (lane >> 1) & 3;
        if (P.n_layers <= 4) {
            const double t = P.sig_sf * fast_rcp(P.halfH + layer0 * P.delta + zc);
            const double t2 = t * t, t4 = t2 * t2;
            const double term = 0.2 * (t4 * t4 * t2) - 0.5 * t4;
            acc = (layer0 < P.n_layers) ? term : 0.0;
        } else {
            for (int layer = layer0; layer < P.n_layers; layer += 4) {
                const double t = P.sig_sf * fast_rcp(P.halfH + layer * P.delta + zc);
                const double t2 = t * t, t4 = t2 * t2;
                acc += 0.2 * (t4 * t4 * t2) - 0.5 * t4;
            }
        }
        acc += __shfl_xor_sync(MCPM_FULL_MASK, acc, 1);
        acc += __shfl_xor_sync(MCPM_FULL_MASK, acc, 2);
        acc += __shfl_xor_sync(MCPM_FULL_MASK, acc, 4);
    }
    const double zz[4] = {z0, z1, z2, z3};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        w[k] = __shfl_sync(MCPM_FULL_MASK, acc, 8 * k) * P.wall_pref;
        if (P.geometry == MCPM_GEOM_SLIT) {  // out-of-pore guard: finite INT_MAX, wall still added (quirk Q6)
            const double zc = zz[k] - P.halfH;
            if (zc * zc > P.sqr_pore_r) w[k] = MCPM_INT_MAX_D + w[k];
        }
    }
}

template <class Draws, bool PBZ>
__device__ void chain_loop_dual(const ChainParams& P, ChainState& S, Draws& rng, double* xs, double* ys, double* zs, const RunArgs& a,
                                unsigned char* trace, int cap, uint32_t chain) {
    const int lane = threadIdx.x;
    const bool lead = (lane == 0);
    const PairGeom g = make_geom<true>(P);
    mcpm_chain_stats& st = S.st;   // shared memory; written by lane 0 only

    int n = st.n;
    double dr = st.dr, e_pair = st.pair, e_wall = st.wall;
    unsigned long long step = st.steps_done;
    int overflow = 0;
    const double beta = P.beta, eps4 = P.eps4, ln_zzV = P.ln_zzV, zzV = P.zzV;
    const int steps_per_set = (int)a.steps_per_set;
    double* ep = a.ep + (size_t)chain * (size_t)cap;
    {
        const unsigned long long sum = config_checksum<false>(xs, ys, zs, n, nullptr, lane, 32, lane, 0, 1);
        if (!(S.cache_valid && S.cache_sum == sum)) {
            double cp, cw;
            cache_fill<true, PBZ, false>(xs, ys, zs, ep, n, P, g, nullptr, 1, 0, lane, lane, 32, cp, cw);
            e_pair = cp; e_wall = cw;
            if (lead) st.pair_evals += (unsigned long long)n * (unsigned long long)n;
        }
        __syncwarp();
    }
    double sum_n = st.sum_n, sum_n2 = st.sum_n2, sum_e = st.sum_e, sum_e2 = st.sum_e2;   // replicated; written back at the end
    unsigned long long samples = st.samples;
    SetCounters sc;

    for (int set = a.first_set; set < a.first_set + a.n_sets && !overflow; set++) {
        sc.reset();                      // MC::ResetStats, src/MC.cpp:81-98
        if (lead) { st.widom_insert = st.widom_delete = 0.0; st.widom_insertions = st.widom_deletions = 1; }
        const bool production = set > P.n_equil;
        unsigned char* tr = trace ? trace + (unsigned long long)(set - a.first_set) * (unsigned long long)a.steps_per_set : nullptr;
        int s = 0;
        while (s < steps_per_set) {
            // An empty box exposes the stale slot 0, which a rejected insertion attempt overwrites (quirk Q12): one step per round.
            const bool two = (n > 0) && (s + 1 < steps_per_set);
            rng.begin_round(step, lane);
            Prep qa = prepare<Draws, PBZ>(rng, 0, P, xs, ys, zs, ep, n, cap, dr);
            if (qa.type == MV_OVERFLOW) { overflow = 1; step++; break; }
            rng.second_starts_after(qa.count);
            Prep qb;
            if (two) qb = prepare<Draws, PBZ>(rng, 1, P, xs, ys, zs, ep, n, cap, dr);
            else { qb.type = MV_VOL; qb.idx = 0; qb.self = -1; qb.count = 0; qb.scan = false; qb.x = qb.y = qb.z = 0.0; qb.z_old = qb.z_new = P.halfH; qb.ep_old = 0.0; }
            // ---- one partner loop, one wall evaluation, one reduction for both candidates ----
            double va = 0.0, vb = 0.0;
            if (qa.scan && qb.scan) { pair_sum_two<PBZ>(xs, ys, zs, n, g, qa, qb, lane, va, vb); sc.evals += 2ull * (unsigned long long)n; }
            else if (qa.scan) { va = pair_sum1<true, PBZ>(xs, ys, zs, n, qa.self, g, qa.x, qa.y, qa.z, lane, 32); sc.evals += (unsigned long long)n; }
            else if (qb.scan) { vb = pair_sum1<true, PBZ>(xs, ys, zs, n, qb.self, g, qb.x, qb.y, qb.z, lane, 32); sc.evals += (unsigned long long)n; }
            double w[4];
            wall_four(P, qa.z_old, qa.z_new, qb.z_old, qb.z_new, w, lane);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { va += __shfl_xor_sync(MCPM_FULL_MASK, va, o); vb += __shfl_xor_sync(MCPM_FULL_MASK, vb, o); }

            // ---- decide candidate 0, then (if it is rejected) candidate 1; commit the accepted one ----
            int kept = 0, consumed = 0;
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const Prep& q = c ? qb : qa;
                if (c == 1 && !two) break;
                if (q.type == MV_OVERFLOW) { overflow = 1; step++; break; }   // the step is counted, like chain_loop does
                const double v = c ? vb : va, wo = w[2 * c], wn = w[2 * c + 1];
                const double po = eps4 * q.ep_old, pn = eps4 * v;
                bool acc = false;
                int code;
                if (q.type == MV_TRANS) {
                    const double dE = (pn + wn) - (po + wo);                          // newEnergy - oldEnergy, src/moves.cpp:169
                    acc = metropolis_translation(rng.template at<7>(c), dE * beta);   // u < exp(-dE/T), src/moves.cpp:170-172
                    sc.n_disp++; sc.alg += 2ull * (unsigned long long)n;
                    if (acc) {
                        sc.acc_disp++;
                        const int i = q.idx;
                        const double xo = xs[i], yo = ys[i], zo = zs[i];
                        if (n > 0) { e_pair += pn - po; e_wall += wn - wo; }          // full pair change (Q16); n == 0 moves a ghost (Q12)
                        __syncwarp();                                                 // everyone has read the old position
                        bool big = cache_apply<true, PBZ>(xs, ys, zs, ep, n, i, g, true, xo, yo, zo, true, q.x, q.y, q.z, lane, 32);
                        if ((i & 31) == lane) { ep[i] = v; xs[i] = q.x; ys[i] = q.y; zs[i] = q.z; }
                        sc.evals += 2ull * (unsigned long long)n;
                        if (cta_sync_or<false>(big)) {                                // a hard overlap went through the cache: rebuild it
                            double cp, cw;
                            cache_fill<true, PBZ, false>(xs, ys, zs, ep, n, P, g, nullptr, 1, 0, lane, lane, 32, cp, cw);
                            e_pair = cp; e_wall = cw;
                            sc.evals += (unsigned long long)n * (unsigned long long)n;
                            __syncwarp();
                        }
                    }
                    code = acc ? 1 : 0;
                } else if (q.type == MV_INS) {
                    // the trial position is written to the first free slot even if rejected (src/moves.cpp:302-303)
                    if ((n & 31) == lane) { xs[n] = q.x; ys[n] = q.y; zs[n] = q.z; }
                    acc = metropolis_insertion(rng.template at<6>(c), -((pn + wn) * beta), ln_zzV, zzV, n);   // src/moves.cpp:306-307
                    sc.n_ins++; sc.alg += (unsigned long long)n;
                    if (acc) {
                        bool big = cache_apply<true, PBZ>(xs, ys, zs, ep, n, -1, g, false, 0., 0., 0., true, q.x, q.y, q.z, lane, 32);
                        if ((n & 31) == lane) ep[n] = v;
                        sc.evals += (unsigned long long)n;
                        sc.acc_ins++; n++; e_pair += pn; e_wall += wn;
                        if (cta_sync_or<false>(big)) {
                            double cp, cw;
                            cache_fill<true, PBZ, false>(xs, ys, zs, ep, n, P, g, nullptr, 1, 0, lane, lane, 32, cp, cw);
                            e_pair = cp; e_wall = cw;
                            sc.evals += (unsigned long long)n * (unsigned long long)n;
                            __syncwarp();
                        }
                    }
                    code = (1 << 1) | (acc ? 1 : 0);
                } else if (q.type == MV_DEL) {
                    acc = metropolis_deletion(rng.template at<4>(c), (po + wo) * beta, ln_zzV, zzV, n);       // src/moves.cpp:326-327
                    sc.n_del++; sc.alg += (unsigned long long)n;
                    if (acc) {
                        sc.acc_del++;
                        const int o = q.idx;
                        const double xo = xs[o], yo = ys[o], zo = zs[o];
                        const double xl = xs[n - 1], yl = ys[n - 1], zl = zs[n - 1];  // the particle that fills the hole
                        __syncwarp();
                        bool big = cache_apply<true, PBZ>(xs, ys, zs, ep, n, o, g, true, xo, yo, zo, false, 0., 0., 0., lane, 32);
                        sc.evals += (unsigned long long)n;
                        big = cta_sync_or<false>(big);                                // entry n-1 is final
                        if ((o & 31) == lane) { ep[o] = ep[n - 1]; xs[o] = xl; ys[o] = yl; zs[o] = zl; }      // swap-with-last, src/moves.cpp:329
                        n--;
                        e_pair -= po; e_wall -= wo;
                        __syncwarp();
                        if (big) {
                            double cp, cw;
                            cache_fill<true, PBZ, false>(xs, ys, zs, ep, n, P, g, nullptr, 1, 0, lane, lane, 32, cp, cw);
                            e_pair = cp; e_wall = cw;
                            sc.evals += (unsigned long long)n * (unsigned long long)n;
                            __syncwarp();
                        }
                    }
                    code = (2 << 1) | (acc ? 1 : 0);
                } else if (q.type == MV_EMPTY) {
                    sc.n_empty++;
                    code = (3 << 1);
                } else {
                    code = (4 << 1);   // volume moves are out of scope (n_vol is rejected by the host API)
                }
                if (lead && tr) tr[s + c] = (unsigned char)code;
                if (production) {
                    const double dn = (double)n, e = e_pair + e_wall;
                    sum_n += dn; sum_n2 += dn * dn; sum_e += e; sum_e2 += e * e;
                    samples++;
                }
                kept++; consumed += q.count;
                if (acc) break;        // the later candidate was evaluated against a configuration that no longer exists
            }
            rng.end_round(consumed);
            step += (unsigned long long)kept;
            s += kept;
            __syncwarp();              // a stale-slot write of this round is visible to the next selection (n == 0, quirk Q12)
            if (overflow) break;
        }
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
    __syncwarp();

    double c_pair = 0.0, c_wall = 0.0;
    if (a.check_energy) { cache_fill<true, PBZ, false>(xs, ys, zs, ep, n, P, g, nullptr, 1, 0, lane, lane, 32, c_pair, c_wall); __syncwarp(); }
    const unsigned long long end_sum = config_checksum<false>(xs, ys, zs, n, nullptr, lane, 32, lane, 0, 1);
    if (lead) {
        st.n = n; st.overflow = overflow; st.dr = dr; st.steps_done = step;
        st.pair = e_pair; st.wall = e_wall; st.energy = e_pair + e_wall;
        st.sum_n = sum_n; st.sum_n2 = sum_n2; st.sum_e = sum_e; st.sum_e2 = sum_e2; st.samples = samples;
        if (a.check_energy) { st.pair_check = c_pair; st.wall_check = c_wall; st.energy_check = c_pair + c_wall; }
        if (a.check_energy == 2) { st.pair = c_pair; st.wall = c_wall; st.energy = c_pair + c_wall; }   // adopt, like BoxEnergy does
        S.replay_pos = rng.cursor();
        st.replay_consumed = rng.cursor();
        S.cache_valid = overflow ? 0 : 1;
        S.cache_sum = end_sum;
    }
}

}  // namespace

template <int RNG_MODE, bool PBZ>
__global__ void __launch_bounds__(32, 12) k2_dual(const ChainParams* __restrict__ params, ChainState* __restrict__ states, const int* __restrict__ order,
                                                  double* X, double* Y, double* Z, RunArgs a) {
    extern __shared__ double smem[];
    __shared__ ChainParams P;
    __shared__ ChainState S;
    const int chain = order[blockIdx.x];
    const int cap = a.cap;
    const size_t off = (size_t)chain * (size_t)cap;
    double* xs = smem;
    double* ys = smem + cap;
    double* zs = smem + 2 * cap;
    const int lane = threadIdx.x;
    if (lane == 0) { P = params[chain]; S = states[chain]; }
    // stage every slot (not only the n live ones): stale slots are observable through quirk Q12
    for (int j = lane; j < cap; j += 32) { xs[j] = X[off + j]; ys[j] = Y[off + j]; zs[j] = Z[off + j]; }
    __syncwarp();
    unsigned char* trace = a.trace ? a.trace + (unsigned long long)chain * a.trace_stride : nullptr;
    if (RNG_MODE == MCPM_RNG_PHILOX) {
        DualPhilox rng; rng.init(a.seed, S.stream_id, S.st.steps_done, lane);
        chain_loop_dual<DualPhilox, PBZ>(P, S, rng, xs, ys, zs, a, trace, cap, (uint32_t)chain);
    } else {
        DualReplay rng; rng.init(a.replay_u + (unsigned long long)chain * a.replay_stride, a.replay_stride);
        chain_loop_dual<DualReplay, PBZ>(P, S, rng, xs, ys, zs, a, trace, cap, (uint32_t)chain);
    }
    __syncwarp();
    for (int j = lane; j < cap; j += 32) { X[off + j] = xs[j]; Y[off + j] = ys[j]; Z[off + j] = zs[j]; }
    if (lane == 0) states[chain] = S;
}

size_t k2_dual_smem_bytes(int cap) { return (size_t)(3 * cap) * sizeof(double); }

cudaError_t launch_k2_dual(int rng_mode, bool pbz, int n_chains, const ChainParams* params, ChainState* states, const int* order, double* X,
                           double* Y, double* Z, const RunArgs& a, cudaStream_t stream) {
    const size_t smem = k2_dual_smem_bytes(a.cap);
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<n_chains, 32, smem, stream>>>(params, states, order, X, Y, Z, a);
        return cudaGetLastError();
    };
    if (rng_mode == MCPM_RNG_PHILOX) return pbz ? go(k2_dual<MCPM_RNG_PHILOX, true>) : go(k2_dual<MCPM_RNG_PHILOX, false>);
    return pbz ? go(k2_dual<MCPM_RNG_REPLAY, true>) : go(k2_dual<MCPM_RNG_REPLAY, false>);
}

}  // namespace mcpm
