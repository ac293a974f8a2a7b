// K2x — speculative moves: the LATENCY kernel for few chains (a 40-point isotherm with one chain per state point is
// 40 sequential Markov chains: 148 SMs cannot be filled with chains, so each chain should finish its steps sooner).
//
// A Markov chain is sequential only through its ACCEPTED moves.  Most trial moves are rejected (the reference adapts the
// step size to ~20 % translation acceptance; GCMC insertions/deletions in a dense pore succeed a few % of the time), and
// a rejected move leaves the configuration untouched.  One CTA of KW warps therefore evaluates the next KW steps of ITS
// chain concurrently, every warp one step against the SAME configuration, as if all earlier steps of the round were
// rejected.  The first accepted step k* is committed, the steps after k* are discarded and evaluated again in the next
// round with the same uniforms (the stream is counter-based: a step's draws depend only on its index), so the chain is
// exactly the sequential one.  Expected progress per round is (1-(1-p)^KW)/p steps (5.5 at p = 0.11, KW = 8).
// A warp scans all partners of its candidate alone (old energies come from the per-particle energy cache, so a
// candidate costs at most one scan); the commit of the accepted move (second pass over the partners) uses the whole CTA.
//
// Critical path of a round: read 8 uniforms from a shared-memory ring -> scan -> Metropolis -> publish {type, accepted}
// as ONE BYTE per candidate -> barrier -> every thread loads the 8 bytes as one word and finds the first accepted
// candidate with a bit scan -> commit.  Everything else lives in two helper warps that evaluate no candidate and work
// while the candidate warps are already in the next round: lane 0 of the BOOKKEEPING warp keeps the per-set counters
// (popc over the code word: O(1) per round), the trace, the production accumulators and the stale-slot writes of
// rejected insertions; the UNIFORMS warp refills the ring (Philox4x32-10 for the steps up to 2*KW ahead).
//
// Restates the same reference code as chain_loop in mcpm_kernels.cu: step loop src/main.cpp:66-71, MoveParticle
// src/moves.cpp:118-186, SwapParticle src/moves.cpp:284-341, ResetStats src/MC.cpp:81-98, AdjustMCMoves src/moves.cpp:68-78.
#include <algorithm>

#include "mcpm_chain.cuh"
#include "mcpm_device.cuh"
#include "mcpm_kernels.h"

namespace mcpm {

namespace {

// One-hot move types: the round's eight codes are read as one 64-bit word and counted with popc on bit planes.
enum { CAND_TRANS = 0x01, CAND_INS = 0x02, CAND_DEL = 0x04, CAND_EMPTY = 0x08, CAND_VOL = 0x10, CAND_OVERFLOW = 0x20, CAND_ACCEPTED = 0x80 };
#define MCPM_PLANE(b) (0x0101010101010101ull * (unsigned long long)(b))

struct Cand {            // one warp's candidate step, published in shared memory
    int type, acc, idx, pad;
    double px, py, pz;   // trial position (translation, insertion)
    double ox, oy, oz;   // position of the selected particle (translation, deletion)
    double v;            // unscaled pair sum of the trial position
    double po, pn, wo, wn;
};

// The 8 uniforms of global step `step` of this chain: lanes 0..3 hold the step's four Philox calls.
struct StepPhilox {
    double ua, ub;
    __device__ __forceinline__ void load(unsigned long long seed, uint32_t chain, unsigned long long step, int lane) {
        philox_call(seed, chain, 4ull * step + (unsigned long long)(lane & 3), ua, ub);
    }
    template <int K>
    __device__ __forceinline__ double at() const { return __shfl_sync(MCPM_FULL_MASK, (K & 1) ? ub : ua, K >> 1); }
};
struct StepRing {        // uniforms prepared by the bookkeeping warp
    const double* u;
    template <int K>
    __device__ __forceinline__ double at() const { return u[K]; }
};
struct StepReplay {
    const double* u;
    unsigned long long pos, n;
    template <int K>
    __device__ __forceinline__ double at() const { return (pos + K < n) ? __ldg(u + pos + K) : 0.0; }
};

#ifdef MCPM_SPEC_TIMING
__device__ __forceinline__ long long spec_clock() {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory");
    return t;
}
// clock read that cannot issue before x is available (the branch needs x)
__device__ __forceinline__ long long spec_probe(double x) {
    if (__double_as_longlong(x) == 0x7ff8dead00000001ll) asm volatile("trap;");
    return spec_clock();
}
#define SPEC_PROBE(k, x) tp[k] = spec_probe(x)
#else
#define SPEC_PROBE(k, x)
#endif

// Evaluates one step read-only.  Called by one full warp; every lane returns the same candidate.
template <class Draws, bool PBZ, bool WIDE>
__device__ __forceinline__ Cand evaluate_candidate(const Draws& rng, const ChainParams& P, const PairGeom& g, const double* xs, const double* ys,
                                                   const double* zs, const double* ep, int n, int cap, double dr, int lane, long long* tp) {
    Cand c;
    SPEC_PROBE(0, dr);
    c.type = CAND_EMPTY; c.acc = 0; c.idx = 0; c.pad = 0;
    c.px = c.py = c.pz = c.ox = c.oy = c.oz = c.v = c.po = c.pn = c.wo = c.wn = 0.0;
    const double r = rng.template at<0>() * P.wsum;                                       // src/main.cpp:67
    if (r <= P.thr_disp) {
        SPEC_PROBE(1, r);
        const double u_i = rng.template at<3>(), u_x = rng.template at<4>(), u_y = rng.template at<5>(), u_z = rng.template at<6>(),
                     u_a = rng.template at<7>();
        const int i = (int)(u_i * (double)n);                                             // n == 0: the stale slot 0 (quirk Q12)
        const double xo = xs[i], yo = ys[i], zo = zs[i];
        const double ep_i = (n > 0) ? ep[i] : 0.0;
        const double xr = __dadd_rn(xo, __dmul_rn(u_x - 0.5, dr)), yr = __dadd_rn(yo, __dmul_rn(u_y - 0.5, dr));
        const double zr = __dadd_rn(zo, __dmul_rn(u_z - 0.5, dr));
        bool odd = false;
        double xn = wrap_select(xr, P.L[0], odd), yn = wrap_select(yr, P.L[1], odd), zn = PBZ ? wrap_select(zr, P.L[2], odd) : zr;
        if (odd) { xn = wrap_floor(xr, P.L[0]); yn = wrap_floor(yr, P.L[1]); if (PBZ) zn = wrap_floor(zr, P.L[2]); }
        SPEC_PROBE(2, xn + yn + zn);
        double v = !WIDE ? pair_sum1<true, PBZ>(xs, ys, zs, n, i, g, xn, yn, zn, lane, 32)
                 : (n <= 128) ? pair_sum_wide<true, PBZ, 4>(xs, ys, zs, n, i, g, xn, yn, zn, lane, 32)
                              : pair_sum_wide<true, PBZ, 8>(xs, ys, zs, n, i, g, xn, yn, zn, lane, 32);
        SPEC_PROBE(3, v);
        double wo, wn;
        wall_two_and_sum(P, zo, zn, wo, wn, v, lane);      // wall and scan reductions interleaved
        SPEC_PROBE(4, wo + wn);
        SPEC_PROBE(5, v);
        const double po = P.eps4 * ep_i, pn = P.eps4 * v;
        const double dE = (pn + wn) - (po + wo);                                          // src/moves.cpp:169
        c.type = CAND_TRANS; c.acc = metropolis_translation(u_a, dE * P.beta) ? 1 : 0; c.idx = i;
        SPEC_PROBE(6, (double)c.acc);
        c.px = xn; c.py = yn; c.pz = zn; c.ox = xo; c.oy = yo; c.oz = zo; c.v = v; c.po = po; c.pn = pn; c.wo = wo; c.wn = wn;
    } else if (r <= P.thr_vol) {
        c.type = CAND_VOL;
    } else if (rng.template at<2>() < 0.5) {
        if (n >= cap) { c.type = CAND_OVERFLOW; c.acc = 1; }
        else {
            const double u_x = rng.template at<3>(), u_y = rng.template at<4>(), u_z = rng.template at<5>(), u_a = rng.template at<6>();
            const double xi = u_x * P.L[0], yi = u_y * P.L[1], zi = u_z * P.L[2];         // src/moves.cpp:46-54
            double v = !WIDE ? pair_sum1<true, PBZ>(xs, ys, zs, n, -1, g, xi, yi, zi, lane, 32)
                     : (n <= 128) ? pair_sum_wide<true, PBZ, 4>(xs, ys, zs, n, -1, g, xi, yi, zi, lane, 32)
                                  : pair_sum_wide<true, PBZ, 8>(xs, ys, zs, n, -1, g, xi, yi, zi, lane, 32);
            double wi, wd;
            wall_two_and_sum(P, zi, zi, wi, wd, v, lane);
            const double pi_ = P.eps4 * v;
            c.type = CAND_INS; c.acc = metropolis_insertion(u_a, -((pi_ + wi) * P.beta), P.ln_zzV, P.zzV, n) ? 1 : 0; c.idx = n;   // src/moves.cpp:306
            c.px = xi; c.py = yi; c.pz = zi; c.v = v; c.pn = pi_; c.wn = wi;
        }
    } else if (n > 0) {
        const double u_o = rng.template at<3>(), u_a = rng.template at<4>();
        const int o = (int)(u_o * (double)n + 1.0) - 1;
        const double xo = xs[o], yo = ys[o], zo = zs[o];
        double wo, wd;
        wall_two(P, zo, zo, wo, wd, lane);
        const double po = P.eps4 * ep[o];
        c.type = CAND_DEL; c.acc = metropolis_deletion(u_a, (po + wo) * P.beta, P.ln_zzV, P.zzV, n) ? 1 : 0; c.idx = o;            // src/moves.cpp:326
        c.ox = xo; c.oy = yo; c.oz = zo; c.po = po; c.wo = wo;
    }
    return c;
}

}  // namespace

// LEAN = false: the latency kernel described above (8 candidate warps + 2 helper warps, energy cache in shared memory, one chain
//   per SM).
// LEAN = true : the THROUGHPUT variant for many small chains: KW = 2 candidate warps and nothing else per chain, so that as many
//   chains stay resident per SM as with one warp per chain (same shared memory: the energy cache stays in global memory) but
//   twice the warps are there to hide each other's latencies, and a round advances the chain by 2 - p steps.  Thread 0 keeps
//   the books, the last warp refills the uniform ring 8 steps at a time.
template <int RNG_MODE, bool PBZ, int KW, bool LEAN>
__global__ void __launch_bounds__(32 * (LEAN ? KW : KW + 2), LEAN ? 16 / KW : 1) k2_spec(const ChainParams* __restrict__ params,
                                                                                   ChainState* __restrict__ states,
                                                                                   const int* __restrict__ order, double* X, double* Y, double* Z,
                                                                                   RunArgs a) {
    static_assert(KW >= 2 && KW <= 8, "one code byte per candidate in a 64-bit word");
    constexpr int RS = 16;                           // ring of prepared uniforms, in steps (>= 2*KW + 8 for the lazy refill)
    constexpr int W = LEAN ? KW : KW + 2;            // + bookkeeping warp (KW) + uniforms warp (KW + 1)
    constexpr int BOOK_WARP = LEAN ? 0 : KW, RING_WARP = LEAN ? KW - 1 : KW + 1;
    extern __shared__ double smem[];
    __shared__ ChainParams P;
    __shared__ ChainState S;
    __shared__ Cand cand[2][KW];                     // double-buffered by round parity: a round needs ONE barrier
    __shared__ __align__(8) unsigned char code[2][8];
    __shared__ __align__(16) double ring[RS][8];
    __shared__ unsigned long long rpos[KW + 1];      // replay mode: stream position of every candidate of the round
    __shared__ double sh_dr;
    const int chain = order[blockIdx.x];
    const int cap = a.cap;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
    double* xs = smem;
    double* ys = smem + cap;
    double* zs = smem + 2 * cap;
    const size_t off = (size_t)chain * (size_t)cap;
    double* ep_global = a.ep + off;
    double* ep = LEAN ? ep_global : smem + 3 * cap;  // per-particle pair sums (the energy cache): shared memory, or global when LEAN
    double* partials = smem + (LEAN ? 3 : 4) * cap;  // 128 doubles (LEAN: 16)
    if (tid == 0) { P = params[chain]; S = states[chain]; }
    if (tid < 16) code[tid >> 3][tid & 7] = 0;       // bytes of candidates that do not exist (KW < 8) stay zero
    __syncthreads();
    for (int j = tid; j < cap; j += T) { xs[j] = X[off + j]; ys[j] = Y[off + j]; zs[j] = Z[off + j]; if (!LEAN) ep[j] = ep_global[j]; }

    const bool book = (tid == 32 * BOOK_WARP);       // lane 0 of the bookkeeping warp; the only writer of S.st
    const PairGeom g = make_geom<true>(P);
    mcpm_chain_stats& st = S.st;
    unsigned char* trace = a.trace ? a.trace + (unsigned long long)chain * a.trace_stride : nullptr;
    const double* replay = (RNG_MODE == MCPM_RNG_REPLAY) ? a.replay_u + (unsigned long long)chain * a.replay_stride : nullptr;

    int n = st.n;
    double dr = st.dr, e_pair = st.pair, e_wall = st.wall;
    unsigned long long step = st.steps_done, rcur = 0ull;
    int overflow = 0, par = 0;
    const int steps_per_set = (int)a.steps_per_set;
    unsigned long long filled = step;                // the ring holds the uniforms of steps [.., filled)
    // Uniforms warp: prepare the uniforms of the steps below `upto` (at most RS ahead of the oldest step still needed).
    auto refill = [&](unsigned long long upto) {
        while (filled < upto) {
            const unsigned long long stp = filled + (unsigned long long)(lane >> 2);
            if (stp < upto) {
                double ua, ub;
                philox_call(a.seed, S.stream_id, 4ull * stp + (unsigned long long)(lane & 3), ua, ub);
                double* slot = &ring[(int)(stp & (unsigned long long)(RS - 1))][2 * (lane & 3)];
                slot[0] = ua; slot[1] = ub;
            }
            filled = (filled + 8ull < upto) ? filled + 8ull : upto;
        }
    };
    if (RNG_MODE == MCPM_RNG_PHILOX && warp == RING_WARP) refill(step + (unsigned long long)RS);
    __syncthreads();

    {   // energy cache: reuse the previous launch's if the configuration is unchanged, else rebuild (see chain_loop)
        const unsigned long long sum = config_checksum<true>(xs, ys, zs, n, (unsigned long long*)partials, tid, T, lane, warp, W);
        if (!(S.cache_valid && S.cache_sum == sum)) {
            double cp, cw;
            cache_fill<true, PBZ, true>(xs, ys, zs, ep, n, P, g, partials, W, warp, lane, tid, T, cp, cw);
            e_pair = cp; e_wall = cw;
            if (book) st.pair_evals += (unsigned long long)n * (unsigned long long)n;
        }
    }

    SetCounters sc;                                  // meaningful in the bookkeeping thread only
    double sum_n = st.sum_n, sum_n2 = st.sum_n2, sum_e = st.sum_e, sum_e2 = st.sum_e2;
    unsigned long long samples = st.samples;
#ifdef MCPM_SPEC_TIMING
    double tacc[8] = {0., 0., 0., 0., 0., 0., 0., 0.};
    double tsub[8] = {0., 0., 0., 0., 0., 0., 0., 0.};
#endif
    for (int set = a.first_set; set < a.first_set + a.n_sets && !overflow; set++) {
        sc.reset();                                  // MC::ResetStats, src/MC.cpp:81-98
        if (book) { st.widom_insert = st.widom_delete = 0.0; st.widom_insertions = st.widom_deletions = 1; }
        const bool production = set > P.n_equil;
        int s = 0;
        while (s < steps_per_set && !overflow) {
            // An empty box exposes the stale slot 0, which rejected insertion attempts overwrite (quirk Q12): serial rounds.
            const int B = (n == 0) ? 1 : min(KW, steps_per_set - s);
#ifdef MCPM_SPEC_TIMING
            const long long tq0 = spec_clock();
#endif
            if (RNG_MODE == MCPM_RNG_REPLAY) {       // where each candidate's draws start: 1 + 7/6/4/2 draws per step
                if (tid == 0) {
                    unsigned long long p = rcur;
                    for (int w = 0; w < B; w++) {
                        rpos[w] = p;
                        const double u0 = (p < a.replay_stride) ? replay[p] : 0.0, u2 = (p + 2 < a.replay_stride) ? replay[p + 2] : 0.0;
                        const double r = u0 * P.wsum;
                        p += (r <= P.thr_disp) ? 8 : ((r <= P.thr_vol) ? 1 : ((u2 < 0.5) ? 7 : (n > 0 ? 5 : 3)));
                    }
                    rpos[B] = p;
                }
                __syncthreads();
            }
            // ---- every candidate warp evaluates one step against the current configuration ----
            if (warp < B) {
                Cand c;
                long long tp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                (void)tp;
                if (RNG_MODE == MCPM_RNG_PHILOX) {
                    StepRing rng; rng.u = ring[(int)((step + (unsigned long long)warp) & (unsigned long long)(RS - 1))];
                    c = evaluate_candidate<StepRing, PBZ, !LEAN>(rng, P, g, xs, ys, zs, ep, n, cap, dr, lane, tp);
                } else {
                    StepReplay rng; rng.u = replay; rng.pos = rpos[warp]; rng.n = a.replay_stride;
                    c = evaluate_candidate<StepReplay, PBZ, !LEAN>(rng, P, g, xs, ys, zs, ep, n, cap, dr, lane, tp);
                }
                if (lane == 0) { cand[par][warp] = c; code[par][warp] = (unsigned char)(c.type | (c.acc ? CAND_ACCEPTED : 0)); }
#ifdef MCPM_SPEC_TIMING
                if (lane == 0 && c.type == CAND_TRANS) {
                    const long long tend = spec_clock();
                    for (int k = 0; k < 6; k++) tsub[k] += (double)(tp[k + 1] - tp[k]);
                    tsub[6] += (double)(tend - tp[6]); tsub[7] += 1.0;
                }
#endif
            } else if (warp < KW && lane == 0) {
                code[par][warp] = 0;
            }
#ifdef MCPM_SPEC_TIMING
            const long long tq1 = spec_clock();
#endif
            __syncthreads();
#ifdef MCPM_SPEC_TIMING
            const long long tq2 = spec_clock();
#endif
            // ---- the first accepted candidate wins; everything before it is a rejected step that stands ----
            const unsigned long long lo = *reinterpret_cast<const unsigned long long*>(&code[par][0]);
            const Cand* cd = cand[par];
            par ^= 1;
            const unsigned long long am = lo & MCPM_PLANE(CAND_ACCEPTED);
            const int kstar = am ? ((__ffsll((long long)am) - 1) >> 3) : -1;
            const int tk = (kstar >= 0) ? (int)((lo >> (8 * kstar)) & 0x7Full) : 0;
            const bool stop = (tk == CAND_OVERFLOW);
            const bool committed = (kstar >= 0) && !stop;
            const int kept = stop ? kstar : (kstar >= 0 ? kstar + 1 : B);
            const int n_before = n;
            const double e_before = e_pair + e_wall;
            unsigned long long commit_evals = 0ull;
            // ---- commit the accepted move with the whole CTA ----
            if (committed) {
                const Cand& c = cd[kstar];
                bool big = false;
                if (tk == CAND_TRANS) {
                    const int i = c.idx;
                    const double px = c.px, py = c.py, pz = c.pz;
                    big = cache_apply<true, PBZ>(xs, ys, zs, ep, n, i, g, true, c.ox, c.oy, c.oz, true, px, py, pz, tid, T);
                    if (tid == 0) { ep[i] = c.v; xs[i] = px; ys[i] = py; zs[i] = pz; }         // readers of slot i mask it out (self)
                    if (n > 0) { e_pair += c.pn - c.po; e_wall += c.wn - c.wo; }   // full pair change (Q16); n == 0 moves a ghost (Q12)
                    commit_evals = 2ull * (unsigned long long)n;
                } else if (tk == CAND_INS) {
                    const double px = c.px, py = c.py, pz = c.pz;
                    big = cache_apply<true, PBZ>(xs, ys, zs, ep, n, -1, g, false, 0., 0., 0., true, px, py, pz, tid, T);
                    if (tid == 0) { ep[n] = c.v; xs[n] = px; ys[n] = py; zs[n] = pz; }
                    e_pair += c.pn; e_wall += c.wn;
                    commit_evals = (unsigned long long)n;
                    n++;
                } else {                                                            // CAND_DEL: swap-with-last, src/moves.cpp:329
                    const int o = c.idx;
                    const double xl = xs[n - 1], yl = ys[n - 1], zl = zs[n - 1];
                    big = cache_apply<true, PBZ>(xs, ys, zs, ep, n, o, g, true, c.ox, c.oy, c.oz, false, 0., 0., 0., tid, T);
                    commit_evals = (unsigned long long)n;
                    __syncthreads();                                                // entry n-1 is final, everyone has read slot n-1
                    if (tid == 0) { ep[o] = ep[n - 1]; xs[o] = xl; ys[o] = yl; zs[o] = zl; }
                    e_pair -= c.po; e_wall -= c.wo;
                    n--;
                }
                if (__syncthreads_or(big ? 1 : 0)) {   // a hard overlap went through the cache: rebuild it
                    double cp, cw;
                    cache_fill<true, PBZ, true>(xs, ys, zs, ep, n, P, g, partials, W, warp, lane, tid, T, cp, cw);
                    e_pair = cp; e_wall = cw;
                    commit_evals += (unsigned long long)n * (unsigned long long)n;
                }
            }
            if (RNG_MODE == MCPM_RNG_REPLAY) rcur = rpos[kept];
            step += (unsigned long long)(kept + (stop ? 1 : 0));
#ifdef MCPM_SPEC_TIMING
            const long long tq3 = spec_clock();
#endif
            // ---- off the critical path: uniforms of later rounds; counters, trace, accumulators ----
            if (RNG_MODE == MCPM_RNG_PHILOX && warp == RING_WARP) {
                if (!LEAN) refill(step + (unsigned long long)RS);
                else if (filled < step + 2ull * KW) refill(filled + 8ull);          // lazily, one full Philox pass (8 steps) at a time
            }
            if (book) {
                const unsigned long long kl = (kept >= 8) ? lo : (lo & ((1ull << (8 * kept)) - 1ull));   // the kept steps' codes
                const int nd = __popcll(kl & MCPM_PLANE(CAND_TRANS)), ni = __popcll(kl & MCPM_PLANE(CAND_INS)), nx = __popcll(kl & MCPM_PLANE(CAND_DEL));
                sc.n_disp += nd; sc.n_ins += ni; sc.n_del += nx; sc.n_empty += __popcll(kl & MCPM_PLANE(CAND_EMPTY));
                if (committed) { sc.acc_disp += (tk == CAND_TRANS); sc.acc_ins += (tk == CAND_INS); sc.acc_del += (tk == CAND_DEL); }
                sc.evals += commit_evals + (unsigned long long)__popcll(lo & MCPM_PLANE(CAND_TRANS | CAND_INS)) * (unsigned long long)n_before;   // executed, kept or not
                sc.alg += (unsigned long long)(2 * nd + ni + nx) * (unsigned long long)n_before;
                if (trace) {
                    for (int w = 0; w < kept; w++) {
                        const int cw = (int)((lo >> (8 * w)) & 0xFFull), ac = cw >> 7, ty = cw & 0x7F;
                        const int tc = ty == CAND_TRANS ? ac : (ty == CAND_INS ? (2 | ac) : (ty == CAND_DEL ? (4 | ac) : (ty == CAND_EMPTY ? 6 : 8)));
                        trace[(unsigned long long)(set - a.first_set) * (unsigned long long)a.steps_per_set + (unsigned long long)(s + w)] = (unsigned char)tc;
                    }
                }
                if (production) {   // rejected steps see the old state, the accepted one (the last kept) the new state
                    const double m = (double)(kept - (committed ? 1 : 0)), dn = (double)n_before;
                    sum_n += m * dn; sum_n2 += m * (dn * dn); sum_e = fma(m, e_before, sum_e); sum_e2 = fma(m, e_before * e_before, sum_e2);
                    if (committed) { const double da = (double)n, e = e_pair + e_wall; sum_n += da; sum_n2 += da * da; sum_e += e; sum_e2 += e * e; }
                    samples += (unsigned long long)kept;
                }
                // the trial position of a rejected insertion attempt stays in the first free slot (src/moves.cpp:302-303)
                const unsigned long long im = kl & MCPM_PLANE(CAND_INS);
                if (im) {
                    const int last_ins = (63 - __clzll((long long)im)) >> 3;
                    if (last_ins != kstar) { xs[n_before] = cd[last_ins].px; ys[n_before] = cd[last_ins].py; zs[n_before] = cd[last_ins].pz; }
                }
            }
            s += kept;
            if (stop) overflow = 1;
#ifdef MCPM_SPEC_TIMING
            if (lane == 0) {
                const long long tq4 = spec_clock();
                tacc[0] += (double)(tq1 - tq0); tacc[1] += (double)(tq2 - tq1); tacc[2] += (double)(tq3 - tq2); tacc[3] += (double)(tq4 - tq3);
                tacc[4] += 1.0; tacc[5] += committed ? 1.0 : 0.0; tacc[6] += (double)kept;
            }
#endif
            if (RNG_MODE == MCPM_RNG_REPLAY || n_before == 0 || n == 0) __syncthreads();   // rpos[] / slot 0 are reused at once
        }
        // MC::AdjustMCMoves displacement part, src/moves.cpp:68-78 (ResetStats starts nDisplacements at ONE)
        if (book) {
            st.dr_used = dr;                        // the amplitude this set ran with (src/main.cpp:77-83 prints before adjusting)
            if (!overflow && n > 0 && set <= P.n_equil) {
                const double ratio = (double)sc.acc_disp * 1. / (1. * (double)(sc.n_disp + 1));
                if (ratio < 0.2) dr /= 1.1;
                else dr *= 1.1;
            }
            sh_dr = dr;
            st.acceptance = sc.acc_disp; st.rejection = sc.n_disp - sc.acc_disp; st.n_displacements = 1 + sc.n_disp;
            st.accept_swap = sc.acc_ins + sc.acc_del; st.reject_swap = (sc.n_ins + sc.n_del) - (sc.acc_ins + sc.acc_del);
            st.n_swaps = 1 + sc.n_ins + sc.n_del;
            st.tot_disp += sc.n_disp; st.tot_disp_acc += sc.acc_disp; st.tot_ins += sc.n_ins; st.tot_ins_acc += sc.acc_ins;
            st.tot_del += sc.n_del; st.tot_del_acc += sc.acc_del; st.tot_empty += sc.n_empty; st.pair_evals += sc.evals; st.pair_evals_algorithmic += sc.alg;
        }
        __syncthreads();
        dr = sh_dr;
    }
    __syncthreads();

    double c_pair = 0.0, c_wall = 0.0;
    if (a.check_energy) cache_fill<true, PBZ, true>(xs, ys, zs, ep, n, P, g, partials, W, warp, lane, tid, T, c_pair, c_wall);
    const unsigned long long end_sum = config_checksum<true>(xs, ys, zs, n, (unsigned long long*)partials, tid, T, lane, warp, W);
    if (book) {
        st.n = n; st.overflow = overflow; st.dr = dr; st.steps_done = step;
        st.pair = e_pair; st.wall = e_wall; st.energy = e_pair + e_wall;
        st.sum_n = sum_n; st.sum_n2 = sum_n2; st.sum_e = sum_e; st.sum_e2 = sum_e2; st.samples = samples;
        if (a.check_energy) { st.pair_check = c_pair; st.wall_check = c_wall; st.energy_check = c_pair + c_wall; }
        if (a.check_energy == 2) { st.pair = c_pair; st.wall = c_wall; st.energy = c_pair + c_wall; }
        S.replay_pos = rcur; st.replay_consumed = rcur;
        S.cache_valid = overflow ? 0 : 1; S.cache_sum = end_sum;
    }
    __syncthreads();
    for (int j = tid; j < cap; j += T) { X[off + j] = xs[j]; Y[off + j] = ys[j]; Z[off + j] = zs[j]; if (!LEAN) ep_global[j] = ep[j]; }
    if (book) states[chain] = S;
#ifdef MCPM_SPEC_TIMING
    if (lane == 0 && a.rdf) for (int k = 0; k < 8; k++) a.rdf[(size_t)chain * MCPM_RDF_BINS + warp * 8 + k] = tacc[k];   // debug build only
    if (tid == 32 && a.rdf) for (int k = 0; k < 8; k++) a.rdf[(size_t)chain * MCPM_RDF_BINS + 80 + k] = tsub[k];
#endif
}

size_t k2_spec_smem_bytes(int cap) { return (size_t)(4 * cap + 128) * sizeof(double); }
size_t k2_pair_smem_bytes(int cap) { return (size_t)(3 * cap + 16) * sizeof(double); }

cudaError_t launch_k2_spec(int rng_mode, bool pbz, int lean /* 0, or candidate warps of the lean variant: 2 or 4 */, int n_chains, const ChainParams* params, ChainState* states, const int* order,
                           double* X, double* Y, double* Z, const RunArgs& a, cudaStream_t stream) {
    const size_t smem = lean ? k2_pair_smem_bytes(a.cap) : k2_spec_smem_bytes(a.cap);
    auto go = [&](auto kern, int threads) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<n_chains, threads, smem, stream>>>(params, states, order, X, Y, Z, a);
        return cudaGetLastError();
    };
    const bool ph = rng_mode == MCPM_RNG_PHILOX;
    if (lean == 4) {
        if (ph) return pbz ? go(k2_spec<MCPM_RNG_PHILOX, true, 4, true>, 128) : go(k2_spec<MCPM_RNG_PHILOX, false, 4, true>, 128);
        return pbz ? go(k2_spec<MCPM_RNG_REPLAY, true, 4, true>, 128) : go(k2_spec<MCPM_RNG_REPLAY, false, 4, true>, 128);
    }
    if (lean) {
        if (ph) return pbz ? go(k2_spec<MCPM_RNG_PHILOX, true, 2, true>, 64) : go(k2_spec<MCPM_RNG_PHILOX, false, 2, true>, 64);
        return pbz ? go(k2_spec<MCPM_RNG_REPLAY, true, 2, true>, 64) : go(k2_spec<MCPM_RNG_REPLAY, false, 2, true>, 64);
    }
    if (ph) return pbz ? go(k2_spec<MCPM_RNG_PHILOX, true, 8, false>, 320) : go(k2_spec<MCPM_RNG_PHILOX, false, 8, false>, 320);
    return pbz ? go(k2_spec<MCPM_RNG_REPLAY, true, 8, false>, 320) : go(k2_spec<MCPM_RNG_REPLAY, false, 8, false>, 320);
}

}  // namespace mcpm
