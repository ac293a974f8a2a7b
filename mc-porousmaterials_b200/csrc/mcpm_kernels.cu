// CUDA kernels of the engine (sm_100a, fp64):
//   (K1, MC::EnergyOfParticle for a host-driven loop, lives in mcpm_k1.cu)
//   K2 k2_chains            — persistent batched Metropolis chains, one CTA per chain
//   K3 k3_box_energy        — MC::BoxEnergy, per-particle and box totals, O(N^2) tiled through smem
//   fp64_peak_kernel        — DFMA microbenchmark for the roofline denominator
#include <algorithm>

#include "mcpm_chain.cuh"
#include "mcpm_device.cuh"
#include "mcpm_kernels.h"

namespace mcpm {

// ================================================================================================
// K2 — batched persistent chains.
//
// One CTA advances one chain for n_sets*steps_per_set trial moves without any host round trip:
// the body of the reference step loop (src/main.cpp:66-71) with MoveParticle (src/moves.cpp:118-186),
// SwapParticle's single-box branch (src/moves.cpp:284-341), ResetStats (src/MC.cpp:81-98) and
// AdjustMCMoves (src/moves.cpp:68-78).  CorrectEnergy's heuristic recompute (src/energy.cpp:35-42,
// quirk Q4) is replaced by exact bookkeeping plus an optional full recompute at the end.
//
// Layout: positions live in shared memory as SoA fp64 (x[cap] | y[cap] | z[cap]); thread t owns the
// slots j == t (mod T) — it is the only thread that reads them in the partner loop and the only one
// that writes them.  Uniforms, the Metropolis decision and all scalar state are computed
// REDUNDANTLY by every thread from identical inputs (same uniform stream, partial sums added in the
// same order), so the decision needs no broadcast and a step costs ONE __syncthreads:
//   phase A  select the move, read the selected particle (all cross-thread position reads happen
//            here; a slot written in the previous phase D is forwarded from registers instead);
//   phase B  partner loop over the thread's own slots, warp xor-reduction, partials -> smem[parity];
//   barrier
//   phase D  every thread sums the W partials, decides, owner thread writes the slot.
// ================================================================================================

// n, dr, the running energies, the per-set counters and the decision itself are REPLICATED in every
// thread's registers: all threads compute them from identical inputs, so no broadcast is ever needed and
// nothing but the partner loop's partial sums crosses threads.  (A variant that kept counters and
// energies in shared memory for thread 0 only used 80 registers instead of 128 but was 10-15 % slower at
// every occupancy measured on B200: the extra LDS/STS sit on the per-step critical path.)
#ifdef MCPM_K2_TIMING   // phase timing of a translation step (scripts/k2_timing.py); the probes serialise the phases they separate
__device__ __forceinline__ long long k2_probe(double x) {   // clock read that cannot issue before x is available
    if (__double_as_longlong(x) == 0x7ff8dead00000001ll) asm volatile("trap;");
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory");
    return t;
}
#define K2_PROBE(k, x) tp[k] = k2_probe(x)
#else
#define K2_PROBE(k, x)
#endif

// POOL (k2_pool below): the chain lives in a shared-memory SLOT of cap_slot <= cap particles; a chain that outgrows its slot
// PAUSES (position in the run and the counters of the interrupted set go to S.res) and is resumed by a later launch in a larger slot.
template <class Source, bool FAST, bool MULTI, bool PBZ, bool SAMPLE, bool CACHE, bool POOL = false>
__device__ void chain_loop(const ChainParams& P, ChainState& S, Source& rng, double* xs, double* ys, double* zs, double* partials,
                           const RunArgs& a, unsigned char* trace, int cap, uint32_t chain, double* hist, int cap_slot = 0) {
    if (!POOL) cap_slot = cap;
    const uint32_t sid = S.stream_id;   // Philox streams are keyed by the chain's global id, not its index inside this context
    // (pooled kernel: blocks are (32, warps) — threadIdx.x is the lane and blockDim.x == 32, so every warp of the CTA is a "CTA" of its own here)
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, W = T >> 5;
    const int tmask = T - 1;  // T is a power of two
    const bool lead = (tid == 0);
    PairGeom g = make_geom<FAST>(P);   // (mutable only for NPT volume moves in the general kernels)
    mcpm_chain_stats& st = S.st;   // shared memory; written by thread 0 only
    const bool widom_on = SAMPLE && (P.n_swap == 0) && !a.no_sampling;   // ComputeWidom acts only without swap moves, src/moves.cpp:415

    int n = st.n;
    double nd = (double)n;            // n as a double, kept in step (the conversion would sit on every step's critical path)
    double dr = st.dr, e_pair = st.pair, e_wall = st.wall;
    unsigned long long step = st.steps_done;
    int overflow = 0;
    // (round 2: reading these from shared memory where they are used, and the production accumulators below from shared memory too,
    // frees ~20 registers but was measured 1-5 % slower in this kernel and only +8 % in the 16-warp pooled one: profiles/r02_pool.md)
    const double wsum = P.wsum, thr_disp = P.thr_disp, thr_vol = P.thr_vol;
    const double beta = P.beta, eps4 = P.eps4, ln_zzV = P.ln_zzV, zzV = P.zzV;
    double Lx = P.L[0], Ly = P.L[1], Lz = P.L[2];
    // ---- general kernels only (row f4): NPT state.  MC::ChangeVolume compares the fresh BoxEnergy with the reference's RUNNING box
    // energy (src/moves.cpp:219-220), which as shipped moves by HALF the pair change of an accepted translation (quirk Q16) and is
    // re-anchored by CorrectEnergy's ratio test (src/energy.cpp:35-42, Q4).  Both are tracked here beside the exact energies: the
    // recompute CorrectEnergy would do returns what the exact bookkeeping already holds, so it costs nothing.
    const bool npt = !FAST && P.n_vol > 0;
    // hard spheres: every stored particle carries its own INT_MAX in MC::BoxEnergy (quirk Q17), i.e. half of it in the halved box total
    const double hs_self = (!FAST && g.hs) ? 0.5 * MCPM_INT_MAX_D : 0.0;
    double volume = npt ? Lz * Lz * Lz : 0.0, dv = st.dv, b_energy = S.b_valid ? S.b_energy : st.pair + st.wall, b_old = S.b_valid ? S.b_old : 0.0;
    int acc_vol = 0, rej_vol = 0, n_volch = 1;
    Fwd fwd; fwd.slot = -1; fwd.x = fwd.y = fwd.z = 0.0;
    // One-warp chains on the energy cache synchronise after every accepted move (the cache entries must be visible), so a
    // slot is never read before its last write is visible: no register forwarding, and the hole-filling particle of a
    // deletion is read only when the deletion is accepted.
    constexpr bool NOFWD = CACHE && !MULTI && !SAMPLE;
    auto load_sel = [&](const double* px, const double* py, const double* pz, int i, const Fwd& f, double& x, double& y, double& z) {
        if (NOFWD) { x = px[i]; y = py[i]; z = pz[i]; }
        else load_slot(px, py, pz, i, f, x, y, z);
    };
    int parity = 0;
    SetCounters sc;
    double sum_n = 0.0, sum_n2 = 0.0, sum_e = 0.0, sum_e2 = 0.0;   // production accumulators (non-sampling kernels), written back at the end
    unsigned long long samples = 0ull;
    if (!SAMPLE) { sum_n = st.sum_n; sum_n2 = st.sum_n2; sum_e = st.sum_e; sum_e2 = st.sum_e2; samples = st.samples; }
    const int steps_per_set = (int)a.steps_per_set;
    double* ep = CACHE ? a.ep + (size_t)chain * (size_t)cap : nullptr;
    if (CACHE) {
        // The cache left by the previous launch is reused if the configuration is still the one it was built for
        // (checksum); otherwise it is rebuilt, which also re-anchors the running energies (what MC::BoxEnergy returns).
        const unsigned long long sum = config_checksum<MULTI>(xs, ys, zs, n, (unsigned long long*)partials, tid, T, lane, warp, W);
        if (!(S.cache_valid && S.cache_sum == sum)) {
            double c_pair0, c_wall0;
            cache_fill<FAST, PBZ, MULTI>(xs, ys, zs, ep, n, P, g, partials, W, warp, lane, tid, T, c_pair0, c_wall0);
            e_pair = c_pair0; e_wall = c_wall0;
            if (lead) st.pair_evals += (unsigned long long)n * (unsigned long long)n;
        }
    }

    int paused = 0, set0 = a.first_set, s0 = 0, p_set = 0, p_step = 0;
    const bool resuming = POOL && S.res.paused;
    if (resuming) { set0 = S.res.set; s0 = S.res.step; }
    for (int set = set0; set < a.first_set + a.n_sets && !overflow && !paused; set++) {
        if (resuming && set == set0) {   // the interrupted set goes on with the counters it had
            sc.n_disp = S.res.n_disp; sc.acc_disp = S.res.acc_disp; sc.n_ins = S.res.n_ins; sc.acc_ins = S.res.acc_ins;
            sc.n_del = S.res.n_del; sc.acc_del = S.res.acc_del; sc.n_empty = S.res.n_empty; sc.evals = S.res.evals; sc.alg = S.res.alg;
        } else {
            sc.reset();                  // MC::ResetStats, src/MC.cpp:81-98
            acc_vol = rej_vol = 0; n_volch = 1;
            if (lead) { st.widom_insert = st.widom_delete = 0.0; st.widom_insertions = st.widom_deletions = 1; }
        }
        const bool production = set > P.n_equil;
        for (int s = (resuming && set == set0) ? s0 : 0; s < steps_per_set; s++) {
#ifdef MCPM_K2_TIMING
            long long tp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            K2_PROBE(0, e_pair);
#endif
            rng.begin_step(a.seed, sid, step, lane);
            step++;
            int code;
            double sx = 0., sy = 0., sz = 0., s_po = 0., s_wo = 0., mx = 0., my = 0., mz = 0.;   // this step's MoveParticle, for Widom's slot 0
            bool s_acc = false;
            const double r = rng.template at<0>() * wsum;                         // src/main.cpp:67
            if (r <= thr_disp) {
                // ---------------- MoveParticle, src/moves.cpp:118-186 ----------------
                // draws 1,2 (box, species selection) are consumed but unused: single box, single species
                const double u_i = rng.template at<3>(), u_x = rng.template at<4>(), u_y = rng.template at<5>(),
                             u_z = rng.template at<6>(), u_a = rng.template at<7>();
                rng.end_step(8);
                const int i = (int)(u_i * nd);                             // int(u*n)+1 on 1-based slots; n==0 -> stale slot (Q12)
                double xo, yo, zo;
                load_sel(xs, ys, zs, i, fwd, xo, yo, zo);
                const double xr = __dadd_rn(xo, __dmul_rn(u_x - 0.5, dr)), yr = __dadd_rn(yo, __dmul_rn(u_y - 0.5, dr));
                const double zr = __dadd_rn(zo, __dmul_rn(u_z - 0.5, dr));
                bool odd = false;                                                 // MC::PBC: periodic axes only (hot kernels: x, y and, in bulk, z)
                const bool wx = FAST || g.px, wy = FAST || g.py, wz = FAST ? PBZ : (g.pz != 0);
                double xn = wx ? wrap_select(xr, Lx, odd) : xr, yn = wy ? wrap_select(yr, Ly, odd) : yr, zn = wz ? wrap_select(zr, Lz, odd) : zr;
                if (odd) { if (wx) xn = wrap_floor(xr, Lx); if (wy) yn = wrap_floor(yr, Ly); if (wz) zn = wrap_floor(zr, Lz); }
                double v[2];
                double wo, wn;
                K2_PROBE(1, xn + yn + zn);
                if (CACHE) {
                    const double ep_i = (n > 0) ? ep[i] : 0.0;                    // old pair sum from the cache (issued before the scan)
                    v[1] = pair_sum1<FAST, PBZ>(xs, ys, zs, n, i, g, xn, yn, zn, tid, T);
                    K2_PROBE(2, v[1]);
                    double w1[1] = {v[1]};
                    if (MULTI) {
                        wall_pair<FAST>(P, xo, yo, zo, xn, yn, zn, wo, wn, lane);
                        cta_allsum<1, MULTI>(w1, partials, parity, W, warp, lane); parity ^= 1;
                    } else {
                        wall_two_and_sum(P, zo, zn, wo, wn, w1[0], lane);     // one warp: wall and scan reductions interleaved
                        __syncwarp();
                    }
                    v[0] = ep_i; v[1] = w1[0];
                    sc.evals += (unsigned long long)n;
                } else {
                    pair_sum2<FAST, PBZ>(xs, ys, zs, n, i, g, xo, yo, zo, xn, yn, zn, tid, T, v[0], v[1]);
                    wall_pair<FAST>(P, xo, yo, zo, xn, yn, zn, wo, wn, lane);
                    cta_allsum<2, MULTI>(v, partials, parity, W, warp, lane); parity ^= 1;
                    sc.evals += 2ull * (unsigned long long)n;
                }
                K2_PROBE(3, v[1] + wn + wo);
                const double po = eps4 * v[0], pn = eps4 * v[1];
                const double dE = (pn + wn) - (po + wo);                          // newEnergy - oldEnergy, src/moves.cpp:169
                const bool acc = metropolis_translation(u_a, dE * beta);          // u < exp(-dE/T), src/moves.cpp:170-172
                K2_PROBE(4, acc ? 1.0 : 2.0);
                sc.n_disp++; sc.alg += 2ull * (unsigned long long)n;
                fwd.slot = -1;
                if (npt) { b_old = b_energy; if (acc) b_energy += 0.5 * (pn - po) + (wn - wo); }   // src/moves.cpp:149,173-180 as shipped
                if (acc) {
                    sc.acc_disp++;
                    if (n > 0) {   // exact bookkeeping: the halved box total moves by the FULL pair change (the
                                   // reference's 0.5*dp, src/moves.cpp:178, is quirk Q16); n==0 moves a ghost (Q12)
                        e_pair += pn - po; e_wall += wn - wo;
                    }
                    bool big = false;
                    if (CACHE) {   // partners' entries follow the move; the moved particle's entry is the fresh sum
                        big = cache_apply<FAST, PBZ>(xs, ys, zs, ep, n, i, g, true, xo, yo, zo, true, xn, yn, zn, tid, T);
                        if ((i & tmask) == tid) ep[i] = v[1];
                        sc.evals += 2ull * (unsigned long long)n;
                    }
                    if ((i & tmask) == tid) { xs[i] = xn; ys[i] = yn; zs[i] = zn; }
                    fwd.slot = i; fwd.x = xn; fwd.y = yn; fwd.z = zn;
                    if (CACHE && cta_sync_or<MULTI>(big)) {                        // entries visible before the next selection
                        double cp, cw;                                             // a hard overlap went through the cache: rebuild it
                        cache_fill<FAST, PBZ, MULTI>(xs, ys, zs, ep, n, P, g, partials, W, warp, lane, tid, T, cp, cw);
                        e_pair = cp; e_wall = cw; fwd.slot = -1;
                        sc.evals += (unsigned long long)n * (unsigned long long)n;
                    }
                }
                code = acc ? 1 : 0;
                if (SAMPLE) { sx = xo; sy = yo; sz = zo; s_po = po; s_wo = wo; mx = xn; my = yn; mz = zn; s_acc = acc; }
            } else if (r <= thr_vol) {
                code = (4 << 1);
                if (!npt) rng.end_step(1);   // (the hot kernels never see volume moves: the host routes n_vol > 0 to the general kernels)
                else {
                    // ---------------- ChangeVolume, NPT branch, single box: src/moves.cpp:191-231 ----------------
                    // draw 1 picks the box.  RescaleCenterOfMass runs BEFORE the widths change (ratio 1): the coordinates are NOT
                    // rescaled (quirk Q18); all three widths become ComputeBoxWidth(V') = cbrt V' (bulk, src/read.cpp:19-25).
                    const double u_v = rng.template at<2>(), u_a = rng.template at<3>();
                    rng.end_step(4);
                    n_volch++;
                    const double v_new = exp(log(volume) + (u_v - 0.5) * dv);
                    const double l_new = cbrt(v_new);
                    PairGeom g2 = g;
                    geom_set_cube(g2, l_new);
                    cta_sync<MULTI>();                                            // the previous step's slot write is visible
                    double t_pair, t_wall;
                    cta_box_energy<FAST, PBZ, MULTI>(xs, ys, zs, n, P, g2, partials, W, warp, lane, tid, T, t_pair, t_wall);   // BoxEnergy, :219
                    fwd.slot = -1;
                    sc.evals += (unsigned long long)n * (unsigned long long)n;
                    const double e_new = t_pair + t_wall;
                    const double arg0 = (e_new - b_energy) + P.press_k * (v_new - volume) - (nd + 1.0) * log(v_new / volume) * P.T;   // :220-222
                    if (u_a > exp(-arg0 / P.T)) rej_vol++;                        // rejected: box = oldBox0
                    else {
                        acc_vol++;
                        volume = v_new; Lx = Ly = Lz = l_new; g = g2;
                        e_pair = t_pair; e_wall = t_wall; b_energy = e_new;
                        code |= 1;
                    }
                }
            } else {
                // ---------------- SwapParticle, single box, src/moves.cpp:284-341 ----------------
                // draw 1 (species selection) is consumed but unused
                if (rng.template at<2>() < 0.5) {
                    if (n >= cap_slot) {
                        if (POOL && cap_slot < cap) { paused = 1; p_set = set; p_step = s; step--; }   // this step runs again after the move to a larger slot
                        else overflow = 1;
                        break;
                    }
                    const double u_x = rng.template at<3>(), u_y = rng.template at<4>(), u_z = rng.template at<5>(), u_a = rng.template at<6>();
                    rng.end_step(7);
                    double xi = u_x * Lx, yi = u_y * Ly, zi = u_z * Lz;           // InsertParticle, src/moves.cpp:46-54
                    if (!FAST && P.geometry == MCPM_GEOM_SPHERE) {                // :25-36: radius, theta, phi (not uniform in volume: as shipped)
                        const double rad = u_x * Lz * 0.5, th = u_y * 3.14159265358979323846, ph = u_z * 2 * 3.14159265358979323846;
                        xi = rad * sin(th) * cos(ph) + 0.5 * Lx; yi = rad * sin(th) * sin(ph) + 0.5 * Ly; zi = rad * cos(th) + 0.5 * Lz;
                    } else if (!FAST && P.geometry == MCPM_GEOM_CYLINDER) {       // :37-45: rho, phi, then x
                        const double rho = u_x * Lz * 0.5, ph = u_y * 2 * 3.14159265358979323846;
                        xi = (u_z - 0.5) * Lx + 0.5 * Lx; yi = rho * sin(ph) + 0.5 * Ly; zi = rho * cos(ph) + 0.5 * Lz;
                    }
                    // the trial position is written to the first free slot even if rejected (src/moves.cpp:302-303)
                    if ((n & tmask) == tid) { xs[n] = xi; ys[n] = yi; zs[n] = zi; }
                    double v[1];
                    v[0] = pair_sum1<FAST, PBZ>(xs, ys, zs, n, -1, g, xi, yi, zi, tid, T);
                    double wi, wdummy;
                    if (MULTI) {
                        wall_pair<FAST>(P, xi, yi, zi, xi, yi, zi, wi, wdummy, lane);
                        cta_allsum<1, MULTI>(v, partials, parity, W, warp, lane); parity ^= 1;
                    } else {
                        wall_two_and_sum(P, zi, zi, wi, wdummy, v[0], lane);
                        __syncwarp();
                    }
                    const double pi_ = eps4 * v[0];
                    const double e_in = pi_ + wi;
                    const bool acc = metropolis_insertion(u_a, -(e_in * beta), ln_zzV, zzV, n);   // src/moves.cpp:306-307
                    sc.n_ins++; sc.evals += (unsigned long long)n; sc.alg += (unsigned long long)n;
                    fwd.slot = -1;
                    if (acc) {
                        bool big = false;
                        if (CACHE) {
                            big = cache_apply<FAST, PBZ>(xs, ys, zs, ep, n, -1, g, false, 0., 0., 0., true, xi, yi, zi, tid, T);
                            if ((n & tmask) == tid) ep[n] = v[0];
                            sc.evals += (unsigned long long)n;
                        }
                        sc.acc_ins++; n++; nd += 1.0; e_pair += pi_ + hs_self; e_wall += wi;
                        if (CACHE && cta_sync_or<MULTI>(big)) {
                            double cp, cw;
                            cache_fill<FAST, PBZ, MULTI>(xs, ys, zs, ep, n, P, g, partials, W, warp, lane, tid, T, cp, cw);
                            e_pair = cp; e_wall = cw; fwd.slot = -1;
                            sc.evals += (unsigned long long)n * (unsigned long long)n;
                        }
                    }
                    code = (1 << 1) | (acc ? 1 : 0);
                } else if (n > 0) {
                    const double u_o = rng.template at<3>(), u_a = rng.template at<4>();
                    rng.end_step(5);
                    const int o = (int)(u_o * nd + 1.0) - 1;               // int(u*n+1) on 1-based slots
                    double xo, yo, zo, xl = 0., yl = 0., zl = 0.;
                    load_sel(xs, ys, zs, o, fwd, xo, yo, zo);
                    if (!NOFWD) load_slot(xs, ys, zs, n - 1, fwd, xl, yl, zl);    // the particle that would fill the hole
                    double v[1];
                    double wo, wdummy;
                    if (CACHE) {
                        v[0] = ep[o];                                             // no partner scan at all
                        wall_pair<FAST>(P, xo, yo, zo, xo, yo, zo, wo, wdummy, lane);
                    } else {
                        v[0] = pair_sum1<FAST, PBZ>(xs, ys, zs, n, o, g, xo, yo, zo, tid, T);
                        wall_pair<FAST>(P, xo, yo, zo, xo, yo, zo, wo, wdummy, lane);
                        cta_allsum<1, MULTI>(v, partials, parity, W, warp, lane); parity ^= 1;
                        fwd.slot = -1;
                        sc.evals += (unsigned long long)n;
                    }
                    const double po = eps4 * v[0];
                    const double e_out = po + wo;
                    const bool acc = metropolis_deletion(u_a, e_out * beta, ln_zzV, zzV, n);      // src/moves.cpp:326-327
                    sc.n_del++; sc.alg += (unsigned long long)n;
                    if (acc) {
                        sc.acc_del++;
                        bool big = false;
                        if (CACHE) {   // the leaving particle drops out of every partner's sum; then the last entry fills the hole
                            big = cache_apply<FAST, PBZ>(xs, ys, zs, ep, n, o, g, true, xo, yo, zo, false, 0., 0., 0., tid, T);
                            sc.evals += (unsigned long long)n;
                            big = cta_sync_or<MULTI>(big);
                            if ((o & tmask) == tid) ep[o] = ep[n - 1];
                            fwd.slot = -1;                                          // barrier passed: earlier slot writes are visible
                        }
                        if (NOFWD && (o & tmask) == tid) { xl = xs[n - 1]; yl = ys[n - 1]; zl = zs[n - 1]; }   // read only when needed, by the writer
                        if ((o & tmask) == tid) { xs[o] = xl; ys[o] = yl; zs[o] = zl; }  // swap-with-last, src/moves.cpp:329
                        fwd.slot = o; fwd.x = xl; fwd.y = yl; fwd.z = zl;
                        n--; nd -= 1.0;
                        e_pair -= po - hs_self; e_wall -= wo;   // exact bookkeeping with the fresh energies (quirks Q3, Q16 fixed)
                        if (CACHE) {
                            cta_sync<MULTI>();                                      // the moved entry and slot are visible
                            if (big) {
                                double cp, cw;
                                cache_fill<FAST, PBZ, MULTI>(xs, ys, zs, ep, n, P, g, partials, W, warp, lane, tid, T, cp, cw);
                                e_pair = cp; e_wall = cw; fwd.slot = -1;
                                sc.evals += (unsigned long long)n * (unsigned long long)n;
                            }
                        }
                    }
                    code = (2 << 1) | (acc ? 1 : 0);
                } else {
                    rng.end_step(3);
                    sc.n_empty++;
                    code = (3 << 1);   // no barrier executed: fwd stays valid
                }
            }
            if (SAMPLE && production && widom_on) {
                // ---------------- ComputeWidom, src/moves.cpp:410-449 (NVT branch) ----------------
                rng.load_extra(a.seed, sid, step - 1, lane);
                double energy;                                                   // draw 0 (species) is consumed but unused
                if (rng.template extra<1>() < 0.5) {                             // ghost insertion
                    double gx = rng.template extra<2>() * Lx, gy = rng.template extra<3>() * Ly, gz = rng.template extra<4>() * Lz;
                    if (!FAST && P.geometry == MCPM_GEOM_SPHERE) {                // InsertParticle's sphere branch, src/moves.cpp:25-36
                        const double rad = rng.template extra<2>() * Lz * 0.5, th = rng.template extra<3>() * 3.14159265358979323846, ph = rng.template extra<4>() * 2 * 3.14159265358979323846;
                        gx = rad * sin(th) * cos(ph) + 0.5 * Lx; gy = rad * sin(th) * sin(ph) + 0.5 * Ly; gz = rad * cos(th) + 0.5 * Lz;
                    }
                    rng.end_step(5);
                    double v[1];
                    v[0] = pair_sum1<FAST, PBZ>(xs, ys, zs, n, -1, g, gx, gy, gz, tid, T);
                    double wg, wdummy;
                    wall_pair<FAST>(P, gx, gy, gz, gx, gy, gz, wg, wdummy, lane);
                    cta_allsum<1, MULTI>(v, partials, parity, W, warp, lane); parity ^= 1;
                    fwd.slot = -1;
                    energy = eps4 * v[0] + wg;
                    if (lead) { st.widom_insertions++; st.widom_insert += bar_weight(energy, beta) * exp(-0.5 * energy * beta); st.pair_evals += (unsigned long long)n; st.pair_evals_algorithmic += (unsigned long long)n; }
                } else {                                                         // virtual deletion
                    const int slot1 = (int)(rng.template extra<2>() * (double)n);   // 1-based slot in [0, n-1]: 0 is the scratch slot (Q14)
                    rng.end_step(3);
                    if (slot1 == 0) {
                        // slot 0 holds the old position of this step's MoveParticle (src/moves.cpp:152).  Rejected move: the
                        // particle sits there again and is skipped by position equality -> its old energy.  Accepted
                        // move: a ghost at the old position that also sees the moved particle.
                        double extra = 0.0;
                        if (s_acc) pair_add<FAST>(extra, dist2<FAST, PBZ>(g, sx, sy, sz, mx, my, mz), g, 0, -1);
                        energy = (s_po + eps4 * extra) + s_wo;
                    } else {
                        const int k = slot1 - 1;
                        double xk, yk, zk;
                        load_slot(xs, ys, zs, k, fwd, xk, yk, zk);
                        double v[1];
                        double wk, wdummy;
                        if (CACHE) {
                            v[0] = ep[k];                                         // a real particle's pair sum is in the cache
                            wall_pair<FAST>(P, xk, yk, zk, xk, yk, zk, wk, wdummy, lane);
                        } else {
                            v[0] = pair_sum1<FAST, PBZ>(xs, ys, zs, n, k, g, xk, yk, zk, tid, T);
                            wall_pair<FAST>(P, xk, yk, zk, xk, yk, zk, wk, wdummy, lane);
                            cta_allsum<1, MULTI>(v, partials, parity, W, warp, lane); parity ^= 1;
                            fwd.slot = -1;
                            if (lead) st.pair_evals += (unsigned long long)n;
                        }
                        energy = eps4 * v[0] + wk;
                    }
                    if (lead) {
                        st.pair_evals_algorithmic += (unsigned long long)n;
                        st.widom_deletions++;
                        const double bw = bar_weight(energy, beta);
                        if (exp(0.5 * energy * beta) < 1) st.widom_delete += bw * exp(0.5 * energy * beta);
                        else st.widom_delete += bw * exp(-0.5 * energy * beta);   // the reference's boundary for deletion statistics
                    }
                }
            }
            if (SAMPLE && production && a.sample_rdf) {
                __syncthreads();                                                 // this step's slot write is visible to everyone
                fwd.slot = -1;
                cta_rdf<FAST, PBZ>(xs, ys, zs, n, P, hist, reinterpret_cast<unsigned int*>(hist + 104), tid, T, warp, W);
                __syncthreads();
            }
#ifdef MCPM_K2_TIMING
            if (lane == 0 && warp == 0 && tp[4] != 0 && a.rdf) {
                const long long tend = k2_probe(e_pair + e_wall);
                double* o = a.rdf + (size_t)chain * MCPM_RDF_BINS;
                o[0] += (double)(tp[1] - tp[0]); o[1] += (double)(tp[2] - tp[1]); o[2] += (double)(tp[3] - tp[2]); o[3] += (double)(tp[4] - tp[3]);
                o[4] += (double)(tend - tp[4]); o[5] += 1.0;
            }
#endif
            if (npt && b_energy / b_old > 1) b_energy = e_pair + e_wall;        // MC::CorrectEnergy (src/main.cpp:71): BoxEnergy == the exact energy
            if (lead && trace) trace[(unsigned long long)(set - a.first_set) * (unsigned long long)a.steps_per_set + (unsigned long long)s] = (unsigned char)code;
            if (production) {
                const double dn = nd, e = e_pair + e_wall;
                if (!SAMPLE) {   // replicated in registers like the energies: no divergent read-modify-write of shared memory per step
                    sum_n += dn; sum_n2 += dn * dn; sum_e += e; sum_e2 += e * e;
                    samples++;
                } else if (lead) {   // the sampling kernels are short of registers (Widom state): thread 0 keeps the sums in shared memory
                    st.sum_n += dn; st.sum_n2 += dn * dn; st.sum_e += e; st.sum_e2 += e * e;
                    st.samples++;
                }
            }
        }
        if (POOL && paused) break;          // the set is not over: no adjustment, no flush of its counters
        if (lead) st.dr_used = dr;          // the amplitude this set ran with (statistics are printed before the adjustment, src/main.cpp:77-83)
        // MC::AdjustMCMoves displacement part, src/moves.cpp:68-78 (ResetStats starts nDisplacements at ONE)
        if (!overflow && n > 0 && set <= P.n_equil) {
            const double ratio = (double)sc.acc_disp * 1. / (1. * (double)(sc.n_disp + 1));
            if (ratio < 0.2) dr /= 1.1;
            else dr *= 1.1;
        }
        if (npt && !overflow && n > 0) {   // AdjustMCMoves volume part, src/moves.cpp:79-84: every set, not only during equilibration
            const double ratio = (double)acc_vol * 1. / (1. * (double)n_volch);
            if (ratio < 0.5 && dv > 1e-3) dv /= 1.1;
            else dv *= 1.1;
        }
        if (lead) {
            if (npt) { st.accept_vol = acc_vol; st.reject_vol = rej_vol; st.n_vol_changes = n_volch; }
            st.acceptance = sc.acc_disp; st.rejection = sc.n_disp - sc.acc_disp; st.n_displacements = 1 + sc.n_disp;
            st.accept_swap = sc.acc_ins + sc.acc_del; st.reject_swap = (sc.n_ins + sc.n_del) - (sc.acc_ins + sc.acc_del);
            st.n_swaps = 1 + sc.n_ins + sc.n_del;
            st.tot_disp += sc.n_disp; st.tot_disp_acc += sc.acc_disp; st.tot_ins += sc.n_ins; st.tot_ins_acc += sc.acc_ins;
            st.tot_del += sc.n_del; st.tot_del_acc += sc.acc_del; st.tot_empty += sc.n_empty; st.pair_evals += sc.evals; st.pair_evals_algorithmic += sc.alg;
        }
    }
    cta_sync<MULTI>();

    double c_pair = 0.0, c_wall = 0.0;
    if (a.check_energy && !paused) {
        if (CACHE) cache_fill<FAST, PBZ, MULTI>(xs, ys, zs, ep, n, P, g, partials, W, warp, lane, tid, T, c_pair, c_wall);   // also refreshes the cache
        else cta_box_energy<FAST, PBZ, MULTI>(xs, ys, zs, n, P, g, partials, W, warp, lane, tid, T, c_pair, c_wall);
    }
    unsigned long long end_sum = 0ull;
    if (CACHE) end_sum = config_checksum<MULTI>(xs, ys, zs, n, (unsigned long long*)partials, tid, T, lane, warp, W);

    if (lead) {
        st.n = n; st.overflow = overflow; st.dr = dr; st.steps_done = step;
        st.pair = e_pair; st.wall = e_wall; st.energy = e_pair + e_wall;
        if (npt) { st.volume = volume; st.width = Lz; st.dv = dv; st.energy_as_shipped = b_energy; S.b_energy = b_energy; S.b_old = b_old; S.b_valid = 1; }
        if (!SAMPLE) { st.sum_n = sum_n; st.sum_n2 = sum_n2; st.sum_e = sum_e; st.sum_e2 = sum_e2; st.samples = samples; }
        if (a.check_energy && !paused) { st.pair_check = c_pair; st.wall_check = c_wall; st.energy_check = c_pair + c_wall; }
        if (a.check_energy == 2 && !paused) { st.pair = c_pair; st.wall = c_wall; st.energy = c_pair + c_wall; }   // adopt, like BoxEnergy does
        S.replay_pos = rng.cursor();
        st.replay_consumed = rng.cursor();
        S.cache_valid = (CACHE && !overflow) ? 1 : 0;   // kernels that do not maintain the cache leave it stale
        S.cache_sum = end_sum;
        if (POOL) {
            S.res.paused = paused;
            if (paused) {
                S.res.set = p_set; S.res.step = p_step;
                S.res.n_disp = sc.n_disp; S.res.acc_disp = sc.acc_disp; S.res.n_ins = sc.n_ins; S.res.acc_ins = sc.acc_ins;
                S.res.n_del = sc.n_del; S.res.acc_del = sc.acc_del; S.res.n_empty = sc.n_empty; S.res.evals = sc.evals; S.res.alg = sc.alg;
            }
        }
    }
}

// GLOBAL = false: the chain's positions are staged into shared memory for the whole launch (the normal case).
// GLOBAL = true : chains too large for one CTA's shared memory (capacity > ~9.5k particles, e.g. config C4 with
//                 N = 20 000) are advanced in place in global memory; their 24 B x N working set stays L2-resident
//                 (480 KB at N = 20k vs 126 MB of L2).  Plain (coherent) loads: a slot written by its owner thread
//                 is visible to the CTA after the step's __syncthreads.
template <int RNG_MODE, int MAXT, int MINB, bool GLOBAL, bool FAST, bool SAMPLE, bool CACHE>
__global__ void __launch_bounds__(MAXT, MINB) k2_chains(const ChainParams* __restrict__ params, ChainState* __restrict__ states,
                                                  const int* __restrict__ order, double* X, double* Y, double* Z, RunArgs a) {
    extern __shared__ double smem[];
    const int chain = order[blockIdx.x];
    const int cap = a.cap;
    const size_t off = (size_t)chain * (size_t)cap;
    double* xs = GLOBAL ? X + off : smem;
    double* ys = GLOBAL ? Y + off : smem + cap;
    double* zs = GLOBAL ? Z + off : smem + 2 * cap;
    // after the positions: [sampling kernels: RDF histogram, 104 doubles + 16 per-warp count histograms of 128 unsigned ints]
    // [multi-warp CTAs: 2 parities x 32 warps x 2 partial sums]
    double* hist = GLOBAL ? smem : smem + 3 * cap;
    double* partials = hist + (SAMPLE ? MCPM_RDF_SMEM_DOUBLES : 0);
    __shared__ ChainParams P;
    __shared__ ChainState S;
    const int tid = threadIdx.x, T = blockDim.x;
    if (tid == 0) { P = params[chain]; S = states[chain]; }
    if (SAMPLE) for (int k = threadIdx.x; k < MCPM_RDF_SMEM_DOUBLES; k += blockDim.x) hist[k] = 0.0;   // (also zeroes the per-warp counts)
    __syncthreads();
    if (!GLOBAL) {
        // stage every slot (not only the n live ones): stale slots are observable through quirk Q12
        for (int j = tid; j < cap; j += T) { xs[j] = X[off + j]; ys[j] = Y[off + j]; zs[j] = Z[off + j]; }
        __syncthreads();
    }

    unsigned char* trace = a.trace ? a.trace + (unsigned long long)chain * a.trace_stride : nullptr;
    // FAST (all chains of the launch have their coordinates inside the box) is decided by the host per launch;
    // only the smallest CTA variant can run with a single warp, so MULTI is a compile-time fact elsewhere.
    const bool multi = (MAXT > 64) || T > 32, pbz = P.pbc[2] != 0;
#define MCPM_GO(SRC, M, Z) chain_loop<SRC, FAST, M, Z, SAMPLE, CACHE>(P, S, rng, xs, ys, zs, partials, a, trace, cap, (uint32_t)chain, hist)
#define MCPM_DISPATCH(SRC)                                                                           \
    do {                                                                                             \
        if (MAXT > 64 || multi) { if (pbz) MCPM_GO(SRC, true, true); else MCPM_GO(SRC, true, false); } \
        else { if (pbz) MCPM_GO(SRC, false, true); else MCPM_GO(SRC, false, false); }                  \
    } while (0)
    if (RNG_MODE == MCPM_RNG_PHILOX) {
        if (MAXT <= 64 && !GLOBAL && !multi) {   // one warp per chain: uniforms through shared memory (the partials area is free)
            PhiloxSmemSource rng; rng.init(a.seed, S.stream_id, S.st.steps_done, tid & 31, partials);
            if (pbz) MCPM_GO(PhiloxSmemSource, false, true); else MCPM_GO(PhiloxSmemSource, false, false);
        } else {
            PhiloxSource rng; rng.init(a.seed, S.stream_id, S.st.steps_done, tid & 31);
            MCPM_DISPATCH(PhiloxSource);
        }
    } else {
        ReplaySource rng; rng.init(a.replay_u + (unsigned long long)chain * a.replay_stride, 0ull, a.replay_stride);
        MCPM_DISPATCH(ReplaySource);
    }
#undef MCPM_DISPATCH
#undef MCPM_GO
    __syncthreads();
    if (!GLOBAL) for (int j = tid; j < cap; j += T) { X[off + j] = xs[j]; Y[off + j] = ys[j]; Z[off + j] = zs[j]; }
    if (SAMPLE && a.sample_rdf && a.rdf) for (int k = tid; k < MCPM_RDF_BINS; k += T) a.rdf[(size_t)chain * MCPM_RDF_BINS + k] += hist[k];
    if (tid == 0) states[chain] = S;
}

#ifdef MCPM_EXPERIMENTS
// ================================================================================================
// K2p — k2_pool (EXPERIMENT, built only by `make EXPERIMENTS=1`; measured not faster: profiles/r02_pool.md):
// the one-warp-per-chain hot kernel as ONE PERSISTENT CTA PER SM with a shared-memory pool.
//
// k2_chains sizes every CTA's shared memory for the LARGEST chain of the launch (17 KB at capacity 704), which caps the residency at
// 12 chains per SM although the isotherm's mean loading is half of that and a third of its state points hold < 150 particles.  Here a CTA
// of W warps owns the SM's whole shared memory and carves it into W slots of up to three capacity classes (host: pool_layout in
// mcpm_api.cu); every warp runs the unchanged one-warp chain loop in its slot and, when its chain is done, claims the next one from
// the queue of its class (or of a smaller class): 16 warps at 128 registers — the register file, no longer shared memory, is the
// limit — and no wave quantisation (chains are handed out longest first as warps fall free).
// A chain that grows beyond its slot during the launch pauses (ChainResume) and is finished by a follow-up launch in full-size slots.
// ================================================================================================
template <int RNG_MODE, int WARPS, bool PBZ>
__global__ void __launch_bounds__(32 * WARPS, 1) k2_pool(const ChainParams* __restrict__ params, ChainState* __restrict__ states,
                                                         const int* __restrict__ order, double* X, double* Y, double* Z, RunArgs a, PoolLayout L) {
    // blockDim = (32, WARPS): threadIdx.x is the lane, threadIdx.y the warp / slot.  chain_loop then sees tid = lane and T = 32 exactly as in
    // a one-warp CTA of k2_chains, and compiles to the same hot loop.
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x, w = threadIdx.y;
    if (w >= L.n_slots) return;
    const int cap = a.cap;
    // everything the claim loop needs lives in shared memory / the constant bank and is re-read after a chain: nothing of it has to
    // stay in registers across chain_loop
    volatile int* wq = reinterpret_cast<volatile int*>(smem + L.state_off + (size_t)w * L.state_stride) + (((sizeof(ChainParams) + 15) / 16) * 16 + ((sizeof(ChainState) + 15) / 16) * 16) / 4;
    for (;;) {
        const int cap_slot = L.slot_cap[w];
        double* xs = smem + L.slot_off[w];
        double* ys = xs + cap_slot;
        double* zs = ys + cap_slot;
        double* ubuf = zs + cap_slot;                                    // 64 doubles: this warp's batch of uniforms
        ChainParams& P = *reinterpret_cast<ChainParams*>(smem + L.state_off + (size_t)w * L.state_stride);
        ChainState& S = *reinterpret_cast<ChainState*>(reinterpret_cast<char*>(&P) + ((sizeof(ChainParams) + 15) / 16) * 16);
        // ---- claim a chain: own class first, then the smaller ones ----
        int item = -1;
        if (lane == 0) {
            for (int c = L.slot_cls[w]; c < L.n_classes && item < 0; c++) {
                if (L.q_begin[c] >= L.q_end[c]) continue;
                const int k = L.q_begin[c] + atomicAdd(&L.q_next[c], 1);
                if (k < L.q_end[c]) item = k;
            }
        }
        item = __shfl_sync(MCPM_FULL_MASK, item, 0);
        if (item < 0) return;
        int chain = order[item];
        if (lane == 0) wq[0] = chain;
        {
            const size_t off = (size_t)chain * (size_t)cap;
            const int* ps = reinterpret_cast<const int*>(params + chain); int* pd = reinterpret_cast<int*>(&P);
            for (int k = lane; k < (int)(sizeof(ChainParams) / 4); k += 32) pd[k] = ps[k];
            const int* ss = reinterpret_cast<const int*>(states + chain); int* sd = reinterpret_cast<int*>(&S);
            for (int k = lane; k < (int)(sizeof(ChainState) / 4); k += 32) sd[k] = ss[k];
            for (int j = lane; j < cap_slot; j += 32) { xs[j] = X[off + j]; ys[j] = Y[off + j]; zs[j] = Z[off + j]; }
        }
        __syncwarp();
        if (S.st.n > cap_slot) {            // cannot happen with the host's class assignment; never touch memory beyond the slot
            if (lane == 0) { S.res.paused = 1; S.res.set = a.first_set; S.res.step = 0; S.res.n_disp = S.res.acc_disp = S.res.n_ins = S.res.acc_ins = 0;
                             S.res.n_del = S.res.acc_del = S.res.n_empty = 0; S.res.evals = S.res.alg = 0ull; states[chain].res = S.res; }
            __syncwarp();
            continue;
        }
        {
            unsigned char* trace = a.trace ? a.trace + (unsigned long long)chain * a.trace_stride : nullptr;
            if (RNG_MODE == MCPM_RNG_PHILOX) {
                PhiloxSmemSource rng; rng.init(a.seed, S.stream_id, S.st.steps_done, lane, ubuf);
                chain_loop<PhiloxSmemSource, true, false, PBZ, false, true, true>(P, S, rng, xs, ys, zs, ubuf, a, trace, cap, (uint32_t)chain, nullptr, cap_slot);
            } else {
                ReplaySource rng; rng.init(a.replay_u + (unsigned long long)chain * a.replay_stride, S.res.paused ? S.replay_pos : 0ull, a.replay_stride);
                chain_loop<ReplaySource, true, false, PBZ, false, true, true>(P, S, rng, xs, ys, zs, ubuf, a, trace, cap, (uint32_t)chain, nullptr, cap_slot);
            }
        }
        __syncwarp();
        chain = wq[0];                                                   // re-read: not kept in a register across the chain
        {
            const size_t off = (size_t)chain * (size_t)cap;
            for (int j = lane; j < cap_slot; j += 32) { X[off + j] = xs[j]; Y[off + j] = ys[j]; Z[off + j] = zs[j]; }
            const int* sd = reinterpret_cast<const int*>(&S); int* ss = reinterpret_cast<int*>(states + chain);
            for (int k = lane; k < (int)(sizeof(ChainState) / 4); k += 32) ss[k] = sd[k];
        }
        __syncwarp();
    }
}

size_t pool_state_stride_doubles() { return (((sizeof(ChainParams) + 15) / 16) * 16 + ((sizeof(ChainState) + 15) / 16) * 16 + 16) / sizeof(double); }   // + 16 bytes of claim-loop scratch

cudaError_t launch_k2_pool(int rng_mode, bool pbz, int warps, int n_ctas, size_t smem, const ChainParams* params, ChainState* states, const int* order,
                           double* X, double* Y, double* Z, const RunArgs& a, const PoolLayout& L, cudaStream_t stream) {
    auto go = [&](auto kern, int threads) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<n_ctas, dim3(32, threads / 32), smem, stream>>>(params, states, order, X, Y, Z, a, L);
        return cudaGetLastError();
    };
    const bool ph = rng_mode == MCPM_RNG_PHILOX;
#define MCPM_POOL_GO(W)                                                                                                          \
    do {                                                                                                                         \
        if (ph) return pbz ? go(k2_pool<MCPM_RNG_PHILOX, W, true>, 32 * W) : go(k2_pool<MCPM_RNG_PHILOX, W, false>, 32 * W);     \
        return pbz ? go(k2_pool<MCPM_RNG_REPLAY, W, true>, 32 * W) : go(k2_pool<MCPM_RNG_REPLAY, W, false>, 32 * W);             \
    } while (0)
    if (warps <= 12) MCPM_POOL_GO(12);     // 168 registers per thread (experiment: register pressure)
    if (warps <= 14) MCPM_POOL_GO(14);     // 144 registers per thread
    if (warps <= 16) MCPM_POOL_GO(16);     // 128 registers per thread
    MCPM_POOL_GO(20);                      // 96 registers per thread: more warps per scheduler, less instruction-level parallelism
#undef MCPM_POOL_GO
}

#endif  // MCPM_EXPERIMENTS

// ================================================================================================
// K3 — full box energy (MC::BoxEnergy, src/energy.cpp:106-124).
// grid = (i-tiles, chains).  Thread t of a CTA owns particle i = tile*blockDim + t in registers and
// streams all j through shared-memory tiles (every thread reads the same j: smem broadcast).
// Per-particle pairPotE / boxE are written out as BoxEnergy leaves them in the particle records;
// CTA partial sums go to scratch, the last CTA of each chain adds them in tile order.
// ================================================================================================
#define K3_TILE 512
template <bool FAST>
__global__ void __launch_bounds__(128) k3_box_energy(const ChainParams* __restrict__ params, ChainState* __restrict__ states,
                                                     const double* __restrict__ X, const double* __restrict__ Y, const double* __restrict__ Z,
                                                     int cap, int chain0, double* __restrict__ per_pair, double* __restrict__ per_wall,
                                                     double* __restrict__ scratch, unsigned int* __restrict__ counters, double* __restrict__ out,
                                                     int want_fast) {
    __shared__ ChainParams P;
    __shared__ double tx[K3_TILE], ty[K3_TILE], tz[K3_TILE];
    __shared__ double red[2][4];
    __shared__ bool last;
    const int chain = chain0 + blockIdx.y;
    if (states[chain].fast_image != want_fast) return;   // the other instantiation handles this chain
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) P = params[chain];
    __syncthreads();
    const PairGeom g = make_geom<FAST>(P);
    const int n = states[chain].st.n;
    const int tiles = (n + blockDim.x - 1) / blockDim.x;
    const size_t off = (size_t)chain * (size_t)cap;
    const double* xs = X + off; const double* ys = Y + off; const double* zs = Z + off;
    double* sc = scratch + (size_t)blockIdx.y * 2 * gridDim.x;
    const int i = blockIdx.x * blockDim.x + tid;
    const bool live = (i < n) && ((int)blockIdx.x < tiles);
    double pair_i = 0.0, wall_i = 0.0;
    if ((int)blockIdx.x < tiles) {
        double xi = 0, yi = 0, zi = 0;
        if (live) { xi = xs[i]; yi = ys[i]; zi = zs[i]; }
        double s0 = 0.0, s1 = 0.0;
        for (int j0 = 0; j0 < n; j0 += K3_TILE) {
            const int m = min(K3_TILE, n - j0);
            __syncthreads();
            for (int k = tid; k < m; k += blockDim.x) { tx[k] = xs[j0 + k]; ty[k] = ys[j0 + k]; tz[k] = zs[j0 + k]; }
            __syncthreads();
            if (P.pbc[2]) {
                int k = 0;
                for (; k + 1 < m; k += 2) {
                    pair_add<FAST>(s0, dist2<FAST, true>(g, xi, yi, zi, tx[k], ty[k], tz[k]), g, j0 + k, i);
                    pair_add<FAST>(s1, dist2<FAST, true>(g, xi, yi, zi, tx[k + 1], ty[k + 1], tz[k + 1]), g, j0 + k + 1, i);
                }
                if (k < m) pair_add<FAST>(s0, dist2<FAST, true>(g, xi, yi, zi, tx[k], ty[k], tz[k]), g, j0 + k, i);
            } else {
                int k = 0;
                for (; k + 1 < m; k += 2) {
                    pair_add<FAST>(s0, dist2<FAST, false>(g, xi, yi, zi, tx[k], ty[k], tz[k]), g, j0 + k, i);
                    pair_add<FAST>(s1, dist2<FAST, false>(g, xi, yi, zi, tx[k + 1], ty[k + 1], tz[k + 1]), g, j0 + k + 1, i);
                }
                if (k < m) pair_add<FAST>(s0, dist2<FAST, false>(g, xi, yi, zi, tx[k], ty[k], tz[k]), g, j0 + k, i);
            }
        }
        if (live) {
            pair_i = P.eps4 * (s0 + s1);
            wall_i = wall_single<FAST>(P, xi, yi, zi);
            if (per_pair) per_pair[(size_t)blockIdx.y * cap + i] = pair_i;
            if (per_wall) per_wall[(size_t)blockIdx.y * cap + i] = wall_i;
        }
    }
    double sp = warp_allsum(live ? pair_i : 0.0), sw = warp_allsum(live ? wall_i : 0.0);
    if (lane == 0) { red[0][warp] = sp; red[1][warp] = sw; }
    __syncthreads();
    if (tid == 0) {
        double a = 0, b = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { a += red[0][w]; b += red[1][w]; }
        sc[2 * blockIdx.x] = a; sc[2 * blockIdx.x + 1] = b;
        __threadfence();
        last = (atomicAdd(&counters[blockIdx.y], 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && tid == 0) {
        __threadfence();
        double a = 0, b = 0;
        for (unsigned t = 0; t < gridDim.x; t++) { a += ((volatile double*)sc)[2 * t]; b += ((volatile double*)sc)[2 * t + 1]; }
        double* o = out + 4 * (size_t)blockIdx.y;
        const double ph = 0.5 * a;                       // box.pairPotE += 0.5*pairPotE, src/energy.cpp:120
        o[0] = ph; o[1] = 0.0; o[2] = b; o[3] = ph + b;
        mcpm_chain_stats& st = states[chain].st;         // BoxEnergy also resets the running box energies
        st.pair = ph; st.wall = b; st.energy = ph + b;
        st.pair_check = ph; st.wall_check = b; st.energy_check = ph + b;
        counters[blockIdx.y] = 0u;
    }
}

// ================================================================================================
// K3c — full box energy through a cell list (bulk-like systems whose box holds >= 3 cut-off lengths per periodic
// axis, e.g. config C4: N = 20 000, L = 100 A, rcut = 12 A -> 8x8x8 cells, ~1050 candidate partners instead of
// 20 000).  Same sums as K3 / MC::BoxEnergy: every in-range pair is visited twice, the pair total is halved; only
// the order of the floating-point additions differs.  Four small kernels per call:
//   k3c_bin      cell of every particle + per-cell counts
//   k3c_scan     exclusive scan of the counts (<= 4096 cells: one CTA)
//   k3c_scatter  positions copied into cell order (coalesced reads in the energy kernel)
//   k3c_energy   one thread per particle, 27 neighbour cells
// ================================================================================================
__global__ void k3c_bin(const ChainParams* __restrict__ params, const ChainState* __restrict__ states, const double* __restrict__ X,
                        const double* __restrict__ Y, const double* __restrict__ Z, int cap, int chain0, int* __restrict__ cell_of,
                        int* __restrict__ counts) {
    const int chain = chain0 + blockIdx.y;
    const ChainParams& P = params[chain];
    const int n = states[chain].st.n;
    const CellGrid g = cell_grid(P);
    const size_t off = (size_t)chain * cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = (cell_coord(Z[off + i], P.L[2], g.nz) * g.ny + cell_coord(Y[off + i], P.L[1], g.ny)) * g.nx + cell_coord(X[off + i], P.L[0], g.nx);
        cell_of[(size_t)blockIdx.y * cap + i] = c;
        atomicAdd(&counts[(size_t)blockIdx.y * (K3C_MAXC * K3C_MAXC * K3C_MAXC) + c], 1);
    }
}

__global__ void __launch_bounds__(1024) k3c_scan(const ChainParams* __restrict__ params, int chain0, int* __restrict__ counts, int* __restrict__ starts) {
    __shared__ int part[1024];
    const int chain = chain0 + blockIdx.x;
    const CellGrid g = cell_grid(params[chain]);
    const int nc = g.nx * g.ny * g.nz, tid = threadIdx.x;
    int* cnt = counts + (size_t)blockIdx.x * (K3C_MAXC * K3C_MAXC * K3C_MAXC);
    int* st = starts + (size_t)blockIdx.x * (K3C_MAXC * K3C_MAXC * K3C_MAXC + 1);
    const int per = (nc + 1023) / 1024;
    int local = 0;
    for (int k = 0; k < per; k++) { const int c = tid * per + k; if (c < nc) local += cnt[c]; }
    part[tid] = local;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) { int v = tid >= o ? part[tid - o] : 0; __syncthreads(); part[tid] += v; __syncthreads(); }
    int run = part[tid] - local;
    for (int k = 0; k < per; k++) { const int c = tid * per + k; if (c < nc) { st[c] = run; run += cnt[c]; cnt[c] = 0; } }
    if (tid == 1023) st[nc] = part[1023];
}

__global__ void k3c_scatter(const ChainState* __restrict__ states, const double* __restrict__ X, const double* __restrict__ Y,
                            const double* __restrict__ Z, int cap, int chain0, const int* __restrict__ cell_of, int* __restrict__ counts,
                            const int* __restrict__ starts, double* __restrict__ sx, double* __restrict__ sy, double* __restrict__ sz,
                            int* __restrict__ orig) {
    const int chain = chain0 + blockIdx.y;
    const int n = states[chain].st.n;
    const size_t off = (size_t)chain * cap, so = (size_t)blockIdx.y * cap;
    int* cnt = counts + (size_t)blockIdx.y * (K3C_MAXC * K3C_MAXC * K3C_MAXC);
    const int* st = starts + (size_t)blockIdx.y * (K3C_MAXC * K3C_MAXC * K3C_MAXC + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = cell_of[so + i];
        const int slot = st[c] + atomicAdd(&cnt[c], 1);
        sx[so + slot] = X[off + i]; sy[so + slot] = Y[off + i]; sz[so + slot] = Z[off + i];
        orig[so + slot] = i;
    }
}

template <bool PBZ>
__global__ void __launch_bounds__(128) k3c_energy(const ChainParams* __restrict__ params, ChainState* __restrict__ states, int cap, int chain0,
                                                  const int* __restrict__ starts, const double* __restrict__ sx, const double* __restrict__ sy,
                                                  const double* __restrict__ sz, const int* __restrict__ orig, double* __restrict__ per_pair,
                                                  double* __restrict__ per_wall, double* __restrict__ scratch, unsigned int* __restrict__ counters,
                                                  double* __restrict__ out) {
    __shared__ ChainParams P;
    __shared__ double red[2][4];
    __shared__ bool last;
    const int chain = chain0 + blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) P = params[chain];
    __syncthreads();
    if ((P.pbc[2] != 0) != PBZ) return;                 // the other instantiation handles this chain
    const PairGeom g = make_geom<true>(P);
    const CellGrid cg = cell_grid(P);
    const int n = states[chain].st.n;
    const size_t so = (size_t)blockIdx.y * cap;
    const int* st = starts + (size_t)blockIdx.y * (K3C_MAXC * K3C_MAXC * K3C_MAXC + 1);
    double* sc = scratch + (size_t)blockIdx.y * 2 * gridDim.x;
    const int p = blockIdx.x * blockDim.x + tid;
    double pair_i = 0.0, wall_i = 0.0;
    if (p < n) {
        const double xi = sx[so + p], yi = sy[so + p], zi = sz[so + p];
        const int cx = cell_coord(xi, P.L[0], cg.nx), cy = cell_coord(yi, P.L[1], cg.ny), cz = cell_coord(zi, P.L[2], cg.nz);
        double s0 = 0.0, s1 = 0.0;
        for (int dz = -1; dz <= 1; dz++) {
            int z = cz + dz;
            if (PBZ) z = (z + cg.nz) % cg.nz; else if (z < 0 || z >= cg.nz) continue;
            for (int dy = -1; dy <= 1; dy++) {
                const int y = (cy + dy + cg.ny) % cg.ny;
                for (int dx = -1; dx <= 1; dx++) {
                    const int x = (cx + dx + cg.nx) % cg.nx;
                    const int c = (z * cg.ny + y) * cg.nx + x;
                    int q = st[c];
                    const int qe = st[c + 1];
                    for (; q + 1 < qe; q += 2) {
                        pair_acc(s0, dist2<true, PBZ>(g, xi, yi, zi, sx[so + q], sy[so + q], sz[so + q]), g, q, p);
                        pair_acc(s1, dist2<true, PBZ>(g, xi, yi, zi, sx[so + q + 1], sy[so + q + 1], sz[so + q + 1]), g, q + 1, p);
                    }
                    if (q < qe) pair_acc(s0, dist2<true, PBZ>(g, xi, yi, zi, sx[so + q], sy[so + q], sz[so + q]), g, q, p);
                }
            }
        }
        pair_i = P.eps4 * (s0 + s1);
        wall_i = wall_one(P, zi);
        const int i = orig[so + p];
        if (per_pair) per_pair[(size_t)blockIdx.y * cap + i] = pair_i;
        if (per_wall) per_wall[(size_t)blockIdx.y * cap + i] = wall_i;
    }
    double sp = warp_allsum(pair_i), sw = warp_allsum(wall_i);
    if (lane == 0) { red[0][warp] = sp; red[1][warp] = sw; }
    __syncthreads();
    if (tid == 0) {
        double a = 0, b = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { a += red[0][w]; b += red[1][w]; }
        sc[2 * blockIdx.x] = a; sc[2 * blockIdx.x + 1] = b;
        __threadfence();
        last = (atomicAdd(&counters[blockIdx.y], 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && tid == 0) {
        __threadfence();
        double a = 0, b = 0;
        for (unsigned t = 0; t < gridDim.x; t++) { a += ((volatile double*)sc)[2 * t]; b += ((volatile double*)sc)[2 * t + 1]; }
        double* o = out + 4 * (size_t)blockIdx.y;
        const double ph = 0.5 * a;
        o[0] = ph; o[1] = 0.0; o[2] = b; o[3] = ph + b;
        mcpm_chain_stats& stt = states[chain].st;
        stt.pair = ph; stt.wall = b; stt.energy = ph + b;
        stt.pair_check = ph; stt.wall_check = b; stt.energy_check = ph + b;
        counters[blockIdx.y] = 0u;
    }
}

// ================================================================================================
__global__ void fp64_peak_kernel(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.9999999, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

__global__ void philox_fill_kernel(unsigned long long seed, unsigned int chain, unsigned long long first_step, int n_steps, double* out) {
    const int lane = threadIdx.x & 31;
    PhiloxSource rng; rng.init(seed, chain, first_step, lane);
    for (int s = 0; s < n_steps; s++) {
        rng.begin_step(seed, chain, first_step + (unsigned long long)s, lane);
        double v[8] = {rng.at<0>(), rng.at<1>(), rng.at<2>(), rng.at<3>(), rng.at<4>(), rng.at<5>(), rng.at<6>(), rng.at<7>()};
        if (lane == (s & 31)) for (int k = 0; k < 8; k++) out[(size_t)s * 8 + k] = v[k];
    }
}

// The device Metropolis tests (mcpm_device.cuh) on caller-supplied inputs, for the parity tests of the screen:
// kind 0: translation (x = dE/T), 1: insertion (x = -E/T), 2: deletion (x = E/T).
__global__ void metropolis_probe_kernel(int kind, int count, const double* __restrict__ u, const double* __restrict__ x, const int* __restrict__ n,
                                        double ln_zzV, double zzV, int* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    bool r;
    if (kind == 0) r = metropolis_translation(u[k], x[k]);
    else if (kind == 1) r = metropolis_insertion(u[k], x[k], ln_zzV, zzV, n[k]);
    else r = metropolis_deletion(u[k], x[k], ln_zzV, zzV, n[k]);
    out[k] = r ? 1 : 0;
}

// ================================================================================================
// launchers
// ================================================================================================
// positions + (multi-warp CTAs) 128 doubles of partial sums or (one warp) 64 doubles of uniforms + (sampling kernels) the RDF histogram
static size_t k2_smem_bytes(int cap, bool multi, bool sample) { return (size_t)(3 * cap + (multi ? 128 : 64) + (sample ? MCPM_RDF_SMEM_DOUBLES : 0)) * sizeof(double); }

size_t k2_smem_limit_cap(int max_smem_optin) { return ((size_t)max_smem_optin / sizeof(double) - 128 - MCPM_RDF_SMEM_DOUBLES - 64) / 3; }

template <int RNG_MODE>
#ifndef MCPM_K2_MINB
#define MCPM_K2_MINB 8      // CTAs per SM ptxas plans for in the one-warp hot kernel: 8 -> 128 registers
#endif
static cudaError_t launch_k2_rng(int n_chains, int threads, bool global, bool fast, bool sample, bool cache, size_t smem,
                                 const ChainParams* params, ChainState* states, const int* order, double* X, double* Y, double* Z,
                                 const RunArgs& a, cudaStream_t stream) {
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<n_chains, threads, smem, stream>>>(params, states, order, X, Y, Z, a);
        return cudaSuccess;
    };
    // template arguments: <rng, max threads, min CTAs/SM, GLOBAL, FAST, SAMPLE, CACHE>
    // The hot kernels (shared-memory positions, coordinates inside the box, no in-kernel sampling) exist for CTAs of up
    // to 64 / 256 / 512 threads (128 registers each; ptxas fits the 32-thread kernel into 96 registers without spills,
    // but measured on B200 that costs more ILP than the extra occupancy returns: C2 7.2e8 vs 7.7e8 moves/s), each with
    // and without the per-particle energy cache.  The sampling kernels (Widom + RDF after every production step), the
    // global-memory kernels and the general-minimum-image kernels (coordinates outside the box, never cached) exist for
    // one CTA shape each; k2_threads_supported() clamps the thread count accordingly.
    if (!fast) return global ? go(k2_chains<RNG_MODE, 512, 1, true, false, true, false>) : go(k2_chains<RNG_MODE, 256, 2, false, false, true, false>);
    if (sample) {
        if (global) return cache ? go(k2_chains<RNG_MODE, 512, 1, true, true, true, true>) : go(k2_chains<RNG_MODE, 512, 1, true, true, true, false>);
        return cache ? go(k2_chains<RNG_MODE, 256, 2, false, true, true, true>) : go(k2_chains<RNG_MODE, 256, 2, false, true, true, false>);
    }
    if (global) return cache ? go(k2_chains<RNG_MODE, 512, 1, true, true, false, true>) : go(k2_chains<RNG_MODE, 512, 1, true, true, false, false>);
    if (cache) {
        if (threads <= 64) return go(k2_chains<RNG_MODE, 64, MCPM_K2_MINB, false, true, false, true>);
        if (threads <= 256) return go(k2_chains<RNG_MODE, 256, 2, false, true, false, true>);
        return go(k2_chains<RNG_MODE, 512, 1, false, true, false, true>);
    }
    if (threads <= 64) return go(k2_chains<RNG_MODE, 64, 8, false, true, false, false>);
    if (threads <= 256) return go(k2_chains<RNG_MODE, 256, 2, false, true, false, false>);
    return go(k2_chains<RNG_MODE, 512, 1, false, true, false, false>);
}

int k2_threads_supported(int threads, bool global, bool fast, bool sample) {
    if ((!fast || sample) && !global) return std::min(threads, 256);
    return std::min(threads, 512);
}

bool k2_cache_supported(bool global, bool fast, bool sample) { (void)global; (void)sample; return fast; }

cudaError_t launch_k2(int rng_mode, int n_chains, int threads, bool fast, bool sample, bool cache, const ChainParams* params,
                      ChainState* states, const int* order, double* X, double* Y, double* Z, const RunArgs& a, int max_smem_optin,
                      cudaStream_t stream) {
    const bool global = (size_t)a.cap > k2_smem_limit_cap(max_smem_optin);
    const bool multi = threads > 32 || !fast || sample;   // only the smallest hot kernel runs single-warp CTAs
    const size_t smem = k2_smem_bytes(global ? 0 : a.cap, multi, !fast || sample);
    cache = cache && k2_cache_supported(global, fast, sample) && a.ep != nullptr;
    cudaError_t err = rng_mode == MCPM_RNG_PHILOX
                          ? launch_k2_rng<MCPM_RNG_PHILOX>(n_chains, threads, global, fast, sample, cache, smem, params, states, order, X, Y, Z, a, stream)
                          : launch_k2_rng<MCPM_RNG_REPLAY>(n_chains, threads, global, fast, sample, cache, smem, params, states, order, X, Y, Z, a, stream);
    if (err != cudaSuccess) return err;
    return cudaGetLastError();
}

cudaError_t launch_k3(int tiles, int n_chains, int chain0, const ChainParams* params, ChainState* states, const double* X, const double* Y,
                      const double* Z, int cap, double* per_pair, double* per_wall, double* scratch, unsigned int* counters, double* out,
                      cudaStream_t stream) {
    dim3 grid(tiles, n_chains);
    // two instantiations (fast / general minimum image); each CTA exits at once if the chain is the other kind
    k3_box_energy<true><<<grid, 128, 0, stream>>>(params, states, X, Y, Z, cap, chain0, per_pair, per_wall, scratch, counters, out, 1);
    k3_box_energy<false><<<grid, 128, 0, stream>>>(params, states, X, Y, Z, cap, chain0, per_pair, per_wall, scratch, counters, out, 0);
    return cudaGetLastError();
}

// fast_image flag of every chain after a bulk upload: all coordinates on periodic axes inside [0, L] (and finite)
__global__ void in_box_kernel(const ChainParams* __restrict__ params, const int* __restrict__ n_of, const double* __restrict__ X,
                              const double* __restrict__ Y, const double* __restrict__ Z, int cap, int* __restrict__ flags) {
    const int chain = blockIdx.x;
    const ChainParams& P = params[chain];
    const int n = n_of[chain];
    const size_t off = (size_t)chain * cap;
    int bad = P.exotic ? 1 : 0;   // exotic systems always take the general kernels
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const double x = X[off + j], y = Y[off + j], z = Z[off + j];
        if (P.pbc[0] && !(x >= 0.0 && x <= P.L[0])) bad = 1;
        if (P.pbc[1] && !(y >= 0.0 && y <= P.L[1])) bad = 1;
        if (P.pbc[2] && !(z >= 0.0 && z <= P.L[2])) bad = 1;
    }
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) flags[chain] = bad ? 0 : 1;
}
cudaError_t launch_in_box(int n_chains, const ChainParams* params, const int* n_of, const double* X, const double* Y, const double* Z, int cap,
                          int* flags, cudaStream_t stream) {
    in_box_kernel<<<n_chains, 128, 0, stream>>>(params, n_of, X, Y, Z, cap, flags);
    return cudaGetLastError();
}

size_t k3c_cells_per_chain() { return (size_t)K3C_MAXC * K3C_MAXC * K3C_MAXC; }

// Cell-list K3 for chains [chain0, chain0+n_chains): all of them must qualify (see mcpm_api.cu::cell_list_ok).
cudaError_t launch_k3_cells(int maxn, int n_chains, int chain0, const ChainParams* params, ChainState* states, const double* X,
                            const double* Y, const double* Z, int cap, int* cell_of, int* counts, int* starts, double* sx, double* sy,
                            double* sz, int* orig, double* per_pair, double* per_wall, double* scratch, unsigned int* counters, double* out,
                            cudaStream_t stream) {
    const int blocks = (maxn + 255) / 256, tiles = (maxn + 127) / 128;
    cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int) * k3c_cells_per_chain() * n_chains, stream);
    if (e != cudaSuccess) return e;
    k3c_bin<<<dim3(blocks, n_chains), 256, 0, stream>>>(params, states, X, Y, Z, cap, chain0, cell_of, counts);
    k3c_scan<<<n_chains, 1024, 0, stream>>>(params, chain0, counts, starts);
    k3c_scatter<<<dim3(blocks, n_chains), 256, 0, stream>>>(states, X, Y, Z, cap, chain0, cell_of, counts, starts, sx, sy, sz, orig);
    k3c_energy<true><<<dim3(tiles, n_chains), 128, 0, stream>>>(params, states, cap, chain0, starts, sx, sy, sz, orig, per_pair, per_wall, scratch, counters, out);
    k3c_energy<false><<<dim3(tiles, n_chains), 128, 0, stream>>>(params, states, cap, chain0, starts, sx, sy, sz, orig, per_pair, per_wall, scratch, counters, out);
    return cudaGetLastError();
}

cudaError_t launch_fp64_peak(double* out, int blocks, int threads, int iters, cudaStream_t stream) {
    fp64_peak_kernel<<<blocks, threads, 0, stream>>>(out, iters, 1.0);
    return cudaGetLastError();
}

cudaError_t launch_metropolis_probe(int kind, int count, const double* u, const double* x, const int* n, double ln_zzV, double zzV, int* out,
                                    cudaStream_t stream) {
    metropolis_probe_kernel<<<(count + 127) / 128, 128, 0, stream>>>(kind, count, u, x, n, ln_zzV, zzV, out);
    return cudaGetLastError();
}

cudaError_t launch_philox_fill(unsigned long long seed, unsigned int chain, unsigned long long first_step, int n_steps, double* out,
                               cudaStream_t stream) {
    philox_fill_kernel<<<1, 32, 0, stream>>>(seed, chain, first_step, n_steps, out);
    return cudaGetLastError();
}

}  // namespace mcpm
