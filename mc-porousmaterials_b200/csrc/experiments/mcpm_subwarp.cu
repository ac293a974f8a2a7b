// K2s — sub-warp chains: the throughput kernel for many chains.
//
// One CTA = one warp = 1, 2 or 4 chains (32, 16 or 8 lanes each), chosen per CTA by the host from the chains' loadings
// so that every CTA needs the same shared memory (cap_cta particles in total) and ALL CTAs of a launch run the same
// code.  The per-step work that does not scale with N (uniforms, selection, wall term, reductions, Metropolis test,
// bookkeeping: ~450 instructions, most of a step for N < 300) is executed once per WARP, so packing 2 or 4 small
// chains into a warp divides it by 2 or 4.  To keep the groups of a warp in lockstep the step is written without
// move-type branches: translation, insertion and deletion are one code path with per-group selects
//   selected slot | trial position | partner scan of the trial position (length 0 for a deletion) | wall terms |
//   ln f and the bounds test | exp only if some group is undecided | second pass over the partners if some group accepted
// which is possible because of the per-particle energy cache (see mcpm_kernels.cu): old energies are read, not summed.
// Semantics are those of chain_loop<..., CACHE = true> — same uniforms, same decisions, same positions bit for bit —
// and are tested against the same golden traces of the compiled reference.
#include <algorithm>

#include "../mcpm_device.cuh"
#include "../mcpm_kernels.h"

namespace mcpm {

namespace {

// xor-butterfly sum over the G lanes of a group (G = 8, 16, 32): every lane of the group gets the same bits.
__device__ __forceinline__ double group_allsum(double v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(MCPM_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long group_allsum_u64(unsigned long long v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(MCPM_FULL_MASK, v, o);
    return v;
}

// Per-group Philox batch: lane l of a group evaluates call 4*(batch_step + l/4) + l%4 of ITS chain; G/4 steps per batch.
struct SubPhilox {
    double ua, ub;
    int bi;   // local steps since the batch was filled (same in every group: the groups run in lockstep)
    __device__ __forceinline__ void init() { ua = ub = 0.0; bi = 1 << 30; }
    __device__ __forceinline__ void begin_step(unsigned long long seed, uint32_t chain, unsigned long long step, int l, int G) {
        if (bi >= (G >> 2)) {
            philox_call(seed, chain, 4ull * (step + (unsigned long long)(l >> 2)) + (unsigned long long)(l & 3), ua, ub);
            bi = 0;
        }
    }
    template <int K>
    __device__ __forceinline__ double at(int G) const { return __shfl_sync(MCPM_FULL_MASK, (K & 1) ? ub : ua, bi * 4 + (K >> 1), G); }
    __device__ __forceinline__ void end_step(int) { bi++; }
    __device__ __forceinline__ unsigned long long cursor() const { return 0ull; }
};
struct SubReplay {
    const double* u;
    unsigned long long pos, n;
    __device__ __forceinline__ void init() {}
    __device__ __forceinline__ void begin_step(unsigned long long, uint32_t, unsigned long long, int, int) {}
    template <int K>
    __device__ __forceinline__ double at(int) const { return (pos + K < n) ? __ldg(u + pos + K) : 0.0; }
    __device__ __forceinline__ void end_step(int count) { pos += (unsigned long long)count; }
    __device__ __forceinline__ unsigned long long cursor() const { return pos; }
};

// Partner scan of one position by the l-th lane of a group: own slots j == l (mod G), j < n_scan.
template <bool PBZ>
__device__ __forceinline__ double scan_one(const double* xs, const double* ys, const double* zs, int n_scan, int n_loop, int self,
                                           const PairGeom& g, double px, double py, double pz, int l, int G, int last_slot) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int j = l;
    // n_loop = the largest n_scan in the warp: all groups run the same trip count; slots beyond a group's own n_scan are
    // masked by making them look like "self" (their loads are clamped to a valid slot)
    for (; j + 3 * G < n_loop; j += 4 * G) {
        const int j0 = min(j, last_slot), j1 = min(j + G, last_slot), j2 = min(j + 2 * G, last_slot), j3 = min(j + 3 * G, last_slot);
        pair_acc(s0, dist2<true, PBZ>(g, px, py, pz, xs[j0], ys[j0], zs[j0]), g, (j < n_scan) ? j : self, self);
        pair_acc(s1, dist2<true, PBZ>(g, px, py, pz, xs[j1], ys[j1], zs[j1]), g, (j + G < n_scan) ? j + G : self, self);
        pair_acc(s2, dist2<true, PBZ>(g, px, py, pz, xs[j2], ys[j2], zs[j2]), g, (j + 2 * G < n_scan) ? j + 2 * G : self, self);
        pair_acc(s3, dist2<true, PBZ>(g, px, py, pz, xs[j3], ys[j3], zs[j3]), g, (j + 3 * G < n_scan) ? j + 3 * G : self, self);
    }
    for (; j < n_loop; j += G) {
        const int j0 = min(j, last_slot);
        pair_acc(s0, dist2<true, PBZ>(g, px, py, pz, xs[j0], ys[j0], zs[j0]), g, (j < n_scan) ? j : self, self);
    }
    return (s0 + s1) + (s2 + s3);
}

// Cache rebuild / full energy of the groups flagged by `doit` (per-group flag, identical within a group).
template <bool PBZ>
__device__ void sub_cache_fill(const double* xs, const double* ys, const double* zs, double* ep, int n, bool doit, const ChainParams& P,
                               const PairGeom& g, int l, int G, double& pair_half, double& wall) {
    double vp = 0.0, vw = 0.0;
    const int nn = doit ? n : 0;
    for (int i = l; i < nn; i += G) {
        const double xi = xs[i], yi = ys[i], zi = zs[i];
        double s0 = 0.0, s1 = 0.0;
        int j = 0;
        for (; j + 1 < nn; j += 2) {
            pair_acc(s0, dist2<true, PBZ>(g, xi, yi, zi, xs[j], ys[j], zs[j]), g, j, i);
            pair_acc(s1, dist2<true, PBZ>(g, xi, yi, zi, xs[j + 1], ys[j + 1], zs[j + 1]), g, j + 1, i);
        }
        if (j < nn) pair_acc(s0, dist2<true, PBZ>(g, xi, yi, zi, xs[j], ys[j], zs[j]), g, j, i);
        ep[i] = s0 + s1;
        vp += s0 + s1;
        vw += wall_one(P, zi);
    }
    __syncwarp();
    vp = group_allsum(vp, G); vw = group_allsum(vw, G);
    pair_half = 0.5 * (P.eps4 * vp);
    wall = vw;
}

__device__ __forceinline__ unsigned long long sub_checksum(const double* xs, const double* ys, const double* zs, int n, int l, int G) {
    unsigned long long h = 0ull;
    for (int j = l; j < n; j += G)
        h += (unsigned long long)__double_as_longlong(xs[j]) * 0x9E3779B97F4A7C15ull + (unsigned long long)__double_as_longlong(ys[j]) * 0xC2B2AE3D27D4EB4Full +
             (unsigned long long)__double_as_longlong(zs[j]) * 0x165667B19E3779F9ull + (unsigned long long)(j + 1) * 0xD6E8FEB86659FD93ull;
    // NOTE: must equal config_checksum<false> of mcpm_kernels.cu for the same configuration (cache shared between kernels)
    return group_allsum_u64(h, G) + (unsigned long long)n;
}

}  // namespace

template <int RNG_MODE, bool PBZ>
__global__ void __launch_bounds__(32, 16) k2_sub(const ChainParams* __restrict__ params, ChainState* __restrict__ states,
                                                 const int* __restrict__ order, const int2* __restrict__ ctas, double* __restrict__ X,
                                                 double* __restrict__ Y, double* __restrict__ Z, RunArgs a, int cap_cta) {
    extern __shared__ double smem[];
    __shared__ ChainParams Ps[4];
    __shared__ ChainState Ss[4];
    const int lane = threadIdx.x;
    const int2 cd = ctas[blockIdx.x];
    const int k = cd.y & 0xff, valid = (cd.y >> 8) & 0xff;       // chains per CTA (1, 2, 4) and how many of them exist
    const int G = 32 / k, g = lane / G, l = lane - g * G;
    const bool active = g < valid;
    const int chain = order[cd.x + (active ? g : 0)];
    const int cap_sub = min(cap_cta / k, a.cap);                   // slots of this chain held in shared memory
    if (active && l == 0) { Ps[g] = params[chain]; Ss[g] = states[chain]; }
    __syncwarp();
    const ChainParams& P = Ps[active ? g : 0];
    ChainState& S = Ss[active ? g : 0];
    mcpm_chain_stats& st = S.st;
    double* xs = smem + (size_t)g * 3 * (cap_cta / k);
    double* ys = xs + cap_cta / k;
    double* zs = ys + cap_cta / k;
    const size_t off = (size_t)chain * (size_t)a.cap;
    if (active) for (int j = l; j < cap_sub; j += G) { xs[j] = X[off + j]; ys[j] = Y[off + j]; zs[j] = Z[off + j]; }
    __syncwarp();

    const bool lead = active && l == 0;
    const PairGeom geo = make_geom<true>(P);
    double* ep = a.ep + off;
    unsigned char* trace = a.trace ? a.trace + (unsigned long long)chain * a.trace_stride : nullptr;

    int n = st.n;
    double dr = st.dr, e_pair = st.pair, e_wall = st.wall;
    unsigned long long step = st.steps_done;
    bool alive = active;
    int overflow = 0, paused = 0;
    int set0 = a.first_set, s0 = 0;
    int c_ndisp = 0, c_adisp = 0, c_nins = 0, c_ains = 0, c_ndel = 0, c_adel = 0, c_nempty = 0;
    unsigned long long c_evals = 0ull, c_alg = 0ull;
    if (k == 1 && S.res.paused) {       // continue a chain that outgrew its sub-warp slot in an earlier launch of this call
        set0 = S.res.set; s0 = S.res.step;
        c_ndisp = S.res.n_disp; c_adisp = S.res.acc_disp; c_nins = S.res.n_ins; c_ains = S.res.acc_ins;
        c_ndel = S.res.n_del; c_adel = S.res.acc_del; c_nempty = S.res.n_empty; c_evals = S.res.evals; c_alg = S.res.alg;
    }
    if (alive && n >= cap_sub && cap_sub < a.cap) {   // wrong class: hand back untouched
        alive = false; paused = 1;
        if (lead) { S.res.set = set0; S.res.step = s0; }
    }

    // energy cache: reuse the previous launch's if it belongs to this configuration, else rebuild (and re-anchor)
    {
        const unsigned long long sum = sub_checksum(xs, ys, zs, alive ? n : 0, l, G);
        const bool rebuild = alive && !(S.cache_valid && S.cache_sum == sum);
        if (__any_sync(MCPM_FULL_MASK, rebuild)) {
            double cp, cw;
            sub_cache_fill<PBZ>(xs, ys, zs, ep, n, rebuild, P, geo, l, G, cp, cw);
            if (rebuild) { e_pair = cp; e_wall = cw; }
            if (rebuild && lead) st.pair_evals += (unsigned long long)n * (unsigned long long)n;
            __syncwarp();
        }
    }

    SubPhilox prng; prng.init();
    SubReplay rrng; rrng.u = a.replay_u + (unsigned long long)chain * a.replay_stride; rrng.pos = (k == 1 && S.res.paused) ? S.replay_pos : 0ull; rrng.n = a.replay_stride;
#define DRAW(K) (RNG_MODE == MCPM_RNG_PHILOX ? prng.template at<K>(G) : rrng.template at<K>(G))

    const int steps_per_set = (int)a.steps_per_set;
    const double wsum = P.wsum, thr_disp = P.thr_disp;
    const double beta = P.beta, eps4 = P.eps4, ln_zzV = P.ln_zzV, zzV = P.zzV;
    const double Lx = P.L[0], Ly = P.L[1], Lz = P.L[2];
    const bool has_wall = (P.wall_kind == MCPM_WALL_SLIT_LJ && P.geometry == MCPM_GEOM_SLIT);
    const bool slit = (P.geometry == MCPM_GEOM_SLIT);
    const int n_terms = has_wall ? 4 * P.n_layers : 0;

    for (int set = set0; set < a.first_set + a.n_sets; set++) {
        if (!(set == set0 && s0 > 0)) { c_ndisp = c_adisp = c_nins = c_ains = c_ndel = c_adel = c_nempty = 0; c_evals = 0ull; c_alg = 0ull;
                                        if (set != set0) { /* counters of the previous set were flushed below */ } }
        const bool production = set > P.n_equil;
        for (int s = (set == set0 ? s0 : 0); s < steps_per_set; s++) {
            if (RNG_MODE == MCPM_RNG_PHILOX) prng.begin_step(a.seed, S.stream_id, step, l, G);
            // ---- what kind of step is this (src/main.cpp:67-70, src/moves.cpp:299,320) ----
            const double r = DRAW(0) * wsum;
            const bool is_t = alive && (r <= thr_disp);
            const bool swap = alive && !(r <= thr_disp);
            const double u2 = DRAW(2), u3 = DRAW(3), u4 = DRAW(4), u5 = DRAW(5), u6 = DRAW(6), u7 = DRAW(7);
            const bool ins = swap && (u2 < 0.5);
            bool is_i = ins;
            const bool is_d = swap && !ins && n > 0;
            const bool is_e = swap && !ins && n == 0;
            if (is_i && n >= cap_sub) {                 // no free slot for the trial particle
                if (cap_sub < a.cap) paused = 1; else overflow = 1;
                alive = false; is_i = false;
                if (lead) { S.res.set = set; S.res.step = s; }
            }
            const bool live = is_t || is_i || is_d;     // the group evaluates a move in this step
            // ---- selected slot and trial position ----
            const int idx = is_t ? (int)(u3 * (double)n) : (is_d ? (int)(u3 * (double)n + 1.0) - 1 : n);
            const int idc = min(max(idx, 0), cap_sub - 1);
            const double xo = xs[idc], yo = ys[idc], zo = zs[idc];
            const int lastc = max(n - 1, 0);
            const double xl = xs[lastc], yl = ys[lastc], zl = zs[lastc];     // fills the hole of a deletion
            const double xr = __dadd_rn(xo, __dmul_rn(u4 - 0.5, dr)), yr = __dadd_rn(yo, __dmul_rn(u5 - 0.5, dr));
            const double zr = __dadd_rn(zo, __dmul_rn(u6 - 0.5, dr));
            bool odd = false;
            double tx = wrap_select(xr, Lx, odd), ty = wrap_select(yr, Ly, odd), tz = PBZ ? wrap_select(zr, Lz, odd) : zr;
            if (__any_sync(MCPM_FULL_MASK, odd && is_t)) { tx = wrap_floor(xr, Lx); ty = wrap_floor(yr, Ly); if (PBZ) tz = wrap_floor(zr, Lz); }
            const double px = is_t ? tx : u3 * Lx, py = is_t ? ty : u4 * Ly, pz = is_t ? tz : u5 * Lz;   // InsertParticle: u*L
            // the trial position of an insertion is written to the first free slot even if rejected (src/moves.cpp:302-303)
            if (is_i && (n % G) == l) { xs[n] = px; ys[n] = py; zs[n] = pz; }
            // ---- partner scan of the trial position (translation, insertion); deletions read the cache ----
            const int n_scan = (is_t || is_i) ? n : 0;
            const int n_loop = __reduce_max_sync(MCPM_FULL_MASK, n_scan);
            const int self = is_t ? idx : -1;
            double v = scan_one<PBZ>(xs, ys, zs, n_scan, n_loop, self, geo, px, py, pz, l, G, lastc);
            const double ep_old = ((is_t && n > 0) || is_d) ? ep[idc] : 0.0;
            // ---- wall terms of the old and the trial position, lane-parallel within the group ----
            double a_old = 0.0, a_new = 0.0;
            for (int t = l; __any_sync(MCPM_FULL_MASK, t < n_terms); t += G) {
                const bool on = t < n_terms;
                const bool is_new = t & 1;
                double zc = (is_new ? pz : zo) - P.halfH;
                if (t & 2) zc = -zc;
                const double q = P.sig_sf * fast_rcp(P.halfH + (t >> 2) * P.delta + zc);
                const double q2 = q * q, q4 = q2 * q2;
                const double term = on ? 0.2 * (q4 * q4 * q2) - 0.5 * q4 : 0.0;
                a_new += is_new ? term : 0.0;
                a_old += is_new ? 0.0 : term;
            }
            v = group_allsum(v, G); a_old = group_allsum(a_old, G); a_new = group_allsum(a_new, G);
            double wo = a_old * P.wall_pref, wn = a_new * P.wall_pref;
            if (slit) {      // out-of-pore guard: finite INT_MAX, wall term still added (quirk Q6)
                const double za = zo - P.halfH, zb = pz - P.halfH;
                if (za * za > P.sqr_pore_r) wo = MCPM_INT_MAX_D + wo;
                if (zb * zb > P.sqr_pore_r) wn = MCPM_INT_MAX_D + wn;
            }
            // ---- Metropolis: ln f up to the ln(N) term, bounds, exp only for the undecided ----
            const double po = eps4 * ep_old, pn = eps4 * v;
            const double e_old = po + wo, e_new = pn + wn;
            const double xarg = is_t ? -((e_new - e_old) * beta) : (is_i ? -(e_new * beta) : e_old * beta);   // argument of exp
            const double tl = is_t ? xarg : (is_i ? ln_zzV + xarg : xarg - ln_zzV);
            const double hi = is_i ? 21.0 : 0.0, lo = is_d ? -58.0 : -37.0;
            const double u_acc = is_t ? u7 : (is_i ? u6 : u4);
            bool acc = live && (tl >= hi);
            const bool undecided = live && !(tl >= hi) && ((tl >= lo) || (u_acc == 0.0 && tl == tl));
            if (__any_sync(MCPM_FULL_MASK, undecided)) {
                const double ex = exp(xarg);
                const double f = is_t ? ex : (is_i ? zzV * ex / (double)(n + 1) : (double)n * ex / zzV);   // src/moves.cpp:171,306,326
                if (undecided) acc = u_acc < f;
            }
            // ---- accepted: second pass over the partners (cache), then the slot, the counts and the energies ----
            bool big = false;
            if (__any_sync(MCPM_FULL_MASK, acc)) {
                const bool has_a = acc && (is_t || is_d), has_b = acc && (is_t || is_i);
                const int n_app = acc ? n : 0;
                const int n_app_loop = __reduce_max_sync(MCPM_FULL_MASK, n_app);
                const int self2 = (is_t || is_d) ? idx : -1;
                for (int j = l; j < n_app_loop; j += G) {
                    const int jc = min(j, lastc);
                    const double xj = xs[jc], yj = ys[jc], zj = zs[jc];
                    const int jm = (j < n_app) ? j : self2;
                    double tb = pair_term(dist2<true, PBZ>(geo, px, py, pz, xj, yj, zj), geo, jm, self2);
                    double ta = pair_term(dist2<true, PBZ>(geo, xo, yo, zo, xj, yj, zj), geo, jm, self2);
                    tb = has_b ? tb : 0.0; ta = has_a ? ta : 0.0;
                    big = big || !(fabs(tb) <= 1e6) || !(fabs(ta) <= 1e6);
                    if (j < n_app && j != self2) ep[j] += tb - ta;
                }
                __syncwarp();
                if (acc) {
                    if (is_t) {
                        if ((idx % G) == l) { ep[idx] = v; xs[idx] = px; ys[idx] = py; zs[idx] = pz; }
                        if (n > 0) { e_pair += pn - po; e_wall += wn - wo; }       // n == 0 moves a ghost (Q12)
                        c_evals += 2ull * (unsigned long long)n;
                    } else if (is_i) {
                        if ((n % G) == l) ep[n] = v;
                        e_pair += pn; e_wall += wn;
                        c_evals += (unsigned long long)n;
                        n++;
                    } else {
                        if ((idx % G) == l) { ep[idx] = ep[n - 1]; xs[idx] = xl; ys[idx] = yl; zs[idx] = zl; }   // swap-with-last, src/moves.cpp:329
                        e_pair -= po; e_wall -= wo;
                        c_evals += (unsigned long long)n;
                        n--;
                    }
                }
                __syncwarp();
                big = (group_allsum(big ? 1.0 : 0.0, G) != 0.0) && acc;
                if (__any_sync(MCPM_FULL_MASK, big)) {      // a hard overlap went through the cache: rebuild it
                    double cp, cw;
                    sub_cache_fill<PBZ>(xs, ys, zs, ep, n, big, P, geo, l, G, cp, cw);
                    if (big) { e_pair = cp; e_wall = cw; c_evals += (unsigned long long)n * (unsigned long long)n; }
                    __syncwarp();
                }
            } else {
                __syncwarp();       // the trial slot written above is visible before the next step reads it
            }
            // ---- counters (replicated), trace and production accumulators (thread 0 of the group) ----
            if (is_t) { c_ndisp++; c_adisp += acc; c_evals += (unsigned long long)n_scan; c_alg += 2ull * (unsigned long long)n_scan; }
            if (is_i) { c_nins++; c_ains += acc; c_evals += (unsigned long long)n_scan; c_alg += (unsigned long long)n_scan; }
            if (is_d) { c_ndel++; c_adel += acc; c_alg += (unsigned long long)(acc ? n + 1 : n); }
            if (is_e) c_nempty++;
            const int count = is_t ? 8 : (is_i ? 7 : (is_d ? 5 : 3));
            if (alive) step++;
            if (RNG_MODE == MCPM_RNG_PHILOX) prng.end_step(count); else if (alive) rrng.end_step(count);
            if (lead && alive) {
                if (trace) trace[(unsigned long long)(set - a.first_set) * (unsigned long long)a.steps_per_set + (unsigned long long)s] =
                    (unsigned char)(is_t ? (acc ? 1 : 0) : (is_i ? (2 | (acc ? 1 : 0)) : (is_d ? (4 | (acc ? 1 : 0)) : 6)));
                if (production) {
                    const double dn = (double)n, e = e_pair + e_wall;
                    st.sum_n += dn; st.sum_n2 += dn * dn; st.sum_e += e; st.sum_e2 += e * e;
                    st.samples++;
                }
            }
        }
        // MC::AdjustMCMoves displacement part, src/moves.cpp:68-78 (ResetStats starts nDisplacements at ONE)
        if (lead && alive) st.dr_used = dr;
        if (alive && n > 0 && set <= P.n_equil) {
            const double ratio = (double)c_adisp * 1. / (1. * (double)(c_ndisp + 1));
            if (ratio < 0.2) dr /= 1.1;
            else dr *= 1.1;
        }
        if (lead && alive) {
            st.acceptance = c_adisp; st.rejection = c_ndisp - c_adisp; st.n_displacements = 1 + c_ndisp;
            st.accept_swap = c_ains + c_adel; st.reject_swap = (c_nins + c_ndel) - (c_ains + c_adel);
            st.n_swaps = 1 + c_nins + c_ndel;
            st.widom_insert = st.widom_delete = 0.0; st.widom_insertions = st.widom_deletions = 1;
            st.tot_disp += c_ndisp; st.tot_disp_acc += c_adisp; st.tot_ins += c_nins; st.tot_ins_acc += c_ains;
            st.tot_del += c_ndel; st.tot_del_acc += c_adel; st.tot_empty += c_nempty; st.pair_evals += c_evals; st.pair_evals_algorithmic += c_alg;
        }
        s0 = 0;
        if (set == set0) set0 = -1;     // only the first set of a resumed chain starts in the middle
    }
#undef DRAW
    __syncwarp();

    // end of launch: optional full recompute (also refreshes the cache), checksum, write back
    double c_pair = 0.0, c_wall = 0.0;
    const bool done = alive;            // paused / overflowed groups keep their state as it was when they stopped
    if (a.check_energy && __any_sync(MCPM_FULL_MASK, done)) {
        sub_cache_fill<PBZ>(xs, ys, zs, ep, n, done, P, geo, l, G, c_pair, c_wall);
        __syncwarp();
    }
    const unsigned long long end_sum = sub_checksum(xs, ys, zs, active ? n : 0, l, G);
    if (active) for (int j = l; j < cap_sub; j += G) { X[off + j] = xs[j]; Y[off + j] = ys[j]; Z[off + j] = zs[j]; }
    if (lead) {
        st.n = n; st.overflow = overflow; st.dr = dr; st.steps_done = step;
        st.pair = e_pair; st.wall = e_wall; st.energy = e_pair + e_wall;
        if (a.check_energy && done) {
            st.pair_check = c_pair; st.wall_check = c_wall; st.energy_check = c_pair + c_wall;
            if (a.check_energy == 2) { st.pair = c_pair; st.wall = c_wall; st.energy = c_pair + c_wall; }
        }
        S.replay_pos = rrng.cursor(); st.replay_consumed = rrng.cursor();
        S.cache_valid = overflow ? 0 : 1; S.cache_sum = end_sum;
        S.res.paused = paused;
        if (paused) {       // where to pick up, and the counters of the interrupted set
            S.res.n_disp = c_ndisp; S.res.acc_disp = c_adisp; S.res.n_ins = c_nins; S.res.acc_ins = c_ains;
            S.res.n_del = c_ndel; S.res.acc_del = c_adel; S.res.n_empty = c_nempty; S.res.evals = c_evals; S.res.alg = c_alg;
        }
        states[chain] = S;
    }
}

size_t k2_sub_smem_bytes(int cap_cta) { return (size_t)3 * cap_cta * sizeof(double); }

cudaError_t launch_k2_sub(int rng_mode, bool pbz, int n_ctas, const ChainParams* params, ChainState* states, const int* order, const int2* ctas,
                          double* X, double* Y, double* Z, const RunArgs& a, int cap_cta, cudaStream_t stream) {
    const size_t smem = k2_sub_smem_bytes(cap_cta);
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<n_ctas, 32, smem, stream>>>(params, states, order, ctas, X, Y, Z, a, cap_cta);
        return cudaGetLastError();
    };
    if (rng_mode == MCPM_RNG_PHILOX) return pbz ? go(k2_sub<MCPM_RNG_PHILOX, true>) : go(k2_sub<MCPM_RNG_PHILOX, false>);
    return pbz ? go(k2_sub<MCPM_RNG_REPLAY, true>) : go(k2_sub<MCPM_RNG_REPLAY, false>);
}

}  // namespace mcpm
