// Chain-level device helpers shared by the K2 kernels (mcpm_kernels.cu, mcpm_spec.cu): slot forwarding, partner scans,
// CTA reductions, the per-particle energy cache, configuration checksum, in-kernel RDF.
#pragma once
#include "mcpm_device.cuh"

namespace mcpm {

struct Fwd {
    int slot;
    double x, y, z;
};

__device__ __forceinline__ void load_slot(const double* xs, const double* ys, const double* zs, int i, const Fwd& f,
                                          double& x, double& y, double& z) {
    if (i == f.slot) { x = f.x; y = f.y; z = f.z; }
    else { x = xs[i]; y = ys[i]; z = zs[i]; }
}

// Partner loops over the thread's own slots j == tid (mod T).  `self` is the index excluded as "the
// particle itself" (-1: none).  Independent accumulators give the scheduler 4 dependent chains to
// interleave (each pair evaluation is a ~17-deep chain of fp64 ops).
template <bool FAST, bool PBZ>
__device__ __forceinline__ void pair_sum2(const double* xs, const double* ys, const double* zs,
                                          int n, int self, const PairGeom& g, double xa, double ya, double za, double xb, double yb,
                                          double zb, int tid, int T, double& ea, double& eb) {
    double a0 = 0.0, b0 = 0.0, a1 = 0.0, b1 = 0.0;
    int j = tid;
    for (; j + T < n; j += 2 * T) {
        const double x0 = xs[j], y0 = ys[j], z0 = zs[j];
        const double x1 = xs[j + T], y1 = ys[j + T], z1 = zs[j + T];
        pair_add<FAST>(a0, dist2<FAST, PBZ>(g, xa, ya, za, x0, y0, z0), g, j, self);
        pair_add<FAST>(b0, dist2<FAST, PBZ>(g, xb, yb, zb, x0, y0, z0), g, j, self);
        pair_add<FAST>(a1, dist2<FAST, PBZ>(g, xa, ya, za, x1, y1, z1), g, j + T, self);
        pair_add<FAST>(b1, dist2<FAST, PBZ>(g, xb, yb, zb, x1, y1, z1), g, j + T, self);
    }
    if (j < n) {
        const double x0 = xs[j], y0 = ys[j], z0 = zs[j];
        pair_add<FAST>(a0, dist2<FAST, PBZ>(g, xa, ya, za, x0, y0, z0), g, j, self);
        pair_add<FAST>(b0, dist2<FAST, PBZ>(g, xb, yb, zb, x0, y0, z0), g, j, self);
    }
    ea = a0 + a1; eb = b0 + b1;
}
template <bool FAST, bool PBZ>
__device__ __forceinline__ double pair_sum1(const double* xs, const double* ys, const double* zs,
                                            int n, int self, const PairGeom& g, double xa, double ya, double za, int tid, int T) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int j = tid;
    for (; j + 3 * T < n; j += 4 * T) {
        pair_add<FAST>(s0, dist2<FAST, PBZ>(g, xa, ya, za, xs[j], ys[j], zs[j]), g, j, self);
        pair_add<FAST>(s1, dist2<FAST, PBZ>(g, xa, ya, za, xs[j + T], ys[j + T], zs[j + T]), g, j + T, self);
        pair_add<FAST>(s2, dist2<FAST, PBZ>(g, xa, ya, za, xs[j + 2 * T], ys[j + 2 * T], zs[j + 2 * T]), g, j + 2 * T, self);
        pair_add<FAST>(s3, dist2<FAST, PBZ>(g, xa, ya, za, xs[j + 3 * T], ys[j + 3 * T], zs[j + 3 * T]), g, j + 3 * T, self);
    }
    // (round 2: replacing this serial tail by one block of up to four interleaved, index-clamped and self-masked evaluations was measured
    // 17 % SLOWER on C2 — the kernel sits on a register-pressure cliff at 128 registers — and a two-partners-per-trip cache_apply neutral)
    for (; j < n; j += T) pair_add<FAST>(s0, dist2<FAST, PBZ>(g, xa, ya, za, xs[j], ys[j], zs[j]), g, j, self);
    return (s0 + s1) + (s2 + s3);
}

// Latency-oriented scan for a single warp (or a few) per position: U independent partner evaluations in flight per
// thread and NO serial tail — slots past n are evaluated on a clamped index and mask themselves out (self := j), so a
// chain of N <= U*T particles is one straight-line block of U interleaved dependency chains.
template <bool FAST, bool PBZ, int U>
__device__ __forceinline__ double pair_sum_wide(const double* xs, const double* ys, const double* zs, int n, int self, const PairGeom& g,
                                                double xa, double ya, double za, int tid, int T) {
    double acc[U];
#pragma unroll
    for (int k = 0; k < U; k++) acc[k] = 0.0;
    for (int j0 = tid; j0 < n; j0 += U * T) {
#pragma unroll
        for (int k = 0; k < U; k++) {
            const int j = j0 + k * T;
            const int jc = min(j, n - 1);
            pair_acc(acc[k], dist2<FAST, PBZ>(g, xa, ya, za, xs[jc], ys[jc], zs[jc]), g, j, (j < n) ? self : j);
        }
    }
#pragma unroll
    for (int w = 1; w < U; w *= 2)
#pragma unroll
        for (int k = 0; k + w < U; k += 2 * w) acc[k] += acc[k + w];
    return acc[0];
}

// All-reduce of NV per-thread values over the CTA; every thread returns the same bits.
template <int NV, bool MULTI>
__device__ __forceinline__ void cta_allsum(double (&v)[NV], double* partials, int parity, int W, int warp, int lane) {
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = warp_allsum(v[k]);
    if (!MULTI) { __syncwarp(); return; }
    double* p = partials + parity * 64;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) p[warp * 2 + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double s = 0.0;
        for (int w = 0; w < W; w++) s += p[w * 2 + k];
        v[k] = s;
    }
}

// Barrier that also ORs a flag over the CTA (one warp: a vote).
template <bool MULTI>
__device__ __forceinline__ bool cta_sync_or(bool flag) {
    if (MULTI) return __syncthreads_or(flag ? 1 : 0) != 0;
    const bool r = __any_sync(MCPM_FULL_MASK, flag);
    __syncwarp();
    return r;
}
template <bool MULTI>
__device__ __forceinline__ void cta_sync() { if (MULTI) __syncthreads(); else __syncwarp(); }

// Full energy of the chain held in shared memory: per-particle sums as MC::BoxEnergy forms them
// (src/energy.cpp:106-124; every pair is visited twice and the pair total halved).
template <bool FAST, bool PBZ, bool MULTI>
__device__ void cta_box_energy(const double* xs, const double* ys, const double* zs, int n, const ChainParams& P, const PairGeom& g,
                               double* partials, int W, int warp, int lane, int tid, int T, double& pair_half, double& wall) {
    double v[2] = {0.0, 0.0};
    for (int i = tid; i < n; i += T) {
        double xi = xs[i], yi = ys[i], zi = zs[i];
        double s = 0.0;
        for (int j = 0; j < n; j++) pair_add<FAST>(s, dist2<FAST, PBZ>(g, xi, yi, zi, xs[j], ys[j], zs[j]), g, j, i);
        v[0] += s;
        v[1] += wall_single<FAST>(P, xi, yi, zi);
    }
    cta_sync<MULTI>();                                     // (one warp per chain: warp-level only — several chains may share a CTA)
    cta_allsum<2, MULTI>(v, partials, 0, W, warp, lane);   // single-warp CTAs have no partials area
    cta_sync<MULTI>();
    pair_half = 0.5 * (P.eps4 * v[0]);
    wall = v[1];
}

// MC::ComputeRDF same-species branch, src/sample.cpp:21-31: every pair i<j adds 2 to bin int(r/deltaR)+1 (if < NBINS).
// Binning must agree with the reference at bin edges, so the distance is formed exactly as MC::NeighDistance does
// (true division, round(), no FMA contraction, sqrt) rather than with the fast minimum image of the energy path.
//   * pairs are tiled so that every thread has work in every trip: row i (n-1-i partners) is walked together with row n-2-i (i+1
//     partners) — n items per combined row, strided over the CTA;
//   * every warp counts into its OWN 32-bit histogram in shared memory (native shared-memory integer atomics, no contention between
//     warps; a pair's contribution of 2 is an exact count), and the counts are folded into the fp64 histogram once per step.
// Called by the whole CTA between two barriers.  whist: W x 128 unsigned ints, zero on entry and on exit.
//   * exactness without the divisions: with both coordinates inside the box (FAST) |d| <= L, and round(fl(d/L)) is +-1 exactly when
//     |d| >= L/2 (fl is monotone and L/2 is exact), so the image is chosen by comparison and subtracted with the reference's single
//     rounding; pairs with r^2 > rcut^2 can never reach a bin below NBINS and are dropped before the square root; the bin
//     int(fl(r/deltaR)) is taken from a reciprocal multiply unless the product lies within 1e-9 of an integer, where the true
//     division decides.  Bins equal the reference's bit for bit (tests compare whole histograms).
template <bool FAST, bool PBZ>
__device__ void cta_rdf(const double* xs, const double* ys, const double* zs, int n, const ChainParams& P, double* hist, unsigned int* whist, int tid, int T,
                        int warp, int W) {
    const double deltaR = sqrt(P.rc2) / (1. * MCPM_RDF_BINS), inv_delta = 1.0 / deltaR;
    const double Lx = P.L[0], Ly = P.L[1], Lz = P.L[2], hx = 0.5 * Lx, hy = 0.5 * Ly, hz = 0.5 * Lz, rc2 = P.rc2;
    unsigned int* mine = whist + warp * 128;
    const int rows = (n - 1 + 1) / 2;                      // combined rows r = 0 .. ceil((n-1)/2) - 1
    for (int r = 0; r < rows; r++) {
        const int i2 = n - 2 - r;                          // partner row; equals r for the middle row of an odd count of rows
        const int len1 = n - 1 - r, len = len1 + ((i2 != r) ? r + 1 : 0);   // row i2 has n-1-i2 = r+1 partners
        for (int k = tid; k < len; k += T) {
            const int i = (k < len1) ? r : i2;
            const int j = (k < len1) ? r + 1 + k : i2 + 1 + (k - len1);
            double dx = xs[i] - xs[j], dy = ys[i] - ys[j], dz = zs[i] - zs[j];
            if (FAST) {
                dx = (dx >= hx) ? __dsub_rn(dx, Lx) : ((dx <= -hx) ? __dadd_rn(dx, Lx) : dx);
                dy = (dy >= hy) ? __dsub_rn(dy, Ly) : ((dy <= -hy) ? __dadd_rn(dy, Ly) : dy);
                if (PBZ) dz = (dz >= hz) ? __dsub_rn(dz, Lz) : ((dz <= -hz) ? __dadd_rn(dz, Lz) : dz);
            } else {
                if (P.pbc[0]) dx = __dsub_rn(dx, __dmul_rn(round(__ddiv_rn(dx, Lx)), Lx));
                if (P.pbc[1]) dy = __dsub_rn(dy, __dmul_rn(round(__ddiv_rn(dy, Ly)), Ly));
                if (P.pbc[2]) dz = __dsub_rn(dz, __dmul_rn(round(__ddiv_rn(dz, Lz)), Lz));
            }
            const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            if (r2 > rc2) continue;                        // bin >= NBINS: never counted
            const double rr = sqrt(r2);
            const double q = rr * inv_delta;
            int bin = (int)q;
            const double frac = q - (double)bin;
            if (frac < 1e-9 || frac > 1.0 - 1e-9) bin = (int)__ddiv_rn(rr, deltaR);   // at a bin edge the reference's own expression decides
            bin += 1;
            if (bin < MCPM_RDF_BINS) atomicAdd(&mine[bin], 1u);
        }
    }
    __syncthreads();
    for (int k = tid; k < MCPM_RDF_BINS; k += T) {
        unsigned int c = 0;
        for (int w = 0; w < W; w++) { c += whist[w * 128 + k]; whist[w * 128 + k] = 0u; }
        hist[k] += 2.0 * (double)c;
    }
}

// MC::BarWeight, src/moves.cpp:409, with barC = 1 (ResetStats, src/MC.cpp:88)
__device__ __forceinline__ double bar_weight(double energy, double beta) { return 1. / cosh(0.5 * (energy - 1.) * beta); }

// ---- energy cache -------------------------------------------------------------------------------------------------
// ep[j] = sum over partners k of s6(s6-1) for particle j (unscaled by 4*eps), kept in global memory (L2-resident, 8 B per
// particle: it would cost a third more shared memory and with it occupancy).  With it a trial move needs the partner scan
// only for the NEW position (translation, insertion) or not at all (deletion): the old energy is ep[i].  An ACCEPTED move
// pays a second pass that applies the per-pair changes to every partner's entry.  At the reference's adapted ~20 %
// translation acceptance this is ~0.95 N pair evaluations per step instead of 1.5 N.  The cache is rebuilt from scratch
// at the start of every launch (N^2 evaluations), which also re-anchors the running energies.
template <bool FAST, bool PBZ, bool MULTI>
__device__ void cache_fill(const double* xs, const double* ys, const double* zs, double* ep, int n, const ChainParams& P, const PairGeom& g,
                           double* partials, int W, int warp, int lane, int tid, int T, double& pair_half, double& wall) {
    double v[2] = {0.0, 0.0};
    for (int i = tid; i < n; i += T) {
        const double xi = xs[i], yi = ys[i], zi = zs[i];
        double s0 = 0.0, s1 = 0.0;
        int j = 0;
        for (; j + 1 < n; j += 2) {
            pair_acc(s0, dist2<FAST, PBZ>(g, xi, yi, zi, xs[j], ys[j], zs[j]), g, j, i);
            pair_acc(s1, dist2<FAST, PBZ>(g, xi, yi, zi, xs[j + 1], ys[j + 1], zs[j + 1]), g, j + 1, i);
        }
        if (j < n) pair_acc(s0, dist2<FAST, PBZ>(g, xi, yi, zi, xs[j], ys[j], zs[j]), g, j, i);
        ep[i] = s0 + s1;
        v[0] += s0 + s1;
        v[1] += wall_single<FAST>(P, xi, yi, zi);
    }
    cta_sync<MULTI>();                                     // (one warp per chain: warp-level only — several chains may share a CTA)
    cta_allsum<2, MULTI>(v, partials, 0, W, warp, lane);   // single-warp CTAs have no partials area
    cta_sync<MULTI>();
    pair_half = 0.5 * (P.eps4 * v[0]);
    wall = v[1];
}

// Second pass of an accepted move: every thread updates the cache entries of its own slots.
//   has_a: particle at (xa,ya,za) disappears from every partner's sum;  has_b: one at (xb,yb,zb) appears.
// Returns true if this thread saw a term so large (|s6(s6-1)| > 1e6, i.e. a hard overlap of ~5e8 K) that adding and later
// removing it would leave a visible rounding residue in the cache: the caller then rebuilds the cache at once.
template <bool FAST, bool PBZ>
__device__ __forceinline__ bool cache_apply(const double* xs, const double* ys, const double* zs, double* ep, int n, int self, const PairGeom& g,
                                            bool has_a, double xa, double ya, double za, bool has_b, double xb, double yb, double zb,
                                            int tid, int T) {
    bool big = false;
    for (int j = tid; j < n; j += T) {
        const double xj = xs[j], yj = ys[j], zj = zs[j];
        double tb = 0.0, ta = 0.0;
        if (has_b) tb = pair_term(dist2<FAST, PBZ>(g, xb, yb, zb, xj, yj, zj), g, j, self);
        if (has_a) ta = pair_term(dist2<FAST, PBZ>(g, xa, ya, za, xj, yj, zj), g, j, self);
        big = big || !(fabs(tb) <= 1e6) || !(fabs(ta) <= 1e6);
        if (j != self) ep[j] += tb - ta;
    }
    return big;
}

// Order-independent checksum of the live configuration (all threads return the same value; contains barriers).
template <bool MULTI>
__device__ unsigned long long config_checksum(const double* xs, const double* ys, const double* zs, int n, unsigned long long* scratch,
                                              int tid, int T, int lane, int warp, int W) {
    unsigned long long h = 0ull;
    for (int j = tid; j < n; j += T)
        h += (unsigned long long)__double_as_longlong(xs[j]) * 0x9E3779B97F4A7C15ull + (unsigned long long)__double_as_longlong(ys[j]) * 0xC2B2AE3D27D4EB4Full +
             (unsigned long long)__double_as_longlong(zs[j]) * 0x165667B19E3779F9ull + (unsigned long long)(j + 1) * 0xD6E8FEB86659FD93ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(MCPM_FULL_MASK, h, o);
    if (!MULTI) return h + (unsigned long long)n;          // single-warp CTAs have no scratch area
    __syncthreads();
    if (lane == 0) scratch[warp] = h;
    __syncthreads();
    unsigned long long tot = (unsigned long long)n;
    for (int w = 0; w < W; w++) tot += scratch[w];
    __syncthreads();
    return tot;
}

// ---- cell grid shared by the cell-list kernels (K3c in mcpm_kernels.cu, K2c in mcpm_cells.cu) ------------------------
#define K3C_MAXC 16   // cells per axis

struct CellGrid { int nx, ny, nz; };

__device__ __forceinline__ CellGrid cell_grid(const ChainParams& P) {
    const double rc = sqrt(P.rc2);
    CellGrid g;
    g.nx = min(K3C_MAXC, (int)(P.L[0] / rc)); g.ny = min(K3C_MAXC, (int)(P.L[1] / rc)); g.nz = min(K3C_MAXC, (int)(P.L[2] / rc));
    if (!P.pbc[2]) g.nz = max(1, g.nz);
    return g;
}
__device__ __forceinline__ int cell_coord(double x, double L, int n) { return min(n - 1, max(0, (int)(x * ((double)n / L)))); }

// Per-set counters replicated in registers (cheap ints); flushed to the 64-bit totals in shared memory
// by thread 0 at the end of every set.
struct SetCounters {
    int n_disp, acc_disp, n_ins, acc_ins, n_del, acc_del, n_empty;
    unsigned long long evals;   // executed partner-loop iterations
    unsigned long long alg;     // what the reference evaluates: 2n per translation, n per swap
    __device__ __forceinline__ void reset() { n_disp = acc_disp = n_ins = acc_ins = n_del = acc_del = n_empty = 0; evals = 0ull; alg = 0ull; }
};

}  // namespace mcpm
