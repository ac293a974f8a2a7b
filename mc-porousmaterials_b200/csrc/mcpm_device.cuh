// Device-side building blocks: pair term, minimum image, wall potential, uniform sources,
// warp/CTA reductions.  sm_100a, fp64 CUDA-core pipe only (no tensor cores: nothing here is a
// dense contraction).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "mcpm_types.h"

#define MCPM_FULL_MASK 0xffffffffu
#define MCPM_INT_MAX_D 2147483647.0  // the reference's finite out-of-pore penalty, src/energy.cpp:65-67

namespace mcpm {

// ------------------------------------------------------------------------------------------------
// fp64 reciprocal without the slow-path branch of '/': MUFU.RCP64H seed + Newton steps.
// The seed is good to ~2^-20 or better; one cubic step (error^3) followed by one quadratic step
// leaves <= 1 ulp.  r2 == 0 gives inf/NaN — callers mask such pairs out.
__device__ __forceinline__ double fast_rcp(double a) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    double e = fma(-a, x, 1.0);
    double e2 = fma(e, e, e);
    x = fma(x, e2, x);
    e = fma(-a, x, 1.0);
    x = fma(x, e, x);
    return x;
}

// s6*(s6-1) with s6 = (sigma^2/r^2)^3; the factor 4*eps is applied once per sum.
// Restates the summand of MC::LJ126_Pot, src/potentials.cpp:69, in r^2 form.
__device__ __forceinline__ double lj_core(double r2, double sig2) {
    double s2 = sig2 * fast_rcp(r2);
    double s6 = s2 * s2 * s2;
    return fma(s6, s6, -s6);
}

struct PairGeom {
    double Lx, Ly, Lz, invLx, invLy, invLz, rc2, sig2;
    int pbz;  // z periodic (bulk)
};

__device__ __forceinline__ PairGeom make_geom(const ChainParams& P) {
    PairGeom g;
    g.Lx = P.L[0]; g.Ly = P.L[1]; g.Lz = P.L[2];
    g.invLx = P.invL[0]; g.invLy = P.invL[1]; g.invLz = P.invL[2];
    g.rc2 = P.rc2; g.sig2 = P.sig2; g.pbz = P.pbc[2];
    return g;
}

// Minimum image, MC::NeighDistance src/potentials.cpp:22-33.
//  FAST: both positions lie in [0,L] (true for every chain the engine itself advances: insertion
//        draws u*L and translations are wrapped by floor, src/moves.cpp:46-54, src/energy.cpp:126-130),
//        so |d| <= L and the nearest image is min(|d|, L-|d|) — 2 fp64 ops instead of 4.
//  else: d -= rint(d/L)*L for arbitrary coordinates (rint by the 1.5*2^52 trick; differs from the
//        reference's round() only at exact ties, where both images are at the same distance).
template <bool FAST>
__device__ __forceinline__ double image(double d, double L, double invL) {
    if (FAST) {
        double a = fabs(d);
        return fmin(a, L - a);
    } else {
        const double M = 6755399441055744.0;
        double n = (d * invL + M) - M;
        return fma(-n, L, d);
    }
}

template <bool FAST>
__device__ __forceinline__ double dist2(const PairGeom& g, double xi, double yi, double zi, double xj, double yj, double zj) {
    double dx = image<FAST>(xi - xj, g.Lx, g.invLx);
    double dy = image<FAST>(yi - yj, g.Ly, g.invLy);
    double dz = zi - zj;
    if (g.pbz) dz = image<FAST>(dz, g.Lz, g.invLz);
    return fma(dz, dz, fma(dy, dy, dx * dx));
}

// One partner's contribution.  Inclusive cutoff (quirk Q8); r2 == 0 (self, or exactly coincident
// distinct particles, which the reference skips by position equality, quirk Q7) is masked out.
__device__ __forceinline__ double pair_masked(double r2, const PairGeom& g, bool not_self) {
    bool ok = not_self && (r2 <= g.rc2) && (__double_as_longlong(r2) != 0ll);
    double e = lj_core(r2, g.sig2);
    return ok ? e : 0.0;
}

// ------------------------------------------------------------------------------------------------
// Wall term of MC::EnergyOfParticle (src/energy.cpp:53-77) for up to two z values at once.
// Lanes 0-7 take the terms of z_a, lanes 8-15 those of z_b: lane = which*8 + layer*2 + side.
// SlitLJ_Pot, src/potentials.cpp:381-402:  pref * sum_layers sum_sides [0.2 t^10 - 0.5 t^4],
// t = sig/(H/2 + i*delta +- (z - H/2)).  Must be called by all 32 lanes.
__device__ __forceinline__ void wall_two(const ChainParams& P, double z_a, double z_b, double& w_a, double& w_b, int lane) {
    double acc = 0.0;
    if (P.wall_kind == MCPM_WALL_SLIT_LJ && P.geometry == MCPM_GEOM_SLIT) {
        if (lane < 16) {
            double zc = ((lane & 8) ? z_b : z_a) - P.halfH;
            if (lane & 1) zc = -zc;
            for (int layer = (lane >> 1) & 3; layer < P.n_layers; layer += 4) {
                double t = P.sig_sf / (P.halfH + layer * P.delta + zc);
                double t2 = t * t, t4 = t2 * t2;
                acc += 0.2 * (t4 * t4 * t2) - 0.5 * t4;
            }
        }
        acc += __shfl_xor_sync(MCPM_FULL_MASK, acc, 1);
        acc += __shfl_xor_sync(MCPM_FULL_MASK, acc, 2);
        acc += __shfl_xor_sync(MCPM_FULL_MASK, acc, 4);
    }
    w_a = __shfl_sync(MCPM_FULL_MASK, acc, 0) * P.wall_pref;
    w_b = __shfl_sync(MCPM_FULL_MASK, acc, 8) * P.wall_pref;
    if (P.geometry == MCPM_GEOM_SLIT) {  // out-of-pore guard: finite INT_MAX, wall still added (quirk Q6)
        double za = z_a - P.halfH, zb = z_b - P.halfH;
        if (za * za > P.sqr_pore_r) w_a = MCPM_INT_MAX_D + w_a;
        if (zb * zb > P.sqr_pore_r) w_b = MCPM_INT_MAX_D + w_b;
    }
}

// Serial single-thread version (K3 computes one wall term per particle per thread).
__device__ __forceinline__ double wall_one(const ChainParams& P, double z) {
    double boxE = 0.0;
    double zc = z - P.halfH;
    if (P.geometry == MCPM_GEOM_SLIT && zc * zc > P.sqr_pore_r) boxE = MCPM_INT_MAX_D;
    if (P.wall_kind == MCPM_WALL_SLIT_LJ && P.geometry == MCPM_GEOM_SLIT) {
        double usf = 0.0;
        for (int i = 0; i < P.n_layers; i++) {
            double ta = P.sig_sf / (P.halfH + i * P.delta + zc);
            double tb = P.sig_sf / (P.halfH + i * P.delta - zc);
            double a2 = ta * ta, a4 = a2 * a2, b2 = tb * tb, b4 = b2 * b2;
            usf += 0.2 * (a4 * a4 * a2 + b4 * b4 * b2) - 0.5 * (a4 + b4);
        }
        boxE += usf * P.wall_pref;
    }
    return boxE;
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: stream = f(seed, chain, step, slot).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t& o0, uint32_t& o1, uint32_t& o2, uint32_t& o3) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

// 53-bit uniform in [0,1) scaled by 0.9999 — the reference's MC::Random() range (quirk Q1, src/MC.h:33).
__device__ __forceinline__ double u53(uint32_t lo, uint32_t hi) {
    unsigned long long w = (((unsigned long long)hi << 32) | lo) >> 11;
    return __dmul_rn(0.9999, __ull2double_rn(w) * 1.1102230246251565e-16);
}

// Call index = 4*step + slot/2; ctr = (call_lo, call_hi, chain, 'MCPM'), key = (seed_lo, seed_hi).
__device__ __forceinline__ void philox_call(unsigned long long seed, uint32_t chain, unsigned long long call, double& ua, double& ub) {
    uint32_t o0, o1, o2, o3;
    philox4x32_10((uint32_t)call, (uint32_t)(call >> 32), chain, 0x4D43504Du, (uint32_t)seed, (uint32_t)(seed >> 32), o0, o1, o2, o3);
    ua = u53(o0, o1);
    ub = u53(o2, o3);
}

// Warp-batched Philox source: every 8 steps each lane evaluates ONE of the 32 calls of the batch
// (4 calls x 8 steps); draws are fetched by shuffle.  All lanes of all warps hold the same stream,
// so every thread takes the same Metropolis decision without any broadcast.
struct PhiloxSource {
    unsigned long long seed;
    uint32_t chain;
    double ua, ub;
    unsigned long long batch;  // step index of the batch start (multiple of 8), ~0 = none
    int base, slot;
    __device__ __forceinline__ void init(unsigned long long s, uint32_t c) { seed = s; chain = c; batch = ~0ull; ua = ub = 0.0; base = slot = 0; }
    __device__ __forceinline__ void begin_step(unsigned long long step, int lane) {
        unsigned long long b = step & ~7ull;
        if (b != batch) {
            batch = b;
            philox_call(seed, chain, 4ull * b + (unsigned long long)lane, ua, ub);
        }
        base = (int)(step & 7ull) * 4;
        slot = 0;
    }
    __device__ __forceinline__ double draw() {
        double v = (slot & 1) ? ub : ua;
        v = __shfl_sync(MCPM_FULL_MASK, v, base + (slot >> 1));
        slot++;
        return v;
    }
    __device__ __forceinline__ unsigned long long cursor() const { return 0ull; }
};

// Replay source: consumes a caller-supplied uniform stream exactly as the reference consumes
// MC::Random() — sequentially, 1 + 7/6/4/2 draws per step.
struct ReplaySource {
    const double* u;
    unsigned long long pos, n;
    __device__ __forceinline__ void init(const double* p, unsigned long long start, unsigned long long len) { u = p; pos = start; n = len; }
    __device__ __forceinline__ void begin_step(unsigned long long, int) {}
    __device__ __forceinline__ double draw() {
        double v = (pos < n) ? __ldg(u + pos) : 0.0;
        pos++;
        return v;
    }
    __device__ __forceinline__ unsigned long long cursor() const { return pos; }
};

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_allsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MCPM_FULL_MASK, v, o);
    return v;
}

// MC::PBC, src/energy.cpp:126-130: x -= floor(x/L)*L, without FMA contraction so that the wrapped
// coordinate is bit-identical to the host reference.
__device__ __forceinline__ double wrap_floor(double x, double L) {
    return __dsub_rn(x, __dmul_rn(floor(__ddiv_rn(x, L)), L));
}

}  // namespace mcpm
