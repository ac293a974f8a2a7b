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
// fp64 reciprocal without the slow-path branch of '/': MUFU.RCP64H seed (measured 2^-19.9 relative on
// B200, scripts/rcp_probe.cu) + ONE cubic Newton step x(1 + e + e^2), e = 1 - a*x: error e^3 ~ 2^-60
// before rounding; measured worst case 2.2e-16 relative (<= 1 ulp) over r^2 in [1,1024].
// r2 == 0 gives inf/NaN — callers mask such pairs out.
__device__ __forceinline__ double fast_rcp(double a) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    double e = fma(-a, x, 1.0);
    return fma(x, fma(e, e, e), x);
}

// s6*(s6-1) with s6 = (sigma^2/r^2)^3; the factor 4*eps is applied once per sum.
// Restates the summand of MC::LJ126_Pot, src/potentials.cpp:69, in r^2 form.
__device__ __forceinline__ double lj_core(double r2, double sig2) {
    double s2 = sig2 * fast_rcp(r2);
    double s6 = s2 * s2 * s2;
    return fma(s6, s6, -s6);
}

struct PairGeom {
    double Lx, Ly, Lz;     // FAST: HALF widths; general: full widths
    double iLx, iLy, iLz;  // general path only: 1/L
    double rc2, sig2;
    // general path only (dead in the FAST instantiations): which axes are periodic (src/read.cpp:146-162) and the pair potential
    int px, py, pz, hs;
    double sig;
};

template <bool FAST>
__device__ __forceinline__ PairGeom make_geom(const ChainParams& P) {
    PairGeom g;
    const double f = FAST ? 0.5 : 1.0;
    g.Lx = f * P.L[0]; g.Ly = f * P.L[1]; g.Lz = f * P.L[2];
    g.iLx = P.invL[0]; g.iLy = P.invL[1]; g.iLz = P.invL[2];
    g.rc2 = P.rc2; g.sig2 = P.sig2;
    g.px = P.pbc[0]; g.py = P.pbc[1]; g.pz = P.pbc[2]; g.hs = (P.pair_kind == MCPM_PAIR_HS); g.sig = P.sig_ff;
    return g;
}
// A cubic box of a new width (NPT volume moves, general path)
__device__ __forceinline__ void geom_set_cube(PairGeom& g, double L) { g.Lx = g.Ly = g.Lz = L; g.iLx = g.iLy = g.iLz = 1.0 / L; }

// Minimum image, MC::NeighDistance src/potentials.cpp:22-33.
//  FAST: both positions lie in [0,L] (true for every chain the engine itself advances: insertion
//        draws u*L and translations are wrapped by floor, src/moves.cpp:46-54, src/energy.cpp:126-130),
//        so a = |d| <= L and the nearest image is min(a, L-a) = L/2 - |a - L/2|: two fp64 adds with
//        |.| operand modifiers, no compare/select.  (a - L/2 is exact for a >= L/4 by Sterbenz; below
//        that the result carries an absolute error <= ulp(L/2), ~1e-16 relative at contact distance.)
//  else: d -= rint(d/L)*L for arbitrary coordinates (rint by the 1.5*2^52 trick; differs from the
//        reference's round() only at exact ties, where both images are at the same distance).
template <bool FAST>
__device__ __forceinline__ double image(double d, double L, double invL) {
    if (FAST) {
        return L - fabs(fabs(d) - L);   // L is the HALF width here
    } else {
        const double M = 6755399441055744.0;
        double n = (d * invL + M) - M;
        return fma(-n, L, d);
    }
}

// PBZ: z is periodic (bulk geometry); in a slit only x and y are (src/read.cpp:146-162).
template <bool FAST, bool PBZ>
__device__ __forceinline__ double dist2(const PairGeom& g, double xi, double yi, double zi, double xj, double yj, double zj) {
    if (FAST) {
        double dx = image<FAST>(xi - xj, g.Lx, g.iLx);
        double dy = image<FAST>(yi - yj, g.Ly, g.iLy);
        double dz = zi - zj;
        if (PBZ) dz = image<FAST>(dz, g.Lz, g.iLz);
        return fma(dz, dz, fma(dy, dy, dx * dx));
    } else {   // general path: the periodic axes follow the geometry (bulk xyz, slit xy, cylinder x, sphere none)
        double dx = xi - xj, dy = yi - yj, dz = zi - zj;
        if (g.px) dx = image<false>(dx, g.Lx, g.iLx);
        if (g.py) dy = image<false>(dy, g.Ly, g.iLy);
        if (g.pz) dz = image<false>(dz, g.Lz, g.iLz);
        return fma(dz, dz, fma(dy, dy, dx * dx));
    }
}

// One partner's masked term (0 when out of range / self / coincident), for the energy-cache updates.
__device__ __forceinline__ double pair_term(double r2, const PairGeom& g, int j, int self);

// Adds partner j's contribution to acc, fully predicated (no branch, no select):
//   r2 <= rc2   inclusive cutoff (quirk Q8)
//   hi(r2) != 0 exactly coincident distinct particles, which the reference skips by position equality
//               (quirk Q7); only the high word is inspected, i.e. r2 < 2^-1022 counts as coincident
//   j != self   the particle itself
__device__ __forceinline__ void pair_acc(double& acc, double r2, const PairGeom& g, int j, int self) {
    const double e = lj_core(r2, g.sig2);
    asm("{\n\t.reg .pred p;\n\t"
        "setp.le.f64 p, %1, %2;\n\t"
        "setp.ne.and.s32 p, %3, 0, p;\n\t"
        "setp.ne.and.s32 p, %4, %5, p;\n\t"
        "@p add.f64 %0, %0, %6;\n\t}"
        : "+d"(acc)
        : "d"(r2), "d"(g.rc2), "r"(__double2hiint(r2)), "r"(j), "r"(self), "d"(e));
}

__device__ __forceinline__ double pair_term(double r2, const PairGeom& g, int j, int self) {
    double t = 0.0;
    pair_acc(t, r2, g, j, self);
    return t;
}

// Partner term of the GENERAL kernels: Lennard-Jones as above, or MC::HardSphere_Pot (src/potentials.cpp:40-54): the finite INT_MAX
// for every partner within sigma, and NO self-exclusion — a stored particle meets its own record at distance 0 and counts itself,
// wherever the trial position is (the record moves with it, src/moves.cpp:163): quirk Q17.  Sums of hard-sphere terms are energies
// already (the host sets the 4*eps prefactor to 1).
template <bool FAST>
__device__ __forceinline__ void pair_add(double& acc, double r2, const PairGeom& g, int j, int self) {
    if (FAST || !g.hs) pair_acc(acc, r2, g, j, self);
    else if (j == self || sqrt(r2) <= g.sig) acc += MCPM_INT_MAX_D;
}

// ------------------------------------------------------------------------------------------------
// Wall term of MC::EnergyOfParticle (src/energy.cpp:53-77) for up to two z values at once.
// Lanes 0-7 take the terms of z_a, lanes 8-15 those of z_b: lane = which*8 + layer*2 + side.
// SlitLJ_Pot, src/potentials.cpp:381-402:  pref * sum_layers sum_sides [0.2 t^10 - 0.5 t^4],
// t = sig/(H/2 + i*delta +- (z - H/2)).  Must be called by all 32 lanes.
__device__ __forceinline__ void wall_two(const ChainParams& P, double z_a, double z_b, double& w_a, double& w_b, int lane) {
    double acc = 0.0;
    if (P.wall_kind == MCPM_WALL_SLIT_LJ && P.geometry == MCPM_GEOM_SLIT) {
        double zc = ((lane & 8) ? z_b : z_a) - P.halfH;
        if (lane & 1) zc = -zc;
        const int layer0 = (lane >> 1) & 3;
        if (P.n_layers <= 4) {   // the usual case (3 graphite layers): straight-line code, one term per lane, no loop
            const double t = P.sig_sf * fast_rcp(P.halfH + layer0 * P.delta + zc);
            const double t2 = t * t, t4 = t2 * t2;
            const double term = 0.2 * (t4 * t4 * t2) - 0.5 * t4;
            acc = (lane < 16 && layer0 < P.n_layers) ? term : 0.0;
        } else if (lane < 16) {
            for (int layer = layer0; layer < P.n_layers; layer += 4) {
                double t = P.sig_sf * fast_rcp(P.halfH + layer * P.delta + zc);
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

// wall_two fused with the warp all-reduce of one scan partial `v` (one-warp chains): the two shuffle trees are independent, so
// their stages are issued together and the wall's reduction hides behind the longer one of v.  (butterfly in ascending order: every
// lane still ends with the same bits.)
__device__ __forceinline__ void wall_two_and_sum(const ChainParams& P, double z_a, double z_b, double& w_a, double& w_b, double& v, int lane) {
    double acc = 0.0;
    const bool slit_lj = (P.wall_kind == MCPM_WALL_SLIT_LJ && P.geometry == MCPM_GEOM_SLIT);
    if (slit_lj) {
        double zc = ((lane & 8) ? z_b : z_a) - P.halfH;
        if (lane & 1) zc = -zc;
        const int layer0 = (lane >> 1) & 3;
        if (P.n_layers <= 4) {
            const double t = P.sig_sf * fast_rcp(P.halfH + layer0 * P.delta + zc);
            const double t2 = t * t, t4 = t2 * t2;
            const double term = 0.2 * (t4 * t4 * t2) - 0.5 * t4;
            acc = (lane < 16 && layer0 < P.n_layers) ? term : 0.0;
        } else if (lane < 16) {
            for (int layer = layer0; layer < P.n_layers; layer += 4) {
                double t = P.sig_sf * fast_rcp(P.halfH + layer * P.delta + zc);
                double t2 = t * t, t4 = t2 * t2;
                acc += 0.2 * (t4 * t4 * t2) - 0.5 * t4;
            }
        }
    }
#pragma unroll
    for (int o = 1; o <= 4; o <<= 1) {
        const double pv = __shfl_xor_sync(MCPM_FULL_MASK, v, o), pa = __shfl_xor_sync(MCPM_FULL_MASK, acc, o);
        v += pv; acc += pa;
    }
    const double a0 = __shfl_sync(MCPM_FULL_MASK, acc, 0), a8 = __shfl_sync(MCPM_FULL_MASK, acc, 8);
    v += __shfl_xor_sync(MCPM_FULL_MASK, v, 8);
    v += __shfl_xor_sync(MCPM_FULL_MASK, v, 16);
    w_a = a0 * P.wall_pref;
    w_b = a8 * P.wall_pref;
    if (P.geometry == MCPM_GEOM_SLIT) {  // out-of-pore guard: finite INT_MAX, wall still added (quirk Q6)
        double za = z_a - P.halfH, zb = z_b - P.halfH;
        if (za * za > P.sqr_pore_r) w_a = MCPM_INT_MAX_D + w_a;
        if (zb * zb > P.sqr_pore_r) w_b = MCPM_INT_MAX_D + w_b;
    }
}

// Tools::Pow, src/tools.h:53-58
__device__ __forceinline__ double ipow_d(double v, int p) { double r = 1.0; for (int i = 1; i <= p; i++) r *= v; return r; }

// MC::SphericalLJ_Pot, src/potentials.cpp:474-504 (Baksh & Yang 1991), restated term for term.
__device__ __forceinline__ double sphere_wall(const ChainParams& P, double xc, double yc, double zc) {
    const double sig = P.sig_sf;
    const double dens = P.rho_s * sig * sig, R = P.halfH / sig;
    const double r = sqrt(xc * xc + yc * yc + zc * zc) / sig;
    const double x = R - r, x2 = x - 2.0 * R;
    double u1 = 0.0, u2 = 0.0, sgn = 1.0;
    for (int i = 0; i < 10; i++) {
        const double ri = ipow_d(R, i);
        u1 += 1. / ri / ipow_d(x, 10 - i);
        u1 += sgn / ri / ipow_d(x2, 10 - i);
        if (i < 4) {
            u2 += 1. / ri / ipow_d(x, 4 - i);
            u2 += sgn / ri / ipow_d(x2, 4 - i);
        }
        sgn = -sgn;
    }
    return ((2. / 5.) * u1 - u2) * (2. * 3.14159265358979323846 * P.eps_sf * dens);
}

// ---- cylindrical walls.  The reference evaluates 2F1(a,b;1;z) with the vendored Numerical-Recipes hypgeo (series / ODE integration,
// src/NumRecipes/hypgeo.h:46; for z > 1, outside the pore, integrated through the upper half plane, real part returned).  Here in closed
// form: F(1/2,1/2) = 2K/pi, F(-1/2,1/2) = 2E/pi, F(-1/2,-1/2) = (2/pi)(2E - (1-z)K) with the complete elliptic integrals by the
// arithmetic-geometric mean (z > 1: real parts of their continuation from above), then Gauss' contiguous relation (DLMF 15.5.11) lowers
// one parameter at a time down to -9/2.  Agreement with the compiled reference inside the pore 2e-12 (tests/test_oracle_f4.py).
__device__ __forceinline__ void ellip_ke01(double m, double& K, double& E) {   // 0 <= m < 1
    double a = 1.0, b = sqrt(1.0 - m), c = sqrt(m), sum = 0.5 * c * c, pw = 1.0;
    for (int n = 0; n < 12; n++) {
        const double an = 0.5 * (a + b), bn = sqrt(a * b);
        c = 0.5 * (a - b); a = an; b = bn;
        sum += pw * c * c; pw *= 2.0;
    }
    K = 3.14159265358979323846 / (2.0 * a);
    E = K * (1.0 - sum);
}
__device__ __forceinline__ void ellip_ke(double m, double& K, double& E) {
    if (m < 1.0) { ellip_ke01(m, K, E); return; }
    double k, e;
    const double s = sqrt(m);
    ellip_ke01(1.0 / m, k, e);
    K = k / s;
    E = s * e - (m - 1.0) / s * k;
}
// 2F1(-9/2,-9/2;1;z), 2F1(-3/2,-3/2;1;z), 2F1(-3/2,-1/2;1;z) from one pass over the table T[i][j] = 2F1(1/2-i, 1/2-j; 1; z), kept column by column
__device__ __forceinline__ void hyp_wall(double z, double& f_m45_m45, double& f_m15_m15, double& f_m15_m05) {
    double K, E;
    ellip_ke(z, K, E);
    const double p = 2.0 / 3.14159265358979323846;
    double row0[6], row1[6];   // T[i][0], T[i][1]: the first two entries of every later column by symmetry
    double col[6];
    for (int j = 0; j < 6; j++) {
        const double bb = 0.5 - j;
        if (j == 0) { col[0] = p * K; col[1] = p * E; }
        else if (j == 1) { col[0] = p * E; col[1] = p * (2 * E - (1 - z) * K); }
        else { col[0] = row0[j]; col[1] = row1[j]; }
        for (int i = 2; i < 6; i++) {
            const double aa = 0.5 - (i - 1);
            col[i] = -((2 * aa - 1 + (bb - aa) * z) * col[i - 1] + aa * (z - 1) * col[i - 2]) / (1 - aa);
        }
        if (j == 0) { for (int i = 0; i < 6; i++) row0[i] = col[i]; }
        if (j == 1) { for (int i = 0; i < 6; i++) row1[i] = col[i]; f_m15_m05 = col[2]; }
        if (j == 2) f_m15_m15 = col[2];
        if (j == 5) f_m45_m45 = col[5];
    }
}
// MC::CylindricalLJ10_4, src/potentials.cpp:412-444
__device__ __forceinline__ double cyl_lj104(const ChainParams& P, double yc, double zc) {
    const double sig = P.sig_sf, dens = P.rho_s * sig * sig, R = P.halfH / sig, delta = P.delta / sig;
    const double rho = sqrt(yc * yc + zc * zc) / sig, x = R - rho;
    double u1 = 0.0, u2 = 0.0;
    for (int i = 0; i < P.n_layers; i++) {
        const double Ri = R + i * delta, arg = ipow_d(1 - x / Ri, 2);
        double f45, f15, f05;
        hyp_wall(arg, f45, f15, f05);
        u1 += 63. / 32. * 1. / (ipow_d(x, 10) * ipow_d(2.0 - x / Ri, 10)) * f45;
        u2 += 3. / (ipow_d(x, 4) * ipow_d(2.0 - x / Ri, 4)) * f15;
    }
    return (u1 - u2) * 3.14159265358979323846 * 3.14159265358979323846 * dens * P.eps_sf;
}
// MC::CylindricalSteele10_4_3, src/potentials.cpp:445-469 (alpha = 0.61; the Gamma-function prefactors are constants)
__device__ __forceinline__ double cyl_steele(const ChainParams& P, double yc, double zc) {
    const double sig = P.sig_sf, R = P.halfH, Ra = R + 0.61 * P.delta;
    const double rho = sqrt(yc * yc + zc * zc);
    const double q = ipow_d(rho / R, 2), qa = ipow_d(rho / Ra, 2);
    double f45, f15, f05, g45, g15, g05;
    hyp_wall(q, f45, f15, f05);
    hyp_wall(qa, g45, g15, g05);
    const double psi6 = 3.0925052683774523 * ipow_d(sig / R, 10) / ipow_d(1 - q, 10) * f45;
    const double psi3 = 4.712388980384691 * ipow_d(sig / R, 4) / ipow_d(1 - q, 4) * f15;
    const double phi3 = 1.5707963267948966 * ipow_d(sig / Ra, 3) / ipow_d(1 - qa, 3) * g05;
    return 2 * 3.14159265358979323846 * P.rho_s * P.delta * sig * sig * P.eps_sf * (psi6 - psi3 - sig / P.delta * phi3);
}

// Serial single-thread version (K3 computes one wall term per particle per thread).
__device__ __forceinline__ double wall_one(const ChainParams& P, double z) {
    double boxE = 0.0;
    double zc = z - P.halfH;
    if (P.geometry == MCPM_GEOM_SLIT && zc * zc > P.sqr_pore_r) boxE = MCPM_INT_MAX_D;
    if (P.wall_kind == MCPM_WALL_SLIT_LJ && P.geometry == MCPM_GEOM_SLIT) {
        double usf = 0.0;
        for (int i = 0; i < P.n_layers; i++) {
            double ta = P.sig_sf * fast_rcp(P.halfH + i * P.delta + zc);
            double tb = P.sig_sf * fast_rcp(P.halfH + i * P.delta - zc);
            double a2 = ta * ta, a4 = a2 * a2, b2 = tb * tb, b4 = b2 * b2;
            usf += 0.2 * (a4 * a4 * a2 + b4 * b4 * b2) - 0.5 * (a4 + b4);
        }
        boxE += usf * P.wall_pref;
    }
    return boxE;
}

// boxE of MC::EnergyOfParticle (src/energy.cpp:53-77) for ANY supported geometry / wall: out-of-pore guard by geometry (sphere r^2,
// cylinder y^2+z^2, slit z^2 against (H/2)^2; finite INT_MAX, the wall term is still added: quirk Q6), then the wall dispatch.
__device__ __forceinline__ double wall_any(const ChainParams& P, double x, double y, double z) {
    const double xc = x - 0.5 * P.L[0], yc = y - 0.5 * P.L[1], zc = z - P.halfH;
    double d2 = 0.0;
    if (P.geometry == MCPM_GEOM_SPHERE) d2 = xc * xc + yc * yc + zc * zc;
    else if (P.geometry == MCPM_GEOM_CYLINDER) d2 = yc * yc + zc * zc;
    else if (P.geometry == MCPM_GEOM_SLIT) d2 = zc * zc;
    double boxE = (d2 > P.sqr_pore_r || d2 < 0.0) ? MCPM_INT_MAX_D : 0.0;
    if (P.wall_kind == MCPM_WALL_SPHERE_LJ && P.geometry == MCPM_GEOM_SPHERE) boxE += sphere_wall(P, xc, yc, zc);
    else if (P.wall_kind == MCPM_WALL_CYL_LJ104 && P.geometry == MCPM_GEOM_CYLINDER) boxE += cyl_lj104(P, yc, zc);
    else if (P.wall_kind == MCPM_WALL_CYL_STEELE && P.geometry == MCPM_GEOM_CYLINDER) boxE += cyl_steele(P, yc, zc);
    else if (P.wall_kind == MCPM_WALL_SLIT_LJ && P.geometry == MCPM_GEOM_SLIT) {
        double usf = 0.0;
        for (int i = 0; i < P.n_layers; i++) {
            double ta = P.sig_sf * fast_rcp(P.halfH + i * P.delta + zc);
            double tb = P.sig_sf * fast_rcp(P.halfH + i * P.delta - zc);
            double a2 = ta * ta, a4 = a2 * a2, b2 = tb * tb, b4 = b2 * b2;
            usf += 0.2 * (a4 * a4 * a2 + b4 * b4 * b2) - 0.5 * (a4 + b4);
        }
        boxE += usf * P.wall_pref;
    }
    return boxE;
}
// Wall terms of two positions: the hot kernels (FAST) know only slits and use the lane-parallel wall_two on z; the general kernels
// take the lane-parallel path too unless the system is exotic.  Must be called by all 32 lanes.
template <bool FAST>
__device__ __forceinline__ void wall_pair(const ChainParams& P, double xa, double ya, double za, double xb, double yb, double zb, double& wa, double& wb,
                                          int lane) {
    if (FAST || !P.exotic) wall_two(P, za, zb, wa, wb, lane);
    else { wa = wall_any(P, xa, ya, za); wb = wall_any(P, xb, yb, zb); }
}
template <bool FAST>
__device__ __forceinline__ double wall_single(const ChainParams& P, double x, double y, double z) {
    if (FAST || !P.exotic) return wall_one(P, z);
    return wall_any(P, x, y, z);
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
// tag = 'MCPM' for the move's draws, 'MCPN' for the sampling (Widom) draws of the same step.
__device__ __forceinline__ void philox_call(unsigned long long seed, uint32_t chain, unsigned long long call, double& ua, double& ub,
                                            uint32_t tag = 0x4D43504Du) {
    uint32_t o0, o1, o2, o3;
    philox4x32_10((uint32_t)call, (uint32_t)(call >> 32), chain, tag, (uint32_t)seed, (uint32_t)(seed >> 32), o0, o1, o2, o3);
    ua = u53(o0, o1);
    ub = u53(o2, o3);
}

// Warp-batched Philox source: every 8 steps each lane evaluates ONE of the 32 calls of the batch
// (4 calls x 8 steps); a step's k-th draw is fetched by shuffle from lane 4*(step%8) + k/2.  All lanes
// of all warps hold the same stream, so every thread takes the same Metropolis decision without any
// broadcast.  Draws are addressed by their position k in the step (not consumed sequentially), so a
// branch can fetch all the uniforms it needs up front, off the critical path.
struct PhiloxSource {
    double ua, ub;   // this lane's call of the current batch
    int base;
    __device__ __forceinline__ void init(unsigned long long seed, uint32_t chain, unsigned long long step, int lane) {
        refill(seed, chain, step & ~7ull, lane);   // a launch may start in the middle of a batch
        base = 0;
    }
    __device__ __forceinline__ void refill(unsigned long long seed, uint32_t chain, unsigned long long batch_step, int lane) {
        philox_call(seed, chain, 4ull * batch_step + (unsigned long long)lane, ua, ub);
    }
    __device__ __forceinline__ void begin_step(unsigned long long seed, uint32_t chain, unsigned long long step, int lane) {
        if ((step & 7ull) == 0ull) refill(seed, chain, step, lane);
        base = (int)(step & 7ull) * 4;
    }
    template <int K>
    __device__ __forceinline__ double at() const {
        return __shfl_sync(MCPM_FULL_MASK, (K & 1) ? ub : ua, base + (K >> 1));
    }
    __device__ __forceinline__ void end_step(int) {}
    __device__ __forceinline__ unsigned long long cursor() const { return 0ull; }
    // Sampling draws of a step (positions 8.. in the reference's sequential stream): lanes 0..3 evaluate the four
    // calls of the step on the second stream; only production steps of NVT runs pay for this.
    double ea, eb;
    __device__ __forceinline__ void load_extra(unsigned long long seed, uint32_t chain, unsigned long long step, int lane) {
        philox_call(seed, chain, 4ull * step + (unsigned long long)(lane & 3), ea, eb, 0x4D43504Eu);
    }
    template <int K>
    __device__ __forceinline__ double extra() const { return __shfl_sync(MCPM_FULL_MASK, (K & 1) ? eb : ea, K >> 1); }
};

// The same stream for ONE-WARP chains, staged through shared memory: every 8 steps each lane stores its call's two uniforms
// (slot 2*lane = position 8*(step%8) + k of the batch) and a step's k-th draw is ONE broadcast load instead of two
// shuffles, a select and a lane computation per draw (~25 fewer instructions per step).  buf: 64 doubles of this warp.
struct PhiloxSmemSource {
    double* buf;
    int base;
    __device__ __forceinline__ void init(unsigned long long seed, uint32_t chain, unsigned long long step, int lane, double* b) {
        buf = b;
        refill(seed, chain, step & ~7ull, lane);   // a launch may start in the middle of a batch
        base = 0;
    }
    __device__ __forceinline__ void refill(unsigned long long seed, uint32_t chain, unsigned long long batch_step, int lane) {
        double ua, ub;
        philox_call(seed, chain, 4ull * batch_step + (unsigned long long)lane, ua, ub);
        __syncwarp();                              // every lane is done with the previous batch
        buf[2 * lane] = ua; buf[2 * lane + 1] = ub;   // (two stores: the buffer is only 8-byte aligned for odd capacities)
        __syncwarp();
    }
    __device__ __forceinline__ void begin_step(unsigned long long seed, uint32_t chain, unsigned long long step, int lane) {
        if ((step & 7ull) == 0ull) refill(seed, chain, step, lane);
        base = (int)(step & 7ull) * 8;
    }
    template <int K>
    __device__ __forceinline__ double at() const { return buf[base + K]; }
    __device__ __forceinline__ void end_step(int) {}
    __device__ __forceinline__ unsigned long long cursor() const { return 0ull; }
    double ea, eb;
    __device__ __forceinline__ void load_extra(unsigned long long seed, uint32_t chain, unsigned long long step, int lane) {
        philox_call(seed, chain, 4ull * step + (unsigned long long)(lane & 3), ea, eb, 0x4D43504Eu);
    }
    template <int K>
    __device__ __forceinline__ double extra() const { return __shfl_sync(MCPM_FULL_MASK, (K & 1) ? eb : ea, K >> 1); }
};

// Replay source: consumes a caller-supplied uniform stream exactly as the reference consumes
// MC::Random() — sequentially, 1 + 7/6/4/2 draws per step.
struct ReplaySource {
    const double* u;
    unsigned long long pos, n;
    __device__ __forceinline__ void init(const double* p, unsigned long long start, unsigned long long len) { u = p; pos = start; n = len; }
    __device__ __forceinline__ void begin_step(unsigned long long, uint32_t, unsigned long long, int) {}
    template <int K>
    __device__ __forceinline__ double at() const { return (pos + K < n) ? __ldg(u + pos + K) : 0.0; }
    __device__ __forceinline__ void end_step(int count) { pos += (unsigned long long)count; }
    __device__ __forceinline__ unsigned long long cursor() const { return pos; }
    __device__ __forceinline__ void load_extra(unsigned long long, uint32_t, unsigned long long, int) {}
    template <int K>
    __device__ __forceinline__ double extra() const { return at<K>(); }   // the stream simply continues
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
    // For -L <= x < 2L the quotient floor(x/L) is -1, 0 or 1 and can be read off by comparison
    // (x/L rounds to < 1 for every double x < L, and to >= 1 for x >= L), so the IEEE division is only
    // needed when a (never-capped, src/moves.cpp:77) step size has grown beyond the box.
    double q;
    if (x >= -L && x < L + L) q = (x < 0.0) ? -1.0 : ((x >= L) ? 1.0 : 0.0);
    else q = floor(__ddiv_rn(x, L));
    return __dsub_rn(x, __dmul_rn(q, L));
}

// Common case of wrap_floor as pure selects; `odd` is raised when x lies outside [-L, 2L) and the caller must redo
// the coordinate with wrap_floor (one rare, warp-uniform branch for all three coordinates).
__device__ __forceinline__ double wrap_select(double x, double L, bool& odd) {
    odd = odd || !(x >= -L && x < L + L);
    // x - 0*L == x exactly, so the in-box case (by far the most frequent) is a plain select: the add/subtract of L is evaluated
    // beside the two comparisons instead of behind a compare -> select q -> multiply -> subtract chain
    const double up = __dadd_rn(x, L), down = __dsub_rn(x, L);   // == x - (-1)*L and x - 1*L (the products are exact)
    return (x < 0.0) ? up : ((x >= L) ? down : x);
}

// Metropolis tests without a logarithm and (mostly) without an exponential.  The reference tests `u < f` with
// f = exp(-dE/T), zz V exp(-E/T)/(N+1) or N exp(E/T)/(zz V).  Uniforms lie in (0, 0.9999) (quirk Q1; a 53-bit zero
// has probability 1e-16), so ln f >= 0 always accepts and ln f < -37 (< ln of the smallest non-zero uniform) always
// rejects; only in between is f evaluated, with the reference's own expression.  All inputs are identical in every
// thread, so the branches are warp-uniform.
// In between, a single-precision screen decides all but ~1e-4 of the cases: f32 = exp(t) evaluated with ex2.approx
// carries a relative error < 1e-5 for |t| <= 58 (CUDA documents 2 + floor(1.16|t|) ulp for __expf, plus the rounding of
// t to float and one multiply/divide), so u outside f32 (1 +- 2e-4) has the same outcome as the exact test; only inside
// that band is the reference's fp64 expression evaluated.  The decision is therefore always the reference's.
// The comparison itself is done in single precision too ((float)u is off the critical path: u is known at the start of the
// step): rounding u adds 6e-8 to the relative error budget, still 10x inside the band.
#define MCPM_SCREEN 2e-4f
__device__ __forceinline__ int screen(double u, float f32) {   // +1 accept, -1 reject, 0 undecided
    const float uf = (float)u;
    if (uf < f32 * (1.0f - MCPM_SCREEN)) return 1;
    if (uf > f32 * (1.0f + MCPM_SCREEN)) return -1;
    return 0;
}
__device__ __forceinline__ bool metropolis_translation(double u, double arg /* dE/T */) {
    if (!(arg > 0.0)) return arg == arg;               // downhill: always (NaN: never, as u < NaN is false)
    if (arg > 37.0) return u == 0.0 && exp(-arg) > 0.0;
    const int sc = screen(u, __expf(-(float)arg));
    if (sc != 0) return sc > 0;
    return u < exp(-arg);
}
__device__ __forceinline__ bool metropolis_insertion(double u, double x /* -E/T */, double ln_zzV, double zzV, int n) {
    const double t = ln_zzV + x;                        // ln f = t - ln(n+1), 0 <= ln(n+1) < 21
    if (t >= 21.0) return true;
    if (!(t >= -37.0)) return (t == t) && u == 0.0 && zzV * exp(x) / (double)(n + 1) > 0.0;
    const int sc = screen(u, __fdividef(__expf((float)t), (float)(n + 1)));
    if (sc != 0) return sc > 0;
    return u < zzV * exp(x) / (double)(n + 1);          // src/moves.cpp:306
}
__device__ __forceinline__ bool metropolis_deletion(double u, double y /* E/T */, double ln_zzV, double zzV, int n) {
    const double t = y - ln_zzV;                        // ln f = t + ln(n), 0 <= ln(n) < 21
    if (t >= 0.0) return true;
    if (!(t >= -58.0)) return (t == t) && u == 0.0 && (double)n * exp(y) / zzV > 0.0;
    const int sc = screen(u, (float)n * __expf((float)t));
    if (sc != 0) return sc > 0;
    return u < (double)n * exp(y) / zzV;                // src/moves.cpp:326
}

// ln(u) for the log-domain Metropolis test  ln u < -dE/T  <=>  u < exp(-dE/T).  u == 0 maps to the log of
// the smallest positive double, so that (like the reference's `u < exp(-arg)` with an underflowed
// exponential) a move whose Boltzmann factor underflows is never accepted.
__device__ __forceinline__ double log_uniform(double u) { return (u > 0.0) ? log(u) : -745.1332191019412; }

}  // namespace mcpm
