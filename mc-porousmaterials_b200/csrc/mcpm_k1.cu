// K1 — single-particle energy, MC::EnergyOfParticle (reference src/energy.cpp:43-105), for a HOST-driven move loop: the energy seam
// of INTEGRATION.md §1 (callers src/moves.cpp:153,165 translation, :304,324 insertion / deletion, :421,429 Widom).
//
// One launch answers a BATCH of requests against the present configuration of one chain.  A request names a stored particle
// (index >= 0) and/or a trial position: `cur` is the energy of the stored particle where it is, `moved` its energy displaced to the
// trial position (itself excluded: what MoveParticle evaluates twice), or — index -1 — the energy of a ghost / inserted particle.
//
// sm_100a design, as north_star asks of this kernel:
//   * the chain's SoA fp64 positions reach SHARED MEMORY by 1D TMA bulk copies (cp.async.bulk global -> shared, completion on an
//     mbarrier): one elected thread issues three copies of up to 8 KB (x, y, z rows of the CTA's slice), the copy engine moves
//     them as coalesced 128-byte lines while the other threads set up — no per-thread load instructions at all;
//   * thread t owns the staged partners j == t (mod 256): each is read once from shared memory into registers and evaluated
//     against BOTH positions of up to NT requests (2*NT independent accumulators: the latency of one ~17-deep fp64 dependency
//     chain hides behind the others);
//   * warp-shuffle butterflies reduce the 2*NT partial sums, one shared-memory stage adds the 8 warps (block-wide reduction);
//     slices of more than TILE particles use several CTAs, the last one to finish adds the slices in order (deterministic);
//   * the finished energies are written to device memory and, for the latency path, straight into MAPPED pinned host memory
//     followed by a sequence flag (system-scope fence in between): the host spins on the flag instead of paying a device->host
//     copy and a stream synchronisation per call.
#include <cuda/ptx>

#include "mcpm_chain.cuh"
#include "mcpm_device.cuh"
#include "mcpm_kernels.h"

namespace mcpm {
namespace ptx = cuda::ptx;

#define K1_TILE 1024      // particles per CTA slice: 3 x 8 KB of shared memory
#define K1_THREADS 256

template <int NT, bool FAST, bool PBZ>
__device__ __forceinline__ void k1_scan(const double* tx, const double* ty, const double* tz, int m, int j0, const PairGeom& g, const K1Request (&rq)[NT],
                                        const double (&ax)[NT], const double (&ay)[NT], const double (&az)[NT], double (&sa)[NT], double (&sb)[NT], int tid) {
    for (int k = tid; k < m; k += K1_THREADS) {
        const double xj = tx[k], yj = ty[k], zj = tz[k];
        const int j = j0 + k;
#pragma unroll
        for (int r = 0; r < NT; r++) {
            pair_add<FAST>(sa[r], dist2<FAST, PBZ>(g, ax[r], ay[r], az[r], xj, yj, zj), g, j, rq[r].index);
            pair_add<FAST>(sb[r], dist2<FAST, PBZ>(g, rq[r].t[0], rq[r].t[1], rq[r].t[2], xj, yj, zj), g, j, rq[r].index);
        }
    }
}

// grid = (request groups of NT, slices of K1_TILE particles).  out: [count][8] = {pairA, 0, wallA, EA, pairB, 0, wallB, EB}.
template <int NT, bool FAST>
__global__ void __launch_bounds__(K1_THREADS) k1_energy(const ChainParams* __restrict__ params, const ChainState* __restrict__ states,
                                                        const double* __restrict__ X, const double* __restrict__ Y, const double* __restrict__ Z,
                                                        int chain, int cap, int count, const K1Request* __restrict__ reqs, K1Request single,
                                                        double* __restrict__ scratch, unsigned int* __restrict__ counters, double* __restrict__ out,
                                                        double* __restrict__ host_out, unsigned long long* host_flag, unsigned long long seq,
                                                        unsigned int* __restrict__ done_groups) {
    __shared__ __align__(128) double tx[K1_TILE];
    __shared__ __align__(128) double ty[K1_TILE];
    __shared__ __align__(128) double tz[K1_TILE];
    __shared__ __align__(8) uint64_t bar;
    __shared__ ChainParams P;
    __shared__ double red[K1_THREADS / 32][2 * NT];
    __shared__ double fin[2 * NT];
    __shared__ bool last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = states[chain].st.n;
    const int j0 = blockIdx.y * K1_TILE;
    const int m = max(0, min(K1_TILE, n - j0));
    const size_t off = (size_t)chain * (size_t)cap;
    if (tid == 0) {
        P = params[chain];
        ptx::mbarrier_init(&bar, 1);
        ptx::fence_proxy_async(ptx::space_shared);          // the barrier's initialisation is visible to the copy engine
    }
    __syncthreads();
    if (tid == 0 && m > 0) {
        // rows are 256-byte aligned (capacity is a multiple of 32) and j0 is a multiple of K1_TILE; sizes are multiples of 16 bytes
        const uint32_t bytes = (uint32_t)((m + 1) & ~1) * (uint32_t)sizeof(double);
        ptx::mbarrier_arrive_expect_tx(ptx::sem_release, ptx::scope_cta, ptx::space_shared, &bar, 3u * bytes);
        ptx::cp_async_bulk(ptx::space_cluster, ptx::space_global, tx, X + off + j0, bytes, &bar);
        ptx::cp_async_bulk(ptx::space_cluster, ptx::space_global, ty, Y + off + j0, bytes, &bar);
        ptx::cp_async_bulk(ptx::space_cluster, ptx::space_global, tz, Z + off + j0, bytes, &bar);
    }
    // ---- the requests of this group, while the tile is in flight ----
    K1Request rq[NT];
    double ax[NT], ay[NT], az[NT], sa[NT], sb[NT];
    const int r0 = blockIdx.x * NT;
#pragma unroll
    for (int r = 0; r < NT; r++) {
        if (reqs) rq[r] = reqs[min(r0 + r, count - 1)]; else rq[r] = single;
        ax[r] = ay[r] = az[r] = 0.0;
        if (rq[r].index >= 0) { ax[r] = X[off + rq[r].index]; ay[r] = Y[off + rq[r].index]; az[r] = Z[off + rq[r].index]; }
        if (!rq[r].has_trial) { rq[r].t[0] = ax[r]; rq[r].t[1] = ay[r]; rq[r].t[2] = az[r]; }
        if (rq[r].index < 0) { ax[r] = rq[r].t[0]; ay[r] = rq[r].t[1]; az[r] = rq[r].t[2]; }
        sa[r] = sb[r] = 0.0;
    }
    const PairGeom g = make_geom<FAST>(P);
    if (m > 0) {
        int spins = 0;
        while (!ptx::mbarrier_try_wait_parity(&bar, 0)) { if (++spins > (1 << 24)) __trap(); }   // a lost copy must not hang the device
        if (P.pbc[2]) k1_scan<NT, FAST, true>(tx, ty, tz, m, j0, g, rq, ax, ay, az, sa, sb, tid);
        else k1_scan<NT, FAST, false>(tx, ty, tz, m, j0, g, rq, ax, ay, az, sa, sb, tid);
    }
    // ---- block-wide reduction: shuffle butterflies, then one shared-memory stage over the 8 warps ----
#pragma unroll
    for (int r = 0; r < NT; r++) { sa[r] = warp_allsum(sa[r]); sb[r] = warp_allsum(sb[r]); }
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < NT; r++) { red[warp][2 * r] = sa[r]; red[warp][2 * r + 1] = sb[r]; }
    }
    __syncthreads();
    const int S = gridDim.y;
    double* sc = scratch + (size_t)blockIdx.x * (size_t)S * (2 * NT);
    if (tid < 2 * NT) {
        double t = 0.0;
        for (int w = 0; w < K1_THREADS / 32; w++) t += red[w][tid];
        if (S == 1) fin[tid] = t; else sc[(size_t)blockIdx.y * (2 * NT) + tid] = t;
    }
    if (S > 1) {
        __threadfence();
        __syncthreads();
        if (tid == 0) last = (atomicAdd(&counters[blockIdx.x], 1u) == (unsigned)S - 1u);
        __syncthreads();
        if (!last) return;
        __threadfence();
        if (tid < 2 * NT) {
            double t = 0.0;
            for (int s = 0; s < S; s++) t += ((volatile double*)sc)[(size_t)s * (2 * NT) + tid];   // slices in order: deterministic
            fin[tid] = t;
        }
        if (tid == 0) counters[blockIdx.x] = 0u;
    }
    __syncthreads();
    // ---- wall terms and the result records (warp 0) ----
    if (warp == 0) {
        for (int r = 0; r < NT; r++) {
            if (r0 + r >= count) break;
            double wa, wb;
            wall_pair<FAST>(P, ax[r], ay[r], az[r], rq[r].t[0], rq[r].t[1], rq[r].t[2], wa, wb, lane);
            if (lane == 0) {
                double pa = P.eps4 * fin[2 * r], pb = P.eps4 * fin[2 * r + 1];
                if (rq[r].index < 0) { pa = 0.0; wa = 0.0; }
                const double rec[8] = {pa, 0.0, wa, pa + wa, pb, 0.0, wb, pb + wb};
                double* o = out + (size_t)(r0 + r) * 8;
#pragma unroll
                for (int k = 0; k < 8; k++) o[k] = rec[k];
                if (host_out) {
                    double* h = host_out + (size_t)(r0 + r) * 8;
#pragma unroll
                    for (int k = 0; k < 8; k++) h[k] = rec[k];
                }
            }
        }
        if (lane == 0 && host_flag) {
            __threadfence_system();                          // the records are in host memory before the flag is
            if (atomicAdd(done_groups, 1u) == gridDim.x - 1u) {   // last group of the launch: raise the flag
                *done_groups = 0u;
                __threadfence_system();
                *((volatile unsigned long long*)host_flag) = seq;
            }
        }
    }
}

cudaError_t launch_k1(bool fast, int n, int count, const ChainParams* params, const ChainState* states, const double* X, const double* Y,
                      const double* Z, int chain, int cap, const K1Request* reqs, const K1Request& single, double* scratch,
                      unsigned int* counters, double* out, double* host_out, unsigned long long* host_flag, unsigned long long seq,
                      unsigned int* done_groups, cudaStream_t stream) {
    const int slices = std::max(1, (n + K1_TILE - 1) / K1_TILE);
    if (count == 1) {
        dim3 grid(1, slices);
        if (fast) k1_energy<1, true><<<grid, K1_THREADS, 0, stream>>>(params, states, X, Y, Z, chain, cap, 1, reqs, single, scratch, counters, out, host_out, host_flag, seq, done_groups);
        else k1_energy<1, false><<<grid, K1_THREADS, 0, stream>>>(params, states, X, Y, Z, chain, cap, 1, reqs, single, scratch, counters, out, host_out, host_flag, seq, done_groups);
    } else {
        dim3 grid((count + 3) / 4, slices);
        if (fast) k1_energy<4, true><<<grid, K1_THREADS, 0, stream>>>(params, states, X, Y, Z, chain, cap, count, reqs, single, scratch, counters, out, host_out, host_flag, seq, done_groups);
        else k1_energy<4, false><<<grid, K1_THREADS, 0, stream>>>(params, states, X, Y, Z, chain, cap, count, reqs, single, scratch, counters, out, host_out, host_flag, seq, done_groups);
    }
    return cudaGetLastError();
}

int k1_slices(int n) { return std::max(1, (n + K1_TILE - 1) / K1_TILE); }

}  // namespace mcpm
