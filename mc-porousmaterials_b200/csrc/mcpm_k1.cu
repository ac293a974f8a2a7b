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


// ================================================================================================
// K1r — k1_server: K1 as a RESIDENT service for one chain (what north_star asks of K1, taken one step further: the chain's SoA fp64
// positions are staged into shared memory ONCE and stay there across calls).  One CTA of 256 threads; thread 0 polls the mailbox
// (system-scope acquire load of mapped host memory), the request is broadcast through shared memory, every thread scans its partners
// j == t (mod 256) for the stored AND the trial position, warp-shuffle butterflies + one shared-memory stage reduce the two sums, and
// thread 0 writes the record and then the sequence number straight into mapped host memory (system-scope release).  Accepted moves of
// the host's loop (mcpm_apply_move / mcpm_append / mcpm_remove_swap_last) arrive through the same mailbox and are written through to
// global memory, so every other kernel keeps seeing the current configuration.  The kernel leaves on STOP or after `idle` cycles
// without a request (it can never outlive its caller by more than that); the host restarts it lazily.
// Chains too large for one SM's shared memory are served from global memory (L2).
// ================================================================================================
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

template <bool FAST, bool PBZ, bool STAGED>
__global__ void __launch_bounds__(K1_THREADS) k1_server(const ChainParams* __restrict__ params, const ChainState* __restrict__ states,
                                                        double* X, double* Y, double* Z, int chain, int cap, K1Mail* mail, K1Reply* reply,
                                                        double* host_box, unsigned long long seq0, long long idle) {
    extern __shared__ double srv_pos[];                   // STAGED: x[cap] | y[cap] | z[cap]
    __shared__ ChainParams P;
    __shared__ int cmd, rq_index, rq_has_trial;
    __shared__ double rq_t[3];
    __shared__ double red[K1_THREADS / 32][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t off = (size_t)chain * (size_t)cap;
    double* gx = X + off; double* gy = Y + off; double* gz = Z + off;
    double* xs = STAGED ? srv_pos : gx;
    double* ys = STAGED ? srv_pos + cap : gy;
    double* zs = STAGED ? srv_pos + 2 * cap : gz;
    if (tid == 0) P = params[chain];
    if (STAGED) for (int j = tid; j < cap; j += K1_THREADS) { xs[j] = gx[j]; ys[j] = gy[j]; zs[j] = gz[j]; }   // every slot: stale ones are observable (Q12)
    int n = states[chain].st.n;
    __syncthreads();
    const PairGeom g = make_geom<FAST>(P);
    unsigned long long expected = seq0 + 1ull;
    for (;;) {
        if (warp == 0) {
            // poll: the 8 words of the mailbox line in ONE warp-wide load (lanes 8..31 repeat the addresses of lanes 0..7)
            const volatile unsigned long long* line = reinterpret_cast<const volatile unsigned long long*>(mail);
            const long long t0 = clock64();
            int c = 0;
            for (;;) {
                const unsigned long long v = line[lane & 7];
                const unsigned long long head = __shfl_sync(MCPM_FULL_MASK, v, 0), tail = __shfl_sync(MCPM_FULL_MASK, v, 7);
                if (head == expected && tail == expected) {
                    const unsigned long long what = __shfl_sync(MCPM_FULL_MASK, v, 1);
                    const unsigned long long t0b = __shfl_sync(MCPM_FULL_MASK, v, 2), t1b = __shfl_sync(MCPM_FULL_MASK, v, 3), t2b = __shfl_sync(MCPM_FULL_MASK, v, 4);
                    c = (int)(what & 0xffull);
                    if (lane == 0) {
                        rq_has_trial = (int)((what >> 8) & 1ull); rq_index = (int)(unsigned)(what >> 32);
                        rq_t[0] = __longlong_as_double((long long)t0b); rq_t[1] = __longlong_as_double((long long)t1b); rq_t[2] = __longlong_as_double((long long)t2b);
                    }
                    break;
                }
                if (__shfl_sync(MCPM_FULL_MASK, clock64() - t0, 0) > idle) break;   // c == 0: nobody has asked for a while
            }
            if (lane == 0) cmd = c;
        }
        __syncthreads();
        const int c = cmd;
        if (c == 0 || c == K1_CMD_STOP) {
            if (tid == 0) {
                ((volatile K1Reply*)reply)->alive = 0ull;
                __threadfence_system();
                if (c == K1_CMD_STOP) { st_release_sys(&reply->seq_tail, expected); st_release_sys(&reply->seq, expected); }
            }
            return;
        }
        const int index = rq_index;
        if (c == K1_CMD_ENERGY) {
            double ax = 0., ay = 0., az = 0., bx = rq_t[0], by = rq_t[1], bz = rq_t[2];
            if (index >= 0) { ax = xs[index]; ay = ys[index]; az = zs[index]; }
            if (!rq_has_trial) { bx = ax; by = ay; bz = az; }
            if (index < 0) { ax = bx; ay = by; az = bz; }
            double sa = 0.0, sb = 0.0;
            pair_sum2<FAST, PBZ>(xs, ys, zs, n, index, g, ax, ay, az, bx, by, bz, tid, K1_THREADS, sa, sb);
            sa = warp_allsum(sa); sb = warp_allsum(sb);
            if (lane == 0) { red[warp][0] = sa; red[warp][1] = sb; }
            __syncthreads();
            if (warp == 0) {
                double wa, wb;
                wall_pair<FAST>(P, ax, ay, az, bx, by, bz, wa, wb, lane);
                double ta = 0.0, tb = 0.0;
                for (int w = 0; w < K1_THREADS / 32; w++) { ta += red[w][0]; tb += red[w][1]; }
                double pa = P.eps4 * ta, pb = P.eps4 * tb;
                if (index < 0) { pa = 0.0; wa = 0.0; }
                // the sequence number, the six numbers and the sequence number again leave in ONE warp-wide store of the 64-byte line
                const double mine = lane == 1 ? pa : lane == 2 ? wa : lane == 3 ? pa + wa : lane == 4 ? pb : lane == 5 ? wb : pb + wb;
                const unsigned long long word = (lane == 0 || lane == 7) ? expected : (unsigned long long)__double_as_longlong(mine);
                if (lane < 8) reinterpret_cast<volatile unsigned long long*>(reply)[lane] = word;
            }
        } else if (c == K1_CMD_BOX) {
            // MC::BoxEnergy (src/energy.cpp:106-124) for a host loop that asks for it every few steps (CorrectEnergy as shipped): per-particle
            // pair and wall energies into mapped host memory (host_box: [cap] pair | [cap] wall), the box totals in the reply line
            double tp = 0.0, tw = 0.0;
            for (int i = tid; i < n; i += K1_THREADS) {
                const double xi = xs[i], yi = ys[i], zi = zs[i];
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;           // four independent dependency chains per thread
                int j = 0;
                for (; j + 3 < n; j += 4) {
                    pair_add<FAST>(s0, dist2<FAST, PBZ>(g, xi, yi, zi, xs[j], ys[j], zs[j]), g, j, i);
                    pair_add<FAST>(s1, dist2<FAST, PBZ>(g, xi, yi, zi, xs[j + 1], ys[j + 1], zs[j + 1]), g, j + 1, i);
                    pair_add<FAST>(s2, dist2<FAST, PBZ>(g, xi, yi, zi, xs[j + 2], ys[j + 2], zs[j + 2]), g, j + 2, i);
                    pair_add<FAST>(s3, dist2<FAST, PBZ>(g, xi, yi, zi, xs[j + 3], ys[j + 3], zs[j + 3]), g, j + 3, i);
                }
                for (; j < n; j++) pair_add<FAST>(s0, dist2<FAST, PBZ>(g, xi, yi, zi, xs[j], ys[j], zs[j]), g, j, i);
                const double sum = (s0 + s1) + (s2 + s3);
                const double pe = P.eps4 * sum, we = wall_single<FAST>(P, xi, yi, zi);
                ((volatile double*)host_box)[i] = pe; ((volatile double*)host_box)[cap + i] = we;
                tp += pe; tw += we;
            }
            tp = warp_allsum(tp); tw = warp_allsum(tw);
            if (lane == 0) { red[warp][0] = tp; red[warp][1] = tw; }
            __threadfence_system();                              // the per-particle records are in host memory before the reply line is
            __syncthreads();
            if (warp == 0) {
                double ta = 0.0, tb = 0.0;
                for (int w = 0; w < K1_THREADS / 32; w++) { ta += red[w][0]; tb += red[w][1]; }
                const double half = 0.5 * ta;
                const double mine = lane == 1 ? half : lane == 2 ? tb : lane == 3 ? half + tb : 0.0;
                const unsigned long long word = (lane == 0 || lane == 7) ? expected : (unsigned long long)__double_as_longlong(mine);
                if (lane < 8) reinterpret_cast<volatile unsigned long long*>(reply)[lane] = word;
            }
        } else {
            // state changes of the host's move loop, written through to global memory
            if (tid == 0) {
                if (c == K1_CMD_APPLY || c == K1_CMD_APPEND) {
                    const int slot = (c == K1_CMD_APPEND) ? n : index;
                    xs[slot] = rq_t[0]; ys[slot] = rq_t[1]; zs[slot] = rq_t[2];
                    if (STAGED) { gx[slot] = rq_t[0]; gy[slot] = rq_t[1]; gz[slot] = rq_t[2]; }
                } else if (c == K1_CMD_REMOVE && index != n - 1) {   // swap-with-last, src/moves.cpp:329
                    const double lx = xs[n - 1], ly = ys[n - 1], lz = zs[n - 1];
                    xs[index] = lx; ys[index] = ly; zs[index] = lz;
                    if (STAGED) { gx[index] = lx; gy[index] = ly; gz[index] = lz; }
                }
                __threadfence_system();
                st_release_sys(&reply->seq_tail, expected);
                st_release_sys(&reply->seq, expected);
            }
            if (c == K1_CMD_APPEND) n++;
            if (c == K1_CMD_REMOVE) n--;
        }
        __syncthreads();                                     // the request buffer and `red` are free again; slot writes are visible
        expected++;
    }
}

size_t k1_server_smem_bytes(int cap, int max_smem_optin) {
    const size_t need = (size_t)3 * cap * sizeof(double);
    return need + 4096 <= (size_t)max_smem_optin ? need : 0;   // 0: serve from global memory
}

cudaError_t launch_k1_server(bool fast, bool pbz, const ChainParams* params, const ChainState* states, double* X, double* Y, double* Z, int chain,
                             int cap, size_t smem, K1Mail* mail, K1Reply* reply, double* host_box, unsigned long long seq0, long long idle_cycles,
                             cudaStream_t stream) {
    auto go = [&](auto kern) -> cudaError_t {
        if (smem) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kern<<<1, K1_THREADS, smem, stream>>>(params, states, X, Y, Z, chain, cap, mail, reply, host_box, seq0, idle_cycles);
        return cudaGetLastError();
    };
    if (smem) {
        if (!fast) return go(k1_server<false, false, true>);
        return pbz ? go(k1_server<true, true, true>) : go(k1_server<true, false, true>);
    }
    if (!fast) return go(k1_server<false, false, false>);
    return pbz ? go(k1_server<true, true, false>) : go(k1_server<true, false, false>);
}

int k1_slices(int n) { return std::max(1, (n + K1_TILE - 1) / K1_TILE); }

}  // namespace mcpm
