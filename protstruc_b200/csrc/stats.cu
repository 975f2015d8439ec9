// K4 — masked per-structure statistics and the per-structure affine maps around them.
//
// Replaces StructureBatch.standardize / unstandardize (protstruc/protstruc.py:696-744),
// center_of_mass (:746-757) and the in-place translation of center_at / translate
// (:662-679, 759-788).
//
// Roofline: HBM read + write of the (B,L,A,3) coordinates (25 B per atom with a bool mask).
// One thread-block CLUSTER per structure (1, 2, 4 or 8 CTAs, chosen so that a CTA keeps >= ~1k atoms): every CTA
// reduces its contiguous share of the atoms, parks the partial sums in its shared memory and, after a cluster
// barrier, reads the partials of all its peers through distributed shared memory (summed in rank order, so every
// CTA holds the same totals).  Pass 1 (masked sum, count) and pass 2 (masked squared deviation) read the CTA's
// share, which stays in L1/L2, pass 3 writes the normalised coordinates, so HBM sees one read and one write.
// A few large structures (the bench shape: 16 x 7680 atoms) thus occupy 128 CTAs instead of 16.
// Warp-shuffle + shared-memory block reductions;
// partial sums are accumulated in fp64 (the reference sums in fp32 with ATen's cascade summation —
// both are well inside the 1e-5 relative parity tolerance, fp64 is simply the more exact of the two).

#include <cooperative_groups.h>

#include "common.cuh"

namespace ps {

namespace {

constexpr int kStatsMaxThreads = 512;

// torch.nan_to_num(x, nan=0.0): NaN -> 0, +inf -> FLT_MAX, -inf -> -FLT_MAX.  One compare on the common (finite)
// path: |v| <= FLT_MAX is false for infinities and for NaN.
__device__ __forceinline__ float nan_to_num0(float v) {
    constexpr float kMax = 3.4028234663852886e38f;
    if (fabsf(v) <= kMax) return v;
    if (v != v) return 0.f;
    return v > 0.f ? kMax : -kMax;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sums 4 doubles per thread across the block; result valid in every thread.
__device__ __forceinline__ void block_sum4(double (&v)[4], double (*scratch)[4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();  // scratch reuse across calls
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) scratch[warp][k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        double t = lane < nwarps ? scratch[lane][k] : 0.0;
        v[k] = warp_sum(t);
    }
}

// Block sum for the register-resident kernel, whose threads each hold ONE axis (threadIdx.x % 3) and whose block size
// is a multiple of 96: lanes l, l + 3, l + 6, ... of a warp share an axis, so four shuffle steps at strides 24, 12, 6, 3
// leave the three per-axis sums of the warp in lanes 0, 1, 2 (instead of five steps on each of four values); the warps'
// sums meet in shared memory, four threads add them up, everybody reads the total of ITS axis and the total of `extra`
// (the atom count, exact in fp32).  scratch: [warps][4] partials followed by one row of totals.
__device__ __forceinline__ void block_sum_by_axis(double v, float extra, int axis, double (*scratch)[4], double& axis_total,
                                                  double& extra_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int off = 24; off >= 3; off >>= 1) {
        const double a = __shfl_down_sync(0xffffffffu, v, off);
        const float b = __shfl_down_sync(0xffffffffu, extra, off);
        if (lane + off < 32) {
            v += a;
            extra += b;
        }
    }
    __syncthreads();  // scratch reuse across calls
    if (lane < 3) {
        scratch[warp][axis] = v;
        if (axis == 0) scratch[warp][3] = static_cast<double>(extra);
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < nwarps; ++w) t += scratch[w][threadIdx.x];
        scratch[kStatsMaxThreads / 32][threadIdx.x] = t;
    }
    __syncthreads();
    axis_total = scratch[kStatsMaxThreads / 32][axis];
    extra_total = scratch[kStatsMaxThreads / 32][3];
}

template <int MASK_DTYPE>
__device__ __forceinline__ float mask_value(const void* __restrict__ m, long long idx) {
    if (MASK_DTYPE == PS_MASK_BOOL)
        return __ldg(static_cast<const uint8_t*>(m) + idx) != 0 ? 1.f : 0.f;
    return __ldg(static_cast<const float*>(m) + idx);
}

// Totals of 4 doubles over the cluster: own partial -> shared memory, cluster barrier, read every rank's partial
// through distributed shared memory in rank order.  `slot` must differ between the two uses inside one kernel
// (a peer may still be reading the previous slot).
__device__ __forceinline__ void cluster_sum4(double (&v)[4], double (*exchange)[4], int slot) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned ranks = cluster.num_blocks();
    if (ranks == 1) return;
    if (threadIdx.x < 4) exchange[slot][threadIdx.x] = v[threadIdx.x];
    cluster.sync();
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = 0.0;
    for (unsigned r = 0; r < ranks; ++r) {
        const double* peer = cluster.map_shared_rank(&exchange[slot][0], r);
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] += peer[k];
    }
}

// CLUSTER = false is the plain one-CTA-per-structure kernel (the common case: many structures); CLUSTER = true
// is launched with a cluster dimension > 1 for a few large structures and also unrolls the atom loops (the loads of
// kStatsUnroll atoms are issued before the first use: with few CTAs in flight, latency is what there is to hide).
template <int MASK_DTYPE, bool CLUSTER>
__global__ void __launch_bounds__(kStatsMaxThreads) masked_stats_kernel(
    const float* __restrict__ xyz, const void* __restrict__ atom_mask, int atoms_per_struct,
    float* __restrict__ mu_out, float* __restrict__ sd_out, float* __restrict__ xyz_out) {
    namespace cg = cooperative_groups;
    constexpr int kStatsUnroll = CLUSTER ? 4 : 1;
    __shared__ double scratch[kStatsMaxThreads / 32][4];
    __shared__ double exchange[CLUSTER ? 2 : 1][4];
    const int ranks = CLUSTER ? static_cast<int>(cg::this_cluster().num_blocks()) : 1;
    const int rank = CLUSTER ? static_cast<int>(cg::this_cluster().block_rank()) : 0;
    const long long b = blockIdx.x / ranks;
    const float* __restrict__ x = xyz + b * atoms_per_struct * 3;
    const long long m0 = b * atoms_per_struct;
    // this CTA's contiguous share of the structure's atoms
    const int share = (atoms_per_struct + ranks - 1) / ranks;
    const int t_begin = rank * share;
    const int t_end = t_begin + share < atoms_per_struct ? t_begin + share : atoms_per_struct;

    // pass 1: sum(nan_to_num(x * m)) per axis, sum(m)
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if constexpr (!CLUSTER) {
        for (int t = threadIdx.x; t < atoms_per_struct; t += blockDim.x) {
            const float m = mask_value<MASK_DTYPE>(atom_mask, m0 + t);
            acc[0] += static_cast<double>(nan_to_num0(__fmul_rn(__ldg(x + 3 * t + 0), m)));
            acc[1] += static_cast<double>(nan_to_num0(__fmul_rn(__ldg(x + 3 * t + 1), m)));
            acc[2] += static_cast<double>(nan_to_num0(__fmul_rn(__ldg(x + 3 * t + 2), m)));
            acc[3] += static_cast<double>(m);
        }
    } else {
        for (int t0 = t_begin + threadIdx.x; t0 < t_end; t0 += kStatsUnroll * blockDim.x) {
            float m[kStatsUnroll], c[kStatsUnroll][3];
#pragma unroll
            for (int u = 0; u < kStatsUnroll; ++u) {
                const int t = t0 + u * blockDim.x;
                const bool ok = t < t_end;
                m[u] = ok ? mask_value<MASK_DTYPE>(atom_mask, m0 + t) : 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) c[u][k] = ok ? __ldg(x + 3 * t + k) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < kStatsUnroll; ++u) {
                if (t0 + u * blockDim.x >= t_end) break;
                acc[0] += static_cast<double>(nan_to_num0(__fmul_rn(c[u][0], m[u])));
                acc[1] += static_cast<double>(nan_to_num0(__fmul_rn(c[u][1], m[u])));
                acc[2] += static_cast<double>(nan_to_num0(__fmul_rn(c[u][2], m[u])));
                acc[3] += static_cast<double>(m[u]);
            }
        }
    }
    block_sum4(acc, scratch);
    if (CLUSTER) cluster_sum4(acc, exchange, 0);
    const float count = static_cast<float>(acc[3]);
    float mu[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) mu[k] = __fdiv_rn(static_cast<float>(acc[k]), count);

    // pass 2: sum((nan_to_num(x) - mu)^2 * m) per axis
    double dev[4] = {0.0, 0.0, 0.0, 0.0};
    if constexpr (!CLUSTER) {
        for (int t = threadIdx.x; t < atoms_per_struct; t += blockDim.x) {
            const float m = mask_value<MASK_DTYPE>(atom_mask, m0 + t);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float d = __fsub_rn(nan_to_num0(__ldg(x + 3 * t + k)), mu[k]);
                dev[k] += static_cast<double>(__fmul_rn(__fmul_rn(d, d), m));
            }
        }
    } else {
        for (int t0 = t_begin + threadIdx.x; t0 < t_end; t0 += kStatsUnroll * blockDim.x) {
            float m[kStatsUnroll], c[kStatsUnroll][3];
#pragma unroll
            for (int u = 0; u < kStatsUnroll; ++u) {
                const int t = t0 + u * blockDim.x;
                const bool ok = t < t_end;
                m[u] = ok ? mask_value<MASK_DTYPE>(atom_mask, m0 + t) : 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) c[u][k] = ok ? __ldg(x + 3 * t + k) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < kStatsUnroll; ++u) {
                if (t0 + u * blockDim.x >= t_end) break;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float d = __fsub_rn(nan_to_num0(c[u][k]), mu[k]);
                    dev[k] += static_cast<double>(__fmul_rn(__fmul_rn(d, d), m[u]));
                }
            }
        }
    }
    block_sum4(dev, scratch);
    if (CLUSTER) cluster_sum4(dev, exchange, 1);
    float sd[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) sd[k] = __fsqrt_rn(__fdiv_rn(static_cast<float>(dev[k]), count));

    if (rank == 0 && threadIdx.x < 3) {
        mu_out[b * 3 + threadIdx.x] = mu[threadIdx.x];
        sd_out[b * 3 + threadIdx.x] = sd[threadIdx.x];
    }

    // pass 3: (x - mu) / sd on every atom of the share (masked or not, NaN stays NaN)
    if (xyz_out) {
        float* __restrict__ o = xyz_out + b * atoms_per_struct * 3;
        if constexpr (!CLUSTER) {
            const int n = atoms_per_struct * 3;
            if (blockDim.x % 3 == 0) {
                // the host launches a multiple of 96 threads: a thread's elements all lie on ONE axis, so its mean
                // and deviation are picked once, outside the loop
                const int k = threadIdx.x % 3;
                const float m = k == 0 ? mu[0] : (k == 1 ? mu[1] : mu[2]);
                const float s = k == 0 ? sd[0] : (k == 1 ? sd[1] : sd[2]);
                for (int e = threadIdx.x; e < n; e += blockDim.x) o[e] = __fdiv_rn(__fsub_rn(x[e], m), s);
            } else {
                for (int e = threadIdx.x; e < n; e += blockDim.x) {
                    const int k = e % 3;
                    const float m = k == 0 ? mu[0] : (k == 1 ? mu[1] : mu[2]);
                    const float s = k == 0 ? sd[0] : (k == 1 ? sd[1] : sd[2]);
                    o[e] = __fdiv_rn(__fsub_rn(x[e], m), s);
                }
            }
        } else {
            for (int e0 = t_begin * 3 + threadIdx.x; e0 < t_end * 3; e0 += kStatsUnroll * blockDim.x) {
                float v[kStatsUnroll];
#pragma unroll
                for (int u = 0; u < kStatsUnroll; ++u) {
                    const int e = e0 + u * blockDim.x;
                    v[u] = e < t_end * 3 ? x[e] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < kStatsUnroll; ++u) {
                    const int e = e0 + u * blockDim.x;
                    if (e >= t_end * 3) break;
                    const int k = e % 3;
                    const float m = k == 0 ? mu[0] : (k == 1 ? mu[1] : mu[2]);
                    const float s = k == 0 ? sd[0] : (k == 1 ? sd[1] : sd[2]);
                    o[e] = __fdiv_rn(__fsub_rn(v[u], m), s);
                }
            }
        }
    }
    // a CTA's shared memory must stay valid until every peer has read its partial sums
    if (CLUSTER) cg::this_cluster().sync();
}

// ------------------------------------------------------------------------------------------------------------------
// Register-resident variant (the default whenever a CTA's share of a structure fits in registers).
//
// The three-pass kernel above re-reads its share from L1/L2 for the deviation and for the normalisation, and each pass
// is a chain of dependent scalar loads: ncu showed it latency-bound (issue-active 38 %, top stall long_scoreboard),
// not instruction-bound.  Here every thread issues ALL of its loads up front — E independent coalesced 32-bit loads of
// coordinates plus the matching mask loads, i.e. the whole share of the CTA is in flight at once — and keeps the
// values in registers through mean -> deviation -> normalise, so the structure is read exactly once.
// Element mapping: thread t owns floats t, t + T, t + 2 T, ... of the share with T a multiple of 3, so all of a
// thread's elements lie on ONE axis (t % 3): its mean and deviation are scalars, and the atom index of element k is
// t / 3 + k * T / 3 (no division in the loops).  Works for any alignment and any atom count.
// Arithmetic per element is unchanged (reference op order, fp64 partial sums).
template <int MASK_DTYPE, int E, bool CLUSTER>
__global__ void __launch_bounds__(kStatsMaxThreads) masked_stats_regs_kernel(
    const float* __restrict__ xyz, const void* __restrict__ atom_mask, int atoms_per_struct,
    float* __restrict__ mu_out, float* __restrict__ sd_out, float* __restrict__ xyz_out) {
    namespace cg = cooperative_groups;
    __shared__ double scratch[kStatsMaxThreads / 32 + 1][4];
    __shared__ double exchange[CLUSTER ? 2 : 1][4];
    const int ranks = CLUSTER ? static_cast<int>(cg::this_cluster().num_blocks()) : 1;
    const int rank = CLUSTER ? static_cast<int>(cg::this_cluster().block_rank()) : 0;
    const long long b = blockIdx.x / ranks;
    const int share = (atoms_per_struct + ranks - 1) / ranks;  // atoms per CTA
    const int a_begin = rank * share;
    const int a_end = a_begin + share < atoms_per_struct ? a_begin + share : atoms_per_struct;
    const int n = (a_end - a_begin) * 3;  // floats of this CTA (may be <= 0 for the last ranks of a short structure)
    const int T = blockDim.x, axis = threadIdx.x % 3, astep = T / 3;
    // this thread's first element / first atom; element k is T floats (astep atoms) further on
    const float* __restrict__ x = xyz + (b * atoms_per_struct + a_begin) * 3 + threadIdx.x;
    const long long m_first = b * atoms_per_struct + a_begin + threadIdx.x / 3;

    float v[E], m[E];
#pragma unroll
    for (int k = 0; k < E; ++k) {
        const bool ok = static_cast<int>(threadIdx.x) + k * T < n;
        v[k] = ok ? __ldg(x + k * T) : 0.f;
        m[k] = ok ? mask_value<MASK_DTYPE>(atom_mask, m_first + k * astep) : 0.f;
    }
    // pass 1: sum(nan_to_num(x * m)) on this thread's axis, sum(m) counted by the axis-0 thread of each atom.
    // A thread's <= E values are added in fp32 and enter the fp64 reduction as ONE value: fp32 -> fp64 conversions
    // run on the quarter-rate XU pipe, and one per element (as in the three-pass kernel) was what bounded both kernels.
    float s = 0.f, c = 0.f;
#pragma unroll
    for (int k = 0; k < E; ++k) {
        s += nan_to_num0(__fmul_rn(v[k], m[k]));
        c += m[k];
    }
    double sum_axis, sum_count;
    block_sum_by_axis(static_cast<double>(s), axis == 0 ? c : 0.f, axis, scratch, sum_axis, sum_count);
    if (CLUSTER) {
        // totals of the whole structure: the three axis sums and the count of every CTA of the cluster
        double acc[4] = {scratch[kStatsMaxThreads / 32][0], scratch[kStatsMaxThreads / 32][1],
                         scratch[kStatsMaxThreads / 32][2], scratch[kStatsMaxThreads / 32][3]};
        cluster_sum4(acc, exchange, 0);
        sum_axis = axis == 0 ? acc[0] : (axis == 1 ? acc[1] : acc[2]);
        sum_count = acc[3];
    }
    // every thread needs the statistics of ITS axis only: one IEEE division here, one division + square root below
    const float count = static_cast<float>(sum_count);
    const float my_mu = __fdiv_rn(static_cast<float>(sum_axis), count);

    // pass 2: sum((nan_to_num(x) - mu)^2 * m)
    float d2 = 0.f;
#pragma unroll
    for (int k = 0; k < E; ++k) {
        const float d = __fsub_rn(nan_to_num0(v[k]), my_mu);
        d2 += __fmul_rn(__fmul_rn(d, d), m[k]);
    }
    double dev_axis, unused;
    block_sum_by_axis(static_cast<double>(d2), 0.f, axis, scratch, dev_axis, unused);
    if (CLUSTER) {
        double acc[4] = {scratch[kStatsMaxThreads / 32][0], scratch[kStatsMaxThreads / 32][1],
                         scratch[kStatsMaxThreads / 32][2], 0.0};
        cluster_sum4(acc, exchange, 1);
        dev_axis = axis == 0 ? acc[0] : (axis == 1 ? acc[1] : acc[2]);
    }
    const float my_sd = __fsqrt_rn(__fdiv_rn(static_cast<float>(dev_axis), count));
    if (rank == 0 && threadIdx.x < 3) {  // threads 0, 1, 2 hold axes 0, 1, 2
        mu_out[b * 3 + threadIdx.x] = my_mu;
        sd_out[b * 3 + threadIdx.x] = my_sd;
    }

    // pass 3: (x - mu) / sd on every element of the share (masked or not, NaN stays NaN), straight from registers
    if (xyz_out) {
        float* __restrict__ o = xyz_out + (b * atoms_per_struct + a_begin) * 3 + threadIdx.x;
#pragma unroll
        for (int k = 0; k < E; ++k)
            if (static_cast<int>(threadIdx.x) + k * T < n) o[k * T] = __fdiv_rn(__fsub_rn(v[k], my_mu), my_sd);
    }
    if (CLUSTER) cg::this_cluster().sync();  // peers may still be reading this CTA's partial sums
}

// ------------------------------------------------------------------------------------------------------------------
// Quad variant of the register-resident kernel: the default whenever a structure holds a multiple of 4 atoms and the
// arrays are 16-byte aligned (A = 15 with L a multiple of 4, A = 4 / 8 / 12 / 16 / 20 ... with any L).
//
// ncu of the scalar-mapped kernel above at BASELINE config 4: 314 thread-instructions per atom, IPC 2.1, i.e. bound by
// the NUMBER of instructions: 32-bit loads / stores with their 64-bit addressing, one byte load + conversion per
// element for the mask, an IEEE division (12 instructions) per element, nan_to_num twice per element.  Here a thread
// owns Q groups of FOUR CONSECUTIVE ATOMS: 12 floats = three 128-bit loads with a fixed axis pattern
// (x0 y0 z0 x1 | y1 z1 x2 y2 | z2 x3 y3 z3), the four mask bytes = one 32-bit load, results leave as three 128-bit
// stores; finiteness is tested once per element (a finite coordinate times a 0 / 1 mask needs no nan_to_num); the
// division by the deviation is a reciprocal + one correction step (the fast path of the IEEE division without its
// range checks; non-finite quotients are redone with the IEEE division); mean / deviation are computed by three lanes
// per warp and shuffled.  Arithmetic per element is otherwise unchanged (reference op order).
__device__ __forceinline__ float quad_get(const float4 (&v)[3], int s) {
    const float4 q = v[s >> 2];
    return (s & 3) == 0 ? q.x : ((s & 3) == 1 ? q.y : ((s & 3) == 2 ? q.z : q.w));
}
__device__ __forceinline__ void quad_set(float4 (&v)[3], int s, float val) {
    float4& q = v[s >> 2];
    if ((s & 3) == 0) q.x = val;
    else if ((s & 3) == 1) q.y = val;
    else if ((s & 3) == 2) q.z = val;
    else q.w = val;
}

// Sums three fp32 per-thread partials and a count over the block; totals valid in every thread.  scratch: [warps + 1][4].
// The five shuffle levels inside a warp stay in fp32 (a pairwise tree over 32 partials of <= 16 values each; a double
// costs two SHFL per level), the warps' sums meet in fp64.
__device__ __forceinline__ void block_sum3f(const float (&p)[3], float c, double (*scratch)[4], double (&total)[3],
                                            float& count) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float v0 = p[0], v1 = p[1], v2 = p[2];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        v2 += __shfl_xor_sync(0xffffffffu, v2, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    __syncthreads();  // scratch reuse across calls
    if (lane == 0) {
        scratch[warp][0] = static_cast<double>(v0);
        scratch[warp][1] = static_cast<double>(v1);
        scratch[warp][2] = static_cast<double>(v2);
        scratch[warp][3] = static_cast<double>(c);
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < nwarps; ++w) t += scratch[w][threadIdx.x];
        scratch[kStatsMaxThreads / 32][threadIdx.x] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 3; ++k) total[k] = scratch[kStatsMaxThreads / 32][k];
    count = static_cast<float>(scratch[kStatsMaxThreads / 32][3]);
}

__device__ __forceinline__ float max3_abs(float a, float b, float c) {  // max(a, |b|, |c|) ignoring NaN: one FMNMX3
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(fabsf(b)), "f"(fabsf(c)));
    return r;
}

template <int MASK_DTYPE, int Q, bool CLUSTER, int MAXT = kStatsMaxThreads, int MINB = 1>
__global__ void __launch_bounds__(MAXT, MINB) masked_stats_quad_kernel(
    const float* __restrict__ xyz, const void* __restrict__ atom_mask, int atoms_per_struct,
    float* __restrict__ mu_out, float* __restrict__ sd_out, float* __restrict__ xyz_out) {
    namespace cg = cooperative_groups;
    __shared__ double scratch[kStatsMaxThreads / 32 + 1][4];
    __shared__ double exchange[CLUSTER ? 2 : 1][4];
    const int ranks = CLUSTER ? static_cast<int>(cg::this_cluster().num_blocks()) : 1;
    const int rank = CLUSTER ? static_cast<int>(cg::this_cluster().block_rank()) : 0;
    const long long b = blockIdx.x / ranks;
    const int quads = atoms_per_struct >> 2;
    const int share = (quads + ranks - 1) / ranks;  // quads per CTA
    const int q_begin = rank * share;
    const int nq = (q_begin + share < quads ? q_begin + share : quads) - q_begin;  // may be <= 0 for the last ranks
    const long long first_atom = b * atoms_per_struct + 4ll * q_begin;
    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(xyz + first_atom * 3);
    const int T = blockDim.x;

    float4 v[Q][3];
    float m[Q][4];
#pragma unroll
    for (int k = 0; k < Q; ++k) {
        const int q = threadIdx.x + k * T;
        const bool ok = q < nq;
#pragma unroll
        for (int i = 0; i < 3; ++i) v[k][i] = ok ? __ldg(x4 + 3 * q + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (MASK_DTYPE == PS_MASK_BOOL) {
            const uint32_t w = ok ? __ldg(reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(atom_mask) + first_atom) + q) : 0u;
#pragma unroll
            for (int a = 0; a < 4; ++a) m[k][a] = ((w >> (8 * a)) & 0xFFu) != 0u ? 1.f : 0.f;
        } else {
            const float4 w = ok ? __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(atom_mask) + first_atom) + q)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            m[k][0] = w.x; m[k][1] = w.y; m[k][2] = w.z; m[k][3] = w.w;
        }
    }
    constexpr float kMax = 3.4028234663852886e38f;
    constexpr bool kBoolMask = MASK_DTYPE == PS_MASK_BOOL;
    // pass 1: sum(nan_to_num(x * m)) per axis, sum(m); per-thread partials in fp32 (<= 4 Q values per axis).
    // Straight-line: NaN (a missing atom — half of the atoms of a real batch) becomes 0 by a select, infinities only
    // raise a flag, and a thread that saw one redoes its sums through the full nan_to_num.  `inf_x` is reused by pass 2.
    // With a 0 / 1 mask the product x * m is exact, so  s + nan_to_num(x * m)  is ONE fused multiply-add of the
    // NaN-cleaned x (same bits as the reference's multiply, nan_to_num, add), and |x * m| <= |x|: one running maximum
    // of |x| (an FMNMX3 per two elements) replaces the two infinity tests per element.
    float s[3] = {0.f, 0.f, 0.f}, c = 0.f;
    bool inf_x = false, inf_xm = false;
    float amax = 0.f;  // max |x| over this thread's elements, NaN ignored (bool masks)
#pragma unroll
    for (int k = 0; k < Q; ++k) {
        if (kBoolMask) {
#pragma unroll
            for (int e = 0; e < 12; e += 2) {
                const float x0 = quad_get(v[k], e), x1 = quad_get(v[k], e + 1);
                s[e % 3] = __fmaf_rn((x0 == x0) ? x0 : 0.f, m[k][e / 3], s[e % 3]);
                s[(e + 1) % 3] = __fmaf_rn((x1 == x1) ? x1 : 0.f, m[k][(e + 1) / 3], s[(e + 1) % 3]);
                amax = max3_abs(amax, x0, x1);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 12; ++e) {
                const float x = quad_get(v[k], e);
                const float t = __fmul_rn(x, m[k][e / 3]);
                inf_x |= fabsf(x) > kMax;
                inf_xm |= fabsf(t) > kMax;
                s[e % 3] += (t == t) ? t : 0.f;
            }
        }
        c += (m[k][0] + m[k][1]) + (m[k][2] + m[k][3]);
    }
    if (kBoolMask) inf_x = inf_xm = amax > kMax;
    if (inf_xm) {
        s[0] = s[1] = s[2] = 0.f;
#pragma unroll
        for (int k = 0; k < Q; ++k)
#pragma unroll
            for (int e = 0; e < 12; ++e) s[e % 3] += nan_to_num0(__fmul_rn(quad_get(v[k], e), m[k][e / 3]));
    }
    double acc[3];
    block_sum3f(s, c, scratch, acc, c);
    if (CLUSTER) {
        double a4[4] = {acc[0], acc[1], acc[2], static_cast<double>(c)};
        cluster_sum4(a4, exchange, 0);
        acc[0] = a4[0]; acc[1] = a4[1]; acc[2] = a4[2];
        c = static_cast<float>(a4[3]);
    }
    // mean: lanes 0 .. 2 of every warp take one IEEE division each, the warp shares the results
    const int lane = threadIdx.x & 31;
    const float count = c;
    const float my_num = static_cast<float>(lane % 3 == 0 ? acc[0] : (lane % 3 == 1 ? acc[1] : acc[2]));
    const float my_mu = __fdiv_rn(my_num, count);
    float mu[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) mu[k] = __shfl_sync(0xffffffffu, my_mu, k);

    // pass 2: sum((nan_to_num(x) - mu)^2 * m) per axis  (0 / 1 mask: (d * d) * m is exact -> fused into the add)
    float d2[3] = {0.f, 0.f, 0.f};
    if (!inf_x) {
#pragma unroll
        for (int k = 0; k < Q; ++k) {
#pragma unroll
            for (int e = 0; e < 12; ++e) {
                const float x = quad_get(v[k], e);
                const float d = __fsub_rn((x == x) ? x : 0.f, mu[e % 3]);
                if (kBoolMask) d2[e % 3] = __fmaf_rn(__fmul_rn(d, d), m[k][e / 3], d2[e % 3]);
                else d2[e % 3] += __fmul_rn(__fmul_rn(d, d), m[k][e / 3]);
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < Q; ++k) {
#pragma unroll
            for (int e = 0; e < 12; ++e) {
                const float d = __fsub_rn(nan_to_num0(quad_get(v[k], e)), mu[e % 3]);
                d2[e % 3] += __fmul_rn(__fmul_rn(d, d), m[k][e / 3]);
            }
        }
    }
    double dev[3];
    float unused = 0.f;
    block_sum3f(d2, 0.f, scratch, dev, unused);
    if (CLUSTER) {
        double a4[4] = {dev[0], dev[1], dev[2], 0.0};
        cluster_sum4(a4, exchange, 1);
        dev[0] = a4[0]; dev[1] = a4[1]; dev[2] = a4[2];
    }
    const float my_dev = static_cast<float>(lane % 3 == 0 ? dev[0] : (lane % 3 == 1 ? dev[1] : dev[2]));
    const float my_sd = __fsqrt_rn(__fdiv_rn(my_dev, count));
    const float my_rcp = __frcp_rn(my_sd);
    float sd[3], rcp[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        sd[k] = __shfl_sync(0xffffffffu, my_sd, k);
        rcp[k] = __shfl_sync(0xffffffffu, my_rcp, k);
    }
    if (rank == 0 && threadIdx.x < 3) {
        mu_out[b * 3 + threadIdx.x] = my_mu;
        sd_out[b * 3 + threadIdx.x] = my_sd;
    }

    // pass 3: (x - mu) / sd on every element (masked or not, NaN stays NaN), straight from registers
    if (xyz_out) {
        float4* __restrict__ o4 = reinterpret_cast<float4*>(xyz_out + first_atom * 3);
        // d / sd as reciprocal + one correction of the quotient (what the IEEE division does for operands in range).
        // Bool masks: whether that is safe is decided ONCE per thread — its coordinates are below 1e18 in magnitude
        // (the running maximum of pass 1), the means too, the deviations within [1e-18, 1e18] — so that no quotient
        // can overflow or meet a zero / infinite / NaN deviation; anything else takes the IEEE division.
        bool fast = false;
        if (kBoolMask) {
            fast = amax <= 1e18f;
#pragma unroll
            for (int k = 0; k < 3; ++k) fast = fast && (fabsf(mu[k]) <= 1e18f) && (sd[k] >= 1e-18f) && (sd[k] <= 1e18f);
        }
#pragma unroll
        for (int k = 0; k < Q; ++k) {
            const int q = threadIdx.x + k * T;
            if (q >= nq) continue;
            float4 out[3];
            if (fast) {
#pragma unroll
                for (int e = 0; e < 12; ++e) {
                    const float d = __fsub_rn(quad_get(v[k], e), mu[e % 3]);
                    float qt = __fmul_rn(d, rcp[e % 3]);
                    qt = __fmaf_rn(__fmaf_rn(-qt, sd[e % 3], d), rcp[e % 3], qt);
                    quad_set(out, e, qt);
                }
            } else {
                // per-element test: if an outcome is not finite although its coordinate is a number — sd = 0 or NaN,
                // infinite operands — the group is redone with the IEEE division
                bool redo = false;
#pragma unroll
                for (int e = 0; e < 12; ++e) {
                    const float d = __fsub_rn(quad_get(v[k], e), mu[e % 3]);
                    float qt = __fmul_rn(d, rcp[e % 3]);
                    qt = __fmaf_rn(__fmaf_rn(-qt, sd[e % 3], d), rcp[e % 3], qt);
                    redo |= (d == d) & !(fabsf(qt) <= kMax);  // a NaN coordinate stays NaN: nothing to redo
                    quad_set(out, e, qt);
                }
                if (redo) {
#pragma unroll
                    for (int e = 0; e < 12; ++e)
                        quad_set(out, e, __fdiv_rn(__fsub_rn(quad_get(v[k], e), mu[e % 3]), sd[e % 3]));
                }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) o4[3 * q + i] = out[i];
        }
    }
    if (CLUSTER) cg::this_cluster().sync();  // peers may still be reading this CTA's partial sums
}

// The per-structure elementwise maps run on a 2-D grid: blockIdx.y walks the structures, blockIdx.x / threadIdx.x
// the structure's L*A*3 floats with fully coalesced scalar accesses.  The axis of an element is its 32-bit offset
// inside the structure mod 3 (a multiply-shift); no per-element 64-bit division.

// Four independent loads per thread are issued before the first use (kUnroll), so a thread keeps 4 x 128 B per
// warp in flight instead of one dependent load -> store chain.
constexpr int kUnroll = 4;

// out = x * scale[b, axis] + shift[b, axis]  (unstandardize) — two rounded ops.
__global__ void __launch_bounds__(256) scale_shift_kernel(const float* __restrict__ x,
                                                          const float* __restrict__ scale,
                                                          const float* __restrict__ shift,
                                                          int per_b, int B, float* __restrict__ out) {
    const unsigned n = static_cast<unsigned>(per_b), step = gridDim.x * blockDim.x;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const float* __restrict__ xb = x + static_cast<long long>(b) * per_b;
        float* __restrict__ ob = out + static_cast<long long>(b) * per_b;
        const float sc0 = __ldg(scale + b * 3), sc1 = __ldg(scale + b * 3 + 1), sc2 = __ldg(scale + b * 3 + 2);
        const float sh0 = __ldg(shift + b * 3), sh1 = __ldg(shift + b * 3 + 1), sh2 = __ldg(shift + b * 3 + 2);
        for (unsigned e0 = blockIdx.x * blockDim.x + threadIdx.x; e0 < n; e0 += kUnroll * step) {
            float v[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) v[u] = e0 + u * step < n ? xb[e0 + u * step] : 0.f;
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const unsigned e = e0 + u * step;
                if (e >= n) break;
                const unsigned k = e % 3u;
                const float sc = k == 0 ? sc0 : (k == 1 ? sc1 : sc2);
                const float sh = k == 0 ? sh0 : (k == 1 ? sh1 : sh2);
                ob[e] = __fadd_rn(__fmul_rn(v[u], sc), sh);
            }
        }
    }
}

// out = x + t[b or 0, axis].  x and out may be the SAME buffer (center_at works in place like the reference's `+=`),
// so neither is __restrict__: a thread reads an element before it writes that same element and touches no other.
__global__ void __launch_bounds__(256) translate_kernel(const float* x,
                                                        const float* __restrict__ t, int t_rows,
                                                        int per_b, int B, float* out) {
    const unsigned n = static_cast<unsigned>(per_b), step = gridDim.x * blockDim.x;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const float* xb = x + static_cast<long long>(b) * per_b;
        float* ob = out + static_cast<long long>(b) * per_b;
        const float* __restrict__ tb = t + (t_rows == 1 ? 0 : b * 3);
        const float t0 = __ldg(tb), t1 = __ldg(tb + 1), t2 = __ldg(tb + 2);
        for (unsigned e0 = blockIdx.x * blockDim.x + threadIdx.x; e0 < n; e0 += kUnroll * step) {
            float v[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) v[u] = e0 + u * step < n ? xb[e0 + u * step] : 0.f;
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const unsigned e = e0 + u * step;
                if (e >= n) break;
                const unsigned k = e % 3u;
                ob[e] = __fadd_rn(v[u], k == 0 ? t0 : (k == 1 ? t1 : t2));
            }
        }
    }
}

// float4 flavour of the two maps above for structures of a multiple of 4 floats on 16-byte aligned arrays (ncu: the
// scalar kernels take 15-16 us for BASELINE config 4's 23.6 MB where ATen's vectorised elementwise kernels take 9.2).
// Float c of float4 f of a structure is coordinate axis (f + c) mod 3.  The block size (192) and therefore the grid
// stride are multiples of 3, so a thread meets ONE rotation: it rotates the three per-axis coefficients once per
// structure and then runs load.128 -> 4 x (multiply, add) -> store.128, four float4 in flight.
// SHIFT_ONLY: out = x + t (translate; x and out may alias), else out = x * scale + shift (two rounded operations).
template <bool SHIFT_ONLY>
__global__ void __launch_bounds__(192) axis_map_vec_kernel(const float* x, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, int shift_rows, int per_b4,
                                                           int B, float* out) {
    const unsigned n4 = static_cast<unsigned>(per_b4), step = gridDim.x * blockDim.x;
    const unsigned first = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned rot = first % 3u;  // axis of float 0 of every float4 of this thread
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const float4* xb = reinterpret_cast<const float4*>(x) + static_cast<long long>(b) * per_b4;
        float4* ob = reinterpret_cast<float4*>(out) + static_cast<long long>(b) * per_b4;
        const float* __restrict__ sh = shift + (shift_rows == 1 ? 0 : b * 3);
        const float h0 = __ldg(sh), h1 = __ldg(sh + 1), h2 = __ldg(sh + 2);
        // coefficients by position c mod 3: position p holds axis (rot + p) mod 3
        const float hp0 = rot == 0 ? h0 : (rot == 1 ? h1 : h2);
        const float hp1 = rot == 0 ? h1 : (rot == 1 ? h2 : h0);
        const float hp2 = rot == 0 ? h2 : (rot == 1 ? h0 : h1);
        float sp0 = 1.f, sp1 = 1.f, sp2 = 1.f;
        if (!SHIFT_ONLY) {
            const float c0 = __ldg(scale + b * 3), c1 = __ldg(scale + b * 3 + 1), c2 = __ldg(scale + b * 3 + 2);
            sp0 = rot == 0 ? c0 : (rot == 1 ? c1 : c2);
            sp1 = rot == 0 ? c1 : (rot == 1 ? c2 : c0);
            sp2 = rot == 0 ? c2 : (rot == 1 ? c0 : c1);
        }
        for (unsigned f0 = first; f0 < n4; f0 += kUnroll * step) {
            float4 v[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                v[u] = f0 + u * step < n4 ? xb[f0 + u * step] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const unsigned f = f0 + u * step;
                if (f >= n4) break;
                float4 r;
                if (SHIFT_ONLY) {
                    r = make_float4(__fadd_rn(v[u].x, hp0), __fadd_rn(v[u].y, hp1), __fadd_rn(v[u].z, hp2), __fadd_rn(v[u].w, hp0));
                } else {
                    r = make_float4(__fadd_rn(__fmul_rn(v[u].x, sp0), hp0), __fadd_rn(__fmul_rn(v[u].y, sp1), hp1),
                                    __fadd_rn(__fmul_rn(v[u].z, sp2), hp2), __fadd_rn(__fmul_rn(v[u].w, sp0), hp0));
                }
                ob[f] = r;
            }
        }
    }
}

// Grid of the float4 maps: x = CTAs of 192 threads so that a thread handles about four float4 of its structure.
dim3 per_structure_grid_vec(int per_b4, int B) {
    int gx = (per_b4 + 4 * 192 - 1) / (4 * 192);
    if (gx < 1) gx = 1;
    if (gx > 32) gx = 32;
    return dim3(static_cast<unsigned>(gx), static_cast<unsigned>(B < 65535 ? B : 65535), 1);
}

// nanmean over residues of one atom slot.  One CTA per structure (the first version gave a structure to ONE warp:
// 6 % of the warp slots busy at 256 structures, every lane a chain of dependent strided loads): a thread issues the
// three loads of up to kComUnroll residues before the first use, per-thread fp32 NaN-skipping sums go through fp64
// warp-shuffle / shared-memory reductions.  The loads are strided by the residue (A * 12 B), so a 32-byte sector
// carries 12 useful bytes: the algorithmic 12 B per residue cost ~32 B of L2 traffic whatever the kernel does.
constexpr int kComUnroll = 4;

__global__ void __launch_bounds__(256) center_of_mass_kernel(const float* __restrict__ xyz, int L, int A, int slot,
                                                             float* __restrict__ out) {
    __shared__ double scratch[8][6];
    const long long b = blockIdx.x;
    const float* __restrict__ x = xyz + (b * L * A + slot) * 3;
    const long long stride = static_cast<long long>(A) * 3;
    double s[3] = {0.0, 0.0, 0.0}, n[3] = {0.0, 0.0, 0.0};
    for (int l0 = threadIdx.x; l0 < L; l0 += kComUnroll * blockDim.x) {
        float v[kComUnroll][3];
#pragma unroll
        for (int u = 0; u < kComUnroll; ++u) {
            const int l = l0 + u * blockDim.x;
#pragma unroll
            for (int k = 0; k < 3; ++k) v[u][k] = l < L ? __ldg(x + l * stride + k) : __int_as_float(0x7fc00000);
        }
#pragma unroll
        for (int u = 0; u < kComUnroll; ++u)
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (v[u][k] == v[u][k]) {
                    s[k] += static_cast<double>(v[u][k]);
                    n[k] += 1.0;
                }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        s[k] = warp_sum(s[k]);
        n[k] = warp_sum(n[k]);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            scratch[warp][k] = s[k];
            scratch[warp][3 + k] = n[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double sk = 0.0, nk = 0.0;
        for (int w = 0; w < nwarps; ++w) {
            sk += scratch[w][threadIdx.x];
            nk += scratch[w][3 + threadIdx.x];
        }
        // torch.nanmean = nansum / count, both in fp32
        out[b * 3 + threadIdx.x] = __fdiv_rn(static_cast<float>(sk), static_cast<float>(nk));
    }
}

// Grid of the per-structure elementwise kernels: y = structures (capped, the kernel loops), x = enough CTAs of 256
// threads that a thread handles about four elements of its structure.
dim3 per_structure_grid(int per_b, int B) {
    int gx = (per_b + 1023) / 1024;
    if (gx < 1) gx = 1;
    if (gx > 32) gx = 32;
    return dim3(static_cast<unsigned>(gx), static_cast<unsigned>(B < 65535 ? B : 65535), 1);
}

}  // namespace

// legacy: 0 = default (quad kernel where the shape allows, else the scalar-mapped register kernel), 1 = the three-pass
// kernel of round 1, 2 = the scalar-mapped register kernel, 3 = the quad kernel without the dense 4-quads-per-thread
// configuration (comparison hooks, ps_masked_stats_ex).
int masked_stats_variant_impl(const float* xyz, const void* atom_mask, int mask_dtype, int B, int L, int A,
                              float* mu, float* sd, float* xyz_out, int legacy, cudaStream_t stream) {
    PS_REQUIRE(B > 0 && L > 0 && A > 0, PS_ERR_BAD_SHAPE, "masked_stats: B=%d L=%d A=%d must be > 0",
               B, L, A);
    PS_REQUIRE(xyz && atom_mask && mu && sd, PS_ERR_NULL_POINTER, "masked_stats: NULL pointer");
    PS_REQUIRE(static_cast<long long>(L) * A * 3 < (1ll << 31), PS_ERR_BAD_SHAPE,
               "masked_stats: structure too large (L*A*3 >= 2^31)");
    const int atoms = L * A;
    PS_REQUIRE(mask_dtype == PS_MASK_BOOL || mask_dtype == PS_MASK_F32, PS_ERR_BAD_DTYPE,
               "masked_stats: unknown mask_dtype %d", mask_dtype);
    const int sms = sm_count_for_current_device();
    if (sms < 0) return sms;
    PS_REQUIRE(static_cast<long long>(B) * 8 < (1ll << 31), PS_ERR_BAD_SHAPE, "masked_stats: B=%d too large", B);
    const bool is_bool = mask_dtype == PS_MASK_BOOL;

    // ---- quad kernel: structures of a multiple of 4 atoms, 16-byte aligned arrays (the usual case)
    auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    if ((legacy == 0 || legacy == 3) && atoms % 4 == 0 && aligned16(xyz) && aligned16(xyz_out) && aligned16(atom_mask)) {
        constexpr int kMaxQ = 4, kMaxTq = 512;
        const int quads = atoms / 4;
        int ranks_q = 1;
        // a cluster (DSMEM exchange, three cluster barriers) only when there are too few structures to give every
        // other SM a CTA, or a share would not fit in registers: at 256 structures of 7680 atoms a cluster of 2 ran
        // at 39 us against ~20 for one 480-thread CTA per structure
        while (ranks_q < 8 && quads / (ranks_q * 2) >= 32 &&
               (static_cast<long long>(B) * ranks_q * 2 <= sms || (quads + ranks_q - 1) / ranks_q > kMaxQ * kMaxTq))
            ranks_q *= 2;
        const int share_q = (quads + ranks_q - 1) / ranks_q;
        if (share_q <= kMaxQ * kMaxTq) {
            // two quads (24 floats) per thread once a CTA's share allows >= 128 threads that way: fewer warps in the
            // block reductions for the same data in flight
            int q_per_thread = share_q >= 256 ? 2 : 1;
            while ((share_q + q_per_thread - 1) / q_per_thread > kMaxTq) q_per_thread *= 2;
            int threads = ((share_q + q_per_thread - 1) / q_per_thread + 31) / 32 * 32;
            if (threads < 32) threads = 32;
            cudaLaunchConfig_t config = {};
            config.gridDim = dim3(static_cast<unsigned>(B) * ranks_q, 1, 1);
            config.blockDim = dim3(threads, 1, 1);
            config.dynamicSmemBytes = 0;
            config.stream = stream;
            cudaLaunchAttribute attribute[1];
            attribute[0].id = cudaLaunchAttributeClusterDimension;
            attribute[0].val.clusterDim.x = ranks_q;
            attribute[0].val.clusterDim.y = 1;
            attribute[0].val.clusterDim.z = 1;
            config.attrs = attribute;
            config.numAttrs = ranks_q > 1 ? 1 : 0;
            cudaError_t err = cudaSuccess;
#define PS_STATS_QUAD(Q)                                                                                              \
    do {                                                                                                              \
        if (ranks_q > 1)                                                                                              \
            err = is_bool ? cudaLaunchKernelEx(&config, masked_stats_quad_kernel<PS_MASK_BOOL, Q, true>, xyz,        \
                                               atom_mask, atoms, mu, sd, xyz_out)                                    \
                          : cudaLaunchKernelEx(&config, masked_stats_quad_kernel<PS_MASK_F32, Q, true>, xyz,         \
                                               atom_mask, atoms, mu, sd, xyz_out);                                   \
        else                                                                                                          \
            err = is_bool ? cudaLaunchKernelEx(&config, masked_stats_quad_kernel<PS_MASK_BOOL, Q, false>, xyz,       \
                                               atom_mask, atoms, mu, sd, xyz_out)                                    \
                          : cudaLaunchKernelEx(&config, masked_stats_quad_kernel<PS_MASK_F32, Q, false>, xyz,        \
                                               atom_mask, atoms, mu, sd, xyz_out);                                   \
    } while (0)
            // Waves: a CTA's life is a serial chain (loads -> two block reductions -> normalise -> stores), so a launch
            // that needs a second wave of CTAs pays that chain twice (config 4: 1024 CTAs of 256 threads at 62
            // registers = 1.73 waves of 592; 256 x 7680 atoms: 480 threads at 127 registers = one CTA per SM, 1.73
            // waves of 148).  Two more builds of the 4-quads-per-thread kernel trade registers for residency — 72
            // registers (7 CTAs of <= 128 threads per SM) and 64 registers (2 CTAs of <= 512 threads, a few spills) —
            // and the one with the fewest waves is taken (ties: the fewest spills).
            int dense = 0;  // 0: the choice above, 1: 72-register build, 2: 64-register build
            // (0 / 1 masks only: the fp32-mask path keeps four more values per group and spills at these budgets)
            if (ranks_q == 1 && legacy != 3 && share_q > 128 && is_bool) {
                auto waves = [&](int regs, int t) {
                    int per_sm = 65536 / (regs * t);
                    if (per_sm > 2048 / t) per_sm = 2048 / t;
                    if (per_sm > 32) per_sm = 32;
                    if (per_sm < 1) per_sm = 1;
                    const long long slots = static_cast<long long>(per_sm) * sms;
                    return (B + slots - 1) / slots;
                };
                const int threads4 = ((share_q + 3) / 4 + 31) / 32 * 32;
                const int regs_now = q_per_thread == 1 ? 57 : (q_per_thread == 2 ? 64 : 127);
                long long best = waves(regs_now, threads);
                if (threads4 <= 128 && waves(72, threads4) < best) {
                    best = waves(72, threads4);
                    dense = 1;
                }
                if (threads4 <= 512 && waves(64, threads4) < best) dense = 2;
                if (dense) config.blockDim = dim3(threads4, 1, 1);
            }
            if (dense == 1) {
                err = cudaLaunchKernelEx(&config, masked_stats_quad_kernel<PS_MASK_BOOL, 4, false, 128, 7>, xyz, atom_mask,
                                         atoms, mu, sd, xyz_out);
            } else if (dense == 2) {
                err = cudaLaunchKernelEx(&config, masked_stats_quad_kernel<PS_MASK_BOOL, 4, false, 512, 2>, xyz, atom_mask,
                                         atoms, mu, sd, xyz_out);
            } else if (q_per_thread == 1) PS_STATS_QUAD(1);
            else if (q_per_thread == 2) PS_STATS_QUAD(2);
            else PS_STATS_QUAD(4);
#undef PS_STATS_QUAD
            if (err != cudaSuccess) return cuda_fail(err, "cudaLaunchKernelEx(masked_stats_quad_kernel)");
            return check_launch("masked_stats_quad_kernel");
        }
    }

    // ---- register-resident kernel: one cluster of `ranks` CTAs per structure, each thread holds E floats.
    // ranks: split a structure over a thread-block cluster until the grid holds >= 2 CTAs per SM (few structures)
    // or a CTA's share fits in registers (large structures); a CTA keeps >= 128 atoms.
    constexpr int kMaxE = 32, kMaxT = 480;
    int ranks = 1;
    while (ranks < 8 && atoms / (ranks * 2) >= 128 &&
           (static_cast<long long>(B) * ranks < 2ll * sms || (atoms + ranks - 1) / ranks * 3 > kMaxE * kMaxT))
        ranks *= 2;
    const int share_floats = (atoms + ranks - 1) / ranks * 3;
    if (share_floats <= kMaxE * kMaxT && legacy != 1) {
        // threads: a multiple of 96 (whole warps, every thread on one axis), ~8 floats per thread for small shares
        int threads = (share_floats / 8 + 95) / 96 * 96;
        if (threads < 96) threads = 96;
        if (threads > kMaxT) threads = kMaxT;
        const int need = (share_floats + threads - 1) / threads;
        cudaLaunchConfig_t config = {};
        config.gridDim = dim3(static_cast<unsigned>(B) * ranks, 1, 1);
        config.blockDim = dim3(threads, 1, 1);
        config.dynamicSmemBytes = 0;
        config.stream = stream;
        cudaLaunchAttribute attribute[1];
        attribute[0].id = cudaLaunchAttributeClusterDimension;
        attribute[0].val.clusterDim.x = ranks;
        attribute[0].val.clusterDim.y = 1;
        attribute[0].val.clusterDim.z = 1;
        config.attrs = attribute;
        config.numAttrs = ranks > 1 ? 1 : 0;
        cudaError_t err = cudaSuccess;
#define PS_STATS_REGS(E)                                                                                              \
    do {                                                                                                              \
        if (ranks > 1)                                                                                                \
            err = is_bool ? cudaLaunchKernelEx(&config, masked_stats_regs_kernel<PS_MASK_BOOL, E, true>, xyz,        \
                                               atom_mask, atoms, mu, sd, xyz_out)                                    \
                          : cudaLaunchKernelEx(&config, masked_stats_regs_kernel<PS_MASK_F32, E, true>, xyz,         \
                                               atom_mask, atoms, mu, sd, xyz_out);                                   \
        else                                                                                                          \
            err = is_bool ? cudaLaunchKernelEx(&config, masked_stats_regs_kernel<PS_MASK_BOOL, E, false>, xyz,       \
                                               atom_mask, atoms, mu, sd, xyz_out)                                    \
                          : cudaLaunchKernelEx(&config, masked_stats_regs_kernel<PS_MASK_F32, E, false>, xyz,        \
                                               atom_mask, atoms, mu, sd, xyz_out);                                   \
    } while (0)
        if (need <= 4) PS_STATS_REGS(4);
        else if (need <= 8) PS_STATS_REGS(8);
        else if (need <= 12) PS_STATS_REGS(12);
        else if (need <= 16) PS_STATS_REGS(16);
        else if (need <= 24) PS_STATS_REGS(24);
        else PS_STATS_REGS(32);
#undef PS_STATS_REGS
        if (err != cudaSuccess) return cuda_fail(err, "cudaLaunchKernelEx(masked_stats_regs_kernel)");
        return check_launch("masked_stats_regs_kernel");
    }

    // ---- structures too large for registers even as a cluster of 8 (> 40,960 atoms): the three-pass kernel
    ranks = 1;
    while (ranks < 8 && static_cast<long long>(B) * ranks * 2 <= sms && atoms / (ranks * 2) >= 1024) ranks *= 2;
    const int share = (atoms + ranks - 1) / ranks;
    // ~8 atoms per thread (two unrolled iterations); many CTAs stay resident per SM for small structures
    // (a multiple of 96 = 3 x 32: whole warps, and every thread of the normalisation pass stays on one axis)
    int threads = ((share + 7) / 8 + 95) / 96 * 96;
    if (threads < 96) threads = 96;
    if (threads > 480) threads = 480;

    if (ranks == 1) {  // no cluster: the ordinary launch path
        if (mask_dtype == PS_MASK_BOOL)
            masked_stats_kernel<PS_MASK_BOOL, false><<<B, threads, 0, stream>>>(xyz, atom_mask, atoms, mu, sd, xyz_out);
        else
            masked_stats_kernel<PS_MASK_F32, false><<<B, threads, 0, stream>>>(xyz, atom_mask, atoms, mu, sd, xyz_out);
        return check_launch("masked_stats_kernel");
    }
    cudaLaunchConfig_t config = {};
    config.gridDim = dim3(static_cast<unsigned>(B) * ranks, 1, 1);
    config.blockDim = dim3(threads, 1, 1);
    config.dynamicSmemBytes = 0;
    config.stream = stream;
    cudaLaunchAttribute attribute[1];
    attribute[0].id = cudaLaunchAttributeClusterDimension;
    attribute[0].val.clusterDim.x = ranks;
    attribute[0].val.clusterDim.y = 1;
    attribute[0].val.clusterDim.z = 1;
    config.attrs = attribute;
    config.numAttrs = 1;
    cudaError_t err;
    if (mask_dtype == PS_MASK_BOOL)
        err = cudaLaunchKernelEx(&config, masked_stats_kernel<PS_MASK_BOOL, true>, xyz, atom_mask, atoms, mu, sd, xyz_out);
    else
        err = cudaLaunchKernelEx(&config, masked_stats_kernel<PS_MASK_F32, true>, xyz, atom_mask, atoms, mu, sd, xyz_out);
    if (err != cudaSuccess) return cuda_fail(err, "cudaLaunchKernelEx(masked_stats_kernel)");
    return check_launch("masked_stats_kernel");
}

int masked_stats_impl(const float* xyz, const void* atom_mask, int mask_dtype, int B, int L, int A,
                      float* mu, float* sd, float* xyz_out, cudaStream_t stream) {
    return masked_stats_variant_impl(xyz, atom_mask, mask_dtype, B, L, A, mu, sd, xyz_out, 0, stream);
}

int scale_shift_impl(const float* xyz, const float* scale, const float* shift, int B, int L, int A,
                     float* xyz_out, cudaStream_t stream) {
    PS_REQUIRE(B > 0 && L > 0 && A > 0, PS_ERR_BAD_SHAPE, "scale_shift: B=%d L=%d A=%d must be > 0",
               B, L, A);
    PS_REQUIRE(xyz && scale && shift && xyz_out, PS_ERR_NULL_POINTER, "scale_shift: NULL pointer");
    PS_REQUIRE(static_cast<long long>(L) * A * 3 < (1ll << 31), PS_ERR_BAD_SHAPE,
               "scale_shift: L*A*3=%lld floats per structure exceed 2^31", static_cast<long long>(L) * A * 3);
    const int per_b = L * A * 3;
    if (per_b % 4 == 0 && ((reinterpret_cast<uintptr_t>(xyz) | reinterpret_cast<uintptr_t>(xyz_out)) & 15u) == 0) {
        axis_map_vec_kernel<false><<<per_structure_grid_vec(per_b / 4, B), 192, 0, stream>>>(xyz, scale, shift, B, per_b / 4, B,
                                                                                            xyz_out);
        return check_launch("axis_map_vec_kernel");
    }
    scale_shift_kernel<<<per_structure_grid(per_b, B), 256, 0, stream>>>(xyz, scale, shift, per_b, B, xyz_out);
    return check_launch("scale_shift_kernel");
}

int translate_impl(const float* xyz, const float* t, int t_rows, int B, int L, int A,
                   float* xyz_out, cudaStream_t stream) {
    PS_REQUIRE(B > 0 && L > 0 && A > 0, PS_ERR_BAD_SHAPE, "translate: B=%d L=%d A=%d must be > 0", B,
               L, A);
    PS_REQUIRE(xyz && t && xyz_out, PS_ERR_NULL_POINTER, "translate: NULL pointer");
    PS_REQUIRE(t_rows == 1 || t_rows == B, PS_ERR_BAD_SHAPE, "translate: t_rows=%d must be 1 or B=%d",
               t_rows, B);
    PS_REQUIRE(static_cast<long long>(L) * A * 3 < (1ll << 31), PS_ERR_BAD_SHAPE,
               "translate: L*A*3=%lld floats per structure exceed 2^31", static_cast<long long>(L) * A * 3);
    const int per_b = L * A * 3;
    if (per_b % 4 == 0 && ((reinterpret_cast<uintptr_t>(xyz) | reinterpret_cast<uintptr_t>(xyz_out)) & 15u) == 0) {
        axis_map_vec_kernel<true><<<per_structure_grid_vec(per_b / 4, B), 192, 0, stream>>>(xyz, nullptr, t, t_rows, per_b / 4, B,
                                                                                           xyz_out);
        return check_launch("axis_map_vec_kernel");
    }
    translate_kernel<<<per_structure_grid(per_b, B), 256, 0, stream>>>(xyz, t, t_rows, per_b, B, xyz_out);
    return check_launch("translate_kernel");
}

int center_of_mass_impl(const float* xyz, int B, int L, int A, int slot, float* out,
                        cudaStream_t stream) {
    PS_REQUIRE(B > 0 && L > 0 && A > 0, PS_ERR_BAD_SHAPE,
               "center_of_mass: B=%d L=%d A=%d must be > 0", B, L, A);
    PS_REQUIRE(xyz && out, PS_ERR_NULL_POINTER, "center_of_mass: NULL pointer");
    PS_REQUIRE(slot >= 0 && slot < A, PS_ERR_BAD_SLOT, "center_of_mass: slot %d outside [0,%d)",
               slot, A);
    int threads = (L + 31) / 32 * 32;  // one residue per thread up to 256 threads, then several
    if (threads > 256) threads = 256;
    center_of_mass_kernel<<<B, threads, 0, stream>>>(xyz, L, A, slot, out);
    return check_launch("center_of_mass_kernel");
}

}  // namespace ps
