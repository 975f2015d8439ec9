// K5 — forward diffusion of the coordinates.
//
// Replaces StructureBatch.diffuse_xyz (protstruc/protstruc.py:864-878):
//     noise = randn_like(xyz) * sqrt(beta_b);  xyz = sqrt(1 - beta_b) * xyz + noise
// which the reference evaluates as three separately rounded fp32 ops (no FMA); __fmul_rn /
// __fadd_rn below pin that rounding so that, given the same noise tensor, results are bit-equal.
//
// Random numbers: counter-based Philox4x32-10 (Salmon et al., SC'11) + Box-Muller.  Element e of
// the GLOBAL (unsharded) tensor takes output lane (e & 3) of the counter (e >> 2, step); the key is
// the seed.  The stream is therefore a pure function of (seed, step, global element index):
// independent of the launch shape and identical between the single-step kernel and the fused
// multi-step kernel.  A shard may start at ANY global element (elem_offset need not be a multiple
// of 4): thread g of a launch owns global group (elem_offset >> 2) + g, i.e. the local elements
// 4 g - (elem_offset & 3) + k, k = 0..3, so the result never depends on the number of GPUs.
//
// Roofline: single step = HBM/L2 read + write of 4 B per element (the noise is never
// materialised); the 300-step schedule of BASELINE config 4 (23.6 MB of state) is launch/L2-bound
// when run as 300 launches, so ps_diffuse_steps keeps the state in registers for all T steps and
// touches HBM once (ALU-bound on Philox + log/sincos).

#include "common.cuh"

namespace ps {

namespace {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

struct U4 {
    uint32_t x, y, z, w;
};

// The ten round keys depend on the seed only; the kernels expand them once per thread (twenty
// registers) instead of bumping them inside every call.
struct PhiloxKeys {
    uint32_t k0[10], k1[10];
};

__device__ __forceinline__ PhiloxKeys philox_keys(uint64_t seed) {
    PhiloxKeys k;
    uint32_t a = static_cast<uint32_t>(seed), b = static_cast<uint32_t>(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        k.k0[r] = a;
        k.k1[r] = b;
        // opaque to the optimiser, which would otherwise re-derive every key from the seed inside the step loop
        asm volatile("" : "+r"(k.k0[r]), "+r"(k.k1[r]));
        a += kPhiloxW0;
        b += kPhiloxW1;
    }
    return k;
}

__device__ __forceinline__ U4 philox4x32_10(U4 c, const PhiloxKeys& k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(kPhiloxM0, c.x), lo0 = kPhiloxM0 * c.x;
        const uint32_t hi1 = __umulhi(kPhiloxM1, c.z), lo1 = kPhiloxM1 * c.z;
        c = U4{hi1 ^ c.y ^ k.k0[r], lo1, hi0 ^ c.w ^ k.k1[r], lo0};
    }
    return c;
}

// Same function with the round keys bumped on the fly (uniform-datapath adds): cheaper when a thread makes a
// single call, as in the one-step kernel.
__device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(kPhiloxM0, c.x), lo0 = kPhiloxM0 * c.x;
        const uint32_t hi1 = __umulhi(kPhiloxM1, c.z), lo1 = kPhiloxM1 * c.z;
        c = U4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
        k0 += kPhiloxW0;
        k1 += kPhiloxW1;
    }
    return c;
}

// Two random words -> two N(0,1) (Box-Muller).  The generator is the bottleneck of the fused T-step kernel and
// the quarter-rate XU pipe (MUFU, I2F) its scarcest resource, so
//  * the uniforms are built without an int->float conversion: the top 23 bits of the word become the mantissa of
//    a float in [1, 2); u = f - (1 - 2^-24) = (k + 1/2) * 2^-23 in (0, 1), exact in fp32 (LOP3/SHF + one FADD);
//  * the transcendental steps are the single-MUFU intrinsics: MUFU.LG2 (lg2.approx, abs error 2^-22 on [0.5, 2]), MUFU.SQRT,
//    and __sincosf on the angle shifted into (-pi, pi) (abs error 2^-21.4 there); sin(t - pi) = -sin t, so the
//    stream definition z = (r sin 2*pi*u2, r cos 2*pi*u2) is unchanged.
// Deviation from an fp64 evaluation of the same stream is < 1e-5 for 99.9 % of the samples
// (tests/test_gpu_parity.py).  23-bit uniforms bound |z| by sqrt(-2 ln 2^-24) = 5.77.
__device__ __forceinline__ float unit_open(uint32_t word) {
    return __fsub_rn(__uint_as_float((word >> 9) | 0x3F800000u), 0.99999994f);
}

__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    const float u1 = unit_open(a);
    const float u2 = unit_open(b);
    // -2 ln u1 = (-2 ln 2) log2 u1: u1 >= 2^-24 is never denormal, so the raw MUFU.LG2 needs none of __logf's
    // range fix-ups (FSETP + two predicated ops per call), and the two constants fold into one FMUL
    float l, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u1));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l * -1.38629436f));
    float s, c;
    __sincosf(fmaf(u2, 6.28318548f, -3.14159274f), &s, &c);
    return make_float2(-r * s, -r * c);
}

__device__ __forceinline__ void normal4(uint64_t group, uint64_t step, const PhiloxKeys& keys, float (&z)[4]) {
    const U4 ctr{static_cast<uint32_t>(group), static_cast<uint32_t>(group >> 32),
                 static_cast<uint32_t>(step), static_cast<uint32_t>(step >> 32)};
    const U4 r = philox4x32_10(ctr, keys);
    const float2 a = box_muller(r.x, r.y);
    const float2 b = box_muller(r.z, r.w);
    z[0] = a.x;
    z[1] = a.y;
    z[2] = b.x;
    z[3] = b.y;
}

__device__ __forceinline__ void normal4(uint64_t group, uint64_t step, uint64_t seed, float (&z)[4]) {
    const U4 ctr{static_cast<uint32_t>(group), static_cast<uint32_t>(group >> 32),
                 static_cast<uint32_t>(step), static_cast<uint32_t>(step >> 32)};
    const U4 r = philox4x32_10(ctr, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    const float2 a = box_muller(r.x, r.y);
    const float2 b = box_muller(r.z, r.w);
    z[0] = a.x;
    z[1] = a.y;
    z[2] = b.x;
    z[3] = b.y;
}

__device__ __forceinline__ float diffuse_one(float x, float z, float sa, float sb) {
    return __fadd_rn(__fmul_rn(sa, x), __fmul_rn(z, sb));
}

// Injected-noise variant: purely elementwise.  2-D grid: blockIdx.y walks the structures (one beta, two IEEE
// square roots per thread and structure), x the structure's floats, fully coalesced.
__global__ void __launch_bounds__(256) diffuse_noise_kernel(const float* __restrict__ x,
                                                            const float* __restrict__ beta,
                                                            const float* __restrict__ noise,
                                                            long long per_b, int B,
                                                            float* __restrict__ out) {
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const float bt = __ldg(beta + b);
        const float sa = __fsqrt_rn(__fsub_rn(1.0f, bt));
        const float sb = __fsqrt_rn(bt);
        const long long base = b * per_b;
        const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
        for (long long e0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e0 < per_b; e0 += 4 * step) {
            float v[4], z[4];  // four independent load pairs in flight per thread
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long e = e0 + u * step;
                v[u] = e < per_b ? x[base + e] : 0.f;
                z[u] = e < per_b ? __ldg(noise + base + e) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (e0 + u * step < per_b) out[base + e0 + u * step] = diffuse_one(v[u], z[u], sa, sb);
        }
    }
}

// Philox variant: one thread per group of 4 consecutive elements, T steps in registers.  betas is (T, B).
// A CTA takes blocks of 256 consecutive groups (1024 elements); such a block touches at most kTableSlots
// structures when per_b >= 1024 / (kTableSlots - 1), and then the CTA first tabulates (sqrt(1 - beta), sqrt(beta))
// of those structures for all T steps in shared memory — a few IEEE square roots per thread instead of 2 T —
// and the step loop reads them with one broadcast LDS.64.  Smaller structures take the roots in the loop.
constexpr int kTableSlots = 2;

__global__ void __launch_bounds__(256) diffuse_philox_kernel(const float* __restrict__ x,
                                                             const float* __restrict__ betas, int T,
                                                             int B, uint64_t seed, uint64_t step0,
                                                             uint64_t group_offset, int shift, long long per_b,
                                                             long long total, int use_table,
                                                             float* __restrict__ out) {
    extern __shared__ float2 schedule[];  // [T][kTableSlots] when use_table
    const PhiloxKeys keys = philox_keys(seed);
    const long long groups = (total + shift + 3) / 4;
    const long long num_blocks = (groups + blockDim.x - 1) / blockDim.x;
    for (long long blk = blockIdx.x; blk < num_blocks; blk += gridDim.x) {
        const long long g = blk * blockDim.x + threadIdx.x;
        int b_lo = 0, nb = 0;
        if (use_table) {  // structures this block of 1024 elements touches
            long long block_e0 = blk * blockDim.x * 4 - shift;
            long long block_e1 = block_e0 + static_cast<long long>(blockDim.x) * 4 - 1;
            if (block_e0 < 0) block_e0 = 0;
            if (block_e1 > total - 1) block_e1 = total - 1;
            b_lo = static_cast<int>(block_e0 / per_b);
            nb = static_cast<int>(block_e1 / per_b) - b_lo + 1;
        }
        const bool table = use_table && nb <= kTableSlots;  // block-uniform
        if (table) {
            __syncthreads();  // the previous block's readers are done
            for (int idx = threadIdx.x; idx < T * nb; idx += blockDim.x) {
                const int t = idx / nb, slot = idx - t * nb;
                const float bt = __ldg(betas + static_cast<long long>(t) * B + b_lo + slot);
                schedule[t * kTableSlots + slot] = make_float2(__fsqrt_rn(__fsub_rn(1.0f, bt)), __fsqrt_rn(bt));
            }
            __syncthreads();
        }
        if (g >= groups) continue;
        const long long e0 = g * 4 - shift;  // first local element of the group (negative only for g = 0)
        const long long e0c = e0 < 0 ? 0 : e0;
        float v[4];
        int bidx[4];
        // structure of each of the four elements: ONE division per thread (32-bit whenever the element count
        // allows), then the group either stays inside the structure or steps over its end
        long long b0, rem0;
        if (total <= 0xFFFFFFFFll && per_b <= 0xFFFFFFFFll) {
            const unsigned q = static_cast<unsigned>(e0c) / static_cast<unsigned>(per_b);
            b0 = q;
            rem0 = static_cast<unsigned>(e0c) - q * static_cast<unsigned>(per_b);
        } else {
            b0 = e0c / per_b;
            rem0 = e0c - b0 * per_b;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long e = e0 + k;
            const bool ok = e >= 0 && e < total;
            v[k] = ok ? x[e] : 0.f;
            long long bk = b0 + (rem0 + (e - e0c) >= per_b ? 1 : 0);
            if (per_b < 4 && ok) bk = e / per_b;  // degenerate structures of fewer than four floats
            bidx[k] = ok ? static_cast<int>(bk) : static_cast<int>(b0);
        }
        const bool same_b = bidx[0] == bidx[3];
        for (int t = 0; t < T; ++t) {
            float z[4];
            normal4(static_cast<uint64_t>(g) + group_offset, step0 + static_cast<uint64_t>(t), keys, z);
            const float* __restrict__ bt_row = betas + static_cast<long long>(t) * B;
            if (table && same_b) {
                const float2 ab = schedule[t * kTableSlots + (bidx[0] - b_lo)];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = diffuse_one(v[k], z[k], ab.x, ab.y);
            } else if (table) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 ab = schedule[t * kTableSlots + (bidx[k] - b_lo)];
                    v[k] = diffuse_one(v[k], z[k], ab.x, ab.y);
                }
            } else if (same_b) {
                const float bt = __ldg(bt_row + bidx[0]);
                const float sa = __fsqrt_rn(__fsub_rn(1.0f, bt));
                const float sb = __fsqrt_rn(bt);
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = diffuse_one(v[k], z[k], sa, sb);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float bt = __ldg(bt_row + bidx[k]);
                    v[k] = diffuse_one(v[k], z[k], __fsqrt_rn(__fsub_rn(1.0f, bt)), __fsqrt_rn(bt));
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (e0 + k >= 0 && e0 + k < total) out[e0 + k] = v[k];
    }
}

// One step (what a caller of the reference's diffuse_xyz does T times): the lean form of the kernel above — one
// group of four elements per thread, no schedule table, no key expansion; the same stream and the same arithmetic,
// so T calls of it equal one T-step launch bit for bit.
__global__ void __launch_bounds__(256) diffuse_philox_step_kernel(const float* __restrict__ x,
                                                                  const float* __restrict__ beta, uint64_t seed,
                                                                  uint64_t step, uint64_t group_offset, int shift,
                                                                  long long per_b, long long total,
                                                                  float* __restrict__ out) {
    const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long e0 = g * 4 - shift;
    if (e0 >= total) return;
    float z[4];
    normal4(static_cast<uint64_t>(g) + group_offset, step, seed, z);
    if (e0 < 0) {  // the first group of a shard that starts inside a group of four
        for (int k = static_cast<int>(-e0); k < 4 && e0 + k < total; ++k) {
            const float bt = __ldg(beta + (e0 + k) / per_b);
            out[e0 + k] = diffuse_one(x[e0 + k], z[k], __fsqrt_rn(__fsub_rn(1.0f, bt)), __fsqrt_rn(bt));
        }
        return;
    }
    const bool fits_32_bits = total <= 0xFFFFFFFFll && per_b <= 0xFFFFFFFFll;
    const long long b0 = index_div(e0, per_b, fits_32_bits);
    const long long rem0 = e0 - b0 * per_b;
    if (rem0 + 3 < per_b && e0 + 3 < total) {  // the whole group lies in structure b0 (the common case)
        const float bt = __ldg(beta + b0);
        const float sa = __fsqrt_rn(__fsub_rn(1.0f, bt));
        const float sb = __fsqrt_rn(bt);
#pragma unroll
        for (int k = 0; k < 4; ++k) out[e0 + k] = diffuse_one(x[e0 + k], z[k], sa, sb);
    } else {
        for (int k = 0; k < 4 && e0 + k < total; ++k) {
            const float bt = __ldg(beta + (e0 + k) / per_b);
            out[e0 + k] = diffuse_one(x[e0 + k], z[k], __fsqrt_rn(__fsub_rn(1.0f, bt)), __fsqrt_rn(bt));
        }
    }
}

// The same step for structures of a multiple of four floats on 16-byte aligned arrays whose shard starts on a group
// boundary (the usual case: config 4 has 5,760 floats per structure).  2-D grid — y walks the structures, x the
// structure's groups of four — so a thread takes beta and its two IEEE square roots ONCE per structure instead of
// once per group, needs no division to find its structure, and moves its groups with 128-bit loads / stores.  Same
// counters, same Box-Muller, same rounding: bit-identical to the kernel above (ncu: ~280 instructions per group there,
// of which the generator and the update are ~95).
__global__ void __launch_bounds__(256) diffuse_philox_step_vec_kernel(const float4* __restrict__ x,
                                                                      const float* __restrict__ beta, uint64_t seed,
                                                                      uint64_t step, uint64_t group_offset,
                                                                      unsigned groups_per_b, int B,
                                                                      float4* __restrict__ out) {
    const unsigned stride = gridDim.x * blockDim.x;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const float bt = __ldg(beta + b);
        const float sa = __fsqrt_rn(__fsub_rn(1.0f, bt));
        const float sb = __fsqrt_rn(bt);
        const unsigned long long base = static_cast<unsigned long long>(b) * groups_per_b;
        for (unsigned gl = blockIdx.x * blockDim.x + threadIdx.x; gl < groups_per_b; gl += stride) {
            const float4 v = x[base + gl];
            float z[4];
            normal4(base + gl + group_offset, step, seed, z);
            out[base + gl] = make_float4(diffuse_one(v.x, z[0], sa, sb), diffuse_one(v.y, z[1], sa, sb),
                                         diffuse_one(v.z, z[2], sa, sb), diffuse_one(v.w, z[3], sa, sb));
        }
    }
}

// One diffusion step on the Philox stream: picks the vector kernel when the layout allows it.
int launch_philox_step(const float* x, const float* beta, uint64_t seed, uint64_t step, uint64_t elem_offset, int B,
                       long long per_b, float* out, cudaStream_t stream) {
    const int shift = static_cast<int>(elem_offset & 3u);
    const long long total = per_b * B;
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    if (shift == 0 && (per_b & 3) == 0 && aligned && per_b / 4 < (1ll << 31)) {
        const unsigned groups_per_b = static_cast<unsigned>(per_b / 4);
        // ~2 groups per thread along x keeps the generator busy while the 128-bit loads are in flight
        unsigned gx = (groups_per_b + 511) / 512;
        if (gx < 1) gx = 1;
        if (gx > 64) gx = 64;
        const dim3 grid(gx, static_cast<unsigned>(B < 65535 ? B : 65535), 1);
        diffuse_philox_step_vec_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(x), beta, seed, step,
                                                                 elem_offset / 4, groups_per_b, B,
                                                                 reinterpret_cast<float4*>(out));
        return PS_OK;
    }
    const long long groups = (total + shift + 3) / 4;
    diffuse_philox_step_kernel<<<static_cast<unsigned>((groups + 255) / 256), 256, 0, stream>>>(
        x, beta, seed, step, elem_offset / 4, shift, per_b, total, out);
    return PS_OK;
}

__global__ void __launch_bounds__(256) philox_normal_kernel(float* __restrict__ out, long long n,
                                                            uint64_t seed, uint64_t step,
                                                            uint64_t group_offset, int shift) {
    const PhiloxKeys keys = philox_keys(seed);
    const long long groups = (n + shift + 3) / 4;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; g < groups;
         g += stride) {
        float z[4];
        normal4(static_cast<uint64_t>(g) + group_offset, step, keys, z);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long e = g * 4 - shift + k;
            if (e >= 0 && e < n) out[e] = z[k];
        }
    }
}

int grid_for(long long work_items, int* grid) {
    const int sms = sm_count_for_current_device();
    if (sms < 0) return sms;
    long long g = (work_items + 255) / 256;
    const long long cap = static_cast<long long>(sms) * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    *grid = static_cast<int>(g);
    return PS_OK;
}

}  // namespace

int diffuse_impl(const float* x, const float* betas, int T, const float* noise, uint64_t seed,
                 uint64_t step0, uint64_t elem_offset, float* out, int B, long long per_b,
                 cudaStream_t stream) {
    PS_REQUIRE(B > 0 && per_b > 0 && T > 0, PS_ERR_BAD_SHAPE,
               "diffuse: B=%d per_b=%lld T=%d must be > 0", B, per_b, T);
    PS_REQUIRE(x && betas && out, PS_ERR_NULL_POINTER, "diffuse: NULL pointer");
    const int shift = static_cast<int>(elem_offset & 3u);  // position of the shard's first element in its group
    const long long total = per_b * B;
    int grid = 0;
    if (noise) {
        PS_REQUIRE(T == 1, PS_ERR_BAD_SHAPE, "diffuse: injected noise supports a single step");
        long long gx = (per_b + 1023) / 1024;
        if (gx > 32) gx = 32;
        const dim3 grid2(static_cast<unsigned>(gx), static_cast<unsigned>(B < 65535 ? B : 65535), 1);
        diffuse_noise_kernel<<<grid2, 256, 0, stream>>>(x, betas, noise, per_b, B, out);
        return check_launch("diffuse_noise_kernel");
    }
    if (T == 1) {
        const long long groups = (total + shift + 3) / 4;
        PS_REQUIRE((groups + 255) / 256 < (1ll << 31), PS_ERR_BAD_SHAPE, "diffuse: %lld elements", total);
        launch_philox_step(x, betas, seed, step0, elem_offset, B, per_b, out, stream);
        return check_launch("diffuse_philox_step_kernel");
    }
    int rc = grid_for((total + shift + 3) / 4, &grid);
    if (rc != PS_OK) return rc;
    // schedule table: worth it from a handful of steps on, as long as it fits the default 48 KB of shared memory
    const size_t table_bytes = static_cast<size_t>(T) * kTableSlots * sizeof(float2);
    const int use_table = (T >= 4 && table_bytes <= 40 * 1024 && per_b >= 1024 / (kTableSlots - 1)) ? 1 : 0;
    diffuse_philox_kernel<<<grid, 256, use_table ? table_bytes : 0, stream>>>(x, betas, T, B, seed, step0, elem_offset / 4,
                                                                              shift, per_b, total, use_table, out);
    return check_launch("diffuse_philox_kernel");
}

// T steps as T back-to-back launches of the one-step kernel, each writing its result into its own slice of a
// (T, B, per_b) trajectory buffer (what the reference's tutorial loop collects for its animation:
// `for t in range(T): sb.diffuse_xyz(beta[t]); frames.append(sb.get_xyz())`).  The loop runs HERE, on the host side of
// the C-ABI: ~2 us per launch instead of the ~15 us a Python-level call costs, so the 300-step schedule of BASELINE
// config 4 is bound by the GPU time of the steps again.  Same noise stream as ps_diffuse / ps_diffuse_steps.
int diffuse_trajectory_impl(const float* x, const float* betas, int T, uint64_t seed, uint64_t step0,
                            uint64_t elem_offset, float* trajectory, int B, long long per_b, cudaStream_t stream) {
    PS_REQUIRE(B > 0 && per_b > 0 && T > 0, PS_ERR_BAD_SHAPE,
               "diffuse_trajectory: B=%d per_b=%lld T=%d must be > 0", B, per_b, T);
    PS_REQUIRE(x && betas && trajectory, PS_ERR_NULL_POINTER, "diffuse_trajectory: NULL pointer");
    const long long total = per_b * B;
    const int shift = static_cast<int>(elem_offset & 3u);
    const long long groups = (total + shift + 3) / 4;
    PS_REQUIRE((groups + 255) / 256 < (1ll << 31), PS_ERR_BAD_SHAPE, "diffuse_trajectory: %lld elements", total);
    const float* src = x;
    for (int t = 0; t < T; ++t) {
        float* dst = trajectory + static_cast<long long>(t) * total;
        launch_philox_step(src, betas + static_cast<long long>(t) * B, seed, step0 + static_cast<uint64_t>(t), elem_offset,
                           B, per_b, dst, stream);
        src = dst;
    }
    return check_launch("diffuse_philox_step_kernel");
}

int philox_normal_impl(float* out, long long n, uint64_t seed, uint64_t step, uint64_t elem_offset,
                       cudaStream_t stream) {
    PS_REQUIRE(n >= 0, PS_ERR_BAD_SHAPE, "philox_normal: n=%lld", n);
    if (n == 0) return PS_OK;
    PS_REQUIRE(out, PS_ERR_NULL_POINTER, "philox_normal: out is NULL");
    const int shift = static_cast<int>(elem_offset & 3u);
    int grid = 0;
    int rc = grid_for((n + shift + 3) / 4, &grid);
    if (rc != PS_OK) return rc;
    philox_normal_kernel<<<grid, 256, 0, stream>>>(out, n, seed, step, elem_offset / 4, shift);
    return check_launch("philox_normal_kernel");
}

}  // namespace ps
