// Shared pieces of the staged pair-distance tile kernels (pair_dist.cu: column-strip / lock-step kernel for every staged
// atom count; pair_sweep.cu: the linear-sweep kernel with TMA-staged residue j): output kinds, tile geometry, the
// square-root flavours, the word-wise mask-block writer, the per-tile named barrier.
#pragma once

#include "common.cuh"

namespace ps {

// What the most recent pair-distance launch of this host thread chose (ps_pair_dist_last_plan): tests assert on it
// that a shape took the path — and the tile schedule — its parity claim is about.
struct PairDistPlan {
    long long path = -1;          // 0 staged tile kernel, 1 any-A tile kernel, 2 row kernel
    long long lockstep = 0;       // staged kernel: 1 = linear lock-step sweep, 0 = (chunk, strip) cells
    long long ctas = 0;
    long long tile_buffers = 0;   // tile buffers (workers) of the grid
    long long active_buffers = 0; // buffers that take part
    long long strip_stride = 0;   // tiles between two members of a strip
    long long tile_pairs = 0;
    long long launches = 0;       // kernel launches of the call
    long long sweep = 0;          // 1 = the linear-sweep kernel (pair_sweep.cu), 0 = the column-strip kernel
};
extern thread_local PairDistPlan g_last_plan;  // defined in pair_dist.cu

namespace {

constexpr int kTilePairs = 32;  // pairs per tile per pass over the lanes (tile = 32 * Q pairs)

// Pairs per lane.  With few atoms per residue a 32-pair tile is only a few KB and the per-tile bookkeeping
// dominates, so small residues take several pairs per lane (lane l owns pairs l, l + 32, ... of the tile).
template <int A>
__host__ __device__ constexpr int pairs_per_lane() {
    return A <= 6 ? 4 : (A <= 10 ? 2 : 1);
}
constexpr int kDefaultWarpsPerTile = 2;  // see pair_tiles_kernel; variant bit 9 selects the other value

// Default is the single-instruction MUFU.SQRT (sqrt.approx.ftz.f32): relative error <= 2^-23, sqrt(0) = 0,
// NaN -> NaN, +inf -> +inf.  "ftz" only matters for a sum of squares below 1.18e-38, i.e. distances
// below 1.1e-19 A, which are returned as 0 (well inside the 1e-4 A tolerance); the non-ftz flavour
// costs three more issue slots per element, the IEEE one eight.
enum SqrtMode { kSqrtApproxFtz = 0, kSqrtApprox = 1, kSqrtRn = 2 };

template <int MODE>
__device__ __forceinline__ float sqrt_mode(float v) {
    float r;
    if (MODE == kSqrtApprox) {
        // MUFU.SQRT with subnormal pre-scaling; max relative error 2^-23 (PTX ISA), sqrt(0)=0,
        // NaN -> NaN, +inf -> +inf.
        asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
    } else if (MODE == kSqrtApproxFtz) {
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    } else {
        r = __fsqrt_rn(v);
    }
    return r;
}

// What one launch of the staged kernel produces.
enum OutKind {
    kDistBoolMask = 0,  // distances + 1-byte mask        (the reference's from_pdb path)
    kDistOnly = 1,      // distances only
    kF32MaskOnly = 2,   // fp32 mask product only          (the reference's from_pdb_id path)
    kBoolMaskOnly = 3   // 1-byte mask only
};

template <int A>
struct TileGeom {
    static constexpr int kPairs = kTilePairs * pairs_per_lane<A>();
    static constexpr int kElemsPerPair = A * A;
    static constexpr int kTileElems = kPairs * A * A;
    static constexpr int kDistBytes = kTileElems * 4;
    static constexpr int kMaskBytes = kTileElems;
    static_assert(kDistBytes % 16 == 0 && kMaskBytes % 16 == 0, "tile must be 16-B granular");
};

template <int KIND>
__host__ __device__ constexpr bool kind_has_f32() {
    return KIND == kDistBoolMask || KIND == kDistOnly || KIND == kF32MaskOnly;
}
template <int KIND>
__host__ __device__ constexpr bool kind_has_u8() {
    return KIND == kDistBoolMask || KIND == kBoolMaskOnly;
}
template <int A, int KIND>
__host__ __device__ constexpr int warp_smem_bytes() {
    return (kind_has_f32<KIND>() ? TileGeom<A>::kDistBytes : 0) +
           (kind_has_u8<KIND>() ? TileGeom<A>::kMaskBytes : 0);
}

struct PairDistParams {
    const float* __restrict__ xyz;
    const void* __restrict__ atom_mask;
    float* __restrict__ dist;    // f32 output (distances, or the fp32 mask for kF32MaskOnly)
    uint8_t* __restrict__ mask;  // 1-byte mask output
    float* __restrict__ omega;   // fused angles (may be null)
    float* __restrict__ theta;
    float* __restrict__ phi;
    float* __restrict__ d_ca;    // compact (B, L, L) copies of dist[..., CA, CA], [..., CB, CB], [..., N, O] for the
    float* __restrict__ d_cb;    // optional NVLink gather of compact features (may be null)
    float* __restrict__ d_no;
    int L;
    long long num_rows;   // B*L residues
    long long num_pairs;  // B*L*L
    long long num_tiles;  // ceil(num_pairs / 32)
    // Column-strip schedule: tiles t and t + strip_stride cover the same 32 residues j (one residue i
    // row further down), so a warp that walks t, t + S, t + 2S, ... keeps residue j in registers.
    long long strip_stride;   // S = L / gcd(L, 32) tiles
    long long strip_members;  // M = ceil(num_tiles / S)
    // Strips are cut into chunks of `chunk_members` consecutive members; a (chunk, strip) cell is the
    // unit a worker walks.  Cells are ordered strip-fastest, so workers that run side by side write
    // ADJACENT tiles (same member, neighbouring strips): the concurrent stores of the whole GPU form a
    // few contiguous fronts in HBM instead of ~900 scattered streams.
    long long chunk_members;  // C
    long long num_cells;      // S * ceil(M / C)
    long long active_workers; // tile buffers that take part (<= gridDim.x * buffers per CTA)
    int lockstep;             // strips mapped 1:1 to workers (linear sweep of the output): 0 auto, 1 on, 2 off
    int stores_only;          // diagnostic: skip the arithmetic, only issue the tile stores (measures the
                              // memory-system ceiling of this write pattern; output content is undefined)
    int l2_hint;              // L2 eviction policy of the tile stores: 0 none, 1 evict_first, 2 evict_last
};

// Gather the A mask bytes of one residue into a bit field (bit a = mask[a] != 0).
template <int A>
__device__ __forceinline__ uint32_t load_mask_bits(const uint8_t* __restrict__ m) {
    uint32_t bits = 0;
#pragma unroll
    for (int a = 0; a < A; ++a) bits |= (__ldg(m + a) != 0 ? 1u : 0u) << a;
    return bits;
}

// Writes one lane's A*A mask bytes (byte a*A + c = mask_i[a] & mask_j[c]) into the warp's mask tile
// with 32-bit shared-memory stores only.
//
// The lane's block starts at byte 225*lane of the tile, i.e. at byte s = lane & 3 of an aligned word.
//  1. R = the 15 bytes of row "mask_j" as four words (a 4-bit -> 4-byte spread is one IMAD + one LOP3);
//  2. the masked rows (R & -mask_i[a]) are concatenated into the unshifted block words B[0..56]; 4
//     consecutive block bytes always come from two consecutive words of the masked-row array, so
//     every B word is a single PRMT with a compile-time selector;
//  3. the block is moved to its byte alignment with one funnel shift per word (runtime s);
//  4. the word that straddles two lanes' blocks is completed with the neighbour's first word
//     (one shuffle) and written by the lower lane only.
// Lane stride is 56.25 words: start words floor(56.25*l) hit 32 distinct banks -> conflict free.
// ~1.2 issue slots per mask byte instead of 3 for byte-wise stores (which also bank-conflict).
// Byte-wise mask block for the other atom counts (A != 15): three issue slots per byte and 2-way bank
// conflicts, acceptable for the secondary fast paths.
template <int A>
__device__ __forceinline__ void write_mask_block_bytes(uint8_t* __restrict__ tile_bytes, int lane,
                                                       uint32_t mi_bits, uint32_t mj_bits) {
    uint8_t* dst = tile_bytes + lane * (A * A);
#pragma unroll
    for (int a = 0; a < A; ++a) {
        const uint32_t row = ((mi_bits >> a) & 1u) ? mj_bits : 0u;
#pragma unroll
        for (int c = 0; c < A; ++c) dst[a * A + c] = static_cast<uint8_t>((row >> c) & 1u);
    }
}

// `r` = the 15 bytes of row "mask_j" as four words (byte 15 zero): from a bit field here, straight from staged mask
// bytes in the linear-sweep kernel.
template <int A>
__device__ __forceinline__ void write_mask_block_rows(uint32_t* __restrict__ tile_words, int lane,
                                                      uint32_t mi_bits, const uint32_t (&r)[4]) {
    static_assert(A == 15, "the word-wise mask writer is laid out for 15 atoms per residue");
    constexpr int kWordsPerRow = 4;              // 15 bytes + 1 pad byte
    constexpr int kBlockWords = (A * A + 3) / 4;  // 57
    uint32_t rm[A * kWordsPerRow + 1];
#pragma unroll
    for (int a = 0; a < A; ++a) {
        const uint32_t keep = 0u - ((mi_bits >> a) & 1u);
#pragma unroll
        for (int w = 0; w < kWordsPerRow; ++w) rm[a * kWordsPerRow + w] = r[w] & keep;
    }
    rm[A * kWordsPerRow] = 0u;
    uint32_t blk[kBlockWords];
#pragma unroll
    for (int k = 0; k < kBlockWords; ++k) {
        // source position of block byte t: row t / 15, column t % 15 -> word 4*row + col/4, byte col%4
        const int t0 = 4 * k;
        const int s0 = (t0 / A) * kWordsPerRow + ((t0 % A) >> 2);
        uint32_t sel = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int t = t0 + q;
            if (t < A * A) {
                const int sw = (t / A) * kWordsPerRow + ((t % A) >> 2);
                sel |= static_cast<uint32_t>((sw - s0) * 4 + ((t % A) & 3)) << (4 * q);
            } else {
                sel |= 7u << (4 * q);  // byte 3 of the zero word rm[60]
            }
        }
        blk[k] = __byte_perm(rm[s0], rm[s0 + 1], sel);
    }
    const int shift_bits = (lane & 3) * 8;
    uint32_t out[kBlockWords];
    out[0] = blk[0] << shift_bits;
#pragma unroll
    for (int k = 1; k < kBlockWords; ++k) out[k] = __funnelshift_l(blk[k - 1], blk[k], shift_bits);
    // the last word of this lane is the first word of lane + 1 unless that lane starts word-aligned
    const uint32_t neighbour_first = __shfl_down_sync(0xffffffffu, out[0], 1);
    if (((lane + 1) & 3) != 0 && lane < 31) out[kBlockWords - 1] |= neighbour_first;
    uint32_t* dst = tile_words + ((A * A * lane) >> 2);
    if ((lane & 3) == 0) dst[0] = out[0];
#pragma unroll
    for (int k = 1; k < kBlockWords; ++k) dst[k] = out[k];
}

template <int A>
__device__ __forceinline__ void write_mask_block(uint32_t* __restrict__ tile_words, int lane,
                                                 uint32_t mi_bits, uint32_t mj_bits) {
    uint32_t r[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) r[w] = (((mj_bits >> (4 * w)) & 0xFu) * 0x00204081u) & 0x01010101u;
    write_mask_block_rows<A>(tile_words, lane, mi_bits, r);
}

// Synchronises the WPT warps that share one tile (named barrier `slot + 1`; plain __syncwarp for WPT = 1).
template <int WPT>
__device__ __forceinline__ void tile_sync(int slot) {
    if (WPT == 1) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" :: "r"(slot + 1), "n"(WPT * 32) : "memory");
    }
}

// Staging area of one warp: two residues of A*3 floats, rounded up to whole 32-float rows, double buffered.
template <int A>
__host__ __device__ constexpr int stage_floats() {
    return (2 * A * 3 + 31) / 32 * 32;
}
template <int A>
__host__ __device__ constexpr int stage_bytes_per_warp() {
    return 2 * stage_floats<A>() * 4;
}

}  // namespace

}  // namespace ps
