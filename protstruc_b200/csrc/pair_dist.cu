// K1 — all-atom pairwise distance matrix + fused pair mask (+ optionally the trRosetta angles).
//
// Replaces StructureBatch.pairwise_distance_matrix (protstruc/protstruc.py:455-484) and, in its
// fused form, StructureBatch.inter_residue_geometry (protstruc/protstruc.py:790-817).
//
// Roofline: store-bound.  Per residue pair (b,i,j) the kernel writes A*A fp32 distances plus A*A
// mask bytes (1125 B at A = 15) and reads ~2 residues of coordinates from L1/L2; there is no
// reuse to exploit and no GEMM shape (K = 3), so the design goal is: as few issued instructions
// per output element as possible and perfectly coalesced, asynchronous stores.
//
// Data layout in HBM (row-major, contiguous):
//   xyz        (B, L, A, 3)      fp32
//   atom_mask  (B, L, A)         bool (1 B) or fp32
//   dist       (B, L, L, A, A)   fp32      -> a flat array of P = B*L*L pair blocks of A*A floats
//   dist_mask  (B, L, L, A, A)   same dtype as atom_mask
//
// Staged kernel (A = 15, the reference's MAX_N_ATOMS_PER_RESIDUE; also built for A = 5, 10, 14):
//   * the pair blocks are a flat list; a TILE is 32 consecutive pair blocks = 7200 elements
//     = 28,800 B of distances + 7,200 B of mask, both multiples of 16 B, so every tile starts
//     16-byte aligned whatever L is (no head/tail peeling for odd L);
//   * persistent grid, one CTA per SM, 4 tile buffers per CTA, two warps per buffer; lane = pair.
//     The lane keeps the 15 atoms of residue j in registers (packed as f32x2 so the
//     FADD2/FMUL2/FFMA2 pipe does two atoms per instruction) and writes its 225 distances into the
//     shared-memory tile at lane*225 + a*15 + c — a stride of 225 words between lanes, which is
//     1 mod 32, hence bank-conflict free;
//   * column-strip schedule: a buffer walks tiles t, t+S, t+2S, ... (S = L / gcd(L, 32)), which all
//     cover the same 32 residues j, so residue j is loaded once per ~150 tiles;
//   * residue i (new for every tile) is fetched one tile ahead with coalesced loads into a small
//     per-warp staging area and read with broadcast LDS; its mask bits come from a ballot;
//   * the mask block is generated word-wise (PRMT + funnel shift), conflict-free 32-bit STS;
//   * the finished tile leaves through the TMA engine: one elected lane issues
//     cp.async.bulk.global.shared::cta (SASS: UBLKCP) for the distance tile and one for the mask
//     tile.  Address generation and coalescing cost no issue slots, stores are 100 % full-line;
//   * with ANGLES the lane also evaluates omega / theta / phi of its pair (inter_residue_geometry
//     becomes one launch).
// Measured: 6.2-6.8 TB/s, the store ceiling of the memory system for non-uniform data.
// Generic kernel (any other A, L < 32, misaligned outputs): one thread per output element, coalesced
// scalar stores.  Slow but shape-agnostic.

#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "pair_tiles.cuh"

namespace ps {

thread_local PairDistPlan g_last_plan;

namespace {

// WPT = warps per tile.  WPT = 1: one warp computes the whole tile.  WPT = 2: two warps share the tile
// buffer — warp 0 takes the first rows and the angle triple, warp 1 the remaining rows and the mask block — which
// doubles the resident warps (12 per SM) for the same shared-memory footprint; the extra thread-level
// parallelism hides the fixed-latency dependency stalls that dominate with 6 warps per SM.
//
// Residue i (the row side of the pair) changes with every tile, so its 45 coordinates + 15 mask
// bytes are a compulsory L2 round trip per tile.  They are therefore fetched ONE TILE AHEAD: while tile
// k is computed, three coalesced loads per lane bring the two residues tile k+1 can touch (a tile of 32
// pairs spans at most two residue-i rows when L >= 32) into registers; they are parked in a small
// double-buffered per-warp staging area and the row loop reads them with broadcast LDS.  The mask bits
// of both residues come from one byte load per lane and a ballot.
template <int A, int KIND, int SQRT, bool ANGLES, int WPT>
__global__ void __launch_bounds__(WPT * 256 > 384 ? (A <= 6 ? 512 : 384) : WPT * 256, 1)
pair_tiles_kernel(const PairDistParams p) {
    using G = TileGeom<A>;
    constexpr int Q = pairs_per_lane<A>();
    static_assert(A >= 1 && 2 * A <= 32, "the mask ballot holds two residues of at most 16 atoms");
    constexpr int kStageFloats = stage_floats<A>();
    constexpr int kStageRows = kStageFloats / 32;
    constexpr int kStageBytesPerWarp = stage_bytes_per_warp<A>();
    constexpr int NP = (A + 1) / 2;  // f32x2 packs per coordinate
    // Row split between the two warps of a tile, balanced against their extra duties: the angle triple
    // (warp 0, ~300 issue slots per tile) and the mask block (warp 1, ~255) versus ~80 per row.
    constexpr int kSplitRow = (WPT == 1) ? A : (ANGLES ? (A - 1) / 2 : (A * 3) / 5);
    constexpr bool kNeedsXyz = (KIND == kDistBoolMask || KIND == kDistOnly);
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int slot = warp / WPT;   // tile buffer of this warp
    const int wsub = warp % WPT;   // role inside the tile
    const int slots_per_cta = (blockDim.x >> 5) / WPT;
    const bool does_rows_lo = (wsub == 0);
    const bool does_mask = kind_has_u8<KIND>() && (wsub == WPT - 1);
    const bool does_angles = ANGLES && (wsub == 0);
    const bool is_issuer = (wsub == 0) && (lane == 0);
    const int row_begin = does_rows_lo ? 0 : kSplitRow;
    const int row_end = (WPT == 1 || !does_rows_lo) ? A : kSplitRow;

    unsigned char* wbase = smem_raw + static_cast<size_t>(slot) * warp_smem_bytes<A, KIND>();
    float* tile_f32 = reinterpret_cast<float*>(wbase);
    uint8_t* tile_u8 = wbase + (kind_has_f32<KIND>() ? G::kDistBytes : 0);
    float* stage = reinterpret_cast<float*>(smem_raw + static_cast<size_t>(slots_per_cta) * warp_smem_bytes<A, KIND>() +
                                            static_cast<size_t>(warp) * kStageBytesPerWarp);

    // Work partition: the linear order u = (cell * C + step) over (chunk, strip) cells is cut into equal
    // contiguous ranges, one per tile buffer of the persistent grid (balanced to +-1 position).
    const long long workers = p.active_workers;
    // consecutive buffers of the grid sit on different SMs, so neighbouring strips are written by different SMs
    const long long worker = static_cast<long long>(slot) * gridDim.x + blockIdx.x;
    const long long positions = p.num_cells * p.chunk_members;
    long long u = positions / workers * worker + (positions % workers) * worker / workers;
    long long u_end = positions / workers * (worker + 1) + (positions % workers) * (worker + 1) / workers;
    if (worker >= workers) u = u_end = 0;
    long long cell = u / p.chunk_members;
    long long step = u - cell * p.chunk_members;
    long long strip = cell % p.strip_stride;
    long long member0 = (cell / p.strip_stride) * p.chunk_members;
    auto next_tile = [&]() -> long long {
        while (u < u_end) {
            if (step == p.chunk_members) {
                step = 0;
                ++cell;
                strip = cell % p.strip_stride;
                member0 = (cell / p.strip_stride) * p.chunk_members;
            }
            const long long member = member0 + step;
            const long long t = strip + p.strip_stride * member;
            ++u;
            ++step;
            if (member < p.strip_members && t < p.num_tiles) return t;  // ragged edges of the cell grid
        }
        return -1;
    };
    // first residue-i row a tile touches
    auto first_row_of = [&](long long t) -> long long {
        const long long first_pair = t * G::kPairs;
        if (p.num_pairs <= 0xFFFFFFFFll) return static_cast<unsigned>(first_pair) / static_cast<unsigned>(p.L);
        return first_pair / p.L;
    };

    // Residue-i prefetch registers: floats lane, lane+32, ... of the 2-residue block, one mask byte.
    float pf[kStageRows];
#pragma unroll
    for (int k = 0; k < kStageRows; ++k) pf[k] = 0.f;
    uint32_t pf_mask_ballot = 0;
    auto prefetch_issue = [&](long long t) {
        const long long r0 = first_row_of(t);
        const long long last_float = p.num_rows * (A * 3) - 1;
        if (kNeedsXyz) {
            const long long base = r0 * (A * 3) + lane;
#pragma unroll
            for (int k = 0; k < kStageRows; ++k) {
                const long long idx = base + 32 * k;
                pf[k] = __ldg(p.xyz + (idx < last_float ? idx : last_float));
            }
        }
        if (kind_has_u8<KIND>()) {
            const uint8_t* am = static_cast<const uint8_t*>(p.atom_mask);
            const long long idx = r0 * A + lane;
            const long long last = p.num_rows * A - 1;
            const bool bit = (lane < 2 * A) && (__ldg(am + (idx < last ? idx : last)) != 0);
            pf_mask_ballot = __ballot_sync(0xffffffffu, bit);
        }
    };
    auto prefetch_commit = [&](float* buf) {
        if (kNeedsXyz) {
#pragma unroll
            for (int k = 0; k < kStageRows; ++k) buf[lane + 32 * k] = pf[k];
        }
    };

    // Residue j of each of this lane's Q pairs: A atoms in registers, SoA, packed two atoms per 64-bit pair.
    float2 xj[Q][NP], yj[Q][NP], zj[Q][NP];
    float mjf[Q][A];  // fp32 mask row (kF32MaskOnly)
    uint32_t mj_bits[Q];
    long long loaded_res_j[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        mj_bits[q] = 0;
        loaded_res_j[q] = -1;
    }

    long long tile = next_tile();
    int parity = 0;
    uint32_t mask_ballot = 0;
    if (tile >= 0) {
        prefetch_issue(tile);
        prefetch_commit(stage);
        mask_ballot = pf_mask_ballot;
        __syncwarp();
    }
    while (tile >= 0) {
        const long long upcoming = next_tile();
        if (upcoming >= 0) prefetch_issue(upcoming);  // consumed after this tile's rows

        const float* __restrict__ xi_stage = stage + parity * kStageFloats;
        const long long pair0 = tile * G::kPairs;
        // lane l owns pairs pair0 + l + 32 q; pair -> (row = b*L + i, j) with one division, the others follow
        long long pair[Q];
        unsigned row[Q];
        long long res_j[Q];
        int which[Q];  // 0 or 1: which of the two staged residues the pair uses (lane 0, q 0 holds the first row)
        {
            long long p0 = pair0 + lane;
            if (p0 >= p.num_pairs) p0 = p.num_pairs - 1;  // tail lanes recompute the last pair
            unsigned r0, j0;
            if (p.num_pairs <= 0xFFFFFFFFll) {  // 32-bit division whenever the pair count allows it
                const unsigned pr = static_cast<unsigned>(p0);
                r0 = pr / static_cast<unsigned>(p.L);
                j0 = pr - r0 * static_cast<unsigned>(p.L);
            } else {
                r0 = static_cast<unsigned>(p0 / p.L);
                j0 = static_cast<unsigned>(p0 - static_cast<long long>(r0) * p.L);
            }
            const unsigned first_row = __shfl_sync(0xffffffffu, r0, 0);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                long long pq = pair0 + lane + kTilePairs * q;
                unsigned rq = r0, jq = j0 + kTilePairs * q;
                if (jq >= static_cast<unsigned>(p.L)) {  // L >= 32 Q: at most one wrap inside a tile
                    jq -= p.L;
                    ++rq;
                }
                if (pq >= p.num_pairs) {
                    pq = p.num_pairs - 1;
                    rq = static_cast<unsigned>(p.num_rows - 1);
                    jq = p.L - 1;
                }
                pair[q] = pq;
                row[q] = rq;
                res_j[q] = static_cast<long long>(rq - rq % static_cast<unsigned>(p.L)) + jq;
                which[q] = static_cast<int>(rq - first_row);
            }
        }

        // Reload residue j only when the strip (or the structure) changed; warp-uniform decision.
        bool same_j = true;
#pragma unroll
        for (int q = 0; q < Q; ++q) same_j = same_j && (res_j[q] == loaded_res_j[q]);
        if (!__all_sync(0xffffffffu, same_j)) {
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                loaded_res_j[q] = res_j[q];
                const float* __restrict__ xj_ptr = p.xyz + res_j[q] * (A * 3);
                if (kNeedsXyz) {
#pragma unroll
                    for (int k = 0; k < NP; ++k) {
                        const int c0 = 2 * k, c1 = (2 * k + 1 < A) ? 2 * k + 1 : 2 * k;
                        xj[q][k] = make_float2(__ldg(xj_ptr + 3 * c0 + 0), __ldg(xj_ptr + 3 * c1 + 0));
                        yj[q][k] = make_float2(__ldg(xj_ptr + 3 * c0 + 1), __ldg(xj_ptr + 3 * c1 + 1));
                        zj[q][k] = make_float2(__ldg(xj_ptr + 3 * c0 + 2), __ldg(xj_ptr + 3 * c1 + 2));
                    }
                }
                if (does_mask)
                    mj_bits[q] = load_mask_bits<A>(static_cast<const uint8_t*>(p.atom_mask) + res_j[q] * A);
                if (KIND == kF32MaskOnly) {
                    const float* am = static_cast<const float*>(p.atom_mask);
#pragma unroll
                    for (int c = 0; c < A; ++c) mjf[q][c] = __ldg(am + res_j[q] * A + c);
                }
            }
        }

        // The previous tile of this buffer must have left shared memory before it is overwritten.
        if (is_issuer) bulk_wait_read_all();
        tile_sync<WPT>(slot);

        if (p.stores_only) {
            // diagnostic path: nothing is computed
        } else if (kNeedsXyz && Q == 1) {
            // Rows (atoms a of residue i) are processed in groups of kRowsPerGroup; the staged coordinates
            // of the next group are read (broadcast LDS) before the current group is computed.
            const float* __restrict__ xi = xi_stage + which[0] * (A * 3);
            float* my_f32 = tile_f32 + lane * G::kElemsPerPair;
            constexpr int kRowsPerGroup = 3;
            float cur[kRowsPerGroup][3], nxt[kRowsPerGroup][3];
#pragma unroll
            for (int r = 0; r < kRowsPerGroup; ++r)
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    cur[r][k] = (row_begin + r < row_end) ? xi[3 * (row_begin + r) + k] : 0.f;
#pragma unroll 1
            for (int a0 = row_begin; a0 < row_end; a0 += kRowsPerGroup) {
#pragma unroll
                for (int r = 0; r < kRowsPerGroup; ++r) {
                    const int an = a0 + kRowsPerGroup + r;
#pragma unroll
                    for (int k = 0; k < 3; ++k) nxt[r][k] = (an < row_end) ? xi[3 * an + k] : 0.f;
                }
#pragma unroll
                for (int r = 0; r < kRowsPerGroup; ++r) {
                    const int a = a0 + r;
                    if (a < row_end) {
                        const float2 nx = make_float2(-cur[r][0], -cur[r][0]);
                        const float2 ny = make_float2(-cur[r][1], -cur[r][1]);
                        const float2 nz = make_float2(-cur[r][2], -cur[r][2]);
                        float* out_row = my_f32 + a * A;
#pragma unroll
                        for (int k = 0; k < NP; ++k) {
                            const float2 dx = __fadd2_rn(xj[0][k], nx);
                            const float2 dy = __fadd2_rn(yj[0][k], ny);
                            const float2 dz = __fadd2_rn(zj[0][k], nz);
                            float2 s = __fmul2_rn(dx, dx);
                            s = __ffma2_rn(dy, dy, s);
                            s = __ffma2_rn(dz, dz, s);
                            out_row[2 * k] = sqrt_mode<SQRT>(s.x);
                            if (2 * k + 1 < A) out_row[2 * k + 1] = sqrt_mode<SQRT>(s.y);
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < kRowsPerGroup; ++r)
#pragma unroll
                    for (int k = 0; k < 3; ++k) cur[r][k] = nxt[r][k];
            }
        } else if (kNeedsXyz) {
            // several pairs per lane (small residues): the Q independent pairs provide the instruction-level
            // parallelism, so the rows are read straight from the staging area
#pragma unroll 1
            for (int a = row_begin; a < row_end; ++a) {
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const float* __restrict__ xi = xi_stage + which[q] * (A * 3) + 3 * a;
                    const float2 nx = make_float2(-xi[0], -xi[0]);
                    const float2 ny = make_float2(-xi[1], -xi[1]);
                    const float2 nz = make_float2(-xi[2], -xi[2]);
                    float* out_row = tile_f32 + (lane + kTilePairs * q) * G::kElemsPerPair + a * A;
#pragma unroll
                    for (int k = 0; k < NP; ++k) {
                        const float2 dx = __fadd2_rn(xj[q][k], nx);
                        const float2 dy = __fadd2_rn(yj[q][k], ny);
                        const float2 dz = __fadd2_rn(zj[q][k], nz);
                        float2 s = __fmul2_rn(dx, dx);
                        s = __ffma2_rn(dy, dy, s);
                        s = __ffma2_rn(dz, dz, s);
                        out_row[2 * k] = sqrt_mode<SQRT>(s.x);
                        if (2 * k + 1 < A) out_row[2 * k + 1] = sqrt_mode<SQRT>(s.y);
                    }
                }
            }
        } else if (KIND == kF32MaskOnly) {
            const float* am = static_cast<const float*>(p.atom_mask);
#pragma unroll 1
            for (int a = row_begin; a < row_end; ++a) {
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const float mi = __ldg(am + static_cast<long long>(row[q]) * A + a);
                    float* out_row = tile_f32 + (lane + kTilePairs * q) * G::kElemsPerPair + a * A;
#pragma unroll
                    for (int c = 0; c < A; ++c) out_row[c] = __fmul_rn(mi, mjf[q][c]);
                }
            }
        }
        if (does_mask && !p.stores_only) {
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const uint32_t mi_bits = (mask_ballot >> (which[q] * A)) & ((1u << A) - 1u);
                if (A == 15)
                    write_mask_block<15>(reinterpret_cast<uint32_t*>(tile_u8), lane, mi_bits, mj_bits[q]);
                else
                    write_mask_block_bytes<A>(tile_u8, lane + kTilePairs * q, mi_bits, mj_bits[q]);
            }
        }

        if constexpr (ANGLES && A >= 5) {
        if (does_angles && !p.stores_only) {
            // trRosetta triple of this lane's pairs, reference definitions
            // (protstruc/protstruc.py:810-815): real CB in slot 4.
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float* __restrict__ xi = xi_stage + which[q] * (A * 3);
                const V3 n_i{xi[0], xi[1], xi[2]}, ca_i{xi[3], xi[4], xi[5]}, cb_i{xi[12], xi[13], xi[14]};
                const V3 ca_j{xj[q][0].y, yj[q][0].y, zj[q][0].y};  // atom 1 = second half of pack 0
                const V3 cb_j{xj[q][2].x, yj[q][2].x, zj[q][2].x};  // atom 4 = first half of pack 2
                if (pair0 + lane + kTilePairs * q < p.num_pairs) {
                    float w, t, f;
                    trrosetta_triple(triple_row_side(n_i, ca_i, cb_i), ca_j, cb_j, p.omega != nullptr,
                                     p.theta != nullptr, p.phi != nullptr, w, t, f);
                    if (p.omega) p.omega[pair[q]] = w;
                    if (p.theta) p.theta[pair[q]] = t;
                    if (p.phi) p.phi[pair[q]] = f;
                }
            }
        }
        }

        // Park the prefetched residue-i data of the upcoming tile in the other staging buffer.
        if (upcoming >= 0) prefetch_commit(stage + (parity ^ 1) * kStageFloats);

        const long long elem0 = pair0 * G::kElemsPerPair;
        if (pair0 + G::kPairs <= p.num_pairs) {
            // Full tile: hand it to the TMA engine.
            fence_proxy_async_smem();
            tile_sync<WPT>(slot);
            if (is_issuer) {
                if (p.l2_hint) {
                    const uint64_t policy = l2_policy(p.l2_hint);
                    if (kind_has_f32<KIND>()) bulk_store_s2g_hint(p.dist + elem0, tile_f32, G::kDistBytes, policy);
                    if (kind_has_u8<KIND>()) bulk_store_s2g_hint(p.mask + elem0, tile_u8, G::kMaskBytes, policy);
                } else {
                    if (kind_has_f32<KIND>()) bulk_store_s2g(p.dist + elem0, tile_f32, G::kDistBytes);
                    if (kind_has_u8<KIND>()) bulk_store_s2g(p.mask + elem0, tile_u8, G::kMaskBytes);
                }
                bulk_commit();
            }
        } else {
            // Tail tile (num_pairs % 32 != 0): byte count is not 16-B granular, copy by hand.
            tile_sync<WPT>(slot);
            const int n = static_cast<int>(p.num_pairs - pair0) * G::kElemsPerPair;
            const int t = wsub * 32 + lane;
            if (kind_has_f32<KIND>())
                for (int e = t; e < n; e += 32 * WPT) p.dist[elem0 + e] = tile_f32[e];
            if (kind_has_u8<KIND>())
                for (int e = t; e < n; e += 32 * WPT) p.mask[elem0 + e] = tile_u8[e];
            tile_sync<WPT>(slot);
        }
        if constexpr (ANGLES && A >= 5) {
            // Compact (B, L, L) copies of dist[..., CA, CA], [..., CB, CB], [..., N, O] for the optional gather of
            // compact features: the tile is complete (barrier above) and stays in shared memory until this buffer's
            // next tile, so the three values of the lane's pair are read back from it and stored coalesced.
            if (does_angles && p.d_ca != nullptr && !p.stores_only) {
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    if (pair0 + lane + kTilePairs * q < p.num_pairs) {
                        const float* blk = tile_f32 + (lane + kTilePairs * q) * G::kElemsPerPair;
                        p.d_ca[pair[q]] = blk[1 * A + 1];
                        p.d_cb[pair[q]] = blk[4 * A + 4];
                        p.d_no[pair[q]] = blk[0 * A + 3];
                    }
                }
            }
        }
        __syncwarp();  // staging buffer of the upcoming tile is complete
        mask_ballot = pf_mask_ballot;
        tile = upcoming;
        parity ^= 1;
    }
    // Shared memory must stay allocated until the engine has read the last tile.
    if (is_issuer) bulk_wait_all();
    __syncwarp();
}

// ------------------------------------------------------------------ any-shape row kernel
// Any A, any L, any pointer alignment.  The unit of work is one residue-i row (b, i) restricted to a range of
// residues j: its outputs are ONE contiguous run of (j1 - j0) * A * A elements.  Thread u of the unit owns the
// column (j, c) = divmod(u, A) — residue j's atom c stays in registers (three strided loads that hit L1/L2: the
// residues of structure b are re-read by every row of b) — and walks the A atoms of residue i, which the CTA
// staged once per unit as float4 (x, y, z, mask) in shared memory and every lane reads with one broadcast
// LDS.128.  No integer division in the element loop; the stores of a warp for one atom a are a few A-element
// segments that the following iterations extend, so every 128-B line leaves L2 complete.
enum RowMaskKind { kRowNoMask = 0, kRowBoolMask = 1, kRowF32Mask = 2 };

template <int SQRT, bool HAS_DIST, int MASK>
__global__ void __launch_bounds__(256) pair_rows_kernel(
    const float* __restrict__ xyz, const void* __restrict__ atom_mask, float* __restrict__ dist,
    void* __restrict__ dist_mask, int L, int A, long long num_rows, int parts, int cols_per_part) {
    extern __shared__ float4 row_stage[];  // 2 x A entries, double buffered across units
    const long long AA = static_cast<long long>(A) * A;
    const long long num_units = num_rows * parts;
    int buf = 0;
    for (long long unit = blockIdx.x; unit < num_units; unit += gridDim.x, buf ^= 1) {
        const unsigned row = static_cast<unsigned>(unit / parts);  // b * L + i  (< 2^31, checked by the host)
        const int part = static_cast<int>(unit - static_cast<long long>(row) * parts);
        const unsigned i = row % static_cast<unsigned>(L);
        const long long first_residue = static_cast<long long>(row - i);  // b * L
        float4* xi = row_stage + buf * A;
        for (int a = threadIdx.x; a < A; a += blockDim.x) {
            const float* src = xyz + (static_cast<long long>(row) * A + a) * 3;
            float m = 0.f;
            if (MASK == kRowBoolMask)
                m = __ldg(static_cast<const uint8_t*>(atom_mask) + static_cast<long long>(row) * A + a) != 0 ? 1.f : 0.f;
            if (MASK == kRowF32Mask) m = __ldg(static_cast<const float*>(atom_mask) + static_cast<long long>(row) * A + a);
            xi[a] = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), m);
        }
        __syncthreads();  // one barrier per unit: the other buffer is only rewritten after the next barrier
        const int col_begin = part * cols_per_part;
        const int col_end = min(col_begin + cols_per_part, L * A);
        for (int u = col_begin + threadIdx.x; u < col_end; u += blockDim.x) {
            const int j = u / A;
            const int c = u - j * A;
            const long long atom_j = (first_residue + j) * A + c;
            float xj = 0.f, yj = 0.f, zj = 0.f;
            if (HAS_DIST) {
                xj = __ldg(xyz + atom_j * 3);
                yj = __ldg(xyz + atom_j * 3 + 1);
                zj = __ldg(xyz + atom_j * 3 + 2);
            }
            float mj_f = 0.f;
            bool mj_b = false;
            if (MASK == kRowBoolMask) mj_b = __ldg(static_cast<const uint8_t*>(atom_mask) + atom_j) != 0;
            if (MASK == kRowF32Mask) mj_f = __ldg(static_cast<const float*>(atom_mask) + atom_j);
            const long long e0 = (static_cast<long long>(row) * L + j) * AA + c;  // element (row, j, a = 0, c)
            float* dp = HAS_DIST ? dist + e0 : nullptr;
            uint8_t* mb = MASK == kRowBoolMask ? static_cast<uint8_t*>(dist_mask) + e0 : nullptr;
            float* mf = MASK == kRowF32Mask ? static_cast<float*>(dist_mask) + e0 : nullptr;
#pragma unroll 4
            for (int a = 0; a < A; ++a) {
                const float4 ri = xi[a];
                if (HAS_DIST) {
                    const float dx = xj - ri.x;
                    const float dy = yj - ri.y;
                    const float dz = zj - ri.z;
                    *dp = sqrt_mode<SQRT>(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
                    dp += A;
                }
                if (MASK == kRowBoolMask) {
                    *mb = (mj_b && ri.w != 0.f) ? 1 : 0;
                    mb += A;
                }
                if (MASK == kRowF32Mask) {
                    *mf = __fmul_rn(ri.w, mj_f);
                    mf += A;
                }
            }
        }
    }
}

template <int SQRT>
int launch_rows_sqrt(const float* xyz, const void* atom_mask, int mask_dtype, float* dist, void* dist_mask,
                     int L, int A, long long num_rows, int parts, int cols_per_part, unsigned grid, int threads,
                     cudaStream_t stream) {
    const size_t smem = 2 * static_cast<size_t>(A) * sizeof(float4);
#define PS_ROWS(HAS_DIST, MASK)                                                                              \
    pair_rows_kernel<SQRT, HAS_DIST, MASK><<<grid, threads, smem, stream>>>(xyz, atom_mask, dist, dist_mask, \
                                                                            L, A, num_rows, parts, cols_per_part)
    if (dist && !dist_mask) PS_ROWS(true, kRowNoMask);
    else if (dist && mask_dtype == PS_MASK_BOOL) PS_ROWS(true, kRowBoolMask);
    else if (dist) PS_ROWS(true, kRowF32Mask);
    else if (mask_dtype == PS_MASK_BOOL) PS_ROWS(false, kRowBoolMask);
    else PS_ROWS(false, kRowF32Mask);
#undef PS_ROWS
    return check_launch("pair_rows_kernel");
}

// ------------------------------------------------------------------ any-A tile kernel
// The staged idea for a run-time atom count: a CTA composes a tile of P consecutive pairs (P * A * A
// elements, contiguous in both outputs) in shared memory and hands it to the TMA engine as one bulk store per
// output, so HBM only ever sees whole lines.  Thread u of the tile owns the column (pair, c) = divmod(u, A):
// atom c of the pair's residue j stays in registers and the A atoms of residue i come from the per-tile staging
// area as broadcast LDS.128 (x, y, z, mask); a warp's shared-memory stores for one atom a are runs of A
// consecutive words.  P is a multiple of the quantum that keeps the fp32 tile's size and address 16-B granular
// (4 / gcd(A*A, 4) pairs); the byte-mask tile takes the engine too whenever its byte range happens to be 16-B
// granular (always, if P is a multiple of 16 / gcd(A*A, 16)) and coalesced word stores otherwise, as do the last,
// partial tile and outputs that are not 16-B aligned.
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

struct ColsParams {
    const float* __restrict__ xyz;
    const void* __restrict__ atom_mask;
    float* __restrict__ out_f32;
    uint8_t* __restrict__ out_u8;
    int L;
    int A;
    unsigned magic_a;  // ceil(2^32 / A): u / A == umulhi(u, magic_a) for u * A < 2^32
    int tile_pairs;    // P
    int bulk_f32;      // f32 output 16-B aligned and P on its quantum: tiles leave through the TMA engine
    int bulk_u8;       // byte output 16-B aligned: tiles whose byte range is 16-B granular leave through the engine
    long long num_pairs;
    long long num_tiles;
    int l2_hint;       // L2 eviction policy of the bulk tile stores: 0 none, 1 evict_first, 2 evict_last
};

__host__ __device__ constexpr int round_up16(int v) { return (v + 15) & ~15; }

// TA > 0: the atom count is a compile-time constant (the popular counts beyond the staged 15: 25 and atom37's 37) — the
// column loop is fully unrolled with immediate shared-memory offsets, which removes the four running-pointer
// increments per two elements and the loop control: about half of the instructions of a column.  TA = 0: any A.
template <int KIND, int SQRT, int TA>
__global__ void __launch_bounds__(384) pair_cols_kernel(const ColsParams p) {
    extern __shared__ __align__(128) unsigned char cols_smem[];
    constexpr bool kF32 = kind_has_f32<KIND>();
    constexpr bool kU8 = kind_has_u8<KIND>();
    constexpr bool kXyz = (KIND == kDistBoolMask || KIND == kDistOnly);
    constexpr bool kMaskIn = (KIND != kDistOnly);
    const int A = TA > 0 ? TA : p.A, L = p.L, P = p.tile_pairs;
    const int AA = A * A;
    float* tile_f32 = reinterpret_cast<float*>(cols_smem);
    uint8_t* tile_u8 = cols_smem + (kF32 ? round_up16(P * AA * 4) : 0);
    // residue-i staging, structure of arrays: per staged row four runs of Ae floats (x, y, z, mask), Ae = A rounded up
    // to even, so that the column loop takes TWO atoms of residue i per LDS.64 and evaluates them as one f32x2 value
    const int Ae = (A + 1) & ~1;
    float* xi = reinterpret_cast<float*>(tile_u8 + (kU8 ? round_up16(P * AA) : 0));
    const int tid = threadIdx.x;
    const bool few_rows = P <= L;  // a tile then touches at most two residue-i rows
    const bool any_bulk = p.bulk_f32 || p.bulk_u8;

    // Tile coordinates are carried from tile to tile (the grid stride is a fixed number of pairs), so the only
    // integer divisions of the kernel happen once per CTA: pair0 = (b0 * L + i0) * L + j_first.
    long long pair0 = static_cast<long long>(blockIdx.x) * P;
    int j_first, i0;
    long long b0;
    {
        const long long row0 = pair0 / L;
        j_first = static_cast<int>(pair0 - row0 * L);
        b0 = row0 / L;
        i0 = static_cast<int>(row0 - b0 * L);
    }
    const long long stride_pairs = static_cast<long long>(gridDim.x) * P;
    const long long stride_rows = stride_pairs / L;
    const int stride_j = static_cast<int>(stride_pairs - stride_rows * L);
    const long long stride_b = stride_rows / L;
    const int stride_i = static_cast<int>(stride_rows - stride_b * L);

    for (; pair0 < p.num_pairs; pair0 += stride_pairs) {
        const long long left = p.num_pairs - pair0;
        const int np = left < P ? static_cast<int>(left) : P;
        const int last_rel = j_first + np - 1;
        const int nrows = few_rows ? (last_rel >= L ? 2 : 1) : last_rel / L + 1;
        const long long structure_atom0 = b0 * L * A;  // first atom of structure b0
        const float* __restrict__ xb = p.xyz + structure_atom0 * 3;
        const uint8_t* __restrict__ mb8 = static_cast<const uint8_t*>(p.atom_mask) + structure_atom0;
        const float* __restrict__ mbf = static_cast<const float*>(p.atom_mask) + structure_atom0;
        // the engine must have read the previous tile before anyone overwrites it
        if (any_bulk && tid == 0) bulk_wait_read_all();
        for (int k = tid; k < nrows * Ae; k += blockDim.x) {
            const int rr = k / Ae, a = k - rr * Ae;
            const int atom = (i0 + rr) * A + a;  // relative to structure b0 (rows may run into structure b0 + 1)
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a < A) {
                if (kXyz) {
                    v.x = __ldg(xb + atom * 3);
                    v.y = __ldg(xb + atom * 3 + 1);
                    v.z = __ldg(xb + atom * 3 + 2);
                }
                if (KIND == kF32MaskOnly) v.w = __ldg(mbf + atom);
                else if (kMaskIn) v.w = __int_as_float(__ldg(mb8 + atom) != 0 ? 1 : 0);  // the INTEGER 0 / 1: see below
            }
            float* row = xi + rr * 4 * Ae + a;
            row[0] = v.x;
            row[Ae] = v.y;
            row[2 * Ae] = v.z;
            row[3 * Ae] = v.w;
        }
        __syncthreads();

        for (int u = tid; u < np * A; u += blockDim.x) {
            const int pl = A == 1 ? u : static_cast<int>(__umulhi(static_cast<unsigned>(u), p.magic_a));
            const int c = u - pl * A;
            const int rel = j_first + pl;
            int r, j, db;
            if (few_rows) {
                r = rel >= L ? 1 : 0;
                j = rel - (r ? L : 0);
                db = (i0 + r) >= L ? 1 : 0;
            } else {
                r = rel / L;
                j = rel - r * L;
                db = (i0 + r) / L;
            }
            const int atom_j = (db * L + j) * A + c;  // relative to structure b0; < (P / L + 2) * L * A (host-checked)
            float xj = 0.f, yj = 0.f, zj = 0.f, mj = 0.f;
            if (kXyz) {
                xj = __ldg(xb + atom_j * 3);
                yj = __ldg(xb + atom_j * 3 + 1);
                zj = __ldg(xb + atom_j * 3 + 2);
            }
            int mjm = 0;  // byte kinds: all ones if atom c of residue j is present — a mask byte is then ONE LOP3
            if (KIND == kF32MaskOnly) mj = __ldg(mbf + atom_j);
            else if (kMaskIn) mjm = __ldg(mb8 + atom_j) != 0 ? -1 : 0;
            const float2* __restrict__ rx = reinterpret_cast<const float2*>(xi + r * 4 * Ae);
            const float2* __restrict__ ry = rx + (Ae >> 1);
            const float2* __restrict__ rz = ry + (Ae >> 1);
            const float2* __restrict__ rw = rz + (Ae >> 1);
            const int off = pl * AA + c;
            float* of = tile_f32 + off;
            uint8_t* ob = tile_u8 + off;
            const float2 xj2 = make_float2(xj, xj), yj2 = make_float2(yj, yj), zj2 = make_float2(zj, zj);
            const float2 mj2 = make_float2(mj, mj);
            // Two atoms a = 2h, 2h + 1 of residue i per step, written at element offsets o0, o0 + A of the column.  The
            // staged residue-i values and the tile live in the same shared memory, so the compiler cannot move a load
            // above an earlier tile store by itself: the loads of step h + 2 are issued, in program order, before the
            // stores of step h (each LDS -> FADD2 otherwise waits its full shared-memory latency).
            struct IAtoms {
                float2 vx, vy, vz, vw;
            };
            auto load_i = [&](int h) {
                IAtoms v;
                v.vx = v.vy = v.vz = v.vw = make_float2(0.f, 0.f);
                if (kXyz) { v.vx = rx[h]; v.vy = ry[h]; v.vz = rz[h]; }
                if (KIND == kF32MaskOnly || kU8) v.vw = rw[h];
                return v;
            };
            auto emit = [&](const IAtoms& v, int o0) {
                if (kXyz) {
                    const float2 dx = __fadd2_rn(xj2, make_float2(-v.vx.x, -v.vx.y));
                    const float2 dy = __fadd2_rn(yj2, make_float2(-v.vy.x, -v.vy.y));
                    const float2 dz = __fadd2_rn(zj2, make_float2(-v.vz.x, -v.vz.y));
                    const float2 ss = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
                    of[o0] = sqrt_mode<SQRT>(ss.x);
                    of[o0 + A] = sqrt_mode<SQRT>(ss.y);
                }
                if (KIND == kF32MaskOnly) {
                    const float2 prod = __fmul2_rn(v.vw, mj2);
                    of[o0] = prod.x;
                    of[o0 + A] = prod.y;
                }
                if (kU8) {
                    ob[o0] = static_cast<uint8_t>(__float_as_int(v.vw.x) & mjm);
                    ob[o0 + A] = static_cast<uint8_t>(__float_as_int(v.vw.y) & mjm);
                }
            };
            const int full = A >> 1;
            if constexpr (TA > 0) {
                constexpr int H = TA / 2;
                IAtoms cur = load_i(0), nxt = load_i(H > 1 ? 1 : 0);
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    IAtoms after = nxt;
                    if (h + 2 < H) after = load_i(h + 2);
                    emit(cur, 2 * h * TA);  // immediate offsets
                    cur = nxt;
                    nxt = after;
                }
            } else {
                // run-time A: running output pointers (no index arithmetic in the loop).  (Loads one step ahead, as in
                // the unrolled loop, measured slower here: 0.83 instead of 0.93 of HBM at A = 20, distances only.)
                const int row_step = 2 * A;
                float* of1 = of + A;
                uint8_t* ob1 = ob + A;
#pragma unroll 2
                for (int h = 0; h < full; ++h) {
                    if (kXyz) {
                        const float2 vx = rx[h], vy = ry[h], vz = rz[h];
                        const float2 dx = __fadd2_rn(xj2, make_float2(-vx.x, -vx.y));
                        const float2 dy = __fadd2_rn(yj2, make_float2(-vy.x, -vy.y));
                        const float2 dz = __fadd2_rn(zj2, make_float2(-vz.x, -vz.y));
                        const float2 ss = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
                        *of = sqrt_mode<SQRT>(ss.x);
                        *of1 = sqrt_mode<SQRT>(ss.y);
                    }
                    if (KIND == kF32MaskOnly || kU8) {
                        const float2 vw = rw[h];
                        if (KIND == kF32MaskOnly) {
                            const float2 prod = __fmul2_rn(vw, mj2);
                            *of = prod.x;
                            *of1 = prod.y;
                        }
                        if (kU8) {
                            *ob = static_cast<uint8_t>(__float_as_int(vw.x) & mjm);
                            *ob1 = static_cast<uint8_t>(__float_as_int(vw.y) & mjm);
                        }
                    }
                    of += row_step;
                    of1 += row_step;
                    ob += row_step;
                    ob1 += row_step;
                }
                of -= full * row_step;
                ob -= full * row_step;
            }
            // the last atom of an odd A is done on its own (its pair partner is the zero padding of the staging area)
            if (A & 1) {
                const int o0 = 2 * full * A;
                if (kXyz) {
                    const float dx = xj - rx[full].x, dy = yj - ry[full].x, dz = zj - rz[full].x;
                    of[o0] = sqrt_mode<SQRT>(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
                }
                if (KIND == kF32MaskOnly) of[o0] = __fmul_rn(rw[full].x, mj);
                if (kU8) ob[o0] = static_cast<uint8_t>(__float_as_int(rw[full].x) & mjm);
            }
        }

        // coordinates of the tile this CTA takes next
        {
            j_first += stride_j;
            const int carry_j = j_first >= L ? 1 : 0;
            j_first -= carry_j ? L : 0;
            i0 += stride_i + carry_j;
            const int carry_i = i0 >= L ? 1 : 0;
            i0 -= carry_i ? L : 0;
            b0 += stride_b + carry_i;
        }
        const long long elem0 = pair0 * AA;
        const int n = np * AA;
        const bool f32_bulk = kF32 && p.bulk_f32 && ((n & 3) == 0);
        const bool u8_bulk = kU8 && p.bulk_u8 && ((elem0 & 15) == 0) && ((n & 15) == 0);
        if (f32_bulk || u8_bulk) fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0 && (f32_bulk || u8_bulk)) {
            if (p.l2_hint) {
                const uint64_t policy = l2_policy(p.l2_hint);
                if (f32_bulk) bulk_store_s2g_hint(p.out_f32 + elem0, tile_f32, static_cast<uint32_t>(n) * 4u, policy);
                if (u8_bulk) bulk_store_s2g_hint(p.out_u8 + elem0, tile_u8, static_cast<uint32_t>(n), policy);
            } else {
                if (f32_bulk) bulk_store_s2g(p.out_f32 + elem0, tile_f32, static_cast<uint32_t>(n) * 4u);
                if (u8_bulk) bulk_store_s2g(p.out_u8 + elem0, tile_u8, static_cast<uint32_t>(n));
            }
            bulk_commit();
        }
        // plain coalesced stores for whatever the engine cannot take (the barrier after the next tile's staging
        // keeps the tile intact until every thread is done reading it)
        if (kF32 && !f32_bulk)
            for (int e = tid; e < n; e += blockDim.x) p.out_f32[elem0 + e] = tile_f32[e];
        if (kU8 && !u8_bulk) {
            uint8_t* dst = p.out_u8 + elem0;
            if ((reinterpret_cast<uintptr_t>(dst) & 3u) == 0) {
                const int words = n >> 2;
                const uint32_t* src = reinterpret_cast<const uint32_t*>(tile_u8);
                for (int e = tid; e < words; e += blockDim.x) reinterpret_cast<uint32_t*>(dst)[e] = src[e];
                for (int e = 4 * words + tid; e < n; e += blockDim.x) dst[e] = tile_u8[e];
            } else {
                for (int e = tid; e < n; e += blockDim.x) dst[e] = tile_u8[e];
            }
        }
    }
    if ((p.bulk_f32 || p.bulk_u8) && tid == 0) bulk_wait_all();  // shared memory must outlive the last bulk read
}

template <int A, int KIND, int SQRT, bool ANGLES, int WPT>
int launch_tiles_wpt(const PairDistParams& p, int slots_override, cudaStream_t stream) {
    constexpr int per_slot = warp_smem_bytes<A, KIND>() + WPT * stage_bytes_per_warp<A>();
    constexpr int kMaxSmem = 227 * 1024;
    constexpr int kMaxWarps = (WPT == 1) ? 8 : (A <= 6 ? 16 : 12);
    int slots = kMaxSmem / per_slot;
    if (slots * WPT > kMaxWarps) slots = kMaxWarps / WPT;
    // Measured on B200 (profiles/r1g_k1_sweep_v6_cells.json): with two warps per tile, four tile buffers
    // (8 warps, 147 KB) sustain 6.1-6.3 TB/s on the distance + mask kernels, five or six buffers 3-8 % less.
    // The small layouts (4 / 5 atoms: 10-16 KB per tile) need MORE buffers to keep enough bytes in flight per SM:
    // four buffers 4.5 / 3.3 TB/s, six 5.7 / 4.4, eight (16 warps, 128 registers) 6.5 / 5.2 with the byte mask and
    // 6.3 / 5.3 without (profiles/r5n_small_layout_tile_buffers.jsonl; 256 x 512, 512 x 256 and 1024 x 128 residues).
    constexpr int kDefaultSlots = A <= 6 ? 8 : 4;
    if (WPT == 2 && kind_has_f32<KIND>() && slots > kDefaultSlots) slots = kDefaultSlots;
    if (slots_override > 0 && slots_override <= kMaxSmem / per_slot && slots_override * WPT <= kMaxWarps)
        slots = slots_override;
    if (slots < 1) {
        set_error("pair_tiles_kernel: a tile of %d B does not fit in shared memory", per_slot);
        return PS_ERR_BAD_SHAPE;
    }
    const int sms = sm_count_for_current_device();
    if (sms < 0) return sms;
    if (A <= 6 && slots_override <= 0 && slots > 4) {
        // small calls keep the four buffers per CTA they always had, so that their tiles spread over as many SMs as before
        const long long per_sm = (p.num_tiles + sms - 1) / sms;
        if (per_sm < slots) slots = per_sm < 4 ? 4 : static_cast<int>(per_sm);
    }
    const int smem = slots * per_slot;
    auto kernel = pair_tiles_kernel<A, KIND, SQRT, ANGLES, WPT>;
    cudaError_t err =
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(pair_tiles_kernel)");
    long long ctas = (p.num_tiles + slots - 1) / slots;
    if (ctas > sms) ctas = sms;
    PairDistParams q = p;
    const long long workers = ctas * slots;
    q.active_workers = workers;
    // Lockstep schedule: widen the strip stride to S' = S * floor(workers / S) (still a multiple of S, so
    // residue j is still reused) and give every strip to one worker.  All workers then advance member by
    // member together and the stores of the whole GPU sweep the output linearly, S' adjacent tiles at a time —
    // the access pattern HBM likes best.
    //
    // It pays when a worker stays inside one structure for several steps (residue j is reloaded whenever the
    // structure changes) and as long as few tile buffers are left without a strip.  Measured on B200
    // (profiles/r1_lockstep_by_length.txt): distances + mask gain 3-8 % at L = 256 / 384 / 1024 and at L = 250 / 300 /
    // 350 / 500 / 510 (11-16 % of the buffers idle), lose 22 % at L = 128 and 6 % at L = 229 (23 % idle).  The fused
    // kernel, which needs every warp for the angle triple, gains 2-7 % at L = 512 / 1024, is inconsistent at
    // L = 384 (+2 % with 28 structures per launch, -6 % with 128) and loses at L = 256 and whenever buffers idle.
    // Default: on when a structure holds at least 1.9 (fused: 12) such wide rows of tiles and at most 16 %
    // (fused: 6 %) of the buffers idle.
    // q.lockstep: 0 = this default, 1 = forced on (same idle limit), 2 = forced off, 3 = forced on with up to 35 % idle
    // buffers (tuning hooks).
    const long long wide = q.strip_stride * (workers / q.strip_stride);
    const long long tiles_per_structure = static_cast<long long>(p.L) * p.L / TileGeom<A>::kPairs;
    // (measured for A = 15 only; the other staged atom counts keep the cell schedule: A = 5 loses 17 % with lock-step)
    // (and for the kinds that write an fp32 tensor; the byte-mask-only kind was not measured)
    const bool automatic = A == 15 && KIND != kBoolMaskOnly && tiles_per_structure * 10 >= (ANGLES ? 120 : 19) * wide;
    const long long idle_pct_allowed = q.lockstep == 3 ? 35 : (ANGLES ? 6 : 16);
    const bool forced = q.lockstep == 1 || q.lockstep == 3;
    if ((forced || (q.lockstep == 0 && automatic)) && wide > 0 && (workers - wide) * 100 <= workers * idle_pct_allowed &&
        q.num_tiles >= 4 * wide) {
        q.strip_stride = wide;
        q.strip_members = (q.num_tiles + wide - 1) / wide;
        q.chunk_members = q.strip_members;
        q.num_cells = wide;
        q.active_workers = wide;
    } else {
        long long chunks_per_strip = (workers + q.strip_stride / 2) / q.strip_stride;  // ~ one cell per worker
        if (chunks_per_strip < 1) chunks_per_strip = 1;
        if (chunks_per_strip > q.strip_members) chunks_per_strip = q.strip_members;
        q.chunk_members = (q.strip_members + chunks_per_strip - 1) / chunks_per_strip;
        q.num_cells = q.strip_stride * ((q.strip_members + q.chunk_members - 1) / q.chunk_members);
    }
    g_last_plan.path = 0;
    g_last_plan.lockstep = (q.num_cells == q.strip_stride && q.chunk_members == q.strip_members && q.active_workers == q.strip_stride) ? 1 : 0;
    g_last_plan.ctas = ctas;
    g_last_plan.tile_buffers = workers;
    g_last_plan.active_buffers = q.active_workers;
    g_last_plan.strip_stride = q.strip_stride;
    g_last_plan.tile_pairs = TileGeom<A>::kPairs;
    ++g_last_plan.launches;
    kernel<<<static_cast<unsigned>(ctas), slots * WPT * 32, smem, stream>>>(q);
    return check_launch("pair_tiles_kernel");
}

template <int A, int KIND, int SQRT, bool ANGLES>
int launch_tiles(const PairDistParams& p, int slots_override, int wpt, cudaStream_t stream) {
    if (wpt == 2 && (KIND == kDistBoolMask || KIND == kDistOnly))
        return launch_tiles_wpt<A, KIND, SQRT, ANGLES, 2>(p, slots_override, stream);
    return launch_tiles_wpt<A, KIND, SQRT, ANGLES, 1>(p, slots_override, stream);
}

template <int A, int KIND, bool ANGLES>
int launch_tiles_sqrt(const PairDistParams& p, int sqrt_mode_id, int slots_override, int wpt,
                      cudaStream_t stream) {
    switch (sqrt_mode_id) {
        case kSqrtApproxFtz:
            return launch_tiles<A, KIND, kSqrtApproxFtz, ANGLES>(p, slots_override, wpt, stream);
        case kSqrtApprox:
            return launch_tiles<A, KIND, kSqrtApprox, ANGLES>(p, slots_override, wpt, stream);
        case kSqrtRn:
            return launch_tiles<A, KIND, kSqrtRn, ANGLES>(p, slots_override, wpt, stream);
        default:
            set_error("unknown sqrt mode %d", sqrt_mode_id);
            return PS_ERR_BAD_DTYPE;
    }
}

template <int KIND, int TA>
int launch_cols_kind_ta(const ColsParams& p, int sqrt_mode_id, unsigned grid, int threads, size_t smem, cudaStream_t stream) {
#define PS_COLS(SQRT)                                                                                              \
    do {                                                                                                           \
        auto kernel = pair_cols_kernel<KIND, SQRT, TA>;                                                            \
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);   \
        if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(pair_cols_kernel)");                   \
        kernel<<<grid, threads, smem, stream>>>(p);                                                                \
    } while (0)
    if (KIND == kDistBoolMask || KIND == kDistOnly) {
        if (sqrt_mode_id == kSqrtRn) PS_COLS(kSqrtRn);
        else if (sqrt_mode_id == kSqrtApprox) PS_COLS(kSqrtApprox);
        else PS_COLS(kSqrtApproxFtz);
    } else {
        PS_COLS(kSqrtApproxFtz);
    }
#undef PS_COLS
    return check_launch("pair_cols_kernel");
}

// Atom counts with an unrolled instantiation (besides the staged kernels of 4 / 5 / 10 / 14 / 15 atoms): 25 (the
// reference's test fixtures) and 37 (atom37).  `generic_a` keeps the run-time-A instantiation (comparison hook).
template <int KIND>
int launch_cols_kind(const ColsParams& p, int sqrt_mode_id, unsigned grid, int threads, size_t smem, bool generic_a,
                     cudaStream_t stream) {
    if (!generic_a && p.A == 25) return launch_cols_kind_ta<KIND, 25>(p, sqrt_mode_id, grid, threads, smem, stream);
    if (!generic_a && p.A == 37) return launch_cols_kind_ta<KIND, 37>(p, sqrt_mode_id, grid, threads, smem, stream);
    return launch_cols_kind_ta<KIND, 0>(p, sqrt_mode_id, grid, threads, smem, stream);
}

// Picks the tile (pairs per tile, threads per CTA) for one output kind and launches.  Returns PS_OK + launched =
// false when no tile of this atom count fits in shared memory (the caller then uses the row kernel).
// `tune` (comparison hook): bits 0-7 pairs per tile in units of the quantum (0 = choose), bits 8-11 warps per CTA,
// bits 12-13 L2 eviction policy of the bulk tile stores
// (0 = choose; 13 = the run-time-A instantiation with the chosen CTA size).
int launch_cols(const float* xyz, const void* atom_mask, int kind, float* out_f32, void* out_u8, int B, int L, int A,
                int sqrt_mode_id, int tune, bool* launched, cudaStream_t stream) {
    *launched = false;
    const long long AA = static_cast<long long>(A) * A;
    const bool f32 = kind != kBoolMaskOnly, u8 = (kind == kDistBoolMask || kind == kBoolMaskOnly);
    const long long bytes_per_pair = AA * ((f32 ? 4 : 0) + (u8 ? 1 : 0));
    auto gcd = [](long long a, long long b) { while (b) { const long long r = a % b; a = b; b = r; } return a; };
    const long long q_f32 = 4 / gcd(AA, 4), q_u8 = 16 / gcd(AA, 16);  // pairs that keep a tile 16-B granular
    const bool f32_aligned = f32 && aligned16(out_f32), u8_aligned = u8 && aligned16(out_u8);
    const long long step = f32 ? (f32_aligned ? q_f32 : 1) : (u8_aligned ? q_u8 : 1);
    const long long kBudget = 56 * 1024, kSmemPerSm = 225 * 1024, kSmemCap = 200 * 1024;
    const long long num_pairs = static_cast<long long>(B) * L * L;
    auto smem_for = [&](long long pairs) -> long long {
        const long long max_rows = (pairs + L - 2) / L + 2;  // residue-i rows a tile can touch
        return (f32 ? round_up16(static_cast<int>(pairs * AA * 4)) : 0) + (u8 ? round_up16(static_cast<int>(pairs * AA)) : 0) +
               max_rows * ((A + 1) & ~1) * static_cast<long long>(sizeof(float4));  // residue-i staging (x, y, z, mask runs)
    };
    if (step * bytes_per_pair > kSmemCap || smem_for(step) > kSmemCap) return PS_OK;  // row kernel
    // Candidates: multiples of `step` up to the budget, 128 or 256 threads.  Score = issue efficiency at the
    // resident warp count / issue slots per element, both fitted to the sweep in
    // profiles/r1r_any_shape_tile_sweep.json: a warp spends ~kPerTile slots per tile and ~kPerColumn per column
    // besides ~kPerElement per element, and the SM needs ~32 resident warps to keep issuing.
    const double kPerTile = 60.0, kPerColumn = 20.0, kPerElement = f32 ? (u8 ? 14.0 : 11.0) : 6.0;
    long long best_pairs = step;
    int best_threads = 256;
    double best_score = -1.0;
    for (long long pairs = step; pairs == step || pairs * bytes_per_pair <= kBudget; pairs += step) {
        if (pairs > step && pairs - step >= num_pairs) break;
        const long long smem = smem_for(pairs);
        if (smem > kSmemCap) break;
        // any whole number of warps (the columns of a tile rarely fill 256 threads) up to 12: 13 warps per CTA measured
        // 0.58 of HBM where 7 gave 0.76 (A = 25; profiles/r2t_any_a_tile_sweep_before_unrolled.json)
        for (int threads = 384; threads >= 64; threads -= 32) {
            const long long cols = pairs * A;
            const long long passes = (cols + threads - 1) / threads;
            long long ctas = kSmemPerSm / (smem + 1024);
            if (ctas > 2048 / threads) ctas = 2048 / threads;
            if (ctas > 32) ctas = 32;
            const double warps = static_cast<double>(ctas * threads / 32);
            const double slots_per_element = (threads / 32) * (kPerTile + passes * (kPerColumn + kPerElement * A)) /
                                             static_cast<double>(pairs * AA);
            double efficiency = 0.4 + 0.6 * warps / 32.0;
            if (efficiency > 1.0) efficiency = 1.0;
            double score = efficiency / slots_per_element;
            if (u8 && u8_aligned && pairs % q_u8 == 0) score *= 1.03;  // byte mask leaves through the engine too
            if (score > best_score) {
                best_score = score;
                best_pairs = pairs;
                best_threads = threads;
            }
        }
    }
    if ((tune & 0xFF) > 0 && smem_for((tune & 0xFF) * step) <= kSmemCap) best_pairs = (tune & 0xFF) * step;
    const bool generic_a = ((tune >> 8) & 15) == 13;
    // The unrolled instantiations run ahead of the fitted model: their best tiles were measured directly
    // (profiles/r2w_any_a_tile_sweep.json: A = 25 0.75 -> 0.87 of HBM with the mask, 0.95 -> 1.03 without).
    if (!generic_a && (tune & 0xFF) == 0 && f32 && (A == 25 || A == 37)) {
        const long long want_pairs = A == 25 ? (u8 ? 16 : 20) : (u8 ? 4 : 12);
        const int want_threads = A == 25 ? (u8 ? 128 : 192) : (u8 ? 160 : 256);
        if (want_pairs % step == 0 && want_pairs <= num_pairs && smem_for(want_pairs) <= kSmemCap) {
            best_pairs = want_pairs;
            best_threads = want_threads;
        }
    }
    if (((tune >> 8) & 15) && !generic_a) best_threads = 32 * ((tune >> 8) & 15);
    PS_REQUIRE(best_pairs * A < (1ll << 24), PS_ERR_BAD_SHAPE, "pair_dist_mask: tile of %lld pairs x %d atoms", best_pairs, A);
    ColsParams p;
    p.xyz = xyz;
    p.atom_mask = atom_mask;
    p.out_f32 = out_f32;
    p.out_u8 = static_cast<uint8_t*>(out_u8);
    p.L = L;
    p.A = A;
    p.magic_a = static_cast<unsigned>(((1ull << 32) + A - 1) / A);
    p.tile_pairs = static_cast<int>(best_pairs);
    p.bulk_f32 = (f32_aligned && best_pairs % q_f32 == 0) ? 1 : 0;
    p.bulk_u8 = u8_aligned ? 1 : 0;
    p.num_pairs = num_pairs;
    p.num_tiles = (num_pairs + best_pairs - 1) / best_pairs;
    p.l2_hint = (tune >> 12) & 3;
    if (p.l2_hint == 0 && kind == kDistBoolMask) p.l2_hint = 1;
    if (p.l2_hint == 3) p.l2_hint = 0;
    const size_t smem = static_cast<size_t>(smem_for(best_pairs));
    const int sms = sm_count_for_current_device();
    if (sms < 0) return sms;
    long long per_sm = kSmemPerSm / static_cast<long long>(smem + 1024);
    if (per_sm > 2048 / best_threads) per_sm = 2048 / best_threads;
    if (per_sm < 1) per_sm = 1;
    long long blocks = p.num_tiles;
    if (blocks > per_sm * sms) blocks = per_sm * sms;
    const unsigned grid = static_cast<unsigned>(blocks);
    *launched = true;
    g_last_plan.path = 1;
    g_last_plan.ctas = blocks;
    g_last_plan.tile_pairs = best_pairs;
    ++g_last_plan.launches;
    switch (kind) {
        case kDistBoolMask: return launch_cols_kind<kDistBoolMask>(p, sqrt_mode_id, grid, best_threads, smem, generic_a, stream);
        case kDistOnly: return launch_cols_kind<kDistOnly>(p, sqrt_mode_id, grid, best_threads, smem, generic_a, stream);
        case kF32MaskOnly: return launch_cols_kind<kF32MaskOnly>(p, sqrt_mode_id, grid, best_threads, smem, generic_a, stream);
        default: return launch_cols_kind<kBoolMaskOnly>(p, sqrt_mode_id, grid, best_threads, smem, generic_a, stream);
    }
}

// Any shape: the any-A tile kernel, one launch per f32 output (as the staged path does for fp32 masks); the row
// kernel when a tile of this atom count cannot fit in shared memory.
int launch_any_shape(const float* xyz, const void* atom_mask, int mask_dtype, float* dist, void* dist_mask,
                     int B, int L, int A, int sqrt_mode_id, bool rows_only, int tune, cudaStream_t stream);

int launch_generic(const float* xyz, const void* atom_mask, int mask_dtype, float* dist,
                   void* dist_mask, int B, int L, int A, int sqrt_mode_id, cudaStream_t stream) {
    PS_REQUIRE(A <= 1024, PS_ERR_BAD_SHAPE, "pair_dist_mask: A=%d atoms per residue exceed 1024", A);
    PS_REQUIRE(static_cast<long long>(L) * A < (1ll << 31), PS_ERR_BAD_SHAPE,
               "pair_dist_mask: L*A=%lld exceeds 2^31", static_cast<long long>(L) * A);
    const long long num_rows = static_cast<long long>(B) * L;
    const int sms = sm_count_for_current_device();
    if (sms < 0) return sms;
    const int cols = L * A;  // (j, c) columns of one row
    const int threads = cols >= 256 ? 256 : (cols + 31) / 32 * 32;
    // Rows are cut into parts (ranges of columns) when there are too few rows to balance the SMs; a part keeps at
    // least two passes of the block so the staging of residue i stays amortised.
    const long long want_units = static_cast<long long>(sms) * 8 * 4;
    int parts = 1;
    if (num_rows < want_units) {
        const long long max_parts = cols / (2 * threads) > 0 ? cols / (2 * threads) : 1;
        const long long need = (want_units + num_rows - 1) / num_rows;
        parts = static_cast<int>(need < max_parts ? need : max_parts);
    }
    const int cols_per_part = (cols + parts - 1) / parts;
    parts = (cols + cols_per_part - 1) / cols_per_part;
    long long blocks = num_rows * parts;
    const long long cap = static_cast<long long>(sms) * 8;
    if (blocks > cap) blocks = cap;
    const unsigned grid = static_cast<unsigned>(blocks);
    g_last_plan.path = 2;
    g_last_plan.ctas = blocks;
    ++g_last_plan.launches;
    if (sqrt_mode_id == kSqrtRn)
        return launch_rows_sqrt<kSqrtRn>(xyz, atom_mask, mask_dtype, dist, dist_mask, L, A, num_rows, parts,
                                         cols_per_part, grid, threads, stream);
    if (sqrt_mode_id == kSqrtApprox)
        return launch_rows_sqrt<kSqrtApprox>(xyz, atom_mask, mask_dtype, dist, dist_mask, L, A, num_rows, parts,
                                             cols_per_part, grid, threads, stream);
    return launch_rows_sqrt<kSqrtApproxFtz>(xyz, atom_mask, mask_dtype, dist, dist_mask, L, A, num_rows, parts,
                                            cols_per_part, grid, threads, stream);
}

int launch_any_shape(const float* xyz, const void* atom_mask, int mask_dtype, float* dist, void* dist_mask,
                     int B, int L, int A, int sqrt_mode_id, bool rows_only, int tune, cudaStream_t stream) {
    const bool tiles_possible = !rows_only && A <= 128 && static_cast<long long>(L) * L < (1ll << 31);
    if (tiles_possible) {
        bool launched = false;
        int rc = PS_OK;
        if (dist_mask == nullptr || mask_dtype == PS_MASK_BOOL) {
            const int kind = dist ? (dist_mask ? kDistBoolMask : kDistOnly) : kBoolMaskOnly;
            rc = launch_cols(xyz, atom_mask, kind, dist, dist_mask, B, L, A, sqrt_mode_id, tune, &launched, stream);
            if (rc != PS_OK || launched) return rc;
        } else {
            bool first = true, second = false;
            if (dist) rc = launch_cols(xyz, atom_mask, kDistOnly, dist, nullptr, B, L, A, sqrt_mode_id, tune, &first, stream);
            if (rc != PS_OK) return rc;
            if (first) {
                rc = launch_cols(xyz, atom_mask, kF32MaskOnly, static_cast<float*>(dist_mask), nullptr, B, L, A,
                                 sqrt_mode_id, tune, &second, stream);
                if (rc != PS_OK || second) return rc;
                // (same atom count, same footprint per element: if the first fitted, so does the second)
            }
        }
    }
    return launch_generic(xyz, atom_mask, mask_dtype, dist, dist_mask, B, L, A, sqrt_mode_id, stream);
}

// Kernel selection for one atom count.  A = 15 carries every tuning variant; the other staged atom counts
// (4 = backbone, 5 = backbone + CB, 10, 14 = atom14) are built for the default configuration only.
template <int A>
int dispatch_tiles(const PairDistParams& p, int mask_dtype, void* dist_mask, bool want_angles, int sqrt_id,
                   int slots_override, int wpt, cudaStream_t stream) {
    constexpr bool kAllVariants = (A == 15);
    (void)sqrt_id;  // only the A = 15 instantiation reads the two tuning arguments
    (void)wpt;
    auto dist_kind = [&](auto kind_tag, bool angles) -> int {
        constexpr int KIND = decltype(kind_tag)::value;
        if constexpr (kAllVariants) {
            return angles ? launch_tiles_sqrt<A, KIND, true>(p, sqrt_id, slots_override, wpt, stream)
                          : launch_tiles_sqrt<A, KIND, false>(p, sqrt_id, slots_override, wpt, stream);
        } else if constexpr (A >= 5) {
            return angles
                       ? launch_tiles_wpt<A, KIND, kSqrtApproxFtz, true, kDefaultWarpsPerTile>(p, slots_override, stream)
                       : launch_tiles_wpt<A, KIND, kSqrtApproxFtz, false, kDefaultWarpsPerTile>(p, slots_override, stream);
        } else {  // no CB slot: the caller has already rejected angle requests
            return launch_tiles_wpt<A, KIND, kSqrtApproxFtz, false, kDefaultWarpsPerTile>(p, slots_override, stream);
        }
    };
    if (mask_dtype == PS_MASK_BOOL || dist_mask == nullptr) {
        if (p.dist && dist_mask) return dist_kind(std::integral_constant<int, kDistBoolMask>{}, want_angles);
        if (p.dist) return dist_kind(std::integral_constant<int, kDistOnly>{}, want_angles);
        PS_REQUIRE(!want_angles, PS_ERR_NULL_POINTER, "fused angles need the distance output");
        return launch_tiles_wpt<A, kBoolMaskOnly, kSqrtApproxFtz, false, 1>(p, slots_override, stream);
    }
    // fp32 mask: distances (+angles) first, then the mask product through the same tile path.
    if (p.dist) {
        const int rc = dist_kind(std::integral_constant<int, kDistOnly>{}, want_angles);
        if (rc != PS_OK) return rc;
    } else {
        PS_REQUIRE(!want_angles, PS_ERR_NULL_POINTER, "fused angles need the distance output");
    }
    PairDistParams pm = p;
    pm.dist = static_cast<float*>(dist_mask);
    pm.mask = nullptr;
    pm.omega = pm.theta = pm.phi = nullptr;
    pm.d_ca = pm.d_cb = pm.d_no = nullptr;
    return launch_tiles_wpt<A, kF32MaskOnly, kSqrtApproxFtz, false, 1>(pm, slots_override, stream);
}

}  // namespace

int trrosetta_angles_variant_impl(const float* xyz, int B, int L, int A, int use_virtual_cb, float* omega,
                                  float* theta, float* phi, int variant, cudaStream_t stream);  // pair_angles.cu
// pair_sweep.cu
bool pair_sweep_supported(const float* xyz, const void* atom_mask, int mask_dtype, const float* dist, const void* dist_mask,
                          int L, int A);
int pair_sweep_impl(const float* xyz, const void* atom_mask, int mask_dtype, float* dist, void* dist_mask, float* omega,
                    float* theta, float* phi, float* d_ca, float* d_cb, float* d_no, int B, int L, int sqrt_id,
                    int slots_override, int stores_only, int pace_ns, int l2_hint, cudaStream_t stream);

// Diagnostic: plain 128-bit stores of a non-uniform pattern, linear sweep (what a copy kernel's write side
// does).  Gives the store ceiling of the memory system for comparison with the TMA bulk-store path.
__global__ void __launch_bounds__(256) debug_fill_pattern_kernel(float4* __restrict__ out, long long n4) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float v = static_cast<float>(i & 0xFFFFF);
        out[i] = make_float4(v, v + 0.25f, v + 0.5f, v + 0.75f);
    }
}

int debug_fill_pattern_impl(float* out, long long n, int blocks_per_sm, cudaStream_t stream) {
    PS_REQUIRE(out != nullptr && n > 0 && (n & 3) == 0, PS_ERR_BAD_SHAPE, "debug_fill_pattern: n=%lld", n);
    const int sms = sm_count_for_current_device();
    if (sms < 0) return sms;
    if (blocks_per_sm < 1) blocks_per_sm = 8;
    debug_fill_pattern_kernel<<<sms * blocks_per_sm, 256, 0, stream>>>(reinterpret_cast<float4*>(out), n / 4);
    return check_launch("debug_fill_pattern_kernel");
}

int pair_dist_last_plan_impl(long long* out, int n) {
    const long long v[9] = {g_last_plan.path, g_last_plan.lockstep, g_last_plan.ctas, g_last_plan.tile_buffers,
                            g_last_plan.active_buffers, g_last_plan.strip_stride, g_last_plan.tile_pairs,
                            g_last_plan.launches, g_last_plan.sweep};
    PS_REQUIRE(out != nullptr && n >= 0, PS_ERR_NULL_POINTER, "pair_dist_last_plan: out is NULL");
    for (int k = 0; k < n && k < 9; ++k) out[k] = v[k];
    return PS_OK;
}

// Host entry used by the C-ABI wrappers (cabi.cu).
int pair_dist_mask_compact_impl(const float* xyz, const void* atom_mask, int mask_dtype, float* dist,
                                void* dist_mask, float* omega, float* theta, float* phi, float* d_ca, float* d_cb,
                                float* d_no, int B, int L, int A, int variant, cudaStream_t stream);

int pair_dist_mask_impl(const float* xyz, const void* atom_mask, int mask_dtype, float* dist,
                        void* dist_mask, float* omega, float* theta, float* phi, int B, int L,
                        int A, int variant, cudaStream_t stream) {
    return pair_dist_mask_compact_impl(xyz, atom_mask, mask_dtype, dist, dist_mask, omega, theta, phi, nullptr, nullptr,
                                       nullptr, B, L, A, variant, stream);
}

// Strided gather of one atom-pair plane of the distance tensor: out[p] = dist[p * A * A + offset] (the compact planes
// when a shape does not take the fused staged kernel).
__global__ void __launch_bounds__(256) gather_plane_kernel(const float* __restrict__ dist, long long num_pairs,
                                                           int block, int off_ca, int off_cb, int off_no,
                                                           float* __restrict__ d_ca, float* __restrict__ d_cb,
                                                           float* __restrict__ d_no) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < num_pairs; p += stride) {
        const float* blk = dist + p * block;
        d_ca[p] = __ldg(blk + off_ca);
        d_cb[p] = __ldg(blk + off_cb);
        d_no[p] = __ldg(blk + off_no);
    }
}

static int launch_gather_planes(const float* dist, long long pairs, int A, float* d_ca, float* d_cb, float* d_no,
                                cudaStream_t stream) {
    const int sms = sm_count_for_current_device();
    if (sms < 0) return sms;
    long long blocks = (pairs + 255) / 256;
    if (blocks > 16ll * sms) blocks = 16ll * sms;
    ++g_last_plan.launches;
    gather_plane_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(dist, pairs, A * A, 1 * A + 1, 4 * A + 4,
                                                                          0 * A + 3, d_ca, d_cb, d_no);
    return check_launch("gather_plane_kernel");
}

int pair_dist_mask_compact_impl(const float* xyz, const void* atom_mask, int mask_dtype, float* dist,
                                void* dist_mask, float* omega, float* theta, float* phi, float* d_ca, float* d_cb,
                                float* d_no, int B, int L, int A, int variant, cudaStream_t stream) {
    PS_REQUIRE(B > 0 && L > 0 && A > 0, PS_ERR_BAD_SHAPE, "pair_dist_mask: B=%d L=%d A=%d must be > 0",
               B, L, A);
    PS_REQUIRE(xyz != nullptr, PS_ERR_NULL_POINTER, "pair_dist_mask: xyz is NULL");
    PS_REQUIRE((atom_mask == nullptr) == (dist_mask == nullptr), PS_ERR_NULL_POINTER,
               "pair_dist_mask: atom_mask and dist_mask must both be given or both be NULL");
    PS_REQUIRE(dist != nullptr || dist_mask != nullptr, PS_ERR_NULL_POINTER,
               "pair_dist_mask: nothing to compute (dist and dist_mask are NULL)");
    PS_REQUIRE(mask_dtype == PS_MASK_BOOL || mask_dtype == PS_MASK_F32, PS_ERR_BAD_DTYPE,
               "pair_dist_mask: unknown mask_dtype %d", mask_dtype);
    PS_REQUIRE(static_cast<long long>(B) * L < (1ll << 31), PS_ERR_BAD_SHAPE,
               "pair_dist_mask: B*L=%lld residues exceed 2^31", static_cast<long long>(B) * L);
    g_last_plan = PairDistPlan();
    const bool want_angles = omega || theta || phi;
    PS_REQUIRE(!want_angles || A >= 5, PS_ERR_BAD_SHAPE,
               "inter_residue_geometry needs the CB slot (A >= 5), got A=%d", A);
    const bool want_compact = d_ca || d_cb || d_no;
    PS_REQUIRE(!want_compact || (d_ca && d_cb && d_no && want_angles && dist), PS_ERR_NULL_POINTER,
               "compact distance planes come as a set of three, with the angles and the distance tensor");

    const int sqrt_id = variant & 3;
    const int warps_override = (variant >> 4) & 15;  // tile buffers (slots) per CTA
    const bool force_generic = (variant >> 8) & 1;
    const int wpt = ((variant >> 9) & 1) ? (3 - kDefaultWarpsPerTile) : kDefaultWarpsPerTile;

    // A = 15 (the reference's atom layout) with distances requested: the linear-sweep kernel, one schedule for every
    // length and one launch also for fp32 masks; variant bit 15 keeps the column-strip kernel (comparison hook)
    // (environment override for A/B runs of whole test suites: PROTSTRUC_B200_K1 = strip | sweep)
    static const int env_choice = [] {
        const char* e = getenv("PROTSTRUC_B200_K1");
        return e == nullptr ? 0 : (strcmp(e, "strip") == 0 ? 1 : (strcmp(e, "sweep") == 0 ? 2 : 0));
    }();
    const bool want_strip = ((variant >> 15) & 1) || (env_choice == 1 && !((variant >> 27) & 1));
    // tuning hook of every tile kernel: L2 eviction policy of the bulk tile stores (0 none, 1 evict_first, 2 evict_last)
    static const int env_l2 = [] {
        const char* e = getenv("PROTSTRUC_B200_L2HINT");
        return e == nullptr ? 0 : (atoi(e) & 7);
    }();
    if (!force_generic && !want_strip && pair_sweep_supported(xyz, atom_mask, mask_dtype, dist, dist_mask, L, A))
    {
        // Pacing defaults of the linear-sweep kernel (profiles/r2_pace_probe.json; the same for every length): HOW
        // FAST a tile buffer comes back with its next tile moves the achieved HBM bandwidth by 5-10 %.  The distance +
        // byte-mask kernel gains 5 % with the non-ftz MUFU square root (three more issue slots per element, spread
        // over the tile: 6.46 -> 6.83 TB/s); the distance + fp32-mask kernel, whose two 28.8 KB stores per tile come
        // back to back, gains 9 % when the issuing lane waits 400 ns after handing them to the engine (6.25 -> 6.8);
        // distances only and the fused kernel are at their best undisturbed.  Bit 26 switches these defaults off.
        int sweep_sqrt = sqrt_id, pace_ns = ((variant >> 28) & 7) * 100;
        if (!((variant >> 26) & 1)) {
            const bool angles = omega || theta || phi;
            if (dist_mask && mask_dtype == PS_MASK_BOOL && !angles && sqrt_id == 0) sweep_sqrt = 1;
            if (dist_mask && mask_dtype == PS_MASK_F32 && pace_ns == 0) pace_ns = 400;
        }
        // L2 eviction policy of the tile stores (variant bits 16-18, or PROTSTRUC_B200_L2HINT for whole processes;
        // 7 = none).  Default: evict_first for distances + byte mask — two interleaved output streams of 28.8 KB and
        // 7.2 KB tiles; written lines that leave L2 early reach HBM closer to the order they were written in:
        // 5.5-6.0 -> 6.6-6.8 TB/s at L = 256 / 384 / 512, unchanged at odd L (profiles/r5c_l2_hint_probe_sweep_box1.json, r5d_*_box2.jsonl: two
        // boxes).  The fused kernel loses 2 % with any hint and the fp32-mask kernel does not care: both stay without.
        int l2_hint = ((variant >> 16) & 7) ? ((variant >> 16) & 7) : env_l2;
        if (l2_hint == 0 && dist_mask && mask_dtype == PS_MASK_BOOL && !(omega || theta || phi)) l2_hint = 1;
        if (l2_hint == 7) l2_hint = 0;
        return pair_sweep_impl(xyz, atom_mask, mask_dtype, dist, dist_mask, omega, theta, phi, d_ca, d_cb, d_no, B, L,
                               sweep_sqrt, warps_override, (variant >> 10) & 1, pace_ns, l2_hint, stream);
    }

    // the staged kernel needs L >= pairs per tile (a tile then touches at most two residue-i rows)
    const bool staged_atom_count = (A == 15) || (A == 5) || (A == 10) || (A == 14) || (A == 4);
    const int tile_pairs = kTilePairs * (A <= 6 ? 4 : (A <= 10 ? 2 : 1));  // = TileGeom<A>::kPairs
    const bool fast = staged_atom_count && (L >= tile_pairs) && !force_generic && aligned16(dist) && aligned16(dist_mask);
    if (!fast) {
        const bool rows_only = (variant >> 12) & 1;
        // any-A kernel: variant bits 16-27 are its tile / CTA shape, bits 28-29 the L2 policy of its bulk stores
        // (3 = none; 0 = the default: evict_first for distances + byte mask, +6 % at 25 atoms, +1.5 % at 37, 0 at 20)
        const int cols_l2 = ((variant >> 28) & 3) ? ((variant >> 28) & 3) : (env_l2 == 7 ? 3 : (env_l2 <= 2 ? env_l2 : 0));
        int rc = launch_any_shape(xyz, atom_mask, mask_dtype, dist, dist_mask, B, L, A, sqrt_id, rows_only,
                                  ((variant >> 16) & 0xFFF) | (cols_l2 << 12), stream);
        if (rc != PS_OK || !want_angles) return rc;
        ++g_last_plan.launches;
        // the exact-sequence angle kernel: the same trrosetta_triple the fused tile kernel evaluates, so
        // inter_residue_geometry returns identical bits whichever way a shape is dispatched
        rc = trrosetta_angles_variant_impl(xyz, B, L, A, 0, omega, theta, phi, 1, stream);
        if (rc != PS_OK || !want_compact) return rc;
        return launch_gather_planes(dist, static_cast<long long>(B) * L * L, A, d_ca, d_cb, d_no, stream);
    }

    PairDistParams p;
    p.xyz = xyz;
    p.atom_mask = atom_mask;
    p.dist = dist;
    p.mask = static_cast<uint8_t*>(dist_mask);
    p.omega = omega;
    p.theta = theta;
    p.phi = phi;
    p.d_ca = d_ca;
    p.d_cb = d_cb;
    p.d_no = d_no;
    p.L = L;
    p.num_rows = static_cast<long long>(L) * B;
    p.num_pairs = static_cast<long long>(L) * L * B;
    p.num_tiles = (p.num_pairs + tile_pairs - 1) / tile_pairs;
    int g = L, h = tile_pairs;  // gcd(L, pairs per tile)
    while (h != 0) {
        const int r = g % h;
        g = h;
        h = r;
    }
    p.strip_stride = L / g;
    p.strip_members = (p.num_tiles + p.strip_stride - 1) / p.strip_stride;
    p.chunk_members = p.strip_members;  // refined per launch once the worker count is known
    p.num_cells = p.strip_stride;
    p.stores_only = (variant >> 10) & 1;
    // strip kernels: variant bits 16-17 (3 = none).  Default: evict_first for the 10- and 14-atom layouts (+7-12 % with
    // and without the byte mask, fused +2 % / +9 %: profiles/r5e_l2_hint_probe_others.json, r5k_*); 5 atoms: +2 % / -3 %
    // (fused: bound by the angle triple, 0 %), none
    p.l2_hint = ((variant >> 16) & 3) ? ((variant >> 16) & 3) : (env_l2 == 7 ? 3 : (env_l2 <= 2 ? env_l2 : 0));
    if (p.l2_hint == 0 && (A == 10 || A == 14)) p.l2_hint = 1;
    if (p.l2_hint == 3) p.l2_hint = 0;
    p.lockstep = ((variant >> 14) & 1) ? 3 : ((variant >> 11) & 1) ? 1 : (((variant >> 13) & 1) ? 2 : 0);
    p.active_workers = 0;

    // Few atoms per residue make the fused launch angle-bound: a tile of the 5-atom layout holds 128 pairs but only
    // 3,200 distances, and the triples of a tile are evaluated by ONE of its warps — 8 warps per SM cannot hide the
    // dependent chains of 4 triples per lane.  For 5 and 10 atoms the launcher therefore splits the call: distance tiles
    // without the angles, then the exact-sequence angle kernel at full occupancy — the SAME trrosetta_triple, so the
    // bits do not depend on the dispatch (as for shapes that take the any-A kernel above).  Measured
    // (profiles/r5l_split_dispatch_probe.json): 256 x 512 x 5 4.78 -> 2.39 ms, 64 x 384 x 10 1.02 -> 0.82 ms, on every
    // input kind; 14 atoms 0.81 -> 0.78 ms (not worth a second launch), 15 atoms take the sweep kernel.  Below 32 k
    // pairs (one structure of 181 residues) the call is launch-bound and stays ONE launch.  Variant bit 19 keeps the
    // fused launch, bit 20 splits any staged atom count and size (comparison hooks).
    const bool split_angles = want_angles && !((variant >> 19) & 1) &&
                              (((A == 5 || A == 10) && p.num_pairs >= 32768) || ((variant >> 20) & 1));
    if (split_angles) {
        p.omega = p.theta = p.phi = nullptr;
        p.d_ca = p.d_cb = p.d_no = nullptr;
    }
    const bool tile_angles = want_angles && !split_angles;
    int rc;
    switch (A) {
        case 4:
            rc = dispatch_tiles<4>(p, mask_dtype, dist_mask, tile_angles, sqrt_id, warps_override, wpt, stream);
            break;
        case 5:
            rc = dispatch_tiles<5>(p, mask_dtype, dist_mask, tile_angles, sqrt_id, warps_override, wpt, stream);
            break;
        case 10:
            rc = dispatch_tiles<10>(p, mask_dtype, dist_mask, tile_angles, sqrt_id, warps_override, wpt, stream);
            break;
        case 14:
            rc = dispatch_tiles<14>(p, mask_dtype, dist_mask, tile_angles, sqrt_id, warps_override, wpt, stream);
            break;
        default:
            rc = dispatch_tiles<15>(p, mask_dtype, dist_mask, tile_angles, sqrt_id, warps_override, wpt, stream);
            break;
    }
    if (rc != PS_OK || !split_angles) return rc;
    ++g_last_plan.launches;
    rc = trrosetta_angles_variant_impl(xyz, B, L, A, 0, omega, theta, phi, 1, stream);
    if (rc != PS_OK || !want_compact) return rc;
    return launch_gather_planes(dist, static_cast<long long>(B) * L * L, A, d_ca, d_cb, d_no, stream);
}

}  // namespace ps
