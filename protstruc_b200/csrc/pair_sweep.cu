// K1, linear-sweep flavour — all-atom pairwise distances + pair mask (+ the fused trRosetta angles) for A = 15 and ANY
// structure length, with every tile buffer of the GPU writing ADJACENT tiles at any moment.
//
// Replaces StructureBatch.pairwise_distance_matrix (protstruc/protstruc.py:455-484) and, fused,
// StructureBatch.inter_residue_geometry (protstruc/protstruc.py:790-817), like pair_dist.cu.
//
// Why a second schedule.  The column-strip kernel of pair_dist.cu keeps residue j of a lane in registers by letting a
// tile buffer walk tiles t, t + S, t + 2S, ... (S = L / gcd(L, 32)); its "lock-step" variant, in which all buffers sweep
// the output linearly (the access pattern HBM rewards with 5-11 %), needs the number of buffers to be a multiple of S —
// for odd L only with 14-23 % of the buffers idle — and was therefore switched on by a table of hand-fitted length /
// kind thresholds.  Here the sweep is linear for every L: buffer w takes tiles w, w + W, w + 2W, ... (W = all buffers
// of the grid), and residue j is simply RE-STAGED for every tile by the TMA engine:
//
//  * a tile's 32 residues j are one contiguous run of the coordinate array (two runs when the tile wraps into the next
//    residue-i row), 5.8 KB; one elected lane issues cp.async.bulk.shared::cluster.global (SASS UBLKCP, completion on
//    an mbarrier with expect_tx) for the 16-byte aligned superset of each run, ONE TILE AHEAD, into a per-buffer
//    staging area — no issue slots, no registers, no scattered loads;
//  * when the tile starts, each lane copies its residue (45 floats at a lane stride of 45 words: conflict-free LDS)
//    into the same packed f32x2 registers the strip kernel uses, the staging area is handed back to the engine for the
//    next tile, and the row loop, the word-wise mask writer, the fused angle triple and the TMA tile store are the
//    strip kernel's, unchanged;
//  * the mask bytes of the 32 residues j (480 B) are staged the same way; the lane's 15 bytes arrive as five aligned
//    words and one funnel shift per word, which IS the "mask row as four words" the word-wise writer starts from;
//  * fp32 masks (the reference's from_pdb_id path) take ONE launch: the distance tile and the mask-product tile are
//    produced from the same staged residues and leave as two bulk stores (8 B per element).
//
// Tiles whose aligned superset would run past the end of an input array (the last < 16 bytes of the batch) are staged
// by the lanes themselves with plain loads.
//
// Roofline: HBM write bandwidth, as for pair_dist.cu.

#include "pair_tiles.cuh"

namespace ps {

namespace {

constexpr int kSweepA = 15;
constexpr int kResidueFloats = kSweepA * 3;         // 45
constexpr int kResidueBytes = kResidueFloats * 4;   // 180
// staging area of one tile buffer: two runs of residues, each with up to 3 leading residues and < 16 trailing bytes
constexpr int kStageXyzBytes = (38 * kResidueBytes + 48 + 15) / 16 * 16;   // 6896
constexpr int kStageMaskU8Bytes = 62 * kSweepA + 48 + 16;                // two runs, up to 15 leading residues each
constexpr int kStageMaskF32Bytes = (38 * kSweepA * 4 + 48 + 15) / 16 * 16;

enum SweepKind { kSweepDistBool = 0, kSweepDistOnly = 1, kSweepDistF32 = 2 };
constexpr int kMaxPeers = 8;

struct SweepParams {
    const float* __restrict__ xyz;
    const void* __restrict__ atom_mask;
    float* __restrict__ dist;
    void* __restrict__ mask_out;  // bool bytes (kSweepDistBool) or fp32 (kSweepDistF32)
    float* __restrict__ omega;
    float* __restrict__ theta;
    float* __restrict__ phi;
    float* __restrict__ d_ca;
    float* __restrict__ d_cb;
    float* __restrict__ d_no;
    int L;
    long long num_rows;
    long long num_pairs;
    long long num_tiles;
    int stores_only;
    int pace_ns;  // tuning hook: the issuing lane sleeps this long after handing a tile to the engine
    int l2_hint;  // tuning hook: L2 eviction policy of the tile stores (0 = none, 1 = evict_first, 2 = evict_last,
                  // 3 = evict_first for the distance tile only, 4 = for the mask tile only)
    // Fused all-gather of the compact features (optional): every rank's gathered buffer (6, world, shard, L, L) as
    // mapped into THIS GPU's address space over NVLink (peer[r], r = 0 .. n_peers-1, own buffer included), or ONE
    // NVSwitch multicast address that reaches all of them (mc).  The angle warp stores omega / theta / phi and the
    // three compact distance planes of its pair straight into them: the transfer rides on the kernel, tile by tile.
    float* peer[kMaxPeers];
    float* mc;
    int n_peers;
    long long peer_plane;   // floats between two feature planes of a gathered buffer: world * shard * L * L
    long long peer_offset;  // this rank's slab inside a plane: rank * shard * L * L
};

template <int KIND>
__host__ __device__ constexpr int sweep_tile_bytes() {
    return TileGeom<kSweepA>::kDistBytes +
           (KIND == kSweepDistBool ? TileGeom<kSweepA>::kMaskBytes : (KIND == kSweepDistF32 ? TileGeom<kSweepA>::kDistBytes : 0));
}
template <int KIND>
__host__ __device__ constexpr int sweep_stage_mask_bytes() {
    return KIND == kSweepDistBool ? (kStageMaskU8Bytes + 15) / 16 * 16 : (KIND == kSweepDistF32 ? kStageMaskF32Bytes : 0);
}
// per tile buffer: tile, residue-j staging (coordinates, mask), residue-i staging of its two warps, mbarrier
template <int KIND>
__host__ __device__ constexpr int sweep_slot_bytes() {
    return sweep_tile_bytes<KIND>() + kStageXyzBytes + sweep_stage_mask_bytes<KIND>() + 2 * stage_bytes_per_warp<kSweepA>() +
           (KIND == kSweepDistF32 ? 2 * 2 * 32 * 4 : 0) + 16;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "SWEEP_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra SWEEP_DONE_%=;\n\t"
        "bra SWEEP_WAIT_%=;\n\t"
        "SWEEP_DONE_%=:\n\t"
        "}"
        :: "r"(addr), "r"(parity) : "memory");
}
// global -> shared bulk copy through the TMA engine, completion counted on `bar`
__device__ __forceinline__ void bulk_load_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(static_cast<uint32_t>(__cvta_generic_to_shared(sdst))), "l"(gsrc), "r"(bytes),
                    "r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar)))
                 : "memory");
}

// One value of one compact feature plane to every rank of the job: a multicast store through the NVSwitch when the
// buffer has a multicast mapping (one packet leaves this GPU, the switch replicates it), else one store per peer.
__device__ __forceinline__ void push_to_peers(const SweepParams& p, int feature, long long pair, float v) {
    const long long at = feature * p.peer_plane + p.peer_offset + pair;
    if (p.mc != nullptr) {
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" :: "l"(p.mc + at), "f"(v) : "memory");
    } else {
        for (int r = 0; r < p.n_peers; ++r) p.peer[r][at] = v;
    }
}

// One contiguous run of residues j of a tile and where its 16-byte aligned superset lands in the staging areas.
struct Run {
    long long first;  // global residue index of the run's first residue
    int count;        // residues in the run (0: no run)
};

// Geometry of a tile: its two residue-j runs.  Lane l < run[0].count belongs to run 0, the others to run 1.
struct TileRuns {
    Run run[2];
    long long row0;   // first residue-i row (b * L + i)
    int j0;           // first residue j within the structure of row0
};

// Where a tile buffer stands in the (row, j) plane.  A buffer advances by a FIXED number of pairs per step (all buffers
// of the grid times 32), so its position is carried from tile to tile with adds and compares only; the divisions by
// L happen once per buffer, in `start`.
struct TileCursor {
    long long pair0;   // first pair of the tile
    unsigned row0;     // its residue-i row b * L + i  (B * L < 2^31, checked by the host)
    int i0;            // i of row0 within its structure
    int j0;            // first residue j of the tile within its structure
    long long d_pair;  // per step: pairs, and the same split into rows / residues
    unsigned d_row;
    int d_i, d_j;

    __device__ __forceinline__ void start(long long tile, long long tiles_per_step, int L) {
        pair0 = tile * kTilePairs;
        const long long r = pair0 / L;
        row0 = static_cast<unsigned>(r);
        j0 = static_cast<int>(pair0 - r * L);
        i0 = static_cast<int>(r % L);
        d_pair = tiles_per_step * kTilePairs;
        const long long dr = d_pair / L;
        d_row = static_cast<unsigned>(dr);
        d_j = static_cast<int>(d_pair - dr * L);
        d_i = static_cast<int>(dr % L);
    }
    __device__ __forceinline__ void advance(int L) {
        pair0 += d_pair;
        j0 += d_j;
        const int carry = j0 >= L ? 1 : 0;
        j0 -= carry ? L : 0;
        row0 += d_row + carry;
        i0 += d_i + carry;
        i0 -= i0 >= L ? L : 0;
    }
    __device__ __forceinline__ TileRuns runs(const SweepParams& p) const {
        TileRuns t;
        t.row0 = row0;
        t.j0 = j0;
        const long long left = p.num_pairs - pair0;
        const int np = left < kTilePairs ? static_cast<int>(left) : kTilePairs;
        const int n0 = min(np, p.L - j0);
        const long long base0 = static_cast<long long>(row0) - i0;  // first residue of the structure of row0
        t.run[0] = Run{base0 + j0, n0};
        // the wrap continues at residue 0 of the NEXT row's structure: the same one unless row0 was its last row
        t.run[1] = Run{i0 + 1 == p.L ? base0 + p.L : base0, np - n0};
        return t;
    }
};

// Aligned superset of a run in a per-residue array of `bytes_per_residue` bytes whose residues are 16-byte aligned
// every `align_residues` residues: byte offset of the superset, its size, and the lead (bytes before the run).
struct Span {
    long long offset;
    uint32_t bytes;
    uint32_t lead;
};
__device__ __forceinline__ Span span_of(Run r, int bytes_per_residue, int align_residues) {
    const long long a = r.first - r.first % align_residues;
    Span s;
    s.offset = a * bytes_per_residue;
    s.lead = static_cast<uint32_t>(r.first - a) * bytes_per_residue;
    s.bytes = (s.lead + static_cast<uint32_t>(r.count) * bytes_per_residue + 15u) & ~15u;
    return s;
}

// The staged kernel with a linear sweep.  Two warps per tile buffer; warp 0: rows 0 .. kSplitRow-1, the angle triple,
// the TMA issue; warp 1: the remaining rows and the mask block.
template <int KIND, int SQRT, bool ANGLES>
__global__ void __launch_bounds__(256, 1) pair_sweep_kernel(const SweepParams p) {
    constexpr int A = kSweepA;
    using G = TileGeom<A>;
    constexpr int WPT = 2;
    constexpr int NP = (A + 1) / 2;
    constexpr int kStageFloats = stage_floats<A>();
    constexpr int kStageRows = kStageFloats / 32;
    constexpr int kSplitRow = ANGLES ? (A - 1) / 2 : (A * 3) / 5;
    constexpr bool kBool = KIND == kSweepDistBool;
    constexpr bool kF32 = KIND == kSweepDistF32;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int slot = warp / WPT;
    const int wsub = warp % WPT;
    const int slots_per_cta = (blockDim.x >> 5) / WPT;
    const bool does_rows_lo = (wsub == 0);
    const bool does_mask = (kBool || kF32) && (wsub == 1);
    const bool does_angles = ANGLES && (wsub == 0);
    const bool is_issuer = (wsub == 0) && (lane == 0);
    const int row_begin = does_rows_lo ? 0 : kSplitRow;
    const int row_end = does_rows_lo ? kSplitRow : A;

    // shared-memory carve-up of this tile buffer
    unsigned char* sbase = smem_raw + static_cast<size_t>(slot) * sweep_slot_bytes<KIND>();
    float* tile_f32 = reinterpret_cast<float*>(sbase);
    unsigned char* tile_mask = sbase + G::kDistBytes;  // bool bytes or fp32 mask products
    unsigned char* stage_xyz = sbase + sweep_tile_bytes<KIND>();
    unsigned char* stage_mask = stage_xyz + kStageXyzBytes;
    float* stage_i = reinterpret_cast<float*>(stage_mask + sweep_stage_mask_bytes<KIND>()) + wsub * (2 * kStageFloats);
    float* stage_mi = reinterpret_cast<float*>(stage_mask + sweep_stage_mask_bytes<KIND>() + 2 * stage_bytes_per_warp<A>()) +
                      wsub * 64;  // fp32 mask values of the two residues i (kF32), double buffered
    uint64_t* bar = reinterpret_cast<uint64_t*>(sbase + sweep_slot_bytes<KIND>() - 16);

    const long long workers = static_cast<long long>(gridDim.x) * slots_per_cta;
    // consecutive buffers of the grid sit on different SMs, so neighbouring tiles are written by different SMs
    const long long worker = static_cast<long long>(slot) * gridDim.x + blockIdx.x;

    if (is_issuer) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const long long xyz_bytes = p.num_rows * kResidueBytes;
    const long long mask_bytes = p.num_rows * A * (kF32 ? 4 : 1);
    const char* xyz_b = reinterpret_cast<const char*>(p.xyz);
    const char* mask_b = reinterpret_cast<const char*>(p.atom_mask);

    // A tile is staged by the engine unless one of its supersets would run past the end of an input array.
    auto engine_ok = [&](const TileRuns& t) -> bool {
        bool ok = true;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (t.run[k].count == 0) continue;
            const Span sx = span_of(t.run[k], kResidueBytes, 4);
            ok = ok && (sx.offset + sx.bytes <= xyz_bytes);
            if (kBool) {
                const Span sm = span_of(t.run[k], A, 16);
                ok = ok && (sm.offset + sm.bytes <= mask_bytes);
            }
            if (kF32) {
                const Span sm = span_of(t.run[k], A * 4, 4);
                ok = ok && (sm.offset + sm.bytes <= mask_bytes);
            }
        }
        return ok;
    };
    // issuer only: arm the barrier and start the bulk loads of a tile
    auto stage_issue = [&](const TileRuns& t) {
        uint32_t total = 0, off_x = 0, off_m = 0;
        Span sx[2], sm[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            sx[k] = span_of(t.run[k], kResidueBytes, 4);
            sm[k] = kBool ? span_of(t.run[k], A, 16) : span_of(t.run[k], A * 4, 4);
            if (t.run[k].count > 0) total += sx[k].bytes + ((kBool || kF32) ? sm[k].bytes : 0u);
        }
        mbar_expect_tx(bar, total);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (t.run[k].count == 0) continue;
            bulk_load_g2s(stage_xyz + off_x, xyz_b + sx[k].offset, sx[k].bytes, bar);
            off_x += sx[k].bytes;
            if (kBool || kF32) {
                bulk_load_g2s(stage_mask + off_m, mask_b + sm[k].offset, sm[k].bytes, bar);
                off_m += sm[k].bytes;
            }
        }
    };

    // Residue-i prefetch registers (as in pair_dist.cu): floats lane, lane + 32, ... of the 2-residue block.
    float pf[kStageRows];
#pragma unroll
    for (int k = 0; k < kStageRows; ++k) pf[k] = 0.f;
    uint32_t pf_mask_ballot = 0;
    float pf_mi = 0.f;
    auto prefetch_issue = [&](long long r0) {
        const long long last_float = p.num_rows * (A * 3) - 1;
        const long long base = r0 * (A * 3) + lane;
#pragma unroll
        for (int k = 0; k < kStageRows; ++k) {
            const long long idx = base + 32 * k;
            pf[k] = __ldg(p.xyz + (idx < last_float ? idx : last_float));
        }
        if (kBool) {
            const uint8_t* am = static_cast<const uint8_t*>(p.atom_mask);
            const long long idx = r0 * A + lane;
            const long long last = p.num_rows * A - 1;
            const bool bit = (lane < 2 * A) && (__ldg(am + (idx < last ? idx : last)) != 0);
            pf_mask_ballot = __ballot_sync(0xffffffffu, bit);
        }
        if (kF32) {
            const float* am = static_cast<const float*>(p.atom_mask);
            const long long idx = r0 * A + lane;
            const long long last = p.num_rows * A - 1;
            pf_mi = __ldg(am + (idx < last ? idx : last));
        }
    };
    auto prefetch_commit = [&](int par) {
        float* buf = stage_i + par * kStageFloats;
#pragma unroll
        for (int k = 0; k < kStageRows; ++k) buf[lane + 32 * k] = pf[k];
        if (kF32) stage_mi[par * 32 + lane] = pf_mi;
    };

    float2 xj[NP], yj[NP], zj[NP];
    float mjf[A];
    uint32_t mj_words[4] = {0u, 0u, 0u, 0u};

    long long tile = worker < p.num_tiles ? worker : -1;
    TileCursor cursor;
    cursor.start(tile >= 0 ? tile : 0, workers, p.L);
    TileRuns cur{};
    bool cur_engine = false;
    int parity = 0;          // residue-i staging buffer
    uint32_t phase = 0;      // mbarrier phase of the residue-j staging area
    uint32_t mask_ballot = 0;
    if (tile >= 0) {
        cur = cursor.runs(p);
        cur_engine = engine_ok(cur);
        if (is_issuer && cur_engine) stage_issue(cur);
        prefetch_issue(cur.row0);
        prefetch_commit(0);
        mask_ballot = pf_mask_ballot;
        __syncwarp();
    }
    while (tile >= 0) {
        const long long upcoming = (tile + workers < p.num_tiles) ? tile + workers : -1;
        TileRuns nxt{};
        bool nxt_engine = false;
        if (upcoming >= 0) {
            cursor.advance(p.L);
            nxt = cursor.runs(p);
            nxt_engine = engine_ok(nxt);
            prefetch_issue(nxt.row0);  // consumed after this tile's rows
        }
        const long long pair0 = tile * kTilePairs;
        const int n0 = cur.run[0].count;
        const int np = n0 + cur.run[1].count;
        // lane -> (which staged residue i, which residue of which run); tail lanes recompute the last pair
        const int lp = lane < np ? lane : np - 1;
        const int which = lp < n0 ? 0 : 1;
        const long long pair = pair0 + lp;

        // ---- residue j of this lane from the staging area
        const Span sx0 = span_of(cur.run[0], kResidueBytes, 4);
        const Span sx1 = span_of(cur.run[1], kResidueBytes, 4);
        if (cur_engine) {
            mbar_wait(bar, phase);
            phase ^= 1u;
        } else {
            // the lanes stage the runs themselves: same layout, plain loads
            tile_sync<WPT>(slot);
            const int t2 = wsub * 32 + lane;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const Run r = cur.run[k];
                if (r.count == 0) continue;
                const Span s = k == 0 ? sx0 : sx1;
                float* dst = reinterpret_cast<float*>(stage_xyz + (k == 0 ? 0u : sx0.bytes) + s.lead);
                const float* src = p.xyz + r.first * kResidueFloats;
                for (int f = t2; f < r.count * kResidueFloats; f += 64) dst[f] = __ldg(src + f);
                if (kBool) {
                    const Span m0 = span_of(cur.run[0], A, 16), m = span_of(r, A, 16);
                    uint8_t* dm = stage_mask + (k == 0 ? 0u : m0.bytes) + m.lead;
                    const uint8_t* sm = static_cast<const uint8_t*>(p.atom_mask) + r.first * A;
                    for (int f = t2; f < r.count * A; f += 64) dm[f] = __ldg(sm + f);
                }
                if (kF32) {
                    const Span m0 = span_of(cur.run[0], A * 4, 4), m = span_of(r, A * 4, 4);
                    float* dm = reinterpret_cast<float*>(stage_mask + (k == 0 ? 0u : m0.bytes) + m.lead);
                    const float* sm = static_cast<const float*>(p.atom_mask) + r.first * A;
                    for (int f = t2; f < r.count * A; f += 64) dm[f] = __ldg(sm + f);
                }
            }
            tile_sync<WPT>(slot);
        }
        {
            const int idx = which == 0 ? lp : lp - n0;
            const float* src = reinterpret_cast<const float*>(stage_xyz + (which == 0 ? sx0.lead : sx0.bytes + sx1.lead)) +
                               idx * kResidueFloats;
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const int c0 = 2 * k, c1 = (2 * k + 1 < A) ? 2 * k + 1 : 2 * k;
                xj[k] = make_float2(src[3 * c0 + 0], src[3 * c1 + 0]);
                yj[k] = make_float2(src[3 * c0 + 1], src[3 * c1 + 1]);
                zj[k] = make_float2(src[3 * c0 + 2], src[3 * c1 + 2]);
            }
            if (kBool && does_mask) {
                // the lane's 15 mask bytes: five aligned words, one funnel shift each -> the row as four words
                const Span m0 = span_of(cur.run[0], A, 16), m1 = span_of(cur.run[1], A, 16);
                const uint32_t byte0 = (which == 0 ? m0.lead : m0.bytes + m1.lead) + static_cast<uint32_t>(idx) * A;
                const uint32_t* w = reinterpret_cast<const uint32_t*>(stage_mask) + (byte0 >> 2);
                const uint32_t sh = (byte0 & 3u) * 8u;
                const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
                mj_words[0] = __funnelshift_r(w0, w1, sh);
                mj_words[1] = __funnelshift_r(w1, w2, sh);
                mj_words[2] = __funnelshift_r(w2, w3, sh);
                mj_words[3] = __funnelshift_r(w3, w4, sh) & 0x00FFFFFFu;
#pragma unroll
                for (int q = 0; q < 4; ++q) {  // any non-zero byte counts as 1, like `mask != 0`
                    const uint32_t v = mj_words[q];
                    mj_words[q] = ((((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) >> 7) & 0x01010101u;
                }
            }
            if (kF32 && does_mask) {
                const Span m0 = span_of(cur.run[0], A * 4, 4), m1 = span_of(cur.run[1], A * 4, 4);
                const float* srcm = reinterpret_cast<const float*>(stage_mask + (which == 0 ? m0.lead : m0.bytes + m1.lead)) + idx * A;
#pragma unroll
                for (int c = 0; c < A; ++c) mjf[c] = srcm[c];
            }
        }

        // The previous tile of this buffer must have left shared memory before it is overwritten; after this barrier
        // both warps hold residue j in registers, so the staging area goes back to the engine for the next tile.
        if (is_issuer) bulk_wait_read_all();
        tile_sync<WPT>(slot);
        if (is_issuer && upcoming >= 0 && nxt_engine) stage_issue(nxt);

        const float* __restrict__ xi_stage = stage_i + parity * kStageFloats;
        if (!p.stores_only) {
            const float* __restrict__ xi = xi_stage + which * (A * 3);
            float* my_f32 = tile_f32 + lane * G::kElemsPerPair;
            constexpr int kRowsPerGroup = 3;
            float cur_r[kRowsPerGroup][3], nxt_r[kRowsPerGroup][3];
#pragma unroll
            for (int r = 0; r < kRowsPerGroup; ++r)
#pragma unroll
                for (int k = 0; k < 3; ++k) cur_r[r][k] = (row_begin + r < row_end) ? xi[3 * (row_begin + r) + k] : 0.f;
#pragma unroll 1
            for (int a0 = row_begin; a0 < row_end; a0 += kRowsPerGroup) {
#pragma unroll
                for (int r = 0; r < kRowsPerGroup; ++r) {
                    const int an = a0 + kRowsPerGroup + r;
#pragma unroll
                    for (int k = 0; k < 3; ++k) nxt_r[r][k] = (an < row_end) ? xi[3 * an + k] : 0.f;
                }
#pragma unroll
                for (int r = 0; r < kRowsPerGroup; ++r) {
                    const int a = a0 + r;
                    if (a < row_end) {
                        const float2 nx = make_float2(-cur_r[r][0], -cur_r[r][0]);
                        const float2 ny = make_float2(-cur_r[r][1], -cur_r[r][1]);
                        const float2 nz = make_float2(-cur_r[r][2], -cur_r[r][2]);
                        float* out_row = my_f32 + a * A;
#pragma unroll
                        for (int k = 0; k < NP; ++k) {
                            const float2 dx = __fadd2_rn(xj[k], nx);
                            const float2 dy = __fadd2_rn(yj[k], ny);
                            const float2 dz = __fadd2_rn(zj[k], nz);
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
                    for (int k = 0; k < 3; ++k) cur_r[r][k] = nxt_r[r][k];
            }
            if (kBool && does_mask) {
                const uint32_t mi_bits = (mask_ballot >> (which * A)) & ((1u << A) - 1u);
                write_mask_block_rows<A>(reinterpret_cast<uint32_t*>(tile_mask), lane, mi_bits, mj_words);
            }
            if (kF32 && does_mask) {
                const float* mi = stage_mi + parity * 32 + which * A;
                float* out = reinterpret_cast<float*>(tile_mask) + lane * G::kElemsPerPair;
#pragma unroll 3
                for (int a = 0; a < A; ++a) {
                    const float m = mi[a];
#pragma unroll
                    for (int c = 0; c < A; ++c) out[a * A + c] = __fmul_rn(m, mjf[c]);
                }
            }
            if constexpr (ANGLES) {
                if (does_angles && lane < np) {
                    // trRosetta triple of this lane's pair, reference definitions (protstruc/protstruc.py:810-815)
                    const V3 n_i{xi[0], xi[1], xi[2]}, ca_i{xi[3], xi[4], xi[5]}, cb_i{xi[12], xi[13], xi[14]};
                    const V3 ca_j{xj[0].y, yj[0].y, zj[0].y};  // atom 1 = second half of pack 0
                    const V3 cb_j{xj[2].x, yj[2].x, zj[2].x};  // atom 4 = first half of pack 2
                    float w, t, f;
                    const bool push = p.n_peers > 0;
                    trrosetta_triple(triple_row_side(n_i, ca_i, cb_i), ca_j, cb_j, push || p.omega != nullptr,
                                     push || p.theta != nullptr, push || p.phi != nullptr, w, t, f);
                    if (p.l2_hint >= 5) {  // tuning hook: streaming (evict-first) stores for the angle planes
                        if (p.omega) __stcs(p.omega + pair, w);
                        if (p.theta) __stcs(p.theta + pair, t);
                        if (p.phi) __stcs(p.phi + pair, f);
                    } else {
                        if (p.omega) p.omega[pair] = w;
                        if (p.theta) p.theta[pair] = t;
                        if (p.phi) p.phi[pair] = f;
                    }
                    if (push) {
                        push_to_peers(p, 0, pair, w);
                        push_to_peers(p, 1, pair, t);
                        push_to_peers(p, 2, pair, f);
                    }
                }
            }
        }

        // Park the prefetched residue-i data of the upcoming tile in the other staging buffer.
        if (upcoming >= 0) prefetch_commit(parity ^ 1);

        const long long elem0 = pair0 * G::kElemsPerPair;
        if (np == kTilePairs) {
            fence_proxy_async_smem();
            tile_sync<WPT>(slot);
            if (is_issuer) {
                if (p.l2_hint && p.l2_hint != 5) {  // 1 / 2: evict_first / evict_last for both tiles; 3 / 4: evict_first
                                                     // for one of them; 5: angle planes only; 6: tiles and angle planes
                    const uint64_t policy = l2_policy(p.l2_hint == 2 ? 2 : 1);
                    if (p.l2_hint != 4)
                        bulk_store_s2g_hint(p.dist + elem0, tile_f32, G::kDistBytes, policy);
                    else
                        bulk_store_s2g(p.dist + elem0, tile_f32, G::kDistBytes);
                    constexpr uint32_t kMaskTileBytes = kBool ? G::kMaskBytes : G::kDistBytes;
                    if (kBool || kF32) {
                        if (p.l2_hint != 3)
                            bulk_store_s2g_hint(static_cast<uint8_t*>(p.mask_out) + elem0 * (kBool ? 1 : 4), tile_mask,
                                                kMaskTileBytes, policy);
                        else
                            bulk_store_s2g(static_cast<uint8_t*>(p.mask_out) + elem0 * (kBool ? 1 : 4), tile_mask,
                                           kMaskTileBytes);
                    }
                } else {
                    bulk_store_s2g(p.dist + elem0, tile_f32, G::kDistBytes);
                    if (kBool) bulk_store_s2g(static_cast<uint8_t*>(p.mask_out) + elem0, tile_mask, G::kMaskBytes);
                    if (kF32) bulk_store_s2g(static_cast<float*>(p.mask_out) + elem0, tile_mask, G::kDistBytes);
                }
                bulk_commit();
                if (p.pace_ns) __nanosleep(p.pace_ns);
            }
        } else {
            // tail tile (num_pairs % 32 != 0): byte count is not 16-B granular, copy by hand
            tile_sync<WPT>(slot);
            const int n = np * G::kElemsPerPair;
            const int t2 = wsub * 32 + lane;
            for (int e = t2; e < n; e += 64) p.dist[elem0 + e] = tile_f32[e];
            if (kBool)
                for (int e = t2; e < n; e += 64) static_cast<uint8_t*>(p.mask_out)[elem0 + e] = tile_mask[e];
            if (kF32)
                for (int e = t2; e < n; e += 64)
                    static_cast<float*>(p.mask_out)[elem0 + e] = reinterpret_cast<const float*>(tile_mask)[e];
            tile_sync<WPT>(slot);
        }
        if constexpr (ANGLES) {
            // compact (B, L, L) copies of dist[..., CA, CA], [..., CB, CB], [..., N, O] (optional gather of compact
            // features): read back from the finished tile, which stays in shared memory until this buffer's next tile
            if (does_angles && (p.d_ca != nullptr || p.n_peers > 0) && !p.stores_only && lane < np) {
                const float* blk = tile_f32 + lane * G::kElemsPerPair;
                const float dca = blk[1 * A + 1], dcb = blk[4 * A + 4], dno = blk[0 * A + 3];
                if (p.d_ca != nullptr) {
                    p.d_ca[pair] = dca;
                    p.d_cb[pair] = dcb;
                    p.d_no[pair] = dno;
                }
                if (p.n_peers > 0) {
                    push_to_peers(p, 3, pair, dca);
                    push_to_peers(p, 4, pair, dcb);
                    push_to_peers(p, 5, pair, dno);
                }
            }
        }
        __syncwarp();  // staging buffer of the upcoming tile is complete
        mask_ballot = pf_mask_ballot;
        tile = upcoming;
        cur = nxt;
        cur_engine = nxt_engine;
        parity ^= 1;
    }
    // Shared memory must stay allocated until the engine has read the last tile.
    if (is_issuer) bulk_wait_all();
    __syncwarp();
}

template <int KIND, int SQRT, bool ANGLES>
int launch_sweep(const SweepParams& p, int slots_override, cudaStream_t stream) {
    constexpr int per_slot = sweep_slot_bytes<KIND>();
    constexpr int kMaxSmem = 227 * 1024;
    int slots = kMaxSmem / per_slot;
    if (slots > 4) slots = 4;
    if (slots_override > 0 && slots_override <= kMaxSmem / per_slot && slots_override <= 4) slots = slots_override;
    const int smem = slots * per_slot;
    auto kernel = pair_sweep_kernel<KIND, SQRT, ANGLES>;
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(pair_sweep_kernel)");
    const int sms = sm_count_for_current_device();
    if (sms < 0) return sms;
    long long ctas = (p.num_tiles + slots - 1) / slots;
    if (ctas > sms) ctas = sms;
    g_last_plan.path = 0;
    g_last_plan.lockstep = 1;
    g_last_plan.ctas = ctas;
    g_last_plan.tile_buffers = ctas * slots;
    g_last_plan.active_buffers = ctas * slots;
    g_last_plan.strip_stride = ctas * slots;
    g_last_plan.tile_pairs = kTilePairs;
    g_last_plan.sweep = 1;
    ++g_last_plan.launches;
    kernel<<<static_cast<unsigned>(ctas), slots * 64, smem, stream>>>(p);
    return check_launch("pair_sweep_kernel");
}

template <int KIND, bool ANGLES>
int launch_sweep_sqrt(const SweepParams& p, int sqrt_id, int slots_override, cudaStream_t stream) {
    switch (sqrt_id) {
        case kSqrtApproxFtz: return launch_sweep<KIND, kSqrtApproxFtz, ANGLES>(p, slots_override, stream);
        case kSqrtApprox: return launch_sweep<KIND, kSqrtApprox, ANGLES>(p, slots_override, stream);
        case kSqrtRn: return launch_sweep<KIND, kSqrtRn, ANGLES>(p, slots_override, stream);
        default: set_error("unknown sqrt mode %d", sqrt_id); return PS_ERR_BAD_DTYPE;
    }
}

}  // namespace

// Whether a request can take the linear-sweep kernel: A = 15, L >= 32 (a tile then touches at most two residue-i rows),
// distances requested, 16-byte aligned inputs and outputs.
bool pair_sweep_supported(const float* xyz, const void* atom_mask, int mask_dtype, const float* dist, const void* dist_mask,
                          int L, int A) {
    auto aligned = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    (void)mask_dtype;
    return A == kSweepA && L >= kTilePairs && dist != nullptr && aligned(xyz) && aligned(atom_mask) && aligned(dist) &&
           aligned(dist_mask);
}

// Peer buffers of the fused all-gather for the NEXT pair_sweep_impl call of this thread (set by the push entry point).
struct PushTarget {
    float* peer[kMaxPeers] = {};
    float* mc = nullptr;
    int n_peers = 0;
    long long plane = 0, offset = 0;
};
thread_local PushTarget g_push_target;

void pair_sweep_set_push_target(void* const* peers, int n_peers, void* multicast, long long plane, long long offset) {
    PushTarget t;
    for (int r = 0; r < n_peers && r < kMaxPeers; ++r) t.peer[r] = static_cast<float*>(peers[r]);
    t.mc = static_cast<float*>(multicast);
    t.n_peers = n_peers;
    t.plane = plane;
    t.offset = offset;
    g_push_target = t;
}

int pair_sweep_max_peers() { return kMaxPeers; }

int pair_sweep_impl(const float* xyz, const void* atom_mask, int mask_dtype, float* dist, void* dist_mask, float* omega,
                    float* theta, float* phi, float* d_ca, float* d_cb, float* d_no, int B, int L, int sqrt_id,
                    int slots_override, int stores_only, int pace_ns, int l2_hint, cudaStream_t stream) {
    SweepParams p;
    p.pace_ns = pace_ns;
    p.l2_hint = l2_hint;
    for (int r = 0; r < kMaxPeers; ++r) p.peer[r] = g_push_target.peer[r];
    p.mc = g_push_target.mc;
    p.n_peers = g_push_target.n_peers;
    p.peer_plane = g_push_target.plane;
    p.peer_offset = g_push_target.offset;
    g_push_target = PushTarget();  // one launch only
    p.xyz = xyz;
    p.atom_mask = atom_mask;
    p.dist = dist;
    p.mask_out = dist_mask;
    p.omega = omega;
    p.theta = theta;
    p.phi = phi;
    p.d_ca = d_ca;
    p.d_cb = d_cb;
    p.d_no = d_no;
    p.L = L;
    p.num_rows = static_cast<long long>(B) * L;
    p.num_pairs = p.num_rows * L;
    p.num_tiles = (p.num_pairs + kTilePairs - 1) / kTilePairs;
    p.stores_only = stores_only;
    const bool angles = omega || theta || phi || p.n_peers > 0;
    if (dist_mask == nullptr)
        return angles ? launch_sweep_sqrt<kSweepDistOnly, true>(p, sqrt_id, slots_override, stream)
                      : launch_sweep_sqrt<kSweepDistOnly, false>(p, sqrt_id, slots_override, stream);
    if (mask_dtype == PS_MASK_BOOL)
        return angles ? launch_sweep_sqrt<kSweepDistBool, true>(p, sqrt_id, slots_override, stream)
                      : launch_sweep_sqrt<kSweepDistBool, false>(p, sqrt_id, slots_override, stream);
    return angles ? launch_sweep<kSweepDistF32, kSqrtApproxFtz, true>(p, slots_override, stream)
                  : launch_sweep<kSweepDistF32, kSqrtApproxFtz, false>(p, slots_override, stream);
}

}  // namespace ps
