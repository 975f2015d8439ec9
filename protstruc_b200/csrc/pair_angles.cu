// K2 — pairwise dihedral / planar angles between residues, and the fused trRosetta triple.
//
// Replaces StructureBatch.pairwise_dihedrals / pairwise_planar_angles
// (protstruc/protstruc.py:620-660) — including the (B, L^2, n, 3) gather the reference
// materialises in _pairwise_xyz (:589-618) — and the three angle calls of
// inter_residue_geometry (:810-815).
//
// Roofline: NOT HBM-bound.  A pair costs ~60 geometric + ~100 transcendental lane-instructions
// (atan2f / acosf, IEEE division and sqrt) against 4 B (12 B fused) written, so the binding roof
// is FP32/SFU issue; HBM fraction is reported but is not the target.  The kernel therefore
// minimises instructions: per-(b,i) invariants are hoisted out of the j loop, the five atoms of
// residue j a thread needs are loaded once, and lanes walk consecutive j so loads/stores coalesce.
//
// Grid: blockIdx.x = (b*L + i) row, threads stride over j.  Rows of L floats are written fully
// coalesced.

#include "common.cuh"

namespace ps {

namespace {

struct SlotList {
    int s[4];   // atom slot of point k
    int from_j[4];  // 0: residue i, 1: residue j
    int n;
};

// NI = how many of the points come from residue i (the first NI of the list).  It is a template parameter so that
// the points of residue i — and everything dihedral4 / angle3 derive from them alone (b0 for NI >= 2; b1, the normal
// n1 and 1 / |b1| for NI >= 3) — are loop invariants the compiler hoists out of the j loop, and so that the point
// array is indexed with compile-time constants only.
template <int KIND, int NI>
__global__ void __launch_bounds__(256) pair_angles_kernel(const float* __restrict__ xyz,
                                                          float* __restrict__ out, int L, int A,
                                                          SlotList sl, long long rows) {
    constexpr int N = KIND == PS_ANGLE_DIHEDRAL ? 4 : 3;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const long long b = row / L;
        const float* __restrict__ xi = xyz + row * A * 3;
        const float* __restrict__ xb = xyz + b * L * A * 3;
        float* __restrict__ out_row = out + row * L;
        V3 pt[4];
#pragma unroll
        for (int k = 0; k < N; ++k) pt[k] = V3{0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < NI; ++k) pt[k] = ld3(xi + sl.s[k] * 3);
        const int stride = A * 3;
        for (int j = threadIdx.x; j < L; j += blockDim.x) {
            const float* __restrict__ xj = xb + j * stride;
#pragma unroll
            for (int k = NI; k < N; ++k) pt[k] = ld3(xj + sl.s[k] * 3);
            float v;
            if (KIND == PS_ANGLE_DIHEDRAL)
                v = dihedral4(pt[0], pt[1], pt[2], pt[3]);
            else
                v = angle3(pt[0], pt[1], pt[2]);
            out_row[j] = v;
        }
    }
}

// omega = dihedral(CA_i, CB_i, CA_j, CB_j); theta = dihedral(N_i, CA_i, CB_i, CB_j);
// phi = angle(CA_i, CB_i, CB_j).  Shares loads and the i-only sub-expressions.
template <bool VIRTUAL_CB>
__global__ void __launch_bounds__(256) trrosetta_kernel(const float* __restrict__ xyz,
                                                        float* __restrict__ omega,
                                                        float* __restrict__ theta,
                                                        float* __restrict__ phi, int L, int A,
                                                        long long rows) {
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const long long b = row / L;
        const float* __restrict__ xi = xyz + row * A * 3;
        const float* __restrict__ xb = xyz + b * L * A * 3;
        const V3 n_i = ld3(xi + 0), ca_i = ld3(xi + 3);
        const V3 cb_i = VIRTUAL_CB ? virtual_cb(n_i, ca_i, ld3(xi + 6)) : ld3(xi + 12);
        const TripleRowSide side = triple_row_side(n_i, ca_i, cb_i);
        for (int j = threadIdx.x; j < L; j += blockDim.x) {
            const float* __restrict__ xj = xb + static_cast<long long>(j) * A * 3;
            const V3 ca_j = ld3(xj + 3);
            const V3 cb_j = VIRTUAL_CB ? virtual_cb(ld3(xj + 0), ca_j, ld3(xj + 6)) : ld3(xj + 12);
            const long long o = row * L + j;
            float w, t, f;
            trrosetta_triple(side, ca_j, cb_j, omega != nullptr, theta != nullptr, phi != nullptr, w, t, f);
            if (omega) omega[o] = w;
            if (theta) theta[o] = t;
            if (phi) phi[o] = f;
        }
    }
}

int grid_for_rows(long long rows, int* grid) {
    const int sms = sm_count_for_current_device();
    if (sms < 0) return sms;
    long long g = rows;
    const long long cap = static_cast<long long>(sms) * 64;
    if (g > cap) g = cap;
    *grid = static_cast<int>(g);
    return PS_OK;
}

// Threads per (b, i) row.  Every thread recomputes the row-only part of the angles (~130 issue slots with
// two IEEE reciprocals / square roots), so a thread should own several j: L / 8 threads, 32..256.
int threads_for_L(int L) {
    int t = ((L + 7) / 8 + 31) / 32 * 32;
    if (t < 32) t = 32;
    if (t > 256) t = 256;
    return t;
}

}  // namespace

int pair_angles_impl(const float* xyz, int B, int L, int A, const int* slots_i, int n_i,
                     const int* slots_j, int n_j, int kind, float* out, cudaStream_t stream) {
    PS_REQUIRE(B > 0 && L > 0 && A > 0, PS_ERR_BAD_SHAPE, "pair_angles: B=%d L=%d A=%d must be > 0",
               B, L, A);
    PS_REQUIRE(xyz && out, PS_ERR_NULL_POINTER, "pair_angles: NULL pointer");
    PS_REQUIRE(kind == PS_ANGLE_DIHEDRAL || kind == PS_ANGLE_PLANAR, PS_ERR_BAD_DTYPE,
               "pair_angles: unknown kind %d", kind);
    const int need = kind == PS_ANGLE_DIHEDRAL ? 4 : 3;
    PS_REQUIRE(n_i >= 0 && n_j >= 0 && n_i + n_j == need, PS_ERR_BAD_SHAPE,
               "pair_angles: kind %d needs %d atoms in total, got %d + %d", kind, need, n_i, n_j);
    PS_REQUIRE((n_i == 0 || slots_i) && (n_j == 0 || slots_j), PS_ERR_NULL_POINTER,
               "pair_angles: NULL slot list");
    SlotList sl;
    sl.n = need;
    for (int k = 0; k < 4; ++k) {
        sl.s[k] = 0;
        sl.from_j[k] = 0;
    }
    for (int k = 0; k < need; ++k) {
        const bool from_j = k >= n_i;
        const int s = from_j ? slots_j[k - n_i] : slots_i[k];
        PS_REQUIRE(s >= 0 && s < A, PS_ERR_BAD_SLOT, "pair_angles: slot %d outside [0,%d)", s, A);
        sl.s[k] = s;
        sl.from_j[k] = from_j ? 1 : 0;
    }
    const long long rows = static_cast<long long>(B) * L;
    int grid = 0;
    int rc = grid_for_rows(rows, &grid);
    if (rc != PS_OK) return rc;
    const int threads = threads_for_L(L);
    PS_REQUIRE(static_cast<long long>(L) * A * 3 < (1ll << 31), PS_ERR_BAD_SHAPE,
               "pair_angles: L*A*3=%lld floats per structure exceed 2^31", static_cast<long long>(L) * A * 3);
#define PS_PAIR_ANGLES(KIND, NI) \
    pair_angles_kernel<KIND, NI><<<grid, threads, 0, stream>>>(xyz, out, L, A, sl, rows)
    if (kind == PS_ANGLE_DIHEDRAL) {
        switch (n_i) {
            case 0: PS_PAIR_ANGLES(PS_ANGLE_DIHEDRAL, 0); break;
            case 1: PS_PAIR_ANGLES(PS_ANGLE_DIHEDRAL, 1); break;
            case 2: PS_PAIR_ANGLES(PS_ANGLE_DIHEDRAL, 2); break;
            case 3: PS_PAIR_ANGLES(PS_ANGLE_DIHEDRAL, 3); break;
            default: PS_PAIR_ANGLES(PS_ANGLE_DIHEDRAL, 4); break;
        }
    } else {
        switch (n_i) {
            case 0: PS_PAIR_ANGLES(PS_ANGLE_PLANAR, 0); break;
            case 1: PS_PAIR_ANGLES(PS_ANGLE_PLANAR, 1); break;
            case 2: PS_PAIR_ANGLES(PS_ANGLE_PLANAR, 2); break;
            default: PS_PAIR_ANGLES(PS_ANGLE_PLANAR, 3); break;
        }
    }
#undef PS_PAIR_ANGLES
    return check_launch("pair_angles_kernel");
}

int trrosetta_angles_impl(const float* xyz, int B, int L, int A, int use_virtual_cb, float* omega,
                          float* theta, float* phi, cudaStream_t stream) {
    PS_REQUIRE(B > 0 && L > 0 && A > 0, PS_ERR_BAD_SHAPE,
               "trrosetta_angles: B=%d L=%d A=%d must be > 0", B, L, A);
    PS_REQUIRE(xyz, PS_ERR_NULL_POINTER, "trrosetta_angles: xyz is NULL");
    PS_REQUIRE(omega || theta || phi, PS_ERR_NULL_POINTER, "trrosetta_angles: no output requested");
    PS_REQUIRE(A >= (use_virtual_cb ? 3 : 5), PS_ERR_BAD_SHAPE,
               "trrosetta_angles: A=%d has no %s slot", A, use_virtual_cb ? "C" : "CB");
    const long long rows = static_cast<long long>(B) * L;
    int grid = 0;
    int rc = grid_for_rows(rows, &grid);
    if (rc != PS_OK) return rc;
    const int threads = threads_for_L(L);
    if (use_virtual_cb)
        trrosetta_kernel<true><<<grid, threads, 0, stream>>>(xyz, omega, theta, phi, L, A, rows);
    else
        trrosetta_kernel<false><<<grid, threads, 0, stream>>>(xyz, omega, theta, phi, L, A, rows);
    return check_launch("trrosetta_kernel");
}

}  // namespace ps
