// K3 — per-residue backbone features: phi/psi/omega dihedrals with chain-terminal handling, the
// 3-bool validity mask, and Gram-Schmidt frames; plus the flat-list geometry free functions.
//
// Replaces StructureBatch.backbone_dihedrals (protstruc/protstruc.py:486-541) with
// get_n_terminal_mask / get_c_terminal_mask (:435-453), StructureBatch.backbone_orientations
// (:543-571) -> geometry.gram_schmidt (protstruc/geometry.py:413-439), and geometry.angle /
// geometry.dihedral on (n,3) point lists (protstruc/geometry.py:39-124).
//
// Roofline: HBM read+write, tiny (~41 B read + 51 B written per residue); one thread per residue,
// neighbours' atoms come from L1.  Latency/launch-bound at the reference's sizes.

#include "common.cuh"

namespace ps {

namespace {

// torch's float `!=`: NaN compares unequal to everything, itself included.
__device__ __forceinline__ bool chain_differs(float a, float b) { return a != b; }

__global__ void __launch_bounds__(256) backbone_kernel(
    const float* __restrict__ xyz, const uint8_t* __restrict__ residue_mask,
    const float* __restrict__ chain_idx, int L, int A, int a1, int a2, int a3,
    float* __restrict__ dihedrals, uint8_t* __restrict__ dihedral_mask, float* __restrict__ frames,
    long long total) {
    const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= total) return;
    const int l = static_cast<int>(r % L);
    const float* __restrict__ x = xyz + r * A * 3;

    if (dihedrals) {
        const float nan = __int_as_float(0x7fc00000);
        const float ch = __ldg(chain_idx + r);
        const float ch_prev = l > 0 ? __ldg(chain_idx + r - 1) : nan;
        const float ch_next = l < L - 1 ? __ldg(chain_idx + r + 1) : nan;
        const bool rm = __ldg(residue_mask + r) != 0;
        const bool nterm = chain_differs(ch_prev, ch) && rm;
        const bool cterm = chain_differs(ch, ch_next) && rm;

        const V3 n = ld3(x + 0), ca = ld3(x + 3), c = ld3(x + 6);
        float phi = 0.f, psi = 0.f, omega = 0.f;
        // The reference evaluates every interior dihedral and then zero-fills the terminal ones
        // (phi[nterm] = 0, psi[cterm] = omega[cterm] = 0); skipping the evaluation is equivalent.
        if (l > 0 && !nterm) {
            const V3 c_prev = ld3(x - A * 3 + 6);
            phi = dihedral4(c_prev, n, ca, c);
        }
        if (l < L - 1 && !cterm) {
            const V3 n_next = ld3(x + A * 3 + 0);
            const V3 ca_next = ld3(x + A * 3 + 3);
            psi = dihedral4(n, ca, c, n_next);
            omega = dihedral4(ca, c, n_next, ca_next);
        }
        dihedrals[r * 3 + 0] = phi;
        dihedrals[r * 3 + 1] = psi;
        dihedrals[r * 3 + 2] = omega;
        dihedral_mask[r * 3 + 0] = (!nterm && rm) ? 1 : 0;
        dihedral_mask[r * 3 + 1] = (!cterm && rm) ? 1 : 0;
        dihedral_mask[r * 3 + 2] = (!cterm && rm) ? 1 : 0;
    }

    if (frames) {
        const V3 a = ld3(x + a1 * 3), b = ld3(x + a2 * 3), c = ld3(x + a3 * 3);
        const V3 v1 = sub3(c, b);
        const V3 e1 = div3(v1, norm3(v1));
        const V3 v2 = sub3(a, b);
        const V3 u2 = sub3(v2, scale3(e1, dot3(e1, v2)));
        const V3 e2 = div3(u2, norm3(u2));
        const V3 e3 = cross3(e1, e2);
        float* __restrict__ f = frames + r * 9;  // R[row][col], column k = e_k
        f[0] = e1.x; f[1] = e2.x; f[2] = e3.x;
        f[3] = e1.y; f[4] = e2.y; f[5] = e3.y;
        f[6] = e1.z; f[7] = e2.z; f[8] = e3.z;
    }
}

constexpr float kRadToDeg = 57.29577951308232f;  // 180/pi rounded to fp32, as torch.rad2deg uses

__global__ void __launch_bounds__(256) geom_angle_kernel(const float* __restrict__ a,
                                                         const float* __restrict__ b,
                                                         const float* __restrict__ c, long long n,
                                                         int to_degree, float* __restrict__ out) {
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n) return;
    float v = angle3(ld3(a + t * 3), ld3(b + t * 3), ld3(c + t * 3));
    if (to_degree) v = __fmul_rn(v, kRadToDeg);
    out[t] = v;
}

__global__ void __launch_bounds__(256) geom_dihedral_kernel(
    const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
    const float* __restrict__ d, long long n, int to_degree, float* __restrict__ out) {
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n) return;
    float v = dihedral4(ld3(a + t * 3), ld3(b + t * 3), ld3(c + t * 3), ld3(d + t * 3));
    if (to_degree) v = __fmul_rn(v, kRadToDeg);
    out[t] = v;
}

__global__ void __launch_bounds__(256) geom_gram_schmidt_kernel(const float* __restrict__ pa,
                                                                const float* __restrict__ pb,
                                                                const float* __restrict__ pc,
                                                                long long n,
                                                                float* __restrict__ out) {
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const V3 a = ld3(pa + t * 3), b = ld3(pb + t * 3), c = ld3(pc + t * 3);
    const V3 v1 = sub3(c, b);
    const V3 e1 = div3(v1, norm3(v1));
    const V3 v2 = sub3(a, b);
    const V3 u2 = sub3(v2, scale3(e1, dot3(e1, v2)));
    const V3 e2 = div3(u2, norm3(u2));
    const V3 e3 = cross3(e1, e2);
    float* __restrict__ f = out + t * 9;
    f[0] = e1.x; f[1] = e2.x; f[2] = e3.x;
    f[3] = e1.y; f[4] = e2.y; f[5] = e3.y;
    f[6] = e1.z; f[7] = e2.z; f[8] = e3.z;
}

unsigned blocks_for(long long n) { return static_cast<unsigned>((n + 255) / 256); }

}  // namespace

int backbone_impl(const float* xyz, const uint8_t* residue_mask, const float* chain_idx, int B,
                  int L, int A, int a1, int a2, int a3, float* dihedrals, uint8_t* dihedral_mask,
                  float* frames, cudaStream_t stream) {
    PS_REQUIRE(B > 0 && L > 0 && A > 0, PS_ERR_BAD_SHAPE, "backbone: B=%d L=%d A=%d must be > 0", B,
               L, A);
    PS_REQUIRE(xyz, PS_ERR_NULL_POINTER, "backbone: xyz is NULL");
    PS_REQUIRE((dihedrals == nullptr) == (dihedral_mask == nullptr), PS_ERR_NULL_POINTER,
               "backbone: dihedrals and dihedral_mask must both be given or both be NULL");
    PS_REQUIRE(dihedrals || frames, PS_ERR_NULL_POINTER, "backbone: no output requested");
    if (dihedrals) {
        PS_REQUIRE(residue_mask && chain_idx, PS_ERR_NULL_POINTER,
                   "backbone: dihedrals need residue_mask and chain_idx");
        PS_REQUIRE(A >= 3, PS_ERR_BAD_SHAPE, "backbone: dihedrals need N, CA, C slots (A >= 3)");
    }
    if (frames) {
        PS_REQUIRE(a1 >= 0 && a1 < A && a2 >= 0 && a2 < A && a3 >= 0 && a3 < A, PS_ERR_BAD_SLOT,
                   "backbone: frame slots (%d,%d,%d) outside [0,%d)", a1, a2, a3, A);
    }
    const long long total = static_cast<long long>(B) * L;
    backbone_kernel<<<blocks_for(total), 256, 0, stream>>>(xyz, residue_mask, chain_idx, L, A, a1,
                                                           a2, a3, dihedrals, dihedral_mask, frames,
                                                           total);
    return check_launch("backbone_kernel");
}

int geom_angle_impl(const float* a, const float* b, const float* c, long long n, int to_degree,
                    float* out, cudaStream_t stream) {
    PS_REQUIRE(n >= 0, PS_ERR_BAD_SHAPE, "geom_angle: n=%lld", n);
    if (n == 0) return PS_OK;
    PS_REQUIRE(a && b && c && out, PS_ERR_NULL_POINTER, "geom_angle: NULL pointer");
    geom_angle_kernel<<<blocks_for(n), 256, 0, stream>>>(a, b, c, n, to_degree, out);
    return check_launch("geom_angle_kernel");
}

int geom_dihedral_impl(const float* a, const float* b, const float* c, const float* d, long long n,
                       int to_degree, float* out, cudaStream_t stream) {
    PS_REQUIRE(n >= 0, PS_ERR_BAD_SHAPE, "geom_dihedral: n=%lld", n);
    if (n == 0) return PS_OK;
    PS_REQUIRE(a && b && c && d && out, PS_ERR_NULL_POINTER, "geom_dihedral: NULL pointer");
    geom_dihedral_kernel<<<blocks_for(n), 256, 0, stream>>>(a, b, c, d, n, to_degree, out);
    return check_launch("geom_dihedral_kernel");
}

int geom_gram_schmidt_impl(const float* a, const float* b, const float* c, long long n, float* out,
                           cudaStream_t stream) {
    PS_REQUIRE(n >= 0, PS_ERR_BAD_SHAPE, "geom_gram_schmidt: n=%lld", n);
    if (n == 0) return PS_OK;
    PS_REQUIRE(a && b && c && out, PS_ERR_NULL_POINTER, "geom_gram_schmidt: NULL pointer");
    geom_gram_schmidt_kernel<<<blocks_for(n), 256, 0, stream>>>(a, b, c, n, out);
    return check_launch("geom_gram_schmidt_kernel");
}

}  // namespace ps
