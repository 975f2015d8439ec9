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

int translate_impl(const float* xyz, const float* t, int t_rows, int B, int L, int A, float* xyz_out,
                   cudaStream_t stream);  // stats.cu

namespace {

struct Frame {
    V3 e1, e2, e3;
};

// geometry.gram_schmidt (protstruc/geometry.py:430-439), cross product on the last axis.
__device__ __forceinline__ Frame gram_schmidt_frame(V3 a, V3 b, V3 c) {
    Frame f;
    const V3 v1 = sub3(c, b);
    f.e1 = div3(v1, norm3(v1));
    const V3 v2 = sub3(a, b);
    const V3 u2 = sub3(v2, scale3(f.e1, dot3(f.e1, v2)));
    f.e2 = div3(u2, norm3(u2));
    f.e3 = cross3(f.e1, f.e2);
    return f;
}

__device__ __forceinline__ void store_frame(float* __restrict__ f, const Frame& fr) {
    // R[row][col], column k = e_k
    f[0] = fr.e1.x; f[1] = fr.e2.x; f[2] = fr.e3.x;
    f[3] = fr.e1.y; f[4] = fr.e2.y; f[5] = fr.e3.y;
    f[6] = fr.e1.z; f[7] = fr.e2.z; f[8] = fr.e3.z;
}

// torch's float `!=`: NaN compares unequal to everything, itself included.
__device__ __forceinline__ bool chain_differs(float a, float b) { return a != b; }

__global__ void __launch_bounds__(256) backbone_kernel(
    const float* __restrict__ xyz, const uint8_t* __restrict__ residue_mask,
    const float* __restrict__ chain_idx, int L, int A, int a1, int a2, int a3,
    float* __restrict__ dihedrals, uint8_t* __restrict__ dihedral_mask, float* __restrict__ frames,
    long long total) {
    const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= total) return;
    const int l = static_cast<int>(r - index_div(r, L, total <= 0xFFFFFFFFll) * L);
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
        store_frame(frames + r * 9, gram_schmidt_frame(ld3(x + a1 * 3), ld3(x + a2 * 3), ld3(x + a3 * 3)));
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
    store_frame(out + t * 9, gram_schmidt_frame(ld3(pa + t * 3), ld3(pb + t * 3), ld3(pc + t * 3)));
}

// geometry.dot / norm / unit (protstruc/geometry.py:24-36) on n rows of D components (D = 3 for points; any D works).
// One thread per row.  dot: separately rounded products added left to right from +0, like ATen's sum over a short last
// axis; norm: ATen's vector norm, a fused multiply-add chain under an IEEE square root (see norm3 in common.cuh);
// unit: IEEE division of every component by the norm (0 / 0 = NaN for a zero vector, like the reference).
// op: 0 = dot(x, y) -> out (n), 1 = norm(x) -> out (n), 2 = unit(x) -> out (n, D).
__global__ void __launch_bounds__(256) geom_rowwise_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                           long long n, int D, int op, float* __restrict__ out) {
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float* __restrict__ xr = x + t * D;
    if (op == 0) {
        const float* __restrict__ yr = y + t * D;
        float acc = 0.0f;
        for (int k = 0; k < D; ++k) acc = __fadd_rn(acc, __fmul_rn(__ldg(xr + k), __ldg(yr + k)));
        out[t] = acc;
        return;
    }
    float ss = 0.0f;
    for (int k = 0; k < D; ++k) {
        const float v = __ldg(xr + k);
        ss = k == 0 ? __fmul_rn(v, v) : __fmaf_rn(v, v, ss);
    }
    const float nrm = __fsqrt_rn(ss);
    if (op == 1) {
        out[t] = nrm;
        return;
    }
    for (int k = 0; k < D; ++k) out[t * D + k] = __fdiv_rn(__ldg(xr + k), nrm);
}

// ---- rigid-frame family ---------------------------------------------------------
// get_local_xyz (protstruc/protstruc.py:347-362): local = R^T x - CA, with R the residue's Gram-Schmidt
// frame and CA the residue's GLOBAL alpha-carbon (the reference subtracts it after rotating; kept).
// A CTA takes kLocalResidues residues: their frames (and CA) are computed once, one thread per residue, and parked
// in shared memory; then one thread per (residue, atom) applies them.  (The first version recomputed the
// Gram-Schmidt frame in every atom thread: A-fold redundant square roots and divisions, issue-bound.)
constexpr int kLocalResidues = 64;

__global__ void __launch_bounds__(256) local_xyz_kernel(const float* __restrict__ xyz, int A, int a1,
                                                        int a2, int a3, int ca_slot, long long num_residues,
                                                        float* __restrict__ out) {
    __shared__ float frame[kLocalResidues][12];  // e1, e2, e3, CA
    const long long r0 = static_cast<long long>(blockIdx.x) * kLocalResidues;
    const long long left = num_residues - r0;
    const int nres = left < kLocalResidues ? static_cast<int>(left) : kLocalResidues;
    if (threadIdx.x < nres) {
        const float* __restrict__ x = xyz + (r0 + threadIdx.x) * A * 3;
        const Frame f = gram_schmidt_frame(ld3(x + a1 * 3), ld3(x + a2 * 3), ld3(x + a3 * 3));
        const V3 ca = ld3(x + ca_slot * 3);
        float* dst = frame[threadIdx.x];
        dst[0] = f.e1.x; dst[1] = f.e1.y; dst[2] = f.e1.z;
        dst[3] = f.e2.x; dst[4] = f.e2.y; dst[5] = f.e2.z;
        dst[6] = f.e3.x; dst[7] = f.e3.y; dst[8] = f.e3.z;
        dst[9] = ca.x; dst[10] = ca.y; dst[11] = ca.z;
    }
    __syncthreads();
    const float* __restrict__ xin = xyz + r0 * A * 3;
    float* __restrict__ xout = out + r0 * A * 3;
    const int atoms = nres * A;
    for (int t = threadIdx.x; t < atoms; t += blockDim.x) {
        const int r = t / A;
        const float* fr = frame[r];
        const V3 p = ld3(xin + t * 3);
        // einsum("bnaji,bnaj->bnai"): out_i = sum_j R[j][i] x_j = e_i . x
        xout[t * 3 + 0] = __fsub_rn(dot3(V3{fr[0], fr[1], fr[2]}, p), fr[9]);
        xout[t * 3 + 1] = __fsub_rn(dot3(V3{fr[3], fr[4], fr[5]}, p), fr[10]);
        xout[t * 3 + 2] = __fsub_rn(dot3(V3{fr[6], fr[7], fr[8]}, p), fr[11]);
    }
}

// rotate (protstruc/protstruc.py:681-694): x' = R_b x with one (3,3) matrix per structure (or one for all).
__global__ void __launch_bounds__(256) rotate_kernel(const float* __restrict__ xyz,
                                                     const float* __restrict__ rot, int rot_rows,
                                                     long long atoms_per_struct, long long total,
                                                     float* __restrict__ out) {
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const float* __restrict__ m =
        rot + (rot_rows == 1 ? 0 : index_div(t, atoms_per_struct, total <= 0xFFFFFFFFll) * 9);
    const V3 p = ld3(xyz + t * 3);
    out[t * 3 + 0] = dot3(V3{__ldg(m + 0), __ldg(m + 1), __ldg(m + 2)}, p);
    out[t * 3 + 1] = dot3(V3{__ldg(m + 3), __ldg(m + 4), __ldg(m + 5)}, p);
    out[t * 3 + 2] = dot3(V3{__ldg(m + 6), __ldg(m + 7), __ldg(m + 8)}, p);
}

// The same map for FOUR consecutive atoms per thread (structures of a multiple of 4 atoms, 16-byte aligned arrays):
// three 128-bit loads and stores instead of twelve 32-bit ones, the matrix fetched once per four atoms, one index
// division per thread (12.4 -> ~9 us for 23.6 MB under ncu; ATen's elementwise kernels move the same bytes in 7.4-8.2).
__global__ void __launch_bounds__(256) rotate_quad_kernel(const float* __restrict__ xyz, const float* __restrict__ rot,
                                                          int rot_rows, long long quads_per_struct, long long total_quads,
                                                          float* __restrict__ out) {
    const long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q >= total_quads) return;
    const float* __restrict__ m =
        rot + (rot_rows == 1 ? 0 : index_div(q, quads_per_struct, total_quads <= 0xFFFFFFFFll) * 9);
    const V3 r0{__ldg(m + 0), __ldg(m + 1), __ldg(m + 2)}, r1{__ldg(m + 3), __ldg(m + 4), __ldg(m + 5)},
        r2{__ldg(m + 6), __ldg(m + 7), __ldg(m + 8)};
    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(xyz) + 3 * q;
    const float4 a = __ldg(x4), b = __ldg(x4 + 1), c = __ldg(x4 + 2);
    const V3 p0{a.x, a.y, a.z}, p1{a.w, b.x, b.y}, p2{b.z, b.w, c.x}, p3{c.y, c.z, c.w};
    float4* __restrict__ o4 = reinterpret_cast<float4*>(out) + 3 * q;
    o4[0] = make_float4(dot3(r0, p0), dot3(r1, p0), dot3(r2, p0), dot3(r0, p1));
    o4[1] = make_float4(dot3(r1, p1), dot3(r2, p1), dot3(r0, p2), dot3(r1, p2));
    o4[2] = make_float4(dot3(r2, p2), dot3(r0, p3), dot3(r1, p3), dot3(r2, p3));
}

// from_backbone_orientations_translations (protstruc/protstruc.py:289-314): atom a < n_ideal is
// R_r ideal[a] + t_r, the remaining slots are zero; the mask is 1 for the placed atoms (fp32).
__global__ void __launch_bounds__(256) frames_to_backbone_kernel(
    const float* __restrict__ orientations, const float* __restrict__ translations,
    const float* __restrict__ ideal, int n_ideal, int A, long long total, float* __restrict__ xyz,
    float* __restrict__ atom_mask) {
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const long long r = index_div(t, A, total <= 0xFFFFFFFFll);
    const int a = static_cast<int>(t - r * A);
    V3 o{0.f, 0.f, 0.f};
    float m = 0.f;
    if (a < n_ideal) {
        const float* __restrict__ R = orientations + r * 9;
        const V3 q = ld3(ideal + a * 3);
        const V3 tr = ld3(translations + r * 3);
        o.x = __fadd_rn(dot3(V3{__ldg(R + 0), __ldg(R + 1), __ldg(R + 2)}, q), tr.x);
        o.y = __fadd_rn(dot3(V3{__ldg(R + 3), __ldg(R + 4), __ldg(R + 5)}, q), tr.y);
        o.z = __fadd_rn(dot3(V3{__ldg(R + 6), __ldg(R + 7), __ldg(R + 8)}, q), tr.z);
        m = 1.f;
    }
    xyz[t * 3 + 0] = o.x;
    xyz[t * 3 + 1] = o.y;
    xyz[t * 3 + 2] = o.z;
    atom_mask[t] = m;
}

// translate (protstruc/protstruc.py:662-679): x += t with a broadcastable translation given by its
// element strides over (structure, residue, atom); stride 0 = broadcast.
// FAST: the residue index comes from a multiply-high by a host-computed reciprocal of the row length (exact while
// e * row < 2^32) and the translation offset is 32-bit arithmetic; the general flavour divides and uses 64-bit
// strides.
template <bool FAST>
// xyz and out may be the SAME buffer (StructureBatch.translate works in place like the reference's `+=`), so neither
// is declared __restrict__: every thread reads an element before it writes that same element and touches no other.
__global__ void __launch_bounds__(256) translate_bcast_kernel(const float* xyz,
                                                              const float* __restrict__ tr,
                                                              long long sb, long long sl, long long sa,
                                                              int B, int A, int per_b, unsigned row_magic,
                                                              float* out) {
    // 2-D grid: blockIdx.y walks the structures, x the structure's floats; (l, a, axis) from the 32-bit offset
    // inside the structure
    const unsigned row = static_cast<unsigned>(A) * 3u;
    const unsigned sl32 = static_cast<unsigned>(sl), sa32 = static_cast<unsigned>(sa);
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const float* xb = xyz + static_cast<long long>(b) * per_b;
        float* ob = out + static_cast<long long>(b) * per_b;
        const float* __restrict__ tb = tr + b * sb;
        const unsigned n = static_cast<unsigned>(per_b), step = gridDim.x * blockDim.x;
        for (unsigned e0 = blockIdx.x * blockDim.x + threadIdx.x; e0 < n; e0 += 4 * step) {
            float v[4], w[4];  // four independent load pairs in flight per thread
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned e = e0 + u * step;
                v[u] = w[u] = 0.f;
                if (e < n) {
                    const unsigned l = FAST ? __umulhi(e, row_magic) : e / row;
                    const unsigned r = e - l * row;
                    const unsigned a = r / 3u;
                    const unsigned k = r - a * 3u;
                    v[u] = xb[e];
                    w[u] = FAST ? __ldg(tb + (l * sl32 + a * sa32 + k)) : __ldg(tb + l * sl + a * sa + k);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (e0 + u * step < n) ob[e0 + u * step] = __fadd_rn(v[u], w[u]);
        }
    }
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

int local_xyz_impl(const float* xyz, int B, int L, int A, int a1, int a2, int a3, int ca_slot,
                   float* out, cudaStream_t stream) {
    PS_REQUIRE(B > 0 && L > 0 && A > 0, PS_ERR_BAD_SHAPE, "local_xyz: B=%d L=%d A=%d must be > 0", B, L, A);
    PS_REQUIRE(xyz && out, PS_ERR_NULL_POINTER, "local_xyz: NULL pointer");
    PS_REQUIRE(a1 >= 0 && a1 < A && a2 >= 0 && a2 < A && a3 >= 0 && a3 < A && ca_slot >= 0 && ca_slot < A,
               PS_ERR_BAD_SLOT, "local_xyz: slot outside [0,%d)", A);
    const long long residues = static_cast<long long>(B) * L;
    const long long blocks = (residues + kLocalResidues - 1) / kLocalResidues;
    PS_REQUIRE(blocks < (1ll << 31), PS_ERR_BAD_SHAPE, "local_xyz: %lld residues", residues);
    local_xyz_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(xyz, A, a1, a2, a3, ca_slot, residues, out);
    return check_launch("local_xyz_kernel");
}

int rotate_impl(const float* xyz, const float* rot, int rot_rows, int B, int L, int A, float* out,
                cudaStream_t stream) {
    PS_REQUIRE(B > 0 && L > 0 && A > 0, PS_ERR_BAD_SHAPE, "rotate: B=%d L=%d A=%d must be > 0", B, L, A);
    PS_REQUIRE(xyz && rot && out, PS_ERR_NULL_POINTER, "rotate: NULL pointer");
    PS_REQUIRE(rot_rows == 1 || rot_rows == B, PS_ERR_BAD_SHAPE, "rotate: %d matrices for %d structures",
               rot_rows, B);
    PS_REQUIRE(xyz != out, PS_ERR_MISALIGNED, "rotate: in-place rotation is not supported");
    const long long per = static_cast<long long>(L) * A;
    if (per % 4 == 0 && ((reinterpret_cast<uintptr_t>(xyz) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0) {
        rotate_quad_kernel<<<blocks_for(per / 4 * B), 256, 0, stream>>>(xyz, rot, rot_rows, per / 4, per / 4 * B, out);
        return check_launch("rotate_quad_kernel");
    }
    rotate_kernel<<<blocks_for(per * B), 256, 0, stream>>>(xyz, rot, rot_rows, per, per * B, out);
    return check_launch("rotate_kernel");
}

int frames_to_backbone_impl(const float* orientations, const float* translations, const float* ideal,
                            int n_ideal, int B, int L, int A, float* xyz, float* atom_mask,
                            cudaStream_t stream) {
    PS_REQUIRE(B > 0 && L > 0 && A > 0, PS_ERR_BAD_SHAPE, "frames_to_backbone: B=%d L=%d A=%d", B, L, A);
    PS_REQUIRE(orientations && translations && ideal && xyz && atom_mask, PS_ERR_NULL_POINTER,
               "frames_to_backbone: NULL pointer");
    PS_REQUIRE(n_ideal > 0 && n_ideal <= A, PS_ERR_BAD_SHAPE, "frames_to_backbone: %d ideal atoms, A=%d",
               n_ideal, A);
    const long long total = static_cast<long long>(B) * L * A;
    frames_to_backbone_kernel<<<blocks_for(total), 256, 0, stream>>>(orientations, translations, ideal,
                                                                     n_ideal, A, total, xyz, atom_mask);
    return check_launch("frames_to_backbone_kernel");
}

int translate_bcast_impl(const float* xyz, const float* tr, long long sb, long long sl, long long sa,
                         int B, int L, int A, float* out, cudaStream_t stream) {
    PS_REQUIRE(B > 0 && L > 0 && A > 0, PS_ERR_BAD_SHAPE, "translate: B=%d L=%d A=%d must be > 0", B, L, A);
    PS_REQUIRE(xyz && tr && out, PS_ERR_NULL_POINTER, "translate: NULL pointer");
    PS_REQUIRE(sb >= 0 && sl >= 0 && sa >= 0, PS_ERR_BAD_SHAPE, "translate: negative stride");
    PS_REQUIRE(static_cast<long long>(L) * A * 3 < (1ll << 31), PS_ERR_BAD_SHAPE,
               "translate: L*A*3=%lld floats per structure exceed 2^31", static_cast<long long>(L) * A * 3);
    const int per_b = L * A * 3;
    // one translation per structure (or one for all): the per-structure map of stats.cu (128-bit accesses when the
    // shape allows) — what align() and center_at() issue
    if (sl == 0 && sa == 0 && (sb == 0 || sb == 3))
        return translate_impl(xyz, tr, sb == 0 ? 1 : B, B, L, A, out, stream);
    int gx = (per_b + 1023) / 1024;
    if (gx > 64) gx = 64;
    const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(B < 65535 ? B : 65535), 1);
    const unsigned long long row = static_cast<unsigned long long>(A) * 3ull;
    const bool fast = static_cast<unsigned long long>(per_b) * row < (1ull << 32) &&
                      static_cast<long long>(L - 1) * sl + static_cast<long long>(A - 1) * sa + 2 < (1ll << 31);
    const unsigned row_magic = static_cast<unsigned>(((1ull << 32) + row - 1) / row);
    if (fast && row > 1)
        translate_bcast_kernel<true><<<grid, 256, 0, stream>>>(xyz, tr, sb, sl, sa, B, A, per_b, row_magic, out);
    else
        translate_bcast_kernel<false><<<grid, 256, 0, stream>>>(xyz, tr, sb, sl, sa, B, A, per_b, 0u, out);
    return check_launch("translate_bcast_kernel");
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

int geom_rowwise_impl(const float* x, const float* y, long long n, int D, int op, float* out, cudaStream_t stream) {
    PS_REQUIRE(n >= 0 && D > 0, PS_ERR_BAD_SHAPE, "geom dot/norm/unit: n=%lld D=%d", n, D);
    PS_REQUIRE(op >= 0 && op <= 2, PS_ERR_BAD_DTYPE, "geom dot/norm/unit: unknown op %d", op);
    if (n == 0) return PS_OK;
    PS_REQUIRE(x && out && (op != 0 || y), PS_ERR_NULL_POINTER, "geom dot/norm/unit: NULL pointer");
    geom_rowwise_kernel<<<blocks_for(n), 256, 0, stream>>>(x, y, n, D, op, out);
    return check_launch("geom_rowwise_kernel");
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
