// Rows f2 and f4 of the scope table: batched Kabsch alignment and the top-k nearest-residue mask.
//
// Kabsch — replaces the per-structure Python loop of StructureBatch.align (protstruc/protstruc.py:880-918)
// around geometry.kabsch (protstruc/geometry.py:442-480): boolean gathers + einsum + 3x3 LAPACK SVD per
// structure become ONE launch, one CTA per structure: masked centroids, masked 3x3 covariance
// H = sum (a - ca)(b - cb)^T (warp-shuffle + shared-memory reductions, fp64 partial sums), then a
// closed 3x3 solve by one thread.  With H = U S V^T the reference forms R = V diag(1,1,sign det(V U^T)) U^T;
// that equals  v0 u0^T + v1 u1^T + (v0 x v1)(u0 x u1)^T  for the two leading singular pairs, which needs no
// sign bookkeeping: v0, v1 come from a Jacobi eigen-decomposition of H^T H, u_k = H v_k / |H v_k|.
// Roofline: HBM read of the two coordinate sets (24 B + 1 B per atom), tiny; latency-bound.
//
// top-k — replaces StructureBatch.get_topk_nearest_residue_mask (protstruc/protstruc.py:819-862):
// CA-to-query distances, min over the queries, invalid residues pushed to 1e9, the k smallest selected.
// Selection is by rank counting (O(L^2) compares, L is a few thousand at most) with index tie-break.

#include "common.cuh"

namespace ps {

namespace {

constexpr int kKabschThreads = 512;
constexpr int kKabschUnroll = 4;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double (*scratch)[N]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = warp_sum_d(v[k]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) scratch[warp][k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double t = lane < nwarps ? scratch[lane][k] : 0.0;
        v[k] = warp_sum_d(t);
    }
}

// Cyclic Jacobi for a symmetric positive semi-definite 3x3 matrix; eigenvectors in the columns of V, eigenvalues in w.
// One thread runs this per structure, so its latency is the kernel's tail: the matrix is scaled to unit trace, and the
// rotation ANGLE is computed in fp32 (MUFU square root and reciprocal instead of three fp64 divisions and two fp64
// square roots per rotation — an inexact angle only slows the convergence of that one rotation from "to zero" to "by
// 1e-7"), while the rotation itself, c = rsqrt(1 + t^2), s = t c, is applied in fp64 so that V stays orthogonal to
// fp64 accuracy.
__device__ void jacobi_eigen3(double (&a)[3][3], double (&V)[3][3], double (&w)[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    const double trace = a[0][0] + a[1][1] + a[2][2];
    const bool scaled = trace > 0.0 && trace < 1e300;
    if (scaled) {
        const double inv = 1.0 / trace;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) a[i][j] *= inv;
    }
    for (int sweep = 0; sweep < 32; ++sweep) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (off <= 1e-32 * diag || off == 0.0) break;
        // (every loop below is unrolled: the matrices are indexed by constants only and live in registers)
#pragma unroll
        for (int p = 0; p < 2; ++p) {
#pragma unroll
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                double t;
                if (scaled) {
                    // t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)), theta = d / (2 a_pq), without dividing by a_pq
                    const float apq2 = 2.0f * static_cast<float>(a[p][q]);
                    const float d = static_cast<float>(a[q][q] - a[p][p]);
                    const float den = fabsf(d) + sqrtf(fmaf(d, d, apq2 * apq2));
                    t = den > 0.f ? static_cast<double>(__fdividef(d >= 0.f ? apq2 : -apq2, den)) : 0.0;
                    if (t == 0.0) continue;  // a_pq below fp32 range relative to the trace: nothing left to rotate
                } else {
                    const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                    t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                }
                const double c = rsqrt(t * t + 1.0), s = t * c;
#pragma unroll
                for (int k = 0; k < 3; ++k) {  // A <- A J
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {  // A <- J^T A
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {  // V <- V J
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) w[i] = a[i][i];  // (scaled by 1 / trace: only their order is used)
}

// Closed 3x3 solve from the sixteen masked sums (first moments 0-5, count 6, raw second moments 7-15) by ONE thread.
__device__ void kabsch_solve(const double (&s)[16], long long b, float* __restrict__ rot, float* __restrict__ trans) {
    const double cnt = s[6];
    const double ca[3] = {s[0] / cnt, s[1] / cnt, s[2] / cnt};
    const double cb[3] = {s[3] / cnt, s[4] / cnt, s[5] / cnt};
    double h[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) h[i * 3 + j] = s[7 + i * 3 + j] - cnt * ca[i] * cb[j];

    {
        double H[3][3] = {{h[0], h[1], h[2]}, {h[3], h[4], h[5]}, {h[6], h[7], h[8]}};
        double K[3][3];  // H^T H
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) K[i][j] = H[0][i] * H[0][j] + H[1][i] * H[1][j] + H[2][i] * H[2][j];
        double V[3][3], w[3];
        jacobi_eigen3(K, V, w);
        int i0 = 0;  // indices of the two largest eigenvalues
        if (w[1] > w[i0]) i0 = 1;
        if (w[2] > w[i0]) i0 = 2;
        int i1 = (i0 + 1) % 3, i2 = (i0 + 2) % 3;
        if (w[i2] > w[i1]) i1 = i2;
        double v0[3] = {V[0][i0], V[1][i0], V[2][i0]}, v1[3] = {V[0][i1], V[1][i1], V[2][i1]};
        double u0[3], u1[3];
        for (int i = 0; i < 3; ++i) {
            u0[i] = H[i][0] * v0[0] + H[i][1] * v0[1] + H[i][2] * v0[2];
            u1[i] = H[i][0] * v1[0] + H[i][1] * v1[1] + H[i][2] * v1[2];
        }
        // Rank-deficient selections (reference: torch.linalg.svd still returns a proper rotation, geometry.py:470-478):
        //   no atom selected                      -> identity motion;
        //   H = 0 (one atom, coincident points)   -> identity rotation, translation between the centroids;
        //   rank 1 (collinear atoms)              -> the second singular pair is any unit vector orthogonal to the
        //                                            first on either side (the rotation about the line is free).
        const double n0 = sqrt(u0[0] * u0[0] + u0[1] * u0[1] + u0[2] * u0[2]);
        if (!(cnt > 0.0) || !(n0 > 0.0)) {
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) rot[b * 9 + i * 3 + j] = i == j ? 1.f : 0.f;
                trans[b * 3 + i] = cnt > 0.0 ? static_cast<float>(cb[i] - ca[i]) : 0.f;
            }
            return;
        }
        for (int i = 0; i < 3; ++i) u0[i] /= n0;
        const double p = u1[0] * u0[0] + u1[1] * u0[1] + u1[2] * u0[2];
        for (int i = 0; i < 3; ++i) u1[i] -= p * u0[i];
        double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
        if (!(n1 > 1e-9 * n0)) {
            int k = 0;  // the coordinate axis least aligned with u0
            if (fabs(u0[1]) < fabs(u0[k])) k = 1;
            if (fabs(u0[2]) < fabs(u0[k])) k = 2;
            const double e[3] = {k == 0 ? 1.0 : 0.0, k == 1 ? 1.0 : 0.0, k == 2 ? 1.0 : 0.0};
            u1[0] = u0[1] * e[2] - u0[2] * e[1];
            u1[1] = u0[2] * e[0] - u0[0] * e[2];
            u1[2] = u0[0] * e[1] - u0[1] * e[0];
            n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
        }
        for (int i = 0; i < 3; ++i) u1[i] /= n1;
        const double v2[3] = {v0[1] * v1[2] - v0[2] * v1[1], v0[2] * v1[0] - v0[0] * v1[2], v0[0] * v1[1] - v0[1] * v1[0]};
        const double u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
        double R[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R[i][j] = v0[i] * u0[j] + v1[i] * u1[j] + v2[i] * u2[j];
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) rot[b * 9 + i * 3 + j] = static_cast<float>(R[i][j]);
            trans[b * 3 + i] = static_cast<float>(cb[i] - (R[i][0] * ca[0] + R[i][1] * ca[1] + R[i][2] * ca[2]));
        }
    }
}

__global__ void __launch_bounds__(kKabschThreads) kabsch_kernel(
    const float* __restrict__ src, const float* __restrict__ dst, const uint8_t* __restrict__ mask,
    int dst_rows, int n_atoms, float* __restrict__ rot, float* __restrict__ trans) {
    __shared__ double scratch[kKabschThreads / 32][16];
    const long long b = blockIdx.x;
    const float* __restrict__ a = src + b * n_atoms * 3;
    const float* __restrict__ t = dst + (dst_rows == 1 ? 0 : b) * static_cast<long long>(n_atoms) * 3;
    const uint8_t* __restrict__ m = mask + b * n_atoms;

    // ONE pass over the selected atoms: first moments (centroids, reference: a.mean(dim=-2), b.mean(dim=-2)) and
    // the raw second moments sum a_i b_j, all in fp64, so that the centred covariance
    //   H[i][j] = sum_k (a_k - ca)_i (b_k - cb)_j = sum a_i b_j - n ca_i cb_j
    // loses nothing that matters (|x| ~ 1e2, n ~ 1e4: the subtraction cancels ~4 of fp64's 16 digits).
    // The loop is latency-bound (one CTA per structure, ~15 atoms per thread): the mask bytes and BOTH coordinate
    // triples of kKabschUnroll atoms are requested before the first use, unconditionally (an unselected atom's
    // coordinates are discarded by the select below, so NaN there is harmless).
    double s[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k0 = threadIdx.x; k0 < n_atoms; k0 += kKabschUnroll * blockDim.x) {
        float pa[kKabschUnroll][3], pb[kKabschUnroll][3];
        bool sel[kKabschUnroll];
#pragma unroll
        for (int u = 0; u < kKabschUnroll; ++u) {
            const int k = k0 + u * blockDim.x;
            const bool in = k < n_atoms;
            const int kk = in ? k : 0;
            sel[u] = in && __ldg(m + kk) != 0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                pa[u][c] = __ldg(a + 3 * kk + c);
                pb[u][c] = __ldg(t + 3 * kk + c);
            }
        }
#pragma unroll
        for (int u = 0; u < kKabschUnroll; ++u) {
            if (sel[u]) {
                const double ax = pa[u][0], ay = pa[u][1], az = pa[u][2];
                const double bx = pb[u][0], by = pb[u][1], bz = pb[u][2];
                s[0] += ax; s[1] += ay; s[2] += az;
                s[3] += bx; s[4] += by; s[5] += bz;
                s[6] += 1.0;
                s[7] += ax * bx; s[8] += ax * by; s[9] += ax * bz;
                s[10] += ay * bx; s[11] += ay * by; s[12] += ay * bz;
                s[13] += az * bx; s[14] += az * by; s[15] += az * bz;
            }
        }
    }
    block_sum<16>(s, scratch);
    if (threadIdx.x == 0) kabsch_solve(s, b, rot, trans);
}

// The same sums with 128-bit loads: four atoms (48 B of each coordinate set, 4 mask bytes) per thread and step as
// 3 + 3 + 1 loads instead of 28, two steps in flight, 256 threads so that two CTAs (structures) share an SM.  Needs
// n_atoms % 4 == 0 and 16-byte aligned coordinate sets (every structure then starts on a 16-byte boundary).
__device__ __forceinline__ void kabsch_accumulate(double (&s)[16], bool sel, float ax, float ay, float az, float bx,
                                                  float by, float bz) {
    if (sel) {
        const double dax = ax, day = ay, daz = az, dbx = bx, dby = by, dbz = bz;
        s[0] += dax; s[1] += day; s[2] += daz;
        s[3] += dbx; s[4] += dby; s[5] += dbz;
        s[6] += 1.0;
        s[7] += dax * dbx; s[8] += dax * dby; s[9] += dax * dbz;
        s[10] += day * dbx; s[11] += day * dby; s[12] += day * dbz;
        s[13] += daz * dbx; s[14] += daz * dby; s[15] += daz * dbz;
    }
}

constexpr int kKabschQuadThreads = 256;

__global__ void __launch_bounds__(kKabschQuadThreads, 2) kabsch_quad_kernel(
    const float* __restrict__ src, const float* __restrict__ dst, const uint8_t* __restrict__ mask,
    int dst_rows, int n_atoms, float* __restrict__ rot, float* __restrict__ trans) {
    __shared__ double scratch[kKabschQuadThreads / 32][16];
    const long long b = blockIdx.x;
    const float4* __restrict__ a4 = reinterpret_cast<const float4*>(src + b * n_atoms * 3);
    const float4* __restrict__ t4 =
        reinterpret_cast<const float4*>(dst + (dst_rows == 1 ? 0 : b) * static_cast<long long>(n_atoms) * 3);
    const uint32_t* __restrict__ m4 = reinterpret_cast<const uint32_t*>(mask + b * n_atoms);
    const int quads = n_atoms >> 2;
    double s[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int q0 = threadIdx.x; q0 < quads; q0 += 2 * blockDim.x) {
        uint32_t m[2];
        float4 a[2][3], t[2][3];
#pragma unroll
        for (int u = 0; u < 2; ++u) {  // both steps' loads are requested before the first use
            const int q = q0 + u * blockDim.x;
            const bool in = q < quads;
            const int qq = in ? q : q0;
            m[u] = in ? __ldg(m4 + qq) : 0u;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                a[u][k] = __ldg(a4 + 3 * qq + k);
                t[u][k] = __ldg(t4 + 3 * qq + k);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            kabsch_accumulate(s, (m[u] & 0x000000ffu) != 0, a[u][0].x, a[u][0].y, a[u][0].z, t[u][0].x, t[u][0].y, t[u][0].z);
            kabsch_accumulate(s, (m[u] & 0x0000ff00u) != 0, a[u][0].w, a[u][1].x, a[u][1].y, t[u][0].w, t[u][1].x, t[u][1].y);
            kabsch_accumulate(s, (m[u] & 0x00ff0000u) != 0, a[u][1].z, a[u][1].w, a[u][2].x, t[u][1].z, t[u][1].w, t[u][2].x);
            kabsch_accumulate(s, (m[u] & 0xff000000u) != 0, a[u][2].y, a[u][2].z, a[u][2].w, t[u][2].y, t[u][2].z, t[u][2].w);
        }
    }
    block_sum<16>(s, scratch);
    if (threadIdx.x == 0) kabsch_solve(s, b, rot, trans);
}

// ---- top-k nearest residues ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) min_query_distance_kernel(const float* __restrict__ xyz,
                                                                 const uint8_t* __restrict__ valid,
                                                                 const float* __restrict__ query, int n_query,
                                                                 int L, int A, int ca_slot,
                                                                 float* __restrict__ dmin) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= L) return;
    const V3 ca = ld3(xyz + (static_cast<long long>(r) * A + ca_slot) * 3);
    float best = __int_as_float(0x7f800000);
    bool any_nan = false;
    for (int q = 0; q < n_query; ++q) {
        const V3 d = sub3(ca, ld3(query + q * 3));
        const float v = norm3(d);
        any_nan |= (v != v);
        best = fminf(best, v);
    }
    if (any_nan) best = __int_as_float(0x7fc00000);  // torch.min propagates NaN
    dmin[r] = __ldg(valid + r) ? best : 1e9f;
}

__global__ void __launch_bounds__(256) rank_select_kernel(const float* __restrict__ dmin, int L, int k,
                                                          uint8_t* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= L) return;
    const float mine = dmin[r];
    int rank = 0;
    for (int s = 0; s < L; ++s) {
        const float other = __ldg(dmin + s);
        // NaN sorts last (torch.topk treats NaN as the largest value)
        const bool other_nan = other != other, mine_nan = mine != mine;
        const bool less = (!other_nan && mine_nan) || (!other_nan && !mine_nan && other < mine) ||
                          ((other == mine || (other_nan && mine_nan)) && s < r);
        rank += less ? 1 : 0;
    }
    out[r] = rank < k ? 1 : 0;
}

}  // namespace

int kabsch_impl(const float* src, const float* dst, const uint8_t* mask, int dst_rows, int B, int n_atoms,
                float* rot, float* trans, cudaStream_t stream) {
    PS_REQUIRE(B > 0 && n_atoms > 0, PS_ERR_BAD_SHAPE, "kabsch: B=%d atoms=%d must be > 0", B, n_atoms);
    PS_REQUIRE(src && dst && mask && rot && trans, PS_ERR_NULL_POINTER, "kabsch: NULL pointer");
    PS_REQUIRE(dst_rows == 1 || dst_rows == B, PS_ERR_BAD_SHAPE, "kabsch: %d targets for %d structures", dst_rows, B);
    // ~16 atoms per thread: small structures get small CTAs, so that many of the serial 3x3 solves (one thread per
    // structure, a few microseconds of dependent fp64 arithmetic) run side by side on an SM
    int threads = (n_atoms / 16 + 31) / 32 * 32;
    if (threads < 64) threads = 64;
    if (threads > kKabschThreads) threads = kKabschThreads;
    const bool quads = (n_atoms % 4 == 0) && n_atoms >= 4 * 64 &&
                       ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0 &&
                       (reinterpret_cast<uintptr_t>(mask) & 3u) == 0;
    if (quads) {
        // CTA size: the largest of 256 / 128 / 64 threads with which all B structures are resident at once (110
        // registers per thread), else about four 4-atom steps per thread
        const int sms = sm_count_for_current_device();
        if (sms < 0) return sms;
        int quad_threads = 0;
        for (int cand : {256, 128, 64}) {
            if (4 * cand > n_atoms) continue;
            int per_sm = 65536 / (110 * cand);
            if (per_sm > 32) per_sm = 32;
            if (static_cast<long long>(per_sm) * sms >= B) { quad_threads = cand; break; }
        }
        if (quad_threads == 0) quad_threads = n_atoms / 16 >= 256 ? 256 : (n_atoms / 16 >= 128 ? 128 : 64);
        kabsch_quad_kernel<<<B, quad_threads, 0, stream>>>(src, dst, mask, dst_rows, n_atoms, rot, trans);
        return check_launch("kabsch_quad_kernel");
    }
    kabsch_kernel<<<B, threads, 0, stream>>>(src, dst, mask, dst_rows, n_atoms, rot, trans);
    return check_launch("kabsch_kernel");
}

int topk_nearest_impl(const float* xyz, const uint8_t* valid, const float* query, int n_query, int L, int A,
                      int ca_slot, int k, float* scratch_dmin, uint8_t* out, cudaStream_t stream) {
    PS_REQUIRE(L > 0 && A > 0 && n_query > 0, PS_ERR_BAD_SHAPE, "topk_nearest: L=%d A=%d queries=%d", L, A, n_query);
    PS_REQUIRE(xyz && valid && query && scratch_dmin && out, PS_ERR_NULL_POINTER, "topk_nearest: NULL pointer");
    PS_REQUIRE(ca_slot >= 0 && ca_slot < A, PS_ERR_BAD_SLOT, "topk_nearest: slot %d outside [0,%d)", ca_slot, A);
    PS_REQUIRE(k >= 0, PS_ERR_BAD_SHAPE, "topk_nearest: k=%d", k);
    const unsigned blocks = static_cast<unsigned>((L + 255) / 256);
    min_query_distance_kernel<<<blocks, 256, 0, stream>>>(xyz, valid, query, n_query, L, A, ca_slot, scratch_dmin);
    int rc = check_launch("min_query_distance_kernel");
    if (rc != PS_OK) return rc;
    rank_select_kernel<<<blocks, 256, 0, stream>>>(scratch_dmin, L, k, out);
    return check_launch("rank_select_kernel");
}

}  // namespace ps
