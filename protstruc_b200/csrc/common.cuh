// Shared device helpers for the protstruc_b200 kernels (sm_100a only).
//
// Arithmetic contract: the angle / frame helpers reproduce the reference's op order
// (protstruc/geometry.py:24-124, 413-439) with explicitly rounded fp32 operations
// (__fmul_rn / __fsub_rn / __fadd_rn are never contracted into FMAs by nvcc), because
// the reference computes cross and dot products as separate numpy / ATen array ops and
// its degenerate cases (diagonal pairs, collinear atoms) depend on exact cancellation.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/protstruc_b200.h"

namespace ps {

// ---------------------------------------------------------------- error plumbing (host)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t err, const char* what);
int check_launch(const char* kernel_name);
int sm_count_for_current_device();

#define PS_REQUIRE(cond, code, ...)     \
    do {                                \
        if (!(cond)) {                  \
            ::ps::set_error(__VA_ARGS__); \
            return (code);              \
        }                               \
    } while (0)

// ---------------------------------------------------------------- small vector type
struct V3 {
    float x, y, z;
};

__device__ __forceinline__ V3 ld3(const float* __restrict__ p) {
    V3 v;
    v.x = __ldg(p);
    v.y = __ldg(p + 1);
    v.z = __ldg(p + 2);
    return v;
}

__device__ __forceinline__ V3 sub3(V3 a, V3 b) {
    return V3{__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)};
}

// np.cross / torch.cross on the last axis: two rounded products, one rounded subtraction.
__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
    V3 r;
    r.x = __fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y));
    r.y = __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z));
    r.z = __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x));
    return r;
}

// (x * y).sum(-1): products are materialised (rounded) first, then accumulated left to right
// starting from +0 like ATen's sum does.  The leading +0 matters: three -0 products (zero-padded
// residues) must sum to +0, otherwise atan2(+0, -0) = pi where the reference returns 0.
__device__ __forceinline__ float dot3(V3 a, V3 b) {
    const float acc = __fadd_rn(0.0f, __fmul_rn(a.x, b.x));
    return __fadd_rn(__fadd_rn(acc, __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}

// x.norm(dim=-1) / torch.norm(x, dim=-1): IEEE sqrt of the sum of squares.  ATen's CPU vector-norm kernel
// accumulates the squares with fused multiply-adds — fma(z, z, fma(y, y, x * x)) — for every shape the reference
// uses it on (probed on 1e6 random vectors and on (n, 3) ... (B, L, L, 3) shapes: bit-identical to this chain,
// 10 % one-ulp mismatches against the separately rounded sum).  The last ulp matters where a quotient by a
// product of norms is fed to an unclamped arccos (geometry.py:64-71): it decides which collinear triples are NaN.
__device__ __forceinline__ float norm3(V3 a) {
    return __fsqrt_rn(__fmaf_rn(a.z, a.z, __fmaf_rn(a.y, a.y, __fmul_rn(a.x, a.x))));
}

__device__ __forceinline__ V3 scale3(V3 a, float s) {
    return V3{__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)};
}

__device__ __forceinline__ V3 div3(V3 a, float s) {
    return V3{__fdiv_rn(a.x, s), __fdiv_rn(a.y, s), __fdiv_rn(a.z, s)};
}

// Single-MUFU primitives refined to ~1 ulp, used for the last scalar steps of the dihedral (see the note at
// "tuned trRosetta triple" below).
__device__ __forceinline__ float rcp_mufu(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_mufu(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// 1/sqrt(s) to ~1 ulp.  s = 0 gives NaN (inf * 0), which is what the reference's 0/0 gives downstream.
__device__ __forceinline__ float rsqrt_refined(float s) {
    const float r = rsqrt_mufu(s);
    return r * fmaf(-0.5f * s * r, r, 1.5f);
}

__device__ __forceinline__ float atan2_tuned(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float sum = ax + ay;
    if (!((mx > 1e-30f) & (mx < 1e30f) & (sum == sum))) {
        // The two everyday special cases answered in place (bit for bit what atan2f returns, without its ~50
        // instructions): a NaN operand, and x = y = 0 — every pair that involves a zero-padded residue of a ragged
        // batch: atan2(+-0, +0) = +-0, atan2(+-0, -0) = +-pi.
        if (!(sum == sum)) return __int_as_float(0x7fc00000);
        if (mx == 0.f) return copysignf(__float_as_int(x) < 0 ? 3.14159274f : 0.f, y);
        return atan2f(y, x);
    }
    const float r = rcp_mufu(mx);
    float t = mn * r;
    t = fmaf(r, fmaf(-t, mx, mn), t);  // quotient refined to ~1 ulp
    const float s = t * t;
    float p = 2.622234402e-03f;
    p = fmaf(p, s, -1.513249334e-02f);
    p = fmaf(p, s, 4.112178832e-02f);
    p = fmaf(p, s, -7.366699725e-02f);
    p = fmaf(p, s, 1.057392880e-01f);
    p = fmaf(p, s, -1.418597400e-01f);
    p = fmaf(p, s, 1.999039650e-01f);
    p = fmaf(p, s, -3.333298564e-01f);
    float a = fmaf(t * s, p, t);
    if (ay > ax) a = 1.57079637f - a;
    if (x < 0.f) a = 3.14159274f - a;
    return copysignf(a, y);
}

__device__ __forceinline__ bool has_nan3(V3 a) { return (a.x != a.x) | (a.y != a.y) | (a.z != a.z); }

// Missing atoms are NaN coordinates in the reference's tensors (protstruc/pdb.py:133-135) and every
// NaN input coordinate makes the angle NaN (each cross / dot product mixes all three components).
// Returning NaN up front is therefore exact, and it keeps NaN-carrying lanes out of the slow paths of
// the IEEE division / square root / atan2f / acosf sequences (a 25 % hit on the fused kernel with half
// of the atoms missing).  The cheap probe is a plain sum; only when it is NaN (a NaN, or +inf and -inf
// together) are the coordinates inspected one by one, so infinities still take the full computation.
__device__ __forceinline__ bool any_nan_coordinate(V3 a, V3 b, V3 c) {
    const float probe = ((a.x + a.y) + (a.z + b.x)) + ((b.y + b.z) + (c.x + c.y)) + c.z;
    if (probe == probe) return false;
    return has_nan3(a) || has_nan3(b) || has_nan3(c);
}

// geometry.dihedral (protstruc/geometry.py:110-124).
__device__ __forceinline__ float dihedral4(V3 a, V3 b, V3 c, V3 d) {
    if (any_nan_coordinate(a, b, c) || has_nan3(d)) return __int_as_float(0x7fc00000);
    const V3 b0 = sub3(a, b);
    const V3 b1 = sub3(c, b);
    const V3 b2 = sub3(d, c);
    const V3 n1 = cross3(b0, b1);
    const V3 n2 = cross3(b2, b1);
    const V3 m = cross3(n1, n2);
    const float x = dot3(n1, n2);
    // (m.b1) / |b1| and atan2 with the tuned primitives: same result to ~1 ulp as the IEEE division / sqrt /
    // libdevice atan2f sequence at a third of the issue slots (special values fall back to atan2f)
    const float y = __fmul_rn(dot3(m, b1), rsqrt_refined(dot3(b1, b1)));
    return atan2_tuned(y, x);
}

// geometry.angle (protstruc/geometry.py:64-71); no clamp, exactly like the reference.
__device__ __forceinline__ float angle3(V3 a, V3 b, V3 c) {
    if (any_nan_coordinate(a, b, c)) return __int_as_float(0x7fc00000);
    const V3 ba = sub3(a, b);
    const V3 bc = sub3(c, b);
    const float cosine = __fdiv_rn(dot3(ba, bc), __fmul_rn(norm3(ba), norm3(bc)));
    return acosf(cosine);
}

// Virtual CB from N, CA, C (protstruc/geometry.py:217-221):
//   b = CA - N, c = C - CA, a = b x c;  CB = -0.58273431 a + 0.56802827 b - 0.54067466 c + CA
// evaluated left to right with separately rounded ops as the torch expression does.
__device__ __forceinline__ V3 virtual_cb(V3 n, V3 ca, V3 c) {
    const V3 vb = sub3(ca, n);
    const V3 vc = sub3(c, ca);
    const V3 va = cross3(vb, vc);
    V3 r;
    r.x = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(-0.58273431f, va.x), __fmul_rn(0.56802827f, vb.x)),
                              __fmul_rn(0.54067466f, vc.x)), ca.x);
    r.y = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(-0.58273431f, va.y), __fmul_rn(0.56802827f, vb.y)),
                              __fmul_rn(0.54067466f, vc.y)), ca.y);
    r.z = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(-0.58273431f, va.z), __fmul_rn(0.56802827f, vb.z)),
                              __fmul_rn(0.54067466f, vc.z)), ca.z);
    return r;
}

// ---------------------------------------------------------------- tuned trRosetta triple
// omega / theta / phi of one residue pair for the two hot kernels (K2f and the fused K1).  Same
// formulas and the same non-contracted cross / dot products as dihedral4 / angle3 above — so exact
// cancellations (diagonal pairs, zero-padded residues) and NaN placement are unchanged — but the final
// scalar steps use single-MUFU primitives refined to ~1 ulp instead of the IEEE division / sqrt /
// libdevice atan2f sequences (which cost more than the geometry itself):
//   * y = (m.b1) / |b1|  ->  (m.b1) * rsqrt(b1.b1), one Newton step on the MUFU.RSQ seed;
//   * atan2 -> min/max quotient by MUFU.RCP + one Newton step, degree-15 odd minimax polynomial on
//     [0, 1] (max error 7e-9 before rounding), octant fix-ups; zeros, infinities, NaN and out-of-range
//     magnitudes fall back to atan2f, so atan2(+-0, -0) = +-pi etc. keep their IEEE values.
// Measured deviation from the reference after this change: see profiles/*parity_report*.json.
__device__ __forceinline__ bool atom_has_nan(V3 a) {
    const float probe = (a.x + a.y) + a.z;
    return (probe != probe) && has_nan3(a);
}

// Note on packed arithmetic: evaluating two pairs per thread on the FADD2 / FMUL2 pipe would halve the issue
// slots of the cross / dot products, but ptxas 12.9 fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with
// explicit .rn modifiers and with -fmad=false (checked on sm_100a), which breaks the exact cancellations the
// reference relies on (a x a = 0 on the diagonal).  The geometry therefore stays on scalar, individually
// rounded operations.
//
// Everything of the triple that depends on residue i only (hoisted out of the j loop by K2f).
struct TripleRowSide {
    V3 cb;        // CB_i
    V3 b0;        // CA_i - CB_i   (omega's b0, phi's ba)
    V3 tb1;       // CB_i - CA_i   (theta's b1)
    V3 tn1;       // (N_i - CA_i) x (CB_i - CA_i)   (theta's n1)
    float inv_tb1_norm;  // 1 / |CB_i - CA_i|
    float inv_ba_norm;   // 1 / |CA_i - CB_i|
    float ba_norm;       // |CA_i - CB_i|  (phi's exact sequence near |cos| = 1)
    bool nan_ca_cb;      // CA_i or CB_i missing
    bool nan_n;          // N_i missing
};

__device__ __forceinline__ TripleRowSide triple_row_side(V3 n_i, V3 ca_i, V3 cb_i) {
    TripleRowSide r;
    r.cb = cb_i;
    r.b0 = sub3(ca_i, cb_i);
    r.tb1 = sub3(cb_i, ca_i);
    r.tn1 = cross3(sub3(n_i, ca_i), r.tb1);
    r.inv_tb1_norm = __frcp_rn(norm3(r.tb1));
    r.ba_norm = norm3(r.b0);
    r.inv_ba_norm = __frcp_rn(r.ba_norm);
    r.nan_ca_cb = atom_has_nan(ca_i) || atom_has_nan(cb_i);
    r.nan_n = atom_has_nan(n_i);
    return r;
}

// omega = dihedral(CA_i, CB_i, CA_j, CB_j); theta = dihedral(N_i, CA_i, CB_i, CB_j); phi = angle(CA_i, CB_i, CB_j)
// (reference protstruc/protstruc.py:810-815).
__device__ __forceinline__ void trrosetta_triple(const TripleRowSide& r, V3 ca_j, V3 cb_j, bool want_omega,
                                                 bool want_theta, bool want_phi, float& omega, float& theta,
                                                 float& phi) {
    const float nan = __int_as_float(0x7fc00000);
    const bool nan_cb_j = atom_has_nan(cb_j);
    const bool nan_ca_j = atom_has_nan(ca_j);
    const V3 bc = sub3(cb_j, r.cb);  // CB_j - CB_i: theta's b2 and phi's bc
    omega = theta = phi = nan;
    if (want_omega && !(r.nan_ca_cb || nan_ca_j || nan_cb_j)) {
        const V3 b1 = sub3(ca_j, r.cb);
        const V3 b2 = sub3(cb_j, ca_j);
        const V3 n1 = cross3(r.b0, b1);
        const V3 n2 = cross3(b2, b1);
        const V3 m = cross3(n1, n2);
        const float x = dot3(n1, n2);
        const float y = __fmul_rn(dot3(m, b1), rsqrt_refined(dot3(b1, b1)));
        omega = atan2_tuned(y, x);
    }
    if (want_theta && !(r.nan_ca_cb || r.nan_n || nan_cb_j)) {
        const V3 n2 = cross3(bc, r.tb1);
        const V3 m = cross3(r.tn1, n2);
        const float x = dot3(r.tn1, n2);
        const float y = __fmul_rn(dot3(m, r.tb1), r.inv_tb1_norm);
        theta = atan2_tuned(y, x);
    }
    if (want_phi && !(r.nan_ca_cb || nan_cb_j)) {
        // cos = (ba.bc) / (|ba| |bc|) without a clamp (reference geometry.py:64-71): acos of 1.0000001 is NaN,
        // so WHERE the rounded quotient exceeds 1 is part of the contract.  Away from |cos| = 1 the quotient is
        // formed with the hoisted 1/|ba| and one refined MUFU.RSQ (<= 3 ulp from the reference's value, i.e.
        // <= 4e-6 rad at sin >= 0.045); within 1e-3 of +-1 — and for anything non-finite (the diagonal's 0/0,
        // zero-padded residues, |bc|^2 below the ftz threshold) — the reference's exact op sequence is issued:
        // ATen's norms (norm3), their rounded product, IEEE division.  NaN placement is then the reference's
        // bit for bit (tests: adversarial collinear triples).
        const float d = dot3(r.b0, bc);
        const float bc2 = dot3(bc, bc);
        float cosine = __fmul_rn(__fmul_rn(d, r.inv_ba_norm), rsqrt_refined(bc2));
        if (!(fabsf(cosine) <= 0.999f)) cosine = __fdiv_rn(d, __fmul_rn(r.ba_norm, norm3(bc)));
        phi = cosine == cosine ? acosf(cosine) : nan;  // (0 / 0 of a zero-padded residue: NaN without the call)
    }
}

// n / d for a non-negative 64-bit index and a positive 32-bit divisor: the 32-bit instruction sequence whenever
// the caller knows (uniformly) that the index range fits — an emulated 64-bit division costs ~120 issue slots.
__device__ __forceinline__ long long index_div(long long n, long long d, bool fits_32_bits) {
    if (fits_32_bits) return static_cast<unsigned>(n) / static_cast<unsigned>(d);
    return n / d;
}

// ---------------------------------------------------------------- bulk async copy (TMA engine)
// smem -> global bulk copy, tracked by the per-thread bulk async-group.
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(ssrc));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(s), "r"(bytes) : "memory");
}
// The same with an L2 eviction policy for the written lines (createpolicy.fractional.L2::evict_first / evict_last).
__device__ __forceinline__ void bulk_store_s2g_hint(void* gdst, const void* ssrc, uint32_t bytes, uint64_t policy) {
    const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(ssrc));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 :: "l"(gdst), "r"(s), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy(int which) {  // 1 = evict_first, 2 = evict_last
    uint64_t policy;
    if (which == 2)
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    else
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    return policy;
}
__device__ __forceinline__ void bulk_commit() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// Wait until the smem source of all committed groups has been read (buffer reusable).
__device__ __forceinline__ void bulk_wait_read_all() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// Wait until all committed groups are complete (writes performed).
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// Make generic-proxy smem writes visible to the async proxy before a bulk copy reads them.
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

}  // namespace ps
