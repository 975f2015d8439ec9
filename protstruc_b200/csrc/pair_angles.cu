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

#include <type_traits>

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

// ------------------------------------------------------------------------------------------------------------------
// K2f, second generation: the trRosetta triple on the packed FP32 pipe.
//
// The first version (trrosetta_kernel above, kept as the exact-sequence variant and for structures that do not fit in
// shared memory) issues the reference's operation sequence literally — separately rounded products, cross product of
// cross products — and was issue-bound at 289 lane-instructions per pair (profiles/r1o_k2f_ncu_summary.txt: 0.20 of
// the HBM roof).  This version spends the tolerance the contract gives (<= 1e-5 rad where min sin(bond angle) >= 0.1)
// on a cheaper evaluation of the SAME angles, and keeps the reference's NaN placement and exact special cases:
//
//  * a thread owns TWO consecutive residues j, j + 1 of a row and evaluates both pairs at once as f32x2 values: every
//    subtraction, product and fused multiply-add of the geometry is one FADD2 / FMUL2 / FFMA2 instruction for two
//    pairs (fused multiply-adds are more accurate than the reference's separately rounded products, never less);
//  * the sine term needs no third cross product:  (n1 x n2) . b1 = -(n1 . b2) |b1|^2  for n1 = b0 x b1, n2 = b2 x b1,
//    hence  y = (m . b1) / |b1| = -(n1 . b2) |b1|;
//  * a CTA stages the structure's CA / CB (real slot 4, or the virtual CB of geometry.py:217-221 evaluated ONCE per
//    residue with the reference's exact sequence) as structure-of-arrays in shared memory, so residue j arrives as
//    three conflict-free LDS.64 per atom instead of six strided global loads, and everything that depends on residue i
//    only (b0, CB_i, the normal n1 of theta, |b0|, 1/|b0|) is computed once per residue, not once per thread;
//  * atan2 = MUFU.RCP quotient + degree-13 odd minimax polynomial evaluated as FFMA2 for the two pairs, octant
//    fix-ups; acos = sqrt(1 - |c|) P(|c|), one range; ONE range test per loop iteration decides whether some lane
//    (zeros, infinities, NaN, out-of-range magnitudes, |cos| near 1) is redone by the IEEE / exact-sequence path;
//  * what bounds the kernel is the FP32 pipe, not HBM: an FFMA2 / FMUL2 / FADD2 saves an issue slot but occupies the
//    pipe for two cycles (ncu: sm__pipe_fma_cycles_active = 2 x sm__inst_executed_pipe_fma), so the per-pair operation
//    count is what matters: ~66 packed instructions per two pairs after moving every product of row-only vectors
//    into the row record (theta needs two dot products per pair, no cross product).
//
// What keeps the special cases exact (all covered by tests/test_gpu_parity.py):
//  * zero-padded residues and coincident atoms: a zero operand makes every fused product exactly zero, so x = y = 0
//    arrives at the atan2f fallback as in the reference (x is given the reference's +0 sign there: ATen's sum starts
//    from +0); norms are formed as v.v * rsqrt(v.v), which is NaN for v = 0 exactly where the reference divides 0 / 0;
//  * the diagonal j = i: b1 = b0 (omega) and b2 = 0 (theta, phi) make every term zero in exact arithmetic, and the
//    values are known per residue: omega = theta = 0 — or NaN if an atom is missing or CA_i = CB_i (0 * 1/|b0|) — and
//    phi = NaN (0 / 0).  They are computed once per row record and written after the row loop, over whatever the
//    straight-line path produced there;
//  * missing atoms (NaN coordinates) propagate through every product and through the select-based min / max of the
//    atan2 (fminf / fmaxf would drop them);
//  * phi: within 1e-4 of |cos| = 1 — where the unclamped arccos of the reference turns a last-ulp excess into NaN —
//    the reference's exact sequence is issued (trrosetta_phi_exact below), as in the fused K1.
template <typename T>
__device__ __forceinline__ float2 f2(T v) { return make_float2(v, v); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }

struct P3 {  // three f32x2 values: one 3-vector for each of the two pairs of a thread
    float2 x, y, z;
};
__device__ __forceinline__ P3 sub_p3(P3 a, P3 b) {
    return P3{__fadd2_rn(a.x, neg2(b.x)), __fadd2_rn(a.y, neg2(b.y)), __fadd2_rn(a.z, neg2(b.z))};
}
__device__ __forceinline__ P3 cross_p3(P3 a, P3 b) {  // a x b, one FMUL2 + one FFMA2 per component
    P3 r;
    r.x = __ffma2_rn(a.y, b.z, neg2(__fmul2_rn(a.z, b.y)));
    r.y = __ffma2_rn(a.z, b.x, neg2(__fmul2_rn(a.x, b.z)));
    r.z = __ffma2_rn(a.x, b.y, neg2(__fmul2_rn(a.y, b.x)));
    return r;
}
__device__ __forceinline__ float2 dot_p3(P3 a, P3 b) {
    return __ffma2_rn(a.z, b.z, __ffma2_rn(a.y, b.y, __fmul2_rn(a.x, b.x)));
}

// max / min that PROPAGATE NaN (fmaxf / fminf return the other operand): one FMNMX each, and a missing atom on either
// side of an atan2 shows up in both of them.
__device__ __forceinline__ float max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float min_nan(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

__device__ __forceinline__ float max3_nan(float a, float b, float c) {  // one FMNMX3 (sm_100+)
    float r;
    asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float min3_nan(float a, float b, float c) {
    float r;
    asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// The packed atan2 below is valid while the larger of |x|, |y| lies in this range (MUFU.RCP of it and the quotient are
// then normal numbers); zero (the diagonal, zero padding, coincident atoms), huge, infinite and NaN operands are not.
constexpr float kAtanLo = 1e-30f, kAtanHi = 1e30f;
// The packed cosine (relative error < 4e-7) decides between a value and NaN of the unclamped arccos only up to this
// |cos|; beyond it the reference's exact sequence is issued.  (1e-3 below 1 in the first version: one pair in a thousand
// — a patch in 6 % of the warp iterations — where one in ten thousand leaves the same 250-fold margin on the decision.)
constexpr float kCosExact = 0.9999f;
__device__ __forceinline__ bool atan2_in_range(float ny, float x) {
    const float mx = max_nan(fabsf(x), fabsf(ny));
    return (mx > kAtanLo) & (mx < kAtanHi);
}

// atan2(-ny, x) of ONE out-of-range operand pair: IEEE atan2f; x + 0 turns a -0 cosine term into the reference's +0
// (ATen's sum starts from +0).  A NaN operand is answered at once instead of being dragged through atan2f.
__device__ __forceinline__ float atan2_slow(float ny, float x) {
    if ((x != x) | (ny != ny)) return __int_as_float(0x7fc00000);
    return atan2f(-ny, x + 0.0f);
}

// atan2(-ny, x) of two operand pairs at once (the callers hold the NEGATED sine term: the sign flip is folded into the
// final sign transfer): MUFU.RCP quotient of the smaller by the larger magnitude, the degree-13 odd minimax polynomial
// of atan on [0, 1] (|error| 3.8e-7 rad; every FFMA2 is two cycles of the FP32 pipe that bounds this kernel, and the
// degree-17 polynomial of the first version bought accuracy below fp32 rounding), octant fix-ups — WITHOUT any range
// test or branch: mx0 / mx1 are max(|x|, |y|) of the lanes, from which the caller decides — once for all the atan2 of
// a loop iteration — whether some lane has to be redone by atan2_slow.  The evaluation is split into prepare / Horner
// / finish so that the caller can interleave the dependent chains of its independent angles statement by statement
// (with six warps per scheduler, a serial Horner chain leaves the FP32 pipe idle between its steps).
struct AtanPair {
    float2 t, s, p;
    float mx0, mx1;
    bool sw0, sw1;
};
__device__ __forceinline__ void atan2_prepare(float2 ny, float2 x, AtanPair& e) {
    const float ax0 = fabsf(x.x), ay0 = fabsf(ny.x), ax1 = fabsf(x.y), ay1 = fabsf(ny.y);
    e.mx0 = max_nan(ax0, ay0);
    e.mx1 = max_nan(ax1, ay1);
    e.sw0 = ay0 > ax0;
    e.sw1 = ay1 > ax1;
    const float mn0 = min_nan(ax0, ay0), mn1 = min_nan(ax1, ay1);
    e.t = __fmul2_rn(make_float2(mn0, mn1), make_float2(rcp_mufu(e.mx0), rcp_mufu(e.mx1)));
    e.s = __fmul2_rn(e.t, e.t);
    e.p = __ffma2_rn(f2(7.353078341e-03f), e.s, f2(-3.545713666e-02f));
}
__device__ __forceinline__ float2 atan2_finish(const AtanPair& e, float2 ny, float2 x) {
    const float2 a = __ffma2_rn(__fmul2_rn(e.t, e.s), e.p, e.t);
    float a0 = a.x, a1 = a.y;
    if (e.sw0) a0 = 1.57079637f - a0;
    if (e.sw1) a1 = 1.57079637f - a1;
    if (x.x < 0.f) a0 = 3.14159274f - a0;
    if (x.y < 0.f) a1 = 3.14159274f - a1;
    // a >= +0 here (or NaN): give it the sign of y = -ny with one LOP3
    a0 = __int_as_float(__float_as_int(a0) | (~__float_as_int(ny.x) & 0x80000000));
    a1 = __int_as_float(__float_as_int(a1) | (~__float_as_int(ny.y) & 0x80000000));
    return make_float2(a0, a1);
}

// phi's cosine by the reference's exact sequence (geometry.py:64-71): separately rounded dot product, ATen norms,
// rounded product of the norms, IEEE division.
__device__ __forceinline__ float trrosetta_phi_exact(V3 ba, V3 bc) {
    return acosf(__fdiv_rn(dot3(ba, bc), __fmul_rn(norm3(ba), norm3(bc))));
}

// acos of two cosines at once (|c| <= 1; anything else, NaN included, gives NaN), one range and no select:
// acos(|c|) = sqrt(1 - |c|) * P(|c|) with the degree-6 minimax polynomial of acos(x) / sqrt(1 - x) on [0, 1]
// (|error| of the product 9e-8 rad; 5e-7 rad with fp32 rounding and the 1-ulp MUFU.SQRT), mirrored for c < 0.
// Split like the atan2 above.
struct AcosPair {
    float2 a, q, p;
};
__device__ __forceinline__ void acos_prepare(float2 c, AcosPair& e) {
    e.a = make_float2(fabsf(c.x), fabsf(c.y));
    const float2 z = __fadd2_rn(f2(1.0f), neg2(e.a));  // negative for |c| > 1: the square root is NaN
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(e.q.x) : "f"(z.x));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(e.q.y) : "f"(z.y));
    e.p = __ffma2_rn(f2(2.6117212e-03f), e.a, f2(-1.2003397e-02f));
}
__device__ __forceinline__ float2 acos_finish(const AcosPair& e, float2 c) {
    const float2 r = __fmul2_rn(e.q, e.p);
    float o0 = r.x, o1 = r.y;
    if (c.x < 0.f) o0 = 3.14159274f - o0;
    if (c.y < 0.f) o1 = 3.14159274f - o1;
    return make_float2(o0, o1);
}

// Row-side record of one residue i (16 floats, 64-byte aligned: four broadcast LDS.128 per row) — everything that depends
// on residue i alone, in the form that leaves the fewest FP32 operations per pair:
//   q0 = b0 = CA - CB (3), 1 / |b0| (inf for b0 = 0)
//   q1 = CB (3), flags (1 = CA_i or CB_i missing: the whole row is NaN; 2 = N_i missing: theta of the row is NaN)
//   q2 = tm = (CB - CA) x tn1 (3) with tn1 = (N - CA) x (CB - CA): theta's cosine term is the triple product
//        tn1 . (bc x (CB - CA)) = bc . tm — one dot product per pair instead of a cross product and a dot product;
//        omega[i, i]
//   q3 = ts = tn1 |b0| (3; NaN for b0 = 0, where the reference divides 0 / 0): theta's NEGATED sine term is bc . ts;
//        theta[i, i]                                  (phi[i, i] is always NaN; the diagonal is written after the loop)
constexpr int kRowRecord = 16;

// omega / theta / phi of one row record against the thread's pair of residues j (both pairs at once).
struct JPair {
    P3 ca, cb, b2;      // CA_j, CB_j, omega's b2 = CB_j - CA_j
    bool nan0, nan1;    // CB_j missing (lane x / lane y)
    int diag_k;         // row (relative to the CTA's first row) whose diagonal entry is lane x of this pair
};

// Everything one (row, pair of j) evaluation produces before any branch: the three angle pairs from the straight-line
// path, what the rare per-lane patches need (negated sine / cosine terms, phi's cosine, bc, ba), and ONE flag saying
// that some lane is out of range of the straight-line path (which lanes is worked out again in the patch: keeping six
// flags alive through the loop body cost more predicate-spill instructions than the geometry has multiplications).
struct RowEval {
    float2 w, t, f;              // omega, theta, phi (lanes x / y = residues j / j + 1)
    float2 nyw, xw, nyt, xt, c;  // atan2 operands of omega and theta (sine terms negated), phi's fast cosine
    P3 bc;                       // CB_j - CB_i
    float bax, bay, baz;         // CA_i - CB_i
    bool bad;
};

// Straight-line part (no branch): independent chains for omega, theta, phi that the scheduler interleaves.
template <bool ALL3>
__device__ __forceinline__ void eval_row_core(const float4 q0, const float4 q1, const float4 q2, const float4 q3,
                                              const JPair& jp, bool want_omega, bool want_theta, bool want_phi,
                                              RowEval& r) {
    const float2 nan2 = f2(__int_as_float(0x7fc00000));
    r.w = r.t = r.f = nan2;
    r.nyw = r.xw = r.nyt = r.xt = f2(1.0f);
    r.c = f2(0.0f);
    r.bax = q0.x; r.bay = q0.y; r.baz = q0.z;
    const P3 b0{f2(q0.x), f2(q0.y), f2(q0.z)};
    const P3 cbi{f2(q1.x), f2(q1.y), f2(q1.z)};
    r.bc = sub_p3(jp.cb, cbi);  // CB_j - CB_i: theta's b2, phi's bc
    // ---- geometry: the operands of the two atan2 and of the acos
    if (ALL3 || want_omega) {
        const P3 b1 = sub_p3(jp.ca, cbi);
        const P3 n1 = cross_p3(b0, b1);
        const P3 n2 = cross_p3(jp.b2, b1);
        r.xw = dot_p3(n1, n2);
        const float2 sn = dot_p3(n1, jp.b2);
        const float2 bb = dot_p3(b1, b1);
        const float2 nb1 = __fmul2_rn(bb, make_float2(rsqrt_mufu(bb.x), rsqrt_mufu(bb.y)));  // |b1|, NaN at 0
        r.nyw = __fmul2_rn(sn, nb1);  // y = -(n1 . b2) |b1|
    }
    if (ALL3 || want_theta) {  // a missing N_i makes tm / ts NaN: the lanes fail the range test and are patched to NaN
        r.xt = dot_p3(P3{f2(q2.x), f2(q2.y), f2(q2.z)}, r.bc);
        r.nyt = dot_p3(P3{f2(q3.x), f2(q3.y), f2(q3.z)}, r.bc);
    }
    if (ALL3 || want_phi) {
        // cos = (b0 . bc) / (|b0| |bc|) with the raw MUFU.RSQ of |bc|^2 (relative error 2^-22.9: 1.3e-6 rad at the
        // sin(phi) = 0.1 edge of the stated tolerance range; a Newton step on it cost four FP32-pipe instructions)
        const float2 d = dot_p3(b0, r.bc);
        const float2 cc = dot_p3(r.bc, r.bc);
        const float2 rs = make_float2(rsqrt_mufu(cc.x), rsqrt_mufu(cc.y));
        r.c = __fmul2_rn(__fmul2_rn(d, f2(q0.w)), rs);
    }
    // ---- the three function evaluations, their Horner chains interleaved step by step
    AtanPair ew, et;
    AcosPair ef;
    atan2_prepare(r.nyw, r.xw, ew);
    atan2_prepare(r.nyt, r.xt, et);
    acos_prepare(r.c, ef);
    ew.p = __ffma2_rn(ew.p, ew.s, f2(8.210079680e-02f));
    et.p = __ffma2_rn(et.p, et.s, f2(8.210079680e-02f));
    ef.p = __ffma2_rn(ef.p, ef.a, f2(2.7762914e-02f));
    ew.p = __ffma2_rn(ew.p, ew.s, f2(-1.339595112e-01f));
    et.p = __ffma2_rn(et.p, et.s, f2(-1.339595112e-01f));
    ef.p = __ffma2_rn(ef.p, ef.a, f2(-4.919744e-02f));
    ew.p = __ffma2_rn(ew.p, ew.s, f2(1.986158291e-01f));
    et.p = __ffma2_rn(et.p, et.s, f2(1.986158291e-01f));
    ef.p = __ffma2_rn(ef.p, ef.a, f2(8.883589e-02f));
    ew.p = __ffma2_rn(ew.p, ew.s, f2(-3.332545806e-01f));
    et.p = __ffma2_rn(et.p, et.s, f2(-3.332545806e-01f));
    ef.p = __ffma2_rn(ef.p, ef.a, f2(-2.1459109e-01f));
    ef.p = __ffma2_rn(ef.p, ef.a, f2(1.5707963f));
    if (ALL3 || want_omega) r.w = atan2_finish(ew, r.nyw, r.xw);
    if (ALL3 || want_theta) r.t = atan2_finish(et, r.nyt, r.xt);
    if (ALL3 || want_phi) r.f = acos_finish(ef, r.c);
    // one range test for the four atan2 and the two acos of the iteration (NaN fails every comparison; an angle that
    // was not asked for was given in-range operands above)
    const float lo = min3_nan(ew.mx0, ew.mx1, min_nan(et.mx0, et.mx1));
    const float hi = max3_nan(ew.mx0, ew.mx1, max_nan(et.mx0, et.mx1));
    const float cm = max_nan(fabsf(r.c.x), fabsf(r.c.y));
    r.bad = !((lo > kAtanLo) & (hi < kAtanHi) & (cm <= kCosExact));
}

// Rare part: lanes out of range (zeros: the diagonal, zero-padded residues, coincident atoms; NaN: a missing atom in one of
// the two pairs or in the row; |cos| within 1e-4 of 1) are found again and redone one by one.  dk = row - jp.diag_k:
// lane x (dk = 0) or lane y (dk = 1) is the diagonal entry, whose three values are written after the row loop.
__device__ __forceinline__ void eval_row_patch(RowEval& r, const JPair& jp, int dk) {
    const bool live0 = dk != 0, live1 = dk != 1;
    if (live0 && !atan2_in_range(r.nyw.x, r.xw.x)) r.w.x = atan2_slow(r.nyw.x, r.xw.x);
    if (live1 && !atan2_in_range(r.nyw.y, r.xw.y)) r.w.y = atan2_slow(r.nyw.y, r.xw.y);
    if (live0 && !atan2_in_range(r.nyt.x, r.xt.x)) r.t.x = atan2_slow(r.nyt.x, r.xt.x);
    if (live1 && !atan2_in_range(r.nyt.y, r.xt.y)) r.t.y = atan2_slow(r.nyt.y, r.xt.y);
    const V3 ba{r.bax, r.bay, r.baz};
    // (a lane whose CB_j is missing is NaN either way and needs no exact evaluation)
    if (live0 && !jp.nan0 && !(fabsf(r.c.x) <= kCosExact)) r.f.x = trrosetta_phi_exact(ba, V3{r.bc.x.x, r.bc.y.x, r.bc.z.x});
    if (live1 && !jp.nan1 && !(fabsf(r.c.y) <= kCosExact)) r.f.y = trrosetta_phi_exact(ba, V3{r.bc.x.y, r.bc.y.y, r.bc.z.y});
}

// Loop order: a thread OWNS pairs of residues j (one pair when L <= 2 * blockDim.x) and walks the CTA's rows with them
// in registers, so per (row, pair of j) the only memory traffic is the broadcast read of the row record and the three
// 64-bit stores; nothing that depends on j alone (its coordinates, its NaN flags, the position of the diagonal) is
// redone per row.  (The first packed version walked j inside a row: with one j-pair per thread and row it re-read the
// row record, the coordinates and the flags for every pair and spent a quarter of its issue slots on addressing.)
// MIN_CTAS = resident CTAs per SM the kernel is compiled for (its register budget): a row is one long dependent chain
// (differences -> cross products -> dot products -> MUFU -> polynomial), so resident warps are what fills the issue
// slots.  Measured at BASELINE config 3 with the round's first loop body: 3 CTAs (77 registers) 0.44 ms, 4 (64) 0.41,
// 5 (51, spills) 0.44, 6 (42) 0.50; two rows per iteration at 120 registers (2 CTAs) 0.53.
template <bool VIRTUAL_CB, bool ALL3, int MIN_CTAS>
__global__ void __launch_bounds__(256, MIN_CTAS)
    trrosetta_fast_kernel(
    const float* __restrict__ xyz, float* __restrict__ omega, float* __restrict__ theta, float* __restrict__ phi, int L,
    int A, int rows_per_cta, int blocks_per_structure, int vector_stores) {
    extern __shared__ __align__(16) float fast_smem[];
    const int Lp = (L + 1) & ~1;  // residues j padded to whole pairs
    float* const srow = fast_smem;  // rows_per_cta records of kRowRecord floats (16-byte aligned)
    float* const sca_x = srow + rows_per_cta * kRowRecord;
    float* const sca_y = sca_x + Lp;
    float* const sca_z = sca_y + Lp;
    float* const scb_x = sca_z + Lp;
    float* const scb_y = scb_x + Lp;
    float* const scb_z = scb_y + Lp;
    // per residue j: bit 0 = CB missing (NaN coordinate): all three angles of the pair are NaN
    unsigned char* const sflag = reinterpret_cast<unsigned char*>(scb_z + Lp);

    const long long b = blockIdx.x / blocks_per_structure;
    const int row0 = (blockIdx.x - static_cast<int>(b) * blocks_per_structure) * rows_per_cta;
    const int nrows = min(rows_per_cta, L - row0);
    const float* __restrict__ xb = xyz + b * L * A * 3;
    const bool want_omega = ALL3 || omega != nullptr, want_theta = ALL3 || theta != nullptr, want_phi = ALL3 || phi != nullptr;

    // ---- stage residue j of the whole structure: CA and CB (real or virtual), structure of arrays
    bool special = false;
    for (int r = threadIdx.x; r < Lp; r += blockDim.x) {
        V3 ca{0.f, 0.f, 0.f}, cb{0.f, 0.f, 0.f};
        if (r < L) {
            const float* __restrict__ xr = xb + static_cast<long long>(r) * A * 3;
            ca = ld3(xr + 3);
            cb = VIRTUAL_CB ? virtual_cb(ld3(xr + 0), ca, ld3(xr + 6)) : ld3(xr + 12);
        }
        sca_x[r] = ca.x; sca_y[r] = ca.y; sca_z[r] = ca.z;
        scb_x[r] = cb.x; scb_y[r] = cb.y; scb_z[r] = cb.z;
        // bit 0: CB missing (all three angles of the pair are NaN) — also set for the padding residue of an odd L;
        // bit 1: CA missing (omega is NaN); bit 2: CA = CB = 0 exactly (a zero-padded residue of a ragged batch: b2 = 0,
        // omega is 0 — or NaN opposite a CB at the origin — by exact cancellation); bit 3: CA = 0 exactly
        const bool ca_zero = ca.x == 0.f && ca.y == 0.f && ca.z == 0.f, cb_zero = cb.x == 0.f && cb.y == 0.f && cb.z == 0.f;
        sflag[r] = (r >= L || atom_has_nan(cb) ? 1 : 0) | (atom_has_nan(ca) ? 2 : 0) | (r < L && ca_zero && cb_zero ? 4 : 0) |
                   (r < L && ca_zero ? 8 : 0);
        special |= sflag[r] != 0;
    }
    // Does the structure hold ANY residue that needs the special handling below (a missing CA / CB, a zero-padded
    // residue, the padding lane of an odd L)?  If not — every structure of a clean, full-length batch — the CTA runs
    // the row loop without the flag tests and the lane masks (8 % of the loop's instructions).
    const bool cta_special = __syncthreads_or(special) != 0;
    // ---- row-side records of this CTA's residues i
    for (int k = threadIdx.x; k < nrows; k += blockDim.x) {
        const int i = row0 + k;
        const V3 n_i = ld3(xb + static_cast<long long>(i) * A * 3);
        const V3 ca{sca_x[i], sca_y[i], sca_z[i]}, cb{scb_x[i], scb_y[i], scb_z[i]};
        const V3 b0 = sub3(ca, cb);                       // omega's b0, phi's ba;  theta's b1 is exactly -b0
        const V3 u = sub3(n_i, ca);
        const V3 tb1{-b0.x, -b0.y, -b0.z};
        V3 tn1;                                           // (N - CA) x (CB - CA), fused
        tn1.x = fmaf(u.y, tb1.z, -(u.z * tb1.y));
        tn1.y = fmaf(u.z, tb1.x, -(u.x * tb1.z));
        tn1.z = fmaf(u.x, tb1.y, -(u.y * tb1.x));
        V3 tm;                                            // (CB - CA) x tn1, fused
        tm.x = fmaf(tb1.y, tn1.z, -(tb1.z * tn1.y));
        tm.y = fmaf(tb1.z, tn1.x, -(tb1.x * tn1.z));
        tm.z = fmaf(tb1.x, tn1.y, -(tb1.y * tn1.x));
        const float bb = fmaf(b0.z, b0.z, fmaf(b0.y, b0.y, b0.x * b0.x));
        const float inv = rsqrt_refined(bb);              // inf * 0 = NaN for bb = 0, NaN for missing atoms
        const float inv_or_inf = bb == 0.f ? __int_as_float(0x7f800000) : inv;
        const float nb0 = bb * inv;                       // |b0|, NaN where the reference divides 0 / 0
        const float omega_ii = 0.0f * inv_or_inf;         // 0, or NaN (missing / coincident CA, CB)
        float* rec = srow + k * kRowRecord;
        rec[0] = b0.x; rec[1] = b0.y; rec[2] = b0.z;
        rec[3] = inv_or_inf;                              // 1 / |b0|: 0 * inf = NaN for phi, as 0 / 0
        rec[4] = cb.x; rec[5] = cb.y; rec[6] = cb.z;
        // flags: 1 = CA_i / CB_i missing (the row is NaN), 2 = N_i missing, 4 = CA_i = CB_i = 0 exactly (a zero-padded
        // residue: omega is 0 — NaN opposite a CA at the origin —, theta and phi are 0 / 0 = NaN), 8 = CB_i = 0 exactly
        const bool ca_zero = ca.x == 0.f && ca.y == 0.f && ca.z == 0.f, cb_zero = cb.x == 0.f && cb.y == 0.f && cb.z == 0.f;
        rec[7] = __int_as_float((atom_has_nan(ca) || atom_has_nan(cb) ? 1 : 0) | (atom_has_nan(n_i) ? 2 : 0) |
                                (ca_zero && cb_zero ? 4 : 0) | (cb_zero ? 8 : 0));
        rec[8] = tm.x; rec[9] = tm.y; rec[10] = tm.z;
        rec[11] = omega_ii;
        rec[12] = tn1.x * nb0; rec[13] = tn1.y * nb0; rec[14] = tn1.z * nb0;
        rec[15] = omega_ii + 0.0f * tn1.x + 0.0f * tn1.y + 0.0f * tn1.z;  // theta[i, i]: NaN also without N_i
    }
    __syncthreads();

    const float2* __restrict__ pca_x = reinterpret_cast<const float2*>(sca_x);
    const float2* __restrict__ pca_y = reinterpret_cast<const float2*>(sca_y);
    const float2* __restrict__ pca_z = reinterpret_cast<const float2*>(sca_z);
    const float2* __restrict__ pcb_x = reinterpret_cast<const float2*>(scb_x);
    const float2* __restrict__ pcb_y = reinterpret_cast<const float2*>(scb_y);
    const float2* __restrict__ pcb_z = reinterpret_cast<const float2*>(scb_z);
    const float4* __restrict__ rows4 = reinterpret_cast<const float4*>(srow);
    const int npairs = Lp >> 1;
    const long long first_out = (b * L + row0) * L;  // element (b, row0, 0) of the outputs
    const float2 nan2 = f2(__int_as_float(0x7fc00000));

    for (int jpi = threadIdx.x; jpi < npairs; jpi += blockDim.x) {
        JPair jp;
        jp.ca = P3{pca_x[jpi], pca_y[jpi], pca_z[jpi]};
        jp.cb = P3{pcb_x[jpi], pcb_y[jpi], pcb_z[jpi]};
        // Missing atoms are NaN coordinates (protstruc/pdb.py:133-135) and make the angle NaN whatever the other atoms
        // are: such pairs (and rows) are answered without arithmetic instead of dragging NaN through the IEEE
        // fall-backs of atan2 / acos.  A thread with ONE lane missing an atom (a glycine next to any other residue: on
        // real data nearly every warp holds such a thread) evaluates that lane on stand-in coordinates — its partner's
        // CB, a point next to CB for a missing CA — so that it never leaves the straight-line path, and overwrites the
        // lane's results with NaN before the store (one predicated branch per row for everybody else).
        const unsigned short fl = reinterpret_cast<const unsigned short*>(sflag)[jpi];
        const bool cbn0 = fl & 0x0001, cbn1 = fl & 0x0100, can0 = fl & 0x0002, can1 = fl & 0x0200;
        // Zero-padded residues (ragged batches: xyz = 0 beyond a structure's length) are the other everyday source of
        // out-of-range lanes: opposite a residue with CA = CB = 0, omega's b2 vanishes and the reference gets x = y = 0
        // -> 0 by exact cancellation (NaN if CB_i is at the origin too: |b1| = 0).  Such a lane gets a stand-in CA for
        // the arithmetic and the known omega at the store (theta and phi are ordinary values there: CB_j = 0 is a point).
        const bool zero0 = fl & 0x0004, zero1 = fl & 0x0400, caz0 = fl & 0x0008, caz1 = fl & 0x0800;
        const bool pair_nan = cbn0 && cbn1;
        const bool lane_nan = (fl & 0x0707) != 0 && !pair_nan;
        if (lane_nan) {
            if (cbn0) { jp.cb.x.x = jp.cb.x.y; jp.cb.y.x = jp.cb.y.y; jp.cb.z.x = jp.cb.z.y; }
            if (cbn1) { jp.cb.x.y = jp.cb.x.x; jp.cb.y.y = jp.cb.y.x; jp.cb.z.y = jp.cb.z.x; }
            if (can0 || cbn0 || zero0) { jp.ca.x.x = jp.cb.x.x + 1.0f; jp.ca.y.x = jp.cb.y.x + 0.25f; jp.ca.z.x = jp.cb.z.x + 0.5f; }
            if (can1 || cbn1 || zero1) { jp.ca.x.y = jp.cb.x.y + 1.0f; jp.ca.y.y = jp.cb.y.y + 0.25f; jp.ca.z.y = jp.cb.z.y + 0.5f; }
        }
        jp.b2 = sub_p3(jp.cb, jp.ca);
        jp.nan0 = jp.nan1 = false;  // no lane carries NaN from the column side any more
        const int j = 2 * jpi;
        jp.diag_k = j - row0;
        const float qnan = __int_as_float(0x7fc00000);
        auto mask_lanes = [&](RowEval& r, int row_flags) {
            if (lane_nan) {
                const float zero_omega = (row_flags & 8) ? qnan : 0.f;  // opposite a zero-padded residue
                if (cbn0) r.w.x = r.t.x = r.f.x = qnan;
                else if (can0) r.w.x = qnan;
                else if (zero0) r.w.x = zero_omega;
                if (cbn1) r.w.y = r.t.y = r.f.y = qnan;
                else if (can1) r.w.y = qnan;
                else if (zero1) r.w.y = zero_omega;
            }
        };
        const bool second = vector_stores || (j + 1 < L);  // lane y is a residue of the structure
        // running output pointers of (row0 + k, j): one 64-bit add per feature and row
        float* pw = (ALL3 || want_omega) ? omega + first_out + j : nullptr;
        float* pt = (ALL3 || want_theta) ? theta + first_out + j : nullptr;
        float* pf = (ALL3 || want_phi) ? phi + first_out + j : nullptr;
        auto store = [&](long long at, float2 w, float2 t, float2 f) {  // at = 0 or L: the row at the pointers, or the next
            if (vector_stores) {  // L even, outputs 8-byte aligned: one 64-bit store per feature
                if (ALL3 || want_omega) *reinterpret_cast<float2*>(pw + at) = w;
                if (ALL3 || want_theta) *reinterpret_cast<float2*>(pt + at) = t;
                if (ALL3 || want_phi) *reinterpret_cast<float2*>(pf + at) = f;
            } else {
                if (ALL3 || want_omega) pw[at] = w.x;
                if (ALL3 || want_theta) pt[at] = t.x;
                if (ALL3 || want_phi) pf[at] = f.x;
                if (second) {
                    if (ALL3 || want_omega) pw[at + 1] = w.y;
                    if (ALL3 || want_theta) pt[at + 1] = t.y;
                    if (ALL3 || want_phi) pf[at + 1] = f.y;
                }
            }
        };
        auto one_row = [&](auto special_tag, const float4* rec, int k, long long at) {
            constexpr bool kSpecial = decltype(special_tag)::value;
            const float4 q1 = rec[1];
            if constexpr (!kSpecial) {  // a structure without missing atoms / zero padding: no flag can be set
                RowEval r;
                eval_row_core<ALL3>(rec[0], q1, rec[2], rec[3], jp, want_omega, want_theta, want_phi, r);
                if (r.bad) eval_row_patch(r, jp, k - jp.diag_k);
                store(at, r.w, r.t, r.f);
                return;
            }
            const int row_flags = __float_as_int(q1.w);
            if (pair_nan || (row_flags & 5)) {  // ONE test on the straight-line path for both kinds of answered rows
                if (pair_nan || (row_flags & 1)) {
                    store(at, nan2, nan2, nan2);
                    return;
                }
                // zero-padded residue i: b0 = 0, so omega is 0 by exact cancellation (NaN where b1 = CA_j vanishes too),
                // theta and phi are 0 / 0; no arithmetic
                RowEval z;
                z.w = make_float2(caz0 ? qnan : 0.f, caz1 ? qnan : 0.f);
                z.t = z.f = nan2;
                mask_lanes(z, row_flags);
                store(at, z.w, z.t, z.f);
                return;
            }
            const float4 q0 = rec[0], q2 = rec[2], q3 = rec[3];
            RowEval r;
            eval_row_core<ALL3>(q0, q1, q2, q3, jp, want_omega, want_theta, want_phi, r);
            if (r.bad) eval_row_patch(r, jp, k - jp.diag_k);
            mask_lanes(r, row_flags);
            store(at, r.w, r.t, r.f);
        };
        const float4* rec = rows4;
        int k = 0;
        if constexpr (MIN_CTAS == 2) {
            // tuning variant: TWO rows per iteration — two independent dependency chains per warp at 16 warps per SM
            for (; k + 1 < nrows; k += 2, rec += 8, pw += 2 * L, pt += 2 * L, pf += 2 * L) {
                const float4 q1a = rec[1], q1b = rec[5];
                if (pair_nan || ((__float_as_int(q1a.w) | __float_as_int(q1b.w)) & 5)) {
                    one_row(std::true_type{}, rec, k, 0);
                    one_row(std::true_type{}, rec + 4, k + 1, L);
                    continue;
                }
                RowEval ra, rb;
                eval_row_core<ALL3>(rec[0], q1a, rec[2], rec[3], jp, want_omega, want_theta, want_phi, ra);
                eval_row_core<ALL3>(rec[4], q1b, rec[6], rec[7], jp, want_omega, want_theta, want_phi, rb);
                if (ra.bad | rb.bad) {
                    if (ra.bad) eval_row_patch(ra, jp, k - jp.diag_k);
                    if (rb.bad) eval_row_patch(rb, jp, k + 1 - jp.diag_k);
                }
                mask_lanes(ra, __float_as_int(q1a.w));
                mask_lanes(rb, __float_as_int(q1b.w));
                store(0, ra.w, ra.t, ra.f);
                store(L, rb.w, rb.t, rb.f);
            }
        }
        if (cta_special) {
            for (; k < nrows; ++k, rec += 4, pw += L, pt += L, pf += L) one_row(std::true_type{}, rec, k, 0);
        } else {
            for (; k < nrows; ++k, rec += 4, pw += L, pt += L, pf += L) one_row(std::false_type{}, rec, k, 0);
        }
        // the diagonal entries (row0 + dk, j) and (row0 + dk + 1, j + 1) of this pair of columns, if the CTA owns them
        const int dk = jp.diag_k;
#pragma unroll
        for (int lane = 0; lane < 2; ++lane) {
            const int k = dk + lane;
            if (k < 0 || k >= nrows || j + lane >= L) continue;
            const float* rec_k = srow + k * kRowRecord;
            const long long o = first_out + static_cast<long long>(k) * L + j + lane;
            if (ALL3 || want_omega) omega[o] = rec_k[11];
            if (ALL3 || want_theta) theta[o] = rec_k[15];
            if (ALL3 || want_phi) phi[o] = nan2.x;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// K2, second generation: pairwise_dihedrals / pairwise_planar_angles for ANY atom-slot lists on the packed FP32 pipe.
//
// pair_angles_kernel above issues the reference's operation sequence literally and is issue-bound at ~110
// lane-instructions per pair for ONE angle (profiles/r3i_all_kernels_ncu_table_before_packed_generic.txt: 0.14 of the HBM roof, issue-active
// 84 %).  This kernel evaluates the same angle the way the packed trRosetta kernel does — a thread owns two consecutive
// residues j and walks the CTA's rows; every subtraction / product is an FADD2 / FMUL2 / FFMA2 for both pairs; the sine
// term is y = -(n1 . b2) |b1| (no third cross product); atan2 / acos are the packed polynomial evaluations with ONE
// range test per iteration and a rare per-lane patch (atan2f on the packed operands; the reference's exact cosine
// sequence within 1e-4 of |cos| = 1, where the unclamped arccos decides between a value and NaN) — under the same
// contract: NaN placement of the reference, <= 1e-5 rad where min sin(bond angle) >= 0.1.
// What keeps the special cases right:
//  * rows / residues j with a missing (NaN) atom among the requested slots are answered NaN without arithmetic;
//  * the diagonal j = i is where two requested points can be THE SAME ATOM (omega's CA_i, CB_i, CA_i, CB_i): exact
//    cancellations (b0 x b0 = 0) that fused multiply-adds do not reproduce, so the diagonal entry of every row is
//    evaluated by the exact-sequence dihedral4 / angle3 after the row loop;
//  * zero-padded residues: a zero vector makes every fused product exactly zero as well, |b1| = b1.b1 * rsqrt(b1.b1)
//    is NaN for b1 = 0 exactly where the reference divides 0 / 0, and x = y = 0 reaches atan2f like in the reference.
// NI (points taken from residue i) is a template parameter: differences of two row-side points are loop invariants of
// the j walk only in the sense of scalar work per row; differences of two column-side points are hoisted by the compiler.
constexpr int kAngleRecord = 12;  // floats per row record: up to three points (9) + flags (bit 0: a point is NaN)

template <int KIND, int NI>
__global__ void __launch_bounds__(256, 3) pair_angles_fast_kernel(
    const float* __restrict__ xyz, float* __restrict__ out, int L, int A, SlotList sl, int rows_per_cta,
    int blocks_per_structure, int vector_stores) {
    constexpr int N = KIND == PS_ANGLE_DIHEDRAL ? 4 : 3;
    constexpr int NJ = N - NI;
    static_assert(NI >= 1 && NJ >= 1 && NI <= 3, "points must come from both residues");
    extern __shared__ __align__(16) float fast_smem[];
    const int Lp = (L + 1) & ~1;
    float* const srow = fast_smem;                                   // rows_per_cta records
    float* const sj = srow + rows_per_cta * kAngleRecord;            // [NJ][3][Lp] structure of arrays
    unsigned char* const sflag = reinterpret_cast<unsigned char*>(sj + NJ * 3 * Lp);

    const long long b = blockIdx.x / blocks_per_structure;
    const int row0 = (blockIdx.x - static_cast<int>(b) * blocks_per_structure) * rows_per_cta;
    const int nrows = min(rows_per_cta, L - row0);
    const float* __restrict__ xb = xyz + b * L * A * 3;

    for (int r = threadIdx.x; r < Lp; r += blockDim.x) {
        bool bad = false;
#pragma unroll
        for (int k = 0; k < NJ; ++k) {
            V3 a{0.f, 0.f, 0.f};
            if (r < L) a = ld3(xb + (static_cast<long long>(r) * A + sl.s[NI + k]) * 3);
            sj[(k * 3 + 0) * Lp + r] = a.x;
            sj[(k * 3 + 1) * Lp + r] = a.y;
            sj[(k * 3 + 2) * Lp + r] = a.z;
            bad |= atom_has_nan(a);
        }
        sflag[r] = (bad || r >= L) ? 1 : 0;  // (the padding residue of an odd L counts as missing)
    }
    for (int k = threadIdx.x; k < nrows; k += blockDim.x) {
        float* rec = srow + k * kAngleRecord;
        bool bad = false;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            V3 v{0.f, 0.f, 0.f};
            if (a < NI) v = ld3(xb + (static_cast<long long>(row0 + k) * A + sl.s[a]) * 3);
            rec[3 * a + 0] = v.x; rec[3 * a + 1] = v.y; rec[3 * a + 2] = v.z;
            bad |= atom_has_nan(v);
        }
        rec[9] = __int_as_float(bad ? 1 : 0);
        rec[10] = rec[11] = 0.f;
        // Everything that depends on residue i alone is formed here, once per row, instead of once per thread and row:
        //   dihedral with three points of residue i: p2, |b1|; b1 = p2 - p1, flags; n1 = (p0 - p1) x b1
        //   planar angle with two points of residue i: p1, |ba|^2; ba = p0 - p1, flags
        if (KIND == PS_ANGLE_DIHEDRAL && NI == 3) {
            const V3 p0{rec[0], rec[1], rec[2]}, p1{rec[3], rec[4], rec[5]}, p2{rec[6], rec[7], rec[8]};
            const V3 b0 = sub3(p0, p1), b1 = sub3(p2, p1);
            const float bb = fmaf(b1.z, b1.z, fmaf(b1.y, b1.y, b1.x * b1.x));
            rec[0] = p2.x; rec[1] = p2.y; rec[2] = p2.z;
            rec[3] = bb * rsqrt_mufu(bb);  // |b1|, NaN where the reference divides 0 / 0
            rec[4] = b1.x; rec[5] = b1.y; rec[6] = b1.z;
            rec[7] = __int_as_float(bad ? 1 : 0);
            rec[8] = fmaf(b0.y, b1.z, -(b0.z * b1.y));
            rec[9] = fmaf(b0.z, b1.x, -(b0.x * b1.z));
            rec[10] = fmaf(b0.x, b1.y, -(b0.y * b1.x));
        } else if (KIND == PS_ANGLE_PLANAR && NI == 2) {
            const V3 p0{rec[0], rec[1], rec[2]}, p1{rec[3], rec[4], rec[5]};
            const V3 ba = sub3(p0, p1);
            rec[0] = p1.x; rec[1] = p1.y; rec[2] = p1.z;
            rec[3] = fmaf(ba.z, ba.z, fmaf(ba.y, ba.y, ba.x * ba.x));
            rec[4] = ba.x; rec[5] = ba.y; rec[6] = ba.z;
            rec[7] = __int_as_float(bad ? 1 : 0);
            rec[8] = p0.x; rec[9] = p0.y; rec[10] = p0.z;  // (the exact-sequence fall-back wants the points themselves)
        }
    }
    __syncthreads();
    constexpr bool kRowSide = (KIND == PS_ANGLE_DIHEDRAL && NI == 3) || (KIND == PS_ANGLE_PLANAR && NI == 2);

    const float4* __restrict__ rows4 = reinterpret_cast<const float4*>(srow);
    const int npairs = Lp >> 1;
    const long long first_out = (b * L + row0) * L;
    const float2 nan2 = f2(__int_as_float(0x7fc00000));

    for (int jpi = threadIdx.x; jpi < npairs; jpi += blockDim.x) {
        P3 pj[NJ];
#pragma unroll
        for (int k = 0; k < NJ; ++k) {
            pj[k].x = reinterpret_cast<const float2*>(sj + (k * 3 + 0) * Lp)[jpi];
            pj[k].y = reinterpret_cast<const float2*>(sj + (k * 3 + 1) * Lp)[jpi];
            pj[k].z = reinterpret_cast<const float2*>(sj + (k * 3 + 2) * Lp)[jpi];
        }
        const unsigned short fl = reinterpret_cast<const unsigned short*>(sflag)[jpi];
        const bool miss0 = fl & 0x0001, miss1 = fl & 0x0100;
        const bool pair_nan = miss0 && miss1;
        // ONE lane missing an atom: it is evaluated on its partner's coordinates (so that it never leaves the
        // straight-line path) and overwritten with NaN before the store
        const bool lane_nan = (miss0 || miss1) && !pair_nan;
        if (lane_nan) {
#pragma unroll
            for (int k = 0; k < NJ; ++k) {
                if (miss0) { pj[k].x.x = pj[k].x.y; pj[k].y.x = pj[k].y.y; pj[k].z.x = pj[k].z.y; }
                else { pj[k].x.y = pj[k].x.x; pj[k].y.y = pj[k].y.x; pj[k].z.y = pj[k].z.x; }
            }
        }
        constexpr bool nan0 = false, nan1 = false;  // no lane carries NaN from the column side any more
        const int j = 2 * jpi;
        const bool second = vector_stores || (j + 1 < L);
        float* po = out + first_out + j;
        auto store = [&](float2 v) {
            if (lane_nan) {
                if (miss0) v.x = __int_as_float(0x7fc00000);
                else v.y = __int_as_float(0x7fc00000);
            }
            if (vector_stores) {
                *reinterpret_cast<float2*>(po) = v;
            } else {
                po[0] = v.x;
                if (second) po[1] = v.y;
            }
        };
        const int dk = j - row0;  // row (relative to the CTA's first) whose diagonal entry is lane x of this pair
        const float4* rec = rows4;
        for (int k = 0; k < nrows; ++k, rec += 3, po += L) {
            const float4 q2 = rec[2], q1 = rec[1];
            if (pair_nan || (__float_as_int(kRowSide ? q1.w : q2.y) & 1)) {
                store(nan2);
                continue;
            }
            const float4 q0 = rec[0];
            const float ri[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x};
            P3 pt[N];
#pragma unroll
            for (int a = 0; a < N; ++a) {
                if (a < NI) pt[a] = P3{f2(ri[3 * a]), f2(ri[3 * a + 1]), f2(ri[3 * a + 2])};
                else pt[a] = pj[a - NI];
            }
            float2 res;
            if (KIND == PS_ANGLE_DIHEDRAL) {
                float2 x, ny;
                if constexpr (NI == 3) {  // row-side record: p2, |b1|, b1, n1
                    const P3 b1{f2(q1.x), f2(q1.y), f2(q1.z)};
                    const P3 n1{f2(q2.x), f2(q2.y), f2(q2.z)};
                    const P3 b2 = sub_p3(pj[0], P3{f2(q0.x), f2(q0.y), f2(q0.z)});
                    const P3 n2 = cross_p3(b2, b1);
                    x = dot_p3(n1, n2);
                    ny = __fmul2_rn(dot_p3(n1, b2), f2(q0.w));  // y = -(n1 . b2) |b1|
                } else {
                    const P3 b0 = sub_p3(pt[0], pt[1]);
                    const P3 b1 = sub_p3(pt[2], pt[1]);
                    const P3 b2 = sub_p3(pt[3], pt[2]);
                    const P3 n1 = cross_p3(b0, b1);
                    const P3 n2 = cross_p3(b2, b1);
                    x = dot_p3(n1, n2);
                    const float2 sn = dot_p3(n1, b2);
                    const float2 bb = dot_p3(b1, b1);
                    const float2 nb1 = __fmul2_rn(bb, make_float2(rsqrt_mufu(bb.x), rsqrt_mufu(bb.y)));  // |b1|, NaN at 0
                    ny = __fmul2_rn(sn, nb1);  // y = -(n1 . b2) |b1|
                }
                AtanPair e;
                atan2_prepare(ny, x, e);
                e.p = __ffma2_rn(e.p, e.s, f2(8.210079680e-02f));
                e.p = __ffma2_rn(e.p, e.s, f2(-1.339595112e-01f));
                e.p = __ffma2_rn(e.p, e.s, f2(1.986158291e-01f));
                e.p = __ffma2_rn(e.p, e.s, f2(-3.332545806e-01f));
                res = atan2_finish(e, ny, x);
                const float lo = min_nan(e.mx0, e.mx1), hi = max_nan(e.mx0, e.mx1);
                if (!((lo > kAtanLo) & (hi < kAtanHi))) {  // rare: zeros, NaN of ONE lane, out-of-range magnitudes
                    // (the diagonal entry — coincident points, zeros — is rewritten after the loop: not patched here)
                    if (k != dk && !atan2_in_range(ny.x, x.x)) res.x = atan2_slow(ny.x, x.x);
                    if (k != dk + 1 && !atan2_in_range(ny.y, x.y)) res.y = atan2_slow(ny.y, x.y);
                }
            } else {
                P3 ba, bc;
                float2 aa;
                if constexpr (NI == 2) {  // row-side record: p1, |ba|^2, ba, p0
                    ba = P3{f2(q1.x), f2(q1.y), f2(q1.z)};
                    bc = sub_p3(pj[0], P3{f2(q0.x), f2(q0.y), f2(q0.z)});
                    aa = f2(q0.w);
                    pt[0] = P3{f2(q2.x), f2(q2.y), f2(q2.z)};
                    pt[1] = P3{f2(q0.x), f2(q0.y), f2(q0.z)};
                } else {
                    ba = sub_p3(pt[0], pt[1]);
                    bc = sub_p3(pt[2], pt[1]);
                    aa = dot_p3(ba, ba);
                }
                const float2 d = dot_p3(ba, bc);
                const float2 nn = __fmul2_rn(aa, dot_p3(bc, bc));  // (|ba| |bc|)^2
                const float2 c = __fmul2_rn(d, make_float2(rsqrt_mufu(nn.x), rsqrt_mufu(nn.y)));
                AcosPair e;
                acos_prepare(c, e);
                e.p = __ffma2_rn(e.p, e.a, f2(2.7762914e-02f));
                e.p = __ffma2_rn(e.p, e.a, f2(-4.919744e-02f));
                e.p = __ffma2_rn(e.p, e.a, f2(8.883589e-02f));
                e.p = __ffma2_rn(e.p, e.a, f2(-2.1459109e-01f));
                e.p = __ffma2_rn(e.p, e.a, f2(1.5707963f));
                res = acos_finish(e, c);
                // safe while the squared norms' product is a normal number and |cos| stays 1e-3 away from 1
                const float cm = max_nan(fabsf(c.x), fabsf(c.y));
                const float nlo = min_nan(nn.x, nn.y), nhi = max_nan(nn.x, nn.y);
                if (!((cm <= kCosExact) & (nlo > kAtanLo) & (nhi < kAtanHi))) {
                    // (a lane whose residue j misses an atom is NaN either way and needs no exact evaluation)
                    // nor does the diagonal entry (0 / 0 for a shared vertex), which is rewritten after the loop
                    if (!nan0 && k != dk && !((fabsf(c.x) <= kCosExact) & (nn.x > kAtanLo) & (nn.x < kAtanHi)))
                        res.x = angle3(V3{pt[0].x.x, pt[0].y.x, pt[0].z.x}, V3{pt[1].x.x, pt[1].y.x, pt[1].z.x},
                                       V3{pt[2].x.x, pt[2].y.x, pt[2].z.x});
                    if (!nan1 && k != dk + 1 && !((fabsf(c.y) <= kCosExact) & (nn.y > kAtanLo) & (nn.y < kAtanHi)))
                        res.y = angle3(V3{pt[0].x.y, pt[0].y.y, pt[0].z.y}, V3{pt[1].x.y, pt[1].y.y, pt[1].z.y},
                                       V3{pt[2].x.y, pt[2].y.y, pt[2].z.y});
                }
            }
            store(res);
        }
        // the diagonal entries (row0 + dk, j) and (row0 + dk + 1, j + 1): every point from the same residue — the
        // exact-sequence evaluation (coincident atoms cancel exactly there)
#pragma unroll
        for (int lane = 0; lane < 2; ++lane) {
            const int k = dk + lane, jj = j + lane;
            if (k < 0 || k >= nrows || jj >= L) continue;
            V3 q[4];
#pragma unroll
            for (int a = 0; a < N; ++a) q[a] = ld3(xb + (static_cast<long long>(jj) * A + sl.s[a]) * 3);
            const float v = KIND == PS_ANGLE_DIHEDRAL ? dihedral4(q[0], q[1], q[2], q[3]) : angle3(q[0], q[1], q[2]);
            out[first_out + static_cast<long long>(k) * L + jj] = v;
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

// variant: 0 = default (the packed kernel when the points come from both residues and the structure fits in shared
// memory), 1 = the exact-sequence kernel (comparison hook, ps_pair_angles_ex).
int pair_angles_variant_impl(const float* xyz, int B, int L, int A, const int* slots_i, int n_i,
                             const int* slots_j, int n_j, int kind, float* out, int variant, cudaStream_t stream) {
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
    PS_REQUIRE(static_cast<long long>(L) * A * 3 < (1ll << 31), PS_ERR_BAD_SHAPE,
               "pair_angles: L*A*3=%lld floats per structure exceed 2^31", static_cast<long long>(L) * A * 3);
    // ---- packed kernel: points from both residues, the structure's column-side atoms fit in shared memory
    if (variant != 1 && n_i >= 1 && n_j >= 1) {
        const int sms = sm_count_for_current_device();
        if (sms < 0) return sms;
        const int Lp = (L + 1) & ~1;
        int fthreads = ((Lp / 2) + 31) / 32 * 32;
        if (fthreads > 256) fthreads = 256;
        int rows_per_cta = 64;
        while (rows_per_cta > 4 && rows / rows_per_cta < 8ll * sms) rows_per_cta /= 2;
        if (rows_per_cta > L) rows_per_cta = L;
        const size_t smem = (static_cast<size_t>(n_j) * 3 * Lp + static_cast<size_t>(rows_per_cta) * kAngleRecord) * sizeof(float) +
                            static_cast<size_t>(Lp + 16);
        const int blocks_per_structure = (L + rows_per_cta - 1) / rows_per_cta;
        const long long ctas = static_cast<long long>(B) * blocks_per_structure;
        if (smem <= 200 * 1024 && ctas < (1ll << 31)) {
            const int vector_stores = (L % 2 == 0) && ((reinterpret_cast<uintptr_t>(out) & 7u) == 0);
#define PS_FAST_ANGLES(KIND, NI)                                                                                      \
    do {                                                                                                              \
        cudaError_t err = cudaFuncSetAttribute(pair_angles_fast_kernel<KIND, NI>,                                     \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);              \
        if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(pair_angles_fast_kernel)");               \
        pair_angles_fast_kernel<KIND, NI><<<static_cast<unsigned>(ctas), fthreads, smem, stream>>>(                   \
            xyz, out, L, A, sl, rows_per_cta, blocks_per_structure, vector_stores);                                   \
    } while (0)
            if (kind == PS_ANGLE_DIHEDRAL) {
                if (n_i == 1) PS_FAST_ANGLES(PS_ANGLE_DIHEDRAL, 1);
                else if (n_i == 2) PS_FAST_ANGLES(PS_ANGLE_DIHEDRAL, 2);
                else PS_FAST_ANGLES(PS_ANGLE_DIHEDRAL, 3);
            } else {
                if (n_i == 1) PS_FAST_ANGLES(PS_ANGLE_PLANAR, 1);
                else PS_FAST_ANGLES(PS_ANGLE_PLANAR, 2);
            }
#undef PS_FAST_ANGLES
            return check_launch("pair_angles_fast_kernel");
        }
    }
    int grid = 0;
    int rc = grid_for_rows(rows, &grid);
    if (rc != PS_OK) return rc;
    const int threads = threads_for_L(L);
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

int pair_angles_impl(const float* xyz, int B, int L, int A, const int* slots_i, int n_i,
                     const int* slots_j, int n_j, int kind, float* out, cudaStream_t stream) {
    return pair_angles_variant_impl(xyz, B, L, A, slots_i, n_i, slots_j, n_j, kind, out, 0, stream);
}

// variant: 0 = default (the packed kernel whenever the structure fits in shared memory), 1 = the exact-sequence
// kernel of round 1, 3 = the packed kernel with two rows per iteration, 4 / 5 / 6 = other register budgets (tuning /
// comparison hooks, ps_trrosetta_angles_ex).
int trrosetta_angles_variant_impl(const float* xyz, int B, int L, int A, int use_virtual_cb, float* omega,
                                  float* theta, float* phi, int variant, cudaStream_t stream) {
    PS_REQUIRE(B > 0 && L > 0 && A > 0, PS_ERR_BAD_SHAPE,
               "trrosetta_angles: B=%d L=%d A=%d must be > 0", B, L, A);
    PS_REQUIRE(xyz, PS_ERR_NULL_POINTER, "trrosetta_angles: xyz is NULL");
    PS_REQUIRE(omega || theta || phi, PS_ERR_NULL_POINTER, "trrosetta_angles: no output requested");
    PS_REQUIRE(A >= (use_virtual_cb ? 3 : 5), PS_ERR_BAD_SHAPE,
               "trrosetta_angles: A=%d has no %s slot", A, use_virtual_cb ? "C" : "CB");
    const long long rows = static_cast<long long>(B) * L;
    const int sms = sm_count_for_current_device();
    if (sms < 0) return sms;
    // Packed kernel: the structure's CA / CB (6 floats per residue) plus the row records must fit in shared memory.
    const int Lp = (L + 1) & ~1;
    // one pair of residues j per thread (several beyond 512 residues)
    int threads = ((Lp / 2) + 31) / 32 * 32;
    if (threads > 256) threads = 256;
    // rows per CTA: as many as possible (the staging of the structure is amortised over them) while the grid still
    // holds >= 8 CTAs per SM (several waves: the CTAs of a launch differ a lot in cost when atoms are missing); at least 4,
    // at most 64.  (A wave-quantisation model — minimise ceil(CTAs / resident slots) x (rows + staging) — was measured
    // and lost: no gain at BASELINE config 3, and 57 -> 81 us at 64 x 384 ragged, where it picked a single wave.)
    int rows_per_cta = 64;
    while (rows_per_cta > 4 && rows / rows_per_cta < 8ll * sms) rows_per_cta /= 2;
    if (rows_per_cta > L) rows_per_cta = L;
    const size_t smem = (static_cast<size_t>(6) * Lp + static_cast<size_t>(rows_per_cta) * kRowRecord) * sizeof(float) +
                        static_cast<size_t>(Lp + 16);  // + one flag byte per residue
    const int blocks_per_structure = (L + rows_per_cta - 1) / rows_per_cta;
    const long long ctas = static_cast<long long>(B) * blocks_per_structure;
    // variant 0 (default): 3 CTAs / SM (80 registers, no spills); 4 / 5 / 6: compiled for 4 / 5 / 6 CTAs per SM;
    // 3: two rows per iteration at 2 CTAs / SM
    const int min_ctas = variant == 4 ? 4 : (variant == 5 ? 5 : (variant == 6 ? 6 : (variant == 3 ? 2 : 3)));
    if (variant != 1 && smem <= 200 * 1024 && ctas < (1ll << 31)) {
        auto aligned8 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; };
        const int vector_stores = (L % 2 == 0) && aligned8(omega) && aligned8(theta) && aligned8(phi);
        const bool all3 = omega && theta && phi;
#define PS_FAST(VCB, ALL, MINC)                                                                                       \
    do {                                                                                                              \
        cudaError_t err = cudaFuncSetAttribute(trrosetta_fast_kernel<VCB, ALL, MINC>,                                 \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);              \
        if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(trrosetta_fast_kernel)");                 \
        trrosetta_fast_kernel<VCB, ALL, MINC><<<static_cast<unsigned>(ctas), threads, smem, stream>>>(                \
            xyz, omega, theta, phi, L, A, rows_per_cta, blocks_per_structure, vector_stores);                         \
    } while (0)
        if (use_virtual_cb) {
            if (all3 && min_ctas == 4) PS_FAST(true, true, 4);
            else if (all3 && min_ctas == 2) PS_FAST(true, true, 2);
            else if (all3 && min_ctas == 5) PS_FAST(true, true, 5);
            else if (all3 && min_ctas == 6) PS_FAST(true, true, 6);
            else if (all3) PS_FAST(true, true, 3);
            else PS_FAST(true, false, 3);
        } else {
            if (all3 && min_ctas == 4) PS_FAST(false, true, 4);
            else if (all3 && min_ctas == 2) PS_FAST(false, true, 2);
            else if (all3 && min_ctas == 5) PS_FAST(false, true, 5);
            else if (all3 && min_ctas == 6) PS_FAST(false, true, 6);
            else if (all3) PS_FAST(false, true, 3);
            else PS_FAST(false, false, 3);
        }
#undef PS_FAST
        return check_launch("trrosetta_fast_kernel");
    }
    int grid = 0;
    int rc = grid_for_rows(rows, &grid);
    if (rc != PS_OK) return rc;
    const int row_threads = threads_for_L(L);
    if (use_virtual_cb)
        trrosetta_kernel<true><<<grid, row_threads, 0, stream>>>(xyz, omega, theta, phi, L, A, rows);
    else
        trrosetta_kernel<false><<<grid, row_threads, 0, stream>>>(xyz, omega, theta, phi, L, A, rows);
    return check_launch("trrosetta_kernel");
}

int trrosetta_angles_impl(const float* xyz, int B, int L, int A, int use_virtual_cb, float* omega,
                          float* theta, float* phi, cudaStream_t stream) {
    return trrosetta_angles_variant_impl(xyz, B, L, A, use_virtual_cb, omega, theta, phi, 0, stream);
}

}  // namespace ps
