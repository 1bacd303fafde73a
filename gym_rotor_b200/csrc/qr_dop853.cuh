// qr_dop853.cuh -- rigid-body right-hand side + scipy's DOP853 controller, one env per lane.
//
// What is reproduced (reference: coupled_yaw_wrapper.py:63 / decoupled_yaw_wrapper.py:76 / quad.py:265 call
// scipy.integrate.solve_ivp(method='DOP853') with default tolerances; SCIPY = scipy/integrate/_ivp):
//   RungeKutta.__init__ ............ SCIPY/rk.py:85-103      f0 = F(y0), select_initial_step
//   select_initial_step ............ SCIPY/common.py:68-134
//   OdeSolver.step / _step_impl .... SCIPY/base.py:179-210, SCIPY/rk.py:111-176   (accept / reject, factors)
//   rk_step ........................ SCIPY/rk.py:14-75
//   DOP853._estimate_error_norm .... SCIPY/rk.py:683-691
//   EoM / decouple_EoM ............. gym_rotor/envs/quad.py:321-335, wrappers/decoupled_yaw_wrapper.py:143-161
//
// B200-first restructuring (results agree with the reference to rounding, ~1e-15, not bit-for-bit):
//   * The ODE is a cascade: W' depends on W; R' on (R, W); v' on R; x' = v.  Nothing depends on x, so the
//     stage values of x are never formed: x_new and its two error estimates are accumulated on the fly from
//     the stage velocities.  W3' = M3/J3 is constant (J1 == J2), so W3 needs no stage storage either and
//     its error estimate is identically zero.  Stage storage is 14 values (v 3, R 9, W12 2) instead of 18.
//   * The kernel is bound by instruction issue and by the shared-memory data pipe, so one attempt is STRAIGHT-LINE code
//     with the tableau as immediates, organised to touch every stage derivative as rarely as possible:
//       - the 14 components are kept in an internal order `z` in which the right-hand side runs on aligned register
//         pairs (packed fma/mul .f32x2: FFMA2 / FMUL2), and all weighted sums are packed too;
//       - stages are processed in PAIRS (s, s+1): a stage derivative fetched from shared memory feeds the sums of both
//         stages, the derivative that was just computed is consumed from registers, and the last pair also feeds the
//         sums of y_new and of the fifth-order error estimate (float32).  13 + 1 vector loads and 7 vector stores per
//         attempt instead of 52 + 11; K1, K9, K10, K11 never leave the registers;
//       - the third-order error estimate is formed as (B sum) + three correction terms: scipy builds E3 from B by
//         changing entries 0, 8 and 11 (dop853_coefficients.py).
//   * One ATTEMPT is a function: the caller owns the accept/reject loop, so that lanes of a warp that need
//     a second attempt do not hold back lanes that are ready for their next env (see qr_kernels.cuh).
//   * F(y_new) is only evaluated when another step follows (t_new < T): scipy evaluates it always but only
//     uses it as the next step's first stage; the RHS has no side effects.  nfev is still reported as
//     scipy counts it (2 + 12 per attempt).
//   * x^(+-1/8) in the step-size controller are square-root chains.
#pragma once
#include "qr_math.cuh"
#include "dop853_tableau.h"

#ifndef QR_HD
#define QR_HD __host__ __device__
#endif
namespace qr {

// ---- tableau as compile-time constants (the stage code is fully unrolled: every call folds to an immediate) ------
QR_HD constexpr double dop_a(int s, int j)
{
    switch (s * 16 + j) {
#define QR_CA(s_, j_) case s_ * 16 + j_: return DOP_A##s_##_##j_;
        QR_CA(1, 0) QR_CA(2, 0) QR_CA(2, 1) QR_CA(3, 0) QR_CA(3, 2) QR_CA(4, 0) QR_CA(4, 2) QR_CA(4, 3)
        QR_CA(5, 0) QR_CA(5, 3) QR_CA(5, 4) QR_CA(6, 0) QR_CA(6, 3) QR_CA(6, 4) QR_CA(6, 5)
        QR_CA(7, 0) QR_CA(7, 3) QR_CA(7, 4) QR_CA(7, 5) QR_CA(7, 6)
        QR_CA(8, 0) QR_CA(8, 3) QR_CA(8, 4) QR_CA(8, 5) QR_CA(8, 6) QR_CA(8, 7)
        QR_CA(9, 0) QR_CA(9, 3) QR_CA(9, 4) QR_CA(9, 5) QR_CA(9, 6) QR_CA(9, 7) QR_CA(9, 8)
        QR_CA(10, 0) QR_CA(10, 3) QR_CA(10, 4) QR_CA(10, 5) QR_CA(10, 6) QR_CA(10, 7) QR_CA(10, 8) QR_CA(10, 9)
        QR_CA(11, 0) QR_CA(11, 3) QR_CA(11, 4) QR_CA(11, 5) QR_CA(11, 6) QR_CA(11, 7) QR_CA(11, 8) QR_CA(11, 9) QR_CA(11, 10)
#undef QR_CA
    }
    return 0.0;
}
QR_HD constexpr double dop_b(int j)
{
    switch (j) { case 0: return DOP_B0; case 5: return DOP_B5; case 6: return DOP_B6; case 7: return DOP_B7; case 8: return DOP_B8;
                 case 9: return DOP_B9; case 10: return DOP_B10; case 11: return DOP_B11; }
    return 0.0;
}
QR_HD constexpr double dop_e5(int j)
{
    switch (j) { case 0: return DOP_E5_0; case 5: return DOP_E5_5; case 6: return DOP_E5_6; case 7: return DOP_E5_7; case 8: return DOP_E5_8;
                 case 9: return DOP_E5_9; case 10: return DOP_E5_10; case 11: return DOP_E5_11; }
    return 0.0;
}
QR_HD constexpr double dop_e3(int j)
{
    switch (j) { case 0: return DOP_E3_0; case 5: return DOP_E3_5; case 6: return DOP_E3_6; case 7: return DOP_E3_7; case 8: return DOP_E3_8;
                 case 9: return DOP_E3_9; case 10: return DOP_E3_10; case 11: return DOP_E3_11; }
    return 0.0;
}
// E3 - B: non-zero for j = 0, 8, 11 only
QR_HD constexpr double dop_d3(int j) { return (j == 0 || j == 8 || j == 11) ? dop_e3(j) - dop_b(j) : 0.0; }
QR_HD constexpr double dop_c(int s)
{
    switch (s) { case 1: return DOP_C1; case 2: return DOP_C2; case 3: return DOP_C3; case 4: return DOP_C4; case 5: return DOP_C5; case 6: return DOP_C6;
                 case 7: return DOP_C7; case 8: return DOP_C8; case 9: return DOP_C9; case 10: return DOP_C10; case 11: return DOP_C11; }
    return 0.0;
}
static_assert(dop_e3(5) == dop_b(5) && dop_e3(6) == dop_b(6) && dop_e3(7) == dop_b(7) && dop_e3(9) == dop_b(9) && dop_e3(10) == dop_b(10),
              "E3 must equal B outside entries 0, 8, 11");

// ---- internal component order ------------------------------------------------------------------------------------
// External order of the 14 integrated components (state rows 3..16): v0 v1 v2 | R0..R8 (column-major: b1 b2 b3) | W1 W2.
// Internal order z, chosen so that the right-hand side works on aligned pairs:
//   z0 z1 = b1.xy   z2 z3 = b2.xy   z4 z5 = b3.xy   z6 z7 = b1.z b2.z   z8 z9 = b3.z v2   z10 z11 = v0 v1   z12 z13 = W1 W2
// ZOF[i] = position in z of external component i.  The permutation is free: everything is unrolled, so it only names registers.
QR_HD constexpr int zof(int i)
{
    switch (i) { case 0: return 10; case 1: return 11; case 2: return 9; case 3: return 0; case 4: return 1; case 5: return 6; case 6: return 2;
                 case 7: return 3; case 8: return 7; case 9: return 4; case 10: return 5; case 11: return 8; case 12: return 12; case 13: return 13; }
    return 0;
}
template <typename T> QR_DEV void to_z(const T* y, T* z)
{
#pragma unroll
    for (int i = 0; i < 14; ++i) z[zof(i)] = y[i];
}
template <typename T> QR_DEV void from_z(const T* z, T* y)
{
#pragma unroll
    for (int i = 0; i < 14; ++i) y[i] = z[zof(i)];
}
// R (column-major 3x3) out of / into a z vector
template <typename T> QR_DEV void z_get_R(const T* z, T* R)
{
    R[0] = z[0]; R[1] = z[1]; R[2] = z[6]; R[3] = z[2]; R[4] = z[3]; R[5] = z[7]; R[6] = z[4]; R[7] = z[5]; R[8] = z[8];
}
template <typename T> QR_DEV void z_set_R(const T* R, T* z)
{
    z[0] = R[0]; z[1] = R[1]; z[6] = R[2]; z[2] = R[3]; z[3] = R[4]; z[7] = R[5]; z[4] = R[6]; z[5] = R[7]; z[8] = R[8];
}

// ---- packed pair arithmetic (float32 on sm_100a: one FFMA2 / FMUL2 / FADD2 per pair; otherwise two scalar ops) ----
template <typename T> QR_DEV void pfma(T c, T k0, T k1, T a0, T a1, T& o0, T& o1)   // o = c * k + a
{
    o0 = num<T>::fma(c, k0, a0); o1 = num<T>::fma(c, k1, a1);
}
template <typename T> QR_DEV void pfma2(T c0, T c1, T k0, T k1, T a0, T a1, T& o0, T& o1)   // o = (c0, c1) * k + a
{
    o0 = num<T>::fma(c0, k0, a0); o1 = num<T>::fma(c1, k1, a1);
}
template <typename T> QR_DEV void pmul(T c, T k0, T k1, T& o0, T& o1) { o0 = c * k0; o1 = c * k1; }
template <typename T> QR_DEV void pmul2(T c0, T c1, T k0, T k1, T& o0, T& o1) { o0 = c0 * k0; o1 = c1 * k1; }
template <typename T> QR_DEV void padd(T a0, T a1, T b0, T b1, T& o0, T& o1) { o0 = a0 + b0; o1 = a1 + b1; }
#if QR_PTX
template <> QR_DEV void pfma<float>(float c, float k0, float k1, float a0, float a1, float& o0, float& o1)
{
    const float2 r = __ffma2_rn(make_float2(c, c), make_float2(k0, k1), make_float2(a0, a1)); o0 = r.x; o1 = r.y;
}
template <> QR_DEV void pfma2<float>(float c0, float c1, float k0, float k1, float a0, float a1, float& o0, float& o1)
{
    const float2 r = __ffma2_rn(make_float2(c0, c1), make_float2(k0, k1), make_float2(a0, a1)); o0 = r.x; o1 = r.y;
}
template <> QR_DEV void pmul<float>(float c, float k0, float k1, float& o0, float& o1)
{
    const float2 r = __fmul2_rn(make_float2(c, c), make_float2(k0, k1)); o0 = r.x; o1 = r.y;
}
template <> QR_DEV void pmul2<float>(float c0, float c1, float k0, float k1, float& o0, float& o1)
{
    const float2 r = __fmul2_rn(make_float2(c0, c1), make_float2(k0, k1)); o0 = r.x; o1 = r.y;
}
template <> QR_DEV void padd<float>(float a0, float a1, float b0, float b1, float& o0, float& o1)
{
    const float2 r = __fadd2_rn(make_float2(a0, a1), make_float2(b0, b1)); o0 = r.x; o1 = r.y;
}
#endif
// 14-vectors (7 pairs)
template <typename T> QR_DEV void v_mul(T c, const T* k, T* out)   // out = c * k
{
#pragma unroll
    for (int i = 0; i < 7; ++i) pmul<T>(c, k[2 * i], k[2 * i + 1], out[2 * i], out[2 * i + 1]);
}
template <typename T> QR_DEV void v_fma(T c, const T* k, T* acc)   // acc += c * k
{
#pragma unroll
    for (int i = 0; i < 7; ++i) pfma<T>(c, k[2 * i], k[2 * i + 1], acc[2 * i], acc[2 * i + 1], acc[2 * i], acc[2 * i + 1]);
}
template <typename T> QR_DEV void v_fma_out(T c, const T* k, const T* base, T* out)   // out = base + c * k
{
#pragma unroll
    for (int i = 0; i < 7; ++i) pfma<T>(c, k[2 * i], k[2 * i + 1], base[2 * i], base[2 * i + 1], out[2 * i], out[2 * i + 1]);
}

// ---- shared-memory stage storage ----------------------------------------------------------------------------
// Per warp: KS[slot][14 components][32 lanes] of T.  Within a slot the 14 components of one lane are
// packed as 3 x (4 consecutive T) + 1 x (2 consecutive T) so that float accesses are LDS/STS.128 + .64:
//   element (c, lane): c < 12 -> ((c >> 2) * 32 + lane) * 4 + (c & 3) ;  c >= 12 -> 384 + lane * 2 + (c - 12)
// Stored derivatives: K2..K8 (float32) / K2..K10 (float64); K5 reuses the slot of K2 (last read by the pair (4,5),
// whose second stage produces K5).  K1 and K11 (and K9, K10 in float32) never leave the registers.
constexpr int QR_NSLOTS = 8;
constexpr int QR_SLOT_ELEMS = 14 * 32;
QR_HD constexpr int k_slot(int j) { return j == 2 ? 0 : (j == 5 ? 0 : (j < 5 ? j - 2 : j - 3)); }   // K2 K3 K4 K5 K6 .. K10 -> 0 1 2 0 3 .. 7

// `kl` = this lane's column in slot 0 (ks + lane * 4).  float: groups g=0..2 at kl + g*128 (16 B each), tail at
// slot + 384 + lane*2 = kl + 384 - lane*2.  double: 7 groups of 2 at slot + (g*32 + lane)*2 = kl + g*64 - lane*2.
template <typename T> QR_DEV void ks_load(const T* kl, unsigned kl_sa, int lane, int slot, T* k)
{
    (void)kl_sa;
    const T* col = kl + slot * QR_SLOT_ELEMS;
    const T* p = col - lane * 2;
#pragma unroll
    for (int g = 0; g < 7; ++g) {
        double2 v = *reinterpret_cast<const double2*>(p + g * 64);
        k[2 * g] = (T)v.x; k[2 * g + 1] = (T)v.y;
    }
}
template <typename T> QR_DEV void ks_store(T* kl, unsigned kl_sa, int lane, int slot, const T* k)
{
    (void)kl_sa;
    T* p = kl + slot * QR_SLOT_ELEMS - lane * 2;
#pragma unroll
    for (int g = 0; g < 7; ++g) *reinterpret_cast<double2*>(p + g * 64) = make_double2((double)k[2 * g], (double)k[2 * g + 1]);
}
// float64 on the device: the same accesses as explicit (volatile) ld/st.shared -- as plain C++ loads the compiler hoists
// them far ahead of their use, and at 255 registers every hoisted stage derivative is spilled (the float64 kernel spent
// most of its time waiting for local memory: 278 local loads per 32 env-steps, long_scoreboard 3.3 per issued instruction)
#if QR_PTX
template <> QR_DEV void ks_load<double>(const double* kl, unsigned kl_sa, int lane, int slot, double* k)
{
    (void)kl;
    const unsigned sa = kl_sa + (unsigned)(slot * QR_SLOT_ELEMS * 8) - (unsigned)lane * 16u;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(k[0]), "=d"(k[1]) : "r"(sa) : "memory");
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+512];" : "=d"(k[2]), "=d"(k[3]) : "r"(sa) : "memory");
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+1024];" : "=d"(k[4]), "=d"(k[5]) : "r"(sa) : "memory");
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+1536];" : "=d"(k[6]), "=d"(k[7]) : "r"(sa) : "memory");
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+2048];" : "=d"(k[8]), "=d"(k[9]) : "r"(sa) : "memory");
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+2560];" : "=d"(k[10]), "=d"(k[11]) : "r"(sa) : "memory");
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+3072];" : "=d"(k[12]), "=d"(k[13]) : "r"(sa) : "memory");
}
template <> QR_DEV void ks_store<double>(double* kl, unsigned kl_sa, int lane, int slot, const double* k)
{
    (void)kl;
    const unsigned sa = kl_sa + (unsigned)(slot * QR_SLOT_ELEMS * 8) - (unsigned)lane * 16u;
    asm volatile("st.shared.v2.f64 [%2], {%0,%1};" :: "d"(k[0]), "d"(k[1]), "r"(sa) : "memory");
    asm volatile("st.shared.v2.f64 [%2+512], {%0,%1};" :: "d"(k[2]), "d"(k[3]), "r"(sa) : "memory");
    asm volatile("st.shared.v2.f64 [%2+1024], {%0,%1};" :: "d"(k[4]), "d"(k[5]), "r"(sa) : "memory");
    asm volatile("st.shared.v2.f64 [%2+1536], {%0,%1};" :: "d"(k[6]), "d"(k[7]), "r"(sa) : "memory");
    asm volatile("st.shared.v2.f64 [%2+2048], {%0,%1};" :: "d"(k[8]), "d"(k[9]), "r"(sa) : "memory");
    asm volatile("st.shared.v2.f64 [%2+2560], {%0,%1};" :: "d"(k[10]), "d"(k[11]), "r"(sa) : "memory");
    asm volatile("st.shared.v2.f64 [%2+3072], {%0,%1};" :: "d"(k[12]), "d"(k[13]), "r"(sa) : "memory");
}
#endif
// float32: explicit ld/st.shared on the 32-bit shared-window address with immediate slot offsets (no 64-bit pointer math)
template <> QR_DEV void ks_load<float>(const float* kl, unsigned kl_sa, int lane, int slot, float* k)
{
#if QR_PTX
    (void)kl;
    const unsigned sa = kl_sa + (unsigned)(slot * QR_SLOT_ELEMS * 4);
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(k[0]), "=f"(k[1]), "=f"(k[2]), "=f"(k[3]) : "r"(sa) : "memory");
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+512];" : "=f"(k[4]), "=f"(k[5]), "=f"(k[6]), "=f"(k[7]) : "r"(sa) : "memory");
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+1024];" : "=f"(k[8]), "=f"(k[9]), "=f"(k[10]), "=f"(k[11]) : "r"(sa) : "memory");
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+1536];" : "=f"(k[12]), "=f"(k[13]) : "r"(sa - (unsigned)lane * 8u) : "memory");
#else
    (void)kl_sa;
    const float* col = kl + slot * QR_SLOT_ELEMS;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(col + g * 128);
        k[4 * g] = v.x; k[4 * g + 1] = v.y; k[4 * g + 2] = v.z; k[4 * g + 3] = v.w;
    }
    const float2 w = *reinterpret_cast<const float2*>(col + 384 - lane * 2);
    k[12] = w.x; k[13] = w.y;
#endif
}
template <> QR_DEV void ks_store<float>(float* kl, unsigned kl_sa, int lane, int slot, const float* k)
{
#if QR_PTX
    (void)kl;
    const unsigned sa = kl_sa + (unsigned)(slot * QR_SLOT_ELEMS * 4);
    asm volatile("st.shared.v4.f32 [%4], {%0,%1,%2,%3};" :: "f"(k[0]), "f"(k[1]), "f"(k[2]), "f"(k[3]), "r"(sa) : "memory");
    asm volatile("st.shared.v4.f32 [%4+512], {%0,%1,%2,%3};" :: "f"(k[4]), "f"(k[5]), "f"(k[6]), "f"(k[7]), "r"(sa) : "memory");
    asm volatile("st.shared.v4.f32 [%4+1024], {%0,%1,%2,%3};" :: "f"(k[8]), "f"(k[9]), "f"(k[10]), "f"(k[11]), "r"(sa) : "memory");
    asm volatile("st.shared.v2.f32 [%2+1536], {%0,%1};" :: "f"(k[12]), "f"(k[13]), "r"(sa - (unsigned)lane * 8u) : "memory");
#else
    (void)kl_sa;
    float* col = kl + slot * QR_SLOT_ELEMS;
#pragma unroll
    for (int g = 0; g < 3; ++g) *reinterpret_cast<float4*>(col + g * 128) = make_float4(k[4 * g], k[4 * g + 1], k[4 * g + 2], k[4 * g + 3]);
    *reinterpret_cast<float2*>(col + 384 - lane * 2) = make_float2(k[12], k[13]);
#endif
}

// Layout of the 14 integrated components kept in registers by the caller (external order): y[0..2] = v, y[3..11] = R
// (column-major), y[12..13] = W1, W2.  x[3] and W3 are carried separately.
template <typename T> struct Dyn {
    T fm;      // f / m          (thrust acceleration magnitude)
    T g;       // gravity
    T Mi0, Mi1;// M1/J1, M2/J1
    T kw0, kw1;// (J1 - J3)/J1 for W1' ; (J3 - J1)/J1 for W2'
    T w3dot;   // M3 / J3 : constant because J1 == J2 (quad.py:378)
};

// One right-hand-side evaluation at a stage point given in EXTERNAL order (ys, W3s) -> k[14] (external order).  Used
// outside the integrator's hot path (explicit Euler of the base env).  Returns ensure_SO3 flags.
template <typename T, bool NEWTON = false, bool CHECK = true> QR_DEV int rhs14(const T* ys, T W3s, const Dyn<T>& d, T* k)
{
    using N = num<T>;
    T R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = ys[3 + i];
    int fl = 0;
    if (CHECK) fl = ensure_so3<T, NEWTON>(R);  // state_decomposition -> ensure_SO3 on every call (quad_utils.py:12-16)
    const T W0 = ys[12], W1 = ys[13], W2 = W3s;
    // v' = g e3 - (f/m) R e3
    k[0] = -d.fm * R[6];
    k[1] = -d.fm * R[7];
    k[2] = N::fma(-d.fm, R[8], d.g);
    // R' = R hat(W)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        T r0 = R[i], r1 = R[i + 3], r2 = R[i + 6];
        k[3 + i] = N::fma(r1, W2, -(r2 * W1));
        k[6 + i] = N::fma(r2, W0, -(r0 * W2));
        k[9 + i] = N::fma(r0, W1, -(r1 * W0));
    }
    // W' = J^-1 (-W x JW + M), J = diag(J1, J1, J3)
    k[12] = N::fma(d.kw0 * W1, W2, d.Mi0);
    k[13] = N::fma(d.kw1 * W0, W2, d.Mi1);
    return fl;
}

// The same right-hand side on the internal order: z -> k (both in z order), no SO(3) test (the callers run it).
// Same operations and roundings as rhs14; the six column-pair products are packed.
template <typename T> QR_DEV void rhs_z(const T* z, T W3, const Dyn<T>& d, T* k)
{
    using N = num<T>;
    const T W1 = z[12], W2 = z[13];
    T t0, t1;
    // b1' = b2 W3 - b3 W2 ; b2' = b3 W1 - b1 W3 ; b3' = b1 W2 - b2 W1      (R' = R hat(W), column by column)
    pmul<T>(-W2, z[4], z[5], t0, t1); pfma<T>(W3, z[2], z[3], t0, t1, k[0], k[1]);
    pmul<T>(-W3, z[0], z[1], t0, t1); pfma<T>(W1, z[4], z[5], t0, t1, k[2], k[3]);
    pmul<T>(-W1, z[2], z[3], t0, t1); pfma<T>(W2, z[0], z[1], t0, t1, k[4], k[5]);
    k[6] = N::fma(z[7], W3, -(z[8] * W2));
    k[7] = N::fma(z[8], W1, -(z[6] * W3));
    k[8] = N::fma(z[6], W2, -(z[7] * W1));
    // v' = g e3 - (f/m) b3
    k[9] = N::fma(-d.fm, z[8], d.g);
    pmul<T>(-d.fm, z[4], z[5], k[10], k[11]);
    // W' = J^-1 (-W x JW + M), J = diag(J1, J1, J3)
    k[12] = N::fma(d.kw0 * W2, W3, d.Mi0);
    k[13] = N::fma(d.kw1 * W1, W3, d.Mi1);
}

// out[0..8] = (M hat(W)) for the 3x3 matrix M held in positions 0..8 of a z vector (the first nine lines of rhs_z).
template <typename T> QR_DEV void rhat_z(const T* m, T W1, T W2, T W3, T* out)
{
    using N = num<T>;
    T t0, t1;
    pmul<T>(-W2, m[4], m[5], t0, t1); pfma<T>(W3, m[2], m[3], t0, t1, out[0], out[1]);
    pmul<T>(-W3, m[0], m[1], t0, t1); pfma<T>(W1, m[4], m[5], t0, t1, out[2], out[3]);
    pmul<T>(-W1, m[2], m[3], t0, t1); pfma<T>(W2, m[0], m[1], t0, t1, out[4], out[5]);
    out[6] = N::fma(m[7], W3, -(m[8] * W2));
    out[7] = N::fma(m[8], W1, -(m[6] * W3));
    out[8] = N::fma(m[6], W2, -(m[7] * W1));
}

// The acceptance test of ensure_SO3 (see so3_ok in qr_math.cuh) on the R part of a z vector.  Same tolerances; the
// column products are packed, every comparison is ordered (a NaN fails), one predicate at the end.
template <typename T> QR_DEV bool so3_ok_z(const T* z, T* mx = nullptr)
{
    using N = num<T>;
    const T tol = (T)1e-5;
    T p0, p1, q0, q1, r0, r1, s0, s1, u0, u1, w0, w1;
    pfma2<T>(z[0], z[1], z[0], z[1], (T)-1, (T)0, p0, p1);   // b1.xy^2 (- 1)
    pfma2<T>(z[2], z[3], z[2], z[3], (T)-1, (T)0, q0, q1);   // b2.xy^2 (- 1)
    pfma2<T>(z[4], z[5], z[4], z[5], (T)-1, (T)0, r0, r1);   // b3.xy^2 (- 1)
    pmul2<T>(z[0], z[1], z[2], z[3], s0, s1);                // b1.xy * b2.xy
    pmul2<T>(z[0], z[1], z[4], z[5], u0, u1);                // b1.xy * b3.xy
    pmul2<T>(z[2], z[3], z[4], z[5], w0, w1);                // b2.xy * b3.xy
    const T e00 = N::fma(z[6], z[6], p0) + p1;
    const T e11 = N::fma(z[7], z[7], q0) + q1;
    const T e22 = N::fma(z[8], z[8], r0) + r1;
    const T e01 = N::fma(z[6], z[7], s0) + s1;
    const T e02 = N::fma(z[6], z[8], u0) + u1;
    const T e12 = N::fma(z[7], z[8], w0) + w1;
    // det R - 1.  float64: b1 . (b2 x b3) - 1.  float32: det R = sqrt(det(I + E)) = 1 + tr(E) / 2 + O(|E|^2) with E = R^T R - I
    // just computed; where the test can pass at all |E| <= 2e-5, so the neglected terms are below 1e-9 -- a hundred times
    // less than the rounding error of the nine-product determinant in float32 (~2e-7), against a threshold of 1e-5
    T dm1;
    if (sizeof(T) == 4) {
        dm1 = (T)0.5 * ((e00 + e11) + e22);
    } else {
        const T cx = N::fma(z[3], z[8], -(z[7] * z[5]));
        const T cy = N::fma(z[7], z[4], -(z[2] * z[8]));
        const T cz = N::fma(z[2], z[5], -(z[3] * z[4]));
        dm1 = N::fma(z[0], cx, N::fma(z[1], cy, N::fma(z[6], cz, (T)-1)));
    }
    if (mx && sizeof(T) == 8) {
        // float64 (no NaN-propagating 3-input maximum there): the seven comparisons per stage, folded into the first maximum
        const bool ok = (N::abs(e00) <= tol + tol) & (N::abs(e11) <= tol + tol) & (N::abs(e22) <= tol + tol) & (N::abs(e01) <= tol) &
                        (N::abs(e02) <= tol) & (N::abs(e12) <= tol) & (N::abs(dm1) <= (T)1e-8 + tol);
        mx[0] = ok ? mx[0] : N::inf();
        return true;
    }
    if (mx) {
        // speculative stages (stage_finish): the seven defects only feed three running maxima (diagonal, off-diagonal,
        // determinant; NaN-propagating), compared once at the end of the attempt instead of seven times per stage
        mx[0] = N::absmax2_nan(N::absmax3_nan(mx[0], e00, e11), e22);
        mx[1] = N::absmax2_nan(N::absmax3_nan(mx[1], e01, e02), e12);
        mx[2] = N::absmax2_nan(mx[2], dm1);
        return true;
    }
    return (N::abs(e00) <= tol + tol) & (N::abs(e11) <= tol + tol) & (N::abs(e22) <= tol + tol) & (N::abs(e01) <= tol) &
           (N::abs(e02) <= tol) & (N::abs(e12) <= tol) & (N::abs(dm1) <= (T)1e-8 + tol);
}

// Integrator state of one lane between attempts.
template <typename T> struct OdeLane {
    T t, h_abs;
    int rejected;   // a rejection happened inside the current scipy step
    int nfev;       // as scipy counts: 2 + 12 * attempts
    int status;     // QR_ST_* bits
    int nproj;      // SO(3) re-projections inside RHS evaluations
    int checked;    // the speculative attempt met a stage matrix that fails the SO(3) test: redo it with dop853_attempt_checked
};

// RungeKutta.__init__: f0 = F(y0) -> K0 (z order), then select_initial_step.  (2 RHS evaluations)
template <typename T>
QR_DEV void dop853_begin(const T* x, const T* y, T W3, const Dyn<T>& d, const T Tend, const T rtol, const T atol, T* K0, OdeLane<T>& o)
{
    using N = num<T>;
    o.t = 0; o.rejected = 0; o.nfev = 2; o.status = 0; o.nproj = 0; o.checked = 0;
    T z[14];
    to_z<T>(y, z);
    rhs_z<T>(z, W3, d, K0);   // y's R was SO(3)-checked by the caller (observation_wrapper)
    // sums of (y/sc)^2 and (f0/sc)^2 over all 18 components; x' = v
    T s0a = 0, s0b = 0, s1a = 0, s1b = 0;
    T isc[14], iscx[3];
#pragma unroll
    for (int i = 0; i < 14; ++i) isc[i] = N::recip(N::fma(N::abs(z[i]), rtol, atol));
#pragma unroll
    for (int i = 0; i < 3; ++i) iscx[i] = N::recip(N::fma(N::abs(x[i]), rtol, atol));
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        T a0, a1, b0, b1;
        pmul2<T>(isc[2 * i], isc[2 * i + 1], z[2 * i], z[2 * i + 1], a0, a1);
        pmul2<T>(isc[2 * i], isc[2 * i + 1], K0[2 * i], K0[2 * i + 1], b0, b1);
        pfma2<T>(a0, a1, a0, a1, s0a, s0b, s0a, s0b);
        pfma2<T>(b0, b1, b0, b1, s1a, s1b, s1a, s1b);
    }
    {
        const T v[3] = {z[10], z[11], z[9]};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            T a = x[i] * iscx[i], b = v[i] * iscx[i];
            s0a = N::fma(a, a, s0a); s1a = N::fma(b, b, s1a);
        }
        T iscw3 = N::recip(N::fma(N::abs(W3), rtol, atol));
        T a = W3 * iscw3, b = d.w3dot * iscw3;
        s0b = N::fma(a, a, s0b); s1b = N::fma(b, b, s1b);
    }
    const T inv_sqrt_n = (T)0.23570226039551584;  // 1/sqrt(18)
    T d0 = N::sqrt(s0a + s0b) * inv_sqrt_n, d1 = N::sqrt(s1a + s1b) * inv_sqrt_n;
    T h0 = (d0 < (T)1e-5 || d1 < (T)1e-5) ? (T)1e-6 : N::div_fast((T)0.01 * d0, d1);
    h0 = (Tend < h0) ? Tend : h0;   // python min(h0, interval): keeps a NaN h0
    // Euler probe y1 = y0 + h0 f0 ; f1 = F(y1)
    T z1[14], k1[14];
    v_fma_out<T>(h0, K0, z, z1);
    T W31 = N::fma(h0, d.w3dot, W3);
    if (sizeof(T) == 8) {
        T R[9];
        z_get_R<T>(z1, R);
        const int fl = ensure_so3<T, true>(R);   // state_decomposition inside EoM; the probe leaves SO(3) in ~12 % of the steps
        o.nproj += fl & 1; if (fl & 2) o.status |= 4;
        T zr[14];
#pragma unroll
        for (int i = 0; i < 14; ++i) zr[i] = z1[i];
        z_set_R<T>(R, zr);
        rhs_z<T>(zr, W31, d, k1);
    } else {
        // float32 mode: the probe's attitude is R (I + h0 hat(W)) with R on SO(3) (checked by the caller), so ensure_SO3's
        // verdict and its result are known in closed form -- no test of nine products, no iteration, no divergence:
        //   R1^T R1 - I = h0^2 (|W|^2 I - W W^T), det R1 = 1 + a2, a2 = h0^2 |W|^2: the determinant test (a2 <= 1e-5 + 1e-8)
        //   is the binding one of the seven;
        //   the projection U V^T (psvd) of R1 is R Rot(W / |W|, atan(h0 |W|)) = R + s R hat(W) + c R hat(W)^2 with
        //   s = h0 / sqrt(1 + a2), c = (1 - cos) / |W|^2 = h0^2 rs^2 / (1 + rs), rs = 1 / sqrt(1 + a2)   (Rodrigues)
        // and R hat(W) is the attitude part of f0.  Lanes whose probe passes use s = h0, c = 0: the Euler probe itself.
        const T W1 = z[12], W2 = z[13];
        const T n2 = N::fma(W3, W3, N::fma(W2, W2, W1 * W1));
        const T hh = h0 * h0, a2 = hh * n2;
        const bool pass = a2 <= (T)1.001e-5;
        const T rs = N::rsqrt((T)1 + a2);
        const T s = pass ? h0 : h0 * rs;
        const T cc = pass ? (T)0 : hh * rs * rs * N::recip((T)1 + rs);
        o.nproj += pass ? 0 : 1;
        T m2[9], zr[14];
        rhat_z<T>(K0, W1, W2, W3, m2);
#pragma unroll
        for (int i = 0; i < 9; ++i) zr[i] = N::fma(cc, m2[i], N::fma(s, K0[i], z[i]));
#pragma unroll
        for (int i = 9; i < 14; ++i) zr[i] = z1[i];
        rhs_z<T>(zr, W31, d, k1);
    }
    T s2a = 0, s2b = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        T a0, a1;
        padd<T>(k1[2 * i], k1[2 * i + 1], -K0[2 * i], -K0[2 * i + 1], a0, a1);
        pmul2<T>(isc[2 * i], isc[2 * i + 1], a0, a1, a0, a1);
        pfma2<T>(a0, a1, a0, a1, s2a, s2b, s2a, s2b);
    }
    {   // f1_x - f0_x = v1 - v0 ; the W3 component of f1 - f0 is exactly zero
        const T dv[3] = {z1[10] - z[10], z1[11] - z[11], z1[9] - z[9]};
#pragma unroll
        for (int i = 0; i < 3; ++i) { T a = dv[i] * iscx[i]; s2a = N::fma(a, a, s2a); }
    }
    T d2 = N::div_fast(N::sqrt(s2a + s2b) * inv_sqrt_n, h0);
    T h1;
    if (d1 <= (T)1e-15 && d2 <= (T)1e-15) h1 = N::max((T)1e-6, h0 * (T)1e-3);
    else h1 = N::root8(N::div_fast((T)0.01, N::max(d1, d2)));
    T h_abs = (T)100 * h0;
    h_abs = (h1 < h_abs) ? h1 : h_abs;
    h_abs = (Tend < h_abs) ? Tend : h_abs;
    // first _step_impl: h_abs is raised to min_step = 10 ulp(t) if smaller
    const T min_step = (T)10 * N::ulp_up((T)0);
    o.h_abs = (h_abs < min_step) ? min_step : h_abs;
}

// ---- the pieces of one attempt ---------------------------------------------------------------------------------------
// Running state of an attempt that the stage code shares (all registers; the struct only keeps the argument lists short).
template <typename T> struct AttCtx {
    T h;              // signed step
    T W3;             // W3 at the start of the step
    T xb[3], x5[3], x3[3];   // B / E5 / (E3 - B) weighted sums of the stage velocities (x' = v): order v0 v1 v2
    T so3mx[3];       // running maxima over the stages of the SO(3) defects |diag(RtR) - 1|, |offdiag(RtR)|, |det R - 1|
    QR_DEV bool all_ok() const   // every stage matrix passed the SO(3) test (a NaN fails it)
    {
        const T tol = (T)1e-5;
        return (so3mx[0] <= tol + tol) & (so3mx[1] <= tol) & (so3mx[2] <= (T)1e-8 + tol);
    }
};

// Stage S: stage point zs = z + h P (P = sum_j a_Sj K_j, complete), SO(3) test, F = K_S = f(zs).  P is consumed.
template <int S, typename T>
QR_DEV void stage_finish(const T* z, T* P, const Dyn<T>& d, AttCtx<T>& c, T* F)
{
    using N = num<T>;
    v_fma_out<T>(c.h, P, z, P);   // P now holds the stage point
    const T W3s = N::fma(c.h * (T)dop_c(S), d.w3dot, c.W3);
    if (dop_b(S) != 0.0) {   // stage velocity into the running sums of x (zero weights for S < 5)
        pfma<T>((T)dop_b(S), P[10], P[11], c.xb[0], c.xb[1], c.xb[0], c.xb[1]); c.xb[2] = N::fma((T)dop_b(S), P[9], c.xb[2]);
        pfma<T>((T)dop_e5(S), P[10], P[11], c.x5[0], c.x5[1], c.x5[0], c.x5[1]); c.x5[2] = N::fma((T)dop_e5(S), P[9], c.x5[2]);
    }
    if (dop_d3(S) != 0.0) {
        pfma<T>((T)dop_d3(S), P[10], P[11], c.x3[0], c.x3[1], c.x3[0], c.x3[1]); c.x3[2] = N::fma((T)dop_d3(S), P[9], c.x3[2]);
    }
    // The reference tests every stage matrix against SO(3) before using it (state_decomposition inside EoM) and
    // re-projects it if the test fails -- which, for stage points of an accepted-size step, essentially never
    // happens.  The stages therefore run SPECULATIVELY: the test is evaluated as plain dataflow (nothing waits
    // for it), and if some stage failed it the attempt is thrown away and redone by dop853_attempt_checked, which
    // re-projects stage by stage exactly like the reference does.
    so3_ok_z<T>(P, c.so3mx);
    rhs_z<T>(P, W3s, d, F);
}

// Stages S and S+1.  On entry F = K_{S-1} (registers); on exit F = K_{S+1}.  FINAL: the pair (10, 11) also accumulates
// the sums of y_new (sb) and of the fifth-order error estimate (s5) from the derivatives it has in registers anyway.
template <int S, bool FINAL, typename T>
QR_DEV void stage_pair(const T* z, const T* K0, const Dyn<T>& d, AttCtx<T>& c, T* F, T* kl, unsigned kl_sa, int lane, T* sb, T* s5)
{
    constexpr int NSTORE = (sizeof(T) == 4) ? 8 : 10;   // derivatives K2 .. K_NSTORE go to shared memory
    T P[14], Q[14];
    v_mul<T>((T)dop_a(S, 0), K0, P);
    v_mul<T>((T)dop_a(S + 1, 0), K0, Q);
    if (dop_a(S, S - 1) != 0.0) v_fma<T>((T)dop_a(S, S - 1), F, P);
    if (dop_a(S + 1, S - 1) != 0.0) v_fma<T>((T)dop_a(S + 1, S - 1), F, Q);
    if (FINAL) {
        v_mul<T>((T)dop_b(0), K0, sb); v_mul<T>((T)dop_e5(0), K0, s5);
        v_fma<T>((T)dop_b(S - 1), F, sb); v_fma<T>((T)dop_e5(S - 1), F, s5);
    }
#pragma unroll
    for (int j = 1; j <= S - 2; ++j) {
        if (dop_a(S, j) != 0.0 || dop_a(S + 1, j) != 0.0) {
            T k[14];
            ks_load<T>(kl, kl_sa, lane, k_slot(j), k);
            if (dop_a(S, j) != 0.0) v_fma<T>((T)dop_a(S, j), k, P);
            if (dop_a(S + 1, j) != 0.0) v_fma<T>((T)dop_a(S + 1, j), k, Q);
            if (FINAL && dop_b(j) != 0.0) { v_fma<T>((T)dop_b(j), k, sb); v_fma<T>((T)dop_e5(j), k, s5); }
        }
    }
    stage_finish<S, T>(z, P, d, c, F);
    if (S >= 2 && S <= NSTORE) ks_store<T>(kl, kl_sa, lane, k_slot(S), F);
    v_fma<T>((T)dop_a(S + 1, S), F, Q);
    if (FINAL) { v_fma<T>((T)dop_b(S), F, sb); v_fma<T>((T)dop_e5(S), F, s5); }
    stage_finish<S + 1, T>(z, Q, d, c, F);
    if (S + 1 <= NSTORE) ks_store<T>(kl, kl_sa, lane, k_slot(S + 1), F);
    if (FINAL) { v_fma<T>((T)dop_b(S + 1), F, sb); v_fma<T>((T)dop_e5(S + 1), F, s5); }
}

// Error norm, accept / reject and the step-size update (rk.py:125-176, 683-691) from the finished sums, for one lane.
// znew: y_new (z order); s5 / s3: the two error estimators' weighted sums; xnew etc. for the position.
// Returns true when the lane is finished with the whole interval.  K0 receives F(y_new) when another step follows.
template <typename T>
QR_DEV bool dop853_conclude(T* x, T* y, T& W3, const Dyn<T>& d, const T Tend, const T rtol, const T atol, T* K0, OdeLane<T>& o,
                            const T h, const T t_new, const bool too_small, const T* z, const T* znew, const T* s5, const T* s3,
                            const T* xnew, const T* x5, const T* x3, const int nproj, const int bad)
{
    using N = num<T>;
    T e5a = 0, e5b = 0, e3a = 0, e3b = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const T isc = N::recip(N::fma(N::max(N::abs(x[i]), N::abs(xnew[i])), rtol, atol));
        const T e5 = x5[i] * isc, e3 = x3[i] * isc;
        e5a = N::fma(e5, e5, e5a); e3a = N::fma(e3, e3, e3a);
    }
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        T sc0, sc1, a0, a1, b0, b1;
        pfma<T>(rtol, N::max(N::abs(z[2 * i]), N::abs(znew[2 * i])), N::max(N::abs(z[2 * i + 1]), N::abs(znew[2 * i + 1])), atol, atol, sc0, sc1);
        const T i0 = N::recip(sc0), i1 = N::recip(sc1);
        pmul2<T>(i0, i1, s5[2 * i], s5[2 * i + 1], a0, a1);
        pmul2<T>(i0, i1, s3[2 * i], s3[2 * i + 1], b0, b1);
        pfma2<T>(a0, a1, a0, a1, e5a, e5b, e5a, e5b);
        pfma2<T>(b0, b1, b0, b1, e3a, e3b, e3a, e3b);
    }
    const T e5n = e5a + e5b, e3n = e3a + e3b;
    T err;
    if (e5n == (T)0 && e3n == (T)0) err = 0;
    else err = N::abs(h) * e5n * N::rsqrt((e5n + (T)0.01 * e3n) * (T)18);

    if (too_small) { o.status |= 2; return true; }   // TOO_SMALL_STEP: keep the last accepted y
    o.h_abs = N::abs(h);
    o.nfev += 12;
    o.nproj += nproj; if (bad) o.status |= 4;
    if (err < (T)1) {   // accept
#pragma unroll
        for (int i = 0; i < 3; ++i) x[i] = xnew[i];
        from_z<T>(znew, y);
        W3 = N::fma(h, d.w3dot, W3);
        o.t = t_new;
        if (!(o.t < Tend)) return true;
        T factor = (err == (T)0) ? (T)10 : N::min((T)10, (T)0.9 * N::inv_root8(err));
        if (o.rejected) factor = N::min((T)1, factor);
        o.h_abs *= factor;
        {   // f_new becomes the next step's first stage (evaluated on the SO(3)-checked copy, like every EoM call)
            T R[9], zr[14];
#pragma unroll
            for (int i = 0; i < 14; ++i) zr[i] = znew[i];
            z_get_R<T>(znew, R);
            const int fl = ensure_so3<T>(R);
            z_set_R<T>(R, zr);
            rhs_z<T>(zr, W3, d, K0);
            o.nproj += fl & 1; if (fl & 2) o.status |= 4;
        }
        // next _step_impl call: fresh rejection flag, h_abs raised to min_step(t) if smaller
        o.rejected = 0;
        const T ms = (T)10 * N::ulp_up(o.t);
        if (o.h_abs < ms) o.h_abs = ms;
        return false;
    }
    // A NaN error norm also lands here (nan < 1 is False).  scipy then shrinks h by 0.2 until TOO_SMALL_STEP
    // and solve_ivp returns the last accepted y; that outcome is produced at once.
    if (err != err) { o.status |= 1 | 2; return true; }
    o.h_abs *= N::max((T)0.2, (T)0.9 * N::inv_root8(err));
    o.rejected = 1;
    return false;
}

// One attempt of _step_impl (rk.py:125-166) for the lane: 11 stages from K0, y_new, error norm,
// accept / reject and step-size update.  Returns true when the lane is finished with the whole interval
// (t reached Tend, or the integrator gave up and keeps the last accepted state).
// ks: this warp's stage storage, lane: lane id.  y in external order, K0 in z order.
//
// The function is executed by ALL lanes of the warp (straight-line code); `live` says whether this lane really has an
// attempt to make.  Lanes without one run on whatever benign state they hold and commit nothing.  If a stage matrix
// fails the SO(3) test the lane commits nothing either and sets o.checked: the caller then runs dop853_attempt_checked.
template <typename T>
QR_DEV bool dop853_attempt(T* x, T* y, T& W3, const Dyn<T>& d, const T Tend, const T rtol, const T atol, T* K0,
                           OdeLane<T>& o, T* ks, const int lane, const bool live)
{
    using N = num<T>;
    constexpr bool MERGE = sizeof(T) == 4;   // float32: y_new / E5 sums ride on the loads of the last stage pair
    const T min_step = (T)10 * N::ulp_up(o.t);
    const bool too_small = o.h_abs < min_step;
    T t_new = o.t + o.h_abs;
    if (t_new - Tend > (T)0) t_new = Tend;
    AttCtx<T> c;
    c.h = t_new - o.t; c.W3 = W3; c.so3mx[0] = c.so3mx[1] = c.so3mx[2] = 0;
    T z[14];
    to_z<T>(y, z);
    {   // stage 0 of the position sums: K0_x = v
        const T v[3] = {z[10], z[11], z[9]};
#pragma unroll
        for (int i = 0; i < 3; ++i) { c.xb[i] = (T)dop_b(0) * v[i]; c.x5[i] = (T)dop_e5(0) * v[i]; c.x3[i] = (T)dop_d3(0) * v[i]; }
    }
    T* const kl = ks + lane * 4;   // this lane's column inside slot 0 (see ks_load / ks_store)
    unsigned kl_sa = (unsigned)__cvta_generic_to_shared(kl);

    T F[14], sb[14], s5[14];
    {   // stage 1
        T P[14];
        v_mul<T>((T)dop_a(1, 0), K0, P);
        stage_finish<1, T>(z, P, d, c, F);
    }
    // (float64 un-paired stages measured slower, 0.76 vs 0.91 G: profiles/r02/r02s_ab.txt)
    stage_pair<2, false, T>(z, K0, d, c, F, kl, kl_sa, lane, sb, s5);
    stage_pair<4, false, T>(z, K0, d, c, F, kl, kl_sa, lane, sb, s5);
    stage_pair<6, false, T>(z, K0, d, c, F, kl, kl_sa, lane, sb, s5);
    stage_pair<8, false, T>(z, K0, d, c, F, kl, kl_sa, lane, sb, s5);
    stage_pair<10, MERGE, T>(z, K0, d, c, F, kl, kl_sa, lane, sb, s5);
    // y_new = y + h * sum_s B_s K_s ; error estimates (rk.py:683-691).  F = K11.
    T s3[14];
    v_mul<T>((T)dop_d3(0), K0, s3);
    v_fma<T>((T)dop_d3(11), F, s3);
    if (!MERGE) {
        v_mul<T>((T)dop_b(0), K0, sb); v_mul<T>((T)dop_e5(0), K0, s5);
        v_fma<T>((T)dop_b(11), F, sb); v_fma<T>((T)dop_e5(11), F, s5);
#pragma unroll
        for (int j = 5; j <= 10; ++j) {
            T k[14];
            ks_load<T>(kl, kl_sa, lane, k_slot(j), k);
            v_fma<T>((T)dop_b(j), k, sb); v_fma<T>((T)dop_e5(j), k, s5);
            if (dop_d3(j) != 0.0) v_fma<T>((T)dop_d3(j), k, s3);
        }
    } else {
        T k[14];
        ks_load<T>(kl, kl_sa, lane, k_slot(8), k);
        v_fma<T>((T)dop_d3(8), k, s3);
    }
    T xnew[3], x3[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { xnew[i] = N::fma(c.h, c.xb[i], x[i]); x3[i] = c.x3[i] + c.xb[i]; }
#pragma unroll
    for (int i = 0; i < 7; ++i) padd<T>(s3[2 * i], s3[2 * i + 1], sb[2 * i], sb[2 * i + 1], s3[2 * i], s3[2 * i + 1]);   // E3 sum = B sum + correction
    v_fma_out<T>(c.h, sb, z, sb);   // y_new
    if (!live) return false;
    if (!c.all_ok()) { o.checked = 1; return false; }   // redo this attempt with per-stage re-projection
    return dop853_conclude<T>(x, y, W3, d, Tend, rtol, atol, K0, o, c.h, t_new, too_small, z, sb, s5, s3, xnew, c.x5, x3, 0, 0);
}

// The redo of an attempt in which some stage matrix failed the SO(3) test (never within the termination limits;
// thousands of times at |W| > 20 rad/s): the reference's per-stage ensure_SO3 with re-projection, stage by stage.
// Out of line, rolled and scalar on purpose -- it is a rare path and must not cost the hot loop registers or
// instruction-cache space; its stage derivatives live in local memory.  Same arguments as dop853_attempt, passed
// through memory (the caller copies its registers in and out).
template <typename T>
__device__ __noinline__ bool dop853_attempt_checked(T* x, T* zio, T* W3p, const Dyn<T>* dp, const T Tend, const T rtol, const T atol, T* K0, OdeLane<T>* op)
{
    // zio: the 14 integrated components in the INTERNAL order, in and out (the caller's copies travel through local memory as
    // 8/16-byte vectors: in the internal order those are the register pairs the hot loop works on, see ensure_so3)
    using N = num<T>;
    T y[14];
    from_z<T>(zio, y);
    const Dyn<T> d = *dp;
    OdeLane<T>& o = *op;
    const T min_step = (T)10 * N::ulp_up(o.t);
    const bool too_small = o.h_abs < min_step;
    T t_new = o.t + o.h_abs;
    if (t_new - Tend > (T)0) t_new = Tend;
    const T h = t_new - o.t;
    const T W3 = *W3p;
    T z[14], K[12][14], xv[12][3];
    to_z<T>(y, z);
#pragma unroll
    for (int i = 0; i < 14; ++i) K[0][i] = K0[i];
    xv[0][0] = z[10]; xv[0][1] = z[11]; xv[0][2] = z[9];
    int nproj = 0, bad = 0;
#pragma unroll 1
    for (int s = 1; s <= 11; ++s) {
        T P[14];
#pragma unroll
        for (int i = 0; i < 14; ++i) P[i] = 0;
#pragma unroll 1
        for (int j = 0; j < s; ++j) {
            const T a = (T)dop_a(s, j);
            if (a == (T)0) continue;
#pragma unroll
            for (int i = 0; i < 14; ++i) P[i] = N::fma(a, K[j][i], P[i]);
        }
#pragma unroll
        for (int i = 0; i < 14; ++i) P[i] = N::fma(h, P[i], z[i]);
        const T W3s = N::fma(h * (T)dop_c(s), d.w3dot, W3);
        xv[s][0] = P[10]; xv[s][1] = P[11]; xv[s][2] = P[9];
        if (!so3_ok_z<T>(P)) {
            T R[9];
            z_get_R<T>(P, R);
            const int pb = project_so3<T>(R);
            z_set_R<T>(R, P);
            nproj += 1; bad |= pb ? 2 : 0;
        }
        T kn[14];
        rhs_z<T>(P, W3s, d, kn);
#pragma unroll
        for (int i = 0; i < 14; ++i) K[s][i] = kn[i];
    }
    T sb[14], s5[14], s3[14], xb[3], x5[3], x3[3];
#pragma unroll
    for (int i = 0; i < 14; ++i) { sb[i] = 0; s5[i] = 0; s3[i] = 0; }
#pragma unroll
    for (int i = 0; i < 3; ++i) { xb[i] = 0; x5[i] = 0; x3[i] = 0; }
#pragma unroll 1
    for (int j = 0; j <= 11; ++j) {
        const T b = (T)dop_b(j), e5 = (T)dop_e5(j), e3 = (T)dop_e3(j);
        if (b == (T)0) continue;
#pragma unroll
        for (int i = 0; i < 14; ++i) { sb[i] = N::fma(b, K[j][i], sb[i]); s5[i] = N::fma(e5, K[j][i], s5[i]); s3[i] = N::fma(e3, K[j][i], s3[i]); }
#pragma unroll
        for (int i = 0; i < 3; ++i) { xb[i] = N::fma(b, xv[j][i], xb[i]); x5[i] = N::fma(e5, xv[j][i], x5[i]); x3[i] = N::fma(e3, xv[j][i], x3[i]); }
    }
    T xnew[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) xnew[i] = N::fma(h, xb[i], x[i]);
#pragma unroll
    for (int i = 0; i < 14; ++i) sb[i] = N::fma(h, sb[i], z[i]);
    o.checked = 0;
    T W3v = W3;
    const bool fin = dop853_conclude<T>(x, y, W3v, d, Tend, rtol, atol, K0, o, h, t_new, too_small, z, sb, s5, s3, xnew, x5, x3, nproj, bad);
    *W3p = W3v;
    to_z<T>(y, zio);
    return fin;
}

}  // namespace qr
