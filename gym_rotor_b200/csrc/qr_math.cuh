// qr_math.cuh -- scalar-type traits and small SO(3) helpers for the step kernels (sm_100a).
//
// Reference semantics restated here (never copied):
//   ensure_SO3 / psvd ........ gym_rotor/envs/quad_utils.py:123-142, 226-240
//   hat ...................... gym_rotor/envs/quad_utils.py:80-85
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define QR_DEV __device__ __forceinline__
// PTX is only emitted in device passes; the host pass of nvcc -- and tests/host_twin, which compiles these headers
// with g++ to unit-test the arithmetic without a GPU -- sees plain C for the few inline-PTX helpers
#ifdef __CUDA_ARCH__
#define QR_PTX 1
#else
#define QR_PTX 0
#endif

namespace qr {

// A copy in a register of its own.  The observation row leaves the kernel as 16-byte vectors, whose registers must be
// consecutive; entries that are plain copies of the attitude (float32 mode: no conversion in between) would otherwise
// tie the persistent state to that grouping, which conflicts with the integrator's register pairs (qr_dop853.cuh) and
// costs moves in every stage instead of once per step.
#if QR_PTX
// (not volatile: a copy nobody reads -- the attitude entries of an observation row that is not stored -- may be dropped)
QR_DEV float own_reg(float v) { float r; asm("mov.b32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
QR_DEV double own_reg(double v) { double r; asm("mov.b64 %0, %1;" : "=d"(r) : "d"(v)); return r; }
#else
QR_DEV float own_reg(float v) { return v; }
QR_DEV double own_reg(double v) { return v; }
#endif

template <typename T> struct num;

template <> struct num<float> {
    static QR_DEV float fma(float a, float b, float c) { return fmaf(a, b, c); }
    static QR_DEV float abs(float a) { return fabsf(a); }
    static QR_DEV float sqrt(float a) { return sqrtf(a); }
    static QR_DEV float rsqrt(float a) { return rsqrtf(a); }
    static QR_DEV float max(float a, float b) { return fmaxf(a, b); }
    static QR_DEV float min(float a, float b) { return fminf(a, b); }
    // atan2 by octant reduction + the degree-16 even polynomial for atan(a)/a on [0, 1] of Abramowitz & Stegun 4.4.49 (|error|
    // <= 2e-8; measured 3.0e-7 against float64 over 2e6 random arguments in float32 arithmetic -- atan2f: 3.2e-7), half the
    // instructions of atan2f and no slow path; float32 mode only (float64 calls ::atan2)
    static QR_DEV float atan2(float y, float x)
    {
        const float ax = fabsf(x), ay = fabsf(y);
        const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
        const float a = (mx > 0.f) ? mn * recip(mx) : 0.f;
        const float s = a * a;
        float p = 0.0028662257f;
        p = fmaf(p, s, -0.0161657367f); p = fmaf(p, s, 0.0429096138f); p = fmaf(p, s, -0.0752896400f);
        p = fmaf(p, s, 0.1065626393f); p = fmaf(p, s, -0.1420889944f); p = fmaf(p, s, 0.1999355085f);
        p = fmaf(p, s, -0.3333314528f); p = fmaf(p, s, 1.0f);
        float r = a * p;
        r = (ay > ax) ? 1.57079632679489662f - r : r;
        r = (x < 0.f) ? 3.14159265358979324f - r : r;
        return copysignf(r, y);
    }
    static QR_DEV float nextafter(float a, float b) { return nextafterf(a, b); }
    // spacing of the floats above t (t >= 0, finite): what scipy takes as |nextafter(t, inf) - t|
    static QR_DEV float ulp_up(float t) { return __int_as_float(__float_as_int(t) + 1) - t; }
    static QR_DEV float inf() { return __int_as_float(0x7f800000); }
    // x^(1/8) and x^(-1/8) by square-root chains: <= 2 ulp, no transcendental (only scales the step size)
#if QR_PTX
    static QR_DEV float asqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else
    static QR_DEV float asqrt(float x) { return sqrtf(x); }
#endif
    static QR_DEV float root8(float x) { return asqrt(asqrt(asqrt(x))); }
    static QR_DEV float inv_root8(float x) { return rsqrtf(asqrt(asqrt(x))); }
    // reciprocal of an error scale (feeds norms that only steer the step size): MUFU.RCP, <= 1 ulp
#if QR_PTX
    static QR_DEV float recip(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else
    static QR_DEV float recip(float x) { return 1.0f / x; }
#endif
    // max(acc, |a|, |b|) that PROPAGATES a NaN (fmaxf drops it): a running maximum whose final comparison fails on a NaN
    // like every comparison of the quantities it absorbed would have
#if QR_PTX
    static QR_DEV float absmax3_nan(float acc, float a, float b)
    {
        float r;
        asm("{\n\t.reg .f32 ta, tb;\n\tabs.f32 ta, %2;\n\tabs.f32 tb, %3;\n\tmax.NaN.f32 %0, %1, ta, tb;\n\t}" : "=f"(r) : "f"(acc), "f"(a), "f"(b));
        return r;
    }
    static QR_DEV float absmax2_nan(float acc, float a)
    {
        float r;
        asm("{\n\t.reg .f32 ta;\n\tabs.f32 ta, %2;\n\tmax.NaN.f32 %0, %1, ta;\n\t}" : "=f"(r) : "f"(acc), "f"(a));
        return r;
    }
#else
    static QR_DEV float absmax3_nan(float acc, float a, float b) { return (acc != acc || a != a || b != b) ? NAN : fmaxf(acc, fmaxf(fabsf(a), fabsf(b))); }
    static QR_DEV float absmax2_nan(float acc, float a) { return (acc != acc || a != a) ? NAN : fmaxf(acc, fabsf(a)); }
#endif
    // quotient where a relative error of ~2 ulp is immaterial (step-size heuristics): MUFU.RCP + multiply, no slow path
    static QR_DEV float div_fast(float a, float b) { return a * recip(b); }
    static constexpr float eps_jacobi = 1e-7f;
    static constexpr float huge = 3.0e38f;
};

template <> struct num<double> {
    static QR_DEV double fma(double a, double b, double c) { return ::fma(a, b, c); }
    static QR_DEV double abs(double a) { return fabs(a); }
    static QR_DEV double sqrt(double a) { return ::sqrt(a); }
    static QR_DEV double rsqrt(double a) { return 1.0 / ::sqrt(a); }
    static QR_DEV double max(double a, double b) { return fmax(a, b); }
    static QR_DEV double min(double a, double b) { return fmin(a, b); }
    static QR_DEV double atan2(double a, double b) { return ::atan2(a, b); }
    static QR_DEV double nextafter(double a, double b) { return ::nextafter(a, b); }
    static QR_DEV double ulp_up(double t) { return __longlong_as_double(__double_as_longlong(t) + 1) - t; }
    static QR_DEV double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
    static QR_DEV double root8(double x) { return ::sqrt(::sqrt(::sqrt(x))); }
    static QR_DEV double inv_root8(double x) { return 1.0 / ::sqrt(::sqrt(::sqrt(x))); }
    static QR_DEV double recip(double x) { return 1.0 / x; }
    static QR_DEV double absmax3_nan(double acc, double a, double b) { return (acc != acc || a != a || b != b) ? (acc + a + b) : fmax(acc, fmax(fabs(a), fabs(b))); }
    static QR_DEV double absmax2_nan(double acc, double a) { return (acc != acc || a != a) ? (acc + a) : fmax(acc, fabs(a)); }
    static QR_DEV double div_fast(double a, double b) { return a / b; }   // float64 is the parity mode: numpy's exact quotient
    static constexpr double eps_jacobi = 1e-16;
    static constexpr double huge = 1.0e300;
};

// R is column-major: R[i + 3*j] = R_ij, i.e. R[0..2] = b1, R[3..5] = b2, R[6..8] = b3.
template <typename T> QR_DEV T det3(const T* A)
{
    return A[0] * (A[4] * A[8] - A[7] * A[5]) - A[3] * (A[1] * A[8] - A[7] * A[2]) + A[6] * (A[1] * A[5] - A[4] * A[2]);
}

// The acceptance test of ensure_SO3 (quad_utils.py:133-136):
//   np.allclose(R.T@R, I, rtol=1e-5, atol=1e-5)  <=> |RtR_ij - I_ij| <= 1e-5 + 1e-5*I_ij   (diag 2e-5, off-diag 1e-5)
//   np.isclose(det R, 1, rtol=1e-5)              <=> |det - 1| <= 1e-8 + 1e-5
// NaNs fail every comparison, exactly like numpy.  This runs inside EVERY right-hand-side evaluation
// (state_decomposition, quad_utils.py:12-16), so it is written branch-free: one predicate at the end.
template <typename T> QR_DEV bool so3_ok(const T* R, T* defect = nullptr)
{
    using N = num<T>;
    const T tol = (T)1e-5;
    T d00 = N::fma(R[2], R[2], N::fma(R[1], R[1], R[0] * R[0]));
    T d11 = N::fma(R[5], R[5], N::fma(R[4], R[4], R[3] * R[3]));
    T d22 = N::fma(R[8], R[8], N::fma(R[7], R[7], R[6] * R[6]));
    T d01 = N::fma(R[2], R[5], N::fma(R[1], R[4], R[0] * R[3]));
    T d02 = N::fma(R[2], R[8], N::fma(R[1], R[7], R[0] * R[6]));
    T d12 = N::fma(R[5], R[8], N::fma(R[4], R[7], R[3] * R[6]));
    T dg = N::max(N::max(N::abs(d00 - (T)1), N::abs(d11 - (T)1)), N::abs(d22 - (T)1));
    T od = N::max(N::max(N::abs(d01), N::abs(d02)), N::abs(d12));
    // float32: det R - 1 = tr(R^T R - I) / 2 + O(1e-9) wherever the test can pass (see so3_ok_z in qr_dop853.cuh)
    T dt = (sizeof(T) == 4) ? N::abs((T)0.5 * (((d00 - (T)1) + (d11 - (T)1)) + (d22 - (T)1))) : N::abs(det3(R) - (T)1);
    // fmax drops NaNs, so test finiteness through a sum that propagates them
    T nanprobe = (d00 + d11 + d22) + (d01 + d02 + d12);
    if (defect) *defect = (nanprobe == nanprobe) ? N::max(dg, od) : (T)1;   // max |RtR - I| (1 if not finite)
    return (dg <= tol + tol) && (od <= tol) && (dt <= (T)1e-8 + tol) && (nanprobe == nanprobe);
}

// psvd + "U @ VT.T" (quad_utils.py:138-140, 226-240): R <- U diag(1,1,det U det Vt) Vt by one-sided
// Jacobi.  Rare path (only the Euler probe of the initial-step selection leaves SO(3) by > 1e-5), kept
// out of line so that it does not add register pressure to the integrator.  Returns 1 on failure.
// ZORDER: the caller hands the matrix over in the integrator's internal component order (qr_so3_zpos), see ensure_so3.
QR_DEV constexpr int qr_so3_zpos(int i) { return i == 0 ? 0 : i == 1 ? 1 : i == 2 ? 6 : i == 3 ? 2 : i == 4 ? 3 : i == 5 ? 7 : i == 6 ? 4 : i == 7 ? 5 : 8; }
template <typename T, bool ZORDER = false> __device__ __noinline__ int project_so3(T* R)
{
    using N = num<T>;
    T A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; ++i) {
        A[i] = R[ZORDER ? qr_so3_zpos(i) : i];
        if (!(N::abs(A[i]) <= N::huge)) return 1;
    }
    for (int sweep = 0; sweep < 30; ++sweep) {
        int rotated = 0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                T al = 0, be = 0, ga = 0;
                for (int k = 0; k < 3; ++k) {
                    al += A[k + 3 * p] * A[k + 3 * p];
                    be += A[k + 3 * q] * A[k + 3 * q];
                    ga += A[k + 3 * p] * A[k + 3 * q];
                }
                if (N::abs(ga) <= N::eps_jacobi * N::sqrt(al * be) || ga == 0) continue;
                rotated = 1;
                T zeta = (be - al) / ((T)2 * ga);
                T t = (zeta >= 0 ? (T)1 : (T)-1) / (N::abs(zeta) + N::sqrt((T)1 + zeta * zeta));
                T c = (T)1 / N::sqrt((T)1 + t * t), s = c * t;
                for (int k = 0; k < 3; ++k) {
                    T ap = A[k + 3 * p], aq = A[k + 3 * q];
                    A[k + 3 * p] = c * ap - s * aq;
                    A[k + 3 * q] = s * ap + c * aq;
                    T vp = V[k + 3 * p], vq = V[k + 3 * q];
                    V[k + 3 * p] = c * vp - s * vq;
                    V[k + 3 * q] = s * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    T sig[3], U[9];
    int bad = 0, jmin = 0;
    for (int j = 0; j < 3; ++j) {
        sig[j] = N::sqrt(A[3 * j] * A[3 * j] + A[1 + 3 * j] * A[1 + 3 * j] + A[2 + 3 * j] * A[2 + 3 * j]);
        if (sig[j] < sig[jmin]) jmin = j;
    }
    for (int j = 0; j < 3; ++j) {
        if (!(sig[j] > 0)) { bad = 1; for (int k = 0; k < 3; ++k) U[k + 3 * j] = 0; continue; }
        for (int k = 0; k < 3; ++k) U[k + 3 * j] = A[k + 3 * j] / sig[j];
    }
    if (bad) {
        int a = (jmin + 1) % 3, b = (jmin + 2) % 3;
        if (!(sig[a] > 0) || !(sig[b] > 0)) return 1;
        U[0 + 3 * jmin] = U[1 + 3 * a] * U[2 + 3 * b] - U[2 + 3 * a] * U[1 + 3 * b];
        U[1 + 3 * jmin] = U[2 + 3 * a] * U[0 + 3 * b] - U[0 + 3 * a] * U[2 + 3 * b];
        U[2 + 3 * jmin] = U[0 + 3 * a] * U[1 + 3 * b] - U[1 + 3 * a] * U[0 + 3 * b];
    }
    T sgn = det3(U) * det3(V);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            T s = 0;
            for (int k = 0; k < 3; ++k) {
                T u = U[i + 3 * k];
                if (k == jmin) u *= sgn;
                s += u * V[j + 3 * k];
            }
            R[ZORDER ? qr_so3_zpos(i + 3 * j) : i + 3 * j] = s;
        }
    return bad;
}

// Polar factor of a NEAR-orthogonal matrix by Newton's iteration X <- (X + X^-T)/2 (Higham).  For det X > 0
// the limit is U V^T, i.e. exactly what psvd + "U @ VT.T" returns (the sign correction is the identity);
// convergence is quadratic, so an orthogonality defect of 1e-2 needs three steps to reach float64 rounding.
// This is the common case (the Euler probe of select_initial_step leaves SO(3) by (h0 |W|)^2 ~ 1e-4..1e-2 in
// ~12 % of env-steps); it stays in registers.  Returns false if the input is not in that regime.
// float32: the number of steps is fixed a priori from the defect d = max |X^T X - I| measured by so3_ok (singular
// values within 1.5 d of 1; one step maps an error e to e^2/2): 1 step for d <= 2e-4, 2 for d <= 1.5e-2, else 3
// reach float32 rounding -- no extra step just to observe convergence (all lanes of a warp pay for the slowest).
template <typename T> QR_DEV bool polar_newton(T* R, T defect)
{
    using N = num<T>;
    T X[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) X[i] = R[i];
    const bool apriori = sizeof(T) == 4;
    const int iters = (sizeof(T) == 8) ? 5 : (defect <= (T)2e-4 ? 1 : (defect <= (T)1.5e-2 ? 2 : 3));
    T delta = 0;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        // cofactor matrix C (C_ij = d det / d X_ij); X^-T = C / det
        T C[9];
        // column-major: X[i + 3j] = X_ij ; C[i + 3j] = cofactor of X_ij
        C[0] = X[4] * X[8] - X[7] * X[5]; C[3] = X[7] * X[2] - X[1] * X[8]; C[6] = X[1] * X[5] - X[4] * X[2];
        C[1] = X[6] * X[5] - X[3] * X[8]; C[4] = X[0] * X[8] - X[6] * X[2]; C[7] = X[3] * X[2] - X[0] * X[5];
        C[2] = X[3] * X[7] - X[6] * X[4]; C[5] = X[6] * X[1] - X[0] * X[7]; C[8] = X[0] * X[4] - X[3] * X[1];
        const T det = X[0] * C[0] + X[1] * C[1] + X[2] * C[2];
        if (!(det > (T)0.5)) return false;
        const T hid = (T)0.5 / det;
        delta = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const T xn = N::fma(C[i], hid, (T)0.5 * X[i]);
            if (!apriori) delta = N::max(delta, N::abs(xn - X[i]));
            X[i] = xn;
        }
        if (!apriori && delta <= (T)1e-15) break;
    }
    if (!apriori && !(delta <= (T)1e-12)) return false;
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = X[i];
    return true;
}

// ensure_SO3 on a register-resident R; flags: bit0 = projected, bit1 = projection failed
// NEWTON: try the in-register Newton polar iteration first (used where the projection is expected to fire:
// the Euler probe); everywhere else the rare projection goes straight to the out-of-line Jacobi routine.
template <typename T, bool NEWTON = false> QR_DEV int ensure_so3(T* R)
{
    T defect;
    if (so3_ok(R, &defect)) return 0;
    if (NEWTON) { if (defect < (T)0.05 && polar_newton<T>(R, defect)) return 1; }
    // (rare path) the matrix travels to the out-of-line routine through local memory as 16-byte vectors, i.e. in groups of
    // four consecutive registers.  It is handed over in the integrator's internal order (qr_dop853.cuh: b1.xy b2.xy | b3.xy
    // b1.z b2.z | b3.z), so that those groups are unions of the register PAIRS the packed instructions work on: in
    // column-major order ptxas laid the caller's persistent attitude out for these stores and assembled three of the seven
    // pairs with two moves at each of their uses in the hot loop
    T tmp[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) tmp[qr_so3_zpos(i)] = R[i];
    int bad = project_so3<T, true>(tmp);
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = own_reg(tmp[qr_so3_zpos(i)]);
    return 1 | (bad << 1);
}

}  // namespace qr
