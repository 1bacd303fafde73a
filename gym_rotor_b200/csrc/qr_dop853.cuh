// qr_dop853.cuh -- rigid-body right-hand side + scipy's DOP853 controller, one env per thread, all
// stage derivatives resident in registers.
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
//   * Tableau sparsity: stages 1,2 die after stage 4, so at most 10 stage vectors are live.
//   * F(y_new) is only evaluated when another step follows (t_new < T): scipy evaluates it always but only
//     uses it as the next step's first stage; the RHS has no side effects.  nfev is still reported as
//     scipy counts it (2 + 12 per attempt).
//   * x^(+-1/8) in the step-size controller are square-root chains.
#pragma once
#include "qr_math.cuh"
#include "dop853_tableau.h"

namespace qr {

#define QR_A(s, j) ((T)DOP_A##s##_##j)

// linear combinations of stage derivatives K[j][i] for component i, by tableau row
#define QR_COMB1(K, i) (QR_A(1, 0) * K[0][i])
#define QR_COMB2(K, i) (N::fma(QR_A(2, 1), K[1][i], QR_A(2, 0) * K[0][i]))
#define QR_COMB3(K, i) (N::fma(QR_A(3, 2), K[2][i], QR_A(3, 0) * K[0][i]))
#define QR_COMB4(K, i) (N::fma(QR_A(4, 3), K[3][i], N::fma(QR_A(4, 2), K[2][i], QR_A(4, 0) * K[0][i])))
#define QR_COMB5(K, i) (N::fma(QR_A(5, 4), K[4][i], N::fma(QR_A(5, 3), K[3][i], QR_A(5, 0) * K[0][i])))
#define QR_COMB6(K, i) (N::fma(QR_A(6, 5), K[5][i], N::fma(QR_A(6, 4), K[4][i], N::fma(QR_A(6, 3), K[3][i], QR_A(6, 0) * K[0][i]))))
#define QR_COMB7(K, i) (N::fma(QR_A(7, 6), K[6][i], N::fma(QR_A(7, 5), K[5][i], N::fma(QR_A(7, 4), K[4][i], N::fma(QR_A(7, 3), K[3][i], QR_A(7, 0) * K[0][i])))))
#define QR_COMB8(K, i) (N::fma(QR_A(8, 7), K[7][i], N::fma(QR_A(8, 6), K[6][i], N::fma(QR_A(8, 5), K[5][i], N::fma(QR_A(8, 4), K[4][i], N::fma(QR_A(8, 3), K[3][i], QR_A(8, 0) * K[0][i]))))))
#define QR_COMB9(K, i) (N::fma(QR_A(9, 8), K[8][i], N::fma(QR_A(9, 7), K[7][i], N::fma(QR_A(9, 6), K[6][i], N::fma(QR_A(9, 5), K[5][i], N::fma(QR_A(9, 4), K[4][i], N::fma(QR_A(9, 3), K[3][i], QR_A(9, 0) * K[0][i])))))))
#define QR_COMB10(K, i) (N::fma(QR_A(10, 9), K[9][i], N::fma(QR_A(10, 8), K[8][i], N::fma(QR_A(10, 7), K[7][i], N::fma(QR_A(10, 6), K[6][i], N::fma(QR_A(10, 5), K[5][i], N::fma(QR_A(10, 4), K[4][i], N::fma(QR_A(10, 3), K[3][i], QR_A(10, 0) * K[0][i]))))))))
#define QR_COMB11(K, i) (N::fma(QR_A(11, 10), K[10][i], N::fma(QR_A(11, 9), K[9][i], N::fma(QR_A(11, 8), K[8][i], N::fma(QR_A(11, 7), K[7][i], N::fma(QR_A(11, 6), K[6][i], N::fma(QR_A(11, 5), K[5][i], N::fma(QR_A(11, 4), K[4][i], N::fma(QR_A(11, 3), K[3][i], QR_A(11, 0) * K[0][i])))))))))
#define QR_COMB_W(K, i, P) (N::fma((T)P##11, K[11][i], N::fma((T)P##10, K[10][i], N::fma((T)P##9, K[9][i], N::fma((T)P##8, K[8][i], N::fma((T)P##7, K[7][i], N::fma((T)P##6, K[6][i], N::fma((T)P##5, K[5][i], (T)P##0 * K[0][i]))))))))

// Layout of the 14 integrated components kept in registers: y[0..2] = v, y[3..11] = R (column-major),
// y[12..13] = W1, W2.  x[3] and W3 are carried separately.
template <typename T> struct Dyn {
    T fm;      // f / m          (thrust acceleration magnitude)
    T g;       // gravity
    T Mi0, Mi1;// M1/J1, M2/J1
    T kw0, kw1;// (J1 - J3)/J1 for W1' ; (J3 - J1)/J1 for W2'
    T w3dot;   // M3 / J3 : constant because J1 == J2 (quad.py:378)
};

// One right-hand-side evaluation at stage point (ys, W3s) -> k[14].  Returns ensure_SO3 flags.
template <typename T> QR_DEV int rhs14(const T* ys, T W3s, const Dyn<T>& d, T* k)
{
    using N = num<T>;
    T R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = ys[3 + i];
    int fl = ensure_so3<T>(R);  // state_decomposition -> ensure_SO3 on every call (quad_utils.py:12-16)
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

template <typename T> struct StepResult {
    int nfev;    // as scipy counts: 2 + 12 * attempts
    int status;  // QR_ST_* bits
    int nproj;   // SO(3) re-projections that fired inside RHS evaluations
};

// Integrates (x, y14, W3) over [0, Tend] in place with scipy's DOP853 driver.
template <typename T>
QR_DEV StepResult<T> dop853_step(T* x, T* y, T& W3, const Dyn<T>& d, const T Tend, const T rtol, const T atol)
{
    using N = num<T>;
    StepResult<T> res;
    res.nfev = 0; res.status = 0; res.nproj = 0;
    T K[12][14];
    int fl;

    // ---- RungeKutta.__init__: f0 and select_initial_step --------------------------------------------
    fl = rhs14<T>(y, W3, d, K[0]); res.nfev++;
    res.nproj += fl & 1; if (fl & 2) res.status |= 4;
    T h_abs;
    {
        T s0 = 0, s1 = 0;   // sums of (y/sc)^2 and (f0/sc)^2 over all 18 components
        T isc[14], iscx[3], iscw3;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            iscx[i] = N::recip(N::fma(N::abs(x[i]), rtol, atol));
            T a = x[i] * iscx[i], b = y[i] * iscx[i];   // x' = v
            s0 = N::fma(a, a, s0); s1 = N::fma(b, b, s1);
        }
#pragma unroll
        for (int i = 0; i < 14; ++i) {
            isc[i] = N::recip(N::fma(N::abs(y[i]), rtol, atol));
            T a = y[i] * isc[i], b = K[0][i] * isc[i];
            s0 = N::fma(a, a, s0); s1 = N::fma(b, b, s1);
        }
        {
            iscw3 = N::recip(N::fma(N::abs(W3), rtol, atol));
            T a = W3 * iscw3, b = d.w3dot * iscw3;
            s0 = N::fma(a, a, s0); s1 = N::fma(b, b, s1);
        }
        const T inv_sqrt_n = (T)0.23570226039551584;  // 1/sqrt(18)
        T d0 = N::sqrt(s0) * inv_sqrt_n, d1 = N::sqrt(s1) * inv_sqrt_n;
        T h0 = (d0 < (T)1e-5 || d1 < (T)1e-5) ? (T)1e-6 : (T)0.01 * d0 / d1;
        h0 = (Tend < h0) ? Tend : h0;   // python min(h0, interval): keeps a NaN h0
        // Euler probe y1 = y0 + h0 f0 ; f1 = F(y1)
        T y1[14], k1[14];
#pragma unroll
        for (int i = 0; i < 14; ++i) y1[i] = N::fma(h0, K[0][i], y[i]);
        T W31 = N::fma(h0, d.w3dot, W3);
        fl = rhs14<T>(y1, W31, d, k1); res.nfev++;
        res.nproj += fl & 1; if (fl & 2) res.status |= 4;
        T s2 = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {   // f1_x - f0_x = v1 - v0
            T a = (y1[i] - y[i]) * iscx[i];
            s2 = N::fma(a, a, s2);
        }
#pragma unroll
        for (int i = 0; i < 14; ++i) {
            T a = (k1[i] - K[0][i]) * isc[i];
            s2 = N::fma(a, a, s2);
        }
        // the W3 component of f1 - f0 is exactly zero
        T d2 = N::sqrt(s2) * inv_sqrt_n / h0;
        T h1;
        if (d1 <= (T)1e-15 && d2 <= (T)1e-15) {
            h1 = N::max((T)1e-6, h0 * (T)1e-3);
        } else {
            h1 = N::root8((T)0.01 / N::max(d1, d2));
        }
        h_abs = (T)100 * h0;
        h_abs = (h1 < h_abs) ? h1 : h_abs;
        h_abs = (Tend < h_abs) ? Tend : h_abs;
    }

    // ---- while t < Tend: OdeSolver.step -> _step_impl -----------------------------------------------
    T t = 0;
    while (t < Tend) {
        const T min_step = (T)10 * N::abs(N::nextafter(t, N::inf()) - t);
        if (h_abs < min_step) h_abs = min_step;
        bool rejected = false;
        for (;;) {
            if (h_abs < min_step) { res.status |= 2; return res; }  // TOO_SMALL_STEP: keep last accepted y
            T t_new = t + h_abs;
            if (t_new - Tend > (T)0) t_new = Tend;
            const T h = t_new - t;
            h_abs = N::abs(h);
            res.nfev += 12;

            // running sums for x (x' = v): B, E5 and E3 weighted stage velocities
            T xb[3], x5[3], x3[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) { xb[i] = (T)DOP_B0 * y[i]; x5[i] = (T)DOP_E5_0 * y[i]; x3[i] = (T)DOP_E3_0 * y[i]; }

            T ys[14];
#define QR_STAGE(S, COMB, CS, USE_X, BS, E5S, E3S)                                                         \
    {                                                                                                      \
        _Pragma("unroll") for (int i = 0; i < 14; ++i) ys[i] = N::fma(h, COMB(K, i), y[i]);                 \
        T W3s = N::fma(h * (T)(CS), d.w3dot, W3);                                                           \
        if (USE_X) {                                                                                       \
            _Pragma("unroll") for (int i = 0; i < 3; ++i) {                                                 \
                xb[i] = N::fma((T)(BS), ys[i], xb[i]);                                                      \
                x5[i] = N::fma((T)(E5S), ys[i], x5[i]);                                                     \
                x3[i] = N::fma((T)(E3S), ys[i], x3[i]);                                                     \
            }                                                                                              \
        }                                                                                                  \
        fl = rhs14<T>(ys, W3s, d, K[S]);                                                                    \
        res.nproj += fl & 1; if (fl & 2) res.status |= 4;                                                   \
    }
            QR_STAGE(1, QR_COMB1, DOP_C1, 0, 0, 0, 0)
            QR_STAGE(2, QR_COMB2, DOP_C2, 0, 0, 0, 0)
            QR_STAGE(3, QR_COMB3, DOP_C3, 0, 0, 0, 0)
            QR_STAGE(4, QR_COMB4, DOP_C4, 0, 0, 0, 0)
            QR_STAGE(5, QR_COMB5, DOP_C5, 1, DOP_B5, DOP_E5_5, DOP_E3_5)
            QR_STAGE(6, QR_COMB6, DOP_C6, 1, DOP_B6, DOP_E5_6, DOP_E3_6)
            QR_STAGE(7, QR_COMB7, DOP_C7, 1, DOP_B7, DOP_E5_7, DOP_E3_7)
            QR_STAGE(8, QR_COMB8, DOP_C8, 1, DOP_B8, DOP_E5_8, DOP_E3_8)
            QR_STAGE(9, QR_COMB9, DOP_C9, 1, DOP_B9, DOP_E5_9, DOP_E3_9)
            QR_STAGE(10, QR_COMB10, DOP_C10, 1, DOP_B10, DOP_E5_10, DOP_E3_10)
            QR_STAGE(11, QR_COMB11, DOP_C11, 1, DOP_B11, DOP_E5_11, DOP_E3_11)
#undef QR_STAGE

            // y_new = y + h * sum_s B_s K_s ; error estimates (rk.py:683-691)
            T ynew[14], xnew[3];
            T e5n = 0, e3n = 0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                xnew[i] = N::fma(h, xb[i], x[i]);
                T isc = N::recip(N::fma(N::max(N::abs(x[i]), N::abs(xnew[i])), rtol, atol));
                T e5 = x5[i] * isc, e3 = x3[i] * isc;
                e5n = N::fma(e5, e5, e5n); e3n = N::fma(e3, e3, e3n);
            }
#pragma unroll
            for (int i = 0; i < 14; ++i) {
                ynew[i] = N::fma(h, QR_COMB_W(K, i, DOP_B), y[i]);
                T isc = N::recip(N::fma(N::max(N::abs(y[i]), N::abs(ynew[i])), rtol, atol));
                T e5 = QR_COMB_W(K, i, DOP_E5_) * isc, e3 = QR_COMB_W(K, i, DOP_E3_) * isc;
                e5n = N::fma(e5, e5, e5n); e3n = N::fma(e3, e3, e3n);
            }
            T err;
            if (e5n == (T)0 && e3n == (T)0) err = 0;
            else err = N::abs(h) * e5n * N::rsqrt((e5n + (T)0.01 * e3n) * (T)18);

            if (err < (T)1) {
                // accept
#pragma unroll
                for (int i = 0; i < 3; ++i) x[i] = xnew[i];
#pragma unroll
                for (int i = 0; i < 14; ++i) y[i] = ynew[i];
                W3 = N::fma(h, d.w3dot, W3);
                t = t_new;
                if (t < Tend) {
                    T factor = (err == (T)0) ? (T)10 : N::min((T)10, (T)0.9 * N::inv_root8(err));
                    if (rejected) factor = N::min((T)1, factor);
                    h_abs *= factor;
                    fl = rhs14<T>(y, W3, d, K[0]);  // f_new becomes the next step's first stage
                    res.nproj += fl & 1; if (fl & 2) res.status |= 4;
                }
                break;
            }
            // A NaN error norm also lands here (nan < 1 is False).  scipy then shrinks h by 0.2 until
            // TOO_SMALL_STEP and solve_ivp returns the last accepted y; that outcome is produced at once.
            if (err != err) { res.status |= 1 | 2; return res; }
            h_abs *= N::max((T)0.2, (T)0.9 * N::inv_root8(err));
            rejected = true;
        }
    }
    return res;
}

}  // namespace qr
