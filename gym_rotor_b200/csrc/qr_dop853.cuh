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
//   * Stage derivatives K1..K11 live in SHARED MEMORY ([slot][component][lane], conflict-free 128/64-bit
//     accesses), K0 in registers; the stage loop is ROLLED with the tableau in constant memory.  A fully
//     unrolled register version (round-1 capture A) was 206 KB of SASS and instruction-fetch bound.
//     Tableau sparsity: row s uses K0 and the contiguous range K[jlo(s)..s-1]; K1/K3 and K2/K4 share slots.
//   * One ATTEMPT is a function: the caller owns the accept/reject loop, so that lanes of a warp that need
//     a second attempt do not hold back lanes that are ready for their next env (see qr_kernels.cuh).
//   * F(y_new) is only evaluated when another step follows (t_new < T): scipy evaluates it always but only
//     uses it as the next step's first stage; the RHS has no side effects.  nfev is still reported as
//     scipy counts it (2 + 12 per attempt).
//   * x^(+-1/8) in the step-size controller are square-root chains.
#pragma once
#include "qr_math.cuh"
#include "dop853_tableau.h"

// experimental (off: not measured yet): E3 equals B except in entries 0, 8 and 11 (scipy builds it that way,
// dop853_coefficients.py), so the third-order error sums are formed as (B-weighted sum) + (three correction terms);
// the x sums skip stages 1-4, whose B / E5 / E3 weights are zero.  ~1.5 % fewer instructions per attempt.
#ifndef QR_E3_FROM_B
#define QR_E3_FROM_B 0
#endif
#ifndef QR_PIPE
#define QR_PIPE 1
#endif
namespace qr {

// ---- tableau in constant memory (uniform-indexed loads in the rolled stage loop) ------------------------
// P / Ps: the 45 non-zero couplings A[s][j], j >= 1, flattened in stage order (entries Ps[s] .. Ps[s+1]-1 belong
// to stage s), each with the byte offset of K_j's slot inside a warp's stage storage: the inner loop of the
// stage sums is one 8/16-byte constant load + one add per K vector instead of index arithmetic.
template <typename T> struct TabEntry { T c; int off; };
struct Tableau {
    double A[12][12];
    double B[12], E5[12], E3[12], C[12];
    TabEntry<double> P[48];
    int Ps[16];
};
struct TableauF {
    float A[12][12];
    float B[12], E5[12], E3[12], C[12];
    TabEntry<float> P[48];
    int Ps[16];
};
__constant__ Tableau c_tab64;
__constant__ TableauF c_tab32;

inline void fill_tableau(Tableau& t)
{
    for (int i = 0; i < 12; ++i) {
        for (int j = 0; j < 12; ++j) t.A[i][j] = 0;
        t.B[i] = t.E5[i] = t.E3[i] = t.C[i] = 0;
    }
#define QR_SETA(s, j) t.A[s][j] = DOP_A##s##_##j
    QR_SETA(1, 0); QR_SETA(2, 0); QR_SETA(2, 1); QR_SETA(3, 0); QR_SETA(3, 2); QR_SETA(4, 0); QR_SETA(4, 2); QR_SETA(4, 3);
    QR_SETA(5, 0); QR_SETA(5, 3); QR_SETA(5, 4); QR_SETA(6, 0); QR_SETA(6, 3); QR_SETA(6, 4); QR_SETA(6, 5);
    QR_SETA(7, 0); QR_SETA(7, 3); QR_SETA(7, 4); QR_SETA(7, 5); QR_SETA(7, 6);
    QR_SETA(8, 0); QR_SETA(8, 3); QR_SETA(8, 4); QR_SETA(8, 5); QR_SETA(8, 6); QR_SETA(8, 7);
    QR_SETA(9, 0); QR_SETA(9, 3); QR_SETA(9, 4); QR_SETA(9, 5); QR_SETA(9, 6); QR_SETA(9, 7); QR_SETA(9, 8);
    QR_SETA(10, 0); QR_SETA(10, 3); QR_SETA(10, 4); QR_SETA(10, 5); QR_SETA(10, 6); QR_SETA(10, 7); QR_SETA(10, 8); QR_SETA(10, 9);
    QR_SETA(11, 0); QR_SETA(11, 3); QR_SETA(11, 4); QR_SETA(11, 5); QR_SETA(11, 6); QR_SETA(11, 7); QR_SETA(11, 8); QR_SETA(11, 9); QR_SETA(11, 10);
#undef QR_SETA
    t.B[0] = DOP_B0; t.B[5] = DOP_B5; t.B[6] = DOP_B6; t.B[7] = DOP_B7; t.B[8] = DOP_B8; t.B[9] = DOP_B9; t.B[10] = DOP_B10; t.B[11] = DOP_B11;
    t.E5[0] = DOP_E5_0; t.E5[5] = DOP_E5_5; t.E5[6] = DOP_E5_6; t.E5[7] = DOP_E5_7; t.E5[8] = DOP_E5_8; t.E5[9] = DOP_E5_9; t.E5[10] = DOP_E5_10; t.E5[11] = DOP_E5_11;
    t.E3[0] = DOP_E3_0; t.E3[5] = DOP_E3_5; t.E3[6] = DOP_E3_6; t.E3[7] = DOP_E3_7; t.E3[8] = DOP_E3_8; t.E3[9] = DOP_E3_9; t.E3[10] = DOP_E3_10; t.E3[11] = DOP_E3_11;
    t.C[1] = DOP_C1; t.C[2] = DOP_C2; t.C[3] = DOP_C3; t.C[4] = DOP_C4; t.C[5] = DOP_C5; t.C[6] = DOP_C6;
    t.C[7] = DOP_C7; t.C[8] = DOP_C8; t.C[9] = DOP_C9; t.C[10] = DOP_C10; t.C[11] = DOP_C11;
}

template <typename T> struct tab;
template <> struct tab<double> {
    static QR_DEV double A(int s, int j) { return c_tab64.A[s][j]; }
    static QR_DEV double B(int j) { return c_tab64.B[j]; }
    static QR_DEV double E5(int j) { return c_tab64.E5[j]; }
    static QR_DEV double E3(int j) { return c_tab64.E3[j]; }
    static QR_DEV double C(int j) { return c_tab64.C[j]; }
    static QR_DEV TabEntry<double> P(int p) { return c_tab64.P[p]; }
    static QR_DEV int Ps(int s) { return c_tab64.Ps[s]; }
};
template <> struct tab<float> {
    static QR_DEV float A(int s, int j) { return c_tab32.A[s][j]; }
    static QR_DEV float B(int j) { return c_tab32.B[j]; }
    static QR_DEV float E5(int j) { return c_tab32.E5[j]; }
    static QR_DEV float E3(int j) { return c_tab32.E3[j]; }
    static QR_DEV float C(int j) { return c_tab32.C[j]; }
    static QR_DEV TabEntry<float> P(int p) { return c_tab32.P[p]; }
    static QR_DEV int Ps(int s) { return c_tab32.Ps[s]; }
};

// ---- shared-memory stage storage ----------------------------------------------------------------------------
// Per warp: KS[slot 0..7][14 components][32 lanes] of T.  Within a slot the 14 components of one lane are
// packed as 3 x (4 consecutive T) + 1 x (2 consecutive T) so that float accesses are LDS/STS.128 + .64:
//   element (c, lane): c < 12 -> ((c >> 2) * 32 + lane) * 4 + (c & 3) ;  c >= 12 -> 384 + lane * 2 + (c - 12)
constexpr int QR_NSLOTS = 8;
constexpr int QR_SLOT_ELEMS = 14 * 32;
QR_DEV int k_slot(int j) { return j < 5 ? ((j + 1) & 1) : (j == 11 ? 0 : j - 3); }
inline int k_slot_host(int j) { return j < 5 ? ((j + 1) & 1) : (j == 11 ? 0 : j - 3); }
// K1,K3 -> 0 ; K2,K4 -> 1 ; K5..K10 -> 2..7 ; K11 -> 0 again (K3 is last read by the stage that produces K11)

template <typename T> struct vec4 { T a, b, c, d; };
template <typename T> struct vec2 { T a, b; };

// `col` = slot base + lane * 4 elements.  float: groups g=0..2 at col + g*128 (16 B each), tail at
// slot + 384 + lane*2 = col + 384 - lane*2.  double: 7 groups of 2 at slot + (g*32 + lane)*2 = col + g*64 - lane*2.
template <typename T> QR_DEV void ks_load_lane(const T* col, int lane, T* k)
{
    if (sizeof(T) == 4) {
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            float4 v = *reinterpret_cast<const float4*>(col + g * 128);
            k[4 * g] = (T)v.x; k[4 * g + 1] = (T)v.y; k[4 * g + 2] = (T)v.z; k[4 * g + 3] = (T)v.w;
        }
        float2 w = *reinterpret_cast<const float2*>(col + 384 - lane * 2);
        k[12] = (T)w.x; k[13] = (T)w.y;
    } else {
        const T* p = col - lane * 2;
#pragma unroll
        for (int g = 0; g < 7; ++g) {
            double2 v = *reinterpret_cast<const double2*>(p + g * 64);
            k[2 * g] = (T)v.x; k[2 * g + 1] = (T)v.y;
        }
    }
}
// Same as ks_load_lane for float, from a 32-bit shared-window address (`sa` = address of the lane's column in the
// slot, `lane8` = lane * 8): explicit ld.shared with immediate offsets, so the loop needs one add per K vector.
QR_DEV void ks_load_lane_sa(unsigned sa, unsigned lane8, float* k)
{
#if QR_PTX
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(k[0]), "=f"(k[1]), "=f"(k[2]), "=f"(k[3]) : "r"(sa) : "memory");
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+512];" : "=f"(k[4]), "=f"(k[5]), "=f"(k[6]), "=f"(k[7]) : "r"(sa) : "memory");
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+1024];" : "=f"(k[8]), "=f"(k[9]), "=f"(k[10]), "=f"(k[11]) : "r"(sa) : "memory");
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+1536];" : "=f"(k[12]), "=f"(k[13]) : "r"(sa - lane8) : "memory");
#else
    (void)sa; (void)lane8; (void)k;   // host builds take the generic-pointer loader (see dop853_attempt)
#endif
}
QR_DEV void ks_load_lane_sa(unsigned, unsigned, double*) {}   // float64 uses the generic-pointer loader

template <typename T> QR_DEV void ks_store_lane(T* col, int lane, const T* k)
{
    if (sizeof(T) == 4) {
#pragma unroll
        for (int g = 0; g < 3; ++g)
            *reinterpret_cast<float4*>(col + g * 128) = make_float4((float)k[4 * g], (float)k[4 * g + 1], (float)k[4 * g + 2], (float)k[4 * g + 3]);
        *reinterpret_cast<float2*>(col + 384 - lane * 2) = make_float2((float)k[12], (float)k[13]);
    } else {
        T* p = col - lane * 2;
#pragma unroll
        for (int g = 0; g < 7; ++g) *reinterpret_cast<double2*>(p + g * 64) = make_double2((double)k[2 * g], (double)k[2 * g + 1]);
    }
}

// acc[0..13] += c * k[0..13].  float32 on sm_100a: seven packed FFMA2 (fma.rn.f32x2) instead of fourteen FFMA --
// the kernel is issue bound, not FMA-pipe bound, so halving the instruction count of the stage sums pays.
template <typename T> QR_DEV void axpy14(T c, const T* k, T* acc)
{
#pragma unroll
    for (int i = 0; i < 14; ++i) acc[i] = num<T>::fma(c, k[i], acc[i]);
}
template <> QR_DEV void axpy14<float>(float c, const float* k, float* acc)
{
    const float2 cc = make_float2(c, c);
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        float2 r = __ffma2_rn(cc, make_float2(k[2 * i], k[2 * i + 1]), make_float2(acc[2 * i], acc[2 * i + 1]));
        acc[2 * i] = r.x; acc[2 * i + 1] = r.y;
    }
}
// out[0..13] = base[0..13] + c * k[0..13]   (out of place: no copy of `base` first)
template <typename T> QR_DEV void axpy14_out(T c, const T* k, const T* base, T* out)
{
#pragma unroll
    for (int i = 0; i < 14; ++i) out[i] = num<T>::fma(c, k[i], base[i]);
}
// (no packed specialisation: `base` is the persistent state, whose registers are not pair-aligned -- packing
//  it costs one MOV per element, more than the FFMA2 saves; the results land directly in the pair-aligned `out`)
// three weighted sums at once (y_new and the two error estimators)
template <typename T> QR_DEV void axpy14x3(T b, T e5, T e3, const T* k, T* sb, T* s5, T* s3)
{
    axpy14<T>(b, k, sb); axpy14<T>(e5, k, s5); axpy14<T>(e3, k, s3);
}

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
// CHECK = false: the caller has just run ensure_SO3 on this very matrix (it passed, or it was re-projected and
// passes now), so the reference's test inside this evaluation is known to succeed and is not repeated.
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

// Integrator state of one lane between attempts.
template <typename T> struct OdeLane {
    T t, h_abs;
    int rejected;   // a rejection happened inside the current scipy step
    int nfev;       // as scipy counts: 2 + 12 * attempts
    int status;     // QR_ST_* bits
    int nproj;      // SO(3) re-projections inside RHS evaluations
    int checked;    // the next attempt re-projects failing stage matrices (redo of a speculative attempt)
};

// RungeKutta.__init__: f0 = F(y0) -> K0, then select_initial_step.  (2 RHS evaluations)
template <typename T>
QR_DEV void dop853_begin(const T* x, const T* y, T W3, const Dyn<T>& d, const T Tend, const T rtol, const T atol, T* K0, OdeLane<T>& o)
{
    using N = num<T>;
    o.t = 0; o.rejected = 0; o.nfev = 2; o.status = 0; o.nproj = 0; o.checked = 0;
    int fl = rhs14<T, false, false>(y, W3, d, K0);   // y's R was SO(3)-checked by the caller (observation_wrapper)
    o.nproj += fl & 1; if (fl & 2) o.status |= 4;
    T s0 = 0, s1 = 0;   // sums of (y/sc)^2 and (f0/sc)^2 over all 18 components
    T isc[14], iscx[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        iscx[i] = N::recip(N::fma(N::abs(x[i]), rtol, atol));
        T a = x[i] * iscx[i], b = y[i] * iscx[i];   // x' = v
        s0 = N::fma(a, a, s0); s1 = N::fma(b, b, s1);
    }
#pragma unroll
    for (int i = 0; i < 14; ++i) {
        isc[i] = N::recip(N::fma(N::abs(y[i]), rtol, atol));
        T a = y[i] * isc[i], b = K0[i] * isc[i];
        s0 = N::fma(a, a, s0); s1 = N::fma(b, b, s1);
    }
    {
        T iscw3 = N::recip(N::fma(N::abs(W3), rtol, atol));
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
    for (int i = 0; i < 14; ++i) y1[i] = N::fma(h0, K0[i], y[i]);
    T W31 = N::fma(h0, d.w3dot, W3);
    fl = rhs14<T, true>(y1, W31, d, k1);
    o.nproj += fl & 1; if (fl & 2) o.status |= 4;
    T s2 = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {   // f1_x - f0_x = v1 - v0
        T a = (y1[i] - y[i]) * iscx[i];
        s2 = N::fma(a, a, s2);
    }
#pragma unroll
    for (int i = 0; i < 14; ++i) {
        T a = (k1[i] - K0[i]) * isc[i];
        s2 = N::fma(a, a, s2);
    }
    // the W3 component of f1 - f0 is exactly zero
    T d2 = N::sqrt(s2) * inv_sqrt_n / h0;
    T h1;
    if (d1 <= (T)1e-15 && d2 <= (T)1e-15) h1 = N::max((T)1e-6, h0 * (T)1e-3);
    else h1 = N::root8((T)0.01 / N::max(d1, d2));
    T h_abs = (T)100 * h0;
    h_abs = (h1 < h_abs) ? h1 : h_abs;
    h_abs = (Tend < h_abs) ? Tend : h_abs;
    // first _step_impl: h_abs is raised to min_step = 10 ulp(t) if smaller
    const T min_step = (T)10 * N::ulp_up((T)0);
    o.h_abs = (h_abs < min_step) ? min_step : h_abs;
}

// One attempt of _step_impl (rk.py:125-166) for the lane: 11 stages from K0, y_new, error norm,
// accept / reject and step-size update.  Returns true when the lane is finished with the whole interval
// (t reached Tend, or the integrator gave up and keeps the last accepted state).
// ks: this warp's stage storage, lane: lane id.
//
// The function is executed by ALL lanes of the warp (control flow stays warp-uniform, so loop counters and
// tableau loads live in the uniform datapath); `live` says whether this lane really has an attempt to make.
// Lanes without one run on whatever benign state they hold and commit nothing.
template <typename T>
QR_DEV bool dop853_attempt(T* x, T* y, T& W3, const Dyn<T>& d, const T Tend, const T rtol, const T atol, T* K0,
                           OdeLane<T>& o, T* ks, const int lane, const bool live)
{
    using N = num<T>;
    using TB = tab<T>;
    const T min_step = (T)10 * N::ulp_up(o.t);
    const bool too_small = o.h_abs < min_step;   // TOO_SMALL_STEP: keep the last accepted y
    T t_new = o.t + o.h_abs;
    if (t_new - Tend > (T)0) t_new = Tend;
    const T h = t_new - o.t;

    // running sums for x (x' = v): B, E5 and E3 weighted stage velocities
    T xb[3], x5[3], x3[3];
#pragma unroll
#if QR_E3_FROM_B
    for (int i = 0; i < 3; ++i) { xb[i] = TB::B(0) * y[i]; x5[i] = TB::E5(0) * y[i]; x3[i] = (TB::E3(0) - TB::B(0)) * y[i]; }   // x3: E3 - B part only
#else
    for (int i = 0; i < 3; ++i) { xb[i] = TB::B(0) * y[i]; x5[i] = TB::E5(0) * y[i]; x3[i] = TB::E3(0) * y[i]; }
#endif
    int nproj = 0, bad = 0;
    T* const kl = ks + lane * 4;   // this lane's column inside every slot (see ks_load / ks_store)
    unsigned kl_sa = (unsigned)__cvta_generic_to_shared(kl);
    asm volatile("" : "+r"(kl_sa));   // keep it in a register: the compiler otherwise re-derives it per K vector

    // The reference tests every stage matrix against SO(3) before using it (state_decomposition inside EoM) and
    // re-projects it if the test fails -- which, for stage points of an accepted-size step, essentially never
    // happens.  The stages therefore run SPECULATIVELY: the test is evaluated as plain dataflow (nothing waits
    // for it), and if some stage failed it the attempt is thrown away and redone by this lane in `checked` mode,
    // where a failing stage is re-projected exactly like the reference does.  Same results as testing stage by
    // stage, without a data-dependent branch per stage.
    const bool checked = o.checked != 0;
    bool all_ok = true;
#pragma unroll 1
    for (int s = 1; s <= 11; ++s) {
        // ys = y + sum_j (h a_sj) K_j
        T ys[14];
        {
            const T ha0 = h * TB::A(s, 0);
            axpy14_out<T>(ha0, K0, y, ys);
        }
        // couplings A[s][j], j >= 1, from the flattened table (constant bank, uniform datapath)
#if QR_PIPE
        // two K vectors in flight: the loads of the next one are issued before the sums of the current one
        if (QR_PTX && sizeof(T) == 4) {
            int p = TB::Ps(s);
            int n = TB::Ps(s + 1) - p;
            if (n > 0) {
                T ka[14], kb[14];
                TabEntry<T> te = TB::P(p);
                T ca = h * te.c, cb;
                ks_load_lane_sa(kl_sa + (unsigned)te.off, (unsigned)lane * 8u, ka);
#pragma unroll 1
                for (;;) {
                    if (n == 1) { axpy14<T>(ca, ka, ys); break; }
                    te = TB::P(p + 1); cb = h * te.c;
                    ks_load_lane_sa(kl_sa + (unsigned)te.off, (unsigned)lane * 8u, kb);
                    axpy14<T>(ca, ka, ys);
                    if (n == 2) { axpy14<T>(cb, kb, ys); break; }
                    te = TB::P(p + 2); ca = h * te.c;
                    ks_load_lane_sa(kl_sa + (unsigned)te.off, (unsigned)lane * 8u, ka);
                    axpy14<T>(cb, kb, ys);
                    p += 2; n -= 2;
                }
            }
        } else
#endif
        {
            const int p1 = TB::Ps(s + 1);
#pragma unroll 1
            for (int p = TB::Ps(s); p < p1; ++p) {
                const TabEntry<T> te = TB::P(p);
                const T c = h * te.c;
                T k[14];
                if (QR_PTX && sizeof(T) == 4) ks_load_lane_sa(kl_sa + (unsigned)te.off, (unsigned)lane * 8u, k);
                else ks_load_lane<T>(reinterpret_cast<const T*>(reinterpret_cast<const char*>(kl) + te.off), lane, k);
                axpy14<T>(c, k, ys);
            }
        }
        const T W3s = N::fma(h * TB::C(s), d.w3dot, W3);
#if QR_E3_FROM_B
        if (s >= 5) {   // (warp-uniform)
            const T bs = TB::B(s), e5s = TB::E5(s), d3s = TB::E3(s) - bs;
#pragma unroll
            for (int i = 0; i < 3; ++i) { xb[i] = N::fma(bs, ys[i], xb[i]); x5[i] = N::fma(e5s, ys[i], x5[i]); }
            if (d3s != (T)0) {
#pragma unroll
                for (int i = 0; i < 3; ++i) x3[i] = N::fma(d3s, ys[i], x3[i]);
            }
        }
#else
        {
            const T bs = TB::B(s), e5s = TB::E5(s), e3s = TB::E3(s);   // zero for s < 5: no branch
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                xb[i] = N::fma(bs, ys[i], xb[i]); x5[i] = N::fma(e5s, ys[i], x5[i]); x3[i] = N::fma(e3s, ys[i], x3[i]);
            }
        }
#endif
        const bool ok = so3_ok<T>(ys + 3);
        all_ok = all_ok && ok;
        if (checked && !ok) {   // slow path of a redone attempt: the reference's re-projection
            T tmp[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) tmp[i] = ys[3 + i];
            const int pb = project_so3<T>(tmp);
#pragma unroll
            for (int i = 0; i < 9; ++i) ys[3 + i] = tmp[i];
            nproj += 1; bad |= pb ? 2 : 0;
        }
        T kn[14];
        rhs14<T, false, false>(ys, W3s, d, kn);
        ks_store_lane<T>(kl + k_slot(s) * QR_SLOT_ELEMS, lane, kn);
    }

    // y_new = y + h * sum_s B_s K_s ; error estimates (rk.py:683-691)
    T sb[14], s5[14], s3[14];
    {
        const T b0 = TB::B(0), e50 = TB::E5(0), e30 = TB::E3(0);
#pragma unroll
#if QR_E3_FROM_B
        for (int i = 0; i < 14; ++i) { sb[i] = b0 * K0[i]; s5[i] = e50 * K0[i]; s3[i] = (e30 - b0) * K0[i]; }   // s3: E3 - B part only
#else
        for (int i = 0; i < 14; ++i) { sb[i] = b0 * K0[i]; s5[i] = e50 * K0[i]; s3[i] = e30 * K0[i]; }
#endif
    }
#pragma unroll 1
    for (int j = 5; j <= 11; ++j) {
        const T bj = TB::B(j), e5j = TB::E5(j), e3j = TB::E3(j);
        T k[14];
        ks_load_lane<T>(kl + (j == 11 ? 0 : j - 3) * QR_SLOT_ELEMS, lane, k);
#if QR_E3_FROM_B
        axpy14<T>(bj, k, sb); axpy14<T>(e5j, k, s5);
        if (e3j - bj != (T)0) axpy14<T>(e3j - bj, k, s3);   // j = 8, 11 (warp-uniform)
#else
        axpy14x3<T>(bj, e5j, e3j, k, sb, s5, s3);
#endif
    }
    T e5n = 0, e3n = 0;
    T xnew[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        xnew[i] = N::fma(h, xb[i], x[i]);
        T isc = N::recip(N::fma(N::max(N::abs(x[i]), N::abs(xnew[i])), rtol, atol));
#if QR_E3_FROM_B
        T e5 = x5[i] * isc, e3 = (x3[i] + xb[i]) * isc;
#else
        T e5 = x5[i] * isc, e3 = x3[i] * isc;
#endif
        e5n = N::fma(e5, e5, e5n); e3n = N::fma(e3, e3, e3n);
    }
#pragma unroll
    for (int i = 0; i < 14; ++i) {
#if QR_E3_FROM_B
        s3[i] = s3[i] + sb[i];            // E3 sum = B sum + correction
#endif
        sb[i] = N::fma(h, sb[i], y[i]);   // y_new
        T isc = N::recip(N::fma(N::max(N::abs(y[i]), N::abs(sb[i])), rtol, atol));
        T e5 = s5[i] * isc, e3 = s3[i] * isc;
        e5n = N::fma(e5, e5, e5n); e3n = N::fma(e3, e3, e3n);
    }
    T err;
    if (e5n == (T)0 && e3n == (T)0) err = 0;
    else err = N::abs(h) * e5n * N::rsqrt((e5n + (T)0.01 * e3n) * (T)18);

    if (!live) return false;
    if (!all_ok && !checked) { o.checked = 1; return false; }   // redo this attempt with per-stage re-projection
    o.checked = 0;
    if (too_small) { o.status |= 2; return true; }
    o.h_abs = N::abs(h);
    o.nfev += 12;
    o.nproj += nproj; if (bad) o.status |= 4;
    if (err < (T)1) {   // accept
#pragma unroll
        for (int i = 0; i < 3; ++i) x[i] = xnew[i];
#pragma unroll
        for (int i = 0; i < 14; ++i) y[i] = sb[i];
        W3 = N::fma(h, d.w3dot, W3);
        o.t = t_new;
        if (!(o.t < Tend)) return true;
        T factor = (err == (T)0) ? (T)10 : N::min((T)10, (T)0.9 * N::inv_root8(err));
        if (o.rejected) factor = N::min((T)1, factor);
        o.h_abs *= factor;
        int fl = rhs14<T>(y, W3, d, K0);   // f_new becomes the next step's first stage
        o.nproj += fl & 1; if (fl & 2) o.status |= 4;
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

}  // namespace qr
