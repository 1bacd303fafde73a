// qr_env.cuh -- per-env pieces around the integrator: action maps, observation + integral errors,
// reward / termination, Philox resets and the mode-0 goal generator.  One env per thread, registers only.
//
// Reference semantics restated (paths relative to the gym-rotor tree):
//   action wrappers .............. gym_rotor/envs/quad.py:225-242, wrappers/coupled_yaw_wrapper.py:44-53,
//                                  wrappers/decoupled_yaw_wrapper.py:49-59, 68-73
//   get_norm_error_state ......... gym_rotor/envs/quad.py:421-466; quad_utils.py:20-26, 38-63; wrapper_utils.py:3-28
//   reward / done ................ coupled_yaw_wrapper.py:78-110, decoupled_yaw_wrapper.py:92-140, quad.py:150-166, 274-318
//   reset ........................ quad.py:171-222, 338-404; coupled_yaw_wrapper.py:27-41
//   goal generator, mode 0 ....... utils/trajectory_generator.py:113-173, 196-221
#pragma once
#include "qr_dop853.cuh"

namespace qr {

// ---- arithmetic that must NOT be contracted into FMAs (numpy evaluates these op by op) --------------
template <typename T> struct rn;
template <> struct rn<float> {
    // float32 mode is not the bit-parity mode (bar: 1e-5 of the reference): plain operators, so that the compiler contracts
    // multiply-add chains into FMAs; the float64 instantiation below keeps numpy's op-by-op roundings
    static QR_DEV float mul(float a, float b) { return a * b; }
    static QR_DEV float add(float a, float b) { return a + b; }
    static QR_DEV float sub(float a, float b) { return a - b; }
    static QR_DEV float div(float a, float b) { return __fdiv_rn(a, b); }
    // division by a launch constant: multiply by its host-computed reciprocal (float32 mode is not the
    // bit-parity mode; the float64 instantiation below keeps numpy's exact quotient)
    static QR_DEV float divc(float a, float b, float inv_b) { (void)b; return __fmul_rn(a, inv_b); }
    // a b + c d: WHICH product is fused is spelled out -- left to the compiler it depends on the use counts of the products,
    // i.e. on the kernel instantiation, and multi-step and single-step launches must agree bit for bit
    static QR_DEV float dot2(float a, float b, float c, float d) { return fmaf(a, b, __fmul_rn(c, d)); }
    static QR_DEV float madd(float a, float b, float c) { return fmaf(a, b, c); }   // a b + c, likewise
};
template <> struct rn<double> {
    static QR_DEV double mul(double a, double b) { return __dmul_rn(a, b); }
    static QR_DEV double add(double a, double b) { return __dadd_rn(a, b); }
    static QR_DEV double sub(double a, double b) { return __dsub_rn(a, b); }
    static QR_DEV double div(double a, double b) { return __ddiv_rn(a, b); }
    static QR_DEV double divc(double a, double b, double inv_b) { (void)inv_b; return __ddiv_rn(a, b); }
    static QR_DEV double dot2(double a, double b, double c, double d) { return __dadd_rn(__dmul_rn(a, b), __dmul_rn(c, d)); }
    static QR_DEV double madd(double a, double b, double c) { return __dadd_rn(__dmul_rn(a, b), c); }
};

// ---- kernel arguments --------------------------------------------------------------------------------
template <typename T> struct EnvConst {
    T dt, g, rtol, atol, x_lim, v_lim, W_lim, eIx_lim, eIb1_lim, sat, alpha, beta, min_force, euler_lim;
    T inv_x_lim, inv_v_lim, inv_W_lim, inv_eIx_lim, inv_eIb1_lim;   // host-computed reciprocals (float32 mode only)
    float nCx, nCIx, nCv, nCb1, nCIb1, nCW, nCw12, nCW3;   // negated reward coefficients, as float32 (numpy weak scalars)
    double Cx, Cv, Cb1, CW;                                // base Quad-v0 reward is evaluated in float64
    double rmin, rmin1, rmin2, slope, slope1, slope2, udm;
    int mode, integrator, autoreset, goal_mode, env_type, max_episode_steps, diagnostics, round_returns;
};

// ---- Philox4x32-10 (Salmon et al., SC'11) ---------------------------------------------------------------
struct Philox {
    uint32_t k0, k1;
    // rolled on purpose: resets and in-kernel action draws sit in divergent, once-per-episode paths where
    // code size (instruction cache) matters more than the ten-round latency
    QR_DEV void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t* out) const { run<false>(c0, c1, c2, c3, out); }
    // UNROLLED: the per-step action draw of the synthetic workload sits in the hot loop (3.4 % of its instructions when rolled)
    template <bool UNROLLED> QR_DEV void run(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t* out) const
    {
        uint32_t a = k0, b = k1;
#pragma unroll(UNROLLED ? 10 : 1)
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            uint32_t n0 = hi1 ^ c1 ^ a, n2 = hi0 ^ c3 ^ b;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};
// u32 -> uniform in (0,1): (k + 0.5) * 2^-32, exact in double
QR_DEV double u01(uint32_t k) { return ((double)k + 0.5) * 2.3283064365386963e-10; }
#define QR_DOMAIN_RESET 0u
#define QR_DOMAIN_ACTION 0x40000000u

// ---- per-env register state -----------------------------------------------------------------------------
template <typename T> struct EnvRegs {
    T x[3];     // position
    T y[14];    // v(3) | R column-major (9) | W1 W2
    T W3;
    T I[8];     // eIx.error(3) eIx.integrand(3) eIb1.error eIb1.integrand
    T m, d, J1, J3, c_tf, c_tw;
    T goal[12]; // xd vd b1d Wd
};

// ---- env.reset(env_type) with counter-based randomness ---------------------------------------------------
// Uniform order: m d J1 J3 c_tf c_tw | yaw | origin-spawn coin | x(3) | v(3) | W(3) | roll pitch | theta(goal)
// The arithmetic runs in T: float64 mode reproduces the oracle's reset to rounding; float32 mode keeps the
// (once per episode, divergent) reset path cheap.
template <typename T> QR_DEV T u01t(uint32_t k)
{
    if (sizeof(T) == 8) return (T)(((double)k + 0.5) * 2.3283064365386963e-10);
    return (T)(((float)(k >> 8) + 0.5f) * 5.9604644775390625e-8f);   // 24 bits: exact in float32, never 0 or 1
}
template <typename T> QR_DEV void sincos_t(T a, T* s, T* c);
template <> QR_DEV void sincos_t<double>(double a, double* s, double* c) { sincos(a, s, c); }
// float32 mode: arguments are bounded by pi, so the range-reduction slow path of sincosf is dead weight
// (hundreds of instructions in a divergent path); __sincosf is accurate to ~4e-7 there
template <> QR_DEV void sincos_t<float>(float a, float* s, float* c) { __sincosf(a, s, c); }

template <typename T>
QR_DEV void reset_env(EnvRegs<T>& e, const Philox& ph, uint64_t gid, uint32_t episode, int env_type, double udm, T* theta_out)
{
    uint32_t r[20];
#pragma unroll
    for (int j = 0; j < 5; ++j) ph((uint32_t)gid, (uint32_t)(gid >> 32), episode, QR_DOMAIN_RESET + j, r + 4 * j);
    const T nom[6] = {(T)2.15, (T)0.23, (T)0.022, (T)0.035, (T)0.0135, (T)2.2};  // quad.py:28-32
    T p[6] = {nom[0], nom[1], nom[2], nom[3], nom[4], nom[5]};
    if (env_type == 0) {
        const T rr = (T)(udm / 100.0);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            T w = (j == 5) ? nom[j] * (rr / (T)2) : nom[j] * rr;  // c_tw: half the range (quad.py:375)
            T lo = nom[j] - w, hi = nom[j] + w;
            p[j] = lo + (hi - lo) * u01t<T>(r[j]);
        }
    }
    const T PI = (T)3.14159265358979323846;
    T yaw = -PI + (PI - (-PI)) * u01t<T>(r[6]);
    T ix, iv, iR, iW;
    if (env_type == 0) {
        if (u01t<T>(r[7]) < (T)0.2) { ix = 0; iv = 0; iR = 0; iW = 0; }                              // quad.py:342-346
        else { ix = (T)0.6; iv = (T)4.0 * (T)0.5; iR = (T)50 * (PI / (T)180); iW = (T)2 * PI * (T)0.5; }  // quad.py:347-351
    } else { ix = (T)0.4; iv = 0; iR = 0; iW = 0; }                                                 // quad.py:352-356
    T xs[3], vs[3], Ws[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        xs[i] = -ix + (ix - (-ix)) * u01t<T>(r[8 + i]);
        vs[i] = -iv + (iv - (-iv)) * u01t<T>(r[11 + i]);
        Ws[i] = -iW + (iW - (-iW)) * u01t<T>(r[14 + i]);
    }
    T roll = -iR + (iR - (-iR)) * u01t<T>(r[17]), pitch = -iR + (iR - (-iR)) * u01t<T>(r[18]);
    T sr, cr, sp, cp, sy, cy;
    sincos_t<T>(roll, &sr, &cr); sincos_t<T>(pitch, &sp, &cp); sincos_t<T>(yaw, &sy, &cy);
    // R = Rz(yaw) Ry(pitch) Rx(roll)   (Rotation.from_euler('xyz'), quad.py:199)
    T R[9] = {cy * cp, sy * cp, -sp,
              cy * sp * sr - sy * cr, sy * sp * sr + cy * cr, cp * sr,
              cy * sp * cr + sy * sr, sy * sp * cr - cy * sr, cp * cr};
#pragma unroll
    for (int i = 0; i < 3; ++i) { e.x[i] = xs[i]; e.y[i] = vs[i]; }
#pragma unroll
    for (int i = 0; i < 9; ++i) e.y[3 + i] = R[i];
    e.y[12] = Ws[0]; e.y[13] = Ws[1]; e.W3 = Ws[2];
#pragma unroll
    for (int i = 0; i < 8; ++i) e.I[i] = (T)0;
    e.m = p[0]; e.d = p[1]; e.J1 = p[2]; e.J3 = p[3]; e.c_tf = p[4]; e.c_tw = p[5];
    *theta_out = ((T)-25 + (T)50 * u01t<T>(r[19])) * (PI / (T)180);   // trajectory_generator.py:144
}

// ---- goal generator, mode 0 ---------------------------------------------------------------------------------
// Wd = [0, 0, b3 . (b1c x b1c_dot)] with b1d_dot = 0 (trajectory_generator.py:165-172)
template <typename T> QR_DEV void traj_wd(const T* R, const T* W, const T* b1d, T* Wd)
{
    const T* b3 = R + 6;
    // b3_dot = R hat(W) e3
    T b3d[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) b3d[i] = R[i] * W[1] - R[i + 3] * W[0];
    T dp = b1d[0] * b3[0] + b1d[1] * b3[1] + b1d[2] * b3[2];
    T dq = b1d[0] * b3d[0] + b1d[1] * b3d[1] + b1d[2] * b3d[2];
    T b1c[3], b1cd[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        b1c[i] = b1d[i] - dp * b3[i];
        b1cd[i] = (T)0 - (((T)0 * b3[i] + dq * b3[i]) + dp * b3d[i]);
    }
    T oc0 = b1c[1] * b1cd[2] - b1c[2] * b1cd[1];
    T oc1 = b1c[2] * b1cd[0] - b1c[0] * b1cd[2];
    T oc2 = b1c[0] * b1cd[1] - b1c[1] * b1cd[0];
    Wd[0] = 0; Wd[1] = 0; Wd[2] = b3[0] * oc0 + b3[1] * oc1 + b3[2] * oc2;
}

// mark_traj_start + first get_desired(mode 0) on the float32-cast reset state (main.py:226-229):
// xd = vd = 0, b1d = Rz(theta) [cos psi, sin psi, 0], psi = heading of b1; Wd from that same f32 state.
template <typename T> QR_DEV void init_goal_mode0(EnvRegs<T>& e, T theta)
{
    T R[9], W[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = (T)(float)e.y[3 + i];
    W[0] = (T)(float)e.y[12]; W[1] = (T)(float)e.y[13]; W[2] = (T)(float)e.W3;
    ensure_so3<T>(R);
    T psi = num<T>::atan2(R[1], R[0]);
    T sps, cps, sth, cth;
    sincos_t<T>(psi, &sps, &cps); sincos_t<T>(theta, &sth, &cth);
    T b1d[3] = {cth * cps - sth * sps, sth * cps + cth * sps, (T)0};
    T Wd[3];
    traj_wd<T>(R, W, b1d, Wd);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        e.goal[i] = 0; e.goal[3 + i] = 0; e.goal[6 + i] = b1d[i]; e.goal[9 + i] = Wd[i];
    }
}

// ---- get_norm_error_state ------------------------------------------------------------------------------------
// Writes the float32 observation (COUPLED 23, DECOUPLED 15+3) and advances the integral errors once.
template <typename T> QR_DEV int norm_error_state(EnvRegs<T>& e, const EnvConst<T>& c, float* o, const int mode)
{
    using N = num<T>;
    using A = rn<T>;
    T R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = e.y[3 + i];
    int fl = ensure_so3<T>(R);   // state_normalization (quad_utils.py:20-26)
    const T W[3] = {e.y[12], e.y[13], e.W3};
    T ex[3], ev[3], eW[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ex[i] = A::sub(A::divc(e.x[i], c.x_lim, c.inv_x_lim), A::divc(e.goal[i], c.x_lim, c.inv_x_lim));
        ev[i] = A::sub(A::divc(e.y[i], c.v_lim, c.inv_v_lim), A::divc(e.goal[3 + i], c.v_lim, c.inv_v_lim));
        eW[i] = A::sub(A::divc(W[i], c.W_lim, c.inv_W_lim), A::divc(e.goal[9 + i], c.W_lim, c.inv_W_lim));
    }
    const T* b1 = R; const T* b2 = R + 3; const T* b3 = R + 6;
    const T* b1d = e.goal + 6;
    // numpy's 3-vector dot is an FMA chain (OpenBLAS ddot)
    T d3 = N::fma(b1d[2], b3[2], N::fma(b1d[1], b3[1], A::mul(b1d[0], b3[0])));
    T b1c[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) b1c[i] = A::madd(-d3, b3[i], b1d[i]);
    T dn = N::fma(b1c[2], b2[2], N::fma(b1c[1], b2[1], A::mul(b1c[0], b2[0])));
    T dd = N::fma(b1c[2], b1[2], N::fma(b1c[1], b1[1], A::mul(b1c[0], b1[0])));
    const T PI = (T)3.14159265358979323846;
    T eb1n = A::divc(N::atan2(-dn, dd), PI, (T)0.31830988618379067154);
    T eIxn[3], eIb1n;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        T gnew = A::dot2(-c.alpha, e.I[i], ex[i], c.x_lim);
        e.I[i] = A::madd(A::mul(A::add(e.I[3 + i], gnew), c.dt), (T)0.5, e.I[i]);
        e.I[3 + i] = gnew;
        T q = A::divc(e.I[i], c.eIx_lim, c.inv_eIx_lim);
        eIxn[i] = q < -c.sat ? -c.sat : (q > c.sat ? c.sat : q);
    }
    {
        T gnew = A::dot2(-c.beta, e.I[6], eb1n, PI);
        e.I[6] = A::madd(A::mul(A::add(e.I[7], gnew), c.dt), (T)0.5, e.I[6]);
        e.I[7] = gnew;
        T q = A::divc(e.I[6], c.eIb1_lim, c.inv_eIb1_lim);
        eIb1n = q < -c.sat ? -c.sat : (q > c.sat ? c.sat : q);
    }
    if (mode == 2) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            o[i] = (float)ex[i]; o[3 + i] = (float)eIxn[i]; o[6 + i] = (float)ev[i]; o[9 + i] = own_reg((float)b3[i]);
            o[12 + i] = (float)A::dot2(eW[0], b1[i], eW[1], b2[i]);
        }
        o[15] = (float)eb1n; o[16] = (float)eIb1n; o[17] = (float)eW[2];
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            o[i] = (float)ex[i]; o[3 + i] = (float)eIxn[i]; o[6 + i] = (float)ev[i]; o[20 + i] = (float)eW[i];
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) o[9 + i] = own_reg((float)R[i]);
        o[18] = (float)eb1n; o[19] = (float)eIb1n;
    }
    return fl;
}

// ---- reward and termination on the float32 observation --------------------------------------------------------
// numpy: norm(v,2) of a float32 3-vector = sqrtf(float(double-accumulated float32 products)) (OpenBLAS sdot);
// `**2` restated as a correctly rounded square (numpy's powf(x,2) differs by 1 ulp in ~0.07 % of cases).
// EXACT = float64 (parity) mode.  In float32 mode the observation already differs from numpy's at rounding
// level, so the float64 accumulation (three conversions through the FP64 pipe) is replaced by a float32 sum.
template <bool EXACT> QR_DEV float norm2sq_f32(const float* v)
{
    float n;
    if (EXACT) {
        double s = (double)__fmul_rn(v[0], v[0]) + (double)__fmul_rn(v[1], v[1]) + (double)__fmul_rn(v[2], v[2]);
        n = __fsqrt_rn((float)s);
    } else {
        // float32 mode: norm(v)**2 without the round trip through the square root (within 1 ulp of numpy's value)
        return fmaf(v[2], v[2], fmaf(v[1], v[1], v[0] * v[0]));
    }
    return __fmul_rn(n, n);
}
QR_DEV double interp01(double r, double rmin, double slope)
{
    // np.interp(r, [rmin, 0], [0, 1]); slope = 1 / (0 - rmin), computed once on the host
    if (r != r) return r;
    if (r <= rmin) return 0.0;
    if (r >= 0.0) return 1.0;
    return __dadd_rn(__dmul_rn(slope, __dsub_rn(r, rmin)), 0.0);
}

// float32 mode: the same interpolation in float32 (no trip through the FP64 pipe)
// (two selects, no branch; a NaN r fails both comparisons and comes back through the product)
QR_DEV float interp01f(float r, float rmin, float slope)
{
    float v = slope * (r - rmin);
    v = (r >= 0.0f) ? 1.0f : v;
    return (r <= rmin) ? 0.0f : v;
}
template <typename T> QR_DEV T interp01t(float r, double rmin, double slope)
{
    if (sizeof(T) == 8) return (T)interp01((double)r, rmin, slope);
    return (T)interp01f(r, (float)rmin, (float)slope);
}

// RT: the type the reward is handed on in -- T in the step kernel (float32 mode: no trip through the FP64 pipe), double elsewhere
template <typename T, typename RT> QR_DEV void reward_done(const EnvConst<T>& c, const float* o, RT* rew, int* dn, const int mode)
{
    dn[0] = 0; dn[1] = 0;
    if (mode == 1) {
        float rx = __fmul_rn(c.nCx, norm2sq_f32<sizeof(T) == 8>(o)), rix = __fmul_rn(c.nCIx, norm2sq_f32<sizeof(T) == 8>(o + 3));
        float rv = __fmul_rn(c.nCv, norm2sq_f32<sizeof(T) == 8>(o + 6)), rw = __fmul_rn(c.nCW, norm2sq_f32<sizeof(T) == 8>(o + 20));
        float a18 = fabsf(o[18]), a19 = fabsf(o[19]);
        float rb = __fmul_rn(c.nCb1, a18), rib = __fmul_rn(c.nCIb1, __fmul_rn(a19, a19));
        float r = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(rx, rix), rv), rb), rib), rw);
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if (fabsf(o[i]) >= 1.0f || fabsf(o[6 + i]) >= 1.0f || fabsf(o[20 + i]) >= 1.0f) dn[0] = 1;
        rew[0] = dn[0] ? (RT)-1 : (RT)interp01t<T>(r, c.rmin, c.slope);
        rew[1] = 0;
    } else {
        float rx = __fmul_rn(c.nCx, norm2sq_f32<sizeof(T) == 8>(o)), rix = __fmul_rn(c.nCIx, norm2sq_f32<sizeof(T) == 8>(o + 3));
        float rv = __fmul_rn(c.nCv, norm2sq_f32<sizeof(T) == 8>(o + 6)), rw = __fmul_rn(c.nCw12, norm2sq_f32<sizeof(T) == 8>(o + 12));
        float r1 = __fadd_rn(__fadd_rn(__fadd_rn(rx, rix), rv), rw);
        float a0 = fabsf(o[15]), a1 = fabsf(o[16]), a2 = fabsf(o[17]);
        float r2 = __fadd_rn(__fadd_rn(__fmul_rn(c.nCb1, a0), __fmul_rn(c.nCIb1, __fmul_rn(a1, a1))),
                             __fmul_rn(c.nCW3, __fmul_rn(a2, a2)));
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if (fabsf(o[i]) >= 1.0f || fabsf(o[6 + i]) >= 1.0f || fabsf(o[12 + i]) >= 1.0f) dn[0] = 1;
        if (a2 >= 1.0f) dn[1] = 1;
        rew[0] = dn[0] ? (RT)-1 : (RT)interp01t<T>(r1, c.rmin1, c.slope1);
        rew[1] = dn[1] ? (RT)-1 : (RT)interp01t<T>(r2, c.rmin2, c.slope2);
    }
}

// Base Quad-v0 reward / done on the float64 next state (quad.py:274-318)
template <typename T> QR_DEV void reward_done_quad(const EnvRegs<T>& e, const EnvConst<T>& c, double* rew, int* dn)
{
    double R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = (double)e.y[3 + i];
    ensure_so3<double>(R);
    const double W[3] = {(double)e.y[12], (double)e.y[13], (double)e.W3};
    double eX2 = 0, eV2 = 0, W2 = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double a = (double)e.x[i] - (double)e.goal[i], b = (double)e.y[i] - (double)e.goal[3 + i];
        eX2 = fma(a, a, eX2); eV2 = fma(b, b, eV2); W2 = fma(W[i], W[i], W2);
    }
    double nx = sqrt(eX2), nv = sqrt(eV2), nw = sqrt(W2);
    double th = atan2(R[1], R[0]);
    double cb[3] = {cos(th), sin(th), 0.0};
    double b1d[3] = {(double)e.goal[6], (double)e.goal[7], (double)e.goal[8]};
    double nd = sqrt(fma(b1d[2], b1d[2], fma(b1d[1], b1d[1], b1d[0] * b1d[0])));
    double nc = sqrt(fma(cb[2], cb[2], fma(cb[1], cb[1], cb[0] * cb[0])));
    double du[3] = {b1d[0] / nd, b1d[1] / nd, b1d[2] / nd}, cu[3] = {cb[0] / nc, cb[1] / nc, cb[2] / nc};
    double dp = fma(du[2], cu[2], fma(du[1], cu[1], du[0] * cu[0]));
    dp = dp < -1 ? -1 : (dp > 1 ? 1 : dp);
    double ang = acos(dp);
    if (du[0] * cu[1] - du[1] * cu[0] < 0) ang = -ang;
    double eb1 = ang / 3.14159265358979323846;
    double r = 0.0 + (((-c.Cx * (nx * nx) + -c.Cb1 * fabs(eb1)) + -c.Cv * (nv * nv)) + -c.CW * (nw * nw));
    const double R2D = 180.0 / 3.14159265358979323846;
    double roll = atan2(R[5], R[8]) * R2D;
    double pitch = atan2(-R[2], sqrt(R[0] * R[0] + R[1] * R[1])) * R2D;
    int d = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (fabs((double)e.x[i]) >= (double)c.x_lim) d = 1;
        if (fabs((double)e.y[i]) >= (double)c.v_lim) d = 1;
        if (fabs(W[i]) >= (double)c.W_lim) d = 1;
    }
    if (fabs(roll) >= (double)c.euler_lim || fabs(pitch) >= (double)c.euler_lim) d = 1;
    dn[0] = d; dn[1] = 0;
    rew[0] = d ? -1.0 : interp01(r, c.rmin, c.slope);
    rew[1] = 0.0;
}

// ---- action wrappers -> thrust f and moment M -----------------------------------------------------------------
// a[]: normalised action (already converted to T); act_f32: the caller's array was float32, in which case
// numpy evaluates the thrust scaling in float32 (python-float * np.float32 -> float32, NEP 50).
template <typename T>
QR_DEV void action_to_fM(const EnvRegs<T>& e, const EnvConst<T>& c, const T* a, bool act_f32, T& f, T* M, const int mode)
{
    using A = rn<T>;
    using N = num<T>;
    const T hover = A::mul(A::mul(e.m, c.g), (T)0.25);        // quad.py:390  (x/4 == x*0.25 exactly)
    const T maxf = A::mul(e.c_tw, hover);                     // quad.py:392
    const T avrg = A::mul(A::add(c.min_force, maxf), (T)0.5); // quad.py:403  (x/2 == x*0.5 exactly)
    const T scale = A::sub(maxf, avrg);                       // quad.py:404
    if (mode == 0) {
        T Tm[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (act_f32 && sizeof(T) == 8) {
                float tv = __fadd_rn(__fmul_rn((float)scale, (float)a[i]), (float)avrg);
                tv = tv < (float)c.min_force ? (float)c.min_force : (tv > (float)maxf ? (float)maxf : tv);
                Tm[i] = (T)tv;
            } else {
                T tv = A::madd(scale, a[i], avrg);
                Tm[i] = tv < c.min_force ? c.min_force : (tv > maxf ? maxf : tv);
            }
        }
        f = A::add(A::add(A::add(Tm[0], Tm[1]), Tm[2]), Tm[3]);   // forces_to_fM @ T, quad.py:396-401,238
        M[0] = A::dot2(-e.d, Tm[1], e.d, Tm[3]);
        M[1] = A::dot2(e.d, Tm[0], -e.d, Tm[2]);
        M[2] = A::madd(e.c_tf, Tm[3], A::madd(-e.c_tf, Tm[2], A::dot2(-e.c_tf, Tm[0], e.c_tf, Tm[1])));
        return;
    }
    if (act_f32 && sizeof(T) == 8) {
        float fv = __fmul_rn(4.0f, __fadd_rn(__fmul_rn((float)scale, (float)a[0]), (float)avrg));
        float lo = (float)A::mul((T)4, c.min_force), hi = (float)A::mul((T)4, maxf);
        fv = fv < lo ? lo : (fv > hi ? hi : fv);
        f = (T)fv;
    } else {
        T fv = A::mul((T)4, A::madd(scale, a[0], avrg));
        T lo = A::mul((T)4, c.min_force), hi = A::mul((T)4, maxf);
        f = fv < lo ? lo : (fv > hi ? hi : fv);
    }
    if (mode == 1) {
        M[0] = a[1]; M[1] = a[2]; M[2] = a[3];
    } else {
        // decoupled:68-73 on the pre-step (already SO(3)-checked) R and W
        const T* b1 = e.y + 3; const T* b2 = e.y + 6;
        T t1 = N::fma(b1[2], a[3], N::fma(b1[1], a[2], A::mul(b1[0], a[1])));
        T t2 = N::fma(b2[2], a[3], N::fma(b2[1], a[2], A::mul(b2[0], a[1])));
        M[0] = A::madd(A::mul(e.J3, e.W3), e.y[13], t1);
        M[1] = A::madd(-A::mul(e.J3, e.W3), e.y[12], t2);
        M[2] = a[4];
    }
}

}  // namespace qr
