// qr_traj.cuh -- on-device goal generation for trajectory modes 1 (hover), 2 (take-off), 3 (land), 4 (stay),
// 5 (circle), >= 6 (figure eight) and the manual-mode fallback after a trajectory completes.
//
// Restates utils/trajectory_generator.py:113-173 (get_desired / calculate_desired / Wd), 176-229 (mark_traj_start,
// clock), 232-277 (manual, hovering), 280-357 (takeoff, waypoint_reached, land, stay), 359-412 (circle), 415-505
// (eight_shaped_curve).  Mode 0 lives in qr_env.cuh
// (it is evaluated inside the step kernel).  Per-env trajectory state ts[12]:
//   0 t | 1 flags (bit0 trajectory_started, bit1 manual_mode, bit2 manual_mode_init) | 2..4 x_init / centre |
//   5 theta_init | 6 w_b1d | 7 smooth_term | 8 t_traj | 9,10 b1d_dot x,y | 11 unused
// numpy detail reproduced: under main.py's protocol xd / vd are float32 arrays for the whole trajectory (copies of
// the float32 reset state, :212-215 with main.py:226-228), so every assignment to them rounds to float32, and the
// expressions that only mix the float32 centre with python floats are evaluated in float32 (NEP 50).
#pragma once
#include "qr_env.cuh"

namespace qr {

template <typename T> QR_DEV T f32r(T v) { return (T)(float)v; }
template <typename T> QR_DEV T exp_t(T a);
template <> QR_DEV double exp_t<double>(double a) { return exp(a); }
template <> QR_DEV float exp_t<float>(float a) { return expf(a); }
template <typename T> QR_DEV void sincos_acc(T a, T* s, T* c);
template <> QR_DEV void sincos_acc<double>(double a, double* s, double* c) { sincos(a, s, c); }
// float32 mode: the arguments (w t + theta0) stay within a few tens of radians, so a two-constant Cody-Waite reduction to
// [-pi, pi] followed by the MUFU pair (abs error ~4e-7) replaces sincosf and its Payne-Hanek slow path
template <> QR_DEV void sincos_acc<float>(float a, float* s, float* c)
{
    const float k = rintf(a * 0.15915494309189535f);
    float r = fmaf(k, -6.2831854820251465f, a);
    r = fmaf(k, 1.7484555e-7f, r);   // 2 pi = 6.2831854820251465 - 1.7484555e-7
    __sincosf(r, s, c);
}

// mark_traj_start: clock and flags to zero, initial position / heading from the state handed in
template <typename T> QR_DEV void traj_start(const T* x, const T* R_so3, T* ts)
{
#pragma unroll
    for (int i = 0; i < 12; ++i) ts[i] = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) ts[2 + i] = x[i];
    ts[5] = num<T>::atan2(R_so3[1], R_so3[0]);
}

// One get_desired(state, mode) call.  R must already be ensure_SO3'd.  u_ttraj / u_w: uniforms for the hover's
// t_traj ~ U(2,5) and w_b1d ~ U(-0.15 pi, 0.15 pi), only read when the trajectory starts.
// ALL_MODES = false leaves out take-off / land / stay (modes 2-4): that instantiation is the one inlined into the step
// kernel's auto reset, whose code -- and with it the register allocation of the whole persistent loop -- stays exactly
// what was profiled (profiles/r01u_*); qr_create therefore refuses autoreset together with those three goal modes.
template <typename T, bool ALL_MODES = true>
QR_DEV void traj_desired(int mode, const T* x, const T* v, const T* R, const T* W, T* ts, T* goal, T u_ttraj, T u_w, T dt)
{
    using N = num<T>;
    const T PI = (T)3.14159265358979323846;
    T* xd = goal; T* vd = goal + 3; T* b1d = goal + 6; T* Wd = goal + 9;
    int flags = (int)ts[1];
    // get_current_b1 = atan2(R[1], R[0]) is only needed when a trajectory (or the manual mode after it) starts: evaluated there
    T sn, cs;
    if (flags & 2) {   // manual(): calculate_desired returns before the Wd block
        if (!(flags & 4)) {
#pragma unroll
            for (int i = 0; i < 3; ++i) { xd[i] = x[i]; vd[i] = v[i]; }
            ts[5] = N::atan2(R[1], R[0]);
            flags |= 4;
        }
        vd[0] = vd[1] = vd[2] = 0;
        sincos_acc<T>(ts[5], &sn, &cs);
        b1d[0] = cs; b1d[1] = sn; b1d[2] = 0;
        ts[1] = (T)flags;
        return;
    }
    if (!(flags & 1)) {   // set_desired_states_to_current + per-mode start
#pragma unroll
        for (int i = 0; i < 3; ++i) { xd[i] = x[i]; vd[i] = v[i]; ts[2 + i] = x[i]; }
        sincos_acc<T>(N::atan2(R[1], R[0]), &sn, &cs);
        b1d[0] = cs; b1d[1] = sn; b1d[2] = 0;
        flags |= 1;
        if (mode == 1) {
            ts[8] = (T)2 + ((T)5 - (T)2) * u_ttraj;
            ts[7] = (T)6.907755278982137 / ts[8];                       // -log(0.001) / t_traj
            ts[6] = (T)-0.15 * PI + ((T)0.15 * PI - ((T)-0.15 * PI)) * u_w;
        } else if (ALL_MODES && mode == 2) {
            // takeoff(): set_desired_states_to_zero (fresh float64 arrays), horizontal position held; t_traj is a
            // float32 expression (python float - np.float32, NEP 50)
#pragma unroll
            for (int i = 0; i < 3; ++i) { xd[i] = 0; vd[i] = 0; }
            xd[0] = x[0]; xd[1] = x[1];
            ts[8] = (T)__fdiv_rn(__fsub_rn(-0.5f, (float)x[2]), -0.05f);   // (takeoff_end_height - z) / takeoff_velocity
        } else if (ALL_MODES && mode == 3) {
            ts[8] = (T)__fdiv_rn(__fsub_rn(-0.25f, (float)x[2]), 1.0f);    // (landing_motor_cutoff_height - z) / landing_velocity
        } else if (ALL_MODES && mode == 4) {
            // stay(): nothing else
        } else if (mode == 5) {
            ts[8] = (T)0.7 / (T)0.4 + (T)2 * (T)2 * PI / (T)0.4;         // radius / v + num_circles * 2 pi / W
        } else {
            ts[8] = (T)3 * (T)9; ts[6] = (T)0.349066;                   // num_of_eights * T ; eight_w_b1d
        }
    }
    if (!ALL_MODES || mode != 4) ts[0] = ts[0] + dt;   // update_current_time() is called by every mode function except stay()
    const T t = ts[0];
    if (ALL_MODES && mode == 2) {
        // x_init is the float32 state handed in at the start: "x_init[2] + v t" is float32 arithmetic and "t < t_traj"
        // a float32 comparison (the python floats are cast, NEP 50)
        if ((float)t < (float)ts[8]) {
            xd[2] = (T)__fadd_rn((float)ts[4], (float)((T)-0.05 * t));
        } else {
            const T d0 = xd[0] - x[0], d1 = xd[1] - x[1], d2 = xd[2] - x[2];
            if (N::sqrt(N::fma(d2, d2, N::fma(d1, d1, d0 * d0))) < (T)0.04) {   // waypoint_reached(xd, x, 0.04)
                xd[2] = (T)-0.5; vd[2] = 0;
                flags |= 2;   // mark_traj_end(True): manual mode from the next call on
            }
        }
    } else if (ALL_MODES && mode == 3) {
        if ((float)t < (float)ts[8]) {
            xd[2] = (T)__fadd_rn((float)ts[4], (float)((T)1 * t));
        } else if (x[2] > (T)-0.25) {
            xd[2] = (T)-0.25; vd[2] = 0;   // mark_traj_end(False): stays in this branch, no manual mode
        } else {
            xd[2] = (T)-0.25; vd[2] = (T)1;
        }
    } else if (ALL_MODES && mode == 4) {
        flags |= 2;   // mark_traj_end(True)
    } else if (mode == 1) {
        const T k = ts[7], w = ts[6], ek = exp_t<T>(-k * t);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            xd[i] = f32r<T>((ts[2 + i] - (T)0) * ek + (T)0);
            vd[i] = f32r<T>(-(ts[2 + i] - (T)0) * k * ek);
        }
        sincos_acc<T>(w * t + ts[5], &sn, &cs);
        b1d[0] = cs; b1d[1] = sn; b1d[2] = 0;
        ts[9] = -w * sn; ts[10] = w * cs;
    } else if (mode == 5) {
        const T r = (T)0.7, lv = (T)0.4, cw = (T)0.4;
        if (t < r / lv) {
            xd[0] = (T)((float)ts[2] + (float)(lv * t));                 // np.float32 + python float: float32 arithmetic
            vd[0] = f32r<T>(lv);
        } else if (t < ts[8]) {
            const T tt = t - r / lv, th = cw * tt;
            sincos_acc<T>(th, &sn, &cs);
            xd[0] = f32r<T>(r * cs + ts[2]);
            vd[0] = f32r<T>(-r * cw * sn);
            xd[1] = f32r<T>(r * sn + ts[3]);
            vd[1] = f32r<T>(r * cw * cs);
            sincos_acc<T>(cw * tt + PI, &sn, &cs);
            b1d[0] = cs; b1d[1] = sn; b1d[2] = 0;
            ts[9] = -cw * sn; ts[10] = cw * cs;
        } else {
            flags |= 2;   // mark_traj_end(True): manual mode from the next call on
        }
    } else {
        const T A1 = (T)1.5, A2 = (T)1.0, Tp = (T)9;
        const T w1 = (T)2 * PI / Tp, w2 = (T)4 * PI / Tp;
        const T kxy = (T)4.605170185988091 / Tp;                         // -log(0.01) / T
        if (t < ts[8]) {
            const T en = exp_t<T>(-kxy * t), ex = (T)1 - en, dex = kxy * en;
            T s1, c1, s2, c2;
            sincos_acc<T>(w1 * t, &s1, &c1); sincos_acc<T>(w2 * t, &s2, &c2);
            xd[0] = f32r<T>(A2 * (s2 * ex) + ts[2]);
            vd[0] = f32r<T>(A2 * ((w2 * c2 * ex) + (s2 * dex)));
            xd[1] = f32r<T>(A1 * (c1 - (T)1) * ex + ts[3]);
            vd[1] = f32r<T>(A1 * ((w1 * -s1 * ex) + (c1 - (T)1) * dex));
            const float za = ((float)ts[4] - (float)(T)-0.6) / 2.0f;    // float32 centre meets python floats only
            xd[2] = (T)__fadd_rn(__fmul_rn(za, (float)((T)1 - c1)), (float)ts[4]);   // numpy: separate float32 multiply and add
            vd[2] = (T)__fmul_rn(__fmul_rn(za, (float)w1), (float)s1);
            const T wt = ts[6] * t * ex + ts[5], dwt = ts[6] * (ex + t * dex);
            sincos_acc<T>(wt, &sn, &cs);
            b1d[0] = cs; b1d[1] = sn; b1d[2] = 0;
            ts[9] = -sn * dwt; ts[10] = cs * dwt;
        } else {
            flags |= 2;
        }
    }
    ts[1] = (T)flags;
    // Wd = [0, 0, b3 . (b1c x b1c_dot)]  (:165-172) with the current b1d_dot
    const T* b3 = R + 6;
    T b3d[3];
    const T bdd[3] = {ts[9], ts[10], (T)0};
#pragma unroll
    for (int i = 0; i < 3; ++i) b3d[i] = R[i] * W[1] - R[i + 3] * W[0];
    const T dp = b1d[0] * b3[0] + b1d[1] * b3[1] + b1d[2] * b3[2];
    const T dq = b1d[0] * b3d[0] + b1d[1] * b3d[1] + b1d[2] * b3d[2];
    const T dr = bdd[0] * b3[0] + bdd[1] * b3[1] + bdd[2] * b3[2];
    T b1c[3], b1cd[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        b1c[i] = b1d[i] - dp * b3[i];
        b1cd[i] = bdd[i] - ((dr * b3[i] + dq * b3[i]) + dp * b3d[i]);
    }
    const T oc0 = b1c[1] * b1cd[2] - b1c[2] * b1cd[1];
    const T oc1 = b1c[2] * b1cd[0] - b1c[0] * b1cd[2];
    const T oc2 = b1c[0] * b1cd[1] - b1c[1] * b1cd[0];
    Wd[0] = 0; Wd[1] = 0; Wd[2] = b3[0] * oc0 + b3[1] * oc1 + b3[2] * oc2;
}

// reference mode numbers for the goal_mode values of the C ABI:
// QR_GOAL_TRAJ_HOVER/CIRCLE/EIGHT/TAKEOFF/LAND/STAY = 2/3/4/5/6/7 -> trajectory_generator modes 1/5/6/2/3/4
template <bool ALL_MODES = true> QR_DEV int traj_ref_mode(int goal_mode)
{
    if (!ALL_MODES) return goal_mode == 2 ? 1 : (goal_mode == 3 ? 5 : 6);
    return goal_mode == 2 ? 1 : (goal_mode == 3 ? 5 : (goal_mode == 4 ? 6 : (goal_mode == 5 ? 2 : (goal_mode == 6 ? 3 : 4))));
}

// trajectory start from the float32-cast state (main.py:226-228): mark_traj_start + the first get_desired
template <typename T, bool ALL_MODES = true>
QR_DEV void traj_restart(int goal_mode, const EnvRegs<T>& e, T* ts, T* goal, T u_ttraj, T u_w, T dt)
{
    T x[3], v[3], R[9], W[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { x[i] = (T)(float)e.x[i]; v[i] = (T)(float)e.y[i]; }
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = (T)(float)e.y[3 + i];
    W[0] = (T)(float)e.y[12]; W[1] = (T)(float)e.y[13]; W[2] = (T)(float)e.W3;
    ensure_so3<T>(R);
    traj_start<T>(x, R, ts);
    goal[6] = 1; goal[7] = 0; goal[8] = 0;
    traj_desired<T, ALL_MODES>(traj_ref_mode<ALL_MODES>(goal_mode), x, v, R, W, ts, goal, u_ttraj, u_w, dt);
}

}  // namespace qr
