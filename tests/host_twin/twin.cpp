// twin.cpp -- TEST INFRASTRUCTURE (see cuda_shim.h).  The per-env device functions of the step kernel, compiled for
// the host in float64 and strung together the way qr::k_step strings them for one lane:
//   A3  ensure_SO3 of the incoming R, goal (mode 0: Wd from the pre-step state), action -> (f, M), RHS constants,
//       dop853_begin
//   B   dop853_attempt until the interval is done (stage storage = a plain array, lane 0)
//   A1  norm_error_state, reward / done
// so that the arithmetic of the kernel can be checked against the reference's golden vectors on a machine without a
// GPU.  The warp-level orchestration (lane refill, stash, auto-reset queue, stores) is NOT covered here: that is what
// the -m gpu tests are for.  The package never loads this file: the product has no CPU path.
#include "cuda_shim.h"
#include "qr_traj.cuh"
#include "../../include/quadrotor_b200.h"

namespace {

using namespace qr;

template <typename T> EnvConst<T> make_const(const qr_config& c)   // same assignments as make_args() in quadrotor_b200.cu
{
    EnvConst<T> e;
    memset(&e, 0, sizeof(e));
    e.dt = (T)c.dt; e.g = (T)c.g; e.rtol = (T)c.rtol; e.atol = (T)c.atol;
    e.x_lim = (T)c.x_lim; e.v_lim = (T)c.v_lim; e.W_lim = (T)c.W_lim; e.eIx_lim = (T)c.eIx_lim; e.eIb1_lim = (T)c.eIb1_lim; e.sat = (T)c.sat_sigma;
    e.alpha = (T)c.alpha; e.beta = (T)c.beta; e.min_force = (T)c.min_force; e.euler_lim = (T)c.euler_lim_deg;
    e.inv_x_lim = (T)(1.0 / c.x_lim); e.inv_v_lim = (T)(1.0 / c.v_lim); e.inv_W_lim = (T)(1.0 / c.W_lim);
    e.inv_eIx_lim = (T)(1.0 / c.eIx_lim); e.inv_eIb1_lim = (T)(1.0 / c.eIb1_lim);
    e.nCx = (float)(-c.Cx); e.nCIx = (float)(-c.CIx); e.nCv = (float)(-c.Cv); e.nCb1 = (float)(-c.Cb1);
    e.nCIb1 = (float)(-c.CIb1); e.nCW = (float)(-c.CW); e.nCw12 = (float)(-c.Cw12); e.nCW3 = (float)(-c.CW3);
    e.Cx = c.Cx; e.Cv = c.Cv; e.Cb1 = c.Cb1; e.CW = c.CW;
    e.rmin = c.reward_min; e.rmin1 = c.reward_min_1; e.rmin2 = c.reward_min_2; e.udm = c.udm_pct;
    e.slope = 1.0 / (0.0 - c.reward_min); e.slope1 = 1.0 / (0.0 - c.reward_min_1); e.slope2 = 1.0 / (0.0 - c.reward_min_2);
    e.mode = c.mode; e.integrator = c.integrator; e.autoreset = c.autoreset; e.goal_mode = c.goal_mode;
    e.env_type = c.env_type; e.max_episode_steps = c.max_episode_steps; e.diagnostics = c.reserved0;
    return e;
}

template <typename T>
int step_one(const qr_config* cfg, const double* state, const double* integ, const double* params, double* goal_io,
             const double* action, int act_is_f32, double* state_out, double* integ_out, float* obs_out, double* reward_out,
             int* done_out, int* nfev_out, int* nproj_out)
{
    const EnvConst<T> c = make_const<T>(*cfg);
    const int MODE = c.mode;
    const int O = (MODE == 1) ? 23 : 18;
    T x[3], y[14], W3, K0[14];
    for (int i = 0; i < 3; ++i) { x[i] = (T)state[i]; y[i] = (T)state[3 + i]; }
    for (int i = 0; i < 9; ++i) y[3 + i] = (T)state[6 + i];
    y[12] = (T)state[15]; y[13] = (T)state[16]; W3 = (T)state[17];
    for (int i = 0; i < 14; ++i) K0[i] = 0;
    // ---- A3
    int fl = ensure_so3<T>(y + 3);
    EnvRegs<T> r;
    for (int i = 0; i < 3; ++i) r.x[i] = x[i];
    for (int i = 0; i < 14; ++i) r.y[i] = y[i];
    r.W3 = W3;
    r.m = (T)params[0]; r.d = (T)params[1]; r.J1 = (T)params[2]; r.J3 = (T)params[3]; r.c_tf = (T)params[4]; r.c_tw = (T)params[5];
    if (c.goal_mode == 1) {
        const T Wv[3] = {y[12], y[13], W3};
        const T b1d[3] = {(T)goal_io[6], (T)goal_io[7], (T)goal_io[8]};
        T Wd[3];
        traj_wd<T>(y + 3, Wv, b1d, Wd);
        for (int i = 0; i < 3; ++i) goal_io[9 + i] = (double)Wd[i];
    }
    const int A = (MODE == 2) ? 5 : 4;
    T act[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < A; ++i) act[i] = (T)action[i];
    T f, M[3];
    action_to_fM<T>(r, c, act, act_is_f32 != 0, f, M, MODE);
    Dyn<T> d;
    {
        const T rm = (T)1 / r.m, rJ1 = (T)1 / r.J1, rJ3 = (T)1 / r.J3;
        d.fm = f * rm; d.g = c.g;
        d.Mi0 = M[0] * rJ1; d.Mi1 = M[1] * rJ1;
        d.kw0 = (r.J1 - r.J3) * rJ1; d.kw1 = (r.J3 - r.J1) * rJ1;
        d.w3dot = M[2] * rJ3;
    }
    OdeLane<T> ode;
    ode.t = 0; ode.h_abs = c.dt; ode.rejected = 0; ode.nfev = 0; ode.status = 0; ode.nproj = 0; ode.checked = 0;
    bool finite = true;
    for (int i = 0; i < 3; ++i) finite = finite && (num<T>::abs(x[i]) <= num<T>::huge);
    for (int i = 0; i < 14; ++i) finite = finite && (num<T>::abs(y[i]) <= num<T>::huge);
    finite = finite && (num<T>::abs(W3) <= num<T>::huge);
    bool fin = false;
    if (!finite) {
        ode.t = c.dt; ode.h_abs = 0; ode.status = 1; fin = true;
    } else if (MODE == 0 && c.integrator == 1) {
        T kk[14];
        rhs14<T>(y, W3, d, kk);
        for (int i = 0; i < 3; ++i) x[i] = num<T>::fma(y[i], c.dt, x[i]);
        for (int i = 0; i < 14; ++i) y[i] = num<T>::fma(kk[i], c.dt, y[i]);
        W3 = num<T>::fma(d.w3dot, c.dt, W3);
        ode.nfev = 1; fin = true;
    } else {
        dop853_begin<T>(x, y, W3, d, c.dt, c.rtol, c.atol, K0, ode);
    }
    if (fl & 2) ode.status |= 4;
    ode.nproj += fl & 1;
    // ---- B
    static T ks[QR_NSLOTS * QR_SLOT_ELEMS];
    int guard = 0;
    while (!fin && guard++ < 100000) {
        fin = dop853_attempt<T>(x, y, W3, d, c.dt, c.rtol, c.atol, K0, ode, ks, 0, true);
        if (ode.checked) {   // as phase B of k_step: the redo takes and returns the components in the internal order
            T tz[14];
            to_z<T>(y, tz);
            fin = dop853_attempt_checked<T>(x, tz, &W3, &d, c.dt, c.rtol, c.atol, K0, &ode);
            from_z<T>(tz, y);
        }
    }
    // ---- A1
    for (int i = 0; i < 3; ++i) r.x[i] = x[i];
    for (int i = 0; i < 14; ++i) r.y[i] = y[i];
    r.W3 = W3;
    for (int i = 0; i < 8; ++i) r.I[i] = (T)integ[i];
    for (int i = 0; i < 12; ++i) r.goal[i] = (T)goal_io[i];
    float o[23];
    double rew[2] = {0, 0};
    int dn[2] = {0, 0};
    int st = ode.status;
    if (MODE == 0) {
        for (int i = 0; i < 3; ++i) o[i] = (float)x[i];
        for (int i = 0; i < 12; ++i) o[3 + i] = (float)y[i];
        o[15] = (float)y[12]; o[16] = (float)y[13]; o[17] = (float)W3;
        reward_done_quad<T>(r, c, rew, dn);
    } else {
        int f2 = norm_error_state<T>(r, c, o, MODE);
        if (f2 & 2) st |= 4;
        reward_done<T, double>(c, o, rew, dn, MODE);
    }
    for (int i = 0; i < 3; ++i) { state_out[i] = (double)x[i]; state_out[3 + i] = (double)y[i]; }
    for (int i = 0; i < 9; ++i) state_out[6 + i] = (double)y[3 + i];
    state_out[15] = (double)y[12]; state_out[16] = (double)y[13]; state_out[17] = (double)W3;
    for (int i = 0; i < 8; ++i) integ_out[i] = (MODE == 0) ? integ[i] : (double)r.I[i];
    for (int i = 0; i < O; ++i) obs_out[i] = o[i];
    reward_out[0] = rew[0]; reward_out[1] = rew[1];
    done_out[0] = dn[0]; done_out[1] = dn[1];
    *nfev_out = ode.nfev; *nproj_out = ode.nproj;
    return st;
}


}  // namespace

extern "C" {

// One env.step().  state 18 (x v R-colmajor W), integ 8, params 6 (m d J1 J3 c_tf c_tw), goal 12; float64 interface,
// computed in float64 (tw_step) or float32 (tw_step_f32: the headline arithmetic, with exact 1/x and sqrt in place of
// the MUFU approximations).  goal_mode 1 recomputes Wd from the pre-step state (written back to goal_io).  Returns
// the integrator status bits.
#define TW_ARGS const qr_config* cfg, const double* state, const double* integ, const double* params, double* goal_io, \
    const double* action, int act_is_f32, double* state_out, double* integ_out, float* obs_out, double* reward_out, \
    int* done_out, int* nfev_out, int* nproj_out
#define TW_PASS cfg, state, integ, params, goal_io, action, act_is_f32, state_out, integ_out, obs_out, reward_out, done_out, nfev_out, nproj_out
int tw_step(TW_ARGS) { return step_one<double>(TW_PASS); }
int tw_step_f32(TW_ARGS) { return step_one<float>(TW_PASS); }

// env.get_norm_error_state(): observation from state + goal, integral errors advanced once (in place)
void tw_norm_error_state(const qr_config* cfg, const double* state, double* integ_io, const double* goal, float* obs_out)
{
    const EnvConst<double> c = make_const<double>(*cfg);
    EnvRegs<double> r;
    memset(&r, 0, sizeof(r));
    for (int i = 0; i < 3; ++i) { r.x[i] = state[i]; r.y[i] = state[3 + i]; }
    for (int i = 0; i < 9; ++i) r.y[3 + i] = state[6 + i];
    r.y[12] = state[15]; r.y[13] = state[16]; r.W3 = state[17];
    for (int i = 0; i < 8; ++i) r.I[i] = integ_io[i];
    for (int i = 0; i < 12; ++i) r.goal[i] = goal[i];
    float o[23];
    norm_error_state<double>(r, c, o, c.mode);
    for (int i = 0; i < 8; ++i) integ_io[i] = r.I[i];
    const int O = (c.mode == 1) ? 23 : 18;
    for (int i = 0; i < O; ++i) obs_out[i] = o[i];
}

// env.reset(env_type) for global env id `gid`, episode index `episode` (+ the mode-0 goal when goal_mode == 1)
void tw_reset(const qr_config* cfg, uint64_t gid, uint32_t episode, int env_type, double* state_out, double* integ_out,
              double* params_out, double* goal_out)
{
    const Philox ph{(uint32_t)cfg->seed, (uint32_t)(cfg->seed >> 32)};
    EnvRegs<double> r;
    memset(&r, 0, sizeof(r));
    double theta;
    reset_env<double>(r, ph, gid, episode, env_type, cfg->udm_pct, &theta);
    for (int i = 0; i < 12; ++i) r.goal[i] = 0;
    r.goal[6] = 1.0;
    if (cfg->goal_mode == 1) init_goal_mode0<double>(r, theta);
    for (int i = 0; i < 3; ++i) { state_out[i] = r.x[i]; state_out[3 + i] = r.y[i]; }
    for (int i = 0; i < 9; ++i) state_out[6 + i] = r.y[3 + i];
    state_out[15] = r.y[12]; state_out[16] = r.y[13]; state_out[17] = r.W3;
    for (int i = 0; i < 8; ++i) integ_out[i] = r.I[i];
    params_out[0] = r.m; params_out[1] = r.d; params_out[2] = r.J1; params_out[3] = r.J3; params_out[4] = r.c_tf; params_out[5] = r.c_tw;
    for (int i = 0; i < 12; ++i) goal_out[i] = r.goal[i];
}

// trajectory_generator.get_desired for the on-device modes (1 hover, 5 circle, 6 eight): ts[12] is carried by the caller
void tw_traj_start(const double* x, const double* R_so3, double* ts) { traj_start<double>(x, R_so3, ts); }
void tw_traj_desired(int mode, const double* x, const double* v, const double* R, const double* W, double* ts, double* goal,
                     double u_ttraj, double u_w, double dt)
{
    traj_desired<double>(mode, x, v, R, W, ts, goal, u_ttraj, u_w, dt);
}

}  // extern "C"
