/* quad_oracle.c -- CPU restatement of gym-rotor's env.step() hot path (see quad_oracle.h).
 * TEST INFRASTRUCTURE ONLY.  Build: make -C oracle  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>
#include "quad_oracle.h"
#include "dop853_tableau.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---- tableau as dense arrays (zeros skipped at use) ------------------------------------------------ */
#define TAB(T, name)                                                                                       \
    static const T dopA##name[12][12] = {                                                                  \
        {0},                                                                                               \
        {(T)DOP_A1_0},                                                                                     \
        {(T)DOP_A2_0, (T)DOP_A2_1},                                                                        \
        {(T)DOP_A3_0, 0, (T)DOP_A3_2},                                                                     \
        {(T)DOP_A4_0, 0, (T)DOP_A4_2, (T)DOP_A4_3},                                                        \
        {(T)DOP_A5_0, 0, 0, (T)DOP_A5_3, (T)DOP_A5_4},                                                     \
        {(T)DOP_A6_0, 0, 0, (T)DOP_A6_3, (T)DOP_A6_4, (T)DOP_A6_5},                                        \
        {(T)DOP_A7_0, 0, 0, (T)DOP_A7_3, (T)DOP_A7_4, (T)DOP_A7_5, (T)DOP_A7_6},                           \
        {(T)DOP_A8_0, 0, 0, (T)DOP_A8_3, (T)DOP_A8_4, (T)DOP_A8_5, (T)DOP_A8_6, (T)DOP_A8_7},              \
        {(T)DOP_A9_0, 0, 0, (T)DOP_A9_3, (T)DOP_A9_4, (T)DOP_A9_5, (T)DOP_A9_6, (T)DOP_A9_7, (T)DOP_A9_8}, \
        {(T)DOP_A10_0, 0, 0, (T)DOP_A10_3, (T)DOP_A10_4, (T)DOP_A10_5, (T)DOP_A10_6, (T)DOP_A10_7,         \
         (T)DOP_A10_8, (T)DOP_A10_9},                                                                      \
        {(T)DOP_A11_0, 0, 0, (T)DOP_A11_3, (T)DOP_A11_4, (T)DOP_A11_5, (T)DOP_A11_6, (T)DOP_A11_7,         \
         (T)DOP_A11_8, (T)DOP_A11_9, (T)DOP_A11_10}};                                                      \
    static const T dopB##name[12] = {(T)DOP_B0, 0, 0, 0, 0, (T)DOP_B5, (T)DOP_B6, (T)DOP_B7,               \
                                     (T)DOP_B8, (T)DOP_B9, (T)DOP_B10, (T)DOP_B11};                        \
    static const T dopE3##name[13] = {(T)DOP_E3_0, 0, 0, 0, 0, (T)DOP_E3_5, (T)DOP_E3_6, (T)DOP_E3_7,      \
                                      (T)DOP_E3_8, (T)DOP_E3_9, (T)DOP_E3_10, (T)DOP_E3_11, 0};            \
    static const T dopE5##name[13] = {(T)DOP_E5_0, 0, 0, 0, 0, (T)DOP_E5_5, (T)DOP_E5_6, (T)DOP_E5_7,      \
                                      (T)DOP_E5_8, (T)DOP_E5_9, (T)DOP_E5_10, (T)DOP_E5_11, 0};
TAB(double, _f64)
TAB(float, _f32)

/* test instrumentation: SO(3) re-projections that happened inside DOP853 stages / f_new (not the Euler probe) */
static long long qo_stage_projections = 0;
long long qo_stage_projection_count(int reset) { long long v = qo_stage_projections; if (reset) qo_stage_projections = 0; return v; }

/* ---- instantiate for double ------------------------------------------------------------------------ */
#define REAL double
#define RWD double
#define SUF _f64
#define FMA fma
#define FABS fabs
#define SQRT sqrt
#define POW pow
#define ATAN2 atan2
#define COS cos
#define SIN sin
#define ACOS acos
#define NEXTAFTER nextafter
#define FMAX fmax
#include "quad_oracle_impl.inc"
#undef REAL
#undef RWD
#undef SUF
#undef FMA
#undef FABS
#undef SQRT
#undef POW
#undef ATAN2
#undef COS
#undef SIN
#undef ACOS
#undef NEXTAFTER
#undef FMAX

/* ---- instantiate for float ------------------------------------------------------------------------- */
#define REAL float
#define RWD float
#define SUF _f32
#define FMA fmaf
#define FABS fabsf
#define SQRT sqrtf
#define POW powf
#define ATAN2 atan2f
#define COS cosf
#define SIN sinf
#define ACOS acosf
#define NEXTAFTER nextafterf
#define FMAX fmaxf
#include "quad_oracle_impl.inc"
#undef REAL
#undef RWD
#undef SUF

/* ---- public API ------------------------------------------------------------------------------------ */

/* plain pthreads over contiguous env blocks (no OpenMP runtime: keeps the oracle free of libgomp clashes
 * with the interpreter that loads it) */
static int g_threads = 1;
void qo_set_threads(int n) { g_threads = n > 0 ? n : 1; }
int qo_get_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct qo_job {
    void (*fn)(void* ctx, int64_t lo, int64_t hi);
    void* ctx;
    int64_t lo, hi;
} qo_job;

static void* qo_job_main(void* a)
{
    qo_job* j = (qo_job*)a;
    j->fn(j->ctx, j->lo, j->hi);
    return 0;
}

static void qo_parallel_for(int64_t n, void (*fn)(void*, int64_t, int64_t), void* ctx)
{
    int nt = g_threads;
    if (nt > 256) nt = 256;
    if ((int64_t)nt > n) nt = (int)(n > 0 ? n : 1);
    if (nt <= 1) { fn(ctx, 0, n); return; }
    pthread_t th[256];
    qo_job jobs[256];
    int64_t chunk = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].lo = t * chunk; jobs[t].hi = (t + 1) * chunk < n ? (t + 1) * chunk : n;
        if (jobs[t].lo > n) jobs[t].lo = n;
        pthread_create(&th[t], 0, qo_job_main, &jobs[t]);
    }
    for (int t = 0; t < nt; ++t) pthread_join(th[t], 0);
}

int qo_obs_dim(int mode) { return mode == QO_MODE_COUPLED ? 23 : 18; }
int qo_act_dim(int mode) { return mode == QO_MODE_DECOUPLED ? 5 : 4; }
int qo_num_agents(int mode) { return mode == QO_MODE_DECOUPLED ? 2 : 1; }

void qo_default_config(qo_config* c, int mode)
{
    memset(c, 0, sizeof(*c));
    c->mode = mode;
    c->integrator = QO_INT_DOP853;
    c->act_f32 = 0;
    c->dt = 1. / 200;          /* quad.py:60-61 */
    c->g = 9.81;               /* quad.py:33 */
    c->rtol = 1e-3;            /* scipy rk.py:86 */
    c->atol = 1e-6;
    c->x_lim = 1.0;            /* quad.py:104-106 */
    c->v_lim = 4.0;
    c->W_lim = 2 * M_PI;
    c->eIx_lim = 3.0;          /* coupled:23-24 */
    c->eIb1_lim = 3.0;
    c->sat_sigma = 1.;         /* quad.py:91 */
    c->alpha = 0.01;           /* args_parse.py:27 */
    c->beta = 0.05;            /* args_parse.py:32 */
    c->Cx = 6.0; c->CIx = 0.1; c->Cv = 0.4; c->Cw12 = 0.6;   /* args_parse.py:23-26 */
    c->Cb1 = 6.0; c->CIb1 = 0.1; c->CW3 = 0.1;               /* args_parse.py:29-31 */
    c->CW = c->Cw12;           /* quad.py:80 */
    c->reward_min = -ceil(c->Cx + c->CIx + c->Cv + c->Cb1 + c->CIb1 + c->CW);  /* quad.py:81 */
    c->reward_min_1 = -ceil(c->Cx + c->CIx + c->Cv + c->Cw12);                 /* quad.py:85 */
    c->reward_min_2 = -ceil(c->Cb1 + c->CW3 + c->CIb1);                        /* quad.py:88 */
    c->min_force = 0.5;        /* quad.py:39 */
    c->euler_lim_deg = 85;     /* quad.py:107 */
}

typedef struct step_ctx {
    const qo_config* c;
    void *state, *integ; const void *params, *goal, *action; float* obs; void* reward; uint8_t* done;
    int32_t* nfev; uint8_t* status;
} step_ctx;

static void step_range_f64(void* vp, int64_t lo, int64_t hi)
{
    step_ctx* s = (step_ctx*)vp;
    const qo_config* c = s->c;
    const int A = qo_act_dim(c->mode), O = qo_obs_dim(c->mode), G = qo_num_agents(c->mode);
    for (int64_t e = lo; e < hi; ++e)
        step_one_f64(c, (double*)s->state + 18 * e, (double*)s->integ + 8 * e, (const double*)s->params + 6 * e,
                     (const double*)s->goal + 12 * e, (const double*)s->action + A * e, s->obs + O * e,
                     (double*)s->reward + G * e, s->done + G * e, s->nfev ? s->nfev + e : 0,
                     s->status ? s->status + e : 0);
}

static void step_range_f32(void* vp, int64_t lo, int64_t hi)
{
    step_ctx* s = (step_ctx*)vp;
    const qo_config* c = s->c;
    const int A = qo_act_dim(c->mode), O = qo_obs_dim(c->mode), G = qo_num_agents(c->mode);
    for (int64_t e = lo; e < hi; ++e)
        step_one_f32(c, (float*)s->state + 18 * e, (float*)s->integ + 8 * e, (const float*)s->params + 6 * e,
                     (const float*)s->goal + 12 * e, (const float*)s->action + A * e, s->obs + O * e,
                     (float*)s->reward + G * e, s->done + G * e, s->nfev ? s->nfev + e : 0,
                     s->status ? s->status + e : 0);
}

int qo_step_f64(const qo_config* c, int64_t n, double* state, double* integ, const double* params,
                const double* goal, const double* action, float* obs, double* reward, uint8_t* done,
                int32_t* nfev, uint8_t* status)
{
    step_ctx s = {c, state, integ, params, goal, action, obs, reward, done, nfev, status};
    qo_parallel_for(n, step_range_f64, &s);
    return 0;
}

int qo_step_f32(const qo_config* c, int64_t n, float* state, float* integ, const float* params,
                const float* goal, const float* action, float* obs, float* reward, uint8_t* done,
                int32_t* nfev, uint8_t* status)
{
    step_ctx s = {c, state, integ, params, goal, action, obs, reward, done, nfev, status};
    qo_parallel_for(n, step_range_f32, &s);
    return 0;
}

int qo_norm_error_state_f64(const qo_config* c, int64_t n, const double* state, double* integ,
                            const double* goal, float* obs)
{
    const int O = qo_obs_dim(c->mode);
    for (int64_t e = 0; e < n; ++e) {
        int bad = 0;
        norm_error_state_f64(c, state + 18 * e, integ + 8 * e, goal + 12 * e, obs + O * e, &bad);
    }
    return 0;
}

int qo_rhs_f64(int64_t n, const double* y, const double* params, const double* fM, double* ydot)
{
    for (int64_t e = 0; e < n; ++e) {
        rhs_par_f64 p;
        p.m = params[6 * e]; p.J1 = params[6 * e + 2]; p.J3 = params[6 * e + 3]; p.g = 9.81;
        p.f = fM[4 * e]; p.M[0] = fM[4 * e + 1]; p.M[1] = fM[4 * e + 2]; p.M[2] = fM[4 * e + 3];
        p.n_svd = 0; p.svd_bad = 0; p.n_svd_stage = 0; p.in_probe = 0;
        rhs_f64(y + 18 * e, ydot + 18 * e, &p);
    }
    return 0;
}

int64_t qo_ensure_so3_f64(int64_t n, double* R)
{
    int64_t k = 0;
    for (int64_t e = 0; e < n; ++e) k += ensure_so3_f64(R + 9 * e, 0);
    return k;
}

/* ---- reset (quad.py:171-222, 338-404) with the uniforms supplied by the caller ------------------------ */

static double lerp_u(double lo, double hi, double u) { return lo + (hi - lo) * u; }

int qo_reset_from_uniforms_f64(const qo_config* c, int env_type, double udm_pct, int64_t n, const double* u,
                               double* state, double* integ, double* params)
{
    (void)c;
    const double m0 = 2.15, d0 = 0.23, J10 = 0.022, J30 = 0.035, ctf0 = 0.0135, ctw0 = 2.2; /* quad.py:28-32 */
    for (int64_t e = 0; e < n; ++e) {
        const double* q = u + 20 * e;
        double* p = params + 6 * e;
        p[0] = m0; p[1] = d0; p[2] = J10; p[3] = J30; p[4] = ctf0; p[5] = ctw0;
        if (env_type == QO_ENV_TRAIN) { /* quad.py:368-386: U(nom -+ 10 %), c_tw -+ 5 % */
            double r = udm_pct / 100.0;
            p[0] = lerp_u(m0 - m0 * r, m0 + m0 * r, q[0]);
            p[1] = lerp_u(d0 - d0 * r, d0 + d0 * r, q[1]);
            p[2] = lerp_u(J10 - J10 * r, J10 + J10 * r, q[2]);
            p[3] = lerp_u(J30 - J30 * r, J30 + J30 * r, q[3]);
            p[4] = lerp_u(ctf0 - ctf0 * r, ctf0 + ctf0 * r, q[4]);
            p[5] = lerp_u(ctw0 - ctw0 * (r / 2.), ctw0 + ctw0 * (r / 2.), q[5]);
        }
        double yaw = lerp_u(-M_PI, M_PI, q[6]); /* quad.py:339 */
        double ix, iv, iR, iW;                  /* quad.py:340-356 */
        if (env_type == QO_ENV_TRAIN) {
            if (q[7] < 0.2) { ix = 0; iv = 0; iR = 0; iW = 0; }
            else { ix = 0.6; iv = 4.0 * 0.5; iR = 50 * (M_PI / 180.); iW = 2 * M_PI * 0.5; }
        } else { ix = 0.4; iv = 0; iR = 0; iW = 0; }
        double* s = state + 18 * e;
        for (int i = 0; i < 3; ++i) {
            s[i] = lerp_u(-ix, ix, q[8 + i]);
            s[3 + i] = lerp_u(-iv, iv, q[11 + i]);
            s[15 + i] = lerp_u(-iW, iW, q[14 + i]);
        }
        double roll = lerp_u(-iR, iR, q[17]), pitch = lerp_u(-iR, iR, q[18]);
        double cr = cos(roll), sr = sin(roll), cp = cos(pitch), sp = sin(pitch), cy = cos(yaw), sy = sin(yaw);
        /* R = Rz(yaw) Ry(pitch) Rx(roll)  (scipy Rotation.from_euler('xyz'), quad.py:199), column-major */
        s[6] = cy * cp;  s[9]  = cy * sp * sr - sy * cr;  s[12] = cy * sp * cr + sy * sr;
        s[7] = sy * cp;  s[10] = sy * sp * sr + cy * cr;  s[13] = sy * sp * cr - cy * sr;
        s[8] = -sp;      s[11] = cp * sr;                 s[14] = cp * cr;
        for (int i = 0; i < 8; ++i) integ[8 * e + i] = 0.0;
    }
    return 0;
}

/* ---- trajectory generator, mode 0 (utils/trajectory_generator.py:113-173, 196-221) ------------------------ */

/* Wd = [0, 0, b3 . (b1c x b1c_dot)] from the CURRENT state and the stored b1d (b1d_dot = 0 in mode 0):
 * trajectory_generator.py:165-172.  get_desired first runs state_decomposition (ensure_SO3) on the state. */
int qo_traj_wd_f64(int64_t n, const double* state, const double* b1d, double* Wd)
{
    for (int64_t e = 0; e < n; ++e) {
        double R[9];
        for (int i = 0; i < 9; ++i) R[i] = state[18 * e + 6 + i];
        ensure_so3_f64(R, 0);
        const double* W = state + 18 * e + 15;
        const double* bd = b1d + 3 * e;
        const double* b3 = R + 6;
        double b3d[3];
        for (int i = 0; i < 3; ++i) b3d[i] = R[i] * W[1] - R[i + 3] * W[0];      /* R hat(W) e3 */
        double dp = fma(bd[2], b3[2], fma(bd[1], b3[1], bd[0] * b3[0]));
        double dq = fma(bd[2], b3d[2], fma(bd[1], b3d[1], bd[0] * b3d[0]));
        double b1c[3], b1cd[3];
        for (int i = 0; i < 3; ++i) {
            b1c[i] = bd[i] - dp * b3[i];
            b1cd[i] = 0.0 - ((0.0 * b3[i] + dq * b3[i]) + dp * b3d[i]);
        }
        double oc0 = b1c[1] * b1cd[2] - b1c[2] * b1cd[1];
        double oc1 = b1c[2] * b1cd[0] - b1c[0] * b1cd[2];
        double oc2 = b1c[0] * b1cd[1] - b1c[1] * b1cd[0];
        Wd[3 * e] = 0; Wd[3 * e + 1] = 0;
        Wd[3 * e + 2] = fma(b3[2], oc2, fma(b3[1], oc1, b3[0] * oc0));
    }
    return 0;
}

/* mark_traj_start + first get_desired(mode 0) (main.py:226-229): b1d = Rz(theta) [cos psi, sin psi, 0] with psi the
 * heading of b1 of the (float32-cast) reset state; theta is the caller's draw of U(-25 deg, 25 deg). */
int qo_traj_init_mode0_f64(int64_t n, const double* state, const double* theta, double* b1d)
{
    for (int64_t e = 0; e < n; ++e) {
        double R[9];
        for (int i = 0; i < 9; ++i) R[i] = (double)(float)state[18 * e + 6 + i];
        ensure_so3_f64(R, 0);
        double psi = atan2(R[1], R[0]);
        double cps = cos(psi), sps = sin(psi), cth = cos(theta[e]), sth = sin(theta[e]);
        b1d[3 * e] = cth * cps - sth * sps; b1d[3 * e + 1] = sth * cps + cth * sps; b1d[3 * e + 2] = 0.0;
    }
    return 0;
}

/* ---- trajectory generator, modes 1 / 5 / 6 and the manual fallback (utils/trajectory_generator.py:113-173, 232-505) ----
 * Per-env trajectory state ts[n][12]:
 *   0 t | 1 flags (bit0 trajectory_started, bit1 manual_mode, bit2 manual_mode_init) | 2..4 x_init / circle or eight centre |
 *   5 theta_init | 6 w_b1d | 7 smooth_term (hover) | 8 t_traj | 9,10 b1d_dot x,y | 11 unused                                  */

/* mark_traj_start (trajectory_generator.py:176-192): clock and flags to zero, initial heading from the state handed in
 * (the float32 reset state in main.py:226-227). */
int qo_traj_start_f64(int64_t n, const double* state, double* ts)
{
    for (int64_t e = 0; e < n; ++e) {
        double R[9];
        for (int i = 0; i < 9; ++i) R[i] = state[18 * e + 6 + i];
        ensure_so3_f64(R, 0);
        double* s = ts + 12 * e;
        for (int i = 0; i < 12; ++i) s[i] = 0.0;
        for (int i = 0; i < 3; ++i) s[2 + i] = state[18 * e + i];
        s[5] = atan2(R[1], R[0]);
    }
    return 0;
}

/* One get_desired(state, mode) call (trajectory_generator.py:113-173) for mode 1 (hover), 2 (take-off), 3 (land),
 * 4 (stay), 5 (circle), >= 6 (figure eight).  goal[n][12] = xd vd b1d Wd is read and updated in place (several branches leave components untouched).
 * numpy detail that is part of the observable behaviour: set_desired_states_to_current makes xd / vd COPIES of the
 * state's x / v (:212-215); under main.py's protocol the trajectory starts from the float32 reset state
 * (main.py:226-228), so xd and vd are float32 arrays for the whole trajectory and every element assignment rounds
 * to float32 (F32 below).  After a switch to manual mode they are float64 copies of the then-current state.
 * draws[n][2]: uniforms in [0,1) for the hover's t_traj ~ U(2,5) and w_b1d ~ U(-0.15 pi, 0.15 pi).  */
#define F32(v) ((double)(float)(v))
int qo_traj_desired_f64(int mode, int64_t n, const double* state, double* ts, double* goal, const double* draws, double dt)
{
    const double circle_radius = 0.7, circle_linear_v = 0.4, circle_W = 0.4;   /* :87-90 */
    const int num_circles = 2;
    const double eight_A1 = 1.5, eight_A2 = 1.0, eight_T = 9.0, eight_w_b1d = 0.349066;   /* :93-98 */
    const int num_of_eights = 3;
    const double eight_w1 = 2 * M_PI / eight_T, eight_w2 = 4 * M_PI / eight_T;
    const double eight_exp_xy = -log(0.01) / eight_T, eight_alt_d = -0.6;       /* :101-105 */
    for (int64_t e = 0; e < n; ++e) {
        const double* y = state + 18 * e;
        double R[9];
        for (int i = 0; i < 9; ++i) R[i] = y[6 + i];
        ensure_so3_f64(R, 0);
        const double* x = y; const double* v = y + 3; const double* W = y + 15;
        double* s = ts + 12 * e;
        double* g = goal + 12 * e;
        double* xd = g; double* vd = g + 3; double* b1d = g + 6; double* Wd = g + 9;
        int flags = (int)s[1];
        const double th_cur = atan2(R[1], R[0]);   /* get_current_b1 */
        if (flags & 2) {   /* manual() :232-249; calculate_desired returns before the Wd block */
            if (!(flags & 4)) {
                for (int i = 0; i < 3; ++i) { xd[i] = x[i]; vd[i] = v[i]; }
                s[5] = th_cur;
                flags |= 4;
            }
            vd[0] = vd[1] = vd[2] = 0.0;
            b1d[0] = cos(s[5]); b1d[1] = sin(s[5]); b1d[2] = 0.0;
            s[1] = (double)flags;
            continue;
        }
        if (mode == 1) {   /* hovering :252-277 */
            if (!(flags & 1)) {
                for (int i = 0; i < 3; ++i) { xd[i] = x[i]; vd[i] = v[i]; s[2 + i] = x[i]; }
                b1d[0] = cos(th_cur); b1d[1] = sin(th_cur); b1d[2] = 0.0;
                flags |= 1;
                s[8] = 2.0 + (5.0 - 2.0) * draws[2 * e];
                s[7] = -log(0.001) / s[8];
                s[6] = -0.15 * M_PI + (0.15 * M_PI - (-0.15 * M_PI)) * draws[2 * e + 1];
            }
            s[0] = s[0] + dt;
            const double t = s[0], k = s[7], w = s[6];
            for (int i = 0; i < 3; ++i) {
                xd[i] = F32((s[2 + i] - 0.0) * exp(-k * t) + 0.0);
                vd[i] = F32(-(s[2 + i] - 0.0) * k * exp(-k * t));
            }
            b1d[0] = cos(w * t + s[5]); b1d[1] = sin(w * t + s[5]); b1d[2] = 0.0;
            s[9] = -w * sin(w * t + s[5]); s[10] = w * cos(w * t + s[5]);
        } else if (mode == 5) {   /* circle :359-412 */
            if (!(flags & 1)) {
                for (int i = 0; i < 3; ++i) { xd[i] = x[i]; vd[i] = v[i]; s[2 + i] = x[i]; }
                b1d[0] = cos(th_cur); b1d[1] = sin(th_cur); b1d[2] = 0.0;
                flags |= 1;
                s[8] = circle_radius / circle_linear_v + num_circles * 2 * M_PI / circle_W;
            }
            s[0] = s[0] + dt;
            const double t = s[0];
            if (t < circle_radius / circle_linear_v) {
                /* np.float32 + python float -> float32 arithmetic (NEP 50): both operands rounded first */
                xd[0] = (double)((float)s[2] + (float)(circle_linear_v * t));
                vd[0] = F32(circle_linear_v);
            } else if (t < s[8]) {
                const double tt = t - circle_radius / circle_linear_v, th = circle_W * tt;
                xd[0] = F32(circle_radius * cos(th) + s[2]);
                vd[0] = F32(-circle_radius * circle_W * sin(th));
                xd[1] = F32(circle_radius * sin(th) + s[3]);
                vd[1] = F32(circle_radius * circle_W * cos(th));
                const double thb = circle_W * tt + M_PI;
                b1d[0] = cos(thb); b1d[1] = sin(thb); b1d[2] = 0.0;
                s[9] = -circle_W * sin(thb); s[10] = circle_W * cos(thb);
            } else {
                flags |= 2;   /* mark_traj_end(True): manual mode from the next call on */
            }
        } else if (mode == 2) {   /* takeoff :280-309 */
            /* set_desired_states_to_zero makes xd / vd fresh FLOAT64 arrays; x_init is the (float32) state handed in
             * at the start, so "x_init[2] + takeoff_velocity * t" and t_traj are float32 expressions (NEP 50: the
             * python-float operand is cast to float32), and "t < t_traj" compares in float32. */
            const double takeoff_end_height = -0.5, takeoff_velocity = -0.05;   /* :82-83 */
            if (!(flags & 1)) {
                for (int i = 0; i < 3; ++i) { xd[i] = 0.0; vd[i] = 0.0; s[2 + i] = x[i]; }
                xd[0] = x[0]; xd[1] = x[1];
                s[8] = (double)(((float)takeoff_end_height - (float)x[2]) / (float)takeoff_velocity);
                b1d[0] = cos(th_cur); b1d[1] = sin(th_cur); b1d[2] = 0.0;
                flags |= 1;
            }
            s[0] = s[0] + dt;
            const double t = s[0];
            if ((float)t < (float)s[8]) {
                xd[2] = (double)((float)s[4] + (float)(takeoff_velocity * t));
            } else {
                const double d0 = xd[0] - x[0], d1 = xd[1] - x[1], d2 = xd[2] - x[2];
                if (sqrt(d0 * d0 + d1 * d1 + d2 * d2) < 0.04) {   /* waypoint_reached(xd, x, 0.04) */
                    xd[2] = takeoff_end_height; vd[2] = 0.0;
                    flags |= 2;   /* mark_traj_end(True) */
                }
            }
        } else if (mode == 3) {   /* land :322-349 */
            const double landing_velocity = 1.0, cutoff = -0.25;   /* :86-87 */
            if (!(flags & 1)) {
                for (int i = 0; i < 3; ++i) { xd[i] = x[i]; vd[i] = v[i]; s[2 + i] = x[i]; }
                s[8] = (double)(((float)cutoff - (float)x[2]) / (float)landing_velocity);
                b1d[0] = cos(th_cur); b1d[1] = sin(th_cur); b1d[2] = 0.0;
                flags |= 1;
            }
            s[0] = s[0] + dt;
            const double t = s[0];
            if ((float)t < (float)s[8]) {
                xd[2] = (double)((float)s[4] + (float)(landing_velocity * t));
            } else if (x[2] > cutoff) {
                xd[2] = cutoff; vd[2] = 0.0;   /* mark_traj_end(False): no manual mode; the branch repeats */
            } else {
                xd[2] = cutoff; vd[2] = F32(landing_velocity);
            }
        } else if (mode == 4) {   /* stay :352-357: no clock update; manual mode from the next call on */
            if (!(flags & 1)) {
                for (int i = 0; i < 3; ++i) { xd[i] = x[i]; vd[i] = v[i]; s[2 + i] = x[i]; }
                b1d[0] = cos(th_cur); b1d[1] = sin(th_cur); b1d[2] = 0.0;
                flags |= 1;
            }
            flags |= 2;
        } else {   /* eight_shaped_curve :415-505 */
            if (!(flags & 1)) {
                for (int i = 0; i < 3; ++i) { xd[i] = x[i]; vd[i] = v[i]; s[2 + i] = x[i]; }
                b1d[0] = cos(th_cur); b1d[1] = sin(th_cur); b1d[2] = 0.0;
                flags |= 1;
                s[8] = num_of_eights * eight_T;
                s[6] = eight_w_b1d;
            }
            s[0] = s[0] + dt;
            const double t = s[0];
            if (t < s[8]) {
                const double ex = 1. - exp(-eight_exp_xy * t), dex = eight_exp_xy * exp(-eight_exp_xy * t);
                xd[0] = F32(eight_A2 * (sin(eight_w2 * t) * ex) + s[2]);
                vd[0] = F32(eight_A2 * ((eight_w2 * cos(eight_w2 * t) * ex) + (sin(eight_w2 * t) * dex)));
                xd[1] = F32(eight_A1 * (cos(eight_w1 * t) - 1.) * ex + s[3]);
                vd[1] = F32(eight_A1 * ((eight_w1 * -sin(eight_w1 * t) * ex) + (cos(eight_w1 * t) - 1.) * dex));
                /* the altitude terms start from the float32 centre and only meet python floats: all float32 */
                const float za = ((float)s[4] - (float)eight_alt_d) / 2.0f;
                xd[2] = (double)(za * (float)(1 - cos(eight_w1 * t)) + (float)s[4]);
                vd[2] = (double)((za * (float)eight_w1) * (float)sin(eight_w1 * t));
                const double wt = s[6] * t * ex + s[5], dwt = s[6] * (ex + t * dex);
                b1d[0] = cos(wt); b1d[1] = sin(wt); b1d[2] = 0.0;
                s[9] = -sin(wt) * dwt; s[10] = cos(wt) * dwt;
            } else {
                flags |= 2;
            }
        }
        s[1] = (double)flags;
        /* Wd (:165-172) with the current b1d_dot */
        const double* b3 = R + 6;
        double b3d[3], bdd[3] = {s[9], s[10], 0.0};
        for (int i = 0; i < 3; ++i) b3d[i] = R[i] * W[1] - R[i + 3] * W[0];
        double dp = fma(b1d[2], b3[2], fma(b1d[1], b3[1], b1d[0] * b3[0]));
        double dq = fma(b1d[2], b3d[2], fma(b1d[1], b3d[1], b1d[0] * b3d[0]));
        double dr = fma(bdd[2], b3[2], fma(bdd[1], b3[1], bdd[0] * b3[0]));
        double b1c[3], b1cd[3];
        for (int i = 0; i < 3; ++i) {
            b1c[i] = b1d[i] - dp * b3[i];
            b1cd[i] = bdd[i] - ((dr * b3[i] + dq * b3[i]) + dp * b3d[i]);
        }
        double oc0 = b1c[1] * b1cd[2] - b1c[2] * b1cd[1];
        double oc1 = b1c[2] * b1cd[0] - b1c[0] * b1cd[2];
        double oc2 = b1c[0] * b1cd[1] - b1c[1] * b1cd[0];
        Wd[0] = 0; Wd[1] = 0; Wd[2] = fma(b3[2], oc2, fma(b3[1], oc1, b3[0] * oc0));
    }
    return 0;
}

/* ---- Philox4x32-10 ----------------------------------------------------------------------------------- */

void qo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
