/* quadrotor_b200.h -- C ABI of the B200-native batched quadrotor simulator.
 *
 * Drop-in boundary for ONE hot path of fdcl-gwu/gym-rotor: env.step() (+ reset / get_norm_error_state)
 * of Quad-v0, CoupledWrapper and DecoupledWrapper.  The reference has no FFI: its boundary is a
 * duck-typed Python object (main.py:42,52,126-129,145-147,164,226-230).  Each entry point below names
 * the reference interface it replaces; INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - plain C: opaque handle, plain pointers and sizes, int return codes (0 = QR_OK), no torch types;
 *   - one handle per device shard; all calls are stream-ordered on the `stream` argument
 *     (a cudaStream_t passed as void*; NULL = the legacy default stream); no internal synchronisation
 *     except in the *_host calls, which return when the host buffers are valid;
 *   - a handle is not thread-safe; distinct handles are independent;
 *   - the library owns the per-env device buffers (qr_get_buffers exposes them, zero copy);
 *     the action buffer passed to qr_step is caller-owned device memory.
 *
 * Device layouts (T = float or double according to qr_config.dtype; N = n_envs)
 *   state  [18][N] T   x(3) | v(3) | R column-major = b1|b2|b3 (9) | W(3)      (quad.py:146)
 *   integ  [ 8][N] T   eIx.error(3) | eIx.integrand(3) | eIb1.error | eIb1.integrand (quad_utils.py:38-63)
 *   params [ 6][N] T   m | d | J1(=J2) | J3 | c_tf | c_tw                      (quad.py:359-404)
 *   goal   [12][N] T   xd(3) | vd(3) | b1d(3) | Wd(3)                          (quad.py:413-418)
 *   obs    [N][O] f32  COUPLED O=23: ex eIx ev R(9) eb1 eIb1 eW ; DECOUPLED O=18: obs1[15] | obs2[3]
 *                      (quad.py:453-464, wrapper_utils.py:3-28); QUAD O=18: next state cast to f32
 *   reward [N][G] T    G = 1, or 2 for DECOUPLED; [0,1] or -1 on crash        (quad.py:150-166)
 *   done   [N][G] u8   per-agent termination                                   (coupled:95-110, decoupled:116-140)
 *   actions[N][A] f32|f64  A = 4, or 5 for DECOUPLED; normalised to [-1,1]     (coupled:44-53, decoupled:49-59)
 */
#ifndef QUADROTOR_B200_H
#define QUADROTOR_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QR_ABI_VERSION 2

enum { QR_OK = 0, QR_ERR_INVALID = 1, QR_ERR_CUDA = 2, QR_ERR_NOMEM = 3, QR_ERR_NO_DEVICE = 4 };
enum { QR_MODE_QUAD = 0, QR_MODE_COUPLED = 1, QR_MODE_DECOUPLED = 2 };   /* Quad-v0 | CoupledWrapper | DecoupledWrapper */
enum { QR_F32 = 0, QR_F64 = 1 };
/* act_dtype of qr_rollout only: no action array, the reference's shipped TD3 actor is evaluated inside the kernel */
enum { QR_ACT_POLICY = 2 };
enum { QR_INT_DOP853 = 0, QR_INT_EULER = 1 };                            /* quad.py:62 */
enum { QR_ENV_TRAIN = 0, QR_ENV_EVAL = 1 };                              /* reset(env_type=...) quad.py:171 */
/* set_goal_state | on-device trajectory_generator: mode 0 (idle, evaluated inside qr_step), 1 hover, 5 circle, 6 figure
 * eight, 2 take-off, 3 land, 4 stay (utils/trajectory_generator.py:113-173, 252-505).  The last three cannot be
 * combined with autoreset (reset such envs with qr_reset + qr_init_goal). */
enum { QR_GOAL_EXTERNAL = 0, QR_GOAL_TRAJ_MODE0 = 1, QR_GOAL_TRAJ_HOVER = 2, QR_GOAL_TRAJ_CIRCLE = 3, QR_GOAL_TRAJ_EIGHT = 4,
       QR_GOAL_TRAJ_TAKEOFF = 5, QR_GOAL_TRAJ_LAND = 6, QR_GOAL_TRAJ_STAY = 7 };
/* per-env status bits (the reference raises / ignores sol.status instead: coupled:63-64) */
enum { QR_ST_NONFINITE = 1, QR_ST_TOO_SMALL_STEP = 2, QR_ST_SVD = 4 };
/* indices into the 20-double statistics vector of qr_stats.  BENCH_REWARD: sum over env-steps of benchmark_reward_func
 * (utils/utils.py:21-47) on the step's observation; SOLVED_AT_LIMIT: episodes that hit max_episode_steps with
 * |ex| <= 0.03 m and reward != -1 -- what the trainer relabels done_n[0] = True (main.py:169-173). */
enum {
    QR_STAT_EPISODES = 0, QR_STAT_RETURN0 = 1, QR_STAT_RETURN1 = 2, QR_STAT_LENGTH = 3, QR_STAT_CRASHED = 4,
    QR_STAT_TRUNCATED = 5, QR_STAT_RETURN0_SQ = 6, QR_STAT_STEPS = 7, QR_STAT_BAD_STATUS = 8, QR_STAT_NFEV = 9,
    QR_STAT_ATTEMPTS_1 = 10, QR_STAT_ATTEMPTS_2 = 11, QR_STAT_ATTEMPTS_3 = 12, QR_STAT_ATTEMPTS_4P = 13,
    QR_STAT_REWARD0 = 14, QR_STAT_SO3_PROJECTIONS = 15, QR_STAT_BENCH_REWARD = 16, QR_STAT_SOLVED_AT_LIMIT = 17,
    QR_NUM_STATS = 20                                                      /* 18, 19: reserved (zero) */
};

typedef struct qr_handle qr_handle;

/* Everything the reference reads from argparse defaults and constructor constants (args_parse.py:14-35,
 * quad.py:28-41,60-61,75-91,104-107, coupled:21-24), plus what is new for a batched device env. */
typedef struct qr_config {
    int64_t n_envs;             /* envs owned by this handle (this GPU's shard), 1 .. 2^30 */
    int64_t env_id_offset;      /* global id of local env 0: Philox streams do not depend on the sharding */
    uint64_t seed;
    int32_t mode;               /* QR_MODE_* */
    int32_t dtype;              /* QR_F32 | QR_F64: arithmetic and storage type of state/integ/params/goal/reward */
    int32_t integrator;         /* QR_INT_* (EULER only meaningful for QR_MODE_QUAD, quad.py:252-262) */
    int32_t autoreset;          /* 0: like the reference, step() never resets (quad.py:168); 1: reset in-kernel */
    int32_t goal_mode;          /* QR_GOAL_* */
    int32_t env_type;           /* QR_ENV_* used by in-kernel auto resets */
    int32_t max_episode_steps;  /* truncation limit (main.py:169, args_parse.py:16); 0 = none */
    int32_t reserved0;          /* diagnostics (default 1): write nfev per env and keep the per-step statistics ATTEMPTS_*, REWARD0,
                                 * BENCH_REWARD, SO3_PROJECTIONS (0: they stay zero; the mean attempt count is (NFEV / STEPS - 2) / 12) */
    int32_t round_returns;      /* 1: running episode returns are rounded to 4 decimals after every step, as the trainer
                                 * keeps them (main.py:180); 0: plain sums */
    int32_t reserved1;
    double dt, g, rtol, atol;
    double x_lim, v_lim, W_lim, eIx_lim, eIb1_lim, sat_sigma, alpha, beta;
    double Cx, CIx, Cv, Cb1, CIb1, CW, Cw12, CW3;
    double reward_min, reward_min_1, reward_min_2;
    double min_force, euler_lim_deg, udm_pct;
} qr_config;

/* zero-copy views of the library-owned device buffers (see layouts above) */
typedef struct qr_buffers {
    void* state; void* integ; void* params; void* goal;
    float* obs; void* reward; uint8_t* done;
    uint8_t* terminated;        /* [N] any(done) over agents (main.py:212) */
    uint8_t* truncated;         /* [N] episode hit max_episode_steps this step */
    float* final_obs;           /* [N][O] terminal observation of envs that were auto-reset this step */
    int32_t* nfev;              /* [N] RHS evaluations of the last step as scipy counts them: 2 + 12*attempts */
    uint8_t* status;            /* [N] QR_ST_* bits, sticky until qr_reset */
    void* ep_return;            /* [G][N] T running episode return */
    int32_t* ep_length;         /* [N] steps in the running episode */
    uint32_t* ep_index;         /* [N] episode counter = Philox stream index */
    double* stats;              /* [QR_NUM_STATS] device accumulators */
    int32_t obs_dim, act_dim, n_agents, elem_size;
    int64_t n_envs;
    void* traj;                 /* [12][N] T trajectory-generator state: t | flags | centre(3) | theta_init | w_b1d | smooth | t_traj | b1d_dot(2) | - */
} qr_buffers;

/* Fills *c with the reference's defaults for `mode`/`dtype` (args_parse.py, quad.py:28-107). */
int qr_default_config(qr_config* c, int mode, int dtype);

/* Replaces the env constructors CoupledWrapper() / DecoupledWrapper() / QuadEnv() (main.py:42,52). */
int qr_create(const qr_config* c, int device, qr_handle** out);
int qr_destroy(qr_handle* h);
int qr_get_config(const qr_handle* h, qr_config* out);
int qr_get_buffers(qr_handle* h, qr_buffers* out);

/* env.reset(env_type) (coupled:27-41, quad.py:171-222): re-draws parameters and initial state with
 * Philox4x32-10 keyed by (seed, global env id, episode index), zeroes integrals.  mask: device u8[N] or
 * NULL (= all).  Like the reference it does NOT compute an observation. */
int qr_reset(qr_handle* h, const uint8_t* mask, int env_type, void* stream);

/* trajectory_generator.mark_traj_start + get_desired(mode 0) after a reset (main.py:127-128,227-229;
 * trajectory_generator.py:141-148,165-172): b1d = Rz(theta) [cos psi, sin psi, 0], xd = vd = 0.
 * For the other on-device goal modes: mark_traj_start + the first get_desired of that mode.  mask as in qr_reset. */
int qr_init_goal(qr_handle* h, const uint8_t* mask, void* stream);

/* One trajectory_generator.get_desired(env.get_current_state(), mode) + env.set_goal_state for goal_mode HOVER / CIRCLE /
 * EIGHT / TAKEOFF / LAND / STAY (utils/trajectory_generator.py:252-505, manual fallback 232-249) WITHOUT stepping (it
 * advances the trajectory clock like every get_desired call).  qr_step / qr_rollout / qr_step_host make this call
 * themselves before every env.step, inside the step kernel, exactly where the trainer does (main.py:145-147); use this
 * entry point only to evaluate the generator on its own.  No-op for external goals and for mode 0. */
int qr_goal_update(qr_handle* h, void* stream);

/* env.get_norm_error_state(framework) (quad.py:421-466): writes obs from the CURRENT state and goal and,
 * like the reference, advances the integral terms once (main.py:129,230,314). */
int qr_norm_error_state(qr_handle* h, const uint8_t* mask, void* stream);

/* env.step(action) (quad.py:142-168) for every env.  actions: device [N][A], act_dtype QR_F32 or QR_F64
 * (float32 actions make numpy compute the thrust in float32, coupled:46-48 -- reproduced).
 * Outputs land in the qr_buffers views. */
int qr_step(qr_handle* h, const void* actions, int act_dtype, void* stream);

/* agent.choose_action(obs, explor_noise_std=0) (algos/td3/td3.py:93-96) for the TD3 actors the reference ships
 * (models/TD3_*.pth, loaded at main.py:101-110): reads the handle's current observations, writes float32
 * actions [N][A] (device).  The actors are compiled from their effective weights (tools/gen_actor_kernels.py);
 * COUPLED uses the monolithic actor, DECOUPLED module 1 (f, tau) and module 2 (M3). */
int qr_policy_td3(qr_handle* h, float* actions, void* stream);

/* `n_steps` consecutive env.step() calls fused in one launch, state resident in registers.
 * actions: device [n_steps][N][A] or NULL = U(-1,1) actions drawn in-kernel with Philox (the synthetic
 * random-action workload).  act_dtype = QR_ACT_POLICY (actions = NULL): every step's action is the shipped TD3 actor
 * (as qr_policy_td3) evaluated in the kernel on the env's latest observation -- the evaluation loop of
 * main.py:304-365 (obs -> agent.choose_action -> env.step) in one launch; the observation buffer must hold the
 * current observations (qr_norm_error_state after a reset).  obs_out/reward_out/done_out: device
 * [n_steps][N][..] or NULL to keep only the last step's outputs in the qr_buffers views.  1 <= n_steps <= 32767.
 * With autoreset, an env whose episode ends inside the launch is reset there (batched per warp) and goes on stepping:
 * its row of obs_out at that sub-step holds the first observation of the new episode, like qr_step's obs.
 * All on-device goal modes are evaluated inside the launch before every sub-step. */
int qr_rollout(qr_handle* h, int n_steps, const void* actions, int act_dtype, float* obs_out, void* reward_out,
               uint8_t* done_out, void* stream);

/* Same as qr_step with HOST buffers (pinned or pageable): copies actions host->device, steps, copies
 * obs (dense [N][obs_dim]) / reward / done device->host and returns when they are valid.  This is the call the
 * end-to-end benchmark times.  Any output pointer may be NULL.  The work is pipelined in chunks over two streams
 * owned by the handle; it is ordered after everything enqueued on `stream` so far (pass the stream of the caller's
 * earlier qr_* calls; NULL = the legacy default stream).  (Eight chunks; the environment variable QR_HOST_CHUNKS, read
 * once, overrides the count for measurements: throughput is flat from 4 to 16, profiles/r02/r02be_e2e_chunks.txt.) */
int qr_step_host(qr_handle* h, const void* actions_host, int act_dtype, float* obs_host, void* reward_host,
                 uint8_t* done_host, void* stream);

/* Host-layout access for C callers and tests: row-major [N][18] / [N][8] / [N][6] / [N][12] doubles
 * (env.state, env.eIx/eIb1, env.m/J/..., env.xd/vd/b1d/Wd).  NULL pointers are skipped.  Synchronous. */
int qr_set_state_host(qr_handle* h, const double* state, const double* integ, const double* params, const double* goal);
int qr_get_state_host(qr_handle* h, double* state, double* integ, double* params, double* goal);

/* Copies the QR_NUM_STATS statistics accumulators to host (synchronous on `stream`); reset_after != 0 zeroes them.
 * Multi-GPU callers all-reduce this vector (the only collective of the path). */
int qr_stats(qr_handle* h, double* out, int reset_after, void* stream);

/* Number of kernels this library has launched so far in this process (for the benchmark's gpu_launches). */
int64_t qr_launch_count(void);
const char* qr_last_error(void);
int qr_abi_version(void);
/* Row stride, in floats, of the obs and final_obs buffers of qr_get_buffers (= obs_dim unless the library was built
 * with padded observation rows). */
int qr_obs_stride(const qr_handle* h);

#ifdef __cplusplus
}
#endif
#endif
