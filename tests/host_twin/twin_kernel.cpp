// twin_kernel.cpp -- TEST INFRASTRUCTURE.  The REAL step kernel qr::k_step (gym_rotor_b200/csrc/qr_kernels.cuh: lane
// state machine, stash, fetch-ahead, reset queue and parked reset call, observation paths, statistics) compiled for
// the host and run on the one-warp SIMT emulator of simt.h, warp after warp, over plain host arrays that have the
// layout of the device buffers.  tests/test_host_twin_kernel.py drives it against the golden vectors and the oracle,
// so that the kernel's control logic can be exercised -- and a divergent warp collective caught -- without a GPU.
// Nothing in the package loads this: the product has no CPU path.
#include "cuda_shim.h"
#include "simt.h"
#define __launch_bounds__(...)
#ifndef __grid_constant__
#define __grid_constant__
#endif
#include "qr_kernels.cuh"
#include "../../include/quadrotor_b200.h"

namespace qr {
alignas(16) unsigned char smem_raw[256 * 1024];   // the kernel's `extern __shared__` array
}

namespace {
using namespace qr;

template <typename T> void fill_const(EnvConst<T>& e, const qr_config& c)   // as make_args() in quadrotor_b200.cu
{
    memset(&e, 0, sizeof(e));
    e.dt = (T)c.dt; e.g = (T)c.g; e.rtol = (T)c.rtol; e.atol = (T)c.atol;
    e.x_lim = (T)c.x_lim; e.v_lim = (T)c.v_lim; e.W_lim = (T)c.W_lim; e.eIx_lim = (T)c.eIx_lim; e.eIb1_lim = (T)c.eIb1_lim;
    e.sat = (T)c.sat_sigma; e.alpha = (T)c.alpha; e.beta = (T)c.beta; e.min_force = (T)c.min_force; e.euler_lim = (T)c.euler_lim_deg;
    e.inv_x_lim = (T)(1.0 / c.x_lim); e.inv_v_lim = (T)(1.0 / c.v_lim); e.inv_W_lim = (T)(1.0 / c.W_lim);
    e.inv_eIx_lim = (T)(1.0 / c.eIx_lim); e.inv_eIb1_lim = (T)(1.0 / c.eIb1_lim);
    e.nCx = (float)(-c.Cx); e.nCIx = (float)(-c.CIx); e.nCv = (float)(-c.Cv); e.nCb1 = (float)(-c.Cb1);
    e.nCIb1 = (float)(-c.CIb1); e.nCW = (float)(-c.CW); e.nCw12 = (float)(-c.Cw12); e.nCW3 = (float)(-c.CW3);
    e.Cx = c.Cx; e.Cv = c.Cv; e.Cb1 = c.Cb1; e.CW = c.CW;
    e.rmin = c.reward_min; e.rmin1 = c.reward_min_1; e.rmin2 = c.reward_min_2; e.udm = c.udm_pct;
    e.slope = 1.0 / (0.0 - c.reward_min); e.slope1 = 1.0 / (0.0 - c.reward_min_1); e.slope2 = 1.0 / (0.0 - c.reward_min_2);
    e.mode = c.mode; e.integrator = c.integrator; e.autoreset = c.autoreset; e.goal_mode = c.goal_mode;
    e.env_type = c.env_type; e.max_episode_steps = c.max_episode_steps; e.diagnostics = c.reserved0;
    e.round_returns = c.round_returns;
}

template <typename T, int MODE, bool MULTI, bool GOAL1, bool POLICY = false> void lane_body(void* p) { k_step<T, MODE, MULTI, GOAL1, POLICY>(*(const StepArgs<T>*)p); }

}  // namespace

// The host arrays a launch works on: same layouts as qr_buffers (state [18][n] ... obs [n][O] ...), element type per dtype.
struct tw_arrays {
    void *state, *integ, *params, *goal, *traj;
    float* obs; void* reward; uint8_t *done, *terminated, *truncated; float* final_obs;
    int32_t* nfev; uint8_t* status; void* ep_return; int32_t* ep_length; uint32_t* ep_index; double* stats;
    const void* actions; int act_f32;
    float* obs_roll; void* reward_roll; uint8_t* done_roll;
};

extern "C" int tw_kstep(const qr_config* cfg, const tw_arrays* b, int64_t env_lo, int64_t env_hi, int n_steps, int warps, int policy)
{
    unsigned long long tile_counter[2] = {0, 0};
    const bool multi = n_steps > 1 || policy || b->obs_roll || b->reward_roll || b->done_roll, goal1 = cfg->goal_mode == QR_GOAL_TRAJ_MODE0;   // as launch_step()
    if (policy && cfg->mode == QR_MODE_QUAD) return -2;
    if (warps < 1 || warps > 12) return -1;
    blockDim.x = (unsigned)warps * 32; gridDim.x = 1; blockIdx.x = 0;
#define TW_FILL(T)                                                                                                            \
    StepArgs<T> a; memset(&a, 0, sizeof(a)); fill_const<T>(a.c, *cfg);                                                       \
    a.n = cfg->n_envs; a.env_lo = env_lo; a.env_hi = env_hi; a.env_id_offset = cfg->env_id_offset;                            \
    a.key0 = (uint32_t)cfg->seed; a.key1 = (uint32_t)(cfg->seed >> 32); a.tile_counter = tile_counter;                       \
    a.state = (T*)b->state; a.integ = (T*)b->integ; a.params = (T*)b->params; a.goal = (T*)b->goal; a.traj = (T*)b->traj;    \
    a.obs = b->obs; a.reward = (T*)b->reward; a.done = b->done; a.terminated = b->terminated; a.truncated = b->truncated;    \
    a.final_obs = b->final_obs; a.nfev = b->nfev; a.status = b->status; a.ep_return = (T*)b->ep_return;                      \
    a.ep_length = b->ep_length; a.ep_index = b->ep_index; a.stats = b->stats;                                                \
    a.actions = b->actions; a.act_f32 = b->act_f32; a.n_steps = n_steps;                                                      \
    a.obs_roll = b->obs_roll; a.reward_roll = (T*)b->reward_roll; a.done_roll = b->done_roll;
#define TW_RUN(T, MODE)                                                                                                       \
    {                                                                                                                         \
        void (*body)(void*) = policy ? (goal1 ? lane_body<T, MODE, true, true, true> : lane_body<T, MODE, true, false, true>)   \
                            : multi ? (goal1 ? lane_body<T, MODE, true, true> : lane_body<T, MODE, true, false>)             \
                                    : (goal1 ? lane_body<T, MODE, false, true> : lane_body<T, MODE, false, false>);           \
        for (int w = 0; w < warps; ++w) simt::run_warp(body, &a, (unsigned)w * 32);                                           \
    }
    if (cfg->dtype == QR_F64) {
        TW_FILL(double)
        if (cfg->mode == QR_MODE_COUPLED) TW_RUN(double, 1) else if (cfg->mode == QR_MODE_DECOUPLED) TW_RUN(double, 2)
        else { void (*body)(void*) = multi ? lane_body<double, 0, true, false> : lane_body<double, 0, false, false>;
               for (int w = 0; w < warps; ++w) simt::run_warp(body, &a, (unsigned)w * 32); }
    } else {
        TW_FILL(float)
        if (cfg->mode == QR_MODE_COUPLED) TW_RUN(float, 1) else if (cfg->mode == QR_MODE_DECOUPLED) TW_RUN(float, 2)
        else { void (*body)(void*) = multi ? lane_body<float, 0, true, false> : lane_body<float, 0, false, false>;
               for (int w = 0; w < warps; ++w) simt::run_warp(body, &a, (unsigned)w * 32); }
    }
    return 0;
}

// The per-thread companion kernels (no warp collectives): k_reset (0), k_init_goal (1), k_goal_update (2),
// k_norm_error_state (3), run thread by thread over [env_lo, env_hi).
extern "C" int tw_companion(const qr_config* cfg, const tw_arrays* b, int which, const uint8_t* mask, int env_type)
{
    unsigned long long tile_counter[2] = {0, 0};
    const int64_t env_lo = 0, env_hi = cfg->n_envs;
    const int n_steps = 1;
    blockDim.x = QR_BLOCK; gridDim.x = (unsigned)((cfg->n_envs + QR_BLOCK - 1) / QR_BLOCK);
#define TW_COMPANION(T)                                                                            \
    {                                                                                              \
        TW_FILL(T)                                                                                 \
        for (unsigned blk = 0; blk < gridDim.x; ++blk)                                             \
            for (unsigned t = 0; t < (unsigned)QR_BLOCK; ++t) {                                    \
                blockIdx.x = blk; threadIdx.x = t;                                                 \
                if (which == 0) k_reset<T>(a, mask, env_type);                                     \
                else if (which == 1) k_init_goal<T>(a, mask);                                      \
                else if (which == 2) k_goal_update<T>(a);                                          \
                else k_norm_error_state<T>(a, mask);                                               \
            }                                                                                      \
    }
    if (cfg->dtype == QR_F64) TW_COMPANION(double) else TW_COMPANION(float)
    blockIdx.x = 0; threadIdx.x = 0;
    return 0;
}

// the compiled actors on one observation row (what k_actor_td3 computes per env): mode 1 -> 4 actions, mode 2 -> 5
extern "C" void tw_actor(int mode, const float* obs, float* act)
{
    if (mode == 1) actor_td3_mono(obs, act);
    else { actor_td3_modul1(obs, act); actor_td3_modul2(obs + 15, act + 4); }
}

// row stride of the obs / final_obs arrays this build of the kernels expects (rows are padded to a multiple of 4 floats)
extern "C" int tw_obs_stride(int mode) { return obs_stride_of(mode == 1 ? 23 : 18); }
