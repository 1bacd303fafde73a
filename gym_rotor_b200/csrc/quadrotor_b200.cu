// quadrotor_b200.cu -- C ABI (include/quadrotor_b200.h) over the sm_100a kernels in qr_kernels.cuh.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
//        (gym_rotor_b200/build.py).  No torch, no CPU fallback: every entry point needs a CUDA device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <atomic>
#include <new>
#include <string>
#include <vector>

#include "../../include/quadrotor_b200.h"
#include "qr_kernels.cuh"

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

#define QR_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return fail(QR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));               \
    } while (0)

}  // namespace

struct qr_handle {
    qr_config cfg;
    int device;
    int elem;  // sizeof(T)
    int O, A, G;
    // device buffers
    void *state, *integ, *params, *goal, *reward, *ep_return, *traj;
    float *obs, *final_obs;
    uint8_t *done, *terminated, *truncated, *status;
    int32_t *nfev, *ep_length;
    uint32_t* ep_index;
    double* stats;
    // staging for the *_host calls
    void* d_actions; size_t d_actions_bytes;
    double* d_stage; size_t d_stage_bytes;
    float* d_obs_dense; size_t d_obs_dense_bytes;   // qr_step_host: observation rows without the padding, chunk by chunk
    cudaStream_t io_stream[2];      // qr_step_host pipelines its chunks over these two
    cudaEvent_t io_event[3];        // [0]: the caller's stream at entry; [1], [2]: the io streams at exit
    unsigned long long* tile_counter;   // [3]: launches on the caller's stream, on io_stream[0], on io_stream[1]
    int num_sms; int smem_optin; int attr_set[16];
};

namespace {

template <typename T> qr::StepArgs<T> make_args(const qr_handle* h)
{
    qr::StepArgs<T> a;
    memset(&a, 0, sizeof(a));
    const qr_config& c = h->cfg;
    a.c.dt = (T)c.dt; a.c.g = (T)c.g; a.c.rtol = (T)c.rtol; a.c.atol = (T)c.atol;
    a.c.x_lim = (T)c.x_lim; a.c.v_lim = (T)c.v_lim; a.c.W_lim = (T)c.W_lim;
    a.c.eIx_lim = (T)c.eIx_lim; a.c.eIb1_lim = (T)c.eIb1_lim; a.c.sat = (T)c.sat_sigma;
    a.c.alpha = (T)c.alpha; a.c.beta = (T)c.beta; a.c.min_force = (T)c.min_force; a.c.euler_lim = (T)c.euler_lim_deg;
    a.c.inv_x_lim = (T)(1.0 / c.x_lim); a.c.inv_v_lim = (T)(1.0 / c.v_lim); a.c.inv_W_lim = (T)(1.0 / c.W_lim);
    a.c.inv_eIx_lim = (T)(1.0 / c.eIx_lim); a.c.inv_eIb1_lim = (T)(1.0 / c.eIb1_lim);
    a.c.nCx = (float)(-c.Cx); a.c.nCIx = (float)(-c.CIx); a.c.nCv = (float)(-c.Cv); a.c.nCb1 = (float)(-c.Cb1);
    a.c.nCIb1 = (float)(-c.CIb1); a.c.nCW = (float)(-c.CW); a.c.nCw12 = (float)(-c.Cw12); a.c.nCW3 = (float)(-c.CW3);
    a.c.Cx = c.Cx; a.c.Cv = c.Cv; a.c.Cb1 = c.Cb1; a.c.CW = c.CW;
    a.c.rmin = c.reward_min; a.c.rmin1 = c.reward_min_1; a.c.rmin2 = c.reward_min_2; a.c.udm = c.udm_pct;
    a.c.slope = 1.0 / (0.0 - c.reward_min); a.c.slope1 = 1.0 / (0.0 - c.reward_min_1); a.c.slope2 = 1.0 / (0.0 - c.reward_min_2);
    a.c.mode = c.mode; a.c.integrator = c.integrator; a.c.autoreset = c.autoreset; a.c.goal_mode = c.goal_mode;
    a.c.env_type = c.env_type; a.c.max_episode_steps = c.max_episode_steps; a.c.diagnostics = c.reserved0;
    a.c.round_returns = c.round_returns;
    a.n = c.n_envs; a.env_lo = 0; a.env_hi = c.n_envs; a.env_id_offset = c.env_id_offset;
    a.key0 = (uint32_t)c.seed; a.key1 = (uint32_t)(c.seed >> 32);
    a.state = (T*)h->state; a.integ = (T*)h->integ; a.params = (T*)h->params; a.goal = (T*)h->goal; a.traj = (T*)h->traj;
    a.obs = h->obs; a.reward = (T*)h->reward; a.done = h->done; a.terminated = h->terminated; a.truncated = h->truncated;
    a.final_obs = h->final_obs; a.nfev = h->nfev; a.status = h->status; a.ep_return = (T*)h->ep_return;
    a.ep_length = h->ep_length; a.ep_index = h->ep_index; a.stats = h->stats;
    a.n_steps = 1;
    return a;
}

inline unsigned blocks_for(int64_t n) { return (unsigned)((n + qr::QR_BLOCK - 1) / qr::QR_BLOCK); }

template <typename T>
int launch_step(qr_handle* h, int64_t lo, int64_t hi, const void* actions, int act_dtype, int n_steps, float* obs_roll,
                void* reward_roll, uint8_t* done_roll, cudaStream_t s)
{
    if (hi <= lo) return QR_OK;
    qr::StepArgs<T> a = make_args<T>(h);
    a.env_lo = lo; a.env_hi = hi;
    const bool policy = act_dtype == QR_ACT_POLICY;   // actions from the shipped actor, evaluated inside the kernel
    a.actions = policy ? nullptr : actions; a.act_f32 = (act_dtype == QR_F32); a.n_steps = n_steps;
    a.obs_roll = obs_roll; a.reward_roll = (T*)reward_roll; a.done_roll = done_roll;
    // launches that may overlap must not share a tile counter: the two streams of qr_step_host have their own.  (Launches
    // of one handle on several CALLER streams at once are not supported: they would race on the env state anyway.)
    a.tile_counter = h->tile_counter + ((s == h->io_stream[0]) ? 1 : (s == h->io_stream[1]) ? 2 : 0);
    QR_CUDA(cudaMemsetAsync(a.tile_counter, 0, sizeof(unsigned long long), s));
    // persistent warps: one CTA per SM, as many warps as the stage storage in shared memory allows
    const size_t per_warp = qr::warp_smem<T>::bytes;
    int warps = (int)((size_t)h->smem_optin / per_warp);
    if (warps > qr::step_threads<T>::value / 32) warps = qr::step_threads<T>::value / 32;
    if (warps < 1) return fail(QR_ERR_CUDA, "not enough shared memory per block for the step kernel");
    const int64_t ntiles = (hi - lo + 31) / 32;
    int64_t grid = (ntiles + warps - 1) / warps;
    if (grid > h->num_sms) grid = h->num_sms;
    if (ntiles < (int64_t)warps) warps = (int)ntiles;
    const size_t smem = per_warp * warps;
    // the multi-step kernel (the env keeps stepping in its lane) also serves every launch that writes the caller's rollout
    // storage -- the single-step kernel has no code for it -- and the policy variants, which exist in that flavour only
    const bool multi = n_steps > 1 || policy || obs_roll || reward_roll || done_roll;
    const bool goal1 = h->cfg.goal_mode == QR_GOAL_TRAJ_MODE0;   // only with a wrapper mode (checked in qr_create)
    // the kernel for this configuration (compiled in its own translation unit, qr_step_tu.cu)
    const qr::step_kernel_t<T> kern = qr::step_kernel<T>(h->cfg.mode, multi, goal1, policy && h->cfg.mode != QR_MODE_QUAD);
    const int attr_idx = (policy ? 8 : 0) + (sizeof(T) == 8 ? 4 : 0) + (multi ? 2 : 0) + (goal1 ? 1 : 0);
    if (!h->attr_set[attr_idx]) {
        QR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)h->smem_optin / per_warp * per_warp)));
        h->attr_set[attr_idx] = 1;
    }
    kern<<<(unsigned)grid, warps * 32, smem, s>>>(a);
    g_launches++;
    QR_CUDA(cudaGetLastError());
    return QR_OK;
}

int check(const qr_handle* h)
{
    if (!h) return fail(QR_ERR_INVALID, "null handle");
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return fail(QR_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return QR_OK;
}

}  // namespace

extern "C" {

int qr_abi_version(void) { return QR_ABI_VERSION; }
int qr_obs_stride(const qr_handle* h) { return h ? obs_stride_of(h->O) : 0; }
const char* qr_last_error(void) { return g_err.c_str(); }
int64_t qr_launch_count(void) { return (int64_t)g_launches.load(); }

int qr_default_config(qr_config* c, int mode, int dtype)
{
    if (!c || mode < 0 || mode > 2 || (dtype != QR_F32 && dtype != QR_F64)) return fail(QR_ERR_INVALID, "qr_default_config: bad arguments");
    memset(c, 0, sizeof(*c));
    c->n_envs = 1; c->env_id_offset = 0; c->seed = 1992;      /* args_parse.py:6 */
    c->mode = mode; c->dtype = dtype; c->integrator = QR_INT_DOP853;
    c->autoreset = 0; c->goal_mode = QR_GOAL_EXTERNAL; c->env_type = QR_ENV_TRAIN;
    c->max_episode_steps = 0; c->reserved0 = 1;               /* reserved0 = diagnostics (write nfev) */
    c->round_returns = 0; c->reserved1 = 0;
    c->dt = 1. / 200; c->g = 9.81; c->rtol = 1e-3; c->atol = 1e-6;
    c->x_lim = 1.0; c->v_lim = 4.0; c->W_lim = 2 * 3.14159265358979323846;
    c->eIx_lim = 3.0; c->eIb1_lim = 3.0; c->sat_sigma = 1.; c->alpha = 0.01; c->beta = 0.05;
    c->Cx = 6.0; c->CIx = 0.1; c->Cv = 0.4; c->Cw12 = 0.6; c->Cb1 = 6.0; c->CIb1 = 0.1; c->CW3 = 0.1;
    c->CW = c->Cw12;
    c->reward_min = -ceil(c->Cx + c->CIx + c->Cv + c->Cb1 + c->CIb1 + c->CW);
    c->reward_min_1 = -ceil(c->Cx + c->CIx + c->Cv + c->Cw12);
    c->reward_min_2 = -ceil(c->Cb1 + c->CW3 + c->CIb1);
    c->min_force = 0.5; c->euler_lim_deg = 85; c->udm_pct = 10;
    return QR_OK;
}

int qr_create(const qr_config* c, int device, qr_handle** out)
{
    if (!c || !out) return fail(QR_ERR_INVALID, "qr_create: null argument");
    if (c->n_envs <= 0) return fail(QR_ERR_INVALID, "qr_create: n_envs must be positive");
    if (c->mode < 0 || c->mode > 2) return fail(QR_ERR_INVALID, "qr_create: bad mode");
    if (c->dtype != QR_F32 && c->dtype != QR_F64) return fail(QR_ERR_INVALID, "qr_create: bad dtype");
    // (the step kernel keeps env indices in 32-bit registers; 2^30 envs are 393 GB of env state in float32 mode anyway)
    if (c->n_envs > ((int64_t)1 << 30)) return fail(QR_ERR_INVALID, "qr_create: n_envs must not exceed 2^30 per handle");
    if (c->integrator == QR_INT_EULER && c->mode != QR_MODE_QUAD)
        return fail(QR_ERR_INVALID, "qr_create: the Euler integrator exists only for the base Quad-v0 env (quad.py:252)");
    if (c->goal_mode < QR_GOAL_EXTERNAL || c->goal_mode > QR_GOAL_TRAJ_STAY) return fail(QR_ERR_INVALID, "qr_create: bad goal_mode");
    if (c->autoreset && c->goal_mode >= QR_GOAL_TRAJ_TAKEOFF)
        return fail(QR_ERR_INVALID, "qr_create: autoreset is not available with the take-off / land / stay goal modes (reset them with qr_reset + qr_init_goal)");
    if (c->goal_mode != QR_GOAL_EXTERNAL && c->mode == QR_MODE_QUAD)
        return fail(QR_ERR_INVALID, "qr_create: on-device goal generation needs a wrapper mode");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(QR_ERR_NO_DEVICE, std::string("no CUDA device: this library has no CPU path (") + cudaGetErrorString(e) + ")");
    if (device < 0 || device >= ndev) return fail(QR_ERR_INVALID, "qr_create: bad device index");
    QR_CUDA(cudaSetDevice(device));
    qr_handle* h = new (std::nothrow) qr_handle();
    if (!h) return fail(QR_ERR_NOMEM, "qr_create: out of host memory");
    memset(h, 0, sizeof(*h));
    h->cfg = *c; h->device = device;
    h->elem = (c->dtype == QR_F64) ? 8 : 4;
    h->O = (c->mode == QR_MODE_COUPLED) ? 23 : 18;
    h->A = (c->mode == QR_MODE_DECOUPLED) ? 5 : 4;
    h->G = (c->mode == QR_MODE_DECOUPLED) ? 2 : 1;
    const size_t n = (size_t)c->n_envs, E = (size_t)h->elem;
    struct { void** p; size_t bytes; } allocs[] = {
        {&h->state, 18 * n * E}, {&h->integ, 8 * n * E}, {&h->params, 6 * n * E}, {&h->goal, 12 * n * E}, {&h->traj, 12 * n * E},
        {(void**)&h->obs, n * obs_stride_of(h->O) * 4}, {&h->reward, n * h->G * E}, {(void**)&h->done, n * h->G},
        {(void**)&h->terminated, n}, {(void**)&h->truncated, n}, {(void**)&h->final_obs, n * obs_stride_of(h->O) * 4},
        {(void**)&h->nfev, n * 4}, {(void**)&h->status, n}, {&h->ep_return, 2 * n * E}, {(void**)&h->ep_length, n * 4},
        {(void**)&h->ep_index, n * 4}, {(void**)&h->stats, QR_NUM_STATS * sizeof(double)},
        {(void**)&h->tile_counter, 3 * sizeof(unsigned long long)}};
    for (auto& al : allocs) {
        cudaError_t ce = cudaMalloc(al.p, al.bytes);
        if (ce != cudaSuccess) {
            std::string msg = std::string("cudaMalloc: ") + cudaGetErrorString(ce);
            qr_destroy(h);
            return fail(QR_ERR_NOMEM, msg);
        }
        cudaMemset(*al.p, 0, al.bytes);
    }
    for (auto& st : h->io_stream) QR_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (auto& ev : h->io_event) QR_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    QR_CUDA(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
    QR_CUDA(cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    h->smem_optin -= 256;    // reserve
    // identity attitude, nominal parameters, b1d = e1: a defined state before the first reset
    {
        std::vector<double> st(18 * n, 0.0), par(6 * n), gl(12 * n, 0.0);
        const double nom[6] = {2.15, 0.23, 0.022, 0.035, 0.0135, 2.2};
        for (size_t i = 0; i < n; ++i) {
            st[18 * i + 6] = 1; st[18 * i + 10] = 1; st[18 * i + 14] = 1;
            for (int j = 0; j < 6; ++j) par[6 * i + j] = nom[j];
            gl[12 * i + 6] = 1;
        }
        int rc = qr_set_state_host(h, st.data(), nullptr, par.data(), gl.data());
        if (rc != QR_OK) { qr_destroy(h); return rc; }
    }
    QR_CUDA(cudaDeviceSynchronize());
    *out = h;
    return QR_OK;
}

int qr_destroy(qr_handle* h)
{
    if (!h) return QR_OK;
    cudaSetDevice(h->device);
    void* ptrs[] = {h->tile_counter, h->traj, h->state, h->integ, h->params, h->goal, h->obs, h->reward, h->done, h->terminated, h->truncated,
                    h->final_obs, h->nfev, h->status, h->ep_return, h->ep_length, h->ep_index, h->stats, h->d_actions, h->d_stage, h->d_obs_dense};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (auto st : h->io_stream) if (st) cudaStreamDestroy(st);
    for (auto ev : h->io_event) if (ev) cudaEventDestroy(ev);
    delete h;
    return QR_OK;
}

int qr_get_config(const qr_handle* h, qr_config* out)
{
    if (!h || !out) return fail(QR_ERR_INVALID, "qr_get_config: null argument");
    *out = h->cfg;
    return QR_OK;
}

int qr_get_buffers(qr_handle* h, qr_buffers* b)
{
    if (!h || !b) return fail(QR_ERR_INVALID, "qr_get_buffers: null argument");
    b->state = h->state; b->integ = h->integ; b->params = h->params; b->goal = h->goal;
    b->obs = h->obs; b->reward = h->reward; b->done = h->done; b->terminated = h->terminated; b->truncated = h->truncated;
    b->final_obs = h->final_obs; b->nfev = h->nfev; b->status = h->status; b->ep_return = h->ep_return;
    b->ep_length = h->ep_length; b->ep_index = h->ep_index; b->stats = h->stats;
    b->traj = h->traj;
    b->obs_dim = h->O; b->act_dim = h->A; b->n_agents = h->G; b->elem_size = h->elem; b->n_envs = h->cfg.n_envs;
    return QR_OK;
}

int qr_reset(qr_handle* h, const uint8_t* mask, int env_type, void* stream)
{
    int rc = check(h); if (rc) return rc;
    if (env_type != QR_ENV_TRAIN && env_type != QR_ENV_EVAL) return fail(QR_ERR_INVALID, "qr_reset: bad env_type");
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned nb = blocks_for(h->cfg.n_envs);
    if (h->cfg.dtype == QR_F64) qr::k_reset<double><<<nb, qr::QR_BLOCK, 0, s>>>(make_args<double>(h), mask, env_type);
    else qr::k_reset<float><<<nb, qr::QR_BLOCK, 0, s>>>(make_args<float>(h), mask, env_type);
    g_launches++;
    QR_CUDA(cudaGetLastError());
    return QR_OK;
}

int qr_init_goal(qr_handle* h, const uint8_t* mask, void* stream)
{
    int rc = check(h); if (rc) return rc;
    if (h->cfg.mode == QR_MODE_QUAD) return fail(QR_ERR_INVALID, "qr_init_goal: needs a wrapper mode");
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned nb = blocks_for(h->cfg.n_envs);
    if (h->cfg.dtype == QR_F64) qr::k_init_goal<double><<<nb, qr::QR_BLOCK, 0, s>>>(make_args<double>(h), mask);
    else qr::k_init_goal<float><<<nb, qr::QR_BLOCK, 0, s>>>(make_args<float>(h), mask);
    g_launches++;
    QR_CUDA(cudaGetLastError());
    return QR_OK;
}

int qr_goal_update(qr_handle* h, void* stream)
{
    int rc = check(h); if (rc) return rc;
    if (h->cfg.goal_mode < QR_GOAL_TRAJ_HOVER) return QR_OK;   // external goals / mode 0 (evaluated inside qr_step)
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned nb = blocks_for(h->cfg.n_envs);
    if (h->cfg.dtype == QR_F64) qr::k_goal_update<double><<<nb, qr::QR_BLOCK, 0, s>>>(make_args<double>(h));
    else qr::k_goal_update<float><<<nb, qr::QR_BLOCK, 0, s>>>(make_args<float>(h));
    g_launches++;
    QR_CUDA(cudaGetLastError());
    return QR_OK;
}

int qr_norm_error_state(qr_handle* h, const uint8_t* mask, void* stream)
{
    int rc = check(h); if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned nb = blocks_for(h->cfg.n_envs);
    if (h->cfg.dtype == QR_F64) qr::k_norm_error_state<double><<<nb, qr::QR_BLOCK, 0, s>>>(make_args<double>(h), mask);
    else qr::k_norm_error_state<float><<<nb, qr::QR_BLOCK, 0, s>>>(make_args<float>(h), mask);
    g_launches++;
    QR_CUDA(cudaGetLastError());
    return QR_OK;
}

int qr_policy_td3(qr_handle* h, float* actions, void* stream)
{
    int rc = check(h); if (rc) return rc;
    if (!actions) return fail(QR_ERR_INVALID, "qr_policy_td3: null actions");
    if (h->cfg.mode == QR_MODE_QUAD) return fail(QR_ERR_INVALID, "qr_policy_td3: the reference ships no actor for the base Quad-v0 env");
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned nb = blocks_for(h->cfg.n_envs);
    if (h->cfg.mode == QR_MODE_COUPLED) qr::k_actor_td3<1><<<nb, qr::QR_BLOCK, 0, s>>>(h->obs, actions, h->cfg.n_envs);
    else qr::k_actor_td3<2><<<nb, qr::QR_BLOCK, 0, s>>>(h->obs, actions, h->cfg.n_envs);
    g_launches++;
    QR_CUDA(cudaGetLastError());
    return QR_OK;
}

int qr_step(qr_handle* h, const void* actions, int act_dtype, void* stream)
{
    int rc = check(h); if (rc) return rc;
    if (!actions) return fail(QR_ERR_INVALID, "qr_step: null actions (use qr_rollout for in-kernel random actions)");
    if (act_dtype != QR_F32 && act_dtype != QR_F64) return fail(QR_ERR_INVALID, "qr_step: bad act_dtype");
    cudaStream_t s = (cudaStream_t)stream;
    if (h->cfg.dtype == QR_F64) return launch_step<double>(h, 0, h->cfg.n_envs, actions, act_dtype, 1, nullptr, nullptr, nullptr, s);
    return launch_step<float>(h, 0, h->cfg.n_envs, actions, act_dtype, 1, nullptr, nullptr, nullptr, s);
}

int qr_rollout(qr_handle* h, int n_steps, const void* actions, int act_dtype, float* obs_out, void* reward_out,
               uint8_t* done_out, void* stream)
{
    int rc = check(h); if (rc) return rc;
    if (n_steps <= 0 || n_steps > 32767) return fail(QR_ERR_INVALID, "qr_rollout: n_steps must be in 1 .. 32767");
    if (act_dtype == QR_ACT_POLICY) {
        if (actions) return fail(QR_ERR_INVALID, "qr_rollout: act_dtype QR_ACT_POLICY takes no action array");
        if (h->cfg.mode == QR_MODE_QUAD) return fail(QR_ERR_INVALID, "qr_rollout: the shipped actors exist for the wrapper modes only");
    } else if (actions && act_dtype != QR_F32 && act_dtype != QR_F64) return fail(QR_ERR_INVALID, "qr_rollout: bad act_dtype");
    cudaStream_t s = (cudaStream_t)stream;
    if (h->cfg.dtype == QR_F64) return launch_step<double>(h, 0, h->cfg.n_envs, actions, act_dtype, n_steps, obs_out, reward_out, done_out, s);
    return launch_step<float>(h, 0, h->cfg.n_envs, actions, act_dtype, n_steps, obs_out, reward_out, done_out, s);
}

int qr_step_host(qr_handle* h, const void* actions_host, int act_dtype, float* obs_host, void* reward_host, uint8_t* done_host,
                 void* stream)
{
    int rc = check(h); if (rc) return rc;
    if (!actions_host) return fail(QR_ERR_INVALID, "qr_step_host: null actions");
    if (act_dtype != QR_F32 && act_dtype != QR_F64) return fail(QR_ERR_INVALID, "qr_step_host: bad act_dtype");
    const int64_t n = h->cfg.n_envs;
    const size_t asz = (act_dtype == QR_F64) ? 8 : 4;
    const size_t abytes = (size_t)n * h->A * asz;
    if (h->d_actions_bytes < abytes) {
        if (h->d_actions) cudaFree(h->d_actions);
        h->d_actions = nullptr; h->d_actions_bytes = 0;
        QR_CUDA(cudaMalloc(&h->d_actions, abytes));
        h->d_actions_bytes = abytes;
    }
    const int OS = obs_stride_of(h->O);
    const bool dense_copy = obs_host && OS != h->O;   // padded device rows: compact them on the device, then ONE flat copy per chunk
    if (dense_copy && h->d_obs_dense_bytes < (size_t)n * h->O * 4) {
        if (h->d_obs_dense) cudaFree(h->d_obs_dense);
        h->d_obs_dense = nullptr; h->d_obs_dense_bytes = 0;
        QR_CUDA(cudaMalloc((void**)&h->d_obs_dense, (size_t)n * h->O * 4));
        h->d_obs_dense_bytes = (size_t)n * h->O * 4;
    }
    // order the pipeline after whatever the caller has in flight on its stream (a reset, a goal, an earlier step)
    QR_CUDA(cudaEventRecord(h->io_event[0], (cudaStream_t)stream));
    for (auto st : h->io_stream) QR_CUDA(cudaStreamWaitEvent(st, h->io_event[0], 0));
    // Chunked pipeline over the handle's two streams: copy-in, step and copy-out of chunk i overlap those of chunk i+1
    // through the copy engines; chunk boundaries are multiples of the block size so rows stay line aligned.
    // (QR_HOST_CHUNKS: measurement override, tools/e2e_chunks.sh)
    static const int64_t target_chunks = [] { const char* e = getenv("QR_HOST_CHUNKS"); const long v = e ? atol(e) : 0; return (int64_t)(v >= 1 && v <= 256 ? v : 8); }();
    int64_t chunk = ((n + target_chunks - 1) / target_chunks + qr::QR_BLOCK - 1) / qr::QR_BLOCK * qr::QR_BLOCK;
    if (chunk < 16384) chunk = 16384;
    int idx = 0;
    for (int64_t lo = 0; lo < n; lo += chunk, ++idx) {
        const int64_t hi = (lo + chunk < n) ? lo + chunk : n;
        cudaStream_t cs = h->io_stream[idx & 1];
        QR_CUDA(cudaMemcpyAsync((char*)h->d_actions + (size_t)lo * h->A * asz, (const char*)actions_host + (size_t)lo * h->A * asz,
                                (size_t)(hi - lo) * h->A * asz, cudaMemcpyHostToDevice, cs));
        if (h->cfg.dtype == QR_F64) rc = launch_step<double>(h, lo, hi, h->d_actions, act_dtype, 1, nullptr, nullptr, nullptr, cs);
        else rc = launch_step<float>(h, lo, hi, h->d_actions, act_dtype, 1, nullptr, nullptr, nullptr, cs);
        if (rc) return rc;
        if (obs_host) {
            const float* src = h->obs + (size_t)lo * OS;
            if (dense_copy) {
                const int64_t tot = (hi - lo) * h->O;
                qr::k_dense_rows<<<(unsigned)((tot + 255) / 256), 256, 0, cs>>>(src, h->d_obs_dense + (size_t)lo * h->O, hi - lo, h->O, OS);
                g_launches++;
                QR_CUDA(cudaGetLastError());
                src = h->d_obs_dense + (size_t)lo * h->O;
            }
            QR_CUDA(cudaMemcpyAsync(obs_host + (size_t)lo * h->O, src, (size_t)(hi - lo) * h->O * 4, cudaMemcpyDeviceToHost, cs));
        }
        if (reward_host) QR_CUDA(cudaMemcpyAsync((char*)reward_host + (size_t)lo * h->G * h->elem, (char*)h->reward + (size_t)lo * h->G * h->elem,
                                                 (size_t)(hi - lo) * h->G * h->elem, cudaMemcpyDeviceToHost, cs));
        if (done_host) QR_CUDA(cudaMemcpyAsync(done_host + (size_t)lo * h->G, h->done + (size_t)lo * h->G, (size_t)(hi - lo) * h->G, cudaMemcpyDeviceToHost, cs));
    }
    // later work on the caller's stream is ordered after the pipeline; the call itself returns when the host buffers are valid
    for (int i = 0; i < 2; ++i) {
        QR_CUDA(cudaEventRecord(h->io_event[1 + i], h->io_stream[i]));
        QR_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, h->io_event[1 + i], 0));
    }
    for (auto st : h->io_stream) QR_CUDA(cudaStreamSynchronize(st));
    return QR_OK;
}

static int ensure_stage(qr_handle* h, size_t bytes)
{
    if (h->d_stage_bytes >= bytes) return QR_OK;
    if (h->d_stage) cudaFree(h->d_stage);
    h->d_stage = nullptr; h->d_stage_bytes = 0;
    QR_CUDA(cudaMalloc((void**)&h->d_stage, bytes));
    h->d_stage_bytes = bytes;
    return QR_OK;
}

int qr_set_state_host(qr_handle* h, const double* state, const double* integ, const double* params, const double* goal)
{
    int rc = check(h); if (rc) return rc;
    const int64_t n = h->cfg.n_envs;
    rc = ensure_stage(h, (size_t)n * 18 * sizeof(double)); if (rc) return rc;
    struct { const double* src; void* dst; int C; } items[] = {{state, h->state, 18}, {integ, h->integ, 8}, {params, h->params, 6}, {goal, h->goal, 12}};
    for (auto& it : items) {
        if (!it.src) continue;
        QR_CUDA(cudaMemcpy(h->d_stage, it.src, (size_t)n * it.C * sizeof(double), cudaMemcpyHostToDevice));
        const int64_t tot = n * it.C;
        const unsigned nb = (unsigned)((tot + 255) / 256);
        if (h->cfg.dtype == QR_F64) qr::k_aos_to_soa<double><<<nb, 256>>>(h->d_stage, (double*)it.dst, n, it.C);
        else qr::k_aos_to_soa<float><<<nb, 256>>>(h->d_stage, (float*)it.dst, n, it.C);
        g_launches++;
        QR_CUDA(cudaGetLastError());
        QR_CUDA(cudaDeviceSynchronize());
    }
    return QR_OK;
}

int qr_get_state_host(qr_handle* h, double* state, double* integ, double* params, double* goal)
{
    int rc = check(h); if (rc) return rc;
    const int64_t n = h->cfg.n_envs;
    rc = ensure_stage(h, (size_t)n * 18 * sizeof(double)); if (rc) return rc;
    struct { double* dst; void* src; int C; } items[] = {{state, h->state, 18}, {integ, h->integ, 8}, {params, h->params, 6}, {goal, h->goal, 12}};
    QR_CUDA(cudaDeviceSynchronize());
    for (auto& it : items) {
        if (!it.dst) continue;
        const int64_t tot = n * it.C;
        const unsigned nb = (unsigned)((tot + 255) / 256);
        if (h->cfg.dtype == QR_F64) qr::k_soa_to_aos<double><<<nb, 256>>>((const double*)it.src, h->d_stage, n, it.C);
        else qr::k_soa_to_aos<float><<<nb, 256>>>((const float*)it.src, h->d_stage, n, it.C);
        g_launches++;
        QR_CUDA(cudaGetLastError());
        QR_CUDA(cudaMemcpy(it.dst, h->d_stage, (size_t)n * it.C * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return QR_OK;
}

int qr_stats(qr_handle* h, double* out, int reset_after, void* stream)
{
    int rc = check(h); if (rc) return rc;
    if (!out) return fail(QR_ERR_INVALID, "qr_stats: null output");
    cudaStream_t s = (cudaStream_t)stream;
    QR_CUDA(cudaMemcpyAsync(out, h->stats, QR_NUM_STATS * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (reset_after) QR_CUDA(cudaMemsetAsync(h->stats, 0, QR_NUM_STATS * sizeof(double), s));
    QR_CUDA(cudaStreamSynchronize(s));
    return QR_OK;
}

}  // extern "C"
