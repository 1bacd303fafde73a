// qr_kernels.cuh -- the fused env.step() kernel and its small companions (reset, goal init, observation).
//
// One env per thread, structure-of-arrays state ([component][env], coalesced), state resident in registers
// across `n_steps` fused sub-steps, float32 observations staged through shared memory so that the
// row-major [N][O] output is written as full 128-byte lines.  No tensor cores: the dynamics are not a
// dense contraction; the roofline that binds is FP32/FP64 issue.
//
// Replaces QuadEnv.step (gym_rotor/envs/quad.py:142-168) with its wrappers' overrides
// (coupled_yaw_wrapper.py:44-110, decoupled_yaw_wrapper.py:49-161) and the trainer's reset protocol
// (main.py:212-230) when autoreset is on.
#pragma once
#include "qr_env.cuh"

namespace qr {

constexpr int QR_BLOCK = 128;

template <typename T> struct StepArgs {
    EnvConst<T> c;
    int64_t n;                 // envs in this handle (array stride)
    int64_t env_lo, env_hi;    // range processed by this launch
    int64_t env_id_offset;     // global id of local env 0
    uint32_t key0, key1;       // Philox key = seed
    T *state, *integ, *params, *goal;
    float* obs; T* reward; uint8_t *done, *terminated, *truncated; float* final_obs;
    int32_t* nfev; uint8_t* status; T* ep_return; int32_t* ep_length; uint32_t* ep_index; double* stats;
    const void* actions;       // [n_steps][n][A] f32|f64, or nullptr -> Philox U(-1,1)
    int act_f32, n_steps;
    float* obs_roll; T* reward_roll; uint8_t* done_roll;   // optional [n_steps][n][..] rollout storage
};

template <typename T> QR_DEV void load_env(EnvRegs<T>& r, const StepArgs<T>& a, int64_t e)
{
    const int64_t N = a.n;
#pragma unroll
    for (int i = 0; i < 3; ++i) r.x[i] = a.state[i * N + e];
#pragma unroll
    for (int i = 0; i < 12; ++i) r.y[i] = a.state[(3 + i) * N + e];
    r.y[12] = a.state[15 * N + e]; r.y[13] = a.state[16 * N + e]; r.W3 = a.state[17 * N + e];
#pragma unroll
    for (int i = 0; i < 8; ++i) r.I[i] = a.integ[i * N + e];
    r.m = a.params[0 * N + e]; r.d = a.params[1 * N + e]; r.J1 = a.params[2 * N + e];
    r.J3 = a.params[3 * N + e]; r.c_tf = a.params[4 * N + e]; r.c_tw = a.params[5 * N + e];
#pragma unroll
    for (int i = 0; i < 12; ++i) r.goal[i] = a.goal[i * N + e];
}

template <typename T> QR_DEV void store_state(const EnvRegs<T>& r, const StepArgs<T>& a, int64_t e)
{
    const int64_t N = a.n;
#pragma unroll
    for (int i = 0; i < 3; ++i) a.state[i * N + e] = r.x[i];
#pragma unroll
    for (int i = 0; i < 12; ++i) a.state[(3 + i) * N + e] = r.y[i];
    a.state[15 * N + e] = r.y[12]; a.state[16 * N + e] = r.y[13]; a.state[17 * N + e] = r.W3;
#pragma unroll
    for (int i = 0; i < 8; ++i) a.integ[i * N + e] = r.I[i];
}

template <typename T> QR_DEV void store_params_goal(const EnvRegs<T>& r, const StepArgs<T>& a, int64_t e, bool params, bool goal)
{
    const int64_t N = a.n;
    if (params) {
        a.params[0 * N + e] = r.m; a.params[1 * N + e] = r.d; a.params[2 * N + e] = r.J1;
        a.params[3 * N + e] = r.J3; a.params[4 * N + e] = r.c_tf; a.params[5 * N + e] = r.c_tw;
    }
    if (goal) {
#pragma unroll
        for (int i = 0; i < 12; ++i) a.goal[i * N + e] = r.goal[i];
    }
}

QR_DEV double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct LocalStats {
    double ret0, ret1, ret0sq, rew0;
    int episodes, length, crashed, truncated, steps, bad, nfev, a1, a2, a3, a4, proj;
};

// ---- the step kernel ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(QR_BLOCK) k_step(const StepArgs<T> a)
{
    extern __shared__ float s_tile[];              // [QR_BLOCK][O] observation staging
    __shared__ double s_stats[16];
    const int tid = threadIdx.x;
    const int64_t e0 = a.env_lo + (int64_t)blockIdx.x * QR_BLOCK;
    const int64_t e = e0 + tid;
    const bool live = e < a.env_hi;
    const int64_t N = a.n;
    const EnvConst<T>& c = a.c;
    const int O = (c.mode == 1) ? 23 : 18;
    const int A = (c.mode == 2) ? 5 : 4;
    const int G = (c.mode == 2) ? 2 : 1;
    const Philox ph{a.key0, a.key1};
    const uint64_t gid = (uint64_t)(a.env_id_offset + e);

    if (tid < 16) s_stats[tid] = 0.0;

    EnvRegs<T> r;
    T ep_ret[2] = {0, 0};
    int ep_len = 0;
    uint32_t ep_idx = 0;
    LocalStats ls = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (live) {
        load_env(r, a, e);
        ep_ret[0] = a.ep_return[e];
        if (G == 2) ep_ret[1] = a.ep_return[N + e];
        ep_len = a.ep_length[e];
        ep_idx = a.ep_index[e];
    }

    for (int k = 0; k < a.n_steps; ++k) {
        float o[23];
        const bool last = (k == a.n_steps - 1);
        if (live) {
            // ---- goal (pre-step state), main.py:145-147 ----
            if (c.goal_mode == 1) {
                const T W[3] = {r.y[12], r.y[13], r.W3};
                T Rg[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) Rg[i] = r.y[3 + i];
                ensure_so3<T>(Rg);   // get_desired -> state_decomposition
                traj_wd<T>(Rg, W, r.goal + 6, r.goal + 9);
            }
            // ---- action ----
            T act[5];
            bool act_f32 = a.act_f32 != 0;
            if (a.actions) {
                const int64_t base = ((int64_t)k * N + e) * A;
                if (a.act_f32) {
                    const float* p = (const float*)a.actions + base;
                    for (int i = 0; i < A; ++i) act[i] = (T)__ldg(p + i);
                } else {
                    const double* p = (const double*)a.actions + base;
                    for (int i = 0; i < A; ++i) act[i] = (T)__ldg(p + i);
                }
            } else {
                uint32_t rnd[8];
                ph((uint32_t)gid, (uint32_t)(gid >> 32), ep_idx, QR_DOMAIN_ACTION + 2u * (uint32_t)ep_len, rnd);
                if (A == 5) ph((uint32_t)gid, (uint32_t)(gid >> 32), ep_idx, QR_DOMAIN_ACTION + 2u * (uint32_t)ep_len + 1u, rnd + 4);
                for (int i = 0; i < A; ++i) act[i] = (T)(2.0 * u01(rnd[i]) - 1.0);
                act_f32 = false;
            }
            // ---- observation_wrapper: SO(3) check of the incoming R, then integrate ----
            int st = 0;
            int fl = ensure_so3<T>(r.y + 3);
            if (fl & 2) st |= 4;
            ls.proj += fl & 1;
            T f, M[3];
            action_to_fM<T>(r, c, act, act_f32, f, M);
            Dyn<T> d;
            d.fm = f / r.m; d.g = c.g;
            d.Mi0 = M[0] / r.J1; d.Mi1 = M[1] / r.J1;
            d.kw0 = (r.J1 - r.J3) / r.J1; d.kw1 = (r.J3 - r.J1) / r.J1;
            d.w3dot = M[2] / r.J3;
            int nf;
            bool finite = true;
#pragma unroll
            for (int i = 0; i < 3; ++i) finite = finite && (num<T>::abs(r.x[i]) <= num<T>::huge);
#pragma unroll
            for (int i = 0; i < 14; ++i) finite = finite && (num<T>::abs(r.y[i]) <= num<T>::huge);
            finite = finite && (num<T>::abs(r.W3) <= num<T>::huge);
            if (!finite) {
                st |= 1; nf = 0;     // scipy raises ValueError on a non-finite y0; flagged instead
            } else if (c.integrator == 1 && c.mode == 0) {
                // explicit Euler (quad.py:252-262), base env only
                T kk[14];
                rhs14<T>(r.y, r.W3, d, kk);
#pragma unroll
                for (int i = 0; i < 3; ++i) r.x[i] = num<T>::fma(r.y[i], c.dt, r.x[i]);
#pragma unroll
                for (int i = 0; i < 14; ++i) r.y[i] = num<T>::fma(kk[i], c.dt, r.y[i]);
                r.W3 = num<T>::fma(d.w3dot, c.dt, r.W3);
                nf = 1;
            } else {
                StepResult<T> res = dop853_step<T>(r.x, r.y, r.W3, d, c.dt, c.rtol, c.atol);
                st |= res.status; nf = res.nfev; ls.proj += res.nproj;
            }
            // ---- obs, reward, done ----
            double rew[2]; int dn[2];
            if (c.mode == 0) {
#pragma unroll
                for (int i = 0; i < 3; ++i) o[i] = (float)r.x[i];
#pragma unroll
                for (int i = 0; i < 12; ++i) o[3 + i] = (float)r.y[i];
                o[15] = (float)r.y[12]; o[16] = (float)r.y[13]; o[17] = (float)r.W3;
                reward_done_quad<T>(r, c, rew, dn);
            } else {
                fl = norm_error_state<T>(r, c, o);
                if (fl & 2) st |= 4;
                reward_done<T>(c, o, rew, dn);
            }
            // ---- episode accounting ----
            ep_ret[0] += (T)rew[0]; ep_ret[1] += (T)rew[1];
            ep_len += 1;
            const bool term = (dn[0] | dn[1]) != 0;
            const bool trunc = c.max_episode_steps > 0 && ep_len >= c.max_episode_steps;
            ls.steps += 1; ls.nfev += nf; ls.rew0 += rew[0]; ls.bad += (st != 0);
            { int att = (nf - 2) / 12; ls.a1 += att == 1; ls.a2 += att == 2; ls.a3 += att == 3; ls.a4 += att >= 4; }
            // ---- per-step outputs ----
            {
                T* rw = a.reward_roll ? a.reward_roll + ((int64_t)k * N + e) * G : (last ? a.reward + e * G : nullptr);
                uint8_t* dd = a.done_roll ? a.done_roll + ((int64_t)k * N + e) * G : (last ? a.done + e * G : nullptr);
                if (rw) { rw[0] = (T)rew[0]; if (G == 2) rw[1] = (T)rew[1]; }
                if (dd) { dd[0] = (uint8_t)dn[0]; if (G == 2) dd[1] = (uint8_t)dn[1]; }
                if (a.reward_roll && last) { a.reward[e * G] = (T)rew[0]; if (G == 2) a.reward[e * G + 1] = (T)rew[1]; }
                if (a.done_roll && last) { a.done[e * G] = (uint8_t)dn[0]; if (G == 2) a.done[e * G + 1] = (uint8_t)dn[1]; }
                if (last) { a.terminated[e] = (uint8_t)term; a.truncated[e] = (uint8_t)trunc; }
                if (c.diagnostics && last) a.nfev[e] = nf;
                if (st) a.status[e] |= (uint8_t)st;
            }
            // ---- auto reset (main.py:212-230) ----
            if (c.autoreset && (term || trunc)) {
                ls.episodes += 1; ls.length += ep_len; ls.crashed += term; ls.truncated += (trunc && !term);
                ls.ret0 += (double)ep_ret[0]; ls.ret1 += (double)ep_ret[1]; ls.ret0sq += (double)ep_ret[0] * (double)ep_ret[0];
                if (last) for (int i = 0; i < O; ++i) a.final_obs[e * O + i] = o[i];
                ep_idx += 1;
                double theta;
                reset_env<T>(r, ph, gid, ep_idx, c.env_type, c.udm, &theta);
                if (c.goal_mode == 1) init_goal_mode0<T>(r, theta);
                store_params_goal(r, a, e, true, c.goal_mode == 1);
                ep_ret[0] = 0; ep_ret[1] = 0; ep_len = 0;
                if (c.mode == 0) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) o[i] = (float)r.x[i];
#pragma unroll
                    for (int i = 0; i < 12; ++i) o[3 + i] = (float)r.y[i];
                    o[15] = (float)r.y[12]; o[16] = (float)r.y[13]; o[17] = (float)r.W3;
                } else {
                    norm_error_state<T>(r, c, o);   // first obs of the new episode, main.py:230
                }
            } else if (c.goal_mode == 1 && last) {
#pragma unroll
                for (int i = 0; i < 3; ++i) a.goal[(9 + i) * N + e] = r.goal[9 + i];
            }
        }
        // ---- observation tile: registers -> shared (conflict free, O is odd/even-safe) -> full lines ----
        float* dst = a.obs_roll ? a.obs_roll + ((int64_t)k * N + e0) * O : (last ? a.obs + e0 * O : nullptr);
        if (dst || (a.obs_roll && last)) {
            __syncthreads();
            if (live) for (int i = 0; i < O; ++i) s_tile[tid * O + i] = o[i];
            __syncthreads();
            const int64_t rem = a.env_hi - e0;
            const int nvalid = (int)(rem < QR_BLOCK ? rem : QR_BLOCK) * O;
            if (dst) for (int i = tid; i < nvalid; i += QR_BLOCK) dst[i] = s_tile[i];
            if (a.obs_roll && last) { float* d2 = a.obs + e0 * O; for (int i = tid; i < nvalid; i += QR_BLOCK) d2[i] = s_tile[i]; }
        }
    }

    if (live) {
        store_state(r, a, e);
        a.ep_return[e] = ep_ret[0];
        if (G == 2) a.ep_return[N + e] = ep_ret[1];
        a.ep_length[e] = ep_len;
        a.ep_index[e] = ep_idx;
    }

    // ---- statistics: warp shuffle reduce -> one shared atomic per warp -> one global atomic per block ----
    {
        double v[16] = {(double)ls.episodes, ls.ret0, ls.ret1, (double)ls.length, (double)ls.crashed, (double)ls.truncated,
                        ls.ret0sq, (double)ls.steps, (double)ls.bad, (double)ls.nfev, (double)ls.a1, (double)ls.a2,
                        (double)ls.a3, (double)ls.a4, ls.rew0, (double)ls.proj};
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            double s = warp_sum(v[i]);
            if ((tid & 31) == 0 && s != 0.0) atomicAdd(&s_stats[i], s);
        }
        __syncthreads();
        if (tid < 16 && s_stats[tid] != 0.0) atomicAdd(&a.stats[tid], s_stats[tid]);
    }
}

// ---- env.reset(env_type) -----------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(QR_BLOCK) k_reset(const StepArgs<T> a, const uint8_t* mask, int env_type)
{
    const int64_t e = a.env_lo + (int64_t)blockIdx.x * QR_BLOCK + threadIdx.x;
    if (e >= a.env_hi) return;
    if (mask && !mask[e]) return;
    const Philox ph{a.key0, a.key1};
    const uint64_t gid = (uint64_t)(a.env_id_offset + e);
    EnvRegs<T> r;
    uint32_t ep = a.ep_index[e] + 1;
    double theta;
    reset_env<T>(r, ph, gid, ep, env_type, a.c.udm, &theta);
    store_state(r, a, e);
    store_params_goal(r, a, e, true, false);
    a.ep_index[e] = ep;
    a.ep_length[e] = 0;
    a.ep_return[e] = 0;
    if (a.c.mode == 2) a.ep_return[a.n + e] = 0;
    a.status[e] = 0;
}

// ---- trajectory_generator: mark_traj_start + get_desired(mode 0) after a reset ---------------------------------
template <typename T>
__global__ void __launch_bounds__(QR_BLOCK) k_init_goal(const StepArgs<T> a, const uint8_t* mask)
{
    const int64_t e = a.env_lo + (int64_t)blockIdx.x * QR_BLOCK + threadIdx.x;
    if (e >= a.env_hi) return;
    if (mask && !mask[e]) return;
    const Philox ph{a.key0, a.key1};
    const uint64_t gid = (uint64_t)(a.env_id_offset + e);
    EnvRegs<T> r;
    load_env(r, a, e);
    uint32_t rnd[4];
    ph((uint32_t)gid, (uint32_t)(gid >> 32), a.ep_index[e], QR_DOMAIN_RESET + 4u, rnd);
    double theta = (-25.0 + 50.0 * u01(rnd[3])) * (3.14159265358979323846 / 180.);
    init_goal_mode0<T>(r, theta);
    store_params_goal(r, a, e, false, true);
}

// ---- env.get_norm_error_state(framework) ---------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(QR_BLOCK) k_norm_error_state(const StepArgs<T> a, const uint8_t* mask)
{
    const int64_t e = a.env_lo + (int64_t)blockIdx.x * QR_BLOCK + threadIdx.x;
    if (e >= a.env_hi) return;
    if (mask && !mask[e]) return;
    EnvRegs<T> r;
    load_env(r, a, e);
    float o[23];
    const int O = (a.c.mode == 1) ? 23 : 18;
    if (a.c.mode == 0) {
        for (int i = 0; i < 3; ++i) o[i] = (float)r.x[i];
        for (int i = 0; i < 12; ++i) o[3 + i] = (float)r.y[i];
        o[15] = (float)r.y[12]; o[16] = (float)r.y[13]; o[17] = (float)r.W3;
    } else {
        int fl = norm_error_state<T>(r, a.c, o);
        if (fl & 2) a.status[e] |= 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) a.integ[i * a.n + e] = r.I[i];
    }
    for (int i = 0; i < O; ++i) a.obs[e * O + i] = o[i];
}

// ---- host-layout <-> device-layout (row-major [n][C] doubles <-> [C][n] T) -----------------------------------
template <typename T> __global__ void k_aos_to_soa(const double* aos, T* soa, int64_t n, int C)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * C) return;
    const int64_t e = i / C; const int c = (int)(i % C);
    soa[(int64_t)c * n + e] = (T)aos[i];
}
template <typename T> __global__ void k_soa_to_aos(const T* soa, double* aos, int64_t n, int C)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * C) return;
    const int64_t e = i / C; const int c = (int)(i % C);
    aos[i] = (double)soa[(int64_t)c * n + e];
}

}  // namespace qr
