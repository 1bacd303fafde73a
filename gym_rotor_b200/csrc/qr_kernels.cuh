// qr_kernels.cuh -- the fused env.step() kernel and its small companions (reset, goal init, observation).
//
// Step kernel: persistent warps, one env per lane, structure-of-arrays state ([component][env], coalesced),
// state resident in registers across `n_steps` fused sub-steps, DOP853 stage derivatives in shared memory,
// float32 observations staged through a per-warp shared tile so that the row-major [N][O] output leaves as
// full 128-byte lines.  No tensor cores: the dynamics are not a dense contraction; the roofline that binds is
// FP32/FP64 instruction issue.
//
// Replaces QuadEnv.step (gym_rotor/envs/quad.py:142-168) with its wrappers' overrides
// (coupled_yaw_wrapper.py:44-110, decoupled_yaw_wrapper.py:49-161) and the trainer's reset protocol
// (main.py:212-230) when autoreset is on.
#pragma once
#include "qr_env.cuh"
#include "qr_traj.cuh"
#include "generated/actor_td3.cuh"

#ifndef QR_DEFER_RESET
#define QR_DEFER_RESET 1
#endif
#ifndef QR_RESET_BATCH
#define QR_RESET_BATCH 24
#endif
namespace qr {

constexpr int QR_BLOCK = 128;        // companion kernels (reset, goal init, observation)
constexpr int QR_MAX_THREADS = 384;  // step kernel: up to 12 persistent warps per SM (float32; 6 warps in float64)
#ifndef QR_STEP_THREADS_F32
#define QR_STEP_THREADS_F32 384
#endif
template <typename T> struct step_threads { static constexpr int value = sizeof(T) == 8 ? 192 : QR_STEP_THREADS_F32; };

template <typename T> struct StepArgs {
    EnvConst<T> c;
    int64_t n;                 // envs in this handle (array stride)
    int64_t env_lo, env_hi;    // range processed by this launch
    int64_t env_id_offset;     // global id of local env 0
    uint32_t key0, key1;       // Philox key = seed
    unsigned long long* tile_counter;   // zeroed before every launch: next 32-env tile to hand out
    T *state, *integ, *params, *goal;
    T* traj;                   // [12][n] trajectory-generator state (goal modes hover / circle / eight)
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

QR_DEV void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

QR_DEV double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- auto reset, out of line (rare: once per episode) ---------------------------------------------------------
// env.reset -> trajectory_generator.mark_traj_start/get_desired -> set_goal_state -> get_norm_error_state
// (main.py:226-230).  Works through global memory so that the hot loop's registers are not affected; the
// caller re-loads the env afterwards.  `o` receives the first observation of the new episode.
template <typename T, int MODE>
__device__ __noinline__ void auto_reset_env(const StepArgs<T>* ap, int64_t e, uint32_t episode, float* orow, T* scratch)
{
    const StepArgs<T>& a = *ap;
    constexpr int O = (MODE == 1) ? 23 : 18;
    float o[23];
    const EnvConst<T>& c = a.c;
    const Philox ph{a.key0, a.key1};
    const uint64_t gid = (uint64_t)(a.env_id_offset + e);
    EnvRegs<T> r;
    T theta;
    reset_env<T>(r, ph, gid, episode, c.env_type, c.udm, &theta);
    if (c.goal_mode == 1) init_goal_mode0<T>(r, theta);
    else if (c.goal_mode >= 2) {
        T ts[12];
        uint32_t rnd[4];
        ph((uint32_t)gid, (uint32_t)(gid >> 32), episode, QR_DOMAIN_RESET + 5u, rnd);
        traj_restart<T>(c.goal_mode, r, ts, r.goal, u01t<T>(rnd[0]), u01t<T>(rnd[1]), c.dt);
#pragma unroll
        for (int i = 0; i < 12; ++i) a.traj[i * a.n + e] = ts[i];
    } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) r.goal[i] = a.goal[i * a.n + e];
    }
    if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 3; ++i) o[i] = (float)r.x[i];
#pragma unroll
        for (int i = 0; i < 12; ++i) o[3 + i] = (float)r.y[i];
        o[15] = (float)r.y[12]; o[16] = (float)r.y[13]; o[17] = (float)r.W3;
    } else {
        norm_error_state<T>(r, c, o, MODE);   // first obs of the new episode; advances the integrals once
    }
    // new state / integrals / parameters go back to the caller through its shared scratch (the caller keeps
    // them in registers and writes the state arrays when it releases the env); parameters and the goal are
    // only ever written here
#pragma unroll
    for (int i = 0; i < 3; ++i) scratch[i] = r.x[i];
#pragma unroll
    for (int i = 0; i < 14; ++i) scratch[3 + i] = r.y[i];
    scratch[17] = r.W3;
#pragma unroll
    for (int i = 0; i < 8; ++i) scratch[18 + i] = r.I[i];
    scratch[26] = r.m; scratch[27] = r.d; scratch[28] = r.J1; scratch[29] = r.J3; scratch[30] = r.c_tf; scratch[31] = r.c_tw;
    store_params_goal(r, a, e, true, c.goal_mode >= 1);
#pragma unroll
    for (int i = 0; i < O; ++i) orow[i] = o[i];   // replaces the terminal observation in the caller's tile row
}

QR_DEV float warp_sum_f(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- the step kernel ---------------------------------------------------------------------------------------
// Persistent warps.  Every lane runs the state machine
//     [A] finish the previous env.step (observation, reward, done, outputs, auto reset) -> take the next
//         sub-step of the same env or the next env of this warp's sequence -> goal, action, SO(3) check,
//         f0 and scipy's initial step size
//     [B] one DOP853 attempt (11 stages through shared memory)
// and the warp iterates A/B until its envs are exhausted.  A lane whose attempt was rejected or whose first
// step was shorter than dt simply goes through B again while its neighbours pass through A: the adaptive
// controller costs the extra attempts it needs (about 6 % under random actions) instead of doubling the
// work of the whole warp.
//
// Warp w of the grid owns the 32-env tiles w, w + W, w + 2W, ... (W = warps in the grid); lanes take
// consecutive envs from that sequence, so loads and stores of a refill are coalesced.  Observations go
// through a per-warp shared tile and leave as full 128-byte lines.  Episode statistics are reduced with
// warp votes / REDUX into per-warp shared accumulators (no per-lane counters: registers are the scarce
// resource at 12 warps per SM) and flushed with one atomic per statistic and warp at the end.
//
// Per-warp shared memory: KS[9][14][32] T (stage derivatives) | OS[32][O] f32 (observation rows, by lane) |
//                         WS[16] f64 (statistics) | RQ[32] i32 (envs whose reset is pending, see below)
template <typename T> struct warp_smem {
    static constexpr size_t ks_bytes = (size_t)QR_NSLOTS * QR_SLOT_ELEMS * sizeof(T);
    static constexpr size_t os_bytes = 32 * 23 * sizeof(float);
    static constexpr size_t ws_bytes = 16 * sizeof(double);
    static constexpr size_t rq_bytes = 32 * sizeof(int32_t);
    static constexpr size_t bytes = ks_bytes + os_bytes + ws_bytes + rq_bytes;   // multiple of 16
};

template <typename T, int MODE>
__global__ void __launch_bounds__(step_threads<T>::value, 1) k_step(const __grid_constant__ StepArgs<T> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int O = (MODE == 1) ? 23 : 18;
    constexpr int A = (MODE == 2) ? 5 : 4;
    constexpr int G = (MODE == 2) ? 2 : 1;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const EnvConst<T>& c = a.c;
    const int64_t N = a.n;
    unsigned char* wbase = smem_raw + warp * warp_smem<T>::bytes;
    T* ks = reinterpret_cast<T*>(wbase);
    float* os = reinterpret_cast<float*>(wbase + warp_smem<T>::ks_bytes);
    double* ws = reinterpret_cast<double*>(wbase + warp_smem<T>::ks_bytes + warp_smem<T>::os_bytes);
    int32_t* rq = reinterpret_cast<int32_t*>(wbase + warp_smem<T>::ks_bytes + warp_smem<T>::os_bytes + warp_smem<T>::ws_bytes);
    const Philox ph{a.key0, a.key1};

    if (lane < 16) ws[lane] = 0.0;
    __syncwarp();

    // 32-env tiles are handed out dynamically (one atomic per tile on a per-launch counter): warps that drew
    // cheap envs simply take more tiles, so the persistent grid drains evenly
    (void)wpb;
    const int64_t n_range = a.env_hi - a.env_lo;
    const int64_t ntiles = (n_range + 31) >> 5;
    int64_t tile_base = 0;    // warp-uniform: first env of the tile currently being handed out
    int tile_pos = 32;        // warp-uniform: envs of that tile already taken (32 = none left)
    bool exhausted = false;   // warp-uniform: the counter ran past the last tile

    // per-lane persistent state
    bool busy = false, fin = false, need_init = false;
    int64_t e = 0;
    int k = 0;
    T x[3], y[14], W3 = 0, I[8], K0[14];
    T p_m = 1, p_d = 0, p_J1 = 1, p_J3 = 1, p_ctf = 0, p_ctw = 1;
    Dyn<T> d;
    OdeLane<T> ode;
    T ep_ret[2] = {0, 0};
    int ep_len = 0;
    uint32_t ep_idx = 0;
    T g_b1d[3] = {1, 0, 0}, g_Wd[3] = {0, 0, 0};   // mode-0 goal of the step in flight (xd = vd = 0): no reload at the end
#pragma unroll
    for (int i = 0; i < 3; ++i) x[i] = 0;
#pragma unroll
    for (int i = 0; i < 14; ++i) { y[i] = 0; K0[i] = 0; }
    y[3] = 1; y[7] = 1; y[11] = 1;   // idle lanes run the (warp-uniform) attempt on a benign state: R = I
#pragma unroll
    for (int i = 0; i < 8; ++i) I[i] = 0;
    d.fm = d.g = d.Mi0 = d.Mi1 = d.kw0 = d.kw1 = d.w3dot = 0;
    ode.t = 0; ode.h_abs = c.dt; ode.rejected = 0; ode.nfev = 0; ode.status = 0; ode.nproj = 0; ode.checked = 0;

    // queued resets (single-step launches): lane i resets the env in RQ[i] -- all queued envs at once -- and
    // writes the new episode's state, integrals and first observation straight to the arrays.  The lanes' own
    // env registers are not involved: the reset works through shared scratch (the stage storage, free in phase A).
    int rq_n = 0;                                                          // warp-uniform: entries in RQ
    const bool defer_ok = a.n_steps == 1 && N < ((int64_t)1 << 31);        // kernel-uniform
    auto flush_resets = [&]() {
        __syncwarp();
        if (lane < rq_n) {
            const int64_t er = (int64_t)rq[lane];
            const uint32_t ep = __ldcg(a.ep_index + er);   // written when the env was released (already incremented)
            auto_reset_env<T, MODE>(&a, er, ep, os + lane * O, ks + lane * 32);
            const T* sc = ks + lane * 32;
#pragma unroll
            for (int i = 0; i < 18; ++i) a.state[i * N + er] = sc[i];   // scratch order = state row order
#pragma unroll
            for (int i = 0; i < 8; ++i) a.integ[i * N + er] = sc[18 + i];
            float* d1 = a.obs_roll ? a.obs_roll + er * O : a.obs + er * O;
#pragma unroll
            for (int i = 0; i < O; ++i) d1[i] = os[lane * O + i];
            if (a.obs_roll) {
#pragma unroll 1
                for (int i = 0; i < O; ++i) a.obs[er * O + i] = os[lane * O + i];
            }
        }
        __syncwarp();
        rq_n = 0;
    };

    for (;;) {
        // =============================== phase A ===============================
        // ---- A1: finish the env.step that just completed ----
        const unsigned finmask = __ballot_sync(FULL, fin);
        if (finmask) {
            __syncwarp();   // phase B is over for every lane: the stage storage may be reused as reset scratch
            bool did_reset = false, ep_done = false, deferred = false, term = false, trunc = false;
            int nf = 0, st = 0, nproj = 0, ep_len_done = 0;
            float rew0f = 0.f;
            T ret_done0 = 0, ret_done1 = 0;
            const bool last = (k == a.n_steps - 1);
            if (fin) {
                float o[23];
                st = ode.status; nf = ode.nfev; nproj = ode.nproj;
                EnvRegs<T> r;
#pragma unroll
                for (int i = 0; i < 3; ++i) r.x[i] = x[i];
#pragma unroll
                for (int i = 0; i < 14; ++i) r.y[i] = y[i];
                r.W3 = W3;
#pragma unroll
                for (int i = 0; i < 8; ++i) r.I[i] = I[i];
                if (c.goal_mode == 1) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) { r.goal[i] = 0; r.goal[3 + i] = 0; r.goal[6 + i] = g_b1d[i]; r.goal[9 + i] = g_Wd[i]; }
                } else {
#pragma unroll
                    for (int i = 0; i < 12; ++i) r.goal[i] = a.goal[i * N + e];
                }
                double rew[2]; int dn[2];
                if (MODE == 0) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) o[i] = (float)x[i];
#pragma unroll
                    for (int i = 0; i < 12; ++i) o[3 + i] = (float)y[i];
                    o[15] = (float)y[12]; o[16] = (float)y[13]; o[17] = (float)W3;
                    reward_done_quad<T>(r, c, rew, dn);
                } else {
                    int fl = norm_error_state<T>(r, c, o, MODE);
                    if (fl & 2) st |= 4;
#pragma unroll
                    for (int i = 0; i < 8; ++i) I[i] = r.I[i];
                    reward_done<T>(c, o, rew, dn, MODE);
                }
                // the observation row leaves the registers at once (tile row = lane; see the copy below)
#pragma unroll
                for (int i = 0; i < O; ++i) os[lane * O + i] = o[i];
                rew0f = (float)rew[0];
                ep_ret[0] += (T)rew[0];
                if (G == 2) ep_ret[1] += (T)rew[1];
                ep_len += 1;
                term = (dn[0] | dn[1]) != 0;
                trunc = c.max_episode_steps > 0 && ep_len >= c.max_episode_steps;
                // per-step scalar outputs
                T* rw = a.reward_roll ? a.reward_roll + ((int64_t)k * N + e) * G : (last ? a.reward + e * G : nullptr);
                uint8_t* dd = a.done_roll ? a.done_roll + ((int64_t)k * N + e) * G : (last ? a.done + e * G : nullptr);
                if (rw) { rw[0] = (T)rew[0]; if (G == 2) rw[1] = (T)rew[1]; }
                if (dd) { dd[0] = (uint8_t)dn[0]; if (G == 2) dd[1] = (uint8_t)dn[1]; }
                if (a.reward_roll && last) { a.reward[e * G] = (T)rew[0]; if (G == 2) a.reward[e * G + 1] = (T)rew[1]; }
                if (a.done_roll && last) { a.done[e * G] = (uint8_t)dn[0]; if (G == 2) a.done[e * G + 1] = (uint8_t)dn[1]; }
                if (last) { a.terminated[e] = (uint8_t)term; a.truncated[e] = (uint8_t)trunc; }
                if (c.diagnostics && last) a.nfev[e] = nf;
                if (st) a.status[e] |= (uint8_t)st;
                if (c.autoreset && (term || trunc)) {
                    ep_len_done = ep_len; ret_done0 = ep_ret[0]; ret_done1 = ep_ret[1];
                    if (last) {
#pragma unroll
                        for (int i = 0; i < O; ++i) a.final_obs[e * O + i] = o[i];
                    }
                    ep_idx += 1;
                    ep_ret[0] = 0; ep_ret[1] = 0; ep_len = 0;
                    ep_done = true;
                }
            }
            // ---- auto reset.  A reset is ~1 000 instructions for one lane of the warp; with single-step launches
            // the env leaves the lane anyway, so its reset is QUEUED (per warp) and a whole batch of queued envs is
            // reset by all lanes at once further down.  Multi-step launches (the env continues in this lane) and
            // queue overflow (e.g. a common time limit hitting every env at once) reset on the spot. ----
            {
                const unsigned wmask = __ballot_sync(FULL, ep_done);
                if (wmask) {
                    if (QR_DEFER_RESET && defer_ok) {
                        const int pos = rq_n + __popc(wmask & ((1u << lane) - 1u));
                        if (ep_done && pos < 32) { rq[pos] = (int32_t)e; deferred = true; }
                        rq_n = min(32, rq_n + __popc(wmask));
                    }
                    if (ep_done && !deferred) {
                        // the new episode's first observation replaces the terminal one in this lane's staging row
                        auto_reset_env<T, MODE>(&a, e, ep_idx, os + lane * O, ks + lane * 32);   // the stage storage is free in phase A
                        did_reset = true;
                    }
                }
            }
            __syncwarp();
            // ---- statistics of the finished lanes: votes and REDUX, accumulated by lane 0 in shared memory ----
            {
                const int att = (nf - 2) / 12;
                const int s_nfev = __reduce_add_sync(FULL, nf);
                const int s_proj = __reduce_add_sync(FULL, nproj);
                const unsigned m1 = __ballot_sync(FULL, fin && att == 1), m2 = __ballot_sync(FULL, fin && att == 2);
                const unsigned m3 = __ballot_sync(FULL, fin && att == 3), m4 = __ballot_sync(FULL, fin && att >= 4);
                const unsigned mbad = __ballot_sync(FULL, fin && st != 0);
                const float s_rew = warp_sum_f(rew0f);
                const unsigned mres = __ballot_sync(FULL, ep_done);
                if (lane == 0) {
                    ws[7] += (double)__popc(finmask); ws[9] += (double)s_nfev; ws[15] += (double)s_proj;
                    ws[10] += (double)__popc(m1); ws[11] += (double)__popc(m2); ws[12] += (double)__popc(m3); ws[13] += (double)__popc(m4);
                    ws[8] += (double)__popc(mbad); ws[14] += (double)s_rew;
                }
                if (mres) {   // once per episode
                    const int s_len = __reduce_add_sync(FULL, ep_len_done);
                    const unsigned mterm = __ballot_sync(FULL, ep_done && term);
                    const double r0 = (double)warp_sum_f(ep_done ? (float)ret_done0 : 0.f);
                    const double r1 = (double)warp_sum_f(ep_done ? (float)ret_done1 : 0.f);
                    const double r0sq = (double)warp_sum_f(ep_done ? (float)ret_done0 * (float)ret_done0 : 0.f);
                    if (lane == 0) {
                        ws[0] += (double)__popc(mres); ws[3] += (double)s_len; ws[4] += (double)__popc(mterm);
                        ws[5] += (double)(__popc(mres) - __popc(mterm)); ws[1] += r0; ws[2] += r1; ws[6] += r0sq;
                    }
                }
            }
            // ---- observation rows: each finished lane writes its own row from the shared staging row (the new
            // episode's first observation if the env was just reset).  Scattered 4-byte stores, but the rows of
            // neighbouring lanes are adjacent in memory, so L2 assembles full sectors; measured 5 % faster than
            // routing the rows through a coalescing copy loop (profiles/r01_summary.md, r01n).
            if (fin && !deferred) {
                const bool lst = (k == a.n_steps - 1);
                float* d1 = a.obs_roll ? a.obs_roll + ((int64_t)k * N + e) * O : (lst ? a.obs + e * O : nullptr);
                float* d2 = (a.obs_roll && lst) ? a.obs + e * O : nullptr;
                if (d1) {
#pragma unroll
                    for (int i = 0; i < O; ++i) d1[i] = os[lane * O + i];
                }
                if (d2) {
#pragma unroll 1
                    for (int i = 0; i < O; ++i) d2[i] = os[lane * O + i];
                }
            }
            __syncwarp();
            if (fin) {
                fin = false;
                if (did_reset) {
                    const T* sc = ks + lane * 32;   // what auto_reset_env left in this lane's scratch
#pragma unroll
                    for (int i = 0; i < 3; ++i) x[i] = sc[i];
#pragma unroll
                    for (int i = 0; i < 14; ++i) y[i] = sc[3 + i];
                    W3 = sc[17];
#pragma unroll
                    for (int i = 0; i < 8; ++i) I[i] = sc[18 + i];
                    p_m = sc[26]; p_J1 = sc[28]; p_J3 = sc[29]; p_ctw = sc[31];
                    if (MODE == 0) { p_d = sc[27]; p_ctf = sc[30]; }
                }
                k += 1;
                if (k < a.n_steps) need_init = true;
                else {
                    // release the env: state back to HBM (not the terminal state of an env whose reset is queued)
                    if (!deferred) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) a.state[i * N + e] = x[i];
#pragma unroll
                        for (int i = 0; i < 12; ++i) a.state[(3 + i) * N + e] = y[i];
                        a.state[15 * N + e] = y[12]; a.state[16 * N + e] = y[13]; a.state[17 * N + e] = W3;
#pragma unroll
                        for (int i = 0; i < 8; ++i) a.integ[i * N + e] = I[i];
                    }
                    a.ep_return[e] = ep_ret[0];
                    if (G == 2) a.ep_return[N + e] = ep_ret[1];
                    a.ep_length[e] = ep_len;
                    a.ep_index[e] = ep_idx;
                    busy = false;
                }
            }
            if (QR_DEFER_RESET && rq_n >= QR_RESET_BATCH) flush_resets();
        }
        // ---- A2: idle lanes take the next envs of the warp's sequence ----
        {
            const unsigned need = __ballot_sync(FULL, !busy);
            const int cnt = __popc(need);
            if (cnt && !(exhausted && tile_pos >= 32)) {
                const int rank = __popc(need & ((1u << lane) - 1u));
                const int rem = 32 - tile_pos;
                int64_t base2 = -1;
                if (cnt > rem && !exhausted) {
                    unsigned long long t = 0;
                    if (lane == 0) t = atomicAdd(a.tile_counter, 1ULL);
                    t = __shfl_sync(FULL, t, 0);
                    if ((int64_t)t < ntiles) base2 = a.env_lo + ((int64_t)t << 5);
                    else exhausted = true;
                }
                int64_t ee = a.env_hi;
                if (!busy) {
                    if (rank < rem) ee = tile_base + tile_pos + rank;
                    else if (base2 >= 0) ee = base2 + (rank - rem);
                }
                if (cnt > rem) { tile_base = base2; tile_pos = (base2 >= 0) ? cnt - rem : 32; }
                else tile_pos += cnt;
                {
                    if (ee < a.env_hi) {
                        e = ee; k = 0; busy = true; need_init = true;
#pragma unroll
                        for (int i = 0; i < 3; ++i) x[i] = a.state[i * N + e];
#pragma unroll
                        for (int i = 0; i < 12; ++i) y[i] = a.state[(3 + i) * N + e];
                        y[12] = a.state[15 * N + e]; y[13] = a.state[16 * N + e]; W3 = a.state[17 * N + e];
#pragma unroll
                        for (int i = 0; i < 8; ++i) I[i] = a.integ[i * N + e];
                        p_m = a.params[0 * N + e]; p_J1 = a.params[2 * N + e]; p_J3 = a.params[3 * N + e]; p_ctw = a.params[5 * N + e];
                        if (MODE == 0) { p_d = a.params[1 * N + e]; p_ctf = a.params[4 * N + e]; }
                        ep_ret[0] = a.ep_return[e];
                        ep_ret[1] = (G == 2) ? a.ep_return[N + e] : (T)0;
                        ep_len = a.ep_length[e];
                        ep_idx = a.ep_index[e];
                    }
                }
            }
        }
        if (!__any_sync(FULL, busy)) {
            if (QR_DEFER_RESET && rq_n) flush_resets();
            break;
        }
        // ---- A3: start the next env.step: goal, action, SO(3) check, f0 and the initial step size ----
        if (busy && need_init) {
            need_init = false;
            // every load of this phase is issued before the first consumer (in-order issue: a stalled
            // consumer would otherwise delay the independent loads behind it by a full memory round trip)
            T act[5], b1d[3];
            bool act_f32 = a.act_f32 != 0;
            if (a.actions) {
                const int64_t base = ((int64_t)k * N + e) * A;
                if (a.act_f32) {
                    const float* p = (const float*)a.actions + base;
                    if (A == 4) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(p));
                        act[0] = (T)v.x; act[1] = (T)v.y; act[2] = (T)v.z; act[3] = (T)v.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < A; ++i) act[i] = (T)__ldg(p + i);
                    }
                } else {
                    const double* p = (const double*)a.actions + base;
#pragma unroll
                    for (int i = 0; i < A; ++i) act[i] = (T)__ldg(p + i);
                }
            }
            if (c.goal_mode == 1) {
#pragma unroll
                for (int i = 0; i < 3; ++i) b1d[i] = a.goal[(6 + i) * N + e];
            }
            if (c.goal_mode != 1) {   // the observation at the end of this step reads the goal: have it in L2 by then
#pragma unroll
                for (int i = 0; i < 12; ++i) prefetch_l2(a.goal + i * N + e);
            }
            if (!a.actions) {
                const uint64_t gid = (uint64_t)(a.env_id_offset + e);
                uint32_t rnd[8];
                ph((uint32_t)gid, (uint32_t)(gid >> 32), ep_idx, QR_DOMAIN_ACTION + 2u * (uint32_t)ep_len, rnd);
                if (A == 5) ph((uint32_t)gid, (uint32_t)(gid >> 32), ep_idx, QR_DOMAIN_ACTION + 2u * (uint32_t)ep_len + 1u, rnd + 4);
#pragma unroll
                for (int i = 0; i < A; ++i) act[i] = (T)(2.0 * u01(rnd[i]) - 1.0);
                act_f32 = false;
            }
            // state_decomposition of the incoming state: get_desired (trajectory_generator.py:115) and
            // observation_wrapper (coupled:58) both run ensure_SO3 on the same R; once is enough
            int fl = ensure_so3<T>(y + 3);
            EnvRegs<T> r;
#pragma unroll
            for (int i = 0; i < 3; ++i) r.x[i] = x[i];
#pragma unroll
            for (int i = 0; i < 14; ++i) r.y[i] = y[i];
            r.W3 = W3;
            r.m = p_m; r.d = p_d; r.J1 = p_J1; r.J3 = p_J3; r.c_tf = p_ctf; r.c_tw = p_ctw;
            if (c.goal_mode == 1) {   // goal from the pre-step state, main.py:145-147
                const T Wv[3] = {y[12], y[13], W3};
                T Wd[3];
                traj_wd<T>(y + 3, Wv, b1d, Wd);
#pragma unroll
                for (int i = 0; i < 3; ++i) { g_b1d[i] = b1d[i]; g_Wd[i] = Wd[i]; }
                if (k == a.n_steps - 1) {   // visible in the goal buffer like env.Wd after set_goal_state
#pragma unroll
                    for (int i = 0; i < 3; ++i) a.goal[(9 + i) * N + e] = Wd[i];
                }
            }
            T f, M[3];
            action_to_fM<T>(r, c, act, act_f32, f, M, MODE);
            {
                const T rm = (T)1 / p_m, rJ1 = (T)1 / p_J1, rJ3 = (T)1 / p_J3;   // inv(J) as the reference forms it (quad.py:329)
                d.fm = f * rm; d.g = c.g;
                d.Mi0 = M[0] * rJ1; d.Mi1 = M[1] * rJ1;
                d.kw0 = (p_J1 - p_J3) * rJ1; d.kw1 = (p_J3 - p_J1) * rJ1;
                d.w3dot = M[2] * rJ3;
            }
            bool finite = true;
#pragma unroll
            for (int i = 0; i < 3; ++i) finite = finite && (num<T>::abs(x[i]) <= num<T>::huge);
#pragma unroll
            for (int i = 0; i < 14; ++i) finite = finite && (num<T>::abs(y[i]) <= num<T>::huge);
            finite = finite && (num<T>::abs(W3) <= num<T>::huge);
            if (!finite) {
                // scipy raises ValueError on a non-finite y0; flagged instead, state left as it is
                ode.t = c.dt; ode.h_abs = 0; ode.rejected = 0; ode.nfev = 0; ode.status = 1; ode.nproj = 0;
                fin = true;
            } else if (MODE == 0 && c.integrator == 1) {
                // explicit Euler (quad.py:252-262), base env only
                T kk[14];
                rhs14<T>(y, W3, d, kk);
#pragma unroll
                for (int i = 0; i < 3; ++i) x[i] = num<T>::fma(y[i], c.dt, x[i]);
#pragma unroll
                for (int i = 0; i < 14; ++i) y[i] = num<T>::fma(kk[i], c.dt, y[i]);
                W3 = num<T>::fma(d.w3dot, c.dt, W3);
                ode.status = 0; ode.nproj = 0; ode.nfev = 1;
                fin = true;
            } else {
                dop853_begin<T>(x, y, W3, d, c.dt, c.rtol, c.atol, K0, ode);
            }
            if (fl & 2) ode.status |= 4;
            ode.nproj += fl & 1;
        }
        // =============================== phase B ===============================
        __syncwarp();       // scratch reads of phase A are done before any lane writes stage derivatives again
        {
            const bool live = busy && !fin;
            const bool f2 = dop853_attempt<T>(x, y, W3, d, c.dt, c.rtol, c.atol, K0, ode, ks, lane, live);
            if (live) fin = f2;
        }
    }

    // ---- flush this warp's statistics: one atomic per non-zero statistic ----
    __syncwarp();
    if (lane < 16 && ws[lane] != 0.0) atomicAdd(&a.stats[lane], ws[lane]);
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
    T theta;
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
    if (a.c.goal_mode >= 2) {
        T ts[12];
        ph((uint32_t)gid, (uint32_t)(gid >> 32), a.ep_index[e], QR_DOMAIN_RESET + 5u, rnd);
        traj_restart<T>(a.c.goal_mode, r, ts, r.goal, u01t<T>(rnd[0]), u01t<T>(rnd[1]), a.c.dt);
#pragma unroll
        for (int i = 0; i < 12; ++i) a.traj[i * a.n + e] = ts[i];
    } else {
        T theta = ((T)-25 + (T)50 * u01t<T>(rnd[3])) * ((T)3.14159265358979323846 / (T)180);
        init_goal_mode0<T>(r, theta);
    }
    store_params_goal(r, a, e, false, true);
}

// ---- trajectory_generator.get_desired(state, mode) before every step, modes hover / circle / eight (main.py:145-147) --
template <typename T>
__global__ void __launch_bounds__(QR_BLOCK) k_goal_update(const StepArgs<T> a)
{
    const int64_t e = a.env_lo + (int64_t)blockIdx.x * QR_BLOCK + threadIdx.x;
    if (e >= a.env_hi) return;
    const int64_t N = a.n;
    T x[3], v[3], R[9], W[3], ts[12], goal[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) { x[i] = a.state[i * N + e]; v[i] = a.state[(3 + i) * N + e]; W[i] = a.state[(15 + i) * N + e]; }
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = a.state[(6 + i) * N + e];
#pragma unroll
    for (int i = 0; i < 12; ++i) { ts[i] = a.traj[i * N + e]; goal[i] = a.goal[i * N + e]; }
    ensure_so3<T>(R);   // get_desired -> state_decomposition
    traj_desired<T>(traj_ref_mode(a.c.goal_mode), x, v, R, W, ts, goal, (T)0, (T)0, a.c.dt);
#pragma unroll
    for (int i = 0; i < 12; ++i) { a.traj[i * N + e] = ts[i]; a.goal[i * N + e] = goal[i]; }
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
        int fl = norm_error_state<T>(r, a.c, o, a.c.mode);
        if (fl & 2) a.status[e] |= 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) a.integ[i * a.n + e] = r.I[i];
    }
    for (int i = 0; i < O; ++i) a.obs[e * O + i] = o[i];
}

// ---- the reference's shipped TD3 actors, obs -> action on device (agent.choose_action(obs, explor_noise_std=0),
// algos/td3/td3.py:93-96 with the checkpoints of main.py:101-110) -------------------------------------------------
// One env per thread; the row-major observation tile of the block goes through shared memory so that the
// global loads are full lines.  MODE 1: monolithic actor 23 -> 4.  MODE 2: module 1 (15 -> 4) and module 2 (3 -> 1).
template <int MODE>
__global__ void __launch_bounds__(QR_BLOCK) k_actor_td3(const float* __restrict__ obs, float* __restrict__ act, int64_t n)
{
    constexpr int O = (MODE == 1) ? 23 : 18;
    constexpr int A = (MODE == 2) ? 5 : 4;
    __shared__ float tile[QR_BLOCK * O];
    const int tid = threadIdx.x;
    const int64_t e0 = (int64_t)blockIdx.x * QR_BLOCK;
    const int64_t rem = n - e0;
    const int nvalid = (int)(rem < QR_BLOCK ? rem : QR_BLOCK);
    for (int i = tid; i < nvalid * O; i += QR_BLOCK) tile[i] = obs[e0 * O + i];
    __syncthreads();
    if (tid >= nvalid) return;
    float x[O], a[A];
#pragma unroll
    for (int i = 0; i < O; ++i) x[i] = tile[tid * O + i];
    if (MODE == 1) {
        actor_td3_mono(x, a);
        *reinterpret_cast<float4*>(act + (e0 + tid) * 4) = make_float4(a[0], a[1], a[2], a[3]);
    } else {
        actor_td3_modul1(x, a);
        actor_td3_modul2(x + 15, a + 4);
#pragma unroll
        for (int i = 0; i < A; ++i) act[(e0 + tid) * A + i] = a[i];
    }
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
