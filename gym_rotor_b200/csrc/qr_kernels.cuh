// qr_kernels.cuh -- the fused env.step() kernel and its small companions (reset, goal init, observation).
//
// Step kernel: persistent warps, one env per lane, structure-of-arrays state ([component][env], coalesced),
// state resident in registers across `n_steps` fused sub-steps, DOP853 stage derivatives and the values that
// are only needed at the end of a step in shared memory, the next env fetched ahead of the end-of-step work.
// No tensor cores: the dynamics are not a dense contraction; the roofline that binds is FP32/FP64 instruction
// issue.
//
// Replaces QuadEnv.step (gym_rotor/envs/quad.py:142-168) with its wrappers' overrides
// (coupled_yaw_wrapper.py:44-110, decoupled_yaw_wrapper.py:49-161) and the trainer's reset protocol
// (main.py:212-230) when autoreset is on.
#pragma once
#include "qr_env.cuh"
#include "qr_traj.cuh"
#include "generated/actor_td3.cuh"

// Rows of the obs / final_obs buffers are padded to a multiple of 4 floats (23 -> 24, 18 -> 20), so that a lane writes its row as
// 16-byte vectors instead of scattered 4-byte stores (measured +13 %, profiles/r02a_ab.txt).  qr_obs_stride() reports the row
// stride to the host side; rollout storage handed in by the caller stays dense.
__host__ __device__ constexpr int obs_stride_of(int O) { return (O + 3) & ~3; }
#ifndef QR_RESET_BATCH
#define QR_RESET_BATCH 24
#endif
#define QR_NSTATS 20   // QR_NUM_STATS of include/quadrotor_b200.h
// 1: the next env's state is fetched global -> shared (cp.async into the park area of the stage storage, free until the reset
// section) instead of into dead registers, so that no scoreboard of the end-of-step code is shared with loads on their way to
// HBM (+2 % at one step per launch, profiles/r02/r02p_ab.txt; 0: round 1's loads into the dead registers of K0 and d, which
// the float64 kernels keep: r02r_f64_variants.txt)
#ifndef QR_PREFETCH_KS
#define QR_PREFETCH_KS 1
#endif
#ifndef QR_KS_SLOTS_F32
#define QR_KS_SLOTS_F32 6   // float32 keeps K2..K8 in slots 0..5 (qr_dop853.cuh); with 6 the block fits the 196 KB shared-memory configuration (60 KB of L1 instead of 28: +5 %, profiles/r02/r02e_ab.txt)
#endif
namespace qr {

constexpr int QR_BLOCK = 128;        // companion kernels (reset, goal init, observation)
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

#if QR_PTX
QR_DEV void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else   // host pass / tests/host_twin (see QR_PTX in qr_math.cuh)
QR_DEV void prefetch_l2(const void*) {}
#endif

// global -> shared without a register in between (LDGSTS); completion is awaited by the issuing thread
#if QR_PTX
template <int BYTES> QR_DEV void cp_async(void* smem, const void* gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem), "n"(BYTES) : "memory");
}
QR_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
QR_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> QR_DEV void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#else   // host pass / tests/host_twin: the copy happens at once, the waits are no-ops
template <int BYTES> QR_DEV void cp_async(void* smem, const void* gmem) { memcpy(smem, gmem, BYTES); }
QR_DEV void cp_async_wait_all() {}
QR_DEV void cp_async_commit() {}
template <int N> QR_DEV void cp_async_wait_group() {}
#endif
// Stores behind a predicate that can never become a branch.  The warp runs phase A with three warps per scheduler and
// nothing to hide a fetch redirect behind: every TAKEN branch on the common path (a skipped option, a skipped rare case)
// measured ~0.3 % of the launch, about ten instructions' worth (profiles/r02/r02ap_taken_branches.txt).
#if QR_PTX
QR_DEV void st_if(bool p, float* g, float v) { asm volatile("{ .reg .pred q; setp.ne.s32 q, %0, 0; @q st.global.f32 [%1], %2; }" ::"r"((int)p), "l"(g), "f"(v) : "memory"); }
QR_DEV void st_if(bool p, double* g, double v) { asm volatile("{ .reg .pred q; setp.ne.s32 q, %0, 0; @q st.global.f64 [%1], %2; }" ::"r"((int)p), "l"(g), "d"(v) : "memory"); }
QR_DEV void sts_if(bool p, double* s, double v)
{
    asm volatile("{ .reg .pred q; setp.ne.s32 q, %0, 0; @q st.shared.f64 [%1], %2; }" ::"r"((int)p), "r"((unsigned)__cvta_generic_to_shared(s)), "d"(v) : "memory");
}
#else
QR_DEV void st_if(bool p, float* g, float v) { if (p) *g = v; }
QR_DEV void st_if(bool p, double* g, double v) { if (p) *g = v; }
QR_DEV void sts_if(bool p, double* s, double v) { if (p) *s = v; }
#endif
template <typename T> QR_DEV int32_t& stash_i32(T* sh, int slot) { return *reinterpret_cast<int32_t*>(sh + slot * 32); }

// ---- auto reset, out of line (rare: once per episode) ---------------------------------------------------------
// env.reset -> trajectory_generator.mark_traj_start/get_desired -> set_goal_state -> get_norm_error_state
// (main.py:226-230).  Works through global memory so that the hot loop's registers are not affected; the
// caller re-loads the env afterwards.  `orow` (and `orow2`) receive the first observation of the new episode.
template <typename T, int MODE>
__device__ __noinline__ void auto_reset_env(const StepArgs<T>* ap, int64_t e, uint32_t episode, float* orow, float* orow2, T* scratch)
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
        traj_restart<T, false>(c.goal_mode, r, ts, r.goal, u01t<T>(rnd[0]), u01t<T>(rnd[1]), c.dt);   // hover / circle / eight (see qr_traj.cuh)
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
    for (int i = 0; i < O; ++i) orow[i] = o[i];   // replaces the terminal observation of the step that ended the episode
    if (orow2) {
#pragma unroll 1
        for (int i = 0; i < O; ++i) orow2[i] = o[i];
    }
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
// Warps draw 32-env tiles from a per-launch atomic counter; lanes take consecutive envs of the warp's tile
// sequence, so the loads and stores of a refill are coalesced.  A lane writes its observation row itself, or
// -- when the whole warp finishes 32 consecutive envs together -- the rows leave as one contiguous block
// through a shared tile.  Episode statistics are reduced with warp votes / REDUX into per-warp shared
// accumulators (no per-lane counters: registers are the scarce resource at 12 warps per SM) and flushed with
// one atomic per statistic and warp at the end.  Phase A in order: A0 fetch the next env ahead, A1 finish the
// step, (parked) auto reset, A2 adopt the fetched env, A3 start the next step.
//
// Per-warp shared memory: KS[8][14][32] T (stage derivatives) | STASH[36][32] T (per-lane values that are only needed
//                         when a step ends: integrals, goal, episode counters; and the landing zone of the next
//                         env's action / parameters / goal, fetched ahead) | WS[16] f64 (statistics) |
//                         RQ[64] i32 (envs whose reset is queued, see `parked reset`).
// Multi-step launches keep two more small arrays in the four unused stash slots (32..35): the sub-step at which a queued
// env goes on (i16), and CQ[64] (env i32, step i16), the reset envs that wait for a free lane to go on stepping.
// The float32 total is 16 512 bytes per warp: 12 warps (all the registers allow) + the 1 KB the system reserves fit the
// 196 KB shared-memory configuration, which leaves 60 KB of L1 for the loop's few spilled values and the row stores; at
// 19 328 bytes (228 KB configuration, 28 KB of L1) the kernel was 5 % slower, and beyond that it loses a warp per SM (-9 %).
template <typename T> struct warp_smem {
    static constexpr size_t park_elems = 1024 + 58 * 32;   // phase A's other use of the stage storage: reset scratch + parked lane state
    static constexpr size_t slot_elems = (size_t)(sizeof(T) == 4 ? QR_KS_SLOTS_F32 : QR_NSLOTS) * QR_SLOT_ELEMS;
    static constexpr size_t ks_bytes = (slot_elems > park_elems ? slot_elems : park_elems) * sizeof(T);
    static constexpr size_t os_bytes = 32 * 36 * sizeof(T);   // the stash: 36 slots per lane
    static constexpr size_t ws_bytes = QR_NSTATS * sizeof(double);
    static constexpr int rq_cap = 64, cq_cap = 64;   // envs out of their lane never exceed QR_RESET_BATCH - 1 + 32 (see `parked reset`)
    static constexpr size_t rq_bytes = rq_cap * sizeof(int32_t);
    static constexpr size_t bytes = ks_bytes + os_bytes + ws_bytes + rq_bytes;   // multiple of 16
    static_assert(sizeof(T) == 8 || bytes * 12 + 1024 <= 196 * 1024, "float32: 12 warps per SM within the 196 KB shared-memory configuration");
    static_assert(4 * 32 * sizeof(T) >= cq_cap * 6 + rq_cap * 2, "the queues of multi-step launches live in stash slots 32..35");
};

// GOAL1: the goal is generated on the device in trajectory mode 0 (config goal_mode == 1) -- a template parameter so
// that this configuration has no global load at the end of a step (it would share a scoreboard with the loads that
// fetch the next env, and wait for them).
// POLICY: the action of every (sub-)step is the reference's shipped TD3 actor evaluated on the env's latest observation
// (agent.choose_action(obs, explor_noise_std=0), algos/td3/td3.py:93-96; the evaluation loop of main.py:304-365 fused
// into the launch).  Only instantiated for the wrapper modes with MULTI = true.
template <typename T, int MODE, bool MULTI, bool GOAL1, bool POLICY = false>
__global__ void __launch_bounds__(step_threads<T>::value, 1) k_step(const __grid_constant__ StepArgs<T> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int O = (MODE == 1) ? 23 : 18;
    constexpr int OS = obs_stride_of(O);   // row stride of a.obs / a.final_obs
    constexpr int A = (MODE == 2) ? 5 : 4;
    constexpr int G = (MODE == 2) ? 2 : 1;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const EnvConst<T>& c = a.c;
    const int64_t N = a.n;
    const int NS = MULTI ? a.n_steps : 1;   // single-step kernels: the sub-step counter k folds to the constant 0
    constexpr bool PREFETCH_KS = QR_PREFETCH_KS != 0 && sizeof(T) == 4;   // float64: the loads into dead registers measured 4 % faster
    unsigned char* wbase = smem_raw + warp * warp_smem<T>::bytes;
    T* ks = reinterpret_cast<T*>(wbase);
    T* const sh = reinterpret_cast<T*>(wbase + warp_smem<T>::ks_bytes) + lane;   // stash slot j of this lane: sh[j * 32]
    double* ws = reinterpret_cast<double*>(wbase + warp_smem<T>::ks_bytes + warp_smem<T>::os_bytes);
    int32_t* rq = reinterpret_cast<int32_t*>(wbase + warp_smem<T>::ks_bytes + warp_smem<T>::os_bytes + warp_smem<T>::ws_bytes);
    // multi-step launches: CQ (env, step) and the steps of RQ's entries, in stash slots 32..35
    int32_t* const cq = reinterpret_cast<int32_t*>(reinterpret_cast<T*>(wbase + warp_smem<T>::ks_bytes) + 32 * 32);
    int16_t* const cqk = reinterpret_cast<int16_t*>(cq + warp_smem<T>::cq_cap);
    int16_t* const rqk = cqk + warp_smem<T>::cq_cap;
    const Philox ph{a.key0, a.key1};

    if (lane < QR_NSTATS) ws[lane] = 0.0;
    __syncwarp();

    // 32-env tiles are handed out dynamically (one atomic per tile on a per-launch counter): warps that drew
    // cheap envs simply take more tiles, so the persistent grid drains evenly
    const int64_t n_range = a.env_hi - a.env_lo;
    const int64_t ntiles = (n_range + 31) >> 5;
    // (dealing the tiles round robin instead -- no atomic, whose answer is the most stalled-on instruction of single-step
    //  launches -- measured 1.5 % to 3.7 % slower: profiles/r02/r02aa_ab_static_tiles.txt)
    // (env indices are 32-bit in the kernel: qr_create refuses more than 2^30 envs per handle -- 366 B each; three registers)
    int tile_base = 0;        // warp-uniform: first env of the tile currently being handed out
    int tile_pos = 32;        // warp-uniform: envs of that tile already taken (32 = none left, 33 = the counter ran past the last tile)

    // per-lane persistent state (registers).  Everything that is only needed when a step ENDS -- integral
    // errors, the goal of the step in flight, episode return / length / index -- lives in the lane's stash in
    // shared memory instead: 168 registers are all a thread gets at 12 warps per SM.
    bool busy = false, fin = false, need_init = false;
    int e = 0;
    int k = 0;
    T x[3], y[14], W3 = 0, K0[14];
    Dyn<T> d;
    OdeLane<T> ode;
    constexpr int S_I = 0, S_B1D = 8, S_WD = 11, S_RET0 = 14, S_LEN = 15, S_IDX = 16, S_RET1 = 17;   // stash slots: env in flight
    constexpr int S_ACT = 18, S_PAR = 23, S_NB1D = 29;   // next env, fetched ahead: action (<= 5), parameters (6), b1d (3)
    bool has_next = false;   // K0 / d hold the NEXT env's state (loads in flight), the stash its action etc.
    bool fresh = false;      // the env was adopted in this round's A2: its action / parameters / b1d are in the stash
    bool r_ok = false;       // multi-step launches: the attitude passed ensure_SO3 in the observation that ended the previous sub-step
    bool obs_in_tile = false;   // policy rollouts: this round's A1 left the lane's observation row in the shared tile (stage storage)
    int e_next = 0;
    int k_next = 0;          // multi-step launches: the sub-step at which the fetched env goes on (a reset env resumes mid-rollout)
#pragma unroll
    for (int i = 0; i < 3; ++i) x[i] = 0;
#pragma unroll
    for (int i = 0; i < 14; ++i) { y[i] = 0; K0[i] = 0; }
    y[3] = 1; y[7] = 1; y[11] = 1;   // idle lanes run the (warp-uniform) attempt on a benign state: R = I
    d.fm = d.g = d.Mi0 = d.Mi1 = d.kw0 = d.kw1 = d.w3dot = 0;
    ode.t = 0; ode.h_abs = c.dt; ode.rejected = 0; ode.nfev = 0; ode.status = 0; ode.nproj = 0; ode.checked = 0;

    int rq_n = 0;   // warp-uniform: entries in RQ
    int cq_n = 0;   // warp-uniform: entries in CQ (multi-step launches)

    for (;;) {
        // =============================== phase A ===============================
        if (PREFETCH_KS) __syncwarp();   // the stage storage is written again below (A0's fetch-ahead lands in its park area): phase B must be over for every lane
        const unsigned finmask = __ballot_sync(FULL, fin);
        // ---- A0: lanes that are idle, or about to release their env, are given the next env of the warp's sequence
        // NOW and start fetching it with cp.async (no register in between): the state into the park area of the stage
        // storage, action / parameters / goal into the stash.  The round trip to HBM overlaps the end-of-step work
        // below instead of stalling the start of the next step.
        {
            const bool leaving = fin && (k == NS - 1);
            const unsigned need = __ballot_sync(FULL, (!busy || leaving) && !has_next);
            const int cnt = __popc(need);
            const int ncq = MULTI ? min(cnt, cq_n) : 0;   // reset envs waiting to go on stepping are served first
            if (cnt && (ncq > 0 || tile_pos <= 32)) {
                const bool mine = (need >> lane) & 1u;
                const int rank = __popc(need & ((1u << lane) - 1u)) - ncq;   // < 0: this lane takes a waiting env
                int ee = 0x7fffffff;
                int kk = 0;
                if (MULTI && mine && rank < 0) { ee = cq[cq_n + rank]; kk = cqk[cq_n + rank]; }
                cq_n -= ncq;
                const int cnt_t = cnt - ncq;   // lanes served from the tile sequence
                if (cnt_t > 0 && tile_pos <= 32) {
                    const int rem = 32 - tile_pos;
                    int base2 = -1;
                    if (cnt_t > rem) {
                        unsigned long long t = 0;
                        if (lane == 0) t = atomicAdd(a.tile_counter, 1ULL);
                        t = __shfl_sync(FULL, t, 0);
                        if ((int64_t)t < ntiles) base2 = (int)(a.env_lo + ((int64_t)t << 5));
                    }
                    if (mine && rank >= 0) {
                        if (rank < rem) ee = tile_base + tile_pos + rank;
                        else if (base2 >= 0) ee = base2 + (rank - rem);
                    }
                    if (cnt_t > rem) { tile_base = base2; tile_pos = (base2 >= 0) ? cnt_t - rem : 33; }
                    else tile_pos += cnt_t;
                }
                if (ee < a.env_hi) {
                    has_next = true; e_next = ee;
                    if (MULTI) k_next = kk;
                    // K0 and d are dead for a lane that is idle or has finished its step (A1 only reads ode's counters)
                    // (K0 is kept in the integrator's internal order, qr_dop853.cuh: the fetched state lands in the
                    //  matching positions, so that y and K0 agree on which components form a register pair)
                    if (PREFETCH_KS) {   // (phase B of every lane is over: the __syncwarp at the top of the round)
#pragma unroll
                        for (int i = 0; i < 18; ++i) cp_async<sizeof(T)>(ks + 1024 + i * 32 + lane, a.state + i * N + ee);
                    } else {
                        d.fm = a.state[0 * N + ee]; d.g = a.state[1 * N + ee]; d.Mi0 = a.state[2 * N + ee];
#pragma unroll
                        for (int i = 0; i < 14; ++i) K0[zof(i)] = a.state[(3 + i) * N + ee];
                        d.Mi1 = a.state[17 * N + ee];
                    }
                    if (a.actions) {
                        if (a.act_f32) {
                            const float* p = (const float*)a.actions + ((int64_t)kk * N + ee) * A;
#pragma unroll
                            for (int i = 0; i < A; ++i) cp_async<4>(sh + (S_ACT + i) * 32, p + i);
                        } else if (sizeof(T) == 8) {
                            const double* p = (const double*)a.actions + ((int64_t)kk * N + ee) * A;
#pragma unroll
                            for (int i = 0; i < A; ++i) cp_async<8>(sh + (S_ACT + i) * 32, p + i);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        if (MODE == 0 || (i != 1 && i != 4)) cp_async<sizeof(T)>(sh + (S_PAR + i) * 32, a.params + i * N + ee);
                    }
                    if (GOAL1) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) cp_async<sizeof(T)>(sh + (S_NB1D + i) * 32, a.goal + (6 + i) * N + ee);
                    }
                }
            }
            cp_async_commit();
        }
        // ---- A1: finish the env.step that just completed ----
        // One test per lane for everything that does not happen on every sub-step of a multi-step launch (rows and scalar
        // outputs of the last sub-step, caller's rollout storage, the end of an episode, the release of the env, a status
        // flag, rounded returns), one warp-uniform test for what follows it (rows of a lock-step warp, the reset queue,
        // episode statistics, diagnostics); what every step needs is straight-line code.  Single-step launches take the
        // first block unconditionally (`last` and `!stay` are compile-time constants there).
        if (finmask) {
            __syncwarp();   // phase B is over for every lane: the stage storage may be reused as reset scratch
            bool ep_done = false, term = false, trunc = false;
            int nf = 0, st = 0, nproj = 0, ep_len_done = 0;
            float rew0f = 0.f;
            float bex0 = 0.f, bex1 = 0.f, bex2 = 0.f, beb1 = 0.f;   // what benchmark_reward_func reads (diagnostics)
            bool solved = false;
            T ret_done0 = 0, ret_done1 = 0;
            const bool finl = fin;
            const bool last = (k == NS - 1);
            // all 32 lanes finish the same sub-step of 32 consecutive envs, first one 4-aligned (16-byte aligned rows block)
            const int e_first = __shfl_sync(FULL, e, 0);
            const int k_first = __shfl_sync(FULL, k, 0);
            const bool same = __all_sync(FULL, e - lane == e_first && k == k_first);   // (no warp primitive behind a short-circuit)
            // padded rows are 16-byte aligned; dense rollout rows: per lane.  (`same` makes `last` warp-uniform.)
            const bool coop = finmask == FULL && same && !(MULTI && a.obs_roll) && (last || POLICY);
            if (fin) {
                float o[23];
                st = ode.status; nf = ode.nfev; nproj = ode.nproj;
                cp_async_wait_group<1>();   // A2's copies into the stash, long done (all but this round's A0 group)
                EnvRegs<T> r;
#pragma unroll
                for (int i = 0; i < 3; ++i) r.x[i] = x[i];
#pragma unroll
                for (int i = 0; i < 14; ++i) r.y[i] = y[i];
                r.W3 = W3;
#pragma unroll
                for (int i = 0; i < 8; ++i) r.I[i] = sh[(S_I + i) * 32];
                T ep_ret0 = sh[S_RET0 * 32], ep_ret1 = 0;   // episode accumulators
                if (G == 2) ep_ret1 = sh[S_RET1 * 32];
                int ep_len = stash_i32(sh, S_LEN);
                uint32_t ep_idx = (uint32_t)stash_i32(sh, S_IDX);
                if (GOAL1) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) { r.goal[i] = 0; r.goal[3 + i] = 0; r.goal[6 + i] = sh[(S_B1D + i) * 32]; r.goal[9 + i] = sh[(S_WD + i) * 32]; }
                } else if (MODE != 0 && c.goal_mode >= 2) {   // trajectory goal of this step: left in the stash by A3 (kernel-uniform)
#pragma unroll
                    for (int i = 0; i < 3; ++i) { r.goal[i] = sh[(S_B1D + i) * 32]; r.goal[3 + i] = sh[(S_WD + i) * 32]; }
                    r.goal[6] = sh[(S_NB1D + 0) * 32]; r.goal[7] = sh[(S_NB1D + 1) * 32]; r.goal[8] = 0;
                    r.goal[9] = 0; r.goal[10] = 0; r.goal[11] = sh[(S_NB1D + 2) * 32];
                } else {
#pragma unroll
                    for (int i = 0; i < 12; ++i) r.goal[i] = a.goal[i * N + e];
                }
                T rew[2]; int dn[2];   // (float32 mode: the reward is a float32 value anyway)
                if (MODE == 0) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) o[i] = (float)x[i];
#pragma unroll
                    for (int i = 0; i < 12; ++i) o[3 + i] = (float)y[i];
                    o[15] = (float)y[12]; o[16] = (float)y[13]; o[17] = (float)W3;
                    double rq[2];
                    reward_done_quad<T>(r, c, rq, dn);
                    rew[0] = (T)rq[0]; rew[1] = (T)rq[1];
                } else {
                    int fl = norm_error_state<T>(r, c, o, MODE);
                    if (fl & 2) st |= 4;
                    if (MULTI) r_ok = (fl == 0);
                    reward_done<T, T>(c, o, rew, dn, MODE);
                }
                rew0f = (float)rew[0];
                T ret0 = ep_ret0 + rew[0], ret1 = (G == 2) ? ep_ret1 + rew[1] : (T)0;   // (rounded returns: replaced below)
                ep_len += 1;
                term = (dn[0] | dn[1]) != 0;
                trunc = c.max_episode_steps > 0 && ep_len >= c.max_episode_steps;
                if (MODE != 0) {
                    // benchmark_reward_func(ex, eb1) = interp(-|ex| - |eb1|, [-2, 0], [0, 1]) (utils/utils.py:21-47) on this
                    // step's observation (a per-step statistic: with diagnostics only, evaluated in the statistics block), and
                    // the trainer's "solved" relabel at the time limit (main.py:169-173)
                    bex0 = o[0] * (float)c.x_lim; bex1 = o[1] * (float)c.x_lim; bex2 = o[2] * (float)c.x_lim;
                    beb1 = o[MODE == 1 ? 18 : 15];
                    solved = trunc && fabsf(bex0) <= 0.03f && fabsf(bex1) <= 0.03f && fabsf(bex2) <= 0.03f && rew[0] != (T)-1;
                }
                const bool epd = c.autoreset && (term || trunc);
                const bool stay = MULTI && k + 1 < NS && !epd;   // the env keeps its lane for the next sub-step
                const bool want_row = last || POLICY || (MULTI && a.obs_roll);
                if (want_row || !stay || st != 0 || c.round_returns || (MULTI && (a.reward_roll || a.done_roll))) {
                    if (c.round_returns) {   // the trainer's running return, main.py:180: float('{:.4f}'.format(ret + r)) every step
                        ret0 = (T)(rint(((double)ep_ret0 + (double)rew[0]) * 1e4) / 1e4);   // k / 10^4 correctly rounded = float('0.dddd')
                        if (G == 2) ret1 = (T)(rint(((double)ep_ret1 + (double)rew[1]) * 1e4) / 1e4);
                    }
                    // ---- the observation row.  General case: each lane writes its own row (the rows of neighbouring lanes
                    // are adjacent in memory, so L2 assembles full sectors; measured 5 % faster than a general coalescing
                    // copy through shared memory).  When the whole warp finishes 32 consecutive envs together (`coop`: the
                    // lock-step regime of a trained policy), the rows go through a shared tile and leave as 16-byte stores
                    // of one contiguous block.
                    if (want_row) {
                        float *obs1 = nullptr, *obs2 = nullptr;   // where this step's observation row goes
                        if (MULTI && a.obs_roll) {   // kernel-uniform: caller's rollout storage (dense rows), plus the handle's row where it is read back
                            obs1 = a.obs_roll + ((int64_t)k * N + e) * O;
                            if (last || POLICY) obs2 = a.obs + (int64_t)e * OS;   // POLICY: the actor reads a.obs at the next sub-step
                        } else obs1 = a.obs + (int64_t)e * OS;
                        if (coop || (POLICY && MULTI)) {
                            // (policy rollouts: the actor of the next sub-step reads the row from here, not from HBM -- unless a
                            //  reset batch reuses the stage storage in between, see A3)
                            float* tile = reinterpret_cast<float*>(ks);   // the stage storage is free in phase A
#pragma unroll
                            for (int i = 0; i < OS / 4; ++i)
                                reinterpret_cast<float4*>(tile + lane * OS)[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], (4 * i + 3 < O) ? o[4 * i + 3] : 0.f);
                            if (POLICY && MULTI) obs_in_tile = true;
                        }
                        if (!coop) {   // rows inside a.obs are 16-byte aligned and padded: vector stores; caller's rollout storage: scalar
                            float op[OS];
#pragma unroll
                            for (int i = 0; i < OS; ++i) op[i] = (i < O) ? o[i] : 0.f;
                            float* rows[2] = {obs1, obs2};
#pragma unroll
                            for (int w = 0; w < 2; ++w) {
                                float* row = rows[w];
                                if (!row) continue;
                                if (MULTI && a.obs_roll && w == 0) {
#pragma unroll
                                    for (int i = 0; i < O; ++i) row[i] = o[i];
                                } else {
#pragma unroll
                                    for (int i = 0; i < OS / 4; ++i)
                                        reinterpret_cast<float4*>(row)[i] = make_float4(op[4 * i], op[4 * i + 1], op[4 * i + 2], op[4 * i + 3]);
                                }
                            }
                        }
                    }
                    // per-step scalar outputs
                    if (MULTI && a.reward_roll) {   // (kernel-uniform)
                        T* rw = a.reward_roll + ((int64_t)k * N + e) * G;
                        rw[0] = rew[0]; if (G == 2) rw[1] = rew[1];
                    }
                    if (MULTI && a.done_roll) {
                        uint8_t* dd = a.done_roll + ((int64_t)k * N + e) * G;
                        dd[0] = (uint8_t)dn[0]; if (G == 2) dd[1] = (uint8_t)dn[1];
                    }
                    if (last) {
                        a.reward[(int64_t)e * G] = rew[0]; if (G == 2) a.reward[(int64_t)e * G + 1] = rew[1];
                        a.done[(int64_t)e * G] = (uint8_t)dn[0]; if (G == 2) a.done[(int64_t)e * G + 1] = (uint8_t)dn[1];
                        a.terminated[e] = (uint8_t)term; a.truncated[e] = (uint8_t)trunc;
                        if (c.diagnostics) a.nfev[e] = nf;
                        if (epd) {
#pragma unroll
                            for (int i = 0; i < O; ++i) a.final_obs[(int64_t)e * OS + i] = o[i];
                        }
                    }
                    if (st) a.status[e] |= (uint8_t)st;
                    if (!stay) {
                        // release the env: state back to HBM -- but not the terminal state of an env whose reset is queued (the
                        // reset writes the new one; in a multi-step launch another lane then goes on stepping it, see CQ)
                        if (!epd) {
#pragma unroll
                            for (int i = 0; i < 3; ++i) a.state[i * N + e] = x[i];
#pragma unroll
                            for (int i = 0; i < 12; ++i) a.state[(3 + i) * N + e] = y[i];
                            a.state[15 * N + e] = y[12]; a.state[16 * N + e] = y[13]; a.state[17 * N + e] = W3;
#pragma unroll
                            for (int i = 0; i < 8; ++i) a.integ[i * N + e] = r.I[i];
                        }
                        a.ep_return[e] = epd ? (T)0 : ret0;
                        if (G == 2) a.ep_return[N + e] = epd ? (T)0 : ret1;
                        a.ep_length[e] = epd ? 0 : ep_len;
                        a.ep_index[e] = ep_idx + (epd ? 1u : 0u);
                        busy = false;
                    }
                }
                // the end of an episode, without a branch: only noted here, the reset is carried out further down
                ep_done = epd;
                ep_len_done = epd ? ep_len : 0; ret_done0 = epd ? ret0 : (T)0; ret_done1 = epd ? ret1 : (T)0;
                fin = false;
                if (MULTI) k += 1;
                need_init = stay;
                if (stay) {   // end-of-step values back into the stash
#pragma unroll
                    for (int i = 0; i < 8; ++i) sh[(S_I + i) * 32] = r.I[i];
                    sh[S_RET0 * 32] = ret0;
                    if (G == 2) sh[S_RET1 * 32] = ret1;
                    stash_i32(sh, S_LEN) = ep_len;
                    stash_i32(sh, S_IDX) = (int32_t)ep_idx;
                }
            }
            // ---- statistics of the finished lanes: votes and REDUX; lane 0 accumulates in shared memory (every lane forms
            // the sums, one stores: no divergent region) ----
            const unsigned mres = __ballot_sync(FULL, ep_done);
            {
                const int s_nfev = __reduce_add_sync(FULL, nf);
                const unsigned mbad = __ballot_sync(FULL, finl && st != 0);
                const unsigned msolved = (MODE != 0) ? __ballot_sync(FULL, solved) : 0u;
                const bool l0 = lane == 0;
                const double v7 = ws[7] + (double)__popc(finmask), v9 = ws[9] + (double)s_nfev, v8 = ws[8] + (double)__popc(mbad);
                const double v17 = (MODE != 0) ? ws[17] + (double)__popc(msolved) : 0.0;
                __syncwarp();   // every lane has read; lane 0 writes (the next reads are a round away, behind the barrier that ends A1)
                sts_if(l0, ws + 7, v7); sts_if(l0, ws + 9, v9); sts_if(l0, ws + 8, v8);
                if (MODE != 0) sts_if(l0, ws + 17, v17);
            }
            if (coop || mres != 0 || c.diagnostics) {   // (warp-uniform)
                if (coop) {
                    __syncwarp();
                    const float4* tile4 = reinterpret_cast<const float4*>(ks);
                    constexpr int NV = 32 * OS / 4;   // float4 elements of the 32-row block (rows padded)
                    float4* g1 = reinterpret_cast<float4*>(a.obs + (int64_t)e_first * OS);
#pragma unroll
                    for (int it = 0; it < (NV + 31) / 32; ++it) {
                        const int q = it * 32 + lane;
                        if (q < NV) g1[q] = tile4[q];
                    }
                    __syncwarp();
                }
                if (mres) {   // once per episode
                    // auto reset: single-step launches queue the env too (it leaves the lane anyway), so that a whole batch is
                    // reset at once.  RQ holds 64: at most QR_RESET_BATCH - 1 entries from earlier rounds + 32 new ones
                    if (ep_done) {
                        const int q = rq_n + __popc(mres & ((1u << lane) - 1u));
                        rq[q] = e; if (MULTI) rqk[q] = (int16_t)k;   // the env goes on at the next sub-step (k: already advanced)
                    }
                    rq_n += __popc(mres);
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
                if (c.diagnostics) {
                    // per-step sums -- attempt histogram, reward, benchmark reward, projections -- with diagnostics only
                    // (kernel-uniform); the mean attempt count is always available from the nfev sum (2 + 12 per attempt),
                    // returns from the episode statistics
                    const int att = (nf - 2) / 12;
                    const int s_proj = __reduce_add_sync(FULL, nproj);
                    const unsigned m1 = __ballot_sync(FULL, finl && att == 1), m2 = __ballot_sync(FULL, finl && att == 2);
                    const unsigned m3 = __ballot_sync(FULL, finl && att == 3), m4 = __ballot_sync(FULL, finl && att >= 4);
                    const float s_rew = warp_sum_f(rew0f);
                    float s_brew = 0.f;
                    if (MODE != 0) {
                        const float rb = -sqrtf(fmaf(bex2, bex2, fmaf(bex1, bex1, bex0 * bex0))) - fabsf(beb1 * 3.14159265358979f);
                        s_brew = warp_sum_f(finl ? fminf(fmaxf(fmaf(rb, 0.5f, 1.0f), 0.f), 1.f) : 0.f);
                    }
                    if (lane == 0) {
                        ws[15] += (double)s_proj;
                        ws[10] += (double)__popc(m1); ws[11] += (double)__popc(m2); ws[12] += (double)__popc(m3); ws[13] += (double)__popc(m4);
                        ws[14] += (double)s_rew; ws[16] += (double)s_brew;
                    }
                }
            }
            __syncwarp();
        }
        // ---- A2: idle lanes adopt the env fetched in A0 ----
        if (has_next && !busy) {
            has_next = false;
            e = e_next; k = MULTI ? k_next : 0; busy = true; need_init = true; if (MULTI) fresh = true;
            if (PREFETCH_KS) {
                cp_async_wait_all();   // this round's A0 group, issued a whole end-of-step ago
                const T* pf = ks + 1024 + lane;
#pragma unroll
                for (int i = 0; i < 3; ++i) x[i] = pf[i * 32];
#pragma unroll
                for (int i = 0; i < 14; ++i) y[i] = pf[(3 + i) * 32];
                W3 = pf[17 * 32];
            } else {
                x[0] = d.fm; x[1] = d.g; x[2] = d.Mi0; W3 = d.Mi1;
#pragma unroll
                for (int i = 0; i < 14; ++i) y[i] = K0[zof(i)];
            }
            // end-of-step values go global -> stash without passing through registers (needed when the step ends)
#pragma unroll
            for (int i = 0; i < 8; ++i) cp_async<sizeof(T)>(sh + (S_I + i) * 32, a.integ + i * N + e);
            cp_async<sizeof(T)>(sh + S_RET0 * 32, a.ep_return + e);
            if (G == 2) cp_async<sizeof(T)>(sh + S_RET1 * 32, a.ep_return + N + e);
            cp_async<4>(&stash_i32(sh, S_LEN), a.ep_length + e);
            cp_async<4>(&stash_i32(sh, S_IDX), a.ep_index + e);
        }
        cp_async_commit();
        const bool drained = !__any_sync(FULL, busy);
        // ---- parked reset.  A reset is ~1 000 instructions and out of line; a call from inside this loop would put
        // every value that lives across it into local memory FOR THE WHOLE LOOP (measured: ~2.5 M local accesses per
        // launch, in-flight loads serialised behind them).  So the lane state is parked in the stage storage (free in
        // phase A) around the call and re-defined from there afterwards: nothing is live across it.  A whole
        // batch of queued envs is reset at once, one per lane (the env left its lane when its episode ended).  In a
        // multi-step launch a reset env that still has sub-steps to go then waits in CQ for the next free lane: a reset in
        // the env's own lane -- one lane working, 31 waiting, in 40 % of the rounds -- cost 14 % of the launch. ----
        {
            bool do_reset = false;
            int r_e = 0; uint32_t r_ep = 0; float *r_o1 = nullptr, *r_o2 = nullptr;
            int r_k = 0;   // the sub-step that ended the episode
            if (rq_n >= QR_RESET_BATCH || (drained && rq_n > 0)) {   // (warp-uniform; at least one lane resets)
                // at most 32 per pass; a burst (e.g. a common time limit) leaves the rest for the next round.  Neither queue can
                // overflow: envs in flight (in a lane, in RQ or in CQ) only increase when a lane takes an env from the tile
                // sequence, which it does only when CQ is empty, i.e. when they number at most 32 + QR_RESET_BATCH - 1.
                const int n_now = min(rq_n, 32);
                if (lane < n_now) {
                    do_reset = true; r_e = rq[rq_n - n_now + lane];
                    if (MULTI) r_k = rqk[rq_n - n_now + lane] - 1;
                    const bool r_last = r_k == NS - 1;
                    r_ep = __ldcg(a.ep_index + r_e);   // written when the env was released (already incremented)
                    r_o1 = (MULTI && a.obs_roll) ? a.obs_roll + ((int64_t)r_k * N + r_e) * O : ((r_last || POLICY) ? a.obs + (int64_t)r_e * OS : nullptr);
                    r_o2 = (MULTI && a.obs_roll && (r_last || POLICY)) ? a.obs + (int64_t)r_e * OS : nullptr;
                }
                rq_n -= n_now;
                obs_in_tile = false;   // the reset scratch and the park area overwrite the tile
                __syncwarp();
                T* const pk = ks + 1024 + lane;   // park slot j of this lane: pk[j * 32] (the reset scratch is ks[0 .. 1023])
#define QR_PKI(j) (*reinterpret_cast<int32_t*>(pk + (j) * 32))
#pragma unroll
                for (int i = 0; i < 3; ++i) pk[i * 32] = x[i];
#pragma unroll
                for (int i = 0; i < 14; ++i) { pk[(3 + i) * 32] = y[i]; pk[(18 + i) * 32] = K0[i]; }
                pk[17 * 32] = W3;
                pk[32 * 32] = d.fm; pk[33 * 32] = d.g; pk[34 * 32] = d.Mi0; pk[35 * 32] = d.Mi1; pk[36 * 32] = d.kw0; pk[37 * 32] = d.kw1; pk[38 * 32] = d.w3dot;
                pk[39 * 32] = ode.t; pk[40 * 32] = ode.h_abs;
                QR_PKI(41) = ode.rejected; QR_PKI(42) = ode.nfev; QR_PKI(43) = ode.status; QR_PKI(44) = ode.nproj; QR_PKI(45) = ode.checked;
                if (MULTI) QR_PKI(46) = k;
                QR_PKI(47) = (int)busy | ((int)fin << 1) | ((int)need_init << 2) | ((int)has_next << 3) | ((int)(MULTI && fresh) << 5) | ((int)(MULTI && r_ok) << 6);
                QR_PKI(48) = e;
                QR_PKI(50) = e_next;
                QR_PKI(52) = tile_base;
                QR_PKI(54) = tile_pos; QR_PKI(55) = rq_n;
                if (MULTI) { QR_PKI(56) = cq_n; QR_PKI(57) = k_next; }
                if (do_reset) {
                    float dummy[23];
                    // the new episode's first observation replaces the terminal one in the step's output row
                    auto_reset_env<T, MODE>(&a, r_e, r_ep, r_o1 ? r_o1 : dummy, r_o2, ks + lane * 32);
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) x[i] = pk[i * 32];
#pragma unroll
                for (int i = 0; i < 14; ++i) { y[i] = pk[(3 + i) * 32]; K0[i] = pk[(18 + i) * 32]; }
                W3 = pk[17 * 32];
                d.fm = pk[32 * 32]; d.g = pk[33 * 32]; d.Mi0 = pk[34 * 32]; d.Mi1 = pk[35 * 32]; d.kw0 = pk[36 * 32]; d.kw1 = pk[37 * 32]; d.w3dot = pk[38 * 32];
                ode.t = pk[39 * 32]; ode.h_abs = pk[40 * 32];
                ode.rejected = QR_PKI(41); ode.nfev = QR_PKI(42); ode.status = QR_PKI(43); ode.nproj = QR_PKI(44); ode.checked = QR_PKI(45);
                if (MULTI) k = QR_PKI(46);
                {
                    const int fl = QR_PKI(47);
                    busy = fl & 1; fin = (fl >> 1) & 1; need_init = (fl >> 2) & 1; has_next = (fl >> 3) & 1; fresh = (fl >> 5) & 1; r_ok = (fl >> 6) & 1;
                }
                e = QR_PKI(48);
                e_next = QR_PKI(50);
                tile_base = QR_PKI(52);
                tile_pos = QR_PKI(54); rq_n = QR_PKI(55);
                if (MULTI) { cq_n = QR_PKI(56); k_next = QR_PKI(57); }
#undef QR_PKI
                if (do_reset) {
                    const T* sc = ks + lane * 32;   // what auto_reset_env left in this lane's scratch
#pragma unroll
                    for (int i = 0; i < 18; ++i) a.state[i * N + r_e] = sc[i];   // scratch order = state row order
#pragma unroll
                    for (int i = 0; i < 8; ++i) a.integ[i * N + r_e] = sc[18 + i];
                }
                if (MULTI) {   // envs with sub-steps left wait in CQ for a free lane (A0 of the coming rounds)
                    const bool cont = do_reset && (r_k + 1 < NS);
                    const unsigned cm = __ballot_sync(FULL, cont);
                    if (cont) {
                        const int q = cq_n + __popc(cm & ((1u << lane) - 1u));
                        cq[q] = r_e; cqk[q] = (int16_t)(r_k + 1);
                    }
                    cq_n += __popc(cm);
                }
                __syncwarp();
            }
        }
        if (drained && rq_n == 0 && cq_n == 0) break;
        // ---- A3: start the next env.step: goal, action, SO(3) check, f0 and the initial step size ----
        if (busy && need_init) {
            need_init = false;
#pragma unroll
            for (int i = 0; i < 14; ++i) K0[i] = 0;   // dead here on every path (liveness hint, as in A1)
            // every load of this phase is issued before the first consumer (in-order issue: a stalled
            // consumer would otherwise delay the independent loads behind it by a full memory round trip)
            T act[5], b1d[3];
            T p_m, p_J1, p_J3, p_ctw, p_d = 0, p_ctf = 0;
            bool act_f32 = a.act_f32 != 0;
            const bool staged_act = a.actions && (a.act_f32 || sizeof(T) == 8);   // what A0 can stage (kernel-uniform)
            const bool staged = MULTI ? fresh : true;   // adopted in this round's A2: action, parameters and b1d were fetched ahead into the stash
            fresh = false;
            // everything but A2's group (end-of-step values, not needed yet).  (A lane that keeps its env must not wait here:
            //  on its last sub-step the group it would wait for is the fetch-ahead of its NEXT env, issued a moment ago.)
            if (staged) cp_async_wait_group<1>();
            // the parameters stay in the landing zone of the stash for as long as the env stays in this lane (only a lane that is
            // idle or on its last sub-step is handed a next env, A0): later sub-steps read them from there, not from HBM
            p_m = sh[(S_PAR + 0) * 32]; p_J1 = sh[(S_PAR + 2) * 32]; p_J3 = sh[(S_PAR + 3) * 32]; p_ctw = sh[(S_PAR + 5) * 32];
            if (MODE == 0) { p_d = sh[(S_PAR + 1) * 32]; p_ctf = sh[(S_PAR + 4) * 32]; }
            if (GOAL1) {   // fetched ahead -- or the b1d of the step before (constant within an episode in mode 0), kept for the observation
                const T* const bp = sh + (staged ? S_NB1D : S_B1D) * 32;
#pragma unroll
                for (int i = 0; i < 3; ++i) b1d[i] = bp[i * 32];
            }
            if (a.actions) {   // (one test on the path of in-kernel actions)
                if (staged && staged_act) {
#pragma unroll
                    for (int i = 0; i < A; ++i)
                        act[i] = a.act_f32 ? (T)*reinterpret_cast<const float*>(sh + (S_ACT + i) * 32) : sh[(S_ACT + i) * 32];
                } else {
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
            }
            if (!GOAL1 && !(MODE != 0 && c.goal_mode >= 2)) {   // the observation at the end of this step reads the goal: have it in L2 by then
#pragma unroll
                for (int i = 0; i < 12; ++i) prefetch_l2(a.goal + i * N + e);
            }
            if (POLICY) {
                // obs -> action with the compiled actor(s); the row was written by this env's previous step (by this
                // thread, or by its warp through the shared tile / by an earlier launch)
                float xo[24], af[5];
                if (MULTI && !staged && obs_in_tile) {
                    const float4* row = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ks) + lane * OS);
#pragma unroll
                    for (int i = 0; i < OS / 4; ++i) { const float4 v = row[i]; xo[4 * i] = v.x; xo[4 * i + 1] = v.y; xo[4 * i + 2] = v.z; xo[4 * i + 3] = v.w; }
                } else {
#pragma unroll
                    for (int i = 0; i < O; ++i) xo[i] = a.obs[(int64_t)e * OS + i];
                }
                obs_in_tile = false;
                if (MODE == 1) actor_td3_mono(xo, af);
                else { actor_td3_modul1(xo, af); actor_td3_modul2(xo + 15, af + 4); }
#pragma unroll
                for (int i = 0; i < A; ++i) act[i] = (T)af[i];
                act_f32 = true;   // the torch actor returns float32 actions (numpy then computes the thrust in float32)
            } else if (!a.actions) {
                const uint64_t gid = (uint64_t)(a.env_id_offset + e);
                uint32_t rnd[8];
                cp_async_wait_all();
                const uint32_t ep_idx = (uint32_t)stash_i32(sh, S_IDX);
                const int ep_len = stash_i32(sh, S_LEN);
                ph.template run<true>((uint32_t)gid, (uint32_t)(gid >> 32), ep_idx, QR_DOMAIN_ACTION + 2u * (uint32_t)ep_len, rnd);
                if (A == 5) ph.template run<true>((uint32_t)gid, (uint32_t)(gid >> 32), ep_idx, QR_DOMAIN_ACTION + 2u * (uint32_t)ep_len + 1u, rnd + 4);
#pragma unroll
                for (int i = 0; i < A; ++i) act[i] = (sizeof(T) == 8) ? (T)(2.0 * u01(rnd[i]) - 1.0) : (T)2 * u01t<T>(rnd[i]) - (T)1;   // U(-1, 1)
                act_f32 = false;
            }
            // state_decomposition of the incoming state: get_desired (trajectory_generator.py:115) and
            // observation_wrapper (coupled:58) both run ensure_SO3 on the same R; once is enough
            // ... and in a multi-step launch the observation that ended the previous sub-step (get_norm_error_state ->
            // state_normalization) has just run the same test on the same matrix: if it passed there, it passes here
            int fl = 0;
            if (!(MULTI && MODE != 0 && !staged && r_ok)) fl = ensure_so3<T>(y + 3);
            EnvRegs<T> r;
#pragma unroll
            for (int i = 0; i < 3; ++i) r.x[i] = x[i];
#pragma unroll
            for (int i = 0; i < 14; ++i) r.y[i] = y[i];
            r.W3 = W3;
            r.m = p_m; r.d = p_d; r.J1 = p_J1; r.J3 = p_J3; r.c_tf = p_ctf; r.c_tw = p_ctw;
            if (!GOAL1 && MODE != 0 && c.goal_mode >= 2) {
                // trajectory_generator.get_desired(state, mode) + set_goal_state on the pre-step state, as the trainer calls
                // them before every env.step (main.py:145-147): hover / circle / eight / take-off / land / stay (kernel-uniform
                // branch).  The goal of the step (xd, vd, b1d.xy, Wd.z: the other three are zero in every mode) and what
                // changes per call of the trajectory state (clock, flags, b1d_dot) live in nine + four stash slots that this
                // configuration does not use otherwise; an env that keeps its lane reads only the trajectory's constants from
                // HBM and writes nothing back before its last sub-step.
                T ts[12], gl[12];
                const bool from_hbm = !MULTI || staged;
                if (from_hbm) {
#pragma unroll
                    for (int i = 0; i < 12; ++i) { ts[i] = a.traj[i * N + e]; gl[i] = a.goal[i * N + e]; }
                } else {
#pragma unroll
                    for (int i = 2; i < 9; ++i) ts[i] = a.traj[i * N + e];
                    ts[0] = sh[(S_ACT + 0) * 32]; ts[1] = sh[(S_ACT + 1) * 32]; ts[9] = sh[(S_ACT + 2) * 32]; ts[10] = sh[(S_ACT + 3) * 32];
                    ts[11] = 0;
#pragma unroll
                    for (int i = 0; i < 3; ++i) { gl[i] = sh[(S_B1D + i) * 32]; gl[3 + i] = sh[(S_WD + i) * 32]; }
                    gl[6] = sh[(S_NB1D + 0) * 32]; gl[7] = sh[(S_NB1D + 1) * 32]; gl[8] = 0;
                    gl[9] = 0; gl[10] = 0; gl[11] = sh[(S_NB1D + 2) * 32];
                }
                const T Wv[3] = {y[12], y[13], W3};
                const int fl0 = (int)ts[1];
                traj_desired<T>(traj_ref_mode<>(c.goal_mode), x, y, y + 3, Wv, ts, gl, (T)0, (T)0, c.dt);
#pragma unroll
                for (int i = 0; i < 3; ++i) { sh[(S_B1D + i) * 32] = gl[i]; sh[(S_WD + i) * 32] = gl[3 + i]; }
                sh[(S_NB1D + 0) * 32] = gl[6]; sh[(S_NB1D + 1) * 32] = gl[7]; sh[(S_NB1D + 2) * 32] = gl[11];
                if (MULTI) { sh[(S_ACT + 0) * 32] = ts[0]; sh[(S_ACT + 1) * 32] = ts[1]; sh[(S_ACT + 2) * 32] = ts[9]; sh[(S_ACT + 3) * 32] = ts[10]; }
                if (k == NS - 1) {   // the env leaves the lane after this step: goal and trajectory state back to their arrays
#pragma unroll
                    for (int i = 0; i < 12; ++i) a.goal[i * N + e] = gl[i];
                    a.traj[0 * N + e] = ts[0]; a.traj[1 * N + e] = ts[1]; a.traj[9 * N + e] = ts[9]; a.traj[10 * N + e] = ts[10];
                }
                // the rest of the trajectory state is set when a trajectory (or the manual mode after it) starts
                if (((int)ts[1] ^ fl0) & 5) {
#pragma unroll
                    for (int i = 2; i < 9; ++i) a.traj[i * N + e] = ts[i];
                }
            }
            if (GOAL1) {   // goal from the pre-step state, main.py:145-147
                const T Wv[3] = {y[12], y[13], W3};
                T Wd[3];
                traj_wd<T>(y + 3, Wv, b1d, Wd);
#pragma unroll
                for (int i = 0; i < 3; ++i) { sh[(S_B1D + i) * 32] = b1d[i]; sh[(S_WD + i) * 32] = Wd[i]; }   // for the observation at the end
#pragma unroll
                for (int i = 0; i < 3; ++i) st_if(k == NS - 1, a.goal + (9 + i) * N + e, Wd[i]);   // visible in the goal buffer like env.Wd after set_goal_state
            }
            T f, M[3];
            action_to_fM<T>(r, c, act, act_f32, f, M, MODE);
            {
                // inv(J) as the reference forms it (quad.py:329); float32 mode: MUFU reciprocals (<= 1 ulp)
                const T rm = num<T>::recip(p_m), rJ1 = num<T>::recip(p_J1), rJ3 = num<T>::recip(p_J3);
                d.fm = f * rm; d.g = c.g;
                d.Mi0 = M[0] * rJ1; d.Mi1 = M[1] * rJ1;
                d.kw0 = (p_J1 - p_J3) * rJ1; d.kw1 = (p_J3 - p_J1) * rJ1;
                d.w3dot = M[2] * rJ3;
            }
            // all 18 state words finite?  0 * v is (+-)0 for a finite v and NaN otherwise: one probe sum instead of 18 comparisons
            // (pairs as the integrator's internal order holds them, qr_dop853.cuh)
            T fp0 = 0, fp1 = 0;
            pfma<T>((T)0, y[3], y[4], fp0, fp1, fp0, fp1); pfma<T>((T)0, y[6], y[7], fp0, fp1, fp0, fp1);
            pfma<T>((T)0, y[9], y[10], fp0, fp1, fp0, fp1); pfma<T>((T)0, y[5], y[8], fp0, fp1, fp0, fp1);
            pfma<T>((T)0, y[11], y[2], fp0, fp1, fp0, fp1); pfma<T>((T)0, y[0], y[1], fp0, fp1, fp0, fp1);
            pfma<T>((T)0, y[12], y[13], fp0, fp1, fp0, fp1); pfma<T>((T)0, x[0], x[1], fp0, fp1, fp0, fp1);
            pfma<T>((T)0, x[2], W3, fp0, fp1, fp0, fp1);
            const bool finite = (fp0 + fp1) == (T)0;
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
            // rare: some stage matrix of the speculative attempt failed the SO(3) test -> the lane redoes the attempt out
            // of line with the reference's per-stage re-projection.  The call works on copies, so that the loop's
            // registers never have their address taken.
            if (__any_sync(FULL, live && ode.checked != 0)) {
                if (live && ode.checked != 0) {
                    T tx[3], ty[14], tK[14], tW3 = W3;   // ty: internal order (see dop853_attempt_checked)
                    Dyn<T> td = d;
                    OdeLane<T> to = ode;
#pragma unroll
                    for (int i = 0; i < 3; ++i) tx[i] = x[i];
#pragma unroll
                    for (int i = 0; i < 14; ++i) { ty[zof(i)] = y[i]; tK[i] = K0[i]; }
                    fin = dop853_attempt_checked<T>(tx, ty, &tW3, &td, c.dt, c.rtol, c.atol, tK, &to);
#pragma unroll
                    for (int i = 0; i < 3; ++i) x[i] = tx[i];
#pragma unroll
                    for (int i = 0; i < 14; ++i) { y[i] = ty[zof(i)]; K0[i] = tK[i]; }
                    W3 = tW3; ode = to;
                }
            }
        }
    }

    // ---- flush this warp's statistics: one atomic per non-zero statistic ----
    __syncwarp();
    if (lane < QR_NSTATS && ws[lane] != 0.0) atomicAdd(&a.stats[lane], ws[lane]);
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
    traj_desired<T>(traj_ref_mode<>(a.c.goal_mode), x, v, R, W, ts, goal, (T)0, (T)0, a.c.dt);
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
    const int OS = obs_stride_of(O);
    for (int i = 0; i < O; ++i) a.obs[e * OS + i] = o[i];
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
    constexpr int OS = obs_stride_of(O);
    __shared__ float tile[QR_BLOCK * OS];
    const int tid = threadIdx.x;
    const int64_t e0 = (int64_t)blockIdx.x * QR_BLOCK;
    const int64_t rem = n - e0;
    const int nvalid = (int)(rem < QR_BLOCK ? rem : QR_BLOCK);
    for (int i = tid; i < nvalid * OS; i += QR_BLOCK) tile[i] = obs[e0 * OS + i];
    __syncthreads();
    if (tid >= nvalid) return;
    float x[O], a[A];
#pragma unroll
    for (int i = 0; i < O; ++i) x[i] = tile[tid * OS + i];
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

// ---- observation rows without their padding ([rows][OS] -> [rows][O]); qr_step_host copies the dense block to the host ----
static __global__ void k_dense_rows(const float* __restrict__ src, float* __restrict__ dst, int64_t rows, int O, int OS)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * O) return;
    const int64_t r = i / O; const int c = (int)(i - r * O);
    dst[i] = src[r * OS + c];
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

// ---- the step kernel's instantiations live in their own translation units (qr_step_tu.cu, compiled in parallel) ----
template <typename T> using step_kernel_t = void (*)(const StepArgs<T>);
#define QR_DECL_STEP(T, M, P) step_kernel_t<T> step_kernel_##T##_m##M##_p##P(bool multi, bool goal1);
QR_DECL_STEP(float, 0, 0) QR_DECL_STEP(float, 1, 0) QR_DECL_STEP(float, 2, 0) QR_DECL_STEP(float, 1, 1) QR_DECL_STEP(float, 2, 1)
QR_DECL_STEP(double, 0, 0) QR_DECL_STEP(double, 1, 0) QR_DECL_STEP(double, 2, 0) QR_DECL_STEP(double, 1, 1) QR_DECL_STEP(double, 2, 1)
#undef QR_DECL_STEP
template <typename T> inline step_kernel_t<T> step_kernel(int mode, bool multi, bool goal1, bool policy);
template <> inline step_kernel_t<float> step_kernel<float>(int mode, bool multi, bool goal1, bool policy)
{
    if (policy) return mode == 1 ? step_kernel_float_m1_p1(multi, goal1) : step_kernel_float_m2_p1(multi, goal1);
    return mode == 1 ? step_kernel_float_m1_p0(multi, goal1) : mode == 2 ? step_kernel_float_m2_p0(multi, goal1) : step_kernel_float_m0_p0(multi, goal1);
}
template <> inline step_kernel_t<double> step_kernel<double>(int mode, bool multi, bool goal1, bool policy)
{
    if (policy) return mode == 1 ? step_kernel_double_m1_p1(multi, goal1) : step_kernel_double_m2_p1(multi, goal1);
    return mode == 1 ? step_kernel_double_m1_p0(multi, goal1) : mode == 2 ? step_kernel_double_m2_p0(multi, goal1) : step_kernel_double_m0_p0(multi, goal1);
}

}  // namespace qr
