// ubench_attempt.cu -- the integrator alone: every lane runs `iters` full env.step() integrations (dop853_begin + attempts
// until the 5 ms interval is done) on a state kept in registers, with the stage storage in shared memory exactly as in
// qr::k_step (12 warps per CTA, one CTA per SM), but WITHOUT phase A (no loads, stores, observations, rewards, resets).
// It gives the ceiling the step kernel would reach if everything around the integrator were free, and a small target for
// ncu when tuning the stage loop (60 % of k_step's instructions).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Igym_rotor_b200/csrc -o /tmp/ubench_attempt tools/ubench_attempt.cu
//   /tmp/ubench_attempt            # prints integrations/s; compare with bench.py's env-steps/s
#include "qr_env.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

using namespace qr;

__global__ void __launch_bounds__(384, 1) k_attempts(const float* __restrict__ init, float* __restrict__ out, int iters, unsigned long long* attempts)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* ks = reinterpret_cast<float*>(smem_raw) + (size_t)warp * QR_NSLOTS * QR_SLOT_ELEMS;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float x0[3], y0[14], W30;
    for (int i = 0; i < 3; ++i) x0[i] = init[tid * 24 + i];
    for (int i = 0; i < 14; ++i) y0[i] = init[tid * 24 + 3 + i];
    W30 = init[tid * 24 + 17];
    Dyn<float> d;
    d.fm = init[tid * 24 + 18]; d.g = 9.81f; d.Mi0 = init[tid * 24 + 19]; d.Mi1 = init[tid * 24 + 20];
    d.kw0 = -0.59f; d.kw1 = 0.59f; d.w3dot = init[tid * 24 + 21];
    float acc = 0;
    unsigned long long natt = 0;
    for (int it = 0; it < iters; ++it) {
        float x[3], y[14], W3 = W30, K0[14];
        for (int i = 0; i < 3; ++i) x[i] = x0[i];
        for (int i = 0; i < 14; ++i) y[i] = y0[i];
        OdeLane<float> ode;
        dop853_begin<float>(x, y, W3, d, 0.005f, 1e-3f, 1e-6f, K0, ode);
        bool fin = false;
        // warp-uniform loop as in k_step: a lane that is done keeps executing attempts on its state without committing
        while (__any_sync(0xffffffffu, !fin)) {
            const bool live = !fin;
            const bool f2 = dop853_attempt<float>(x, y, W3, d, 0.005f, 1e-3f, 1e-6f, K0, ode, ks, lane, live);
            if (live) { fin = f2; natt += 1; }
        }
        acc += y[0] + x[2] + W3;
        d.Mi0 = -d.Mi0;   // vary the problem a little from one integration to the next
    }
    out[tid] = acc;
    natt = __reduce_add_sync(0xffffffffu, (unsigned)natt);
    if (lane == 0) atomicAdd(attempts, natt);
}

int main()
{
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

    const int threads = 384, iters = 64;
    const int64_t n = (int64_t)sms * threads;
    std::vector<float> h(n * 24, 0.f);
    srand(1);
    auto u = [] { return rand() / (float)RAND_MAX * 2.f - 1.f; };
    for (int64_t t = 0; t < n; ++t) {
        float* p = &h[t * 24];
        for (int i = 0; i < 3; ++i) { p[i] = 0.5f * u(); p[3 + i] = 2.f * u(); }
        const float yaw = 3.14f * u(), c = cosf(yaw), s = sinf(yaw);
        p[6] = c; p[7] = s; p[8] = 0; p[9] = -s; p[10] = c; p[11] = 0; p[12] = 0; p[13] = 0; p[14] = 1;   // R = Rz(yaw), column-major
        p[15] = 3.f * u(); p[16] = 3.f * u(); p[17] = 3.f * u();
        p[18] = 9.81f * (1.f + 0.8f * u()); p[19] = 40.f * u(); p[20] = 40.f * u(); p[21] = 25.f * u();   // f/m, M1/J1, M2/J1, M3/J3
    }
    float *d_init, *d_out; unsigned long long* d_att;
    CK(cudaMalloc(&d_init, n * 24 * sizeof(float))); CK(cudaMalloc(&d_out, n * sizeof(float))); CK(cudaMalloc(&d_att, 8));
    CK(cudaMemcpy(d_init, h.data(), n * 24 * sizeof(float), cudaMemcpyHostToDevice));
    const size_t smem = (size_t)(threads / 32) * QR_NSLOTS * QR_SLOT_ELEMS * sizeof(float);
    CK(cudaFuncSetAttribute(k_attempts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(d_att, 0, 8));
        CK(cudaEventRecord(a));
        k_attempts<<<sms, threads, smem>>>(d_init, d_out, iters, d_att);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); CK(cudaGetLastError());
        float ms = 0; CK(cudaEventElapsedTime(&ms, a, b));
        unsigned long long att = 0; CK(cudaMemcpy(&att, d_att, 8, cudaMemcpyDeviceToHost));
        printf("integrations %lld x %d in %.3f ms -> %.3f G integrations/s, %.3f attempts per integration\n",
               (long long)n, iters, ms, n * (double)iters / (ms * 1e-3) / 1e9, att / (double)(n * iters));
    }
    return 0;
}
