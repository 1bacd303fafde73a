// ubench_layout.cu -- micro-benchmarks for the memory-layout decisions queued in profiles/r01_next_steps.md.
// Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_layout tools/ubench_layout.cu && /tmp/ubench_layout
//
// Each kernel moves what one env.step() launch of 2^21 envs moves for ONE of the two traffic classes, with the access
// pattern of a candidate layout, one env per thread (consecutive threads = consecutive envs, as in k_step's common case):
//   obs rows   : (a) 23 scattered 4-byte stores per env, row stride 92 B (today's general path)
//                (b) 6 x 16-byte stores per env, row stride 96 B (rows padded to 24 floats)
//                (c) rows staged in shared memory and written as one contiguous block per warp (today's lock-step path)
//   env state  : (d) 32 scalar loads from structure-of-arrays [32][N] with 64-bit address chains (today)
//                (e) the same components in tiles [N/32][32][32] ("AoSoA"): one base per env + immediate offsets
//                (f) array of 128-byte records [N][32]: 8 x 16-byte loads per env
// The numbers say how much of the step's ~650 us each pattern costs on its own (they overlap with compute in k_step,
// so they bound the gain from above), and whether the L2/LSU request count or the DRAM bytes dominate.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void obs_scatter(float* obs, int64_t n)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float v = (float)e;
#pragma unroll
    for (int i = 0; i < 23; ++i) obs[e * 23 + i] = v + i;
}
__global__ void obs_padded(float* obs, int64_t n)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float v = (float)e;
    float4* row = reinterpret_cast<float4*>(obs + e * 24);
#pragma unroll
    for (int i = 0; i < 6; ++i) row[i] = make_float4(v + 4 * i, v + 4 * i + 1, v + 4 * i + 2, v + 4 * i + 3);
}
__global__ void obs_tile(float* obs, int64_t n)
{
    __shared__ __align__(16) float tile[8][32 * 23];   // 256 threads = 8 warps
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float v = (float)e;
#pragma unroll
    for (int i = 0; i < 23; ++i) tile[w][lane * 23 + i] = v + i;
    __syncwarp();
    const int64_t e0 = e - lane;
    if (e0 + 32 <= n) {
        const float4* t4 = reinterpret_cast<const float4*>(tile[w]);
        float4* g = reinterpret_cast<float4*>(obs + e0 * 23);
#pragma unroll
        for (int it = 0; it < 6; ++it) { const int q = it * 32 + lane; if (q < 184) g[q] = t4[q]; }
    }
}
__global__ void state_soa(const float* __restrict__ st, float* out, int64_t n)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += st[i * n + e];
    if (s == 123.456f) out[e] = s;   // keep the loads alive without storing
}
__global__ void state_aosoa(const float* __restrict__ st, float* out, int64_t n)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const float* p = st + (e >> 5) * (32 * 32) + (e & 31);
    float s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += p[i * 32];
    if (s == 123.456f) out[e] = s;
}
__global__ void state_aos(const float* __restrict__ st, float* out, int64_t n)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const float4* p = reinterpret_cast<const float4*>(st + e * 32);
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float4 v = p[i]; s += v.x + v.y + v.z + v.w; }
    if (s == 123.456f) out[e] = s;
}

template <typename F> float time_ms(F launch, int reps)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

int main()
{
    const int64_t n = 1 << 21;
    const int B = 256, G = (int)((n + B - 1) / B), reps = 50;
    float *obs, *st, *out, *flush;
    CK(cudaMalloc(&obs, n * 24 * sizeof(float)));
    CK(cudaMalloc(&st, n * 32 * sizeof(float)));
    CK(cudaMalloc(&out, n * sizeof(float)));
    CK(cudaMalloc(&flush, 256u << 20));
    CK(cudaMemset(st, 0, n * 32 * sizeof(float)));
    auto fl = [&]() { CK(cudaMemsetAsync(flush, 1, 256u << 20)); };   // evict L2 between launches (126 MB)
    struct { const char* name; double bytes; float ms; } r[6];
    r[0] = {"obs  (a) 23 x STG.32, stride 92 B ", (double)n * 92, time_ms([&] { fl(); obs_scatter<<<G, B>>>(obs, n); }, reps)};
    r[1] = {"obs  (b) 6 x STG.128, stride 96 B ", (double)n * 96, time_ms([&] { fl(); obs_padded<<<G, B>>>(obs, n); }, reps)};
    r[2] = {"obs  (c) shared tile, contiguous  ", (double)n * 92, time_ms([&] { fl(); obs_tile<<<G, B>>>(obs, n); }, reps)};
    r[3] = {"state (d) SoA [32][N]             ", (double)n * 128, time_ms([&] { fl(); state_soa<<<G, B>>>(st, out, n); }, reps)};
    r[4] = {"state (e) AoSoA [N/32][32][32]    ", (double)n * 128, time_ms([&] { fl(); state_aosoa<<<G, B>>>(st, out, n); }, reps)};
    r[5] = {"state (f) AoS [N][32], LDG.128    ", (double)n * 128, time_ms([&] { fl(); state_aos<<<G, B>>>(st, out, n); }, reps)};
    const float t_flush = time_ms([&] { fl(); }, reps);
    printf("L2 flush alone: %.1f us (subtracted below)\n", t_flush * 1e3f);
    for (auto& x : r) printf("%s %8.1f us  %7.1f GB/s\n", x.name, (x.ms - t_flush) * 1e3f, x.bytes / ((x.ms - t_flush) * 1e-3) / 1e9);
    return 0;
}
