// ubench_pipes.cu -- issue / pipe rates that decide the design of the step kernel's integrator (sm_100a):
// scalar FFMA vs packed FFMA2 (fma.rn.f32x2) vs mixes with ALU / LDS work, at 4, 8 and 12 warps per SM.
// Prints per SM and clock: warp instructions issued and FP32 lane-FMAs (FFMA2 counts 2 per lane).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_ab/ubench_pipes tools/ubench_pipes.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ILP = 8;
constexpr int INNER = 32;

// every operation is an `asm volatile` so that nvcc keeps exactly the instruction mix written here
__device__ __forceinline__ void op_ffma(float& a, float b, float c) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c)); }
__device__ __forceinline__ void op_ffma2(unsigned long long& a, unsigned long long b, unsigned long long c) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(b), "l"(c)); }
__device__ __forceinline__ void op_fmul2(unsigned long long& a, unsigned long long b) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b)); }
__device__ __forceinline__ void op_alu(unsigned& u, unsigned v) { asm volatile("lop3.b32 %0, %0, %1, 0x5a5a5a5a, 0x96;" : "+r"(u) : "r"(v)); }
__device__ __forceinline__ void op_lds(float4& v, unsigned addr) { asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr)); }

// MODE 0: FFMA ; 1: FFMA2 ; 2: FFMA2 + FFMA alternating ; 3: FFMA2 + LOP3 alternating ; 4: 4 FFMA2 + 1 LDS.128 ;
// 5: FFMA + LOP3 alternating ; 6: FMUL2 ; 7: FFMA2 + 2 LOP3 ; 8: FFMA + FFMA + LOP3... (2:1) ; 9: FFMA2 with a shared multiplier register
template <int MODE> __global__ void __launch_bounds__(384, 1) k_pipe(float* out, int iters, float seed)
{
    __shared__ float4 sm[512];
    sm[threadIdx.x] = make_float4(seed, seed, seed, seed);
    __syncthreads();
    float a[ILP], b[ILP];
    unsigned long long A[ILP], B[ILP], C;
    unsigned u[ILP], w[ILP];
    for (int i = 0; i < ILP; ++i) {
        a[i] = seed + i; b[i] = seed * 0.999f + 1e-7f * i; u[i] = threadIdx.x * 7 + i; w[i] = threadIdx.x * 3 + i;
        A[i] = ((unsigned long long)__float_as_uint(seed + i) << 32) | __float_as_uint(seed - i);
        B[i] = ((unsigned long long)__float_as_uint(seed * 0.999f + 1e-7f * i) << 32) | __float_as_uint(seed * 0.998f - 1e-7f * i);
    }
    C = ((unsigned long long)__float_as_uint(seed * 1e-9f) << 32) | __float_as_uint(seed * 2e-9f);
    const float c = seed * 1e-9f;
    float4 ld = make_float4(0, 0, 0, 0);
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < INNER; ++j) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (MODE == 0) op_ffma(a[i], b[i], c);
                if (MODE == 1) op_ffma2(A[i], B[i], C);
                if (MODE == 2) { if (i & 1) op_ffma2(A[i], B[i], C); else op_ffma(a[i], b[i], c); }
                if (MODE == 3) { if (i & 1) op_ffma2(A[i], B[i], C); else op_alu(u[i], w[i]); }
                if (MODE == 4) { op_ffma2(A[i], B[i], C); if ((i & 3) == 0) op_lds(ld, sa + ((j & 7) * 512)); }
                if (MODE == 5) { if (i & 1) op_ffma(a[i], b[i], c); else op_alu(u[i], w[i]); }
                if (MODE == 6) op_fmul2(A[i], B[i]);
                if (MODE == 7) { op_ffma2(A[i], B[i], C); op_alu(u[i], w[i]); op_alu(w[i], u[i]); }
                if (MODE == 8) { op_ffma(a[i], b[i], c); op_ffma(b[i], a[i], c); op_alu(u[i], w[i]); }
                if (MODE == 9) op_ffma2(A[i], B[0], C);
            }
        }
    }
    float s = ld.x + ld.y;
    for (int i = 0; i < ILP; ++i) s += a[i] + b[i] + (float)u[i] + (float)w[i] + (float)(A[i] >> 32) + (float)(A[i] & 0xffffffffu);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char* name, int sms, float clock_ghz, double fma_per_thread_iter, double inst_per_thread_iter)
{
    float* out; CK(cudaMalloc(&out, sizeof(float) * sms * 384));
    const int iters = 2000;
    for (int warps : {4, 8, 12}) {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        k_pipe<MODE><<<sms, warps * 32>>>(out, 50, 1.0f);
        CK(cudaEventRecord(e0));
        k_pipe<MODE><<<sms, warps * 32>>>(out, iters, 1.0f);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double cycles = ms * 1e-3 * clock_ghz * 1e9;
        const double lane_fma = fma_per_thread_iter * iters * warps * 32 / cycles;       // per SM per clock
        const double winst = inst_per_thread_iter * iters * warps / cycles;
        printf("%-28s warps %2d  %.3f ms  lane-FMA/clk/SM %6.1f  warp-inst/clk/SM %.2f\n", name, warps, ms, lane_fma, winst);
    }
    CK(cudaFree(out));
}

int main()
{
    int dev = 0, sms = 0, khz = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    const float ghz = khz * 1e-6f;
    printf("SMs %d, clock %.3f GHz (rates below assume that clock)\n", sms, ghz);
    const double N = (double)INNER * ILP;   // per thread and outer iteration: (lane FMAs, instructions)
    run<0>("FFMA", sms, ghz, N, N);
    run<1>("FFMA2", sms, ghz, 2 * N, N);
    run<9>("FFMA2, shared multiplier", sms, ghz, 2 * N, N);
    run<6>("FMUL2", sms, ghz, 2 * N, N);
    run<2>("FFMA2 + FFMA 1:1", sms, ghz, 1.5 * N, N);
    run<3>("FFMA2 + LOP3 1:1", sms, ghz, N, N);
    run<7>("FFMA2 + 2 LOP3", sms, ghz, 2 * N, 3 * N);
    run<5>("FFMA + LOP3 1:1", sms, ghz, 0.5 * N, N);
    run<8>("2 FFMA + LOP3", sms, ghz, 2 * N, 3 * N);
    run<4>("4 FFMA2 + 1 LDS.128", sms, ghz, 2 * N, 1.25 * N);
    return 0;
}
