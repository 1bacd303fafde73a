// sass_probe.cu -- compiles ONLY the headline instantiations of the step kernel, for quick -Xptxas -v / SASS checks while
// editing the kernel headers (the product library has 28 instantiations and takes minutes):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Igym_rotor_b200/csrc -Xptxas -v -cubin -o /tmp/probe/probe.cubin tools/sass_probe.cu
#include "qr_kernels.cuh"
#ifndef PROBE_T
#define PROBE_T float
#endif
#ifndef PROBE_MULTI
#define PROBE_MULTI false
#endif
template __global__ void qr::k_step<PROBE_T, 1, PROBE_MULTI, true, false>(const __grid_constant__ qr::StepArgs<PROBE_T>);
