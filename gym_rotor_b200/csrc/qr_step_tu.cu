// qr_step_tu.cu -- one translation unit per (dtype, mode, policy) group of step-kernel instantiations, so that the
// 28 instantiations of qr::k_step compile in parallel (gym_rotor_b200/build.py passes QR_TU_T / QR_TU_MODE / QR_TU_POLICY).
// Each unit exports one selector that returns the kernel for (multi, goal1); quadrotor_b200.cu launches through it.
#include "qr_kernels.cuh"

#ifndef QR_TU_T
#error "QR_TU_T (float|double), QR_TU_MODE (0|1|2) and QR_TU_POLICY (0|1) are set by the build"
#endif

namespace qr {

#define QR_TU_CAT2(a, b, c, d) step_kernel_##a##_m##b##_p##c
#define QR_TU_CAT(a, b, c) QR_TU_CAT2(a, b, c, 0)
#define QR_TU_FN QR_TU_CAT(QR_TU_T, QR_TU_MODE, QR_TU_POLICY)

step_kernel_t<QR_TU_T> QR_TU_FN(bool multi, bool goal1)
{
#if QR_TU_POLICY
    (void)multi;   // the policy variants exist for the in-kernel reset flavour only
    return goal1 ? k_step<QR_TU_T, QR_TU_MODE, true, true, true> : k_step<QR_TU_T, QR_TU_MODE, true, false, true>;
#elif QR_TU_MODE == 0
    (void)goal1;   // on-device goal generation needs a wrapper mode
    return multi ? k_step<QR_TU_T, 0, true, false> : k_step<QR_TU_T, 0, false, false>;
#else
    return multi ? (goal1 ? k_step<QR_TU_T, QR_TU_MODE, true, true> : k_step<QR_TU_T, QR_TU_MODE, true, false>)
                 : (goal1 ? k_step<QR_TU_T, QR_TU_MODE, false, true> : k_step<QR_TU_T, QR_TU_MODE, false, false>);
#endif
}

}  // namespace qr
