// awq_gemm.cu -- placeholder translation unit for the tcgen05/TMEM AWQ loss GEMM (filled in by the next milestone).
#include "../../include/b200q.h"
#include "common.cuh"
using namespace b200q;
extern "C" {
int64_t b200q_awq_gemm_loss_workspace(int64_t, int64_t, int64_t, int32_t) { return 0; }
int b200q_awq_gemm_loss(const void*, int64_t, int64_t, const void*, const void*, int64_t, int32_t, float*, void*, int64_t, void*) {
    set_error("b200q_awq_gemm_loss: not built yet");
    return B200Q_ENOSYS;
}
}
