// tcgen05 / TMEM / TMA bf16-split GEMM path (FA_GEMM_BF16X1 / X3 / X6).  Placeholder until the kernel lands:
// every call reports FA_ERR_UNSUPPORTED so that nothing silently falls back to another path.
#include "common.cuh"
#include "kernels.h"

namespace fa {

size_t gemm_tc_scratch_bytes(int64_t, int, int) { return 0; }

int gemm_tc_launch(const float*, int64_t, int64_t, const FaLinear&, int, const float*, int64_t, const float*, int64_t,
                   float*, int64_t, int, Arena*, cudaStream_t) {
  return FA_ERR_UNSUPPORTED;
}

}  // namespace fa
