// qpwc_corr_tiled.cu -- register-tiled correlation kernels (placeholder: filled in next commit).
#include "qpwc_common.cuh"

namespace qpwc {

int launch_corr_fwd_tiled(const float*, const float*, const float*, int, float*, int, int, int, int,
                          int, float, long long, cudaStream_t) {
  return QPWC_ERR_UNSUPPORTED;
}
int launch_corr_bwd_tiled(const float*, const float*, const float*, const float*, float*, float*,
                          int, int, int, int, int, float, long long, cudaStream_t) {
  return QPWC_ERR_UNSUPPORTED;
}

}  // namespace qpwc
