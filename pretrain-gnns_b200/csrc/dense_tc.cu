// Tensor-core (tcgen05, 3xTF32) dense path — placeholder until the kernels land: every entry reports
// PGNN_EUNSUPPORTED so that precision=1 requests take the fp32 FFMA path of dense.cu.
#include "common.cuh"

int pgnn_tc_linear_fwd(const float*, int64_t, const float*, const float*, int64_t, int64_t, int64_t, int, float*, int64_t,
                       cudaStream_t) { return PGNN_EUNSUPPORTED; }
int pgnn_tc_linear_bwd_x(const float*, int64_t, const float*, int64_t, int64_t, int64_t, const float*, int64_t, float*, int64_t,
                         cudaStream_t) { return PGNN_EUNSUPPORTED; }
int pgnn_tc_linear_bwd_w(const float*, int64_t, const float*, int64_t, int64_t, int64_t, int64_t, float*, float*, cudaStream_t) {
  return PGNN_EUNSUPPORTED;
}
