// lm.cu — Levenberg-Marquardt refinement of sphere / cylinder / cone (K7). PLACEHOLDER: copies the
// un-refined model and reports status -100 until the device LM lands.
#include "pitt_common.cuh"
#include "sac.cuh"

namespace pitt {
__global__ void lm_copy_kernel(const float* __restrict__ m, float* __restrict__ r, int* __restrict__ info) {
  if (threadIdx.x < 8) r[threadIdx.x] = m[threadIdx.x];
  if (threadIdx.x == 0) { info[0] = -100; info[1] = 0; }
}
int lm_refine(pitt_ctx* ctx, const pitt_cloud*, int, const float* d_model, const int*, const int*, int, float* d_refined,
              int* d_lm_info) {
  lm_copy_kernel<<<1, 32, 0, ctx->stream>>>(d_model, d_refined, d_lm_info);
  ctx->launches++;
  return PITT_OK;
}
}  // namespace pitt
