// A9: LayerNorm over the embedding dimension of the concatenated TPV token sequence (tpvformer_encoder_layer.py:185-196,
// nn.LayerNorm(C), eps 1e-5), one warp per token row, the row held in registers; optional fused pre-add (x + r).
// HBM-bound: 2 x rows x C x 4 bytes.
#include "common.cuh"

namespace so {

template <int PER_LANE>   // C <= 32 * PER_LANE
__global__ void __launch_bounds__(256) layer_norm_kernel(const float* __restrict__ x, const float* __restrict__ add,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         float* __restrict__ y, long long rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * C;
  float v[PER_LANE];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    int c = lane + 32 * i;
    float t = c < C ? xr[c] : 0.f;
    if (add && c < C) t += add[row * C + c];
    v[i] = t;
    sum += t;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    int c = lane + 32 * i;
    float d = c < C ? v[i] - mean : 0.f;
    sq = fmaf(d, d, sq);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)C + eps);
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    int c = lane + 32 * i;
    if (c < C) y[row * C + c] = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
  }
}

}  // namespace so

using namespace so;

extern "C" int so_layer_norm(const float* x, const float* add, const float* gamma, const float* beta, float* y, int64_t rows,
                             int32_t C, float eps, void* stream) {
  if (!x || !gamma || !beta || !y || rows < 0 || C < 1) return SO_ERR_INVALID_ARG;
  if (C > 256) return SO_ERR_UNSUPPORTED;
  if (rows == 0) return SO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned grid = (unsigned)ceil_div64(rows, 8);
  if (C <= 96) layer_norm_kernel<3><<<grid, 256, 0, st>>>(x, add, gamma, beta, y, rows, C, eps);
  else if (C <= 128) layer_norm_kernel<4><<<grid, 256, 0, st>>>(x, add, gamma, beta, y, rows, C, eps);
  else layer_norm_kernel<8><<<grid, 256, 0, st>>>(x, add, gamma, beta, y, rows, C, eps);
  note_launch(1);
  return check_launch();
}
