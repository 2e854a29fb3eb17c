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

// ---- A3: FPN level -> token rows (tpvformer_encoder.py:261-277) ------------------------------------------------------
// feat [N, C, hw] of one level -> out[n, level_start + p, :] = (feat[n, :, p] + cams_embeds[n, :]) + level_embed[:]
// (the reference's two adds in its order), i.e. flatten(3).permute + both embeddings + the concat over levels in ONE pass:
// a 32 x 32 shared-memory tile transpose, coalesced on both sides.  HBM-bound: 2 x N x hw x C x 4 bytes per level.
namespace so {
__global__ void __launch_bounds__(256) flatten_level_kernel(const float* __restrict__ feat, const float* __restrict__ cams,
                                                            const float* __restrict__ lvl, float* __restrict__ out, int C, int hw,
                                                            long long level_start, long long total) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8 threads
  const float* src = feat + (long long)n * C * hw;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = c0 + ty + 8 * j, p = p0 + tx;
    if (c < C && p < hw) tile[ty + 8 * j][tx] = src[(long long)c * hw + p];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int p = p0 + ty + 8 * j, c = c0 + tx;
    if (c < C && p < hw)
      out[((long long)n * total + level_start + p) * C + c] = __fadd_rn(__fadd_rn(tile[tx][ty + 8 * j], __ldg(cams + n * C + c)), __ldg(lvl + c));
  }
}
}  // namespace so

extern "C" int so_flatten_level(const float* feat, const float* cams_embeds, const float* level_embed, float* out, int32_t N,
                                int32_t C, int32_t hw, int64_t level_start, int64_t total, void* stream) {
  if (!feat || !cams_embeds || !level_embed || !out || N < 1 || C < 1 || hw < 1 || level_start < 0 || level_start + hw > total)
    return SO_ERR_INVALID_ARG;
  if (N > 65535 || (C + 31) / 32 > 65535) return SO_ERR_UNSUPPORTED;
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)N);
  so::flatten_level_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(feat, cams_embeds, level_embed, out, C, hw, level_start, total);
  so::note_launch(1);
  return so::check_launch();
}
