// Device-side DepthMetric step (utils/metric_util.py:247-279,311-349 and the same computation inlined in
// eval_novel_depth.py:174-200): the rendered depth map is sampled at the LiDAR pixel locations with
// F.grid_sample(bilinear, padding_mode='border', align_corners=True) and reduced to the seven error sums per camera,
// without the boolean-mask indexing (a host sync per camera) of the reference.
#include "common.cuh"
#include <math.h>

namespace so {

// depth_pred [N, h, w], loc [N, n, 2] in [0, 1] (x, y) -> sampled [N, n]
__global__ void __launch_bounds__(256) depth_sample_kernel(const float* __restrict__ pred, const float* __restrict__ loc, int N, int n,
                                                           int h, int w, float* __restrict__ out) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)N * n) return;
  int cam = (int)(t / n);
  // metric_util.py:319-324: grid = loc * 2 - 1; ATen unnormalise (align_corners): ((g + 1) / 2) * (size - 1); border clip
  float gx = __fsub_rn(__fmul_rn(loc[2 * t], 2.f), 1.f), gy = __fsub_rn(__fmul_rn(loc[2 * t + 1], 2.f), 1.f);
  float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(w - 1));
  float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), (float)(h - 1));
  ix = fminf((float)(w - 1), fmaxf(ix, 0.f));
  iy = fminf((float)(h - 1), fmaxf(iy, 0.f));
  float x0f = floorf(ix), y0f = floorf(iy);
  int x0 = (int)x0f, y0 = (int)y0f;
  float tx = __fsub_rn(ix, x0f), ty = __fsub_rn(iy, y0f);               // ix - ix_nw
  float ux = __fsub_rn(__fadd_rn(x0f, 1.f), ix), uy = __fsub_rn(__fadd_rn(y0f, 1.f), iy);   // ix_se - ix
  const float* p = pred + (long long)cam * h * w;
  bool xb = x0 + 1 < w, yb = y0 + 1 < h;
  float nw = p[y0 * w + x0], ne = xb ? p[y0 * w + x0 + 1] : 0.f;
  float sw = yb ? p[(y0 + 1) * w + x0] : 0.f, se = (xb && yb) ? p[(y0 + 1) * w + x0 + 1] : 0.f;
  float r = __fmul_rn(nw, __fmul_rn(ux, uy));
  r = __fadd_rn(r, __fmul_rn(ne, __fmul_rn(tx, uy)));
  r = __fadd_rn(r, __fmul_rn(sw, __fmul_rn(ux, ty)));
  r = __fadd_rn(r, __fmul_rn(se, __fmul_rn(tx, ty)));
  out[t] = r;
}

// One CTA per camera (a camera sees a few thousand LiDAR points): deterministic tree reduction of the error sums of
// cal_depth_metric (metric_util.py:247-279) over the masked points.  sums [N, 8] =
// (abs_rel, sq_rel, squared error, squared log error, a1, a2, a3, count); pred' = clamp(scale[cam] * pred, 1e-3, 80).
__global__ void __launch_bounds__(512) depth_metric_kernel(const float* __restrict__ sampled, const float* __restrict__ gt,
                                                           const unsigned char* __restrict__ mask, const float* __restrict__ scale,
                                                           int n, float* __restrict__ sums) {
  const int cam = blockIdx.x;
  const float sc = scale ? scale[cam] : 1.f;
  double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    long long k = (long long)cam * n + i;
    if (!mask[k]) continue;
    float g = gt[k];
    float p = fminf(fmaxf(sc * sampled[k], 1e-3f), 80.f);
    float th = fmaxf(g / p, p / g);
    float d = g - p, l = logf(g) - logf(p);
    a[0] += fabsf(d) / g; a[1] += d * d / g; a[2] += d * d; a[3] += l * l;
    a[4] += th < 1.25f ? 1.0 : 0.0; a[5] += th < 1.25f * 1.25f ? 1.0 : 0.0; a[6] += th < 1.25f * 1.25f * 1.25f ? 1.0 : 0.0;
    a[7] += 1.0;
  }
  __shared__ double sm[8][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    double v = a[j];
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == 0) sm[j][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double v = 0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v += sm[threadIdx.x][wv];
    sums[cam * 8 + threadIdx.x] = (float)v;
  }
}

}  // namespace so

using namespace so;

extern "C" int so_depth_metric_sample(const float* depth_pred, const float* loc, int32_t N, int32_t n, int32_t h, int32_t w,
                                      float* sampled, void* stream) {
  if (!depth_pred || !loc || !sampled || N < 1 || n < 0 || h < 1 || w < 1) return SO_ERR_INVALID_ARG;
  if (n == 0) return SO_OK;
  depth_sample_kernel<<<(unsigned)ceil_div64((int64_t)N * n, 256), 256, 0, (cudaStream_t)stream>>>(depth_pred, loc, N, n, h, w, sampled);
  note_launch(1);
  return check_launch();
}

extern "C" int so_depth_metric_sums(const float* sampled, const float* depth_gt, const uint8_t* mask, const float* scale, int32_t N,
                                    int32_t n, float* sums, void* stream) {
  if (!sampled || !depth_gt || !mask || !sums || N < 1 || n < 0) return SO_ERR_INVALID_ARG;
  depth_metric_kernel<<<N, 512, 0, (cudaStream_t)stream>>>(sampled, depth_gt, mask, scale, n, sums);
  note_launch(1);
  return check_launch();
}
