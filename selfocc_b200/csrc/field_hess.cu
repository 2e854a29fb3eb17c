// `second_grad` training output (neus_head.py:177,703-706; consumer loss/second_grad_loss.py:19-20: |second_grad|.mean()).
// DECLARED ASSUMPTION -- the quantity is defined inside the un-vendored sdfstudio fork (cuda_gridsample_grad2); what is
// restated here is the standard double-backward idiom  second_grad = d( sum_j d sdf / d x_j ) / d x  of the trilinear field,
// i.e. the ROW SUMS of the Hessian of the interpolant in metres.  Inside a cell the pure second derivatives vanish and
//     d2 s / dh dw = sum_k wz_k (a11k - a10k - a01k + a00k)          (and cyclic),
// so  second_grad = (Hxy + Hxz, Hxy + Hyz, Hxz + Hyz),  Hxy = kw kh d2s/dw dh, ...  (x <-> w, y <-> h, z <-> d).
// Zero padding outside the volume like F.grid_sample(padding_mode='zeros', align_corners=True).  Opt-in (return_second_grad).
#include "common.cuh"

namespace so {

struct Corner8 { float a[2][2][2]; };

__device__ __forceinline__ void hess_weights(const Taps& t, float kh, float kw, float kd, float ghw, float ghz, float gwz,
                                             float c[2][2][2]) {
  // contribution of corner (i, j, k) to  ghw * s_hw + ghz * s_hz + gwz * s_wz  (signs: +1 for index 1, -1 for index 0)
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        float si = i ? 1.f : -1.f, sj = j ? 1.f : -1.f, sk = k ? 1.f : -1.f;
        float wh = i ? t.fh : 1.f - t.fh, ww = j ? t.fw : 1.f - t.fw, wz = k ? t.fz : 1.f - t.fz;
        float m = (i ? t.mh1 : t.mh0) * (j ? t.mw1 : t.mw0) * (k ? t.mz1 : t.mz0);
        c[i][j][k] = m * (ghw * si * sj * wz + ghz * si * sk * ww + gwz * sj * sk * wh);
      }
}

__device__ __forceinline__ size_t corner_index(const VolumeDev& v, const Taps& t, int i, int j, int k) {
  int h = min(max(t.h0 + i, 0), v.H - 1), w = min(max(t.w0 + j, 0), v.W - 1), z = min(max(t.z0 + k, 0), v.Z - 1);
  return ((size_t)h * v.W + w) * v.zpitch + z;
}

__global__ void __launch_bounds__(256) field_hess_kernel(VolumeDev V, const float* __restrict__ pts, long long n,
                                                         float* __restrict__ out) {
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  float x = pts[3 * p], y = pts[3 * p + 1], z = pts[3 * p + 2];
  float kh, kw, kd;
  float gh = axis_m2g(V.ax[0], y, kh), gw = axis_m2g(V.ax[1], x, kw), gd = axis_m2g(V.ax[2], z, kd);
  Taps t = make_taps(V, gh, gw, gd);
  float s_hw = 0.f, s_hz = 0.f, s_wz = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        float m = (i ? t.mh1 : t.mh0) * (j ? t.mw1 : t.mw0) * (k ? t.mz1 : t.mz0);
        float a = m * __ldg(V.sdf + corner_index(V, t, i, j, k));
        float si = i ? 1.f : -1.f, sj = j ? 1.f : -1.f, sk = k ? 1.f : -1.f;
        s_hw = fmaf(si * sj * (k ? t.fz : 1.f - t.fz), a, s_hw);
        s_hz = fmaf(si * sk * (j ? t.fw : 1.f - t.fw), a, s_hz);
        s_wz = fmaf(sj * sk * (i ? t.fh : 1.f - t.fh), a, s_wz);
      }
  float Hxy = kw * kh * s_hw, Hxz = kw * kd * s_wz, Hyz = kh * kd * s_hz;
  out[3 * p] = Hxy + Hxz;
  out[3 * p + 1] = Hxy + Hyz;
  out[3 * p + 2] = Hxz + Hyz;
}

__global__ void __launch_bounds__(256) field_hess_bwd_kernel(VolumeDev V, const float* __restrict__ pts, long long n,
                                                             const float* __restrict__ g, float* __restrict__ gvs) {
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  float x = pts[3 * p], y = pts[3 * p + 1], z = pts[3 * p + 2];
  float kh, kw, kd;
  float gh = axis_m2g(V.ax[0], y, kh), gw = axis_m2g(V.ax[1], x, kw), gd = axis_m2g(V.ax[2], z, kd);
  Taps t = make_taps(V, gh, gw, gd);
  float g0 = g[3 * p], g1 = g[3 * p + 1], g2 = g[3 * p + 2];
  float c[2][2][2];
  hess_weights(t, kh, kw, kd, (g0 + g1) * kw * kh, (g1 + g2) * kh * kd, (g0 + g2) * kw * kd, c);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (c[i][j][k] != 0.f) atomicAdd(gvs + corner_index(V, t, i, j, k), c[i][j][k]);
}

}  // namespace so

using namespace so;

extern "C" int so_field_second_grad(const float* vol_sdf, const so_volume_desc* vol_host, const float* points, int64_t n,
                                    float* second_grad, void* stream) {
  if (!vol_sdf || !points || !second_grad || n < 0) return SO_ERR_INVALID_ARG;
  int rc = validate_volume(vol_host);
  if (rc) return rc;
  if (n == 0) return SO_OK;
  VolumeDev V = make_volume(*vol_host, vol_sdf, nullptr);
  field_hess_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(V, points, n, second_grad);
  note_launch(1);
  return check_launch();
}

extern "C" int so_field_second_grad_backward(const so_volume_desc* vol_host, const float* points, int64_t n, const float* g_second_grad,
                                             float* g_vol_sdf, void* stream) {
  if (!points || !g_second_grad || !g_vol_sdf || n < 0) return SO_ERR_INVALID_ARG;
  int rc = validate_volume(vol_host);
  if (rc) return rc;
  if (n == 0) return SO_OK;
  VolumeDev V = make_volume(*vol_host, nullptr, nullptr);
  field_hess_bwd_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(V, points, n, g_second_grad, g_vol_sdf);
  note_launch(1);
  return check_launch();
}
