// Training-form render: NeuSHead.forward must emit per-sample tensors
// (weights / ts / deltas / eik_grad [/ sample_sdf], neus_head.py:667-682) and be differentiable w.r.t. the
// decoded volume and the NeuS deviation parameter.
//
// Mapping: one WARP per ray, lane = sample (s = 32*k + lane, k = 0..S/32-1), so every per-sample tensor is written
// fully coalesced (32 consecutive floats per store) -- this form is HBM-write bound (~6 KB/ray).  The
// transmittance is an exclusive product scan along the sample dimension done with warp shuffles (5 steps per
// 32-sample chunk + a carried prefix); the backward needs the matching suffix sums and runs the chunks in reverse.
#include "render_common.cuh"
#ifndef SO_TRAIN_FWD_MIN_CTAS
#define SO_TRAIN_FWD_MIN_CTAS 8   // 64 registers: 0.122 ms vs 0.153 ms at 4 CTAs/SM (cfg-5 sizes); the kernel is latency-bound
#endif

namespace so {

constexpr int kTrainMaxChunks = 8;  // S <= 256

struct TrainOut {
  float *depth, *acc, *fars, *rgb, *sem, *max_depth;      // per ray
  float *weights, *ts, *deltas, *eik, *sdf;               // per sample
};

struct TrainGrad {
  const float *g_depth, *g_acc, *g_rgb, *g_sem, *g_weights, *g_eik, *g_sdf;
  float *g_vol_sdf, *g_vol_feat, *g_inv_s;
};

// one sample of one ray: geometry + field + alpha.  `s` must be < S.
struct Sample {
  float mid, delta, sdf, gx, gy, gz, alpha;
  // pieces the backward needs
  float half, pa, pb, tc;
  Taps t;
  float kh, kw, kd;
};

// Per-ray constants of the training kernels (hoisted out of the per-sample work)
struct RayCtx {
  float o[3], d[3], nrm, inv_nrm, tn, tf;
  float gh0, gdh, gw0, gdw, gd0, gdd;   // affine grid-space ray (valid when the mapping has no outer ring)
  bool affine;
  const float* u;                       // this ray's jitter row or nullptr
};

__device__ __forceinline__ void make_ctx(const VolumeDev& V, const RayDev& R, const RenderDev& P, long long gid, RayCtx& c) {
  make_ray(R, gid, c.o, c.d, c.nrm);
  slab(P, c.o, c.d, c.tn, c.tf);
  c.inv_nrm = 1.0f / c.nrm;
  c.affine = V.ax[0].k1 == 0.f && V.ax[1].k1 == 0.f && V.ax[2].k1 == 0.f;
  c.gh0 = fmaf(c.o[1] - V.ax[0].start, V.ax[0].k0, V.ax[0].offset); c.gdh = c.d[1] * V.ax[0].k0;
  c.gw0 = fmaf(c.o[0] - V.ax[1].start, V.ax[1].k0, V.ax[1].offset); c.gdw = c.d[0] * V.ax[1].k0;
  c.gd0 = fmaf(c.o[2] - V.ax[2].start, V.ax[2].k0, V.ax[2].offset); c.gdd = c.d[2] * V.ax[2].k0;
  c.u = P.jitter ? P.jitter + gid * (long long)(P.S + 1) : nullptr;
}

// one sample of one ray (lane = sample): geometry + field + alpha.  `s` must be < S.  Edges are shared between
// neighbouring lanes with one shuffle (lane 31 computes its own right edge).
__device__ __forceinline__ void eval_sample(const VolumeDev& V, const RenderDev& P, const RayCtx& c, int s, int lane, Sample& q) {
  const int S = P.S;
  const float step = 1.0f / (float)S;
  float e0 = edge_t(bin_edge01_jit(s, S, step, c.u), c.tn, c.tf);
  float e1 = __shfl_down_sync(0xffffffffu, e0, 1);
  if (lane == 31 || s + 1 >= S) e1 = edge_t(bin_edge01_jit(min(s + 1, S), S, step, c.u), c.tn, c.tf);
  q.mid = __fmul_rn(__fadd_rn(e0, e1), 0.5f);
  q.delta = __fsub_rn(e1, e0);
  float tq = P.anchor_mid ? q.mid : e0;
  float gh, gw, gd;
  if (c.affine) {
    gh = fmaf(c.gdh, tq, c.gh0); gw = fmaf(c.gdw, tq, c.gw0); gd = fmaf(c.gdd, tq, c.gd0);
    q.kh = V.ax[0].k0; q.kw = V.ax[1].k0; q.kd = V.ax[2].k0;
  } else {
    float x = fmaf(c.d[0], tq, c.o[0]), y = fmaf(c.d[1], tq, c.o[1]), z = fmaf(c.d[2], tq, c.o[2]);
    gh = axis_m2g(V.ax[0], y, q.kh); gw = axis_m2g(V.ax[1], x, q.kw); gd = axis_m2g(V.ax[2], z, q.kd);
  }
  q.t = make_taps(V, gh, gw, gd);
  float dgh, dgw, dgd;
  const bool interior = (unsigned)q.t.h0 < (unsigned)(V.H - 1) && (unsigned)q.t.w0 < (unsigned)(V.W - 1) &&
                        (unsigned)q.t.z0 < (unsigned)(V.Z - 1);
  if (__all_sync(0xffffffffu, interior)) gather_sdf_interior(V, q.t.h0, q.t.w0, q.t.z0, q.t.fh, q.t.fw, q.t.fz, q.sdf, dgh, dgw, dgd);
  else gather_sdf(V, q.t, q.sdf, dgh, dgw, dgd);
  q.gx = dgw * q.kw; q.gy = dgh * q.kh; q.gz = dgd * q.kd;
  q.tc = c.d[0] * q.gx + c.d[1] * q.gy + c.d[2] * q.gz;
  float ic = -(fmaxf(fmaf(-q.tc, 0.5f, 0.5f), 0.f) * (1.0f - P.cos_anneal) + fmaxf(-q.tc, 0.f) * P.cos_anneal);
  q.half = ic * q.delta * 0.5f;
  float a = (q.sdf - q.half) * P.inv_s, b = (q.sdf + q.half) * P.inv_s;
  q.pa = sigmoid_fast(a);
  q.pb = sigmoid_fast(b);
  float diff = q.pa * sigmoid_fast(-b) * one_minus_exp_neg(-2.0f * q.half * P.inv_s);
  q.alpha = __saturatef(__fdividef(diff + 1e-5f, q.pa + 1e-5f));
}

__device__ __forceinline__ float warp_incl_prod(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= n;
  }
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// reverse inclusive sum: out[lane] = sum_{i >= lane} v[i]
__device__ __forceinline__ float warp_rev_incl_sum(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += n;
  }
  return v;
}

__device__ __forceinline__ void sample_colour(const VolumeDev& V, const RenderDev& P, const Taps& t, float col[3], float raw[3]) {
  float f[3];
  gather_feat<3>(V, t, 0, f);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    raw[c] = f[c] * kC0;
    col[c] = P.sh_act == 0 ? fmaxf(raw[c] + 0.5f, 0.f) : sigmoidf_acc(raw[c]);
  }
}

// FAST = affine metre->grid map, S a power of two (multiple of 32), cos-anneal finished, mid-point anchor: lean per-sample
// path (closed-form jittered edges, interior gather chosen by a warp vote, base-2 alpha), ~2.4x fewer instructions.
template <bool HAS_RGB, bool HAS_SEM, bool FAST>
__global__ void __launch_bounds__(128, SO_TRAIN_FWD_MIN_CTAS) render_train_fwd_kernel(VolumeDev V, RayDev R, RenderDev P, const float* __restrict__ ws,
                                                               const float* __restrict__ bkgd_rand, TrainOut O) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  const int S = P.S;
  const int K = (S + 31) >> 5;
  const int n_sem = HAS_SEM ? V.n_feat - 3 : 0;
  const float eps = 1.1920928955078125e-07f;
  for (long long ray = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); ray < R.ray_count; ray += warps) {
    long long gid = R.ray_begin + ray;
    RayCtx c;
    make_ctx(V, R, P, gid, c);
    const float nrm = c.nrm, tf = c.tf;
    float carry = 1.0f, acc = 0.f, dsum = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
    float best = -INFINITY, best_ts = 0.f;
    int best_i = 0x7fffffff;
    float sem_acc[HAS_SEM ? kMaxSem : 1];
    if (HAS_SEM)
      for (int c = 0; c < kMaxSem; ++c) sem_acc[c] = 0.f;
    const float step = 1.0f / (float)S, span = c.tf - c.tn, k_log2 = P.inv_s * 1.4426950408889634f;
    // the jitter value of the NEXT chunk is requested one iteration ahead, so its DRAM latency is off the critical path
    float u_cur = (FAST && c.u) ? __ldg(c.u + lane) : 0.f;
    for (int k = 0; k < K; ++k) {
      int s = k * 32 + lane;
      bool live = s < S;
      Sample q;
      if (FAST) {
        const int s_nxt = s + 32;
        const float u_nxt = (c.u && s_nxt <= S) ? __ldg(c.u + s_nxt) : 0.f;
        // edge i is re-drawn inside [max(i - 1/2, 0), min(i + 1/2, S)] / S (exact arithmetic for power-of-two S)
        auto edge = [&](int i, float u) {
          float fi = (float)i;
          float lo_b = fmaxf(fi - 0.5f, 0.f) * step, up_b = fminf(fi + 0.5f, (float)S) * step;
          float b = c.u ? fmaf(up_b - lo_b, u, lo_b) : fi * step;
          return fmaf(b, span, c.tn);
        };
        float e0 = edge(s, u_cur);
        float e_next = edge(k * 32 + 32, __shfl_sync(0xffffffffu, u_nxt, 0));   // right edge of lane 31 = first edge of the next chunk
        float e1 = __shfl_down_sync(0xffffffffu, e0, 1);
        if (lane == 31) e1 = e_next;
        u_cur = u_nxt;
        q.mid = 0.5f * (e0 + e1);
        q.delta = e1 - e0;
        float gh = fmaf(c.gdh, q.mid, c.gh0), gw = fmaf(c.gdw, q.mid, c.gw0), gd = fmaf(c.gdd, q.mid, c.gd0);
        q.kh = V.ax[0].k0; q.kw = V.ax[1].k0; q.kd = V.ax[2].k0;
        float flh = floorf(gh), flw = floorf(gw), flz = floorf(gd);
        int h0 = (int)flh, w0 = (int)flw, z0 = (int)flz;
        bool interior = (unsigned)h0 < (unsigned)(V.H - 1) && (unsigned)w0 < (unsigned)(V.W - 1) && (unsigned)z0 < (unsigned)(V.Z - 1);
        float dgh, dgw, dgd;
        if (__all_sync(0xffffffffu, interior)) {
          gather_sdf_interior(V, h0, w0, z0, gh - flh, gw - flw, gd - flz, q.sdf, dgh, dgw, dgd);
          if (HAS_RGB) q.t = make_taps(V, gh, gw, gd);
        } else {
          q.t = make_taps(V, gh, gw, gd);
          gather_sdf(V, q.t, q.sdf, dgh, dgw, dgd);
        }
        q.gx = dgw * q.kw; q.gy = dgh * q.kh; q.gz = dgd * q.kd;
        float tc = c.d[0] * q.gx + c.d[1] * q.gy + c.d[2] * q.gz;
        q.alpha = neus_alpha_log2(q.sdf * k_log2, fminf(tc, 0.f) * (q.delta * (0.5f * k_log2)));
      } else {
        eval_sample(V, P, c, live ? s : S - 1, lane, q);
      }
      float alpha = live ? q.alpha : 0.f;
      float f = live ? (1.0f - alpha + 1e-7f) : 1.0f;
      float incl = warp_incl_prod(f, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      float T = carry * (lane == 0 ? 1.0f : excl);
      carry *= __shfl_sync(0xffffffffu, incl, 31);
      float w = alpha * T;
      if (live) {
        long long oidx = ray * S + s;
        float ts = q.mid * c.inv_nrm, dl = q.delta * c.inv_nrm;     // neus_head.py:571-577 (reciprocal multiply, <= 1 ulp)
        if (O.weights) O.weights[oidx] = w;
        if (O.ts) O.ts[oidx] = ts;
        if (O.deltas) O.deltas[oidx] = dl;
        if (O.sdf) O.sdf[oidx] = q.sdf;
        if (O.eik) { O.eik[3 * oidx] = q.gx; O.eik[3 * oidx + 1] = q.gy; O.eik[3 * oidx + 2] = q.gz; }
        acc += w;
        dsum = fmaf(w, q.mid, dsum);
        float cand = (dl < eps ? 0.f : w) * __fdividef(1.0f, fmaxf(dl, eps));  // neus_head.py:579-587
        if (cand > best) { best = cand; best_i = s; best_ts = ts; }
        if (HAS_RGB) {
          float col[3], raw[3];
          sample_colour(V, P, q.t, col, raw);
          cr = fmaf(w, col[0], cr); cg = fmaf(w, col[1], cg); cb = fmaf(w, col[2], cb);
        }
        if (HAS_SEM) {
          float lg[kMaxSem];
          float mx = -INFINITY;
          for (int c = 0; c < n_sem; ++c) { float f1[1]; gather_feat<1>(V, q.t, 3 + c, f1); lg[c] = f1[0]; mx = fmaxf(mx, f1[0]); }
          float den = 0.f;
          for (int c = 0; c < n_sem; ++c) { lg[c] = expf(lg[c] - mx); den += lg[c]; }
          float sc = w / den;
          for (int c = 0; c < n_sem; ++c) sem_acc[c] = fmaf(sc, lg[c], sem_acc[c]);
        }
      }
    }
    acc = warp_sum(acc);
    dsum = warp_sum(dsum);
    // first-max argmax across lanes: larger score wins, ties go to the smaller sample index
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, off);
      int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
      float ot = __shfl_xor_sync(0xffffffffu, best_ts, off);
      if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; best_ts = ot; }
    }
    if (HAS_RGB) { cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb); }
    if (HAS_SEM)
      for (int c = 0; c < n_sem; ++c) sem_acc[c] = warp_sum(sem_acc[c]);
    if (lane == 0) {
      long long chunk = R.chunk_len > 0 ? gid / R.chunk_len : 0;
      float lo = ws[2 * chunk], hi = ws[2 * chunk + 1];
      float dd = fminf(fmaxf(dsum / (acc + 1e-10f), lo), hi);
      if (O.depth) O.depth[ray] = dd / nrm;
      if (O.acc) O.acc[ray] = acc;
      if (O.fars) O.fars[ray] = tf / nrm;
      if (O.max_depth) O.max_depth[ray] = best_ts;
      if (HAS_RGB && O.rgb) {
        float b0, b1, b2;
        if (P.bkgd_mode == 2) { b0 = bkgd_rand[3 * ray]; b1 = bkgd_rand[3 * ray + 1]; b2 = bkgd_rand[3 * ray + 2]; }
        else { b0 = b1 = b2 = (P.bkgd_mode == 1) ? 1.f : 0.f; }
        float rem = 1.0f - acc;
        float r = fmaf(b0, rem, cr), g = fmaf(b1, rem, cg), b = fmaf(b2, rem, cb);
        if (P.eval_clamp) { r = __saturatef(r); g = __saturatef(g); b = __saturatef(b); }
        O.rgb[3 * ray] = r; O.rgb[3 * ray + 1] = g; O.rgb[3 * ray + 2] = b;
      }
      if (HAS_SEM && O.sem)
        for (int c = 0; c < n_sem; ++c) O.sem[ray * n_sem + c] = sem_acc[c];
    }
  }
}

// scatter d(loss)/d(sdf value), d/d(metre-gradient) of one sample into the 8 corners of the sdf volume
__device__ __forceinline__ void scatter_sdf(const VolumeDev& V, float* __restrict__ gvol, const Taps& t, float g_s, float g_gh,
                                            float g_gw, float g_gd) {
  // interpolant:  s = sum_c Wh(c) Ww(c) Wz(c) v_c ;  d s / d gh = sum_c Wh'(c) Ww Wz v_c  etc. (masks = zero padding)
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int dh = k >> 2, dw = (k >> 1) & 1, dz = k & 1;
    float mh = dh ? t.mh1 : t.mh0, mw = dw ? t.mw1 : t.mw0, mz = dz ? t.mz1 : t.mz0;
    float m = mh * mw * mz;
    if (m == 0.f) continue;
    float wh = dh ? t.fh : 1.f - t.fh, ww = dw ? t.fw : 1.f - t.fw, wz = dz ? t.fz : 1.f - t.fz;
    float sh = dh ? 1.f : -1.f, sw = dw ? 1.f : -1.f, sz = dz ? 1.f : -1.f;
    float g = g_s * wh * ww * wz + g_gh * sh * ww * wz + g_gw * wh * sw * wz + g_gd * wh * ww * sz;
    atomicAdd(gvol + ((size_t)(t.h0 + dh) * V.W + (t.w0 + dw)) * V.zpitch + (t.z0 + dz), g);
  }
}

template <int N>
__device__ __forceinline__ void scatter_feat(const VolumeDev& V, float* __restrict__ gfeat, const Taps& t, int c0, const float g[N]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int dh = k >> 2, dw = (k >> 1) & 1, dz = k & 1;
    float m = (dh ? t.mh1 : t.mh0) * (dw ? t.mw1 : t.mw0) * (dz ? t.mz1 : t.mz0);
    if (m == 0.f) continue;
    float wgt = (dh ? t.fh : 1.f - t.fh) * (dw ? t.fw : 1.f - t.fw) * (dz ? t.fz : 1.f - t.fz);
    float* p = gfeat + (((size_t)(t.h0 + dh) * V.W + (t.w0 + dw)) * V.Z + (t.z0 + dz)) * V.feat_pitch + c0;
#pragma unroll
    for (int i = 0; i < N; ++i) atomicAdd(p + i, wgt * g[i]);
  }
}

template <bool HAS_RGB, bool HAS_SEM>
__global__ void __launch_bounds__(128) render_train_bwd_kernel(VolumeDev V, RayDev R, RenderDev P, const float* __restrict__ ws,
                                                               const float* __restrict__ bkgd_rand, TrainGrad G) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  const int S = P.S;
  const int K = (S + 31) >> 5;
  const int n_sem = HAS_SEM ? V.n_feat - 3 : 0;
  float g_invs_local = 0.f;
  for (long long ray = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); ray < R.ray_count; ray += warps) {
    long long gid = R.ray_begin + ray;
    RayCtx c;
    make_ctx(V, R, P, gid, c);
    const float nrm = c.nrm;
    const float* d = c.d;
    // ---- pass 1: transmittance per sample, ray sums
    float Tk[kTrainMaxChunks], Ak[kTrainMaxChunks];
    float carry = 1.0f, acc = 0.f, dsum = 0.f;
#pragma unroll
    for (int k = 0; k < kTrainMaxChunks; ++k) {
      if (k >= K) break;
      int s = k * 32 + lane;
      bool live = s < S;
      Sample q;
      eval_sample(V, P, c, live ? s : S - 1, lane, q);
      float alpha = live ? q.alpha : 0.f;
      float f = live ? (1.0f - alpha + 1e-7f) : 1.0f;
      float incl = warp_incl_prod(f, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      float T = carry * (lane == 0 ? 1.0f : excl);
      carry *= __shfl_sync(0xffffffffu, incl, 31);
      Tk[k] = T; Ak[k] = alpha;
      acc += alpha * T;
      dsum = fmaf(alpha * T, q.mid, dsum);
    }
    acc = warp_sum(acc);
    dsum = warp_sum(dsum);
    long long chunk = R.chunk_len > 0 ? gid / R.chunk_len : 0;
    float lo = ws[2 * chunk], hi = ws[2 * chunk + 1];
    float draw = dsum / (acc + 1e-10f);
    bool clipped = draw < lo || draw > hi;
    float gd = (G.g_depth && !clipped) ? G.g_depth[ray] / nrm : 0.f;   // d depth / d depth_raw (depth = clip(raw)/|dir|)
    float ga = G.g_acc ? G.g_acc[ray] : 0.f;
    float gr[3] = {0.f, 0.f, 0.f}, bg[3] = {0.f, 0.f, 0.f};
    if (HAS_RGB && G.g_rgb) {
      gr[0] = G.g_rgb[3 * ray]; gr[1] = G.g_rgb[3 * ray + 1]; gr[2] = G.g_rgb[3 * ray + 2];
      if (P.bkgd_mode == 2) { bg[0] = bkgd_rand[3 * ray]; bg[1] = bkgd_rand[3 * ray + 1]; bg[2] = bkgd_rand[3 * ray + 2]; }
      else bg[0] = bg[1] = bg[2] = (P.bkgd_mode == 1) ? 1.f : 0.f;
    }
    // ---- pass 2: reverse over chunks
    float tail = 0.f;   // sum_{j in later chunks} G_j w_j
#pragma unroll
    for (int kk = kTrainMaxChunks - 1; kk >= 0; --kk) {
      if (kk >= K) continue;
      int s = kk * 32 + lane;
      bool live = s < S;
      Sample q;
      eval_sample(V, P, c, live ? s : S - 1, lane, q);
      float T = Tk[kk], alpha = Ak[kk];
      float w = alpha * T;
      long long oidx = ray * S + (live ? s : S - 1);
      // dL/dw_s
      float Gw = (G.g_weights ? G.g_weights[oidx] : 0.f) + ga + gd * (q.mid - draw) / (acc + 1e-10f);
      float col[3] = {0.f, 0.f, 0.f}, raw[3] = {0.f, 0.f, 0.f};
      float lg[HAS_SEM ? kMaxSem : 1];
      float gdot = 0.f;
      if (HAS_RGB && G.g_rgb) {
        sample_colour(V, P, q.t, col, raw);
        Gw += gr[0] * (col[0] - bg[0]) + gr[1] * (col[1] - bg[1]) + gr[2] * (col[2] - bg[2]);
      }
      if (HAS_SEM && G.g_sem) {
        float mx = -INFINITY;
        for (int c = 0; c < n_sem; ++c) { float f1[1]; gather_feat<1>(V, q.t, 3 + c, f1); lg[c] = f1[0]; mx = fmaxf(mx, f1[0]); }
        float den = 0.f;
        for (int c = 0; c < n_sem; ++c) { lg[c] = expf(lg[c] - mx); den += lg[c]; }
        for (int c = 0; c < n_sem; ++c) { lg[c] /= den; gdot += G.g_sem[ray * n_sem + c] * lg[c]; }
        Gw += gdot;
      }
      if (!live) Gw = 0.f;
      float gw_w = live ? Gw * w : 0.f;
      float rinc = warp_rev_incl_sum(gw_w, lane);
      float B = tail + rinc - gw_w;                         // sum_{j > s} G_j w_j
      tail += __shfl_sync(0xffffffffu, rinc, 0);
      float dalpha = Gw * T - B / (1.0f - alpha + 1e-7f);
      if (!live) dalpha = 0.f;   // (the raw alpha lies in (0, 1] by construction, so the clip never cuts a gradient)
      // alpha = (Pa - Pb + e) / (Pa + e):  d/da = Pa' Pb / (Pa+e)^2,  d/db = -Pb' / (Pa+e)
      float den = q.pa + 1e-5f;
      float da = q.pa * (1.0f - q.pa) * q.pb / (den * den);
      float db = -q.pb * (1.0f - q.pb) / den;
      float g_sdf = dalpha * (da + db) * P.inv_s;
      float g_half = dalpha * (db - da) * P.inv_s;
      g_invs_local += dalpha * (da * (q.sdf - q.half) + db * (q.sdf + q.half));
      // half = ic * delta / 2,  ic = -(relu(-tc/2 + 1/2)(1-r) + relu(-tc) r)
      float dic = 0.f;
      if (fmaf(-q.tc, 0.5f, 0.5f) > 0.f) dic += 0.5f * (1.0f - P.cos_anneal);
      if (-q.tc > 0.f) dic += P.cos_anneal;
      float g_tc = g_half * 0.5f * q.delta * dic;
      float ggx = g_tc * d[0], ggy = g_tc * d[1], ggz = g_tc * d[2];
      if (live) {
        if (G.g_sdf) g_sdf += G.g_sdf[oidx];
        if (G.g_eik) { ggx += G.g_eik[3 * oidx]; ggy += G.g_eik[3 * oidx + 1]; ggz += G.g_eik[3 * oidx + 2]; }
        // metre gradient (gx, gy, gz) = (dgw kw, dgh kh, dgd kd)
        scatter_sdf(V, G.g_vol_sdf, q.t, g_sdf, ggy * q.kh, ggx * q.kw, ggz * q.kd);
        if (HAS_RGB && G.g_rgb && G.g_vol_feat) {
          float gf[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float dact = P.sh_act == 0 ? (raw[c] + 0.5f > 0.f ? 1.f : 0.f) : col[c] * (1.f - col[c]);
            gf[c] = w * gr[c] * dact * kC0;
          }
          scatter_feat<3>(V, G.g_vol_feat, q.t, 0, gf);
        }
        if (HAS_SEM && G.g_sem && G.g_vol_feat) {
          for (int c = 0; c < n_sem; ++c) {
            float gl[1] = {w * lg[c] * (G.g_sem[ray * n_sem + c] - gdot)};
            scatter_feat<1>(V, G.g_vol_feat, q.t, 3 + c, gl);
          }
        }
      }
    }
  }
  if (G.g_inv_s) {
    g_invs_local = warp_sum(g_invs_local);
    if (lane == 0 && g_invs_local != 0.f) atomicAdd(G.g_inv_s, g_invs_local);
  }
}

__global__ void __launch_bounds__(256) field_query_bwd_kernel(VolumeDev V, const float* __restrict__ pts, long long n,
                                                              const float* __restrict__ g_sdf, const float* __restrict__ g_grad,
                                                              const float* __restrict__ g_feat, float* __restrict__ gvs,
                                                              float* __restrict__ gvf) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
  float kh, kw, kd;
  float gh = axis_m2g(V.ax[0], y, kh), gw = axis_m2g(V.ax[1], x, kw), gd = axis_m2g(V.ax[2], z, kd);
  Taps t = make_taps(V, gh, gw, gd);
  float gs = g_sdf ? g_sdf[i] : 0.f;
  float gx = g_grad ? g_grad[3 * i] : 0.f, gy = g_grad ? g_grad[3 * i + 1] : 0.f, gz = g_grad ? g_grad[3 * i + 2] : 0.f;
  scatter_sdf(V, gvs, t, gs, gy * kh, gx * kw, gz * kd);
  if (g_feat && gvf)
    for (int c = 0; c < V.n_feat; ++c) { float g1[1] = {g_feat[i * V.n_feat + c]}; scatter_feat<1>(V, gvf, t, c, g1); }
}

static int train_common_checks(const float* vol_sdf, const float* vol_feat, const so_volume_desc* d, const float* cam_mats,
                               const so_ray_desc* rd, const so_render_params* pr, const float* workspace, bool want_rgb,
                               bool want_sem, const float* bkgd_rand) {
  if (!vol_sdf || !cam_mats || !rd || !pr || !workspace) return SO_ERR_INVALID_ARG;
  int rc = validate_volume(d);
  if (rc) return rc;
  if (pr->num_samples < 1) return SO_ERR_INVALID_ARG;
  if (pr->num_samples > 32 * kTrainMaxChunks) return SO_ERR_UNSUPPORTED;
  if (want_rgb && (d->n_feat < 3 || !vol_feat)) return SO_ERR_INVALID_ARG;
  if (want_sem && (d->n_feat <= 3 || !vol_feat)) return SO_ERR_INVALID_ARG;
  if (want_sem && d->n_feat - 3 > kMaxSem) return SO_ERR_UNSUPPORTED;
  if (pr->bkgd_mode == 2 && want_rgb && !bkgd_rand) return SO_ERR_INVALID_ARG;
  if (pr->bkgd_mode < 0 || pr->bkgd_mode > 2 || pr->sh_act < 0 || pr->sh_act > 1) return SO_ERR_INVALID_ARG;
  return SO_OK;
}

}  // namespace so

using namespace so;

extern "C" int so_render_train_forward(const float* vol_sdf, const float* vol_feat, const so_volume_desc* vol_host,
                                       const float* cam_mats, const float* pix, const so_ray_desc* rd,
                                       const so_render_params* pr, const float* jitter, const float* bkgd_rand,
                                       float* depth, float* acc, float* fars, float* rgb, float* sem, float* max_depth,
                                       float* weights, float* ts, float* deltas, float* eik_grad, float* sample_sdf,
                                       float* workspace, void* stream) {
  bool want_rgb = rgb != nullptr, want_sem = sem != nullptr;
  int rc = train_common_checks(vol_sdf, vol_feat, vol_host, cam_mats, rd, pr, workspace, want_rgb, want_sem, bkgd_rand);
  if (rc) return rc;
  RayDev R;
  if ((rc = make_ray_dev(rd, cam_mats, pix, &R))) return rc;
  if (R.ray_count == 0) return SO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  VolumeDev V = make_volume(*vol_host, vol_sdf, vol_feat);
  RenderDev P = make_render_dev(*pr, jitter);
  if ((rc = launch_depth_bounds(R, P, workspace, st))) return rc;
  TrainOut O{depth, acc, fars, rgb, sem, max_depth, weights, ts, deltas, eik_grad, sample_sdf};
  // persistent-style grid: warps stride over rays; 4 warps per CTA, enough CTAs to fill 148 SMs several times over
  long long ctas = ceil_div64(R.ray_count, 4);
  unsigned grid = (unsigned)(ctas < (long long)kNumSMs * 16 ? ctas : (long long)kNumSMs * 16);
  ProfScope prof(6, st);
  const bool fast = V.ax[0].k1 == 0.f && V.ax[1].k1 == 0.f && V.ax[2].k1 == 0.f && (P.S & (P.S - 1)) == 0 && P.S >= 32 &&
                    P.cos_anneal == 1.0f && P.anchor_mid;
#define SO_TRAIN_FWD(RGB, SEM, F) render_train_fwd_kernel<RGB, SEM, F><<<grid, 128, 0, st>>>(V, R, P, workspace, bkgd_rand, O)
  if (want_sem) { if (fast) SO_TRAIN_FWD(true, true, true); else SO_TRAIN_FWD(true, true, false); }
  else if (want_rgb) { if (fast) SO_TRAIN_FWD(true, false, true); else SO_TRAIN_FWD(true, false, false); }
  else { if (fast) SO_TRAIN_FWD(false, false, true); else SO_TRAIN_FWD(false, false, false); }
#undef SO_TRAIN_FWD
  note_launch(1);
  return check_launch();
}

extern "C" int so_render_train_backward(const float* vol_sdf, const float* vol_feat, const so_volume_desc* vol_host,
                                        const float* cam_mats, const float* pix, const so_ray_desc* rd,
                                        const so_render_params* pr, const float* jitter, const float* bkgd_rand,
                                        const float* g_depth, const float* g_acc, const float* g_rgb, const float* g_sem,
                                        const float* g_weights, const float* g_eik, const float* g_sdf, float* g_vol_sdf,
                                        float* g_vol_feat, float* g_inv_s, float* workspace, void* stream) {
  bool want_rgb = g_rgb != nullptr, want_sem = g_sem != nullptr;
  if (!g_vol_sdf) return SO_ERR_INVALID_ARG;
  int rc = train_common_checks(vol_sdf, vol_feat, vol_host, cam_mats, rd, pr, workspace, want_rgb, want_sem, bkgd_rand);
  if (rc) return rc;
  if ((want_rgb || want_sem) && !g_vol_feat) return SO_ERR_INVALID_ARG;
  RayDev R;
  if ((rc = make_ray_dev(rd, cam_mats, pix, &R))) return rc;
  if (R.ray_count == 0) return SO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  VolumeDev V = make_volume(*vol_host, vol_sdf, vol_feat);
  RenderDev P = make_render_dev(*pr, jitter);
  if ((rc = launch_depth_bounds(R, P, workspace, st))) return rc;
  TrainGrad G{g_depth, g_acc, g_rgb, g_sem, g_weights, g_eik, g_sdf, g_vol_sdf, g_vol_feat, g_inv_s};
  long long ctas = ceil_div64(R.ray_count, 4);
  unsigned grid = (unsigned)(ctas < (long long)kNumSMs * 16 ? ctas : (long long)kNumSMs * 16);
  ProfScope prof(7, st);
  if (want_sem) render_train_bwd_kernel<true, true><<<grid, 128, 0, st>>>(V, R, P, workspace, bkgd_rand, G);
  else if (want_rgb) render_train_bwd_kernel<true, false><<<grid, 128, 0, st>>>(V, R, P, workspace, bkgd_rand, G);
  else render_train_bwd_kernel<false, false><<<grid, 128, 0, st>>>(V, R, P, workspace, bkgd_rand, G);
  note_launch(1);
  return check_launch();
}

extern "C" int so_field_query_backward(const so_volume_desc* vol_host, const float* points, int64_t n, const float* g_sdf,
                                       const float* g_grad, const float* g_feat, float* g_vol_sdf, float* g_vol_feat,
                                       void* stream) {
  if (!points || n < 0 || !g_vol_sdf) return SO_ERR_INVALID_ARG;
  int rc = validate_volume(vol_host);
  if (rc) return rc;
  if (g_feat && !g_vol_feat) return SO_ERR_INVALID_ARG;
  if (n == 0) return SO_OK;
  VolumeDev V = make_volume(*vol_host, nullptr, nullptr);
  field_query_bwd_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(V, points, n, g_sdf, g_grad, g_feat,
                                                                                         g_vol_sdf, g_vol_feat);
  note_launch(1);
  return check_launch();
}
