// HBM-bound passes of the UNet backbone on NHWC bf16 tensors: BatchNorm2d (training statistics come from the
// convolution epilogue) + exact-erf GELU apply and backward, bilinear x2 upsample forward / backward, and the
// per-channel sum used for bias gradients.
//
// Reference: ConvWithNorms (OpenSceneFlow/src/models/basic/__init__.py:61-79: Conv2d -> BatchNorm2d(eps 1e-5,
// momentum 0.1) -> GELU), BilinearDecoder (OpenSceneFlow/src/models/basic/unet.py:8-18:
// F.interpolate(scale_factor=2, mode="bilinear", align_corners=False)).
#include "common.cuh"
#include "../../include/deflow_b200.h"

namespace dfb {

struct bf8 { uint4 u; };

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(p[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

// 8 consecutive channels of element-vector `e`: one 16-byte load for bf16 tensors, two for fp32 (parity mode)
template <bool F32>
__device__ __forceinline__ void ld8v(const void* p, long long e, float* f) {
  if (F32) {
    const float4* q = reinterpret_cast<const float4*>(p) + 2 * e;
    const float4 a = __ldg(q), b = __ldg(q + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    unpack8(__ldg(reinterpret_cast<const uint4*>(p) + e), f);
  }
}
template <bool F32>
__device__ __forceinline__ void st8v(void* p, long long e, const float* f) {
  if (F32) {
    float4* q = reinterpret_cast<float4*>(p) + 2 * e;
    q[0] = make_float4(f[0], f[1], f[2], f[3]);
    q[1] = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    reinterpret_cast<uint4*>(p)[e] = pack8(f);
  }
}

// Exact-erf GELU and its derivative share one exponential: with u = |z|/sqrt(2), erf(u) = 1 - P(t) exp(-u^2),
// t = 1/(1 + p u) (Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7 -- far below bf16 resolution), and the Gaussian
// density of the derivative is the same exp(-z^2/2).  ~20 instructions instead of erff + expf (~60): these passes are
// otherwise instruction-bound, not HBM-bound.
__device__ __forceinline__ void gelu_parts(float z, float& cdf, float& pdf) {
  const float u = fabsf(z) * 0.70710678118654752f;
  const float e = __expf(-u * u);                       // exp(-z^2 / 2)
  const float t = __fdividef(1.0f, fmaf(0.3275911f, u, 1.0f));
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
  const float erf_abs = fmaf(-poly, e, 1.0f);
  cdf = 0.5f * (1.0f + copysignf(erf_abs, z));
  pdf = 0.3989422804014327f * e;
}
__device__ __forceinline__ float gelu_f(float z) { float c, p; gelu_parts(z, c, p); return z * c; }
__device__ __forceinline__ float gelu_grad_f(float z) { float c, p; gelu_parts(z, c, p); return fmaf(z, p, c); }

// ---------------------------------------------------------------- BatchNorm2d parameters from batch statistics
// stats [2][C] (double): sum, sum of squares over count elements per channel.  bn [4][C]: a = gamma*rstd,
// b = beta - mean*a, mean, rstd.  Training updates the running statistics (unbiased variance) in place.
__global__ void k_bn2d_finalize(const double* __restrict__ stats, double count, int C, int training, float eps,
                                float momentum, const float* __restrict__ gamma, const float* __restrict__ beta,
                                float* __restrict__ running_mean, float* __restrict__ running_var,
                                float* __restrict__ bn) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (training) {
    const double m = stats[c] / count;
    double v = stats[C + c] / count - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    var = (float)v;
    if (running_mean) {
      const double unbiased = count > 1.0 ? v * count / (count - 1.0) : v;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const float rstd = 1.0f / sqrtf(var + eps);
  const float a = gamma[c] * rstd;
  bn[c] = a;
  bn[C + c] = beta[c] - mean * a;
  bn[2 * C + c] = mean;
  bn[3 * C + c] = rstd;
}

// In all three passes a thread keeps ONE channel octet for its whole grid-stride loop (the stride is a multiple of
// the vectors-per-pixel count).  The per-channel parameters are read from shared memory inside the loop rather than
// held in registers: these kernels are latency-bound on HBM, so what matters is resident warps x bytes in flight
// (<= 40 registers per thread -> 6 CTAs per SM, 4 x 16-byte loads in flight per thread).
struct BnSmem { float a[256], b[256], mu[256], rs[256], m1[256], m2[256]; };

__device__ __forceinline__ void ld8(const float* p, float* f) {
  const float4 u = *reinterpret_cast<const float4*>(p), v = *reinterpret_cast<const float4*>(p + 4);
  f[0] = u.x; f[1] = u.y; f[2] = u.z; f[3] = u.w; f[4] = v.x; f[5] = v.y; f[6] = v.z; f[7] = v.w;
}

// y = GELU(a*x + b)
template <bool F32>
__global__ void __launch_bounds__(256, 4) k_bn_gelu_apply(const void* __restrict__ x, const float* __restrict__ bn, int C,
                                                          long long n_vec, void* __restrict__ y) {
  __shared__ __align__(16) float sa[256], sb[256];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { sa[i] = bn[i]; sb[i] = bn[C + i]; }
  __syncthreads();
  const int vpp = C >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long e0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)(e0 % vpp) << 3;
  for (long long e = e0; e < n_vec; e += 4 * stride) {
    float f[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j) { const long long ej = e + j * stride; if (ej < n_vec) ld8v<F32>(x, ej, f[j]); }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long ej = e + j * stride;
      if (ej < n_vec) {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[j][i] = gelu_f(fmaf(f[j][i], sa[c0 + i], sb[c0 + i]));
        st8v<F32>(y, ej, f[j]);
      }
    }
  }
}

// Backward pass 1: red[0][c] = sum g1, red[1][c] = sum g1 * xhat, with g1 = gy * GELU'(a*x + b)
template <bool F32>
__global__ void __launch_bounds__(256, 3) k_bn_gelu_bwd_reduce(const void* __restrict__ x, const void* __restrict__ gy,
                                                               const float* __restrict__ bn, int C, long long n_vec,
                                                               double* __restrict__ red) {
  __shared__ __align__(16) float sa[256], sb[256], smu[256], srs[256];
  __shared__ float acc1[256], acc2[256];
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    sa[i] = bn[i]; sb[i] = bn[C + i]; smu[i] = bn[2 * C + i]; srs[i] = bn[3 * C + i];
    acc1[i] = 0.f; acc2[i] = 0.f;
  }
  __syncthreads();
  const int vpp = C >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long e0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)(e0 % vpp) << 3;
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
  for (long long e = e0; e < n_vec; e += 2 * stride) {
    const long long e2 = e + stride;
    const bool two = e2 < n_vec;
    float fx[2][8], fg[2][8];
    ld8v<F32>(x, e, fx[0]); ld8v<F32>(gy, e, fg[0]);
    if (two) { ld8v<F32>(x, e2, fx[1]); ld8v<F32>(gy, e2, fg[1]); }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (j == 0 || two) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float g1 = fg[j][i] * gelu_grad_f(fmaf(fx[j][i], sa[c0 + i], sb[c0 + i]));
          s1[i] += g1;
          s2[i] = fmaf(g1, (fx[j][i] - smu[c0 + i]) * srs[c0 + i], s2[i]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { atomicAdd(&acc1[c0 + i], s1[i]); atomicAdd(&acc2[c0 + i], s2[i]); }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(&red[i], (double)acc1[i]);
    atomicAdd(&red[C + i], (double)acc2[i]);
  }
}

// Backward pass 2: gx = a * (g1 - m1 - xhat*m2) (training) or a * g1 (eval)
template <bool F32>
__global__ void __launch_bounds__(256, 3) k_bn_gelu_bwd_apply(const void* __restrict__ x, const void* __restrict__ gy,
                                                              const float* __restrict__ bn, const double* __restrict__ red,
                                                              double count, int training, int C, long long n_vec,
                                                              void* __restrict__ gx) {
  __shared__ __align__(16) float sa[256], sb[256], smu[256], srs[256], sm1[256], sm2[256];
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    sa[i] = bn[i]; sb[i] = bn[C + i]; smu[i] = bn[2 * C + i]; srs[i] = bn[3 * C + i];
    sm1[i] = training ? (float)(red[i] / count) : 0.f;
    sm2[i] = training ? (float)(red[C + i] / count) : 0.f;
  }
  __syncthreads();
  const int vpp = C >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long e0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)(e0 % vpp) << 3;
  for (long long e = e0; e < n_vec; e += 2 * stride) {
    const long long e2 = e + stride;
    const bool two = e2 < n_vec;
    float fx[2][8], fg[2][8];
    ld8v<F32>(x, e, fx[0]); ld8v<F32>(gy, e, fg[0]);
    if (two) { ld8v<F32>(x, e2, fx[1]); ld8v<F32>(gy, e2, fg[1]); }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (j == 0 || two) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float g1 = fg[j][i] * gelu_grad_f(fmaf(fx[j][i], sa[c0 + i], sb[c0 + i]));
          fx[j][i] = sa[c0 + i] * (g1 - sm1[c0 + i] - (fx[j][i] - smu[c0 + i]) * srs[c0 + i] * sm2[c0 + i]);
        }
        st8v<F32>(gx, j == 0 ? e : e2, fx[j]);
      }
    }
  }
}

// d gamma = sum g1*xhat, d beta = sum g1; conv bias gradient = sum gx = 0 (training) or a * sum g1 (eval)
__global__ void k_bn_param_grads(const double* __restrict__ red, const float* __restrict__ bn, int C, int training,
                                 float* __restrict__ g_gamma, float* __restrict__ g_beta, float* __restrict__ g_bias) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  g_gamma[c] += (float)red[C + c];
  g_beta[c] += (float)red[c];
  if (g_bias) g_bias[c] += training ? 0.f : bn[c] * (float)red[c];
}

// ---------------------------------------------------------------- per-channel sum (bias gradients) / sum of squares
template <bool F32>
__global__ void __launch_bounds__(256) k_channel_sum(const void* __restrict__ g, int C, long long n_pix,
                                                     float* __restrict__ out, double* __restrict__ stats2) {
  __shared__ float acc[256], acc2[256];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { acc[i] = 0.f; acc2[i] = 0.f; }
  __syncthreads();
  const int vpp = C >> 3;
  const int oct = threadIdx.x % vpp, prow = threadIdx.x / vpp, prows = blockDim.x / vpp;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  for (long long p = (long long)blockIdx.x * prows + prow; p < n_pix; p += (long long)gridDim.x * prows) {
    float f[8];
    ld8v<F32>(g, p * vpp + oct, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { atomicAdd(&acc[(oct << 3) + i], s[i]); if (stats2) atomicAdd(&acc2[(oct << 3) + i], q[i]); }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    if (out) atomicAdd(&out[i], acc[i]);
    if (stats2) { atomicAdd(&stats2[i], (double)acc[i]); atomicAdd(&stats2[C + i], (double)acc2[i]); }
  }
}

// ---------------------------------------------------------------- bilinear x2 (align_corners = False)
// out[2i] = 0.25 in[i-1] + 0.75 in[i], out[2i+1] = 0.75 in[i] + 0.25 in[i+1], indices clamped to the edge.
template <bool F32>
__global__ void __launch_bounds__(256) k_upsample2x(const void* __restrict__ in, int n, int h, int w, int C,
                                                    void* __restrict__ out) {
  const int vpp = C >> 3;
  const long long total = (long long)n * 4 * h * w * vpp;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(e % vpp);
    long long p = e / vpp;
    const int ox = (int)(p % (2 * w)); p /= 2 * w;
    const int oy = (int)(p % (2 * h));
    const int b = (int)(p / (2 * h));
    const int iy = oy >> 1, ix = ox >> 1;
    const int y2 = (oy & 1) ? min(iy + 1, h - 1) : max(iy - 1, 0);
    const int x2 = (ox & 1) ? min(ix + 1, w - 1) : max(ix - 1, 0);
    const long long base = (long long)b * h * w * vpp + v;
    float a[8], bq[8], c[8], d[8];
    ld8v<F32>(in, base + ((long long)iy * w + ix) * vpp, a);
    ld8v<F32>(in, base + ((long long)iy * w + x2) * vpp, bq);
    ld8v<F32>(in, base + ((long long)y2 * w + ix) * vpp, c);
    ld8v<F32>(in, base + ((long long)y2 * w + x2) * vpp, d);
#pragma unroll
    for (int i = 0; i < 8; ++i)  // rows first (like ATen: y-lerp of x-lerps), weights 0.75 / 0.25
      a[i] = 0.75f * (0.75f * a[i] + 0.25f * bq[i]) + 0.25f * (0.75f * c[i] + 0.25f * d[i]);
    st8v<F32>(out, e, a);
  }
}

__device__ __forceinline__ float up_w(int i, int o, int len) {
  // weight of in[i] in out[o] along one axis (len = input length)
  float wgt = 0.f;
  const int c = o >> 1;                                   // centre tap of out[o]
  const int nb = (o & 1) ? min(c + 1, len - 1) : max(c - 1, 0);
  if (c == i) wgt += 0.75f;
  if (nb == i) wgt += 0.25f;
  return wgt;
}

template <bool F32>
__global__ void __launch_bounds__(256) k_upsample2x_bwd(const void* __restrict__ gout, int n, int h, int w, int C,
                                                        void* __restrict__ gin) {
  const int vpp = C >> 3;
  const long long total = (long long)n * h * w * vpp;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(e % vpp);
    long long p = e / vpp;
    const int ix = (int)(p % w); p /= w;
    const int iy = (int)(p % h);
    const int b = (int)(p / h);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const long long base = (long long)b * 4 * h * w * vpp + v;
    for (int oy = max(2 * iy - 1, 0); oy <= min(2 * iy + 2, 2 * h - 1); ++oy) {
      const float wy = up_w(iy, oy, h);
      if (wy == 0.f) continue;
      for (int ox = max(2 * ix - 1, 0); ox <= min(2 * ix + 2, 2 * w - 1); ++ox) {
        const float wx = up_w(ix, ox, w);
        if (wx == 0.f) continue;
        float f[8];
        ld8v<F32>(gout, base + ((long long)oy * 2 * w + ox) * vpp, f);
        const float wt = wy * wx;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(wt, f[i], acc[i]);
      }
    }
    st8v<F32>(gin, e, acc);
  }
}

static int grid_for_elems(long long work, int block, int mult = 16) {
  long long b = (work + block - 1) / block;
  const long long cap = (long long)sm_count() * mult;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace dfb

using namespace dfb;

extern "C" int dfb_bn2d_finalize(const double* stats, double count, int C, int training, float eps, float momentum,
                                 const float* gamma, const float* beta, float* running_mean, float* running_var,
                                 float* bn, void* stream_) {
  if (C <= 0 || C > 256) { set_error("dfb_bn2d_finalize: C must be in 1..256"); return DFB_ERR_ARG; }
  k_bn2d_finalize<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream_>>>(stats, count, C, training, eps, momentum, gamma, beta,
                                                                      running_mean, running_var, bn);
  add_launches(1);
  return check_launch("dfb_bn2d_finalize");
}

extern "C" int dfb_bn_gelu_apply(const void* x, const float* bn, int C, long long n_pix, void* y, int f32, void* stream_) {
  if (C % 8 || C > 256) { set_error("dfb_bn_gelu_apply: C must be a multiple of 8 and <= 256"); return DFB_ERR_ARG; }
  const long long n_vec = n_pix * (C >> 3);
  const int g = grid_for_elems(n_vec, 1024, 12);
  if (f32) k_bn_gelu_apply<true><<<g, 256, 0, (cudaStream_t)stream_>>>(x, bn, C, n_vec, y);
  else k_bn_gelu_apply<false><<<g, 256, 0, (cudaStream_t)stream_>>>(x, bn, C, n_vec, y);
  add_launches(1);
  return check_launch("dfb_bn_gelu_apply");
}

extern "C" int dfb_bn_gelu_backward(const void* x, const void* gy, const float* bn, int C, long long n_pix, int training,
                                    double* red, void* gx, float* g_gamma, float* g_beta, float* g_bias, int f32,
                                    void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (C % 8 || C > 256 || 256 % (C >> 3)) { set_error("dfb_bn_gelu_backward: unsupported channel count %d", C); return DFB_ERR_ARG; }
  cudaMemsetAsync(red, 0, sizeof(double) * 2 * C, st);
  const long long n_vec = n_pix * (C >> 3);
  const int g = grid_for_elems(n_vec, 512, 8);
  if (f32) {
    k_bn_gelu_bwd_reduce<true><<<g, 256, 0, st>>>(x, gy, bn, C, n_vec, red);
    k_bn_gelu_bwd_apply<true><<<g, 256, 0, st>>>(x, gy, bn, red, (double)n_pix, training, C, n_vec, gx);
  } else {
    k_bn_gelu_bwd_reduce<false><<<g, 256, 0, st>>>(x, gy, bn, C, n_vec, red);
    k_bn_gelu_bwd_apply<false><<<g, 256, 0, st>>>(x, gy, bn, red, (double)n_pix, training, C, n_vec, gx);
  }
  k_bn_param_grads<<<(C + 127) / 128, 128, 0, st>>>(red, bn, C, training, g_gamma, g_beta, g_bias);
  add_launches(3);
  return check_launch("dfb_bn_gelu_backward");
}

// out[c] += sum over pixels (may be NULL); stats2 (may be NULL): [2][C] += sum, sum of squares (BatchNorm statistics of
// an fp32 tensor in parity mode)
extern "C" int dfb_channel_sum(const void* g, int C, long long n_pix, float* out, double* stats2, int f32, void* stream_) {
  if (C % 8 || C > 256 || 256 % (C >> 3)) { set_error("dfb_channel_sum: unsupported channel count %d", C); return DFB_ERR_ARG; }
  const int prows = 256 / (C >> 3);
  const int gr = grid_for_elems(n_pix, prows, 8);
  if (f32) k_channel_sum<true><<<gr, 256, 0, (cudaStream_t)stream_>>>(g, C, n_pix, out, stats2);
  else k_channel_sum<false><<<gr, 256, 0, (cudaStream_t)stream_>>>(g, C, n_pix, out, stats2);
  add_launches(1);
  return check_launch("dfb_channel_sum");
}

extern "C" int dfb_upsample2x(const void* in, int n, int h, int w, int C, void* out, int backward, int f32, void* stream_) {
  if (C % 8) { set_error("dfb_upsample2x: C must be a multiple of 8"); return DFB_ERR_ARG; }
  const long long vec_in = (long long)n * h * w * (C >> 3);
  cudaStream_t st = (cudaStream_t)stream_;
  if (!backward) {
    const int g = grid_for_elems(vec_in * 4, 256);
    if (f32) k_upsample2x<true><<<g, 256, 0, st>>>(in, n, h, w, C, out);
    else k_upsample2x<false><<<g, 256, 0, st>>>(in, n, h, w, C, out);
  } else {
    const int g = grid_for_elems(vec_in, 256);
    if (f32) k_upsample2x_bwd<true><<<g, 256, 0, st>>>(in, n, h, w, C, out);
    else k_upsample2x_bwd<false><<<g, 256, 0, st>>>(in, n, h, w, C, out);
  }
  add_launches(1);
  return check_launch("dfb_upsample2x");
}
