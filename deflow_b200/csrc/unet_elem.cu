// HBM-bound passes of the UNet backbone on NHWC bf16 tensors: BatchNorm2d (training statistics come from the
// convolution epilogue) + exact-erf GELU apply and backward, bilinear x2 upsample forward / backward, and the
// per-channel sum used for bias gradients.
//
// Reference: ConvWithNorms (OpenSceneFlow/src/models/basic/__init__.py:61-79: Conv2d -> BatchNorm2d(eps 1e-5,
// momentum 0.1) -> GELU), BilinearDecoder (OpenSceneFlow/src/models/basic/unet.py:8-18:
// F.interpolate(scale_factor=2, mode="bilinear", align_corners=False)).
#include "common.cuh"
#include "stream_pipe.cuh"
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

// ================================================================ bf16 streaming variants (perf mode)
// The same three BatchNorm/GELU passes and the per-channel sum for bf16 tensors, restructured for B200:
//   * inputs arrive through the bulk-async ring of stream_pipe.cuh (bytes in flight independent of the ALU phase);
//   * the arithmetic is packed fp32x2 (FFMA2 / FMUL2 / FADD2: two channels per instruction), the per-channel
//     parameters of the thread's fixed channel octet live in registers, bf16 <-> fp32 is a shift / one cvt per pair.
// ~13 issue slots + 2 MUFU per element instead of ~30 + 2: the passes become HBM-bound instead of issue-bound.
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float2 f2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 bf2_to_f2(unsigned w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
__device__ __forceinline__ unsigned f2_to_bf2(float2 v) {
  const __nv_bfloat162 p = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<const unsigned*>(&p);
}

// cdf = Phi(z), pdf = phi(z) for two channels (same Abramowitz & Stegun 7.1.26 form as gelu_parts; signs of the
// polynomial folded into its coefficients, exp through ex2 with the 1/2 and log2(e) folded into one constant)
__device__ __forceinline__ void gelu_parts2(float2 z, float2& cdf, float2& pdf) {
  const float2 e2 = __fmul2_rn(__fmul2_rn(z, z), f2(-0.72134752044448170368f));   // -z^2/2 * log2(e)
  const float2 e = make_float2(ex2_approx(e2.x), ex2_approx(e2.y));                // exp(-z^2/2)
  const float2 az = make_float2(fabsf(z.x), fabsf(z.y));
  const float2 den = __ffma2_rn(az, f2(0.3275911f * 0.70710678118654752f), f2(1.0f));
  const float2 t = make_float2(rcp_approx(den.x), rcp_approx(den.y));
  float2 q = __ffma2_rn(t, f2(-1.061405429f), f2(1.453152027f));
  q = __ffma2_rn(t, q, f2(-1.421413741f));
  q = __ffma2_rn(t, q, f2(0.284496736f));
  q = __ffma2_rn(t, q, f2(-0.254829592f));
  q = __fmul2_rn(q, t);                                                            // -P(t)
  const float2 ea = __ffma2_rn(q, e, f2(1.0f));                                    // erf(|z|/sqrt2)
  const float2 es = make_float2(copysignf(ea.x, z.x), copysignf(ea.y, z.y));
  cdf = __ffma2_rn(es, f2(0.5f), f2(0.5f));
  pdf = __fmul2_rn(e, f2(0.3989422804014327f));
}

// per-thread parameter octet: 4 float2 per array
__device__ __forceinline__ void ld_oct(const float* p, float2 (&d)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) d[i] = make_float2(p[2 * i], p[2 * i + 1]);
}

constexpr int SP_STAGES = 4;

// y = GELU(a*x + b)
__global__ void __launch_bounds__(sp::THREADS, 4) k_bn_gelu_apply_s(const void* __restrict__ x, const float* __restrict__ bn,
                                                                   int C, long long n_vec, uint4* __restrict__ y) {
  extern __shared__ __align__(128) uint8_t sp_smem[];
  sp::Ring<1, SP_STAGES> ring(sp_smem);
  ring.init();
  if (threadIdx.x >= sp::CONSUMERS) {
    if (threadIdx.x == sp::CONSUMERS) { const void* const in[1] = {x}; ring.produce(in, n_vec); }
    return;
  }
  const int c0 = (threadIdx.x % (C >> 3)) << 3;
  float2 a[4], b[4];
  ld_oct(bn + c0, a); ld_oct(bn + C + c0, b);
  const long long n_chunks = (n_vec + sp::CHUNK_VEC - 1) / sp::CHUNK_VEC;
  int it = 0;
  for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x, ++it) {
    ring.wait_full(it);
    const uint4* sx = ring.slot(it % SP_STAGES, 0);
    const uint4 u0 = sx[threadIdx.x], u1 = sx[threadIdx.x + 256];
    ring.release(it);
    const long long e = ch * sp::CHUNK_VEC + threadIdx.x;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const uint4 u = v ? u1 : u0;
      const unsigned w[4] = {u.x, u.y, u.z, u.w};
      unsigned o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 z = __ffma2_rn(bf2_to_f2(w[i]), a[i], b[i]);
        float2 cdf, pdf;
        gelu_parts2(z, cdf, pdf);
        o[i] = f2_to_bf2(__fmul2_rn(z, cdf));
      }
      if (e + 256 * v < n_vec) y[e + 256 * v] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// Backward pass 1: red[0][c] = sum g1, red[1][c] = sum g1 * xhat, with g1 = gy * GELU'(a*x + b)
__global__ void __launch_bounds__(sp::THREADS, 2) k_bn_gelu_bwd_reduce_s(const void* __restrict__ x, const void* __restrict__ gy,
                                                                        const float* __restrict__ bn, int C, long long n_vec,
                                                                        double* __restrict__ red) {
  extern __shared__ __align__(128) uint8_t sp_smem[];
  __shared__ float acc1[256], acc2[256];
  sp::Ring<2, SP_STAGES> ring(sp_smem);
  for (int i = threadIdx.x; i < C; i += blockDim.x) { acc1[i] = 0.f; acc2[i] = 0.f; }
  ring.init();
  if (threadIdx.x >= sp::CONSUMERS) {
    if (threadIdx.x == sp::CONSUMERS) { const void* const in[2] = {x, gy}; ring.produce(in, n_vec); }
  } else {
    const int c0 = (threadIdx.x % (C >> 3)) << 3;
    float2 a[4], b[4], rs[4], nmr[4], s1[4], s2[4];
    ld_oct(bn + c0, a); ld_oct(bn + C + c0, b); ld_oct(bn + 3 * C + c0, rs);
    {
      float2 mu[4];
      ld_oct(bn + 2 * C + c0, mu);
#pragma unroll
      for (int i = 0; i < 4; ++i) { nmr[i] = make_float2(-mu[i].x * rs[i].x, -mu[i].y * rs[i].y); s1[i] = f2(0.f); s2[i] = f2(0.f); }
    }
    const long long n_chunks = (n_vec + sp::CHUNK_VEC - 1) / sp::CHUNK_VEC;
    int it = 0;
    for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x, ++it) {
      ring.wait_full(it);
      const uint4* sx = ring.slot(it % SP_STAGES, 0);
      const uint4* sg = ring.slot(it % SP_STAGES, 1);
      const bool two = ch * sp::CHUNK_VEC + threadIdx.x + 256 < n_vec, one = ch * sp::CHUNK_VEC + threadIdx.x < n_vec;
      const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
      // a partial last chunk leaves stale bytes behind the copied range: read those vectors as zeros
      const uint4 x0 = one ? sx[threadIdx.x] : z4, x1 = two ? sx[threadIdx.x + 256] : z4;
      const uint4 g0 = one ? sg[threadIdx.x] : z4, g1v = two ? sg[threadIdx.x + 256] : z4;
      ring.release(it);
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const uint4 ux = v ? x1 : x0, ug = v ? g1v : g0;
        const unsigned wx[4] = {ux.x, ux.y, ux.z, ux.w}, wg[4] = {ug.x, ug.y, ug.z, ug.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 xf = bf2_to_f2(wx[i]);
          const float2 z = __ffma2_rn(xf, a[i], b[i]);
          float2 cdf, pdf;
          gelu_parts2(z, cdf, pdf);
          const float2 g1 = __fmul2_rn(bf2_to_f2(wg[i]), __ffma2_rn(z, pdf, cdf));
          s1[i] = __fadd2_rn(s1[i], g1);
          s2[i] = __ffma2_rn(g1, __ffma2_rn(xf, rs[i], nmr[i]), s2[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(&acc1[c0 + 2 * i], s1[i].x); atomicAdd(&acc1[c0 + 2 * i + 1], s1[i].y);
      atomicAdd(&acc2[c0 + 2 * i], s2[i].x); atomicAdd(&acc2[c0 + 2 * i + 1], s2[i].y);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(&red[i], (double)acc1[i]);
    atomicAdd(&red[C + i], (double)acc2[i]);
  }
}

// Backward pass 2: gx = a * (g1 - m1 - xhat*m2) (training) or a * g1 (eval), as a*g1 + (c1*x + c0) with
// c1 = -a*m2*rstd, c0 = -a*(m1 - m2*mean*rstd)
__global__ void __launch_bounds__(sp::THREADS, 2) k_bn_gelu_bwd_apply_s(const void* __restrict__ x, const void* __restrict__ gy,
                                                                       const float* __restrict__ bn, const double* __restrict__ red,
                                                                       double count, int training, int C, long long n_vec,
                                                                       uint4* __restrict__ gx) {
  extern __shared__ __align__(128) uint8_t sp_smem[];
  sp::Ring<2, SP_STAGES> ring(sp_smem);
  ring.init();
  if (threadIdx.x >= sp::CONSUMERS) {
    if (threadIdx.x == sp::CONSUMERS) { const void* const in[2] = {x, gy}; ring.produce(in, n_vec); }
    return;
  }
  const int c0 = (threadIdx.x % (C >> 3)) << 3;
  float2 a[4], b[4], k0[4], k1[4];
  ld_oct(bn + c0, a); ld_oct(bn + C + c0, b);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v0[2], v1[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = c0 + 2 * i + h;
      const float m1 = training ? (float)(red[c] / count) : 0.f, m2 = training ? (float)(red[C + c] / count) : 0.f;
      const float ac = bn[c], mu = bn[2 * C + c], r = bn[3 * C + c];
      v1[h] = -ac * m2 * r;
      v0[h] = -ac * (m1 - m2 * mu * r);
    }
    k0[i] = make_float2(v0[0], v0[1]); k1[i] = make_float2(v1[0], v1[1]);
  }
  const long long n_chunks = (n_vec + sp::CHUNK_VEC - 1) / sp::CHUNK_VEC;
  int it = 0;
  for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x, ++it) {
    ring.wait_full(it);
    const uint4* sx = ring.slot(it % SP_STAGES, 0);
    const uint4* sg = ring.slot(it % SP_STAGES, 1);
    const uint4 x0 = sx[threadIdx.x], x1 = sx[threadIdx.x + 256];
    const uint4 g0 = sg[threadIdx.x], g1v = sg[threadIdx.x + 256];
    ring.release(it);
    const long long e = ch * sp::CHUNK_VEC + threadIdx.x;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const uint4 ux = v ? x1 : x0, ug = v ? g1v : g0;
      const unsigned wx[4] = {ux.x, ux.y, ux.z, ux.w}, wg[4] = {ug.x, ug.y, ug.z, ug.w};
      unsigned o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 xf = bf2_to_f2(wx[i]);
        const float2 z = __ffma2_rn(xf, a[i], b[i]);
        float2 cdf, pdf;
        gelu_parts2(z, cdf, pdf);
        const float2 g1 = __fmul2_rn(bf2_to_f2(wg[i]), __ffma2_rn(z, pdf, cdf));
        o[i] = f2_to_bf2(__ffma2_rn(a[i], g1, __ffma2_rn(k1[i], xf, k0[i])));
      }
      if (e + 256 * v < n_vec) gx[e + 256 * v] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// out[c] += sum over pixels (bias gradients of the un-normalised decoder convolutions and of the GRU gates)
__global__ void __launch_bounds__(sp::THREADS, 4) k_channel_sum_s(const void* __restrict__ g, int C, long long n_vec,
                                                                 float* __restrict__ out) {
  extern __shared__ __align__(128) uint8_t sp_smem[];
  __shared__ float acc[256];
  sp::Ring<1, SP_STAGES> ring(sp_smem);
  for (int i = threadIdx.x; i < C; i += blockDim.x) acc[i] = 0.f;
  ring.init();
  if (threadIdx.x >= sp::CONSUMERS) {
    if (threadIdx.x == sp::CONSUMERS) { const void* const in[1] = {g}; ring.produce(in, n_vec); }
  } else {
    const int c0 = (threadIdx.x % (C >> 3)) << 3;
    float2 s[4] = {f2(0.f), f2(0.f), f2(0.f), f2(0.f)};
    const long long n_chunks = (n_vec + sp::CHUNK_VEC - 1) / sp::CHUNK_VEC;
    int it = 0;
    for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x, ++it) {
      ring.wait_full(it);
      const uint4* sg = ring.slot(it % SP_STAGES, 0);
      const long long e = ch * sp::CHUNK_VEC + threadIdx.x;
      const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
      const uint4 u0 = e < n_vec ? sg[threadIdx.x] : z4, u1 = e + 256 < n_vec ? sg[threadIdx.x + 256] : z4;
      ring.release(it);
      s[0] = __fadd2_rn(s[0], __fadd2_rn(bf2_to_f2(u0.x), bf2_to_f2(u1.x)));
      s[1] = __fadd2_rn(s[1], __fadd2_rn(bf2_to_f2(u0.y), bf2_to_f2(u1.y)));
      s[2] = __fadd2_rn(s[2], __fadd2_rn(bf2_to_f2(u0.z), bf2_to_f2(u1.z)));
      s[3] = __fadd2_rn(s[3], __fadd2_rn(bf2_to_f2(u0.w), bf2_to_f2(u1.w)));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { atomicAdd(&acc[c0 + 2 * i], s[i].x); atomicAdd(&acc[c0 + 2 * i + 1], s[i].y); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&out[i], acc[i]);
}

// ---------------------------------------------------------------- bilinear x2, bf16 streaming variants
// Forward: a thread owns one INPUT pixel octet and writes the 2x2 output quad it centres: 9 loads (mostly L1 hits,
// the 3x3 neighbourhoods of adjacent threads overlap) for 4 stores, 32-bit index arithmetic, packed fp32x2 lerps.
// Same rounding order as the generic kernel: x-lerp (0.75 centre + 0.25 neighbour) then y-lerp.
__device__ __forceinline__ void unpack4(const uint4& u, float2 (&f)[4]) {
  f[0] = bf2_to_f2(u.x); f[1] = bf2_to_f2(u.y); f[2] = bf2_to_f2(u.z); f[3] = bf2_to_f2(u.w);
}
__device__ __forceinline__ uint4 pack4(const float2 (&f)[4]) {
  return make_uint4(f2_to_bf2(f[0]), f2_to_bf2(f[1]), f2_to_bf2(f[2]), f2_to_bf2(f[3]));
}
__device__ __forceinline__ void lerp4(const float2 (&c)[4], const float2 (&nb)[4], float2 (&o)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = __ffma2_rn(c[i], f2(0.75f), __fmul2_rn(nb[i], f2(0.25f)));
}

__global__ void __launch_bounds__(256) k_upsample2x_q(const uint4* __restrict__ in, int n, int h, int w, int vshift,
                                                      uint4* __restrict__ out) {
  const int vpp = 1 << vshift;
  const unsigned total = (unsigned)n * h * w * vpp;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int v = e & (vpp - 1);
    unsigned p = e >> vshift;
    const int ix = p % w; p /= w;
    const int iy = p % h;
    const int b = p / h;
    const int xm = max(ix - 1, 0), xp = min(ix + 1, w - 1);
    const int ym = max(iy - 1, 0), yp = min(iy + 1, h - 1);
    const uint4* base = in + (((size_t)b * h * w) << vshift) + v;
    float2 L[3][4], R[3][4];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int y = r == 0 ? ym : r == 1 ? iy : yp;
      const uint4* row = base + (((size_t)y * w) << vshift);
      float2 c[4], l[4], rr[4];
      unpack4(__ldg(row + ((size_t)ix << vshift)), c);
      unpack4(__ldg(row + ((size_t)xm << vshift)), l);
      unpack4(__ldg(row + ((size_t)xp << vshift)), rr);
      lerp4(c, l, L[r]);
      lerp4(c, rr, R[r]);
    }
    float2 o[4];
    uint4* ob = out + ((((size_t)b * 2 * h + 2 * iy) * 2 * w + 2 * ix) << vshift) + v;
    const size_t orow = ((size_t)2 * w) << vshift;
    lerp4(L[1], L[0], o); ob[0] = pack4(o);
    lerp4(R[1], R[0], o); ob[vpp] = pack4(o);
    lerp4(L[1], L[2], o); ob[orow] = pack4(o);
    lerp4(R[1], R[2], o); ob[orow + vpp] = pack4(o);
  }
}

// Backward: gin[iy][ix] = sum over the 4x4 output window 2i-1 .. 2i+2 of wy*wx*gout (weights 0.25/0.75/0.75/0.25 inside,
// folded at the edges by up_w); rows combined horizontally first.
__global__ void __launch_bounds__(256) k_upsample2x_bwd_q(const uint4* __restrict__ gout, int n, int h, int w, int vshift,
                                                          uint4* __restrict__ gin) {
  const int vpp = 1 << vshift;
  const unsigned total = (unsigned)n * h * w * vpp;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int v = e & (vpp - 1);
    unsigned p = e >> vshift;
    const int ix = p % w; p /= w;
    const int iy = p % h;
    const int b = p / h;
    float wx[4], wy[4];
    int ox[4], oy[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = 2 * ix - 1 + j, y = 2 * iy - 1 + j;
      ox[j] = min(max(x, 0), 2 * w - 1); oy[j] = min(max(y, 0), 2 * h - 1);
      wx[j] = (x < 0 || x >= 2 * w) ? 0.f : up_w(ix, x, w);
      wy[j] = (y < 0 || y >= 2 * h) ? 0.f : up_w(iy, y, h);
    }
    const uint4* base = gout + (((size_t)b * 4 * h * w) << vshift) + v;
    float2 acc[4] = {f2(0.f), f2(0.f), f2(0.f), f2(0.f)};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint4* row = base + (((size_t)oy[r] * 2 * w) << vshift);
      float2 t[4] = {f2(0.f), f2(0.f), f2(0.f), f2(0.f)};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 g[4];
        unpack4(__ldg(row + ((size_t)ox[j] << vshift)), g);
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = __ffma2_rn(g[i], f2(wx[j]), t[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = __ffma2_rn(t[i], f2(wy[r]), acc[i]);
    }
    gin[e] = pack4(acc);
  }
}

// one wave: CTAs per SM from the occupancy calculator (cached per kernel), never more CTAs than chunks
template <class K>
static int stream_grid(K kernel, int smem, long long n_vec, int* cache) {
  if (*cache == 0) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, sp::THREADS, smem);
    *cache = per_sm > 0 ? per_sm : 1;
  }
  long long g = (long long)sm_count() * *cache;
  const long long n_chunks = (n_vec + sp::CHUNK_VEC - 1) / sp::CHUNK_VEC;
  if (g > n_chunks) g = n_chunks;
  return (int)(g < 1 ? 1 : g);
}

static bool stream_ok(int C, const void* p0, const void* p1 = nullptr, const void* p2 = nullptr) {
  return C % 8 == 0 && C <= 256 && 256 % (C >> 3) == 0 && !(((uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2) & 15);
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
  if (!f32 && stream_ok(C, x, y)) {
    static int occ = 0;
    constexpr int smem = sp::Ring<1, SP_STAGES>::BYTES;
    const int g = stream_grid(k_bn_gelu_apply_s, smem, n_vec, &occ);
    k_bn_gelu_apply_s<<<g, sp::THREADS, smem, (cudaStream_t)stream_>>>(x, bn, C, n_vec, (uint4*)y);
    add_launches(1);
    return check_launch("dfb_bn_gelu_apply");
  }
  const int g = grid_for_elems(n_vec, 1024, 12);
  if (f32) k_bn_gelu_apply<true><<<g, 256, 0, (cudaStream_t)stream_>>>(x, bn, C, n_vec, y);
  else k_bn_gelu_apply<false><<<g, 256, 0, (cudaStream_t)stream_>>>(x, bn, C, n_vec, y);
  add_launches(1);
  return check_launch("dfb_bn_gelu_apply");
}

// phase 0: everything.  SyncBatchNorm callers split the call around their all-reduce of `red`:
//   phase 1: red[2][C] = this rank's sums (sum g1, sum g1 * xhat) and the parameter gradients from them (local sums, as
//            torch.nn.SyncBatchNorm leaves them to the data-parallel gradient mean);
//   phase 2: gx from `red` (now summed over all ranks) and count_total = pixels per channel over all ranks.
static int bn_gelu_backward_impl(const void* x, const void* gy, const float* bn, int C, long long n_pix, int training,
                                 double* red, void* gx, float* g_gamma, float* g_beta, float* g_bias, int f32, int phase,
                                 double count_total, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (C % 8 || C > 256 || 256 % (C >> 3)) { set_error("dfb_bn_gelu_backward: unsupported channel count %d", C); return DFB_ERR_ARG; }
  if (phase < 0 || phase > 2) { set_error("dfb_bn_gelu_backward: phase must be 0, 1 or 2"); return DFB_ERR_ARG; }
  const bool do_reduce = phase != 2, do_apply = phase != 1;
  if (do_reduce) cudaMemsetAsync(red, 0, sizeof(double) * 2 * C, st);
  const long long n_vec = n_pix * (C >> 3);
  const int g = grid_for_elems(n_vec, 512, 8);
  const double cnt = phase == 2 ? count_total : (double)n_pix;
  int launches = 0;
  if (!f32 && stream_ok(C, x, gy, gx)) {
    static int occ_r = 0, occ_a = 0;
    constexpr int smem = sp::Ring<2, SP_STAGES>::BYTES;
    const int gr = stream_grid(k_bn_gelu_bwd_reduce_s, smem, n_vec, &occ_r);
    const int ga = stream_grid(k_bn_gelu_bwd_apply_s, smem, n_vec, &occ_a);
    if (do_reduce) { k_bn_gelu_bwd_reduce_s<<<gr, sp::THREADS, smem, st>>>(x, gy, bn, C, n_vec, red); ++launches; }
    if (do_apply) { k_bn_gelu_bwd_apply_s<<<ga, sp::THREADS, smem, st>>>(x, gy, bn, red, cnt, training, C, n_vec, (uint4*)gx); ++launches; }
  } else if (f32) {
    if (do_reduce) { k_bn_gelu_bwd_reduce<true><<<g, 256, 0, st>>>(x, gy, bn, C, n_vec, red); ++launches; }
    if (do_apply) { k_bn_gelu_bwd_apply<true><<<g, 256, 0, st>>>(x, gy, bn, red, cnt, training, C, n_vec, gx); ++launches; }
  } else {
    if (do_reduce) { k_bn_gelu_bwd_reduce<false><<<g, 256, 0, st>>>(x, gy, bn, C, n_vec, red); ++launches; }
    if (do_apply) { k_bn_gelu_bwd_apply<false><<<g, 256, 0, st>>>(x, gy, bn, red, cnt, training, C, n_vec, gx); ++launches; }
  }
  if (do_reduce) { k_bn_param_grads<<<(C + 127) / 128, 128, 0, st>>>(red, bn, C, training, g_gamma, g_beta, g_bias); ++launches; }
  add_launches(launches);
  return check_launch("dfb_bn_gelu_backward");
}

extern "C" int dfb_bn_gelu_backward(const void* x, const void* gy, const float* bn, int C, long long n_pix, int training,
                                    double* red, void* gx, float* g_gamma, float* g_beta, float* g_bias, int f32,
                                    void* stream_) {
  return bn_gelu_backward_impl(x, gy, bn, C, n_pix, training, red, gx, g_gamma, g_beta, g_bias, f32, 0, 0.0, stream_);
}

extern "C" int dfb_bn_gelu_backward_phase(const void* x, const void* gy, const float* bn, int C, long long n_pix, int training,
                                          double* red, void* gx, float* g_gamma, float* g_beta, float* g_bias, int f32,
                                          int phase, double count_total, void* stream_) {
  return bn_gelu_backward_impl(x, gy, bn, C, n_pix, training, red, gx, g_gamma, g_beta, g_bias, f32, phase, count_total, stream_);
}

// out[c] += sum over pixels (may be NULL); stats2 (may be NULL): [2][C] += sum, sum of squares (BatchNorm statistics of
// an fp32 tensor in parity mode)
extern "C" int dfb_channel_sum(const void* g, int C, long long n_pix, float* out, double* stats2, int f32, void* stream_) {
  if (C % 8 || C > 256 || 256 % (C >> 3)) { set_error("dfb_channel_sum: unsupported channel count %d", C); return DFB_ERR_ARG; }
  if (!f32 && !stats2 && out && stream_ok(C, g)) {
    static int occ = 0;
    constexpr int smem = sp::Ring<1, SP_STAGES>::BYTES;
    const long long n_vec = n_pix * (C >> 3);
    const int gr = stream_grid(k_channel_sum_s, smem, n_vec, &occ);
    k_channel_sum_s<<<gr, sp::THREADS, smem, (cudaStream_t)stream_>>>(g, C, n_vec, out);
    add_launches(1);
    return check_launch("dfb_channel_sum");
  }
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
  const int vpp = C >> 3;
  if (!f32 && (vpp & (vpp - 1)) == 0 && vec_in * 4 < (1ll << 31) && !(((uintptr_t)in | (uintptr_t)out) & 15)) {
    int vshift = 0;
    while ((1 << vshift) < vpp) ++vshift;
    const int g = grid_for_elems(vec_in, 256, 8);
    if (!backward) k_upsample2x_q<<<g, 256, 0, st>>>((const uint4*)in, n, h, w, vshift, (uint4*)out);
    else k_upsample2x_bwd_q<<<g, 256, 0, st>>>((const uint4*)in, n, h, w, vshift, (uint4*)out);
    add_launches(1);
    return check_launch("dfb_upsample2x");
  }
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

// ------------------------------------------------------------------------------------------------
// Per-channel sums of a 3x3 / stride 1 / pad 1 data gradient WITHOUT touching the data gradient: the bias gradient of the
// convolution that produced this layer's input (UpsampleSkip, OpenSceneFlow/src/models/basic/unet.py:25-37: conv -> conv with
// bias and nothing in between).  gx[p] = sum_taps W_tap^T gy[p - off_tap] with gy zero outside the image, hence
//   sum_p gx[p, ci] = sum_{ky,kx} sum_co W[co, ci, ky, kx] * S_{ky,kx}[co],
//   S_{ky,kx} = T - [ky == 0] top row - [ky == 2] bottom row - [kx == 0] left column - [kx == 2] right column + corner,
// where T = sum of gy over all pixels (this layer's own bias gradient, already known) and the border terms are sums of gy
// over one image row / column / corner pixel (summed over the batch).  Taking these sums in the epilogue of the 64-channel
// row-pair data-gradient kernel cost +50 % of that launch (B200: 0.235 -> 0.355 ms at 512^2); the border sums read
// 2(H + W) pixels per image.
namespace dfb {

// out [8][C] (+=): top row, bottom row, left column, right column, corners TL, TR, BL, BR
template <bool F32>
__global__ void __launch_bounds__(256) k_border_sums(const void* __restrict__ gy, int n, int H, int W, int C, float* __restrict__ out) {
  extern __shared__ float s_acc[];    // [8][C]
  for (int i = threadIdx.x; i < 8 * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int per_img = 2 * W + 2 * H;
  const long long total = (long long)n * per_img;
  const int lanes_per_pix = C;                       // thread = (pixel slot, channel)
  const int pix_per_iter = blockDim.x / lanes_per_pix > 0 ? blockDim.x / lanes_per_pix : 1;
  const int c = threadIdx.x % lanes_per_pix, slot = threadIdx.x / lanes_per_pix;
  if (slot < pix_per_iter) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long b = (long long)blockIdx.x * pix_per_iter + slot; b < total; b += (long long)gridDim.x * pix_per_iter) {
      const int img = (int)(b / per_img);
      int r = (int)(b % per_img);
      int kind, y, x;
      if (r < W) { kind = 0; y = 0; x = r; }
      else if (r < 2 * W) { kind = 1; y = H - 1; x = r - W; }
      else if (r < 2 * W + H) { kind = 2; y = r - 2 * W; x = 0; }
      else { kind = 3; y = r - 2 * W - H; x = W - 1; }
      const size_t off = (((size_t)img * H + y) * W + x) * C + c;
      const float v = F32 ? reinterpret_cast<const float*>(gy)[off] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(gy)[off]);
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] += (kind == k) ? v : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) if (acc[k] != 0.f) atomicAdd(&s_acc[k * C + c], acc[k]);
  }
  // corners: block 0, thread = channel (looped), summed over the batch
  if (blockIdx.x == 0) {
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      for (int img = 0; img < n; ++img) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int y = (k & 2) ? H - 1 : 0, x = (k & 1) ? W - 1 : 0;
          const size_t off = (((size_t)img * H + y) * W + x) * C + ch;
          t[k] += F32 ? reinterpret_cast<const float*>(gy)[off] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(gy)[off]);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) s_acc[(4 + k) * C + ch] += t[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * C; i += blockDim.x) {
    if (s_acc[i] != 0.f) atomicAdd(out + i, s_acc[i]);
  }
}

// colsum[ci] = sum_taps sum_co W[co, cin_off + ci, ky, kx] * S_tap[co]; one block per ci
__global__ void __launch_bounds__(128) k_colsum_from_borders(const float* __restrict__ border, const float* __restrict__ total,
                                                             const float* __restrict__ w, int cout, int cin_total, int cin_off,
                                                             float* __restrict__ colsum) {
  const int ci = blockIdx.x;
  double acc = 0.0;
  for (int co = threadIdx.x; co < cout; co += blockDim.x) {
    const float T = total[co];
    const float top = border[co], bot = border[cout + co], lef = border[2 * cout + co], rig = border[3 * cout + co];
    const float tl = border[4 * cout + co], tr = border[5 * cout + co], bl = border[6 * cout + co], br = border[7 * cout + co];
    const float* wp = w + ((size_t)co * cin_total + cin_off + ci) * 9;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        double s = T;
        if (ky == 0) s -= top;
        if (ky == 2) s -= bot;
        if (kx == 0) s -= lef;
        if (kx == 2) s -= rig;
        if (ky == 0 && kx == 0) s += tl;
        if (ky == 0 && kx == 2) s += tr;
        if (ky == 2 && kx == 0) s += bl;
        if (ky == 2 && kx == 2) s += br;
        acc += (double)wp[ky * 3 + kx] * s;
      }
    }
  }
  acc = warp_sum(acc);
  __shared__ double red[4];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) colsum[ci] = (float)(red[0] + red[1] + red[2] + red[3]);
}

}  // namespace dfb

extern "C" int dfb_conv3x3_dgrad_colsum(const void* gy, int f32, int n, int H, int W, int cout, const float* gy_total,
                                        const float* w, int cin_total, int cin_off, int cin, float* border_ws,
                                        float* colsum, void* stream_) {
  using namespace dfb;
  cudaStream_t st = (cudaStream_t)stream_;
  if (n <= 0 || H <= 0 || W <= 0 || cout <= 0 || cout > 1024 || cin <= 0 || cin_off < 0 || cin_off + cin > cin_total) {
    set_error("dfb_conv3x3_dgrad_colsum: bad sizes");
    return DFB_ERR_ARG;
  }
  cudaMemsetAsync(border_ws, 0, sizeof(float) * 8 * (size_t)cout, st);
  const long long border_px = (long long)n * (2 * W + 2 * H);
  const int ppi = 256 / cout > 0 ? 256 / cout : 1;
  long long blocks = (border_px + ppi * 8 - 1) / (ppi * 8);
  if (blocks > sm_count() * 4) blocks = sm_count() * 4;
  if (blocks < 1) blocks = 1;
  const int threads = cout > 256 ? 1024 : 256;
  const size_t smem = sizeof(float) * 8 * (size_t)cout;
  if (f32) k_border_sums<true><<<(int)blocks, threads, smem, st>>>(gy, n, H, W, cout, border_ws);
  else k_border_sums<false><<<(int)blocks, threads, smem, st>>>(gy, n, H, W, cout, border_ws);
  k_colsum_from_borders<<<cin, 128, 0, st>>>(border_ws, gy_total, w, cout, cin_total, cin_off, colsum);
  add_launches(2);
  return check_launch("dfb_conv3x3_dgrad_colsum");
}
