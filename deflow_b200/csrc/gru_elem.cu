// Gate arithmetic of the per-point ConvGRU decoder around the tensor-core gate GEMMs.
//
// Reference: ConvGRU.forward (OpenSceneFlow/src/models/basic/decoder.py:184-193)
//   z = sigmoid(Wz [h,x] + bz); r = sigmoid(Wr [h,x] + br); q = tanh(Wq [r*h, x] + bq); h' = (1-z) h + z q
// and the MLP head of ConvGRUDecoder.forward_single (:236-237) flow = W2 GELU(W1 [h,x] + b1) + b2.
// The three k=1 Conv1d are [n,192] x [192,128] GEMMs and run as 1x1 tcgen05 convolutions over the point list
// (csrc/conv_igemm.cu); the kernels here are the HBM-bound elementwise stages between them, forward and
// backward.  Rows are points; every buffer has n_pad >= n rows and rows >= n are written as zeros so that they
// contribute nothing to the weight-gradient reductions.
#include "common.cuh"
#include "../../include/deflow_b200.h"

namespace dfb {

__device__ __forceinline__ void g_unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(p[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 g_pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }

// 8 consecutive values of vector index e: bf16 (one 16-byte access) or fp32 (two) -- F32 = parity mode tensors
template <bool F32>
__device__ __forceinline__ void ld8(const void* p, size_t e, float* f) {
  if (F32) {
    const float4* q = reinterpret_cast<const float4*>(p) + 2 * e;
    const float4 a = __ldg(q), b = __ldg(q + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    g_unpack8(__ldg(reinterpret_cast<const uint4*>(p) + e), f);
  }
}
template <bool F32>
__device__ __forceinline__ void st8(void* p, size_t e, const float* f) {
  if (F32) {
    float4* q = reinterpret_cast<float4*>(p) + 2 * e;
    q[0] = make_float4(f[0], f[1], f[2], f[3]);
    q[1] = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    reinterpret_cast<uint4*>(p)[e] = g_pack8(f);
  }
}

constexpr int HID = 128;  // hidden width; 16 vectors of 8 per row

// x = W_off o + b_off (Linear(3, 64|128)), written as bf16; also converts h0 fp32 -> bf16 copy.
template <bool F32>
__global__ void __launch_bounds__(256) k_offset_encode(const float* __restrict__ offs, const float* __restrict__ w,
                                                       const float* __restrict__ b, int n, int n_pad, int cx,
                                                       void* __restrict__ x) {
  const int vpr = cx >> 3;
  const long long total = (long long)n_pad * vpr;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(e / vpr), c0 = (int)(e % vpr) << 3;
    float f[8];
    if (row < n) {
      const float o0 = offs[3 * (size_t)row], o1 = offs[3 * (size_t)row + 1], o2 = offs[3 * (size_t)row + 2];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float* wr = w + 3 * (c0 + i);
        f[i] = fmaf(o2, wr[2], fmaf(o1, wr[1], fmaf(o0, wr[0], b[c0 + i])));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = 0.f;
    }
    st8<F32>(x, e, f);
  }
}

// fp32 [n, C] -> bf16 [n_pad, C] with zero padding rows
__global__ void __launch_bounds__(256) k_to_bf16_pad(const float* __restrict__ src, int n, int n_pad, int C,
                                                     uint4* __restrict__ dst) {
  const int vpr = C >> 3;
  const long long total = (long long)n_pad * vpr;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(e / vpr);
    float f[8];
    if (row < n) {
      const float4 a = *reinterpret_cast<const float4*>(src + e * 8), c = *reinterpret_cast<const float4*>(src + e * 8 + 4);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = c.x; f[5] = c.y; f[6] = c.z; f[7] = c.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = 0.f;
    }
    dst[e] = g_pack8(f);
  }
}

// rh = bf16(sigmoid(r_pre) * h);  zr_pre [n_pad, 256] bf16 (z | r), h fp32 [n_pad, 128]
template <bool F32>
__global__ void __launch_bounds__(256) k_gru_rh(const void* __restrict__ zr_pre, const float* __restrict__ h, int n,
                                                int n_pad, void* __restrict__ rh) {
  const long long total = (long long)n_pad * 16;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(e >> 4), v = (int)(e & 15);
    float f[8];
    if (row < n) {
      float r[8];
      ld8<F32>(zr_pre, (size_t)row * 32 + 16 + v, r);
      const float4 a = *reinterpret_cast<const float4*>(h + e * 8), c = *reinterpret_cast<const float4*>(h + e * 8 + 4);
      const float hv[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = sigmoid_f(r[i]) * hv[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = 0.f;
    }
    st8<F32>(rh, e, f);
  }
}

// h' = (1 - z) h + z q;  writes fp32 h' and its bf16 copy
template <bool F32>
__global__ void __launch_bounds__(256) k_gru_update(const void* __restrict__ zr_pre, const void* __restrict__ q_pre,
                                                    const float* __restrict__ h, int n, int n_pad,
                                                    float* __restrict__ h_new, uint4* __restrict__ hb_new) {
  const long long total = (long long)n_pad * 16;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(e >> 4), v = (int)(e & 15);
    float f[8];
    if (row < n) {
      float z[8], q[8];
      ld8<F32>(zr_pre, (size_t)row * 32 + v, z);
      ld8<F32>(q_pre, (size_t)e, q);
      const float4 a = *reinterpret_cast<const float4*>(h + e * 8), c = *reinterpret_cast<const float4*>(h + e * 8 + 4);
      const float hv[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float zz = sigmoid_f(z[i]);
        f[i] = (1.0f - zz) * hv[i] + zz * tanhf(q[i]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = 0.f;
    }
    *reinterpret_cast<float4*>(h_new + e * 8) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(h_new + e * 8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
    if (hb_new) hb_new[e] = g_pack8(f);
  }
}

// Backward stage 1 (given dh' = dL/dh'): dq_pre = dh' z (1 - q^2) [bf16], dz_pre = dh' (q - h) z (1 - z) [bf16, into
// dzr_pre[:, 0:128]], dh_acc = dh' (1 - z) [fp32].
template <bool F32>
__global__ void __launch_bounds__(256) k_gru_bwd1(const void* __restrict__ zr_pre, const void* __restrict__ q_pre,
                                                  const float* __restrict__ h, const float* __restrict__ dh_new, int n,
                                                  int n_pad, void* __restrict__ dq_pre, void* __restrict__ dzr_pre,
                                                  float* __restrict__ dh_acc) {
  const long long total = (long long)n_pad * 16;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(e >> 4), v = (int)(e & 15);
    float dq[8], dz[8], da[8];
    if (row < n) {
      float z[8], q[8];
      ld8<F32>(zr_pre, (size_t)row * 32 + v, z);
      ld8<F32>(q_pre, (size_t)e, q);
      const float4 a = *reinterpret_cast<const float4*>(h + e * 8), c = *reinterpret_cast<const float4*>(h + e * 8 + 4);
      const float hv[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
      const float4 g0 = *reinterpret_cast<const float4*>(dh_new + e * 8), g1 = *reinterpret_cast<const float4*>(dh_new + e * 8 + 4);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float zz = sigmoid_f(z[i]), qq = tanhf(q[i]);
        dq[i] = g[i] * zz * (1.0f - qq * qq);
        dz[i] = g[i] * (qq - hv[i]) * zz * (1.0f - zz);
        da[i] = g[i] * (1.0f - zz);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) { dq[i] = 0.f; dz[i] = 0.f; da[i] = 0.f; }
    }
    st8<F32>(dq_pre, (size_t)e, dq);
    st8<F32>(dzr_pre, (size_t)row * 32 + v, dz);
    *reinterpret_cast<float4*>(dh_acc + e * 8) = make_float4(da[0], da[1], da[2], da[3]);
    *reinterpret_cast<float4*>(dh_acc + e * 8 + 4) = make_float4(da[4], da[5], da[6], da[7]);
  }
}

// Backward stage 2 (given d_rh = dL/d(r*h) from the Wq data gradient, bf16): dr_pre = d_rh h r (1 - r) [into
// dzr_pre[:, 128:256]], dh_acc += d_rh r.
template <bool F32>
__global__ void __launch_bounds__(256) k_gru_bwd2(const void* __restrict__ zr_pre, const float* __restrict__ h,
                                                  const void* __restrict__ d_rh, int n, int n_pad,
                                                  void* __restrict__ dzr_pre, float* __restrict__ dh_acc) {
  const long long total = (long long)n_pad * 16;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(e >> 4), v = (int)(e & 15);
    float dr[8];
    if (row < n) {
      float r[8], g[8];
      ld8<F32>(zr_pre, (size_t)row * 32 + 16 + v, r);
      ld8<F32>(d_rh, (size_t)e, g);
      const float4 a = *reinterpret_cast<const float4*>(h + e * 8), c = *reinterpret_cast<const float4*>(h + e * 8 + 4);
      const float hv[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
      float4 d0 = *reinterpret_cast<float4*>(dh_acc + e * 8), d1 = *reinterpret_cast<float4*>(dh_acc + e * 8 + 4);
      float da[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float rr = sigmoid_f(r[i]);
        dr[i] = g[i] * hv[i] * rr * (1.0f - rr);
        da[i] = fmaf(g[i], rr, da[i]);
      }
      *reinterpret_cast<float4*>(dh_acc + e * 8) = make_float4(da[0], da[1], da[2], da[3]);
      *reinterpret_cast<float4*>(dh_acc + e * 8 + 4) = make_float4(da[4], da[5], da[6], da[7]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) dr[i] = 0.f;
    }
    st8<F32>(dzr_pre, (size_t)row * 32 + 16 + v, dr);
  }
}

// acc fp32 [n_pad, C] += bf16 a (+ bf16 b)
template <bool F32>
__global__ void __launch_bounds__(256) k_acc_bf16(float* __restrict__ acc, const void* __restrict__ a,
                                                  const void* __restrict__ b, long long n_vec) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_vec; e += (long long)gridDim.x * blockDim.x) {
    float fa[8], fb[8];
    ld8<F32>(a, (size_t)e, fa);
    if (b) ld8<F32>(b, (size_t)e, fb);
    float4 d0 = *reinterpret_cast<float4*>(acc + e * 8), d1 = *reinterpret_cast<float4*>(acc + e * 8 + 4);
    float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] += fa[i] + (b ? fb[i] : 0.f);
    *reinterpret_cast<float4*>(acc + e * 8) = make_float4(d[0], d[1], d[2], d[3]);
    *reinterpret_cast<float4*>(acc + e * 8 + 4) = make_float4(d[4], d[5], d[6], d[7]);
  }
}

// MLP head tail: flow = W2 GELU(y1) + b2 with y1 [n_pad, 32] bf16 (pre-activation, bias included)
template <bool F32>
__global__ void __launch_bounds__(256) k_head_out(const void* __restrict__ y1, const float* __restrict__ w2,
                                                  const float* __restrict__ b2, int n, float* __restrict__ flow) {
  __shared__ float sw[96], sb[3];
  if (threadIdx.x < 96) sw[threadIdx.x] = w2[threadIdx.x];
  if (threadIdx.x < 3) sb[threadIdx.x] = b2[threadIdx.x];
  __syncthreads();
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += gridDim.x * blockDim.x) {
    float o0 = sb[0], o1 = sb[1], o2 = sb[2];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      float f[8];
      ld8<F32>(y1, (size_t)row * 4 + v, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float z = f[i];
        const float a = 0.5f * z * (1.0f + erff(z * 0.70710678118654752f));
        const int c = v * 8 + i;
        o0 = fmaf(a, sw[c], o0); o1 = fmaf(a, sw[32 + c], o1); o2 = fmaf(a, sw[64 + c], o2);
      }
    }
    flow[3 * (size_t)row] = o0; flow[3 * (size_t)row + 1] = o1; flow[3 * (size_t)row + 2] = o2;
  }
}

// backward: dy1 = (W2^T dflow) GELU'(y1) [bf16, zero rows >= n]; gW2 [3,32] += dflow^T GELU(y1); gb2 += sum dflow
template <bool F32>
__global__ void __launch_bounds__(256) k_head_out_bwd(const void* __restrict__ y1, const float* __restrict__ w2,
                                                      const float* __restrict__ dflow, int n, int n_pad,
                                                      void* __restrict__ dy1, float* __restrict__ gw2,
                                                      float* __restrict__ gb2) {
  __shared__ float sw[96];
  __shared__ float acc[99];
  if (threadIdx.x < 96) sw[threadIdx.x] = w2[threadIdx.x];
  if (threadIdx.x < 99) acc[threadIdx.x] = 0.f;
  __syncthreads();
  float lw[96];
  float lb0 = 0.f, lb1 = 0.f, lb2 = 0.f;
#pragma unroll
  for (int i = 0; i < 96; ++i) lw[i] = 0.f;
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n_pad; row += gridDim.x * blockDim.x) {
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
    const bool live = row < n;
    if (live) { d0 = dflow[3 * (size_t)row]; d1 = dflow[3 * (size_t)row + 1]; d2 = dflow[3 * (size_t)row + 2]; }
    lb0 += d0; lb1 += d1; lb2 += d2;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      float f[8], o[8];
      ld8<F32>(y1, (size_t)row * 4 + v, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float z = live ? f[i] : 0.f;
        const float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752f));
        const float a = z * cdf;
        const float ga = cdf + z * 0.3989422804014327f * expf(-0.5f * z * z);
        const int c = v * 8 + i;
        o[i] = (d0 * sw[c] + d1 * sw[32 + c] + d2 * sw[64 + c]) * ga;
        lw[c] = fmaf(d0, a, lw[c]); lw[32 + c] = fmaf(d1, a, lw[32 + c]); lw[64 + c] = fmaf(d2, a, lw[64 + c]);
      }
      st8<F32>(dy1, (size_t)row * 4 + v, o);
    }
  }
#pragma unroll
  for (int i = 0; i < 96; ++i) {
    const float s = warp_sum(lw[i]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&acc[i], s);
  }
  lb0 = warp_sum(lb0); lb1 = warp_sum(lb1); lb2 = warp_sum(lb2);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&acc[96], lb0); atomicAdd(&acc[97], lb1); atomicAdd(&acc[98], lb2); }
  __syncthreads();
  if (threadIdx.x < 96) atomicAdd(&gw2[threadIdx.x], acc[threadIdx.x]);
  else if (threadIdx.x < 99) atomicAdd(&gb2[threadIdx.x - 96], acc[threadIdx.x]);
}

// offset encoder backward: gW [cx,3] += dx^T o, gb [cx] += sum dx, with dx fp32 [n_pad, cx] (rows >= n are zero)
__global__ void __launch_bounds__(256) k_offset_encode_bwd(const float* __restrict__ dx, const float* __restrict__ offs,
                                                           int n, int cx, float* __restrict__ gw, float* __restrict__ gb) {
  __shared__ float acc[128 * 4];
  for (int i = threadIdx.x; i < cx * 4; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int c = threadIdx.x % cx, prow = threadIdx.x / cx, prows = blockDim.x / cx;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (long long row = (long long)blockIdx.x * prows + prow; row < n; row += (long long)gridDim.x * prows) {
    const float g = dx[row * cx + c];
    s0 = fmaf(g, offs[3 * row], s0); s1 = fmaf(g, offs[3 * row + 1], s1); s2 = fmaf(g, offs[3 * row + 2], s2);
    s3 += g;
  }
  atomicAdd(&acc[c * 4], s0); atomicAdd(&acc[c * 4 + 1], s1); atomicAdd(&acc[c * 4 + 2], s2); atomicAdd(&acc[c * 4 + 3], s3);
  __syncthreads();
  for (int i = threadIdx.x; i < cx; i += blockDim.x) {
    atomicAdd(&gw[3 * i], acc[i * 4]); atomicAdd(&gw[3 * i + 1], acc[i * 4 + 1]); atomicAdd(&gw[3 * i + 2], acc[i * 4 + 2]);
    atomicAdd(&gb[i], acc[i * 4 + 3]);
  }
}

static int gridv(long long work, int block = 256, int mult = 16) {
  long long b = (work + block - 1) / block;
  const long long cap = (long long)sm_count() * mult;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace dfb

using namespace dfb;
#define ST ((cudaStream_t)stream_)
#define DFB_F32_DISPATCH(kernel, grid, ...)                          \
  do {                                                               \
    if (f32) kernel<true><<<grid, 256, 0, ST>>>(__VA_ARGS__);        \
    else kernel<false><<<grid, 256, 0, ST>>>(__VA_ARGS__);           \
  } while (0)

extern "C" int dfb_offset_encode(const float* offs, const float* w, const float* b, int n, int n_pad, int cx, void* x,
                                 int f32, void* stream_) {
  if (cx % 8 || n_pad < n) { set_error("dfb_offset_encode: bad sizes"); return DFB_ERR_ARG; }
  DFB_F32_DISPATCH(k_offset_encode, gridv((long long)n_pad * (cx >> 3)), offs, w, b, n, n_pad, cx, x);
  add_launches(1);
  return check_launch("dfb_offset_encode");
}
extern "C" int dfb_offset_encode_backward(const float* dx, const float* offs, int n, int cx, float* gw, float* gb,
                                          void* stream_) {
  if (cx != 64 && cx != 128) { set_error("dfb_offset_encode_backward: width must be 64 or 128"); return DFB_ERR_ARG; }
  k_offset_encode_bwd<<<gridv((long long)n, 256 / cx, 4), 256, 0, ST>>>(dx, offs, n, cx, gw, gb);
  add_launches(1);
  return check_launch("dfb_offset_encode_backward");
}
extern "C" int dfb_to_bf16_pad(const float* src, int n, int n_pad, int C, void* dst, void* stream_) {
  if (C % 8 || n_pad < n) { set_error("dfb_to_bf16_pad: bad sizes"); return DFB_ERR_ARG; }
  k_to_bf16_pad<<<gridv((long long)n_pad * (C >> 3)), 256, 0, ST>>>(src, n, n_pad, C, (uint4*)dst);
  add_launches(1);
  return check_launch("dfb_to_bf16_pad");
}
extern "C" int dfb_gru_rh(const void* zr_pre, const float* h, int n, int n_pad, void* rh, int f32, void* stream_) {
  DFB_F32_DISPATCH(k_gru_rh, gridv((long long)n_pad * 16), zr_pre, h, n, n_pad, rh);
  add_launches(1);
  return check_launch("dfb_gru_rh");
}
extern "C" int dfb_gru_update(const void* zr_pre, const void* q_pre, const float* h, int n, int n_pad, float* h_new,
                              void* hb_new, int f32, void* stream_) {
  DFB_F32_DISPATCH(k_gru_update, gridv((long long)n_pad * 16), zr_pre, q_pre, h, n, n_pad, h_new, (uint4*)hb_new);
  add_launches(1);
  return check_launch("dfb_gru_update");
}
extern "C" int dfb_gru_bwd1(const void* zr_pre, const void* q_pre, const float* h, const float* dh_new, int n, int n_pad,
                            void* dq_pre, void* dzr_pre, float* dh_acc, int f32, void* stream_) {
  DFB_F32_DISPATCH(k_gru_bwd1, gridv((long long)n_pad * 16), zr_pre, q_pre, h, dh_new, n, n_pad, dq_pre, dzr_pre, dh_acc);
  add_launches(1);
  return check_launch("dfb_gru_bwd1");
}
extern "C" int dfb_gru_bwd2(const void* zr_pre, const float* h, const void* d_rh, int n, int n_pad, void* dzr_pre,
                            float* dh_acc, int f32, void* stream_) {
  DFB_F32_DISPATCH(k_gru_bwd2, gridv((long long)n_pad * 16), zr_pre, h, d_rh, n, n_pad, dzr_pre, dh_acc);
  add_launches(1);
  return check_launch("dfb_gru_bwd2");
}
extern "C" int dfb_acc_bf16(float* acc, const void* a, const void* b, long long n_elems, int f32, void* stream_) {
  if (n_elems % 8) { set_error("dfb_acc_bf16: element count must be a multiple of 8"); return DFB_ERR_ARG; }
  DFB_F32_DISPATCH(k_acc_bf16, gridv(n_elems / 8), acc, a, b, n_elems / 8);
  add_launches(1);
  return check_launch("dfb_acc_bf16");
}
extern "C" int dfb_head_out(const void* y1, const float* w2, const float* b2, int n, float* flow, int f32, void* stream_) {
  if (n <= 0) return DFB_OK;
  DFB_F32_DISPATCH(k_head_out, gridv(n), y1, w2, b2, n, flow);
  add_launches(1);
  return check_launch("dfb_head_out");
}
extern "C" int dfb_head_out_backward(const void* y1, const float* w2, const float* dflow, int n, int n_pad, void* dy1,
                                     float* gw2, float* gb2, int f32, void* stream_) {
  DFB_F32_DISPATCH(k_head_out_bwd, gridv(n_pad, 256, 2), y1, w2, dflow, n, n_pad, dy1, gw2, gb2);
  add_launches(1);
  return check_launch("dfb_head_out_backward");
}
