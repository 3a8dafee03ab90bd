// Pillar -> point gather of the flow decoders and its backward.
//
// Reference: ConvGRUDecoder.forward_single / LinearDecoder.forward_single
// (OpenSceneFlow/src/models/basic/decoder.py:215-225 and :86-94) index two NCHW images with
// img[:, y, x] -- a channel-strided read (1 MB between channels at 512x512) -- and autograd's
// backward is index_put_(accumulate=True) with float atomics.  Here the images are NHWC, so a point
// reads two/three contiguous rows, and the backward is an atomic-free segment sum over the CSR point
// list of every pc0 pillar.
//
// h0[p] = [ img[b, y, x, 0:32] | img[B + b, y, x, 0:32] | unet[b, y, x, 0:64] ]   (deflow.py:92-94)
#include "common.cuh"
#include "../../include/deflow_b200.h"

namespace dfb {

template <bool BF>
__device__ __forceinline__ float4 load4(const void* base, size_t elem) {
  if (BF) {
    const uint2 r = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&r.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&r.y);
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  } else {
    return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem);
  }
}

template <bool BF>
__device__ __forceinline__ void store4(void* base, size_t elem, float4 v) {
  if (BF) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<unsigned*>(&a);
    r.y = *reinterpret_cast<unsigned*>(&b);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + elem) = r;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + elem) = v;
  }
}

// One warp per point; lane l owns channels 4l..4l+3 of the 128-wide row.
template <bool IN_BF, bool OUT_BF>
__global__ void __launch_bounds__(256) k_decoder_gather(const void* __restrict__ img, const void* __restrict__ unet,
                                                        int B, int HW, const int* __restrict__ counts, int F,
                                                        const int* __restrict__ pt_pillar,
                                                        const int* __restrict__ pil_pix, void* __restrict__ h0,
                                                        int n_cap) {
  const int n = min(counts[2 * F + B], n_cap);  // pc0 points = frames 0..B-1
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < n; p += warps) {
    const int pix = pil_pix[pt_pillar[p]];  // b*HW + y*W + x
    float4 v;
    if (lane < 8) v = load4<IN_BF>(img, (size_t)pix * 32 + lane * 4);
    else if (lane < 16) v = load4<IN_BF>(img, ((size_t)pix + (size_t)B * HW) * 32 + (lane - 8) * 4);
    else v = load4<IN_BF>(unet, (size_t)pix * 64 + (lane - 16) * 4);
    store4<OUT_BF>(h0, (size_t)p * 128 + lane * 4, v);
  }
}

// One warp per pc0 pillar: sum the 128-wide gradient rows of its points, write the three NHWC rows.  Pillar populations
// are heavy-tailed (median 4, a few with hundreds to thousands of points), and a warp walking a 4000-point pillar alone
// was the whole kernel's duration (0.6 ms at config 2 for 270 MB of rows): a block owns a contiguous range of pillars, its
// warps take the light ones one each, and the heavy ones (> GB_HEAVY points) are summed by all eight warps together.
constexpr int GB_HEAVY = 96;
constexpr int GB_LIST = 64;

template <bool IN_BF>
__device__ __forceinline__ void gather_rows_sum(const void* __restrict__ grad_h0, const int* __restrict__ sorted_pt, int j0, int s1,
                                                int stride, int lane, float4& a) {
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  // up to eight rows in flight (all indices first, then all rows)
  for (int j = j0; j < s1; j += stride) {
    int pt[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pt[k] = j + k < s1 ? sorted_pt[j + k] : -1;
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      v[k] = pt[k] >= 0 ? load4<IN_BF>(grad_h0, (size_t)pt[k] * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      a.x += v[k].x; a.y += v[k].y; a.z += v[k].z; a.w += v[k].w;
      b.x += v[k + 1].x; b.y += v[k + 1].y; b.z += v[k + 1].z; b.w += v[k + 1].w;
    }
  }
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}

template <bool OUT_BF>
__device__ __forceinline__ void gather_rows_write(const float4& a, int q, int pix, int lane, int B, int HW, void* __restrict__ grad_img,
                                                  void* __restrict__ grad_unet, float* __restrict__ img_rows) {
  if (lane < 16) {
    if (img_rows) *reinterpret_cast<float4*>(img_rows + (size_t)q * 64 + lane * 4) = a;
    else if (lane < 8) store4<OUT_BF>(grad_img, (size_t)pix * 32 + lane * 4, a);
    else store4<OUT_BF>(grad_img, ((size_t)pix + (size_t)B * HW) * 32 + (lane - 8) * 4, a);
  } else {
    store4<OUT_BF>(grad_unet, (size_t)pix * 64 + (lane - 16) * 4, a);
  }
}

template <bool IN_BF, bool OUT_BF>
__global__ void __launch_bounds__(256) k_decoder_gather_bwd(const void* __restrict__ grad_h0, int B, int HW,
                                                            const int* __restrict__ counts, int F,
                                                            const int* __restrict__ pil_pix,
                                                            const int* __restrict__ pil_start,
                                                            const int* __restrict__ sorted_pt,
                                                            void* __restrict__ grad_img, void* __restrict__ grad_unet,
                                                            int pil_cap, float* __restrict__ img_rows,
                                                            float* __restrict__ unet_colsum) {
  // unet_colsum != NULL: f32[64] += the per-channel sums of grad_unet (= the bias gradient of the convolution that produced
  // the UNet output) taken from the pillar sums at hand instead of a pass over the dense, mostly zero tensor.
  // img_rows != NULL (then grad_img == NULL): the 64 image channels of every pc0 pillar's sum go to the compact fp32 buffer
  // img_rows[q][64] instead of a dense zero-filled image gradient; k_gather_img_rows_add adds them into the image gradient
  // later, once the other consumers of the pseudo-image have written theirs.
  __shared__ int heavy[GB_LIST];
  __shared__ int n_heavy;
  __shared__ float4 part[8][32];
  const int M0 = min(counts[3 * F + 1 + B], pil_cap);  // pillars of the pc0 frames
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = (M0 + (int)gridDim.x - 1) / (int)gridDim.x;
  const int qa = min(M0, (int)blockIdx.x * per), qb = min(M0, qa + per);
  if (threadIdx.x == 0) n_heavy = 0;
  __syncthreads();
  float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);   // lanes 16..31: running sum of this warp's UNet-channel rows
  for (int q = qa + warp; q < qb; q += 8) {
    const int s0 = pil_start[q], s1 = pil_start[q + 1];
    if (s1 - s0 > GB_HEAVY) {
      int slot = 0;
      if (lane == 0) slot = atomicAdd(&n_heavy, 1);
      slot = __shfl_sync(0xffffffffu, slot, 0);
      if (slot < GB_LIST) {
        if (lane == 0) heavy[slot] = q;
        continue;
      }                                            // list full: this warp walks the pillar alone
    }
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    gather_rows_sum<IN_BF>(grad_h0, sorted_pt, s0, s1, 8, lane, a);
    gather_rows_write<OUT_BF>(a, q, pil_pix[q], lane, B, HW, grad_img, grad_unet, img_rows);
    tot.x += a.x; tot.y += a.y; tot.z += a.z; tot.w += a.w;
  }
  __syncthreads();
  const int nh = min(n_heavy, GB_LIST);
  for (int h = 0; h < nh; ++h) {
    const int q = heavy[h];
    const int s0 = pil_start[q], s1 = pil_start[q + 1];
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    gather_rows_sum<IN_BF>(grad_h0, sorted_pt, s0 + warp * 8, s1, 64, lane, a);   // warp w: rows s0 + 8w + 64i + [0, 8)
    part[warp][lane] = a;
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int w = 1; w < 8; ++w) { const float4 t = part[w][lane]; a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
      gather_rows_write<OUT_BF>(a, q, pil_pix[q], lane, B, HW, grad_img, grad_unet, img_rows);
      tot.x += a.x; tot.y += a.y; tot.z += a.z; tot.w += a.w;
    }
    __syncthreads();
  }
  if (unet_colsum) {   // block total of the 64 UNet channels -> one atomic per channel and block
    part[warp][lane] = tot;
    __syncthreads();
    if (warp == 0 && lane >= 16) {
      float4 t = part[0][lane];
#pragma unroll
      for (int w = 1; w < 8; ++w) { const float4 u = part[w][lane]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
      float* dst = unet_colsum + (lane - 16) * 4;
      if (qb > qa) { atomicAdd(dst, t.x); atomicAdd(dst + 1, t.y); atomicAdd(dst + 2, t.z); atomicAdd(dst + 3, t.w); }
    }
  }
}

// grad_img[b, pix, :] += img_rows[q][0:32], grad_img[B + b, pix, :] += img_rows[q][32:64] for every pc0 pillar q: every
// (frame, pixel) row belongs to exactly one pillar, so this is a plain read-modify-write of two 32-channel rows per pillar.
template <bool OUT_BF>
__global__ void __launch_bounds__(256) k_gather_img_rows_add(const float* __restrict__ img_rows, int B, int HW,
                                                             const int* __restrict__ counts, int F,
                                                             const int* __restrict__ pil_pix, void* __restrict__ grad_img,
                                                             int pil_cap) {
  const int M0 = min(counts[3 * F + 1 + B], pil_cap);
  const long long total = (long long)M0 * 16;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(t >> 4), l = (int)(t & 15);
    const int pix = pil_pix[q];
    float4 a = *reinterpret_cast<const float4*>(img_rows + (size_t)q * 64 + l * 4);
    const size_t e = l < 8 ? (size_t)pix * 32 + l * 4 : ((size_t)pix + (size_t)B * HW) * 32 + (l - 8) * 4;
    const float4 o = load4<OUT_BF>(grad_img, e);
    a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
    store4<OUT_BF>(grad_img, e, a);
  }
}

}  // namespace dfb

using namespace dfb;

extern "C" int dfb_decoder_gather(const void* img, const void* unet, int in_bf16, int B, int H, int W,
                                  const int* counts, int F, const int* pt_pillar, const int* pil_pix, void* h0,
                                  int out_bf16, int n_cap, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (B <= 0 || F < B || H <= 0 || W <= 0) { set_error("dfb_decoder_gather: bad sizes"); return DFB_ERR_ARG; }
  if (n_cap <= 0) return DFB_OK;
  long long blocks = ((long long)n_cap * 32 + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  const int g = (int)blocks, HW = H * W;
  if (in_bf16 && out_bf16) k_decoder_gather<true, true><<<g, 256, 0, st>>>(img, unet, B, HW, counts, F, pt_pillar, pil_pix, h0, n_cap);
  else if (in_bf16) k_decoder_gather<true, false><<<g, 256, 0, st>>>(img, unet, B, HW, counts, F, pt_pillar, pil_pix, h0, n_cap);
  else if (out_bf16) k_decoder_gather<false, true><<<g, 256, 0, st>>>(img, unet, B, HW, counts, F, pt_pillar, pil_pix, h0, n_cap);
  else k_decoder_gather<false, false><<<g, 256, 0, st>>>(img, unet, B, HW, counts, F, pt_pillar, pil_pix, h0, n_cap);
  add_launches(1);
  return check_launch("dfb_decoder_gather");
}

static int gather_backward_impl(const void* grad_h0, int grad_bf16, int B, int H, int W, const int* counts, int F,
                                const int* pil_pix, const int* pil_start, const int* sorted_pt, void* grad_img,
                                void* grad_unet, int out_bf16, int pil_cap, float* img_rows, float* unet_colsum, cudaStream_t st) {
  if (B <= 0 || F < B || H <= 0 || W <= 0) { set_error("dfb_decoder_gather_backward: bad sizes"); return DFB_ERR_ARG; }
  const size_t HW = (size_t)H * W, es = out_bf16 ? 2 : 4;
  // dense gradients: zero everywhere except the pc0 pillars (index_put_ accumulate into zeros)
  if (grad_img) cudaMemsetAsync(grad_img, 0, (size_t)2 * B * HW * 32 * es, st);
  cudaMemsetAsync(grad_unet, 0, (size_t)B * HW * 64 * es, st);
  if (pil_cap > 0) {
    long long blocks = ((long long)pil_cap * 32 + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    const int g = (int)blocks;
    if (grad_bf16 && out_bf16) k_decoder_gather_bwd<true, true><<<g, 256, 0, st>>>(grad_h0, B, (int)HW, counts, F, pil_pix, pil_start, sorted_pt, grad_img, grad_unet, pil_cap, img_rows, unet_colsum);
    else if (grad_bf16) k_decoder_gather_bwd<true, false><<<g, 256, 0, st>>>(grad_h0, B, (int)HW, counts, F, pil_pix, pil_start, sorted_pt, grad_img, grad_unet, pil_cap, img_rows, unet_colsum);
    else if (out_bf16) k_decoder_gather_bwd<false, true><<<g, 256, 0, st>>>(grad_h0, B, (int)HW, counts, F, pil_pix, pil_start, sorted_pt, grad_img, grad_unet, pil_cap, img_rows, unet_colsum);
    else k_decoder_gather_bwd<false, false><<<g, 256, 0, st>>>(grad_h0, B, (int)HW, counts, F, pil_pix, pil_start, sorted_pt, grad_img, grad_unet, pil_cap, img_rows, unet_colsum);
    add_launches(1);
  }
  return check_launch("dfb_decoder_gather_backward");
}

extern "C" int dfb_decoder_gather_backward(const void* grad_h0, int grad_bf16, int B, int H, int W,
                                           const int* counts, int F, const int* pil_pix, const int* pil_start,
                                           const int* sorted_pt, void* grad_img, void* grad_unet, int out_bf16,
                                           int pil_cap, void* stream_) {
  if (!grad_img || !grad_unet) { set_error("dfb_decoder_gather_backward: both outputs are required (see dfb_decoder_gather_backward_rows)"); return DFB_ERR_ARG; }
  return gather_backward_impl(grad_h0, grad_bf16, B, H, W, counts, F, pil_pix, pil_start, sorted_pt, grad_img, grad_unet, out_bf16,
                              pil_cap, nullptr, nullptr, (cudaStream_t)stream_);
}

extern "C" int dfb_decoder_gather_backward_rows(const void* grad_h0, int grad_bf16, int B, int H, int W,
                                                const int* counts, int F, const int* pil_pix, const int* pil_start,
                                                const int* sorted_pt, float* img_rows, void* grad_unet, int out_bf16,
                                                int pil_cap, float* unet_colsum, void* stream_) {
  if (!img_rows || !grad_unet) { set_error("dfb_decoder_gather_backward_rows: img_rows and grad_unet are required"); return DFB_ERR_ARG; }
  return gather_backward_impl(grad_h0, grad_bf16, B, H, W, counts, F, pil_pix, pil_start, sorted_pt, nullptr, grad_unet, out_bf16,
                              pil_cap, img_rows, unet_colsum, (cudaStream_t)stream_);
}

extern "C" int dfb_gather_img_rows_add(const float* img_rows, int B, int H, int W, const int* counts, int F, const int* pil_pix,
                                       void* grad_img, int out_bf16, int pil_cap, void* stream_) {
  if (!img_rows || !grad_img || B <= 0 || F < B) { set_error("dfb_gather_img_rows_add: bad arguments"); return DFB_ERR_ARG; }
  if (pil_cap <= 0) return DFB_OK;
  long long blocks = ((long long)pil_cap * 16 + 255) / 256;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  if (out_bf16) k_gather_img_rows_add<true><<<(int)blocks, 256, 0, (cudaStream_t)stream_>>>(img_rows, B, H * W, counts, F, pil_pix, grad_img, pil_cap);
  else k_gather_img_rows_add<false><<<(int)blocks, 256, 0, (cudaStream_t)stream_>>>(img_rows, B, H * W, counts, F, pil_pix, grad_img, pil_cap);
  add_launches(1);
  return check_launch("dfb_gather_img_rows_add");
}

// out[0 : n) = a0 + b0, out[n : 2n) = a1 + b1 (16-byte vectors): the gradient of the pseudo-image from its two consumers per
// frame half (first encoder convolution, last skip convolution), concatenated -- one pass instead of two additions and a
// concatenation.  b0 / b1 may be NULL (a consumer without gradient).
template <bool BF>
__global__ void __launch_bounds__(256) k_add_cat2(const uint4* __restrict__ a0, const uint4* __restrict__ b0,
                                                  const uint4* __restrict__ a1, const uint4* __restrict__ b1, long long n16,
                                                  uint4* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n16; i += (long long)gridDim.x * blockDim.x) {
    const bool hi = i >= n16;
    const long long j = hi ? i - n16 : i;
    const uint4* a = hi ? a1 : a0;
    const uint4* b = hi ? b1 : b0;
    uint4 x = a[j];
    if (b) {
      const uint4 y = b[j];
      if (BF) {
        __nv_bfloat162* xp = reinterpret_cast<__nv_bfloat162*>(&x);
        const __nv_bfloat162* yp = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 p = __bfloat1622float2(xp[k]), q = __bfloat1622float2(yp[k]);
          xp[k] = __floats2bfloat162_rn(p.x + q.x, p.y + q.y);
        }
      } else {
        float* xp = reinterpret_cast<float*>(&x);
        const float* yp = reinterpret_cast<const float*>(&y);
#pragma unroll
        for (int k = 0; k < 4; ++k) xp[k] += yp[k];
      }
    }
    out[i] = x;
  }
}

extern "C" int dfb_add_cat2(const void* a0, const void* b0, const void* a1, const void* b1, long long bytes_per_half, int bf16,
                            void* out, void* stream_) {
  if (!a0 || !a1 || !out || bytes_per_half <= 0 || bytes_per_half % 16) { set_error("dfb_add_cat2: bad arguments"); return DFB_ERR_ARG; }
  if ((b0 == nullptr) != (b1 == nullptr)) { set_error("dfb_add_cat2: both or neither of the second addends"); return DFB_ERR_ARG; }
  const long long n16 = bytes_per_half / 16;
  long long blocks = (2 * n16 + 255) / 256;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  if (bf16) k_add_cat2<true><<<(int)blocks, 256, 0, (cudaStream_t)stream_>>>((const uint4*)a0, (const uint4*)b0, (const uint4*)a1, (const uint4*)b1, n16, (uint4*)out);
  else k_add_cat2<false><<<(int)blocks, 256, 0, (cudaStream_t)stream_>>>((const uint4*)a0, (const uint4*)b0, (const uint4*)a1, (const uint4*)b1, n16, (uint4*)out);
  add_launches(1);
  return check_launch("dfb_add_cat2");
}
