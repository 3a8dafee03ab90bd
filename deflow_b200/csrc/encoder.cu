// Fused pillar feature net (forward + backward) for all frames of a step.
//
// Reference path: DynamicPillarFeatureNet.forward (OpenSceneFlow/src/models/basic/encoder.py:430-475)
//   cluster_scatter (mean xyz per pillar)  -> map_voxel_center_to_point -> decorations [N,9]
//   -> Linear(9,32,no bias) -> BatchNorm1d(32, eps 1e-3, momentum 0.01) -> ReLU -> pfn_scatter (mean)
// followed by PointPillarsScatter.forward_single (encoder.py:126-147).
//
// Here the [N,9] and [N,32] per-point tensors are never written to HBM: every pass recomputes the
// 9 decorations from the 12-byte point and the pillar mean/centre.  Passes:
//   pillar_mean  (CSR segment sum, 8 lanes per pillar)
//   moments      (per-frame first / second moments of the 9 decorated features -> BatchNorm batch statistics)
//   bn_finalize  (scale/shift per frame + the 2B sequential running-stat updates)
//   pfn_points   (thread per point: Linear+BN+ReLU; warp per 32 points: pillar means straight into the NHWC image)
//   straddlers   (pillars that cross a 32-point group boundary)
// Backward: one pillar-centric pass over the saved ReLU masks + a finalize block.
#include "common.cuh"
#include <stdlib.h>
#include "../../include/deflow_b200.h"

namespace dfb {

constexpr int PFN_C = 32;
constexpr int PFN_K = 9;

__constant__ float c_pfn_w[PFN_C * PFN_K];   // Linear(9,32,bias=False) weight [c][k] of the running forward (k_pfn_points)

struct PfnGeom {
  float vx, vy, vz;
  float ox, oy, oz;  // centre offsets v/2 + range_min (fp32-rounded python doubles)
};

// decorations of encoder.py:446-465; every product / sum individually rounded like the torch ops
__device__ __forceinline__ void decorate(float x, float y, float z, float mx, float my, float mz, int cz, int cy,
                                         int cx, const PfnGeom& G, float* f) {
  f[0] = x; f[1] = y; f[2] = z;
  f[3] = __fsub_rn(x, mx); f[4] = __fsub_rn(y, my); f[5] = __fsub_rn(z, mz);
  f[6] = __fsub_rn(x, __fadd_rn(__fmul_rn((float)cx, G.vx), G.ox));
  f[7] = __fsub_rn(y, __fadd_rn(__fmul_rn((float)cy, G.vy), G.oy));
  f[8] = __fsub_rn(z, __fadd_rn(__fmul_rn((float)cz, G.vz), G.oz));
}

// All passes stream through csr_rec: (x, y, z, pillar id) of every valid point in CSR (pillar-sorted) order, written
// by the pillar index (pillar_index.cu:k_fill_csr).  One coalesced 16-byte load per point, no indirection.
__device__ __forceinline__ int rec_q(const float4& r) { return __float_as_int(r.w); }

// ---------------------------------------------------------------- pillar mean (cluster_scatter)
// Also writes the 48-byte pillar header every later pass reads instead of six scattered arrays:
//   h[0] = (first CSR position, point count, pixel index, frame)   (ints)
//   h[1] = (mean x, mean y, mean z, centre x)                       h[2] = (centre y, centre z, -, -)
// with centre = c * voxel + (voxel / 2 + range_min), rounded exactly like encoder.py:452-457.
// Four lanes per pillar (median pillar: 4 points), two pillars per lane group and iteration with every first-round
// load of both issued before the first use -- the pass is latency-bound, not bandwidth-bound.  The four lanes split the
// header stores between them.
__global__ void __launch_bounds__(256) k_pillar_mean(const int* __restrict__ counts, int F, int HW, PfnGeom G,
                                                     const float4* __restrict__ rec,
                                                     const int* __restrict__ pil_start, const int* __restrict__ pil_pix,
                                                     const int* __restrict__ pil_coor, float* __restrict__ pil_mean,
                                                     float4* __restrict__ pil_hdr) {
  const int M = counts[3 * F + 1 + F];
  const int Mr = (M + 7) & ~7;                    // all 8 lane groups of a warp run the same number of iterations
  const int sub = threadIdx.x & 3;
  const int groups = (gridDim.x * blockDim.x) >> 2;
  constexpr int U = 2;
  for (int q0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 2; q0 < Mr; q0 += U * groups) {
    int s0[U], s1[U], pix[U], cz[U], cy[U], cx[U];
    float4 r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int q = q0 + u * groups;
      s0[u] = 0; s1[u] = 0; pix[u] = 0; cz[u] = cy[u] = cx[u] = 0;
      if (q < M) {
        s0[u] = pil_start[q]; s1[u] = pil_start[q + 1];
        pix[u] = pil_pix[q];
        cz[u] = pil_coor[3 * (size_t)q]; cy[u] = pil_coor[3 * (size_t)q + 1]; cx[u] = pil_coor[3 * (size_t)q + 2];
      }
    }
    // pillars with more than 64 points (the heavy tail: up to ~750 on real sweeps) are summed by the whole warp below;
    // left to a 4-lane group they would set the duration of the kernel
    bool heavy[U];
#pragma unroll
    for (int u = 0; u < U; ++u) heavy[u] = s1[u] - s0[u] > 64;
#pragma unroll
    for (int u = 0; u < U; ++u)
      r[u] = (!heavy[u] && s0[u] + sub < s1[u]) ? __ldg(rec + s0[u] + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int q = q0 + u * groups;
      float sx = r[u].x, sy = r[u].y, sz = r[u].z;
      if (!heavy[u]) {
        for (int j = s0[u] + sub + 4; j < s1[u]; j += 4) {
          const float4 t = __ldg(rec + j);
          sx += t.x; sy += t.y; sz += t.z;
        }
      }
#pragma unroll
      for (int o = 2; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
      }
      unsigned hm = __ballot_sync(0xffffffffu, heavy[u] && sub == 0);
      while (hm) {                                   // warp-uniform
        const int src = __ffs(hm) - 1;
        hm &= hm - 1;
        const int hs0 = __shfl_sync(0xffffffffu, s0[u], src), hs1 = __shfl_sync(0xffffffffu, s1[u], src);
        float ax = 0.f, ay = 0.f, az = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
        const int lane = threadIdx.x & 31;
        int j = hs0 + lane;
        for (; j + 32 < hs1; j += 64) {
          const float4 t0 = __ldg(rec + j), t1 = __ldg(rec + j + 32);
          ax += t0.x; ay += t0.y; az += t0.z; bx += t1.x; by += t1.y; bz += t1.z;
        }
        if (j < hs1) { const float4 t0 = __ldg(rec + j); ax += t0.x; ay += t0.y; az += t0.z; }
        ax = warp_sum(ax + bx); ay = warp_sum(ay + by); az = warp_sum(az + bz);
        if ((lane & ~3) == src) { sx = ax; sy = ay; sz = az; }
      }
      if (q < M) {
        const float n = (float)(s1[u] - s0[u]);  // reduced_feats /= count.to(float) (scatter_points_cuda.cu:59-60)
        const float mx = __fdiv_rn(sx, n), my = __fdiv_rn(sy, n), mz = __fdiv_rn(sz, n);
        const float ox = __fadd_rn(__fmul_rn((float)cx[u], G.vx), G.ox), oy = __fadd_rn(__fmul_rn((float)cy[u], G.vy), G.oy),
                    oz = __fadd_rn(__fmul_rn((float)cz[u], G.vz), G.oz);
        float4* h = pil_hdr + 3 * (size_t)q;
        if (sub == 0) h[0] = make_float4(__int_as_float(s0[u]), __int_as_float(s1[u] - s0[u]), __int_as_float(pix[u]),
                                         __int_as_float(pix[u] / HW));
        else if (sub == 1) h[1] = make_float4(mx, my, mz, ox);
        else if (sub == 2) h[2] = make_float4(oy, oz, 0.f, 0.f);
        else { pil_mean[3 * (size_t)q] = mx; pil_mean[3 * (size_t)q + 1] = my; pil_mean[3 * (size_t)q + 2] = mz; }
      }
    }
  }
}

// decorations of encoder.py:446-465 from the pillar header; every sum individually rounded like the torch ops
__device__ __forceinline__ void decorate_hdr(const float4& r, const float4& h1, const float4& h2, float* f) {
  f[0] = r.x; f[1] = r.y; f[2] = r.z;
  f[3] = __fsub_rn(r.x, h1.x); f[4] = __fsub_rn(r.y, h1.y); f[5] = __fsub_rn(r.z, h1.z);
  f[6] = __fsub_rn(r.x, h1.w); f[7] = __fsub_rn(r.y, h2.x); f[8] = __fsub_rn(r.z, h2.y);
}

// ---------------------------------------------------------------- BatchNorm batch statistics from moments
// y = W f is linear in the 9 decorated features, so the per-frame mean / variance of all 32 channels follow from
// S1 = sum f (9) and S2 = sum f f^T (45 upper-triangular entries): one light pass, 54 accumulators per thread,
// instead of two passes over the 32 outputs.  mom layout [F][64] (double): S1[9] | S2[45] (row-major upper triangle).
constexpr int PFN_MOM = 54;
constexpr int PFN_MOM_PITCH = 64;

__global__ void __launch_bounds__(256) k_pfn_moments(const int* __restrict__ counts, int F,
                                                     const float4* __restrict__ rec, const float4* __restrict__ pil_hdr,
                                                     double* __restrict__ mom) {
  __shared__ float red[8][PFN_MOM];
  const int f = blockIdx.y;
  const int p0 = counts[2 * F + f], p1 = counts[2 * F + f + 1];
  float acc[PFN_MOM];
#pragma unroll
  for (int i = 0; i < PFN_MOM; ++i) acc[i] = 0.f;
  for (int j = p0 + blockIdx.x * blockDim.x + threadIdx.x; j < p1; j += gridDim.x * blockDim.x) {
    const float4 r = __ldg(rec + j);
    const int q = rec_q(r);
    float fe[PFN_K];
    decorate_hdr(r, __ldg(pil_hdr + 3 * (size_t)q + 1), __ldg(pil_hdr + 3 * (size_t)q + 2), fe);
    int o = PFN_K;
#pragma unroll
    for (int jj = 0; jj < PFN_K; ++jj) {
      acc[jj] += fe[jj];
#pragma unroll
      for (int k = jj; k < PFN_K; ++k) { acc[o] = fmaf(fe[jj], fe[k], acc[o]); ++o; }
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < PFN_MOM; ++i) {
    const float s = warp_sum(acc[i]);
    if (lane == 0) red[w][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < PFN_MOM) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += (double)red[i][threadIdx.x];
    if (s != 0.0) atomicAdd(&mom[(size_t)f * PFN_MOM_PITCH + threadIdx.x], s);
  }
}

// One block, thread = (frame slot, channel): per-frame mean / variance of W f from the moments (double) and the
// scale/shift rows, 32 frames at a time; then one warp applies the sequential running-statistics updates in the
// reference's call order pc0[0..B-1], pc1[0..B-1] (encoder.py:624-627, DeFlow.forward deflow.py:82-83).
__global__ void __launch_bounds__(1024) k_bn_finalize(const int* __restrict__ counts, int F, int training, float eps,
                                                      float momentum, const double* __restrict__ mom,
                                                      const float* __restrict__ weight, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, float* __restrict__ running_mean,
                                                      float* __restrict__ running_var, float* __restrict__ bn_params,
                                                      const int* __restrict__ sync_counts) {
  __shared__ float s_mean[32][PFN_C], s_unb[32][PFN_C];
  __shared__ int s_n[32];
  const int c = threadIdx.x & 31, slot = threadIdx.x >> 5;
  float rm = running_mean ? running_mean[c] : 0.f, rv = running_var ? running_var[c] : 1.f;
  const float rm0 = rm, rv0 = rv;
  const float g = gamma[c], b = beta[c];
  double w[PFN_K];
#pragma unroll
  for (int k = 0; k < PFN_K; ++k) w[k] = (double)weight[c * PFN_K + k];
  for (int f0 = 0; f0 < F; f0 += 32) {
    const int f = f0 + slot;
    if (f < F) {
      // SyncBatchNorm: `mom` then holds the moments summed over all ranks and sync_counts the pooled point counts
      const int n = sync_counts ? sync_counts[f] : counts[f];
      float mean = rm0, var = rv0;
      float unb = 0.f;
      if (training && n > 0) {
        const double* m = mom + (size_t)f * PFN_MOM_PITCH;
        double s1 = 0.0, s2 = 0.0;
        int o = PFN_K;
#pragma unroll
        for (int j = 0; j < PFN_K; ++j) {
          s1 += w[j] * m[j];
#pragma unroll
          for (int k = j; k < PFN_K; ++k) s2 += (j == k ? 1.0 : 2.0) * w[j] * w[k] * m[o++];
        }
        const double mu = s1 / n;
        double ssq = s2 - s1 * mu;  // sum (y - mean)^2
        if (ssq < 0.0) ssq = 0.0;
        mean = (float)mu;
        var = (float)(ssq / n);  // biased, used to normalise
        unb = n > 1 ? (float)(ssq / (n - 1)) : 0.f;
      }
      const float rstd = 1.0f / sqrtf(var + eps);
      const float a = g * rstd;
      float* o = bn_params + (size_t)f * 4 * PFN_C;
      o[c] = a;
      o[PFN_C + c] = b - mean * a;
      o[2 * PFN_C + c] = mean;
      o[3 * PFN_C + c] = rstd;
      s_mean[slot][c] = mean; s_unb[slot][c] = unb;
      if (c == 0) s_n[slot] = n;
    }
    __syncthreads();
    if (slot == 0 && training) {
      const int lim = min(32, F - f0);
      for (int i = 0; i < lim; ++i) {
        if (s_n[i] > 1) {
          rm = (1.f - momentum) * rm + momentum * s_mean[i][c];
          rv = (1.f - momentum) * rv + momentum * s_unb[i][c];
        }
      }
    }
    __syncthreads();
  }
  if (slot == 0 && training && running_mean) { running_mean[c] = rm; running_var[c] = rv; }
}

// ---------------------------------------------------------------- point pass (forward)
// A warp owns a group of 32 consecutive CSR positions, lane = point, no shared-memory staging and no block barriers:
//   1. all 32 channels of Linear + BN + ReLU with packed fp32x2 FMAs (weights broadcast from shared memory); the 32 ReLU
//      decisions of the point are saved as one mask word for the backward;
//   2. a segmented warp scan (segments = pillars; their boundaries come from two ballots) leaves, in the lane of a
//      pillar's LAST point, the 32 channel sums of that pillar in registers;
//   3. that lane finishes the pillar on the spot -- mean, voxel feature row, NHWC image row (one 64 / 128-byte row
//      store) -- with no atomics and no later pass.
// Pillars that straddle group boundaries leave per-group partial rows (part[G][0] = the leading segment that began in
// an earlier group, part[G][1] = the trailing segment that continues into the next one); k_pfn_straddlers adds them up
// in order.  Deterministic.  (The mean is sum * (1 / count) with a correctly rounded reciprocal: <= 1 ulp from the
// reference's division.)
template <bool BF16>
__device__ __forceinline__ void store_pillar(float v, int q, int pix, int c, float* __restrict__ pil_feats,
                                             void* __restrict__ image) {
  if (pil_feats) pil_feats[(size_t)q * PFN_C + c] = v;
  if (BF16) reinterpret_cast<__nv_bfloat16*>(image)[(size_t)pix * PFN_C + c] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(image)[(size_t)pix * PFN_C + c] = v;
}

__device__ __forceinline__ void store_row32(float* __restrict__ dst, const float (&v)[PFN_C]) {
#pragma unroll
  for (int c4 = 0; c4 < PFN_C; c4 += 4)
    *reinterpret_cast<float4*>(dst + c4) = make_float4(v[c4], v[c4 + 1], v[c4 + 2], v[c4 + 3]);
}

template <bool BF16>
__global__ void __launch_bounds__(256, 3) k_pfn_points(const int* __restrict__ counts, int F,
                                                       const float4* __restrict__ rec, const float4* __restrict__ pil_hdr,
                                                       const float* __restrict__ bn_params,
                                                       unsigned* __restrict__ pt_mask, float* __restrict__ part,
                                                       float* __restrict__ pil_feats, void* __restrict__ image) {
  extern __shared__ __align__(16) float sAB[];              // [F][2][32] scale / shift per frame
  const int n = counts[2 * F + F];
  for (int i = threadIdx.x; i < F * 2 * PFN_C; i += blockDim.x) {
    const int f = i / (2 * PFN_C), r = i % (2 * PFN_C);
    sAB[i] = bn_params[(size_t)f * 4 * PFN_C + r];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int n_groups = (n + 31) >> 5;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; grp < n_groups; grp += warps) {
    const int j = grp * 32 + lane;
    const bool live = j < n;
    int q = 0, pix = 0, cnt = 1;
    bool first = false, last = false;
    float v[PFN_C];
#pragma unroll
    for (int c = 0; c < PFN_C; ++c) v[c] = 0.f;
    if (live) {
      const float4 r = __ldg(rec + j);
      q = rec_q(r);
      const float4 h0 = __ldg(pil_hdr + 3 * (size_t)q), h1 = __ldg(pil_hdr + 3 * (size_t)q + 1),
                   h2 = __ldg(pil_hdr + 3 * (size_t)q + 2);
      const int s0 = __float_as_int(h0.x);
      cnt = __float_as_int(h0.y);
      pix = __float_as_int(h0.z);
      first = j == s0;
      last = j + 1 == s0 + cnt;
      float fe[PFN_K];
      decorate_hdr(r, h1, h2, fe);
      // weights are FFMA constant-bank operands (c_pfn_w[c][k], the torch layout): no load instructions at all
      float2 y2[PFN_C / 2];
#pragma unroll
      for (int c = 0; c < PFN_C / 2; ++c) {
        float y0 = 0.f, y1 = 0.f;
#pragma unroll
        for (int k = 0; k < PFN_K; ++k) {
          y0 = fmaf(fe[k], c_pfn_w[(2 * c) * PFN_K + k], y0);
          y1 = fmaf(fe[k], c_pfn_w[(2 * c + 1) * PFN_K + k], y1);
        }
        y2[c] = make_float2(y0, y1);
      }
      const float* ab = sAB + (size_t)__float_as_int(h0.w) * 2 * PFN_C;
      unsigned m = 0u;
#pragma unroll
      for (int c4 = 0; c4 < PFN_C; c4 += 4) {
        const float4 a = *reinterpret_cast<const float4*>(ab + c4), b = *reinterpret_cast<const float4*>(ab + PFN_C + c4);
        const float2 v0 = __ffma2_rn(y2[c4 / 2], make_float2(a.x, a.y), make_float2(b.x, b.y));
        const float2 v1 = __ffma2_rn(y2[c4 / 2 + 1], make_float2(a.z, a.w), make_float2(b.z, b.w));
        v[c4] = fmaxf(v0.x, 0.f); v[c4 + 1] = fmaxf(v0.y, 0.f); v[c4 + 2] = fmaxf(v1.x, 0.f); v[c4 + 3] = fmaxf(v1.y, 0.f);
        // relu(x) > 0  <=>  x > 0; the bits of a non-negative float are a positive int exactly when it is > 0
        m |= ((unsigned)(-__float_as_int(v[c4])) >> 31) << c4;
        m |= ((unsigned)(-__float_as_int(v[c4 + 1])) >> 31) << (c4 + 1);
        m |= ((unsigned)(-__float_as_int(v[c4 + 2])) >> 31) << (c4 + 2);
        m |= ((unsigned)(-__float_as_int(v[c4 + 3])) >> 31) << (c4 + 3);
      }
      pt_mask[j] = m;
    }
    const unsigned firstmask = __ballot_sync(0xffffffffu, first), lastmask = __ballot_sync(0xffffffffu, last);
    const int nvalid = min(32, n - grp * 32);
    // segmented inclusive scan over the lanes; a segment starts at lane 0 (possibly the tail of an earlier group's
    // pillar) and at every first point of a pillar
    const unsigned starts = firstmask | 1u;
    const int seg_start = 31 - __clz(starts & (0xffffffffu >> (31 - lane)));
    const int dist = lane - seg_start;
    const unsigned far = __ballot_sync(0xffffffffu, dist >= 4);
#pragma unroll
    for (int c = 0; c < PFN_C; ++c) {
      float x = v[c], t;
      t = __shfl_up_sync(0xffffffffu, x, 1); if (dist >= 1) x += t;
      t = __shfl_up_sync(0xffffffffu, x, 2); if (dist >= 2) x += t;
      v[c] = x;
    }
    if (far) {   // some segment of this group is longer than 4 points (warp-uniform)
#pragma unroll
      for (int c = 0; c < PFN_C; ++c) {
        float x = v[c], t;
        t = __shfl_up_sync(0xffffffffu, x, 4); if (dist >= 4) x += t;
        t = __shfl_up_sync(0xffffffffu, x, 8); if (dist >= 8) x += t;
        t = __shfl_up_sync(0xffffffffu, x, 16); if (dist >= 16) x += t;
        v[c] = x;
      }
    }
    const bool begins = (firstmask >> seg_start) & 1u;   // the segment this lane closes began inside this group
    float* prow = part + (size_t)grp * 2 * PFN_C;
    if (live && last) {
      if (begins) {
        const float inv = __frcp_rn((float)cnt);
#pragma unroll
        for (int c = 0; c < PFN_C; ++c) v[c] *= inv;
        if (pil_feats) store_row32(pil_feats + (size_t)q * PFN_C, v);
        if (BF16) {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(image) + (size_t)pix * PFN_C);
#pragma unroll
          for (int c8 = 0; c8 < PFN_C; c8 += 8) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(v[c8], v[c8 + 1]), p1 = __floats2bfloat162_rn(v[c8 + 2], v[c8 + 3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(v[c8 + 4], v[c8 + 5]), p3 = __floats2bfloat162_rn(v[c8 + 6], v[c8 + 7]);
            uint4 u;
            u.x = *reinterpret_cast<unsigned*>(&p0); u.y = *reinterpret_cast<unsigned*>(&p1);
            u.z = *reinterpret_cast<unsigned*>(&p2); u.w = *reinterpret_cast<unsigned*>(&p3);
            dst[c8 / 8] = u;
          }
        } else {
          store_row32(reinterpret_cast<float*>(image) + (size_t)pix * PFN_C, v);
        }
      } else {
        store_row32(prow, v);            // leading segment of a pillar that began in an earlier group
      }
    } else if (live && lane == nvalid - 1) {
      store_row32(prow + (begins ? PFN_C : 0), v);   // the pillar continues into the next group
    }
  }
}

// warp per group boundary: the group in which a straddling pillar BEGINS adds up that pillar's partial sums
template <bool BF16>
__global__ void __launch_bounds__(256) k_pfn_straddlers(const int* __restrict__ counts, int F,
                                                        const float4* __restrict__ rec, const float4* __restrict__ pil_hdr,
                                                        const float* __restrict__ part,
                                                        float* __restrict__ pil_feats, void* __restrict__ image) {
  const int n = counts[2 * F + F];
  const int c = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int n_groups = (n + 31) >> 5;
  for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_groups; g += warps) {
    const int jl = 32 * g + 31;
    if (jl >= n) break;                        // a partial last group has nothing after it
    const int q = rec_q(__ldg(rec + jl));
    const float4 h0 = __ldg(pil_hdr + 3 * (size_t)q);
    const int s0 = __float_as_int(h0.x), cnt = __float_as_int(h0.y), s1 = s0 + cnt;
    if (s1 <= jl + 1 || s0 < 32 * g) continue; // ends here, or began earlier (then an earlier group owns it)
    float tot = part[((size_t)g * 2 + 1) * PFN_C + c];
    const int g_last = (s1 - 1) >> 5;
    for (int g2 = g + 1; g2 <= g_last; ++g2) tot += part[(size_t)g2 * 2 * PFN_C + c];
    store_pillar<BF16>(tot * __frcp_rn((float)cnt), q, __float_as_int(h0.z), c, pil_feats, image);
  }
}

// ---------------------------------------------------------------- backward
// ONE pass over the points.  With gy^[p,c] = relu'(.) * grad_pillar[q(p),c] / count_q the only sums that need the
// points are, per frame and channel,   A1[c] = sum_p gy^[p,c]   and   T[c][k] = sum_p gy^[p,c] f[p,k]   (10 values):
//   sum_p gy^ lin  = W_c . T_c                       (lin = W f is linear)
//   sum_p xhat f_k = rstd (sum_j W_cj S2[j,k] - mean S1[k])          (S1, S2 = the forward's feature moments)
// so  A2 = sum gy^ xhat = rstd (W_c . T_c - mean A1),  grad_gamma = A2,  grad_beta = A1  and
//   grad_W[c,k] = a_c (T[c][k] - (A1/N) S1[k] - (A2/N) sum_p xhat f_k)      (training; eval: a_c T[c][k]).
// grid (X, F): a warp owns 32 consecutive CSR positions of frame f.  Phase 1, lane = point: coalesced loads of the
// point records, the ReLU mask words the forward saved and the pillar headers into a shared slab.  Then the image
// gradient rows of the (at most 32) pillars the group touches are fetched, all loads in flight together -- the kernel
// is latency-bound, not bandwidth-bound.  Phase 2, lane = channel: per pillar segment the masked sums of (1, x, y, z),
// expanded once per segment.  Everything is linear in the segment sums, so pillars that straddle groups need no
// special handling.  bwd_acc layout [F][32][10] (double): A1 | T[9].
struct BwdGeom {       // pixel of a point, recomputed from its coordinates exactly as the index pass did
  float lox, loy, vx, vy;
  int W, HW;
};

template <bool BF16>
struct BwdSlab {
  static constexpr int ROW_BYTES = BF16 ? 64 : 128;
  float4 pt[32];        // x, y, z of the group's points
  float4 h1[32], h2[32];// per segment: mean xyz, centre x | centre y, centre z
  float ic[32];         // per segment: 1 / point count
  __align__(16) unsigned char g[32][ROW_BYTES];   // per segment: image-gradient row, raw (cp.async destination)
};

// transpose the 32 x 32 bit matrix held one row per lane (5 butterfly exchanges): lane c ends up with bit r = row r's bit c
__device__ __forceinline__ unsigned transpose_bits(unsigned x, int lane) {
#pragma unroll
  for (int k = 16; k >= 1; k >>= 1) {
    const unsigned m0 = k == 16 ? 0x0000ffffu : k == 8 ? 0x00ff00ffu : k == 4 ? 0x0f0f0f0fu : k == 2 ? 0x33333333u : 0x55555555u;
    const unsigned o = __shfl_xor_sync(0xffffffffu, x, k);
    x = (lane & k) ? ((x & ~m0) | ((o >> k) & m0)) : ((x & m0) | ((o << k) & ~m0));
  }
  return x;
}

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Latency plan per group of 32 CSR positions: the point records and ReLU mask words of the NEXT group are already in
// registers (prefetched during the previous compute phase); the pillar headers (segment starts only) and the image
// gradient rows are fetched TOGETHER -- the row address comes from the point itself (its pixel is recomputed from xyz),
// not from the header -- and the rows go straight to shared memory with cp.async (16 bytes per lane, 8 / 4 rows per
// instruction), so no load result is waited for before the next load is issued.
template <bool BF16>
__global__ void __launch_bounds__(256) k_pfn_bwd(const int* __restrict__ counts, int F, BwdGeom Gm,
                                                 const float4* __restrict__ rec, const unsigned* __restrict__ pt_mask,
                                                 const float4* __restrict__ pil_hdr,
                                                 const void* __restrict__ grad_image, double* __restrict__ bwd_acc) {
  using Slab = BwdSlab<BF16>;
  __shared__ Slab slabs[8];
  static_assert(sizeof(Slab) * 8 >= sizeof(float) * 8 * PFN_C * (PFN_K + 2), "the block reduction reuses the slabs");
  float (*red)[PFN_C][PFN_K + 2] = reinterpret_cast<float (*)[PFN_C][PFN_K + 2]>(slabs);
  const int f = blockIdx.y;
  const int p0 = counts[2 * F + f], p1 = counts[2 * F + f + 1];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  Slab& S = slabs[wib];
  const int warps = (gridDim.x * blockDim.x) >> 5;
  float acc[PFN_K + 1];   // A1 | T[9] of channel `lane`
#pragma unroll
  for (int k = 0; k < PFN_K + 1; ++k) acc[k] = 0.f;
  const int n_groups = (p1 - p0 + 31) >> 5;
  constexpr int LPR = Slab::ROW_BYTES / 16;   // lanes per row
  constexpr int RPI = 32 / LPR;               // rows per cp.async instruction
  int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned mrow = 0u;
  if (grp < n_groups && p0 + grp * 32 + lane < p1) { r = __ldg(rec + p0 + grp * 32 + lane); mrow = __ldg(pt_mask + p0 + grp * 32 + lane); }
  for (; grp < n_groups; grp += warps) {
    const int j = p0 + grp * 32 + lane;
    const bool live = j < p1;
    const int q = live ? rec_q(r) : -1;
    S.pt[lane] = r;
    const int qprev = __shfl_up_sync(0xffffffffu, q, 1);
    const bool start = live && (lane == 0 || q != qprev);
    const unsigned startmask = __ballot_sync(0xffffffffu, start);
    const int slot = __popc(startmask & ((1u << lane) - 1u));
    const int nseg = __popc(startmask);
    // pixel of the point = f*H*W + cy*W + cx with the index pass's own voxel formula (the point is known to be in range)
    const int pix = f * Gm.HW + voxel_coord(r.y, Gm.loy, Gm.vy) * Gm.W + voxel_coord(r.x, Gm.lox, Gm.vx);
    if (start) {
      const float4* h = pil_hdr + 3 * (size_t)q;
      const float4 h0 = __ldg(h);
      S.h1[slot] = __ldg(h + 1);
      S.h2[slot] = __ldg(h + 2);
      S.ic[slot] = __frcp_rn((float)__float_as_int(h0.y));   // mean backward: grad / count (scatter_points_cuda_kernel.cuh:134-137)
    }
    // image-gradient rows of the group's segments -> shared memory, asynchronously
    for (int s0 = 0; s0 < nseg; s0 += RPI) {
      const int s = s0 + lane / LPR;
      // lane of the s-th set bit of startmask
      const int src = s < nseg ? __fns(startmask, 0, s + 1) : 0;
      const int px = __shfl_sync(0xffffffffu, pix, src);
      if (s < nseg)
        cp_async16(&S.g[s][(lane % LPR) * 16],
                   reinterpret_cast<const unsigned char*>(grad_image) + (size_t)px * Slab::ROW_BYTES + (lane % LPR) * 16);
    }
    // prefetch the next group's records while this one is processed
    const unsigned col = transpose_bits(mrow, lane);   // bit i = ReLU decision of point i for channel `lane`
    {
      const int jn = j + warps * 32;
      r = make_float4(0.f, 0.f, 0.f, 0.f); mrow = 0u;
      if (grp + warps < n_groups && jn < p1) { r = __ldg(rec + jn); mrow = __ldg(pt_mask + jn); }
    }
    cp_async_wait_all();
    __syncwarp();
    // lane = channel: masked sums of (1, x, y, z) over each pillar segment, expanded once per segment
    float u0 = 0.f, ux = 0.f, uy = 0.f, uz = 0.f;
    int s = -1;
    auto flush = [&]() {
      const float4 m = S.h1[s], c2 = S.h2[s];
      const float graw = BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(S.g[s])[lane])
                              : reinterpret_cast<const float*>(S.g[s])[lane];
      const float g = graw * S.ic[s];
      // offsets of the three decoration groups: raw (0), cluster mean (m.xyz), pillar centre (m.w, c2.x, c2.y)
      acc[0] = fmaf(g, u0, acc[0]);
      acc[1] = fmaf(g, ux, acc[1]); acc[2] = fmaf(g, uy, acc[2]); acc[3] = fmaf(g, uz, acc[3]);
      acc[4] = fmaf(g, fmaf(-m.x, u0, ux), acc[4]); acc[5] = fmaf(g, fmaf(-m.y, u0, uy), acc[5]);
      acc[6] = fmaf(g, fmaf(-m.z, u0, uz), acc[6]);
      acc[7] = fmaf(g, fmaf(-m.w, u0, ux), acc[7]); acc[8] = fmaf(g, fmaf(-c2.x, u0, uy), acc[8]);
      acc[9] = fmaf(g, fmaf(-c2.y, u0, uz), acc[9]);
      u0 = 0.f; ux = 0.f; uy = 0.f; uz = 0.f;
    };
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if ((startmask >> i) & 1u) {   // warp-uniform
        if (s >= 0) flush();
        ++s;
      }
      const float4 pt = S.pt[i];
      if ((col >> i) & 1u) { u0 += 1.f; ux += pt.x; uy += pt.y; uz += pt.z; }
    }
    if (s >= 0) flush();
    __syncwarp();
  }
  __syncthreads();   // every warp is done with its slab
#pragma unroll
  for (int k = 0; k < PFN_K + 1; ++k) red[wib][lane][k] = acc[k];
  __syncthreads();
  for (int i = threadIdx.x; i < PFN_C * (PFN_K + 1); i += blockDim.x) {
    const int c = i / (PFN_K + 1), k = i % (PFN_K + 1);
    double s = 0.0;
    for (int j = 0; j < 8; ++j) s += (double)red[j][c][k];
    if (s != 0.0) atomicAdd(&bwd_acc[((size_t)f * PFN_C + c) * (PFN_K + 1) + k], s);
  }
}

// One block, thread = (frame slot, channel): the algebra above in double, frames strided over the 32 slots, then a
// reduction over the slots through shared memory.
__device__ __forceinline__ int tri_index(int j, int k) {  // position of S2[j][k], j <= k, in the packed upper triangle
  return PFN_K + j * PFN_K - (j * (j - 1)) / 2 + (k - j);
}

__global__ void __launch_bounds__(512) k_pfn_bwd_finalize(const int* __restrict__ counts, int F, int training,
                                                           const double* __restrict__ bwd_acc,
                                                           const double* __restrict__ mom,
                                                           const float* __restrict__ bn_params,
                                                           const float* __restrict__ weight,
                                                           float* __restrict__ grad_weight, float* __restrict__ grad_gamma,
                                                           float* __restrict__ grad_beta,
                                                           const double* __restrict__ sync_bwd_acc,
                                                           const int* __restrict__ sync_counts) {
  // SyncBatchNorm (sync_* given): bwd_acc / mom / counts are THIS rank's sums, sync_bwd_acc / sync_counts the sums over
  // all ranks.  grad_gamma, grad_beta and grad_W stay local sums over this rank's points (the data-parallel gradient
  // mean follows), but the two batch means inside the BatchNorm backward are the pooled ones:
  //   grad_W[c,k] = a_c (T_loc[c][k] - (A1_glob/N_glob) S1_loc[k] - (A2_glob/N_glob) sum_{p local} xhat f_k)
  __shared__ double sred[16][PFN_C + 1];
  const int c = threadIdx.x & 31, slot = threadIdx.x >> 5;
  double w[PFN_K], out[PFN_K + 2];  // gw[9] | gg | gb
#pragma unroll
  for (int k = 0; k < PFN_K; ++k) w[k] = (double)weight[c * PFN_K + k];
#pragma unroll
  for (int k = 0; k < PFN_K + 2; ++k) out[k] = 0.0;
  for (int f = slot; f < F; f += 16) {
    if (counts[f] <= 0) continue;
    const int n = sync_counts ? sync_counts[f] : counts[f];
    const double* A = bwd_acc + ((size_t)f * PFN_C + c) * (PFN_K + 1);
    const float* bp = bn_params + (size_t)f * 4 * PFN_C;
    const double a = bp[c], mu = bp[2 * PFN_C + c], rstd = bp[3 * PFN_C + c];
    const double A1 = A[0];
    double wt = 0.0;
#pragma unroll
    for (int k = 0; k < PFN_K; ++k) wt += w[k] * A[1 + k];
    const double A2 = rstd * (wt - mu * A1);
    out[PFN_K] += A2;
    out[PFN_K + 1] += A1;
    if (training) {
      const double* m = mom + (size_t)f * PFN_MOM_PITCH;  // S1[9] | S2 upper triangle
      double A1g = A1, A2g = A2;
      if (sync_bwd_acc) {
        const double* Ag = sync_bwd_acc + ((size_t)f * PFN_C + c) * (PFN_K + 1);
        double wtg = 0.0;
#pragma unroll
        for (int k = 0; k < PFN_K; ++k) wtg += w[k] * Ag[1 + k];
        A1g = Ag[0];
        A2g = rstd * (wtg - mu * A1g);
      }
      const double m1 = A1g / n, m2 = A2g / n;
#pragma unroll
      for (int k = 0; k < PFN_K; ++k) {
        double ws2 = 0.0;
#pragma unroll
        for (int j = 0; j < PFN_K; ++j) ws2 += w[j] * m[j <= k ? tri_index(j, k) : tri_index(k, j)];
        const double sxf = rstd * (ws2 - mu * m[k]);  // sum_p xhat[p,c] f[p,k]
        out[k] += a * (A[1 + k] - m1 * m[k] - m2 * sxf);
      }
    } else {
#pragma unroll
      for (int k = 0; k < PFN_K; ++k) out[k] += a * A[1 + k];
    }
  }
#pragma unroll
  for (int k = 0; k < PFN_K + 2; ++k) {
    sred[slot][c] = out[k];
    __syncthreads();
    if (slot == 0) {
      double s = 0.0;
      for (int i = 0; i < 16; ++i) s += sred[i][c];
      if (k < PFN_K) grad_weight[c * PFN_K + k] += (float)s;
      else if (k == PFN_K) grad_gamma[c] += (float)s;
      else grad_beta[c] += (float)s;
    }
    __syncthreads();
  }
}

// Dense zero fill with ONE small block per SM (16-byte stores, grid-stride): enough store traffic to saturate HBM while
// leaving almost all warp slots free, so that it can run under other kernels on a second stream (a library memset /
// fill launches enough blocks to occupy every SM until it is done).
__global__ void __launch_bounds__(256) k_zero_fill(uint4* __restrict__ p, size_t n16, unsigned char* __restrict__ tail,
                                                   int ntail) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) { p[i] = z; p[i + stride] = z; p[i + 2 * stride] = z; p[i + 3 * stride] = z; }
  for (; i < n16; i += stride) p[i] = z;
  if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail[threadIdx.x] = 0;
}

// rows of 16-byte vectors: row r = pix[r], v16 vectors each
__global__ void __launch_bounds__(256) k_clear_rows(uint4* __restrict__ image, int v16, const int* __restrict__ pix,
                                                    const int* __restrict__ counts, int count_index, long long cap) {
  const long long n = min((long long)counts[count_index], cap);
  const long long total = n * v16;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / v16;
    image[(size_t)__ldg(pix + r) * v16 + (e - r * v16)] = z;
  }
}

static PfnGeom make_geom(const dfb_pfn_args* a) {
  PfnGeom G;
  G.vx = a->voxel_size[0]; G.vy = a->voxel_size[1]; G.vz = a->voxel_size[2];
  G.ox = a->center_off[0]; G.oy = a->center_off[1]; G.oz = a->center_off[2];
  return G;
}

}  // namespace dfb

using namespace dfb;

extern "C" int dfb_zero_fill(void* ptr, long long bytes, int blocks_per_sm, void* stream_) {
  if (bytes < 0 || ((uintptr_t)ptr & 15)) { set_error("dfb_zero_fill: pointer must be 16-byte aligned"); return DFB_ERR_ARG; }
  if (bytes == 0) return DFB_OK;
  const size_t n16 = (size_t)bytes / 16;
  if (blocks_per_sm < 1) blocks_per_sm = 1;
  k_zero_fill<<<sm_count() * blocks_per_sm, 256, 0, (cudaStream_t)stream_>>>((uint4*)ptr, n16, (unsigned char*)ptr + n16 * 16,
                                                                            (int)(bytes - (long long)n16 * 16));
  add_launches(1);
  return check_launch("dfb_zero_fill");
}

extern "C" int dfb_clear_rows(void* image, int row_bytes, const int* pix, const int* counts, int count_index, long long cap,
                              void* stream_) {
  if (!image || !pix || !counts || row_bytes <= 0 || row_bytes % 16 || ((uintptr_t)image & 15)) {
    set_error("dfb_clear_rows: rows must be 16-byte multiples of a 16-byte aligned image"); return DFB_ERR_ARG;
  }
  if (cap <= 0) return DFB_OK;
  k_clear_rows<<<sm_count() * 8, 256, 0, (cudaStream_t)stream_>>>((uint4*)image, row_bytes / 16, pix, counts, count_index, cap);
  add_launches(1);
  return check_launch("dfb_clear_rows");
}

// a->phase: 0 = everything; 1 = up to the per-frame feature moments (a->stats); 2 = from the BatchNorm finalisation on.
// SyncBatchNorm callers run phase 1, sum a->stats and the per-frame counts over the ranks (into a->sync_stats /
// a->sync_counts) and run phase 2; a->stats keeps THIS rank's moments for the backward.
extern "C" int dfb_pfn_forward(const dfb_pfn_args* a, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (!a || a->F <= 0 || a->H <= 0 || a->W <= 0) { set_error("dfb_pfn_forward: bad sizes"); return DFB_ERR_ARG; }
  if (a->F > 256) { set_error("dfb_pfn_forward: at most 256 frames per call"); return DFB_ERR_UNSUPPORTED; }
  if (!a->csr_rec || !a->pt_mask || !a->partials || !a->pil_hdr) { set_error("dfb_pfn_forward: csr_rec / pt_mask / partials / pil_hdr missing"); return DFB_ERR_ARG; }
  if (a->phase < 0 || a->phase > 2) { set_error("dfb_pfn_forward: phase must be 0, 1 or 2"); return DFB_ERR_ARG; }
  const int F = a->F, HW = a->H * a->W;
  const PfnGeom G = make_geom(a);
  const int sms = sm_count();
  const float4* rec = (const float4*)a->csr_rec;
  float4* hdr = (float4*)a->pil_hdr;
  int launches = 0;
  if (a->phase != 2) {
    const size_t img_bytes = (size_t)F * HW * PFN_C * (a->image_bf16 ? 2 : 4);
    // the dense zero canvas (PointPillarsScatter, encoder.py:135-139).  A caller that zero-fills the image itself on
    // another stream (to overlap the fill with the index kernels) passes the event that marks its completion.
    if (!a->image_ready_event) {
      const size_t n16 = img_bytes / 16;   // F*H*W*32 elements: always a multiple of 16 bytes
      k_zero_fill<<<sms * 2, 256, 0, st>>>((uint4*)a->image, n16, (unsigned char*)a->image + n16 * 16, (int)(img_bytes - n16 * 16));
      ++launches;
    }
    cudaMemsetAsync(a->stats, 0, sizeof(double) * (size_t)F * PFN_MOM_PITCH, st);
    k_pillar_mean<<<sms * 8, 256, 0, st>>>(a->counts, F, HW, G, rec, a->pil_start, a->pil_pix, a->pil_coor, a->pil_mean, hdr);
    ++launches;
    if (a->training) {
      int bx = (sms * 3) / F;   // 80 registers -> 3 resident blocks per SM: one wave, no tail
      if (bx < 1) bx = 1;
      dim3 g(bx, F);
      k_pfn_moments<<<g, 256, 0, st>>>(a->counts, F, rec, hdr, a->stats);
      ++launches;
    }
  }
  if (a->phase != 1) {
    k_bn_finalize<<<1, 1024, 0, st>>>(a->counts, F, a->training, a->eps, a->momentum,
                                      a->sync_stats ? (const double*)a->sync_stats : (const double*)a->stats, a->weight, a->gamma,
                                      a->beta, a->running_mean, a->running_var, a->bn_params, (const int*)a->sync_counts);
    if (a->image_ready_event) cudaStreamWaitEvent(st, (cudaEvent_t)a->image_ready_event, 0);
    const int dyn = F * 2 * PFN_C * (int)sizeof(float);
    static bool configured = false;
    if (!configured) {
      cudaFuncSetAttribute(k_pfn_points<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 2 * PFN_C * (int)sizeof(float));
      cudaFuncSetAttribute(k_pfn_points<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 2 * PFN_C * (int)sizeof(float));
      configured = true;
    }
    // Linear(9,32) weights -> constant bank (device-to-device, stream-ordered): the point pass reads them as FFMA
    // constant operands instead of 72 shared-memory broadcast loads per point (ncu: the pass was LSU-bound).  One
    // constant copy per device: concurrent forwards of DIFFERENT feature nets must not run on different streams.
    cudaMemcpyToSymbolAsync(c_pfn_w, a->weight, sizeof(float) * PFN_C * PFN_K, 0, cudaMemcpyDeviceToDevice, st);
    if (a->image_bf16)
      k_pfn_points<true><<<sms * 3, 256, dyn, st>>>(a->counts, F, rec, hdr, a->bn_params, a->pt_mask, a->partials,
                                                    a->pil_feats, a->image);
    else
      k_pfn_points<false><<<sms * 3, 256, dyn, st>>>(a->counts, F, rec, hdr, a->bn_params, a->pt_mask, a->partials,
                                                     a->pil_feats, a->image);
    if (a->image_bf16) k_pfn_straddlers<true><<<sms * 8, 256, 0, st>>>(a->counts, F, rec, hdr, a->partials, a->pil_feats, a->image);
    else k_pfn_straddlers<false><<<sms * 8, 256, 0, st>>>(a->counts, F, rec, hdr, a->partials, a->pil_feats, a->image);
    launches += 3;
  }
  add_launches(launches);
  return check_launch("dfb_pfn_forward");
}

// b->phase: 0 = everything; 1 = the point pass only (b->bwd_stats = this rank's A1 | T sums); 2 = the finalisation only,
// with b->sync_bwd_stats = those sums over all ranks (SyncBatchNorm; b->fwd.sync_counts = pooled counts).
extern "C" int dfb_pfn_backward(const dfb_pfn_bwd_args* b, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (!b) { set_error("dfb_pfn_backward: null args"); return DFB_ERR_ARG; }
  const dfb_pfn_args* a = &b->fwd;
  if (!a->csr_rec || !a->pt_mask || !a->pil_hdr) { set_error("dfb_pfn_backward: csr_rec / pt_mask / pil_hdr missing"); return DFB_ERR_ARG; }
  if (b->phase < 0 || b->phase > 2) { set_error("dfb_pfn_backward: phase must be 0, 1 or 2"); return DFB_ERR_ARG; }
  const int F = a->F;
  const int sms = sm_count();
  const float4* rec = (const float4*)a->csr_rec;
  int launches = 0;
  if (b->phase != 2) {
    cudaMemsetAsync(b->bwd_stats, 0, sizeof(double) * (size_t)F * PFN_C * (PFN_K + 1), st);
    if (((uintptr_t)b->grad_image & 15)) { set_error("dfb_pfn_backward: grad_image must be 16-byte aligned"); return DFB_ERR_ARG; }
    BwdGeom Gm;
    Gm.lox = a->range_min[0]; Gm.loy = a->range_min[1]; Gm.vx = a->voxel_size[0]; Gm.vy = a->voxel_size[1];
    Gm.W = a->W; Gm.HW = a->H * a->W;
    // 30 KB of shared memory / 48 registers (bf16) -> 5 resident blocks per SM, 46 KB / 64 registers (fp32) -> 4: one wave
    int bx = (sms * (a->image_bf16 ? 5 : 4)) / F;
    if (bx < 1) bx = 1;
    dim3 g(bx, F);
    if (a->image_bf16)
      k_pfn_bwd<true><<<g, 256, 0, st>>>(a->counts, F, Gm, rec, a->pt_mask, (const float4*)a->pil_hdr, b->grad_image, b->bwd_stats);
    else
      k_pfn_bwd<false><<<g, 256, 0, st>>>(a->counts, F, Gm, rec, a->pt_mask, (const float4*)a->pil_hdr, b->grad_image, b->bwd_stats);
    ++launches;
  }
  if (b->phase != 1) {
    k_pfn_bwd_finalize<<<1, 512, 0, st>>>(a->counts, F, a->training, b->bwd_stats, a->stats, a->bn_params, a->weight,
                                           b->grad_weight, b->grad_gamma, b->grad_beta, (const double*)b->sync_bwd_stats,
                                           (const int*)a->sync_counts);
    ++launches;
  }
  add_launches(launches);
  return check_launch("dfb_pfn_backward");
}
