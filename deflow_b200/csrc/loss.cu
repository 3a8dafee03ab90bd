// deflowLoss / ff3dLoss / zeroflowLoss with the trainer's ground-truth construction fused in, for all samples of a
// step, without host synchronisation.
//
// Reference: deflowLoss (OpenSceneFlow/src/lossfuncs.py:102-125), zeroflowLoss (:128-145), ff3dLoss (:148-157) and the loop of
// ModelWrapper.training_step (OpenSceneFlow/src/trainer.py:120-142):
//   gt[p]  = flow[b][idx[p]] - pose_flow[b][idx[p]]
//   deflow : speed = |gt| / 0.1; mean |est - gt| over the buckets speed < 0.4, 0.4 <= speed <= 1.0,
//            speed > 1.0; empty buckets (NaN mean) are skipped; the per-sample losses are SUMMED.
//   ff3d   : mean(|est - gt| * (0.1 + 0.9 * [class > 0])).
//   zeroflow (kind 2): mean over the finite points of |est - gt| * clamp(1.8 * (|gt| * 10) - 0.8, 0.1, 1.0).
// The reference launches three boolean-mask kernels and three isnan() host syncs per sample.
#include "common.cuh"
#include "../../include/deflow_b200.h"

namespace dfb {

// bucket_ws layout per sample: [0..2] error sums, [3..5] counts (as doubles), [6] ff3d sum, [7] n
constexpr int LW = 8;

__device__ __forceinline__ bool finite3(float a, float b, float c) { return isfinite(a) && isfinite(b) && isfinite(c); }

__device__ __forceinline__ int frame_of(const int* __restrict__ pt_off, int B, int p) {
  // B is small (<= 64): linear search of the frame offsets
  int b = 0;
  while (b + 1 < B && p >= pt_off[b + 1]) ++b;
  return b;
}

struct PointTerm {
  float ex, ey, ez, err;
  int bucket;  // -1 = masked out
  float w;     // ff3d weight
};

__device__ __forceinline__ PointTerm point_term(int kind, int p, int b, const float* __restrict__ est,
                                                const float* __restrict__ flow_gt,
                                                const float* __restrict__ pose_flow,
                                                const unsigned char* __restrict__ classes,
                                                const long long* __restrict__ pt_idx, int Nmax) {
  PointTerm t;
  const size_t src = (size_t)b * Nmax + (size_t)pt_idx[p];
  const float gx = __fsub_rn(flow_gt[3 * src], pose_flow[3 * src]);
  const float gy = __fsub_rn(flow_gt[3 * src + 1], pose_flow[3 * src + 1]);
  const float gz = __fsub_rn(flow_gt[3 * src + 2], pose_flow[3 * src + 2]);
  const float px = est[3 * (size_t)p], py = est[3 * (size_t)p + 1], pz = est[3 * (size_t)p + 2];
  t.ex = px - gx; t.ey = py - gy; t.ez = pz - gz;
  t.err = sqrtf(t.ex * t.ex + t.ey * t.ey + t.ez * t.ez);
  t.w = 1.f;
  if (kind == 0) {
    // mask = ~isnan & ~isinf on both tensors (lossfuncs.py:107-110); a point is dropped as a whole
    if (!finite3(gx, gy, gz) || !finite3(px, py, pz)) { t.bucket = -1; return t; }
    const float speed = __fdiv_rn(sqrtf(gx * gx + gy * gy + gz * gz), 0.1f);
    t.bucket = speed < 0.4f ? 0 : (speed <= 1.0f ? 1 : 2);
  } else if (kind == 2) {
    if (!finite3(gx, gy, gz) || !finite3(px, py, pz)) { t.bucket = -1; return t; }      // lossfuncs.py:131-134
    const float speed = __fmul_rn(sqrtf(gx * gx + gy * gy + gz * gz), 10.0f);             // :138
    t.w = fmaxf(0.1f, fminf(__fsub_rn(__fmul_rn(1.8f, speed), 0.8f), 1.0f));              // :140-142
    t.bucket = 0;
  } else {
    t.bucket = 0;
    t.w = (classes && classes[src] > 0) ? 1.0f : 0.1f;  // 0.1 + 0.9 * [class > 0] (lossfuncs.py:154-156)
  }
  return t;
}

__global__ void __launch_bounds__(256) k_loss_accumulate(int kind, const float* __restrict__ est,
                                                         const float* __restrict__ flow_gt,
                                                         const float* __restrict__ pose_flow,
                                                         const unsigned char* __restrict__ classes,
                                                         const long long* __restrict__ pt_idx,
                                                         const int* __restrict__ counts, int F, int B, int Nmax,
                                                         double* __restrict__ ws, int n_cap) {
  __shared__ double sh[8][LW];
  const int b = blockIdx.y;
  const int p0 = counts[2 * F + b], p1 = min(counts[2 * F + b + 1], n_cap);
  float s[3] = {0.f, 0.f, 0.f};
  int c[3] = {0, 0, 0};
  for (int p = p0 + blockIdx.x * blockDim.x + threadIdx.x; p < p1; p += gridDim.x * blockDim.x) {
    const PointTerm t = point_term(kind, p, b, est, flow_gt, pose_flow, classes, pt_idx, Nmax);
    if (t.bucket >= 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (t.bucket == k) { s[k] += t.err * t.w; c[k] += 1; }
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double ss = warp_sum((double)s[k]);
    const double cc = warp_sum((double)c[k]);
    if (lane == 0) { sh[w][k] = ss; sh[w][3 + k] = cc; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double a = 0.0;
    for (int i = 0; i < 8; ++i) a += sh[i][threadIdx.x];
    if (a != 0.0) atomicAdd(&ws[(size_t)b * LW + threadIdx.x], a);
  }
}

__global__ void k_loss_finalize(int kind, int B, const int* __restrict__ counts, int F, const double* __restrict__ ws,
                                float* __restrict__ loss) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float total = 0.f;
  for (int b = 0; b < B; ++b) {
    const double* w = ws + (size_t)b * LW;
    if (kind == 0) {
      // order of the reference's accumulation: speed > 1.0, then < 0.4, then the middle bucket
      float l = 0.f;
      if (w[5] > 0) l += (float)(w[2] / w[5]);
      if (w[3] > 0) l += (float)(w[0] / w[3]);
      if (w[4] > 0) l += (float)(w[1] / w[4]);
      total += l;
    } else if (kind == 2) {
      total += w[3] > 0 ? (float)(w[0] / w[3]) : __int_as_float(0x7fc00000);   // mean over the finite points
    } else {
      const int n = counts[b];
      total += n > 0 ? (float)(w[0] / (double)n) : __int_as_float(0x7fc00000);  // mean of empty = NaN
    }
  }
  loss[0] = total;
}

__global__ void __launch_bounds__(256) k_loss_grad(int kind, const float* __restrict__ est,
                                                   const float* __restrict__ flow_gt,
                                                   const float* __restrict__ pose_flow,
                                                   const unsigned char* __restrict__ classes,
                                                   const long long* __restrict__ pt_idx,
                                                   const int* __restrict__ counts, int F, int B, int Nmax,
                                                   const double* __restrict__ ws, float* __restrict__ grad_est,
                                                   int n_cap) {
  const int n = min(counts[2 * F + B], n_cap);
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const int b = frame_of(counts + 2 * F, B, p);
    const PointTerm t = point_term(kind, p, b, est, flow_gt, pose_flow, classes, pt_idx, Nmax);
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (t.bucket >= 0 && t.err > 0.f) {  // d|v|/dv = v/|v|, 0 at v = 0 (torch vector_norm backward)
      const double denom = kind != 1 ? ws[(size_t)b * LW + 3 + t.bucket] : (double)counts[b];
      const float s = t.w / ((float)denom * t.err);
      gx = t.ex * s; gy = t.ey * s; gz = t.ez * s;
    }
    grad_est[3 * (size_t)p] = gx;
    grad_est[3 * (size_t)p + 1] = gy;
    grad_est[3 * (size_t)p + 2] = gz;
  }
}

}  // namespace dfb

using namespace dfb;

extern "C" int dfb_flow_loss(int kind, const float* est, const float* flow_gt, const float* pose_flow,
                             const unsigned char* classes, const long long* pt_idx, const int* counts, int F, int B,
                             int Nmax, double* bucket_ws, float* loss, float* grad_est, int n_cap, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (kind < 0 || kind > 2) { set_error("dfb_flow_loss: unknown loss kind %d (0 deflowLoss, 1 ff3dLoss, 2 zeroflowLoss)", kind); return DFB_ERR_ARG; }
  if (B <= 0 || F < B) { set_error("dfb_flow_loss: bad sizes"); return DFB_ERR_ARG; }
  cudaMemsetAsync(bucket_ws, 0, sizeof(double) * (size_t)B * LW, st);
  int bx = (sm_count() * 4 + B - 1) / B;
  if (bx < 1) bx = 1;
  dim3 g(bx, B);
  k_loss_accumulate<<<g, 256, 0, st>>>(kind, est, flow_gt, pose_flow, classes, pt_idx, counts, F, B, Nmax, bucket_ws, n_cap);
  k_loss_finalize<<<1, 32, 0, st>>>(kind, B, counts, F, bucket_ws, loss);
  int launches = 2;
  if (grad_est && n_cap > 0) {
    long long blocks = ((long long)n_cap + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    k_loss_grad<<<(int)blocks, 256, 0, st>>>(kind, est, flow_gt, pose_flow, classes, pt_idx, counts, F, B, Nmax, bucket_ws, grad_est, n_cap);
    ++launches;
  }
  add_launches(launches);
  return check_launch("dfb_flow_loss");
}
