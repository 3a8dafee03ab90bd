// On-device evaluation metrics: EPE three-way + dynamic IoU, bucketed (class x speed) EPE, range-wise EPE -- one pass over
// the points of a frame, no device->host copy of any per-point tensor.
//
// Reference: evaluate_leaderboard / evaluate_leaderboard_v2 / evaluate_ssf (OpenSceneFlow/src/utils/eval_metric.py:28-106)
// build ~10 boolean masks per frame in torch, move every tensor to the host (.cpu().numpy()) and reduce in float64 numpy:
// compute_metrics (src/utils/av2_eval.py:460-553), compute_bucketed_epe (:839-870), compute_ssf_metrics (:872-915);
// called per validation frame from ModelWrapper (src/trainer.py:224-266).  Arithmetic kept as there: thresholds and
// rigid-flow subtraction in fp32, norms / sums in float64.
//
// acc layout (doubles, all += so that frames can also be pooled):
//   [0,8)    three-way subset counts   (class {background, foreground} x motion {dynamic, static} x distance {close, far})
//   [8,16)   three-way subset sums of EPE
//   [16,19)  TP, FP, FN of the dynamic segmentation
//   [19,784) bucketed: [meta class 5][speed bucket 51][count, sum error, sum speed]   (BACKGROUND uses bucket 0 only)
//   [784,814) range-wise: [range 5][Static, Dynamic][count, sum error, sum distance]
#include "common.cuh"
#include "../../include/deflow_b200.h"

namespace dfb {

constexpr int EV_THREADS = 256;
constexpr int EV_V2 = 19, EV_SSF = 784, EV_TOTAL = DFB_EVAL_ACC_DOUBLES;
static_assert(EV_TOTAL == 814, "accumulator layout");

__device__ __forceinline__ bool has_nan3(const float* p) { return isnan(p[0]) || isnan(p[1]) || isnan(p[2]); }
// fp32 2-norm as torch.linalg.vector_norm reduces three fp32 values
__device__ __forceinline__ float norm3f(float x, float y, float z) { return sqrtf(fmaf(z, z, fmaf(y, y, __fmul_rn(x, x)))); }
// float64 2-norm as numpy: sqrt(x^2 + y^2 + z^2), no contraction
__device__ __forceinline__ double norm3d(double x, double y, double z) {
  return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
}

__global__ void __launch_bounds__(EV_THREADS) k_eval_accumulate(
    const float* __restrict__ est, const float* __restrict__ rigid, const float* __restrict__ pc0, int pc_stride,
    const float* __restrict__ gt, const unsigned char* __restrict__ valid, const unsigned char* __restrict__ cls, long long n,
    const __grid_constant__ dfb_eval_tables T, double* __restrict__ acc) {
  __shared__ double s_acc[EV_TOTAL];
  for (int i = threadIdx.x; i < EV_TOTAL; i += EV_THREADS) s_acc[i] = 0.0;
  __syncthreads();
  // the hot bins live in registers: the 8 three-way subsets, TP/FP/FN, and the BACKGROUND row of the bucketed matrix
  int cnt8[8];
  double sum8[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { cnt8[i] = 0; sum8[i] = 0.0; }
  int tp = 0, fp = 0, fn = 0, bg_cnt = 0;
  double bg_err = 0.0, bg_spd = 0.0;
  for (long long i = (long long)blockIdx.x * EV_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * EV_THREADS) {
    const float* e = est + 3 * i;
    const float* r = rigid + 3 * i;
    const float* p = pc0 + (size_t)pc_stride * i;
    const float* g = gt + 3 * i;
    if (has_nan3(e) || has_nan3(r) || has_nan3(p) || has_nan3(g)) continue;     // eval_metric.py:30-31, 62, 86
    const bool v = valid[i] != 0;
    if (!v) continue;               // every one of the three metric families drops invalid points before reducing
    const int c = cls[i];
    const float ex = e[0], ey = e[1], ez = e[2], rx = r[0], ry = r[1], rz = r[2], gx = g[0], gy = g[1], gz = g[2];
    // ---- three-way (eval_metric.py:29-54, av2_eval.py:485-553)
    {
      const bool gt_dyn = norm3f(__fsub_rn(gx, rx), __fsub_rn(gy, ry), __fsub_rn(gz, rz)) >= 0.05f;
      const bool est_dyn = norm3f(__fsub_rn(ex, rx), __fsub_rn(ey, ry), __fsub_rn(ez, rz)) >= 0.05f;
      const bool close = fabsf(p[0]) <= 35.0f && fabsf(p[1]) <= 35.0f;
      const int k3 = T.fg_bg[c];    // 0 background, 1 foreground, anything else: in no subset
      if (k3 < 2) {
        const int s = k3 * 4 + (gt_dyn ? 0 : 2) + (close ? 0 : 1);
        const double epe = norm3d((double)ex - (double)gx, (double)ey - (double)gy, (double)ez - (double)gz);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (q == s) { cnt8[q] += 1; sum8[q] += epe; }
        }
        tp += (est_dyn && gt_dyn); fp += (est_dyn && !gt_dyn); fn += (!est_dyn && gt_dyn);
      }
    }
    // ---- relative flows shared by the bucketed and the range-wise metric (eval_metric.py:64-66, 91-94: fp32 subtract)
    const double ax = (double)__fsub_rn(ex, rx), ay = (double)__fsub_rn(ey, ry), az = (double)__fsub_rn(ez, rz);
    const double bx = (double)__fsub_rn(gx, rx), by = (double)__fsub_rn(gy, ry), bz = (double)__fsub_rn(gz, rz);
    const double speed = norm3d(bx, by, bz);
    const double err = norm3d(ax - bx, ay - by, az - bz);
    // ---- bucketed (eval_metric.py:57-78, av2_eval.py:839-870): points within 35 m in xy
    if (norm3f(p[0], p[1], 0.0f) <= 35.0f) {
      const int row = T.meta[c];    // 0 BACKGROUND, 1 CAR, 2 OTHER_VEHICLES, 3 PEDESTRIAN, 4 WHEELED_VRU, else none
      if (row == 0) { bg_cnt += 1; bg_err += err; bg_spd += speed; }
      else if (row < 5) {
        int b = -1;
        for (int j = 0; j < T.n_speed; ++j) {
          if (speed >= T.speed_splits[j] && speed < T.speed_splits[j + 1]) { b = j; break; }
        }
        if (b >= 0) {
          double* a = s_acc + EV_V2 + (row * 51 + b) * 3;
          atomicAdd(a, 1.0); atomicAdd(a + 1, err); atomicAdd(a + 2, speed);
        }
      }
    }
    // ---- range-wise (eval_metric.py:81-106, av2_eval.py:872-915): 3-D distance, dynamic = gt speed * 10 Hz >= 1.4 m/s
    {
      const double dist = (double)norm3f(p[0], p[1], p[2]);
      int rb = -1;
      for (int j = 0; j < T.n_dist; ++j) {
        if (dist >= T.dist_splits[j] && dist < T.dist_splits[j + 1]) { rb = j; break; }
      }
      if (rb >= 0) {
        const int dyn = __dmul_rn(speed, 10.0) >= 1.4 ? 1 : 0;
        double* a = s_acc + EV_SSF + (rb * 2 + dyn) * 3;
        atomicAdd(a, 1.0); atomicAdd(a + 1, err); atomicAdd(a + 2, dist);
      }
    }
  }
  // register bins -> warp sums -> shared
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const double c = warp_sum((double)cnt8[q]), s = warp_sum(sum8[q]);
    if ((threadIdx.x & 31) == 0 && c != 0.0) { atomicAdd(s_acc + q, c); atomicAdd(s_acc + 8 + q, s); }
  }
  {
    const double a = warp_sum((double)tp), b = warp_sum((double)fp), c = warp_sum((double)fn);
    const double d = warp_sum((double)bg_cnt), e2 = warp_sum(bg_err), f = warp_sum(bg_spd);
    if ((threadIdx.x & 31) == 0) {
      if (a != 0.0) atomicAdd(s_acc + 16, a);
      if (b != 0.0) atomicAdd(s_acc + 17, b);
      if (c != 0.0) atomicAdd(s_acc + 18, c);
      if (d != 0.0) { atomicAdd(s_acc + EV_V2, d); atomicAdd(s_acc + EV_V2 + 1, e2); atomicAdd(s_acc + EV_V2 + 2, f); }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < EV_TOTAL; i += EV_THREADS) {
    if (s_acc[i] != 0.0) atomicAdd(acc + i, s_acc[i]);
  }
}

}  // namespace dfb

using namespace dfb;

extern "C" int dfb_eval_accumulate(const float* est_flow, const float* rigid_flow, const float* pc0, int pc_stride,
                                   const float* gt_flow, const unsigned char* is_valid, const unsigned char* cls, long long n,
                                   const dfb_eval_tables* tables, double* acc, void* stream_) {
  if (n < 0 || pc_stride < 3 || !tables || !acc) { set_error("dfb_eval_accumulate: bad arguments"); return DFB_ERR_ARG; }
  if (tables->n_speed < 1 || tables->n_speed > 51 || tables->n_dist < 1 || tables->n_dist > 5) {
    set_error("dfb_eval_accumulate: 1..51 speed buckets and 1..5 distance ranges");
    return DFB_ERR_ARG;
  }
  if (n == 0) return DFB_OK;
  long long blocks = (n + EV_THREADS * 4 - 1) / (EV_THREADS * 4);
  if (blocks > sm_count() * 4) blocks = sm_count() * 4;
  if (blocks < 1) blocks = 1;
  k_eval_accumulate<<<(int)blocks, EV_THREADS, 0, (cudaStream_t)stream_>>>(est_flow, rigid_flow, pc0, pc_stride, gt_flow, is_valid,
                                                                         cls, n, *tables, acc);
  add_launches(1);
  return check_launch("dfb_eval_accumulate");
}
