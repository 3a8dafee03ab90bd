// Ego-motion compensation of pc0 for every sample of the batch in one launch.
//
// Reference: cal_pose0to1 (OpenSceneFlow/src/models/basic/__init__.py:4-15) and the per-sample loop
// of DeFlow.forward (OpenSceneFlow/src/models/deflow.py:60-77):
//   pose1_inv = [R1^T | (R1^T * -t1).sum(axis=1)]   (translation formed in fp32, stored in fp64)
//   pose_0to1 = fp32(pose1_inv @ fp64(pose0))
//   pc0'      = pc0 @ R^T + t ;  pose_flow = pc0' - pc0       (NaN padding rows stay NaN)
#include "common.cuh"
#include "../../include/deflow_b200.h"

namespace dfb {

__global__ void __launch_bounds__(256) k_ego_warp(const float* __restrict__ pose0, const float* __restrict__ pose1,
                                                  const float* __restrict__ ego, const float* __restrict__ pc0,
                                                  int Nmax, float* __restrict__ warped, long long warped_stride_b,
                                                  float* __restrict__ pose_flow, float* __restrict__ pose_out) {
  __shared__ float T[12];  // rows 0..2 of pose_0to1
  const int b = blockIdx.y;
  if (threadIdx.x < 12) {
    const int i = threadIdx.x >> 2, j = threadIdx.x & 3;
    float v;
    if (ego) {
      v = ego[b * 16 + i * 4 + j];
    } else {
      const float* P0 = pose0 + b * 16;
      const float* P1 = pose1 + b * 16;
      // row i of pose1_inv: R1^T[i, :] = P1[0..2][i], translation summed in fp32 in index order
      const float ti = __fadd_rn(__fadd_rn(__fmul_rn(P1[0 * 4 + i], -P1[0 * 4 + 3]), __fmul_rn(P1[1 * 4 + i], -P1[1 * 4 + 3])),
                                 __fmul_rn(P1[2 * 4 + i], -P1[2 * 4 + 3]));
      double acc = 0.0;
      for (int k = 0; k < 3; ++k) acc += (double)P1[k * 4 + i] * (double)P0[k * 4 + j];
      acc += (double)ti * (double)P0[3 * 4 + j];
      v = (float)acc;
    }
    T[threadIdx.x] = v;
    if (pose_out) pose_out[b * 16 + threadIdx.x] = v;
  }
  if (threadIdx.x < 4 && pose_out) {
    float v;
    if (ego) v = ego[b * 16 + 12 + threadIdx.x];
    else {
      // last row of pose1_inv is (0,0,0,1): row 3 of the product is row 3 of pose0
      v = (float)(double)pose0[b * 16 + 12 + threadIdx.x];
    }
    pose_out[b * 16 + 12 + threadIdx.x] = v;
  }
  __syncthreads();
  const float* src = pc0 + (size_t)b * Nmax * 3;
  float* dst = warped + (size_t)b * warped_stride_b;
  float* pf = pose_flow + (size_t)b * Nmax * 3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Nmax; i += gridDim.x * blockDim.x) {
    const float x = src[3 * (size_t)i], y = src[3 * (size_t)i + 1], z = src[3 * (size_t)i + 2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float acc = __fmul_rn(x, T[r * 4]);
      acc = fmaf(y, T[r * 4 + 1], acc);
      acc = fmaf(z, T[r * 4 + 2], acc);
      const float w = __fadd_rn(acc, T[r * 4 + 3]);
      const float in = r == 0 ? x : (r == 1 ? y : z);
      dst[3 * (size_t)i + r] = w;
      pf[3 * (size_t)i + r] = __fsub_rn(w, in);
    }
  }
}

}  // namespace dfb

using namespace dfb;

extern "C" int dfb_ego_warp(const float* pose0, const float* pose1, const float* ego, const float* pc0, int B,
                            int Nmax, float* pc0_warped, long long warped_stride_b, float* pose_flow,
                            float* pose_0to1, void* stream_) {
  if (B <= 0 || Nmax < 0) { set_error("dfb_ego_warp: bad sizes"); return DFB_ERR_ARG; }
  if (!ego && (!pose0 || !pose1)) { set_error("dfb_ego_warp: need pose0/pose1 or ego_motion"); return DFB_ERR_ARG; }
  int bx = (Nmax + 255) / 256;
  const int cap = (sm_count() * 8 + B - 1) / B;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 g(bx, B);
  k_ego_warp<<<g, 256, 0, (cudaStream_t)stream_>>>(pose0, pose1, ego, pc0, Nmax, pc0_warped, warped_stride_b,
                                                   pose_flow, pose_0to1);
  add_launches(1);
  return check_launch("dfb_ego_warp");
}
