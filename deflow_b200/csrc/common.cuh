// Shared device/host helpers for the deflow_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "deflow_b200 kernels target sm_100a (B200) only"
#endif

#define DFB_OK 0
#define DFB_ERR_ARG 1
#define DFB_ERR_CUDA 2
#define DFB_ERR_UNSUPPORTED 3

namespace dfb {

void set_error(const char* fmt, ...);   // api.cu: thread-local message for dfb_last_error()
int check_launch(const char* what);     // api.cu: cudaGetLastError -> status
void add_launches(int n);               // api.cu: kernel launches issued by this library (dfb_launch_count)
int sm_count();                         // api.cu: cached multiprocessor count of the current device

// Voxel grid parameters, computed on the host exactly as the reference launcher does
// (voxelization_cuda.cu:259-271): fp32 voxel size / range, grid = round((max-min)/voxel).
struct VoxelParams {
  float vx, vy, vz;
  float lox, loy, loz;
  int gx, gy, gz;
};

// c = (int)floorf((p - lo) / v) with IEEE fp32 subtract and divide (never contracted or
// replaced by a reciprocal multiply) -- voxelization_cuda_kernel.cuh:26,32,39.
__device__ __forceinline__ int voxel_coord(float p, float lo, float v) {
  return (int)floorf(__fdiv_rn(__fsub_rn(p, lo), v));
}

// status: 0 in range, 1 x failed, 2 y failed, 3 z failed (x -> y -> z order with early exit)
__device__ __forceinline__ int voxel_coords(float x, float y, float z, const VoxelParams& P, int& cx, int& cy,
                                            int& cz) {
  cx = voxel_coord(x, P.lox, P.vx);
  if (cx < 0 || cx >= P.gx) return 1;
  cy = voxel_coord(y, P.loy, P.vy);
  if (cy < 0 || cy >= P.gy) return 2;
  cz = voxel_coord(z, P.loz, P.vz);
  if (cz < 0 || cz >= P.gz) return 3;
  return 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Exclusive block scan of one int per thread.  smem must hold 33 ints.  Returns the exclusive
// prefix of v and the block total.  Ends with a barrier so smem can be reused immediately.
template <int BLOCK>
__device__ __forceinline__ int block_excl_scan(int v, int* smem, int& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = lane < (BLOCK / 32) ? smem[lane] : 0;
    int si = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, si, o);
      if (lane >= o) si += t;
    }
    smem[lane] = si - s;
    if (lane == 31) smem[32] = si;
  }
  __syncthreads();
  int res = smem[w] + inc - v;
  total = smem[32];
  __syncthreads();
  return res;
}

}  // namespace dfb
