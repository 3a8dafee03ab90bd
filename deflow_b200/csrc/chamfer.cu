// Chamfer nearest-neighbour search between two point clouds, both directions, forward + backward.
//
// Reference: NmDistanceKernel / NmDistanceGradKernel (OpenSceneFlow/assets/cuda/chamfer3D/chamfer3D.cu:33-124), bound as
// chamfer3D.forward / .backward (chamfer3D_cuda.cpp) and used by seflowLoss (OpenSceneFlow/src/lossfuncs.py:22-100):
//   dist0[i] = min_j |pc0[i] - pc1[j]|^2  (squared, fp32),  idx0[i] = the LOWEST j attaining it (strict `<` scan);
//   same for pc1 against pc0.  backward: grad_pc0[i] += 2 g0[i] (pc0[i] - pc1[idx0[i]]), grad_pc1[idx0[i]] -= the same,
//   and symmetrically for the other direction.
// The reference runs one thread per query over the whole target cloud (one block per 256 queries; queries of a partial
// last block read target tiles that the exited threads never staged -- undefined there, defined here).  Here the
// (query block x target slice) plane is tiled over ~4 CTAs per SM: a CTA stages its target slice through shared memory as
// float4 (one broadcast LDS.128 per target), every thread scans it for TWO queries, and the partial minima meet in one
// 64-bit atomicMin per query on the key (distance bits << 32 | index): non-negative floats order like their bit patterns,
// so the minimum key is the smallest distance and, among equal distances, the smallest index -- the reference's
// tie-break.  Distance arithmetic is pinned to what nvcc makes of the reference's expression (checked in its SASS: FMUL on
// dy, then FFMA with dx, then FFMA with dz): fma(dz,dz, fma(dx,dx, dy*dy)).
#include "common.cuh"
#include "../../include/deflow_b200.h"

namespace dfb {

constexpr int CH_THREADS = 128;
constexpr int CH_Q = 2;                       // queries per thread
constexpr int CH_TILE = 1024;                 // targets staged per shared-memory tile (16 KB as float4)

__global__ void __launch_bounds__(CH_THREADS) k_chamfer_nn(const float* __restrict__ q_xyz, int nq, const float* __restrict__ t_xyz,
                                                           int nt, int slice, unsigned long long* __restrict__ best_key) {
  __shared__ float4 tile[CH_TILE];
  const int q0 = (blockIdx.x * CH_THREADS + threadIdx.x) * CH_Q;
  float qx[CH_Q], qy[CH_Q], qz[CH_Q], best[CH_Q];
  int best_i[CH_Q];
#pragma unroll
  for (int k = 0; k < CH_Q; ++k) {
    const int q = q0 + k;
    const bool ok = q < nq;
    qx[k] = ok ? q_xyz[3 * (size_t)q] : 0.f; qy[k] = ok ? q_xyz[3 * (size_t)q + 1] : 0.f; qz[k] = ok ? q_xyz[3 * (size_t)q + 2] : 0.f;
    best[k] = 1e20f; best_i[k] = -1;            // chamfer3D.cu:45-46
  }
  const int t_lo = blockIdx.y * slice, t_hi = min(nt, t_lo + slice);
  for (int base = t_lo; base < t_hi; base += CH_TILE) {
    const int cnt = min(CH_TILE, t_hi - base);
    for (int j = threadIdx.x; j < cnt; j += CH_THREADS) {
      const float* p = t_xyz + 3 * (size_t)(base + j);
      tile[j] = make_float4(p[0], p[1], p[2], 0.f);
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const float4 t = tile[j];
#pragma unroll
      for (int k = 0; k < CH_Q; ++k) {
        const float dx = __fsub_rn(t.x, qx[k]), dy = __fsub_rn(t.y, qy[k]), dz = __fsub_rn(t.z, qz[k]);
        const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
        if (d < best[k]) { best[k] = d; best_i[k] = base + j; }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < CH_Q; ++k) {
    const int q = q0 + k;
    if (q < nq && best_i[k] >= 0) {
      const unsigned long long key = ((unsigned long long)__float_as_uint(best[k]) << 32) | (unsigned)best_i[k];
      atomicMin(best_key + q, key);
    }
  }
}

__global__ void __launch_bounds__(256) k_chamfer_init(unsigned long long* __restrict__ key, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // (1e20f, -1): what the reference returns for an empty target cloud; any real candidate has a smaller key
  if (i < n) key[i] = ((unsigned long long)__float_as_uint(1e20f) << 32) | 0xffffffffull;
}

__global__ void __launch_bounds__(256) k_chamfer_unpack(const unsigned long long* __restrict__ key, int n, float* __restrict__ dist,
                                                        int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const unsigned long long k = key[i];
    dist[i] = __uint_as_float((unsigned)(k >> 32));
    idx[i] = (int)(unsigned)(k & 0xffffffffull);
  }
}

// one direction of the backward (chamfer3D.cu:92-114): the query's own gradient is written once (no atomic needed: the
// buffer was zeroed and every query appears once per direction), the matched target's gradient is a float reduction
__global__ void __launch_bounds__(256) k_chamfer_grad(const float* __restrict__ q_xyz, int nq, const float* __restrict__ t_xyz, int nt,
                                                      const float* __restrict__ grad_dist, const int* __restrict__ idx,
                                                      float* __restrict__ grad_q, float* __restrict__ grad_t) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += gridDim.x * blockDim.x) {
    const int j = idx[i];
    if (j < 0 || j >= nt) continue;
    const float g = grad_dist[i] * 2.f;
    const float gx = g * (q_xyz[3 * (size_t)i] - t_xyz[3 * (size_t)j]);
    const float gy = g * (q_xyz[3 * (size_t)i + 1] - t_xyz[3 * (size_t)j + 1]);
    const float gz = g * (q_xyz[3 * (size_t)i + 2] - t_xyz[3 * (size_t)j + 2]);
    atomicAdd(grad_q + 3 * (size_t)i, gx); atomicAdd(grad_q + 3 * (size_t)i + 1, gy); atomicAdd(grad_q + 3 * (size_t)i + 2, gz);
    atomicAdd(grad_t + 3 * (size_t)j, -gx); atomicAdd(grad_t + 3 * (size_t)j + 1, -gy); atomicAdd(grad_t + 3 * (size_t)j + 2, -gz);
  }
}

static void nn_one_direction(const float* q, int nq, const float* t, int nt, unsigned long long* key, float* dist, int* idx,
                             cudaStream_t st) {
  if (nq <= 0) return;
  k_chamfer_init<<<(nq + 255) / 256, 256, 0, st>>>(key, nq);
  if (nt > 0) {
    const int qblocks = (nq + CH_THREADS * CH_Q - 1) / (CH_THREADS * CH_Q);
    // enough target slices for ~4 CTAs per SM, each at least one tile long
    int slices = (sm_count() * 4 + qblocks - 1) / qblocks;
    const int max_slices = (nt + CH_TILE - 1) / CH_TILE;
    if (slices > max_slices) slices = max_slices;
    if (slices < 1) slices = 1;
    int slice = (nt + slices - 1) / slices;
    slice = (slice + CH_TILE - 1) / CH_TILE * CH_TILE;
    slices = (nt + slice - 1) / slice;
    dim3 g(qblocks, slices);
    k_chamfer_nn<<<g, CH_THREADS, 0, st>>>(q, nq, t, nt, slice, key);
  }
  k_chamfer_unpack<<<(nq + 255) / 256, 256, 0, st>>>(key, nq, dist, idx);
  add_launches(nt > 0 ? 3 : 2);
}

}  // namespace dfb

using namespace dfb;

extern "C" int dfb_chamfer_forward(const float* pc0, int n0, const float* pc1, int n1, float* dist0, float* dist1, int* idx0,
                                   int* idx1, unsigned long long* workspace, void* stream_) {
  if (n0 < 0 || n1 < 0 || !workspace) { set_error("dfb_chamfer_forward: bad arguments"); return DFB_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream_;
  nn_one_direction(pc0, n0, pc1, n1, workspace, dist0, idx0, st);
  nn_one_direction(pc1, n1, pc0, n0, workspace + n0, dist1, idx1, st);
  return check_launch("dfb_chamfer_forward");
}

extern "C" int dfb_chamfer_backward(const float* pc0, int n0, const float* pc1, int n1, const int* idx0, const int* idx1,
                                    const float* grad_dist0, const float* grad_dist1, float* grad_pc0, float* grad_pc1,
                                    void* stream_) {
  if (n0 < 0 || n1 < 0) { set_error("dfb_chamfer_backward: bad sizes"); return DFB_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream_;
  if (n0 > 0) cudaMemsetAsync(grad_pc0, 0, sizeof(float) * 3 * (size_t)n0, st);
  if (n1 > 0) cudaMemsetAsync(grad_pc1, 0, sizeof(float) * 3 * (size_t)n1, st);
  const int cap = sm_count() * 8;
  if (n0 > 0 && n1 > 0) {
    int b0 = (n0 + 255) / 256, b1 = (n1 + 255) / 256;
    if (b0 > cap) b0 = cap;
    if (b1 > cap) b1 = cap;
    k_chamfer_grad<<<b0, 256, 0, st>>>(pc0, n0, pc1, n1, grad_dist0, idx0, grad_pc0, grad_pc1);
    k_chamfer_grad<<<b1, 256, 0, st>>>(pc1, n1, pc0, n0, grad_dist1, idx1, grad_pc1, grad_pc0);
    add_launches(2);
  }
  return check_launch("dfb_chamfer_backward");
}
