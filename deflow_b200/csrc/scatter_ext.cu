// Drop-in kernels for the three mmcv._ext functions on the DeFlow path
// (OpenSceneFlow/assets/cuda/mmcv/pybind.cpp:32-50) -- generic N x C / arbitrary int32 coords.
//
//   dynamic_voxelize_forward        voxelization_cuda_kernel.cuh:13-50
//   dynamic_point_to_voxel_forward  scatter_points_cuda.cu:9-66   (unique_dim + atomic reduce)
//   dynamic_point_to_voxel_backward scatter_points_cuda.cu:68-132
//
// The forward is sort-free: bitmap occupancy over the coordinate extent, popcount-scan ranks
// (== unique_dim's lexicographic order), a counting sort into CSR segments and atomic-free
// segment reductions with coalesced row reads.
#include "common.cuh"
#include "../../include/deflow_b200.h"

namespace dfb {

VoxelParams make_voxel_params(const float* vs, const float* rng);  // pillar_index.cu

// ---------------------------------------------------------------- single-launch scans of the generic (per-call) path
// One 1024-thread block per frame: (a) exclusive scan of the per-block valid counts,
// (b) exclusive scan of the bitmap popcounts.
__global__ void __launch_bounds__(1024) k_scan_frame_1(const unsigned* __restrict__ bitmap, int Wd,
                                                     int* __restrict__ blk_cnt, int nblk,
                                                     int* __restrict__ word_rank, int* __restrict__ counts, int F) {
  __shared__ int sm[33];
  const int f = blockIdx.x;
  int carry = 0;
  for (int base = 0; base < nblk; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nblk ? blk_cnt[f * nblk + i] : 0;
    int tot;
    const int ex = block_excl_scan<1024>(v, sm, tot);
    if (i < nblk) blk_cnt[f * nblk + i] = carry + ex;  // in place: becomes the block offset
    carry += tot;
  }
  if (threadIdx.x == 0) counts[f] = carry;  // n_valid[f]
  carry = 0;
  for (int base = 0; base < Wd; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < Wd ? __popc(bitmap[(size_t)f * Wd + i]) : 0;
    int tot;
    const int ex = block_excl_scan<1024>(v, sm, tot);
    if (i < Wd) word_rank[(size_t)f * Wd + i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) counts[F + f] = carry;  // n_pil[f]
}

// frame offsets.  counts = n_valid[F] | n_pil[F] | pt_off[F+1] | pil_off[F+1]
__global__ void k_frame_offsets_1(int* __restrict__ counts, int F) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int a = 0, b = 0;
    for (int f = 0; f < F; ++f) {
      counts[2 * F + f] = a;
      counts[3 * F + 1 + f] = b;
      a += counts[f];
      b += counts[F + f];
    }
    counts[2 * F + F] = a;
    counts[3 * F + 1 + F] = b;
  }
}

// ---------------------------------------------------------------- CSR offsets of the voxels
// One block per frame; pil_start[q] = pt_off[f] + exclusive scan of the counts of that frame.
__global__ void __launch_bounds__(1024) k_pillar_scan_1(const int* __restrict__ pil_cnt, const int* __restrict__ counts,
                                                      int F, int* __restrict__ pil_start) {
  __shared__ int sm[33];
  const int f = blockIdx.x;
  const int q0 = counts[3 * F + 1 + f], q1 = counts[3 * F + 1 + f + 1];
  int carry = counts[2 * F + f];
  for (int base = q0; base < q1; base += 1024) {
    const int q = base + threadIdx.x;
    const int v = q < q1 ? pil_cnt[q] : 0;
    int tot;
    const int ex = block_excl_scan<1024>(v, sm, tot);
    if (q < q1) pil_start[q] = carry + ex;
    carry += tot;
  }
  if (f == F - 1 && threadIdx.x == 0) pil_start[q1] = counts[2 * F + F];
}


// ---------------------------------------------------------------- dynamic_voxelize_forward
__global__ void __launch_bounds__(256) k_dynamic_voxelize(const float* __restrict__ points, int n, int nf,
                                                          VoxelParams P, int* __restrict__ coors) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* p = points + (size_t)i * nf;
    int* c = coors + (size_t)i * 3;
    int cx, cy, cz;
    const int st = voxel_coords(p[0], p[1], p[2], P, cx, cy, cz);
    // partial -1 pattern of the reference (voxelization_cuda_kernel.cuh:27-43): untouched
    // columns keep the caller's zero initialisation
    if (st == 1) { c[0] = -1; }
    else if (st == 2) { c[0] = -1; c[1] = -1; }
    else if (st == 3) { c[0] = -1; c[1] = -1; c[2] = -1; }
    else { c[0] = cz; c[1] = cy; c[2] = cx; }
  }
}

// ---------------------------------------------------------------- scatter index (generic coords)
constexpr int SX_BLOCK = 256, SX_ITEMS = 4, SX_CHUNK = SX_BLOCK * SX_ITEMS;

__global__ void __launch_bounds__(SX_BLOCK) k_mark_coors(const int* __restrict__ coors, int n, int Ez, int Ey, int Ex,
                                                         int* __restrict__ keys, unsigned* __restrict__ bitmap,
                                                         int* __restrict__ blk_cnt) {
  int total = 0;
#pragma unroll
  for (int j = 0; j < SX_ITEMS; ++j) {
    const int i = blockIdx.x * SX_CHUNK + j * SX_BLOCK + threadIdx.x;
    int key = -1;
    if (i < n) {
      const int z = coors[3 * (size_t)i], y = coors[3 * (size_t)i + 1], x = coors[3 * (size_t)i + 2];
      // rows with ANY negative component are invalid (scatter_points_cuda.cu:24)
      if (z >= 0 && y >= 0 && x >= 0 && z < Ez && y < Ey && x < Ex) {
        key = (z * Ey + y) * Ex + x;
        atomicOr(&bitmap[key >> 5], 1u << (key & 31));
      }
      keys[i] = key;
    }
    total += __syncthreads_count(key >= 0);
  }
  if (threadIdx.x == 0) blk_cnt[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) k_rank_coors(int n, int Ey, int Ex, const unsigned* __restrict__ bitmap,
                                                    const int* __restrict__ word_rank, int* __restrict__ map,
                                                    int* __restrict__ slot, int* __restrict__ cnt,
                                                    int* __restrict__ voxel_coors) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int key = map[i];  // holds the key on entry
    if (key < 0) continue;
    const unsigned word = bitmap[key >> 5];
    const int r = word_rank[key >> 5] + __popc(word & ((1u << (key & 31)) - 1u));
    map[i] = r;
    const int s = atomicAdd(&cnt[r], 1);
    slot[i] = s;
    if (s == 0) {
      const int x = key % Ex, t = key / Ex;
      voxel_coors[3 * (size_t)r] = t / Ey;
      voxel_coors[3 * (size_t)r + 1] = t % Ey;
      voxel_coors[3 * (size_t)r + 2] = x;
    }
  }
}

__global__ void __launch_bounds__(256) k_fill_csr_generic(int n, const int* __restrict__ map,
                                                          const int* __restrict__ slot,
                                                          const int* __restrict__ pil_start,
                                                          int* __restrict__ sorted_pt) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = map[i];
    if (r >= 0) sorted_pt[pil_start[r] + slot[i]] = i;
  }
}

// ---------------------------------------------------------------- segment reductions
// MODE 0 sum, 1 mean, 2 max.  Warp per pillar; lane = channel (stride 32): every point row is one
// coalesced read.
template <int MODE>
__global__ void __launch_bounds__(256) k_scatter_reduce_wide(const float* __restrict__ feats, int c,
                                                             const int* __restrict__ pil_start,
                                                             const int* __restrict__ sorted_pt,
                                                             const int* __restrict__ num_voxels, int max_voxels,
                                                             float* __restrict__ out) {
  const int M = min(*num_voxels, max_voxels);
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < M; q += warps) {
    const int s0 = pil_start[q], s1 = pil_start[q + 1];
    for (int ch = lane; ch < c; ch += 32) {
      float acc = MODE == 2 ? -INFINITY : 0.f;
      int j = s0;
      // two independent gathers in flight per lane
      for (; j + 1 < s1; j += 2) {
        const int p0 = sorted_pt[j], p1 = sorted_pt[j + 1];
        const float v0 = feats[(size_t)p0 * c + ch], v1 = feats[(size_t)p1 * c + ch];
        if (MODE == 2) acc = fmaxf(acc, fmaxf(v0, v1)); else acc += v0 + v1;
      }
      if (j < s1) {
        const float v0 = feats[(size_t)sorted_pt[j] * c + ch];
        if (MODE == 2) acc = fmaxf(acc, v0); else acc += v0;
      }
      if (MODE == 1) acc = __fdiv_rn(acc, (float)(s1 - s0));
      out[(size_t)q * c + ch] = acc;
    }
  }
}

// c <= 4: 8 lanes per pillar striding over its points, channels in registers.
template <int MODE>
__global__ void __launch_bounds__(256) k_scatter_reduce_narrow(const float* __restrict__ feats, int c,
                                                               const int* __restrict__ pil_start,
                                                               const int* __restrict__ sorted_pt,
                                                               const int* __restrict__ num_voxels, int max_voxels,
                                                               float* __restrict__ out) {
  const int M = min(*num_voxels, max_voxels);
  const int sub = threadIdx.x & 7;
  const int groups = (gridDim.x * blockDim.x) >> 3;
  for (int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; q < ((M + 3) & ~3); q += groups) {
    float a[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = MODE == 2 ? -INFINITY : 0.f;
    int s0 = 0, s1 = 0;
    if (q < M) { s0 = pil_start[q]; s1 = pil_start[q + 1]; }
    for (int j = s0 + sub; j < s1; j += 8) {
      const float* row = feats + (size_t)sorted_pt[j] * c;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < c) { if (MODE == 2) a[k] = fmaxf(a[k], row[k]); else a[k] += row[k]; }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float t = __shfl_xor_sync(0xffffffffu, a[k], o);
        if (MODE == 2) a[k] = fmaxf(a[k], t); else a[k] += t;
      }
    if (sub == 0 && q < M) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < c) out[(size_t)q * c + k] = MODE == 1 ? __fdiv_rn(a[k], (float)(s1 - s0)) : a[k];
    }
  }
}

// ---------------------------------------------------------------- backward
template <int MODE>
__global__ void __launch_bounds__(256) k_scatter_bwd(float* __restrict__ gf, const float* __restrict__ gr,
                                                     const int* __restrict__ map, const int* __restrict__ cnt,
                                                     long long total, int c) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / c), ch = (int)(e - (long long)i * c);
    const int r = map[i];
    float g = 0.f;
    if (r >= 0) {
      g = gr[(size_t)r * c + ch];
      if (MODE == 1) g = __fdiv_rn(g, (float)cnt[r]);
    }
    gf[e] = g;
  }
}

// max: the smallest point index attaining the maximum receives the gradient
// (scatter_points_cuda_kernel.cuh:143-185): atomicMin trace-back, then scatter.
__global__ void __launch_bounds__(256) k_max_traceback(const float* __restrict__ feats,
                                                       const float* __restrict__ red, const int* __restrict__ map,
                                                       int* __restrict__ reduce_from, long long total, int c) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / c), ch = (int)(e - (long long)i * c);
    const int r = map[i];
    if (r >= 0 && feats[e] == red[(size_t)r * c + ch]) atomicMin(&reduce_from[(size_t)r * c + ch], i);
  }
}

__global__ void __launch_bounds__(256) k_max_scatter_grad(float* __restrict__ gf, const float* __restrict__ gr,
                                                          const int* __restrict__ reduce_from, long long total_mc,
                                                          int c, int n) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total_mc;
       e += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(e % c);
    const int src = reduce_from[e];
    if (src < n) gf[(size_t)src * c + ch] = gr[e];
  }
}

__global__ void __launch_bounds__(256) k_fill_int(int* __restrict__ p, long long n, int v) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    p[e] = v;
}

}  // namespace dfb

using namespace dfb;

static int grid_for(long long work, int block, int cap_mult = 16) {
  long long b = (work + block - 1) / block;
  const long long cap = (long long)sm_count() * cap_mult;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" int dfb_dynamic_voxelize_forward(const float* points, int n, int num_features, const float* voxel_size,
                                            const float* range, int* coors, void* stream_) {
  if (n < 0 || num_features < 3) { set_error("dynamic_voxelize_forward: points must be [n, >=3]"); return DFB_ERR_ARG; }
  if (n == 0) return DFB_OK;
  const VoxelParams P = make_voxel_params(voxel_size, range);
  k_dynamic_voxelize<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream_>>>(points, n, num_features, P, coors);
  add_launches(1);
  return check_launch("dfb_dynamic_voxelize_forward");
}

extern "C" int dfb_scatter_index(const int* coors, int n, const int* extent_zyx, unsigned* bitmap, int* word_rank,
                                 int* blk, int* slot, int* point2voxel_map, int* voxel_coors,
                                 int* voxel_points_count, int* pil_start, int* sorted_pt, int* counts6,
                                 void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n <= 0) { set_error("dfb_scatter_index: n must be > 0 (the n == 0 case is handled by the caller)"); return DFB_ERR_ARG; }
  const long long cells = (long long)extent_zyx[0] * extent_zyx[1] * extent_zyx[2];
  if (cells <= 0 || cells >= (1ll << 31)) { set_error("dfb_scatter_index: extent %lld cells unsupported", cells); return DFB_ERR_UNSUPPORTED; }
  const int Wd = (int)((cells + 31) / 32);
  const int nblk = (n + SX_CHUNK - 1) / SX_CHUNK;
  cudaMemsetAsync(bitmap, 0, sizeof(unsigned) * (size_t)Wd, st);
  cudaMemsetAsync(voxel_points_count, 0, sizeof(int) * (size_t)n, st);
  k_mark_coors<<<nblk, SX_BLOCK, 0, st>>>(coors, n, extent_zyx[0], extent_zyx[1], extent_zyx[2], point2voxel_map,
                                          bitmap, blk);
  k_scan_frame_1<<<1, 1024, 0, st>>>(bitmap, Wd, blk, nblk, word_rank, counts6, 1);
  k_frame_offsets_1<<<1, 32, 0, st>>>(counts6, 1);
  k_rank_coors<<<grid_for(n, 256), 256, 0, st>>>(n, extent_zyx[1], extent_zyx[2], bitmap, word_rank, point2voxel_map,
                                                 slot, voxel_points_count, voxel_coors);
  k_pillar_scan_1<<<1, 1024, 0, st>>>(voxel_points_count, counts6, 1, pil_start);
  k_fill_csr_generic<<<grid_for(n, 256), 256, 0, st>>>(n, point2voxel_map, slot, pil_start, sorted_pt);
  add_launches(6);
  return check_launch("dfb_scatter_index");
}

extern "C" int dfb_scatter_reduce(const float* feats, int n, int c, const int* pil_start, const int* sorted_pt,
                                  const int* num_voxels, int max_voxels, int reduce_type, float* voxel_feats,
                                  void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (reduce_type < 0 || reduce_type > 2) { set_error("do not support reduce type %d", reduce_type); return DFB_ERR_ARG; }
  if (n <= 0 || c <= 0 || max_voxels <= 0) return DFB_OK;
  if (c <= 4) {
    const int g = grid_for((long long)max_voxels * 8, 256, 8);
    if (reduce_type == 0) k_scatter_reduce_narrow<0><<<g, 256, 0, st>>>(feats, c, pil_start, sorted_pt, num_voxels, max_voxels, voxel_feats);
    else if (reduce_type == 1) k_scatter_reduce_narrow<1><<<g, 256, 0, st>>>(feats, c, pil_start, sorted_pt, num_voxels, max_voxels, voxel_feats);
    else k_scatter_reduce_narrow<2><<<g, 256, 0, st>>>(feats, c, pil_start, sorted_pt, num_voxels, max_voxels, voxel_feats);
  } else {
    const int g = grid_for((long long)max_voxels * 32, 256, 8);
    if (reduce_type == 0) k_scatter_reduce_wide<0><<<g, 256, 0, st>>>(feats, c, pil_start, sorted_pt, num_voxels, max_voxels, voxel_feats);
    else if (reduce_type == 1) k_scatter_reduce_wide<1><<<g, 256, 0, st>>>(feats, c, pil_start, sorted_pt, num_voxels, max_voxels, voxel_feats);
    else k_scatter_reduce_wide<2><<<g, 256, 0, st>>>(feats, c, pil_start, sorted_pt, num_voxels, max_voxels, voxel_feats);
  }
  add_launches(1);
  return check_launch("dfb_scatter_reduce");
}

extern "C" int dfb_dynamic_point_to_voxel_backward(float* grad_feats, const float* grad_reduced_feats,
                                                   const float* feats, const float* reduced_feats,
                                                   const int* point2voxel_map, const int* voxel_points_count, int n,
                                                   int m, int c, int reduce_type, int* workspace_mc, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (reduce_type < 0 || reduce_type > 2) { set_error("do not support reduce type %d", reduce_type); return DFB_ERR_ARG; }
  if (n <= 0 || c <= 0) return DFB_OK;
  const long long total = (long long)n * c;
  if (m <= 0) { cudaMemsetAsync(grad_feats, 0, sizeof(float) * (size_t)total, st); return check_launch("scatter_backward"); }
  const int g = grid_for(total, 256);
  if (reduce_type == 0) k_scatter_bwd<0><<<g, 256, 0, st>>>(grad_feats, grad_reduced_feats, point2voxel_map, voxel_points_count, total, c);
  else if (reduce_type == 1) k_scatter_bwd<1><<<g, 256, 0, st>>>(grad_feats, grad_reduced_feats, point2voxel_map, voxel_points_count, total, c);
  else {
    if (!workspace_mc) { set_error("scatter backward (max) needs an int32[m*c] workspace"); return DFB_ERR_ARG; }
    const long long mc = (long long)m * c;
    cudaMemsetAsync(grad_feats, 0, sizeof(float) * (size_t)total, st);
    k_fill_int<<<grid_for(mc, 256), 256, 0, st>>>(workspace_mc, mc, n);
    k_max_traceback<<<g, 256, 0, st>>>(feats, reduced_feats, point2voxel_map, workspace_mc, total, c);
    k_max_scatter_grad<<<grid_for(mc, 256), 256, 0, st>>>(grad_feats, grad_reduced_feats, workspace_mc, mc, c, n);
  }
  add_launches(reduce_type == 2 ? 3 : 1);
  return check_launch("dfb_dynamic_point_to_voxel_backward");
}
