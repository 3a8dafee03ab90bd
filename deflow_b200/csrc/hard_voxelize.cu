// hard_voxelize_forward: fixed-capacity voxelisation (max_points per voxel, max_voxels) with the reference's deterministic
// semantics, without its O(N^2) search and its single-thread pass.
//
// Reference: HardVoxelizeForwardCUDAKernelLauncher (OpenSceneFlow/assets/cuda/mmcv/voxelization_cuda.cu:8-148) and its
// kernels (voxelization_cuda_kernel.cuh:52-170): per point the number of EARLIER points in the same voxel (a linear scan
// over all previous points, point_to_voxelidx_kernel) and a <<<1,1>>> loop over the points that numbers the voxels in order
// of first appearance (determin_voxel_num).  Restated: a voxel's id is the rank of its first point among all first points;
// ids >= max_voxels are dropped; a point is kept when its rank inside its voxel is < max_points; voxels[id][rank] =
// features, coors[id] = voxel coordinates (z, y, x), num_points_per_voxel[id] = min(count, max_points).
//
// Here the unique-voxel index of the dynamic scatter (dfb_scatter_index: CSR point lists per voxel) already groups the
// points; a point's rank is the number of smaller indices in its own list, "first point" flags are scanned over the points
// (three launches, the pattern of csrc/collate.cu) and one pass writes the outputs.  Integer work, bit-exact.
#include "common.cuh"
#include "../../include/deflow_b200.h"

namespace dfb {

constexpr int HV_THREADS = 256;
constexpr int HV_PER = 4;
constexpr int HV_CHUNK = HV_THREADS * HV_PER;

// rank[i] = number of points j < i in the same voxel (-1 for invalid points); chunk_cnt[c] = first points in chunk c
__global__ void __launch_bounds__(HV_THREADS) k_hv_rank(const int* __restrict__ cmap, int n, const int* __restrict__ pil_start,
                                                        const int* __restrict__ sorted_pt, int* __restrict__ rank,
                                                        int* __restrict__ chunk_cnt) {
  int firsts = 0;
#pragma unroll
  for (int k = 0; k < HV_PER; ++k) {
    const int i = blockIdx.x * HV_CHUNK + threadIdx.x * HV_PER + k;
    if (i >= n) continue;
    const int v = cmap[i];
    int r = -1;
    if (v >= 0) {
      r = 0;
      const int a = pil_start[v], b = pil_start[v + 1];
      for (int p = a; p < b; ++p) r += (sorted_pt[p] < i);
      firsts += (r == 0);
    }
    rank[i] = r;
  }
  __shared__ int red[HV_THREADS / 32];
  for (int o = 16; o > 0; o >>= 1) firsts += __shfl_xor_sync(0xffffffffu, firsts, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = firsts;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < HV_THREADS / 32; ++w) s += red[w];
    chunk_cnt[blockIdx.x] = s;
  }
}

// single block: exclusive scan of the chunk counts; voxel_num = min(total, max_voxels)
__global__ void __launch_bounds__(1024) k_hv_scan(int* __restrict__ chunk_cnt, int n_chunks, int max_voxels, int* __restrict__ voxel_num) {
  __shared__ int smem[33];
  int run = 0;
  for (int c0 = 0; c0 < n_chunks; c0 += 1024) {
    const int c = c0 + threadIdx.x;
    const int v = c < n_chunks ? chunk_cnt[c] : 0;
    int total;
    const int ex = block_excl_scan<1024>(v, smem, total);
    if (c < n_chunks) chunk_cnt[c] = run + ex;
    run += total;
  }
  if (threadIdx.x == 0) voxel_num[0] = min(run, max_voxels);
}

// voxel id of every first point -> vid_of_voxel[sorted-unique voxel] (or -1 beyond max_voxels), coors and counts of kept voxels
__global__ void __launch_bounds__(HV_THREADS) k_hv_number(const int* __restrict__ cmap, const int* __restrict__ rank, int n,
                                                          const int* __restrict__ chunk_off, const int* __restrict__ coors_in,
                                                          const int* __restrict__ pil_start, int max_points, int max_voxels,
                                                          int* __restrict__ vid_of_voxel, int* __restrict__ coors_out,
                                                          int* __restrict__ num_points_per_voxel) {
  __shared__ int smem[33];
  const int i0 = blockIdx.x * HV_CHUNK + threadIdx.x * HV_PER;
  int mine = 0;
  bool first[HV_PER];
#pragma unroll
  for (int k = 0; k < HV_PER; ++k) {
    first[k] = (i0 + k < n) && rank[i0 + k] == 0;
    mine += first[k];
  }
  int total;
  int id = chunk_off[blockIdx.x] + block_excl_scan<HV_THREADS>(mine, smem, total);
#pragma unroll
  for (int k = 0; k < HV_PER; ++k) {
    if (!first[k]) continue;
    const int i = i0 + k, v = cmap[i];
    if (id < max_voxels) {
      vid_of_voxel[v] = id;
      coors_out[3 * id] = coors_in[3 * i]; coors_out[3 * id + 1] = coors_in[3 * i + 1]; coors_out[3 * id + 2] = coors_in[3 * i + 2];
      num_points_per_voxel[id] = min(pil_start[v + 1] - pil_start[v], max_points);
    } else {
      vid_of_voxel[v] = -1;
    }
    ++id;
  }
}

__global__ void __launch_bounds__(256) k_hv_assign(const float* __restrict__ points, int n, int c, const int* __restrict__ cmap,
                                                   const int* __restrict__ rank, const int* __restrict__ vid_of_voxel,
                                                   int max_points, float* __restrict__ voxels) {
  const long long total = (long long)n * c;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / c), k = (int)(e % c);
    const int r = rank[i];
    if (r < 0 || r >= max_points) continue;
    const int id = vid_of_voxel[cmap[i]];
    if (id < 0) continue;
    voxels[((size_t)id * max_points + r) * c + k] = points[e];
  }
}

}  // namespace dfb

using namespace dfb;

extern "C" long long dfb_hard_voxelize_workspace(int n) {
  const int chunks = (n + HV_CHUNK - 1) / HV_CHUNK;
  return (long long)n /*rank*/ + (chunks > 0 ? chunks : 1) + n /*vid_of_voxel (<= n voxels)*/;
}

extern "C" int dfb_hard_voxelize_assign(const float* points, int n, int num_features, const int* coors, const int* point2voxel_map,
                                        const int* pil_start, const int* sorted_pt, int max_points, int max_voxels,
                                        float* voxels, int* voxel_coors, int* num_points_per_voxel, int* voxel_num,
                                        int* workspace, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n <= 0 || num_features < 3 || max_points <= 0 || max_voxels <= 0) { set_error("dfb_hard_voxelize_assign: bad sizes"); return DFB_ERR_ARG; }
  const int chunks = (n + HV_CHUNK - 1) / HV_CHUNK;
  int* rank = workspace;
  int* chunk_cnt = workspace + n;
  int* vid = chunk_cnt + chunks;
  k_hv_rank<<<chunks, HV_THREADS, 0, st>>>(point2voxel_map, n, pil_start, sorted_pt, rank, chunk_cnt);
  k_hv_scan<<<1, 1024, 0, st>>>(chunk_cnt, chunks, max_voxels, voxel_num);
  k_hv_number<<<chunks, HV_THREADS, 0, st>>>(point2voxel_map, rank, n, chunk_cnt, coors, pil_start, max_points, max_voxels, vid,
                                            voxel_coors, num_points_per_voxel);
  long long blocks = ((long long)n * num_features + 255) / 256;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  k_hv_assign<<<(int)blocks, 256, 0, st>>>(points, n, num_features, point2voxel_map, rank, vid, max_points, voxels);
  add_launches(4);
  return check_launch("dfb_hard_voxelize_assign");
}
