// Sort-free, bit-exact pillar indexing for a batch of NaN-padded frames.
//
// Replaces, for all 2B frames of a step at once, what the reference does per sample with
//   DynamicVoxelizer.forward              (OpenSceneFlow/src/models/basic/encoder.py:567-600)
//   dynamic_voxelize_kernel               (assets/cuda/mmcv/voxelization_cuda_kernel.cuh:13-50)
//   at::unique_dim + inverse + counts     (assets/cuda/mmcv/scatter_points_cuda.cu:24-37)
//
// A linear key (z*gy + y)*gx + x preserves unique_dim's lexicographic (z,y,x) order, so the rank
// of a cell among the occupied cells of its frame IS the reference's voxel id.  Occupancy lives
// in a bitmap (1 bit per cell: 32 KB per 512x512 frame, L2 resident); ranks come from a popcount
// scan; valid points are compacted in their original order (stable), which is what the
// reference's boolean indexing produces.  A counting sort (one integer atomic per point) builds a
// CSR list of the points of every pillar so that all later reductions are atomic-free segment sums.
#include "common.cuh"
#include <stdlib.h>
#include "../../include/deflow_b200.h"

namespace dfb {

constexpr int IDX_BLOCK = 256;
constexpr int IDX_ITEMS = 4;
constexpr int IDX_CHUNK = IDX_BLOCK * IDX_ITEMS;  // points per block

// ---------------------------------------------------------------- K1: occupancy bitmap + valid counts
// All twelve loads of a thread's four points are issued before the first use (the kernel is latency-bound: what
// matters is bytes in flight).  No per-point key array: K3 recomputes the voxel of a point from its coordinates.
__device__ __forceinline__ int point_key(float x, float y, float z, const VoxelParams& P) {
  // NaN rows are dropped before voxelisation (encoder.py:576-577)
  if (isnan(x) || isnan(y) || isnan(z)) return -1;
  int cx, cy, cz;
  if (voxel_coords(x, y, z, P, cx, cy, cz) != 0) return -1;
  return (cz * P.gy + cy) * P.gx + cx;
}

// The block's 1024 x (x, y, z) rows through shared memory: fully coalesced global loads (one sector fetched once; the
// strided per-point loads re-fetched every sector ~3x from L2, ncu r01), then conflict-free stride-3 shared reads.
// Thread t receives points t, t + 256, t + 512, t + 768 of the chunk; rows past Nmax read as NaN (= dropped).
__device__ __forceinline__ void load_chunk(const float* __restrict__ p, int Nmax, int stride, int blk, float* stage,
                                           float (&x)[IDX_ITEMS], float (&y)[IDX_ITEMS], float (&z)[IDX_ITEMS]) {
  const float qnan = __int_as_float(0x7fc00000);
  if (stride == 3) {
    const float* src = p + (size_t)blk * IDX_CHUNK * 3;
    const int nfl = min(IDX_CHUNK, Nmax - blk * IDX_CHUNK) * 3;
    if (nfl == IDX_CHUNK * 3 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k)
        reinterpret_cast<float4*>(stage)[k * IDX_BLOCK + threadIdx.x] = __ldg(reinterpret_cast<const float4*>(src) + k * IDX_BLOCK + threadIdx.x);
    } else {
#pragma unroll
      for (int k = 0; k < 3 * IDX_ITEMS; ++k) {
        const int e = k * IDX_BLOCK + threadIdx.x;
        stage[e] = e < nfl ? src[e] : qnan;
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < IDX_ITEMS; ++j) {
      const int e = 3 * (j * IDX_BLOCK + threadIdx.x);
      x[j] = stage[e]; y[j] = stage[e + 1]; z[j] = stage[e + 2];
    }
  } else {
#pragma unroll
    for (int j = 0; j < IDX_ITEMS; ++j) {
      const int i = blk * IDX_CHUNK + j * IDX_BLOCK + threadIdx.x;
      x[j] = y[j] = z[j] = qnan;
      if (i < Nmax) { x[j] = p[(size_t)i * stride]; y[j] = p[(size_t)i * stride + 1]; z[j] = p[(size_t)i * stride + 2]; }
    }
  }
}

__global__ void __launch_bounds__(IDX_BLOCK) k_mark_points(const float* __restrict__ pts, int Nmax, int stride,
                                                          VoxelParams P, int Wd, int nblk,
                                                          unsigned* __restrict__ bitmap, unsigned char* __restrict__ occ,
                                                          int* __restrict__ blk_cnt) {
  __shared__ int s_total;
  __shared__ __align__(16) float stage[IDX_CHUNK * 3];
  const int f = blockIdx.y, blk = blockIdx.x;
  const float* p = pts + (size_t)f * Nmax * stride;
  if (threadIdx.x == 0) s_total = 0;
  float x[IDX_ITEMS], y[IDX_ITEMS], z[IDX_ITEMS];
  load_chunk(p, Nmax, stride, blk, stage, x, y, z);
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < IDX_ITEMS; ++j) {
    const int key = point_key(x[j], y[j], z[j], P);
    if (key >= 0) {
      // occupancy: a plain byte store (every writer stores the same value, so no atomic is needed; the scan kernel packs
      // the bytes to bits) -- 5 M atomicOr per step were the bound of this kernel (ncu r01: 1.3 TB/s of DRAM traffic)
      if (occ) occ[((size_t)f * Wd << 5) + key] = 1;
      else atomicOr(&bitmap[(size_t)f * Wd + (key >> 5)], 1u << (key & 31));
      ++cnt;
    }
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_total, cnt);
  __syncthreads();
  if (threadIdx.x == 0) blk_cnt[f * nblk + blk] = s_total;
}

// ---------------------------------------------------------------- K2: bitmap popcount scan (two levels, one launch)
// One block per 1024-word segment of a frame's bitmap: exclusive scan of the popcounts inside the segment
// (word_rank holds segment-local ranks) and the segment total.  The LAST block to finish (threadfence + ticket)
// scans, warp per frame, the segment totals (-> seg_base) and the per-block valid counts of K1 (-> block offsets in
// place), then writes the frame offsets: counts = n_valid[F] | n_pil[F] | pt_off[F+1] | pil_off[F+1].
constexpr int SEG_WORDS = 1024;

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// 32 occupancy bytes (0 / 1) -> one bitmap word: per 4-byte group (x * 0x01020408) >> 24 gathers the four low bits
__device__ __forceinline__ unsigned pack_occ(const uint4& a, const uint4& b) {
  const unsigned w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  unsigned r = 0u;
#pragma unroll
  for (int k = 0; k < 8; ++k) r |= (((w[k] * 0x01020408u) >> 24) & 0xFu) << (4 * k);
  return r;
}

__global__ void __launch_bounds__(1024) k_bitmap_scan(unsigned* __restrict__ bitmap, const unsigned char* __restrict__ occ, int Wd, int S,
                                                      int* __restrict__ word_rank, int* __restrict__ seg_tot,
                                                      int* __restrict__ seg_base, int* __restrict__ blk_cnt, int nblk,
                                                      int* __restrict__ counts, int F, unsigned* __restrict__ ticket) {
  __shared__ int sm[33];
  __shared__ int is_last;
  const int s = blockIdx.x, f = blockIdx.y;
  const int w = s * SEG_WORDS + threadIdx.x;
  unsigned word = 0u;
  if (w < Wd) {
    if (occ) {
      const uint4* o = reinterpret_cast<const uint4*>(occ + (((size_t)f * Wd + w) << 5));
      word = pack_occ(__ldg(o), __ldg(o + 1));
      bitmap[(size_t)f * Wd + w] = word;
    } else {
      word = bitmap[(size_t)f * Wd + w];
    }
  }
  const int v = __popc(word);
  int tot;
  const int ex = block_excl_scan<1024>(v, sm, tot);
  if (w < Wd) word_rank[(size_t)f * Wd + w] = ex;
  if (threadIdx.x == 0) {
    seg_tot[f * S + s] = tot;
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == (unsigned)(S * F - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  for (int ff = wp; ff < F; ff += 32) {
    int carry = 0;
    for (int base = 0; base < S; base += 32) {
      const int i = base + lane;
      const int t = i < S ? __ldcg(&seg_tot[ff * S + i]) : 0;
      const int inc = warp_incl_scan(t, lane);
      if (i < S) seg_base[ff * S + i] = carry + inc - t;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) counts[F + ff] = carry;  // n_pil[f]
    carry = 0;
    for (int base = 0; base < nblk; base += 32) {
      const int i = base + lane;
      const int t = i < nblk ? blk_cnt[ff * nblk + i] : 0;
      const int inc = warp_incl_scan(t, lane);
      if (i < nblk) blk_cnt[ff * nblk + i] = carry + inc - t;  // in place: becomes the block offset
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) counts[ff] = carry;  // n_valid[f]
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int a = 0, b = 0;
    for (int ff = 0; ff < F; ++ff) {
      counts[2 * F + ff] = a;
      counts[3 * F + 1 + ff] = b;
      a += counts[ff];
      b += counts[F + ff];
    }
    counts[2 * F + F] = a;
    counts[3 * F + 1 + F] = b;
  }
}

// ---------------------------------------------------------------- K3: stable compaction + rank + slot
// Thread t owns points t, t + 256, t + 512, t + 768 of the block's 1024-point chunk (consecutive lanes = consecutive
// points, so the compacted rows a warp stores are contiguous).  All twelve coordinate loads are issued up front, the
// four stable destinations come from four ballots + one shared-memory table of per-(item, warp) totals (one barrier
// instead of twelve), and the rank gathers, counting-sort atomics and stores of the four points are independent.
__global__ void __launch_bounds__(IDX_BLOCK) k_compact(
    const float* __restrict__ pts, int Nmax, int stride, VoxelParams P, int Wd,
    int nblk, int F, int HW, const unsigned* __restrict__ bitmap, const int* __restrict__ word_rank,
    const int* __restrict__ seg_base, int S, const int* __restrict__ blk_off, const int* __restrict__ counts, float* __restrict__ pt_xyz,
    int* __restrict__ pt_coor, long long* __restrict__ pt_idx, float* __restrict__ pt_offs,
    int* __restrict__ pt_pillar, int* __restrict__ pt_slot, int* __restrict__ pil_cnt, int* __restrict__ pil_coor,
    int* __restrict__ pil_pix) {
  constexpr int NW = IDX_BLOCK / 32;
  __shared__ int wtot[IDX_ITEMS][NW];
  __shared__ __align__(16) float stage[IDX_CHUNK * 3];
  const int f = blockIdx.y, blk = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float* p = pts + (size_t)f * Nmax * stride;
  float x[IDX_ITEMS], y[IDX_ITEMS], z[IDX_ITEMS];
  load_chunk(p, Nmax, stride, blk, stage, x, y, z);
  __syncthreads();                       // every thread has its rows: the staging buffer becomes the store transposer
  float* wst = stage + w * 96;           // 32 rows x 3 words per warp
  int key[IDX_ITEMS];
  unsigned bal[IDX_ITEMS];
#pragma unroll
  for (int j = 0; j < IDX_ITEMS; ++j) {
    key[j] = point_key(x[j], y[j], z[j], P);
    bal[j] = __ballot_sync(0xffffffffu, key[j] >= 0);
    if (lane == 0) wtot[j][w] = __popc(bal[j]);
  }
  __syncthreads();
  const int pt_base = counts[2 * F + f] + blk_off[f * nblk + blk];
  const int pil_base = counts[3 * F + 1 + f];
  int dst[IDX_ITEMS];
  {
    int run = pt_base;
#pragma unroll
    for (int j = 0; j < IDX_ITEMS; ++j) {
      int before = 0, tot = 0;
#pragma unroll
      for (int ww = 0; ww < NW; ++ww) { const int t = wtot[j][ww]; before += ww < w ? t : 0; tot += t; }
      dst[j] = run + before + __popc(bal[j] & ((1u << lane) - 1u));
      run += tot;
    }
  }
  const float hx = P.vx / 2, hy = P.vy / 2, hz = P.vz / 2;  // voxel_size / 2 (encoder.py:519)
  // rank of the point's cell among the occupied cells of the frame = the reference's voxel id
  int q[IDX_ITEMS], slot[IDX_ITEMS];
#pragma unroll
  for (int j = 0; j < IDX_ITEMS; ++j) {
    q[j] = -1;
    if (key[j] >= 0) {
      const unsigned word = bitmap[(size_t)f * Wd + (key[j] >> 5)];
      q[j] = pil_base + seg_base[f * S + (key[j] >> 15)] + word_rank[(size_t)f * Wd + (key[j] >> 5)] +
             __popc(word & ((1u << (key[j] & 31)) - 1u));
    }
  }
#pragma unroll
  for (int j = 0; j < IDX_ITEMS; ++j) slot[j] = q[j] >= 0 ? atomicAdd(&pil_cnt[q[j]], 1) : 0;
  // the compacted rows of a warp-item are contiguous in the outputs: transpose each 3-word row array through 96 words of
  // shared memory so that a store instruction writes 128 contiguous bytes (the direct 12-byte-stride stores wrote every
  // sector three times: 3.1x L1->L2 write traffic, ncu r01)
#pragma unroll
  for (int j = 0; j < IDX_ITEMS; ++j) {
    const bool ok = key[j] >= 0;
    const int nrow = __popc(bal[j]), r = __popc(bal[j] & ((1u << lane) - 1u));
    const int d = dst[j];
    const int d0 = __shfl_sync(0xffffffffu, d - r, 0);       // destination of the warp-item's first valid row
    const int kk = ok ? key[j] : 0;
    const int cx = kk % P.gx, t = kk / P.gx, cy = t % P.gy, cz = t / P.gy;
    if (ok) { wst[3 * r] = x[j]; wst[3 * r + 1] = y[j]; wst[3 * r + 2] = z[j]; }
    __syncwarp();
    for (int e = lane; e < 3 * nrow; e += 32) pt_xyz[3 * (size_t)d0 + e] = wst[e];
    __syncwarp();
    if (ok) { wst[3 * r] = __int_as_float(cz); wst[3 * r + 1] = __int_as_float(cy); wst[3 * r + 2] = __int_as_float(cx); }
    __syncwarp();
    for (int e = lane; e < 3 * nrow; e += 32) pt_coor[3 * (size_t)d0 + e] = __float_as_int(wst[e]);
    __syncwarp();
    if (ok) {
      // point_offsets = p - ((c * vs + min) + vs / 2), every step rounded to fp32 (encoder.py:516-523)
      wst[3 * r] = __fsub_rn(x[j], __fadd_rn(__fadd_rn(__fmul_rn((float)cx, P.vx), P.lox), hx));
      wst[3 * r + 1] = __fsub_rn(y[j], __fadd_rn(__fadd_rn(__fmul_rn((float)cy, P.vy), P.loy), hy));
      wst[3 * r + 2] = __fsub_rn(z[j], __fadd_rn(__fadd_rn(__fmul_rn((float)cz, P.vz), P.loz), hz));
    }
    __syncwarp();
    for (int e = lane; e < 3 * nrow; e += 32) pt_offs[3 * (size_t)d0 + e] = wst[e];
    __syncwarp();
    if (!ok) continue;
    pt_idx[d] = blk * IDX_CHUNK + j * IDX_BLOCK + threadIdx.x;
    pt_pillar[d] = q[j];
    pt_slot[d] = slot[j];
    if (slot[j] == 0) {  // exactly one point per pillar sees slot 0
      pil_coor[3 * (size_t)q[j]] = cz;
      pil_coor[3 * (size_t)q[j] + 1] = cy;
      pil_coor[3 * (size_t)q[j] + 2] = cx;
      pil_pix[q[j]] = f * HW + cy * P.gx + cx;  // PointPillarsScatter: y * nx + x (encoder.py:141)
    }
  }
}

// ---------------------------------------------------------------- K4: CSR offsets of the pillars
// Points and pillars are both stored frame after frame, so pil_start is ONE global exclusive scan of pil_cnt.  Two
// levels in one launch: block-local scans of 2048 pillars (-> pil_loc, block totals); the last block to finish scans
// the block totals (-> blk_base).  K5 adds the two.
constexpr int PSCAN_ITEMS = 2048;

__global__ void __launch_bounds__(1024) k_pillar_scan(const int* __restrict__ pil_cnt, const int* __restrict__ counts,
                                                      int F, int* __restrict__ pil_loc, int* __restrict__ blk_tot,
                                                      int* __restrict__ blk_base, unsigned* __restrict__ ticket) {
  __shared__ int sm[33];
  __shared__ int is_last;
  const int M = counts[3 * F + 1 + F];
  const int q = blockIdx.x * PSCAN_ITEMS + 2 * threadIdx.x;
  int tot = 0;
  if (blockIdx.x * PSCAN_ITEMS < M) {
    const int v0 = q < M ? pil_cnt[q] : 0, v1 = q + 1 < M ? pil_cnt[q + 1] : 0;
    const int ex = block_excl_scan<1024>(v0 + v1, sm, tot);
    if (q < M) pil_loc[q] = ex;
    if (q + 1 < M) pil_loc[q + 1] = ex + v0;
  }
  if (threadIdx.x == 0) {
    blk_tot[blockIdx.x] = tot;
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int nb = (int)gridDim.x;
  int carry = 0;
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? __ldcg(&blk_tot[i]) : 0;
    int t;
    const int ex = block_excl_scan<1024>(v, sm, t);
    if (i < nb) blk_base[i] = carry + ex;
    carry += t;
  }
}

// ---------------------------------------------------------------- K5: CSR fill
// pil_start (final) for every pillar; for every point its CSR position j = pil_start[q] + slot:
//   sorted_pt[j] = p   and   csr_rec[j] = (x, y, z, q)   -- the 16-byte record every later pass over the points of a
// pillar reads sequentially (no p -> xyz -> q double indirection in the feature-net kernels).
template <int U>
__global__ void __launch_bounds__(256) k_fill_csr(const int* __restrict__ counts, int F, int total_cap,
                                                  const int* __restrict__ pt_pillar, const int* __restrict__ pt_slot,
                                                  const float* __restrict__ pt_xyz, const int* __restrict__ pil_loc,
                                                  const int* __restrict__ blk_base, int* __restrict__ pil_start,
                                                  int* __restrict__ sorted_pt, float4* __restrict__ csr_rec) {
  const int n = min(counts[2 * F + F], total_cap);
  const int M = counts[3 * F + 1 + F];
  const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int q = t0; q < M; q += stride) pil_start[q] = pil_loc[q] + blk_base[q / PSCAN_ITEMS];
  if (t0 == 0) pil_start[M] = n;
  // U points per thread and iteration, every load of the U issued before the first dependent gather
  for (int p0 = t0; p0 < n; p0 += U * stride) {
    int q[U], sl[U];
    float x[U], y[U], z[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * stride;
      q[u] = -1;
      if (p < n) {
        q[u] = pt_pillar[p]; sl[u] = pt_slot[p];
        x[u] = pt_xyz[3 * (size_t)p]; y[u] = pt_xyz[3 * (size_t)p + 1]; z[u] = pt_xyz[3 * (size_t)p + 2];
      }
    }
    int j[U];
#pragma unroll
    for (int u = 0; u < U; ++u) j[u] = q[u] >= 0 ? pil_loc[q[u]] + blk_base[q[u] / PSCAN_ITEMS] + sl[u] : -1;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (j[u] < 0) continue;
      sorted_pt[j[u]] = p0 + u * stride;
      csr_rec[j[u]] = make_float4(x[u], y[u], z[u], __int_as_float(q[u]));
    }
  }
}

VoxelParams make_voxel_params(const float* vs, const float* rng) {
  VoxelParams P;
  P.vx = vs[0]; P.vy = vs[1]; P.vz = vs[2];
  P.lox = rng[0]; P.loy = rng[1]; P.loz = rng[2];
  // const int grid_x = round((coors_x_max - coors_x_min) / voxel_x);  (voxelization_cuda.cu:269)
  P.gx = (int)roundf((rng[3] - rng[0]) / vs[0]);
  P.gy = (int)roundf((rng[4] - rng[1]) / vs[1]);
  P.gz = (int)roundf((rng[5] - rng[2]) / vs[2]);
  return P;
}

}  // namespace dfb

using namespace dfb;

extern "C" int dfb_grid_size(const float* voxel_size, const float* range, int* grid_xyz) {
  VoxelParams P = make_voxel_params(voxel_size, range);
  grid_xyz[0] = P.gx; grid_xyz[1] = P.gy; grid_xyz[2] = P.gz;
  return DFB_OK;
}

extern "C" int dfb_index_workspace(int F, int Nmax, const float* voxel_size, const float* range,
                                   long long* bitmap_words_per_frame, long long* blocks_per_frame) {
  VoxelParams P = make_voxel_params(voxel_size, range);
  if (P.gx <= 0 || P.gy <= 0 || P.gz <= 0) { set_error("dfb_index_workspace: empty voxel grid"); return DFB_ERR_ARG; }
  const long long cells = (long long)P.gx * P.gy * P.gz;
  if (cells >= (1ll << 31)) { set_error("dfb_index_workspace: %lld cells do not fit an int32 key", cells); return DFB_ERR_UNSUPPORTED; }
  *bitmap_words_per_frame = (cells + 31) / 32;
  *blocks_per_frame = (Nmax + IDX_CHUNK - 1) / IDX_CHUNK;
  return DFB_OK;
}

extern "C" long long dfb_index_scan_workspace(int F, long long bitmap_words_per_frame, long long pil_cap) {
  const long long S = (bitmap_words_per_frame + SEG_WORDS - 1) / SEG_WORDS;
  const long long nb = (pil_cap + PSCAN_ITEMS - 1) / PSCAN_ITEMS;
  return 2 * (long long)F * S + 2 * nb + pil_cap;  // seg_tot | seg_base | blk_tot | blk_base | pil_loc   (int32 elements)
}

extern "C" int dfb_pillar_index(const dfb_index_args* a, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (!a || a->F <= 0 || a->Nmax < 0 || a->pt_stride < 3) { set_error("dfb_pillar_index: bad sizes"); return DFB_ERR_ARG; }
  VoxelParams P = make_voxel_params(a->voxel_size, a->range);
  long long Wd_, nblk_;
  int rc = dfb_index_workspace(a->F, a->Nmax, a->voxel_size, a->range, &Wd_, &nblk_);
  if (rc) return rc;
  const int Wd = (int)Wd_, nblk = (int)(nblk_ > 0 ? nblk_ : 1), F = a->F;
  const long long cap = (long long)F * a->Nmax;
  if (cap >= (1ll << 31) || (long long)F * P.gx * P.gy >= (1ll << 31)) {
    set_error("dfb_pillar_index: batch too large for int32 indexing"); return DFB_ERR_UNSUPPORTED;
  }
  const long long pil_cap = a->pil_cap;
  const int S = (Wd + SEG_WORDS - 1) / SEG_WORDS;
  const int nb = (int)((pil_cap + PSCAN_ITEMS - 1) / PSCAN_ITEMS);
  if (!a->scan_ws || !a->tickets || !a->csr_rec) { set_error("dfb_pillar_index: scan workspace / tickets / csr_rec missing"); return DFB_ERR_ARG; }
  int* seg_tot = a->scan_ws;
  int* seg_base = seg_tot + (size_t)F * S;
  int* blk_tot = seg_base + (size_t)F * S;
  int* blk_base = blk_tot + nb;
  int* pil_loc = blk_base + nb;
  if (a->occ && ((uintptr_t)a->occ & 15)) { set_error("dfb_pillar_index: occ must be 16-byte aligned"); return DFB_ERR_ARG; }
  if (a->zero_base) {
    // (occ |) bitmap | pil_cnt | blk_cnt | tickets live in one allocation: one memset
    cudaMemsetAsync(a->zero_base, 0, (size_t)a->zero_bytes, st);
  } else {
    if (a->occ) cudaMemsetAsync(a->occ, 0, (size_t)F * Wd * 32, st);
    cudaMemsetAsync(a->bitmap, 0, sizeof(unsigned) * (size_t)F * Wd, st);
    cudaMemsetAsync(a->pil_cnt, 0, sizeof(int) * (size_t)pil_cap, st);
    cudaMemsetAsync(a->blk_cnt, 0, sizeof(int) * (size_t)F * nblk, st);
    cudaMemsetAsync(a->tickets, 0, sizeof(unsigned) * 2, st);
  }
  if (a->Nmax > 0) {
    dim3 g((unsigned)nblk_, F);
    k_mark_points<<<g, IDX_BLOCK, 0, st>>>(a->pts, a->Nmax, a->pt_stride, P, Wd, nblk, a->bitmap, a->occ, a->blk_cnt);
  }
  {
    dim3 g(S, F);
    k_bitmap_scan<<<g, 1024, 0, st>>>(a->bitmap, a->occ, Wd, S, a->word_rank, seg_tot, seg_base, a->blk_cnt, nblk, a->counts, F,
                                      a->tickets);
  }
  if (a->Nmax > 0) {
    dim3 g((unsigned)nblk_, F);
    k_compact<<<g, IDX_BLOCK, 0, st>>>(a->pts, a->Nmax, a->pt_stride, P, Wd, nblk, F, P.gx * P.gy, a->bitmap,
                                       a->word_rank, seg_base, S, a->blk_cnt, a->counts, a->pt_xyz, a->pt_coor, a->pt_idx,
                                       a->pt_offs, a->pt_pillar, a->pt_slot, a->pil_cnt, a->pil_coor, a->pil_pix);
  }
  k_pillar_scan<<<nb, 1024, 0, st>>>(a->pil_cnt, a->counts, F, pil_loc, blk_tot, blk_base, a->tickets + 1);
  {
    long long work = cap > pil_cap ? cap : pil_cap;
    int blocks = (int)((work + 255) / 256);
    int maxb = sm_count() * 8;
    if (blocks > maxb) blocks = maxb;
    if (blocks < 1) blocks = 1;
    // one point per thread and iteration: two or four in flight measured SLOWER on B200 (r01 sweep: the random 16-byte
    // CSR writes then spread over a wider window and more partially written sectors are evicted before they fill)
    k_fill_csr<1><<<blocks, 256, 0, st>>>(a->counts, F, (int)cap, a->pt_pillar, a->pt_slot, a->pt_xyz, pil_loc, blk_base,
                                          a->pil_start, a->sorted_pt, (float4*)a->csr_rec);
  }
  add_launches(a->Nmax > 0 ? 5 : 3);
  return check_launch("dfb_pillar_index");
}
