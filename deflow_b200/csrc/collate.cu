// Device-side collate_fn_pad: ground-mask strip (stable) + NaN / zero padding of a batch of raw sweeps.
//
// Reference: collate_fn_pad (OpenSceneFlow/src/dataset.py:22-74) -- per sample `pc[~gm]` boolean indexing on the host
// followed by torch.nn.utils.rnn.pad_sequence (NaN for points, 0 for flow / flow_is_valid / flow_category_indices) in
// the DataLoader workers; and ModelWrapper.run_model_wo_ground_data (OpenSceneFlow/src/trainer.py:268-282), the same
// strip for a single validation sample.  Here the raw samples arrive concatenated ("ragged": one H2D copy per field)
// and three launches produce the padded batch: per-chunk keep counts -> per-sample exclusive scan -> stable compaction
// + padding.  Integer / byte work, HBM-bound: 13 B read + (12 [+ 14]) B written per raw point.
#include "common.cuh"
#include "../../include/deflow_b200.h"

namespace dfb {

constexpr int CL_THREADS = 256;
constexpr int CL_PER_THREAD = 4;
constexpr int CL_CHUNK = CL_THREADS * CL_PER_THREAD;   // points per block

__global__ void __launch_bounds__(CL_THREADS) k_collate_count(const unsigned char* __restrict__ ground,
                                                              const int* __restrict__ offs, int max_chunks,
                                                              int* __restrict__ chunk_cnt) {
  const int b = blockIdx.y, c = blockIdx.x;
  const int lo = offs[b], n = offs[b + 1] - lo;
  if (c * CL_CHUNK >= n) {
    if (threadIdx.x == 0) chunk_cnt[b * max_chunks + c] = 0;
    return;
  }
  int keep = 0;
#pragma unroll
  for (int j = 0; j < CL_PER_THREAD; ++j) {
    const int i = c * CL_CHUNK + threadIdx.x * CL_PER_THREAD + j;
    if (i < n) keep += ground[lo + i] ? 0 : 1;
  }
  __shared__ int red[CL_THREADS / 32];
  for (int o = 16; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = keep;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < CL_THREADS / 32; ++w) s += red[w];
    chunk_cnt[b * max_chunks + c] = s;
  }
}

// one warp per sample: chunk counts -> exclusive offsets, total -> keep_counts[b]; flags a sample that exceeds Nmax
__global__ void __launch_bounds__(32) k_collate_scan(int max_chunks, int Nmax, int* __restrict__ chunk_cnt,
                                                     int* __restrict__ keep_counts, int* __restrict__ overflow) {
  const int b = blockIdx.x, lane = threadIdx.x;
  int run = 0;
  for (int c0 = 0; c0 < max_chunks; c0 += 32) {
    const int c = c0 + lane;
    const int v = c < max_chunks ? chunk_cnt[b * max_chunks + c] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (c < max_chunks) chunk_cnt[b * max_chunks + c] = run + inc - v;
    run += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) {
    keep_counts[b] = run;
    if (run > Nmax) atomicExch(overflow, 1);
  }
}

__global__ void __launch_bounds__(CL_THREADS) k_collate_write(
    const float* __restrict__ pts, const unsigned char* __restrict__ ground, const int* __restrict__ offs, int max_chunks,
    int Nmax, const int* __restrict__ chunk_off, const int* __restrict__ keep_counts, const float* __restrict__ flow,
    const unsigned char* __restrict__ valid, const unsigned char* __restrict__ cls, float* __restrict__ pts_out,
    float* __restrict__ flow_out, unsigned char* __restrict__ valid_out, unsigned char* __restrict__ cls_out) {
  __shared__ int scan_smem[33];
  const int b = blockIdx.y, c = blockIdx.x;
  const int lo = offs[b], n = offs[b + 1] - lo;
  const int kept = min(keep_counts[b], Nmax);
  const size_t ob = (size_t)b * Nmax;
  if (c < max_chunks && c * CL_CHUNK < n) {   // block-uniform: the scan below has barriers
    const int i0 = c * CL_CHUNK + threadIdx.x * CL_PER_THREAD;
    bool k[CL_PER_THREAD];
    int mine = 0;
#pragma unroll
    for (int j = 0; j < CL_PER_THREAD; ++j) {
      k[j] = (i0 + j < n) && !ground[lo + i0 + j];
      mine += k[j];
    }
    int total;
    int dst = chunk_off[b * max_chunks + c] + block_excl_scan<CL_THREADS>(mine, scan_smem, total);
#pragma unroll
    for (int j = 0; j < CL_PER_THREAD; ++j) {
      if (!k[j]) continue;
      if (dst < Nmax) {
        const size_t s = (size_t)(lo + i0 + j), o = ob + dst;
        pts_out[3 * o] = pts[3 * s]; pts_out[3 * o + 1] = pts[3 * s + 1]; pts_out[3 * o + 2] = pts[3 * s + 2];
        if (flow_out) { flow_out[3 * o] = flow[3 * s]; flow_out[3 * o + 1] = flow[3 * s + 1]; flow_out[3 * o + 2] = flow[3 * s + 2]; }
        if (valid_out) valid_out[o] = valid[s];
        if (cls_out) cls_out[o] = cls[s];
      }
      ++dst;
    }
  }
  // padding rows [kept, Nmax) of this block's output range: NaN points, zero flow / valid / class (pad_sequence)
  const float qnan = __int_as_float(0x7fc00000);
#pragma unroll
  for (int j = 0; j < CL_PER_THREAD; ++j) {
    const int r = c * CL_CHUNK + j * CL_THREADS + threadIdx.x;
    if (r >= kept && r < Nmax) {
      const size_t o = ob + r;
      pts_out[3 * o] = qnan; pts_out[3 * o + 1] = qnan; pts_out[3 * o + 2] = qnan;
      if (flow_out) { flow_out[3 * o] = 0.f; flow_out[3 * o + 1] = 0.f; flow_out[3 * o + 2] = 0.f; }
      if (valid_out) valid_out[o] = 0;
      if (cls_out) cls_out[o] = 0;
    }
  }
}

}  // namespace dfb

using namespace dfb;

extern "C" long long dfb_collate_workspace(int B, int max_points_per_sample) {
  const int max_chunks = (max_points_per_sample + CL_CHUNK - 1) / CL_CHUNK;
  return (long long)B * (max_chunks > 0 ? max_chunks : 1) + 1;   // ints: chunk counts / offsets + the overflow flag
}

extern "C" int dfb_collate_pad(const float* pts, const unsigned char* ground, const int* offs, int B,
                               int max_points_per_sample, int Nmax, const float* flow, const unsigned char* valid,
                               const unsigned char* cls, float* pts_out, float* flow_out, unsigned char* valid_out,
                               unsigned char* cls_out, int* keep_counts, int* workspace, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (B <= 0 || Nmax < 0 || max_points_per_sample < 0) { set_error("dfb_collate_pad: bad sizes"); return DFB_ERR_ARG; }
  if (!pts || !ground || !offs || !pts_out || !keep_counts || !workspace) { set_error("dfb_collate_pad: null pointer"); return DFB_ERR_ARG; }
  if ((flow_out && !flow) || (valid_out && !valid) || (cls_out && !cls)) { set_error("dfb_collate_pad: output without its input"); return DFB_ERR_ARG; }
  int max_chunks = (max_points_per_sample + CL_CHUNK - 1) / CL_CHUNK;
  if (max_chunks < 1) max_chunks = 1;
  int* overflow = workspace + (size_t)B * max_chunks;
  cudaMemsetAsync(overflow, 0, sizeof(int), st);
  dim3 g1(max_chunks, B);
  k_collate_count<<<g1, CL_THREADS, 0, st>>>(ground, offs, max_chunks, workspace);
  k_collate_scan<<<B, 32, 0, st>>>(max_chunks, Nmax, workspace, keep_counts, overflow);
  int out_chunks = (Nmax + CL_CHUNK - 1) / CL_CHUNK;
  dim3 g3(out_chunks > max_chunks ? out_chunks : max_chunks, B);
  k_collate_write<<<g3, CL_THREADS, 0, st>>>(pts, ground, offs, max_chunks, Nmax, workspace, keep_counts, flow, valid, cls, pts_out,
                                             flow_out, valid_out, cls_out);
  add_launches(3);
  return check_launch("dfb_collate_pad");
}
