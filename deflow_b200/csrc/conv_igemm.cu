// Implicit-GEMM 2-D convolution on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), NHWC bf16.
//
// Reference path: every nn.Conv2d of FastFlow3DUNet (OpenSceneFlow/src/models/basic/unet.py:49-68,
// ConvWithNorms in basic/__init__.py:61-79) -- cuDNN library calls on NCHW fp32/TF32 there.
//
// One kernel serves forward (stride 1 and 2, 1x1 and 3x3, channel-concatenated inputs) and data-gradient
// (stride 1; stride 2 as one launch over the four parity planes of the result): the host builds a table of K-steps, each step
// naming (source tensor map, pixel shift, channel chunk, weight K offset).  Per step the TMA producer
// loads a [16 x 8 pixels] x KC-channel activation box (out-of-bounds = zero padding) and an [N x KC]
// weight box; one thread issues KC/16 tcgen05.mma (M = 128 pixels, N = Cout) into a TMEM accumulator;
// eight epilogue warps read TMEM, add the bias, accumulate BatchNorm batch statistics (sum, sum of
// squares per channel) and store bf16 NHWC through a swizzled staging tile and TMA.  Accumulators are double buffered in TMEM so the epilogue
// of tile i overlaps the MMAs of tile i+1.  Persistent: one CTA per SM, static round-robin over tiles.
#include "tc_common.cuh"
#include "../../include/deflow_b200.h"

#include <mutex>

namespace dfb {
namespace tc {

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tensor_map_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return DFB_ERR_CUDA; }
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu ...)", (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1]); return DFB_ERR_CUDA; }
  return DFB_OK;
}

// ------------------------------------------------------------------------------------------------
constexpr int TILE_H = 16, TILE_W = 8, TILE_M = TILE_H * TILE_W;  // 128 output pixels per tile
constexpr int MAX_STEPS = 224;
constexpr int MAX_AMAPS = 8;   // 4 sources / parity planes x (hi, lo) in split-precision mode
// warp 0 TMA, warp 1 MMA + TMEM owner, warps 2..9 epilogue.  Eight epilogue warps: two per TMEM lane quarter, each
// taking every other block of 32 accumulator columns.  With four (one per scheduler, ~300 dependent instructions per
// column block) the epilogue of a 64-channel tile took longer than its MMAs (r01: 7 k cycles per 256 pixels measured
// against 3.6 k cycles of MMA time) -- the layers with few K steps per tile were epilogue-bound, not tensor-bound.
constexpr int EPI_WARPS = 8;
constexpr int IGEMM_THREADS = 64 + 32 * EPI_WARPS;

struct KStep {
  int8_t map, dy, dx, pad;  // source tensor map, pixel shift (in that source's pixel grid)
  int16_t c0;               // channel offset inside the source
  int16_t wk;               // K offset in the packed weight matrix
};

struct IgemmMaps {
  CUtensorMap a[MAX_AMAPS];
  CUtensorMap b;
  CUtensorMap out;   // bf16 output as [ch, x, y, n] with box [32, 8, 4, 1], SWIZZLE_64B (epilogue TMA stores)
  CUtensorMap out2;  // second output of a split data-gradient launch (accumulator columns >= split_col); follows `out`
  CUtensorMap out3, out4;  // parity-merged stride-2 data gradient: out, out2, out3, out4 = the four parity planes of gx
};

struct IgemmParams {
  int n_img, Ht, Wt;              // tile-space extent (pixels the tiles enumerate)
  int tiles_x, tiles_y, num_tiles;
  int nsteps;
  // epilogue: element (n, y, x, ch) of the tile space goes to out[n*img + (y*sy+oy)*row + (x*sx+ox)*pix + ch]
  void* out;
  int out_fp32;
  long long out_img, out_row, out_pix;
  int sy, sx, oy, ox;
  const float* bias;   // [N] or null
  double* stats;       // [2][N] running sum / sum of squares over all output elements, or null
  int stats_sum_only;  // only the sums are wanted (bias gradient from a data-gradient launch)
  int split_col;       // > 0: accumulator columns >= split_col are the channels of a second output tensor (maps.out2)
  void* out2;          // that tensor [n, Ht, Wt, N - split_col] bf16 (sx = sy = 1, no offsets)
  // halo variant (3x3 stride 1): steps[] holds one entry per 64-channel chunk (wk = K offset of tap 0)
  int halo_pitch;      // pixels per halo row in shared memory (10 or 16)
  int halo_flip;       // data gradient: tap t reads halo offset (2 - t/3, 2 - t%3)
  int tap_k_stride;    // K distance between consecutive taps in the packed weights
  int w_resident;      // all 9 * nsteps weight tiles fit the ring: load them once
  int dbg_base_offset; // descriptor base-offset field = (start >> 7) & 7 (experiment switch)
  // parity-merged stride-2 data gradient (k_conv_igemm only): work item = (region tile, parity plane of gx); the K steps
  // of plane p are steps[seg[p] .. seg[p+1]), its output map is (&maps.out)[p]; num_tiles counts items (4 x region tiles)
  int par_mode;
  int16_t seg[5];
  // k_conv_igemm only.  wide: the epilogue stores 64-channel (128-byte) rows, box [64, 8, 4, 1] SWIZZLE_128B, half as
  // many TMA requests per tile.  b_resident: all nsteps weight tiles are loaded ONCE into the tail of the stage ring
  // (the ring keeps ring_stages stages and moves A tiles only) instead of once per tile and K step.
  // alt: the two groups of four epilogue warps take alternate tiles (igemm_epilogue_alt; bf16 output only)
  int wide, b_resident, ring_stages, alt;
  KStep steps[MAX_STEPS];
};

constexpr int EPI_STAGE_BYTES = 2048;                 // per epilogue warp: 32 pixels x 32 channels bf16 / 32 x 16 fp32
constexpr int EPI_SMEM = EPI_WARPS * EPI_STAGE_BYTES;

// Epilogue shared by the implicit-GEMM kernels: 4 warps = 128 TMEM lanes = the 16 x 8 pixels of a tile; warp q owns
// tile rows 4q .. 4q+3 (32 pixels, lane = pixel).  Per block of 32 channels:
//   * bias from shared memory;
//   * BatchNorm statistics: the warp's 32 x 32 fp32 block goes through a 2 KB XOR-swizzled slab in two halves of 16
//     columns and is summed column-wise (lane = column x row parity) -- conflict-free both ways, ~1/2 the
//     instructions of a shuffle butterfly;
//   * bf16 output: each lane writes its 64-byte row into the same 2 KB region with the 64-byte TMA swizzle and one
//     lane issues a TMA store of the [32 ch x 8 x 4] box -- fully coalesced, clipped at the image border by the
//     hardware.  (A lane-per-pixel st.global touches 32 different 128-byte lines per instruction.)
//   * fp32 output (parity mode): direct stores.
// PAIR (k_conv_igemm_halo_pair): the N = 128 accumulator columns are two 64-channel halves; lane (r, c) holds output
// pixel (2r, c) of a 32 x 8 tile in columns 0..63 and pixel (2r + 1, c) in columns 64..127.  The output tensor map is
// then 5-D [ch, x, row parity, row / 2, n] and a store box is [32 ch, 8, 1, 4, 1].
// Parity-merged launches: item w = blockIdx.x + i * gridDim.x covers region tile w / 4 and parity plane ((w & 3) + i) & 3.
// The grid is a multiple of 4, so the four items of a region share i: their planes are a permutation of 0..3 (every
// (region, plane) exactly once), four neighbouring CTAs work on the four planes of one region at the same time (their
// interleaved 64-byte stores meet in L2 and leave as full lines, gy is fetched from DRAM once), and every CTA cycles
// through the planes, whose K loops differ in length (1, 2, 2, 4 taps of a 3x3 kernel).
__device__ __forceinline__ int par_decode(const IgemmParams& P, int& tile) {
  if (!P.par_mode) return 0;
  const int par = ((tile & 3) + tile / (int)gridDim.x) & 3;
  tile >>= 2;
  return par;
}

// DUAL (k_conv_igemm_halo<N, true>): a work item is a PAIR of consecutive tiles (2 * item, 2 * item + 1) whose two
// accumulators -- TMEM columns (2 * acc + which) * N -- complete together.
template <int N, bool PAIR = false, bool DUAL = false>
__device__ __forceinline__ void igemm_epilogue(const IgemmParams& P, const CUtensorMap* out_map, uint32_t tmem_base,
                                               uint64_t* acc_full, uint64_t* acc_empty, float* s_stats,
                                               const float* s_bias, uint8_t* stage_all, int warp, int lane) {
  const int q = warp & 3;  // TMEM lane quarter this warp may access
  const int cg = (warp - 2) >> 2;          // which of the EPI_WARPS / 4 interleaved column-block sets this warp takes
  const int m = q * 32 + lane;
  const int r = m >> 3, c = m & 7;
  uint8_t* stg = stage_all + (warp - 2) * EPI_STAGE_BYTES;
  float* slab = reinterpret_cast<float*>(stg);
  const int sw16 = (lane >> 1) & 15;        // slab column swizzle of this lane's row
  const int sw4 = (lane >> 1) & 3;          // 64-byte TMA swizzle of this lane's row
  const int rcol = lane & 15, rpar = lane >> 4;
  bool store_pending = false;
  int acc = 0;
  uint32_t acc_phase = 0;
  const int n_items = DUAL ? P.num_tiles / 2 : P.num_tiles;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    mbar_wait(&acc_full[acc], acc_phase);
    tc_fence_after();
#pragma unroll 1
    for (int which = 0; which < (DUAL ? 2 : 1); ++which) {
    int tile = DUAL ? 2 * item + which : item;
    const int par = (PAIR || DUAL) ? 0 : par_decode(P, tile);
    const int tx = tile % P.tiles_x, t2 = tile / P.tiles_x, ty = t2 % P.tiles_y, n = t2 / P.tiles_y;
    const int y0 = PAIR ? ty * 2 * TILE_H + 2 * r : ty * TILE_H + r, x = tx * TILE_W + c;
    const long long off0 = (long long)n * P.out_img + (long long)(y0 * P.sy + P.oy) * P.out_row +
                           (long long)(x * P.sx + P.ox) * P.out_pix;
    const int acc_col = (DUAL ? 2 * acc + which : acc) * N;
#pragma unroll 1
    for (int col = 32 * cg; col < N; col += 32 * (EPI_WARPS / 4)) {
      const int half = PAIR ? col / (N / 2) : 0;       // output row parity of this column block
      const bool second = !PAIR && P.split_col > 0 && col >= P.split_col;   // second output tensor of a split launch
      const int ch0 = PAIR ? col % (N / 2) : col;      // channel index for bias / statistics
      const int oc0 = second ? col - P.split_col : ch0;  // first channel inside the output tensor
      const int y = y0 + half;
      const bool valid = y < P.Ht && x < P.Wt;
      const long long off = off0 + (long long)half * P.out_row;
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc_col + col, v);
      if (P.bias) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(s_bias + ch0 + i);
          v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
        }
      }
      if (store_pending) {            // the previous TMA store must have read the region before it is rewritten
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        store_pending = false;
      }
      if (P.stats) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int i = 0; i < 16; ++i) slab[lane * 16 + (i ^ sw16)] = valid ? v[h * 16 + i] : 0.f;
          __syncwarp();
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {       // rows 2i + rpar of column rcol
            const float t = slab[(2 * i + rpar) * 16 + (rcol ^ i)];
            s1 += t;
            s2 = fmaf(t, t, s2);
          }
          s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
          s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
          if (lane < 16) {
            atomicAdd(&s_stats[ch0 + h * 16 + lane], s1);
            if (!P.stats_sum_only) atomicAdd(&s_stats[(PAIR ? N / 2 : N) + ch0 + h * 16 + lane], s2);
          }
          __syncwarp();
        }
      }
      if (P.out_fp32) {
        if (valid) {
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(P.out) + off + ch0);
#pragma unroll
          for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * i], v[8 * i + 1]), p1 = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
          __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]), p3 = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
          uint4 u;
          u.x = *reinterpret_cast<unsigned*>(&p0); u.y = *reinterpret_cast<unsigned*>(&p1);
          u.z = *reinterpret_cast<unsigned*>(&p2); u.w = *reinterpret_cast<unsigned*>(&p3);
          *reinterpret_cast<uint4*>(stg + lane * 64 + ((i ^ sw4) << 4)) = u;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) tma_store_5d(out_map, stg, ch0, tx * TILE_W, half, ty * TILE_H + 4 * q, n);
          else tma_store_4d(second ? out_map + 1 : out_map + par, stg, oc0, tx * TILE_W, ty * TILE_H + 4 * q, n);
          tma_store_commit();
        }
        store_pending = true;
      }
    }
    }   // which
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&acc_empty[acc]);
    acc ^= 1;
    if (acc == 0) acc_phase ^= 1;
  }
  if (lane == 0) tma_store_wait_all();
}

// The same epilogue with 64-channel stores (k_conv_igemm, N a multiple of 64, bf16 output).  Measured on the general
// kernels (r02 launch lists): a tile costs ~3.5-4 cycles per TMA request -- one request per box row, 64-byte rows cost what
// 128-byte rows cost -- so the layers with one or two K steps per tile were bound by the NUMBER of requests, and half of
// those were the 32-channel store rows.  A warp takes 64 accumulator columns (two tcgen05.ld), takes the statistics of
// both halves through the slab first (it aliases the staging tile), then writes its 32 pixel rows of 128 bytes with the
// 128-byte swizzle and issues ONE store of the [64 ch x 8 x 4] box.
constexpr int EPI_STAGE_BYTES_WIDE = 4096;
constexpr int EPI_SMEM_WIDE = EPI_WARPS * EPI_STAGE_BYTES_WIDE;

template <int N, bool PAIR = false>
__device__ __forceinline__ void igemm_epilogue_wide(const IgemmParams& P, const CUtensorMap* out_map, uint32_t tmem_base,
                                                    uint64_t* acc_full, uint64_t* acc_empty, float* s_stats,
                                                    const float* s_bias, uint8_t* stage_all, int warp, int lane) {
  const int q = warp & 3;
  const int cg = (warp - 2) >> 2;
  const int m = q * 32 + lane;
  const int r = m >> 3, c = m & 7;
  uint8_t* stg = stage_all + (warp - 2) * EPI_STAGE_BYTES_WIDE;
  float* slab = reinterpret_cast<float*>(stg);
  const int sw16 = (lane >> 1) & 15;
  const int sw8 = lane & 7;                 // 128-byte TMA swizzle of this lane's row
  const int rcol = lane & 15, rpar = lane >> 4;
  bool store_pending = false;
  int acc = 0;
  uint32_t acc_phase = 0;
  for (int item = blockIdx.x; item < P.num_tiles; item += gridDim.x) {
    int tile = item;
    const int par = PAIR ? 0 : par_decode(P, tile);
    const int tx = tile % P.tiles_x, t2 = tile / P.tiles_x, ty = t2 % P.tiles_y, n = t2 / P.tiles_y;
    const int y0 = PAIR ? ty * 2 * TILE_H + 2 * r : ty * TILE_H + r, x = tx * TILE_W + c;
    mbar_wait(&acc_full[acc], acc_phase);
    tc_fence_after();
#pragma unroll 1
    for (int col = 64 * cg; col < N; col += 64 * (EPI_WARPS / 4)) {
      // PAIR: N = 128 accumulator columns = two output rows x 64 channels; a 64-column block is one row parity
      const int half = PAIR ? col / 64 : 0;
      const int ch0 = PAIR ? 0 : col;                    // channel index for bias / statistics
      const bool second = !PAIR && P.split_col > 0 && col >= P.split_col;
      const int oc0 = second ? col - P.split_col : ch0;
      const bool valid = y0 + half < P.Ht && x < P.Wt;
      uint32_t pk[32];                     // the row's 64 channels as packed bf16 pairs
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {     // one tcgen05.ld of 32 columns at a time: 32 fp32 + 16 packed registers live
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * N + col + 32 * hh, v);
        if (P.bias) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(s_bias + ch0 + 32 * hh + i);
            v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
          }
        }
        if (hh == 0 && store_pending) {    // the previous TMA store must have read the tile before the slab reuses it
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
          store_pending = false;
        }
        if (P.stats) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int i = 0; i < 16; ++i) slab[lane * 16 + (i ^ sw16)] = valid ? v[h * 16 + i] : 0.f;
            __syncwarp();
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float t = slab[(2 * i + rpar) * 16 + (rcol ^ i)];
              s1 += t;
              s2 = fmaf(t, t, s2);
            }
            s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
            if (lane < 16) {
              atomicAdd(&s_stats[ch0 + 32 * hh + h * 16 + lane], s1);
              if (!P.stats_sum_only) atomicAdd(&s_stats[(PAIR ? N / 2 : N) + ch0 + 32 * hh + h * 16 + lane], s2);
            }
            __syncwarp();
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const __nv_bfloat162 p2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
          pk[16 * hh + i] = *reinterpret_cast<const unsigned*>(&p2);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<uint4*>(stg + lane * 128 + ((i ^ sw8) << 4)) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) tma_store_5d(out_map, stg, 0, tx * TILE_W, half, ty * TILE_H + 4 * q, n);
        else tma_store_4d(second ? out_map + 1 : out_map + par, stg, oc0, tx * TILE_W, ty * TILE_H + 4 * q, n);
        tma_store_commit();
      }
      store_pending = true;
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&acc_empty[acc]);
    acc ^= 1;
    if (acc == 0) acc_phase ^= 1;
  }
  if (lane == 0) tma_store_wait_all();
}

// Incremental tile decode for the persistent loops: tile index t0, t0 + stride, ... -> (tx, ty, n) without the three
// integer divisions per tile (ncu r02: ~90 of the ~130 instructions a single-thread producer spent per tile, at ~8
// cycles per dependent instruction, when a tile has one or two K steps).
struct TileIter {
  int tx, ty, n, dx, dy, dn, tiles_x, tiles_y;
  __device__ __forceinline__ void init(int t0, int stride, int tiles_x_, int tiles_y_) {
    tiles_x = tiles_x_; tiles_y = tiles_y_;
    tx = t0 % tiles_x; int r = t0 / tiles_x; ty = r % tiles_y; n = r / tiles_y;
    dx = stride % tiles_x; r = stride / tiles_x; dy = r % tiles_y; dn = r / tiles_y;
  }
  __device__ __forceinline__ void next() {
    tx += dx;
    int cy = dy;
    if (tx >= tiles_x) { tx -= tiles_x; ++cy; }
    ty += cy;
    n += dn;
    if (ty >= tiles_y) { ty -= tiles_y; ++n; }
  }
};

// One block of 32 * NH accumulator columns for this warp's 32 pixel rows: bias, BatchNorm statistics through the slab
// (it aliases the staging tile, so statistics come first), bf16 rows of NH * 64 bytes with the matching TMA swizzle.
// The caller issues the TMA store.  NH = 1: 32 channels, SWIZZLE_64B; NH = 2: 64 channels, SWIZZLE_128B.
template <int NH>
__device__ __forceinline__ void epi_block(const IgemmParams& P, uint32_t taddr, const float* bias_blk, float* sum_blk,
                                          float* sq_blk, bool valid, uint8_t* stg, int lane, bool& store_pending) {
  float* slab = reinterpret_cast<float*>(stg);
  const int sw16 = (lane >> 1) & 15;
  const int rcol = lane & 15, rpar = lane >> 4;
  uint32_t pk[16 * NH];
#pragma unroll
  for (int hh = 0; hh < NH; ++hh) {
    float v[32];
    tmem_ld32(taddr + 32 * hh, v);
    if (P.bias) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias_blk + 32 * hh + i);
        v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
      }
    }
    if (hh == 0 && store_pending) {    // the previous TMA store must have read the tile before it is rewritten
      if (lane == 0) tma_store_wait_read();
      __syncwarp();
      store_pending = false;
    }
    if (P.stats) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int i = 0; i < 16; ++i) slab[lane * 16 + (i ^ sw16)] = valid ? v[h * 16 + i] : 0.f;
        __syncwarp();
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float t = slab[(2 * i + rpar) * 16 + (rcol ^ i)];
          s1 += t;
          s2 = fmaf(t, t, s2);
        }
        s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
        if (lane < 16) {
          atomicAdd(&sum_blk[32 * hh + h * 16 + lane], s1);
          if (!P.stats_sum_only) atomicAdd(&sq_blk[32 * hh + h * 16 + lane], s2);
        }
        __syncwarp();
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const __nv_bfloat162 p2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      pk[16 * hh + i] = *reinterpret_cast<const unsigned*>(&p2);
    }
  }
  const int sw = NH == 2 ? (lane & 7) : ((lane >> 1) & 3);
#pragma unroll
  for (int i = 0; i < 4 * NH; ++i)
    *reinterpret_cast<uint4*>(stg + lane * (64 * NH) + ((i ^ sw) << 4)) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
  fence_proxy_async();
  __syncwarp();
}

// Epilogue of the general kernel (k_conv_igemm, bf16 output): the two groups of four epilogue warps take ALTERNATE
// tiles (group g owns accumulator g), each warp all column blocks of its 32 pixel rows.  With one or two K steps per
// tile the kernel is bound by per-tile latency chains, not by bytes or MMAs (ncu r02: 30 % issue slots busy, every warp
// 8 cycles per instruction, epilogue warps ~350 instructions per tile): two tiles in the epilogue at a time, no
// divisions in the tile decode, and (WIDE) half as many TMA stores.
template <int N, bool WIDE>
__device__ __forceinline__ void igemm_epilogue_alt(const IgemmParams& P, const CUtensorMap* out_map, uint32_t tmem_base,
                                                   uint64_t* acc_full, uint64_t* acc_empty, float* s_stats,
                                                   const float* s_bias, uint8_t* stage_all, int warp, int lane) {
  constexpr int NH = WIDE ? 2 : 1, BW = 32 * NH;
  const int q = warp & 3;
  const int grp = (warp - 2) >> 2;
  const int m = q * 32 + lane;
  const int r = m >> 3, c = m & 7;
  uint8_t* stg = stage_all + (warp - 2) * EPI_STAGE_BYTES_WIDE;
  const int G = (int)gridDim.x;
  const int item0 = (int)blockIdx.x + grp * G;
  TileIter ti;
  ti.init(P.par_mode ? item0 >> 2 : item0, P.par_mode ? G >> 1 : 2 * G, P.tiles_x, P.tiles_y);
  bool store_pending = false;
  uint32_t acc_phase = 0;
  int it = grp;                                   // iteration index of the CTA's item sequence
  for (int item = item0; item < P.num_tiles; item += 2 * G, it += 2, ti.next()) {
    const int par = P.par_mode ? (((int)blockIdx.x & 3) + it) & 3 : 0;   // par_decode without the division
    const int y = ti.ty * TILE_H + r, x = ti.tx * TILE_W + c;
    const bool valid = y < P.Ht && x < P.Wt;
    mbar_wait(&acc_full[grp], acc_phase);
    tc_fence_after();
#pragma unroll 1
    for (int col = 0; col < N; col += BW) {
      const bool second = P.split_col > 0 && col >= P.split_col;
      const int oc0 = second ? col - P.split_col : col;
      epi_block<NH>(P, tmem_base + ((uint32_t)(q * 32) << 16) + grp * N + col, s_bias + col, s_stats + col, s_stats + N + col,
                    valid, stg, lane, store_pending);
      if (lane == 0) {
        tma_store_4d(second ? out_map + 1 : out_map + par, stg, oc0, ti.tx * TILE_W, ti.ty * TILE_H + 4 * q, ti.n);
        tma_store_commit();
      }
      store_pending = true;
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&acc_empty[grp]);
    acc_phase ^= 1;
  }
  if (lane == 0) tma_store_wait_all();
}

template <int N, int KC>
struct IgemmCfg {
  static constexpr int A_BYTES = TILE_M * KC * 2;
  static constexpr int B_BYTES = N * KC * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BUDGET = 190 * 1024;          // operand tiles: a ring of [A | B] stages, or (b_resident) A ring + all B tiles
  static constexpr int STAGES_RAW = BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM = BUDGET + EPI_SMEM_WIDE + 1024 /*alignment slack*/ + 256 /*barriers*/ + 3 * N * 4 + 64;
  static constexpr int TMEM_COLS = 2 * N < 32 ? 32 : 2 * N;
  static constexpr int SWIZZLE = KC * 2;             // bytes per pixel row: 128 or 64
  static constexpr int LAYOUT = KC == 64 ? 2 : 4;    // UMMA layout type
  static constexpr int SBO = 8 * KC * 2;             // 8 pixel rows of KC channels
};

template <int N, int KC>
__global__ void __launch_bounds__(IGEMM_THREADS, 1) k_conv_igemm(const __grid_constant__ IgemmMaps maps,
                                                                 const __grid_constant__ IgemmParams P) {
  using Cfg = IgemmCfg<N, KC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* tiles = smem;  // [STAGES][A | B], every tile 1024-byte aligned (A_BYTES, B_BYTES are multiples of 1024)
  uint8_t* epi_stage = smem + Cfg::BUDGET;   // 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + EPI_SMEM_WIDE);
  uint64_t* full = bars;                       // [STAGES]
  uint64_t* empty = bars + Cfg::STAGES;        // [STAGES]
  uint64_t* acc_full = bars + 2 * Cfg::STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint64_t* b_full = reinterpret_cast<uint64_t*>(tmem_slot + 2);   // resident weights have landed
  float* s_stats = reinterpret_cast<float*>(epi_stage + EPI_SMEM_WIDE + 256);  // [2][N]
  float* s_bias = s_stats + 2 * N;                                         // [N]
  // b_resident: the ring holds A tiles only (ring_stages of them), the nsteps weight tiles sit behind it
  const int ring = P.b_resident ? P.ring_stages : Cfg::STAGES;
  const int ring_stride = P.b_resident ? Cfg::A_BYTES : Cfg::STAGE_BYTES;
  uint8_t* b_res = tiles + ring * Cfg::A_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < MAX_AMAPS; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    // alt: each accumulator is drained by ONE group of four epilogue warps (igemm_epilogue_alt)
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], P.alt ? EPI_WARPS / 2 : EPI_WARPS); }
    mbar_init(b_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) s_stats[i] = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s_bias[i] = P.bias ? P.bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int G = (int)gridDim.x;
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (P.b_resident) {
        mbar_arrive_expect_tx(b_full, (uint32_t)(P.nsteps * Cfg::B_BYTES));
        for (int s = 0; s < P.nsteps; ++s) tma_load_2d(b_res + s * Cfg::B_BYTES, &maps.b, b_full, P.steps[s].wk, 0);
      }
      TileIter ti;
      ti.init(P.par_mode ? (int)blockIdx.x >> 2 : (int)blockIdx.x, P.par_mode ? G >> 2 : G, P.tiles_x, P.tiles_y);
      int it = 0;
      for (int item = blockIdx.x; item < P.num_tiles; item += G, ++it, ti.next()) {
        const int par = P.par_mode ? (((int)blockIdx.x & 3) + it) & 3 : 0;   // = par_decode(item), without the division
        const int s0 = P.par_mode ? P.seg[par] : 0, s1 = P.par_mode ? P.seg[par + 1] : P.nsteps;
        const int x0 = ti.tx * TILE_W, y0 = ti.ty * TILE_H, n = ti.n;
        for (int s = s0; s < s1; ++s) {
          const KStep st = P.steps[s];
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* a_dst = tiles + stage * ring_stride;
          uint8_t* b_dst = a_dst + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full[stage], P.b_resident ? Cfg::A_BYTES : Cfg::STAGE_BYTES);
          tma_load_4d(a_dst, &maps.a[st.map], &full[stage], st.c0, x0 + st.dx, y0 + st.dy, n);
          if (!P.b_resident) tma_load_2d(b_dst, &maps.b, &full[stage], st.wk, 0);
          if (++stage == ring) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: the whole warp runs the loop (warp-uniform
    // control flow, descriptors as (lo, hi) halves), one elected lane issues -- see k_conv_igemm_halo
    constexpr uint32_t idesc = make_idesc_bf16(TILE_M, N, 0, 0);
    const uint64_t proto = make_smem_desc(0, 16, Cfg::SBO, Cfg::LAYOUT);
    const uint32_t d_hi = (uint32_t)(proto >> 32);
    const uint32_t lo0 = (uint32_t)proto | ((smem_u32(tiles) & 0x3FFFFu) >> 4);
    const uint32_t b_res_lo = (uint32_t)proto | ((smem_u32(b_res) & 0x3FFFFu) >> 4);
    const uint32_t stride16 = (uint32_t)ring_stride >> 4;
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    if (P.b_resident) {
      mbar_wait(b_full, 0);
      tc_fence_after();
    }
    int it = 0;
    for (int item = blockIdx.x; item < P.num_tiles; item += G, ++it) {
      const int par = P.par_mode ? (((int)blockIdx.x & 3) + it) & 3 : 0;
      const int s0 = P.par_mode ? P.seg[par] : 0, s1 = P.par_mode ? P.seg[par + 1] : P.nsteps;
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + acc * N;
      for (int s = s0; s < s1; ++s) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t a_lo = lo0 + (uint32_t)stage * stride16;
        const uint32_t b_lo = P.b_resident ? b_res_lo + (uint32_t)s * (uint32_t)(Cfg::B_BYTES >> 4) : a_lo + (uint32_t)(Cfg::A_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < KC / 16; ++k)  // +32 bytes (2 x 16 B) along K per UMMA_K = 16 bf16
          umma_bf16_lohi_warp(d, a_lo + 2 * k, d_hi, b_lo + 2 * k, d_hi, idesc, (uint32_t)(((s - s0) | k) != 0));
        umma_commit_warp(&empty[stage]);  // frees the smem stage once these MMAs have read it
        if (++stage == ring) { stage = 0; phase ^= 1; }
      }
      umma_commit_warp(&acc_full[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    if (P.alt) {
      if constexpr (N % 64 == 0) {
        if (P.wide) igemm_epilogue_alt<N, true>(P, &maps.out, tmem_base, acc_full, acc_empty, s_stats, s_bias, epi_stage, warp, lane);
        else igemm_epilogue_alt<N, false>(P, &maps.out, tmem_base, acc_full, acc_empty, s_stats, s_bias, epi_stage, warp, lane);
      } else {
        igemm_epilogue_alt<N, false>(P, &maps.out, tmem_base, acc_full, acc_empty, s_stats, s_bias, epi_stage, warp, lane);
      }
    } else {
      igemm_epilogue<N>(P, &maps.out, tmem_base, acc_full, acc_empty, s_stats, s_bias, epi_stage, warp, lane);
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (P.stats) {
    for (int i = threadIdx.x; i < (P.stats_sum_only ? N : 2 * N); i += blockDim.x) atomicAdd(&P.stats[i], (double)s_stats[i]);
  }
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// Halo variant for 3x3 stride-1 convolutions (forward and data gradient).  Instead of nine shifted activation boxes
// per 64-channel chunk (each re-read from L2), ONE [18 x pitch pixels] x 64-channel halo box lands in shared memory
// and the nine taps are nine UMMA descriptors into it: tap (dy, dx) starts (dy*pitch + dx) pixel rows (128 B each)
// into the box, 8-pixel groups are `pitch` pixel rows apart (stride byte offset = pitch * 128).  The 128-byte swizzle
// is a function of the shared-memory address, so a descriptor that starts at a 128-byte (not 1024-byte) boundary
// reads exactly what TMA wrote.  Activation traffic from L2 drops ~4-6x; weights stream through their own ring, or
// stay resident for the whole kernel when all taps fit (64 -> 64 channels).
// DUAL: a work item is a pair of consecutive tiles that share every weight tile -- two halo boxes per stage, two
// accumulators per buffer (4 N TMEM columns, N <= 128).  The 128 / 256-channel layers stream their weights per tile
// (18 / 36 tiles of 16 / 32 KB for 128 pixels) and ran at 11 TB/s of L2 -> SM traffic (ncu l1tex__m_xbar2l1tex_read_bytes,
// profiles/r02_ncu_tensor_pipe_final.txt: 698 MB per launch for 134 MB of activations), the fabric's limit, with the
// tensor pipe 70 % active: M = 256 per weight tile halves that traffic.
template <int N, bool DUAL = false>
struct HaloCfg {
  static constexpr int HALO_BYTES = 23 * 1024;                     // 18 x 10 pixels x 128 B = 23040, 1024-aligned
  static constexpr int SLOT_BYTES = (DUAL ? 2 : 1) * HALO_BYTES;   // one stage: the halo of every tile of the item
  static constexpr int W_BYTES = N * 128;                          // one tap x one 64-channel chunk
  static constexpr int HALO_STAGES = DUAL ? 2 : 3;
  static constexpr int W_STAGES_RAW = (225 * 1024 - 2048 - 3 * N * 4 - EPI_SMEM - HALO_STAGES * SLOT_BYTES) / W_BYTES;
  static constexpr int W_STAGES = W_STAGES_RAW > 18 ? 18 : W_STAGES_RAW;
  static constexpr int TILE_BYTES = HALO_STAGES * SLOT_BYTES + W_STAGES * W_BYTES;
  static constexpr int SMEM = TILE_BYTES + EPI_SMEM + 1024 + 512 + 3 * N * 4;
  static constexpr int TMEM_COLS = (DUAL ? 4 : 2) * N < 32 ? 32 : (DUAL ? 4 : 2) * N;
};

template <int N, bool DUAL = false>
__global__ void __launch_bounds__(IGEMM_THREADS, 1) k_conv_igemm_halo(const __grid_constant__ IgemmMaps maps,
                                                                      const __grid_constant__ IgemmParams P) {
  using Cfg = HaloCfg<N, DUAL>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* halos = smem;
  uint8_t* wts = smem + Cfg::HALO_STAGES * Cfg::SLOT_BYTES;
  uint8_t* epi_stage = smem + Cfg::TILE_BYTES;              // 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + EPI_SMEM);
  uint64_t* h_full = bars;                                  // [HALO_STAGES]
  uint64_t* h_empty = h_full + Cfg::HALO_STAGES;            // [HALO_STAGES]
  uint64_t* w_full = h_empty + Cfg::HALO_STAGES;            // [W_STAGES]
  uint64_t* w_empty = w_full + Cfg::W_STAGES;               // [W_STAGES]
  uint64_t* acc_full = w_empty + Cfg::W_STAGES;             // [2]
  uint64_t* acc_empty = acc_full + 2;                       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_stats = reinterpret_cast<float*>(epi_stage + EPI_SMEM + 512);
  float* s_bias = s_stats + 2 * N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t halo_tx = 18u * (uint32_t)P.halo_pitch * 128u;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < MAX_AMAPS; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    for (int i = 0; i < Cfg::HALO_STAGES; ++i) { mbar_init(&h_full[i], 1); mbar_init(&h_empty[i], 1); }
    for (int i = 0; i < Cfg::W_STAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) s_stats[i] = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s_bias[i] = P.bias ? P.bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int hs = 0, ws = 0;
      uint32_t hph = 0, wph = 0;
      bool first_tile = true;
      const int n_items = DUAL ? P.num_tiles / 2 : P.num_tiles;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int tile = DUAL ? 2 * item : item;
        const int tx = tile % P.tiles_x, t2 = tile / P.tiles_x, ty = t2 % P.tiles_y, n = t2 / P.tiles_y;
        const int x0 = tx * TILE_W - 1, y0 = ty * TILE_H - 1;
        int x0b = 0, y0b = 0, nb = 0;             // DUAL: the second tile of the pair
        if (DUAL) {
          const int tb = tile + 1;
          const int txb = tb % P.tiles_x, t2b = tb / P.tiles_x;
          x0b = txb * TILE_W - 1; y0b = (t2b % P.tiles_y) * TILE_H - 1; nb = t2b / P.tiles_y;
        }
        for (int s = 0; s < P.nsteps; ++s) {
          const KStep st = P.steps[s];
          mbar_wait(&h_empty[hs], hph ^ 1);
          mbar_arrive_expect_tx(&h_full[hs], (DUAL ? 2u : 1u) * halo_tx);
          tma_load_4d(halos + hs * Cfg::SLOT_BYTES, &maps.a[st.map], &h_full[hs], st.c0, x0, y0, n);
          if (DUAL) tma_load_4d(halos + hs * Cfg::SLOT_BYTES + Cfg::HALO_BYTES, &maps.a[st.map], &h_full[hs], st.c0, x0b, y0b, nb);
          if (++hs == Cfg::HALO_STAGES) { hs = 0; hph ^= 1; }
          if (!P.w_resident || first_tile) {
            for (int t = 0; t < 9; ++t) {
              mbar_wait(&w_empty[ws], wph ^ 1);
              mbar_arrive_expect_tx(&w_full[ws], Cfg::W_BYTES);
              tma_load_2d(wts + ws * Cfg::W_BYTES, &maps.b, &w_full[ws], st.wk + t * P.tap_k_stride, 0);
              if (++ws == Cfg::W_STAGES) { ws = 0; wph ^= 1; }
            }
          }
        }
        first_tile = false;
      }
    }
  } else if (warp == 1) {
    {
      // The issue loop is on the critical path of the small-N tiles (an N = 64 MMA takes ~70 cycles).  The whole warp
      // runs the loop (warp-uniform control flow -> uniform registers), one elected lane issues; descriptors are
      // (lo, hi) register pairs whose hi words never change: per MMA one 32-bit add per operand.
      constexpr uint32_t idesc = make_idesc_bf16(TILE_M, N, 0, 0);
      const uint64_t a_proto = make_smem_desc(0, 16, (uint32_t)P.halo_pitch * 128u, 2, 0);
      const uint64_t b_proto = make_smem_desc(0, 16, 1024, 2);
      const uint32_t a_hi = (uint32_t)(a_proto >> 32), b_hi = (uint32_t)(b_proto >> 32);
      const uint32_t a_lo0 = (uint32_t)a_proto | ((smem_u32(halos) & 0x3FFFFu) >> 4);
      const uint32_t b_lo0 = (uint32_t)b_proto | ((smem_u32(wts) & 0x3FFFFu) >> 4);
      uint32_t tap_off[9];   // start of tap t inside the halo box, in 16-byte units
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int ty3 = t / 3, tx3 = t - 3 * ty3;
        const int hy = P.halo_flip ? 2 - ty3 : ty3, hx = P.halo_flip ? 2 - tx3 : tx3;
        tap_off[t] = (uint32_t)(hy * P.halo_pitch + hx) * 8u;
      }
      const bool resident = P.w_resident != 0;
      int hs = 0, ws = 0, acc = 0;
      uint32_t hph = 0, wph = 0, acc_phase = 0;
      bool first_tile = true;
      const int n_items = DUAL ? P.num_tiles / 2 : P.num_tiles;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + (DUAL ? 2 * acc : acc) * N;
        uint32_t accum = 0u;
        if (resident) ws = 0;
        for (int s = 0; s < P.nsteps; ++s) {
          mbar_wait(&h_full[hs], hph);
          tc_fence_after();
          const uint32_t ah = a_lo0 + (uint32_t)hs * (uint32_t)(Cfg::SLOT_BYTES >> 4);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            if (!resident || first_tile) { mbar_wait(&w_full[ws], wph); tc_fence_after(); }
            const uint32_t a_lo = ah + tap_off[t];
            const uint32_t b_lo = b_lo0 + (uint32_t)ws * (uint32_t)(Cfg::W_BYTES >> 4);
            umma_bf16_lohi_warp(d, a_lo, a_hi, b_lo, b_hi, idesc, accum);
            umma_bf16_lohi_warp(d, a_lo + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
            umma_bf16_lohi_warp(d, a_lo + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
            umma_bf16_lohi_warp(d, a_lo + 6, a_hi, b_lo + 6, b_hi, idesc, 1u);
            if (DUAL) {   // the second tile of the pair against the same weight tile
              const uint32_t a2 = a_lo + (uint32_t)(Cfg::HALO_BYTES >> 4);
              umma_bf16_lohi_warp(d + N, a2, a_hi, b_lo, b_hi, idesc, accum);
              umma_bf16_lohi_warp(d + N, a2 + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
              umma_bf16_lohi_warp(d + N, a2 + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
              umma_bf16_lohi_warp(d + N, a2 + 6, a_hi, b_lo + 6, b_hi, idesc, 1u);
            }
            accum = 1u;
            if (!resident) umma_commit_warp(&w_empty[ws]);
            if (++ws == Cfg::W_STAGES) { ws = 0; wph ^= 1; }
          }
          umma_commit_warp(&h_empty[hs]);
          if (++hs == Cfg::HALO_STAGES) { hs = 0; hph ^= 1; }
        }
        umma_commit_warp(&acc_full[acc]);
        first_tile = false;
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    igemm_epilogue<N, false, DUAL>(P, &maps.out, tmem_base, acc_full, acc_empty, s_stats, s_bias, epi_stage, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (P.stats) {
    for (int i = threadIdx.x; i < (P.stats_sum_only ? N : 2 * N); i += blockDim.x) atomicAdd(&P.stats[i], (double)s_stats[i]);
  }
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// Row-pair halo variant for 64 output channels.  An M = 128 x N = 64 MMA is bound by the A-operand fetch from shared
// memory (~70 cycles for 32 cycles of math, tools/mma_chain_bench.cu), so the N = 64 layers ran at half the tensor
// rate.  Here one MMA produces TWO output rows: the 128 A rows are the 16 EVEN rows x 8 columns of a 32 x 8 pixel tile
// (8-pixel groups 2 halo rows apart), the accumulator has 128 columns = [output row 2r | output row 2r + 1] x 64
// channels, and the B operand of halo row offset hy is [W(ky = hy) | W(ky = hy - 1)] -- two taps that sit next to each
// other in shared memory, one N = 128 operand.  hy = 0 and hy = 3 touch only one of the two output rows (N = 64 MMAs
// into one half of the accumulator).  Per 64-channel chunk and kx: 2 x N=128 + 2 x N=64 MMA groups instead of 6 x N=64
// for the same 256 pixels.  Weight slot p of a kx triple holds ky = 2 - p (forward) or ky = p (data gradient, flipped
// taps), so that hy = 1 reads slots (1, 2), hy = 2 slots (0, 1), hy = 0 slot 2 and hy = 3 slot 0 in both directions.
// WIDE: 64-channel epilogue stores (igemm_epilogue_wide: one 128-byte row per output pixel instead of two 64-byte ones);
// the 16 KB of extra staging come out of the weight ring (four kx triples instead of five; the nine taps of a
// single-chunk layer still stay resident).
template <bool WIDE>
struct HaloPairCfgT {
  static constexpr int N = 64;                                     // output channels
  static constexpr int ROWS = 2 * TILE_H + 2;                      // 34 halo rows
  static constexpr int HALO_BYTES = 43 * 1024;                     // 34 x 10 pixels x 128 B = 43520, 1024-aligned
  static constexpr int W_BYTES = N * 128;                          // one tap x one 64-channel chunk
  static constexpr int HALO_STAGES = 2;
  static constexpr int W_STAGES = WIDE ? 12 : 15;                  // kx triples
  static constexpr int EPI_BYTES = WIDE ? EPI_SMEM_WIDE : EPI_SMEM;
  static constexpr int TILE_BYTES = HALO_STAGES * HALO_BYTES + W_STAGES * W_BYTES;
  static constexpr int SMEM = TILE_BYTES + EPI_BYTES + 1024 + 512 + 3 * N * 4;
  static constexpr int TMEM_COLS = 4 * N;                          // two accumulators of 2 x 64 columns
};
using HaloPairCfg = HaloPairCfgT<false>;

template <bool WIDE>
__global__ void __launch_bounds__(IGEMM_THREADS, 1) k_conv_igemm_halo_pair(const __grid_constant__ IgemmMaps maps,
                                                                           const __grid_constant__ IgemmParams P) {
  using Cfg = HaloPairCfgT<WIDE>;
  constexpr int N = Cfg::N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* halos = smem;
  uint8_t* wts = smem + Cfg::HALO_STAGES * Cfg::HALO_BYTES;
  uint8_t* epi_stage = smem + Cfg::TILE_BYTES;              // 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + Cfg::EPI_BYTES);
  uint64_t* h_full = bars;                                  // [HALO_STAGES]
  uint64_t* h_empty = h_full + Cfg::HALO_STAGES;            // [HALO_STAGES]
  uint64_t* w_full = h_empty + Cfg::HALO_STAGES;            // [W_STAGES]
  uint64_t* w_empty = w_full + Cfg::W_STAGES;               // [W_STAGES]
  uint64_t* acc_full = w_empty + Cfg::W_STAGES;             // [2]
  uint64_t* acc_empty = acc_full + 2;                       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_stats = reinterpret_cast<float*>(epi_stage + Cfg::EPI_BYTES + 512);
  float* s_bias = s_stats + 2 * N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t halo_tx = (uint32_t)Cfg::ROWS * (uint32_t)P.halo_pitch * 128u;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < MAX_AMAPS; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    for (int i = 0; i < Cfg::HALO_STAGES; ++i) { mbar_init(&h_full[i], 1); mbar_init(&h_empty[i], 1); }
    for (int i = 0; i < Cfg::W_STAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) s_stats[i] = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s_bias[i] = P.bias ? P.bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int hs = 0, ws = 0;
      uint32_t hph = 0, wph = 0;
      bool first_tile = true;
      for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
        const int tx = tile % P.tiles_x, t2 = tile / P.tiles_x, ty = t2 % P.tiles_y, n = t2 / P.tiles_y;
        const int x0 = tx * TILE_W - 1, y0 = ty * 2 * TILE_H - 1;
        for (int s = 0; s < P.nsteps; ++s) {
          const KStep st = P.steps[s];
          mbar_wait(&h_empty[hs], hph ^ 1);
          mbar_arrive_expect_tx(&h_full[hs], halo_tx);
          tma_load_4d(halos + hs * Cfg::HALO_BYTES, &maps.a[st.map], &h_full[hs], st.c0, x0, y0, n);
          if (++hs == Cfg::HALO_STAGES) { hs = 0; hph ^= 1; }
          if (!P.w_resident || first_tile) {
            for (int hx = 0; hx < 3; ++hx) {
              for (int p = 0; p < 3; ++p) {
                const int ky = P.halo_flip ? p : 2 - p, kx = P.halo_flip ? 2 - hx : hx;
                mbar_wait(&w_empty[ws], wph ^ 1);
                mbar_arrive_expect_tx(&w_full[ws], Cfg::W_BYTES);
                tma_load_2d(wts + ws * Cfg::W_BYTES, &maps.b, &w_full[ws], st.wk + (ky * 3 + kx) * P.tap_k_stride, 0);
                if (++ws == Cfg::W_STAGES) { ws = 0; wph ^= 1; }
              }
            }
          }
        }
        first_tile = false;
      }
    }
  } else if (warp == 1) {
    // whole-warp issue loop, (lo, hi) descriptors (see k_conv_igemm_halo)
    constexpr uint32_t idesc128 = make_idesc_bf16(TILE_M, 2 * N, 0, 0), idesc64 = make_idesc_bf16(TILE_M, N, 0, 0);
    const uint64_t a_proto = make_smem_desc(0, 16, 2u * (uint32_t)P.halo_pitch * 128u, 2, 0);   // 8-pixel groups two halo rows apart
    const uint64_t b_proto = make_smem_desc(0, 16, 1024, 2);
    const uint32_t a_hi = (uint32_t)(a_proto >> 32), b_hi = (uint32_t)(b_proto >> 32);
    const uint32_t a_lo0 = (uint32_t)a_proto | ((smem_u32(halos) & 0x3FFFFu) >> 4);
    const uint32_t b_lo0 = (uint32_t)b_proto | ((smem_u32(wts) & 0x3FFFFu) >> 4);
    const uint32_t row16 = (uint32_t)P.halo_pitch * 8u;      // one halo row in 16-byte units
    constexpr uint32_t slot16 = (uint32_t)(Cfg::W_BYTES >> 4);
    const bool resident = P.w_resident != 0;
    int hs = 0, ws = 0, acc = 0;
    uint32_t hph = 0, wph = 0, acc_phase = 0;
    bool first_tile = true;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + acc * 2 * N;
      uint32_t accum = 0u;
      if (resident) ws = 0;
      for (int s = 0; s < P.nsteps; ++s) {
        mbar_wait(&h_full[hs], hph);
        tc_fence_after();
        const uint32_t ah = a_lo0 + (uint32_t)hs * (uint32_t)(Cfg::HALO_BYTES >> 4);
#pragma unroll
        for (int hx = 0; hx < 3; ++hx) {
          if (!resident || first_tile) {
            mbar_wait(&w_full[ws], wph); mbar_wait(&w_full[ws + 1], wph); mbar_wait(&w_full[ws + 2], wph);
            tc_fence_after();
          }
          const uint32_t b0 = b_lo0 + (uint32_t)ws * slot16;
          const uint32_t a0 = ah + (uint32_t)hx * 8u;
#pragma unroll
          for (int k = 0; k < 4; ++k)   // hy = 1: rows (2r | 2r+1) <- taps in slots (1 | 2); the very first MMA overwrites
            umma_bf16_lohi_warp(d, a0 + row16 + 2 * k, a_hi, b0 + slot16 + 2 * k, b_hi, idesc128, accum | (uint32_t)(k > 0));
#pragma unroll
          for (int k = 0; k < 4; ++k)   // hy = 2: slots (0 | 1)
            umma_bf16_lohi_warp(d, a0 + 2 * row16 + 2 * k, a_hi, b0 + 2 * k, b_hi, idesc128, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k)   // hy = 0: even output rows only <- slot 2
            umma_bf16_lohi_warp(d, a0 + 2 * k, a_hi, b0 + 2 * slot16 + 2 * k, b_hi, idesc64, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k)   // hy = 3: odd output rows only <- slot 0
            umma_bf16_lohi_warp(d + N, a0 + 3 * row16 + 2 * k, a_hi, b0 + 2 * k, b_hi, idesc64, 1u);
          accum = 1u;
          if (!resident) { umma_commit_warp(&w_empty[ws]); umma_commit_warp(&w_empty[ws + 1]); umma_commit_warp(&w_empty[ws + 2]); }
          ws += 3;
          if (ws == Cfg::W_STAGES) { ws = 0; wph ^= 1; }
        }
        umma_commit_warp(&h_empty[hs]);
        if (++hs == Cfg::HALO_STAGES) { hs = 0; hph ^= 1; }
      }
      umma_commit_warp(&acc_full[acc]);
      first_tile = false;
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    if constexpr (WIDE) igemm_epilogue_wide<2 * N, true>(P, &maps.out, tmem_base, acc_full, acc_empty, s_stats, s_bias, epi_stage, warp, lane);
    else igemm_epilogue<2 * N, true>(P, &maps.out, tmem_base, acc_full, acc_empty, s_stats, s_bias, epi_stage, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (P.stats) {
    for (int i = threadIdx.x; i < (P.stats_sum_only ? N : 2 * N); i += blockDim.x) atomicAdd(&P.stats[i], (double)s_stats[i]);
  }
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

static bool env_on(const char* name) {
  const char* e = getenv(name);
  return !(e && atoi(e) == 0);
}

// bf16 output of the row-pair kernel as [ch, x, row parity, row / 2, n], box [32 (wide: 64), 8, 1, 4, 1], SWIZZLE_64B (128B)
static int build_out_map_pair(IgemmMaps& maps, const IgemmParams& P, int N, bool wide) {
  const uint64_t dims[5] = {(uint64_t)N, (uint64_t)P.Wt, 2, (uint64_t)P.Ht / 2, (uint64_t)P.n_img};
  const uint64_t str[4] = {(uint64_t)P.out_pix * 2, (uint64_t)P.out_row * 2, (uint64_t)P.out_row * 4, (uint64_t)P.out_img * 2};
  const uint32_t box[5] = {wide ? 64u : 32u, TILE_W, 1, 4, 1};
  return make_tensor_map_bf16(&maps.out, P.out, 5, dims, str, box, wide ? 128 : 64);
}

template <bool WIDE>
static int launch_igemm_halo_pair_t(IgemmMaps& maps, IgemmParams& P, cudaStream_t st) {
  using Cfg = HaloPairCfgT<WIDE>;
  if (int rc = build_out_map_pair(maps, P, Cfg::N, WIDE)) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_igemm_halo_pair<WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) { set_error("igemm_halo_pair: cannot reserve %d bytes of shared memory: %s", Cfg::SMEM, cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    configured = true;
  }
  P.tiles_y = (P.Ht + 2 * TILE_H - 1) / (2 * TILE_H);
  P.num_tiles = P.tiles_x * P.tiles_y * P.n_img;
  P.w_resident = (9 * P.nsteps <= Cfg::W_STAGES) ? 1 : 0;
  int grid = sm_count();
  if (grid > P.num_tiles) grid = P.num_tiles;
  k_conv_igemm_halo_pair<WIDE><<<grid, IGEMM_THREADS, Cfg::SMEM, st>>>(maps, P);
  add_launches(1);
  return check_launch("conv_igemm_halo_pair");
}

// 64-channel epilogue stores in the row-pair kernel: 3.02 -> 2.95 ms per step over its 19 launches
// (profiles/r02_tma_requests_ab.txt); DFB_PAIR_WIDE=0 -> 32-channel stores and a weight ring of five kx triples (A/B)
static int launch_igemm_halo_pair(IgemmMaps& maps, IgemmParams& P, cudaStream_t st) {
  if (env_on("DFB_PAIR_WIDE")) return launch_igemm_halo_pair_t<true>(maps, P, st);
  return launch_igemm_halo_pair_t<false>(maps, P, st);
}

// DFB_HALO_PAIR=0 keeps the 64-channel layers on k_conv_igemm_halo<64> (A/B comparisons); default = row-pair kernel
static bool halo_pair_enabled() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("DFB_HALO_PAIR");
    mode = e ? atoi(e) : 1;
  }
  return mode != 0;
}

// k_conv_igemm A/B switches (read per call; tests toggle them): DFB_EPI_ALT=0 -> all eight epilogue warps on every tile
// (and 32-channel stores), DFB_EPI_WIDE=0 -> 32-channel epilogue stores, DFB_IGEMM_B_RESIDENT=0 -> weight tiles re-loaded
// with every A tile.
static bool igemm_alt(const IgemmParams& P) { return !P.out_fp32 && P.halo_pitch == 0 && env_on("DFB_EPI_ALT"); }
static bool igemm_wide(const IgemmParams& P, int N) {
  return igemm_alt(P) && N % 64 == 0 && P.split_col % 64 == 0 && env_on("DFB_EPI_WIDE");
}
static bool igemm_resident_enabled() { return env_on("DFB_IGEMM_B_RESIDENT"); }

// Output tensor map of a launch: element (n, y, x, ch) of the tile space lives at
// out + n*img + (y*sy+oy)*row + (x*sx+ox)*pix + ch  (bf16) -> dims [ch, x, y, n], box [32, 8, 4, 1], SWIZZLE_64B.
static int build_out_map(IgemmMaps& maps, const IgemmParams& P, int N) {
  if (P.out_fp32) return DFB_OK;
  const uint64_t dims[4] = {(uint64_t)(P.split_col > 0 ? P.split_col : N), (uint64_t)P.Wt, (uint64_t)P.Ht, (uint64_t)P.n_img};
  const uint64_t str[3] = {(uint64_t)P.out_pix * P.sx * 2, (uint64_t)P.out_row * P.sy * 2, (uint64_t)P.out_img * 2};
  const uint32_t box[4] = {P.wide ? 64u : 32u, TILE_W, 4, 1};
  const int swz = P.wide ? 128 : 64;
  const char* base = (const char*)P.out + ((size_t)P.oy * P.out_row + (size_t)P.ox * P.out_pix) * 2;
  if (int rc = make_tensor_map_bf16(&maps.out, base, 4, dims, str, box, swz)) return rc;
  if (P.split_col > 0) {   // second output: the remaining N - split_col channels, densely packed NHWC
    const uint64_t c2 = (uint64_t)(N - P.split_col);
    const uint64_t d2[4] = {c2, (uint64_t)P.Wt, (uint64_t)P.Ht, (uint64_t)P.n_img};
    const uint64_t s2[3] = {c2 * 2, c2 * 2 * P.Wt, c2 * 2 * P.Wt * P.Ht};
    return make_tensor_map_bf16(&maps.out2, P.out2, 4, d2, s2, box, swz);
  }
  return DFB_OK;
}

template <int N, bool DUAL = false>
static int launch_igemm_halo(IgemmMaps& maps, IgemmParams& P, cudaStream_t st) {
  using Cfg = HaloCfg<N, DUAL>;
  if (int rc = build_out_map(maps, P, N)) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_igemm_halo<N, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) { set_error("igemm_halo<%d>: cannot reserve %d bytes of shared memory: %s", N, Cfg::SMEM, cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    configured = true;
  }
  P.w_resident = (9 * P.nsteps <= Cfg::W_STAGES) ? 1 : 0;
  const int items = DUAL ? P.num_tiles / 2 : P.num_tiles;
  int grid = sm_count();
  if (grid > items) grid = items;
  k_conv_igemm_halo<N, DUAL><<<grid, IGEMM_THREADS, Cfg::SMEM, st>>>(maps, P);
  add_launches(1);
  return check_launch("conv_igemm_halo");
}

// measured (profiles/r02_halo_dual_ab.txt): k_conv_igemm_halo<128> 2.82 -> 2.67 ms per step over its 25 launches, step
// 28.4 -> 28.25 ms.  The rest of the gap to the tensor peak is shared-memory bandwidth: an M = 128 x N = 128 x K = 16 MMA reads
// 8 KB of operands for 64 cycles of math (128 B per cycle, the limit) -- the remedy is cta_group::2 (half of B per CTA).
static bool halo_dual_enabled() { return env_on("DFB_HALO_DUAL"); }   // read per call: tests toggle it

static int dispatch_igemm_halo(int N, IgemmMaps& maps, IgemmParams& P, cudaStream_t st) {
  if (N == 32) return launch_igemm_halo<32>(maps, P, st);
  if (N == 64) return launch_igemm_halo<64>(maps, P, st);
  // 128 output channels: pairs of tiles share every weight tile (HaloCfg DUAL); DFB_HALO_DUAL=0 -> one tile per item (A/B)
  if (N == 128 && P.num_tiles % 2 == 0 && P.num_tiles >= 2 && halo_dual_enabled()) return launch_igemm_halo<128, true>(maps, P, st);
  if (N == 128) return launch_igemm_halo<128>(maps, P, st);
  if (N == 256) return launch_igemm_halo<256>(maps, P, st);
  set_error("conv_igemm_halo: unsupported N=%d", N);
  return DFB_ERR_UNSUPPORTED;
}

// DFB_CONV_HALO=0 falls back to per-tap activation boxes (k_conv_igemm) for A/B comparisons; default = halo.
// Measured on B200 (tools/diag_conv.py): descriptors that start at a 128-byte (non-1024-byte) boundary read the
// TMA-written swizzle correctly with the base-offset field left at 0, for both 1280- and 2048-byte group strides;
// setting base-offset = (start >> 7) & 7 gives wrong results -- the swizzle XOR is taken from the absolute address.
// DFB_DGRAD_PAR_MERGE=0: the stride-2 data gradient as four launches, one per parity plane (A/B); default = one launch.
static bool par_merge_enabled() {
  const char* e = getenv("DFB_DGRAD_PAR_MERGE");   // read per call (six launches per step): tests toggle it
  return !(e && atoi(e) == 0);
}

static int halo_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("DFB_CONV_HALO");
    mode = e ? atoi(e) : 2;
  }
  return mode;
}

// ------------------------------------------------------------------------------------------------
// Weight packing: torch [Cout, Cin, kh, kw] fp32 -> bf16 GEMM-B matrices, K contiguous.
//   forward:  Wf[co][tap * Cin + ci]            (tap = ky * kw + kx)
//   dgrad:    Wd[ci][tap * Cout + co]           (same tap numbering; the step table flips the shifts)
__global__ void __launch_bounds__(256) k_pack_weights(const float* __restrict__ w, int Cout, int Cin, int taps, int split3,
                                                      __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd) {
  // split3: every row holds its hi part followed by its lo part (lo = bf16(w - hi)): rows are 2K long
  const long long total = (long long)Cout * Cin * taps;
  const long long kf = (long long)taps * Cin, kd = (long long)taps * Cout;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(e % taps);
    const int ci = (int)((e / taps) % Cin);
    const int co = (int)(e / ((long long)taps * Cin));
    const float f = w[e];
    const __nv_bfloat16 v = __float2bfloat16_rn(f);
    const __nv_bfloat16 lo = __float2bfloat16_rn(f - __bfloat162float(v));
    if (wf) {
      const long long o = (long long)co * kf * (split3 ? 2 : 1) + (long long)t * Cin + ci;
      wf[o] = v;
      if (split3) wf[o + kf] = lo;
    }
    if (wd) {
      const long long o = (long long)ci * kd * (split3 ? 2 : 1) + (long long)t * Cout + co;
      wd[o] = v;
      if (split3) wd[o + kd] = lo;
    }
  }
}

// All convolution weights of a model in ONE launch (run at the start of every forward: the optimizer has just changed the
// fp32 masters).  table[i].first = index of weight i's first element in the concatenated element space; a thread finds its
// weight by binary search over the (<= 64 entry) table held in shared memory.
__global__ void __launch_bounds__(256) k_pack_weights_multi(const dfb_pack_desc* __restrict__ table, int n_w, long long total,
                                                            int split3) {
  __shared__ dfb_pack_desc sd[64];
  for (int i = threadIdx.x; i < n_w; i += blockDim.x) sd[i] = table[i];
  __syncthreads();
  const int m = split3 ? 2 : 1;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    int lo_i = 0, hi_i = n_w - 1;
    while (lo_i < hi_i) {
      const int mid = (lo_i + hi_i + 1) >> 1;
      if (sd[mid].first <= g) lo_i = mid; else hi_i = mid - 1;
    }
    const dfb_pack_desc& d = sd[lo_i];
    const long long e = g - d.first;
    const int taps = d.ksize * d.ksize, Cin = d.cin, Cout = d.cout;
    const int t = (int)(e % taps);
    const int ci = (int)((e / taps) % Cin);
    const int co = (int)(e / ((long long)taps * Cin));
    const float f = d.w[e];
    const __nv_bfloat16 v = __float2bfloat16_rn(f);
    const __nv_bfloat16 l = __float2bfloat16_rn(f - __bfloat162float(v));
    const long long kf = (long long)taps * Cin, kd = (long long)taps * Cout;
    __nv_bfloat16* wf = (__nv_bfloat16*)d.w_fwd;
    __nv_bfloat16* wd = (__nv_bfloat16*)d.w_dgrad;
    if (wf) {
      const long long o = (long long)co * kf * m + (long long)t * Cin + ci;
      wf[o] = v;
      if (split3) wf[o + kf] = l;
    }
    if (wd) {
      const long long o = (long long)ci * kd * m + (long long)t * Cout + co;
      wd[o] = v;
      if (split3) wd[o + kd] = l;
    }
  }
}

// fp32 -> (hi, lo) bf16 pair with hi + lo = x to ~16 significant bits
__global__ void __launch_bounds__(256) k_split_bf16x2(const float4* __restrict__ x, long long n4, uint2* __restrict__ hi,
                                                      uint2* __restrict__ lo) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + e);
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
    const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
    uint2 a, b;
    a.x = *reinterpret_cast<const unsigned*>(&h0); a.y = *reinterpret_cast<const unsigned*>(&h1);
    b.x = *reinterpret_cast<const unsigned*>(&l0); b.y = *reinterpret_cast<const unsigned*>(&l1);
    hi[e] = a; lo[e] = b;
  }
}

template <int N, int KC>
static int launch_igemm(IgemmMaps& maps, const IgemmParams& P, cudaStream_t st) {
  using Cfg = IgemmCfg<N, KC>;
  IgemmParams Q = P;
  Q.alt = igemm_alt(Q) ? 1 : 0;
  Q.wide = igemm_wide(Q, N) ? 1 : 0;
  if (!Q.par_mode) {   // (parity-merged launches come with their four output maps built)
    if (int rc = build_out_map(maps, Q, N)) return rc;
  }
  Q.b_resident = 0;
  Q.ring_stages = Cfg::STAGES;
  if (!Q.out_fp32 && igemm_resident_enabled()) {
    // weights once per CTA instead of once per tile and K step, when they leave room for >= 3 A tiles in the ring
    const int res = Q.nsteps * Cfg::B_BYTES;
    const int ring = (Cfg::BUDGET - res) / Cfg::A_BYTES;
    if (res < Cfg::BUDGET && ring >= 3) { Q.b_resident = 1; Q.ring_stages = ring > Cfg::STAGES ? Cfg::STAGES : ring; }
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_igemm<N, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) { set_error("igemm<%d,%d>: cannot reserve %d bytes of shared memory: %s", N, KC, Cfg::SMEM, cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    configured = true;
  }
  int grid = sm_count();
  if (grid > P.num_tiles) grid = P.num_tiles;
  if (Q.par_mode) grid &= ~3;   // par_decode: the four parity planes of a region must share the iteration index
  k_conv_igemm<N, KC><<<grid, IGEMM_THREADS, Cfg::SMEM, st>>>(maps, Q);
  add_launches(1);
  return check_launch("conv_igemm");
}

static int dispatch_igemm(int N, int KC, IgemmMaps& maps, const IgemmParams& P, cudaStream_t st) {
#define DFB_CASE(n, kc) if (N == n && KC == kc) return launch_igemm<n, kc>(maps, P, st)
  DFB_CASE(32, 64); DFB_CASE(64, 64); DFB_CASE(128, 64); DFB_CASE(256, 64);
  DFB_CASE(32, 32); DFB_CASE(64, 32); DFB_CASE(128, 32); DFB_CASE(256, 32);
#undef DFB_CASE
  set_error("conv_igemm: unsupported tile N=%d KC=%d (N in {32,64,128,256}, KC in {32,64})", N, KC);
  return DFB_ERR_UNSUPPORTED;
}

}  // namespace tc
}  // namespace dfb

using namespace dfb;
using namespace dfb::tc;

extern "C" int dfb_split_bf16x2(const float* x, long long n, void* hi, void* lo, void* stream_) {
  if (n % 4) { set_error("dfb_split_bf16x2: element count must be a multiple of 4"); return DFB_ERR_ARG; }
  if (n == 0) return DFB_OK;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  k_split_bf16x2<<<(int)blocks, 256, 0, (cudaStream_t)stream_>>>((const float4*)x, n / 4, (uint2*)hi, (uint2*)lo);
  add_launches(1);
  return check_launch("dfb_split_bf16x2");
}

extern "C" int dfb_conv_pack_weights_multi(const dfb_pack_desc* table_dev, int n_weights, long long total_elems, int split3,
                                           void* stream_) {
  if (n_weights <= 0 || n_weights > 64 || total_elems <= 0 || !table_dev) {
    set_error("dfb_conv_pack_weights_multi: 1..64 weights and a device table are required");
    return DFB_ERR_ARG;
  }
  long long blocks = (total_elems + 255) / 256;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  k_pack_weights_multi<<<(int)blocks, 256, 0, (cudaStream_t)stream_>>>(table_dev, n_weights, total_elems, split3);
  add_launches(1);
  return check_launch("dfb_conv_pack_weights_multi");
}

extern "C" int dfb_conv_pack_weights(const float* w, int cout, int cin, int ksize, int split3, void* w_fwd, void* w_dgrad,
                                     void* stream_) {
  if (cout <= 0 || cin <= 0 || (ksize != 1 && ksize != 3)) { set_error("dfb_conv_pack_weights: bad sizes"); return DFB_ERR_ARG; }
  const long long total = (long long)cout * cin * ksize * ksize;
  long long blocks = (total + 255) / 256;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  k_pack_weights<<<(int)blocks, 256, 0, (cudaStream_t)stream_>>>(w, cout, cin, ksize * ksize, split3, (__nv_bfloat16*)w_fwd,
                                                                 (__nv_bfloat16*)w_dgrad);
  add_launches(1);
  return check_launch("dfb_conv_pack_weights");
}

// mode 0: forward  y = conv(x; W) (+ bias)           x: n_src sources [n,H,W,cin_s] concatenated along channels
// mode 1: dgrad    gx = conv^T(gy; W)                 x: ONE source = gy [n,Ho,Wo,cout]; output = gx [n,H,W,cin]
//                                                      (for concatenated inputs call once per source with
//                                                       cin_off / cin_total selecting the weight slice)
extern "C" int dfb_conv2d(const dfb_conv_args* a, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (!a) { set_error("dfb_conv2d: null args"); return DFB_ERR_ARG; }
  const int k = a->ksize, s = a->stride, taps = k * k;
  if ((k != 1 && k != 3) || (s != 1 && s != 2) || a->n_src < 1 || a->n_src > 2) { set_error("dfb_conv2d: ksize must be 1 or 3, stride 1 or 2, 1-2 sources"); return DFB_ERR_ARG; }
  const int H = a->H, W = a->W;                       // input (x) spatial size
  const int Ho = (H + 2 * (k / 2) - k) / s + 1, Wo = (W + 2 * (k / 2) - k) / s + 1;
  int cin_total = 0;
  for (int i = 0; i < a->n_src; ++i) cin_total += a->cin[i];
  if (a->cin_total > 0) cin_total = a->cin_total;

  IgemmMaps maps;
  IgemmParams P;
  memset(&maps, 0, sizeof(maps));
  memset(&P, 0, sizeof(P));
  P.n_img = a->n;
  P.out = a->y;
  P.out_fp32 = a->y_fp32;
  P.bias = a->bias;
  P.stats = a->stats;
  P.stats_sum_only = a->stats_sum_only;
  int N, KC, rc, nsteps = 0;
  // split-precision ("bf16x3") mode: every operand is a (hi, lo) bf16 pair, x = hi + lo to ~16 significant bits; the
  // K loop runs hi*hi + hi*lo + lo*hi into the same fp32 accumulator.  lo tensor maps live at index + 4, the lo half of
  // the packed weights at K offset k_hi.
  const int S3 = a->split3 ? 1 : 0;
  int src_slot = 0;  // which of a->x / a->x_lo the map being built belongs to
  auto mk_a = [&](int idx, const void* base, const uint64_t* dims, const uint64_t* str, const uint32_t* bx, int sw) -> int {
    int r = make_tensor_map_bf16(&maps.a[idx], base, 4, dims, str, bx, sw);
    if (r || !S3) return r;
    const char* lo = (const char*)a->x_lo[src_slot] + ((const char*)base - (const char*)a->x[src_slot]);
    return make_tensor_map_bf16(&maps.a[idx + 4], lo, 4, dims, str, bx, sw);
  };
  auto expand3 = [&](int n, int k_hi) -> int {
    if (!S3) return n;
    if (3 * n > MAX_STEPS) return -1;
    for (int i = n - 1; i >= 0; --i) {
      const KStep st0 = P.steps[i];
      KStep hl = st0, lh = st0;
      hl.wk = (int16_t)(st0.wk + k_hi);
      lh.map = (int8_t)(st0.map + 4);
      P.steps[3 * i] = st0; P.steps[3 * i + 1] = hl; P.steps[3 * i + 2] = lh;
    }
    return 3 * n;
  };

  if (a->mode == 0) {
    // ---------------------------------------------------------------- forward
    const int K_HI = taps * cin_total;
    N = a->cout;
    KC = 64;
    for (int i = 0; i < a->n_src; ++i) if (a->cin[i] % 64 != 0) KC = 32;
    for (int i = 0; i < a->n_src; ++i) if (a->cin[i] % KC != 0) { set_error("dfb_conv2d: input channels must be multiples of 32"); return DFB_ERR_UNSUPPORTED; }
    if (s == 2 && a->n_src != 1) { set_error("dfb_conv2d: stride-2 convolution takes a single source"); return DFB_ERR_UNSUPPORTED; }
    P.Ht = Ho; P.Wt = Wo;
    P.out_img = (long long)Ho * Wo * a->cout; P.out_row = (long long)Wo * a->cout; P.out_pix = a->cout;
    P.sy = P.sx = 1; P.oy = P.ox = 0;
    const uint32_t box[4] = {(uint32_t)KC, TILE_W, TILE_H, 1};
    if (s == 1 && k == 3 && KC == 64 && halo_mode() > 0) {
      // halo variant: one [18 x pitch] pixel box per 64-channel chunk, nine taps as shifted descriptors
      const int hm = halo_mode();
      (void)hm;
      P.halo_pitch = 10;
      P.dbg_base_offset = 0;
      P.halo_flip = 0;
      P.tap_k_stride = cin_total;
      const bool pair = N == 64 && !P.out_fp32 && !S3 && Ho % 2 == 0 && halo_pair_enabled();
      const uint32_t hbox[4] = {64, (uint32_t)P.halo_pitch, pair ? (uint32_t)HaloPairCfg::ROWS : 18u, 1};
      int coff = 0;
      for (int i = 0; i < a->n_src; ++i) {
        const uint64_t C = a->cin[i];
        const uint64_t dims[4] = {C, (uint64_t)W, (uint64_t)H, (uint64_t)a->n};
        const uint64_t str[3] = {C * 2, C * 2 * W, C * 2 * W * H};
        src_slot = i;
        if ((rc = mk_a(i, a->x[i], dims, str, hbox, 128))) return rc;
        for (int c0 = 0; c0 < a->cin[i]; c0 += 64) P.steps[nsteps++] = KStep{(int8_t)i, 0, 0, 0, (int16_t)c0, (int16_t)(coff + c0)};
        coff += a->cin[i];
      }
      const uint64_t bd[2] = {(uint64_t)taps * cin_total * (1 + S3), (uint64_t)a->cout};
      const uint64_t bs[1] = {(uint64_t)taps * cin_total * 2 * (1 + S3)};
      const uint32_t bb[2] = {64, (uint32_t)N};
      if ((rc = make_tensor_map_bf16(&maps.b, a->w, 2, bd, bs, bb, 128))) return rc;
      if ((nsteps = expand3(nsteps, K_HI)) < 0) { set_error("dfb_conv2d: too many K steps"); return DFB_ERR_UNSUPPORTED; }
      P.nsteps = nsteps;
      P.tiles_x = (P.Wt + TILE_W - 1) / TILE_W; P.tiles_y = (P.Ht + TILE_H - 1) / TILE_H;
      P.num_tiles = P.tiles_x * P.tiles_y * a->n;
      return pair ? launch_igemm_halo_pair(maps, P, st) : dispatch_igemm_halo(N, maps, P, st);
    }
    if (s == 1) {
      for (int i = 0; i < a->n_src; ++i) {
        const uint64_t C = a->cin[i];
        const uint64_t dims[4] = {C, (uint64_t)W, (uint64_t)H, (uint64_t)a->n};
        const uint64_t str[3] = {C * 2, C * 2 * W, C * 2 * W * H};
        src_slot = i;
        if ((rc = mk_a(i, a->x[i], dims, str, box, KC * 2))) return rc;
      }
      for (int t = 0; t < taps; ++t) {
        const int dy = k == 3 ? t / 3 - 1 : 0, dx = k == 3 ? t % 3 - 1 : 0;
        int coff = 0;
        for (int i = 0; i < a->n_src; ++i) {
          for (int c0 = 0; c0 < a->cin[i]; c0 += KC) {
            if (nsteps >= MAX_STEPS) { set_error("dfb_conv2d: too many K steps"); return DFB_ERR_UNSUPPORTED; }
            P.steps[nsteps++] = KStep{(int8_t)i, (int8_t)dy, (int8_t)dx, 0, (int16_t)c0, (int16_t)(t * cin_total + coff + c0)};
          }
          coff += a->cin[i];
        }
      }
    } else {
      // stride 2: the four parity planes of x are strided views; tap d in {0,1,2} along an axis reads plane
      // (d + 1) & 1 at shift (d == 0 ? -1 : 0)   [input index 2*o + d - 1]
      const uint64_t C = a->cin[0];
      if (H % 2 || W % 2) { set_error("dfb_conv2d: stride-2 convolution needs even H, W"); return DFB_ERR_UNSUPPORTED; }
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
          const uint64_t dims[4] = {C, (uint64_t)W / 2, (uint64_t)H / 2, (uint64_t)a->n};
          const uint64_t str[3] = {C * 2 * 2, C * 2 * W * 2, C * 2 * W * H};
          const char* base = (const char*)a->x[0] + ((size_t)py * W + px) * C * 2;
          if ((rc = mk_a(py * 2 + px, base, dims, str, box, KC * 2))) return rc;
        }
      for (int t = 0; t < taps; ++t) {
        const int ky = k == 3 ? t / 3 : 1, kx = k == 3 ? t % 3 : 1;
        const int py = (ky + 1) & 1, px = (kx + 1) & 1, dy = ky == 0 ? -1 : 0, dx = kx == 0 ? -1 : 0;
        for (int c0 = 0; c0 < a->cin[0]; c0 += KC)
          P.steps[nsteps++] = KStep{(int8_t)(py * 2 + px), (int8_t)dy, (int8_t)dx, 0, (int16_t)c0, (int16_t)(t * cin_total + c0)};
      }
    }
    const uint64_t bd[2] = {(uint64_t)taps * cin_total * (1 + S3), (uint64_t)a->cout};
    const uint64_t bs[1] = {(uint64_t)taps * cin_total * 2 * (1 + S3)};
    const uint32_t bb[2] = {(uint32_t)KC, (uint32_t)N};
    if ((rc = make_tensor_map_bf16(&maps.b, a->w, 2, bd, bs, bb, KC * 2))) return rc;
    if ((nsteps = expand3(nsteps, K_HI)) < 0) { set_error("dfb_conv2d: too many K steps"); return DFB_ERR_UNSUPPORTED; }
    P.nsteps = nsteps;
    P.tiles_x = (P.Wt + TILE_W - 1) / TILE_W; P.tiles_y = (P.Ht + TILE_H - 1) / TILE_H;
    P.num_tiles = P.tiles_x * P.tiles_y * a->n;
    return dispatch_igemm(N, KC, maps, P, st);
  }

  if (a->mode == 1) {
    // ---------------------------------------------------------------- data gradient
    // source = gy [n, Ho, Wo, cout]; output gx [n, H, W, cin_slice]; weights Wd[ci][tap*cout + co] (rows = all cin_total
    // channels; the slice starts at row cin_off).
    const int cout = a->cout;
    const bool split_out = a->y2 != nullptr;
    if (split_out && (k != 1 || s != 1 || a->y_fp32 || S3 || a->cin[0] % 32 || a->cin2 % 32 || a->cin2 <= 0)) {
      set_error("dfb_conv2d dgrad: a second output needs a 1x1 stride-1 bf16 launch with 32-channel multiples"); return DFB_ERR_UNSUPPORTED;
    }
    const int cin = a->cin[0] + (split_out ? a->cin2 : 0);   // accumulator columns
    const int K_HI = taps * cout;
    N = cin;
    KC = cout % 64 == 0 ? 64 : 32;
    if (cout % KC != 0) { set_error("dfb_conv2d dgrad: output channels must be multiples of 32"); return DFB_ERR_UNSUPPORTED; }
    const uint64_t C = cout;
    const uint64_t dims[4] = {C, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)a->n};
    const uint64_t str[3] = {C * 2, C * 2 * Wo, C * 2 * Wo * Ho};
    const uint32_t box[4] = {(uint32_t)KC, TILE_W, TILE_H, 1};
    if ((rc = mk_a(0, a->x[0], dims, str, box, KC * 2))) return rc;
    const uint64_t bd[2] = {(uint64_t)taps * cout * (1 + S3), (uint64_t)cin};
    const uint64_t bs[1] = {(uint64_t)taps * cout * 2 * (1 + S3)};
    const uint32_t bb[2] = {(uint32_t)KC, (uint32_t)N};
    const char* wbase = (const char*)a->w + (size_t)a->cin_off * taps * cout * 2 * (1 + S3);
    if ((rc = make_tensor_map_bf16(&maps.b, wbase, 2, bd, bs, bb, KC * 2))) return rc;
    P.out_img = (long long)H * W * a->cin[0]; P.out_row = (long long)W * a->cin[0]; P.out_pix = a->cin[0];
    if (split_out) { P.split_col = a->cin[0]; P.out2 = a->y2; }
    if (s == 1 && k == 3 && KC == 64 && halo_mode() > 0) {
      const int hm = halo_mode();
      P.Ht = H; P.Wt = W; P.sy = P.sx = 1; P.oy = P.ox = 0;
      (void)hm;
      P.halo_pitch = 10;
      P.dbg_base_offset = 0;
      P.halo_flip = 1;                       // gx[y, x] = sum_t W[t]^T gy[y - dy_t, x - dx_t]
      P.tap_k_stride = cout;
      const bool pair = N == 64 && !P.out_fp32 && !S3 && H % 2 == 0 && halo_pair_enabled();
      const uint32_t hbox[4] = {64, (uint32_t)P.halo_pitch, pair ? (uint32_t)HaloPairCfg::ROWS : 18u, 1};
      if ((rc = mk_a(0, a->x[0], dims, str, hbox, 128))) return rc;
      for (int c0 = 0; c0 < cout; c0 += 64) P.steps[nsteps++] = KStep{0, 0, 0, 0, (int16_t)c0, (int16_t)c0};
      if ((nsteps = expand3(nsteps, K_HI)) < 0) { set_error("dfb_conv2d: too many K steps"); return DFB_ERR_UNSUPPORTED; }
      P.nsteps = nsteps;
      P.tiles_x = (P.Wt + TILE_W - 1) / TILE_W; P.tiles_y = (P.Ht + TILE_H - 1) / TILE_H;
      P.num_tiles = P.tiles_x * P.tiles_y * a->n;
      return pair ? launch_igemm_halo_pair(maps, P, st) : dispatch_igemm_halo(N, maps, P, st);
    }
    if (s == 1) {
      // gx[y, x] = sum_t W[t]^T gy[y - dy_t, x - dx_t]
      P.Ht = H; P.Wt = W; P.sy = P.sx = 1; P.oy = P.ox = 0;
      for (int t = 0; t < taps; ++t) {
        const int dy = k == 3 ? t / 3 - 1 : 0, dx = k == 3 ? t % 3 - 1 : 0;
        for (int c0 = 0; c0 < cout; c0 += KC)
          P.steps[nsteps++] = KStep{0, (int8_t)(-dy), (int8_t)(-dx), 0, (int16_t)c0, (int16_t)(t * cout + c0)};
      }
      if ((nsteps = expand3(nsteps, K_HI)) < 0) { set_error("dfb_conv2d: too many K steps"); return DFB_ERR_UNSUPPORTED; }
      P.nsteps = nsteps;
      P.tiles_x = (P.Wt + TILE_W - 1) / TILE_W; P.tiles_y = (P.Ht + TILE_H - 1) / TILE_H;
      P.num_tiles = P.tiles_x * P.tiles_y * a->n;
      return dispatch_igemm(N, KC, maps, P, st);
    }
    // stride 2: input row i = 2a + py receives from output rows o with 2o + ky - 1 = i:
    //   py = 0: ky = 1, o = a;      py = 1: ky = 0, o = a + 1  and  ky = 2, o = a.   One launch per parity plane.
    if (H % 2 || W % 2) { set_error("dfb_conv2d dgrad: stride-2 needs even H, W"); return DFB_ERR_UNSUPPORTED; }
    if (k == 3 && !S3 && !a->y_fp32 && par_merge_enabled()) {
      // all four parity planes in ONE launch (par_decode): gy is read from DRAM once instead of four times and the
      // interleaved planes of gx are written while their lines are still in L2
      nsteps = 0;
      P.Ht = H / 2; P.Wt = W / 2; P.sy = P.sx = 2;
      P.wide = igemm_wide(P, N) ? 1 : 0;
      CUtensorMap* outs[4] = {&maps.out, &maps.out2, &maps.out3, &maps.out4};
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
          const int p = py * 2 + px;
          P.seg[p] = (int16_t)nsteps;
          for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx) {
              if (((ky + 1) & 1) != py || ((kx + 1) & 1) != px) continue;
              const int dy = ky == 0 ? 1 : 0, dx = kx == 0 ? 1 : 0;
              const int t = ky * 3 + kx;
              for (int c0 = 0; c0 < cout; c0 += KC) {
                if (nsteps >= MAX_STEPS) { set_error("dfb_conv2d: too many K steps"); return DFB_ERR_UNSUPPORTED; }
                P.steps[nsteps++] = KStep{0, (int8_t)dy, (int8_t)dx, 0, (int16_t)c0, (int16_t)(t * cout + c0)};
              }
            }
          P.oy = py; P.ox = px;
          IgemmMaps tmp;
          if ((rc = build_out_map(tmp, P, N))) return rc;
          *outs[p] = tmp.out;
        }
      P.seg[4] = (int16_t)nsteps;
      P.oy = P.ox = 0;
      P.nsteps = nsteps;
      P.par_mode = 1;
      P.tiles_x = (P.Wt + TILE_W - 1) / TILE_W; P.tiles_y = (P.Ht + TILE_H - 1) / TILE_H;
      P.num_tiles = 4 * P.tiles_x * P.tiles_y * a->n;
      return dispatch_igemm(N, KC, maps, P, st);
    }
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        nsteps = 0;
        for (int ky = 0; ky < k; ++ky)
          for (int kx = 0; kx < k; ++kx) {
            const int kyy = k == 3 ? ky : 1, kxx = k == 3 ? kx : 1;
            if (((kyy + 1) & 1) != py || ((kxx + 1) & 1) != px) continue;
            const int dy = kyy == 0 ? 1 : 0, dx = kxx == 0 ? 1 : 0;
            const int t = ky * k + kx;
            for (int c0 = 0; c0 < cout; c0 += KC)
              P.steps[nsteps++] = KStep{0, (int8_t)dy, (int8_t)dx, 0, (int16_t)c0, (int16_t)(t * cout + c0)};
          }
        P.Ht = H / 2; P.Wt = W / 2; P.sy = P.sx = 2; P.oy = py; P.ox = px;
        P.tiles_x = (P.Wt + TILE_W - 1) / TILE_W; P.tiles_y = (P.Ht + TILE_H - 1) / TILE_H;
        P.num_tiles = P.tiles_x * P.tiles_y * a->n;
        if (nsteps > 0 && (nsteps = expand3(nsteps, K_HI)) < 0) { set_error("dfb_conv2d: too many K steps"); return DFB_ERR_UNSUPPORTED; }
        P.nsteps = nsteps;
        if (nsteps == 0) {  // 1x1 stride-2: planes that no tap reaches get zero gradient
          set_error("dfb_conv2d dgrad: 1x1 stride-2 is not on the DeFlow path"); return DFB_ERR_UNSUPPORTED;
        }
        if ((rc = dispatch_igemm(N, KC, maps, P, st))) return rc;
      }
    return DFB_OK;
  }
  set_error("dfb_conv2d: unknown mode %d", a->mode);
  return DFB_ERR_ARG;
}

// ================================================================================================
// Weight gradient:  gW[t][co][ci] = sum over pixels of gy[n, y, x, co] * x[n, y*s + dy_t, x*s + dx_t, ci]
//
// A GEMM whose K dimension is the pixel index, so both operands are MN-major in shared memory: a TMA box of
// [8 x 8 pixels] x 64 channels lands as 64 rows (K) of 128 bytes (64 channels, M or N).  One CTA owns a pair
// of (tap, 64-channel chunk) groups of x  -> M = 128 accumulator rows -- times all Cout columns, over a range
// of pixel tiles; at the end it adds its partial result into an fp32 workspace [tap][co][ci] with coalesced
// red.global.add.  Split-K over pixel ranges gives every SM work.  k_unpack_wgrad writes the torch layout.
namespace dfb {
namespace tc {

constexpr int WG_PIX = 64;          // pixels per stage (8 x 8 patch)
constexpr int WG_THREADS = 192;
constexpr int MAX_GROUPS = 80;

struct WGroup { int8_t map, dy, dx, tap; int16_t c0; int16_t ci_glob; };  // ci_glob: channel index in the layer's cin_total

struct WgradMaps {
  CUtensorMap a[MAX_AMAPS];  // x sources / parity planes
  CUtensorMap b;             // gy
};

struct WgradParams {
  int n_img, Ht, Wt, tiles_x, tiles_y, num_tiles;  // gy pixel grid in 8x8 tiles
  int n_pairs, splits;                              // grid = n_pairs * splits
  int n_groups, cin_total, cout;
  float* wacc;                                      // [taps][cout][cin_total] fp32, += by red.add
  float* grad_bias;                                 // [cout] fp32 += sum over pixels of gy, or NULL (odd group counts only)
  WGroup groups[MAX_GROUPS];
};

template <int N>
struct WgradCfg {
  static constexpr int A_BYTES = 2 * WG_PIX * 128;        // two x groups
  static constexpr int B_BYTES = (N / 64) * WG_PIX * 128;  // gy, N/64 channel chunks
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = (200 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = N < 32 ? 32 : N;
};

template <int N>
__global__ void __launch_bounds__(WG_THREADS, 1) k_conv_wgrad(const __grid_constant__ WgradMaps maps,
                                                              const __grid_constant__ WgradParams P) {
  using Cfg = WgradCfg<N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* acc_full = bars + 2 * Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int pair = blockIdx.x % P.n_pairs, split = blockIdx.x / P.n_pairs;
  const int g0 = pair * 2, g1 = g0 + 1 < P.n_groups ? g0 + 1 : -1;
  const int per = (P.num_tiles + P.splits - 1) / P.splits;
  const int t_begin = split * per, t_end = min(P.num_tiles, t_begin + per);
  // Bias gradient for free: an odd group count leaves the second 64 rows of the last pair's M = 128 operand unused.
  // Filled with ones (once, in every stage; the producer then loads only the first half), those accumulator rows
  // become sum over pixels of 1 * gy[pixel, co] -- the per-channel sums of gy, i.e. the bias gradient of the layer.
  const bool ones_half = P.grad_bias != nullptr && g1 < 0;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < MAX_AMAPS; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  if (ones_half) {
    const uint4 one8 = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);   // eight bf16 1.0
    for (int i = threadIdx.x; i < Cfg::STAGES * (WG_PIX * 128 / 16); i += blockDim.x) {
      const int stage = i / (WG_PIX * 128 / 16), o = i % (WG_PIX * 128 / 16);
      *reinterpret_cast<uint4*>(smem + stage * Cfg::STAGE_BYTES + WG_PIX * 128 + o * 16) = one8;
    }
    fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const WGroup ga = P.groups[g0];
      const WGroup gb = P.groups[g1 >= 0 ? g1 : g0];
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        const int tx = tile % P.tiles_x, t2 = tile / P.tiles_x, ty = t2 % P.tiles_y, n = t2 / P.tiles_y;
        const int x0 = tx * 8, y0 = ty * 8;
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* a_dst = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* b_dst = a_dst + Cfg::A_BYTES;
        mbar_arrive_expect_tx(&full[stage], ones_half ? Cfg::STAGE_BYTES - WG_PIX * 128 : Cfg::STAGE_BYTES);
        tma_load_4d(a_dst, &maps.a[ga.map], &full[stage], ga.c0, x0 + ga.dx, y0 + ga.dy, n);
        // an odd group count leaves the second half of the last pair unused: load the same box, discard the rows
        // (or keep the ones written at kernel start: the bias-gradient rows)
        if (!ones_half) tma_load_4d(a_dst + WG_PIX * 128, &maps.a[gb.map], &full[stage], gb.c0, x0 + gb.dx, y0 + gb.dy, n);
#pragma unroll
        for (int j = 0; j < N / 64; ++j) tma_load_4d(b_dst + j * WG_PIX * 128, &maps.b, &full[stage], j * 64, x0, y0, n);
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // whole-warp issue loop, (lo, hi) descriptor halves -- see k_conv_wgrad_halo
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 1, 1);  // both operands MN-major
    // MN-major SWIZZLE_128B: 64-element MN groups LBO apart, 8-row K groups SBO = 1024 bytes apart
    const uint32_t d_hi = (uint32_t)(make_smem_desc(0, 0, 1024, 2) >> 32);
    const uint32_t lbo = (uint32_t)((WG_PIX * 128) >> 4) << 16;
    const uint32_t a_lo0 = ((smem_u32(smem) & 0x3FFFFu) >> 4) + lbo;
    const uint32_t b_lo0 = a_lo0 + (uint32_t)(Cfg::A_BYTES >> 4);
    int stage = 0;
    uint32_t phase = 0, accum = 0u;
    for (int tile = t_begin; tile < t_end; ++tile) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t soff = (uint32_t)stage * (uint32_t)(Cfg::STAGE_BYTES >> 4);
#pragma unroll
      for (int k = 0; k < WG_PIX / 16; ++k)   // 16 pixels = 2 K groups = 2048 bytes per UMMA
        umma_bf16_lohi_warp(tmem_base, a_lo0 + soff + (uint32_t)(k * 128), d_hi, b_lo0 + soff + (uint32_t)(k * 128), d_hi, idesc,
                            accum | (uint32_t)(k > 0));
      accum = 1u;
      umma_commit_warp(&empty[stage]);
      if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
    }
    umma_commit_warp(acc_full);
  } else if (t_end > t_begin) {
    const int q = warp & 3;
    const int m = q * 32 + lane;            // accumulator row: group (m >> 6), channel (m & 63) of that group
    const int g = m < 64 ? g0 : g1;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const WGroup gr = P.groups[g >= 0 ? g : g0];
    float* dst = P.wacc + ((size_t)gr.tap * P.cout) * P.cin_total + gr.ci_glob + (m & 63);
    const bool live = g >= 0 && (gr.ci_glob + (m & 63)) < P.cin_total;
#pragma unroll 1
    for (int col = 0; col < N && col < P.cout; col += 32) {  // cout = 32 rides in a zero-filled 64-wide tile
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + col, v);
      if (live) {
#pragma unroll
        for (int i = 0; i < 32; ++i) atomicAdd(dst + (size_t)(col + i) * P.cin_total, v[i]);  // lanes = consecutive ci
      }
      if (ones_half && m == 64) {   // every row 64..127 holds the same per-channel sums of gy: one of them reports
#pragma unroll
        for (int i = 0; i < 32; ++i) atomicAdd(P.grad_bias + col + i, v[i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// Halo weight gradient for 3x3 stride-1 convolutions.  One [10 x 10 pixel] x 64-channel box of x serves all taps of
// a work item: tap (dy, dx) is the MN-major descriptor that starts (dy*10 + dx) pixel rows into the box, and a PAIR of
// taps forms one M = 128 operand whose two 64-channel groups are (offset(t2) - offset(t1)) bytes apart (the leading
// byte offset).  A work item = (64-channel chunk of x, set of <= 512/N tap pairs, range of pixel tiles); gy tiles are
// loaded once per item and tile instead of once per tap pair, x once instead of once per tap.
constexpr int WH_HALO_BYTES = 13 * 1024;  // 100 pixel rows x 128 B = 12800, 1024-aligned
constexpr int WH_MAX_ITEMS = 32;

struct WHItem { int8_t map, npairs, dup_first, pad1; int16_t c0, ci_glob; int8_t t1[5], t2[5]; int16_t pad2; int16_t blk0, nsplit; };  // dup_first: the last pair's first tap repeats one already covered (odd tap count); blk0 / nsplit: the CTAs of this item

struct WgradHaloParams {
  int n_img, Ht, Wt, tiles_x, tiles_y, num_tiles;
  int n_items, splits;
  int cin_total, cout;
  float* wacc;
  WHItem items[WH_MAX_ITEMS];
};

template <int N>
struct WgradHaloCfg {
  static constexpr int B_BYTES = (N / 64) * WG_PIX * 128;
  static constexpr int STAGE_BYTES = WH_HALO_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = (208 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int MAX_PAIRS = 512 / N > 5 ? 5 : 512 / N;
  static constexpr int TMEM_COLS = 512;
};

template <int N>
__global__ void __launch_bounds__(WG_THREADS, 1) k_conv_wgrad_halo(const __grid_constant__ WgradMaps maps,
                                                                   const __grid_constant__ WgradHaloParams P) {
  using Cfg = WgradHaloCfg<N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* acc_full = bars + 2 * Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // CTAs are dealt to the work items in proportion to their MMA count (tap pairs): an item with one pair loads the same
  // x box and gy tile per pixel tile as an item with four and would otherwise only wait on L2
  int item_id = 0;
  for (int i = 1; i < P.n_items; ++i) if ((int)blockIdx.x >= P.items[i].blk0) item_id = i;
  const WHItem it = P.items[item_id];
  const int split = (int)blockIdx.x - it.blk0;
  const int per = (P.num_tiles + it.nsplit - 1) / it.nsplit;
  const int t_begin = split * per, t_end = min(P.num_tiles, t_begin + per);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < MAX_AMAPS; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        const int tx = tile % P.tiles_x, t2 = tile / P.tiles_x, ty = t2 % P.tiles_y, n = t2 / P.tiles_y;
        const int x0 = tx * 8, y0 = ty * 8;
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* a_dst = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* b_dst = a_dst + WH_HALO_BYTES;
        mbar_arrive_expect_tx(&full[stage], 100 * 128 + Cfg::B_BYTES);
        tma_load_4d(a_dst, &maps.a[it.map], &full[stage], it.c0, x0 - 1, y0 - 1, n);
#pragma unroll
        for (int j = 0; j < N / 64; ++j) tma_load_4d(b_dst + j * WG_PIX * 128, &maps.b, &full[stage], j * 64, x0, y0, n);
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // Whole-warp issue loop (warp-uniform control flow, one elected lane issues) with (lo, hi) descriptor halves: only
    // the start-address field of the low word changes from MMA to MMA.  Under `if (lane == 0)` the loop cost ~30 SASS
    // instructions per MMA on one thread (ncu r01: tensor pipe active 14 %) -- more than the 64-128 cycles an MMA takes.
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 1, 1);
    const uint32_t a_hi = (uint32_t)(make_smem_desc(0, 0, 1280, 2) >> 32), b_hi = (uint32_t)(make_smem_desc(0, 0, 1024, 2) >> 32);
    const uint32_t smem_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
    // per tap pair: start (o1 pixel rows into the box) | leading byte offset ((o2 - o1) pixel rows), in 16-byte units
    uint32_t pair_lo[Cfg::MAX_PAIRS];
#pragma unroll
    for (int pr = 0; pr < Cfg::MAX_PAIRS; ++pr) {
      const int t1 = pr < it.npairs ? it.t1[pr] : 0, t2 = pr < it.npairs ? it.t2[pr] : 1;
      const int o1 = (t1 / 3) * 10 + t1 % 3, o2 = (t2 / 3) * 10 + t2 % 3;
      pair_lo[pr] = smem_lo + (uint32_t)o1 * 8u + (((uint32_t)(o2 - o1) * 8u) << 16);
    }
    const uint32_t b_lo0 = smem_lo + (uint32_t)(WH_HALO_BYTES >> 4) + ((uint32_t)((WG_PIX * 128) >> 4) << 16);
    const int npairs = it.npairs;
    int stage = 0;
    uint32_t phase = 0, accum = 0u;
    for (int tile = t_begin; tile < t_end; ++tile) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t soff = (uint32_t)stage * (uint32_t)(Cfg::STAGE_BYTES >> 4);
#pragma unroll
      for (int pr = 0; pr < Cfg::MAX_PAIRS; ++pr) {
        if (pr < npairs) {
          // two 64-channel M groups = two taps; 16 pixels = 2 patch rows: A advances 2 halo rows (160 units), B 128 units
#pragma unroll
          for (int k = 0; k < WG_PIX / 16; ++k)
            umma_bf16_lohi_warp(tmem_base + pr * N, pair_lo[pr] + soff + (uint32_t)(k * 160), a_hi,
                                b_lo0 + soff + (uint32_t)(k * 128), b_hi, idesc, accum | (uint32_t)(k > 0));
        }
      }
      accum = 1u;
      umma_commit_warp(&empty[stage]);
      if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
    }
    umma_commit_warp(acc_full);
  } else if (t_end > t_begin) {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const bool live_c = (it.ci_glob + (m & 63)) < P.cin_total;
    for (int pr = 0; pr < it.npairs; ++pr) {
      const int tap = m < 64 ? it.t1[pr] : it.t2[pr];
      const bool live = live_c && !(m < 64 && it.dup_first && pr == it.npairs - 1);  // drop the repeated tap
      float* dst = P.wacc + ((size_t)tap * P.cout) * P.cin_total + it.ci_glob + (m & 63);
#pragma unroll 1
      for (int col = 0; col < N && col < P.cout; col += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + pr * N + col, v);
        if (live) {
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(dst + (size_t)(col + i) * P.cin_total, v[i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int N>
static int launch_wgrad_halo(const WgradMaps& maps, const WgradHaloParams& P, cudaStream_t st) {
  using Cfg = WgradHaloCfg<N>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_wgrad_halo<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) { set_error("wgrad_halo<%d>: cannot reserve %d bytes of shared memory: %s", N, Cfg::SMEM, cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    configured = true;
  }
  k_conv_wgrad_halo<N><<<P.splits, WG_THREADS, Cfg::SMEM, st>>>(maps, P);   // P.splits = total CTAs
  add_launches(1);
  return check_launch("conv_wgrad_halo");
}

// ------------------------------------------------------------------------------------------------
// Cross-shift weight gradient for 3x3 stride-1 convolutions with 64 output channels (the default for those layers since
// r02: validated on B200 against cuDNN fp32, 0.195 -> 0.162 ms per launch = 970 -> 1164 TFLOP/s; DFB_WGRAD_X=0 selects
// k_conv_wgrad_halo<64> for A/B runs).
//   gW[(ky,kx)][co][ci] = sum_p gy[p, co] x[p + (ky-1, kx-1), ci] = sum_p' gy[p' - (ky-1, 0), co] x[p' + (0, kx-1), ci]
// so the three kx taps are HORIZONTAL shifts of an x box ([8 rows x 10 cols] x 64 channels, taps stacked along M like
// in k_conv_wgrad_halo) and the three ky taps are VERTICAL shifts of a gy box ([10 rows x 8 cols] x 64 channels)
// stacked along N: N group g starts g box rows (1024 B) in = tap ky = 2 - g.  Per 16-pixel K slice: two M = 128 x
// N = 192 MMAs (kx pairs (0,1) and (1,2), the repeated kx = 1 rows dropped) instead of five M = 128 x N = 64 ones,
// which are bound by the A-operand fetch (70 cycles for 32 cycles of math) -- ~1.75x by the measured MMA times.
// Out-of-range box rows / columns are zero-filled by TMA: exactly the padding (x) and the missing output pixels (gy).
constexpr int WX_BOX_BYTES = 80 * 128;      // 80 pixel rows of 128 B (10 KB, 1024-aligned) for either box
struct WgradXCfg {
  static constexpr int STAGE_BYTES = 2 * WX_BOX_BYTES;
  static constexpr int STAGES = 8;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 512;     // two accumulators of 192 columns
};

struct WgradXParams {
  int n_img, Ht, Wt, tiles_x, tiles_y, num_tiles;
  int n_items, splits;          // grid = n_items * splits
  int cin_total, cout;
  float* wacc;
  int8_t map[8];                // per item: x tensor map
  int16_t c0[8], ci_glob[8];    // per item: channel offset in the source / in the layer's cin_total
};

__global__ void __launch_bounds__(WG_THREADS, 1) k_conv_wgrad_x(const __grid_constant__ WgradMaps maps,
                                                                const __grid_constant__ WgradXParams P) {
  using Cfg = WgradXCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* acc_full = bars + 2 * Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x % P.n_items, split = blockIdx.x / P.n_items;
  const int per = (P.num_tiles + P.splits - 1) / P.splits;
  const int t_begin = split * per, t_end = min(P.num_tiles, t_begin + per);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < MAX_AMAPS; ++i) tma_prefetch_desc(&maps.a[i]);
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        const int tx = tile % P.tiles_x, t2 = tile / P.tiles_x, ty = t2 % P.tiles_y, n = t2 / P.tiles_y;
        const int x0 = tx * 8, y0 = ty * 8;
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* a_dst = smem + stage * Cfg::STAGE_BYTES;
        mbar_arrive_expect_tx(&full[stage], 2 * WX_BOX_BYTES);
        tma_load_4d(a_dst, &maps.a[P.map[item]], &full[stage], P.c0[item], x0 - 1, y0, n);      // x: [8 rows][10 cols]
        tma_load_4d(a_dst + WX_BOX_BYTES, &maps.a[MAX_AMAPS - 1], &full[stage], 0, x0, y0 - 1, n); // gy: [10 rows][8 cols]
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // whole-warp issue loop, (lo, hi) descriptor halves (see k_conv_wgrad_halo)
    constexpr uint32_t idesc = make_idesc_bf16(128, 192, 1, 1);
    const uint32_t a_hi = (uint32_t)(make_smem_desc(0, 0, 1280, 2) >> 32);   // x: 8-pixel K groups one box row (10 px) apart
    const uint32_t b_hi = (uint32_t)(make_smem_desc(0, 0, 1024, 2) >> 32);   // gy: 8-pixel K groups one box row (8 px) apart
    const uint32_t smem_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
    const uint32_t a_lo0 = smem_lo + ((128u >> 4) << 16);                                       // M groups: kx, kx + 1 (one pixel apart)
    const uint32_t b_lo0 = smem_lo + (uint32_t)(WX_BOX_BYTES >> 4) + ((1024u >> 4) << 16);      // N groups: ky = 2, 1, 0 (one row apart)
    int stage = 0;
    uint32_t phase = 0, accum = 0u;
    for (int tile = t_begin; tile < t_end; ++tile) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t soff = (uint32_t)stage * (uint32_t)(Cfg::STAGE_BYTES >> 4);
#pragma unroll
      for (int pr = 0; pr < 2; ++pr) {       // kx pairs (0, 1) and (1, 2)
#pragma unroll
        for (int k = 0; k < 4; ++k)          // 16 pixels = 2 box rows: x advances 2 x 1280 B, gy 2 x 1024 B
          umma_bf16_lohi_warp(tmem_base + pr * 192, a_lo0 + soff + (uint32_t)pr * 8u + (uint32_t)(k * 160), a_hi,
                              b_lo0 + soff + (uint32_t)(k * 128), b_hi, idesc, accum | (uint32_t)(k > 0));
      }
      accum = 1u;
      umma_commit_warp(&empty[stage]);
      if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
    }
    umma_commit_warp(acc_full);
  } else if (t_end > t_begin) {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int ci = P.ci_glob[item] + (m & 63);
    for (int pr = 0; pr < 2; ++pr) {
      const int kx = pr + (m >> 6);
      const bool live = ci < P.cin_total && !(pr == 1 && m < 64);   // kx = 1 was already covered by pair 0
#pragma unroll 1
      for (int col = 0; col < 192; col += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + pr * 192 + col, v);
        const int ky = 2 - col / 64, co0 = col % 64;
        float* dst = P.wacc + ((size_t)(ky * 3 + kx) * P.cout + co0) * P.cin_total + ci;
        if (live) {
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(dst + (size_t)i * P.cin_total, v[i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// wacc [taps][cout][cin] -> torch [cout][cin][taps]; accumulate = grad += (autograd accumulation)
__global__ void __launch_bounds__(256) k_unpack_wgrad(const float* __restrict__ wacc, int cout, int cin, int taps,
                                                      float* __restrict__ gw, int accumulate) {
  const long long total = (long long)cout * cin * taps;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(e % taps);
    const int ci = (int)((e / taps) % cin);
    const int co = (int)(e / ((long long)taps * cin));
    const float v = wacc[((size_t)t * cout + co) * cin + ci];
    gw[e] = accumulate ? gw[e] + v : v;
  }
}

template <int N>
static int launch_wgrad(const WgradMaps& maps, const WgradParams& P, cudaStream_t st) {
  using Cfg = WgradCfg<N>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_wgrad<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) { set_error("wgrad<%d>: cannot reserve %d bytes of shared memory: %s", N, Cfg::SMEM, cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    configured = true;
  }
  k_conv_wgrad<N><<<P.n_pairs * P.splits, WG_THREADS, Cfg::SMEM, st>>>(maps, P);
  add_launches(1);
  return check_launch("conv_wgrad");
}

}  // namespace tc
}  // namespace dfb

// CTAs per weight-gradient launch = wgrad_cta_target(): ONE per SM.  These kernels hold one CTA per SM (shared memory), so a
// larger split-K only serialises in waves and multiplies the red.global.add traffic: measured on B200 (r02 visit 12, whole
// step) 30.22 ms at 1 CTA / SM against 31.16 / 30.72 / 30.94 / 31.36 ms at 1.5 / 2 / 3 / 4.  DFB_WGRAD_CTAS_X10 = tenths of
// CTAs per SM for sweeps.
static int wgrad_cta_target() {
  static int x10 = -1;
  if (x10 < 0) {
    const char* e = getenv("DFB_WGRAD_CTAS_X10");
    x10 = e ? atoi(e) : 10;
    if (x10 < 5 || x10 > 80) x10 = 10;
  }
  return (x10 * dfb::sm_count() + 5) / 10;
}

static bool halo_wgrad_enabled() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("DFB_WGRAD_HALO");
    mode = e ? atoi(e) : 1;
  }
  return mode != 0;
}

// x: the forward's sources (n_src, cin[]), a->w unused; a->y = gy [n,Ho,Wo,cout] bf16 (input here);
// wacc: fp32 workspace [taps][cout][cin_total] (zeroed here); grad_w: torch layout fp32 [cout][cin_total][k][k].
extern "C" int dfb_conv2d_wgrad(const dfb_conv_args* a, float* wacc, float* grad_w, int accumulate, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (!a || !wacc) { set_error("dfb_conv2d_wgrad: null args"); return DFB_ERR_ARG; }
  // accumulate bit 1: wacc already holds partial sums (a shared weight's other call, or the caller zeroed it): no memset;
  // grad_w == NULL: leave the result in wacc ([tap][cout][cin]) for dfb_wgrad_unpack_multi
  const bool keep_wacc = (accumulate & 2) != 0;
  accumulate &= 1;
  const int k = a->ksize, s = a->stride, taps = k * k;
  if ((k != 1 && k != 3) || (s != 1 && s != 2) || a->n_src < 1 || a->n_src > 2) { set_error("dfb_conv2d_wgrad: bad conv geometry"); return DFB_ERR_ARG; }
  if (a->cout != 32 && a->cout != 64 && a->cout != 128 && a->cout != 256) { set_error("dfb_conv2d_wgrad: cout must be 32, 64, 128 or 256"); return DFB_ERR_UNSUPPORTED; }
  const int H = a->H, W = a->W;
  const int Ho = (H + 2 * (k / 2) - k) / s + 1, Wo = (W + 2 * (k / 2) - k) / s + 1;
  int cin_total = 0;
  for (int i = 0; i < a->n_src; ++i) cin_total += a->cin[i];
  WgradMaps maps;
  WgradParams P;
  memset(&maps, 0, sizeof(maps));
  memset(&P, 0, sizeof(P));
  int rc;
  const uint32_t box[4] = {64, 8, 8, 1};
  {
    const uint64_t C = a->cout;
    const uint64_t dims[4] = {C, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)a->n};
    const uint64_t str[3] = {C * 2, C * 2 * Wo, C * 2 * Wo * Ho};
    if ((rc = make_tensor_map_bf16(&maps.b, a->y, 4, dims, str, box, 128))) return rc;
  }
  int ng = 0;
  if (a->grad_bias && (k != 1 || s != 1)) { set_error("dfb_conv2d_wgrad: grad_bias is served for 1x1 stride-1 layers only"); return DFB_ERR_UNSUPPORTED; }
  if (s == 1 && k == 3 && halo_wgrad_enabled()) {
    // ---------------------------------------------------------------- halo variant
    {
      const char* ex = getenv("DFB_WGRAD_X");   // cross-shift kernel (k_conv_wgrad_x), 64 output channels only; "0" disables
      int n_chunks = 0;
      for (int i = 0; i < a->n_src; ++i) n_chunks += (a->cin[i] + 63) / 64;
      if (!(ex && atoi(ex) == 0) && a->cout == 64 && n_chunks <= 8) {
        tc::WgradXParams X;
        memset(&X, 0, sizeof(X));
        const uint32_t xbox[4] = {64, 10, 8, 1}, gbox[4] = {64, 8, 10, 1};
        int ni = 0, coff = 0;
        for (int i = 0; i < a->n_src; ++i) {
          const uint64_t C = a->cin[i];
          const uint64_t dims[4] = {C, (uint64_t)W, (uint64_t)H, (uint64_t)a->n};
          const uint64_t str[3] = {C * 2, C * 2 * W, C * 2 * W * H};
          if ((rc = make_tensor_map_bf16(&maps.a[i], a->x[i], 4, dims, str, xbox, 128))) return rc;
          for (int c0 = 0; c0 < a->cin[i]; c0 += 64) { X.map[ni] = (int8_t)i; X.c0[ni] = (int16_t)c0; X.ci_glob[ni] = (int16_t)(coff + c0); ++ni; }
          coff += a->cin[i];
        }
        {
          const uint64_t C = a->cout;
          const uint64_t dims[4] = {C, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)a->n};
          const uint64_t str[3] = {C * 2, C * 2 * Wo, C * 2 * Wo * Ho};
          if ((rc = make_tensor_map_bf16(&maps.a[tc::MAX_AMAPS - 1], a->y, 4, dims, str, gbox, 128))) return rc;
        }
        X.n_items = ni;
        X.n_img = a->n; X.Ht = Ho; X.Wt = Wo;
        X.tiles_x = (Wo + 7) / 8; X.tiles_y = (Ho + 7) / 8;
        X.num_tiles = X.tiles_x * X.tiles_y * a->n;
        int splits = wgrad_cta_target() / ni;        // floor: never more CTAs than the target (a single wave at the default)
        if (splits > X.num_tiles) splits = X.num_tiles;
        if (splits < 1) splits = 1;
        X.splits = splits;
        X.cin_total = cin_total; X.cout = a->cout; X.wacc = wacc;
        const size_t total = (size_t)taps * a->cout * cin_total;
        if (!keep_wacc) cudaMemsetAsync(wacc, 0, total * sizeof(float), st);
        static bool configured = false;
        if (!configured) {
          cudaError_t e = cudaFuncSetAttribute(tc::k_conv_wgrad_x, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::WgradXCfg::SMEM);
          if (e != cudaSuccess) { set_error("wgrad_x: cannot reserve shared memory: %s", cudaGetErrorString(e)); return DFB_ERR_CUDA; }
          configured = true;
        }
        tc::k_conv_wgrad_x<<<ni * splits, tc::WG_THREADS, tc::WgradXCfg::SMEM, st>>>(maps, X);
        add_launches(1);
        if (grad_w) {
          long long blocks = ((long long)total + 255) / 256;
          if (blocks > sm_count() * 8) blocks = sm_count() * 8;
          tc::k_unpack_wgrad<<<(int)blocks, 256, 0, st>>>(wacc, a->cout, cin_total, taps, grad_w, accumulate);
          add_launches(1);
        }
        return check_launch("dfb_conv2d_wgrad (cross-shift)");
      }
    }
    tc::WgradHaloParams Q;
    memset(&Q, 0, sizeof(Q));
    const uint32_t hbox[4] = {64, 10, 10, 1};
    const int ncol = a->cout <= 64 ? 64 : a->cout;
    const int max_pairs = 512 / ncol > 5 ? 5 : 512 / ncol;
    int ni = 0, coff = 0;
    for (int i = 0; i < a->n_src; ++i) {
      const uint64_t C = a->cin[i];
      const uint64_t dims[4] = {C, (uint64_t)W, (uint64_t)H, (uint64_t)a->n};
      const uint64_t str[3] = {C * 2, C * 2 * W, C * 2 * W * H};
      if ((rc = make_tensor_map_bf16(&maps.a[i], a->x[i], 4, dims, str, hbox, 128))) return rc;
      for (int c0 = 0; c0 < a->cin[i]; c0 += 64) {
        // nine taps = five tap pairs (the last one repeats tap 7); spread over ceil(5 / max_pairs) items as evenly as
        // possible (N = 128: 3 + 2 pairs, not 4 + 1; N = 256: 2 + 2 + 1)
        const int n_it = (5 + max_pairs - 1) / max_pairs;
        int pr0 = 0;
        for (int k = 0; k < n_it; ++k) {
          if (ni >= tc::WH_MAX_ITEMS) { set_error("dfb_conv2d_wgrad: too many work items"); return DFB_ERR_UNSUPPORTED; }
          const int cnt = (5 - pr0 + (n_it - k) - 1) / (n_it - k);
          tc::WHItem& it = Q.items[ni++];
          it.map = (int8_t)i; it.c0 = (int16_t)c0; it.ci_glob = (int16_t)(coff + c0); it.npairs = 0; it.dup_first = 0;
          for (int pr = pr0; pr < pr0 + cnt; ++pr) {
            if (pr < 4) { it.t1[it.npairs] = (int8_t)(2 * pr); it.t2[it.npairs] = (int8_t)(2 * pr + 1); }
            else { it.t1[it.npairs] = 7; it.t2[it.npairs] = 8; it.dup_first = 1; }  // (7, 8), 7 dropped
            ++it.npairs;
          }
          pr0 += cnt;
        }
      }
      coff += a->cin[i];
    }
    Q.n_items = ni;
    Q.n_img = a->n; Q.Ht = Ho; Q.Wt = Wo;
    Q.tiles_x = (Wo + 7) / 8; Q.tiles_y = (Ho + 7) / 8;
    Q.num_tiles = Q.tiles_x * Q.tiles_y * a->n;
    {
      int wsum = 0, blk = 0;
      for (int k = 0; k < ni; ++k) wsum += Q.items[k].npairs;
      const int target = wgrad_cta_target();
      int nsv[tc::WH_MAX_ITEMS];
      for (int k = 0; k < ni; ++k) {
        int ns = (target * Q.items[k].npairs + wsum / 2) / wsum;
        if (ns > Q.num_tiles) ns = Q.num_tiles;
        if (ns < 1) ns = 1;
        nsv[k] = ns;
        blk += ns;
      }
      // rounding must not push the launch past the target (one CTA per SM: a second, nearly empty wave costs a whole wave)
      while (blk > target) {
        int big = 0;
        for (int k = 1; k < ni; ++k) if (nsv[k] > nsv[big]) big = k;
        if (nsv[big] <= 1) break;
        --nsv[big]; --blk;
      }
      blk = 0;
      for (int k = 0; k < ni; ++k) {
        Q.items[k].blk0 = (int16_t)blk;
        Q.items[k].nsplit = (int16_t)nsv[k];
        blk += nsv[k];
      }
      Q.splits = blk;   // total CTAs
    }
    Q.cin_total = cin_total; Q.cout = a->cout; Q.wacc = wacc;
    const size_t total = (size_t)taps * a->cout * cin_total;
    if (!keep_wacc) cudaMemsetAsync(wacc, 0, total * sizeof(float), st);
    if (ncol == 64) rc = tc::launch_wgrad_halo<64>(maps, Q, st);
    else if (ncol == 128) rc = tc::launch_wgrad_halo<128>(maps, Q, st);
    else rc = tc::launch_wgrad_halo<256>(maps, Q, st);
    if (rc) return rc;
    if (grad_w) {
      long long blocks = ((long long)total + 255) / 256;
      if (blocks > sm_count() * 8) blocks = sm_count() * 8;
      tc::k_unpack_wgrad<<<(int)blocks, 256, 0, st>>>(wacc, a->cout, cin_total, taps, grad_w, accumulate);
      add_launches(1);
    }
    return check_launch("dfb_conv2d_wgrad");
  }
  if (s == 1) {
    for (int i = 0; i < a->n_src; ++i) {
      const uint64_t C = a->cin[i];
      const uint64_t dims[4] = {C, (uint64_t)W, (uint64_t)H, (uint64_t)a->n};
      const uint64_t str[3] = {C * 2, C * 2 * W, C * 2 * W * H};
      if ((rc = make_tensor_map_bf16(&maps.a[i], a->x[i], 4, dims, str, box, 128))) return rc;
    }
    for (int t = 0; t < taps; ++t) {
      const int dy = k == 3 ? t / 3 - 1 : 0, dx = k == 3 ? t % 3 - 1 : 0;
      int coff = 0;
      for (int i = 0; i < a->n_src; ++i) {
        for (int c0 = 0; c0 < a->cin[i]; c0 += 64) {  // a 32-channel source is covered by one zero-filled 64-wide box
          if (ng >= tc::MAX_GROUPS) { set_error("dfb_conv2d_wgrad: too many groups"); return DFB_ERR_UNSUPPORTED; }
          P.groups[ng++] = tc::WGroup{(int8_t)i, (int8_t)dy, (int8_t)dx, (int8_t)t, (int16_t)c0, (int16_t)(coff + c0)};
        }
        coff += a->cin[i];
      }
    }
  } else {
    if (a->n_src != 1 || H % 2 || W % 2) { set_error("dfb_conv2d_wgrad: stride 2 takes one source with even H, W"); return DFB_ERR_UNSUPPORTED; }
    const uint64_t C = a->cin[0];
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        const uint64_t dims[4] = {C, (uint64_t)W / 2, (uint64_t)H / 2, (uint64_t)a->n};
        const uint64_t str[3] = {C * 2 * 2, C * 2 * W * 2, C * 2 * W * H};
        const char* base = (const char*)a->x[0] + ((size_t)py * W + px) * C * 2;
        if ((rc = make_tensor_map_bf16(&maps.a[py * 2 + px], base, 4, dims, str, box, 128))) return rc;
      }
    for (int t = 0; t < taps; ++t) {
      const int ky = k == 3 ? t / 3 : 1, kx = k == 3 ? t % 3 : 1;
      const int py = (ky + 1) & 1, px = (kx + 1) & 1, dy = ky == 0 ? -1 : 0, dx = kx == 0 ? -1 : 0;
      for (int c0 = 0; c0 < a->cin[0]; c0 += 64)
        P.groups[ng++] = tc::WGroup{(int8_t)(py * 2 + px), (int8_t)dy, (int8_t)dx, (int8_t)t, (int16_t)c0, (int16_t)c0};
    }
  }
  P.n_groups = ng;
  P.n_pairs = (ng + 1) / 2;
  if (a->grad_bias) {
    if (ng % 2 == 0 || a->cout < 32) { set_error("dfb_conv2d_wgrad: grad_bias needs an odd number of 64-channel input groups (got %d)", ng); return DFB_ERR_UNSUPPORTED; }
    P.grad_bias = a->grad_bias;
  }
  P.n_img = a->n; P.Ht = Ho; P.Wt = Wo;
  P.tiles_x = (Wo + 7) / 8; P.tiles_y = (Ho + 7) / 8;
  P.num_tiles = P.tiles_x * P.tiles_y * a->n;
  int splits = wgrad_cta_target() / P.n_pairs;   // floor: a single wave at the default target of one CTA per SM
  if (splits > P.num_tiles) splits = P.num_tiles;
  if (splits < 1) splits = 1;
  P.splits = splits;
  P.cin_total = cin_total;
  P.cout = a->cout;
  P.wacc = wacc;
  const size_t total = (size_t)taps * a->cout * cin_total;
  if (!keep_wacc) cudaMemsetAsync(wacc, 0, total * sizeof(float), st);
  if (a->cout <= 64) rc = tc::launch_wgrad<64>(maps, P, st);
  else if (a->cout == 128) rc = tc::launch_wgrad<128>(maps, P, st);
  else rc = tc::launch_wgrad<256>(maps, P, st);
  if (rc) return rc;
  if (grad_w) {
    long long blocks = ((long long)total + 255) / 256;
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    tc::k_unpack_wgrad<<<(int)blocks, 256, 0, st>>>(wacc, a->cout, cin_total, taps, grad_w, accumulate);
    add_launches(1);
  }
  return check_launch("dfb_conv2d_wgrad");
}

// Every weight gradient of a model from its [tap][cout][cin] accumulator to the torch layout [cout][cin][tap] in ONE launch
// (the counterpart of dfb_conv_pack_weights_multi): table[i].first = running element offset, as there.
namespace dfb {
namespace tc {
__global__ void __launch_bounds__(256) k_unpack_wgrad_multi(const dfb_unpack_desc* __restrict__ table, int n_w, long long total) {
  __shared__ dfb_unpack_desc sd[64];
  for (int i = threadIdx.x; i < n_w; i += blockDim.x) sd[i] = table[i];
  __syncthreads();
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = n_w - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (sd[mid].first <= g) lo = mid; else hi = mid - 1;
    }
    const dfb_unpack_desc& d = sd[lo];
    const long long e = g - d.first;
    const int taps = d.ksize * d.ksize;
    const int t = (int)(e % taps);
    const int ci = (int)((e / taps) % d.cin);
    const int co = (int)(e / ((long long)taps * d.cin));
    const float v = d.wacc[((size_t)t * d.cout + co) * d.cin + ci];
    d.grad[e] = d.accumulate ? d.grad[e] + v : v;
  }
}
}  // namespace tc
}  // namespace dfb

extern "C" int dfb_wgrad_unpack_multi(const dfb_unpack_desc* table_dev, int n_weights, long long total_elems, void* stream_) {
  if (n_weights <= 0 || n_weights > 64 || total_elems <= 0 || !table_dev) {
    set_error("dfb_wgrad_unpack_multi: 1..64 weights and a device table are required");
    return DFB_ERR_ARG;
  }
  long long blocks = (total_elems + 255) / 256;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  dfb::tc::k_unpack_wgrad_multi<<<(int)blocks, 256, 0, (cudaStream_t)stream_>>>(table_dev, n_weights, total_elems);
  add_launches(1);
  return check_launch("dfb_wgrad_unpack_multi");
}
