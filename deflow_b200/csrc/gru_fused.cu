// Persistent fused ConvGRU decoder, forward: offset encoder -> num_iters GRU iterations -> MLP head, one kernel.
//
// Reference: ConvGRUDecoder.forward_single (OpenSceneFlow/src/models/basic/decoder.py:210-237) and ConvGRU.forward
// (:184-193): per sample, 3 k=1 Conv1d + 2 cats + 5 elementwise launches per iteration, cuDNN / cuBLAS.
//
// One CTA owns a tile of 128 points at a time (persistent over tiles).  All gate weights stay resident in shared
// memory as bf16 K-major SWIZZLE_128B operand tiles (Wz|Wr 96 KB, Wq 48 KB, W1 12 KB, loaded once by TMA).  The
// hidden state h (fp32) lives in the registers of 256 worker threads (thread = point row x half of the channels)
// and never touches HBM between iterations; the x = Linear(3,64)(offsets) operand is computed in-kernel.  Per
// iteration the workers write bf16(h) (then bf16(r*h)) into the A operand tile with the 128-byte swizzle applied by
// hand, one thread issues tcgen05.mma (M = 128 points, N = 256 / 128, K = 192) into TMEM, and the workers read the
// gate pre-activations back with tcgen05.ld for the sigmoid / tanh / lerp.  For the backward pass the kernel saves
// bf16 copies of the state entering every iteration, the final state and the head pre-activation.
#include "tc_common.cuh"
#include <stdlib.h>
#include "../../include/deflow_b200.h"

namespace dfb {
namespace tc {

// Warps 0..GW-1 are workers, warp GW is the control warp (TMA + MMA issue).  GW = 8: thread = point row x 64 hidden
// channels (168 registers); GW = 16: four warps per scheduler and TMEM lane quarter, thread = row x 32 channels (the CTA
// is then allocated as 20 warps: 96 registers per thread).  Selected at launch (DFB_GRU_WARPS, see dfb_gru_fused_*).
constexpr int GF_WZR = 0;                         // 3 x [256 x 128 B]
constexpr int GF_WQ = GF_WZR + 3 * 32768;         // 3 x [128 x 128 B]
constexpr int GF_W1 = GF_WQ + 3 * 16384;          // 3 x [ 32 x 128 B]
constexpr int GF_AH = GF_W1 + 3 * 4096;           // 2 x [128 x 128 B]  bf16(h) / bf16(r*h)
constexpr int GF_AX = GF_AH + 2 * 16384;          // 1 x [128 x 128 B]  bf16(x)
constexpr int GF_PAR = GF_AX + 16384;             // fp32 parameters
constexpr int GF_PAR_FLOATS = 384 + 32 + 96 + 3 + 192 + 64 + 5;  // bz|br|bq, b1, W2, b2, Woff, boff
constexpr int GF_BAR = GF_PAR + ((GF_PAR_FLOATS * 4 + 15) / 16) * 16;
constexpr int GF_SMEM = GF_BAR + 64 + 1024;

struct GruFusedMaps { CUtensorMap wzr, wq, w1, h, x; };

struct GruFusedParams {
  const __nv_bfloat16* h0;  // [n_pad,128] bf16 gathered pillar vectors
  const float* offs;    // [n,3]
  const float* par;     // bz[128] br[128] bq[128] b1[32] W2[3*32] b2[3] Woff[64*3] boff[64]
  int n, n_pad, iters;
  int save;             // training: hsave [iters+1][n_pad][128] / xsave [n_pad][64] are written by TMA from the A tiles
  __nv_bfloat16* y1;    // [n_pad][32] head pre-activation (bias included)
  float* flow;          // [n,3]
};

// One SFU op per gate value: tanh.approx.f32 (relative error ~2^-11, below the bf16 resolution of the operands the
// result feeds); sigmoid(v) = 0.5 tanh(v/2) + 0.5.  The gate stages are SFU-bound otherwise (exp + reciprocal).
__device__ __forceinline__ float ftanh(float v) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float fsigmoid(float v) { return fmaf(0.5f, ftanh(0.5f * v), 0.5f); }

// Packed fp32x2 forms (FADD2 / FMUL2 / FFMA2: two channels per instruction).  The gate stages were issue-bound (ncu r01:
// ~2300 scalar instructions per thread and GRU iteration for 2.5 k cycles of MMA time), the MUFU count is unchanged.
__device__ __forceinline__ float2 g2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 ftanh2(float2 v) { return make_float2(ftanh(v.x), ftanh(v.y)); }
__device__ __forceinline__ float2 fsigmoid2(float2 v) { return __ffma2_rn(ftanh2(__fmul2_rn(v, g2(0.5f))), g2(0.5f), g2(0.5f)); }
// shared-memory accesses with explicit state space (pointers derived from the aligned dynamic smem base are generic
// to the compiler: LD.E / ST.E instead of LDS / STS)
__device__ __forceinline__ float2 lds_f2(uint32_t saddr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_u4(uint32_t saddr, const uint4& u) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}
__device__ __forceinline__ uint32_t bf2_pack(float2 v) {
  const __nv_bfloat162 p = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<const uint32_t*>(&p);
}
__device__ __forceinline__ float2 bf2_unpack(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
// 8 consecutive channels (4 packed pairs) of row `row` -> chunk `chunk` of a [128 x 128 B] SWIZZLE_128B K-major tile
__device__ __forceinline__ void st_tile_chunk2(uint32_t tile_saddr, int row, int chunk, const float2 (&f)[4]) {
  sts_u4(tile_saddr + row * 128 + ((chunk ^ (row & 7)) << 4), make_uint4(bf2_pack(f[0]), bf2_pack(f[1]), bf2_pack(f[2]), bf2_pack(f[3])));
}

// store 8 consecutive channels (one 16-byte chunk) of row `row` into a [128 x 128 B] SWIZZLE_128B K-major tile
__device__ __forceinline__ void st_tile_chunk(uint8_t* tile, int row, int chunk, const float* f) {
  uint4 u;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(tile + row * 128 + ((chunk ^ (row & 7)) << 4)) = u;
}

template <int GW>
__global__ void __launch_bounds__(32 * (GW + 1), 1) k_gru_fused_fwd(const __grid_constant__ GruFusedMaps maps,
                                                                 const __grid_constant__ GruFusedParams P) {
  constexpr int GPARTS = GW / 4;       // channel parts per point row
  constexpr int GCPT = 128 / GPARTS;   // hidden channels per thread
  constexpr int GXPT = 64 / GPARTS;    // x channels per thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* par = reinterpret_cast<float*>(smem + GF_PAR);
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + GF_BAR);
  uint64_t* bar_a = bar_w + 1;
  uint64_t* bar_d = bar_w + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (P.n_pad + 127) / 128;

  for (int i = threadIdx.x; i < GF_PAR_FLOATS - 5; i += blockDim.x) par[i] = P.par[i];
  if (warp == GW) {
    if (lane == 0) {
      tma_prefetch_desc(&maps.wzr); tma_prefetch_desc(&maps.wq); tma_prefetch_desc(&maps.w1);
      mbar_init(bar_w, 1); mbar_init(bar_a, GW); mbar_init(bar_d, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float* bz = par; const float* br = par + 128; const float* bq = par + 256; const float* b1 = par + 384;
  const float* w2 = par + 416; const float* b2 = par + 512; const float* woff = par + 515; const float* boff = par + 707;

  if (warp == GW) {
    // ------------------------------------------------------------------ control: weights once, then MMA issue
    // The whole warp runs this code (warp-uniform control flow keeps descriptors and loop state in uniform registers);
    // lane 0 alone issues TMA, one elected lane issues the MMAs.  Under `if (lane == 0)` every tcgen05.mma cost ~30 SASS
    // instructions on one thread -- more than the MMA itself takes.
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_w, 3 * 32768 + 3 * 16384 + 3 * 4096);
      for (int c = 0; c < 3; ++c) {
        tma_load_2d(smem + GF_WZR + c * 32768, &maps.wzr, bar_w, c * 64, 0);
        tma_load_2d(smem + GF_WQ + c * 16384, &maps.wq, bar_w, c * 64, 0);
        tma_load_2d(smem + GF_W1 + c * 4096, &maps.w1, bar_w, c * 64, 0);
      }
    }
    __syncwarp();
    mbar_wait(bar_w, 0);
    uint32_t use_a = 0;
    // K-major SWIZZLE_128B descriptors as (lo, hi) halves: lo = start >> 4 | LBO(16 B) << 16, hi constant
    const uint32_t d_hi = (uint32_t)(make_smem_desc(0, 16, 1024, 2) >> 32);
    const uint32_t lo0 = (uint32_t)make_smem_desc(0, 16, 1024, 2);
    const uint32_t a_h = lo0 | ((smem_u32(smem + GF_AH) & 0x3FFFFu) >> 4), a_x = lo0 | ((smem_u32(smem + GF_AX) & 0x3FFFFu) >> 4);
    // slot >= 0: the A_H tile holds bf16(h) entering iteration `slot` (or the final state): save it for the backward
    auto gemm = [&](uint32_t w_off, uint32_t w_chunk_bytes, uint32_t idesc, uint32_t dcol, int row0, int slot, bool save_x) {
      mbar_wait(bar_a, use_a & 1); ++use_a;
      tc_fence_after();
      const bool st = P.save && slot >= 0;
      if (st && lane == 0) {
        tma_store_3d(&maps.h, smem + GF_AH, 0, row0, slot);
        tma_store_3d(&maps.h, smem + GF_AH + 16384, 64, row0, slot);
        if (save_x) tma_store_3d(&maps.x, smem + GF_AX, 0, row0, 0);
        tma_store_commit();
      }
      const uint32_t w_lo = lo0 | ((smem_u32(smem + w_off) & 0x3FFFFu) >> 4);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const uint32_t ad = c < 2 ? a_h + (uint32_t)c * (16384u >> 4) : a_x;
        const uint32_t bd = w_lo + (uint32_t)c * (w_chunk_bytes >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_lohi_warp(tmem + dcol, ad + 2 * k, d_hi, bd + 2 * k, d_hi, idesc, (uint32_t)((c | k) != 0));
      }
      if (st && lane == 0) tma_store_wait_read();  // the workers overwrite the tile after the commit below
      __syncwarp();
      umma_commit_warp(bar_d);
    };
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int row0 = tile * 128;
      for (int it = 0; it < P.iters; ++it) {
        gemm(GF_WZR, 32768, make_idesc_bf16(128, 256, 0, 0), 0, row0, it, it == 0);   // z | r pre-activations -> columns 0..255
        gemm(GF_WQ, 16384, make_idesc_bf16(128, 128, 0, 0), 256, row0, -1, false);    // q pre-activation      -> columns 256..383
      }
      gemm(GF_W1, 4096, make_idesc_bf16(128, 32, 0, 0), 384, row0, P.iters, P.iters == 0);  // MLP hidden layer -> columns 384..415
    }
    if (lane == 0) tma_store_wait_all();
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ workers
    const int q = warp & 3, part = warp >> 2;          // TMEM lane quarter (rows), channel part
    const int m = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    constexpr int CH0_CHUNKS = GCPT / 8;               // 16-byte chunks of this thread's channels per row
    const int ch0 = part * GCPT;                       // first hidden channel of this thread
    uint8_t* tile_h = smem + GF_AH + (ch0 >> 6) * 16384;   // the 64-channel tile its channels live in ...
    const int chunk0 = (ch0 & 63) >> 3;                    // ... starting at this 16-byte chunk of a row
    uint8_t* tile_x = smem + GF_AX;
    const uint32_t s_tile_h = smem_u32(tile_h), s_bz = smem_u32(bz), s_br = smem_u32(br), s_bq = smem_u32(bq);
    uint32_t use_d = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int p = tile * 128 + m;
      const bool valid = p < P.n;
      const bool inpad = p < P.n_pad;
      // x = Woff o + boff: this thread's GXPT of the 64 channels
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
      if (valid) { o0 = P.offs[3 * (size_t)p]; o1 = P.offs[3 * (size_t)p + 1]; o2 = P.offs[3 * (size_t)p + 2]; }
#pragma unroll
      for (int c8 = 0; c8 < GXPT / 8; ++c8) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int ch = part * GXPT + c8 * 8 + i;
          f[i] = valid ? fmaf(o2, woff[3 * ch + 2], fmaf(o1, woff[3 * ch + 1], fmaf(o0, woff[3 * ch], boff[ch]))) : 0.f;
        }
        st_tile_chunk(tile_x, m, part * (GXPT / 8) + c8, f);
      }
      float h[GCPT];
      if (valid) {
        const uint4* src = reinterpret_cast<const uint4*>(P.h0 + (size_t)p * 128 + ch0);
#pragma unroll
        for (int i = 0; i < CH0_CHUNKS; ++i) {
          const uint4 u = __ldg(src + i);
          const __nv_bfloat162* pp = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int k = 0; k < 4; ++k) { const float2 t = __bfloat1622float2(pp[k]); h[8 * i + 2 * k] = t.x; h[8 * i + 2 * k + 1] = t.y; }
        }
      } else {
#pragma unroll
        for (int i = 0; i < GCPT; ++i) h[i] = 0.f;
      }
      auto publish_h = [&](int) {  // bf16(h) -> A tile (the control warp saves the tile by TMA when training)
#pragma unroll
        for (int c8 = 0; c8 < CH0_CHUNKS; ++c8) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = valid ? h[c8 * 8 + i] : 0.f;
          st_tile_chunk(tile_h, m, chunk0 + c8, f);
        }
      };
      auto signal_a = [&]() {
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_a);
      };
      auto wait_d = [&]() {
        mbar_wait(bar_d, use_d & 1); ++use_d;
        tc_fence_after();
      };
      for (int it = 0; it < P.iters; ++it) {
        publish_h(it);
        signal_a();
        wait_d();
        // r = sigmoid(r_pre + br); A tile <- bf16(r * h)
#pragma unroll
        for (int cc = 0; cc < GCPT / 32; ++cc) {
          float v[32];
          tmem_ld32(tmem + lane_base + 128 + ch0 + cc * 32, v);
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            float2 f[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = cc * 32 + c8 * 8 + 2 * i;
              const float2 pre = __fadd2_rn(make_float2(v[c8 * 8 + 2 * i], v[c8 * 8 + 2 * i + 1]), lds_f2(s_br + 4 * (ch0 + j)));
              f[i] = __fmul2_rn(fsigmoid2(pre), make_float2(h[j], h[j + 1]));
            }
            st_tile_chunk2(s_tile_h, m, chunk0 + cc * 4 + c8, f);
          }
        }
        signal_a();
        wait_d();
        // h = (1 - z) h + z tanh(q_pre + bq),  z = sigmoid(z_pre + bz)
#pragma unroll
        for (int cc = 0; cc < GCPT / 32; ++cc) {
          float vz[32], vq[32];
          tmem_ld32(tmem + lane_base + ch0 + cc * 32, vz);
          tmem_ld32(tmem + lane_base + 256 + ch0 + cc * 32, vq);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const int j = cc * 32 + i;
            const float2 z = fsigmoid2(__fadd2_rn(make_float2(vz[i], vz[i + 1]), lds_f2(s_bz + 4 * (ch0 + j))));
            const float2 qq = ftanh2(__fadd2_rn(make_float2(vq[i], vq[i + 1]), lds_f2(s_bq + 4 * (ch0 + j))));
            const float2 hj = make_float2(h[j], h[j + 1]);
            const float2 hn = __ffma2_rn(z, __fadd2_rn(qq, make_float2(-hj.x, -hj.y)), hj);
            h[j] = hn.x; h[j + 1] = hn.y;
          }
        }
      }
      publish_h(P.iters);
      signal_a();
      wait_d();
      if (part == 0) {  // 32 hidden units of the MLP head: this thread's row, all 32 columns
        float v[32];
        tmem_ld32(tmem + lane_base + 384, v);
        float f0 = b2[0], f1 = b2[1], f2 = b2[2];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v[i] += b1[i];
          const float a = 0.5f * v[i] * (1.0f + erff(v[i] * 0.70710678118654752f));
          f0 = fmaf(a, w2[i], f0); f1 = fmaf(a, w2[32 + i], f1); f2 = fmaf(a, w2[64 + i], f2);
        }
        if (P.y1 && inpad) {
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            uint4 u;
            __nv_bfloat162* pp = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int i = 0; i < 4; ++i) pp[i] = __floats2bfloat162_rn(valid ? v[c8 * 8 + 2 * i] : 0.f, valid ? v[c8 * 8 + 2 * i + 1] : 0.f);
            *reinterpret_cast<uint4*>(P.y1 + (size_t)p * 32 + c8 * 8) = u;
          }
        }
        if (valid) { P.flow[3 * (size_t)p] = f0; P.flow[3 * (size_t)p + 1] = f1; P.flow[3 * (size_t)p + 2] = f2; }
      }
      tc_fence_before();  // TMEM reads of this tile are complete before the next tile's first arrive
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == GW) tmem_dealloc(tmem, 512);
}

}  // namespace tc
}  // namespace dfb

using namespace dfb;
using namespace dfb::tc;

static int make_rows_map(CUtensorMap* map, const void* base, int cols, int n_pad, int slabs);

// worker warps per CTA of the fused GRU kernels: DFB_GRU_WARPS=8|16 forces both; default = what measured fastest on
// B200 (r01): 8 for the forward (168 registers, no spills), 16 for the backward (its gate stages have more latency to hide)
static int gru_worker_warps(bool backward) {
  static int w = -1;
  if (w < 0) {
    const char* e = getenv("DFB_GRU_WARPS");
    w = e ? atoi(e) : 0;
  }
  if (w == 8 || w == 16) return w;
  return backward ? 16 : 8;
}

// wzr: bf16 [256][192] (Wz rows then Wr rows), wq: bf16 [128][192], w1: bf16 [32][192]; K order = [h(128), x(64)].
extern "C" int dfb_gru_fused_forward(const void* h0, const float* offsets, const void* wzr, const void* wq,
                                     const void* w1, const float* par, int n, int n_pad, int iters, void* hsave,
                                     void* xsave, void* y1, float* flow, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n <= 0 || n_pad < n || iters < 0 || iters > 64) { set_error("dfb_gru_fused_forward: bad sizes"); return DFB_ERR_ARG; }
  GruFusedMaps maps;
  GruFusedParams P;
  memset(&maps, 0, sizeof(maps));
  int rc;
  const uint64_t str[1] = {192 * 2};
  {
    const uint64_t d[2] = {192, 256}; const uint32_t b[2] = {64, 256};
    if ((rc = make_tensor_map_bf16(&maps.wzr, wzr, 2, d, str, b, 128))) return rc;
  }
  {
    const uint64_t d[2] = {192, 128}; const uint32_t b[2] = {64, 128};
    if ((rc = make_tensor_map_bf16(&maps.wq, wq, 2, d, str, b, 128))) return rc;
  }
  {
    const uint64_t d[2] = {192, 32}; const uint32_t b[2] = {64, 32};
    if ((rc = make_tensor_map_bf16(&maps.w1, w1, 2, d, str, b, 128))) return rc;
  }
  P.h0 = (const __nv_bfloat16*)h0; P.offs = offsets; P.par = par; P.n = n; P.n_pad = n_pad; P.iters = iters;
  P.save = (hsave && xsave) ? 1 : 0;
  P.y1 = (__nv_bfloat16*)y1; P.flow = flow;
  if (P.save) {
    if ((rc = make_rows_map(&maps.h, hsave, 128, n_pad, iters + 1))) return rc;
    if ((rc = make_rows_map(&maps.x, xsave, 64, n_pad, 1))) return rc;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_gru_fused_fwd<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, GF_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gru_fused_fwd<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, GF_SMEM);
    if (e != cudaSuccess) { set_error("gru_fused: cannot reserve %d bytes of shared memory: %s", GF_SMEM, cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    configured = true;
  }
  const int tiles = (n_pad + 127) / 128;
  int grid = sm_count();
  if (grid > tiles) grid = tiles;
  if (gru_worker_warps(false) == 16) k_gru_fused_fwd<16><<<grid, 32 * 17, GF_SMEM, st>>>(maps, P);
  else k_gru_fused_fwd<8><<<grid, 32 * 9, GF_SMEM, st>>>(maps, P);
  add_launches(1);
  return check_launch("dfb_gru_fused_forward");
}

// ================================================================================================
// Fused backward of the GRU iterations.  Per tile of 128 points, for t = iters-1 .. 0:
//   recompute  z, r = sigmoid([h_t, x] Wzr^T + b),  rh = r * h_t,  q = tanh([rh, x] Wq^T + bq)      (G1, G2)
//   dq_pre = dh z (1 - q^2);  dz_pre = dh (q - h_t) z (1 - z);  dh <- dh (1 - z)
//   d_rh = dq_pre Wq[:, :128];  d_x += dq_pre Wq[:, 128:] + dz_pre Wz[:, 128:]                      (G3)
//   dr_pre = d_rh h_t r (1 - r);  dh += d_rh r
//   dh += dz_pre Wz[:, :128] + dr_pre Wr[:, :128];  d_x += dr_pre Wr[:, 128:]                        (G3, G4)
// The data-gradient GEMMs read the SAME resident weight tiles as the forward ones, through MN-major descriptors
// (the transposed view).  dh lives in registers across iterations, d_x accumulates in TMEM across iterations.
// rh, dq_pre, dzr_pre are written to HBM in bf16 for the weight-gradient GEMMs (dfb_conv2d_wgrad over the point list).
namespace dfb {
namespace tc {

constexpr int GB_WZR = 0;
constexpr int GB_WQ = GB_WZR + 3 * 32768;
constexpr int GB_P = GB_WQ + 3 * 16384;      // 2 tiles: h_t -> r*h_t -> dq_pre -> dr_pre
constexpr int GB_Q = GB_P + 2 * 16384;       // 2 tiles: dz_pre
constexpr int GB_X = GB_Q + 2 * 16384;       // 1 tile
constexpr int GB_PAR = GB_X + 16384;         // bz | br | bq
constexpr int GB_BAR = GB_PAR + 384 * 4;
constexpr int GB_SMEM = GB_BAR + 64 + 1008;

struct GruBwdMaps { CUtensorMap wzr, wq, h, x, rh, dq, dzr; };

struct GruBwdParams {
  const __nv_bfloat16* dh_in;   // [n_pad][128] gradient w.r.t. the final state (from the MLP head)
  const __nv_bfloat16* dx_in;   // [n_pad][64]  gradient w.r.t. x from the MLP head
  const float* par;             // bz | br | bq
  int n, n_pad, iters;
  __nv_bfloat16* dh0;           // [n_pad][128]  gradient w.r.t. the gathered pillar vectors
  float* dx;                    // [n_pad][64]   gradient w.r.t. x (all iterations + head)
};

__device__ __forceinline__ void ld_bf16x8(const __nv_bfloat16* p, float* f) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* pp = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(pp[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ void st_bf16x8(__nv_bfloat16* p, const float* f) {
  uint4 u;
  __nv_bfloat162* pp = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) pp[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float* f) {
  const __nv_bfloat162* pp = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(pp[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

// All HBM traffic of the iteration loop goes through TMA: the control thread loads the h_t / x tiles straight into
// the swizzled operand tiles and stores the rh / dq_pre / dzr_pre tiles from them (bulk async groups), so the 256
// worker threads touch only TMEM, registers and shared memory inside the loop.
template <int GW>
__global__ void __launch_bounds__(32 * (GW + 1), 1) k_gru_fused_bwd(const __grid_constant__ GruBwdMaps maps,
                                                                 const __grid_constant__ GruBwdParams P) {
  constexpr int GPARTS = GW / 4;       // channel parts per point row
  constexpr int GCPT = 128 / GPARTS;   // hidden channels per thread
  constexpr int GXPT = 64 / GPARTS;    // x channels per thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* par = reinterpret_cast<float*>(smem + GB_PAR);
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + GB_BAR);
  uint64_t* bar_a = bar_w + 1;
  uint64_t* bar_d = bar_w + 2;
  uint64_t* bar_h = bar_w + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (P.n_pad + 127) / 128;

  for (int i = threadIdx.x; i < 384; i += blockDim.x) par[i] = P.par[i];
  if (warp == GW) {
    if (lane == 0) {
      tma_prefetch_desc(&maps.wzr); tma_prefetch_desc(&maps.wq); tma_prefetch_desc(&maps.h); tma_prefetch_desc(&maps.x);
      tma_prefetch_desc(&maps.rh); tma_prefetch_desc(&maps.dq); tma_prefetch_desc(&maps.dzr);
      mbar_init(bar_w, 1); mbar_init(bar_a, GW); mbar_init(bar_d, 1); mbar_init(bar_h, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float* bz = par; const float* br = par + 128; const float* bq = par + 256;
  // TMEM columns: [0,256) z|r pre-activations, later [0,128) = d_h; [256,384) q pre-activation, later d_rh;
  // [448,512) d_x accumulator
  constexpr uint32_t C_ZR = 0, C_Q = 256, C_DX = 448;

  if (warp == GW) {
    // whole-warp control flow, lane 0 issues TMA, one elected lane issues the MMAs (see k_gru_fused_fwd)
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_w, 3 * 32768 + 3 * 16384);
      for (int c = 0; c < 3; ++c) {
        tma_load_2d(smem + GB_WZR + c * 32768, &maps.wzr, bar_w, c * 64, 0);
        tma_load_2d(smem + GB_WQ + c * 16384, &maps.wq, bar_w, c * 64, 0);
      }
    }
    __syncwarp();
    mbar_wait(bar_w, 0);
    uint32_t use_a = 0, use_d = 0, use_h = 0;
    uint8_t* sp = smem + GB_P; uint8_t* sq = smem + GB_Q; uint8_t* sx = smem + GB_X;
    const uint32_t t_p = smem_u32(sp), t_q = smem_u32(sq), t_x = smem_u32(sx);
    const uint32_t wzr = smem_u32(smem + GB_WZR), wq = smem_u32(smem + GB_WQ);
    auto lo_of = [](uint32_t addr, uint32_t lbo_bytes) { return ((addr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16); };
    const uint32_t d_hi = (uint32_t)(make_smem_desc(0, 16, 1024, 2) >> 32);   // SBO 1024, SWIZZLE_128B: shared by all operands
    auto wait_a = [&]() { mbar_wait(bar_a, use_a & 1); ++use_a; tc_fence_after(); };
    auto commit_d = [&]() { umma_commit_warp(bar_d); ++use_d; };
    auto gemm_fwd = [&](uint32_t w_base, uint32_t w_chunk, uint32_t idesc, uint32_t dcol) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const uint32_t ad = lo_of(c < 2 ? t_p + c * 16384 : t_x, 16);
        const uint32_t bd = lo_of(w_base + c * w_chunk, 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_lohi_warp(tmem + dcol, ad + 2 * k, d_hi, bd + 2 * k, d_hi, idesc, (uint32_t)((c | k) != 0));
      }
    };
    // D[128, N'] (+)= G[128, 128] (A tile pair, K-major) x W[rows row0..row0+127][columns of chunks c0..]: the weight
    // tiles are read MN-major (transposed view): K' = weight rows, N' = weight columns.
    auto gemm_dgrad = [&](uint32_t a_tile, uint32_t w_base, uint32_t w_chunk, int row0, int c0, uint32_t idesc, uint32_t dcol,
                          bool accumulate) {
      const uint32_t b0 = w_base + c0 * w_chunk + row0 * 128;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t ad = lo_of(a_tile + (i >> 2) * 16384, 16) + 2 * (i & 3);
        const uint32_t bd = lo_of(b0 + i * 2048, w_chunk);
        umma_bf16_lohi_warp(tmem + dcol, ad, d_hi, bd, d_hi, idesc, (uint32_t)(accumulate || i != 0));
      }
    };
    constexpr uint32_t I_F256 = make_idesc_bf16(128, 256, 0, 0), I_F128 = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t I_D128 = make_idesc_bf16(128, 128, 0, 1), I_D64 = make_idesc_bf16(128, 64, 0, 1);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int row0 = tile * 128;
      for (int it = P.iters - 1; it >= 0; --it) {
        const bool first = it == P.iters - 1;
        // P (and X on the first iteration) are free: the previous G4 was waited for below
        if (lane == 0) {
          mbar_arrive_expect_tx(bar_h, 32768 + (first ? 16384 : 0));
          tma_load_3d(sp, &maps.h, bar_h, 0, row0, it);
          tma_load_3d(sp + 16384, &maps.h, bar_h, 64, row0, it);
          if (first) tma_load_3d(sx, &maps.x, bar_h, 0, row0, 0);
        }
        __syncwarp();
        mbar_wait(bar_h, use_h & 1); ++use_h;
        wait_a();  // workers are done with the TMEM contents of the previous stage
        gemm_fwd(wzr, 32768, I_F256, C_ZR);
        commit_d();
        wait_a();  // P = r * h_t
        if (lane == 0) {
          tma_store_3d(&maps.rh, sp, 0, row0, it); tma_store_3d(&maps.rh, sp + 16384, 64, row0, it);
          tma_store_commit();
        }
        gemm_fwd(wq, 16384, I_F128, C_Q);
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        commit_d();
        wait_a();  // P = dq_pre, Q = dz_pre
        if (lane == 0) {
          tma_store_3d(&maps.dq, sp, 0, row0, it); tma_store_3d(&maps.dq, sp + 16384, 64, row0, it);
          tma_store_3d(&maps.dzr, sq, 0, row0, it); tma_store_3d(&maps.dzr, sq + 16384, 64, row0, it);
          tma_store_commit();
        }
        gemm_dgrad(t_p, wq, 16384, 0, 0, I_D128, C_Q, false);        // d_rh  = dq Wq[:, 0:128]
        gemm_dgrad(t_p, wq, 16384, 0, 2, I_D64, C_DX, !first);       // d_x  += dq Wq[:, 128:192]
        gemm_dgrad(t_q, wzr, 32768, 0, 0, I_D128, C_ZR, false);      // d_h   = dz Wz[:, 0:128]
        gemm_dgrad(t_q, wzr, 32768, 0, 2, I_D64, C_DX, true);        // d_x  += dz Wz[:, 128:192]
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        commit_d();
        wait_a();  // P = dr_pre
        if (lane == 0) {
          tma_store_3d(&maps.dzr, sp, 128, row0, it); tma_store_3d(&maps.dzr, sp + 16384, 192, row0, it);
          tma_store_commit();
        }
        gemm_dgrad(t_p, wzr, 32768, 128, 0, I_D128, C_ZR, true);     // d_h  += dr Wr[:, 0:128]
        gemm_dgrad(t_p, wzr, 32768, 128, 2, I_D64, C_DX, true);      // d_x  += dr Wr[:, 128:192]
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        commit_d();
        mbar_wait(bar_d, (use_d - 1) & 1);  // G4 has finished reading P before the next h_t lands in it
      }
    }
    if (lane == 0) tma_store_wait_all();
    __syncwarp();
  } else {
    const int q = warp & 3, part = warp >> 2;          // TMEM lane quarter (rows), channel part (GCPT channels)
    const int m = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    constexpr int NCH = GCPT / 8;                      // 16-byte chunks of this thread's channels per row
    const int ch0 = part * GCPT;
    const int chunk0 = (ch0 & 63) >> 3;
    uint8_t* tile_p = smem + GB_P + (ch0 >> 6) * 16384;
    uint8_t* tile_q = smem + GB_Q + (ch0 >> 6) * 16384;
    const uint32_t s_tile_p = smem_u32(tile_p), s_tile_q = smem_u32(tile_q);
    const uint32_t s_bz = smem_u32(bz), s_br = smem_u32(br), s_bq = smem_u32(bq);
    uint32_t use_d = 0, use_h = 0;
    auto signal_a = [&]() {
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a);
    };
    auto wait_d = [&]() { mbar_wait(bar_d, use_d & 1); ++use_d; tc_fence_after(); };
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int p = tile * 128 + m;
      const bool valid = p < P.n;
      const bool inpad = p < P.n_pad;
      const size_t prow = inpad ? (size_t)p : 0;
      const float2 vmask = g2(valid ? 1.0f : 0.0f);
      float dh[GCPT];
#pragma unroll
      for (int c8 = 0; c8 < NCH; ++c8) {
        if (valid) ld_bf16x8(P.dh_in + prow * 128 + ch0 + c8 * 8, &dh[c8 * 8]);
        else {
#pragma unroll
          for (int i = 0; i < 8; ++i) dh[c8 * 8 + i] = 0.f;
        }
      }
      for (int it = P.iters - 1; it >= 0; --it) {
        signal_a();  // TMEM of the previous stage has been read
        wait_d();    // z | r pre-activations
        mbar_wait(bar_h, use_h & 1); ++use_h;  // the h_t tile written by TMA is visible to this thread
        uint4 hp[NCH];                          // this row's GCPT channels of h_t, packed bf16
#pragma unroll
        for (int c8 = 0; c8 < NCH; ++c8) hp[c8] = *reinterpret_cast<const uint4*>(tile_p + m * 128 + (((chunk0 + c8) ^ (m & 7)) << 4));
#pragma unroll
        for (int cc = 0; cc < GCPT / 32; ++cc) {
          float v[32];
          tmem_ld32(tmem + lane_base + C_ZR + 128 + ch0 + cc * 32, v);
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            const uint4 hw = hp[cc * 4 + c8];
            const uint32_t hwv[4] = {hw.x, hw.y, hw.z, hw.w};
            float2 f[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = cc * 32 + c8 * 8 + 2 * i;
              const float2 pre = __fadd2_rn(make_float2(v[c8 * 8 + 2 * i], v[c8 * 8 + 2 * i + 1]), lds_f2(s_br + 4 * (ch0 + j)));
              f[i] = __fmul2_rn(__fmul2_rn(fsigmoid2(pre), bf2_unpack(hwv[i])), vmask);
            }
            st_tile_chunk2(s_tile_p, m, chunk0 + cc * 4 + c8, f);
          }
        }
        signal_a();
        wait_d();  // q pre-activation
#pragma unroll
        for (int cc = 0; cc < GCPT / 16; ++cc) {
          float vz[16], vq[16];
          tmem_ld16(tmem + lane_base + C_ZR + ch0 + cc * 16, vz);
          tmem_ld16(tmem + lane_base + C_Q + ch0 + cc * 16, vq);
#pragma unroll
          for (int c8 = 0; c8 < 2; ++c8) {
            const uint4 hw = hp[cc * 2 + c8];
            const uint32_t hwv[4] = {hw.x, hw.y, hw.z, hw.w};
            float2 fq[4], fz[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = cc * 16 + c8 * 8 + 2 * i;
              const float2 z = fsigmoid2(__fadd2_rn(make_float2(vz[c8 * 8 + 2 * i], vz[c8 * 8 + 2 * i + 1]), lds_f2(s_bz + 4 * (ch0 + j))));
              const float2 qq = ftanh2(__fadd2_rn(make_float2(vq[c8 * 8 + 2 * i], vq[c8 * 8 + 2 * i + 1]), lds_f2(s_bq + 4 * (ch0 + j))));
              const float2 nhh = bf2_unpack(hwv[i] ^ 0x80008000u);          // -h_t
              const float2 g = make_float2(dh[j], dh[j + 1]);
              const float2 omz = __ffma2_rn(z, g2(-1.0f), g2(1.0f));         // 1 - z
              const float2 gz = __fmul2_rn(g, z);
              fq[i] = __fmul2_rn(gz, __ffma2_rn(qq, make_float2(-qq.x, -qq.y), g2(1.0f)));   // g z (1 - q^2)
              fz[i] = __fmul2_rn(__fmul2_rn(gz, omz), __fadd2_rn(qq, nhh));                  // g (q - h) z (1 - z)
              const float2 dn = __fmul2_rn(g, omz);
              dh[j] = dn.x; dh[j + 1] = dn.y;
            }
            st_tile_chunk2(s_tile_p, m, chunk0 + cc * 2 + c8, fq);
            st_tile_chunk2(s_tile_q, m, chunk0 + cc * 2 + c8, fz);
          }
        }
        signal_a();
        wait_d();  // d_rh in C_Q, partial d_h in C_ZR[0,128); r pre-activation still in C_ZR[128,256)
#pragma unroll
        for (int cc = 0; cc < GCPT / 16; ++cc) {
          float vr[16], vg[16];
          tmem_ld16(tmem + lane_base + C_ZR + 128 + ch0 + cc * 16, vr);
          tmem_ld16(tmem + lane_base + C_Q + ch0 + cc * 16, vg);
#pragma unroll
          for (int c8 = 0; c8 < 2; ++c8) {
            const uint4 hw = hp[cc * 2 + c8];
            const uint32_t hwv[4] = {hw.x, hw.y, hw.z, hw.w};
            float2 f[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = cc * 16 + c8 * 8 + 2 * i;
              const float2 r = fsigmoid2(__fadd2_rn(make_float2(vr[c8 * 8 + 2 * i], vr[c8 * 8 + 2 * i + 1]), lds_f2(s_br + 4 * (ch0 + j))));
              const float2 g = __fmul2_rn(make_float2(vg[c8 * 8 + 2 * i], vg[c8 * 8 + 2 * i + 1]), vmask);
              const float2 gr = __fmul2_rn(g, r);
              f[i] = __fmul2_rn(__fmul2_rn(gr, bf2_unpack(hwv[i])), __ffma2_rn(r, g2(-1.0f), g2(1.0f)));   // g h r (1 - r)
              dh[j] += gr.x; dh[j + 1] += gr.y;
            }
            st_tile_chunk2(s_tile_p, m, chunk0 + cc * 2 + c8, f);
          }
        }
        signal_a();
        wait_d();  // d_h complete in C_ZR[0,128)
#pragma unroll
        for (int cc = 0; cc < GCPT / 32; ++cc) {
          float v[32];
          tmem_ld32(tmem + lane_base + C_ZR + ch0 + cc * 32, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) dh[cc * 32 + i] += valid ? v[i] : 0.f;
        }
      }
      // outputs of the tile: dh0 (bf16) and d_x = TMEM accumulator + the head's contribution
      if (inpad) {
#pragma unroll
        for (int c8 = 0; c8 < NCH; ++c8) st_bf16x8(P.dh0 + prow * 128 + ch0 + c8 * 8, &dh[c8 * 8]);
      }
      {
        float v[GXPT];
        if (P.iters > 0) {
          if constexpr (GXPT == 32) tmem_ld32(tmem + lane_base + C_DX + part * GXPT, v);
          else tmem_ld16(tmem + lane_base + C_DX + part * GXPT, v);
        }
        else {
#pragma unroll
          for (int i = 0; i < GXPT; ++i) v[i] = 0.f;
        }
        if (inpad) {
#pragma unroll
          for (int c8 = 0; c8 < GXPT / 8; ++c8) {
            float f[8];
            ld_bf16x8(P.dx_in + prow * 64 + part * GXPT + c8 * 8, f);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = valid ? f[i] + v[c8 * 8 + i] : 0.f;
            float4* dst = reinterpret_cast<float4*>(P.dx + prow * 64 + part * GXPT + c8 * 8);
            dst[0] = make_float4(f[0], f[1], f[2], f[3]);
            dst[1] = make_float4(f[4], f[5], f[6], f[7]);
          }
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == GW) tmem_dealloc(tmem, 512);
}

}  // namespace tc
}  // namespace dfb

static int make_rows_map(CUtensorMap* map, const void* base, int cols, int n_pad, int slabs) {
  // bf16 [slabs][n_pad][cols], box = 64 columns x 128 rows x 1 slab
  const uint64_t d[3] = {(uint64_t)cols, (uint64_t)n_pad, (uint64_t)slabs};
  const uint64_t st[2] = {(uint64_t)cols * 2, (uint64_t)cols * 2 * (uint64_t)n_pad};
  const uint32_t b[3] = {64, 128, 1};
  return dfb::tc::make_tensor_map_bf16(map, base, 3, d, st, b, 128);
}

extern "C" int dfb_gru_fused_backward(const void* hsave, const void* xsave, const void* dh_in, const void* dx_in,
                                      const void* wzr, const void* wq, const float* par, int n, int n_pad, int iters,
                                      void* rh, void* dq, void* dzr, void* dh0, float* dx, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n <= 0 || n_pad < n || iters <= 0 || iters > 64) { set_error("dfb_gru_fused_backward: bad sizes"); return DFB_ERR_ARG; }
  GruBwdMaps maps;
  GruBwdParams P;
  memset(&maps, 0, sizeof(maps));
  int rc;
  const uint64_t str[1] = {192 * 2};
  {
    const uint64_t d[2] = {192, 256}; const uint32_t b[2] = {64, 256};
    if ((rc = make_tensor_map_bf16(&maps.wzr, wzr, 2, d, str, b, 128))) return rc;
  }
  {
    const uint64_t d[2] = {192, 128}; const uint32_t b[2] = {64, 128};
    if ((rc = make_tensor_map_bf16(&maps.wq, wq, 2, d, str, b, 128))) return rc;
  }
  if ((rc = make_rows_map(&maps.h, hsave, 128, n_pad, iters + 1))) return rc;
  if ((rc = make_rows_map(&maps.x, xsave, 64, n_pad, 1))) return rc;
  if ((rc = make_rows_map(&maps.rh, rh, 128, n_pad, iters))) return rc;
  if ((rc = make_rows_map(&maps.dq, dq, 128, n_pad, iters))) return rc;
  if ((rc = make_rows_map(&maps.dzr, dzr, 256, n_pad, iters))) return rc;
  P.dh_in = (const __nv_bfloat16*)dh_in; P.dx_in = (const __nv_bfloat16*)dx_in; P.par = par;
  P.n = n; P.n_pad = n_pad; P.iters = iters;
  P.dh0 = (__nv_bfloat16*)dh0; P.dx = dx;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_gru_fused_bwd<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, GB_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gru_fused_bwd<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, GB_SMEM);
    if (e != cudaSuccess) { set_error("gru_fused_bwd: cannot reserve %d bytes of shared memory: %s", GB_SMEM, cudaGetErrorString(e)); return DFB_ERR_CUDA; }
    configured = true;
  }
  const int tiles = (n_pad + 127) / 128;
  int grid = sm_count();
  if (grid > tiles) grid = tiles;
  if (gru_worker_warps(true) == 16) k_gru_fused_bwd<16><<<grid, 32 * 17, GB_SMEM, st>>>(maps, P);
  else k_gru_fused_bwd<8><<<grid, 32 * 9, GB_SMEM, st>>>(maps, P);
  add_launches(1);
  return check_launch("dfb_gru_fused_backward");
}
