// Microbenchmark (diagnostic, not part of the library): cycles per tcgen05.mma (M=128, K=16, bf16) when a chain of MMAs
// accumulates into ONE TMEM accumulator versus round-robin over several, for N = 64 / 128 / 256, with the standard
// K-major SWIZZLE_128B A layout and with the halo layout (128-byte-aligned start, 1280-byte group stride).
#include "../deflow_b200/csrc/tc_common.cuh"
#include <cstdio>
using namespace dfb::tc;

template <int N>
__global__ void __launch_bounds__(128, 1) k_chain(int iters, int nacc, int halo, int same_ab, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < (196 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 64 * 1024);
    const long long t0 = clock64();
    int acc = 0;
    for (int i = 0; i < iters; ++i) {
      // 9 "taps" x 4 K slices, like one 64-channel chunk of the halo kernel
      for (int t = 0; t < 9; ++t) {
        const uint32_t aa = halo ? a_addr + (uint32_t)((t / 3) * 10 + t % 3) * 128u : a_addr + (same_ab ? 0u : (uint32_t)(t % 3) * 16384u);
        const uint64_t adesc = make_smem_desc(aa, 16, halo ? 1280 : 1024, 2);
        const uint64_t bdesc = make_smem_desc(b_addr + (same_ab ? 0u : (uint32_t)(t % 4) * (uint32_t)(N * 128)), 16, 1024, 2);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_bf16(tmem_base + acc * N, adesc + 2 * k, bdesc + 2 * k, idesc, true);
          if (nacc > 1) { if (++acc == nacc) acc = 0; }
        }
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int N>
void run(int nacc, int halo, int same_ab, int blocks) {
  long long* d; cudaMalloc(&d, sizeof(long long) * blocks);
  const int iters = 200, smem = 200 * 1024;
  cudaFuncSetAttribute(k_chain<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_chain<N><<<blocks, 128, smem>>>(iters, nacc, halo, same_ab, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[256]; cudaMemcpy(h, d, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < blocks; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=%3d nacc=%d halo=%d same_ab=%d blocks=%3d : %.1f cycles/MMA (floor %d)  %s\n", N, nacc, halo, same_ab, blocks,
         (double)mx / (iters * 36.0), 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int blocks : {1, 148}) {
    for (int halo : {0, 1}) {
      run<64>(1, halo, 0, blocks); run<64>(2, halo, 0, blocks); run<64>(4, halo, 0, blocks);
      run<128>(1, halo, 0, blocks); run<128>(2, halo, 0, blocks); run<128>(4, halo, 0, blocks);
      run<256>(1, halo, 0, blocks); run<256>(2, halo, 0, blocks);
    }
    run<64>(1, 0, 1, blocks); run<64>(2, 0, 1, blocks);
  }
  return 0;
}
