// Bulk-async streaming of flat HBM arrays through a shared-memory ring (sm_100a).
//
// The HBM-bound passes of the path (BatchNorm/GELU apply and backward, bias sums) are simple per-element maps and
// reductions over tensors of 30-540 MB.  With plain loads the bytes in flight per SM are tied to registers and to
// the phase of each warp (a warp that is computing has nothing in flight); here one producer lane per CTA keeps a
// ring of STAGES x NIN chunks filled with `cp.async.bulk` (the TMA engine's 1-D copy, completion counted on an
// mbarrier), so the bytes in flight are STAGES x NIN x 8 KB per CTA no matter what the consumer warps are doing.
//
// Roles: warps 0..7 consume (256 threads), warp 8 lane 0 produces.  Chunk = 512 vectors of 16 bytes per input;
// consumer thread t owns vectors t and t + 256 of every chunk, so its channel octet (vector index mod C/8) never
// changes as long as 256 % (C/8) == 0.  CTA b streams chunks b, b + gridDim.x, ...
#pragma once
#include "tc_common.cuh"

namespace dfb {
namespace sp {

constexpr int CHUNK_VEC = 512;                 // 16-byte vectors per chunk and input
constexpr int CHUNK_BYTES = CHUNK_VEC * 16;    // 8 KB
constexpr int CONSUMERS = 256;
constexpr int THREADS = CONSUMERS + 32;

__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tc::smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}

template <int NIN, int STAGES>
struct Ring {
  static constexpr int BYTES = STAGES * NIN * CHUNK_BYTES + 2 * STAGES * 8;
  uint8_t* data;      // [STAGES][NIN][CHUNK_BYTES], 128-byte aligned
  uint64_t* full;     // [STAGES]
  uint64_t* empty;    // [STAGES]

  __device__ __forceinline__ Ring(uint8_t* smem) {
    data = smem;
    full = reinterpret_cast<uint64_t*>(smem + STAGES * NIN * CHUNK_BYTES);
    empty = full + STAGES;
  }
  // every thread of the CTA calls this once, before any role-specific code
  __device__ __forceinline__ void init() {
    if (threadIdx.x == 0) {
      for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], CONSUMERS / 32); }
      tc::fence_barrier_init();
    }
    __syncthreads();
  }
  __device__ __forceinline__ const uint4* slot(int s, int input) const {
    return reinterpret_cast<const uint4*>(data + ((size_t)s * NIN + input) * CHUNK_BYTES);
  }
  // producer lane: stream this CTA's chunks of the NIN inputs (flat arrays of n_vec 16-byte vectors each)
  __device__ __forceinline__ void produce(const void* const (&in)[NIN], long long n_vec) {
    const long long n_chunks = (n_vec + CHUNK_VEC - 1) / CHUNK_VEC;
    int it = 0;
    for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x, ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      tc::mbar_wait(&empty[s], ph ^ 1u);   // a fresh barrier passes the wait on parity 1
      const long long v0 = ch * CHUNK_VEC;
      const long long left = n_vec - v0;
      const uint32_t bytes = (uint32_t)(left < CHUNK_VEC ? left : CHUNK_VEC) * 16u;
      tc::mbar_arrive_expect_tx(&full[s], bytes * NIN);
#pragma unroll
      for (int i = 0; i < NIN; ++i)
        bulk_load(data + ((size_t)s * NIN + i) * CHUNK_BYTES, reinterpret_cast<const uint4*>(in[i]) + v0, bytes, &full[s]);
    }
  }
  // consumer side: wait until chunk `it` of this CTA has landed / hand its slot back (whole warp calls both)
  __device__ __forceinline__ void wait_full(int it) const { tc::mbar_wait(&full[it % STAGES], (uint32_t)(it / STAGES) & 1u); }
  __device__ __forceinline__ void release(int it) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) tc::mbar_arrive(&empty[it % STAGES]);
  }
};

}  // namespace sp
}  // namespace dfb
