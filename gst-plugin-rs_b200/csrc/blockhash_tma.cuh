// blockhash_tma.cuh -- videocompare block sums as a TMA-fed streaming reduction (sm_100a).
//
// hashed_image.rs:24-64 -> image_hasher blockhash fast path: sums[by*hw+bx] += (A==0 ? 765 : R+G+B).
// One CTA owns a (bw x rows) tile inside ONE hash block.  Thread 0 posts one cp.async.bulk per tile row
// (row segments are contiguous: bw*4 bytes) against a single mbarrier, so the whole tile (<= 32 KB) is in
// flight at once without holding a single register; all threads then sum the tile out of shared memory with
// conflict-free 128-bit reads (one dp4a per pixel), warp-reduce, and issue ONE atomicAdd per CTA.
// Integer adds are order independent => exact.
#pragma once
#include "tma_pipe.cuh"

namespace b200vfx {

constexpr int kBlockhashTileBytes = 32768;

__global__ void __launch_bounds__(128) blockhash_sums_tma_kernel(const uint8_t *__restrict__ src, long stride, int bw, int bh,
                                                                int hw, int rows_per_cta, uint32_t *__restrict__ sums) {
  extern __shared__ __align__(128) uint8_t bh_smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t wsum[4];
  const int bx = blockIdx.x, by = blockIdx.y;
  const int y0 = by * bh + blockIdx.z * rows_per_cta;
  const int rows = min(rows_per_cta, (by + 1) * bh - y0);
  const uint32_t row_bytes = (uint32_t)bw * 4u;
  if (rows <= 0) return;
  if (threadIdx.x == 0) {
    tma::mbar_init(&bar, 1);
    tma::fence_barrier_init();
    const uint64_t pol = tma::policy_evict_first();
    tma::mbar_expect_tx(&bar, row_bytes * (uint32_t)rows);
    const uint8_t *p = src + (size_t)y0 * stride + (size_t)bx * row_bytes;
    for (int r = 0; r < rows; r++) tma::bulk_load(bh_smem + (size_t)r * row_bytes, p + (size_t)r * stride, row_bytes, &bar, pol);
  }
  __syncthreads();
  tma::mbar_wait(&bar, 0);
  const uint4 *tile = reinterpret_cast<const uint4 *>(bh_smem);
  const int n16 = (int)((row_bytes >> 4) * (uint32_t)rows);
  uint32_t acc = 0;
  for (int i = threadIdx.x; i < n16; i += blockDim.x) {
    const uint4 q = tile[i];
    acc += (q.x >> 24) ? __dp4a(q.x, 0x00010101u, 0u) : 765u;   // A == 0 counts as white (765)
    acc += (q.y >> 24) ? __dp4a(q.y, 0x00010101u, 0u) : 765u;
    acc += (q.z >> 24) ? __dp4a(q.z, 0x00010101u, 0u) : 765u;
    acc += (q.w >> 24) ? __dp4a(q.w, 0x00010101u, 0u) : 765u;
  }
  acc = __reduce_add_sync(0xFFFFFFFFu, acc);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(sums + by * hw + bx, wsum[0] + wsum[1] + wsum[2] + wsum[3]);
}

}  // namespace b200vfx
