// stream_map.cuh -- TMA-pipelined streaming skeleton for 4-byte -> 4-byte pixel maps.
//
// Persistent CTAs; the frame moves HBM -> shared -> HBM through the TMA engine in TILE-byte bulk copies
// (cp.async.bulk + mbarrier ring, STAGES deep) issued by thread 0.  All threads transform the tile IN PLACE
// in shared memory (lane-consecutive 32-bit accesses: conflict free, and every table-gather instruction
// covers 32 consecutive pixels), then thread 0 bulk-stores it.
// Requirements (checked by the launcher): rows 16-byte aligned, row_bytes % 16 == 0.
#pragma once
#include "tma_pipe.cuh"

namespace b200vfx {

template <int TILE, int STAGES>
constexpr int stream_smem_bytes() { return TILE * STAGES + 8 * STAGES + 64; }

template <int TILE, int STAGES, int THREADS, int B, typename PixelOp>
__device__ __forceinline__ void stream_map_u32(const PixelOp &op, const uint8_t *__restrict__ src, long sstride,
                                               uint8_t *__restrict__ dst, long dstride, int row_bytes, int height) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // PDL: frames are independent (see b200vfx.cu launch_k)
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + STAGES * TILE);
  const int tid = threadIdx.x;
  const int tiles_per_row = (row_bytes + TILE - 1) / TILE;
  const long long ntiles = (long long)tiles_per_row * height;
  const long long first = blockIdx.x, step = gridDim.x;
  const long long mine = first < ntiles ? (ntiles - first + step - 1) / step : 0;  // tiles this CTA owns
  uint64_t pol_stream = 0;
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) tma::mbar_init(&bars[s], 1);
    tma::fence_barrier_init();
    pol_stream = tma::policy_evict_first();
  }
  __syncthreads();
  auto tile_geom = [&](long long i, size_t &soff, size_t &doff, uint32_t &nbytes) {
    const long long t = first + i * step;
    const int row = (int)(t / tiles_per_row), c = (int)(t % tiles_per_row);
    const int off = c * TILE;
    nbytes = (uint32_t)min(TILE, row_bytes - off);
    soff = (size_t)row * sstride + off;
    doff = (size_t)row * dstride + off;
  };
  auto issue_load = [&](long long i) {  // thread 0 only
    size_t soff, doff; uint32_t nbytes;
    tile_geom(i, soff, doff, nbytes);
    const int s = (int)(i % STAGES);
    tma::mbar_expect_tx(&bars[s], nbytes);
    tma::bulk_load(smem_raw + s * TILE, src + soff, nbytes, &bars[s], pol_stream);
  };
  if (tid == 0)
    for (long long i = 0; i < mine && i < STAGES - 1; i++) issue_load(i);
  for (long long i = 0; i < mine; i++) {
    const int s = (int)(i % STAGES);
    size_t soff, doff; uint32_t nbytes;
    tile_geom(i, soff, doff, nbytes);
    tma::mbar_wait(&bars[s], (uint32_t)((i / STAGES) & 1));
    uint32_t *tile = reinterpret_cast<uint32_t *>(smem_raw + s * TILE);
    const int npx = (int)(nbytes >> 2);
    for (int j0 = tid; j0 < npx; j0 += B * THREADS) {  // B independent pixel ops (gathers) in flight per thread
      uint32_t px[B], o[B];
#pragma unroll
      for (int k = 0; k < B; k++) px[k] = (j0 + k * THREADS < npx) ? tile[j0 + k * THREADS] : 0u;
#pragma unroll
      for (int k = 0; k < B; k++) o[k] = op(px[k]);
#pragma unroll
      for (int k = 0; k < B; k++)
        if (j0 + k * THREADS < npx) tile[j0 + k * THREADS] = o[k];
    }
    tma::fence_proxy_async();  // my smem writes -> visible to the bulk store
    __syncthreads();
    if (tid == 0) {
      tma::bulk_store(dst + doff, tile, nbytes, pol_stream);
      tma::bulk_commit();
      // the stage used one iteration ago is free once its store has finished reading shared memory
      tma::bulk_wait_read<1>();
      const long long nxt = i + STAGES - 1;
      if (nxt < mine) issue_load(nxt);
    }
  }
  if (tid == 0) tma::bulk_wait_all<0>();
}

struct MemoGatherOp {  // out = memo[px & 0xFFFFFF] | alpha
  const uint32_t *memo;
  uint64_t policy;  // 0 = plain read-only load, else an L2 eviction-priority policy (evict_last)
  __device__ __forceinline__ uint32_t operator()(uint32_t px) const {
    const uint32_t *p = memo + memo_index(px & 0x00FFFFFFu);
    const uint32_t v = policy ? tma::ldg_hint_u32(p, policy) : __ldg(p);
    return v | (px & 0xFF000000u);
  }
};

template <int TILE, int STAGES, int THREADS, int B>
__global__ void __launch_bounds__(THREADS) colorlut_memo_stream_kernel(const uint32_t *__restrict__ memo, int use_hint,
                                                                      const uint8_t *__restrict__ src, long sstride,
                                                                      uint8_t *__restrict__ dst, long dstride,
                                                                      int row_bytes, int height) {
  MemoGatherOp op{memo, use_hint ? tma::policy_evict_last() : 0ull};
  stream_map_u32<TILE, STAGES, THREADS, B>(op, src, sstride, dst, dstride, row_bytes, height);
}

// any 4-byte -> 4-byte pixel op (HsvFilterMemoOp, HsvDetectBitmapOp of kernels.cuh) through the same skeleton: used for
// the zero-copy host path of hsvfilter / hsvdetector (tiles bulk-loaded from and bulk-stored to pinned host memory)
template <int TILE, int STAGES, int THREADS, int B, typename Op>
__global__ void __launch_bounds__(THREADS) map_stream_kernel(Op op, const uint8_t *__restrict__ src, long sstride,
                                                            uint8_t *dst, long dstride, int row_bytes, int height) {
  stream_map_u32<TILE, STAGES, THREADS, B>(op, src, sstride, dst, dstride, row_bytes, height);
}

struct Memo1dOp {  // three 256-byte tables staged in shared memory
  const uint8_t *tab;
  __device__ __forceinline__ uint32_t operator()(uint32_t px) const {
    const uint32_t r = tab[px & 255u], g = tab[256 + ((px >> 8) & 255u)], b = tab[512 + ((px >> 16) & 255u)];
    return r | (g << 8) | (b << 16) | (px & 0xFF000000u);
  }
};

template <int TILE, int STAGES, int THREADS, int B>
__global__ void __launch_bounds__(THREADS) colorlut_memo1d_stream_kernel(const uint8_t *__restrict__ memo1d,
                                                                        const uint8_t *__restrict__ src, long sstride,
                                                                        uint8_t *__restrict__ dst, long dstride,
                                                                        int row_bytes, int height) {
  __shared__ uint8_t tab[768];
  for (int i = threadIdx.x; i < 768 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t *>(tab)[i] = __ldg(reinterpret_cast<const uint32_t *>(memo1d) + i);
  __syncthreads();
  Memo1dOp op{tab};
  stream_map_u32<TILE, STAGES, THREADS, B>(op, src, sstride, dst, dstride, row_bytes, height);
}

}  // namespace b200vfx
