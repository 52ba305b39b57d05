// colorlut on a row tile, fused with the all-gather of the tiles (BASELINE config 5).
//
// Row-tiled colorlut needs no exchange for a host sink.  A DEVICE-side consumer that wants the whole frame on every GPU
// would otherwise run the tile kernel and then an in-place ncclAllGather: the tile is written to local HBM, read back by
// the collective kernel and pushed over NVLink.  Here the map kernel itself stores every result vector into the frame
// buffer of every GPU (its own + the peers', mapped through CUDA IPC / peer access), so the tile crosses HBM once and the
// NVLink traffic overlaps the table gathers.  Frames are independent, the only synchronisation is per call:
//
//   entry   rank r tells every peer "my frame buffer may be overwritten for epoch e" (READY[r] = e in the peer's flag
//           block); a CTA stores to a peer only after it has seen that peer's READY -- a peer whose stream is still
//           reading the previous frame out of the same buffer has not launched its kernel yet.
//   exit    the last CTA to retire (device-scope counter) publishes DONE[r] = e to every peer after a system-scope
//           fence and then waits for every peer's DONE: when the kernel completes on rank r's stream, all N tiles of
//           frame e are visible in rank r's frame buffer -- the same post-condition as the in-place all-gather.
//
// Every wait has a deadline on %globaltimer: a missing peer turns into an error word in the flag block (reported by
// b200vfx_peer_status), never into a hung GPU.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "tma_pipe.cuh"

namespace b200vfx {

constexpr int kMaxPeers = 16;
enum : int { PF_READY = 0, PF_DONE = 16, PF_COUNT = 32, PF_ERR = 33, PF_WORDS = 64 };

struct PeerSet {
  uint8_t *frame[kMaxPeers];    // whole-frame buffer of every rank, as addressable from THIS device
  uint32_t *flags[kMaxPeers];   // PF_WORDS u32 per rank
  uint8_t *mc;                  // optional: ONE NVSwitch multicast address bound to the frame buffer of every rank
  int world, rank;
  uint32_t epoch;               // > 0, the same on every rank for one frame, increasing
  uint32_t timeout_ms;
};

__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// flag >= epoch (wrap-safe); false on deadline.  Relaxed polling: nothing is READ from a peer after the wait (what
// follows are stores, which cannot be speculated, or the end of the kernel), so no acquire -- an acquire load costs an
// L1 invalidation (CCTL.IVALL) per poll and the table gathers live in L1.
__device__ __forceinline__ bool wait_epoch(const uint32_t *p, uint32_t epoch, uint64_t deadline) {
  while ((int32_t)(ld_relaxed_sys(p) - epoch) < 0) {
    if (globaltimer_ns() > deadline) return false;
    __nanosleep(32);
  }
  return true;
}

__device__ __forceinline__ void st_v4(uint8_t *p, uint4 v) {
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// store through a multicast mapping: the switch replicates the 16 bytes into the bound buffer of every GPU
__device__ __forceinline__ void st_mc_v4(uint8_t *p, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_v8(uint8_t *p, uint4 a, uint4 b) {   // STG.E.ENL2.256: 1 KB per warp instruction
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void st_u32(uint8_t *p, uint32_t v) {
  asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// entry handshake (all threads of every CTA call it; returns false when a peer never showed up)
__device__ __forceinline__ bool peer_enter(const PeerSet &ps, uint64_t deadline, int *s_ok) {
  uint32_t *mine = ps.flags[ps.rank];
  const int t = threadIdx.x;
  if (t == 0) *s_ok = 1;
  if (blockIdx.x == 0 && blockIdx.y == 0 && t < ps.world && t != ps.rank)
    st_release_sys(ps.flags[t] + PF_READY + ps.rank, ps.epoch);
  __syncthreads();
  if (t < ps.world && t != ps.rank && !wait_epoch(mine + PF_READY + t, ps.epoch, deadline)) {
    *s_ok = 0;
    atomicExch(mine + PF_ERR, ps.epoch);
  }
  __syncthreads();
  return *s_ok != 0;
}

// exit (all threads of every CTA call it after the CTA's last store has been issued; for bulk-async stores the issuing
// thread must have waited for their completion): one system-scope fence per CTA, cumulative over the CTA's stores
// through the barrier, then the CTA is counted; the last CTA publishes DONE and waits for the peers' DONE.
__device__ __forceinline__ void peer_exit(const PeerSet &ps, uint64_t deadline, int *s_last) {
  uint32_t *mine = ps.flags[ps.rank];
  const int t = threadIdx.x;
  __syncthreads();
  if (t == 0) {
    fence_acq_rel_sys();
    const uint32_t total = gridDim.x * gridDim.y;
    *s_last = (atomicAdd(mine + PF_COUNT, 1u) == total - 1u);
    fence_acq_rel_sys();
  }
  __syncthreads();
  if (*s_last) {
    if (t == 0) mine[PF_COUNT] = 0u;   // the next launch on this stream starts from zero
    const bool ok = (*(volatile uint32_t *)(mine + PF_ERR)) != ps.epoch;
    if (t < ps.world && t != ps.rank) {
      if (ok) st_release_sys(ps.flags[t] + PF_DONE + ps.rank, ps.epoch);
      if (!wait_epoch(mine + PF_DONE + t, ps.epoch, deadline)) atomicExch(mine + PF_ERR, ps.epoch);
      // the polls are relaxed (an acquire per poll would invalidate L1 under the table gathers); ONE acquire fence
      // after the last wait, just before the kernel ends, makes the post-condition formal: everything the peers
      // stored before publishing DONE happens-before whatever follows this kernel on the stream
      fence_acq_rel_sys();
    }
  }
}

// ---- variant 0: register path (LDG / STG) -------------------------------------------------------------------------
// Same thread mapping as colorlut_memo_apply_kernel (a warp owns 32*PX consecutive pixels, PX coherent gathers in
// flight per thread), persistent CTAs looping over (strip, row).  VEC: results are transposed through a warp-private
// shared-memory strip so that every lane stores 16 consecutive bytes -- 512-byte warp stores, 4x fewer store
// instructions per destination, full NVLink packets.
template <int PX, int VEC, bool LUT1D>   // VEC: 0 = 4-byte stores, 1 = 16-byte, 2 = 32-byte (PX == 8, 32-byte aligned rows)
__global__ void __launch_bounds__(256) colorlut_tile_gather_kernel(const uint32_t *__restrict__ memo,
                                                                   const uint8_t *__restrict__ memo1d,
                                                                   const uint8_t *__restrict__ src, long sstride,
                                                                   PeerSet ps, long dstride, long dst_offset,
                                                                   int width, int rows) {
  __shared__ __align__(16) uint32_t strip[8][32 * PX];
  __shared__ uint8_t tab[LUT1D ? 768 : 4];
  __shared__ int s_ok, s_last;
  const uint64_t deadline = globaltimer_ns() + (uint64_t)ps.timeout_ms * 1000000ull;
  const int t = threadIdx.x;
  if (LUT1D)
    for (int i = t; i < 768 / 4; i += blockDim.x)
      reinterpret_cast<uint32_t *>(tab)[i] = __ldg(reinterpret_cast<const uint32_t *>(memo1d) + i);
  if (t == 0) s_last = 0;
  const bool go = peer_enter(ps, deadline, &s_ok);

  const int lane = t & 31, warp = t >> 5;
  const int cx = (width + 8 * 32 * PX - 1) / (8 * 32 * PX);   // CTA-wide strips per row
  const long long nwork = (long long)cx * rows;
  uint32_t *xp = strip[warp];
  for (long long c = blockIdx.x; go && c < nwork; c += gridDim.x) {
    int row, ci;
    item_row_chunk(c, cx, rows, row, ci);
    const int xw = (ci * 8 + warp) * (32 * PX);   // first pixel of this warp's strip
    if (xw >= width) continue;
    const uint32_t *s = reinterpret_cast<const uint32_t *>(src + (size_t)row * sstride);
    const size_t drow = (size_t)dst_offset + (size_t)row * dstride;
    uint32_t px[PX], o[PX];
#pragma unroll
    for (int k = 0; k < PX; k++) px[k] = (xw + lane + 32 * k < width) ? ld_stream_u32(s + xw + lane + 32 * k) : 0u;
#pragma unroll
    for (int k = 0; k < PX; k++) {
      if (LUT1D) {
        const uint32_t r = tab[px[k] & 255u], g = tab[256 + ((px[k] >> 8) & 255u)], b = tab[512 + ((px[k] >> 16) & 255u)];
        o[k] = r | (g << 8) | (b << 16) | (px[k] & 0xFF000000u);
      } else {
        o[k] = __ldg(memo + memo_index(px[k] & 0x00FFFFFFu)) | (px[k] & 0xFF000000u);
      }
    }
    if (VEC == 2) {
      __syncwarp();
#pragma unroll
      for (int k = 0; k < PX; k++) xp[32 * k + lane] = o[k];
      __syncwarp();
      const uint4 a = *reinterpret_cast<const uint4 *>(xp + 8 * lane), b = *reinterpret_cast<const uint4 *>(xp + 8 * lane + 4);
      const int xq = xw + 8 * lane;
      if (xq < width)   // width % 8 == 0 on this path
        for (int i = 0; i < ps.world; i++) {
          int p = ps.rank + i;
          if (p >= ps.world) p -= ps.world;
          st_v8(ps.frame[p] + drow + 4 * (size_t)xq, a, b);
        }
    } else if (VEC) {
      __syncwarp();
#pragma unroll
      for (int k = 0; k < PX; k++) xp[32 * k + lane] = o[k];
      __syncwarp();
      uint4 v[PX / 4];
#pragma unroll
      for (int j = 0; j < PX / 4; j++) v[j] = *reinterpret_cast<const uint4 *>(xp + 4 * (lane + 32 * j));
      if (ps.mc) {   // one store leaves the GPU, the switch fans it out: egress = the tile, not (world - 1) x the tile
        uint8_t *d = ps.mc + drow;
#pragma unroll
        for (int j = 0; j < PX / 4; j++) {
          const int xq = xw + 4 * (lane + 32 * j);
          if (xq < width) st_mc_v4(d + 4 * (size_t)xq, v[j]);
        }
      } else
      for (int i = 0; i < ps.world; i++) {
        int p = ps.rank + i;   // own frame first, then the peers starting at the right-hand neighbour:
        if (p >= ps.world) p -= ps.world;   // at any instant the ranks push towards different destinations
        uint8_t *d = ps.frame[p] + drow;
#pragma unroll
        for (int j = 0; j < PX / 4; j++) {
          const int xq = xw + 4 * (lane + 32 * j);
          if (xq < width) st_v4(d + 4 * (size_t)xq, v[j]);   // width % 4 == 0 on this path
        }
      }
    } else {
      for (int i = 0; i < ps.world; i++) {
        int p = ps.rank + i;
        if (p >= ps.world) p -= ps.world;
        uint8_t *d = ps.frame[p] + drow;
#pragma unroll
        for (int k = 0; k < PX; k++)
          if (xw + lane + 32 * k < width) st_u32(d + 4 * (size_t)(xw + lane + 32 * k), o[k]);
      }
    }
  }
  peer_exit(ps, deadline, &s_last);
}

// ---- variant 1: TMA path ------------------------------------------------------------------------------------------
// The streaming skeleton of stream_map.cuh (bulk load HBM -> smem ring, transform in place) with ONE bulk store per
// destination GPU out of the same shared-memory tile: one thread moves TILE bytes per instruction to every peer, the
// SMs' load/store units only see the table gathers.  Rows 16-byte aligned, row_bytes % 16 == 0.
template <int TILE, int STAGES, int THREADS, int B, bool LUT1D>
__global__ void __launch_bounds__(THREADS) colorlut_tile_gather_tma_kernel(const uint32_t *__restrict__ memo,
                                                                          const uint8_t *__restrict__ memo1d,
                                                                          const uint8_t *__restrict__ src, long sstride,
                                                                          PeerSet ps, long dstride, long dst_offset,
                                                                          int row_bytes, int rows) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + STAGES * TILE);
  __shared__ uint8_t tab[LUT1D ? 768 : 4];
  __shared__ int s_ok, s_last;
  const uint64_t deadline = globaltimer_ns() + (uint64_t)ps.timeout_ms * 1000000ull;
  const int tid = threadIdx.x;
  const int tiles_per_row = (row_bytes + TILE - 1) / TILE;
  const long long ntiles = (long long)tiles_per_row * rows;
  const long long first = blockIdx.x, step = gridDim.x;
  long long mine = first < ntiles ? (ntiles - first + step - 1) / step : 0;  // tiles this CTA owns
  uint64_t pol_stream = 0;
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) tma::mbar_init(&bars[s], 1);
    tma::fence_barrier_init();
    pol_stream = tma::policy_evict_first();
    s_last = 0;
  }
  if (LUT1D)
    for (int i = tid; i < 768 / 4; i += THREADS)
      reinterpret_cast<uint32_t *>(tab)[i] = __ldg(reinterpret_cast<const uint32_t *>(memo1d) + i);
  __syncthreads();
  auto tile_geom = [&](long long i, size_t &soff, size_t &doff, uint32_t &nbytes) {
    const long long t = first + i * step;
    const int row = (int)(t / tiles_per_row), c = (int)(t % tiles_per_row);
    const int off = c * TILE;
    nbytes = (uint32_t)min(TILE, row_bytes - off);
    soff = (size_t)row * sstride + off;
    doff = (size_t)dst_offset + (size_t)row * dstride + off;
  };
  auto issue_load = [&](long long i) {  // thread 0 only
    size_t soff, doff; uint32_t nbytes;
    tile_geom(i, soff, doff, nbytes);
    const int s = (int)(i % STAGES);
    tma::mbar_expect_tx(&bars[s], nbytes);
    tma::bulk_load(smem_raw + s * TILE, src + soff, nbytes, &bars[s], pol_stream);
  };
  // the loads do not depend on the peers: start them before the handshake
  if (tid == 0)
    for (long long i = 0; i < mine && i < STAGES - 1; i++) issue_load(i);
  const bool go = peer_enter(ps, deadline, &s_ok);
  for (long long i = 0; i < mine; i++) {
    const int s = (int)(i % STAGES);
    size_t soff, doff; uint32_t nbytes;
    tile_geom(i, soff, doff, nbytes);
    tma::mbar_wait(&bars[s], (uint32_t)((i / STAGES) & 1));
    uint32_t *tile = reinterpret_cast<uint32_t *>(smem_raw + s * TILE);
    const int npx = (int)(nbytes >> 2);
    for (int j0 = tid; j0 < npx; j0 += B * THREADS) {
      uint32_t px[B], o[B];
#pragma unroll
      for (int k = 0; k < B; k++) px[k] = (j0 + k * THREADS < npx) ? tile[j0 + k * THREADS] : 0u;
#pragma unroll
      for (int k = 0; k < B; k++) {
        if (LUT1D) {
          const uint32_t r = tab[px[k] & 255u], g = tab[256 + ((px[k] >> 8) & 255u)], b = tab[512 + ((px[k] >> 16) & 255u)];
          o[k] = r | (g << 8) | (b << 16) | (px[k] & 0xFF000000u);
        } else {
          o[k] = __ldg(memo + memo_index(px[k] & 0x00FFFFFFu)) | (px[k] & 0xFF000000u);
        }
      }
#pragma unroll
      for (int k = 0; k < B; k++)
        if (j0 + k * THREADS < npx) tile[j0 + k * THREADS] = o[k];
    }
    tma::fence_proxy_async();  // my smem writes -> visible to the bulk stores
    __syncthreads();
    if (tid == 0) {
      if (go)
        for (int q = 0; q < ps.world; q++) {
          int p = ps.rank + q;
          if (p >= ps.world) p -= ps.world;
          tma::bulk_store(ps.frame[p] + doff, tile, nbytes, pol_stream);
        }
      tma::bulk_commit();
      tma::bulk_wait_read<1>();   // the stage used one iteration ago is free once its stores have read shared memory
      const long long nxt = i + STAGES - 1;
      if (nxt < mine) issue_load(nxt);
    }
  }
  if (tid == 0) tma::bulk_wait_all<0>();   // completion of the writes, not just of the shared-memory reads
  peer_exit(ps, deadline, &s_last);
}

}  // namespace b200vfx
