// colordetect.cuh -- the pixel pass of `colordetect`.
//
// Reference: ColorDetect::detect_color (video/videofx/src/colordetect/imp.rs:57-86) hands plane 0 to
// color_thief::get_palette (color-thief 0.2.2, Cargo.lock:2045-2053), whose first -- and only per-pixel --
// step is a 5-bit-per-channel histogram over every `quality`-th pixel of the FLAT plane slice
// (stride padding included, pixels = len / bytes_per_pixel):
//     pos = i * bpp, i = 0, q, 2q, ...;   skip if a < 125 or (r > 250 && g > 250 && b > 250);
//     hist[(r>>3) << 10 | (g>>3) << 5 | (b>>3)] += 1
// The median cut over the 32768 bins is host work (colordetect_host.cpp).  Counts are integers, so the
// result does not depend on the order of the atomic adds: bit-exact against any CPU loop.
//
// B200 mapping: the whole histogram (32768 x u32 = 128 KB) fits the 228 KB shared memory of one SM, so the
// kernel is persistent with ONE 1024-thread CTA per SM; every CTA counts into its private shared-memory copy and
// the CTAs of a cluster add their copies through distributed shared memory before the non-zero sums go to the
// global histogram with atomics.  Runs of equal bins (flat areas, gradients: 8 consecutive values share a 5-bit
// bin) are merged in registers / with one ballot before the shared atomic, so coherent video issues a few
// atomics per warp instead of 32 same-address ones.
//   quality == 1, 4-byte pixels, 16-byte aligned plane: every thread streams 4 x uint4 (16 pixels, 64 KB in flight
//     per SM -- one CTA per SM has to cover the DRAM latency alone);
//   otherwise: lane-consecutive samples, one 4-byte (or 3 x 1-byte) load each.
// ONE launch per frame: after its pixel pass every CTA zeroes its slice of the global histogram and arrives at a grid-wide
// counter; the CTAs wait for that counter right before the global atomics (they finish the pixel pass together).
// (All CTAs of the grid are co-resident by construction -- the launcher caps the grid at the occupancy maximum -- so the
// spin cannot deadlock.)  The last CTA to retire resets the two counters for the next launch.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace b200vfx {

constexpr int kColorDetectBins = 32768;
constexpr int kColorDetectThreads = 1024;
constexpr uint32_t kCdInvalid = 0xFFFFFFFFu;

// FMT: 0 RGB, 1 RGBA, 2 ARGB, 3 BGR, 4 BGRA (color_parts() of color-thief).  A pixel is handled as a little-endian
// word of its bytes; 3-byte pixels get 0xFF in the top byte (alpha = 255).
template <int FMT> struct CdFmt;
template <> struct CdFmt<0> { static constexpr int bpp = 3, r = 0, g = 1, b = 2, a = 3; };
template <> struct CdFmt<1> { static constexpr int bpp = 4, r = 0, g = 1, b = 2, a = 3; };
template <> struct CdFmt<2> { static constexpr int bpp = 4, r = 1, g = 2, b = 3, a = 0; };
template <> struct CdFmt<3> { static constexpr int bpp = 3, r = 2, g = 1, b = 0, a = 3; };
template <> struct CdFmt<4> { static constexpr int bpp = 4, r = 2, g = 1, b = 0, a = 3; };

template <int FMT>
__device__ __forceinline__ uint32_t cd_bin(uint32_t px) {
  using F = CdFmt<FMT>;
  // per-byte unsigned compare: colour bytes > 250, alpha byte > 124 (i.e. a >= 125)
  constexpr uint32_t thr = (250u << (8 * F::r)) | (250u << (8 * F::g)) | (250u << (8 * F::b)) | (124u << (8 * F::a));
  constexpr uint32_t amask = 0xFFu << (8 * F::a), cmask = ~amask;
  const uint32_t m = __vcmpgtu4(px, thr);
  const bool counted = (m & amask) != 0u && (m & cmask) != cmask;
  const uint32_t bin = (((px >> (8 * F::r + 3)) & 31u) << 10) | (((px >> (8 * F::g + 3)) & 31u) << 5) | ((px >> (8 * F::b + 3)) & 31u);
  return counted ? bin : kCdInvalid;
}

// `key` identifies what a lane wants to count (`weight` times); equal keys in neighbouring lanes are merged
__device__ __forceinline__ void cd_count_run(uint32_t *sh_hist, uint32_t key, uint32_t bin, uint32_t weight, int lane) {
  const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, key, 1);
  const bool head = lane == 0 || key != prev;
  const uint32_t heads = __ballot_sync(0xFFFFFFFFu, head);
  if (head && bin != kCdInvalid) {
    const uint32_t above = heads & ~((2u << lane) - 1u);
    const int next = above ? (__ffs((int)above) - 1) : 32;
    atomicAdd(&sh_hist[bin], weight * (uint32_t)(next - lane));
  }
}

// MODE 0: byte loads (3-byte pixels, unaligned planes); 1: one aligned 32-bit load per sample; 2: step == 1, uint4 loads
template <int FMT, int MODE>
__global__ void __launch_bounds__(kColorDetectThreads, 1)
colordetect_hist_kernel(const uint8_t *__restrict__ plane, long long nsamples, int step, uint32_t *__restrict__ hist,
                        unsigned *__restrict__ gsync) {   // gsync[0]: CTAs that have zeroed their slice, gsync[1]: CTAs retired
  using F = CdFmt<FMT>;
  extern __shared__ __align__(16) uint32_t sh_hist[];
  pdl_trigger();   // the next launch may be scheduled; its CTAs take the SMs as ours retire
  for (int i = threadIdx.x; i < kColorDetectBins / 4; i += kColorDetectThreads)
    reinterpret_cast<uint4 *>(sh_hist)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int U = 4;
  if (MODE == 2) {
    // a "unit" is a uint4 = 4 consecutive pixels; a warp instruction covers 128 consecutive pixels
    const long long nunits = nsamples >> 2;                 // the tail (< 4 pixels) is counted by CTA 0 below
    const long long per_iter = (long long)kColorDetectThreads * U;
    for (long long base = (long long)blockIdx.x * per_iter; base < nunits; base += (long long)gridDim.x * per_iter) {
      const uint4 *p = reinterpret_cast<const uint4 *>(plane) + base;
      const int left = (int)min((long long)per_iter, nunits - base);
      const int i0 = warp * (32 * U) + lane;
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = (i0 + 32 * u < left) ? __ldcs(p + i0 + 32 * u) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int u = 0; u < U; u++) {
        const bool in = i0 + 32 * u < left;
        const uint32_t b0 = in ? cd_bin<FMT>(v[u].x) : kCdInvalid, b1 = in ? cd_bin<FMT>(v[u].y) : kCdInvalid;
        const uint32_t b2 = in ? cd_bin<FMT>(v[u].z) : kCdInvalid, b3 = in ? cd_bin<FMT>(v[u].w) : kCdInvalid;
        const bool uniform = b0 == b1 && b1 == b2 && b2 == b3;
        // uniform threads merge with their neighbours; mixed ones get a key nobody shares and count their own runs
        cd_count_run(sh_hist, uniform ? b0 : (0x40000000u | (uint32_t)lane), uniform ? b0 : kCdInvalid, 4u, lane);
        if (!uniform) {
          uint32_t cur = b0, cnt = 1u;
          if (b1 == cur) cnt++; else { if (cur != kCdInvalid) atomicAdd(&sh_hist[cur], cnt); cur = b1; cnt = 1u; }
          if (b2 == cur) cnt++; else { if (cur != kCdInvalid) atomicAdd(&sh_hist[cur], cnt); cur = b2; cnt = 1u; }
          if (b3 == cur) cnt++; else { if (cur != kCdInvalid) atomicAdd(&sh_hist[cur], cnt); cur = b3; cnt = 1u; }
          if (cur != kCdInvalid) atomicAdd(&sh_hist[cur], cnt);
        }
      }
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(nsamples & 3)) {
      const uint32_t b = cd_bin<FMT>(reinterpret_cast<const uint32_t *>(plane)[(nunits << 2) + threadIdx.x]);
      if (b != kCdInvalid) atomicAdd(&sh_hist[b], 1u);
    }
  } else {
    const long long per_iter = (long long)kColorDetectThreads * U;
    const int pix_stride = step * F::bpp;
    for (long long base = (long long)blockIdx.x * per_iter; base < nsamples; base += (long long)gridDim.x * per_iter) {
      const uint8_t *p = plane + base * pix_stride;
      const int left = (int)min((long long)per_iter, nsamples - base);
      const int i0 = warp * (32 * U) + lane;   // lane L takes samples L, L+32, ... of the warp's 32*U
      uint32_t px[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int i = i0 + 32 * u;
        px[u] = 0u;
        if (i < left) {
          const uint8_t *q = p + i * pix_stride;
          if (MODE == 1) px[u] = __ldcs(reinterpret_cast<const uint32_t *>(q));
          else {
            px[u] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16);
            px[u] |= (F::bpp == 4) ? ((uint32_t)q[3] << 24) : 0xFF000000u;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const uint32_t bin = (i0 + 32 * u < left) ? cd_bin<FMT>(px[u]) : kCdInvalid;
        cd_count_run(sh_hist, bin, bin, 1u, lane);
      }
    }
  }
  // cluster-wide merge: CTA `rank` sums bins [rank*per, (rank+1)*per) over the shared-memory copies of all CTAs of the cluster
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned csize = cluster.num_blocks(), crank = cluster.block_rank();
  // everything up to here only read the plane and wrote shared memory: with programmatic dependent launch it ran while the
  // previous launch on the stream drained.  The global histogram and the counters are touched after that launch completed.
  pdl_wait_prior();
  {  // this CTA's slice of the global histogram
    const int per = (kColorDetectBins + (int)gridDim.x - 1) / (int)gridDim.x;
    const int lo = (int)blockIdx.x * per, hi = min(kColorDetectBins, lo + per);
    for (int i = lo + (int)threadIdx.x; i < hi; i += kColorDetectThreads) hist[i] = 0u;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(gsync, 1u);
  cluster.sync();
  if (threadIdx.x == 0) {   // every CTA has zeroed its slice (the CTAs finish their pixel pass at about the same time)
    while (*(volatile unsigned *)gsync < gridDim.x) __nanosleep(32);
    __threadfence();
  }
  __syncthreads();
  // 4 consecutive bins per thread and step, read from every CTA of the cluster with ld.shared::cluster (mapa-translated
  // shared-window addresses: no generic-address loads); the reads of a step are independent of each other
  const int per4 = kColorDetectBins / 4 / (int)csize;
  const uint32_t sh_base = (uint32_t)__cvta_generic_to_shared(sh_hist);
  for (int i = threadIdx.x; i < per4; i += kColorDetectThreads) {
    const int u = (int)crank * per4 + i;
    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 2
    for (unsigned r = 0; r < csize; r++) {
      uint32_t ra;
      uint4 t;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(sh_base + 16u * (uint32_t)u), "r"(r));
      asm volatile("ld.shared::cluster.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(ra) : "memory");
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    if (acc.x) atomicAdd(hist + 4 * u, acc.x);
    if (acc.y) atomicAdd(hist + 4 * u + 1, acc.y);
    if (acc.z) atomicAdd(hist + 4 * u + 2, acc.z);
    if (acc.w) atomicAdd(hist + 4 * u + 3, acc.w);
  }
  cluster.sync();  // nobody leaves while a peer may still read its copy
  if (threadIdx.x == 0 && atomicAdd(gsync + 1, 1u) == gridDim.x - 1u) { gsync[0] = 0u; gsync[1] = 0u; }   // last CTA: re-arm
}

}  // namespace b200vfx
