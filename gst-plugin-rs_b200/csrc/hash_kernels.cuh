// hash_kernels.cuh -- videocompare beyond the integer blockhash fast path (hashed_image.rs:24-106 -> image_hasher 3.1.1,
// image 0.25.10; third-party arithmetic restated as recalled, PARITY UNPINNED -- see oracle/vfx_oracle_hash.c):
//
//   * Mean / Gradient / VertGradient / DoubleGradient: image::imageops::grayscale + imageops::resize(Lanczos3) to
//     8x8 / 9x8 / 8x9 / 5x5.  A 4K frame shrinks by 270-480x, so every output sample is a weighted sum of ~1600 (vertical)
//     resp. ~2900 (horizontal) source samples accumulated IN SOURCE ORDER in f32 (t += v * w, unfused): the serial chain
//     per output sample is part of the result and is kept; the parallelism is across the W x new_h (vertical pass) and
//     new_w x new_h (horizontal pass) independent chains.  The normalised tap weights are computed on the host (they
//     need libm's sinf, which the device's sinf does not reproduce bit for bit) and uploaded once per frame size.
//   * Blockhash for frames whose size is not a multiple of the hash grid: block index by the reference's f32 division.
//     The reference accumulates the block sums in f32 in raster order; while a block's sum stays below 2^24 that is exact
//     integer arithmetic (any order), beyond it every addition rounds and the order matters -> a sequential kernel
//     (one warp per block) reproduces the raster-order chain.
#pragma once
#include "kernels.cuh"

namespace b200vfx {

__device__ __forceinline__ unsigned luma_of(unsigned r, unsigned g, unsigned b) { return (2126u * r + 7152u * g + 722u * b) / 10000u; }

// vertical pass: tmp[oy][x] = sum_i luma(x, left[oy] + i) * w[oy][i].  taps: [nh][max_taps] floats, meta[oy] = {left, n}.
template <int BPP>
__global__ void __launch_bounds__(128) luma_vresize_kernel(const uint8_t *__restrict__ src, long stride, int width,
                                                          const float *__restrict__ taps, const int2 *__restrict__ meta,
                                                          int max_taps, float *__restrict__ tmp) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
  if (x >= width) return;
  const int2 m = meta[oy];
  const float *w = taps + (size_t)oy * max_taps;
  const uint8_t *p = src + (size_t)m.x * stride + (size_t)x * BPP;
  float t = 0.0f;
  // U independent loads in flight, and the NEXT batch is requested before the current one is accumulated (there are only
  // W x new_h threads -- ~1.6 warps per scheduler on a 4K frame -- so memory latency is hidden inside a thread or not at
  // all); the accumulation itself stays strictly in source order
  constexpr int U = 16;
  auto fetch = [&](int i0, uint32_t (&raw)[U], float (&wt)[U]) {
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint8_t *q = p + (size_t)(i0 + u) * stride;
      if (BPP == 4) raw[u] = __ldg(reinterpret_cast<const uint32_t *>(q));
      else raw[u] = (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16);
      wt[u] = __ldg(w + i0 + u);
    }
  };
  int i = 0;
  uint32_t cur[U], nxt[U];
  float cw[U], nw[U];
  if (U <= m.y) fetch(0, cur, cw);
  for (; i + U <= m.y; i += U) {
    const bool more = i + 2 * U <= m.y;
    if (more) fetch(i + U, nxt, nw);
#pragma unroll
    for (int u = 0; u < U; u++) {
      const unsigned l = luma_of(cur[u] & 255u, (cur[u] >> 8) & 255u, (cur[u] >> 16) & 255u);
      t = __fadd_rn(t, __fmul_rn((float)l, cw[u]));
    }
    if (more) {
#pragma unroll
      for (int u = 0; u < U; u++) { cur[u] = nxt[u]; cw[u] = nw[u]; }
    }
  }
  for (; i < m.y; i++) {
    const uint8_t *q = p + (size_t)i * stride;
    const unsigned l = (BPP == 4) ? luma_of(q[0], q[1], q[2]) : luma_of(q[0], q[1], q[2]);
    t = __fadd_rn(t, __fmul_rn((float)l, __ldg(w + i)));
  }
  tmp[(size_t)oy * width + x] = t;
}
// BPP == 4 rows that are not 4-byte aligned take the byte loads of the tail loop everywhere
template <int BPP>
__global__ void __launch_bounds__(128) luma_vresize_bytes_kernel(const uint8_t *__restrict__ src, long stride, int width,
                                                                const float *__restrict__ taps, const int2 *__restrict__ meta,
                                                                int max_taps, float *__restrict__ tmp) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
  if (x >= width) return;
  const int2 m = meta[oy];
  const float *w = taps + (size_t)oy * max_taps;
  const uint8_t *p = src + (size_t)m.x * stride + (size_t)x * BPP;
  float t = 0.0f;
  for (int i = 0; i < m.y; i++) {
    const uint8_t *q = p + (size_t)i * stride;
    t = __fadd_rn(t, __fmul_rn((float)luma_of(q[0], q[1], q[2]), __ldg(w + i)));
  }
  tmp[(size_t)oy * width + x] = t;
}

// horizontal pass on the f32 intermediate + clamp(0,255).round(): one thread per output sample (at most 81 of them)
__global__ void luma_hresize_kernel(const float *__restrict__ tmp, int width, int nw, int nh, const float *__restrict__ taps,
                                    const int2 *__restrict__ meta, int max_taps, uint8_t *__restrict__ out) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= nw * nh) return;
  const int oy = o / nw, ox = o - oy * nw;
  const int2 m = meta[ox];
  const float *w = taps + (size_t)ox * max_taps, *row = tmp + (size_t)oy * width + m.x;
  float t = 0.0f;
  constexpr int U = 16;   // operands of the next 16 steps are loaded before the in-order accumulation of the current 16
  int i = 0;
  for (; i + U <= m.y; i += U) {
    float a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; u++) { a[u] = __ldg(row + i + u); b[u] = __ldg(w + i + u); }
#pragma unroll
    for (int u = 0; u < U; u++) t = __fadd_rn(t, __fmul_rn(a[u], b[u]));
  }
  for (; i < m.y; i++) t = __fadd_rn(t, __fmul_rn(__ldg(row + i), __ldg(w + i)));
  t = t < 0.0f ? 0.0f : (t > 255.0f ? 255.0f : t);
  out[o] = (uint8_t)roundf(t);
}
// nothing to resize (frame already new_w x new_h): grayscale only
template <int BPP>
__global__ void luma_copy_kernel(const uint8_t *__restrict__ src, long stride, int width, int height, uint8_t *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= width * height) return;
  const int y = i / width, x = i - y * width;
  const uint8_t *q = src + (size_t)y * stride + (size_t)x * BPP;
  out[i] = (uint8_t)luma_of(q[0], q[1], q[2]);
}

// ---- blockhash, frame size not a multiple of the hash grid -------------------------------------------------------------
__device__ __forceinline__ unsigned px_sum_of(const uint8_t *q, int bpp) {
  const unsigned s = (unsigned)q[0] + q[1] + q[2];
  return (bpp == 4 && q[3] == 0) ? 765u : s;
}
// block index of a coordinate: floor(c as f32 / block_size) with the IEEE division of the reference
__device__ __forceinline__ int block_of(int c, float block_size) { return (int)floorf(__fdiv_rn((float)c, block_size)); }

// exact integer sums (valid as the f32 result while every block sum < 2^24): one CTA per row chunk, run-merged shared atomics
template <int BPP>
__global__ void __launch_bounds__(256) blockhash_frac_kernel(const uint8_t *__restrict__ src, long stride, int width, int height,
                                                            int hw, int hh, float bwf, float bhf, int rows_per_cta,
                                                            uint32_t *__restrict__ sums) {
  extern __shared__ uint32_t bins[];   // hw * hh
  for (int i = threadIdx.x; i < hw * hh; i += blockDim.x) bins[i] = 0u;
  __syncthreads();
  const int y0 = blockIdx.x * rows_per_cta, y1 = min(height, y0 + rows_per_cta);
  for (int y = y0; y < y1; y++) {
    const int by = block_of(y, bhf);
    const uint8_t *row = src + (size_t)y * stride;
    // a thread owns 8 consecutive pixels: its run of equal block indices is merged before the atomic
    for (int xb = threadIdx.x * 8; xb < width; xb += blockDim.x * 8) {
      int cur = -1;
      uint32_t acc = 0;
      const int xe = min(width, xb + 8);
      for (int x = xb; x < xe; x++) {
        const int bx = block_of(x, bwf);
        if (bx != cur) {
          if (cur >= 0) atomicAdd(&bins[by * hw + cur], acc);
          cur = bx; acc = 0;
        }
        acc += px_sum_of(row + (size_t)x * BPP, BPP);
      }
      if (cur >= 0) atomicAdd(&bins[by * hw + cur], acc);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < hw * hh; i += blockDim.x)
    if (bins[i]) atomicAdd(&sums[i], bins[i]);
}

// the raster-order f32 chain of one hash block, for sums that leave the exact-integer range: one warp per block; lanes
// load 32 consecutive pixel sums, every lane then replays the 32 additions in order (shuffle broadcast).
template <int BPP>
__global__ void __launch_bounds__(32) blockhash_seq_kernel(const uint8_t *__restrict__ src, long stride, int width, int height,
                                                          int hw, int hh, float bwf, float bhf, float *__restrict__ sums) {
  const int b = blockIdx.x, by = b / hw, bx = b - by * hw, lane = threadIdx.x;
  // the pixel range of block (bx, by): all x with block_of(x) == bx -- contiguous; found by scanning around the estimate
  auto range = [](int idx, float bs, int limit, int &lo, int &hi) {
    int g = (int)((float)idx * bs);
    g = max(0, min(limit - 1, g));
    while (g > 0 && block_of(g - 1, bs) >= idx) g--;
    while (g < limit && block_of(g, bs) < idx) g++;
    lo = g;
    while (g < limit && block_of(g, bs) == idx) g++;
    hi = g;
  };
  int x_lo, x_hi, y_lo, y_hi;
  range(bx, bwf, width, x_lo, x_hi);
  range(by, bhf, height, y_lo, y_hi);
  float s = 0.0f;
  for (int y = y_lo; y < y_hi; y++) {
    const uint8_t *row = src + (size_t)y * stride;
    for (int xb = x_lo; xb < x_hi; xb += 32) {
      const int x = xb + lane;
      const float v = (x < x_hi) ? (float)px_sum_of(row + (size_t)x * BPP, BPP) : 0.0f;
      const int n = min(32, x_hi - xb);
      for (int j = 0; j < n; j++) s = __fadd_rn(s, __shfl_sync(0xFFFFFFFFu, v, j));
    }
  }
  if (lane == 0) sums[b] = s;
}

}  // namespace b200vfx
