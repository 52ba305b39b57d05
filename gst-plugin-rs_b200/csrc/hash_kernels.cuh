// hash_kernels.cuh -- videocompare beyond the integer blockhash fast path (hashed_image.rs:24-106 -> image_hasher 3.1.1,
// image 0.25.10; third-party arithmetic restated as recalled, PARITY UNPINNED -- see oracle/vfx_oracle_hash.c):
//
//   * Mean / Gradient / VertGradient / DoubleGradient: image::imageops::grayscale + imageops::resize(Lanczos3) to
//     8x8 / 9x8 / 8x9 / 5x5.  A 4K frame shrinks by 270-480x, so every output sample is a weighted sum of ~1600 (vertical)
//     resp. ~2900 (horizontal) source samples accumulated IN SOURCE ORDER in f32 (t += v * w, unfused): the serial chain
//     of additions per output sample is part of the result and is kept; the products are formed in parallel.  The normalised tap weights are computed on the host (they
//     need libm's sinf, which the device's sinf does not reproduce bit for bit) and uploaded once per frame size.
//   * Blockhash for frames whose size is not a multiple of the hash grid: block index by the reference's f32 division.
//     The reference accumulates the block sums in f32 in raster order; while a block's sum stays below 2^24 that is exact
//     integer arithmetic (any order), beyond it every addition rounds and the order matters -> a sequential kernel
//     (one warp per block) reproduces the raster-order chain.
#pragma once
#include "kernels.cuh"

namespace b200vfx {

__device__ __forceinline__ unsigned luma_of(unsigned r, unsigned g, unsigned b) { return (2126u * r + 7152u * g + 722u * b) / 10000u; }

// 0..255 -> f32 without the conversion pipe: 0x4B000000 | l is the float 2^23 + l, the subtraction is exact
__device__ __forceinline__ float small_uint_as_float(unsigned l) { return __fsub_rn(__uint_as_float(0x4B000000u | l), 8388608.0f); }

// vertical pass: tmp[oy][x] = sum_i luma(x, left[oy] + i) * w[oy][i], the additions IN ORDER.  taps: [nh][max_taps] floats,
// meta[oy] = {left, n}.  Only the chain of additions is serial: the products luma * w are rounded one by one and do not
// depend on each other.  A CTA owns 64 columns of one output row; all 256 threads fetch a 64-row x 64-column block of
// the window (next block already in flight), turn it into products in shared memory, then 64 threads run the 64
// additions of their column in source order.  (Round 2's first version gave every chain its own thread: ~1.6 warps per
// scheduler, latency-bound at 100-240 us per 4K frame.)
constexpr int kVresCols = 64, kVresRows = 64, kVresThreads = 256, kVresPer = kVresCols * kVresRows / kVresThreads;
// (2126 r + 7152 g + 722 b) / 10000 of a pixel word r | g << 8 | b << 16 in five instructions: the weights split into
// bytes (2126 = 8 * 256 + 78, 7152 = 27 * 256 + 240, 722 = 2 * 256 + 210) for two dp4a, the division as a multiply-high
__device__ __forceinline__ unsigned luma_of_word(uint32_t px) {
  const unsigned s = __dp4a(px, 0x00021B08u, 0u) * 256u + __dp4a(px, 0x00D2F04Eu, 0u);   // <= 2 550 000
  return __umulhi(s, 0xD1B71759u) >> 13;                                                  // s / 10000, exact for 32-bit s
}
template <int BPP, bool AL4>
__global__ void __launch_bounds__(kVresThreads) luma_vresize_kernel(const uint8_t *__restrict__ src, long stride, int width,
                                                                   const float *__restrict__ taps, const int2 *__restrict__ meta,
                                                                   int max_taps, float *__restrict__ tmp) {
  __shared__ float prod[kVresRows][kVresCols];
  // thread (c, r0) handles column c of rows r0 * 16 .. r0 * 16 + 15 of every 64-row block
  const int oy = blockIdx.y, c = threadIdx.x & (kVresCols - 1), r0 = (threadIdx.x / kVresCols) * kVresPer;
  const int x = blockIdx.x * kVresCols + c;
  const bool xin = x < width;
  const int2 m = meta[oy];
  const float *w = taps + (size_t)oy * max_taps + r0;
  const uint8_t *p = src + (size_t)(m.x + r0) * stride + (size_t)(xin ? x : 0) * BPP;
  const size_t block_step = (size_t)kVresRows * stride;
  uint32_t raw[kVresPer];
  auto load_px = [&](const uint8_t *q) -> uint32_t {
    if (BPP == 4 && AL4) return __ldg(reinterpret_cast<const uint32_t *>(q));
    return (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16);
  };
  auto fetch = [&](int base) {   // rows base + r0 + k; p already points at row base + r0
    const uint8_t *q = p;
    if (xin && base + kVresRows <= m.y) {
#pragma unroll
      for (int k = 0; k < kVresPer; k++, q += stride) raw[k] = load_px(q);
    } else {
#pragma unroll
      for (int k = 0; k < kVresPer; k++, q += stride) raw[k] = (xin && base + r0 + k < m.y) ? load_px(q) : 0u;
    }
    p += block_step;
  };
  float t = 0.0f;
  fetch(0);
  for (int base = 0; base < m.y; base += kVresRows) {
    const float *wb = w + base;
    if (base + kVresRows <= m.y) {
#pragma unroll
      for (int k = 0; k < kVresPer; k++)
        prod[r0 + k][c] = __fmul_rn(small_uint_as_float(luma_of_word(raw[k] & 0x00FFFFFFu)), __ldg(wb + k));
    } else {
#pragma unroll
      for (int k = 0; k < kVresPer; k++)
        prod[r0 + k][c] = __fmul_rn(small_uint_as_float(luma_of_word(raw[k] & 0x00FFFFFFu)), base + r0 + k < m.y ? __ldg(wb + k) : 0.0f);
    }
    __syncthreads();
    if (base + kVresRows < m.y) fetch(base + kVresRows);   // in flight while the chains below run
    if (threadIdx.x < kVresCols) {
      const int n = min(kVresRows, m.y - base);
      int r = 0;
      for (; r + 8 <= n; r += 8) {
#pragma unroll
        for (int u = 0; u < 8; u++) t = __fadd_rn(t, prod[r + u][c]);
      }
      for (; r < n; r++) t = __fadd_rn(t, prod[r][c]);
    }
    __syncthreads();
  }
  if (threadIdx.x < kVresCols && xin) tmp[(size_t)oy * width + x] = t;
}

// horizontal pass on the f32 intermediate + clamp(0,255).round(): one CTA per output sample (at most 81 of them); all
// threads form the products of a block of taps in shared memory, one thread adds them up in source order
constexpr int kHresBlock = 4096;
__global__ void __launch_bounds__(256) luma_hresize_kernel(const float *__restrict__ tmp, int width, int nw, int nh,
                                                          const float *__restrict__ taps, const int2 *__restrict__ meta,
                                                          int max_taps, uint8_t *__restrict__ out) {
  __shared__ float prod[kHresBlock];
  const int o = blockIdx.x;
  const int oy = o / nw, ox = o - oy * nw;
  const int2 m = meta[ox];
  const float *w = taps + (size_t)ox * max_taps, *row = tmp + (size_t)oy * width + m.x;
  float t = 0.0f;
  for (int base = 0; base < m.y; base += kHresBlock) {
    const int n = min(kHresBlock, m.y - base);
    for (int j = threadIdx.x; j < n; j += blockDim.x) prod[j] = __fmul_rn(__ldg(row + base + j), __ldg(w + base + j));
    __syncthreads();
    if (threadIdx.x == 0) {
      int j = 0;
      for (; j + 16 <= n; j += 16) {
#pragma unroll
        for (int u = 0; u < 16; u++) t = __fadd_rn(t, prod[j + u]);
      }
      for (; j < n; j++) t = __fadd_rn(t, prod[j]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    t = t < 0.0f ? 0.0f : (t > 255.0f ? 255.0f : t);
    out[o] = (uint8_t)roundf(t);
  }
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
