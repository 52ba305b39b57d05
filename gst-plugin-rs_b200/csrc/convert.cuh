// convert.cuh -- the `videoconvert`-class kernels that sit either side of the hot-path elements in real pipelines
// (colorlut/imp.rs:18 wraps colorlut in two videoconverts; hsvdetector -> roundedcorners needs
// RGBA -> I420).  With these a pipeline stays in HBM from the first upload to the last download.
//
//   * packed <-> packed (RGBx xRGB BGRx xBGR RGBA ARGB BGRA ABGR RGB BGR): pure byte permutations -- exact by
//     definition.  A missing alpha/padding byte is written as 255 (what GStreamer's pack functions store).
//   * RGB-family -> I420 / A420 and back: GStreamer's converter is not part of the reference tree, so the arithmetic is
//     SPECIFIED HERE and parity with `videoconvert` is UNPINNED: 8-bit fixed point, coefficients rint(c * 256) of the
//     BT.601 (height <= 576) or BT.709 limited-range matrices, + 128 rounding, >> 8, clamp; chroma = rounded mean of
//     the 2x2 block's per-pixel chroma (edge pixels replicated); I420 -> RGB uses the co-sited (nearest) chroma sample.
//   * A420 = I420 + an A8 plane (roundedcorners): plane copies on the device.
#pragma once
#include "kernels.cuh"

namespace b200vfx {

struct PackedFmt { int bpp, r, g, b, a; };   // byte offsets inside a pixel; a < 0: no alpha (x byte at `x`, or none)

// 4-byte -> 4-byte: one PRMT per pixel.  sel picks bytes of {src pixel (0-3), 0xFFFFFFFF (4-7)}.
__global__ void __launch_bounds__(256) swizzle44_kernel(const uint8_t *__restrict__ src, long sstride, uint8_t *__restrict__ dst,
                                                        long dstride, int width, int height, uint32_t sel, int vec) {
  pdl_trigger();
  const int row = blockIdx.y;
  if (vec) {   // rows 16-byte aligned, width % 4 == 0
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i * 4 >= width) return;
    for (int y = row; y < height; y += gridDim.y) {
      const uint4 p = __ldcs(reinterpret_cast<const uint4 *>(src + (size_t)y * sstride) + i);
      uint4 o;
      o.x = __byte_perm(p.x, 0xFFFFFFFFu, sel); o.y = __byte_perm(p.y, 0xFFFFFFFFu, sel);
      o.z = __byte_perm(p.z, 0xFFFFFFFFu, sel); o.w = __byte_perm(p.w, 0xFFFFFFFFu, sel);
      __stcs(reinterpret_cast<uint4 *>(dst + (size_t)y * dstride) + i, o);
    }
  } else {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= width) return;
    for (int y = row; y < height; y += gridDim.y) {
      const uint8_t *s = src + (size_t)y * sstride + (size_t)x * 4;
      uint8_t *d = dst + (size_t)y * dstride + (size_t)x * 4;
      const uint32_t p = (uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16) | ((uint32_t)s[3] << 24);
      const uint32_t o = __byte_perm(p, 0xFFFFFFFFu, sel);
      d[0] = (uint8_t)o; d[1] = (uint8_t)(o >> 8); d[2] = (uint8_t)(o >> 16); d[3] = (uint8_t)(o >> 24);
    }
  }
}

// any packed -> any packed, byte addressed (3-byte formats, unaligned rows)
__global__ void __launch_bounds__(256) swizzle_generic_kernel(const uint8_t *__restrict__ src, long sstride, PackedFmt sf,
                                                              uint8_t *__restrict__ dst, long dstride, PackedFmt df, int width,
                                                              int height) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= width) return;
  for (int y = blockIdx.y; y < height; y += gridDim.y) {
    const uint8_t *s = src + (size_t)y * sstride + (size_t)x * sf.bpp;
    uint8_t *d = dst + (size_t)y * dstride + (size_t)x * df.bpp;
    const uint8_t r = s[sf.r], g = s[sf.g], b = s[sf.b], a = sf.a >= 0 ? s[sf.a] : (uint8_t)255;
    if (df.bpp == 4) d[6 - df.r - df.g - df.b] = (df.a >= 0) ? a : (uint8_t)255;   // the 4th byte: alpha, or padding = 255
    d[df.r] = r; d[df.g] = g; d[df.b] = b;
  }
}

// fixed-point colour matrices (x 256), limited range.  m[0..2] Y, m[3..5] Cb, m[6..8] Cr from R,G,B; inverse: see kernels.
struct YuvMatrix { int yr, yg, yb, ur, ug, ub, vr, vg, vb; int ry, rv, gy, gu, gv, by, bu; };

__device__ __forceinline__ int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
__device__ __forceinline__ void rgb_to_yuv(const YuvMatrix &m, int r, int g, int b, int &y, int &u, int &v) {
  y = clamp255((m.yr * r + m.yg * g + m.yb * b + (16 << 8) + 128) >> 8);
  u = clamp255((m.ur * r + m.ug * g + m.ub * b + (128 << 8) + 128) >> 8);
  v = clamp255((m.vr * r + m.vg * g + m.vb * b + (128 << 8) + 128) >> 8);
}

// packed RGB-family -> I420 (+ A plane when ap != nullptr: A420).  One thread per 2x2 block.
__global__ void __launch_bounds__(256) rgb_to_i420_kernel(const uint8_t *__restrict__ src, long sstride, PackedFmt sf, int width,
                                                          int height, YuvMatrix m, uint8_t *__restrict__ yp, long ys,
                                                          uint8_t *__restrict__ up, long us, uint8_t *__restrict__ vp, long vs,
                                                          uint8_t *__restrict__ ap, long as) {
  const int cx = blockIdx.x * blockDim.x + threadIdx.x, cy = blockIdx.y * blockDim.y + threadIdx.y;
  const int cw = (width + 1) >> 1, ch = (height + 1) >> 1;
  if (cx >= cw || cy >= ch) return;
  int su = 0, sv = 0;
#pragma unroll
  for (int dy = 0; dy < 2; dy++)
#pragma unroll
    for (int dx = 0; dx < 2; dx++) {
      const int x = min(2 * cx + dx, width - 1), y = min(2 * cy + dy, height - 1);   // edge replication for odd sizes
      const uint8_t *s = src + (size_t)y * sstride + (size_t)x * sf.bpp;
      int Y, U, V;
      rgb_to_yuv(m, s[sf.r], s[sf.g], s[sf.b], Y, U, V);
      su += U; sv += V;
      if (2 * cx + dx < width && 2 * cy + dy < height) {
        yp[(size_t)y * ys + x] = (uint8_t)Y;
        if (ap) ap[(size_t)y * as + x] = sf.a >= 0 ? s[sf.a] : (uint8_t)255;
      }
    }
  up[(size_t)cy * us + cx] = (uint8_t)((su + 2) >> 2);
  vp[(size_t)cy * vs + cx] = (uint8_t)((sv + 2) >> 2);
}

// I420 / A420 -> packed RGB-family, nearest chroma
__global__ void __launch_bounds__(256) i420_to_rgb_kernel(const uint8_t *__restrict__ yp, long ys, const uint8_t *__restrict__ up,
                                                          long us, const uint8_t *__restrict__ vp, long vs,
                                                          const uint8_t *__restrict__ ap, long as, int width, int height,
                                                          YuvMatrix m, uint8_t *__restrict__ dst, long dstride, PackedFmt df) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= width) return;
  for (int y = blockIdx.y; y < height; y += gridDim.y) {
    const int Y = (int)yp[(size_t)y * ys + x] - 16, U = (int)up[(size_t)(y >> 1) * us + (x >> 1)] - 128,
              V = (int)vp[(size_t)(y >> 1) * vs + (x >> 1)] - 128;
    const int r = clamp255((m.ry * Y + m.rv * V + 128) >> 8);
    const int g = clamp255((m.gy * Y + m.gu * U + m.gv * V + 128) >> 8);
    const int b = clamp255((m.by * Y + m.bu * U + 128) >> 8);
    uint8_t *d = dst + (size_t)y * dstride + (size_t)x * df.bpp;
    if (df.bpp == 4) d[6 - df.r - df.g - df.b] = (df.a >= 0 && ap) ? ap[(size_t)y * as + x] : (uint8_t)255;
    d[df.r] = (uint8_t)r; d[df.g] = (uint8_t)g; d[df.b] = (uint8_t)b;
  }
}

// ---- colorlut on I420 / A420 frames ("planar YUV via fused convert") --------------------------------
// "videoconvert ! colorlut ! videoconvert" (the doc pipeline of colorlut/imp.rs:18) on a planar frame = I420 -> RGB (spec
// above), ColorLut::transform_rgba (imp.rs:267-294), RGB -> I420 (spec above).  Fused: 1.5 B/px in, 1.5 B/px out instead of
// 19 B/px through two RGBA intermediates; the result is the composition, bit for bit.  The LUT is the memoised answer table
// (3D) or the three per-channel byte tables (1D).
__device__ __forceinline__ uint32_t yuv_to_rgb24(const YuvMatrix &m, int Y, int U, int V) {   // -> r | g << 8 | b << 16
  const int r = clamp255((m.ry * Y + m.rv * V + 128) >> 8);
  const int g = clamp255((m.gy * Y + m.gu * U + m.gv * V + 128) >> 8);
  const int b = clamp255((m.by * Y + m.bu * U + 128) >> 8);
  return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16);
}
// the same two conversions with the work shared inside a 2x2 block: the chroma terms of the inverse matrix are formed once
// per block, and the forward matrix is three dp4a on the packed table answer (Y weights as unsigned bytes, chroma weights
// as signed bytes; yuv_pack_ok() says whether the coefficients fit -- they do for BT.601 and BT.709).  Integer arithmetic:
// regrouping the sums changes nothing.
struct YuvChroma { int cr, cg, cb; };
__device__ __forceinline__ YuvChroma yuv_chroma_terms(const YuvMatrix &m, int U, int V) {
  return YuvChroma{m.rv * V + 128, m.gu * U + m.gv * V + 128, m.bu * U + 128};
}
__device__ __forceinline__ uint32_t yuv_to_rgb24(const YuvMatrix &m, int Y, const YuvChroma &c) {
  const int yt = m.ry * Y;   // ry == gy == by
  return (uint32_t)clamp255((yt + c.cr) >> 8) | ((uint32_t)clamp255((yt + c.cg) >> 8) << 8) | ((uint32_t)clamp255((yt + c.cb) >> 8) << 16);
}
struct YuvPacked { uint32_t wy, wu, wv; };   // byte 0: R weight, 1: G, 2: B
__device__ __forceinline__ int dp4a_u8_s8(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ void rgb24_to_yuv(const YuvPacked &w, uint32_t rgb, int &y, int &u, int &v) {
  y = clamp255((int)__dp4a(rgb, w.wy, (unsigned)((16 << 8) + 128)) >> 8);
  u = clamp255(dp4a_u8_s8(rgb, w.wu, (128 << 8) + 128) >> 8);
  v = clamp255(dp4a_u8_s8(rgb, w.wv, (128 << 8) + 128) >> 8);
}

template <bool LUT1D>
__device__ __forceinline__ uint32_t lut_rgb24(const uint32_t *__restrict__ memo, const uint8_t *tab, uint32_t c) {
  if (LUT1D) return (uint32_t)tab[c & 255u] | ((uint32_t)tab[256 + ((c >> 8) & 255u)] << 8) | ((uint32_t)tab[512 + (c >> 16)] << 16);
  return __ldg(memo + memo_index(c));
}

// any size, any alignment: one thread per 2x2 block (edge pixels replicated exactly like rgb_to_i420_kernel)
template <bool LUT1D>
__global__ void __launch_bounds__(256) colorlut_i420_kernel(const uint32_t *__restrict__ memo, const uint8_t *__restrict__ memo1d,
                                                            const uint8_t *__restrict__ yp, long ys, const uint8_t *__restrict__ up, long us,
                                                            const uint8_t *__restrict__ vp, long vs, int width, int height, YuvMatrix m, YuvPacked w,
                                                            uint8_t *__restrict__ oy, long oys, uint8_t *__restrict__ ou, long ous,
                                                            uint8_t *__restrict__ ov, long ovs) {
  __shared__ uint8_t tab[LUT1D ? 768 : 4];
  if (LUT1D) {
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < 768; i += blockDim.x * blockDim.y) tab[i] = memo1d[i];
    __syncthreads();
  }
  const int cx = blockIdx.x * blockDim.x + threadIdx.x, cy = blockIdx.y * blockDim.y + threadIdx.y;
  const int cw = (width + 1) >> 1, ch = (height + 1) >> 1;
  if (cx >= cw || cy >= ch) return;
  const YuvChroma ct = yuv_chroma_terms(m, (int)up[(size_t)cy * us + cx] - 128, (int)vp[(size_t)cy * vs + cx] - 128);
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int x = min(2 * cx + (k & 1), width - 1), y = min(2 * cy + (k >> 1), height - 1);
    o[k] = lut_rgb24<LUT1D>(memo, tab, yuv_to_rgb24(m, (int)yp[(size_t)y * ys + x] - 16, ct)) & 0x00FFFFFFu;
  }
  int su = 0, sv = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int x = 2 * cx + (k & 1), y = 2 * cy + (k >> 1);
    int Y2, U2, V2;
    rgb24_to_yuv(w, o[k], Y2, U2, V2);
    su += U2; sv += V2;
    if (x < width && y < height) oy[(size_t)y * oys + x] = (uint8_t)Y2;
  }
  ou[(size_t)cy * ous + cx] = (uint8_t)((su + 2) >> 2);
  ov[(size_t)cy * ovs + cx] = (uint8_t)((sv + 2) >> 2);
}

// width % 8 == 0, height % 2 == 0, Y rows 8-byte and chroma rows 4-byte aligned: a thread owns 8 x 2 pixels (four 2x2
// blocks): two 8-byte Y loads, one 4-byte U and V load, 16 table gathers in flight, the same widths on the way out
template <bool LUT1D>
__global__ void __launch_bounds__(256) colorlut_i420_x8_kernel(const uint32_t *__restrict__ memo, const uint8_t *__restrict__ memo1d,
                                                               const uint8_t *__restrict__ yp, long ys, const uint8_t *__restrict__ up, long us,
                                                               const uint8_t *__restrict__ vp, long vs, int width, int height, YuvMatrix m, YuvPacked w,
                                                               uint8_t *__restrict__ oy, long oys, uint8_t *__restrict__ ou, long ous,
                                                               uint8_t *__restrict__ ov, long ovs) {
  __shared__ uint8_t tab[LUT1D ? 768 : 4];
  if (LUT1D) {
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < 768; i += blockDim.x * blockDim.y) tab[i] = memo1d[i];
    __syncthreads();
  }
  const int gx = blockIdx.x * blockDim.x + threadIdx.x, cy = blockIdx.y * blockDim.y + threadIdx.y;
  if (gx * 8 >= width || cy * 2 >= height) return;
  const uint2 y0 = __ldcs(reinterpret_cast<const uint2 *>(yp + (size_t)(2 * cy) * ys) + gx);
  const uint2 y1 = __ldcs(reinterpret_cast<const uint2 *>(yp + (size_t)(2 * cy + 1) * ys) + gx);
  const uint32_t u4 = __ldcs(reinterpret_cast<const uint32_t *>(up + (size_t)cy * us) + gx);
  const uint32_t v4 = __ldcs(reinterpret_cast<const uint32_t *>(vp + (size_t)cy * vs) + gx);
  const uint32_t yw[2][2] = {{y0.x, y0.y}, {y1.x, y1.y}};
  uint32_t o[2][8];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const YuvChroma ct = yuv_chroma_terms(m, (int)((u4 >> (8 * j)) & 255u) - 128, (int)((v4 >> (8 * j)) & 255u) - 128);
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int d = 0; d < 2; d++) {
        const int px = 2 * j + d;
        const int Y = (int)((yw[r][px >> 2] >> (8 * (px & 3))) & 255u) - 16;
        o[r][px] = lut_rgb24<LUT1D>(memo, tab, yuv_to_rgb24(m, Y, ct)) & 0x00FFFFFFu;
      }
  }
  uint32_t yo[2][2] = {{0u, 0u}, {0u, 0u}}, uo = 0u, vo = 0u;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    int su = 0, sv = 0;
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int d = 0; d < 2; d++) {
        const int px = 2 * j + d;
        int Y2, U2, V2;
        rgb24_to_yuv(w, o[r][px], Y2, U2, V2);
        su += U2; sv += V2;
        yo[r][px >> 2] |= (uint32_t)Y2 << (8 * (px & 3));
      }
    uo |= (uint32_t)((su + 2) >> 2) << (8 * j);
    vo |= (uint32_t)((sv + 2) >> 2) << (8 * j);
  }
  __stcs(reinterpret_cast<uint2 *>(oy + (size_t)(2 * cy) * oys) + gx, make_uint2(yo[0][0], yo[0][1]));
  __stcs(reinterpret_cast<uint2 *>(oy + (size_t)(2 * cy + 1) * oys) + gx, make_uint2(yo[1][0], yo[1][1]));
  __stcs(reinterpret_cast<uint32_t *>(ou + (size_t)cy * ous) + gx, uo);
  __stcs(reinterpret_cast<uint32_t *>(ov + (size_t)cy * ovs) + gx, vo);
}

}  // namespace b200vfx
