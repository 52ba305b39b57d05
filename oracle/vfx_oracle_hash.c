/*
 * vfx_oracle_hash.c -- CPU ORACLE (test infrastructure, NOT product code): videocompare's image hashes.
 *
 * video/videofx/src/videocompare/hashed_image.rs:24-106 hands the frame to third-party crates that are NOT under
 * /root/reference: image_hasher 3.1.1 (Cargo.lock:7459-7461; HashAlg::{Mean, Gradient, VertGradient, DoubleGradient,
 * Blockhash}) which in turn uses image 0.25.10 (imageops::grayscale, imageops::resize with FilterType::Lanczos3).
 * Their algorithms are restated here FROM THEIR PUBLISHED SOURCE AS RECALLED -- nothing in this file could be checked
 * against the crates in this container: PARITY UNPINNED.  What the reference's own tests pin
 * (tests/videocompare.rs:57-139: identical frames -> distance 0; snow vs red -> distance > 0) is asserted in
 * tests/test_oracle_cpu.py for every algorithm.
 *
 * Restated pieces:
 *   image::imageops::grayscale      : luma = (2126 R + 7152 G + 722 B) / 10000, integer arithmetic (alpha ignored)
 *   image::imageops::resize         : two passes, vertical then horizontal; for every output sample the weights
 *                                     w_i = lanczos3((i - (centre - 0.5)) / sratio) over i in [left, right), normalised
 *                                     by their f32 sum; accumulation t += v * w in source order (f32, unfused); the
 *                                     intermediate image is f32; the final sample is clamp(t, 0, 255).round() as u8
 *   image_hasher alg/mod.rs         : resize dimensions (w,h) / (w+1,h) / (w,h+1) / (w/2+1,h/2+1), bit rules
 *   image_hasher alg/blockhash.rs   : integer fast path (vfx_oracle.c) and the f32 path for sizes that are not multiples
 *                                     of the hash grid, including its `x + 1. % block_width` operator-precedence quirk
 *                                     (the fractional weights collapse to 0/1); medians over groups of 4 hash rows.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "vfx_oracle.h"

/* ---- image::imageops::sample ---------------------------------------------------------------------------------- */
static float orc_sinc(float t) {
  const float a = t * 3.14159274101257324f; /* f32::consts::PI */
  return t == 0.0f ? 1.0f : sinf(a) / a;
}
static float orc_lanczos3(float x) { return fabsf(x) < 3.0f ? orc_sinc(x) * orc_sinc(x / 3.0f) : 0.0f; }

/* weights of output sample `out` when `in_len` source samples are resized to `out_len` (vertical_sample /
 * horizontal_sample share this code): returns the number of taps, fills *left and ws[] (normalised). */
int orc_resize_taps(int in_len, int out_len, int out, int *left_out, float *ws, int cap) {
  const float ratio = (float)in_len / (float)out_len;
  const float sratio = ratio < 1.0f ? 1.0f : ratio;
  const float src_support = 3.0f * sratio;
  float input = ((float)out + 0.5f) * ratio;
  long long left = (long long)floorf(input - src_support);
  if (left < 0) left = 0;
  if (left > (long long)in_len - 1) left = (long long)in_len - 1;
  long long right = (long long)ceilf(input + src_support);
  if (right < left + 1) right = left + 1;
  if (right > in_len) right = in_len;
  input = input - 0.5f;
  const int n = (int)(right - left);
  if (n > cap) return -1;
  float sum = 0.0f;
  for (int i = 0; i < n; i++) {
    const float w = orc_lanczos3(((float)(left + i) - input) / sratio);
    ws[i] = w;
    sum += w;
  }
  for (int i = 0; i < n; i++) ws[i] /= sum;
  *left_out = (int)left;
  return n;
}

static inline uint8_t orc_luma(const uint8_t *p) { /* image::color rgb_to_luma: integer, truncating division */
  return (uint8_t)((2126u * p[0] + 7152u * p[1] + 722u * p[2]) / 10000u);
}

/* grayscale + resize to nw x nh (Lanczos3).  fmt: ORC_FMT_RGB or ORC_FMT_RGBA.  out: nw*nh bytes, row-major. */
int orc_luma_resize(int fmt, int width, int height, const uint8_t *src, int stride, int nw, int nh, uint8_t *out) {
  if ((fmt != ORC_FMT_RGB && fmt != ORC_FMT_RGBA) || width <= 0 || height <= 0 || nw <= 0 || nh <= 0) return -1;
  const int bpp = fmt == ORC_FMT_RGB ? 3 : 4;
  if (nw == width && nh == height) { /* resize() copies when the dimensions are unchanged */
    for (int y = 0; y < height; y++)
      for (int x = 0; x < width; x++) out[(size_t)y * nw + x] = orc_luma(src + (size_t)y * stride + (size_t)x * bpp);
    return 0;
  }
  float *tmp = (float *)malloc(sizeof(float) * (size_t)width * (size_t)nh);
  float *ws = (float *)malloc(sizeof(float) * (size_t)((width > height ? width : height) + 8));
  if (!tmp || !ws) { free(tmp); free(ws); return -2; }
  /* vertical_sample: out(x, outy) = sum_i luma(x, left + i) * w_i, accumulated in i order */
  for (int oy = 0; oy < nh; oy++) {
    int left;
    const int n = orc_resize_taps(height, nh, oy, &left, ws, height + 8);
    for (int x = 0; x < width; x++) {
      float t = 0.0f;
      for (int i = 0; i < n; i++) {
        const float v = (float)orc_luma(src + (size_t)(left + i) * stride + (size_t)x * bpp);
        const float prod = v * ws[i];
        t += prod;
      }
      tmp[(size_t)oy * width + x] = t;
    }
  }
  /* horizontal_sample on the f32 intermediate, then clamp(0,255).round() */
  for (int ox = 0; ox < nw; ox++) {
    int left;
    const int n = orc_resize_taps(width, nw, ox, &left, ws, width + 8);
    for (int y = 0; y < nh; y++) {
      float t = 0.0f;
      for (int i = 0; i < n; i++) {
        const float prod = tmp[(size_t)y * width + left + i] * ws[i];
        t += prod;
      }
      t = t < 0.0f ? 0.0f : (t > 255.0f ? 255.0f : t);
      out[(size_t)y * nw + ox] = (uint8_t)roundf(t); /* FloatNearest: round half away from zero */
    }
  }
  free(tmp); free(ws);
  return 0;
}

/* ---- image_hasher alg/mod.rs ---------------------------------------------------------------------------------- */
/* algo: 0 mean, 1 gradient, 2 vertgradient, 3 doublegradient (videocompare/mod.rs:57-92); hash size 8x8 */
void orc_hash_resize_dims(int algo, int *nw, int *nh) {
  switch (algo) {
    case 0: *nw = 8; *nh = 8; break;
    case 1: *nw = 9; *nh = 8; break;
    case 2: *nw = 8; *nh = 9; break;
    default: *nw = 8 / 2 + 1; *nh = 8 / 2 + 1; break;
  }
}
/* bits from the resized luma (row-major nw x nh); returns the number of bits written (bytes of 0/1) */
int orc_hash_bits_from_luma(int algo, const uint8_t *l, int nw, int nh, uint8_t *bits) {
  int n = 0;
  if (algo == 0) { /* mean_hash_u8: mean = (sum / len) as u8; bit = x >= mean */
    unsigned sum = 0;
    for (int i = 0; i < nw * nh; i++) sum += l[i];
    const uint8_t mean = (uint8_t)(sum / (unsigned)(nw * nh));
    for (int i = 0; i < nw * nh; i++) bits[n++] = (uint8_t)(l[i] >= mean);
    return n;
  }
  if (algo == 1 || algo == 3) /* gradient_hash: per row, last < this */
    for (int y = 0; y < nh; y++)
      for (int x = 1; x < nw; x++) bits[n++] = (uint8_t)(l[y * nw + x - 1] < l[y * nw + x]);
  if (algo == 2 || algo == 3) /* vert_gradient_hash: per column, going down */
    for (int x = 0; x < nw; x++)
      for (int y = 1; y < nh; y++) bits[n++] = (uint8_t)(l[(y - 1) * nw + x] < l[y * nw + x]);
  return n;
}

/* ---- image_hasher alg/blockhash.rs: f32 path (width % hw != 0 || height % hh != 0) -------------------------------- */
int orc_blockhash_sums_f32(int fmt, int width, int height, const uint8_t *src, int stride, int hw, int hh, float *blocks) {
  if ((fmt != ORC_FMT_RGB && fmt != ORC_FMT_RGBA) || hw <= 0 || hh <= 0 || width <= 0 || height <= 0) return -1;
  const int bpp = fmt == ORC_FMT_RGB ? 3 : 4;
  const float block_width = (float)width / (float)hw, block_height = (float)height / (float)hh;
  for (int i = 0; i < hw * hh; i++) blocks[i] = 0.0f;
  for (int yi = 0; yi < height; yi++) {
    const uint8_t *p = src + (size_t)yi * (size_t)stride;
    for (int xi = 0; xi < width; xi++, p += bpp) {
      unsigned s = (unsigned)p[0] + p[1] + p[2];
      if (bpp == 4 && p[3] == 0) s = 765;
      const float px_sum = (float)s;
      const float x = (float)xi, y = (float)yi;
      const float block_x = x / block_width, block_y = y / block_height;
      /* `x + 1. % block_width` parses as x + (1. % block_width) */
      const float x_mod = x + fmodf(1.0f, block_width), y_mod = y + fmodf(1.0f, block_height);
      const float weight_left = x_mod - truncf(x_mod), weight_right = 1.0f - weight_left;
      const float weight_top = y_mod - truncf(y_mod), weight_bottom = 1.0f - weight_top;
      const unsigned block_left = (unsigned)floorf(block_x), block_top = (unsigned)floorf(block_y);
      const unsigned block_right = truncf(x_mod) == 0.0f ? (unsigned)ceilf(block_x) : block_left;
      const unsigned block_bottom = truncf(y_mod) == 0.0f ? (unsigned)ceilf(block_y) : block_top;
      blocks[block_top * (unsigned)hw + block_left] += px_sum * weight_left * weight_top;
      blocks[block_bottom * (unsigned)hw + block_left] += px_sum * weight_left * weight_bottom;
      blocks[block_top * (unsigned)hw + block_right] += px_sum * weight_right * weight_top;
      blocks[block_bottom * (unsigned)hw + block_right] += px_sum * weight_right * weight_bottom;
    }
  }
  return 0;
}

static int cmp_f32(const void *a, const void *b) {
  const float x = *(const float *)a, y = *(const float *)b;
  return (x > y) - (x < y);
}
/* gen_hash! for f32 blocks: groups of hw*4 blocks, median = sorted[len/2], bit = b > m || (|b - m| < 0.001 && m > cmp) */
void orc_blockhash_bits_f32(const float *blocks, int hw, int hh, int width, int height, uint8_t *bits_out) {
  const int n = hw * hh, group = hw * 4;
  const float block_width = (float)width / (float)hw, block_height = (float)height / (float)hh;
  const float block_area = block_width * block_height;
  const float cmp_factor = 255.0f * 3.0f * block_area / 2.0f;
  float *scratch = (float *)malloc(sizeof(float) * (size_t)(group > 0 ? group : 1));
  for (int g0 = 0; g0 < n; g0 += group) {
    const int len = (n - g0 < group) ? n - g0 : group;
    memcpy(scratch, blocks + g0, sizeof(float) * (size_t)len);
    qsort(scratch, (size_t)len, sizeof(float), cmp_f32);
    const float m = scratch[len / 2];
    for (int i = 0; i < len; i++) {
      const float v = blocks[g0 + i];
      bits_out[g0 + i] = (uint8_t)(v > m || (fabsf(v - m) < 0.001f && m > cmp_factor));
    }
  }
  free(scratch);
}
