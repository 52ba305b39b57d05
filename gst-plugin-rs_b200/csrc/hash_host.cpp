// hash_host.cpp -- host side of videocompare's image hashes (video/videofx/src/videocompare/hashed_image.rs:24-106 ->
// image_hasher 3.1.1 / image 0.25.10, third-party crates absent from the reference tree: restated from their published
// source AS RECALLED, parity unpinned -- only `distance == 0 for identical frames` and `> 0 for snow vs red`
// (tests/videocompare.rs:57-139) are pinned).  Tiny work on at most 81 bytes / 64 sums per frame; the per-pixel work is
// in hash_kernels.cuh.
#include "hash_host.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../../include/b200vfx.h"

namespace b200vfx {

namespace {
// imageops::sample::{sinc, lanczos3_kernel}
float sinc(float t) {
  const volatile float a = t * 3.14159274101257324f;   // f32::consts::PI; volatile: one rounding per operator
  if (t == 0.0f) return 1.0f;
  const volatile float s = std::sin((float)a);
  const volatile float q = s / a;
  return q;
}
float lanczos3(float x) {
  if (!(std::fabs(x) < 3.0f)) return 0.0f;
  const volatile float x3 = x / 3.0f;
  const volatile float p = sinc(x) * sinc(x3);
  return p;
}
}  // namespace

ResizeTaps make_resize_taps(int in_len, int out_len) {
  ResizeTaps r;
  r.left.resize((size_t)out_len);
  r.count.resize((size_t)out_len);
  const volatile float ratio = (float)in_len / (float)out_len;
  const float sratio = ratio < 1.0f ? 1.0f : (float)ratio;
  const volatile float src_support = 3.0f * sratio;
  std::vector<std::vector<float>> rows((size_t)out_len);
  for (int o = 0; o < out_len; o++) {
    const volatile float centre = ((float)o + 0.5f) * ratio;
    const volatile float lo = centre - src_support, hi = centre + src_support;
    long long left = (long long)std::floor((float)lo);
    left = std::max<long long>(0, std::min<long long>(left, (long long)in_len - 1));
    long long right = (long long)std::ceil((float)hi);
    right = std::max<long long>(left + 1, std::min<long long>(right, in_len));
    const volatile float input = centre - 0.5f;
    std::vector<float> &w = rows[(size_t)o];
    volatile float sum = 0.0f;
    for (long long i = left; i < right; i++) {
      const volatile float d = (float)i - input;
      const volatile float arg = d / sratio;
      const float k = lanczos3((float)arg);
      w.push_back(k);
      sum = sum + k;
    }
    for (float &k : w) { const volatile float q = k / sum; k = q; }
    r.left[(size_t)o] = (int)left;
    r.count[(size_t)o] = (int)w.size();
    r.max_taps = std::max(r.max_taps, (int)w.size());
  }
  r.taps.assign((size_t)out_len * (size_t)r.max_taps, 0.0f);
  for (int o = 0; o < out_len; o++) std::copy(rows[(size_t)o].begin(), rows[(size_t)o].end(), r.taps.begin() + (size_t)o * r.max_taps);
  return r;
}

}  // namespace b200vfx

extern "C" {

// test hook (host logic only): the normalised Lanczos3 taps of output sample `out`; returns the tap count
int b200vfx_debug_resize_taps(int in_len, int out_len, int out, int *left, float *ws, int cap) {
  if (in_len <= 0 || out_len <= 0 || out < 0 || out >= out_len || !left || !ws) return B200VFX_ERR_INVALID;
  const b200vfx::ResizeTaps t = b200vfx::make_resize_taps(in_len, out_len);
  const int n = t.count[(size_t)out];
  if (n > cap) return B200VFX_ERR_INVALID;
  *left = t.left[(size_t)out];
  std::copy(t.taps.begin() + (size_t)out * t.max_taps, t.taps.begin() + (size_t)out * t.max_taps + n, ws);
  return n;
}

int b200vfx_hash_resize_dims(int algo, int *nw, int *nh) {
  // HashAlg::resize_dimensions for the default 8x8 hash (image_hasher alg/mod.rs)
  switch (algo) {
    case B200VFX_HASH_MEAN: *nw = 8; *nh = 8; return 0;
    case B200VFX_HASH_GRADIENT: *nw = 9; *nh = 8; return 0;
    case B200VFX_HASH_VERTGRADIENT: *nw = 8; *nh = 9; return 0;
    case B200VFX_HASH_DOUBLEGRADIENT: *nw = 5; *nh = 5; return 0;
    default: return B200VFX_ERR_INVALID;   // blockhash does not resize
  }
}

int b200vfx_hash_bits_from_luma(int algo, const uint8_t *l, int nw, int nh, uint8_t *bits) {
  int n = 0;
  if (algo == B200VFX_HASH_MEAN) {   // mean_hash_u8: mean = (sum / len) as u8, bit = x >= mean
    unsigned sum = 0;
    for (int i = 0; i < nw * nh; i++) sum += l[i];
    const unsigned mean = (sum / (unsigned)(nw * nh)) & 255u;
    for (int i = 0; i < nw * nh; i++) bits[n++] = l[i] >= mean ? 1 : 0;
    return n;
  }
  if (algo == B200VFX_HASH_GRADIENT || algo == B200VFX_HASH_DOUBLEGRADIENT)      // rows: last < this
    for (int y = 0; y < nh; y++)
      for (int x = 0; x + 1 < nw; x++) bits[n++] = l[y * nw + x] < l[y * nw + x + 1] ? 1 : 0;
  if (algo == B200VFX_HASH_VERTGRADIENT || algo == B200VFX_HASH_DOUBLEGRADIENT)  // columns, downwards
    for (int x = 0; x < nw; x++)
      for (int y = 0; y + 1 < nh; y++) bits[n++] = l[y * nw + x] < l[(y + 1) * nw + x] ? 1 : 0;
  return n;
}

void b200vfx_blockhash_bits_f32(const float *blocks, int hw, int hh, int width, int height, uint8_t *bits_out) {
  // gen_hash! with $valty = f32: groups of `hash width * 4` blocks, median = element len/2 of the sorted group,
  // bit = block > median || (|block - median| < 0.001 && median > 255 * 3 * block_area / 2)
  const int n = hw * hh, group = hw * 4;
  if (group <= 0) return;
  const volatile float bw = (float)width / (float)hw, bh = (float)height / (float)hh;
  const volatile float area = bw * bh;
  const volatile float c0 = 765.0f * area;
  const volatile float cmp = c0 / 2.0f;
  std::vector<float> tmp;
  for (int g0 = 0; g0 < n; g0 += group) {
    const int len = std::min(group, n - g0);
    tmp.assign(blocks + g0, blocks + g0 + len);
    std::nth_element(tmp.begin(), tmp.begin() + len / 2, tmp.end());
    const float m = tmp[(size_t)len / 2];
    for (int i = 0; i < len; i++) {
      const float v = blocks[g0 + i];
      const volatile float d = v - m;
      bits_out[g0 + i] = (uint8_t)(v > m || (std::fabs((float)d) < 0.001f && m > cmp));
    }
  }
}

}  // extern "C"
