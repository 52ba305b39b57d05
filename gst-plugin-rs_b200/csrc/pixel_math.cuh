// pixel_math.cuh -- exact (bit-for-bit) device restatement of the reference's per-pixel HSV f32
// arithmetic (colorlut lives in colorlut_math.cuh).  Every operator is a single IEEE-754 binary32 round-to-nearest-even operation,
// spelled with __f*_rn intrinsics so ptxas can never contract a mul+add into an FMA (the file
// is also compiled with -fmad=false).  References are paths inside gst-plugins-rs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "colorlut_math.cuh"

namespace b200vfx {

// ---------------------------------------------------------------------------------------------
// hsvutils  (video/hsv/src/hsvutils.rs:42-198)
//
// The reference spends 8 IEEE divisions and 3 fmodf per pixel.  The restatement below produces the
// SAME f32 values with 3 divisions and no fmodf on the common path, using only rewrites that are
// exact identities in binary32 (each one is argued where it is used and checked over all 2^24
// colours against the oracle in tests/test_gpu_parity.py):
//   * (u8 as f32) / 255.0  ->  256-entry table of the same IEEE quotients (d255[], shared memory)
//   * |value - c| < 1e-5   ->  integer test "c is the max byte" (neighbouring quotients differ by 1/255)
//   * clamp(sat), clamp(value): no-ops because 0 <= chroma <= value <= 1
//   * hue % 360 with hue in [0,360]: only hue == 360 changes (-> 0)
//   * x % 360 for |x| < 8192 and hp % 2 for hp in [0,6]: exact subtraction chains (Sterbenz)
// ---------------------------------------------------------------------------------------------
struct Hsv { float h, s, v; };

// fmodf(t, 360.0f), bit-exact.  Fast path |t| < 8192: q = floor(|t|/360) is at most 22, q*360 is an
// exact integer, |t| - q*360 is a multiple of ulp(t) below 2^10 -> exactly representable; one exact
// +-360 repairs an off-by-one q.  Sign follows the dividend (fmodf(-360,360) = -0).
__device__ __forceinline__ float fmod360_exact(float t) {
  const float a = fabsf(t);
  if (!(a < 8192.0f)) return fmodf(t, 360.0f);  // huge / inf / NaN: library path (rare)
  const float q = floorf(__fmul_rn(a, 0.0027777778f));
  float r = __fsub_rn(a, __fmul_rn(q, 360.0f));
  if (r < 0.0f) r = __fadd_rn(r, 360.0f);
  if (r >= 360.0f) r = __fsub_rn(r, 360.0f);
  return copysignf(r, t);
}

// from_rgb (hsvutils.rs:44-84); from_bgr is the same with r/b swapped by the caller.
// d255[i] == (float)i / 255.0f (IEEE), staged in shared memory by the kernels.
__device__ __forceinline__ Hsv hsv_from_rgb(const float *__restrict__ d255, unsigned rb, unsigned gb, unsigned bb) {
  const float r = d255[rb], g = d255[gb], b = d255[bb];
  const unsigned mx = max(max(rb, gb), bb), mn = min(min(rb, gb), bb);
  const float value = d255[mx];                       // *max as f32 / 255.0
  const float chroma = __fsub_rn(value, d255[mn]);    // value - (*min as f32 / 255.0)
  float hue = 0.0f;
  if (mx != mn) {                                     // chroma == 0.0  <=>  max == min
    float num, add;
    if (rb == mx) { num = __fsub_rn(g, b); add = 0.0f; }        // |value - r| < EPSILON
    else if (gb == mx) { num = __fsub_rn(b, r); add = 2.0f; }   // |value - g| < EPSILON
    else { num = __fsub_rn(r, g); add = 4.0f; }                 // |value - b| < EPSILON
    const float ratio = __fdiv_rn(num, chroma);
    // 60*(ratio) for the red branch, 60*(2+ratio) / 60*(4+ratio) otherwise; 0.0 + ratio would turn -0 into +0,
    // harmless (hue = +-0 behaves identically below) but keep the red branch literal anyway
    hue = (rb == mx) ? __fmul_rn(60.0f, ratio) : __fmul_rn(60.0f, __fadd_rn(add, ratio));
    if (hue < 0.0f) hue = __fadd_rn(hue, 360.0f);
    if (hue >= 360.0f) hue = __fsub_rn(hue, 360.0f);   // hue % 360.0 for hue in [0, 360]
  }
  Hsv o;
  o.h = hue;
  o.s = (mx == 0u) ? 0.0f : __fdiv_rn(chroma, value);  // clamp(0,1) is a no-op: 0 <= chroma <= value
  o.v = value;                                          // clamp(0,1) is a no-op
  return o;
}

// to_rgb (hsvutils.rs:132-163): returns the three bytes (r,g,b); `as u8` truncates, NaN -> 0.
// Precondition (holds for every caller): in.h is in [0, 360], -0, or NaN.
__device__ __forceinline__ void hsv_to_rgb(const Hsv &in, unsigned &ro, unsigned &go, unsigned &bo) {
  const float c = __fmul_rn(in.v, in.s);
  const float hp = __fdiv_rn(in.h, 60.0f);
  // hp % 2.0 for hp in [0,6]: exact subtractions; NaN falls through to NaN - 6 = NaN like fmodf
  const float hm = (hp < 2.0f) ? hp : ((hp < 4.0f) ? __fsub_rn(hp, 2.0f) : ((hp < 6.0f) ? __fsub_rn(hp, 4.0f) : __fsub_rn(hp, 6.0f)));
  const float x = __fmul_rn(c, __fsub_rn(1.0f, fabsf(__fsub_rn(hm, 1.0f))));
  float p0, p1, p2;
  if (hp < 0.0f) { p0 = 0.0f; p1 = 0.0f; p2 = 0.0f; }
  else if (hp <= 1.0f) { p0 = c; p1 = x; p2 = 0.0f; }
  else if (hp <= 2.0f) { p0 = x; p1 = c; p2 = 0.0f; }
  else if (hp <= 3.0f) { p0 = 0.0f; p1 = c; p2 = x; }
  else if (hp <= 4.0f) { p0 = 0.0f; p1 = x; p2 = c; }
  else if (hp <= 5.0f) { p0 = x; p1 = 0.0f; p2 = c; }
  else if (hp <= 6.0f) { p0 = c; p1 = 0.0f; p2 = x; }
  else { p0 = 0.0f; p1 = 0.0f; p2 = 0.0f; }
  const float m = __fsub_rn(in.v, c);
  // ((p + m) * 255).clamp(0,255) as u8 : cvt.rzi saturates below at 0 and maps NaN to 0; min() is the upper clamp
  ro = min(__float2uint_rz(__fmul_rn(__fadd_rn(p0, m), 255.0f)), 255u);
  go = min(__float2uint_rz(__fmul_rn(__fadd_rn(p1, m), 255.0f)), 255u);
  bo = min(__float2uint_rz(__fmul_rn(__fadd_rn(p2, m), 255.0f)), 255u);
}

struct HsvFilterSettings { float hue_shift, sat_mul, sat_off, val_mul, val_off; };
struct HsvDetectSettings { float hue_ref, hue_var, sat_ref, sat_var, val_ref, val_var; };

// hsv_filter body (hsvfilter/imp.rs:100-117); hsvutils::Clamp = max-then-min (NaN -> 0)
__device__ __forceinline__ void hsvfilter_px(const HsvFilterSettings &s, const float *__restrict__ d255, unsigned &r,
                                             unsigned &g, unsigned &b) {
  Hsv hsv = hsv_from_rgb(d255, r, g, b);
  float h = fmod360_exact(__fadd_rn(hsv.h, s.hue_shift));
  if (h < 0.0f) h = __fadd_rn(h, 360.0f);   // may round up to exactly 360.0 -> hp == 6
  hsv.h = h;
  hsv.s = fminf(fmaxf(__fadd_rn(__fmul_rn(s.sat_mul, hsv.s), s.sat_off), 0.0f), 1.0f);
  hsv.v = fminf(fmaxf(__fadd_rn(__fmul_rn(s.val_mul, hsv.v), s.val_off), 0.0f), 1.0f);
  hsv_to_rgb(hsv, r, g, b);
}

// hsv_detect predicate (hsvdetector/imp.rs:139-157)
__device__ __forceinline__ bool hsvdetect_px(const HsvDetectSettings &s, const float *__restrict__ d255, unsigned r,
                                             unsigned g, unsigned b) {
  const Hsv hsv = hsv_from_rgb(d255, r, g, b);
  float sh = __fadd_rn(hsv.h, __fsub_rn(180.0f, s.hue_ref));
  if (sh < 0.0f) sh = __fadd_rn(sh, 360.0f);
  sh = fmod360_exact(sh);
  return fabsf(__fsub_rn(sh, 180.0f)) <= s.hue_var && fabsf(__fsub_rn(hsv.s, s.sat_ref)) <= s.sat_var &&
         fabsf(__fsub_rn(hsv.v, s.val_ref)) <= s.val_var;
}

// fills the 256-entry (u8 as f32)/255.0 table (IEEE division, done once per CTA)
__device__ __forceinline__ void fill_d255(float *d255) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) d255[i] = __fdiv_rn((float)i, 255.0f);
  __syncthreads();
}

}  // namespace b200vfx
