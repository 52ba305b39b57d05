// hsv_fast.cuh -- the per-pixel arithmetic of the DIRECT hsvfilter / hsvdetector kernels (the ones animated
// properties get: every GObject property of both elements is mutable in PLAYING, hsvfilter/imp.rs:127-156, so a
// GstController can change them on every frame and no answer table survives).
//
// Same f32 values as the reference (hsvutils.rs:44-163, hsvfilter/imp.rs:100-117, hsvdetector/imp.rs:139-157), produced
// branch-free with ~100 instructions per pixel instead of ~260:
//   * the three IEEE divisions become reciprocal-seed + FMA-residual sequences whose results are PROVEN equal to the
//     IEEE quotient by exhaustive enumeration of their (finite) operand domains:
//       num/chroma   : operands are differences of two u8/255 quotients -> 2.8 M sorted byte triples, seed 255/(max-min)
//       chroma/value : 65 536 (max,min) byte pairs, seed 255/max
//       h/60         : every f32 in [2^-100, 360] and +-0 (8e8 values), seed RN(1/60); below 2^-124 the residual is
//                      inexact -- such an h needs 0 < |hue-shift| < 1e-30, which is routed to the general code
//   * `% 360` / `% 2` are exact subtraction chains (as in pixel_math.cuh), hp % 2 from the sector index
//   * the six-way sector select is one byte permute of the three truncated candidates (c+m, x+m, 0+m)
//   * hue-shift = NaN / +-inf / |shift| >= 7800 / 0 < |shift| < 1e-30 are uniform per frame and take the general code
//     (library fmodf, IEEE division)
//
// This file is compiled TWICE: by nvcc into the kernels, and by gcc (tests/models/hsv_fast_model.c, -ffp-contract=off)
// where tests/test_hsv_fast_model.py compares it with the oracle over all 2^24 colours and runs the exhaustive
// division checks.  The host build is test infrastructure only -- the product has no CPU path.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define HF_FN __device__ __forceinline__
#define HF_FN_HD __host__ __device__ __forceinline__
#define HF_ADD(a, b) __fadd_rn((a), (b))
#define HF_SUB(a, b) __fsub_rn((a), (b))
#define HF_MUL(a, b) __fmul_rn((a), (b))
#define HF_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define HF_DIV(a, b) __fdiv_rn((a), (b))
/* a + b / a - b issued as FFMA(.., one, ..) with a run-time 1.0 the compiler cannot fold: the SAME single rounding, but on
 * the FMA pipe -- ncu showed the ALU pipe (FADD, selects, logic) as the limiter of these kernels at 62 % vs 34 % */
#define HF_ADDF(a, b, one) __fmaf_rn((a), (one), (b))
#define HF_SUBF(a, b, one) __fmaf_rn(-(b), (one), (a))
#define HF_ADDF_RD(a, b, one) __fmaf_rd((a), (one), (b))
#define HF_ADD_SAT(a, b) __saturatef(__fadd_rn((a), (b)))   /* FADD.SAT: clamp to [0,1], NaN -> +0 == hsvutils::Clamp */
#define HF_ADDF_SAT(a, b, one) __saturatef(__fmaf_rn((a), (one), (b)))   /* the same on the FMA pipe: FFMA.SAT */
#define HF_FLOOR_UF(x, one) (__float_as_uint(__fmaf_rd((x), (one), 8388608.0f)) & 0x007FFFFFu)   /* HF_FLOOR_U on the FMA pipe */
/* truncation of a value in [0, 256) as the low byte of val + 2^23 rounded toward zero (no F2I: XU pipe) */
#define HF_TRUNC_MAGIC(x, one) __float_as_uint(__fmaf_rz((x), (one), 8388608.0f))
#define HF_PERMUTE2(a, b, sel) __byte_perm((a), (b), (sel))
#define HF_MAX(a, b) fmaxf((a), (b))
#define HF_MIN(a, b) fminf((a), (b))
#define HF_ABS(a) fabsf(a)
#define HF_FLOOR_SMALL(x) __fsub_rn(__fadd_rd((x), 8388608.0f), 8388608.0f) /* exact floor for 0 <= x < 2^22 */
#define HF_FLOOR_U(x) (__float_as_uint(__fadd_rd((x), 8388608.0f)) & 0x007FFFFFu) /* the same as an integer (NaN: garbage) */
#define HF_U2F_SMALL(u) __fsub_rn(__uint_as_float(0x4B000000u | (u)), 8388608.0f)  /* (float)u for u < 2^23, no I2F */
#define HF_PERMUTE(cand, sel) __byte_perm((cand), 0u, (sel))
#define HF_TRUNC_U8(x) min(__float2uint_rz(x), 255u)                          /* `as u8` after clamp(0,255): NaN, <0 -> 0 */
#define HF_FMOD(a, b) fmodf((a), (b))
#define HF_BITS(f) __float_as_uint(f)
#define HF_FLOAT(u) __uint_as_float(u)
#else
#include <math.h>
#include <string.h>
#define HF_FN static inline
#define HF_FN_HD static inline
static inline float hf_opaque(float x) { volatile float v = x; return v; }   /* one rounding, no re-association */
#define HF_ADD(a, b) hf_opaque((a) + (b))
#define HF_SUB(a, b) hf_opaque((a) - (b))
#define HF_MUL(a, b) hf_opaque((a) * (b))
#define HF_FMA(a, b, c) fmaf((a), (b), (c))
#define HF_DIV(a, b) hf_opaque((a) / (b))
#define HF_ADDF(a, b, one) hf_opaque((a) * (one) + (b))       /* one == 1.0f: a * one is exact */
#define HF_SUBF(a, b, one) hf_opaque((a) - (b) * (one))
#define HF_ADDF_RD(a, b, one) hf_opaque(floorf((a) * (one)) + (b))   /* only used as x + 2^23 rounded down, 0 <= x < 2^22 */
static inline float hf_add_sat(float a, float b) { const float s = hf_opaque(a + b); return fminf(fmaxf(s, 0.0f), 1.0f); }
#define HF_ADD_SAT(a, b) hf_add_sat((a), (b))
#define HF_ADDF_SAT(a, b, one) hf_add_sat((a) * (one), (b))
#define HF_FLOOR_UF(x, one) hf_floor_u((x) * (one))
static inline uint32_t hf_trunc_magic(float x) { return 0x4B000000u | ((x != x || x <= 0.0f) ? 0u : (uint32_t)x); }
#define HF_TRUNC_MAGIC(x, one) hf_trunc_magic((x) * (one))
static inline uint32_t hf_permute2(uint32_t a, uint32_t b, uint32_t sel) {   /* PRMT, default mode: bytes 0-3 = a, 4-7 = b */
  const unsigned long long ab = ((unsigned long long)b << 32) | a;
  uint32_t o = 0;
  for (int k = 0; k < 4; k++) o |= (uint32_t)((ab >> (8 * ((sel >> (4 * k)) & 7u))) & 255u) << (8 * k);
  return o;
}
#define HF_PERMUTE2(a, b, sel) hf_permute2((a), (b), (sel))
#define HF_MAX(a, b) fmaxf((a), (b))
#define HF_MIN(a, b) fminf((a), (b))
#define HF_ABS(a) fabsf(a)
#define HF_FLOOR_SMALL(x) floorf(x)
static inline unsigned hf_floor_u(float x) { return (x == x) ? (unsigned)floorf(x) : 0u; }
#define HF_FLOOR_U(x) hf_floor_u(x)
#define HF_U2F_SMALL(u) ((float)(u))
static inline uint32_t hf_permute(uint32_t cand, uint32_t sel) {   /* PRMT with a zero second operand, default mode */
  uint32_t o = 0;
  for (int k = 0; k < 4; k++) {
    const uint32_t n = (sel >> (4 * k)) & 7u;
    o |= (n < 4u ? ((cand >> (8 * n)) & 255u) : 0u) << (8 * k);
  }
  return o;
}
#define HF_PERMUTE(cand, sel) hf_permute((cand), (sel))
static inline unsigned hf_trunc_u8(float x) { return (x != x || x <= 0.0f) ? 0u : (x >= 255.0f ? 255u : (unsigned)x); }
#define HF_TRUNC_U8(x) hf_trunc_u8(x)
#define HF_FMOD(a, b) fmodf((a), (b))
static inline uint32_t hf_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float hf_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#define HF_BITS(f) hf_bits(f)
#define HF_FLOAT(u) hf_float(u)
#endif

#ifdef __cplusplus
namespace b200vfx {
#endif

// per-CTA shared-memory tables (3 KB + 32 B): all entries are IEEE-rounded results of the reference's own operators
struct HsvTables {
  float d255[256];    // i / 255.0
  float rdiff[256];   // 255.0 / i: seed for 1/chroma (index max-min) and for 1/value (index max); [0] = 0
  float v2[256];      // hsvfilter only: Clamp(value_mul * d255[i] + value_off, 0, 1) -- the new value depends on max alone
  uint32_t sel[8];    // byte-permute selectors of the six hue sectors (+ hp == 6, + NaN): see hsvf_to_rgb
};
// hsvf_to_rgb permutes bytes out of two words: w = {c+m, x+m, ..} (source bytes 0, 1) and y0 = bits(trunc(0+m) + 2^23)
// (source byte 4 = the value, byte 5 = 0x00).  Selector nibble k = source of output byte k; output byte 3 = 0.
//   0:(c,x,0) 1:(x,c,0) 2:(0,c,x) 3:(0,x,c) 4:(x,0,c) 5:(c,0,x) 6 (hp == 6.0): as 5, 7 (NaN): (0,0,0)
#define HSVF_SEL_INIT {0x5410u, 0x5401u, 0x5104u, 0x5014u, 0x5041u, 0x5140u, 0x5140u, 0x5444u}
// table entry at BYTE offset off4 = 4 * index (the kernels extract the colour bytes pre-scaled: no address arithmetic)
#define HF_TAB(tab, off4) (*(const float *)((const char *)(tab) + (off4)))

struct HsvF { float h, s, v; unsigned mx4; };

// ---- from_rgb (hsvutils.rs:44-84) -----------------------------------------------------------------------------
// value/chroma/hue/saturation exactly as the reference computes them; from_bgr = caller swaps r and b.
// r4, g4, b4 = 4 * byte value; one = 1.0f at run time (see HF_ADDF).
HF_FN struct HsvF hsvf_from_rgb(const struct HsvTables *T, unsigned r4, unsigned g4, unsigned b4, const float one) {
  const float r = HF_TAB(T->d255, r4), g = HF_TAB(T->d255, g4), b = HF_TAB(T->d255, b4);
  const unsigned mx = r4 > g4 ? (r4 > b4 ? r4 : b4) : (g4 > b4 ? g4 : b4);
  const unsigned mn = r4 < g4 ? (r4 < b4 ? r4 : b4) : (g4 < b4 ? g4 : b4);
  const float value = HF_TAB(T->d255, mx);                          // *max as f32 / 255.0
  const float chroma = HF_SUBF(value, HF_TAB(T->d255, mn), one);    // value - *min as f32 / 255.0
  // |value - c| < EPSILON  <=>  c is the max byte (neighbouring quotients are 1/255 apart); priority R, G, B
  const int isr = r4 == mx, isg = g4 == mx;
  const float na = isr ? g : (isg ? b : r), nb = isr ? b : (isg ? r : g);
  const float add = isr ? -0.0f : (isg ? 2.0f : 4.0f);  // -0.0 + x == x for every x: the red branch has no addend
  const float num = HF_SUBF(na, nb, one);
  // num / chroma: seed y0 = 255/(max-min) (within 2^-15 of 1/chroma), one Newton step, quotient + FMA residual.
  // Grey pixels (chroma == 0): y0 = rdiff[0] = 0 makes ratio = +0 and isr holds, so hue = 60 * (-0 + 0) = 0 as required.
  const float y0 = HF_TAB(T->rdiff, mx - mn);
  const float e = HF_FMA(-chroma, y0, 1.0f);
  const float y = HF_FMA(y0, e, y0);
  const float q0 = HF_MUL(num, y);
  const float rem = HF_FMA(-chroma, q0, num);
  const float ratio = HF_FMA(rem, y, q0);
  float hue = HF_MUL(60.0f, HF_ADDF(ratio, add, one));
  if (hue < 0.0f) hue = HF_ADDF(hue, 360.0f, one);
  if (hue >= 360.0f) hue = HF_ADDF(hue, -360.0f, one);  // hue % 360.0 for hue in [0, 360]
  // chroma / value: seed 255/max (within 2^-23 of 1/value); black (value == 0): seed 0 gives saturation 0 as required
  const float yv = HF_TAB(T->rdiff, mx);
  const float s0 = HF_MUL(chroma, yv);
  const float srem = HF_FMA(-value, s0, chroma);
  struct HsvF o;
  o.h = hue;
  o.s = HF_FMA(srem, yv, s0);                           // clamp(0,1) is a no-op: 0 <= chroma <= value
  o.v = value;
  o.mx4 = mx;
  return o;
}

// h / 60.0 for h in [2^-100, 360], +-0 or NaN: reciprocal multiply + FMA residual, proven == IEEE over that range
// (for h = -0.0 the result is +0.0: the sign of a zero hp is never observed)
HF_FN float hsvf_div60(float h) {
  const float c = 1.0f / 60.0f;
  const float q0 = HF_MUL(h, c);
  const float r = HF_FMA(-q0, 60.0f, h);
  return HF_FMA(r, c, q0);
}

// fmodf(t, 360) followed by `if h < 0 { h += 360 }` (hsvfilter/imp.rs:102-105), for |t| < 8192:
// q = floor(|t|/360) <= 22, q*360 exact, |t| - q*360 exact (a multiple of ulp(t) below 2^10, so the FMA is exact too),
// one +-360 repairs an off-by-one q.  A zero result may come out as +0 where the reference has -0 (never observed).
HF_FN float hsvf_wrap360_small(float t, const float one) {
  const float a = HF_ABS(t);
  const float q = HF_SUBF(HF_ADDF_RD(HF_MUL(a, 0.0027777778f), 8388608.0f, one), 8388608.0f, one);   // floor, 0 <= x < 23
  float r = HF_FMA(-q, 360.0f, a);
  if (r < 0.0f) r = HF_ADDF(r, 360.0f, one);
  if (r >= 360.0f) r = HF_ADDF(r, -360.0f, one);
  // t >= 0: r.  t < 0: fmod = -r, negative unless r == 0 -> -r + 360 = RN(360 - r) (may round to exactly 360.0)
  if (t < 0.0f && r > 0.0f) r = HF_SUBF(360.0f, r, one);
  return r;
}
HF_FN float hsvf_wrap360_general(float t) {
  float h = HF_FMOD(t, 360.0f);
  h = (h < 0.0f) ? HF_ADD(h, 360.0f) : h;
  return h;
}

// to_rgb (hsvutils.rs:132-163) for h in [0, 360], -0 or NaN and s, v in [0, 1]; returns r | g<<8 | b<<16.
// general = 0: h is known not to be NaN (hsvf_shift_class) and not near-denormal.
HF_FN uint32_t hsvf_to_rgb(const struct HsvTables *T, float h, float s, float v, int general, const float one) {
  const float c = HF_MUL(v, s);
  const float hp = general ? HF_DIV(h, 60.0f) : hsvf_div60(h);
  // sector index i = floor(hp) in 0..6; hp % 2.0 == hp - 2*(i>>1) exactly (Sterbenz).  NaN: hm = NaN, selector 7.
  const unsigned i = HF_FLOOR_UF(hp, one) & 7u;
  const float hm = HF_SUBF(hp, HF_U2F_SMALL(i & 6u), one);
  const float x = HF_MUL(c, HF_SUBF(1.0f, HF_ABS(HF_SUBF(hm, 1.0f, one)), one));
  const float m = HF_SUBF(v, c, one);
  // ((p + m) * 255).clamp(0,255) as u8 for p in {c, x, 0}: truncation, NaN -> 0.  0 <= p <= c and m = v - c >= 0, so
  // (p + m) * 255 lies in [0, 255.0001]: the clamp never acts and the byte is the low byte of value + 2^23 rounded
  // toward zero.  (0.0 + m differs from m only for m = -0.0, and both truncate to 0.)
  const uint32_t yc = HF_TRUNC_MAGIC(HF_MUL(HF_ADDF(c, m, one), 255.0f), one);
  const uint32_t yx = HF_TRUNC_MAGIC(HF_MUL(HF_ADDF(x, m, one), 255.0f), one);
  const uint32_t y0 = HF_TRUNC_MAGIC(HF_MUL(m, 255.0f), one);
  // the reference's `<=` sector boundaries are harmless: on an integer hp, x equals c or 0 and both neighbours agree
  const unsigned k4 = (!general || hp == hp) ? 4u * i : 28u;
  const uint32_t w = HF_PERMUTE2(yc, yx, 0x7740u);      // byte 0 = c + m, byte 1 = x + m
  return HF_PERMUTE2(w, y0, *(const uint32_t *)((const char *)T->sel + k4));
}

struct HsvFilterParams { float hue_shift, sat_mul, sat_off, val_mul, val_off; };
struct HsvDetectParams { float hue_ref, hue_var, sat_ref, sat_var, val_ref, val_var; };

// hue-shift class, decided once per call: 0 = the fast code is proven for it -- |shift| < 7800 (|h + shift| stays below
// 8192) and shift is zero or not tiny (h + shift is then 0 or >= 2^-49, never a near-denormal); 1 = anything else
// (huge, +-inf, NaN, 0 < |shift| < 1e-30): library fmodf and IEEE division
HF_FN_HD int hsvf_shift_class(float shift) {
  const float a = fabsf(shift);
  return (a < 7800.0f && (a == 0.0f || a >= 1e-30f)) ? 0 : 1;
}

// entry of the per-launch v2 table: Clamp(value_mul * value + value_off, 0, 1), hsvutils::Clamp = max-then-min (NaN -> 0)
HF_FN float hsvf_v2_entry(const struct HsvFilterParams *p, float value) {
  return HF_ADD_SAT(HF_MUL(p->val_mul, value), p->val_off);   /* table build: 256 entries per launch, pipe does not matter */
}

// hsv_filter body (hsvfilter/imp.rs:100-117).  r4,g4,b4 = 4 * byte; packed bytes r | g<<8 | b<<16 out.
HF_FN uint32_t hsvf_filter_px(const struct HsvTables *T, const struct HsvFilterParams *p, int shift_class, unsigned r4,
                              unsigned g4, unsigned b4, const float one) {
  const struct HsvF a = hsvf_from_rgb(T, r4, g4, b4, one);
  const float t = HF_ADDF(a.h, p->hue_shift, one);
  const float h = shift_class ? hsvf_wrap360_general(t) : hsvf_wrap360_small(t, one);
  const float s = HF_ADDF_SAT(HF_MUL(p->sat_mul, a.s), p->sat_off, one);
  const float v = HF_TAB(T->v2, a.mx4);
  return hsvf_to_rgb(T, h, s, v, shift_class, one);
}

// hsv_detect predicate (hsvdetector/imp.rs:139-157).  ref_class = hsvf_shift_class(180 - hue_ref).
HF_FN int hsvf_detect_px(const struct HsvTables *T, const struct HsvDetectParams *p, int ref_class, unsigned r4, unsigned g4,
                         unsigned b4, const float one) {
  const struct HsvF a = hsvf_from_rgb(T, r4, g4, b4, one);
  float sh = HF_ADDF(a.h, HF_SUB(180.0f, p->hue_ref), one);
  if (sh < 0.0f) sh = HF_ADDF(sh, 360.0f, one);
  if (ref_class) {
    sh = HF_FMOD(sh, 360.0f);
  } else {  // sh % 360 for |sh| < 8192; only |sh - 180| is consumed, so the sign of a zero result does not matter
    const float aa = HF_ABS(sh);
    const float q = HF_SUBF(HF_ADDF_RD(HF_MUL(aa, 0.0027777778f), 8388608.0f, one), 8388608.0f, one);
    float rr = HF_FMA(-q, 360.0f, aa);
    if (rr < 0.0f) rr = HF_ADDF(rr, 360.0f, one);
    if (rr >= 360.0f) rr = HF_ADDF(rr, -360.0f, one);
    sh = (sh < 0.0f) ? -rr : rr;
  }
  return HF_ABS(HF_SUBF(sh, 180.0f, one)) <= p->hue_var && HF_ABS(HF_SUBF(a.s, p->sat_ref, one)) <= p->sat_var &&
         HF_ABS(HF_SUBF(a.v, p->val_ref, one)) <= p->val_var;
}

#ifdef __CUDACC__
// ---- two pixels per f32x2 lane pair (device only) ----------------------------------------------------------------------
// hsvf_filter_px for shift class 0, with every FADD / FMUL / FFMA of the scalar code issued ONCE for two pixels (FFMA2 /
// FADD2.RM / FADD2.RZ).  Lane-wise each packed operation is the scalar operation of hsvf_filter_px -- same operands, same
// single rounding -- so the results are identical by construction (and checked over all 2^24 colours on the GPU):
//   a + b, a - b      : fma2(a, +-one2, b) exactly as HF_ADDF / HF_SUBF (one = 1.0f at run time)
//   products that feed an add/sub (60*(..), a*(1/360), smul*s, v*s, c*w, (..)*255): fma2(x, y, nz2) with nz = -0.0f at run
//                       time -- the correctly rounded product, which ptxas cannot contract with the following add
//   products that feed FMAs only (q0 = num*y, s0 = chroma*yv, h*(1/60)): plain mul2
//   -chroma           : vmin - value (RN is sign-symmetric; for chroma == 0 the zero's sign is immaterial, see scalar code)
// Table look-ups, the R/G/B priority selects, the +-360 fix-ups, the two saturating adds and the byte permutes stay scalar.
HF_FN void hsvf_filter_px2(const struct HsvTables *T, const struct HsvFilterParams *p, const unsigned r4[2], const unsigned g4[2],
                           const unsigned b4[2], const float one, const float nzero, uint32_t out[2]) {
  const float MAGIC = 8388608.0f;
  const f32x2_t one2 = pk2(one, one), mone2 = pk2(-one, -one), nz2 = pk2(nzero, nzero);
  float value[2], vmin[2], na[2], nb[2], add[2], y0[2], yv[2], vnew[2];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const float r = HF_TAB(T->d255, r4[k]), g = HF_TAB(T->d255, g4[k]), b = HF_TAB(T->d255, b4[k]);
    const unsigned mx = max(max(r4[k], g4[k]), b4[k]), mn = min(min(r4[k], g4[k]), b4[k]);
    value[k] = HF_TAB(T->d255, mx);
    vmin[k] = HF_TAB(T->d255, mn);
    const bool isr = r4[k] == mx, isg = g4[k] == mx;
    na[k] = isr ? g : (isg ? b : r);
    nb[k] = isr ? b : (isg ? r : g);
    add[k] = isr ? -0.0f : (isg ? 2.0f : 4.0f);
    y0[k] = HF_TAB(T->rdiff, mx - mn);
    yv[k] = HF_TAB(T->rdiff, mx);
    vnew[k] = HF_TAB(T->v2, mx);
  }
  const f32x2_t value2 = pk2(value[0], value[1]), vmin2 = pk2(vmin[0], vmin[1]), y02 = pk2(y0[0], y0[1]), yv2 = pk2(yv[0], yv[1]);
  // from_rgb
  const f32x2_t chroma2 = fma2_rn(vmin2, mone2, value2), nchroma2 = fma2_rn(value2, mone2, vmin2);
  const f32x2_t num2 = fma2_rn(pk2(nb[0], nb[1]), mone2, pk2(na[0], na[1]));
  const f32x2_t e2 = fma2_rn(nchroma2, y02, pk2(1.0f, 1.0f));
  const f32x2_t y2 = fma2_rn(y02, e2, y02);
  const f32x2_t q02 = mul2_rn(num2, y2);
  const f32x2_t rem2 = fma2_rn(nchroma2, q02, num2);
  const f32x2_t ratio2 = fma2_rn(rem2, y2, q02);
  const f32x2_t hue2 = fma2_rn(pk2(60.0f, 60.0f), fma2_rn(ratio2, one2, pk2(add[0], add[1])), nz2);
  float hue[2];
  unpk2(hue2, hue[0], hue[1]);
#pragma unroll
  for (int k = 0; k < 2; k++) {
    if (hue[k] < 0.0f) hue[k] = HF_ADDF(hue[k], 360.0f, one);
    if (hue[k] >= 360.0f) hue[k] = HF_ADDF(hue[k], -360.0f, one);
  }
  const f32x2_t s02 = mul2_rn(chroma2, yv2), ns02 = mul2_rn(nchroma2, yv2);
  const f32x2_t sat2 = fma2_rn(fma2_rn(value2, ns02, chroma2), yv2, s02);
  // hue + shift, % 360, negative fix-up
  float t[2];
  unpk2(fma2_rn(pk2(hue[0], hue[1]), one2, pk2(p->hue_shift, p->hue_shift)), t[0], t[1]);
  const f32x2_t a2 = pk2(fabsf(t[0]), fabsf(t[1]));
  const f32x2_t fl2 = fma2_rd(fma2_rn(a2, pk2(0.0027777778f, 0.0027777778f), nz2), one2, pk2(MAGIC, MAGIC));
  const f32x2_t q2 = fma2_rn(fl2, one2, pk2(-MAGIC, -MAGIC));
  float r[2];
  unpk2(fma2_rn(q2, pk2(-360.0f, -360.0f), a2), r[0], r[1]);
#pragma unroll
  for (int k = 0; k < 2; k++) {
    if (r[k] < 0.0f) r[k] = HF_ADDF(r[k], 360.0f, one);
    if (r[k] >= 360.0f) r[k] = HF_ADDF(r[k], -360.0f, one);
    if (t[k] < 0.0f && r[k] > 0.0f) r[k] = HF_SUBF(360.0f, r[k], one);
  }
  // saturation / value
  float sm[2];
  unpk2(fma2_rn(pk2(p->sat_mul, p->sat_mul), sat2, nz2), sm[0], sm[1]);
  const f32x2_t s2 = pk2(HF_ADDF_SAT(sm[0], p->sat_off, one), HF_ADDF_SAT(sm[1], p->sat_off, one));
  const f32x2_t v2 = pk2(vnew[0], vnew[1]);
  // to_rgb
  const f32x2_t c2 = fma2_rn(v2, s2, nz2);
  const f32x2_t h2 = pk2(r[0], r[1]), c60 = pk2(1.0f / 60.0f, 1.0f / 60.0f);
  const f32x2_t q0h = mul2_rn(h2, c60);
  const f32x2_t hp2 = fma2_rn(fma2_rn(q0h, pk2(-60.0f, -60.0f), h2), c60, q0h);
  float yb[2];
  unpk2(fma2_rd(hp2, one2, pk2(MAGIC, MAGIC)), yb[0], yb[1]);
  const unsigned i0 = __float_as_uint(yb[0]) & 7u, i1 = __float_as_uint(yb[1]) & 7u;
  const f32x2_t fk2 = fma2_rn(pk2(__uint_as_float(0x4B000000u | (i0 & 6u)), __uint_as_float(0x4B000000u | (i1 & 6u))), one2, pk2(-MAGIC, -MAGIC));
  const f32x2_t hm2 = fma2_rn(fk2, mone2, hp2);
  float d[2];
  unpk2(fma2_rn(hm2, one2, pk2(-1.0f, -1.0f)), d[0], d[1]);
  const f32x2_t w2 = fma2_rn(pk2(fabsf(d[0]), fabsf(d[1])), mone2, pk2(1.0f, 1.0f));
  const f32x2_t x2 = fma2_rn(c2, w2, nz2);
  const f32x2_t m2 = fma2_rn(c2, mone2, v2);
  const f32x2_t k255 = pk2(255.0f, 255.0f), magic2 = pk2(MAGIC, MAGIC);
  float yc[2], yx[2], ym[2];
  unpk2(fma2_rz(fma2_rn(fma2_rn(c2, one2, m2), k255, nz2), one2, magic2), yc[0], yc[1]);
  unpk2(fma2_rz(fma2_rn(fma2_rn(x2, one2, m2), k255, nz2), one2, magic2), yx[0], yx[1]);
  unpk2(fma2_rz(fma2_rn(m2, k255, nz2), one2, magic2), ym[0], ym[1]);
  const uint32_t w0 = __byte_perm(__float_as_uint(yc[0]), __float_as_uint(yx[0]), 0x7740u);
  const uint32_t w1 = __byte_perm(__float_as_uint(yc[1]), __float_as_uint(yx[1]), 0x7740u);
  out[0] = __byte_perm(w0, __float_as_uint(ym[0]), *(const uint32_t *)((const char *)T->sel + 4u * i0));
  out[1] = __byte_perm(w1, __float_as_uint(ym[1]), *(const uint32_t *)((const char *)T->sel + 4u * i1));
}
#endif  // __CUDACC__

#ifdef __cplusplus
}  // namespace b200vfx
#endif
