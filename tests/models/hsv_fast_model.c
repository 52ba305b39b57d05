/* hsv_fast_model.c -- TEST INFRASTRUCTURE: compiles the product header gst-plugin-rs_b200/csrc/hsv_fast.cuh with gcc
 * (every operator one IEEE binary32 rounding, fmaf for the FMAs) so that tests/test_hsv_fast_model.py can compare the
 * arithmetic of the direct hsvfilter / hsvdetector kernels with the oracle on the CPU: all 2^24 colours for several
 * settings, plus exhaustive checks of the three division replacements.  Not part of the product (which has no CPU path). */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../gst-plugin-rs_b200/csrc/hsv_fast.cuh"

static struct HsvTables g_T;
static volatile float g_one_v = 1.0f;
#define g_one ((float)g_one_v)
static int g_init = 0;
static void init_tables(void) {
  if (g_init) return;
  static const uint32_t sel[8] = HSVF_SEL_INIT;
  for (int i = 0; i < 256; i++) {
    volatile float q = (float)i / 255.0f, r = i ? 255.0f / (float)i : 0.0f;
    g_T.d255[i] = q; g_T.rdiff[i] = r;
  }
  memcpy(g_T.sel, sel, sizeof sel);
  g_init = 1;
}

/* hue/sat/value of every colour idx = r | g<<8 | b<<16 in [lo, hi): out[3*(idx-lo) + {0,1,2}] */
void hfm_from_rgb_range(uint32_t lo, uint32_t hi, float *out) {
  init_tables();
#pragma omp parallel for schedule(static)
  for (uint32_t idx = lo; idx < hi; idx++) {
    const struct HsvF a = hsvf_from_rgb(&g_T, 4u * (idx & 255u), 4u * ((idx >> 8) & 255u), 4u * (idx >> 16), g_one);
    out[3 * (size_t)(idx - lo)] = a.h; out[3 * (size_t)(idx - lo) + 1] = a.s; out[3 * (size_t)(idx - lo) + 2] = a.v;
  }
}

/* hsvfilter of every colour in [lo, hi): out[idx-lo] = r' | g'<<8 | b'<<16 */
void hfm_filter_range(uint32_t lo, uint32_t hi, float hue_shift, float sat_mul, float sat_off, float val_mul, float val_off,
                      uint32_t *out) {
  init_tables();
  const struct HsvFilterParams p = {hue_shift, sat_mul, sat_off, val_mul, val_off};
  const int cls = hsvf_shift_class(hue_shift);
  struct HsvTables T = g_T;
  for (int i = 0; i < 256; i++) T.v2[i] = hsvf_v2_entry(&p, T.d255[i]);
#pragma omp parallel for schedule(static)
  for (uint32_t idx = lo; idx < hi; idx++)
    out[idx - lo] = hsvf_filter_px(&T, &p, cls, 4u * (idx & 255u), 4u * ((idx >> 8) & 255u), 4u * (idx >> 16), g_one);
}

void hfm_detect_range(uint32_t lo, uint32_t hi, float hue_ref, float hue_var, float sat_ref, float sat_var, float val_ref,
                      float val_var, uint8_t *out) {
  init_tables();
  const struct HsvDetectParams p = {hue_ref, hue_var, sat_ref, sat_var, val_ref, val_var};
  volatile float off = 180.0f - hue_ref;
  const int cls = hsvf_shift_class(off);
#pragma omp parallel for schedule(static)
  for (uint32_t idx = lo; idx < hi; idx++)
    out[idx - lo] = (uint8_t)hsvf_detect_px(&g_T, &p, cls, 4u * (idx & 255u), 4u * ((idx >> 8) & 255u), 4u * (idx >> 16), g_one);
}

/* exhaustive: hsvf_div60(h) == h / 60.0f for every f32 with bit pattern in [lo_bits, hi_bits]; returns mismatches */
uint64_t hfm_check_div60(uint32_t lo_bits, uint32_t hi_bits, uint32_t *first_bad) {
  uint64_t bad = 0;
  uint32_t fb = 0xFFFFFFFFu;
#pragma omp parallel for schedule(static) reduction(+ : bad) reduction(min : fb)
  for (uint64_t u = lo_bits; u <= (uint64_t)hi_bits; u++) {
    const float h = hf_float((uint32_t)u);
    volatile float ref = h / 60.0f;
    const float got = hsvf_div60(h);
    if (hf_bits(got) != hf_bits(ref)) { bad++; if ((uint32_t)u < fb) fb = (uint32_t)u; }
  }
  if (first_bad) *first_bad = fb;
  return bad;
}

/* wrap360_small(t) vs fmodf + fix-up for every f32 bit pattern in [lo, hi] (callers pass |t| < 8192 ranges) */
uint64_t hfm_check_wrap360(uint32_t lo_bits, uint32_t hi_bits, uint32_t *first_bad) {
  uint64_t bad = 0;
  uint32_t fb = 0xFFFFFFFFu;
#pragma omp parallel for schedule(static) reduction(+ : bad) reduction(min : fb)
  for (uint64_t u = lo_bits; u <= (uint64_t)hi_bits; u++) {
    const float t = hf_float((uint32_t)u);
    const float a = hsvf_wrap360_small(t, g_one), b = hsvf_wrap360_general(t);
    /* identical bits, except that a zero may carry the other sign (the sign of a zero hue is never observed) */
    if (hf_bits(a) != hf_bits(b) && !(a == 0.0f && b == 0.0f)) { bad++; if ((uint32_t)u < fb) fb = (uint32_t)u; }
  }
  if (first_bad) *first_bad = fb;
  return bad;
}
