// pixel_math.cuh -- device glue for the exact (bit-for-bit) restatement of the reference's per-pixel HSV f32 arithmetic
// (hsv_fast.cuh; colorlut lives in colorlut_math.cuh).  Every operator is a single IEEE-754 binary32 round-to-nearest-even
// operation, spelled with __f*_rn intrinsics so ptxas can never contract a mul+add into an FMA (the file is also compiled
// with -fmad=false); the only FMAs are the explicit residual steps of the proven division replacements.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "colorlut_math.cuh"
#include "hsv_fast.cuh"

namespace b200vfx {

// ---------------------------------------------------------------------------------------------
// hsvutils / hsvfilter / hsvdetector per-pixel arithmetic: hsv_fast.cuh (shared with the CPU model in tests/models/,
// where it is verified against the oracle over all 2^24 colours).  Here: the shared-memory tables it reads.
// ---------------------------------------------------------------------------------------------
typedef HsvFilterParams HsvFilterSettings;
typedef HsvDetectParams HsvDetectSettings;

// fills the per-CTA tables (IEEE divisions, once per CTA): d255[i] = i/255, rdiff[i] = 255/i, sector selectors and --
// for hsvfilter (fp != nullptr) -- the new value of every possible old value
__device__ __forceinline__ void fill_hsv_tables(HsvTables *T, const HsvFilterParams *fp = nullptr) {
  constexpr uint32_t sel[8] = HSVF_SEL_INIT;
  constexpr unsigned long long lo = sel[0] | (sel[1] << 16) | ((unsigned long long)(sel[2] | (sel[3] << 16)) << 32);
  constexpr unsigned long long hi = sel[4] | (sel[5] << 16) | ((unsigned long long)(sel[6] | (sel[7] << 16)) << 32);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const float d = __fdiv_rn((float)i, 255.0f);
    T->d255[i] = d;
    T->rdiff[i] = i ? __fdiv_rn(255.0f, (float)i) : 0.0f;
    if (fp) T->v2[i] = hsvf_v2_entry(fp, d);
    if (i < 8) T->sel[i] = (uint32_t)(((i < 4) ? lo : hi) >> (16 * (i & 3))) & 0xFFFFu;
  }
  __syncthreads();
}

// the three colour bytes of c = c0 | c1<<8 | c2<<16 as table BYTE offsets (4 * value)
__device__ __forceinline__ void bytes_x4(uint32_t c, unsigned &o0, unsigned &o1, unsigned &o2) {
  o0 = (c << 2) & 0x3FCu; o1 = (c >> 6) & 0x3FCu; o2 = (c >> 14) & 0x3FCu;
}

// hsv_filter body (hsvfilter/imp.rs:100-117) on the three colour bytes; cls = hsvf_shift_class(hue_shift) (uniform)
__device__ __forceinline__ void hsvfilter_px(const HsvFilterSettings &s, const HsvTables *T, int cls, unsigned &r, unsigned &g,
                                             unsigned &b) {
  const uint32_t o = hsvf_filter_px(T, &s, cls, 4u * r, 4u * g, 4u * b, 1.0f);
  r = o & 255u; g = (o >> 8) & 255u; b = o >> 16;
}
// hsv_detect predicate (hsvdetector/imp.rs:139-157); cls = hsvf_shift_class(180 - hue_ref)
__device__ __forceinline__ bool hsvdetect_px(const HsvDetectSettings &s, const HsvTables *T, int cls, unsigned r, unsigned g,
                                             unsigned b) {
  return hsvf_detect_px(T, &s, cls, 4u * r, 4u * g, 4u * b, 1.0f) != 0;
}

}  // namespace b200vfx
