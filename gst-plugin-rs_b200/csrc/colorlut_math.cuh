// colorlut_math.cuh -- exact device restatement of apply_1d / apply_3d (video/colorlut/src/colorlut/imp.rs:399-543)
// over tables that are EXACT partial evaluations of the reference arithmetic, built once per LUT:
//
//  * axis table  axis[c][v] = {o0, o1, t}  for every possible channel value v (256 for RGBA, 65536 for RGBA64):
//        pos = clamp(v/denom * scale[c] + offset[c], 0, 1) * (size-1)        (norm_comp*, imp.rs:471-479,439-441)
//        i0 = min(floor(pos), size-1); i1 = min(i0+1, size-1); t = pos - i0  (sample_*,   imp.rs:482-503)
//        o0/o1 = i0/i1 pre-multiplied by the axis stride (1, size, size^2); 1D LUTs use stride 1.
//    The IEEE division, the clamps, floor and the index clamps leave the pixel loop; t is the same f32.
//  * x-pair table pair[x + y*size + z*size^2] = {a.r,a.g,a.b, d.r,d.g,d.b, 0,0} (32 bytes = one L2 sector) with
//        a = lut.at(x,y,z), d = lut.at(min(x+1,size-1),y,z) - a   -- the SAME rounded f32 difference that
//        lerp4's `b - a` produces at run time (imp.rs:528-535), so `a + d*tx` is bit-identical to the x-lerp
//        and the 8 corner fetches become 4 sector-sized 256-bit loads (LDG.E.256).
// Every remaining operator is a single RN operation (__f*_rn, compiled with -fmad=false).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200vfx {

struct __align__(32) LutPair { float a[3]; float d[3]; float pad[2]; };

struct LutDev {
  const LutPair *pair;   // 3D: size^3 x-pair entries
  const float *lut1d;    // 1D: 3 planes of `size` floats r[], g[], b[]
  const uint4 *axis;     // [3][axis_len] {o0, o1, t (bits), 0}
  int axis_len;          // 256 (RGBA); RGBA64 evaluates the axis entry arithmetically (see axis_entry_u16)
  int size;
  int kind;              // 1 | 3
  int ident_domain;      // scale == 1 and offset == +-0 on all channels (no DOMAIN_MIN/MAX in the .cube)
  float scale[3], offset[3];
};

// f32::clamp(0,1): NaN-preserving (imp.rs:473,478)
__device__ __forceinline__ float clamp01_nanpass(float x) {
  x = (x < 0.0f) ? 0.0f : x;
  x = (x > 1.0f) ? 1.0f : x;
  return x;
}

// one axis-table entry, computed with the reference's operator sequence (used only by the table-build kernel)
__device__ __forceinline__ uint4 axis_entry(float value, float denom, float scale, float offset, int size, int stride) {
  const float v = __fdiv_rn(value, denom);
  const float n = clamp01_nanpass(__fadd_rn(__fmul_rn(v, scale), offset));
  const float pos = __fmul_rn(n, __fsub_rn((float)size, 1.0f));
  const int m = size - 1;
  const int i0 = min((int)__float2uint_rd(pos), m);  // floor() as usize: NaN -> 0, saturating
  const int i1 = min(i0 + 1, m);
  const float t = __fsub_rn(pos, (float)i0);
  return make_uint4((uint32_t)(i0 * stride), (uint32_t)(i1 * stride), __float_as_uint(t), 0u);
}

// (v as f32) / 65535.0 for an integer v in [0, 65535], correctly rounded, in 3 instructions: reciprocal multiply
// plus one Newton correction with two FMAs (the same scheme the compiler's IEEE division uses, minus its
// special-case handling).  Verified against true division for all 65536 inputs with exact rational arithmetic
//  and on the GPU by tests/test_gpu_parity.py::test_colorlut_rgba64_every_channel_value.
__device__ __forceinline__ float div65535_exact(float v) {
  const float c = 1.0f / 65535.0f;
  const float q0 = __fmul_rn(v, c);
  const float r = __fmaf_rn(-q0, 65535.0f, v);
  return __fmaf_rn(r, c, q0);
}

// RGBA64: a 65536-entry axis table would cost one L1 wavefront per lane (neighbouring 16-bit values are 17 apart
// on a ramp), so the entry is evaluated per pixel with the reference's operator sequence (norm_comp_u16 and
// sample_* index logic, imp.rs:476-479, 496-503)
//
// Integer<->float conversions go through the 2^23 "magic number" instead of I2F/F2I (those run on the quarter-rate
// XU pipe and were the first bottleneck, profiles/r01_ncu_direct64.txt):
//   u16 -> f32 : bits(0x4B000000 | v) - 2^23                          (exact)
//   floor(pos) : y = pos +(round-down) 2^23; i0 = bits(y) & 0x7FFFFF; (float)i0 = y - 2^23   (exact, 0 <= pos < 2^23)
// NaN pos: the reference takes i0 = 0 and t = NaN; here i0 is some in-range index and t = NaN -- every lerp with a
// NaN weight is NaN, so the quantised output (0) is the same.
// IDENT: default DOMAIN (scale 1, offset +-0): q*1 + (+-0) == q and clamp(q) == q for q in [0,1] -- both dropped.
template <bool IDENT>
__device__ __forceinline__ uint4 axis_entry_u16(unsigned v, float scale, float offset, int size, int stride) {
  const float MAGIC = 8388608.0f;
  const float vf = __fsub_rn(__uint_as_float(0x4B000000u | v), MAGIC);
  const float q = div65535_exact(vf);
  const float n = IDENT ? q : clamp01_nanpass(__fadd_rn(__fmul_rn(q, scale), offset));
  const float pos = __fmul_rn(n, __fsub_rn((float)size, 1.0f));
  const unsigned m = (unsigned)(size - 1);
  const float y = __fadd_rd(pos, MAGIC);
  const unsigned i0 = min(__float_as_uint(y) & 0x007FFFFFu, m);
  const float f0 = __fsub_rn(__uint_as_float(0x4B000000u | i0), MAGIC);   // (float)i0 after the clamp to size-1
  const unsigned i1 = min(i0 + 1u, m);
  return make_uint4(i0 * (unsigned)stride, i1 * (unsigned)stride, __float_as_uint(__fsub_rn(pos, f0)), 0u);
}

// a + (b - a) * t, three roundings (imp.rs:528-535)
__device__ __forceinline__ float lerp_exact(float a, float b, float t) {
  return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), t));
}
// same value with the difference pre-rounded at table-build time
__device__ __forceinline__ float lerp_pre(float a, float d, float t) { return __fadd_rn(a, __fmul_rn(d, t)); }

// (v.clamp(0,1) * MAXV).round() as uN  (imp.rs:537-543).
// round-half-away for q >= 0 equals floor(q + 0.5) evaluated exactly; FADD.RM never rounds up across an
// integer, so floor(fadd_rd(q, .5)) is exact (q = 0.49999997 -> 0, q = 0.5 -> 1).
// NaN: fmaxf(NaN,0) = 0 here, the reference keeps NaN and `as u8` maps it to 0 -- same byte.
template <int MAXV>
__device__ __forceinline__ unsigned quantize_round(float v) {
  const float c = fminf(fmaxf(v, 0.0f), 1.0f);
  const float q = __fmul_rn(c, (float)MAXV);
  // floor() via the 2^23 magic number as well (no F2I): two round-down adds, then the low mantissa bits
  return __float_as_uint(__fadd_rd(__fadd_rd(q, 0.5f), 8388608.0f)) & 0x007FFFFFu;
}

__device__ __forceinline__ void ldg256(const LutPair *p, float (&v)[8]) {
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
      : "l"(p));
}

// apply_1d / apply_3d for one pixel whose channel values are (vr, vg, vb); MAXV = 255 (table axis) or 65535 (inline axis)
template <int MAXV>
__device__ __forceinline__ void colorlut_eval(const LutDev &L, unsigned vr, unsigned vg, unsigned vb, unsigned out[3]) {
  uint4 ax, ay, az;
  if (MAXV == 255) {
    ax = __ldg(L.axis + vr); ay = __ldg(L.axis + L.axis_len + vg); az = __ldg(L.axis + 2 * L.axis_len + vb);
  } else {
    const int s1 = (L.kind == 3) ? L.size : 1, s2 = (L.kind == 3) ? L.size * L.size : 1;
    if (L.ident_domain) {
      ax = axis_entry_u16<true>(vr, 1.0f, 0.0f, L.size, 1);
      ay = axis_entry_u16<true>(vg, 1.0f, 0.0f, L.size, s1);
      az = axis_entry_u16<true>(vb, 1.0f, 0.0f, L.size, s2);
    } else {
      ax = axis_entry_u16<false>(vr, L.scale[0], L.offset[0], L.size, 1);
      ay = axis_entry_u16<false>(vg, L.scale[1], L.offset[1], L.size, s1);
      az = axis_entry_u16<false>(vb, L.scale[2], L.offset[2], L.size, s2);
    }
  }
  const float tx = __uint_as_float(ax.z), ty = __uint_as_float(ay.z), tz = __uint_as_float(az.z);
  if (L.kind == 3) {
    float e00[8], e10[8], e01[8], e11[8];
    const LutPair *base = L.pair + ax.x;
    ldg256(base + ay.x + az.x, e00);   // (x0|x1, y0, z0)
    ldg256(base + ay.y + az.x, e10);   // (x0|x1, y1, z0)
    ldg256(base + ay.x + az.y, e01);   // (x0|x1, y0, z1)
    ldg256(base + ay.y + az.y, e11);   // (x0|x1, y1, z1)
#pragma unroll
    for (int k = 0; k < 3; k++) {      // lerp order x (R) -> y (G) -> z (B), imp.rs:514-525
      const float c00 = lerp_pre(e00[k], e00[3 + k], tx), c10 = lerp_pre(e10[k], e10[3 + k], tx);
      const float c01 = lerp_pre(e01[k], e01[3 + k], tx), c11 = lerp_pre(e11[k], e11[3 + k], tx);
      const float c0 = lerp_exact(c00, c10, ty), c1 = lerp_exact(c01, c11, ty);
      out[k] = quantize_round<MAXV>(lerp_exact(c0, c1, tz));
    }
  } else {  // three independent 1D tables (imp.rs:399-429, 482-490)
    const float *r = L.lut1d, *g = L.lut1d + L.size, *b = L.lut1d + 2 * L.size;
    out[0] = quantize_round<MAXV>(lerp_exact(__ldg(r + ax.x), __ldg(r + ax.y), tx));
    out[1] = quantize_round<MAXV>(lerp_exact(__ldg(g + ay.x), __ldg(g + ay.y), ty));
    out[2] = quantize_round<MAXV>(lerp_exact(__ldg(b + az.x), __ldg(b + az.y), tz));
  }
}

}  // namespace b200vfx
