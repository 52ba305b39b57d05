// kernels.cuh -- sm_100a kernels of the per-pixel hot path (one kernel per element).
//
// All of these are HBM-streaming integer/byte kernels: no dense contraction, hence no tensor
// cores.  What matters: coalesced full-sector accesses, enough loads in
// flight per thread, grids that fill 148 SMs, and keeping the per-pixel instruction count
// under the issue budget that the HBM roofline leaves (~55 thread-instr per 8 B pixel).
#pragma once
#include "pixel_math.cuh"
#include "tma_pipe.cuh"

namespace b200vfx {

// streaming (read-once / write-once) accesses: keep them from displacing the L2-resident tables
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream_u32(uint32_t *p, uint32_t v) { __stcs(p, v); }
__device__ __forceinline__ uint2 ld_stream_u2(const uint2 *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream_u2(uint2 *p, uint2 v) { __stcs(p, v); }
// Layout of the 2^24-entry answer tables.  Blocked (default): a 128-byte line = a 4(r) x 4(g) x 2(b) colour block, so
// pixels whose colours differ by sensor noise in all three channels share L1 lines instead of spreading over one line
// per (g,b) pair.  Linear (-DB200VFX_MEMO_BLOCKED=0): index = r | g<<8 | b<<16 (a line = 32 consecutive r).
// Measured (profiles/r01_memo_layout.txt): "natural" frames 20.4 -> 16.6 us, frame A/B unchanged within 3 %.
#ifndef B200VFX_MEMO_BLOCKED
#define B200VFX_MEMO_BLOCKED 1
#endif
#ifndef B200VFX_EXP_MEMO_SHIFT   // timing experiment only (wrong results): table footprint 64 MiB >> shift
#define B200VFX_EXP_MEMO_SHIFT 0
#endif
__device__ __forceinline__ uint32_t memo_index_full(uint32_t c);
__device__ __forceinline__ uint32_t memo_index(uint32_t c) { return memo_index_full(c) >> B200VFX_EXP_MEMO_SHIFT; }
__device__ __forceinline__ uint32_t memo_index_full(uint32_t c) {  // c = r | g<<8 | b<<16
#if B200VFX_MEMO_BLOCKED
  return (c & 3u) | ((c >> 6) & 0xCu) | ((c >> 12) & 0x10u) | ((c << 3) & 0x7E0u) | ((c << 1) & 0x1F800u) | (c & 0xFE0000u);
#else
  return c;
#endif
}
// PDL: let the next (independent) frame's kernel start as soon as every CTA of this one has been scheduled
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// ... and (persistent, capped grids) do not RETIRE before the previous kernel of the stream has completed: such kernels
// complete in launch order however deeply consecutive frames overlap, and their CTAs hold their SM slots until then, which
// bounds how many of them can be in flight (pdl_admit, b200vfx.cu).  No-op for a normal launch.  Costs ~3 us of completion
// latency per kernel, hidden for large frames, not for 640x480 ones -- hence only for capped grids.
#ifndef B200VFX_EXP_NO_TAIL_WAIT   // timing experiment only
#define B200VFX_EXP_NO_TAIL_WAIT 0
#endif
// (row, chunk) of a work item of the persistent kernels.  Packed frames are flattened to ONE row by the launchers, so the
// common case needs no division at all; otherwise a 32-bit division serves every frame below 2^31 items.  (The generic
// 64-bit division is a ~100-instruction subroutine call -- per item and thread it cost 10-25 instructions per pixel.)
__device__ __forceinline__ void item_row_chunk(long long item, int chunks_x, int height, int &row, int &cx) {
  if (height == 1) { row = 0; cx = (int)item; }
  else if (item < (1LL << 31)) {
    const unsigned it = (unsigned)item, r = it / (unsigned)chunks_x;
    row = (int)r; cx = (int)(it - r * (unsigned)chunks_x);
  } else {
    row = (int)(item / chunks_x); cx = (int)(item - (long long)row * chunks_x);
  }
}
__device__ __forceinline__ void pdl_wait_prior() {
#if !B200VFX_EXP_NO_TAIL_WAIT
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// --------------------------------------------------------------------------------------------
// colorlut table builders (once per LUT): axis tables and, on the host, the x-pair table
// --------------------------------------------------------------------------------------------
struct AxisBuildParams { int size, kind, axis_len; float denom; float scale[3], offset[3]; };

__global__ void colorlut_axis_table_kernel(uint4 *axis, AxisBuildParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // 0 .. 3*axis_len-1
  if (i >= 3 * p.axis_len) return;
  const int c = i / p.axis_len, v = i - c * p.axis_len;
  const int stride = (p.kind == 3) ? (c == 0 ? 1 : (c == 1 ? p.size : p.size * p.size)) : 1;
  axis[i] = axis_entry((float)v, p.denom, p.scale[c], p.offset[c], p.size, stride);
}

// --------------------------------------------------------------------------------------------
// colorlut direct evaluation (RGBA64 always; RGBA when mode = direct): one pixel per thread,
// rows looped by a persistent-style grid.   imp.rs:237-397
// FMT 0 = RGBA (u8), 1 = RGBA64_LE, 2 = RGBA64_BE.  ALIGNED: one 32/64-bit access per pixel.
// --------------------------------------------------------------------------------------------
template <int FMT, bool ALIGNED>
__global__ void __launch_bounds__(256) colorlut_direct_kernel(LutDev L, const uint8_t *__restrict__ src,
                                                              long sstride, uint8_t *__restrict__ dst,
                                                              long dstride, int width, int height) {
  pdl_trigger();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= width) return;
  for (int row = blockIdx.y; row < height; row += gridDim.y) {
    if (FMT == 0) {
      const uint8_t *s = src + (size_t)row * sstride + (size_t)x * 4;
      uint8_t *d = dst + (size_t)row * dstride + (size_t)x * 4;
      uint32_t px;
      if (ALIGNED) px = ld_stream_u32(reinterpret_cast<const uint32_t *>(s));
      else px = (uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16) | ((uint32_t)s[3] << 24);
      unsigned o[3];
      colorlut_eval<255>(L, px & 255u, (px >> 8) & 255u, (px >> 16) & 255u, o);
      const uint32_t out = o[0] | (o[1] << 8) | (o[2] << 16) | (px & 0xFF000000u);  // d[3] = s[3]
      if (ALIGNED) st_stream_u32(reinterpret_cast<uint32_t *>(d), out);
      else { d[0] = (uint8_t)out; d[1] = (uint8_t)(out >> 8); d[2] = (uint8_t)(out >> 16); d[3] = (uint8_t)(out >> 24); }
    } else {
      // rows are addressed in u16 units = stride/2 (imp.rs:317-318)
      const uint8_t *s = src + (size_t)row * ((sstride / 2) * 2) + (size_t)x * 8;
      uint8_t *d = dst + (size_t)row * ((dstride / 2) * 2) + (size_t)x * 8;
      uint2 px;
      if (ALIGNED) px = ld_stream_u2(reinterpret_cast<const uint2 *>(s));
      else {
        const uint16_t *s16 = reinterpret_cast<const uint16_t *>(s);
        px.x = (uint32_t)s16[0] | ((uint32_t)s16[1] << 16);
        px.y = (uint32_t)s16[2] | ((uint32_t)s16[3] << 16);
      }
      uint32_t w0 = px.x, w1 = px.y;
      if (FMT == 2) { w0 = __byte_perm(w0, 0, 0x2301); w1 = __byte_perm(w1, 0, 0x2301); }  // from_be on an LE device
      unsigned o[3];
      colorlut_eval<65535>(L, w0 & 0xFFFFu, w0 >> 16, w1 & 0xFFFFu, o);
      uint32_t o0 = o[0] | (o[1] << 16), o1 = o[2];
      if (FMT == 2) { o0 = __byte_perm(o0, 0, 0x2301); o1 = __byte_perm(o1, 0, 0x2301); }
      o1 = (o1 & 0xFFFFu) | (px.y & 0xFFFF0000u);  // alpha word copied raw, never byte-swapped (imp.rs:345,394)
      if (ALIGNED) st_stream_u2(reinterpret_cast<uint2 *>(d), make_uint2(o0, o1));
      else {
        uint16_t *d16 = reinterpret_cast<uint16_t *>(d);
        d16[0] = (uint16_t)o0; d16[1] = (uint16_t)(o0 >> 16); d16[2] = (uint16_t)o1; d16[3] = (uint16_t)(o1 >> 16);
      }
    }
  }
}

// RGBA64 with a 3D LUT, vector path: a thread owns FOUR consecutive pixels (one 256-bit load, one 256-bit store: a warp moves
// 1 KB per instruction) and keeps the LUT cell's corner entries in registers across them (colorlut_eval64_cached).
// Needs 32-byte aligned rows and width % 4 == 0; everything else takes colorlut_direct_kernel.
struct U256 { unsigned long long a, b, c, d; };
__device__ __forceinline__ U256 ld_stream_256(const void *p) {
  U256 v;
  asm volatile("ld.global.cs.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream_256(void *p, const U256 &v) {
  asm volatile("st.global.cs.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(v.a), "l"(v.b), "l"(v.c), "l"(v.d) : "memory");
}
template <int FMT>   // 1 = RGBA64_LE, 2 = RGBA64_BE
__global__ void __launch_bounds__(256) colorlut_direct64x4_kernel(LutDev L, const uint8_t *__restrict__ src, long sstride,
                                                                  uint8_t *__restrict__ dst, long dstride, int width, int height) {
  pdl_trigger();
  // (an item-persistent variant with the next item's pixels prefetched, like the lookup kernels, measured 41.6 us against
  // the 39.0 us of this plain row loop -- 76 -> 64 forced registers and the extra moves cost more than the prefetch hides)
  const int g = blockIdx.x * blockDim.x + threadIdx.x;   // group of 4 pixels
  if (4 * g >= width) return;
  CellCache cache;
  cell_cache_reset(cache);
  for (int row = blockIdx.y; row < height; row += gridDim.y) {
    const U256 in = ld_stream_256(src + (size_t)row * sstride + (size_t)g * 32);
    const unsigned long long w[4] = {in.a, in.b, in.c, in.d};
    unsigned long long o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t w0 = (uint32_t)w[k], w1 = (uint32_t)(w[k] >> 32);
      const uint32_t alpha = w1 & 0xFFFF0000u;           // alpha word copied raw, never byte-swapped (imp.rs:345,394)
      if (FMT == 2) { w0 = __byte_perm(w0, 0, 0x2301); w1 = __byte_perm(w1, 0, 0x2301); }
      unsigned q[3];
      colorlut_eval64_cached(L, w0 & 0xFFFFu, w0 >> 16, w1 & 0xFFFFu, cache, q);
      uint32_t o0 = q[0] | (q[1] << 16), o1 = q[2];
      if (FMT == 2) { o0 = __byte_perm(o0, 0, 0x2301); o1 = __byte_perm(o1, 0, 0x2301); }
      o[k] = (unsigned long long)o0 | ((unsigned long long)((o1 & 0xFFFFu) | alpha) << 32);
    }
    st_stream_256(dst + (size_t)row * dstride + (size_t)g * 32, U256{o[0], o[1], o[2], o[3]});
  }
}

// --------------------------------------------------------------------------------------------
// colorlut memoisation for 8-bit RGBA.  apply_1d/apply_3d are pure functions of the 24-bit
// colour, and `location` is only mutable in READY (imp.rs:72-76), so the element evaluates
// them ONCE per start() for all 2^24 colours with the exact direct evaluator above and keeps
// the 64 MiB answer table resident in B200's 126 MB L2.  Bit-exact by construction.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colorlut_memo_build_kernel(LutDev L, uint32_t *__restrict__ memo) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;  // idx = r | g<<8 | b<<16
  unsigned o[3];
  colorlut_eval<255>(L, idx & 255u, (idx >> 8) & 255u, idx >> 16, o);
  memo[memo_index(idx)] = o[0] | (o[1] << 8) | (o[2] << 16);
}

// per-channel 256-entry answer tables for a 1D LUT (768 bytes, lives in shared memory)
__global__ void colorlut_memo1d_build_kernel(LutDev L, uint8_t *__restrict__ memo1d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 768) return;
  const int c = i >> 8;
  const uint4 a = __ldg(L.axis + i);  // axis_len == 256 here
  const float *tab = L.lut1d + c * L.size;
  memo1d[i] = (uint8_t)quantize_round<255>(lerp_exact(__ldg(tab + a.x), __ldg(tab + a.y), __uint_as_float(a.z)));
}

// The pixel loop for RGBA with a memo table: out = memo[px & 0xFFFFFF] | (px & 0xFF000000).
// A warp owns 32*PX consecutive pixels; lane L touches pixels L, L+32, ... so every gather
// instruction covers 32 CONSECUTIVE pixels (spatially coherent colours -> few distinct L1
// lines per gather) while the frame loads/stores stay fully coalesced 128 B per instruction.
template <int PX>
__global__ void __launch_bounds__(256) colorlut_memo_apply_kernel(const uint32_t *__restrict__ memo,
                                                                  const uint8_t *__restrict__ src, long sstride,
                                                                  uint8_t *__restrict__ dst, long dstride,
                                                                  int width, int height, int linger) {
  pdl_trigger();
  // Persistent: the grid is capped at a few CTAs per SM (option "memo_ctas") and every CTA walks over work items
  // (item = row x 2048*PX/8-pixel chunk).  All CTAs are resident at once, so the NEXT frame's kernel (PDL) starts
  // immediately and runs beside this one: a gather-bound frame and an HBM-bound frame then share the SMs.
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunks_x = (width + 8 * 32 * PX - 1) / (8 * 32 * PX);
  const long long items = (long long)chunks_x * height;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    int row, cx;
    item_row_chunk(item, chunks_x, height, row, cx);
    const int x0 = (cx * 8 + warp) * (32 * PX) + lane;
    if (x0 - lane >= width) continue;
    const uint32_t *s = reinterpret_cast<const uint32_t *>(src + (size_t)row * sstride);
    uint32_t *d = reinterpret_cast<uint32_t *>(dst + (size_t)row * dstride);
    uint32_t px[PX], o[PX];
#pragma unroll
    for (int k = 0; k < PX; k++) px[k] = (x0 + 32 * k < width) ? ld_stream_u32(s + x0 + 32 * k) : 0u;
#pragma unroll
    for (int k = 0; k < PX; k++) o[k] = __ldg(memo + memo_index(px[k] & 0x00FFFFFFu));
#pragma unroll
    for (int k = 0; k < PX; k++)
      if (x0 + 32 * k < width) st_stream_u32(d + x0 + 32 * k, o[k] | (px[k] & 0xFF000000u));
  }
  if (linger) pdl_wait_prior();   // capped grids only (large frames): see pdl_admit in b200vfx.cu
}

// 1D LUT: three 256-byte tables in shared memory.  Item-persistent like the 3D kernel above (item = row x chunk of
// 8 warps x 32*PX pixels): any 1-D grid covers any frame, whether it was flattened to one long row or not.
template <int PX>
__global__ void __launch_bounds__(256) colorlut_memo1d_apply_kernel(const uint8_t *__restrict__ memo1d,
                                                                    const uint8_t *__restrict__ src, long sstride,
                                                                    uint8_t *__restrict__ dst, long dstride,
                                                                    int width, int height, int linger) {
  pdl_trigger();
  __shared__ uint8_t tab[768];
  for (int i = threadIdx.x; i < 768 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t *>(tab)[i] = __ldg(reinterpret_cast<const uint32_t *>(memo1d) + i);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunks_x = (width + 8 * 32 * PX - 1) / (8 * 32 * PX);
  const long long items = (long long)chunks_x * height;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    int row, cx;
    item_row_chunk(item, chunks_x, height, row, cx);
    const int x0 = (cx * 8 + warp) * (32 * PX) + lane;
    if (x0 - lane >= width) continue;
    const uint32_t *s = reinterpret_cast<const uint32_t *>(src + (size_t)row * sstride);
    uint32_t *d = reinterpret_cast<uint32_t *>(dst + (size_t)row * dstride);
    uint32_t px[PX];
#pragma unroll
    for (int k = 0; k < PX; k++) px[k] = (x0 + 32 * k < width) ? ld_stream_u32(s + x0 + 32 * k) : 0u;
#pragma unroll
    for (int k = 0; k < PX; k++) {
      const uint32_t r = tab[px[k] & 255u], g = tab[256 + ((px[k] >> 8) & 255u)], b = tab[512 + ((px[k] >> 16) & 255u)];
      if (x0 + 32 * k < width) st_stream_u32(d + x0 + 32 * k, r | (g << 8) | (b << 16) | (px[k] & 0xFF000000u));
    }
  }
  if (linger) pdl_wait_prior();
}

// generic byte-addressed fallback for rows that are not 4-byte aligned
__global__ void colorlut_memo_apply_bytes_kernel(const uint32_t *__restrict__ memo, const uint8_t *__restrict__ memo1d,
                                                 const uint8_t *__restrict__ src, long sstride,
                                                 uint8_t *__restrict__ dst, long dstride, int width, int height) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= width) return;
  for (int row = blockIdx.y; row < height; row += gridDim.y) {
    const uint8_t *s = src + (size_t)row * sstride + (size_t)x * 4;
    uint8_t *d = dst + (size_t)row * dstride + (size_t)x * 4;
    const uint32_t r = s[0], g = s[1], b = s[2];
    if (memo) {
      const uint32_t o = __ldg(memo + memo_index(r | (g << 8) | (b << 16)));
      d[0] = (uint8_t)o; d[1] = (uint8_t)(o >> 8); d[2] = (uint8_t)(o >> 16);
    } else {
      d[0] = __ldg(memo1d + r); d[1] = __ldg(memo1d + 256 + g); d[2] = __ldg(memo1d + 512 + b);
    }
    d[3] = s[3];
  }
}

// --------------------------------------------------------------------------------------------
// hsvfilter (in place) and hsvdetector.   hsvfilter/imp.rs:76-120, hsvdetector/imp.rs:100-160
// BPP 3|4; COFF = byte offset of the first colour byte (1 for xRGB/ARGB/xBGR/ABGR); BGR = byte order.
// 4-bpp pixels move as one 32-bit word when the rows are 4-byte aligned.
// --------------------------------------------------------------------------------------------
// Memoised variants (MEMO = true): hsvfilter/hsvdetector are pure functions of (settings, 24-bit colour).  Once the
// same settings have processed 2^24 pixels (as much arithmetic as evaluating every colour once), the element builds
//   hsvfilter   : u32[2^24]  answer table  memo[r|g<<8|b<<16] = r'|g'<<8|b'<<16   (64 MiB, L2 resident)
//   hsvdetector : 2^24-bit hit bitmap (2 MiB)
// with the exact per-pixel evaluator below and the frame kernel becomes a table lookup (bit-exact by construction).
__global__ void __launch_bounds__(256) hsvfilter_memo_build_kernel(HsvFilterSettings st, int cls, uint32_t *__restrict__ memo) {
  __shared__ HsvTables T;
  fill_hsv_tables(&T, &st);
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < (1u << 24); idx += gridDim.x * blockDim.x) {
    unsigned r4, g4, b4;
    bytes_x4(idx, r4, g4, b4);
    memo[memo_index(idx)] = hsvf_filter_px(&T, &st, cls, r4, g4, b4, 1.0f);
  }
}

__global__ void __launch_bounds__(256) hsvdetector_bitmap_build_kernel(HsvDetectSettings st, int cls, uint32_t *__restrict__ bitmap) {
  __shared__ HsvTables T;
  fill_hsv_tables(&T);
  // 32 consecutive colours per warp -> one bitmap word; the grid-stride keeps whole warps together (2^24 % (grid*256) == 0)
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < (1u << 24); idx += gridDim.x * blockDim.x) {
    const bool hit = hsvdetect_px(st, &T, cls, idx & 255u, (idx >> 8) & 255u, idx >> 16);
    const unsigned word = __ballot_sync(0xFFFFFFFFu, hit);
    if ((threadIdx.x & 31) == 0) bitmap[idx >> 5] = word;
  }
}

template <int BPP, int COFF, bool BGR, bool ALIGNED, bool MEMO>
__global__ void __launch_bounds__(256) hsvfilter_kernel(HsvFilterSettings st, int cls, const uint32_t *__restrict__ memo,
                                                        uint8_t *__restrict__ data, long stride, int width, int height) {
  __shared__ HsvTables T;
  if (!MEMO) fill_hsv_tables(&T, &st);
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= width) return;
  for (int row = blockIdx.y; row < height; row += gridDim.y) {
    uint8_t *p = data + (size_t)row * stride + (size_t)x * BPP;
    uint32_t px = 0;
    unsigned c0, c1, c2;
    if (BPP == 4 && ALIGNED) {
      px = *reinterpret_cast<const uint32_t *>(p);
      c0 = (px >> (8 * COFF)) & 255u; c1 = (px >> (8 * COFF + 8)) & 255u; c2 = (px >> (8 * COFF + 16)) & 255u;
    } else {
      c0 = p[COFF]; c1 = p[COFF + 1]; c2 = p[COFF + 2];
    }
    unsigned r = BGR ? c2 : c0, g = c1, b = BGR ? c0 : c2;
    if (MEMO) {
      const uint32_t v = __ldg(memo + memo_index(r | (g << 8) | (b << 16)));
      r = v & 255u; g = (v >> 8) & 255u; b = (v >> 16) & 255u;
    } else {
      hsvfilter_px(st, &T, cls, r, g, b);
    }
    c0 = BGR ? b : r; c1 = g; c2 = BGR ? r : b;
    if (BPP == 4 && ALIGNED) {
      const uint32_t keep = COFF ? (px & 0x000000FFu) : (px & 0xFF000000u);
      *reinterpret_cast<uint32_t *>(p) = keep | (c0 << (8 * COFF)) | (c1 << (8 * COFF + 8)) | (c2 << (8 * COFF + 16));
    } else {
      p[COFF] = (uint8_t)c0; p[COFF + 1] = (uint8_t)c1; p[COFF + 2] = (uint8_t)c2;
    }
  }
}

// IBPP/ICOFF/IBGR describe the input pixel; OCOFF/OBGR the 4-byte output pixel (alpha at 3 if OCOFF==0 else 0)
template <int IBPP, int ICOFF, bool IBGR, int OCOFF, bool OBGR, bool ALIGNED, bool MEMO>
__global__ void __launch_bounds__(256) hsvdetector_kernel(HsvDetectSettings st, int cls, const uint32_t *__restrict__ bitmap,
                                                          const uint8_t *__restrict__ src, long sstride,
                                                          uint8_t *__restrict__ dst, long dstride, int width, int height) {
  if (MEMO) pdl_trigger();
  __shared__ HsvTables T;
  if (!MEMO) fill_hsv_tables(&T);
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= width) return;
  for (int row = blockIdx.y; row < height; row += gridDim.y) {
    const uint8_t *ip = src + (size_t)row * sstride + (size_t)x * IBPP;
    uint8_t *op = dst + (size_t)row * dstride + (size_t)x * 4;
    unsigned c0, c1, c2;
    if (IBPP == 4 && ALIGNED) {
      const uint32_t px = ld_stream_u32(reinterpret_cast<const uint32_t *>(ip));
      c0 = (px >> (8 * ICOFF)) & 255u; c1 = (px >> (8 * ICOFF + 8)) & 255u; c2 = (px >> (8 * ICOFF + 16)) & 255u;
    } else {
      c0 = ip[ICOFF]; c1 = ip[ICOFF + 1]; c2 = ip[ICOFF + 2];
    }
    const unsigned r = IBGR ? c2 : c0, g = c1, b = IBGR ? c0 : c2;
    bool hit;
    if (MEMO) {
      const uint32_t idx = r | (g << 8) | (b << 16);
      hit = (__ldg(bitmap + (idx >> 5)) >> (idx & 31u)) & 1u;
    } else {
      hit = hsvdetect_px(st, &T, cls, r, g, b);
    }
    const unsigned a = hit ? 255u : 0u;
    const unsigned o0 = OBGR ? b : r, o1 = g, o2 = OBGR ? r : b;
    const uint32_t out = (o0 << (8 * OCOFF)) | (o1 << (8 * OCOFF + 8)) | (o2 << (8 * OCOFF + 16)) | (OCOFF ? a : (a << 24));
    if (ALIGNED) st_stream_u32(reinterpret_cast<uint32_t *>(op), out);
    else { op[0] = (uint8_t)out; op[1] = (uint8_t)(out >> 8); op[2] = (uint8_t)(out >> 16); op[3] = (uint8_t)(out >> 24); }
  }
}

// --------------------------------------------------------------------------------------------
// Generic 4-byte -> 4-byte table-lookup map (memoised hsvfilter / hsvdetector on 4-bpp formats): same thread
// mapping as colorlut_memo_apply_kernel -- a warp owns 32*PX consecutive pixels, every gather instruction covers 32
// consecutive pixels, PX gathers in flight per thread.  In place (src == dst) is fine: a thread only rewrites its
// own pixels.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t swap_c0_c2(uint32_t c) { return __byte_perm(c, 0u, 0x4012); }  // (c0,c1,c2,_) -> (c2,c1,c0,0)

template <int COFF, bool BGR>
struct HsvFilterMemoOp {  // memo is keyed and valued in R,G,B byte order
  const uint32_t *memo;
  __device__ __forceinline__ uint32_t operator()(uint32_t px) const {
    uint32_t c = (px >> (8 * COFF)) & 0x00FFFFFFu;
    if (BGR) c = swap_c0_c2(c);
    uint32_t v = __ldg(memo + memo_index(c));
    if (BGR) v = swap_c0_c2(v);
    const uint32_t keep = COFF ? (px & 0x000000FFu) : (px & 0xFF000000u);   // x / alpha byte untouched
    return keep | (v << (8 * COFF));
  }
};

template <int ICOFF, bool IBGR, int OCOFF, bool OBGR>
struct HsvDetectBitmapOp {  // bitmap bit index = r | g<<8 | b<<16
  const uint32_t *bitmap;
  __device__ __forceinline__ uint32_t operator()(uint32_t px) const {
    const uint32_t c = (px >> (8 * ICOFF)) & 0x00FFFFFFu;
    const uint32_t idx = IBGR ? swap_c0_c2(c) : c;
    const uint32_t hit = (__ldg(bitmap + (idx >> 5)) >> (idx & 31u)) & 1u;
    const uint32_t oc = (IBGR == OBGR) ? c : swap_c0_c2(c);
    const uint32_t a = hit ? (OCOFF ? 0x000000FFu : 0xFF000000u) : 0u;
    return (oc << (8 * OCOFF)) | a;
  }
};

// colorlut on any 4-byte 8-bit format, input and output byte orders independent: the byte permutation a `videoconvert`
// either side of the element would do is folded into the lookup kernel's load and store (one PRMT each).
//   in_sel : source pixel -> (r, g, b, 0) table key;   out_sel: {table value (bytes 0-2), source pixel | 0xFF.. (bytes 4-7)}
//   -> destination pixel; or_mask sets the 4th byte to 255 when the source has no alpha to copy.
template <bool LUT3D>
struct ColorLutFmtOp {
  const uint32_t *memo;      // 3D: 2^24 answers
  const uint8_t *memo1d;     // 1D: 3 x 256 answers
  uint32_t in_sel, out_sel, src_or;
  __device__ __forceinline__ uint32_t operator()(uint32_t px) const {
    const uint32_t c = __byte_perm(px, 0u, in_sel);
    uint32_t v;
    if (LUT3D) v = __ldg(memo + memo_index(c));
    else v = (uint32_t)__ldg(memo1d + (c & 255u)) | ((uint32_t)__ldg(memo1d + 256 + ((c >> 8) & 255u)) << 8) |
             ((uint32_t)__ldg(memo1d + 512 + (c >> 16)) << 16);
    return __byte_perm(v, px | src_or, out_sel);
  }
};

template <typename Op, int PX>
__global__ void __launch_bounds__(256) map_u32_kernel(Op op, const uint8_t *__restrict__ src, long sstride,
                                                      uint8_t *__restrict__ dst, long dstride, int width, int height,
                                                      int linger) {
  pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunks_x = (width + 8 * 32 * PX - 1) / (8 * 32 * PX);   // persistent like colorlut_memo_apply_kernel
  const long long items = (long long)chunks_x * height;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    int row, cx;
    item_row_chunk(item, chunks_x, height, row, cx);
    const int x0 = (cx * 8 + warp) * (32 * PX) + lane;
    if (x0 - lane >= width) continue;
    const uint32_t *s = reinterpret_cast<const uint32_t *>(src + (size_t)row * sstride);
    uint32_t *d = reinterpret_cast<uint32_t *>(dst + (size_t)row * dstride);
    uint32_t px[PX], o[PX];
#pragma unroll
    for (int k = 0; k < PX; k++) px[k] = (x0 + 32 * k < width) ? ld_stream_u32(s + x0 + 32 * k) : 0u;
#pragma unroll
    for (int k = 0; k < PX; k++) o[k] = op(px[k]);
#pragma unroll
    for (int k = 0; k < PX; k++)
      if (x0 + 32 * k < width) st_stream_u32(d + x0 + 32 * k, o[k]);
  }
  if (linger) pdl_wait_prior();
}

// --------------------------------------------------------------------------------------------
// DIRECT hsvfilter / hsvdetector on 4-byte pixels (what frames get while properties are being animated and no
// answer table can pay for itself): the map_u32_kernel thread mapping with the exact branch-free arithmetic of
// hsv_fast.cuh and its 2 KB of shared-memory tables.  Issue-bound (~100 instructions per pixel): PX independent
// pixels per thread give the scheduler ILP across the dependent FMA chains.
// --------------------------------------------------------------------------------------------
template <int COFF, bool BGR, int CLS>   // CLS = hsvf_shift_class (0: proven fast code, 1: general code), fixed per launch
struct HsvFilterDirectOp {
  HsvFilterSettings st;
  float one;     // 1.0f, opaque to the compiler: keeps the adds that hsv_fast.cuh routes to the FMA pipe as FFMAs
  float nzero;   // -0.0f, opaque: the addend that makes a packed FMA a correctly rounded product (hsvf_filter_px2)
  static constexpr int cls = CLS;
  static constexpr bool has_pair = CLS == 0;
  __device__ __forceinline__ const HsvFilterParams *filter_params() const { return &st; }
  // two pixels at once, FP work in f32x2 lanes (shift class 0 only)
  __device__ __forceinline__ void pair(const HsvTables *T, uint32_t pa, uint32_t pb, uint32_t &oa, uint32_t &ob) const {
    const uint32_t ca = (pa >> (8 * COFF)) & 0x00FFFFFFu, cb = (pb >> (8 * COFF)) & 0x00FFFFFFu;
    unsigned a0, a1, a2, b0, b1, b2;
    bytes_x4(ca, a0, a1, a2);
    bytes_x4(cb, b0, b1, b2);
    const unsigned r4[2] = {BGR ? a2 : a0, BGR ? b2 : b0}, g4[2] = {a1, b1}, b4[2] = {BGR ? a0 : a2, BGR ? b0 : b2};
    uint32_t v[2];
    hsvf_filter_px2(T, &st, r4, g4, b4, one, nzero, v);
    if (BGR) { v[0] = swap_c0_c2(v[0]); v[1] = swap_c0_c2(v[1]); }
    oa = (COFF ? (pa & 0x000000FFu) : (pa & 0xFF000000u)) | (v[0] << (8 * COFF));
    ob = (COFF ? (pb & 0x000000FFu) : (pb & 0xFF000000u)) | (v[1] << (8 * COFF));
  }
  __device__ __forceinline__ uint32_t operator()(const HsvTables *T, uint32_t px) const {
    const uint32_t c = (px >> (8 * COFF)) & 0x00FFFFFFu;
    unsigned c0, c1, c2;
    bytes_x4(c, c0, c1, c2);
    uint32_t v = hsvf_filter_px(T, &st, cls, BGR ? c2 : c0, c1, BGR ? c0 : c2, one);   // r | g<<8 | b<<16
    if (BGR) v = swap_c0_c2(v);
    const uint32_t keep = COFF ? (px & 0x000000FFu) : (px & 0xFF000000u);        // x / alpha byte untouched
    return keep | (v << (8 * COFF));
  }
};
template <int ICOFF, bool IBGR, int OCOFF, bool OBGR, int CLS>
struct HsvDetectDirectOp {
  HsvDetectSettings st;
  float one;
  static constexpr int cls = CLS;
  static constexpr bool has_pair = false;
  __device__ __forceinline__ void pair(const HsvTables *, uint32_t, uint32_t, uint32_t &, uint32_t &) const {}
  __device__ __forceinline__ const HsvFilterParams *filter_params() const { return nullptr; }
  __device__ __forceinline__ uint32_t operator()(const HsvTables *T, uint32_t px) const {
    const uint32_t c = (px >> (8 * ICOFF)) & 0x00FFFFFFu;
    unsigned c0, c1, c2;
    bytes_x4(c, c0, c1, c2);
    const bool hit = hsvf_detect_px(T, &st, cls, IBGR ? c2 : c0, c1, IBGR ? c0 : c2, one) != 0;
    const uint32_t oc = (IBGR == OBGR) ? c : swap_c0_c2(c);
    const uint32_t a = hit ? (OCOFF ? 0x000000FFu : 0xFF000000u) : 0u;
    return (oc << (8 * OCOFF)) | a;
  }
};

template <typename Op, int PX>
__global__ void __launch_bounds__(256) hsv_direct_map_kernel(Op op, const uint8_t *__restrict__ src, long sstride,
                                                             uint8_t *__restrict__ dst, long dstride, int width, int height) {
  __shared__ HsvTables T;
  fill_hsv_tables(&T, op.filter_params());
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunks_x = (width + 8 * 32 * PX - 1) / (8 * 32 * PX);
  const long long items = (long long)chunks_x * height;
  // (a software-prefetched variant -- next item's pixels loaded before this item's arithmetic -- measured 6 % SLOWER:
  // the kernel is issue-bound with 64 resident warps per SM, the extra registers and moves cost more than the
  // exposed load latency they hide)
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    int row, cx;
    item_row_chunk(item, chunks_x, height, row, cx);
    const int x0 = (cx * 8 + warp) * (32 * PX) + lane;
    if (x0 - lane >= width) continue;
    const uint32_t *s = reinterpret_cast<const uint32_t *>(src + (size_t)row * sstride);
    uint32_t *d = reinterpret_cast<uint32_t *>(dst + (size_t)row * dstride);
    uint32_t px[PX], o[PX];
#pragma unroll
    for (int k = 0; k < PX; k++) px[k] = (x0 + 32 * k < width) ? ld_stream_u32(s + x0 + 32 * k) : 0u;
    if (Op::has_pair && (PX % 2) == 0) {
#pragma unroll
      for (int k = 0; k < PX; k += 2) op.pair(&T, px[k], px[k + 1], o[k], o[k + 1]);
    } else {
#pragma unroll
      for (int k = 0; k < PX; k++) o[k] = op(&T, px[k]);
    }
#pragma unroll
    for (int k = 0; k < PX; k++)
      if (x0 + 32 * k < width) st_stream_u32(d + x0 + 32 * k, o[k]);
  }
}

// --------------------------------------------------------------------------------------------
// 3-byte pixels (RGB / BGR) through the same answer tables.  A warp owns 32*G groups of 4 pixels (= 3 words each):
// the 96*G words are loaded with fully coalesced 128-byte instructions, staged in a per-warp shared-memory tile and
// re-read at a stride of 3 words (3 is coprime with the 32 banks: conflict-free), so lane L holds 4 whole pixels per
// group and issues 4*G independent gathers.  OUT4 = false: 3-byte result written back the same way (hsvfilter, in
// place; bytes of the last word that lie beyond the row are read and written back unchanged).  OUT4 = true: 4-byte
// output pixels, one 16-byte store per group (hsvdetector RGB/BGR input).  Rows must be 4-byte aligned.
// --------------------------------------------------------------------------------------------
template <typename Op, int G, bool OUT4, bool DST16>
__global__ void __launch_bounds__(256) map_rgb24_kernel(Op op, const uint8_t *__restrict__ src, long sstride,
                                                        uint8_t *dst, long dstride, int width, int height) {
  pdl_trigger();
  __shared__ uint32_t tile_all[8][96 * G];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t *tile = tile_all[warp];
  const int ngroups = (width + 3) >> 2, nwords = (3 * width + 3) >> 2;
  const int chunk = blockIdx.x * 8 + warp;             // one chunk of 32*G groups per warp
  const int g0 = chunk * 32 * G;
  if (g0 >= ngroups) return;
  const int w0 = g0 * 3;
  for (int row = blockIdx.y; row < height; row += gridDim.y) {
    const uint32_t *s = reinterpret_cast<const uint32_t *>(src + (size_t)row * sstride);
#pragma unroll
    for (int j = 0; j < 3 * G; j++) {
      const int wi = w0 + lane + 32 * j;
      tile[lane + 32 * j] = wi < nwords ? ld_stream_u32(s + wi) : 0u;
    }
    __syncwarp();
    uint32_t px[G][4], o[G][4];
#pragma unroll
    for (int i = 0; i < G; i++) {
      const int q = lane + 32 * i;
      const uint32_t a = tile[3 * q], b = tile[3 * q + 1], c = tile[3 * q + 2];
      px[i][0] = a & 0x00FFFFFFu;
      px[i][1] = (a >> 24) | ((b & 0xFFFFu) << 8);
      px[i][2] = (b >> 16) | ((c & 0xFFu) << 16);
      px[i][3] = c >> 8;
    }
#pragma unroll
    for (int i = 0; i < G; i++)
#pragma unroll
      for (int k = 0; k < 4; k++) o[i][k] = op(px[i][k]);
    if (OUT4) {
      uint8_t *drow = dst + (size_t)row * dstride;
#pragma unroll
      for (int i = 0; i < G; i++) {
        const int x = 4 * (g0 + lane + 32 * i);
        if (x + 3 < width && DST16) {
          __stcs(reinterpret_cast<uint4 *>(drow + (size_t)x * 4), make_uint4(o[i][0], o[i][1], o[i][2], o[i][3]));
        } else {
#pragma unroll
          for (int k = 0; k < 4; k++)
            if (x + k < width) st_stream_u32(reinterpret_cast<uint32_t *>(drow + (size_t)(x + k) * 4), o[i][k]);
        }
      }
      __syncwarp();
    } else {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < G; i++) {
        const int q = lane + 32 * i, x = 4 * (g0 + q);
        uint32_t r[4];
#pragma unroll
        for (int k = 0; k < 4; k++) r[k] = (x + k < width) ? (o[i][k] & 0x00FFFFFFu) : px[i][k];   // bytes beyond the row: unchanged
        tile[3 * q] = r[0] | (r[1] << 24);
        tile[3 * q + 1] = (r[1] >> 8) | (r[2] << 16);
        tile[3 * q + 2] = (r[2] >> 16) | (r[3] << 8);
      }
      __syncwarp();
      uint32_t *d = reinterpret_cast<uint32_t *>(dst + (size_t)row * dstride);
#pragma unroll
      for (int j = 0; j < 3 * G; j++) {
        const int wi = w0 + lane + 32 * j;
        if (wi < nwords) st_stream_u32(d + wi, tile[lane + 32 * j]);
      }
      __syncwarp();
    }
  }
}

// --------------------------------------------------------------------------------------------
// videocompare / blockhash block sums.  hashed_image.rs:24-64 -> image_hasher blockhash fast path.
// One CTA reduces a (bw x rows) tile that lies inside ONE hash block: registers -> warp shuffle
// -> shared -> a single atomicAdd on the bin.  Integer adds are order independent => exact.
// VEC: RGBA rows 16-byte aligned and bw % 4 == 0 -> uint4 loads, R+G+B by one dp4a per pixel.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) { return __reduce_add_sync(0xFFFFFFFFu, v); }

// Up to kBlockhashMaxFrames equally sized frames per launch (videocompare hashes the reference pad's frame and every
// other pad's frame on each aggregate tick): blockIdx.z = frame * zchunks + row chunk, sums of frame f at
// sums + f*hw*hh.  The bins are zeroed by blockhash_zero_kernel, this kernel's PDL primary: it only has to wait for
// it (griddepcontrol.wait) right before its single atomicAdd.
constexpr int kBlockhashMaxFrames = 8;
struct BlockhashFrames { const uint8_t *src[kBlockhashMaxFrames]; long stride[kBlockhashMaxFrames]; };

// ONE launch: every CTA publishes its partial sum, the last CTA to arrive (ticket) adds the partials of each bin in a
// fixed order and writes the final sums -- no zeroing pass, no atomics on the bins, nothing to wait for.
template <int BPP, bool VEC>
__global__ void __launch_bounds__(128) blockhash_sums_kernel(const __grid_constant__ BlockhashFrames fr, int zchunks, int bw, int bh, int hw, int hh,
                                                             int rows_per_cta, uint32_t *__restrict__ sums,
                                                             uint32_t *__restrict__ partials, unsigned *__restrict__ ticket) {
  pdl_trigger();   // the next launch may start reading ITS frame; whatever it writes is ordered by the wait below
  const int bx = blockIdx.x, by = blockIdx.y;
  const int f = (int)blockIdx.z / zchunks, chunk = (int)blockIdx.z - f * zchunks;
  const uint8_t *__restrict__ src = fr.src[f];
  const long stride = fr.stride[f];
  const int y0 = by * bh + chunk * rows_per_cta;
  const int y1 = min(y0 + rows_per_cta, (by + 1) * bh);
  uint32_t acc = 0;
  if (VEC) {  // BPP == 4
    const int n4 = bw >> 2;
    constexpr int U = 8;  // rows in flight per thread: 8 independent 16-byte loads (16 in one batch measured slower: 14.5 vs 11.9 us)
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const uint8_t *col = src + (size_t)bx * bw * 4 + (size_t)i * 16;
      for (int y = y0; y < y1; y += U) {
        uint4 q[U];
#pragma unroll
        for (int r = 0; r < U; r++)
          q[r] = (y + r < y1) ? __ldcs(reinterpret_cast<const uint4 *>(col + (size_t)(y + r) * stride)) : make_uint4(0xFF000000u, 0xFF000000u, 0xFF000000u, 0xFF000000u);
#pragma unroll
        for (int r = 0; r < U; r++) {  // out-of-range rows load an opaque black pixel: contributes 0
          acc += (q[r].x >> 24) ? __dp4a(q[r].x, 0x00010101u, 0u) : 765u;   // A==0 counts as white (765)
          acc += (q[r].y >> 24) ? __dp4a(q[r].y, 0x00010101u, 0u) : 765u;
          acc += (q[r].z >> 24) ? __dp4a(q[r].z, 0x00010101u, 0u) : 765u;
          acc += (q[r].w >> 24) ? __dp4a(q[r].w, 0x00010101u, 0u) : 765u;
        }
      }
    }
  } else {
    for (int y = y0; y < y1; y++) {
      const uint8_t *row = src + (size_t)y * stride + (size_t)bx * bw * BPP;
      for (int i = threadIdx.x; i < bw; i += blockDim.x) {
        const uint8_t *p = row + (size_t)i * BPP;
        uint32_t s = (uint32_t)p[0] + p[1] + p[2];
        if (BPP == 4 && p[3] == 0) s = 765u;
        acc += s;
      }
    }
  }
  __shared__ uint32_t wsum[4];
  __shared__ int s_last;
  // launched with programmatic serialisation: the frame was read while the previous launch on the stream drained; the
  // shared scratch (partials, ticket) and the sums are touched only once that launch has completed
  pdl_wait_prior();
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  const int nbins = hw * hh;                       // per frame
  const unsigned total = gridDim.x * gridDim.y * gridDim.z;
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += wsum[w];
    // partial of (frame f, chunk, bin): chunk-major inside a frame so the final reduction reads contiguous bins
    partials[((size_t)f * zchunks + chunk) * nbins + by * hw + bx] = t;
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == total - 1u;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    const int nframes = (int)gridDim.z / zchunks;
    for (int i = threadIdx.x; i < nframes * nbins; i += blockDim.x) {
      const int ff = i / nbins, bin = i - ff * nbins;
      uint32_t t = 0;
      for (int c = 0; c < zchunks; c++) t += __ldcg(partials + ((size_t)ff * zchunks + c) * nbins + bin);
      sums[i] = t;
    }
    if (threadIdx.x == 0) *ticket = 0u;            // re-armed for the next launch
  }
}

// Row-streaming variant for 16-byte aligned RGBA: a CTA owns `rows_per_cta` WHOLE rows inside one row of hash blocks and
// reads them left to right -- every DRAM page is read once, sequentially, by one CTA -- instead of one CTA per hash block
// reading bw-pixel row segments 'stride' bytes apart.  A thread keeps K fixed columns (16 bytes each, one hash block each)
// over 16/K rows per batch = 16 independent loads in flight; per-column sums go to shared bins once per column group
// (match.any + redux: one shared atomic per distinct block in a warp).  grid = (chunks, hh, frames), 256 threads.
constexpr int kBlockhashRowsMaxHW = 256;
template <int K>
__global__ void __launch_bounds__(256) blockhash_rows_kernel(const __grid_constant__ BlockhashFrames fr, int zchunks, int bw, int bh, int hw, int hh,
                                                             int rows_per_cta, uint32_t *__restrict__ sums,
                                                             uint32_t *__restrict__ partials, unsigned *__restrict__ ticket) {
  constexpr int U = 16 / K;
  __shared__ uint32_t bins[kBlockhashRowsMaxHW];
  __shared__ int s_last;
  pdl_trigger();
  const int chunk = blockIdx.x, by = blockIdx.y, f = blockIdx.z, tid = threadIdx.x, lane = tid & 31;
  const uint8_t *__restrict__ src = fr.src[f];
  const long stride = fr.stride[f];
  const int y0 = by * bh + chunk * rows_per_cta;
  const int y1 = min(y0 + rows_per_cta, (by + 1) * bh);
  for (int i = tid; i < hw; i += 256) bins[i] = 0u;
  __syncthreads();
  const int n4blk = bw >> 2, n4row = n4blk * hw;
  const uint4 black = make_uint4(0xFF000000u, 0xFF000000u, 0xFF000000u, 0xFF000000u);   // opaque black: contributes 0
  for (int i0 = 0; i0 < n4row; i0 += 256 * K) {
    uint32_t acc[K];
    bool okc[K];
    const uint8_t *col[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
      const int i = i0 + tid + 256 * k;
      acc[k] = 0u; okc[k] = i < n4row;
      col[k] = src + (size_t)(okc[k] ? i : 0) * 16;
    }
    for (int y = y0; y < y1; y += U) {
      uint4 q[U][K];
#pragma unroll
      for (int r = 0; r < U; r++)
#pragma unroll
        for (int k = 0; k < K; k++)
          q[r][k] = (okc[k] && y + r < y1) ? __ldcs(reinterpret_cast<const uint4 *>(col[k] + (size_t)(y + r) * stride)) : black;
#pragma unroll
      for (int r = 0; r < U; r++)
#pragma unroll
        for (int k = 0; k < K; k++) {
          const uint4 v = q[r][k];
          acc[k] += (v.x >> 24) ? __dp4a(v.x, 0x00010101u, 0u) : 765u;   // A==0 counts as white (765)
          acc[k] += (v.y >> 24) ? __dp4a(v.y, 0x00010101u, 0u) : 765u;
          acc[k] += (v.z >> 24) ? __dp4a(v.z, 0x00010101u, 0u) : 765u;
          acc[k] += (v.w >> 24) ? __dp4a(v.w, 0x00010101u, 0u) : 765u;
        }
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
      const int b = okc[k] ? (i0 + tid + 256 * k) / n4blk : -1;
      const unsigned grp = __match_any_sync(0xFFFFFFFFu, b);
      const uint32_t t = __reduce_add_sync(grp, acc[k]);
      if (b >= 0 && lane == __ffs(grp) - 1) atomicAdd(&bins[b], t);
    }
  }
  __syncthreads();
  pdl_wait_prior();                                // scratch and sums: only after the previous launch has completed
  const int nbins = hw * hh;                       // per frame
  // partial of (frame f, chunk, bin): the layout blockhash_sums_kernel uses
  for (int i = tid; i < hw; i += 256) partials[((size_t)f * zchunks + chunk) * nbins + by * hw + i] = bins[i];
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y * gridDim.z - 1u;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    const int nframes = (int)gridDim.z;
    for (int i = tid; i < nframes * nbins; i += 256) {
      const int ff = i / nbins, bin = i - ff * nbins;
      uint32_t t = 0;
      for (int c = 0; c < zchunks; c++) t += __ldcg(partials + ((size_t)ff * zchunks + c) * nbins + bin);
      sums[i] = t;
    }
    if (tid == 0) *ticket = 0u;                    // re-armed for the next launch
  }
}

// --------------------------------------------------------------------------------------------
// roundedcorners A8 mask.  border/imp.rs:57-180 (cairo fill + 1px stroke), restated analytically:
// only the four r x r corner boxes are partially covered; coverage on a 16x16 integer sample grid.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned mul_un8(unsigned a, unsigned b) {  // pixman MUL_UN8
  const unsigned t = a * b + 0x80u;
  return ((t >> 8) + t) >> 8;
}

// a8 points at absolute row y0 of the plane; `pitch` is the byte distance between rows of a8,
// `stride` the number of bytes per row that belong to the plane (A420 stride[3]).
__global__ void __launch_bounds__(256) roundmask_kernel(uint8_t *__restrict__ a8, long pitch, int width, int height,
                                                        int stride, int y0, int rows, int r) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= stride || (int)blockIdx.y >= rows) return;
  const int y = y0 + (int)blockIdx.y;
  uint8_t v;
  if (r < 0) v = 0xFF;                       // border-radius-px == 0: alpha_mem.fill(0xff) (:123-128)
  else if (x >= width || y >= height) v = 0; // pre-zeroed memory outside the surface (:130)
  else {
    const int i = (x < r) ? (r - 1 - x) : ((x >= width - r) ? (x - (width - r)) : -1);
    const int j = (y < r) ? (r - 1 - y) : ((y >= height - r) ? (y - (height - r)) : -1);
    if (i < 0 || j < 0) v = 0xFF;
    else {
      const long long S = 16;
      const long long rf = 2 * S * r, ro = rf + S, ri = rf - S;
      const long long R2f = rf * rf, R2o = ro * ro, R2i = ri * ri;
      int nf = 0, ns = 0;
      for (int b = 0; b < 16; b++) {
        const long long dy = 2 * S * j + 2 * b + 1;
        for (int a = 0; a < 16; a++) {
          const long long dx = 2 * S * i + 2 * a + 1;
          const long long d2 = dx * dx + dy * dy;
          nf += (d2 <= R2f);
          ns += (d2 <= R2o && d2 >= R2i);
        }
      }
      const unsigned af = (unsigned)((nf * 255 + 128) / 256), as = (unsigned)((ns * 255 + 128) / 256);
      v = (uint8_t)min(as + mul_un8(af, 255u - as), 255u);
    }
  }
  a8[(size_t)blockIdx.y * pitch + x] = v;
}

}  // namespace b200vfx
