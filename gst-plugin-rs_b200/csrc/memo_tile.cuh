// memo_tile.cuh -- table-lookup map kernel with a per-tile shared-memory copy of the colour sub-cube it needs.
//
// Why: the plain lookup kernel (colorlut_memo_apply_kernel / map_u32_kernel) is bound by the L1 gather path as soon as
// the 32 pixels of a gather touch many 128-byte table lines: every distinct line is one L1 wavefront (~1 per SM
// clock).  Measured on 4K frames: coherent ramps ~3 lines per gather -> HBM-bound (11.9 us), ramps with +-3 sensor
// noise ~20 lines -> 17.4 us = exactly 20 wavefronts x 8.3M/32 gathers / 148 SMs / 1.965 GHz, random pixels 32 lines
// -> 35 us.  Shared memory serves 32 different addresses per clock, the table (64 MiB) does not fit it -- but the
// colours of a spatially compact tile of real video do: a CTA owns a 64 x 64 pixel tile, reduces the per-channel
// min/max of its pixels (byte-SIMD in registers, shuffles, one shared round), and if the bounding colour cube has at
// most kTileCap entries it copies exactly that sub-cube of the answer table into shared memory (neighbouring threads
// fetch neighbouring r: ~1 L1 wavefront per 16-32 entries instead of ~20 per 32 pixels) and serves all 4096 lookups
// from there.  Tiles whose cube is too large (edges, texture, random pixels) use the direct gather -- decided per tile,
// results identical either way (the shared copy holds the very same table words).
//
// Round 2 tried to make this the default for "natural" content and could not: a persistent, software-pipelined variant
// (3 CTAs per SM walking over tiles, next tile's pixels in flight, direct gathers for tiny boxes) measured SLOWER than this
// one-tile-per-CTA kernel on every content but random pixels (ramps 18.2 vs 14.0 us, +-3: 22.4 vs 18.0, +-8: 25.2 vs 20.9;
// noise 48 vs 71), and both lose to the plain lookup kernel up to +-3 (16.6 us): the min/max reduction, the two barriers
// and the 64-pixel-wide access pattern cost 2.5-5 us per frame, more than the L1 wavefronts they save
// (profiles/r02_memo_tile_experiment.jsonl).  The kernel stays an option (memo_tile) for noisy sources (+-5 .. +-10).
#pragma once
#include "kernels.cuh"

namespace b200vfx {

constexpr int kTileW = 64, kTileH = 64;          // pixels per CTA tile: 16 uint4 per row, 16 rows per pass, 4 passes
constexpr int kTileCap = 10240;                  // sub-cube entries held in shared memory (40 KB -> 5 CTAs per SM)
constexpr int kTileFillBatch = 8;                // independent table loads in flight per thread while filling

// COFF/BGR describe where the 24-bit colour sits in the 4-byte pixel (as in HsvFilterMemoOp); the table is keyed and
// valued in R,G,B order.  The byte that is not colour (alpha / x) is copied.
template <int COFF, bool BGR>
__global__ void __launch_bounds__(256) memo_tile_kernel(const uint32_t *__restrict__ memo, const uint8_t *__restrict__ src,
                                                        long sstride, uint8_t *__restrict__ dst, long dstride, int width,
                                                        int height) {
  pdl_trigger();
  __shared__ uint32_t cache[kTileCap];
  __shared__ uint4 s_red[8];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int x = blockIdx.x * kTileW + 4 * (t & 15);
  const int y0 = blockIdx.y * kTileH + (t >> 4);
  uint4 px[4];
  bool ok[4];
  // per-byte min/max as two u16x2 words each (bytes 0,2 and bytes 1,3): VIMNMX3.U16x2 is one instruction for two
  // pixels, the byte-SIMD __vminu4/__vmaxu4 are ~12-instruction emulations on sm_100
  uint32_t mn_e = 0x00FF00FFu, mn_o = 0x00FF00FFu, mx_e = 0u, mx_o = 0u;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int y = y0 + 16 * k;
    ok[k] = x < width && y < height;          // width % 4 == 0: a uint4 is entirely inside or outside the row
    px[k] = ok[k] ? __ldcs(reinterpret_cast<const uint4 *>(src + (size_t)y * sstride + (size_t)x * 4)) : make_uint4(0u, 0u, 0u, 0u);
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (!ok[k]) continue;
    const uint32_t e0 = __byte_perm(px[k].x, 0u, 0x4240), o0 = __byte_perm(px[k].x, 0u, 0x4341);
    const uint32_t e1 = __byte_perm(px[k].y, 0u, 0x4240), o1 = __byte_perm(px[k].y, 0u, 0x4341);
    const uint32_t e2 = __byte_perm(px[k].z, 0u, 0x4240), o2 = __byte_perm(px[k].z, 0u, 0x4341);
    const uint32_t e3 = __byte_perm(px[k].w, 0u, 0x4240), o3 = __byte_perm(px[k].w, 0u, 0x4341);
    mn_e = __vimin3_u16x2(mn_e, __vminu2(e0, e1), __vminu2(e2, e3));
    mn_o = __vimin3_u16x2(mn_o, __vminu2(o0, o1), __vminu2(o2, o3));
    mx_e = __vimax3_u16x2(mx_e, __vmaxu2(e0, e1), __vmaxu2(e2, e3));
    mx_o = __vimax3_u16x2(mx_o, __vmaxu2(o0, o1), __vmaxu2(o2, o3));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn_e = __vminu2(mn_e, __shfl_xor_sync(0xFFFFFFFFu, mn_e, o));
    mn_o = __vminu2(mn_o, __shfl_xor_sync(0xFFFFFFFFu, mn_o, o));
    mx_e = __vmaxu2(mx_e, __shfl_xor_sync(0xFFFFFFFFu, mx_e, o));
    mx_o = __vmaxu2(mx_o, __shfl_xor_sync(0xFFFFFFFFu, mx_o, o));
  }
  if (lane == 0) { s_red[warp] = make_uint4(mn_e, mn_o, mx_e, mx_o); }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 8; w++) {
    const uint4 r = s_red[w];
    mn_e = __vminu2(mn_e, r.x); mn_o = __vminu2(mn_o, r.y); mx_e = __vmaxu2(mx_e, r.z); mx_o = __vmaxu2(mx_o, r.w);
  }
  // back to one word of 4 bytes: byte0 | byte1 << 8 | byte2 << 16 | byte3 << 24
  const uint32_t mn = mn_e | (mn_o << 8), mx = mx_e | (mx_o << 8);
  // colour bytes -> (r, g, b) ranges
  const uint32_t cmn = (mn >> (8 * COFF)) & 0x00FFFFFFu, cmx = (mx >> (8 * COFF)) & 0x00FFFFFFu;
  const int lo0 = cmn & 255u, lo1 = (cmn >> 8) & 255u, lo2 = cmn >> 16;
  const int d0 = (int)(cmx & 255u) - lo0 + 1, d1 = (int)((cmx >> 8) & 255u) - lo1 + 1, d2 = (int)(cmx >> 16) - lo2 + 1;
  const int vol = d0 * d1 * d2;   // <= 0 only for a tile without any pixel (mn > mx): nothing to do then
  const bool cached = vol > 0 && vol <= kTileCap;
  if (cached) {
    // copy the sub-cube: entry i <-> (c0, c1, c2) = (lo0 + i % d0, lo1 + (i / d0) % d1, lo2 + i / (d0*d1)); the divisions
    // are exact multiply-high's by 2^32/d + 1 (i < 2^14, d <= 2^8)
    const uint32_t inv0 = 0xFFFFFFFFu / (uint32_t)d0 + 1u, inv1 = 0xFFFFFFFFu / (uint32_t)d1 + 1u;   // wraps to 0 for d == 1
    // kTileFillBatch independent table loads in flight per thread before the first shared-memory store (a one-load-per-
    // iteration loop pays one L2 round trip per 256 entries)
    for (int base = t; base < vol; base += 256 * kTileFillBatch) {
      uint32_t fv[kTileFillBatch];
#pragma unroll
      for (int j = 0; j < kTileFillBatch; j++) {
        const int i = base + 256 * j;
        const uint32_t q0 = d0 == 1 ? (uint32_t)i : __umulhi((uint32_t)i, inv0), c0 = (uint32_t)i - q0 * (uint32_t)d0;
        const uint32_t q1 = d1 == 1 ? q0 : __umulhi(q0, inv1), c1 = q0 - q1 * (uint32_t)d1;
        const uint32_t cc = (uint32_t)(lo0 + c0) | ((uint32_t)(lo1 + c1) << 8) | ((uint32_t)(lo2 + q1) << 16);   // in pixel byte order
        fv[j] = (i < vol) ? __ldg(memo + memo_index(BGR ? swap_c0_c2(cc) : cc)) : 0u;
      }
#pragma unroll
      for (int j = 0; j < kTileFillBatch; j++)
        if (base + 256 * j < vol) cache[base + 256 * j] = fv[j];
    }
  }
  __syncthreads();
  uint32_t v[4][4];
  if (cached) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t p[4] = {px[k].x, px[k].y, px[k].z, px[k].w};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t c = (p[j] >> (8 * COFF)) & 0x00FFFFFFu;
        const int idx = (int)(c & 255u) - lo0 + d0 * ((int)((c >> 8) & 255u) - lo1 + d1 * ((int)(c >> 16) - lo2));
        v[k][j] = ok[k] ? cache[idx] : 0u;
      }
    }
  } else {   // 16 independent gathers in flight per thread
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t p[4] = {px[k].x, px[k].y, px[k].z, px[k].w};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t c = (p[j] >> (8 * COFF)) & 0x00FFFFFFu;
        v[k][j] = __ldg(memo + memo_index(BGR ? swap_c0_c2(c) : c));   // out-of-frame lanes read entry 0: harmless
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (!ok[k]) continue;
    const uint32_t p[4] = {px[k].x, px[k].y, px[k].z, px[k].w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t vv = BGR ? swap_c0_c2(v[k][j]) : v[k][j];
      const uint32_t keep = COFF ? (p[j] & 0x000000FFu) : (p[j] & 0xFF000000u);
      o[j] = keep | (vv << (8 * COFF));
    }
    const int y = y0 + 16 * k;
    __stcs(reinterpret_cast<uint4 *>(dst + (size_t)y * dstride + (size_t)x * 4), make_uint4(o[0], o[1], o[2], o[3]));
  }
}

}  // namespace b200vfx
