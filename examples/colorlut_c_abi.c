/* colorlut_c_abi.c -- the drop-in boundary used from plain C, the way the Rust shim (or any FFI) uses it:
 * what ColorLut::start / transform_frame / stop do in video/colorlut/src/colorlut/imp.rs:168-224, through libb200vfx.
 *
 *   gcc -std=c99 -Iinclude examples/colorlut_c_abi.c -o /tmp/colorlut_c_abi -Lgst-plugin-rs_b200/lib -lb200vfx -Wl,-rpath,$PWD/gst-plugin-rs_b200/lib
 *   /tmp/colorlut_c_abi file.cube 1920 1080
 *
 * Prints the FNV-1a checksum of the output frame (tests/test_examples.py compares it with the oracle's). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200vfx.h"

static uint64_t pcg_inc;
static uint32_t pcg32_next(uint64_t *state) { /* the generator of b200vfx.synth.pcg32: PCG32 XSH-RR */
  uint64_t old = *state;
  *state = old * 6364136223846793005ULL + pcg_inc;
  uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
  uint32_t rot = (uint32_t)(old >> 59u);
  return (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
}

int main(int argc, char **argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s file.cube width height\n", argv[0]); return 2; }
  const int w = atoi(argv[2]), h = atoi(argv[3]);
  b200vfx_ctx *ctx = NULL;
  if (b200vfx_ctx_create(&ctx, -1) != B200VFX_OK) { fprintf(stderr, "no CUDA device: %s\n", b200vfx_last_error(NULL)); return 1; }
  if (b200vfx_colorlut_load_file(ctx, argv[1]) != B200VFX_OK) {           /* start(): ResourceError::Read */
    fprintf(stderr, "cannot load LUT: %s\n", b200vfx_last_error(ctx));
    b200vfx_ctx_destroy(ctx);
    return 1;
  }
  const size_t stride = (size_t)w * 4, bytes = stride * (size_t)h;
  uint8_t *src = (uint8_t *)b200vfx_host_alloc(bytes), *dst = (uint8_t *)b200vfx_host_alloc(bytes);   /* pinned, as a GstAllocator would */
  if (!src || !dst) { fprintf(stderr, "host alloc failed\n"); return 1; }
  const uint64_t seed = 0x5EED0002ULL;   /* pcg32_srandom(seed, seed) */
  uint64_t st = 0;
  pcg_inc = (seed << 1) | 1u;
  st = st * 6364136223846793005ULL + pcg_inc;
  st += seed;
  st = st * 6364136223846793005ULL + pcg_inc;
  for (size_t i = 0; i < bytes; i += 4) { uint32_t v = pcg32_next(&st); memcpy(src + i, &v, 4); }
  memset(dst, 0, bytes);
  for (int frame = 0; frame < 3; frame++)                                   /* transform_frame(), one call per buffer */
    if (b200vfx_colorlut_process(ctx, B200VFX_FORMAT_RGBA, w, h, src, (int)stride, dst, (int)stride) != B200VFX_OK) {
      fprintf(stderr, "process failed: %s\n", b200vfx_last_error(ctx));
      return 1;
    }
  uint32_t fnv_in = 2166136261u, fnv_out = 2166136261u;
  for (size_t i = 0; i < bytes; i++) { fnv_in = (fnv_in ^ src[i]) * 16777619u; fnv_out = (fnv_out ^ dst[i]) * 16777619u; }
  printf("%dx%d in=%08x out=%08x launches=%llu\n", w, h, fnv_in, fnv_out, (unsigned long long)b200vfx_ctx_kernel_launches(ctx));
  b200vfx_host_free(src);
  b200vfx_host_free(dst);
  b200vfx_colorlut_clear(ctx);                                              /* stop() */
  b200vfx_ctx_destroy(ctx);
  return 0;
}
