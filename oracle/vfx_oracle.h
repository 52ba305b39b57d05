/*
 * vfx_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the per-pixel hot path of sdroege/gst-plugin-rs
 * (video/colorlut, video/hsv, video/videofx).  Only tests/, the smoke check in
 * __graft_entry__.py and the cpu_baseline / --impl reference legs of bench.py
 * may load this library.  The product (gst-plugin-rs_b200/csrc) never links it.
 *
 * PARITY PINNING STATUS (see DESIGN.md "Oracle"):
 *   - .cube parser       : pinned by the reference's own unit tests
 *                          (video/colorlut/src/parser.rs:382-473).
 *   - hsvutils           : pinned by video/hsv/src/hsvutils.rs:237-279.
 *   - colorlut pixel math, hsvfilter / hsvdetector element behaviour:
 *                          "parity unpinned" by the reference (it has no tests
 *                          and the Rust toolchain is absent here); pinned
 *                          instead against an independent numpy-f32
 *                          restatement and SURVEY Appendix C vectors.
 *   - blockhash bit rule, rounded-corner mask, colordetect palette / colour
 *     name: third-party arithmetic (image_hasher 3.1.1, cairo,
 *                          color-thief 0.2.2, color-name 1.2.0) absent from
 *                          /root/reference -> "parity unpinned".
 *
 * All arithmetic is IEEE-754 binary32, one rounding per operator, never fused:
 * build with -ffp-contract=off -fno-fast-math and no -march that enables FMA.
 */
#ifndef VFX_ORACLE_H
#define VFX_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* format codes shared with include/b200vfx.h (same numeric values) */
enum {
  ORC_FMT_RGBX = 0, ORC_FMT_XRGB = 1, ORC_FMT_BGRX = 2, ORC_FMT_XBGR = 3,
  ORC_FMT_RGBA = 4, ORC_FMT_ARGB = 5, ORC_FMT_BGRA = 6, ORC_FMT_ABGR = 7,
  ORC_FMT_RGB = 8, ORC_FMT_BGR = 9, ORC_FMT_RGBA64_LE = 10, ORC_FMT_RGBA64_BE = 11,
  ORC_FMT_I420 = 12, ORC_FMT_A420 = 13
};

/* ---- .cube parser (video/colorlut/src/parser.rs:105-375) ------------------
 * values: malloc'ed n*3 floats in file order ([r,g,b] per data line); caller
 * frees with orc_free().  kind: 1 = LUT_1D, 3 = LUT_3D.
 * returns 0 ok, -1 InvalidLut (err filled), -2 Io error (err filled). */
int orc_cube_parse(const char *text, size_t len, int *kind, int *size,
                   float **values, float scale[3], float offset[3],
                   char *err, size_t errlen);
int orc_cube_parse_file(const char *path, int *kind, int *size, float **values,
                        float scale[3], float offset[3], char *err, size_t errlen);
void orc_free(void *p);

/* ---- colorlut (video/colorlut/src/colorlut/imp.rs:226-543) ----------------
 * values as produced by the parser: 1D -> size x [r,g,b]; 3D -> size^3 x
 * [r,g,b], index x + y*size + z*size^2 with x<-R fastest.
 * fmt: ORC_FMT_RGBA, ORC_FMT_RGBA64_LE, ORC_FMT_RGBA64_BE. */
int orc_colorlut_apply(int kind, int size, const float *values,
                       const float scale[3], const float offset[3], int fmt,
                       int width, int height, const uint8_t *src, int sstride,
                       uint8_t *dst, int dstride, int threads);

/* ---- hsvutils (video/hsv/src/hsvutils.rs:42-198) --------------------------*/
void orc_hsv_from_rgb(const uint8_t in_p[3], float hsv[3]);
void orc_hsv_from_bgr(const uint8_t in_p[3], float hsv[3]);
void orc_hsv_to_rgb(const float hsv[3], uint8_t out_p[3]);
void orc_hsv_to_bgr(const float hsv[3], uint8_t out_p[3]);

/* ---- hsvfilter (video/hsv/src/hsvfilter/imp.rs:76-120,323-376), in place ---*/
int orc_hsvfilter(int fmt, int width, int height, uint8_t *data, int stride,
                  float hue_shift, float sat_mul, float sat_off,
                  float val_mul, float val_off, int threads);

/* ---- hsvdetector (video/hsv/src/hsvdetector/imp.rs:100-160,423-707) --------*/
int orc_hsvdetector(int in_fmt, int out_fmt, int width, int height,
                    const uint8_t *src, int sstride, uint8_t *dst, int dstride,
                    float hue_ref, float hue_var, float sat_ref, float sat_var,
                    float val_ref, float val_var, int threads);

/* ---- videocompare / blockhash (hashed_image.rs:24-130 -> image_hasher) ----
 * sums: hw*hh u32 block sums (integer fast path; requires W%hw==0, H%hh==0,
 * else returns -1).  fmt: ORC_FMT_RGB or ORC_FMT_RGBA. */
int orc_blockhash_sums(int fmt, int width, int height, const uint8_t *src,
                       int stride, int hw, int hh, uint32_t *sums, int threads);
/* median/bit rule (recalled from image_hasher 3.1.1; parity unpinned).
 * bits_out: hw*hh bytes of 0/1. */
void orc_blockhash_bits(const uint32_t *sums, int hw, int hh, int width,
                        int height, uint8_t *bits_out);
int orc_hamming(const uint8_t *a, const uint8_t *b, int n);

/* ---- videocompare, the other HashAlg values and the non-divisible blockhash path (vfx_oracle_hash.c; third-party
 * image_hasher 3.1.1 + image 0.25.10 restated as recalled: PARITY UNPINNED) ----------------------------------------
 * algo: 0 mean, 1 gradient, 2 vertgradient, 3 doublegradient.  fmt: ORC_FMT_RGB / ORC_FMT_RGBA. */
int orc_resize_taps(int in_len, int out_len, int out, int *left_out, float *ws, int cap);
int orc_luma_resize(int fmt, int width, int height, const uint8_t *src, int stride, int nw, int nh, uint8_t *out);
void orc_hash_resize_dims(int algo, int *nw, int *nh);
int orc_hash_bits_from_luma(int algo, const uint8_t *luma, int nw, int nh, uint8_t *bits);
int orc_blockhash_sums_f32(int fmt, int width, int height, const uint8_t *src, int stride, int hw, int hh, float *blocks);
void orc_blockhash_bits_f32(const float *blocks, int hw, int hh, int width, int height, uint8_t *bits_out);

/* ---- roundedcorners mask (video/videofx/src/border/imp.rs:57-180) ---------
 * A8 plane, `stride` bytes per row, rows [0,height); rows up to
 * round_up_2(height) are zero-filled like the reference's pre-zeroed memory.
 * Analytic restatement of the cairo drawing (parity unpinned). */
int orc_roundmask(int width, int height, int stride, unsigned radius_px,
                  uint8_t *a8);

/* ---- colordetect (video/videofx/src/colordetect/imp.rs:57-86 -> color-thief 0.2.2, color-name 1.2.0;
 * both third-party and absent from /root/reference: restated from their published algorithm, PARITY UNPINNED
 * except for tests/colordetect.rs:67 (red frame -> "red")).
 * histogram: the per-pixel pass of get_palette over the flat plane slice (stride*height bytes). */
int orc_colordetect_histogram(int fmt, int width, int height, const uint8_t *src,
                              int stride, int quality, uint32_t *hist /*32768*/);
/* modified median cut; returns the number of palette colours, writes up to cap RGB triples */
int orc_colordetect_palette(const uint32_t *hist, int max_colors, uint8_t *rgb, int cap);
const char *orc_css_similar(unsigned r, unsigned g, unsigned b);

#ifdef __cplusplus
}
#endif
#endif
