/*
 * vfx_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See vfx_oracle.h for the scope and the parity-pinning status.
 *
 * Each function cites the reference file:line (relative to /root/reference)
 * whose arithmetic it restates.  Rust semantics restated explicitly:
 *   - `x as u8/u16/usize` from f32 : saturating, NaN -> 0  (sat_u8/sat_u16/sat_idx)
 *   - inherent f32::clamp          : NaN-preserving        (clampf)
 *   - hsvutils::Clamp trait        : max-then-min, NaN -> lower (clamp_maxmin)
 *   - `%` on f32                   : fmodf (exact, sign of dividend)
 *   - f32::round                   : roundf (half away from zero)
 */
#include "vfx_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void orc_free(void *p) { free(p); }

/* ------------------------------------------------------------------------- */
/* Rust numeric helpers                                                       */
/* ------------------------------------------------------------------------- */
static inline float clampf(float x, float lo, float hi) { /* f32::clamp */
  if (x < lo) x = lo;
  if (x > hi) x = hi;
  return x; /* NaN stays NaN */
}
static inline float clamp_maxmin(float x, float lo, float hi) {
  /* hsvutils.rs:16-38  self.max(lower).min(upper); Rust max/min ignore NaN */
  return fminf(fmaxf(x, lo), hi);
}
static inline uint8_t sat_u8(float v) {
  if (!(v == v)) return 0;
  if (v <= 0.0f) return 0;
  if (v >= 255.0f) return 255;
  return (uint8_t)v; /* truncation toward zero */
}
static inline uint16_t sat_u16(float v) {
  if (!(v == v)) return 0;
  if (v <= 0.0f) return 0;
  if (v >= 65535.0f) return 65535;
  return (uint16_t)v;
}
static inline size_t sat_idx(float v, size_t max_idx) { /* (v as usize).min(max) */
  if (!(v == v)) return 0;
  if (v <= 0.0f) return 0;
  if (v >= (float)max_idx) return max_idx;
  return (size_t)v;
}

/* ------------------------------------------------------------------------- */
/* .cube parser  (video/colorlut/src/parser.rs)                               */
/* ------------------------------------------------------------------------- */
/* decode one UTF-8 scalar; returns length or 0 if invalid */
static int utf8_decode(const unsigned char *s, size_t n, uint32_t *cp) {
  if (n == 0) return 0;
  unsigned char c = s[0];
  if (c < 0x80) { *cp = c; return 1; }
  if (c >= 0xC2 && c <= 0xDF) {
    if (n < 2 || (s[1] & 0xC0) != 0x80) return 0;
    *cp = ((uint32_t)(c & 0x1F) << 6) | (s[1] & 0x3F);
    return 2;
  }
  if (c >= 0xE0 && c <= 0xEF) {
    if (n < 3 || (s[1] & 0xC0) != 0x80 || (s[2] & 0xC0) != 0x80) return 0;
    if (c == 0xE0 && s[1] < 0xA0) return 0;
    if (c == 0xED && s[1] > 0x9F) return 0; /* surrogates */
    *cp = ((uint32_t)(c & 0x0F) << 12) | ((uint32_t)(s[1] & 0x3F) << 6) | (s[2] & 0x3F);
    return 3;
  }
  if (c >= 0xF0 && c <= 0xF4) {
    if (n < 4 || (s[1] & 0xC0) != 0x80 || (s[2] & 0xC0) != 0x80 || (s[3] & 0xC0) != 0x80) return 0;
    if (c == 0xF0 && s[1] < 0x90) return 0;
    if (c == 0xF4 && s[1] > 0x8F) return 0;
    *cp = ((uint32_t)(c & 0x07) << 18) | ((uint32_t)(s[1] & 0x3F) << 12) |
          ((uint32_t)(s[2] & 0x3F) << 6) | (s[3] & 0x3F);
    return 4;
  }
  return 0;
}
static int utf8_valid(const unsigned char *s, size_t n) {
  size_t i = 0;
  while (i < n) {
    uint32_t cp;
    int l = utf8_decode(s + i, n - i, &cp);
    if (!l) return 0;
    i += (size_t)l;
  }
  return 1;
}
/* Unicode White_Space, as used by str::trim / split_whitespace */
static int is_ws(uint32_t cp) {
  return (cp >= 9 && cp <= 13) || cp == 0x20 || cp == 0x85 || cp == 0xA0 || cp == 0x1680 ||
         (cp >= 0x2000 && cp <= 0x200A) || cp == 0x2028 || cp == 0x2029 || cp == 0x202F ||
         cp == 0x205F || cp == 0x3000;
}

typedef struct { const char *p; size_t n; } tok_t;

/* split [s, s+n) (valid UTF-8) into whitespace-separated tokens; returns count (max cap) */
static size_t split_ws(const char *s, size_t n, tok_t *out, size_t cap) {
  size_t cnt = 0, i = 0;
  while (i < n) {
    uint32_t cp;
    int l = utf8_decode((const unsigned char *)s + i, n - i, &cp);
    if (is_ws(cp)) { i += (size_t)l; continue; }
    size_t start = i;
    while (i < n) {
      l = utf8_decode((const unsigned char *)s + i, n - i, &cp);
      if (is_ws(cp)) break;
      i += (size_t)l;
    }
    if (cnt < cap) { out[cnt].p = s + start; out[cnt].n = i - start; }
    cnt++;
  }
  return cnt;
}
static int tok_eq(tok_t t, const char *lit) {
  size_t l = strlen(lit);
  return t.n == l && memcmp(t.p, lit, l) == 0;
}
static int ci_eq(const char *p, size_t n, const char *lit) {
  if (n != strlen(lit)) return 0;
  for (size_t i = 0; i < n; i++) {
    char c = p[i];
    if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
    if (c != lit[i]) return 0;
  }
  return 1;
}
/* Rust `str::parse::<f32>` grammar (core::num::dec2flt): [+-] (inf|infinity|nan |
 * digits[.digits][(e|E)[+-]digits] | .digits[...]) ; correctly rounded. */
static int parse_f32_rust(tok_t t, float *out) {
  const char *p = t.p;
  size_t n = t.n, i = 0;
  if (n == 0) return 0;
  if (p[0] == '+' || p[0] == '-') i = 1;
  if (i >= n) return 0;
  if (ci_eq(p + i, n - i, "inf") || ci_eq(p + i, n - i, "infinity")) {
    *out = (p[0] == '-') ? -INFINITY : INFINITY;
    return 1;
  }
  if (ci_eq(p + i, n - i, "nan")) { *out = NAN; return 1; }
  size_t nd = 0;
  while (i < n && p[i] >= '0' && p[i] <= '9') { i++; nd++; }
  if (i < n && p[i] == '.') {
    i++;
    while (i < n && p[i] >= '0' && p[i] <= '9') { i++; nd++; }
  }
  if (nd == 0) return 0;
  if (i < n && (p[i] == 'e' || p[i] == 'E')) {
    i++;
    if (i < n && (p[i] == '+' || p[i] == '-')) i++;
    size_t ne = 0;
    while (i < n && p[i] >= '0' && p[i] <= '9') { i++; ne++; }
    if (ne == 0) return 0;
  }
  if (i != n) return 0;
  char stackbuf[128];
  char *buf = (n + 1 <= sizeof stackbuf) ? stackbuf : (char *)malloc(n + 1);
  memcpy(buf, p, n);
  buf[n] = 0;
  *out = strtof(buf, NULL); /* glibc strtof is correctly rounded */
  if (buf != stackbuf) free(buf);
  return 1;
}
/* Rust `str::parse::<usize>`: optional '+', >=1 ASCII digits, overflow is an error */
static int parse_usize_rust(tok_t t, size_t *out) {
  size_t i = 0;
  if (t.n == 0) return 0;
  if (t.p[0] == '+') i = 1;
  if (i >= t.n) return 0;
  unsigned long long v = 0;
  for (; i < t.n; i++) {
    if (t.p[i] < '0' || t.p[i] > '9') return 0;
    unsigned d = (unsigned)(t.p[i] - '0');
    if (v > (0xFFFFFFFFFFFFFFFFull - d) / 10ull) return 0;
    v = v * 10ull + d;
  }
  *out = (size_t)v;
  return 1;
}

#define PERR(code, ...)                                   \
  do {                                                    \
    if (err && errlen) snprintf(err, errlen, __VA_ARGS__); \
    free(vals);                                           \
    return (code);                                        \
  } while (0)

int orc_cube_parse(const char *text, size_t len, int *kind, int *size, float **values,
                   float scale[3], float offset[3], char *err, size_t errlen) {
  /* parser.rs:110-281 */
  float dmin[3] = {0.0f, 0.0f, 0.0f}, dmax[3] = {1.0f, 1.0f, 1.0f};
  enum { ST_HEADER, ST_1D, ST_3D } state = ST_HEADER;
  size_t lsize = 0;
  int have_data = 0;
  float *vals = NULL;
  size_t nvals = 0, cap = 0;

  if (!utf8_valid((const unsigned char *)text, len))
    PERR(-2, "IO error: stream did not contain valid UTF-8");

  size_t pos = 0, line_no = 0;
  while (pos < len) { /* str::lines(): split on \n, strip one trailing \r */
    size_t e = pos;
    while (e < len && text[e] != '\n') e++;
    size_t le = e;
    if (le > pos && text[le - 1] == '\r' && e < len) le--;
    else if (le > pos && text[le - 1] == '\r' && e == len) le--; /* "a\r" final line: lines() strips it too */
    const char *ls = text + pos;
    size_t ln = le - pos;
    pos = (e < len) ? e + 1 : e;
    line_no++;

    tok_t tk[8];
    size_t nt = split_ws(ls, ln, tk, 8);
    if (nt == 0) continue;              /* empty after trim */
    if (tk[0].p[0] == '#') continue;    /* trimmed line starts with '#' */
    /* trimmed line text for messages */
    const char *tl = tk[0].p;
    int tln = (int)((ls + ln) - tl);
    while (tln > 0) { /* trim end (ASCII approximation is enough for messages) */
      unsigned char c = (unsigned char)tl[tln - 1];
      if (c == ' ' || (c >= 9 && c <= 13)) tln--; else break;
    }

    int is_kw_title = tok_eq(tk[0], "TITLE"), is_kw_min = tok_eq(tk[0], "DOMAIN_MIN"),
        is_kw_max = tok_eq(tk[0], "DOMAIN_MAX"), is_kw_1d = tok_eq(tk[0], "LUT_1D_SIZE"),
        is_kw_3d = tok_eq(tk[0], "LUT_3D_SIZE");
    if (is_kw_title || is_kw_min || is_kw_max || is_kw_1d || is_kw_3d) {
      /* ensure_header parser.rs:284-303 */
      if (state != ST_HEADER && have_data)
        PERR(-1, "Invalid LUT: Header found after LUT data at line %zu: %.*s", line_no, tln, tl);
      if (is_kw_title) continue;
      if (is_kw_min || is_kw_max) { /* parse_vec3 :320-336 */
        float v[3];
        for (int k = 0; k < 3; k++) {
          if ((size_t)(k + 1) >= nt)
            PERR(-1, "Invalid LUT: Invalid line %zu: %.*s", line_no, tln, tl);
          if (!parse_f32_rust(tk[k + 1], &v[k]))
            PERR(-1, "Invalid LUT: Invalid float at line %zu: %.*s", line_no, tln, tl);
        }
        if (nt > 4) PERR(-1, "Invalid LUT: Invalid line %zu: %.*s", line_no, tln, tl);
        memcpy(is_kw_min ? dmin : dmax, v, sizeof v);
        continue;
      }
      /* LUT_1D_SIZE / LUT_3D_SIZE :141-174 */
      if (state != ST_HEADER)
        PERR(-1, "Invalid LUT: Invalid %s at line %zu: %.*s", is_kw_1d ? "LUT_1D_SIZE" : "LUT_3D_SIZE",
             line_no, tln, tl);
      if (nt < 2) PERR(-1, "Invalid LUT: Invalid line %zu: %.*s", line_no, tln, tl);
      size_t sz;
      if (!parse_usize_rust(tk[1], &sz))
        PERR(-1, "Invalid LUT: Invalid integer at line %zu: %.*s", line_no, tln, tl);
      if (nt > 2) PERR(-1, "Invalid LUT: Invalid line %zu: %.*s", line_no, tln, tl);
      size_t mn = 2, mx = is_kw_1d ? 65536 : 256; /* parser.rs:12-16 */
      if (sz < mn || sz > mx)
        PERR(-1, "Invalid LUT: Invalid LUT size %zu at line %zu, expected %zu..=%zu", sz, line_no, mn, mx);
      state = is_kw_1d ? ST_1D : ST_3D;
      lsize = sz;
      have_data = 0;
      continue;
    }
    /* data line :176-201 */
    if (state == ST_HEADER)
      PERR(-1, "Invalid LUT: LUT data found before LUT size at line %zu: %.*s", line_no, tln, tl);
    have_data = 1;
    float v[3];
    for (int k = 0; k < 3; k++) {
      if ((size_t)k >= nt) PERR(-1, "Invalid LUT: Invalid line %zu: %.*s", line_no, tln, tl);
      if (!parse_f32_rust(tk[k], &v[k]))
        PERR(-1, "Invalid LUT: Invalid float at line %zu: %.*s", line_no, tln, tl);
    }
    if (nt > 3) PERR(-1, "Invalid LUT: Invalid line %zu: %.*s", line_no, tln, tl);
    if (nvals == cap) {
      cap = cap ? cap * 2 : 4096;
      vals = (float *)realloc(vals, cap * 3 * sizeof(float));
    }
    memcpy(vals + nvals * 3, v, sizeof v);
    nvals++;
  }

  if (dmin[0] >= dmax[0] || dmin[1] >= dmax[1] || dmin[2] >= dmax[2]) /* :205-212 */
    PERR(-1, "Invalid LUT: Invalid domain min [%g, %g, %g], max [%g, %g, %g]", dmin[0], dmin[1],
         dmin[2], dmax[0], dmax[1], dmax[2]);
  if (state == ST_HEADER) PERR(-1, "Invalid LUT: Missing LUT size");
  if (state == ST_1D) {
    if (nvals != lsize)
      PERR(-1, "Invalid LUT: Invalid 1D LUT value count, expected %zu, got %zu", lsize, nvals);
  } else {
    size_t expected = lsize * lsize * lsize;
    if (nvals != expected)
      PERR(-1, "Invalid LUT: Invalid 3D LUT value count, expected %zu, got %zu", expected, nvals);
  }
  for (int c = 0; c < 3; c++) { /* :264-274 */
    scale[c] = 1.0f / (dmax[c] - dmin[c]);
    offset[c] = -dmin[c] * scale[c];
  }
  *kind = (state == ST_1D) ? 1 : 3;
  *size = (int)lsize;
  if (!vals) vals = (float *)malloc(4);
  *values = vals;
  return 0;
}

int orc_cube_parse_file(const char *path, int *kind, int *size, float **values, float scale[3],
                        float offset[3], char *err, size_t errlen) {
  /* parser.rs:105-108 fs::read_to_string */
  FILE *f = fopen(path, "rb");
  if (!f) {
    if (err && errlen) snprintf(err, errlen, "IO error: cannot open %s", path);
    return -2;
  }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  char *buf = (char *)malloc((size_t)n + 1);
  size_t rd = fread(buf, 1, (size_t)n, f);
  fclose(f);
  int rc = orc_cube_parse(buf, rd, kind, size, values, scale, offset, err, errlen);
  free(buf);
  return rc;
}

/* ------------------------------------------------------------------------- */
/* colorlut pixel math (video/colorlut/src/colorlut/imp.rs)                   */
/* ------------------------------------------------------------------------- */
typedef struct {
  int kind, size;
  const float *v; /* size x 3 (1D) or size^3 x 3 (3D) */
  float scale[3], offset[3];
} lut_t;

static inline float norm_comp(const lut_t *l, int c, float value, float denom) {
  /* imp.rs:471-479 */
  float v = value / denom;
  float n = v * l->scale[c];
  n = n + l->offset[c];
  return clampf(n, 0.0f, 1.0f);
}
static inline float sample_1d(const lut_t *l, int c, float x) { /* imp.rs:482-490 */
  size_t max_idx = (size_t)l->size - 1;
  size_t x0 = sat_idx(floorf(x), max_idx);
  size_t x1 = x0 + 1 < max_idx ? x0 + 1 : max_idx;
  float t = x - (float)x0;
  float a = l->v[x0 * 3 + (size_t)c], b = l->v[x1 * 3 + (size_t)c];
  float d = b - a;
  float p = d * t;
  return a + p;
}
static inline float lerp1(float a, float b, float t) { /* imp.rs:528-535, per lane */
  float d = b - a;
  float p = d * t;
  return a + p;
}
static inline void sample_3d(const lut_t *l, float x, float y, float z, float out[3]) {
  /* imp.rs:493-526; lane 3 (alpha = 1.0) is never consumed (imp.rs:444-448) */
  size_t n = (size_t)l->size, max_idx = n - 1;
  size_t x0 = sat_idx(floorf(x), max_idx), y0 = sat_idx(floorf(y), max_idx),
         z0 = sat_idx(floorf(z), max_idx);
  size_t x1 = x0 + 1 < max_idx ? x0 + 1 : max_idx;
  size_t y1 = y0 + 1 < max_idx ? y0 + 1 : max_idx;
  size_t z1 = z0 + 1 < max_idx ? z0 + 1 : max_idx;
  float tx = x - (float)x0, ty = y - (float)y0, tz = z - (float)z0;
#define AT(X, Y, Z) (l->v + 3 * ((X) + (Y)*n + (Z)*n * n)) /* parser.rs:43-53 */
  const float *c000 = AT(x0, y0, z0), *c100 = AT(x1, y0, z0), *c010 = AT(x0, y1, z0),
              *c110 = AT(x1, y1, z0), *c001 = AT(x0, y0, z1), *c101 = AT(x1, y0, z1),
              *c011 = AT(x0, y1, z1), *c111 = AT(x1, y1, z1);
#undef AT
  for (int k = 0; k < 3; k++) {
    float c00 = lerp1(c000[k], c100[k], tx);
    float c10 = lerp1(c010[k], c110[k], tx);
    float c01 = lerp1(c001[k], c101[k], tx);
    float c11 = lerp1(c011[k], c111[k], tx);
    float c0 = lerp1(c00, c10, ty);
    float c1 = lerp1(c01, c11, ty);
    out[k] = lerp1(c0, c1, tz);
  }
}
static inline uint8_t float_to_u8(float v) { /* imp.rs:537-539 */
  return sat_u8(roundf(clampf(v, 0.0f, 1.0f) * 255.0f));
}
static inline uint16_t float_to_u16(float v) { /* imp.rs:541-543 */
  return sat_u16(roundf(clampf(v, 0.0f, 1.0f) * 65535.0f));
}
static inline uint16_t bswap16(uint16_t v) { return (uint16_t)((v >> 8) | (v << 8)); }

int orc_colorlut_apply(int kind, int size, const float *values, const float scale[3],
                       const float offset[3], int fmt, int width, int height,
                       const uint8_t *src, int sstride, uint8_t *dst, int dstride, int threads) {
  if ((kind != 1 && kind != 3) || size < 2 || width < 0 || height < 0) return -1;
  if (fmt != ORC_FMT_RGBA && fmt != ORC_FMT_RGBA64_LE && fmt != ORC_FMT_RGBA64_BE) return -1;
  lut_t l;
  l.kind = kind; l.size = size; l.v = values;
  memcpy(l.scale, scale, sizeof l.scale);
  memcpy(l.offset, offset, sizeof l.offset);
  const float sm1 = (float)size - 1.0f; /* `size as f32 - 1.0` */
  if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int row = 0; row < height; row++) {
    if (fmt == ORC_FMT_RGBA) { /* imp.rs:237-294 */
      const uint8_t *s = src + (size_t)row * (size_t)sstride;
      uint8_t *d = dst + (size_t)row * (size_t)dstride;
      for (int px = 0; px < width; px++, s += 4, d += 4) {
        if (kind == 1) {
          for (int c = 0; c < 3; c++) { /* apply_1d imp.rs:399-413 */
            float x = norm_comp(&l, c, (float)s[c], 255.0f) * sm1;
            d[c] = float_to_u8(sample_1d(&l, c, x));
          }
        } else { /* apply_3d imp.rs:431-449 */
          float x = norm_comp(&l, 0, (float)s[0], 255.0f) * sm1;
          float y = norm_comp(&l, 1, (float)s[1], 255.0f) * sm1;
          float z = norm_comp(&l, 2, (float)s[2], 255.0f) * sm1;
          float o[3];
          sample_3d(&l, x, y, z, o);
          d[0] = float_to_u8(o[0]); d[1] = float_to_u8(o[1]); d[2] = float_to_u8(o[2]);
        }
        d[3] = s[3];
      }
    } else { /* imp.rs:307-397; rows addressed in u16 units = stride/2 */
      const int le = (fmt == ORC_FMT_RGBA64_LE);
      const uint16_t *s = (const uint16_t *)(src + (size_t)row * ((size_t)(sstride / 2) * 2));
      uint16_t *d = (uint16_t *)(dst + (size_t)row * ((size_t)(dstride / 2) * 2));
      for (int px = 0; px < width; px++, s += 4, d += 4) {
        uint16_t in[3], out[3];
        for (int c = 0; c < 3; c++) in[c] = le ? s[c] : bswap16(s[c]); /* host is LE */
        if (kind == 1) {
          for (int c = 0; c < 3; c++) { /* apply_1d_u16 imp.rs:415-429 */
            float x = norm_comp(&l, c, (float)in[c], 65535.0f) * sm1;
            out[c] = float_to_u16(sample_1d(&l, c, x));
          }
        } else { /* apply_3d_u16 imp.rs:451-469 */
          float x = norm_comp(&l, 0, (float)in[0], 65535.0f) * sm1;
          float y = norm_comp(&l, 1, (float)in[1], 65535.0f) * sm1;
          float z = norm_comp(&l, 2, (float)in[2], 65535.0f) * sm1;
          float o[3];
          sample_3d(&l, x, y, z, o);
          out[0] = float_to_u16(o[0]); out[1] = float_to_u16(o[1]); out[2] = float_to_u16(o[2]);
        }
        for (int c = 0; c < 3; c++) d[c] = le ? out[c] : bswap16(out[c]);
        d[3] = s[3]; /* alpha copied raw, no byte swap (imp.rs:345,394) */
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* hsvutils (video/hsv/src/hsvutils.rs)                                       */
/* ------------------------------------------------------------------------- */
#define HSV_EPSILON 0.00001f /* hsvutils.rs:40 */

static inline void hsv_from(uint8_t rb, uint8_t gb, uint8_t bb, float hsv[3]) {
  /* hsvutils.rs:44-84 (from_rgb) / :88-128 (from_bgr): identical once r,g,b are named */
  float r = (float)rb / 255.0f, g = (float)gb / 255.0f, b = (float)bb / 255.0f;
  uint8_t mx = rb > gb ? rb : gb; if (bb > mx) mx = bb;
  uint8_t mn = rb < gb ? rb : gb; if (bb < mn) mn = bb;
  float value = (float)mx / 255.0f;
  float chroma = value - ((float)mn / 255.0f);
  float hue;
  if (chroma == 0.0f) hue = 0.0f;
  else if (fabsf(value - r) < HSV_EPSILON) hue = 60.0f * ((g - b) / chroma);
  else if (fabsf(value - g) < HSV_EPSILON) hue = 60.0f * (2.0f + ((b - r) / chroma));
  else if (fabsf(value - b) < HSV_EPSILON) hue = 60.0f * (4.0f + ((r - g) / chroma));
  else hue = 0.0f;
  if (hue < 0.0f) hue += 360.0f;
  float saturation = (value == 0.0f) ? 0.0f : chroma / value;
  hsv[0] = fmodf(hue, 360.0f);
  hsv[1] = clampf(saturation, 0.0f, 1.0f);
  hsv[2] = clampf(value, 0.0f, 1.0f);
}
static inline void hsv_to(const float in_p[3], float rgbp[3]) {
  /* hsvutils.rs:132-163: returns (r',g',b') + m scaled, before the u8 cast */
  float c = in_p[2] * in_p[1];
  float hue_prime = in_p[0] / 60.0f;
  float x = c * (1.0f - fabsf(fmodf(hue_prime, 2.0f) - 1.0f));
  float p0, p1, p2;
  if (hue_prime < 0.0f) { p0 = 0; p1 = 0; p2 = 0; }
  else if (hue_prime <= 1.0f) { p0 = c; p1 = x; p2 = 0; }
  else if (hue_prime <= 2.0f) { p0 = x; p1 = c; p2 = 0; }
  else if (hue_prime <= 3.0f) { p0 = 0; p1 = c; p2 = x; }
  else if (hue_prime <= 4.0f) { p0 = 0; p1 = x; p2 = c; }
  else if (hue_prime <= 5.0f) { p0 = x; p1 = 0; p2 = c; }
  else if (hue_prime <= 6.0f) { p0 = c; p1 = 0; p2 = x; }
  else { p0 = 0; p1 = 0; p2 = 0; } /* also the NaN case: every comparison false */
  float m = in_p[2] - c;
  rgbp[0] = clampf((p0 + m) * 255.0f, 0.0f, 255.0f);
  rgbp[1] = clampf((p1 + m) * 255.0f, 0.0f, 255.0f);
  rgbp[2] = clampf((p2 + m) * 255.0f, 0.0f, 255.0f);
}
void orc_hsv_from_rgb(const uint8_t p[3], float hsv[3]) { hsv_from(p[0], p[1], p[2], hsv); }
void orc_hsv_from_bgr(const uint8_t p[3], float hsv[3]) { hsv_from(p[2], p[1], p[0], hsv); }
void orc_hsv_to_rgb(const float hsv[3], uint8_t o[3]) {
  float f[3]; hsv_to(hsv, f);
  o[0] = sat_u8(f[0]); o[1] = sat_u8(f[1]); o[2] = sat_u8(f[2]); /* `as u8` truncates */
}
void orc_hsv_to_bgr(const float hsv[3], uint8_t o[3]) {
  float f[3]; hsv_to(hsv, f);
  o[0] = sat_u8(f[2]); o[1] = sat_u8(f[1]); o[2] = sat_u8(f[0]);
}

/* format -> (bytes per pixel, colour offset, is_bgr) ; -1 if not an 8-bit packed RGB format */
static int fmt_info(int fmt, int *bpp, int *coff, int *bgr) {
  switch (fmt) {
    case ORC_FMT_RGBX: case ORC_FMT_RGBA: *bpp = 4; *coff = 0; *bgr = 0; return 0;
    case ORC_FMT_XRGB: case ORC_FMT_ARGB: *bpp = 4; *coff = 1; *bgr = 0; return 0;
    case ORC_FMT_BGRX: case ORC_FMT_BGRA: *bpp = 4; *coff = 0; *bgr = 1; return 0;
    case ORC_FMT_XBGR: case ORC_FMT_ABGR: *bpp = 4; *coff = 1; *bgr = 1; return 0;
    case ORC_FMT_RGB: *bpp = 3; *coff = 0; *bgr = 0; return 0;
    case ORC_FMT_BGR: *bpp = 3; *coff = 0; *bgr = 1; return 0;
    default: return -1;
  }
}

int orc_hsvfilter(int fmt, int width, int height, uint8_t *data, int stride, float hue_shift,
                  float sat_mul, float sat_off, float val_mul, float val_off, int threads) {
  /* hsvfilter/imp.rs:76-120 with the closures of :323-376 */
  int bpp, coff, bgr;
  if (fmt_info(fmt, &bpp, &coff, &bgr)) return -1;
  if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int row = 0; row < height; row++) {
    uint8_t *p = data + (size_t)row * (size_t)stride + coff;
    for (int px = 0; px < width; px++, p += bpp) {
      float hsv[3];
      if (bgr) hsv_from(p[2], p[1], p[0], hsv); else hsv_from(p[0], p[1], p[2], hsv);
      float h = hsv[0] + hue_shift;
      h = fmodf(h, 360.0f);
      if (h < 0.0f) h += 360.0f;
      hsv[0] = h;
      float s = sat_mul * hsv[1]; s = s + sat_off;
      hsv[1] = clamp_maxmin(s, 0.0f, 1.0f);
      float v = val_mul * hsv[2]; v = v + val_off;
      hsv[2] = clamp_maxmin(v, 0.0f, 1.0f);
      float f[3];
      hsv_to(hsv, f);
      if (bgr) { p[0] = sat_u8(f[2]); p[1] = sat_u8(f[1]); p[2] = sat_u8(f[0]); }
      else { p[0] = sat_u8(f[0]); p[1] = sat_u8(f[1]); p[2] = sat_u8(f[2]); }
    }
  }
  return 0;
}

int orc_hsvdetector(int in_fmt, int out_fmt, int width, int height, const uint8_t *src,
                    int sstride, uint8_t *dst, int dstride, float hue_ref, float hue_var,
                    float sat_ref, float sat_var, float val_ref, float val_var, int threads) {
  /* hsvdetector/imp.rs:100-160 with the 16 closure pairs of :423-707 */
  int ibpp, icoff, ibgr, obpp, ocoff, obgr;
  if (fmt_info(in_fmt, &ibpp, &icoff, &ibgr)) return -1;
  if (in_fmt == ORC_FMT_RGBA || in_fmt == ORC_FMT_ARGB || in_fmt == ORC_FMT_BGRA ||
      in_fmt == ORC_FMT_ABGR) return -1; /* sink caps :78-87 */
  if (out_fmt != ORC_FMT_RGBA && out_fmt != ORC_FMT_ARGB && out_fmt != ORC_FMT_BGRA &&
      out_fmt != ORC_FMT_ABGR) return -1; /* src caps :89-96 */
  fmt_info(out_fmt, &obpp, &ocoff, &obgr);
  const int aoff = ocoff ? 0 : 3;
  if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int row = 0; row < height; row++) {
    const uint8_t *ip = src + (size_t)row * (size_t)sstride + icoff;
    uint8_t *op = dst + (size_t)row * (size_t)dstride;
    for (int px = 0; px < width; px++, ip += ibpp, op += 4) {
      float hsv[3];
      if (ibgr) hsv_from(ip[2], ip[1], ip[0], hsv); else hsv_from(ip[0], ip[1], ip[2], hsv);
      float ref_hue_offset = 180.0f - hue_ref;
      float sh = hsv[0] + ref_hue_offset;
      if (sh < 0.0f) sh += 360.0f;
      sh = fmodf(sh, 360.0f);
      int hit = fabsf(sh - 180.0f) <= hue_var && fabsf(hsv[1] - sat_ref) <= sat_var &&
                fabsf(hsv[2] - val_ref) <= val_var;
      if (ibgr == obgr) { op[ocoff] = ip[0]; op[ocoff + 1] = ip[1]; op[ocoff + 2] = ip[2]; }
      else { op[ocoff] = ip[2]; op[ocoff + 1] = ip[1]; op[ocoff + 2] = ip[0]; }
      op[aoff] = hit ? 255 : 0;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* videocompare: blockhash (videocompare/hashed_image.rs -> image_hasher)     */
/* ------------------------------------------------------------------------- */
int orc_blockhash_sums(int fmt, int width, int height, const uint8_t *src, int stride, int hw,
                       int hh, uint32_t *sums, int threads) {
  /* hashed_image.rs:110-130 de-strides the frame; image_hasher blockhash fast
   * path (W%hw==0 && H%hh==0): sum[by*hw+bx] += (A==0 ? 765 : R+G+B). */
  (void)threads;
  if ((fmt != ORC_FMT_RGB && fmt != ORC_FMT_RGBA) || hw <= 0 || hh <= 0) return -1;
  if (width <= 0 || height <= 0 || width % hw || height % hh) return -1;
  const int bpp = (fmt == ORC_FMT_RGB) ? 3 : 4;
  const int bw = width / hw, bh = height / hh;
  memset(sums, 0, sizeof(uint32_t) * (size_t)hw * (size_t)hh);
  for (int y = 0; y < height; y++) {
    const uint8_t *p = src + (size_t)y * (size_t)stride;
    uint32_t *rowsums = sums + (size_t)(y / bh) * (size_t)hw;
    for (int x = 0; x < width; x++, p += bpp) {
      uint32_t s = (uint32_t)p[0] + p[1] + p[2];
      if (bpp == 4 && p[3] == 0) s = 765;
      rowsums[x / bw] += s;
    }
  }
  return 0;
}

static int cmp_u32(const void *a, const void *b) {
  uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
  return (x > y) - (x < y);
}
void orc_blockhash_bits(const uint32_t *sums, int hw, int hh, int width, int height,
                        uint8_t *bits_out) {
  /* recalled from image_hasher 3.1.1 alg/blockhash.rs gen_hash! (parity unpinned):
   * groups of `hash width * 4` blocks (= 4 rows of the hash grid: two groups for the 8x8 hash); per group
   * m = sorted[len/2]; bit = block > m || (block == m && m > cmp_factor),
   * cmp_factor = 255 * 3 * bw * bh / 2 (integer). */
  const int n = hw * hh, group = hw * 4;
  const uint32_t half = (uint32_t)(((uint64_t)765 * (uint64_t)(width / hw) * (uint64_t)(height / hh)) / 2);
  uint32_t *scratch = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(group > 0 ? group : 1));
  for (int g0 = 0; group > 0 && g0 < n; g0 += group) {
    const int len = (n - g0 < group) ? n - g0 : group;
    memcpy(scratch, sums + g0, sizeof(uint32_t) * (size_t)len);
    qsort(scratch, (size_t)len, sizeof(uint32_t), cmp_u32);
    uint32_t m = scratch[len / 2];
    for (int i = 0; i < len; i++) {
      uint32_t v = sums[g0 + i];
      bits_out[g0 + i] = (uint8_t)(v > m || (v == m && m > half));
    }
  }
  free(scratch);
}
int orc_hamming(const uint8_t *a, const uint8_t *b, int n) {
  int d = 0;
  for (int i = 0; i < n; i++) d += (a[i] != b[i]);
  return d;
}

/* ------------------------------------------------------------------------- */
/* roundedcorners alpha mask (video/videofx/src/border/imp.rs:57-180)         */
/* ------------------------------------------------------------------------- */
/* pixman MUL_UN8 */
static inline unsigned mul_un8(unsigned a, unsigned b) {
  unsigned t = a * b + 0x80u;
  return ((t >> 8) + t) >> 8;
}
int orc_roundmask(int width, int height, int stride, unsigned radius_px, uint8_t *a8) {
  /* border/imp.rs:108-180.  radius 0 -> fill 0xFF (:123-128).  Otherwise the
   * memory is zeroed (:130) and cairo fills the rounded rectangle and strokes
   * its outline with width 1 (:57-106).  Restated analytically: only the four
   * r x r corner boxes are not fully covered; inside them
   *   fill   coverage = area(pixel ∩ disc(centre, r)),
   *   stroke coverage = area(pixel ∩ annulus r-0.5 <= rho <= r+0.5),
   * each estimated on a 16x16 sample grid in exact integer arithmetic, rounded
   * to 8 bits, and composited stroke OVER fill with pixman's MUL_UN8.
   * r is clamped to min(w,h)/2 (documented deviation: the reference's path
   * self-intersects beyond that). */
  if (width <= 0 || height <= 0 || stride < width) return -1;
  const size_t total = (size_t)stride * (size_t)((height + 1) & ~1);
  if (radius_px == 0) { memset(a8, 0xFF, total); return 0; }
  memset(a8, 0, total);
  long r = (long)radius_px;
  long lim = (width < height ? width : height) / 2;
  if (r > lim) r = lim;
  for (int y = 0; y < height; y++) memset(a8 + (size_t)y * (size_t)stride, 0xFF, (size_t)width);
  if (r == 0) return 0;
  const long S = 16;                       /* samples per axis */
  const long R2f = (2 * S * r) * (2 * S * r);                 /* (r * 2S)^2 */
  const long Ro = 2 * S * r + S, Ri = 2 * S * r - S;          /* (r +- 0.5) * 2S */
  const long R2o = Ro * Ro, R2i = Ri * Ri;
  for (long j = 0; j < r; j++) {     /* j, i = pixel index measured from the arc centre */
    for (long i = 0; i < r; i++) {
      int nf = 0, ns = 0;
      for (long b = 0; b < S; b++) {
        long dy = 2 * S * j + 2 * b + 1; /* (j + (2b+1)/(2S)) * 2S */
        for (long a = 0; a < S; a++) {
          long dx = 2 * S * i + 2 * a + 1;
          long d2 = dx * dx + dy * dy;
          nf += (d2 <= R2f);
          ns += (d2 <= R2o && d2 >= R2i);
        }
      }
      unsigned af = (unsigned)((nf * 255 + 128) / 256);
      unsigned as = (unsigned)((ns * 255 + 128) / 256);
      unsigned v = as + mul_un8(af, 255u - as);
      if (v > 255u) v = 255u;
      long xl = r - 1 - i, xr = width - r + i, yt = r - 1 - j, yb = height - r + j;
      a8[(size_t)yt * (size_t)stride + (size_t)xl] = (uint8_t)v;
      a8[(size_t)yt * (size_t)stride + (size_t)xr] = (uint8_t)v;
      a8[(size_t)yb * (size_t)stride + (size_t)xl] = (uint8_t)v;
      a8[(size_t)yb * (size_t)stride + (size_t)xr] = (uint8_t)v;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* colordetect (video/videofx/src/colordetect/imp.rs:57-86)                   */
/*   get_palette() of color-thief 0.2.2 and Color::similar() of color-name    */
/*   1.2.0 are third-party crates not present under /root/reference: this is  */
/*   a restatement of their published algorithm (MMCQ), PARITY UNPINNED.      */
/* ------------------------------------------------------------------------- */
int orc_colordetect_histogram(int fmt, int width, int height, const uint8_t *src,
                              int stride, int quality, uint32_t *hist) {
  int ro, go, bo, ao, bpp;
  switch (fmt) {               /* color_parts(): byte positions per ColorFormat */
    case ORC_FMT_RGB:  ro = 0; go = 1; bo = 2; ao = -1; bpp = 3; break;
    case ORC_FMT_RGBA: ro = 0; go = 1; bo = 2; ao = 3;  bpp = 4; break;
    case ORC_FMT_ARGB: ro = 1; go = 2; bo = 3; ao = 0;  bpp = 4; break;
    case ORC_FMT_BGR:  ro = 2; go = 1; bo = 0; ao = -1; bpp = 3; break;
    case ORC_FMT_BGRA: ro = 2; go = 1; bo = 0; ao = 3;  bpp = 4; break;
    default: return -1;
  }
  if (quality < 1 || quality > 10) return -1;
  memset(hist, 0, sizeof(uint32_t) * 32768);
  if (width <= 0 || height <= 0) return 0;
  /* frame.plane_data(0): one flat slice of stride*height bytes; pixel i sits at
   * byte i*bpp regardless of the stride (imp.rs:68-69) */
  const size_t len = (size_t)stride * (size_t)height;
  const size_t pixel_count = len / (size_t)bpp;
  for (size_t i = 0; i < pixel_count; i += (size_t)quality) {
    const uint8_t *p = src + i * (size_t)bpp;
    unsigned r = p[ro], g = p[go], b = p[bo], a = ao < 0 ? 255u : p[ao];
    if (a >= 125 && !(r > 250 && g > 250 && b > 250))     /* mostly opaque and not white */
      hist[((r >> 3) << 10) + ((g >> 3) << 5) + (b >> 3)] += 1;
  }
  return 0;
}

typedef struct { int lo[3], hi[3]; int avg[3]; long long count; long long volume; int seq; } orc_vbox;

static long long orc_h(const uint32_t *hist, int r, int g, int b) { return hist[(r << 10) + (g << 5) + b]; }

static void orc_vbox_recalc(const uint32_t *hist, orc_vbox *v) {
  long long n = 0, s[3] = {0, 0, 0};
  for (int r = v->lo[0]; r <= v->hi[0]; r++)
    for (int g = v->lo[1]; g <= v->hi[1]; g++)
      for (int b = v->lo[2]; b <= v->hi[2]; b++) {
        double hv = (double)orc_h(hist, r, g, b);
        n += (long long)hv;
        s[0] += (long long)(hv * (r + 0.5) * 8.0);   /* truncated per bin, like `as i32` */
        s[1] += (long long)(hv * (g + 0.5) * 8.0);
        s[2] += (long long)(hv * (b + 0.5) * 8.0);
      }
  for (int c = 0; c < 3; c++) {
    if (n > 0) v->avg[c] = (int)((s[c] / n) & 255);
    else { int m = 8 * (v->lo[c] + v->hi[c] + 1) / 2; v->avg[c] = m > 255 ? 255 : m; }
  }
  v->count = n;
  v->volume = (long long)(v->hi[0] - v->lo[0] + 1) * (v->hi[1] - v->lo[1] + 1) * (v->hi[2] - v->lo[2] + 1);
}

/* population of the slab `i` along `axis` inside the box */
static long long orc_slab(const uint32_t *hist, const orc_vbox *v, int axis, int i) {
  int lo[3], hi[3];
  for (int c = 0; c < 3; c++) { lo[c] = v->lo[c]; hi[c] = v->hi[c]; }
  lo[axis] = hi[axis] = i;
  long long s = 0;
  for (int r = lo[0]; r <= hi[0]; r++) for (int g = lo[1]; g <= hi[1]; g++) for (int b = lo[2]; b <= hi[2]; b++) s += orc_h(hist, r, g, b);
  return s;
}

static int orc_median_cut(const uint32_t *hist, const orc_vbox *v, orc_vbox *a, orc_vbox *b) {
  if (v->count <= 1) return 0;
  int w[3], axis;
  for (int c = 0; c < 3; c++) w[c] = v->hi[c] - v->lo[c];
  int m = w[0] > w[1] ? w[0] : w[1]; if (w[2] > m) m = w[2];
  axis = (m == w[0]) ? 0 : (m == w[1] ? 1 : 2);
  long long part[32], look[32], total = 0;
  for (int i = 0; i < 32; i++) part[i] = look[i] = -1;
  for (int i = v->lo[axis]; i <= v->hi[axis]; i++) { total += orc_slab(hist, v, axis, i); part[i] = total; }
  for (int i = 0; i < 32; i++) if (part[i] != -1) look[i] = total - part[i];
  const int vmin = v->lo[axis], vmax = v->hi[axis];
  for (int i = vmin; i <= vmax; i++) {
    if (part[i] <= total / 2) continue;
    int left = i - vmin, right = vmax - i, d2;
    if (left <= right) { d2 = i + right / 2; if (d2 > vmax - 1) d2 = vmax - 1; }
    else { d2 = (int)((double)(i - 1) - (double)left / 2.0); if (d2 < vmin) d2 = vmin; }
    while (d2 < 0 || part[d2] <= 0) d2++;
    long long c2 = look[d2];
    while (c2 == 0 && d2 > 0 && part[d2 - 1] > 0) { d2--; c2 = look[d2]; }
    *a = *v; *b = *v;
    a->hi[axis] = d2; b->lo[axis] = d2 + 1;
    orc_vbox_recalc(hist, a); orc_vbox_recalc(hist, b);
    return 1;
  }
  return 0;
}

static int orc_cmp_mode;  /* 0: by count, 1: by count*volume (count ties: by volume); seq keeps the sort stable */
static int orc_vbox_cmp(const void *pa, const void *pb) {
  const orc_vbox *x = (const orc_vbox *)pa, *y = (const orc_vbox *)pb;
  long long kx, ky;
  if (orc_cmp_mode == 0) { kx = x->count; ky = y->count; }
  else if (x->count == y->count) { kx = x->volume; ky = y->volume; }
  else { kx = x->count * x->volume; ky = y->count * y->volume; }
  if (kx != ky) return kx < ky ? -1 : 1;
  return x->seq - y->seq;
}
static void orc_sort(orc_vbox *q, int n, int mode) {
  for (int i = 0; i < n; i++) q[i].seq = i;
  orc_cmp_mode = mode;
  qsort(q, (size_t)n, sizeof(orc_vbox), orc_vbox_cmp);
}
static void orc_iterate(const uint32_t *hist, orc_vbox *q, int *n, int mode, int target) {
  int color = 1;
  for (int it = 0; it < 1000; it++) {
    orc_vbox last = q[*n - 1];
    if (last.count == 0) { orc_sort(q, *n, mode); continue; }
    orc_vbox a, b;
    if (orc_median_cut(hist, &last, &a, &b)) { q[*n - 1] = a; q[*n] = b; (*n)++; color++; }
    orc_sort(q, *n, mode);
    if (color >= target) return;
  }
}

int orc_colordetect_palette(const uint32_t *hist, int max_colors, uint8_t *rgb, int cap) {
  if (max_colors < 2 || max_colors > 255) return -1;
  orc_vbox *q = (orc_vbox *)calloc(2100, sizeof(orc_vbox));
  int n = 1;
  for (int c = 0; c < 3; c++) { q[0].lo[c] = 255; q[0].hi[c] = 0; }
  for (int i = 0; i < 32768; i++)
    if (hist[i]) {
      int v[3] = {i >> 10, (i >> 5) & 31, i & 31};
      for (int c = 0; c < 3; c++) { if (v[c] < q[0].lo[c]) q[0].lo[c] = v[c]; if (v[c] > q[0].hi[c]) q[0].hi[c] = v[c]; }
    }
  orc_vbox_recalc(hist, &q[0]);
  int target = (int)ceil(0.75 * (double)max_colors);
  orc_iterate(hist, q, &n, 0, target);
  orc_sort(q, n, 1);
  orc_iterate(hist, q, &n, 1, max_colors - n);
  for (int i = 0; i < n && i < cap; i++) {         /* reversed: most significant first */
    const orc_vbox *v = &q[n - 1 - i];
    rgb[3 * i] = (uint8_t)v->avg[0]; rgb[3 * i + 1] = (uint8_t)v->avg[1]; rgb[3 * i + 2] = (uint8_t)v->avg[2];
  }
  free(q);
  return n;
}

static const struct { const char *n; unsigned char r, g, b; } orc_css[] = {
  {"aliceblue",240,248,255},{"antiquewhite",250,235,215},{"aqua",0,255,255},{"aquamarine",127,255,212},{"azure",240,255,255},
  {"beige",245,245,220},{"bisque",255,228,196},{"black",0,0,0},{"blanchedalmond",255,235,205},{"blue",0,0,255},
  {"blueviolet",138,43,226},{"brown",165,42,42},{"burlywood",222,184,135},{"cadetblue",95,158,160},{"chartreuse",127,255,0},
  {"chocolate",210,105,30},{"coral",255,127,80},{"cornflowerblue",100,149,237},{"cornsilk",255,248,220},{"crimson",220,20,60},
  {"cyan",0,255,255},{"darkblue",0,0,139},{"darkcyan",0,139,139},{"darkgoldenrod",184,134,11},{"darkgray",169,169,169},
  {"darkgreen",0,100,0},{"darkgrey",169,169,169},{"darkkhaki",189,183,107},{"darkmagenta",139,0,139},{"darkolivegreen",85,107,47},
  {"darkorange",255,140,0},{"darkorchid",153,50,204},{"darkred",139,0,0},{"darksalmon",233,150,122},{"darkseagreen",143,188,143},
  {"darkslateblue",72,61,139},{"darkslategray",47,79,79},{"darkslategrey",47,79,79},{"darkturquoise",0,206,209},{"darkviolet",148,0,211},
  {"deeppink",255,20,147},{"deepskyblue",0,191,255},{"dimgray",105,105,105},{"dimgrey",105,105,105},{"dodgerblue",30,144,255},
  {"firebrick",178,34,34},{"floralwhite",255,250,240},{"forestgreen",34,139,34},{"fuchsia",255,0,255},{"gainsboro",220,220,220},
  {"ghostwhite",248,248,255},{"gold",255,215,0},{"goldenrod",218,165,32},{"gray",128,128,128},{"green",0,128,0},
  {"greenyellow",173,255,47},{"grey",128,128,128},{"honeydew",240,255,240},{"hotpink",255,105,180},{"indianred",205,92,92},
  {"indigo",75,0,130},{"ivory",255,255,240},{"khaki",240,230,140},{"lavender",230,230,250},{"lavenderblush",255,240,245},
  {"lawngreen",124,252,0},{"lemonchiffon",255,250,205},{"lightblue",173,216,230},{"lightcoral",240,128,128},{"lightcyan",224,255,255},
  {"lightgoldenrodyellow",250,250,210},{"lightgray",211,211,211},{"lightgreen",144,238,144},{"lightgrey",211,211,211},{"lightpink",255,182,193},
  {"lightsalmon",255,160,122},{"lightseagreen",32,178,170},{"lightskyblue",135,206,250},{"lightslategray",119,136,153},{"lightslategrey",119,136,153},
  {"lightsteelblue",176,196,222},{"lightyellow",255,255,224},{"lime",0,255,0},{"limegreen",50,205,50},{"linen",250,240,230},
  {"magenta",255,0,255},{"maroon",128,0,0},{"mediumaquamarine",102,205,170},{"mediumblue",0,0,205},{"mediumorchid",186,85,211},
  {"mediumpurple",147,112,219},{"mediumseagreen",60,179,113},{"mediumslateblue",123,104,238},{"mediumspringgreen",0,250,154},{"mediumturquoise",72,209,204},
  {"mediumvioletred",199,21,133},{"midnightblue",25,25,112},{"mintcream",245,255,250},{"mistyrose",255,228,225},{"moccasin",255,228,181},
  {"navajowhite",255,222,173},{"navy",0,0,128},{"oldlace",253,245,230},{"olive",128,128,0},{"olivedrab",107,142,35},
  {"orange",255,165,0},{"orangered",255,69,0},{"orchid",218,112,214},{"palegoldenrod",238,232,170},{"palegreen",152,251,152},
  {"paleturquoise",175,238,238},{"palevioletred",219,112,147},{"papayawhip",255,239,213},{"peachpuff",255,218,185},{"peru",205,133,63},
  {"pink",255,192,203},{"plum",221,160,221},{"powderblue",176,224,230},{"purple",128,0,128},{"red",255,0,0},
  {"rosybrown",188,143,143},{"royalblue",65,105,225},{"saddlebrown",139,69,19},{"salmon",250,128,114},{"sandybrown",244,164,96},
  {"seagreen",46,139,87},{"seashell",255,245,238},{"sienna",160,82,45},{"silver",192,192,192},{"skyblue",135,206,235},
  {"slateblue",106,90,205},{"slategray",112,128,144},{"slategrey",112,128,144},{"snow",255,250,250},{"springgreen",0,255,127},
  {"steelblue",70,130,180},{"tan",210,180,140},{"teal",0,128,128},{"thistle",216,191,216},{"tomato",255,99,71},
  {"turquoise",64,224,208},{"violet",238,130,238},{"wheat",245,222,179},{"white",255,255,255},{"whitesmoke",245,245,245},
  {"yellow",255,255,0},{"yellowgreen",154,205,50},
};
const char *orc_css_similar(unsigned r, unsigned g, unsigned b) {
  long best = 0; const char *name = "";
  for (size_t i = 0; i < sizeof orc_css / sizeof orc_css[0]; i++) {
    long dr = (long)r - orc_css[i].r, dg = (long)g - orc_css[i].g, db = (long)b - orc_css[i].b;
    long d = dr * dr + dg * dg + db * db;
    if (i == 0 || d < best) { best = d; name = orc_css[i].n; }
  }
  return name;
}
