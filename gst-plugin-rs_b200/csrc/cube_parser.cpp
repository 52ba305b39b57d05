// cube_parser.cpp -- Adobe .cube parser of the colorlut element (host side, product code).
//
// Same grammar and the same accept/reject decisions as CubeLut::parse in
// gst-plugins-rs video/colorlut/src/parser.rs:105-375, written as a small line/token scanner:
//   * lines: split on '\n' (a preceding '\r' belongs to the line ending), Unicode-trimmed
//   * blank lines and lines whose first character is '#' are skipped
//   * keywords TITLE, DOMAIN_MIN, DOMAIN_MAX, LUT_1D_SIZE, LUT_3D_SIZE are header lines; a header
//     line after the first data line is an error; a second size keyword is an error
//   * every other line is a data line of exactly three Rust-syntax f32 literals
//   * 1D size 2..=65536, 3D size 2..=256, value count must match, domain min < max per channel
//   * the file must be valid UTF-8 (fs::read_to_string)
#include <locale.h>

#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <string_view>
#include <vector>

#include "../../include/b200vfx.h"

namespace {

using sv = std::string_view;

// decode one scalar value; returns its byte length, 0 on malformed input
int decode_utf8(sv s, char32_t &cp) {
  if (s.empty()) return 0;
  const auto b = [&](size_t i) { return (unsigned char)s[i]; };
  const auto cont = [&](size_t i) { return i < s.size() && (b(i) & 0xC0) == 0x80; };
  const unsigned char c = b(0);
  if (c < 0x80) { cp = c; return 1; }
  if (c >= 0xC2 && c <= 0xDF && cont(1)) { cp = ((c & 0x1Fu) << 6) | (b(1) & 0x3Fu); return 2; }
  if (c >= 0xE0 && c <= 0xEF && cont(1) && cont(2)) {
    if ((c == 0xE0 && b(1) < 0xA0) || (c == 0xED && b(1) > 0x9F)) return 0;
    cp = ((c & 0x0Fu) << 12) | ((b(1) & 0x3Fu) << 6) | (b(2) & 0x3Fu);
    return 3;
  }
  if (c >= 0xF0 && c <= 0xF4 && cont(1) && cont(2) && cont(3)) {
    if ((c == 0xF0 && b(1) < 0x90) || (c == 0xF4 && b(1) > 0x8F)) return 0;
    cp = ((c & 0x07u) << 18) | ((b(1) & 0x3Fu) << 12) | ((b(2) & 0x3Fu) << 6) | (b(3) & 0x3Fu);
    return 4;
  }
  return 0;
}

bool white(char32_t c) {  // Unicode White_Space (char::is_whitespace)
  return (c >= 0x09 && c <= 0x0D) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 ||
         (c >= 0x2000 && c <= 0x200A) || c == 0x2028 || c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}

// str::split_whitespace
std::vector<sv> tokens(sv line) {
  std::vector<sv> out;
  size_t i = 0, start = sv::npos;
  while (i < line.size()) {
    char32_t cp = 0;
    const int n = decode_utf8(line.substr(i), cp);
    if (white(cp)) {
      if (start != sv::npos) { out.push_back(line.substr(start, i - start)); start = sv::npos; }
    } else if (start == sv::npos) {
      start = i;
    }
    i += (size_t)(n > 0 ? n : 1);
  }
  if (start != sv::npos) out.push_back(line.substr(start));
  return out;
}

bool ieq(sv a, const char *lower) {
  const size_t n = std::strlen(lower);
  if (a.size() != n) return false;
  for (size_t i = 0; i < n; i++) {
    char c = a[i];
    if (c >= 'A' && c <= 'Z') c = (char)(c + 32);
    if (c != lower[i]) return false;
  }
  return true;
}

// <f32 as FromStr>: [+-]? ( "inf" | "infinity" | "nan" | digits? ('.' digits?)? ([eE][+-]?digits)? ) with >=1 digit
bool parse_float(sv t, float &out) {
  size_t i = 0;
  bool neg = false;
  if (!t.empty() && (t[0] == '+' || t[0] == '-')) { neg = t[0] == '-'; i = 1; }
  const sv body = t.substr(i);
  if (body.empty()) return false;
  if (ieq(body, "inf") || ieq(body, "infinity")) { out = neg ? -INFINITY : INFINITY; return true; }
  if (ieq(body, "nan")) { out = NAN; return true; }
  size_t digits = 0, j = 0;
  const auto isd = [&](size_t k) { return k < body.size() && body[k] >= '0' && body[k] <= '9'; };
  while (isd(j)) { j++; digits++; }
  if (j < body.size() && body[j] == '.') { j++; while (isd(j)) { j++; digits++; } }
  if (digits == 0) return false;
  if (j < body.size() && (body[j] == 'e' || body[j] == 'E')) {
    j++;
    if (j < body.size() && (body[j] == '+' || body[j] == '-')) j++;
    if (!isd(j)) return false;
    while (isd(j)) j++;
  }
  if (j != body.size()) return false;
  // correctly rounded like Rust's dec2flt -- and, like it, independent of the process locale: a GTK / GStreamer
  // application that called setlocale(LC_ALL, "") under de_DE would make plain strtof stop at the '.'
  static const locale_t c_locale = newlocale(LC_ALL_MASK, "C", (locale_t)0);
  const std::string z(t);
  char *end = nullptr;
  out = c_locale ? strtof_l(z.c_str(), &end, c_locale) : std::strtof(z.c_str(), &end);
  return end == z.c_str() + z.size();   // the grammar check above already accepted the whole token
}

// <usize as FromStr>: '+'? digits+, overflow rejected
bool parse_size(sv t, unsigned long long &out) {
  size_t i = (!t.empty() && t[0] == '+') ? 1 : 0;
  if (i >= t.size()) return false;
  unsigned long long v = 0;
  for (; i < t.size(); i++) {
    if (t[i] < '0' || t[i] > '9') return false;
    const unsigned d = (unsigned)(t[i] - '0');
    if (v > (~0ull - d) / 10ull) return false;
    v = v * 10ull + d;
  }
  out = v;
  return true;
}

struct ParseError { int code; std::string msg; };

ParseError invalid(const std::string &m) { return {B200VFX_ERR_PARSE, "Invalid LUT: " + m}; }

struct Parsed {
  int kind = 0;
  size_t size = 0;
  std::vector<float> values;  // n x 3
  float dmin[3] = {0, 0, 0}, dmax[3] = {1, 1, 1};
};

bool parse_text(sv text, Parsed &out, ParseError &err) {
  {  // read_to_string: the whole file must be UTF-8
    size_t i = 0;
    while (i < text.size()) {
      char32_t cp;
      const int n = decode_utf8(text.substr(i), cp);
      if (n == 0) { err = {B200VFX_ERR_IO, "IO error: stream did not contain valid UTF-8"}; return false; }
      i += (size_t)n;
    }
  }
  enum class St { Header, Sized, Data } st = St::Header;
  size_t line_no = 0, pos = 0;
  while (pos < text.size()) {
    size_t nl = text.find('\n', pos);
    if (nl == sv::npos) nl = text.size();
    sv line = text.substr(pos, nl - pos);
    pos = nl + 1;
    line_no++;
    const std::vector<sv> tk = tokens(line);
    if (tk.empty() || tk[0][0] == '#') continue;
    const std::string where = "line " + std::to_string(line_no) + ": " + std::string(tk.front().data(), (size_t)((tk.back().data() + tk.back().size()) - tk.front().data()));
    const sv kw = tk[0];
    const bool is_title = kw == "TITLE", is_min = kw == "DOMAIN_MIN", is_max = kw == "DOMAIN_MAX",
               is_1d = kw == "LUT_1D_SIZE", is_3d = kw == "LUT_3D_SIZE";
    if (is_title || is_min || is_max || is_1d || is_3d) {
      if (st == St::Data) { err = invalid("Header found after LUT data at " + where); return false; }
      if (is_title) continue;
      if (is_min || is_max) {
        float v[3];
        for (int k = 0; k < 3; k++) {
          if ((size_t)k + 1 >= tk.size()) { err = invalid("Invalid " + where); return false; }
          if (!parse_float(tk[(size_t)k + 1], v[k])) { err = invalid("Invalid float at " + where); return false; }
        }
        if (tk.size() > 4) { err = invalid("Invalid " + where); return false; }
        std::memcpy(is_min ? out.dmin : out.dmax, v, sizeof v);
        continue;
      }
      if (st != St::Header) { err = invalid(std::string("Invalid ") + (is_1d ? "LUT_1D_SIZE" : "LUT_3D_SIZE") + " at " + where); return false; }
      unsigned long long sz = 0;
      if (tk.size() < 2) { err = invalid("Invalid " + where); return false; }
      if (!parse_size(tk[1], sz)) { err = invalid("Invalid integer at " + where); return false; }
      if (tk.size() > 2) { err = invalid("Invalid " + where); return false; }
      const unsigned long long lo = 2, hi = is_1d ? 65536 : 256;
      if (sz < lo || sz > hi) {
        err = invalid("Invalid LUT size " + std::to_string(sz) + " at line " + std::to_string(line_no) + ", expected " +
                      std::to_string(lo) + "..=" + std::to_string(hi));
        return false;
      }
      out.kind = is_1d ? 1 : 3;
      out.size = (size_t)sz;
      st = St::Sized;
      continue;
    }
    if (st == St::Header) { err = invalid("LUT data found before LUT size at " + where); return false; }
    st = St::Data;
    float v[3];
    for (int k = 0; k < 3; k++) {
      if ((size_t)k >= tk.size()) { err = invalid("Invalid " + where); return false; }
      if (!parse_float(tk[(size_t)k], v[k])) { err = invalid("Invalid float at " + where); return false; }
    }
    if (tk.size() > 3) { err = invalid("Invalid " + where); return false; }
    out.values.insert(out.values.end(), v, v + 3);
  }
  for (int c = 0; c < 3; c++)
    if (out.dmin[c] >= out.dmax[c]) { err = invalid("Invalid domain min/max"); return false; }
  if (st == St::Header) { err = invalid("Missing LUT size"); return false; }
  const size_t n = out.values.size() / 3;
  const size_t expected = out.kind == 1 ? out.size : out.size * out.size * out.size;
  if (n != expected) {
    err = invalid(std::string("Invalid ") + (out.kind == 1 ? "1D" : "3D") + " LUT value count, expected " +
                  std::to_string(expected) + ", got " + std::to_string(n));
    return false;
  }
  return true;
}

void put_err(char *err, size_t errlen, const std::string &m) {
  if (err && errlen) { std::snprintf(err, errlen, "%s", m.c_str()); }
}

}  // namespace

extern "C" {

int b200vfx_cube_parse(const char *text, size_t len, int *kind, int *size, float **values, float scale[3],
                       float offset[3], char *err, size_t errlen) {
  if (!text || !kind || !size || !values || !scale || !offset) { put_err(err, errlen, "null argument"); return B200VFX_ERR_INVALID; }
  Parsed p;
  ParseError e{0, ""};
  if (!parse_text(sv(text, len), p, e)) { put_err(err, errlen, e.msg); return e.code; }
  for (int c = 0; c < 3; c++) {  // parser.rs:264-274
    scale[c] = 1.0f / (p.dmax[c] - p.dmin[c]);
    offset[c] = -p.dmin[c] * scale[c];
  }
  float *v = (float *)std::malloc(std::max<size_t>(p.values.size(), 1) * sizeof(float));
  if (!v) { put_err(err, errlen, "out of memory"); return B200VFX_ERR_INVALID; }
  std::memcpy(v, p.values.data(), p.values.size() * sizeof(float));
  *values = v;
  *kind = p.kind;
  *size = (int)p.size;
  return 0;
}

int b200vfx_cube_parse_file(const char *path, int *kind, int *size, float **values, float scale[3], float offset[3],
                            char *err, size_t errlen) {
  if (!path) { put_err(err, errlen, "null path"); return B200VFX_ERR_INVALID; }
  std::FILE *f = std::fopen(path, "rb");
  if (!f) { put_err(err, errlen, std::string("IO error: ") + std::strerror(errno)); return B200VFX_ERR_IO; }
  std::string data;
  char buf[1 << 16];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, n);
  const bool bad = std::ferror(f) != 0;
  std::fclose(f);
  if (bad) { put_err(err, errlen, "IO error: read failed"); return B200VFX_ERR_IO; }
  return b200vfx_cube_parse(data.data(), data.size(), kind, size, values, scale, offset, err, errlen);
}

void b200vfx_cube_free(float *values) { std::free(values); }

}  // extern "C"
