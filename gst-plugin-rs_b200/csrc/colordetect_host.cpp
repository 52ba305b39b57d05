// colordetect_host.cpp -- host half of `colordetect` (include/b200vfx.h): modified median cut over the 32768-bin
// histogram the GPU produced, and the nearest CSS colour name.
//
// Reference call sites: video/videofx/src/colordetect/imp.rs:68-79
//     get_palette(plane, color_format, quality, max_colors)      -> color-thief 0.2.2  (Cargo.lock:2045-2053)
//     color_name::css::Color::similar([r,g,b]).to_lowercase()    -> color-name 1.2.0
// Both crates are third-party code that is NOT under /root/reference; what follows restates their published
// algorithm (MMCQ as ported from Leptonica / quantize.js through color-thief-java) from memory.
// PARITY UNPINNED beyond the reference's only test (tests/colordetect.rs:67: a red frame names "red").
// Everything here runs once per frame on 32768 integers -- it is not on the per-pixel path.
#include "../../include/b200vfx.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

constexpr int kSignalBits = 5;
constexpr int kRightShift = 8 - kSignalBits;
constexpr int kMultiplier = 1 << kRightShift;
constexpr int kVBoxLength = 1 << kSignalBits;
constexpr double kFractionByPopulation = 0.75;
constexpr int kMaxIterations = 1000;

inline int color_index(int r, int g, int b) { return (r << (2 * kSignalBits)) + (g << kSignalBits) + b; }

struct VBox {
  int r_min, r_max, g_min, g_max, b_min, b_max;  // 5-bit channel ranges (r_min 255 / r_max 0 when nothing was counted)
  int avg[3] = {0, 0, 0};
  int volume = 0;
  long long count = 0;

  void recalc(const uint32_t *h) {
    // average: per-bin truncation of hval * (i + 0.5) * 8 to an integer, then integer division by the population
    long long ntot = 0, rs = 0, gs = 0, bs = 0;
    for (int i = r_min; i <= r_max; i++)
      for (int j = g_min; j <= g_max; j++)
        for (int k = b_min; k <= b_max; k++) {
          const double hv = (double)h[color_index(i, j, k)];
          ntot += (long long)hv;
          rs += (long long)(hv * ((double)i + 0.5) * (double)kMultiplier);
          gs += (long long)(hv * ((double)j + 0.5) * (double)kMultiplier);
          bs += (long long)(hv * ((double)k + 0.5) * (double)kMultiplier);
        }
    if (ntot > 0) {
      avg[0] = (int)(rs / ntot) & 255; avg[1] = (int)(gs / ntot) & 255; avg[2] = (int)(bs / ntot) & 255;
    } else {
      avg[0] = std::min(kMultiplier * (r_min + r_max + 1) / 2, 255);
      avg[1] = std::min(kMultiplier * (g_min + g_max + 1) / 2, 255);
      avg[2] = std::min(kMultiplier * (b_min + b_max + 1) / 2, 255);
    }
    count = ntot;
    volume = (r_max - r_min + 1) * (g_max - g_min + 1) * (b_max - b_min + 1);
  }
  int widest() const {  // 0 red, 1 green, 2 blue; ties resolve red, green, blue
    const int rw = r_max - r_min, gw = g_max - g_min, bw = b_max - b_min;
    const int m = std::max(std::max(rw, gw), bw);
    return m == rw ? 0 : (m == gw ? 1 : 2);
  }
};

// returns false when the box cannot be cut (population 0 or 1)
bool median_cut(const uint32_t *h, const VBox &box, VBox *a, VBox *b) {
  if (box.count <= 1) return false;
  const int axis = box.widest();
  long long partial[kVBoxLength], ahead[kVBoxLength];
  for (int i = 0; i < kVBoxLength; i++) partial[i] = ahead[i] = -1;
  long long total = 0;
  int lo, hi;
  if (axis == 0) {
    lo = box.r_min; hi = box.r_max;
    for (int i = lo; i <= hi; i++) {
      long long s = 0;
      for (int j = box.g_min; j <= box.g_max; j++) for (int k = box.b_min; k <= box.b_max; k++) s += h[color_index(i, j, k)];
      total += s; partial[i] = total;
    }
  } else if (axis == 1) {
    lo = box.g_min; hi = box.g_max;
    for (int i = lo; i <= hi; i++) {
      long long s = 0;
      for (int j = box.r_min; j <= box.r_max; j++) for (int k = box.b_min; k <= box.b_max; k++) s += h[color_index(j, i, k)];
      total += s; partial[i] = total;
    }
  } else {
    lo = box.b_min; hi = box.b_max;
    for (int i = lo; i <= hi; i++) {
      long long s = 0;
      for (int j = box.r_min; j <= box.r_max; j++) for (int k = box.g_min; k <= box.g_max; k++) s += h[color_index(j, k, i)];
      total += s; partial[i] = total;
    }
  }
  for (int i = 0; i < kVBoxLength; i++) if (partial[i] != -1) ahead[i] = total - partial[i];
  for (int i = lo; i <= hi; i++) {
    if (partial[i] > total / 2) {
      const int left = i - lo, right = hi - i;
      int d2;
      if (left <= right) d2 = std::min(hi - 1, i + right / 2);
      else d2 = std::max(lo, (int)((double)(i - 1) - (double)left / 2.0));
      while (d2 < 0 || partial[d2] <= 0) d2++;                       // avoid 0-count boxes
      long long c2 = ahead[d2];
      while (c2 == 0 && d2 > 0 && partial[d2 - 1] > 0) { d2--; c2 = ahead[d2]; }
      *a = box; *b = box;
      if (axis == 0) { a->r_max = d2; b->r_min = d2 + 1; }
      else if (axis == 1) { a->g_max = d2; b->g_min = d2 + 1; }
      else { a->b_max = d2; b->b_min = d2 + 1; }
      a->recalc(h); b->recalc(h);
      return true;
    }
  }
  return false;
}

bool by_count(const VBox &x, const VBox &y) { return x.count < y.count; }
bool by_product(const VBox &x, const VBox &y) {
  if (x.count == y.count) return x.volume < y.volume;
  return x.count * (long long)x.volume < y.count * (long long)y.volume;
}

template <typename Cmp>
void iterate(std::vector<VBox> &q, Cmp cmp, int target, const uint32_t *h) {
  int color = 1;
  for (int it = 0; it < kMaxIterations; it++) {
    if (q.empty()) return;
    VBox box = q.back();
    if (box.count == 0) { std::stable_sort(q.begin(), q.end(), cmp); continue; }
    q.pop_back();
    VBox a, b;
    if (median_cut(h, box, &a, &b)) { q.push_back(a); q.push_back(b); color++; }
    else q.push_back(box);
    std::stable_sort(q.begin(), q.end(), cmp);  // slice::sort_by is stable
    if (color >= target) return;
  }
}

struct Css { const char *name; uint8_t r, g, b; };
// CSS Color Module Level 3 extended keywords, alphabetical (the first of two names for one RGB wins a tie)
const Css kCss[] = {
    {"aliceblue", 240, 248, 255}, {"antiquewhite", 250, 235, 215}, {"aqua", 0, 255, 255}, {"aquamarine", 127, 255, 212},
    {"azure", 240, 255, 255}, {"beige", 245, 245, 220}, {"bisque", 255, 228, 196}, {"black", 0, 0, 0},
    {"blanchedalmond", 255, 235, 205}, {"blue", 0, 0, 255}, {"blueviolet", 138, 43, 226}, {"brown", 165, 42, 42},
    {"burlywood", 222, 184, 135}, {"cadetblue", 95, 158, 160}, {"chartreuse", 127, 255, 0}, {"chocolate", 210, 105, 30},
    {"coral", 255, 127, 80}, {"cornflowerblue", 100, 149, 237}, {"cornsilk", 255, 248, 220}, {"crimson", 220, 20, 60},
    {"cyan", 0, 255, 255}, {"darkblue", 0, 0, 139}, {"darkcyan", 0, 139, 139}, {"darkgoldenrod", 184, 134, 11},
    {"darkgray", 169, 169, 169}, {"darkgreen", 0, 100, 0}, {"darkgrey", 169, 169, 169}, {"darkkhaki", 189, 183, 107},
    {"darkmagenta", 139, 0, 139}, {"darkolivegreen", 85, 107, 47}, {"darkorange", 255, 140, 0}, {"darkorchid", 153, 50, 204},
    {"darkred", 139, 0, 0}, {"darksalmon", 233, 150, 122}, {"darkseagreen", 143, 188, 143}, {"darkslateblue", 72, 61, 139},
    {"darkslategray", 47, 79, 79}, {"darkslategrey", 47, 79, 79}, {"darkturquoise", 0, 206, 209}, {"darkviolet", 148, 0, 211},
    {"deeppink", 255, 20, 147}, {"deepskyblue", 0, 191, 255}, {"dimgray", 105, 105, 105}, {"dimgrey", 105, 105, 105},
    {"dodgerblue", 30, 144, 255}, {"firebrick", 178, 34, 34}, {"floralwhite", 255, 250, 240}, {"forestgreen", 34, 139, 34},
    {"fuchsia", 255, 0, 255}, {"gainsboro", 220, 220, 220}, {"ghostwhite", 248, 248, 255}, {"gold", 255, 215, 0},
    {"goldenrod", 218, 165, 32}, {"gray", 128, 128, 128}, {"green", 0, 128, 0}, {"greenyellow", 173, 255, 47},
    {"grey", 128, 128, 128}, {"honeydew", 240, 255, 240}, {"hotpink", 255, 105, 180}, {"indianred", 205, 92, 92},
    {"indigo", 75, 0, 130}, {"ivory", 255, 255, 240}, {"khaki", 240, 230, 140}, {"lavender", 230, 230, 250},
    {"lavenderblush", 255, 240, 245}, {"lawngreen", 124, 252, 0}, {"lemonchiffon", 255, 250, 205}, {"lightblue", 173, 216, 230},
    {"lightcoral", 240, 128, 128}, {"lightcyan", 224, 255, 255}, {"lightgoldenrodyellow", 250, 250, 210}, {"lightgray", 211, 211, 211},
    {"lightgreen", 144, 238, 144}, {"lightgrey", 211, 211, 211}, {"lightpink", 255, 182, 193}, {"lightsalmon", 255, 160, 122},
    {"lightseagreen", 32, 178, 170}, {"lightskyblue", 135, 206, 250}, {"lightslategray", 119, 136, 153}, {"lightslategrey", 119, 136, 153},
    {"lightsteelblue", 176, 196, 222}, {"lightyellow", 255, 255, 224}, {"lime", 0, 255, 0}, {"limegreen", 50, 205, 50},
    {"linen", 250, 240, 230}, {"magenta", 255, 0, 255}, {"maroon", 128, 0, 0}, {"mediumaquamarine", 102, 205, 170},
    {"mediumblue", 0, 0, 205}, {"mediumorchid", 186, 85, 211}, {"mediumpurple", 147, 112, 219}, {"mediumseagreen", 60, 179, 113},
    {"mediumslateblue", 123, 104, 238}, {"mediumspringgreen", 0, 250, 154}, {"mediumturquoise", 72, 209, 204}, {"mediumvioletred", 199, 21, 133},
    {"midnightblue", 25, 25, 112}, {"mintcream", 245, 255, 250}, {"mistyrose", 255, 228, 225}, {"moccasin", 255, 228, 181},
    {"navajowhite", 255, 222, 173}, {"navy", 0, 0, 128}, {"oldlace", 253, 245, 230}, {"olive", 128, 128, 0},
    {"olivedrab", 107, 142, 35}, {"orange", 255, 165, 0}, {"orangered", 255, 69, 0}, {"orchid", 218, 112, 214},
    {"palegoldenrod", 238, 232, 170}, {"palegreen", 152, 251, 152}, {"paleturquoise", 175, 238, 238}, {"palevioletred", 219, 112, 147},
    {"papayawhip", 255, 239, 213}, {"peachpuff", 255, 218, 185}, {"peru", 205, 133, 63}, {"pink", 255, 192, 203},
    {"plum", 221, 160, 221}, {"powderblue", 176, 224, 230}, {"purple", 128, 0, 128}, {"red", 255, 0, 0},
    {"rosybrown", 188, 143, 143}, {"royalblue", 65, 105, 225}, {"saddlebrown", 139, 69, 19}, {"salmon", 250, 128, 114},
    {"sandybrown", 244, 164, 96}, {"seagreen", 46, 139, 87}, {"seashell", 255, 245, 238}, {"sienna", 160, 82, 45},
    {"silver", 192, 192, 192}, {"skyblue", 135, 206, 235}, {"slateblue", 106, 90, 205}, {"slategray", 112, 128, 144},
    {"slategrey", 112, 128, 144}, {"snow", 255, 250, 250}, {"springgreen", 0, 255, 127}, {"steelblue", 70, 130, 180},
    {"tan", 210, 180, 140}, {"teal", 0, 128, 128}, {"thistle", 216, 191, 216}, {"tomato", 255, 99, 71},
    {"turquoise", 64, 224, 208}, {"violet", 238, 130, 238}, {"wheat", 245, 222, 179}, {"white", 255, 255, 255},
    {"whitesmoke", 245, 245, 245}, {"yellow", 255, 255, 0}, {"yellowgreen", 154, 205, 50},
};

}  // namespace

extern "C" {

int b200vfx_colordetect_palette(const uint32_t *hist, int max_colors, uint8_t *palette_rgb, int palette_cap, int *n_colors) {
  if (!hist || !palette_rgb || !n_colors || max_colors < 2 || max_colors > 255) return B200VFX_ERR_INVALID;
  VBox box;
  box.r_min = box.g_min = box.b_min = 255; box.r_max = box.g_max = box.b_max = 0;
  for (int r = 0; r < kVBoxLength; r++)
    for (int g = 0; g < kVBoxLength; g++)
      for (int b = 0; b < kVBoxLength; b++)
        if (hist[color_index(r, g, b)]) {
          box.r_min = std::min(box.r_min, r); box.r_max = std::max(box.r_max, r);
          box.g_min = std::min(box.g_min, g); box.g_max = std::max(box.g_max, g);
          box.b_min = std::min(box.b_min, b); box.b_max = std::max(box.b_max, b);
        }
  box.recalc(hist);
  std::vector<VBox> q{box};
  const int target = (int)std::ceil(kFractionByPopulation * (double)max_colors);
  iterate(q, by_count, target, hist);                       // first set of colours, by population
  std::stable_sort(q.begin(), q.end(), by_product);         // re-sort by population x colour-space volume
  iterate(q, by_product, max_colors - (int)q.size(), hist); // next set, by the product
  std::reverse(q.begin(), q.end());                         // most significant box first
  const int n = (int)q.size();
  *n_colors = n;
  for (int i = 0; i < n && i < palette_cap; i++) {
    palette_rgb[3 * i] = (uint8_t)q[(size_t)i].avg[0];
    palette_rgb[3 * i + 1] = (uint8_t)q[(size_t)i].avg[1];
    palette_rgb[3 * i + 2] = (uint8_t)q[(size_t)i].avg[2];
  }
  return B200VFX_OK;
}

const char *b200vfx_css_color_similar(unsigned r, unsigned g, unsigned b) {
  long best = -1;
  const char *name = "";
  for (const Css &c : kCss) {
    const long dr = (long)r - c.r, dg = (long)g - c.g, db = (long)b - c.b;
    const long d = dr * dr + dg * dg + db * db;
    if (best < 0 || d < best) { best = d; name = c.name; }
  }
  return name;
}

}  // extern "C"
