// elements.cpp -- host-side element layer (include/b200gst.h).
//
// The reference's L2 (element shells in Rust over gstreamer-rs) cannot be built here, so the same
// surface is mirrored in C++ over a GstVideoFrame-like descriptor: factory names, GObject type
// names, property names/types/defaults/ranges, pad-template format lists, transform_caps rules,
// start/stop, set_caps, passthrough, and the per-frame vfuncs.  All pixel work goes through the
// C ABI of b200vfx.h (the element owns one b200vfx_ctx between start() and stop()).
#include "../../include/b200gst.h"
#include "../../include/b200vfx.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace {

enum class PType { String, Float, Double, UInt, Enum };

struct PropSpec {
  std::string name;
  PType type;
  double def, min, max;
  const char *mutability;  // "mutable-ready" | "mutable-playing"
  std::vector<std::pair<int, std::string>> enum_nicks;
};

struct PropValue {
  bool is_null = true;  // strings only
  std::string s;
  double d = 0;
};

const double FMAX = std::numeric_limits<float>::max();
const double DMAX = std::numeric_limits<double>::max();

}  // namespace

struct b200gst_element {
  std::string factory, type_name, plugin;
  std::vector<PropSpec> specs;
  std::map<std::string, PropValue> values;
  b200vfx_ctx *ctx = nullptr;
  bool started = false, passthrough = false;
  std::string err;
  std::deque<std::string> bus;

  virtual ~b200gst_element() { if (ctx) b200vfx_ctx_destroy(ctx); }

  void add_prop(PropSpec s) {
    PropValue v;
    if (s.type == PType::Float) s.def = (double)(float)s.def;  // gfloat ParamSpec: the default is an f32
    if (s.type != PType::String) { v.is_null = false; v.d = s.def; }
    values[s.name] = v;
    specs.push_back(std::move(s));
  }
  float f(const char *n) const { return (float)values.at(n).d; }
  int fail(int code, const std::string &m) { err = m; return code; }
  int ctx_error(int flow) { err = b200vfx_last_error(ctx); return flow; }

  virtual void property_changed(const std::string &, const PropValue & /*old*/) {}
  virtual std::vector<int> pad_formats(int direction) const = 0;
  virtual std::vector<int> transform_caps(int /*direction*/, const std::vector<int> &formats) {
    // default GstVideoFilter behaviour: same caps on both sides, restricted to the pad template
    std::vector<int> out, allowed = pad_formats(B200GST_PAD_SRC);
    for (int f : formats) if (std::find(allowed.begin(), allowed.end(), f) != allowed.end()) out.push_back(f);
    return out;
  }
  virtual int start() {
    if (started) return 0;
    if (b200vfx_ctx_create(&ctx, -1) != 0) return fail(-1, b200vfx_last_error(nullptr));
    started = true;
    return 0;
  }
  virtual int stop() {
    if (ctx) { b200vfx_ctx_destroy(ctx); ctx = nullptr; }
    started = false;
    return 0;
  }
  virtual int set_caps(int, int, int, int) { return 0; }
  virtual int transform_frame(const b200gst_video_frame *, b200gst_video_frame *) {
    return fail(B200GST_FLOW_ERROR, factory + " has no transform_frame");
  }
  virtual int transform_frame_ip(b200gst_video_frame *) {
    return fail(B200GST_FLOW_ERROR, factory + " has no transform_frame_ip");
  }
};

namespace {

bool contains(const std::vector<int> &v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); }

// ---- colorlut ---------------------------------------------------------------------------------
struct ColorLut : b200gst_element {  // video/colorlut/src/colorlut/imp.rs
  ColorLut() {
    factory = "colorlut"; type_name = "GstColorLut"; plugin = "colorlut";
    add_prop({"location", PType::String, 0, 0, 0, "mutable-ready", {}});  // imp.rs:72-76
  }
  std::vector<int> pad_formats(int) const override {  // imp.rs:122-137 (little-endian order)
    return {B200VFX_FORMAT_RGBA64_LE, B200VFX_FORMAT_RGBA64_BE, B200VFX_FORMAT_RGBA};
  }
  int start() override {  // imp.rs:168-194
    const PropValue &loc = values["location"];
    if (loc.is_null) return fail(-1, "ResourceError::Settings: LUT file location is not configured");
    if (int rc = b200gst_element::start()) return rc;
    if (b200vfx_colorlut_load_file(ctx, loc.s.c_str()) != 0) {
      err = std::string("ResourceError::Read: ") + b200vfx_last_error(ctx);
      b200gst_element::stop();
      return -1;
    }
    return 0;
  }
  int transform_frame(const b200gst_video_frame *in, b200gst_video_frame *out) override {  // imp.rs:203-224
    if (!started || !ctx) return fail(B200GST_FLOW_ERROR, "No LUT configured");
    if (!contains(pad_formats(0), in->format) || in->format != out->format) return fail(B200GST_FLOW_NOT_NEGOTIATED, "format not negotiated");
    if (in->width != out->width || in->height != out->height) return fail(B200GST_FLOW_NOT_NEGOTIATED, "size mismatch");
    if (b200vfx_colorlut_process(ctx, in->format, in->width, in->height, in->data[0], in->stride[0], out->data[0], out->stride[0]) != 0)
      return ctx_error(B200GST_FLOW_ERROR);
    return B200GST_FLOW_OK;
  }
};

// ---- hsvfilter --------------------------------------------------------------------------------
struct HsvFilter : b200gst_element {  // video/hsv/src/hsvfilter/imp.rs
  HsvFilter() {
    factory = "hsvfilter"; type_name = "GstHsvFilter"; plugin = "hsv";
    add_prop({"hue-shift", PType::Float, 0.0, -FMAX, FMAX, "mutable-playing", {}});       // imp.rs:127-156
    add_prop({"saturation-mul", PType::Float, 1.0, -FMAX, FMAX, "mutable-playing", {}});
    add_prop({"saturation-off", PType::Float, 0.0, -FMAX, FMAX, "mutable-playing", {}});
    add_prop({"value-mul", PType::Float, 1.0, -FMAX, FMAX, "mutable-playing", {}});
    add_prop({"value-off", PType::Float, 0.0, -FMAX, FMAX, "mutable-playing", {}});
  }
  std::vector<int> pad_formats(int) const override {  // imp.rs:278-289
    return {B200VFX_FORMAT_RGBX, B200VFX_FORMAT_XRGB, B200VFX_FORMAT_BGRX, B200VFX_FORMAT_XBGR, B200VFX_FORMAT_RGBA,
            B200VFX_FORMAT_ARGB, B200VFX_FORMAT_BGRA, B200VFX_FORMAT_ABGR, B200VFX_FORMAT_RGB, B200VFX_FORMAT_BGR};
  }
  int transform_frame_ip(b200gst_video_frame *fr) override {  // imp.rs:323-376, settings snapshot :85
    if (!started) return fail(B200GST_FLOW_ERROR, "not started");
    if (!contains(pad_formats(0), fr->format)) return fail(B200GST_FLOW_NOT_NEGOTIATED, "format not negotiated");
    if (b200vfx_hsvfilter_process(ctx, fr->format, fr->width, fr->height, fr->data[0], fr->stride[0], f("hue-shift"),
                                  f("saturation-mul"), f("saturation-off"), f("value-mul"), f("value-off")) != 0)
      return ctx_error(B200GST_FLOW_ERROR);
    return B200GST_FLOW_OK;
  }
};

// ---- hsvdetector ------------------------------------------------------------------------------
struct HsvDetector : b200gst_element {  // video/hsv/src/hsvdetector/imp.rs
  HsvDetector() {
    factory = "hsvdetector"; type_name = "GstHsvDetector"; plugin = "hsv";
    add_prop({"hue-ref", PType::Float, 0.0, -FMAX, FMAX, "mutable-playing", {}});  // imp.rs:167-212
    add_prop({"hue-var", PType::Float, 10.0, 0.0, 180.0, "mutable-playing", {}});
    add_prop({"saturation-ref", PType::Float, 0.0, 0.0, 1.0, "mutable-playing", {}});
    add_prop({"saturation-var", PType::Float, 0.15, 0.0, 1.0, "mutable-playing", {}});
    add_prop({"value-ref", PType::Float, 0.0, 0.0, 1.0, "mutable-playing", {}});
    add_prop({"value-var", PType::Float, 0.3, 0.0, 1.0, "mutable-playing", {}});
  }
  static std::vector<int> in_formats() {  // imp.rs:78-87
    return {B200VFX_FORMAT_RGBX, B200VFX_FORMAT_XRGB, B200VFX_FORMAT_BGRX, B200VFX_FORMAT_XBGR, B200VFX_FORMAT_RGB, B200VFX_FORMAT_BGR};
  }
  static std::vector<int> out_formats() {  // imp.rs:89-96
    return {B200VFX_FORMAT_RGBA, B200VFX_FORMAT_ARGB, B200VFX_FORMAT_BGRA, B200VFX_FORMAT_ABGR};
  }
  std::vector<int> pad_formats(int dir) const override { return dir == B200GST_PAD_SINK ? in_formats() : out_formats(); }
  std::vector<int> transform_caps(int direction, const std::vector<int> &formats) override {  // imp.rs:386-419
    if (formats.empty()) return {};
    return direction == B200GST_PAD_SRC ? in_formats() : out_formats();  // the format field is replaced wholesale
  }
  int transform_frame(const b200gst_video_frame *in, b200gst_video_frame *out) override {  // imp.rs:423-707
    if (!started) return fail(B200GST_FLOW_ERROR, "not started");
    if (!contains(in_formats(), in->format) || !contains(out_formats(), out->format)) return fail(B200GST_FLOW_NOT_NEGOTIATED, "format not negotiated");
    if (in->width != out->width || in->height != out->height) return fail(B200GST_FLOW_NOT_NEGOTIATED, "size mismatch");
    if (b200vfx_hsvdetector_process(ctx, in->format, out->format, in->width, in->height, in->data[0], in->stride[0], out->data[0],
                                    out->stride[0], f("hue-ref"), f("hue-var"), f("saturation-ref"), f("saturation-var"),
                                    f("value-ref"), f("value-var")) != 0)
      return ctx_error(B200GST_FLOW_ERROR);
    return B200GST_FLOW_OK;
  }
};

// ---- roundedcorners ---------------------------------------------------------------------------
inline int round_up_4(int v) { return (v + 3) & ~3; }

struct RoundedCorners : b200gst_element {  // video/videofx/src/border/imp.rs
  bool changed = false, have_state = false;
  int out_format = B200VFX_FORMAT_I420, width = 0, height = 0, alpha_stride = 0;
  std::vector<uint8_t> alpha_mem;  // State::alpha_mem (:469-475): shared by every output buffer
  unsigned generated = 0;          // how often the mask was (re)generated -- once per caps/radius change
  RoundedCorners() {
    factory = "roundedcorners"; type_name = "GstRoundedCorners"; plugin = "rsvideofx";
    add_prop({"border-radius-px", PType::UInt, 0, 0, 4294967295.0, "mutable-playing", {}});  // :282-287
  }
  void property_changed(const std::string &, const PropValue &old) override {  // :299-310
    if (old.d != values["border-radius-px"].d) changed = true;                  // + reconfigure_src()
  }
  std::vector<int> pad_formats(int dir) const override {  // :343-371
    if (dir == B200GST_PAD_SINK) return {B200VFX_FORMAT_I420};
    return {B200VFX_FORMAT_I420, B200VFX_FORMAT_A420};
  }
  std::vector<int> transform_caps(int direction, const std::vector<int> &formats) override {  // :388-442
    if (formats.empty()) return {};
    if (direction == B200GST_PAD_SRC) return {B200VFX_FORMAT_I420};
    if (values["border-radius-px"].d == 0) return {B200VFX_FORMAT_I420, B200VFX_FORMAT_A420};
    return {B200VFX_FORMAT_A420};
  }
  int set_caps(int in_format, int out_fmt, int w, int h) override {  // :444-480
    if (in_format != B200VFX_FORMAT_I420 || (out_fmt != B200VFX_FORMAT_I420 && out_fmt != B200VFX_FORMAT_A420))
      return fail(-1, "Failed to parse output caps");
    out_format = out_fmt; width = w; height = h;
    if (out_fmt == B200VFX_FORMAT_I420) { passthrough = true; return 0; }
    passthrough = false;
    alpha_stride = round_up_4(w);  // GstVideoInfo stride[3] of A420
    alpha_mem.assign((size_t)alpha_stride * (size_t)((h + 1) & ~1), 0);
    have_state = true;
    changed = true;
    return 0;
  }
  int stop() override { have_state = false; alpha_mem.clear(); return b200gst_element::stop(); }
  int prepare_output(const b200gst_video_frame *in, b200gst_video_frame *out) {  // :482-559
    if (passthrough) { *out = *in; return B200GST_FLOW_OK; }
    if (changed) {
      changed = false;
      if (!have_state) return fail(B200GST_FLOW_NOT_NEGOTIATED, "Have no state yet");
      if (!started) return fail(B200GST_FLOW_ERROR, "not started");
      const unsigned radius = (unsigned)values["border-radius-px"].d;
      if (b200vfx_roundmask_generate(ctx, width, height, alpha_stride, radius, alpha_mem.data()) != 0)
        return ctx_error(B200GST_FLOW_NOT_NEGOTIATED);  // "Failed to generate alpha mask"
      generated++;
    }
    if (!have_state) return fail(B200GST_FLOW_NOT_NEGOTIATED, "Have no state yet");
    *out = *in;  // same memories ...
    out->format = B200VFX_FORMAT_A420;
    out->n_planes = 4;  // ... plus the shared alpha memory appended as plane 3 (add_video_meta :182-268)
    out->data[3] = alpha_mem.data();
    out->stride[3] = alpha_stride;
    return B200GST_FLOW_OK;
  }
};

// ---- videocompare -----------------------------------------------------------------------------
struct VideoCompare : b200gst_element {  // video/videofx/src/videocompare/imp.rs
  std::vector<int> pads;  // request pads sink_%u in creation order
  int next_pad = 0, reference_pad = -1;
  VideoCompare() {
    factory = "videocompare"; type_name = "GstVideoCompare"; plugin = "rsvideofx";
    add_prop({"hash-algo", PType::Enum, 4, 0, 4, "mutable-ready",  // mod.rs:57-92, imp.rs:78-82
              {{0, "mean"}, {1, "gradient"}, {2, "vertgradient"}, {3, "doublegradient"}, {4, "blockhash"}}});
    add_prop({"max-dist-threshold", PType::Double, 0.0, 0.0, DMAX, "mutable-ready", {}});  // imp.rs:83-89
  }
  std::vector<int> pad_formats(int) const override { return {B200VFX_FORMAT_RGB, B200VFX_FORMAT_RGBA}; }  // :159-171
  int request_pad() {  // create_new_pad :216-231: the first requested sink pad becomes the reference
    const int id = next_pad++;
    pads.push_back(id);
    if (reference_pad < 0) reference_pad = id;
    return id;
  }
  int release_pad(int id) {  // release_pad :188-206
    if (!contains(pads, id)) return fail(-1, "no such pad");
    if (reference_pad == id)
      for (int p : pads) if (p != id) reference_pad = p;  // last other pad wins, as in the reference loop
    pads.erase(std::find(pads.begin(), pads.end(), id));
    return 0;
  }
  int hash(const b200gst_video_frame &fr, std::vector<uint8_t> &bits) {   // HasherEngine::hash_image, hashed_image.rs:24-64
    bits.assign(B200VFX_HASH_MAX_BITS, 0);
    int n = 0;
    if (b200vfx_hash_image(ctx, (int)values["hash-algo"].d, fr.format, fr.width, fr.height, fr.data[0], fr.stride[0], bits.data(), &n) != 0)
      return ctx_error(B200GST_FLOW_ERROR);
    bits.resize((size_t)n);
    return 0;
  }
  int aggregate(const b200gst_video_frame *frames, const int *pad_ids, int n, int64_t running_time, b200gst_video_frame *out) {
    // aggregate_frames imp.rs:259-389
    if (!started) return fail(B200GST_FLOW_ERROR, "not started");
    if (reference_pad < 0) return fail(B200GST_FLOW_EOS, "No reference sink pad exists");
    const b200gst_video_frame *ref = nullptr;
    for (int i = 0; i < n; i++) if (pad_ids[i] == reference_pad) ref = &frames[i];
    if (!ref) return B200GST_FLOW_OK;  // reference pad has not produced a buffer: nothing to compare (:283-296)
    if (!contains(pad_formats(0), ref->format)) return fail(B200GST_FLOW_NOT_NEGOTIATED, "format not negotiated");
    if (out && out->data[0]) {  // output = the reference buffer (:310-313)
      const size_t row = (size_t)ref->width * (ref->format == B200VFX_FORMAT_RGB ? 3 : 4);
      if (b200vfx_pointer_is_device(out->data[0]) || b200vfx_pointer_is_device(ref->data[0])) {  // device-resident pipeline
        if (b200vfx_copy_plane(ctx, out->data[0], out->stride[0], ref->data[0], ref->stride[0], row, ref->height) != 0 ||
            b200vfx_ctx_synchronize(ctx) != 0)
          return ctx_error(B200GST_FLOW_ERROR);
      } else {
        for (int y = 0; y < ref->height; y++)
          std::memcpy((uint8_t *)out->data[0] + (size_t)y * out->stride[0], (const uint8_t *)ref->data[0] + (size_t)y * ref->stride[0], row);
      }
    }
    // collect the other pads' frames first (the size check of :337-346 precedes any hashing of that pad)
    std::vector<std::pair<int, const b200gst_video_frame *>> others;
    for (int p : pads) {
      if (p == reference_pad) continue;
      const b200gst_video_frame *fr = nullptr;
      for (int i = 0; i < n; i++) if (pad_ids[i] == p) fr = &frames[i];
      if (!fr) return B200GST_FLOW_OK;  // :326-329
      if (fr->width != ref->width || fr->height != ref->height)
        return fail(B200GST_FLOW_NOT_NEGOTIATED, "Video streams do not have the same sizes (add videoscale and force the sizes to be equal on all sink pads)");
      others.push_back({p, fr});
    }
    std::vector<std::pair<int, double>> distances;
    bool same_fmt = true;
    for (auto &o : others) same_fmt = same_fmt && o.second->format == ref->format;
    if ((int)values["hash-algo"].d == 4 && same_fmt && !others.empty() && others.size() + 1 <= B200VFX_BLOCKHASH_MAX_FRAMES &&
        ref->width % 8 == 0 && ref->height % 8 == 0) {
      // one launch hashes the reference frame and every other pad's frame
      const int nf = (int)others.size() + 1;
      const void *srcs[B200VFX_BLOCKHASH_MAX_FRAMES];
      int strides[B200VFX_BLOCKHASH_MAX_FRAMES];
      srcs[0] = ref->data[0]; strides[0] = ref->stride[0];
      for (int i = 1; i < nf; i++) { srcs[i] = others[(size_t)i - 1].second->data[0]; strides[i] = others[(size_t)i - 1].second->stride[0]; }
      std::vector<uint32_t> sums((size_t)64 * nf);
      if (b200vfx_blockhash_sums_batch(ctx, ref->format, ref->width, ref->height, nf, srcs, strides, 8, 8, sums.data()) != 0)
        return ctx_error(B200GST_FLOW_ERROR);
      std::vector<uint8_t> ref_bits(64), bits(64);
      b200vfx_blockhash_bits(sums.data(), 8, 8, ref->width, ref->height, ref_bits.data());
      for (int i = 1; i < nf; i++) {
        b200vfx_blockhash_bits(sums.data() + (size_t)64 * i, 8, 8, ref->width, ref->height, bits.data());
        distances.push_back({others[(size_t)i - 1].first, (double)b200vfx_hash_distance(ref_bits.data(), bits.data(), 64)});
      }
    } else {
      std::vector<uint8_t> ref_bits, bits;
      if (int rc = hash(*ref, ref_bits)) return rc;
      for (auto &o : others) {
        if (int rc = hash(*o.second, bits)) return rc;
        distances.push_back({o.first, (double)b200vfx_hash_distance(ref_bits.data(), bits.data(), (int)std::min(ref_bits.size(), bits.size()))});
      }
    }
    const double thr = values["max-dist-threshold"].d;
    bool any = false;
    for (auto &d : distances) any = any || d.second <= thr;
    if (any) {  // post_message :361-377; structure layout mod.rs:110-123,148-158
      std::string m = "videocompare, pad-distances=(structure)< ";
      for (size_t i = 0; i < distances.size(); i++) {
        char b[128];
        std::snprintf(b, sizeof b, "\"pad-distance\\,\\ pad\\=sink_%d\\,\\ distance\\=(double)%g\\;\"%s", distances[i].first,
                      distances[i].second, i + 1 < distances.size() ? ", " : " ");
        m += b;
      }
      m += ">, running-time=(guint64)";
      m += running_time >= 0 ? std::to_string(running_time) : std::string("none");
      m += ";";
      bus.push_back(m);
    }
    return B200GST_FLOW_OK;
  }
};

// ---- colordetect ------------------------------------------------------------------------------
struct ColorDetect : b200gst_element {  // video/videofx/src/colordetect/imp.rs
  bool have_state = false;
  int color_format = -1;
  bool have_color = false;
  std::string current_color;  // State::current_color survives set_info (:277-284), cleared by stop (:239-243)
  std::vector<uint32_t> hist;
  ColorDetect() {
    factory = "colordetect"; type_name = "GstColorDetect"; plugin = "rsvideofx";
    add_prop({"quality", PType::UInt, 10, 0, 10, "mutable-playing", {}});       // :126-133
    add_prop({"max-colors", PType::UInt, 2, 2, 255, "mutable-playing", {}});    // :134-141
  }
  std::vector<int> pad_formats(int) const override {  // :214-222
    return {B200VFX_FORMAT_RGB, B200VFX_FORMAT_RGBA, B200VFX_FORMAT_ARGB, B200VFX_FORMAT_BGR, B200VFX_FORMAT_BGRA};
  }
  int set_caps(int in_format, int, int, int) override {  // set_info :253-287
    if (!contains(pad_formats(0), in_format)) return fail(-1, "unsupported format");
    color_format = in_format;
    have_state = true;
    passthrough = true;  // PASSTHROUGH_ON_SAME_CAPS + TRANSFORM_IP_ON_PASSTHROUGH (:236-238)
    return 0;
  }
  int stop() override { have_state = false; have_color = false; current_color.clear(); return b200gst_element::stop(); }
  int transform_frame_ip(b200gst_video_frame *fr) override {  // transform_frame_ip_passthrough :289-298 -> detect_color :57-86
    if (!have_state) return fail(B200GST_FLOW_NOT_NEGOTIATED, "Have no state yet");
    if (!started) return fail(B200GST_FLOW_ERROR, "not started");
    if (fr->format != color_format) return fail(B200GST_FLOW_NOT_NEGOTIATED, "format not negotiated");
    hist.resize(B200VFX_COLORDETECT_BINS);
    if (b200vfx_colordetect_histogram(ctx, fr->format, fr->width, fr->height, fr->data[0], fr->stride[0],
                                      (int)values["quality"].d, hist.data()) != 0)
      return ctx_error(B200GST_FLOW_ERROR);  // get_palette(..).map_err(|_| FlowError::Error)
    uint8_t pal[3 * 600];
    int n = 0;
    if (b200vfx_colordetect_palette(hist.data(), (int)values["max-colors"].d, pal, 600, &n) != 0 || n < 1)
      return fail(B200GST_FLOW_ERROR, "palette extraction failed");
    const std::string name = b200vfx_css_color_similar(pal[0], pal[1], pal[2]);
    if (!have_color || current_color != name) {  // color_changed :88-113
      have_color = true;
      current_color = name;
      std::string m = "colordetect, dominant-color=(string)" + name + ", palette=(uint){ ";
      for (int i = 0; i < n && i < 600; i++) {
        const unsigned v = ((unsigned)pal[3 * i] << 16) | ((unsigned)pal[3 * i + 1] << 8) | pal[3 * i + 2];
        m += std::to_string(v) + (i + 1 < n ? ", " : " ");
      }
      m += "};";
      bus.push_back(m);
    }
    return B200GST_FLOW_OK;
  }
};

b200gst_element *make(const std::string &n) {
  if (n == "colorlut") return new ColorLut();
  if (n == "hsvfilter") return new HsvFilter();
  if (n == "hsvdetector") return new HsvDetector();
  if (n == "roundedcorners") return new RoundedCorners();
  if (n == "videocompare") return new VideoCompare();
  if (n == "colordetect") return new ColorDetect();
  return nullptr;
}

const PropSpec *find_spec(const b200gst_element *el, const char *name) {
  for (const PropSpec &s : el->specs) if (s.name == name) return &s;
  return nullptr;
}

std::string fmt_num(double d) {
  char b[64];
  if (d == std::floor(d) && std::fabs(d) < 1e15) std::snprintf(b, sizeof b, "%.0f", d);
  else std::snprintf(b, sizeof b, "%.9g", d);
  return b;
}

int copy_out(const std::string &s, char *buf, size_t n) {
  if (!buf || n == 0) return -1;
  std::snprintf(buf, n, "%s", s.c_str());
  return (int)s.size();
}

}  // namespace

extern "C" {

b200gst_element *b200gst_element_factory_make(const char *factory_name) { return factory_name ? make(factory_name) : nullptr; }
void b200gst_element_unref(b200gst_element *el) { delete el; }
const char *b200gst_element_factory_name(const b200gst_element *el) { return el->factory.c_str(); }
const char *b200gst_element_type_name(const b200gst_element *el) { return el->type_name.c_str(); }
const char *b200gst_element_plugin_name(const b200gst_element *el) { return el->plugin.c_str(); }
const char *b200gst_element_last_error(const b200gst_element *el) { return el->err.c_str(); }

int b200gst_element_set_property(b200gst_element *el, const char *name, const char *value) {
  const PropSpec *s = find_spec(el, name);
  if (!s) return el->fail(-1, std::string("no property '") + name + "' in element " + el->factory);
  PropValue nv, old = el->values[name];
  if (s->type == PType::String) {
    nv.is_null = value == nullptr;
    if (value) nv.s = value;
  } else {
    if (!value) return el->fail(-1, "null value");
    nv.is_null = false;
    if (s->type == PType::Enum) {
      bool ok = false;
      for (auto &e : s->enum_nicks) if (e.second == value || std::to_string(e.first) == value) { nv.d = e.first; ok = true; }
      if (!ok) return el->fail(-1, std::string("invalid enum value '") + value + "' for property " + name);
    } else {
      char *end = nullptr;
      const double d = std::strtod(value, &end);
      if (end == value || *end != 0) return el->fail(-1, std::string("could not convert '") + value + "' for property " + name);
      double v = d;
      if (s->type == PType::Float) v = (double)(float)d;
      if (s->type == PType::UInt && (d != std::floor(d))) return el->fail(-1, "not an unsigned integer");
      if (!(v >= s->min && v <= s->max))  // g_object_set: out of range -> warning, value not set
        return el->fail(-1, std::string("value '") + value + "' is invalid or out of range for property '" + name + "'");
      nv.d = v;
    }
  }
  el->values[name] = nv;
  el->property_changed(name, old);
  return 0;
}

int b200gst_element_get_property(const b200gst_element *el, const char *name, char *buf, size_t n) {
  const PropSpec *s = find_spec(el, name);
  if (!s) return -1;
  const PropValue &v = el->values.at(name);
  if (s->type == PType::String) return copy_out(v.is_null ? "NULL" : v.s, buf, n);
  if (s->type == PType::Enum) {
    for (auto &e : s->enum_nicks) if (e.first == (int)v.d) return copy_out(e.second, buf, n);
    return -1;
  }
  return copy_out(fmt_num(v.d), buf, n);
}

int b200gst_element_list_properties(const b200gst_element *el, char *buf, size_t n) {
  static const char *tn[] = {"gchararray", "gfloat", "gdouble", "guint", "enum"};
  std::string out;
  for (const PropSpec &s : el->specs) {
    std::string def = s.type == PType::String ? "NULL" : fmt_num(s.def);
    if (s.type == PType::Enum) for (auto &e : s.enum_nicks) if (e.first == (int)s.def) def = e.second;
    out += s.name + "|" + tn[(int)s.type] + "|" + def + "|" + (s.type == PType::String ? "" : fmt_num(s.min)) + "|" +
           (s.type == PType::String ? "" : fmt_num(s.max)) + "|" + s.mutability + "\n";
  }
  return copy_out(out, buf, n);
}

static int copy_formats(const std::vector<int> &v, int *out, int cap) {
  for (int i = 0; i < (int)v.size() && i < cap; i++) out[i] = v[(size_t)i];
  return (int)v.size();
}
int b200gst_element_pad_template_formats(const b200gst_element *el, int direction, int *formats, int cap) {
  return copy_formats(el->pad_formats(direction), formats, cap);
}
int b200gst_element_transform_caps(b200gst_element *el, int direction, const int *formats, int n, int *out, int cap) {
  return copy_formats(el->transform_caps(direction, std::vector<int>(formats, formats + n)), out, cap);
}

int b200gst_element_start(b200gst_element *el) { return el->start(); }
int b200gst_element_stop(b200gst_element *el) { return el->stop(); }
int b200gst_element_set_caps(b200gst_element *el, int in_format, int out_format, int w, int h) { return el->set_caps(in_format, out_format, w, h); }
int b200gst_element_is_passthrough(const b200gst_element *el) { return el->passthrough ? 1 : 0; }
int b200gst_element_transform_frame(b200gst_element *el, const b200gst_video_frame *in, b200gst_video_frame *out) { return el->transform_frame(in, out); }
int b200gst_element_transform_frame_ip(b200gst_element *el, b200gst_video_frame *fr) { return el->transform_frame_ip(fr); }

int b200gst_roundedcorners_prepare_output(b200gst_element *el, const b200gst_video_frame *in, b200gst_video_frame *out) {
  RoundedCorners *rc = dynamic_cast<RoundedCorners *>(el);
  return rc ? rc->prepare_output(in, out) : el->fail(B200GST_FLOW_ERROR, "not a roundedcorners element");
}
int b200gst_videocompare_request_pad(b200gst_element *el) {
  VideoCompare *vc = dynamic_cast<VideoCompare *>(el);
  return vc ? vc->request_pad() : -1;
}
int b200gst_videocompare_release_pad(b200gst_element *el, int pad) {
  VideoCompare *vc = dynamic_cast<VideoCompare *>(el);
  return vc ? vc->release_pad(pad) : -1;
}
int b200gst_videocompare_reference_pad(const b200gst_element *el) {
  const VideoCompare *vc = dynamic_cast<const VideoCompare *>(el);
  return vc ? vc->reference_pad : -1;
}
int b200gst_videocompare_aggregate_frames(b200gst_element *el, const b200gst_video_frame *frames, const int *pad_ids, int n,
                                          int64_t running_time_ns, b200gst_video_frame *out) {
  VideoCompare *vc = dynamic_cast<VideoCompare *>(el);
  return vc ? vc->aggregate(frames, pad_ids, n, running_time_ns, out) : el->fail(B200GST_FLOW_ERROR, "not a videocompare element");
}
int b200gst_element_pop_message(b200gst_element *el, char *buf, size_t n) {
  if (el->bus.empty()) return 0;
  copy_out(el->bus.front(), buf, n);
  el->bus.pop_front();
  return 1;
}

}  // extern "C"
