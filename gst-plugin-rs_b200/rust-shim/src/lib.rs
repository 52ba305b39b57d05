// Plugin shell: same element factory names as the reference plugins (`colorlut`, `hsvfilter`, `hsvdetector`,
// `roundedcorners`, `videocompare`), so existing gst-launch pipelines keep working.  Only `colorlut` is spelled
// out in full (colorlut.rs); hsv.rs / videofx.rs hold the vfunc bodies that replace the reference's pixel loops.
pub mod allocator;
pub mod colorlut;
pub mod ffi;
pub mod hsv;
pub mod videofx;

// Registration, as in video/colorlut/src/lib.rs:19-43 of the reference: the element keeps its factory name and rank, and
// the plugin keeps its name, so `gst_plugin_colorlut_register` / `gst_plugin_colorlut_get_desc` stay the entry symbols a
// static build links against (ci/generate-static-test.py:36-41).  The hsv and rsvideofx plugins register their elements
// the same way (hsvfilter, hsvdetector / roundedcorners, videocompare, colordetect) once their element shells call the
// vfunc bodies in hsv.rs / videofx.rs.  Written, never compiled in this image.
use gst::glib;
use gst::prelude::*;

glib::wrapper! {
    pub struct ColorLut(ObjectSubclass<colorlut::ColorLut>) @extends gst_video::VideoFilter, gst_base::BaseTransform, gst::Element, gst::Object;
}

fn plugin_init(plugin: &gst::Plugin) -> Result<(), glib::BoolError> {
    gst::Element::register(Some(plugin), "colorlut", gst::Rank::NONE, ColorLut::static_type())
}

gst::plugin_define!(
    colorlut,
    "Color LUT (B200 / sm_100a through libb200vfx)",
    plugin_init,
    "0.1.0",
    "MPL-2.0",
    "gst-plugin-b200vfx-shim",
    "gst-plugin-b200vfx-shim",
    "https://gitlab.freedesktop.org/gstreamer/gst-plugins-rs",
    "2026-01-01"
);
