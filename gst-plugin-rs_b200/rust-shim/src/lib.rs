// Plugin shell: same element factory names as the reference plugins (`colorlut`, `hsvfilter`, `hsvdetector`,
// `roundedcorners`, `videocompare`), so existing gst-launch pipelines keep working.  Only `colorlut` is spelled
// out in full (colorlut.rs); hsv.rs / videofx.rs hold the vfunc bodies that replace the reference's pixel loops.
pub mod colorlut;
pub mod ffi;
pub mod hsv;
pub mod videofx;
