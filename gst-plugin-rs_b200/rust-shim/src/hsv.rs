// hsv.rs -- the two vfunc bodies of gst-plugins-rs video/hsv that change when libb200vfx is used.
// Everything else of hsvfilter/imp.rs and hsvdetector/imp.rs (Settings, properties, pad templates,
// transform_caps, BaseTransformMode) stays exactly as it is; the elements only gain a `ctx: Mutex<Option<ffi::Ctx>>`
// created in `start()` and dropped in `stop()`.  Written, never compiled in this image.
use gst_video::prelude::*;

use crate::ffi;

/// replaces HsvFilter::transform_frame_ip (hsvfilter/imp.rs:323-376) and hsv_filter (:76-120):
/// the format -> closure dispatch collapses into the format code, the settings snapshot is passed by value.
pub fn hsvfilter_transform_frame_ip(
    ctx: &ffi::Ctx,
    settings: &HsvFilterSettings, // the `*self.settings.lock().unwrap()` snapshot of imp.rs:85
    frame: &mut gst_video::VideoFrameRef<&mut gst::BufferRef>,
) -> Result<gst::FlowSuccess, gst::FlowError> {
    let fmt = ffi::format_code(frame.format()).ok_or(gst::FlowError::NotNegotiated)?;
    let (w, h, stride) = (frame.width() as i32, frame.height() as i32, frame.plane_stride()[0]);
    let data = frame.plane_data_mut(0).unwrap().as_mut_ptr();
    let rc = unsafe {
        ffi::b200vfx_hsvfilter_process(
            ctx.0,
            fmt,
            w,
            h,
            data as *mut _,
            stride,
            settings.hue_shift,
            settings.saturation_mul,
            settings.saturation_off,
            settings.value_mul,
            settings.value_off,
        )
    };
    if rc != ffi::B200VFX_OK {
        return Err(gst::FlowError::Error);
    }
    Ok(gst::FlowSuccess::Ok)
}

/// replaces HsvDetector::transform_frame (hsvdetector/imp.rs:423-707) and hsv_detect (:100-160):
/// the 16 closure pairs become (in_fmt, out_fmt).
pub fn hsvdetector_transform_frame(
    ctx: &ffi::Ctx,
    settings: &HsvDetectorSettings,
    in_frame: &gst_video::VideoFrameRef<&gst::BufferRef>,
    out_frame: &mut gst_video::VideoFrameRef<&mut gst::BufferRef>,
) -> Result<gst::FlowSuccess, gst::FlowError> {
    let in_fmt = ffi::format_code(in_frame.format()).ok_or(gst::FlowError::NotNegotiated)?;
    let out_fmt = ffi::format_code(out_frame.format()).ok_or(gst::FlowError::NotNegotiated)?;
    let rc = unsafe {
        ffi::b200vfx_hsvdetector_process(
            ctx.0,
            in_fmt,
            out_fmt,
            in_frame.width() as i32,
            in_frame.height() as i32,
            in_frame.plane_data(0).unwrap().as_ptr() as *const _,
            in_frame.plane_stride()[0],
            out_frame.plane_data_mut(0).unwrap().as_mut_ptr() as *mut _,
            out_frame.plane_stride()[0],
            settings.hue_ref,
            settings.hue_var,
            settings.saturation_ref,
            settings.saturation_var,
            settings.value_ref,
            settings.value_var,
        )
    };
    if rc != ffi::B200VFX_OK {
        return Err(gst::FlowError::Error);
    }
    Ok(gst::FlowSuccess::Ok)
}

// the reference's own Settings structs (hsvfilter/imp.rs:31-38, hsvdetector/imp.rs:34-42), unchanged
#[derive(Debug, Clone, Copy)]
pub struct HsvFilterSettings {
    pub hue_shift: f32,
    pub saturation_mul: f32,
    pub saturation_off: f32,
    pub value_mul: f32,
    pub value_off: f32,
}

#[derive(Debug, Clone, Copy)]
pub struct HsvDetectorSettings {
    pub hue_ref: f32,
    pub hue_var: f32,
    pub saturation_ref: f32,
    pub saturation_var: f32,
    pub value_ref: f32,
    pub value_var: f32,
}
