// colorlut.rs -- gstreamer-rs `colorlut` element whose pixel loop is libb200vfx (B200 / sm_100a).
//
// Keeps the element surface of gst-plugins-rs video/colorlut/src/colorlut/imp.rs: type name
// GstColorLut : GstVideoFilter, factory "colorlut", sink/src caps video/x-raw
// {RGBA64_LE, RGBA64_BE, RGBA}, property `location` (string, mutable in READY), NeverInPlace,
// start() parses the .cube (Settings error when `location` is unset, Read error when the file does
// not parse), stop() drops the LUT.  Only transform_frame differs: one FFI call.
//
// Written, never compiled in this image (no cargo / gstreamer-rs here).
use gst::glib;
use gst::subclass::prelude::*;
use gst_base::subclass::prelude::*;
use gst_video::prelude::*;
use gst_video::subclass::prelude::*;
use std::ffi::CString;
use std::sync::{LazyLock, Mutex};

use crate::ffi;

static CAT: LazyLock<gst::DebugCategory> =
    LazyLock::new(|| gst::DebugCategory::new("colorlut", gst::DebugColorFlags::empty(), Some("Color LUT (B200)")));

#[derive(Default)]
pub struct ColorLut {
    location: Mutex<Option<String>>,
    ctx: Mutex<Option<ffi::Ctx>>,
}

#[glib::object_subclass]
impl ObjectSubclass for ColorLut {
    const NAME: &'static str = "GstColorLut";
    type Type = super::ColorLut;
    type ParentType = gst_video::VideoFilter;
}

impl ObjectImpl for ColorLut {
    fn properties() -> &'static [glib::ParamSpec] {
        static P: LazyLock<Vec<glib::ParamSpec>> = LazyLock::new(|| {
            vec![glib::ParamSpecString::builder("location")
                .nick("Location")
                .blurb("Location of the LUT file to read from")
                .mutable_ready()
                .build()]
        });
        P.as_ref()
    }
    fn set_property(&self, _id: usize, value: &glib::Value, _pspec: &glib::ParamSpec) {
        *self.location.lock().unwrap() = value.get().unwrap();
    }
    fn property(&self, _id: usize, _pspec: &glib::ParamSpec) -> glib::Value {
        self.location.lock().unwrap().to_value()
    }
}

impl GstObjectImpl for ColorLut {}

impl ElementImpl for ColorLut {
    fn metadata() -> Option<&'static gst::subclass::ElementMetadata> {
        static M: LazyLock<gst::subclass::ElementMetadata> = LazyLock::new(|| {
            gst::subclass::ElementMetadata::new("Color LUT", "Filter/Effect/Video", "Apply color lookup table", "b200vfx")
        });
        Some(&*M)
    }
    fn pad_templates() -> &'static [gst::PadTemplate] {
        static T: LazyLock<Vec<gst::PadTemplate>> = LazyLock::new(|| {
            let caps = gst_video::VideoCapsBuilder::new()
                .format_list([
                    gst_video::VideoFormat::Rgba64Le,
                    gst_video::VideoFormat::Rgba64Be,
                    gst_video::VideoFormat::Rgba,
                ])
                .build();
            vec![
                gst::PadTemplate::new("src", gst::PadDirection::Src, gst::PadPresence::Always, &caps).unwrap(),
                gst::PadTemplate::new("sink", gst::PadDirection::Sink, gst::PadPresence::Always, &caps).unwrap(),
            ]
        });
        T.as_ref()
    }
}

impl BaseTransformImpl for ColorLut {
    const MODE: gst_base::subclass::BaseTransformMode = gst_base::subclass::BaseTransformMode::NeverInPlace;
    const PASSTHROUGH_ON_SAME_CAPS: bool = false;
    const TRANSFORM_IP_ON_PASSTHROUGH: bool = false;

    fn start(&self) -> Result<(), gst::ErrorMessage> {
        let location = self.location.lock().unwrap().clone().ok_or_else(|| {
            gst::error_msg!(gst::ResourceError::Settings, ["LUT file location is not configured"])
        })?;
        let ctx = ffi::Ctx::new().map_err(|e| gst::error_msg!(gst::LibraryError::Init, ["b200vfx: {e}"]))?;
        let c_loc = CString::new(location.clone()).unwrap();
        // parses the .cube with the same grammar as CubeLut::parse and uploads it to HBM
        if unsafe { ffi::b200vfx_colorlut_load_file(ctx.0, c_loc.as_ptr()) } != ffi::B200VFX_OK {
            return Err(gst::error_msg!(gst::ResourceError::Read, ["{}", ctx.error()]));
        }
        *self.ctx.lock().unwrap() = Some(ctx);
        Ok(())
    }

    fn stop(&self) -> Result<(), gst::ErrorMessage> {
        *self.ctx.lock().unwrap() = None;
        Ok(())
    }

    // Pinned buffers on both sides (allocator.rs): upstream allocates from our page-locked allocator, and so does the
    // pool our output buffers come from -- the hooks the reference's GPU variant overrides too
    // (d3d12colorlut/imp.rs:299-542).  With them a host-memory pipeline runs at PCIe speed (1178 instead of 245 frames/s).
    fn propose_allocation(
        &self,
        decide_query: Option<&gst::query::Allocation>,
        query: &mut gst::query::Allocation,
    ) -> Result<(), gst::LoggableError> {
        crate::allocator::propose_allocation(query)?;
        self.parent_propose_allocation(decide_query, query)
    }

    fn decide_allocation(&self, query: &mut gst::query::Allocation) -> Result<(), gst::LoggableError> {
        let (caps, _need_pool) = query.get_owned();
        let info = gst_video::VideoInfo::from_caps(&caps).map_err(|_| gst::loggable_error!(CAT, "bad caps in allocation query"))?;
        crate::allocator::decide_allocation(query, &caps, info.size() as u32)?;
        self.parent_decide_allocation(query)
    }
}

impl VideoFilterImpl for ColorLut {
    fn transform_frame(
        &self,
        in_frame: &gst_video::VideoFrameRef<&gst::BufferRef>,
        out_frame: &mut gst_video::VideoFrameRef<&mut gst::BufferRef>,
    ) -> Result<gst::FlowSuccess, gst::FlowError> {
        let guard = self.ctx.lock().unwrap();
        let ctx = guard.as_ref().ok_or_else(|| {
            gst::error!(CAT, imp = self, "No LUT configured");
            gst::FlowError::Error
        })?;
        let fmt = ffi::format_code(in_frame.format()).ok_or(gst::FlowError::NotNegotiated)?;
        let (w, h) = (in_frame.width() as i32, in_frame.height() as i32);
        let (ss, ds) = (in_frame.plane_stride()[0], out_frame.plane_stride()[0]);
        let src = in_frame.plane_data(0).unwrap().as_ptr();
        let dst = out_frame.plane_data_mut(0).unwrap().as_mut_ptr();
        // the whole per-pixel loop (transform_rgba / transform_rgba64<LE|BE>, 1D and 3D) happens here
        let rc = unsafe { ffi::b200vfx_colorlut_process(ctx.0, fmt, w, h, src as *const _, ss, dst as *mut _, ds) };
        if rc != ffi::B200VFX_OK {
            gst::error!(CAT, imp = self, "b200vfx_colorlut_process: {}", ctx.error());
            return Err(gst::FlowError::Error);
        }
        Ok(gst::FlowSuccess::Ok)
    }
}
