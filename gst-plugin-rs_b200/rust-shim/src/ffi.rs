// ffi.rs -- the `extern "C"` binding of include/b200vfx.h that the gstreamer-rs elements use.
//
// NOTE: this image has no cargo/rustc/GStreamer, so this crate is written but was never compiled
// here.  The C ABI it binds is exercised by tests/ through ctypes and by the C++ element
// emulator (csrc/elements.cpp), which mirrors these elements one-to-one.
#![allow(non_camel_case_types, dead_code)]

use std::ffi::{c_char, c_float, c_int, c_uint, c_void, CStr};

#[repr(C)]
pub struct b200vfx_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct b200vfx_fence {
    _private: [u8; 0],
}

pub const B200VFX_OK: c_int = 0;
pub const B200VFX_ERR_NOT_NEGOTIATED: c_int = -3;

// GstVideoFormat -> b200vfx_format
pub fn format_code(f: gst_video::VideoFormat) -> Option<c_int> {
    use gst_video::VideoFormat::*;
    Some(match f {
        Rgbx => 0,
        Xrgb => 1,
        Bgrx => 2,
        Xbgr => 3,
        Rgba => 4,
        Argb => 5,
        Bgra => 6,
        Abgr => 7,
        Rgb => 8,
        Bgr => 9,
        Rgba64Le => 10,
        Rgba64Be => 11,
        I420 => 12,
        A420 => 13,
        _ => return None,
    })
}

#[link(name = "b200vfx")]
extern "C" {
    pub fn b200vfx_ctx_create(out: *mut *mut b200vfx_ctx, device: c_int) -> c_int;
    pub fn b200vfx_ctx_destroy(ctx: *mut b200vfx_ctx);
    pub fn b200vfx_last_error(ctx: *const b200vfx_ctx) -> *const c_char;
    pub fn b200vfx_host_alloc(bytes: usize) -> *mut c_void;
    pub fn b200vfx_host_free(p: *mut c_void);
    // device-resident frames: what a memory:CUDAMemory-style allocator of these elements needs
    pub fn b200vfx_device_alloc(ctx: *mut b200vfx_ctx, bytes: usize) -> *mut c_void;
    pub fn b200vfx_device_free(ctx: *mut b200vfx_ctx, p: *mut c_void);
    pub fn b200vfx_upload(ctx: *mut b200vfx_ctx, dev_dst: *mut c_void, dst_stride: c_int, host_src: *const c_void, src_stride: c_int, row_bytes: usize, rows: c_int) -> c_int;
    pub fn b200vfx_download(ctx: *mut b200vfx_ctx, host_dst: *mut c_void, dst_stride: c_int, dev_src: *const c_void, src_stride: c_int, row_bytes: usize, rows: c_int) -> c_int;
    pub fn b200vfx_ctx_synchronize(ctx: *mut b200vfx_ctx) -> c_int;
    pub fn b200vfx_ctx_set_host_async(ctx: *mut b200vfx_ctx, enable: c_int) -> c_int;
    pub fn b200vfx_fence_create(ctx: *mut b200vfx_ctx, fence_out: *mut *mut b200vfx_fence) -> c_int;
    pub fn b200vfx_fence_wait(fence: *mut b200vfx_fence) -> c_int;
    pub fn b200vfx_fence_query(fence: *mut b200vfx_fence) -> c_int;
    pub fn b200vfx_fence_destroy(fence: *mut b200vfx_fence);

    pub fn b200vfx_colorlut_load_file(ctx: *mut b200vfx_ctx, location: *const c_char) -> c_int;
    pub fn b200vfx_colorlut_clear(ctx: *mut b200vfx_ctx) -> c_int;
    pub fn b200vfx_colorlut_process(
        ctx: *mut b200vfx_ctx,
        fmt: c_int,
        width: c_int,
        height: c_int,
        src: *const c_void,
        src_stride: c_int,
        dst: *mut c_void,
        dst_stride: c_int,
    ) -> c_int;

    pub fn b200vfx_hsvfilter_process(
        ctx: *mut b200vfx_ctx,
        fmt: c_int,
        width: c_int,
        height: c_int,
        data: *mut c_void,
        stride: c_int,
        hue_shift: c_float,
        saturation_mul: c_float,
        saturation_off: c_float,
        value_mul: c_float,
        value_off: c_float,
    ) -> c_int;

    pub fn b200vfx_hsvdetector_process(
        ctx: *mut b200vfx_ctx,
        in_fmt: c_int,
        out_fmt: c_int,
        width: c_int,
        height: c_int,
        src: *const c_void,
        src_stride: c_int,
        dst: *mut c_void,
        dst_stride: c_int,
        hue_ref: c_float,
        hue_var: c_float,
        saturation_ref: c_float,
        saturation_var: c_float,
        value_ref: c_float,
        value_var: c_float,
    ) -> c_int;

    pub fn b200vfx_roundmask_generate(
        ctx: *mut b200vfx_ctx,
        width: c_int,
        height: c_int,
        stride: c_int,
        border_radius_px: c_uint,
        a8_out: *mut c_void,
    ) -> c_int;

    pub fn b200vfx_blockhash_sums(
        ctx: *mut b200vfx_ctx,
        fmt: c_int,
        width: c_int,
        height: c_int,
        src: *const c_void,
        stride: c_int,
        hw: c_int,
        hh: c_int,
        sums: *mut u32,
    ) -> c_int;
    // videocompare: reference frame + every other pad's frame in one launch (aggregate_frames, imp.rs:297-353)
    pub fn b200vfx_blockhash_sums_batch(
        ctx: *mut b200vfx_ctx,
        fmt: c_int,
        width: c_int,
        height: c_int,
        n_frames: c_int,
        srcs: *const *const c_void,
        strides: *const c_int,
        hw: c_int,
        hh: c_int,
        sums: *mut u32,
    ) -> c_int;
    pub fn b200vfx_blockhash_bits(sums: *const u32, hw: c_int, hh: c_int, width: c_int, height: c_int, bits_out: *mut u8);
    pub fn b200vfx_hash_distance(a: *const u8, b: *const u8, nbits: c_int) -> c_int;
    // HasherEngine::hash_image for every HashAlgorithm value (videocompare/mod.rs:57-92: mean 0, gradient 1, vertgradient 2,
    // doublegradient 3, blockhash 4) and any frame size; bits_out: 64 bytes of 0/1, *n_bits = 64 (40 for doublegradient)
    pub fn b200vfx_hash_image(
        ctx: *mut b200vfx_ctx,
        algo: c_int,
        fmt: c_int,
        width: c_int,
        height: c_int,
        src: *const c_void,
        stride: c_int,
        bits_out: *mut u8,
        n_bits: *mut c_int,
    ) -> c_int;

    // the videoconvert either side of the elements, on the device (include/b200vfx.h "format conversion")
    pub fn b200vfx_convert_packed(ctx: *mut b200vfx_ctx, src_fmt: c_int, dst_fmt: c_int, width: c_int, height: c_int,
                                  src: *const c_void, src_stride: c_int, dst: *mut c_void, dst_stride: c_int) -> c_int;
    pub fn b200vfx_colorlut_process_fmt(ctx: *mut b200vfx_ctx, in_fmt: c_int, out_fmt: c_int, width: c_int, height: c_int,
                                        src: *const c_void, src_stride: c_int, dst: *mut c_void, dst_stride: c_int) -> c_int;
    pub fn b200vfx_convert_to_planar(ctx: *mut b200vfx_ctx, src_fmt: c_int, dst_fmt: c_int, width: c_int, height: c_int,
                                     src: *const c_void, src_stride: c_int, planes: *const *mut c_void, strides: *const c_int,
                                     matrix: c_int) -> c_int;
    pub fn b200vfx_convert_from_planar(ctx: *mut b200vfx_ctx, src_fmt: c_int, dst_fmt: c_int, width: c_int, height: c_int,
                                       planes: *const *const c_void, strides: *const c_int, dst: *mut c_void, dst_stride: c_int,
                                       matrix: c_int) -> c_int;
    /// colorlut on I420 / A420 planes, both videoconverts inside the kernel
    pub fn b200vfx_colorlut_process_planar(ctx: *mut b200vfx_ctx, fmt: c_int, width: c_int, height: c_int, src_planes: *const *const c_void,
        src_strides: *const c_int, dst_planes: *const *mut c_void, dst_strides: *const c_int, matrix: c_int) -> c_int;
    pub fn b200vfx_a420_append(ctx: *mut b200vfx_ctx, width: c_int, height: c_int, i420_planes: *const *const c_void,
                               i420_strides: *const c_int, a8: *const c_void, a8_stride: c_int,
                               out_planes: *const *mut c_void, out_strides: *const c_int) -> c_int;

    // colordetect (video/videofx/src/colordetect/imp.rs:57-86): get_palette's pixel pass on the GPU,
    // median cut + CSS name on the host
    pub fn b200vfx_colordetect_histogram(
        ctx: *mut b200vfx_ctx,
        fmt: c_int,
        width: c_int,
        height: c_int,
        src: *const c_void,
        stride: c_int,
        quality: c_int,
        hist: *mut u32, // 32768 bins
    ) -> c_int;
    pub fn b200vfx_colordetect_palette(
        hist: *const u32,
        max_colors: c_int,
        palette_rgb: *mut u8,
        palette_cap: c_int,
        n_colors: *mut c_int,
    ) -> c_int;
    pub fn b200vfx_css_color_similar(r: c_uint, g: c_uint, b: c_uint) -> *const c_char;
}

/// Owning wrapper: one context per element instance, created in `start()`, dropped in `stop()`.
pub struct Ctx(pub *mut b200vfx_ctx);
unsafe impl Send for Ctx {}

impl Ctx {
    pub fn new() -> Result<Self, String> {
        let mut p = std::ptr::null_mut();
        let rc = unsafe { b200vfx_ctx_create(&mut p, -1) };
        if rc != B200VFX_OK {
            return Err(last_error(std::ptr::null()));
        }
        Ok(Ctx(p))
    }
    pub fn error(&self) -> String {
        last_error(self.0)
    }
}

impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { b200vfx_ctx_destroy(self.0) }
    }
}

pub fn last_error(ctx: *const b200vfx_ctx) -> String {
    unsafe {
        let p = b200vfx_last_error(ctx);
        if p.is_null() {
            String::new()
        } else {
            CStr::from_ptr(p).to_string_lossy().into_owned()
        }
    }
}
