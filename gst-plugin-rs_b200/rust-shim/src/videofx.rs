// videofx.rs -- call sites inside gst-plugins-rs video/videofx that move to libb200vfx.
// Written, never compiled in this image (no cargo / gstreamer-rs / cairo here).
use crate::ffi;

/// replaces RoundedCorners::generate_alpha_mask + draw_rounded_corners (border/imp.rs:57-180): fills the shared
/// A8 `alpha_mem` (stride = out_info.stride()[3], rows = round_up_2(height)) -- no cairo surface/context needed.
pub fn generate_alpha_mask(
    ctx: &ffi::Ctx,
    alpha_mem: &mut [u8],
    width: i32,
    height: i32,
    alpha_stride: i32,
    border_radius_px: u32,
) -> Result<(), String> {
    assert!(alpha_mem.len() >= (alpha_stride as usize) * (((height + 1) & !1) as usize));
    let rc = unsafe {
        ffi::b200vfx_roundmask_generate(ctx.0, width, height, alpha_stride, border_radius_px, alpha_mem.as_mut_ptr() as *mut _)
    };
    if rc != ffi::B200VFX_OK {
        return Err(ctx.error());
    }
    Ok(())
}

/// replaces HasherEngine::hash_image (videocompare/hashed_image.rs:24-64) for every `hash-algo` value
/// (HashAlgorithm, videocompare/mod.rs:57-92: Mean 0, Gradient 1, VertGradient 2, DoubleGradient 3, Blockhash 4) and any
/// frame size: no `tightly_packed_framebuffer` copy (:110-130) -- the stride is passed through.
pub struct BlockHash(pub [u8; 64], pub usize);

pub fn hash_image(
    ctx: &ffi::Ctx,
    algo: i32,
    frame: &gst_video::VideoFrameRef<&gst::BufferRef>,
) -> Result<BlockHash, gst::FlowError> {
    use gst_video::prelude::*;
    let fmt = ffi::format_code(frame.format()).ok_or(gst::FlowError::NotNegotiated)?;
    let mut bits = [0u8; 64];
    let mut n_bits = 0;
    let rc = unsafe {
        ffi::b200vfx_hash_image(
            ctx.0,
            algo,
            fmt,
            frame.width() as i32,
            frame.height() as i32,
            frame.plane_data(0).unwrap().as_ptr() as *const _,
            frame.plane_stride()[0],
            bits.as_mut_ptr(),
            &mut n_bits,
        )
    };
    if rc != ffi::B200VFX_OK {
        return Err(gst::FlowError::Error);
    }
    Ok(BlockHash(bits, n_bits as usize))
}

/// replaces HasherEngine::compare (hashed_image.rs:66-79): Hamming distance as f64
pub fn compare(a: &BlockHash, b: &BlockHash) -> f64 {
    unsafe { ffi::b200vfx_hash_distance(a.0.as_ptr(), b.0.as_ptr(), a.1.min(b.1) as i32) as f64 }
}

/// aggregate_frames (videocompare/imp.rs:297-353) hashes the reference frame and then every other pad's frame:
/// all of them in ONE launch.  `frames[0]` is the reference pad's frame; returns the distance of every other frame.
pub fn hash_and_compare_all(
    ctx: &ffi::Ctx,
    frames: &[&gst_video::VideoFrameRef<&gst::BufferRef>],
) -> Result<Vec<f64>, gst::FlowError> {
    use gst_video::prelude::*;
    let n = frames.len();
    assert!(n >= 1 && n <= 8);
    let fmt = ffi::format_code(frames[0].format()).ok_or(gst::FlowError::NotNegotiated)?;
    let (w, h) = (frames[0].width() as i32, frames[0].height() as i32);
    let srcs: Vec<*const std::ffi::c_void> = frames.iter().map(|f| f.plane_data(0).unwrap().as_ptr() as *const _).collect();
    let strides: Vec<i32> = frames.iter().map(|f| f.plane_stride()[0]).collect();
    let mut sums = vec![0u32; 64 * n];
    let rc = unsafe {
        ffi::b200vfx_blockhash_sums_batch(ctx.0, fmt, w, h, n as i32, srcs.as_ptr(), strides.as_ptr(), 8, 8, sums.as_mut_ptr())
    };
    if rc != ffi::B200VFX_OK {
        return Err(gst::FlowError::Error);
    }
    let mut hashes = Vec::with_capacity(n);
    for i in 0..n {
        let mut bits = [0u8; 64];
        unsafe { ffi::b200vfx_blockhash_bits(sums[64 * i..].as_ptr(), 8, 8, w, h, bits.as_mut_ptr()) };
        hashes.push(BlockHash(bits, 64));
    }
    Ok(hashes[1..].iter().map(|x| compare(&hashes[0], x)).collect())
}

/// replaces the body of ColorDetect::detect_color (colordetect/imp.rs:57-86): color_thief::get_palette's pixel pass
/// (5-bit histogram of every `quality`-th pixel) runs on the GPU, the median cut and the CSS name on the host.
/// Returns (dominant colour name, palette as 0xRRGGBB) -- the caller keeps the `current_color` comparison (:80-85).
pub fn detect_color(
    ctx: &ffi::Ctx,
    frame: &gst_video::VideoFrameRef<&gst::BufferRef>,
    quality: u32,
    max_colors: u32,
) -> Result<(String, Vec<u32>), gst::FlowError> {
    use gst_video::prelude::*;
    let fmt = ffi::format_code(frame.format()).ok_or(gst::FlowError::NotNegotiated)?;
    let mut hist = vec![0u32; 32768];
    let rc = unsafe {
        ffi::b200vfx_colordetect_histogram(
            ctx.0,
            fmt,
            frame.width() as i32,
            frame.height() as i32,
            frame.plane_data(0).unwrap().as_ptr() as *const _,
            frame.plane_stride()[0],
            quality as i32,
            hist.as_mut_ptr(),
        )
    };
    if rc != ffi::B200VFX_OK {
        return Err(gst::FlowError::Error); // get_palette(..).map_err(|_| FlowError::Error), imp.rs:74
    }
    let mut pal = [0u8; 3 * 600];
    let mut n = 0i32;
    let rc = unsafe { ffi::b200vfx_colordetect_palette(hist.as_ptr(), max_colors as i32, pal.as_mut_ptr(), 600, &mut n) };
    if rc != ffi::B200VFX_OK || n < 1 {
        return Err(gst::FlowError::Error);
    }
    let name = unsafe { std::ffi::CStr::from_ptr(ffi::b200vfx_css_color_similar(pal[0] as u32, pal[1] as u32, pal[2] as u32)) }
        .to_string_lossy()
        .into_owned();
    let palette = (0..n.min(600) as usize)
        .map(|i| ((pal[3 * i] as u32) << 16) | ((pal[3 * i + 1] as u32) << 8) | pal[3 * i + 2] as u32)
        .collect();
    Ok((name, palette))
}
