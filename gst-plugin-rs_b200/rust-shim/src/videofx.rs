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

/// replaces HasherEngine::hash_image for HashAlg::Blockhash (videocompare/hashed_image.rs:24-64): no
/// `tightly_packed_framebuffer` copy (:110-130) -- the stride is passed through.
pub struct BlockHash(pub [u8; 64]);

pub fn hash_image(
    ctx: &ffi::Ctx,
    frame: &gst_video::VideoFrameRef<&gst::BufferRef>,
) -> Result<BlockHash, gst::FlowError> {
    use gst_video::prelude::*;
    let fmt = ffi::format_code(frame.format()).ok_or(gst::FlowError::NotNegotiated)?;
    let (w, h) = (frame.width() as i32, frame.height() as i32);
    let mut sums = [0u32; 64];
    let rc = unsafe {
        ffi::b200vfx_blockhash_sums(
            ctx.0,
            fmt,
            w,
            h,
            frame.plane_data(0).unwrap().as_ptr() as *const _,
            frame.plane_stride()[0],
            8,
            8,
            sums.as_mut_ptr(),
        )
    };
    if rc != ffi::B200VFX_OK {
        return Err(gst::FlowError::Error);
    }
    let mut bits = [0u8; 64];
    unsafe { ffi::b200vfx_blockhash_bits(sums.as_ptr(), 8, 8, w, h, bits.as_mut_ptr()) };
    Ok(BlockHash(bits))
}

/// replaces HasherEngine::compare (hashed_image.rs:66-79): Hamming distance as f64
pub fn compare(a: &BlockHash, b: &BlockHash) -> f64 {
    unsafe { ffi::b200vfx_hash_distance(a.0.as_ptr(), b.0.as_ptr(), 64) as f64 }
}
