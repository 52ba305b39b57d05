// allocator.rs -- the two GstAllocators the b200vfx elements hand out, and the propose/decide_allocation hooks that
// install them.  Mirrors what the reference's own GPU element does for its device memory
// (video/colorlut/src/d3d12colorlut/imp.rs:299-542: decide_allocation / propose_allocation with a D3D12 pool and the
// `memory:D3D12Memory` caps feature :236-246) and what roundedcorners overrides for its video meta
// (video/videofx/src/border/imp.rs:565-572).
//
//   * PinnedAllocator  -- page-locked host memory (b200vfx_host_alloc).  Offered upstream in propose_allocation and chosen
//     for our own output buffers in decide_allocation: frames then cross PCIe by full-speed DMA, or not at all -- the
//     zero-copy kernels read and write pinned frames in place (1178 frames/s end to end on 4K RGBA instead of 245 with
//     pageable buffers, BENCH_r01.json / profiles/README.md).
//   * DeviceAllocator  -- HBM (b200vfx_device_alloc), negotiated with the caps feature `memory:B200Memory` between two
//     adjacent b200vfx elements: the frame never leaves the GPU (hsvfilter -> hsvdetector 2933 vs 1266 frames/s).
//     Mapping such a memory for a CPU element downloads it (like GstCudaMemory does).
//
// Written, never compiled in this image (no cargo / gstreamer-rs here); the C entry points it calls are exercised by
// tests/test_gpu_variants.py::test_device_resident_chain_one_upload_one_download and tests/test_gpu_convert.py.
use gst::glib;
use gst::prelude::*;
use gst::subclass::prelude::*;
use std::sync::LazyLock;

use crate::ffi;

pub const CAPS_FEATURE_MEMORY_B200: &str = "memory:B200Memory";

// ---- pinned host memory ---------------------------------------------------------------------------------------------
// A pinned block can carry a FENCE: in asynchronous host-frame mode (b200vfx_ctx_set_host_async) transform_frame returns
// as soon as the frame's copies and kernel are enqueued, the output buffer is pushed downstream at once, and whoever maps
// it for the CPU waits here -- what the reference's d3d12colorlut does with D3D12Memory::set_fence
// (video/colorlut/src/d3d12colorlut/imp.rs:711-714) and GstCudaMemory does on map(READ).  The upload of frame i+1 then overlaps the download
// of frame i: 1380 instead of 1176 frames/s end to end on 4K RGBA (profiles/r02_e2e_async.jsonl).
struct PinnedBlock {
    ptr: *mut u8,
    len: usize,
    fence: std::sync::Mutex<Option<FencePtr>>,
}
struct FencePtr(*mut ffi::b200vfx_fence);
unsafe impl Send for FencePtr {}
unsafe impl Send for PinnedBlock {}
impl PinnedBlock {
    fn wait(&self) {
        if let Some(f) = self.fence.lock().unwrap().take() {
            unsafe {
                ffi::b200vfx_fence_wait(f.0);
                ffi::b200vfx_fence_destroy(f.0);
            }
        }
    }
    /// called by the element right after the asynchronous *_process call that writes (or reads) this block
    pub(crate) fn set_fence(&self, ctx: *mut ffi::b200vfx_ctx) {
        let mut f: *mut ffi::b200vfx_fence = std::ptr::null_mut();
        if unsafe { ffi::b200vfx_fence_create(ctx, &mut f) } == 0 {
            *self.fence.lock().unwrap() = Some(FencePtr(f));
        } else {
            unsafe { ffi::b200vfx_ctx_synchronize(ctx) };   // no fence: fall back to the synchronous contract
        }
    }
}
impl AsRef<[u8]> for PinnedBlock {
    fn as_ref(&self) -> &[u8] {
        self.wait();
        unsafe { std::slice::from_raw_parts(self.ptr, self.len) }
    }
}
impl AsMut<[u8]> for PinnedBlock {
    fn as_mut(&mut self) -> &mut [u8] {
        self.wait();
        unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) }
    }
}
impl Drop for PinnedBlock {
    fn drop(&mut self) {
        self.wait();   // the DMA engines may still be writing into the block
        unsafe { ffi::b200vfx_host_free(self.ptr as *mut _) }
    }
}

mod pinned_imp {
    use super::*;

    #[derive(Default)]
    pub struct PinnedAllocator;

    #[glib::object_subclass]
    impl ObjectSubclass for PinnedAllocator {
        const NAME: &'static str = "GstB200VfxPinnedAllocator";
        type Type = super::PinnedAllocator;
        type ParentType = gst::Allocator;
    }
    impl ObjectImpl for PinnedAllocator {}
    impl GstObjectImpl for PinnedAllocator {}
    impl AllocatorImpl for PinnedAllocator {
        fn alloc(&self, size: usize, params: Option<&gst::AllocationParams>) -> Result<gst::Memory, glib::BoolError> {
            let (prefix, padding) = params.map(|p| (p.prefix(), p.padding())).unwrap_or((0, 0));
            let total = prefix + size + padding;
            // cudaHostAlloc'ed, portable across contexts; page aligned, which satisfies any GstAllocationParams::align
            let ptr = unsafe { ffi::b200vfx_host_alloc(total) } as *mut u8;
            if ptr.is_null() {
                return Err(glib::bool_error!("b200vfx_host_alloc({total}) failed"));
            }
            let mut mem = gst::Memory::from_mut_slice(PinnedBlock { ptr, len: total, fence: std::sync::Mutex::new(None) });
            mem.get_mut().unwrap().resize(prefix as isize, size);
            Ok(mem)
        }
    }
}
glib::wrapper! {
    pub struct PinnedAllocator(ObjectSubclass<pinned_imp::PinnedAllocator>) @extends gst::Allocator, gst::Object;
}
pub static PINNED_ALLOCATOR: LazyLock<gst::Allocator> = LazyLock::new(|| glib::Object::new::<PinnedAllocator>().upcast());

// ---- the hooks every b200vfx BaseTransform element shares -----------------------------------------------------------
/// BaseTransformImpl::propose_allocation: tell upstream to allocate its buffers from our pinned allocator (and that we
/// handle GstVideoMeta, i.e. arbitrary strides -- every C entry point takes a stride).
pub fn propose_allocation(query: &mut gst::query::Allocation) -> Result<(), gst::LoggableError> {
    query.add_allocation_param(Some(&*PINNED_ALLOCATOR), gst::AllocationParams::default());
    query.add_allocation_meta::<gst_video::VideoMeta>(None);
    Ok(())
}

/// BaseTransformImpl::decide_allocation: our output buffers come from a pool backed by the pinned allocator unless
/// downstream insists on its own allocator.  `size` is GstVideoInfo::size of the negotiated output caps.
pub fn decide_allocation(query: &mut gst::query::Allocation, caps: &gst::Caps, size: u32) -> Result<(), gst::LoggableError> {
    let params = gst::AllocationParams::default();
    if query.allocation_params().is_empty() {
        query.add_allocation_param(Some(&*PINNED_ALLOCATOR), params.clone());
    } else {
        query.set_nth_allocation_param(0, Some(&*PINNED_ALLOCATOR), params.clone());
    }
    let (pool, min, max) = match query.allocation_pools().first() {
        Some((Some(pool), _, min, max)) => (pool.clone(), *min, *max),
        _ => (gst_video::VideoBufferPool::new().upcast(), 2, 0),
    };
    let mut config = pool.config();
    config.set_params(Some(caps), size, min.max(2), max);
    config.set_allocator(Some(&*PINNED_ALLOCATOR), Some(&params));
    config.add_option(gst_video::BUFFER_POOL_OPTION_VIDEO_META.as_ref());
    pool.set_config(config).map_err(|_| gst::loggable_error!(gst::CAT_RUST, "pinned pool rejected its configuration"))?;
    if query.allocation_pools().is_empty() {
        query.add_allocation_pool(Some(&pool), size, min.max(2), max);
    } else {
        query.set_nth_allocation_pool(0, Some(&pool), size, min.max(2), max);
    }
    Ok(())
}

// ---- device memory (`memory:B200Memory`) ---------------------------------------------------------------------------------
/// A frame that lives in HBM.  The C entry points accept its pointer directly (`b200vfx_pointer_is_device`): an element
/// that finds this feature on both pads passes `dev_ptr` / `stride` instead of mapping the buffer, and only the head and
/// tail of a chain call b200vfx_upload / b200vfx_download (which are asynchronous on the context stream; the tail
/// synchronises once per frame before pushing to a CPU element).
pub struct DeviceFrame {
    pub ctx: *mut ffi::b200vfx_ctx,
    pub dev_ptr: *mut std::ffi::c_void,
    pub stride: i32,
    pub size: usize,
}
unsafe impl Send for DeviceFrame {}
impl Drop for DeviceFrame {
    fn drop(&mut self) {
        unsafe { ffi::b200vfx_device_free(self.ctx, self.dev_ptr) }
    }
}
impl DeviceFrame {
    pub fn new(ctx: *mut ffi::b200vfx_ctx, stride: i32, rows: i32) -> Option<Self> {
        let size = stride as usize * rows as usize;
        let dev_ptr = unsafe { ffi::b200vfx_device_alloc(ctx, size) };
        (!dev_ptr.is_null()).then_some(DeviceFrame { ctx, dev_ptr, stride, size })
    }
    /// head of a chain: system-memory GstBuffer -> HBM
    pub fn upload(&self, frame: &gst_video::VideoFrameRef<&gst::BufferRef>, row_bytes: usize) -> bool {
        let src = frame.plane_data(0).unwrap();
        unsafe {
            ffi::b200vfx_upload(self.ctx, self.dev_ptr, self.stride, src.as_ptr() as *const _, frame.plane_stride()[0], row_bytes,
                                frame.height() as i32) == ffi::B200VFX_OK
        }
    }
    /// tail of a chain: HBM -> the (pinned) output buffer, then one synchronisation
    pub fn download(&self, frame: &mut gst_video::VideoFrameRef<&mut gst::BufferRef>, row_bytes: usize) -> bool {
        let (stride, rows) = (frame.plane_stride()[0], frame.height() as i32);
        let dst = frame.plane_data_mut(0).unwrap();
        unsafe {
            ffi::b200vfx_download(self.ctx, dst.as_mut_ptr() as *mut _, stride, self.dev_ptr, self.stride, row_bytes, rows) == ffi::B200VFX_OK
                && ffi::b200vfx_ctx_synchronize(self.ctx) == ffi::B200VFX_OK
        }
    }
}

/// transform_caps helper: offer every raw format also with the device-memory feature (as d3d12colorlut/imp.rs:236-246 does
/// with `memory:D3D12Memory`), so two b200vfx elements negotiate HBM between them and anything else gets system memory.
pub fn with_device_feature(caps: &gst::Caps) -> gst::Caps {
    let mut dev = caps.clone();
    {
        let dev = dev.make_mut();
        for i in 0..dev.size() {
            dev.set_features(i, Some(gst::CapsFeatures::new([CAPS_FEATURE_MEMORY_B200])));
        }
    }
    dev.merge(caps.clone());
    dev
}
