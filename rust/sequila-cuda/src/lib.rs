//! UNCOMPILED — see rust/README.md.
//! Safe wrappers over `sequila-cuda-sys`, shaped for `sequila-core/src/physical_planner/joins/interval_join.rs`:
//!
//! * [`CudaIndex`] is what `IntervalJoinAlgorithm::Cuda` holds inside the `Arc<JoinLeftData>` (interval_join.rs:49-68):
//!   immutable after `build`, probed from every partition's stream at once => `Send + Sync`;
//! * [`CudaStream`] belongs to one `IntervalJoinStream` (one per partition, interval_join.rs:528-556) => `Send` only;
//!   every entry point of the library sets the CUDA device itself, so tokio may move the stream between worker threads;
//! * [`Tile`] is a collected probe tile: pinned buffers that return to the library's pool on drop, cheap to wrap into
//!   Arrow `Buffer`s (`Buffer::from_custom_allocation`) without a copy.
use std::ffi::{CStr, CString};
use std::ptr::{null_mut, NonNull};
use std::sync::Arc;

use datafusion::error::{DataFusionError, Result};
use sequila_cuda_sys as sys;

fn exec_err(msg: *const std::os::raw::c_char) -> DataFusionError {
    // the convention of session_context.rs:121-125: the library's message, verbatim
    DataFusionError::Execution(unsafe { CStr::from_ptr(msg) }.to_string_lossy().into_owned())
}

pub struct CudaContext(NonNull<sys::sq_ctx>);
unsafe impl Send for CudaContext {}
unsafe impl Sync for CudaContext {}

impl CudaContext {
    /// Fails when no CUDA device is usable: there is no CPU fallback behind `alg=Cuda`.
    pub fn new(device: i32) -> Result<Arc<Self>> {
        let mut p = null_mut();
        let rc = unsafe { sys::sq_ctx_create(device, &mut p) };
        if rc != sys::SQ_OK { return Err(exec_err(unsafe { sys::sq_last_error(std::ptr::null()) })); }
        Ok(Arc::new(Self(NonNull::new(p).unwrap())))
    }
    /// `SET sequila.cuda_<name> TO <value>` (the keys listed in sequila_cuda.h)
    pub fn set_option(&self, key: &str, value: &str) -> Result<()> {
        let (k, v) = (CString::new(key).unwrap(), CString::new(value).unwrap());
        match unsafe { sys::sq_ctx_set_option(self.0.as_ptr(), k.as_ptr(), v.as_ptr()) } {
            sys::SQ_OK => Ok(()),
            _ => Err(exec_err(unsafe { sys::sq_last_error(self.0.as_ptr()) })),
        }
    }
}
impl Drop for CudaContext { fn drop(&mut self) { unsafe { sys::sq_ctx_destroy(self.0.as_ptr()) } } }

pub struct CudaIndex { ptr: NonNull<sys::sq_index>, ctx: Arc<CudaContext> }
unsafe impl Send for CudaIndex {}
unsafe impl Sync for CudaIndex {}

impl CudaIndex {
    /// Replaces `update_hashmap` + `IntervalJoinAlgorithm::new` (interval_join.rs:662-683): `key_hash[i]` =
    /// `create_hashes(on_left)` of row i, `start` / `end` = `evaluate_as_i32(..)`; row i is left index i.
    pub fn build(ctx: &Arc<CudaContext>, key_hash: &[u64], start: &[i32], end: &[i32]) -> Result<Self> {
        assert!(key_hash.len() == start.len() && start.len() == end.len());
        let mut p = null_mut();
        let rc = unsafe { sys::sq_index_build(ctx.0.as_ptr(), key_hash.as_ptr(), start.as_ptr(), end.as_ptr(), key_hash.len() as u64, &mut p) };
        if rc != sys::SQ_OK { return Err(exec_err(unsafe { sys::sq_last_error(ctx.0.as_ptr()) })); }
        Ok(Self { ptr: NonNull::new(p).unwrap(), ctx: ctx.clone() })
    }
    /// `build_mem_used` gauge / `MemoryReservation::try_grow` (interval_join.rs:629-631)
    pub fn bytes(&self) -> usize { unsafe { sys::sq_index_bytes(self.ptr.as_ptr()) as usize } }
    /// `SET sequila.cuda_build_ids TO positions` was in force at build time: `left_idx` holds positions in the index's
    /// (key, start) order and the payload columns registered with the index live in that order on the device.  The patch of
    /// interval_join.rs keeps `take` on the host and therefore builds with `rows`; a stream that moves `take` onto the
    /// device (sequila_exec.h does) switches this on and never looks at `left_idx` itself.
    pub fn uses_positions(&self) -> bool { unsafe { sys::sq_index_uses_positions(self.ptr.as_ptr()) != 0 } }
    /// position -> build row, for a caller that wants both the locality and the rows
    pub fn position_rows(&self) -> Result<Vec<u32>> {
        let n = unsafe { sys::sq_index_rows(self.ptr.as_ptr()) } as usize;
        let mut rows = vec![0u32; n];
        let rc = unsafe { sys::sq_index_position_rows(self.ptr.as_ptr(), rows.as_mut_ptr()) };
        if rc != sys::SQ_OK { return Err(exec_err(unsafe { sys::sq_last_error(self.ctx.0.as_ptr()) })); }
        Ok(rows)
    }
}
impl Drop for CudaIndex { fn drop(&mut self) { unsafe { sys::sq_index_free(self.ptr.as_ptr()) } } }

/// One collected probe tile.  `left_idx[k]` is `pos as u32` (interval_join.rs:1590); `counts[i]` is `rle_right`
/// (interval_join.rs:1604) — the stream expands it into `index_right` exactly as interval_join.rs:1611-1618 does, or fuses
/// the expansion into its `take` of the probe columns.
pub struct Tile { out: sys::sq_tile_out, ctx: Arc<CudaContext> }
unsafe impl Send for Tile {}
impl Tile {
    pub fn n_pairs(&self) -> usize { self.out.n_pairs as usize }
    pub fn left_idx(&self) -> &[u32] { if self.out.left_idx.is_null() { &[] } else { unsafe { std::slice::from_raw_parts(self.out.left_idx, self.n_pairs()) } } }
    /// 4-byte counts (tiles submitted without SQ_TILE_COUNTS_U8, or a tile in which some row has more than 255 hits)
    pub fn counts(&self) -> &[u32] { if self.out.counts.is_null() || self.out.counts_width != 4 { &[] } else { unsafe { std::slice::from_raw_parts(self.out.counts, self.out.n_rows as usize) } } }
    /// one-byte counts (SQ_TILE_COUNTS_U8 and every count of the tile < 256): `rle_right` expands from either width
    pub fn counts_u8(&self) -> &[u8] { if self.out.counts.is_null() || self.out.counts_width != 1 { &[] } else { unsafe { std::slice::from_raw_parts(self.out.counts as *const u8, self.out.n_rows as usize) } } }
}
impl Drop for Tile {
    fn drop(&mut self) {
        for p in [self.out.left_idx, self.out.right_idx, self.out.counts] {
            unsafe { sys::sq_host_free(self.ctx.0.as_ptr(), p as *mut _) }
        }
    }
}

pub struct CudaStream { ptr: NonNull<sys::sq_stream>, ctx: Arc<CudaContext> }
unsafe impl Send for CudaStream {}

impl CudaStream {
    pub fn new(ctx: &Arc<CudaContext>) -> Result<Self> {
        let mut p = null_mut();
        let rc = unsafe { sys::sq_stream_create(ctx.0.as_ptr(), &mut p) };
        if rc != sys::SQ_OK { return Err(exec_err(unsafe { sys::sq_last_error(ctx.0.as_ptr()) })); }
        Ok(Self { ptr: NonNull::new(p).unwrap(), ctx: ctx.clone() })
    }
    fn err(&self) -> DataFusionError { exec_err(unsafe { sys::sq_stream_last_error(self.ptr.as_ptr()) }) }

    /// Enqueue one tile (>= 1 coalesced probe batch, order kept).  The slices must stay alive until the ticket is
    /// collected — the stream keeps the batches it coalesced — and should live in pinned memory (`sq_host_alloc`).
    /// `Err` with `SQ_EBUSY` semantics never reaches the caller: the stream collects before it submits (see the patch).
    pub fn submit(&mut self, index: &CudaIndex, key_hash: &[u64], start: &[i32], end: &[i32], flags: u32) -> Result<u64> {
        let mut ticket = 0u64;
        let rc = unsafe { sys::sq_stream_submit(self.ptr.as_ptr(), index.ptr.as_ptr(), key_hash.as_ptr(), start.as_ptr(), end.as_ptr(),
                                                key_hash.len() as u32, flags, &mut ticket) };
        if rc != sys::SQ_OK { return Err(self.err()); }
        Ok(ticket)
    }
    pub fn in_flight(&self) -> usize { unsafe { sys::sq_stream_in_flight(self.ptr.as_ptr()) as usize } }
    /// Wait for the oldest ticket.
    pub fn collect(&mut self, ticket: u64) -> Result<Tile> {
        let mut out = sys::sq_tile_out { n_pairs: 0, n_rows: 0, counts_width: 0, left_idx: null_mut(), right_idx: null_mut(), counts: null_mut() };
        let rc = unsafe { sys::sq_stream_collect(self.ptr.as_ptr(), ticket, &mut out) };
        if rc != sys::SQ_OK { return Err(self.err()); }
        Ok(Tile { out, ctx: self.ctx.clone() })
    }
}
impl Drop for CudaStream { fn drop(&mut self) { unsafe { sys::sq_stream_free(self.ptr.as_ptr()) } } }
