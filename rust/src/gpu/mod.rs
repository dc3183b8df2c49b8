//! `ibu::gpu` — the B200 (sm_100a) path of the bulk record operations, behind the `gpu` feature.
//!
//! The module a fork of noamteyssier/ibu adds next to `io` and `parallel` (src/lib.rs:178-181 there
//! gains `#[cfg(feature = "gpu")] pub mod gpu;`).  Files, `Header` and `Record` are untouched: both
//! are `#[repr(C)]` + `bytemuck::Pod` (src/constructs/header.rs:44-61, record.rs:58-66) and are passed
//! across the boundary as they are.  Everything here is a thin, safe wrapper over `ffi.rs`, which is
//! generated from include/ibu_b200.h; there is no CPU fallback — without a usable B200 every call
//! returns `IbuError::Process("CUDA ...")`.
//!
//! What maps to what in the reference:
//!   * `GpuParallelReader::process_gpu`      <- `ParallelReader::process_parallel` (src/parallel.rs:250-296,
//!                                               src/io/mmap.rs:286-332) for the built-in reductions;
//!   * `Ops::TABLE` / `BarcodeTable`          <- the HashMap<barcode, count> processor (src/parallel.rs:79-98)
//!                                               + distinct UMIs per barcode;
//!   * `load_to_device`                       <- `load_to_vec` (src/io/reader.rs:510-535);
//!   * `GpuGroup`                             <- the thread fan-out of mmap.rs:297-332, one thread per GPU;
//!   * `check`                                <- `IbuError` (src/error.rs:56-128), variant for variant.
//!
//! Not compiled in the build image of the ibu_b200 repository (no rustc there):
//! tests/test_rust_ffi.py checks every `ffi::` call below against the generated declarations.

pub mod ffi;

use crate::{Header, IbuError, MmapReader, Record, Result};
use std::ffi::{CStr, CString};
use std::ops::Range;
use std::os::raw::{c_int, c_void};
use std::path::Path;
use std::ptr;

pub use ffi::ibu_barcode_row_t as BarcodeRow;
pub use ffi::ibu_group_timing_t as GroupTiming;
pub use ffi::ibu_reduce_result_t as ReduceResult;

// `Header` / `Record` and their C twins are the same 32 / 24 bytes: cast, never convert.
const _: () = assert!(std::mem::size_of::<Header>() == std::mem::size_of::<ffi::ibu_header_t>());
const _: () = assert!(std::mem::size_of::<Record>() == std::mem::size_of::<ffi::ibu_record_t>());

fn new_err() -> ffi::ibu_error_t {
    ffi::ibu_error_t { code: 0, sys: 0, a: 0, b: 0, msg: [0; 232] }
}

fn msg_of(e: &ffi::ibu_error_t) -> String {
    let s = unsafe { CStr::from_ptr(e.msg.as_ptr()) }.to_string_lossy().into_owned();
    if s.is_empty() {
        unsafe { CStr::from_ptr(ffi::ibu_strerror(e.code)) }.to_string_lossy().into_owned()
    } else {
        s
    }
}

/// Status + payload -> the crate's error, variant for variant (src/error.rs:56-128).
fn check(rc: c_int, e: &ffi::ibu_error_t) -> Result<()> {
    match rc {
        ffi::IBU_OK => Ok(()),
        ffi::IBU_ERR_IO => Err(IbuError::Io(std::io::Error::from_raw_os_error(e.sys))),
        ffi::IBU_ERR_INVALID_MAGIC => Err(IbuError::InvalidMagicNumber { expected: e.a as u32, actual: e.b as u32 }),
        ffi::IBU_ERR_TRUNCATED_RECORD => Err(IbuError::TruncatedRecord { pos: e.a as usize }),
        ffi::IBU_ERR_INVALID_VERSION => Err(IbuError::InvalidVersion { expected: e.a as u32, actual: e.b as u32 }),
        ffi::IBU_ERR_INVALID_BARCODE_LENGTH => Err(IbuError::InvalidBarcodeLength(e.a as u32)),
        ffi::IBU_ERR_INVALID_UMI_LENGTH => Err(IbuError::InvalidUmiLength(e.a as u32)),
        ffi::IBU_ERR_INVALID_MAP_SIZE => Err(IbuError::InvalidMapSize),
        ffi::IBU_ERR_INVALID_INDEX => Err(IbuError::InvalidIndex { idx: e.a as usize, max: e.b as usize }),
        // IBU_ERR_PROCESS (callback), IBU_ERR_CUDA, IBU_ERR_NCCL, IBU_ERR_ARG, IBU_ERR_NOMEM
        _ => Err(IbuError::Process(msg_of(e).into())),
    }
}

fn c_path<P: AsRef<Path>>(path: P) -> Result<CString> {
    CString::new(path.as_ref().to_string_lossy().as_bytes())
        .map_err(|_| IbuError::Io(std::io::Error::from(std::io::ErrorKind::InvalidInput)))
}

/// Which work one pass over the records does (`IBU_OP_*`); `REDUCE` is always on.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct Ops(pub u32);
impl Ops {
    pub const REDUCE: Ops = Ops(ffi::IBU_OP_REDUCE);
    pub const TABLE: Ops = Ops(ffi::IBU_OP_TABLE);
    pub const KEEP: Ops = Ops(ffi::IBU_OP_KEEP);
    pub const UNPACK: Ops = Ops(ffi::IBU_OP_UNPACK);
}
impl std::ops::BitOr for Ops {
    type Output = Ops;
    fn bitor(self, o: Ops) -> Ops {
        Ops(self.0 | o.0)
    }
}

/// Staging configuration of a context (`ibu_gpu_config_t`); zeros pick the defaults.
#[derive(Clone, Copy, Debug, Default)]
pub struct GpuConfig {
    pub chunk_records: u32,
    pub n_slots: u32,
    pub copy_threads: u32,
}
impl GpuConfig {
    fn raw(&self) -> ffi::ibu_gpu_config_t {
        ffi::ibu_gpu_config_t { chunk_records: self.chunk_records, n_slots: self.n_slots, copy_threads: self.copy_threads, reserved: 0 }
    }
}

/// One GPU: streams, pinned chunk slots, scratch.  `Send`, one call at a time per context.
pub struct GpuContext {
    raw: *mut ffi::ibu_gpu_ctx_t,
    owned: bool,
}
unsafe impl Send for GpuContext {}

impl GpuContext {
    pub fn device_count() -> usize {
        unsafe { ffi::ibu_gpu_device_count() }.max(0) as usize
    }

    pub fn new(device: usize, cfg: GpuConfig) -> Result<Self> {
        let (mut raw, mut e, c) = (ptr::null_mut(), new_err(), cfg.raw());
        check(unsafe { ffi::ibu_gpu_ctx_create(device as c_int, &c, &mut raw, &mut e) }, &e)?;
        Ok(GpuContext { raw, owned: true })
    }

    pub fn device(&self) -> usize {
        unsafe { ffi::ibu_gpu_ctx_device(self.raw) as usize }
    }

    pub fn synchronize(&self) -> Result<()> {
        let mut e = new_err();
        check(unsafe { ffi::ibu_gpu_synchronize(self.raw, ptr::null_mut(), &mut e) }, &e)
    }

    /// Host records -> host ASCII, pipelined H2D / unpack+validate / D2H (bitnuc `from_2bit` order).
    /// `bc` is `[n][bc_len]`, `umi` is `[n][umi_len]`, dense; `flags` (one byte per record) optional.
    pub fn unpack_batch(&self, recs: &[Record], bc_len: u32, umi_len: u32, bc: &mut [u8], umi: &mut [u8],
                        flags: Option<&mut [u8]>) -> Result<ReduceResult> {
        assert!(bc.len() == recs.len() * bc_len as usize && umi.len() == recs.len() * umi_len as usize);
        let (mut out, mut e) = (ReduceResult::default_zero(), new_err());
        let f = flags.map_or(ptr::null_mut(), |f| { assert!(f.len() == recs.len()); f.as_mut_ptr() });
        check(unsafe { ffi::ibu_gpu_unpack_host(self.raw, recs.as_ptr() as *const ffi::ibu_record_t, recs.len() as u64,
                                                bc_len, umi_len, bc.as_mut_ptr(), umi.as_mut_ptr(), f, &mut out, &mut e) }, &e)?;
        Ok(out)
    }

    /// Host ASCII -> host records (bitnuc `as_2bit` order), the source of `Writer::write_batch`
    /// (src/io/writer.rs:315-351).  Non-ACGT bytes are data: counted in the result, flagged per record.
    pub fn pack_batch(&self, bc: &[u8], umi: &[u8], bc_len: u32, umi_len: u32, index_base: u64,
                      out: &mut [Record]) -> Result<ReduceResult> {
        assert!(bc.len() == out.len() * bc_len as usize && umi.len() == out.len() * umi_len as usize);
        let (mut res, mut e) = (ReduceResult::default_zero(), new_err());
        check(unsafe { ffi::ibu_gpu_pack_host(self.raw, bc.as_ptr(), umi.as_ptr(), ptr::null(), index_base, out.len() as u64,
                                              bc_len, umi_len, out.as_mut_ptr() as *mut ffi::ibu_record_t, ptr::null_mut(),
                                              &mut res, &mut e) }, &e)?;
        Ok(res)
    }

    /// Per-barcode (records, distinct UMIs) of device-resident records, rows sorted by barcode.
    pub fn barcode_count(&self, recs: &DeviceRecords, header: &Header) -> Result<BarcodeTable<'_>> {
        let mode = ((header.bc_len & 0x3F) << 8 | (header.umi_len & 0x3F) << 16) as c_int; // IBU_COUNT_LENS
        let (mut t, mut e) = (empty_table(), new_err());
        check(unsafe { ffi::ibu_gpu_barcode_count(self.raw, recs.ptr, recs.len, mode, &mut t, ptr::null_mut(), &mut e) }, &e)?;
        Ok(BarcodeTable { ctx: self, raw: t })
    }

    /// Device sort by `Record`'s `Ord` (src/constructs/record.rs:29-32,58); after it `Header::set_sorted` is truthful.
    pub fn sort_records(&self, recs: &DeviceRecords) -> Result<DeviceRecords> {
        let (mut d, mut e) = (ptr::null_mut(), new_err());
        check(unsafe { ffi::ibu_gpu_malloc(self.raw, recs.len as usize * 24, &mut d, &mut e) }, &e)?;
        let sorted = DeviceRecords { ctx: self.raw, ptr: d as *mut ffi::ibu_record_t, len: recs.len };
        check(unsafe { ffi::ibu_gpu_sort_records(self.raw, recs.ptr, recs.len, sorted.ptr, ptr::null_mut(), &mut e) }, &e)?;
        Ok(sorted)
    }

    /// Device records -> an `.ibu` file written with `Writer::write_batch` semantics.
    pub fn write_records<P: AsRef<Path>>(&self, path: P, header: Header, recs: &DeviceRecords) -> Result<()> {
        let (p, mut w, mut e) = (c_path(path)?, ptr::null_mut(), new_err());
        check(unsafe { ffi::ibu_writer_open(p.as_ptr(), &header as *const Header as *const ffi::ibu_header_t, &mut w, &mut e) }, &e)?;
        let rc = unsafe { ffi::ibu_gpu_write_records(self.raw, w, recs.ptr, recs.len, &mut e) };
        let rc = if rc == ffi::IBU_OK { unsafe { ffi::ibu_writer_finish(w, &mut e) } } else { rc };
        unsafe { ffi::ibu_writer_close(w) };
        check(rc, &e)
    }
}

impl Drop for GpuContext {
    fn drop(&mut self) {
        if self.owned {
            unsafe { ffi::ibu_gpu_ctx_destroy(self.raw) }
        }
    }
}

trait ZeroDefault {
    fn default_zero() -> Self;
}
impl ZeroDefault for ReduceResult {
    fn default_zero() -> Self {
        ReduceResult { n_records: 0, sum_barcode: 0, sum_umi: 0, sum_index: 0, xor_all: 0, n_bad_barcode: 0, n_bad_umi: 0, n_bad_records: 0 }
    }
}

fn empty_table() -> ffi::ibu_barcode_table_t {
    ffi::ibu_barcode_table_t { d_rows: ptr::null_mut(), n_rows: 0, n_records: 0, n_distinct_pairs: 0, input_was_sorted: 0, reserved: 0 }
}

/// Records resident in HBM (the device twin of the `Vec<Record>` of `load_to_vec`).
pub struct DeviceRecords {
    ctx: *mut ffi::ibu_gpu_ctx_t,
    ptr: *mut ffi::ibu_record_t,
    len: u64,
}
impl DeviceRecords {
    pub fn len(&self) -> usize {
        self.len as usize
    }
    pub fn is_empty(&self) -> bool {
        self.len == 0
    }
    pub fn to_vec(&self) -> Result<Vec<Record>> {
        let mut v = vec![Record::new(0, 0, 0); self.len as usize];
        let mut e = new_err();
        check(unsafe { ffi::ibu_gpu_memcpy_d2h(self.ctx, v.as_mut_ptr() as *mut c_void, self.ptr as *const c_void, v.len() * 24, &mut e) }, &e)?;
        Ok(v)
    }
}
impl Drop for DeviceRecords {
    fn drop(&mut self) {
        unsafe { ffi::ibu_gpu_free(self.ctx, self.ptr as *mut c_void) }
    }
}

/// Per-barcode rows on the device; `to_vec` brings them to the host.
pub struct BarcodeTable<'a> {
    ctx: &'a GpuContext,
    raw: ffi::ibu_barcode_table_t,
}
impl BarcodeTable<'_> {
    pub fn n_rows(&self) -> usize {
        self.raw.n_rows as usize
    }
    pub fn n_records(&self) -> u64 {
        self.raw.n_records
    }
    pub fn n_distinct_pairs(&self) -> u64 {
        self.raw.n_distinct_pairs
    }
    pub fn input_was_sorted(&self) -> bool {
        self.raw.input_was_sorted != 0
    }
    pub fn to_vec(&self) -> Result<Vec<BarcodeRow>> {
        let mut v = vec![BarcodeRow { barcode: 0, n_records: 0, n_distinct_umi: 0 }; self.n_rows()];
        let mut e = new_err();
        // (through the context's pinned landing area: a plain D2H into a fresh Vec is a staged copy)
        check(unsafe { ffi::ibu_gpu_table_to_host(self.ctx.raw, &self.raw, v.as_mut_ptr(), &mut e) }, &e)?;
        Ok(v)
    }
}
impl Drop for BarcodeTable<'_> {
    fn drop(&mut self) {
        unsafe { ffi::ibu_gpu_table_free(self.ctx.raw, &mut self.raw) }
    }
}

/// Device path of `load_to_vec` (src/io/reader.rs:510-535): same header validation, same size check.
pub fn load_to_device<P: AsRef<Path>>(ctx: &GpuContext, path: P) -> Result<(Header, DeviceRecords)> {
    let (p, mut h, mut d, mut n, mut e) = (c_path(path)?, Header::new(1, 1), ptr::null_mut(), 0u64, new_err());
    check(unsafe { ffi::ibu_gpu_load_to_device(ctx.raw, p.as_ptr(), 0, u64::MAX, &mut h as *mut Header as *mut ffi::ibu_header_t,
                                               &mut d, &mut n, &mut e) }, &e)?;
    Ok((h, DeviceRecords { ctx: ctx.raw, ptr: d, len: n }))
}

/// What one pass produced: the 8-word reduction always, the table / resident records when asked for.
pub struct Processed<'a> {
    pub result: ReduceResult,
    pub table: Option<BarcodeTable<'a>>,
    pub records: Option<DeviceRecords>,
}

unsafe extern "C" fn chunk_trampoline<F>(user: *mut c_void, start: u64, n: u64, r: *const ReduceResult) -> c_int
where
    F: FnMut(u64, u64, &ReduceResult) -> Result<()>,
{
    let f = &mut *(user as *mut F);
    match std::panic::catch_unwind(std::panic::AssertUnwindSafe(|| f(start, n, &*r))) {
        Ok(Ok(())) => 0,
        _ => 1, // -> IBU_ERR_PROCESS, like an Err from on_batch_complete (src/io/mmap.rs:322-326)
    }
}

/// GPU counterpart of `ParallelReader` (src/parallel.rs:250-296), implemented for `MmapReader`.
pub trait GpuParallelReader {
    /// Records `range` are staged chunk by chunk to the GPU, validated and reduced there;
    /// `on_batch` is `ParallelProcessor::on_batch_complete` (src/parallel.rs:141-151) per staged chunk,
    /// on the calling thread, in chunk order.  Do not call into the same context from `on_batch`.
    fn process_gpu<F>(&self, ctx: &GpuContext, range: Range<usize>, ops: Ops, on_batch: F) -> Result<Processed<'_>>
    where
        F: FnMut(u64, u64, &ReduceResult) -> Result<()>;
}

impl GpuParallelReader for MmapReader {
    fn process_gpu<F>(&self, ctx: &GpuContext, range: Range<usize>, ops: Ops, mut on_batch: F) -> Result<Processed<'_>>
    where
        F: FnMut(u64, u64, &ReduceResult) -> Result<()>,
    {
        // The map is already open in this process: hand the record slice over
        // (MmapReader::slice, src/io/mmap.rs:253-270, applies its own bounds rules).
        let recs = self.slice(range.start, range.end)?;
        let header = self.header();
        let (mut table, mut d_records) = (empty_table(), ptr::null_mut());
        let req = ffi::ibu_process_request_t {
            ops: ops.0 | ffi::IBU_OP_REDUCE,
            table_mode: 0,
            table: &mut table,
            d_records: &mut d_records,
            h_bc_ascii: ptr::null_mut(),
            h_umi_ascii: ptr::null_mut(),
            h_flags: ptr::null_mut(),
        };
        let (mut out, mut e) = (ReduceResult::default_zero(), new_err());
        let rc = unsafe {
            ffi::ibu_gpu_process_host_ops(ctx.raw, recs.as_ptr() as *const ffi::ibu_record_t, recs.len() as u64, header.bc_len,
                                          header.umi_len, &req, &mut out, Some(chunk_trampoline::<F>),
                                          &mut on_batch as *mut F as *mut c_void, &mut e)
        };
        check(rc, &e)?;
        // the borrow of `ctx` outlives `self` in practice; a fork ties both to one lifetime parameter
        let ctx_ref: &GpuContext = unsafe { &*(ctx as *const GpuContext) };
        Ok(Processed {
            result: out,
            table: if ops.0 & ffi::IBU_OP_TABLE != 0 { Some(BarcodeTable { ctx: ctx_ref, raw: table }) } else { None },
            records: if ops.0 & ffi::IBU_OP_KEEP != 0 {
                Some(DeviceRecords { ctx: ctx.raw, ptr: d_records, len: recs.len() as u64 })
            } else {
                None
            },
        })
    }
}

/// Page-locks the pages of `range` of an open file so that repeated passes DMA out of the page cache
/// (`ibu_mmap_pin_range`); the handle is the library's own reader over the same path.
pub struct PinnedFile {
    raw: *mut ffi::ibu_mmap_reader_t,
    range: Range<u64>,
}
impl PinnedFile {
    pub fn open<P: AsRef<Path>>(path: P, range: Option<Range<u64>>) -> Result<Self> {
        let (p, mut raw, mut e) = (c_path(path)?, ptr::null_mut(), new_err());
        check(unsafe { ffi::ibu_mmap_open(p.as_ptr(), &mut raw, &mut e) }, &e)?;
        let range = range.unwrap_or(0..unsafe { ffi::ibu_mmap_len(raw) } as u64);
        let rc = unsafe { ffi::ibu_mmap_pin_range(raw, range.start, range.end, &mut e) };
        if rc != ffi::IBU_OK {
            unsafe { ffi::ibu_mmap_close(raw) };
            check(rc, &e)?;
        }
        Ok(PinnedFile { raw, range })
    }
}
impl Drop for PinnedFile {
    fn drop(&mut self) {
        unsafe {
            ffi::ibu_mmap_unpin_range(self.raw, self.range.start, self.range.end);
            ffi::ibu_mmap_close(self.raw);
        }
    }
}

/// Merged table of a multi-GPU pass, in host memory, rows sorted by barcode.
pub struct HostTable {
    raw: ffi::ibu_host_table_t,
}
impl HostTable {
    pub fn rows(&self) -> &[BarcodeRow] {
        if self.raw.h_rows.is_null() {
            &[]
        } else {
            unsafe { std::slice::from_raw_parts(self.raw.h_rows, self.raw.n_rows as usize) }
        }
    }
    pub fn n_records(&self) -> u64 {
        self.raw.n_records
    }
    pub fn n_distinct_pairs(&self) -> u64 {
        self.raw.n_distinct_pairs
    }
}
impl Drop for HostTable {
    fn drop(&mut self) {
        unsafe { ffi::ibu_free(self.raw.h_rows as *mut c_void) }
    }
}

/// How the shards' de-duplicated (barcode, umi) pairs travel to their owners.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Exchange {
    Auto,
    PeerCopy,
    Host,
    Nccl,
}
impl Exchange {
    fn raw(self) -> u32 {
        match self {
            Exchange::Auto => ffi::IBU_EXCHANGE_AUTO,
            Exchange::PeerCopy => ffi::IBU_EXCHANGE_P2P,
            Exchange::Host => ffi::IBU_EXCHANGE_HOST,
            Exchange::Nccl => ffi::IBU_EXCHANGE_NCCL,
        }
    }
}

/// The GPUs of one box: rank r processes `shard_range(len, r, size)`, the partition of
/// `process_parallel` (src/io/mmap.rs:297-307); results merge like the reference processors'
/// `on_batch_complete` (mmap.rs:365-372), the table by one exchange of de-duplicated pairs.
pub struct GpuGroup {
    raw: *mut ffi::ibu_gpu_group_t,
}
unsafe impl Send for GpuGroup {}

impl GpuGroup {
    pub fn new(devices: &[usize], cfg: GpuConfig) -> Result<Self> {
        let devs: Vec<c_int> = devices.iter().map(|&d| d as c_int).collect();
        let (mut raw, mut e, c) = (ptr::null_mut(), new_err(), cfg.raw());
        check(unsafe { ffi::ibu_gpu_group_create(devs.as_ptr(), devs.len() as u32, &c, &mut raw, &mut e) }, &e)?;
        Ok(GpuGroup { raw })
    }

    pub fn size(&self) -> usize {
        unsafe { ffi::ibu_gpu_group_size(self.raw) as usize }
    }

    /// `len / world` each, the last rank takes the remainder.
    pub fn shard_range(len: u64, rank: u32, world: u32) -> Range<u64> {
        let (mut s, mut t) = (0u64, 0u64);
        unsafe { ffi::ibu_shard_range(len, rank, world, &mut s, &mut t) };
        s..t
    }

    /// One pass over `range` of the file on all GPUs: the merged reduction and, with `Ops::TABLE`,
    /// the exact merged per-barcode table.
    pub fn process(&self, reader: &MmapReader, range: Range<usize>, ops: Ops, exchange: Exchange)
                   -> Result<(ReduceResult, Option<HostTable>, GroupTiming)> {
        let recs = reader.slice(range.start, range.end)?;
        let header = reader.header();
        let mut table = HostTable { raw: ffi::ibu_host_table_t { h_rows: ptr::null_mut(), n_rows: 0, n_records: 0, n_distinct_pairs: 0 } };
        let mut timing: GroupTiming = unsafe { std::mem::zeroed() };
        let req = ffi::ibu_group_request_t {
            ops: (ops.0 | ffi::IBU_OP_REDUCE) & !ffi::IBU_OP_KEEP,
            table_mode: 0,
            exchange: exchange.raw(),
            reserved: 0,
            table: &mut table.raw,
            d_records: ptr::null_mut(),
            shard_records: ptr::null_mut(),
            timing: &mut timing,
        };
        let (mut out, mut e) = (ReduceResult::default_zero(), new_err());
        check(unsafe { ffi::ibu_gpu_group_process_host(self.raw, recs.as_ptr() as *const ffi::ibu_record_t, recs.len() as u64,
                                                       header.bc_len, header.umi_len, &req, &mut out, &mut e) }, &e)?;
        let table = if ops.0 & ffi::IBU_OP_TABLE != 0 { Some(table) } else { None };
        Ok((out, table, timing))
    }
}

impl Drop for GpuGroup {
    fn drop(&mut self) {
        unsafe { ffi::ibu_gpu_group_destroy(self.raw) }
    }
}
