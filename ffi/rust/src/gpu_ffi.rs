//! Thin FFI over include/jne.h (hand-written equivalent of `bindgen include/jne.h include/jne_dat.h`).
//! One `Gpu` per process; a context serves one caller thread at a time.

use std::ffi::{CStr, CString};
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct JneCtx {
    _private: [u8; 0],
}

/// `jne_rows_sink` of include/jne.h: rows of seeds[first .. first + count), valid during the call only.
pub type JneRowsSink = extern "C" fn(user: *mut c_void, first: u64, count: u64, rows: *const f64) -> c_int;

extern "C" {
    fn jne_init(device_ids: *const c_int, n_devices: c_int, out: *mut *mut JneCtx) -> c_int;
    fn jne_shutdown(ctx: *mut JneCtx);
    fn jne_last_error(ctx: *const JneCtx) -> *const c_char;
    fn jne_num_eigs(model: u8, dim: u32) -> c_int;
    fn jne_multi_width(model_mask: u32, dim: u32) -> c_int;
    fn jne_eigs_batch(ctx: *mut JneCtx, model: u8, dim: u32, steps: u32, seeds: *const u32, n: u64, out: *mut f64) -> c_int;
    fn jne_eigs_batch_multi(ctx: *mut JneCtx, model_mask: u32, dim: u32, steps: u32, seeds: *const u32, n: u64, out: *mut f64) -> c_int;
    fn jne_submit(ctx: *mut JneCtx, model: u8, dim: u32, steps: u32, seeds: *const u32, n: u64, out: *mut f64) -> i64;
    fn jne_submit_multi(ctx: *mut JneCtx, model_mask: u32, dim: u32, steps: u32, seeds: *const u32, n: u64, out: *mut f64) -> i64;
    fn jne_wait(ctx: *mut JneCtx, ticket: i64) -> c_int;
    fn jne_eigs_batch_multi_stream(ctx: *mut JneCtx, model_mask: u32, dim: u32, steps: u32, seeds: *const u32, n: u64,
                                   sink: JneRowsSink, user: *mut c_void) -> c_int;
    fn jne_run_models_simulation(ctx: *mut JneCtx, model_mask: u32, dim: u32, steps: u32, num_runs: u64,
                                 filenames: *const *const c_char, quiet: c_int, device_ids: *const c_int,
                                 n_devices: c_int, stats: *mut u64) -> c_int;
}

/// Owns a jne_ctx over every visible GPU.
pub struct Gpu(*mut JneCtx);
unsafe impl Send for Gpu {}

/// A batch in flight (`jne_submit`); the rows are valid after `wait`.
pub struct Pending {
    ticket: i64,
    rows: Vec<f64>,
}

impl Gpu {
    pub fn new() -> Result<Self, String> {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { jne_init(std::ptr::null(), 0, &mut ctx) };
        if rc != 0 {
            return Err(last_error(std::ptr::null()));
        }
        Ok(Gpu(ctx))
    }

    /// eigenvalues per run: dim + 1 for models 1 and 3 (thread_manager.rs:40-44)
    pub fn num_eigs(model: u8, dim: usize) -> usize {
        let p = unsafe { jne_num_eigs(model, dim as u32) };
        assert!(p > 0, "invalid model/dim");
        p as usize
    }

    /// Eigenvalues of `seeds`, row-major n x p, each row descending (johansen_statistics.rs:45).
    pub fn eigs_batch(&self, model: u8, dim: usize, steps: usize, seeds: &[u32]) -> Result<Vec<f64>, String> {
        let mut out = vec![0f64; seeds.len() * Self::num_eigs(model, dim)];
        let rc = unsafe {
            jne_eigs_batch(self.0, model, dim as u32, steps as u32, seeds.as_ptr(), seeds.len() as u64, out.as_mut_ptr())
        };
        if rc != 0 {
            return Err(last_error(self.0));
        }
        Ok(out)
    }

    /// All models of `mask` (bit m = model m) from one pass over each seed's Brownian path; a row holds the
    /// selected models' blocks in ascending model order.
    pub fn eigs_batch_multi(&self, mask: u32, dim: usize, steps: usize, seeds: &[u32]) -> Result<Vec<f64>, String> {
        let w = unsafe { jne_multi_width(mask, dim as u32) };
        if w <= 0 {
            return Err("invalid model mask / dim".into());
        }
        let mut out = vec![0f64; seeds.len() * w as usize];
        let rc = unsafe {
            jne_eigs_batch_multi(self.0, mask, dim as u32, steps as u32, seeds.as_ptr(), seeds.len() as u64, out.as_mut_ptr())
        };
        if rc != 0 {
            return Err(last_error(self.0));
        }
        Ok(out)
    }

    /// Enqueue a batch; overlap the writer with it and call `wait` for the rows (one ticket at a time).
    pub fn submit(&self, model: u8, dim: usize, steps: usize, seeds: &[u32]) -> Result<Pending, String> {
        let mut rows = vec![0f64; seeds.len() * Self::num_eigs(model, dim)];
        let ticket = unsafe {
            jne_submit(self.0, model, dim as u32, steps as u32, seeds.as_ptr(), seeds.len() as u64, rows.as_mut_ptr())
        };
        if ticket <= 0 {
            return Err(last_error(self.0));
        }
        Ok(Pending { ticket, rows })
    }

    /// The same for the fused multi-model batch.
    pub fn submit_multi(&self, mask: u32, dim: usize, steps: usize, seeds: &[u32]) -> Result<Pending, String> {
        let w = unsafe { jne_multi_width(mask, dim as u32) };
        if w <= 0 {
            return Err("invalid model mask / dim".into());
        }
        let mut rows = vec![0f64; seeds.len() * w as usize];
        let ticket = unsafe {
            jne_submit_multi(self.0, mask, dim as u32, steps as u32, seeds.as_ptr(), seeds.len() as u64, rows.as_mut_ptr())
        };
        if ticket <= 0 {
            return Err(last_error(self.0));
        }
        Ok(Pending { ticket, rows })
    }

    pub fn wait(&self, p: Pending) -> Result<Vec<f64>, String> {
        let rc = unsafe { jne_wait(self.0, p.ticket) };
        if rc != 0 {
            return Err(last_error(self.0));
        }
        Ok(p.rows)
    }

    /// Streaming form of the fused batch -- the reference's own protocol (every finished record is SENT to a consumer,
    /// parallel_compute.rs:33-39): `on_rows(first, rows)` receives the rows of `seeds[first ..]` (a multiple of the row
    /// width, `jne_multi_width`) as soon as they are on the host, straight from the library's staging memory.  It runs
    /// on the library's per-device threads, possibly concurrently for disjoint ranges; returning `false` aborts.
    pub fn eigs_batch_multi_stream<F>(&self, mask: u32, dim: usize, steps: usize, seeds: &[u32], on_rows: F) -> Result<(), String>
    where
        F: Fn(usize, &[f64]) -> bool + Sync,
    {
        let w = unsafe { jne_multi_width(mask, dim as u32) };
        if w <= 0 {
            return Err("invalid model mask / dim".into());
        }
        struct Ctx<F> {
            f: F,
            width: usize,
        }
        extern "C" fn tramp<F: Fn(usize, &[f64]) -> bool + Sync>(user: *mut c_void, first: u64, count: u64, rows: *const f64) -> c_int {
            let cx = unsafe { &*(user as *const Ctx<F>) };
            let slice = unsafe { std::slice::from_raw_parts(rows, count as usize * cx.width) };
            // a panic must not unwind across the C ABI: it becomes an aborted batch
            match std::panic::catch_unwind(std::panic::AssertUnwindSafe(|| (cx.f)(first as usize, slice))) {
                Ok(true) => 0,
                _ => 1,
            }
        }
        let cx = Ctx { f: on_rows, width: w as usize };
        let rc = unsafe {
            jne_eigs_batch_multi_stream(self.0, mask, dim as u32, steps as u32, seeds.as_ptr(), seeds.len() as u64,
                                        tramp::<F>, &cx as *const Ctx<F> as *mut c_void)
        };
        if rc != 0 {
            return Err(last_error(self.0));
        }
        Ok(())
    }

    /// main.rs:109-114 as one fused job: every selected model's EIGENVALS_V6 file for (dim, steps, num_runs),
    /// each resumed on its own.  `filenames[m]` is used when bit m of `mask` is set.
    /// Returns per model (completed_before, computed, total_in_file).
    pub fn run_models_simulation(&self, mask: u32, dim: usize, steps: usize, num_runs: u64,
                                 filenames: &[Option<String>; 5], quiet: bool) -> Result<[[u64; 3]; 5], String> {
        let owned: Vec<Option<CString>> = filenames.iter().map(|f| f.as_ref().map(|s| CString::new(s.as_str()).unwrap())).collect();
        let ptrs: Vec<*const c_char> = owned.iter().map(|f| f.as_ref().map_or(std::ptr::null(), |c| c.as_ptr())).collect();
        let mut stats = [0u64; 15];
        let rc = unsafe {
            jne_run_models_simulation(self.0, mask, dim as u32, steps as u32, num_runs, ptrs.as_ptr(), quiet as c_int,
                                      std::ptr::null(), 0, stats.as_mut_ptr())
        };
        if rc != 0 {
            return Err(last_error(self.0));
        }
        let mut out = [[0u64; 3]; 5];
        for m in 0..5 {
            out[m].copy_from_slice(&stats[3 * m..3 * m + 3]);
        }
        Ok(out)
    }
}

impl Drop for Gpu {
    fn drop(&mut self) {
        unsafe { jne_shutdown(self.0) }
    }
}

fn last_error(ctx: *const JneCtx) -> String {
    unsafe { CStr::from_ptr(jne_last_error(ctx)).to_string_lossy().into_owned() }
}
