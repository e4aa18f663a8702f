//! Safe shim over `msgpu-sys` for the reference (argumentcomputer/multi-stark): the `TwoAdicSubgroupDft` slot
//! (`type Dft`, src/types.rs:200) and the commitment half of the `Pcs` slot (`type Pcs`, src/types.rs:85,209-223).
//!
//! SOURCE ONLY: this image has no cargo/rustc and Plonky3 is an un-vendored git dependency of the reference, so this file has
//! not been compiled. It shows the exact calls a maintainer makes; INTEGRATION.md walks through `open` (3a), stage 2 and the
//! claims (3b), the device-resident quotient hook (4) and the multi-GPU entry points (5). The working, tested caller of the
//! same ABI is the C++ driver in `multi_stark_b200/host/` (`msh_prove`).
use std::ffi::CStr;
use std::ptr::null_mut;
use std::sync::Arc;

use msgpu_sys::*;
use p3_field::PrimeField64;
use p3_goldilocks::Goldilocks;
use p3_matrix::dense::RowMajorMatrix;
use p3_matrix::Matrix;

#[derive(Debug)]
pub struct GpuError(pub String);

fn check(code: i32) -> Result<(), GpuError> {
    if code == 0 {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(msgpu_last_error()) }.to_string_lossy().into_owned();
    Err(GpuError(format!("msgpu error {code}: {msg}")))
}

/// One CUDA device + stream. `Send`, not `Sync`: one context per proving thread (single-stream arena).
pub struct Ctx {
    raw: *mut msgpu_ctx,
}
unsafe impl Send for Ctx {}
impl Ctx {
    pub fn new(device: i32) -> Result<Arc<Self>, GpuError> {
        let mut raw = null_mut();
        check(unsafe { msgpu_ctx_create(device, null_mut(), &mut raw) })?;
        Ok(Arc::new(Self { raw }))
    }
}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { msgpu_ctx_destroy(self.raw) }
    }
}

/// The ABI takes canonical u64 (< p). p3's `Goldilocks` is `repr(transparent)` over u64 but may hold non-canonical values.
fn canonical(m: &RowMajorMatrix<Goldilocks>) -> Vec<u64> {
    m.values.iter().map(|v| v.as_canonical_u64()).collect()
}

/// `TwoAdicSubgroupDft<Goldilocks>` (src/prover.rs:440,650,716): `dft_batch(m).bit_reverse_rows()` in one call.
#[derive(Clone)]
pub struct GpuDft {
    pub ctx: Arc<Ctx>,
}
impl GpuDft {
    /// Raw storage of `Radix2DitParallel::dft_batch(m)` (natural frequency k at row rev(k)); wrap it in p3's
    /// `BitReversedMatrixView` to get the trait's `Evaluations`.
    pub fn dft_batch_bitrev(&self, m: &RowMajorMatrix<Goldilocks>) -> Result<Vec<u64>, GpuError> {
        let (h, w) = (m.height() as u64, m.width() as u64);
        let input = canonical(m);
        let mut out = vec![0u64; input.len()];
        check(unsafe { msgpu_dft_batch_bitrev(self.ctx.raw, input.as_ptr(), h, w, out.as_mut_ptr()) })?;
        Ok(out)
    }
    /// `coset_lde_batch(m, added_bits, shift).bit_reverse_rows()` (src/prover.rs:681-692)
    pub fn coset_lde_batch_bitrev(&self, m: &RowMajorMatrix<Goldilocks>, added_bits: u32, shift: Goldilocks) -> Result<Vec<u64>, GpuError> {
        let (h, w) = (m.height() as u64, m.width() as u64);
        let input = canonical(m);
        let mut out = vec![0u64; input.len() << added_bits];
        check(unsafe {
            msgpu_coset_lde_batch_bitrev(self.ctx.raw, input.as_ptr(), h, w, added_bits, shift.as_canonical_u64(), out.as_mut_ptr())
        })?;
        Ok(out)
    }
}

/// `Pcs::ProverData`: the LDE matrices (bit-reversed rows) and every digest layer, resident in HBM.
pub struct GpuProverData {
    ctx: Arc<Ctx>,
    raw: *mut msgpu_pdata,
}
impl Drop for GpuProverData {
    fn drop(&mut self) {
        unsafe { msgpu_pdata_free(self.raw) }
    }
}
impl GpuProverData {
    /// `Pcs::get_evaluations_on_domain` (src/prover.rs:454-468): the first `rows * quotient_degree` stored rows of matrix `idx`,
    /// as a device pointer: no copy. Feed it to `msgpu_quotient`.
    pub fn matrix(&self, idx: u64) -> Result<(*mut u64, u64, u64), GpuError> {
        let (mut p, mut rows, mut cols) = (null_mut(), 0u64, 0u64);
        check(unsafe { msgpu_pdata_matrix(self.raw, idx, &mut p, &mut rows, &mut cols) })?;
        Ok((p, rows, cols))
    }
    /// `Mmcs::open_batch` for many indices at once: (opened rows of every matrix back to back, sibling digests bottom-up)
    pub fn open_batch(&self, indices: &[u64], total_width: usize, depth: usize) -> Result<(Vec<u64>, Vec<u8>), GpuError> {
        let mut opened = vec![0u64; indices.len() * total_width];
        let mut proofs = vec![0u8; indices.len() * depth * 32];
        check(unsafe {
            msgpu_open_batch(self.ctx.raw, self.raw, indices.as_ptr(), indices.len() as u64, opened.as_mut_ptr(), proofs.as_mut_ptr())
        })?;
        Ok((opened, proofs))
    }
}

/// The commitment half of `TwoAdicFriPcs` as configured by `new_pcs` (src/types.rs:209-223); `open` is INTEGRATION.md 3a.
pub struct GpuFriPcs {
    pub ctx: Arc<Ctx>,
    pub log_blowup: u32,
}
impl GpuFriPcs {
    /// `Pcs::commit` (src/prover.rs:350,419; src/system.rs:193): the root is p3's `Hash<Goldilocks, u8, 32>` (cap_height 0).
    pub fn commit(&self, evals: &[RowMajorMatrix<Goldilocks>]) -> Result<([u8; 32], GpuProverData), GpuError> {
        let flat: Vec<Vec<u64>> = evals.iter().map(canonical).collect();
        let ptrs: Vec<*const u64> = flat.iter().map(|v| v.as_ptr()).collect();
        let heights: Vec<u64> = evals.iter().map(|m| m.height() as u64).collect();
        let widths: Vec<u64> = evals.iter().map(|m| m.width() as u64).collect();
        let mut raw = null_mut();
        let mut root = [0u8; 32];
        check(unsafe {
            msgpu_commit(self.ctx.raw, ptrs.as_ptr(), heights.as_ptr(), widths.as_ptr(), evals.len() as u64, self.log_blowup, &mut raw,
                         root.as_mut_ptr())
        })?;
        Ok((root, GpuProverData { ctx: self.ctx.clone(), raw }))
    }
    /// `Pcs::commit_ldes` (src/prover.rs:526) over LDEs that `msgpu_quotient` left on the device; the prover data adopts them.
    pub fn commit_ldes(&self, ldes: &[(*mut u64, u64, u64)]) -> Result<([u8; 32], GpuProverData), GpuError> {
        let ptrs: Vec<*mut u64> = ldes.iter().map(|l| l.0).collect();
        let heights: Vec<u64> = ldes.iter().map(|l| l.1).collect();
        let widths: Vec<u64> = ldes.iter().map(|l| l.2).collect();
        let mut raw = null_mut();
        let mut root = [0u8; 32];
        check(unsafe {
            msgpu_commit_ldes_dev(self.ctx.raw, ptrs.as_ptr(), heights.as_ptr(), widths.as_ptr(), ldes.len() as u64, 1, &mut raw,
                                  root.as_mut_ptr())
        })?;
        Ok((root, GpuProverData { ctx: self.ctx.clone(), raw }))
    }
}
