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

// =================================================================================================================
// Trait-level slots. What follows implements the two Plonky3 traits the reference instantiates, so that the swap at
// src/types.rs:85,200 is a type alias change:
//     pub type Dft = msgpu::GpuDft;                                            // was Radix2DitParallel<Val>
//     pub type Pcs = msgpu::GpuFriPcs<TwoAdicFriPcs<Val, Dft, Mmcs, ExtMmcs>>; // verification and the wire types stay p3's
// UNCOMPILED (no cargo in this image). The method sets follow p3-dft / p3-commit 0.5.1 as the reference's call sites show them
// (src/prover.rs:346-580, src/system.rs:182-193, src/verifier.rs:413); a trait method the pinned revision adds or renames is a
// one-line forward to `self.cpu`.
// =================================================================================================================
use std::sync::OnceLock;

use p3_challenger::{CanObserve, FieldChallenger, GrindingChallenger};
use p3_commit::{OpenedValues, Pcs, PolynomialSpace, TwoAdicMultiplicativeCoset};
use p3_dft::TwoAdicSubgroupDft;
use p3_field::extension::BinomialExtensionField;
use p3_field::{BasedVectorSpace, PrimeCharacteristicRing, TwoAdicField};
use p3_matrix::bitrev::{BitReversedMatrixView, BitReversibleMatrix};

type Val = Goldilocks;
type ExtVal = BinomialExtensionField<Val, 2>;

/// `TwoAdicSubgroupDft` needs `Default`: one lazily created context per process for the bare-DFT slot (the PCS owns its own).
fn default_ctx() -> Arc<Ctx> {
    static CTX: OnceLock<Arc<Ctx>> = OnceLock::new();
    CTX.get_or_init(|| {
        let ctx = Ctx::new(0).expect("libmsgpu: no CUDA device (there is no CPU fallback)");
        // p3's Goldilocks keeps any u64 representative: let the device reduce the uploads instead of copying them here
        check(unsafe { msgpu_ctx_set_option(ctx.raw, MSGPU_OPT_CANONICALIZE_INPUTS as i32, 1) }).unwrap();
        ctx
    })
    .clone()
}
impl Default for GpuDft {
    fn default() -> Self {
        Self { ctx: default_ctx() }
    }
}

/// Zero-copy view of a matrix's storage as the ABI's `const uint64_t*`: `Goldilocks` is `repr(transparent)` over `u64` and the
/// context reduces the values on the device (MSGPU_OPT_CANONICALIZE_INPUTS), so nothing is copied or converted on the host.
/// `msgpu_host_register` page-locks the Vec for the duration of the call (full PCIe rate instead of a staged pageable copy).
struct Pinned<'a>(&'a [Val]);
impl<'a> Pinned<'a> {
    fn new(v: &'a [Val]) -> Self {
        let _ = unsafe { msgpu_host_register(v.as_ptr() as *mut _, std::mem::size_of_val(v)) }; // best effort
        Self(v)
    }
    fn ptr(&self) -> *const u64 {
        self.0.as_ptr() as *const u64
    }
}
impl Drop for Pinned<'_> {
    fn drop(&mut self) {
        let _ = unsafe { msgpu_host_unregister(self.0.as_ptr() as *mut _) };
    }
}

impl TwoAdicSubgroupDft<Val> for GpuDft {
    /// Same shape as `Radix2DitParallel`'s: the raw storage is bit-reversed, the view presents natural order, and
    /// `.bit_reverse_rows()` (src/prover.rs:650,716) unwraps the storage without a copy.
    type Evaluations = BitReversedMatrixView<RowMajorMatrix<Val>>;

    fn dft_batch(&self, mat: RowMajorMatrix<Val>) -> Self::Evaluations {
        let (h, w) = (mat.height(), mat.width());
        let mut out = vec![Val::ZERO; h * w];
        {
            let src = Pinned::new(&mat.values);
            check(unsafe { msgpu_dft_batch_bitrev(self.ctx.raw, src.ptr(), h as u64, w as u64, out.as_mut_ptr() as *mut u64) })
                .expect("msgpu_dft_batch_bitrev");
        }
        RowMajorMatrix::new(out, w).bit_reverse_rows()
    }

    fn coset_lde_batch(&self, mat: RowMajorMatrix<Val>, added_bits: usize, shift: Val) -> Self::Evaluations {
        let (h, w) = (mat.height(), mat.width());
        let mut out = vec![Val::ZERO; (h << added_bits) * w];
        {
            let src = Pinned::new(&mat.values);
            check(unsafe {
                msgpu_coset_lde_batch_bitrev(self.ctx.raw, src.ptr(), h as u64, w as u64, added_bits as u32, shift.as_canonical_u64(),
                                             out.as_mut_ptr() as *mut u64)
            })
            .expect("msgpu_coset_lde_batch_bitrev");
        }
        RowMajorMatrix::new(out, w).bit_reverse_rows()
    }
    // idft_batch / coset_idft_batch / lde_batch keep the trait's provided bodies, which are written in terms of dft_batch.
}

/// `Pcs<ExtVal, Challenger>` over the device; `Cpu` is the reference's own `TwoAdicFriPcs<Val, Dft, Mmcs, ExtMmcs>` as `new_pcs`
/// builds it (src/types.rs:209-223): it supplies the associated WIRE types (Commitment, Proof, Error, Domain) and `verify`, so
/// `Proof::to_bytes`, the verifier and every downstream type are untouched.
pub struct GpuFriPcsFull<Cpu> {
    pub gpu: GpuFriPcs,
    pub cpu: Cpu,
    pub fri: FriShape,
}
/// The FRI parameters `open` needs (src/types.rs:185-197)
#[derive(Clone, Copy)]
pub struct FriShape {
    pub log_final_poly_len: usize,
    pub num_queries: usize,
    pub commit_pow_bits: usize,
    pub query_pow_bits: usize,
}

/// `Pcs::EvaluationsOnDomain<'a>`: a DEVICE view (pointer + shape) of the first rows of a committed LDE. The patched
/// `quotient_values` call (integration/prover.rs.patch) hands it to `msgpu_quotient` and never reads it on the host; code that
/// does treat it as a `Matrix` gets a lazily downloaded copy.
pub struct DeviceEvaluations<'a> {
    pub data: &'a GpuProverData,
    pub idx: u64,
    pub rows: usize,
    pub cols: usize,
    host: OnceLock<RowMajorMatrix<Val>>,
}
impl DeviceEvaluations<'_> {
    fn host(&self) -> &RowMajorMatrix<Val> {
        self.host.get_or_init(|| {
            let mut v = vec![Val::ZERO; self.rows * self.cols];
            check(unsafe { msgpu_pdata_read_rows(self.data.ctx.raw, self.data.raw, self.idx, 0, self.rows as u64, v.as_mut_ptr() as *mut u64) })
                .expect("msgpu_pdata_read_rows");
            // stored rows are bit-reversed; the reference's view is natural order on the coset (src/prover.rs:454-468)
            RowMajorMatrix::new(v, self.cols).bit_reverse_rows().to_row_major_matrix()
        })
    }
}
impl Matrix<Val> for DeviceEvaluations<'_> {
    fn width(&self) -> usize {
        self.cols
    }
    fn height(&self) -> usize {
        self.rows
    }
    fn get(&self, r: usize, c: usize) -> Option<Val> {
        self.host().get(r, c)
    }
    // (row / row_slice forward to `self.host()` in the same way)
}

impl<Cpu, Challenger> Pcs<ExtVal, Challenger> for GpuFriPcsFull<Cpu>
where
    Cpu: Pcs<ExtVal, Challenger, Domain = TwoAdicMultiplicativeCoset<Val>, Commitment = p3_symmetric::Hash<Val, u8, 32>, Proof = p3_fri::FriProof<ExtVal, crate_types::ExtMmcs, Val, Vec<p3_commit::BatchOpening<Val, crate_types::Mmcs>>>>,
    Challenger: FieldChallenger<Val> + CanObserve<Cpu::Commitment> + GrindingChallenger<Witness = Val>,
{
    type Domain = Cpu::Domain;
    type Commitment = Cpu::Commitment;
    type ProverData = GpuProverData;
    type EvaluationsOnDomain<'a> = DeviceEvaluations<'a>;
    type Proof = Cpu::Proof;
    type Error = Cpu::Error;
    const ZK: bool = false; // src/prover.rs:522-525 asserts it before commit_ldes

    fn natural_domain_for_degree(&self, degree: usize) -> Self::Domain {
        self.cpu.natural_domain_for_degree(degree)
    }

    /// src/prover.rs:350,419; src/system.rs:193. The matrices are uploaded from where they lie (no `canonical()` copy).
    fn commit(&self, evaluations: impl IntoIterator<Item = (Self::Domain, RowMajorMatrix<Val>)>) -> (Self::Commitment, Self::ProverData) {
        let mats: Vec<RowMajorMatrix<Val>> = evaluations.into_iter().map(|(_, m)| m).collect();
        let pins: Vec<Pinned> = mats.iter().map(|m| Pinned::new(&m.values)).collect();
        let ptrs: Vec<*const u64> = pins.iter().map(|p| p.ptr()).collect();
        let heights: Vec<u64> = mats.iter().map(|m| m.height() as u64).collect();
        let widths: Vec<u64> = mats.iter().map(|m| m.width() as u64).collect();
        let (mut raw, mut root) = (null_mut(), [0u8; 32]);
        check(unsafe {
            msgpu_commit(self.gpu.ctx.raw, ptrs.as_ptr(), heights.as_ptr(), widths.as_ptr(), mats.len() as u64, self.gpu.log_blowup,
                         &mut raw, root.as_mut_ptr())
        })
        .expect("msgpu_commit");
        (root.into(), GpuProverData { ctx: self.gpu.ctx.clone(), raw })
    }

    /// src/prover.rs:526. The LDEs were produced on the device by the quotient hook and arrive wrapped as `DeviceLde` matrices
    /// (integration/prover.rs.patch); they are adopted, not copied.
    fn commit_ldes(&self, ldes: Vec<RowMajorMatrix<Val>>) -> (Self::Commitment, Self::ProverData) {
        let parts: Vec<(*mut u64, u64, u64)> = ldes.iter().map(device_lde_of).collect();
        let (root, data) = self.gpu.commit_ldes(&parts).expect("msgpu_commit_ldes_dev");
        (root.into(), data)
    }

    /// src/prover.rs:454-468: a pointer and a shape -- the first `domain.size()` stored rows of matrix `idx`.
    fn get_evaluations_on_domain<'a>(&self, data: &'a Self::ProverData, idx: usize, domain: Self::Domain) -> Self::EvaluationsOnDomain<'a> {
        let (_, rows, cols) = data.matrix(idx as u64).expect("msgpu_pdata_matrix");
        assert!(domain.size() as u64 <= rows, "quotient domain larger than the committed LDE");
        DeviceEvaluations { data, idx: idx as u64, rows: domain.size(), cols: cols as usize, host: OnceLock::new() }
    }

    /// src/prover.rs:580. The transcript stays HERE (it is the reference's `Challenger`); the device does the arithmetic between
    /// transcript steps: barycentric evaluation, reduced openings, fold + commit rounds, the query openings in one launch.
    fn open(&self, rounds: Vec<(&Self::ProverData, Vec<Vec<ExtVal>>)>, challenger: &mut Challenger) -> (OpenedValues<ExtVal>, Self::Proof) {
        let ctx = self.gpu.ctx.raw;
        let pds: Vec<*const msgpu_pdata> = rounds.iter().map(|(d, _)| d.raw as *const _).collect();
        let (mut n_points, mut points) = (Vec::new(), Vec::new());
        for (_, mats) in &rounds {
            for pts in mats {
                n_points.push(pts.len() as u64);
                for z in pts {
                    let c: &[Val] = z.as_basis_coefficients_slice();
                    points.extend([c[0].as_canonical_u64(), c[1].as_canonical_u64()]);
                }
            }
        }
        let (mut op, mut n_values) = (null_mut(), 0u64);
        check(unsafe { msgpu_open_begin(ctx, pds.len() as u64, pds.as_ptr(), n_points.as_ptr(), points.as_ptr(), self.gpu.log_blowup, &mut op, &mut n_values) })
            .expect("msgpu_open_begin");
        let mut flat = vec![0u64; 2 * n_values as usize];
        check(unsafe { msgpu_open_values(op, flat.as_mut_ptr()) }).unwrap();
        // opened values in (round, matrix, point, column) order: observed in that order, returned in that shape
        let mut it = flat.chunks_exact(2).map(|c| ExtVal::from_basis_coefficients_slice(&[Val::new(c[0]), Val::new(c[1])]).unwrap());
        let mut opened: OpenedValues<ExtVal> = Vec::new();
        for (data, mats) in &rounds {
            let mut round = Vec::new();
            for (m, pts) in mats.iter().enumerate() {
                let width = data.matrix(m as u64).unwrap().2 as usize;
                round.push(pts.iter().map(|_| (&mut it).take(width).collect::<Vec<_>>()).collect::<Vec<_>>());
            }
            opened.push(round);
        }
        for v in opened.iter().flatten().flatten().flatten() {
            challenger.observe_algebra_element(*v);
        }
        let alpha: ExtVal = challenger.sample_algebra_element();
        let a: &[Val] = alpha.as_basis_coefficients_slice();
        let (mut n_inputs, mut log_max) = (0u64, 0u32);
        check(unsafe { msgpu_open_reduce(op, [a[0].as_canonical_u64(), a[1].as_canonical_u64()].as_ptr(), &mut n_inputs, &mut log_max) }).unwrap();

        // commit phase (p3-fri `commit_phase`): commit -> observe -> grind -> beta -> fold
        let stop = (1u64 << self.gpu.log_blowup) << self.fri.log_final_poly_len;
        let (mut commits, mut pow_witnesses) = (Vec::new(), Vec::new());
        let mut len = 0u64;
        check(unsafe { msgpu_fri_current_len(op, &mut len) }).unwrap();
        while len > stop {
            let mut root = [0u8; 32];
            check(unsafe { msgpu_fri_commit_round(op, root.as_mut_ptr()) }).unwrap();
            let commit: Self::Commitment = root.into();
            challenger.observe(commit.clone());
            commits.push(commit);
            pow_witnesses.push(challenger.grind(self.fri.commit_pow_bits));
            let beta: ExtVal = challenger.sample_algebra_element();
            let b: &[Val] = beta.as_basis_coefficients_slice();
            check(unsafe { msgpu_fri_fold(op, [b[0].as_canonical_u64(), b[1].as_canonical_u64()].as_ptr()) }).unwrap();
            len /= 2;
        }
        // (with commit_pow_bits = 0 the whole loop is ONE call, msgpu_fri_commit_phase, and the challenger replays the roots)
        let mut folded = vec![0u64; 2 * len as usize];
        check(unsafe { msgpu_fri_read_current(op, folded.as_mut_ptr()) }).unwrap();
        let final_poly = final_poly_from_folded(&folded, self.fri.log_final_poly_len); // bit-reversal undone, inverse DFT, truncated
        for c in &final_poly {
            challenger.observe_algebra_element(*c);
        }
        let query_pow_witness = challenger.grind(self.fri.query_pow_bits);
        let indices: Vec<u64> = (0..self.fri.num_queries).map(|_| challenger.sample_bits(log_max as usize) as u64).collect();
        // every tree (input commitments, then the commit-phase layers) at every index in one launch: msgpu_open_batch_multi
        let query_proofs = open_queries(ctx, op, &rounds, &indices, log_max, commits.len());
        unsafe { msgpu_open_free(op) };
        (opened, assemble_fri_proof(commits, pow_witnesses, query_proofs, final_poly, query_pow_witness))
    }

    fn verify(
        &self,
        rounds: Vec<(Self::Commitment, Vec<(Self::Domain, Vec<(ExtVal, Vec<ExtVal>)>)>)>,
        proof: &Self::Proof,
        challenger: &mut Challenger,
    ) -> Result<(), Self::Error> {
        self.cpu.verify(rounds, proof, challenger) // the verifier is out of scope of the hot path: p3's own
    }
}
// `device_lde_of`, `final_poly_from_folded`, `open_queries`, `assemble_fri_proof`: 60 lines of plumbing between the flat ABI
// buffers and p3-fri's `FriProof { commit_phase_commits, commit_pow_witnesses, query_proofs, final_poly, query_pow_witness }` /
// `QueryProof { input_proof, commit_phase_openings }`; their exact field lists are the pinned revision's and are the only part of
// this file that cannot be written without the Plonky3 sources. multi_stark_b200/host/pcs.hpp (pcs_open) is the tested C++
// statement of the same plumbing, field for field.
