//! Safe wrapper over libs2c_b200.so for the Rust side of reclaimprotocol/zk-symmetric-crypto.
//!
//! Two seams:
//!  * product level -- `Ctx::prove_chacha20` / `prove_aes_ctr` return the bincode bytes of `StreamProof` / `AESCtrProof`
//!    (stwo/src/chacha/bitwise/air_stream.rs:30-131, stwo/src/aes/lookup/air_ctr.rs:44-184), byte-identical to
//!    `prove_stream_with_inputs::<Blake2sMerkleChannel>`: replace the body of `generate_chacha20_proof` (wasm_api.rs:577) with
//!    one call and deserialise with the existing serde derives;
//!  * backend-trait level -- `ffi::cb_*` map one-to-one onto the methods of upstream stwo's `Backend` traits; `DeviceColumn` is the
//!    `Col<CudaBackend, BaseField>` a `CudaBackend` would use (sketch at the bottom of this file).
//! NOT BUILT IN THIS REPOSITORY (no cargo in the build image).
pub mod ffi;

use std::ffi::CStr;
use std::ptr;

#[derive(Debug)]
pub struct Error(pub String);
pub type Result<T> = std::result::Result<T, Error>;

/// One backend context = one (GPU, stream).  Not `Sync`: a context must not be used from two threads at once.
pub struct Ctx {
    raw: *mut ffi::cb_ctx,
}
unsafe impl Send for Ctx {}

impl Ctx {
    pub fn new(device: i32) -> Result<Self> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { ffi::cb_init(device, &mut raw) };
        if rc != 0 {
            return Err(Error(format!("cb_init(device={device}) failed with status {rc}: no usable CUDA device (no CPU fallback)")));
        }
        Ok(Ctx { raw })
    }

    fn check(&self, rc: i32) -> Result<()> {
        if rc == 0 {
            return Ok(());
        }
        let msg = unsafe { CStr::from_ptr(ffi::cb_last_error(self.raw)) }.to_string_lossy().into_owned();
        Err(Error(msg))
    }

    fn take(&self, p: *mut u8, n: usize) -> Vec<u8> {
        let v = unsafe { std::slice::from_raw_parts(p, n) }.to_vec();
        unsafe { ffi::s2c_free(p as *mut _) };
        v
    }

    /// bincode(StreamProof) for `plaintext.len() / 64` ChaCha20 blocks starting at `counter`; error strings are the reference's.
    pub fn prove_chacha20(&self, key: &[u8; 32], nonce: &[u8; 12], counter: u32, plaintext: &[u8], ciphertext: &[u8]) -> Result<Vec<u8>> {
        if plaintext.len() != ciphertext.len() {
            return Err(Error(format!("Ciphertext must be same length as plaintext, got {} vs {}", ciphertext.len(), plaintext.len())));
        }
        let (mut p, mut n) = (ptr::null_mut(), 0usize);
        self.check(unsafe {
            ffi::s2c_prove_chacha20_raw(self.raw, key.as_ptr(), nonce.as_ptr(), counter, plaintext.as_ptr(), ciphertext.as_ptr(),
                                        plaintext.len(), &mut p, &mut n)
        })?;
        Ok(self.take(p, n))
    }

    /// bincode(AESCtrProof); key.len() = 16 or 32.
    pub fn prove_aes_ctr(&self, key: &[u8], nonce: &[u8; 12], counter: u32, plaintext: &[u8], ciphertext: &[u8]) -> Result<Vec<u8>> {
        if key.len() != 16 && key.len() != 32 {
            return Err(Error(format!("Key must be 16 or 32 bytes, got {}", key.len())));
        }
        if plaintext.len() != ciphertext.len() {
            return Err(Error(format!("Ciphertext must be same length as plaintext, got {} vs {}", ciphertext.len(), plaintext.len())));
        }
        let (mut p, mut n) = (ptr::null_mut(), 0usize);
        self.check(unsafe {
            ffi::s2c_prove_aes_ctr_raw(self.raw, key.len() as i32, key.as_ptr(), nonce.as_ptr(), counter, plaintext.as_ptr(),
                                       ciphertext.as_ptr(), plaintext.len(), &mut p, &mut n)
        })?;
        Ok(self.take(p, n))
    }

    /// The reference's `prove_stream::<Blake2sMerkleChannel>(log_size, PcsConfig::default())` (air_stream.rs:237-289).
    pub fn prove_stream_testdata(&self, log_size: u32) -> Result<Vec<u8>> {
        let (mut p, mut n) = (ptr::null_mut(), 0usize);
        self.check(unsafe { ffi::s2c_prove_chacha20_stream_testdata(self.raw, log_size as i32, &mut p, &mut n) })?;
        Ok(self.take(p, n))
    }

    /// `prove_bitwise::<Blake2sMerkleChannel>(log_size, PcsConfig::default())` (chacha/bitwise/air.rs:53): returns
    /// `u32 log_size || bincode(stark_proof)`; compare `stark_proof` with `bincode::serialize(&proof.stark_proof)` of the reference.
    pub fn prove_bitwise(&self, log_size: u32) -> Result<Vec<u8>> {
        let (mut p, mut n) = (ptr::null_mut(), 0usize);
        self.check(unsafe { ffi::s2c_prove_chacha20_block(self.raw, log_size as i32, &mut p, &mut n) })?;
        Ok(self.take(p, n))
    }

    pub fn raw(&self) -> *mut ffi::cb_ctx {
        self.raw
    }
}

impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { ffi::cb_destroy(self.raw) }
    }
}

/// Host-side verification (no GPU needed): `Ok(())` or the reference's `{:?}` rendering of its `VerificationError`.
pub fn verify_chacha20(proof: &[u8], nonce: &[u8; 12], counter: u32, plaintext: &[u8], ciphertext: &[u8]) -> std::result::Result<(), String> {
    let mut err = ptr::null_mut();
    let rc = unsafe {
        ffi::s2c_verify_chacha20_raw(proof.as_ptr(), proof.len(), nonce.as_ptr(), counter, plaintext.as_ptr(), plaintext.len(),
                                     ciphertext.as_ptr(), ciphertext.len(), &mut err)
    };
    let msg = if err.is_null() { None } else {
        let s = unsafe { CStr::from_ptr(err) }.to_string_lossy().into_owned();
        unsafe { ffi::s2c_free(err as *mut _) };
        Some(s)
    };
    if rc == 0 { Ok(()) } else { Err(msg.unwrap_or_else(|| format!("status {rc}"))) }
}

/// A device column of 2^log_size M31 words: what `Col<CudaBackend, BaseField>` wraps.
pub struct DeviceColumn<'a> {
    ctx: &'a Ctx,
    pub ptr: *mut u32,
    pub log_size: u32,
}

impl<'a> DeviceColumn<'a> {
    pub fn from_host(ctx: &'a Ctx, values: &[u32]) -> Result<Self> {
        assert!(values.len().is_power_of_two());
        let mut p = ptr::null_mut();
        ctx.check(unsafe { ffi::cb_malloc(ctx.raw, values.len() * 4, &mut p) })?;
        ctx.check(unsafe { ffi::cb_h2d(ctx.raw, p, values.as_ptr() as *const _, values.len() * 4) })?;
        Ok(DeviceColumn { ctx, ptr: p as *mut u32, log_size: values.len().trailing_zeros() })
    }
    pub fn to_cpu(&self) -> Result<Vec<u32>> {
        let mut v = vec![0u32; 1 << self.log_size];
        self.ctx.check(unsafe { ffi::cb_d2h(self.ctx.raw, v.as_mut_ptr() as *mut _, self.ptr as *const _, v.len() * 4) })?;
        Ok(v)
    }
    /// ColumnOps::bit_reverse_column
    pub fn bit_reverse(&mut self) -> Result<()> {
        self.ctx.check(unsafe { ffi::cb_bit_reverse(self.ctx.raw, self.ptr, self.log_size as i32) })
    }
    /// PolyOps::interpolate (in place: evaluations in bit-reversed circle-domain order -> coefficients)
    pub fn interpolate(&mut self) -> Result<()> {
        self.ctx.check(unsafe { ffi::cb_interpolate_columns(self.ctx.raw, self.ptr, 1 << self.log_size, 1, self.log_size as i32) })
    }
}

impl Drop for DeviceColumn<'_> {
    fn drop(&mut self) {
        unsafe { ffi::cb_free(self.ctx.raw, self.ptr as *mut _) };
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Backend-trait sketch (compiles only inside a crate that depends on stwo at rev f117d487 with the `prover` feature; kept as a
// comment because upstream's trait signatures are not vendored in the reference repository):
//
//   #[derive(Copy, Clone, Debug, Default)] pub struct CudaBackend;
//   impl Backend for CudaBackend {}
//   impl ColumnOps<BaseField> for CudaBackend   { type Column = CudaColumn;  fn bit_reverse_column(c) -> cb_bit_reverse }
//   impl FieldOps<BaseField> for CudaBackend    { fn batch_inverse(src, dst)  -> cb_batch_inverse_m31 }
//   impl FieldOps<SecureField> for CudaBackend  { fn batch_inverse(src, dst)  -> cb_batch_inverse_qm31 }
//   impl PolyOps for CudaBackend                { precompute_twiddles -> cb_precompute_twiddles(_coset); interpolate(_columns) ->
//                                                 cb_interpolate_columns; evaluate(_polynomials) -> cb_evaluate_polynomials;
//                                                 extend -> cb_extend; eval_at_point -> cb_eval_at_point;
//                                                 barycentric_weights / barycentric_eval_at_point -> cb_barycentric_* }
//   impl MerkleOpsLifted<Blake2sMerkleHasher>   { build_leaves -> cb_merkle_build_leaves; build_next_layer -> cb_merkle_next_layer }
//   impl MerkleOps<Blake2sMerkleHasher>         { commit_on_layer -> cb_commit_on_layer }
//   impl AccumulationOps for CudaBackend        { accumulate -> cb_accumulate; generate_secure_powers ->
//                                                 cb_generate_secure_powers_rev; lift_and_accumulate -> cb_lift_and_accumulate }
//   impl QuotientOps for CudaBackend            { accumulate_quotients -> cb_accumulate_quotients_batches }
//   impl FriOps for CudaBackend                 { fold_circle_into_line / fold_line -> cb_fold_* }
//   impl GrindOps<Blake2sChannel> for CudaBackend { grind -> cb_grind_blake2s (lowest valid nonce, like SimdBackend) }
//   impl ComponentProver<CudaBackend> for FrameworkComponent<ChaChaStreamEval> { evaluate_constraint_quotients_on_domain ->
//                                                 cb_eval_constraints_chacha_stream }
//   impl ComponentProver<CudaBackend> for FrameworkComponent<AESCtrEval>       { -> cb_eval_constraints_aes_ctr }
//   impl ComponentProver<CudaBackend> for FrameworkComponent<SboxTableEval>    { -> cb_eval_constraints_sbox_table }
//   trace generators: generate_stream_trace -> cb_gen_trace_chacha_stream; generate_aes*_ctr_trace_with_inputs ->
//   cb_gen_trace_aes_ctr; generate_ctr_sbox_interaction_trace -> cb_gen_logup_interaction_aes_ctr.
