//! Raw declarations of the C ABI in include/s2c_b200.h (one-to-one; see the header for the contract of every call).
//! All `*mut u32` / `*const u32` column arguments are DEVICE pointers unless the name ends in `_host`.
#![allow(non_camel_case_types)]
use libc::{c_char, c_int, c_void, size_t};

#[repr(C)]
pub struct cb_ctx {
    _private: [u8; 0],
}

extern "C" {
    // ---- lifecycle
    pub fn cb_init(device: c_int, out: *mut *mut cb_ctx) -> c_int;
    pub fn cb_destroy(ctx: *mut cb_ctx);
    pub fn cb_last_error(ctx: *mut cb_ctx) -> *const c_char;
    pub fn cb_set_stream(ctx: *mut cb_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn cb_sync(ctx: *mut cb_ctx) -> c_int;
    pub fn cb_launch_count(ctx: *mut cb_ctx) -> u64;
    // ---- ColumnOps / FieldOps
    pub fn cb_malloc(ctx: *mut cb_ctx, bytes: size_t, dptr: *mut *mut c_void) -> c_int;
    pub fn cb_free(ctx: *mut cb_ctx, dptr: *mut c_void) -> c_int;
    pub fn cb_h2d(ctx: *mut cb_ctx, dst_dev: *mut c_void, src_host: *const c_void, bytes: size_t) -> c_int;
    pub fn cb_d2h(ctx: *mut cb_ctx, dst_host: *mut c_void, src_dev: *const c_void, bytes: size_t) -> c_int;
    pub fn cb_memset_zero(ctx: *mut cb_ctx, dptr: *mut c_void, bytes: size_t) -> c_int;
    pub fn cb_bit_reverse(ctx: *mut cb_ctx, col: *mut u32, log_size: c_int) -> c_int;
    pub fn cb_col_at(ctx: *mut cb_ctx, col: *const u32, index: size_t, value_out_host: *mut u32) -> c_int;
    pub fn cb_col_set(ctx: *mut cb_ctx, col: *mut u32, index: size_t, value: u32) -> c_int;
    pub fn cb_batch_inverse_m31(ctx: *mut cb_ctx, src: *const u32, dst: *mut u32, n: size_t) -> c_int;
    pub fn cb_batch_inverse_qm31(ctx: *mut cb_ctx, src: *const u32, src_stride: size_t, dst: *mut u32, dst_stride: size_t, n: size_t) -> c_int;
    // ---- PolyOps
    pub fn cb_precompute_twiddles(ctx: *mut cb_ctx, max_log: c_int) -> c_int;
    pub fn cb_precompute_twiddles_coset(ctx: *mut cb_ctx, coset_initial_index: u32, coset_log_size: c_int, twiddles_out: *mut u32,
                                        itwiddles_out: *mut u32) -> c_int;
    pub fn cb_interpolate_columns(ctx: *mut cb_ctx, cols: *mut u32, stride: size_t, n_cols: c_int, log_size: c_int) -> c_int;
    pub fn cb_evaluate_polynomials(ctx: *mut cb_ctx, coeffs: *const u32, stride: size_t, n_cols: c_int, log_size: c_int, log_ext: c_int,
                                   evals: *mut u32, eval_stride: size_t) -> c_int;
    pub fn cb_extend(ctx: *mut cb_ctx, coeffs: *const u32, stride: size_t, n_cols: c_int, log_size: c_int, log_ext: c_int, out: *mut u32,
                     out_stride: size_t) -> c_int;
    pub fn cb_commit_lde(ctx: *mut cb_ctx, src_kind: c_int, src: *const u32, src_stride: size_t, first_col: u32, n_cols: c_int,
                         log_size: c_int, log_ext: c_int, coeffs_out: *mut u32, coeff_stride: size_t, lde_out: *mut u32,
                         lde_stride: size_t) -> c_int;
    pub fn cb_lde_packed(ctx: *mut cb_ctx, src_kind: c_int, src_words: *const u32, n_words: c_int, log_size: c_int, tiles_out: *mut u32) -> c_int;
    pub fn cb_eval_at_point(ctx: *mut cb_ctx, coeffs: *const u32, stride: size_t, n_cols: c_int, log_size: c_int, point_host: *const u32,
                            out_host: *mut u32) -> c_int;
    pub fn cb_barycentric_weights(ctx: *mut cb_ctx, log_size: c_int, point_host: *const u32, weights_out: *mut u32) -> c_int;
    pub fn cb_barycentric_eval_at_point(ctx: *mut cb_ctx, evals: *const u32, stride: size_t, n_cols: c_int, log_size: c_int,
                                        weights: *const u32, out_host: *mut u32) -> c_int;
    // ---- MerkleOps
    pub fn cb_merkle_build_leaves(ctx: *mut cb_ctx, group_base: *const *const u32, group_stride: *const size_t, group_ncols: *const c_int,
                                  group_log_size: *const c_int, n_groups: c_int, lifting_log: c_int, hashes_out: *mut u32) -> c_int;
    pub fn cb_merkle_leaves_absorb(ctx: *mut cb_ctx, cols: *const u32, stride: size_t, n_cols: c_int, log_size: c_int, lifting_log: c_int,
                                   state: *mut u32, bytes_before: u64, is_first: c_int, is_final: c_int, hashes_out: *mut u32) -> c_int;
    pub fn cb_merkle_next_layer(ctx: *mut cb_ctx, prev_hashes: *const u32, n_parents: u32, out_hashes: *mut u32) -> c_int;
    pub fn cb_commit_on_layer(ctx: *mut cb_ctx, log_size: c_int, prev_or_null: *const u32, cols_host: *const *const u32, n_cols: c_int,
                              out: *mut u32) -> c_int;
    // ---- ComponentProver / AccumulationOps
    pub fn cb_generate_secure_powers_rev(ctx: *mut cb_ctx, alpha_host: *const u32, n: c_int, out_dev: *mut u32) -> c_int;
    pub fn cb_eval_constraints_chacha_stream(ctx: *mut cb_ctx, lde: *const u32, stride: size_t, eval_log: c_int, trace_log: c_int,
                                             alpha_pows_rev: *const u32, accum: *mut u32, accum_stride: size_t, accumulate: c_int) -> c_int;
    pub fn cb_accumulate(ctx: *mut cb_ctx, dst: *mut u32, src: *const u32, n_words: size_t) -> c_int;
    pub fn cb_lift_and_accumulate(ctx: *mut cb_ctx, big: *mut u32, big_stride: size_t, big_log: c_int, small_cols: *const u32,
                                  small_log: c_int) -> c_int;
    // ---- AES-CTR AIR stages
    pub fn cb_aes_ctr_layout(key_len: c_int, n_cols: *mut c_int, n_constraints: *mut c_int, n_lookups: *mut c_int,
                             lookup_in_cols: *mut c_int, lookup_out_cols: *mut c_int) -> c_int;
    pub fn cb_gen_trace_aes_ctr(ctx: *mut cb_ctx, key_len: c_int, key: *const u8, nonce: *const u8, counter: u32, pt_host: *const u8,
                                ct_host: *const u8, n_blocks: u32, log_size: c_int, trace_out: *mut u32, stride: size_t,
                                mults_out_host: *mut u32, valid: *mut c_int) -> c_int;
    pub fn cb_gen_logup_interaction_aes_ctr(ctx: *mut cb_ctx, key_len: c_int, trace: *const u32, stride: size_t, log_size: c_int,
                                            z_host: *const u32, alpha_host: *const u32, inter_out: *mut u32, inter_stride: size_t,
                                            claimed_sum_out_host: *mut u32) -> c_int;
    pub fn cb_logup_finalize_last(ctx: *mut cb_ctx, col4: *mut u32, stride: size_t, log_size: c_int, claimed_sum_out_host: *mut u32) -> c_int;
    pub fn cb_eval_constraints_aes_ctr(ctx: *mut cb_ctx, key_len: c_int, lde: *const u32, stride: size_t, inter_lde: *const u32,
                                       inter_stride: size_t, trace_log: c_int, alpha_pows_rev: *const u32, z_host: *const u32,
                                       alpha_host: *const u32, claimed_sum_host: *const u32, accum: *mut u32, accum_stride: size_t) -> c_int;
    pub fn cb_eval_constraints_sbox_table(ctx: *mut cb_ctx, pre_in_lde: *const u32, pre_out_lde: *const u32, mult_lde: *const u32,
                                          inter_lde: *const u32, inter_stride: size_t, z_host: *const u32, alpha_host: *const u32,
                                          claimed_sum_host: *const u32, alpha_pow_host: *const u32, accum: *mut u32) -> c_int;
    // ---- QuotientOps / FriOps / GrindOps / gather / trace generation
    pub fn cb_accumulate_quotients(ctx: *mut cb_ctx, cols: *const u32, stride: size_t, n_cols: c_int, domain_log: c_int,
                                   sampled_host: *const u32, point_host: *const u32, random_coeff_host: *const u32, out: *mut u32,
                                   out_stride: size_t) -> c_int;
    pub fn cb_accumulate_quotients_batches(ctx: *mut cb_ctx, col_ptrs_host: *const *const u32, col_logs_host: *const c_int, n_cols: c_int,
                                           domain_log: c_int, n_batches: c_int, batch_points_host: *const u32,
                                           batch_offsets_host: *const c_int, entry_col_host: *const c_int, entry_value_host: *const u32,
                                           entry_alpha_host: *const u32, out: *mut u32, out_stride: size_t) -> c_int;
    pub fn cb_fold_circle_into_line(ctx: *mut cb_ctx, src: *const u32, src_stride: size_t, src_log: c_int, alpha_host: *const u32,
                                    dst: *mut u32, dst_stride: size_t, dst_is_zero: c_int) -> c_int;
    pub fn cb_fold_line(ctx: *mut cb_ctx, src: *const u32, src_stride: size_t, src_log: c_int, alpha_host: *const u32, dst: *mut u32,
                        dst_stride: size_t) -> c_int;
    pub fn cb_grind_blake2s(ctx: *mut cb_ctx, prefixed_digest_host: *const u8, pow_bits: u32, nonce_out: *mut u64) -> c_int;
    pub fn cb_gather_rows(ctx: *mut cb_ctx, cols: *const u32, stride: size_t, n_cols: c_int, rows_host: *const u32, n_rows: c_int,
                          out_host: *mut u32) -> c_int;
    pub fn cb_gen_trace_chacha_stream(ctx: *mut cb_ctx, key: *const u8, nonce: *const u8, counter: u32, pt_host: *const u8,
                                      ct_host: *const u8, n_blocks: u32, log_size: c_int, words_out: *mut u32, stride: size_t,
                                      valid: *mut c_int) -> c_int;
    // ---- one trace over several GPUs
    pub fn cb_comm_unique_id(id_out: *mut u8) -> c_int;
    pub fn cb_comm_init(ctx: *mut cb_ctx, rank: c_int, world: c_int, id: *const u8) -> c_int;
    pub fn cb_comm_destroy(ctx: *mut cb_ctx) -> c_int;
    // ---- product level (wasm_api.rs exports)
    pub fn s2c_generate_chacha20_proof(ctx: *mut cb_ctx, key: *const u8, key_len: size_t, nonce: *const u8, nonce_len: size_t, counter: u32,
                                       pt: *const u8, pt_len: size_t, ct: *const u8, ct_len: size_t, json_out: *mut *mut c_char,
                                       json_len: *mut size_t) -> c_int;
    pub fn s2c_generate_aes128_ctr_proof(ctx: *mut cb_ctx, key: *const u8, key_len: size_t, nonce: *const u8, nonce_len: size_t, counter: u32,
                                         pt: *const u8, pt_len: size_t, ct: *const u8, ct_len: size_t, json_out: *mut *mut c_char,
                                         json_len: *mut size_t) -> c_int;
    pub fn s2c_generate_aes256_ctr_proof(ctx: *mut cb_ctx, key: *const u8, key_len: size_t, nonce: *const u8, nonce_len: size_t, counter: u32,
                                         pt: *const u8, pt_len: size_t, ct: *const u8, ct_len: size_t, json_out: *mut *mut c_char,
                                         json_len: *mut size_t) -> c_int;
    pub fn s2c_verify_chacha20_proof(proof_b64: *const c_char, proof_b64_len: size_t, nonce: *const u8, nonce_len: size_t, counter: u32,
                                     pt: *const u8, pt_len: size_t, ct: *const u8, ct_len: size_t, json_out: *mut *mut c_char,
                                     json_len: *mut size_t) -> c_int;
    pub fn s2c_verify_aes_ctr_proof(proof_b64: *const c_char, proof_b64_len: size_t, nonce: *const u8, nonce_len: size_t, counter: u32,
                                    pt: *const u8, pt_len: size_t, ct: *const u8, ct_len: size_t, json_out: *mut *mut c_char,
                                    json_len: *mut size_t) -> c_int;
    pub fn s2c_prove_chacha20_raw(ctx: *mut cb_ctx, key: *const u8, nonce: *const u8, counter: u32, plaintext: *const u8,
                                  ciphertext: *const u8, len: size_t, proof_out: *mut *mut u8, proof_len: *mut size_t) -> c_int;
    pub fn s2c_prove_aes_ctr_raw(ctx: *mut cb_ctx, key_len: c_int, key: *const u8, nonce: *const u8, counter: u32, plaintext: *const u8,
                                 ciphertext: *const u8, len: size_t, proof_out: *mut *mut u8, proof_len: *mut size_t) -> c_int;
    pub fn s2c_prove_chacha20_stream_testdata(ctx: *mut cb_ctx, log_size: c_int, proof_out: *mut *mut u8, proof_len: *mut size_t) -> c_int;
    // block AIR of the reference's tests (chacha/bitwise/air.rs prove_bitwise / verify_bitwise): u32 log_size || bincode(StarkProof)
    pub fn s2c_prove_chacha20_block(ctx: *mut cb_ctx, log_size: c_int, proof_out: *mut *mut u8, proof_len: *mut size_t) -> c_int;
    pub fn s2c_verify_chacha20_block(proof: *const u8, proof_len: size_t, err_out: *mut *mut c_char, err_len: *mut size_t) -> c_int;
    // AES-128 block AIR (aes/lookup/air.rs prove_aes_lookup / verify_aes_lookup): u32 log_size || stmt1 || bincode(StarkProof)
    pub fn s2c_prove_aes128_block(ctx: *mut cb_ctx, log_size: c_int, proof_out: *mut *mut u8, proof_len: *mut size_t) -> c_int;
    pub fn s2c_verify_aes128_block(proof: *const u8, proof_len: size_t, err_out: *mut *mut c_char, err_len: *mut size_t) -> c_int;
    pub fn s2c_verify_chacha20_raw(proof: *const u8, proof_len: size_t, nonce: *const u8, counter: u32, pt: *const u8, pt_len: size_t,
                                   ct: *const u8, ct_len: size_t, error_out: *mut *mut c_char) -> c_int;
    pub fn s2c_verify_aes_ctr_raw(proof: *const u8, proof_len: size_t, nonce: *const u8, counter: u32, pt: *const u8, pt_len: size_t,
                                  ct: *const u8, ct_len: size_t, error_out: *mut *mut c_char) -> c_int;
    pub fn s2c_free(p: *mut c_void);
}
