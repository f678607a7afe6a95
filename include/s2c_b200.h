/* s2c_b200.h -- C ABI of the B200 (sm_100a) stwo prover backend for reclaimprotocol/zk-symmetric-crypto.
 *
 * Two layers, both plain C (pointers + sizes, int status, no C++/torch types):
 *
 *  (1) cb_*  : one entry point per backend-trait method the reference's prover reaches through upstream stwo's
 *              `SimdBackend` (the generic parameter `B: BackendForChannel<MC>` of `CommitmentSchemeProver<B,MC>` and
 *              `prove<B,MC>`, /root/reference/stwo/src/chacha/bitwise/air_stream.rs:143-153,185-231;
 *              /root/reference/stwo/src/aes/lookup/air_ctr.rs:297-414).  A Rust `CudaBackend` implements
 *              `PolyOps`, `MerkleOps`/`MerkleOpsLifted`, `QuotientOps`, `FriOps`, `GrindOps`, `ColumnOps` and
 *              `ComponentProver` as one-line FFI calls onto these (INTEGRATION.md shows the `extern "C"` block).
 *              All `uint32_t*` column arguments are DEVICE pointers (column-major, one column = 2^log_size words in
 *              stwo's bit-reversed circle-domain order, M31 values in [0,p)); `*_host` arguments are host pointers.
 *
 *  (2) s2c_* : the product-level functions the reference exports from its WASM build
 *              (/root/reference/stwo/src/wasm_api.rs:467-648,953-1008; JS binding
 *              /root/reference/js/src/stwo/s2circuits.cjs): bytes in, malloc'd JSON string out, `s2c_free` to release --
 *              the same convention as `__wbindgen_malloc/__wbindgen_free` and as the gnark library
 *              (/root/reference/gnark/libraries/prover/libprove.go:26-48).
 *
 * Status codes: 0 = ok, non-zero = error; `cb_last_error(ctx)` returns the message (valid until the next call on ctx).
 * Threading: one cb_ctx per (GPU, stream); a ctx must not be used from two threads at once; many contexts may coexist.
 * There is no CPU fallback: every entry point fails with an error if no CUDA device is usable.
 */
#ifndef S2C_B200_H
#define S2C_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cb_ctx cb_ctx;

/* ---- lifecycle ------------------------------------------------------------------------------------------------ */
int cb_init(int device, cb_ctx** out);
void cb_destroy(cb_ctx* ctx);
const char* cb_last_error(cb_ctx* ctx);
/* Run all subsequent work of this context on an existing CUDA stream (cudaStream_t as void*); NULL = own stream. */
int cb_set_stream(cb_ctx* ctx, void* cuda_stream);
int cb_sync(cb_ctx* ctx);
/* Number of kernels this context has launched so far. */
uint64_t cb_launch_count(cb_ctx* ctx);

/* ---- ColumnOps: device buffers (Col<B,T>::{zeros,uninitialized,to_cpu}, bit_reverse_column) ------------------------ */
int cb_malloc(cb_ctx* ctx, size_t bytes, void** dptr);
int cb_free(cb_ctx* ctx, void* dptr);
int cb_h2d(cb_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int cb_d2h(cb_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
int cb_memset_zero(cb_ctx* ctx, void* dptr, size_t bytes);

/* ColumnOps::bit_reverse_column (in place, 2^log_size words) and Column::{at, set} on a device column. */
int cb_bit_reverse(cb_ctx* ctx, uint32_t* col, int log_size);
int cb_col_at(cb_ctx* ctx, const uint32_t* col, size_t index, uint32_t* value_out_host);
int cb_col_set(cb_ctx* ctx, uint32_t* col, size_t index, uint32_t value);
/* FieldOps::batch_inverse: dst[i] = src[i]^-1.  QM31 columns are 4 coordinate columns `stride` words apart.  Reached by the
 * reference through LogupTraceGenerator (aes/lookup/gen_ctr.rs:648-682). */
int cb_batch_inverse_m31(cb_ctx* ctx, const uint32_t* src, uint32_t* dst, size_t n);
int cb_batch_inverse_qm31(cb_ctx* ctx, const uint32_t* src, size_t src_stride, uint32_t* dst, size_t dst_stride, size_t n);

/* ---- PolyOps ------------------------------------------------------------------------------------------------------ */
/* PolyOps::precompute_twiddles for CanonicCoset(max_log).circle_domain().half_coset (and every smaller canonic domain);
 * cached in the context (air_stream.rs:185-189). */
int cb_precompute_twiddles(cb_ctx* ctx, int max_log);
/* PolyOps::interpolate_columns: n_cols columns of 2^log_size evaluations (stride words apart) -> coefficients, in place. */
int cb_interpolate_columns(cb_ctx* ctx, uint32_t* cols, size_t stride, int n_cols, int log_size);
/* PolyOps::evaluate_polynomials / extend: coefficients (2^log_size) -> evaluations on CanonicCoset(log_size+log_ext). */
int cb_evaluate_polynomials(cb_ctx* ctx, const uint32_t* coeffs, size_t stride, int n_cols, int log_size, int log_ext,
                            uint32_t* evals, size_t eval_stride);
/* Fused TreeBuilder::extend_evals + evaluate for a commitment: evaluations -> coefficients (coeffs_out, may be NULL) and
 * LDE on CanonicCoset(log_size+log_ext).  src_kind: 0 = M31 words, 1 = packed bits (column j = bit j&31 of word j>>5),
 * 2 = packed bytes (column j = byte j&3 of word j>>2); for packed kinds `src_stride` is the word-row stride. */
int cb_commit_lde(cb_ctx* ctx, int src_kind, const uint32_t* src, size_t src_stride, uint32_t first_col, int n_cols, int log_size,
                  int log_ext, uint32_t* coeffs_out, size_t coeff_stride, uint32_t* lde_out, size_t lde_stride);
/* Packed-witness form of the same fused interpolate+extend (blow-up 2): `src_words` = n_words rows of 2^log_size packed
 * words (src_kind 1: 32 one-bit columns per word, 2: 4 byte columns per word); tiles_out = n_words tiles of
 * [32 or 4][2^(log_size+1)] LDE values.  This is the transform the streaming provers use (coefficients stay on chip). */
int cb_lde_packed(cb_ctx* ctx, int src_kind, const uint32_t* src_words, int n_words, int log_size, uint32_t* tiles_out);
/* PolyOps::extend: coefficient vectors zero-extended to 2^(log_size+log_ext) words. */
int cb_extend(cb_ctx* ctx, const uint32_t* coeffs, size_t stride, int n_cols, int log_size, int log_ext, uint32_t* out, size_t out_stride);
/* PolyOps::barycentric_weights / barycentric_eval_at_point: weights = 4 coordinate columns of 2^log_size words (device) such
 * that f(point) = sum_i evals[i] * weights[i] for every f of the canonic domain of that size (evals in storage order);
 * out_host = n_cols x 4 words. */
int cb_barycentric_weights(cb_ctx* ctx, int log_size, const uint32_t point_host[8], uint32_t* weights_out);
int cb_barycentric_eval_at_point(cb_ctx* ctx, const uint32_t* evals, size_t stride, int n_cols, int log_size, const uint32_t* weights,
                                 uint32_t* out_host);
/* PolyOps::precompute_twiddles for an arbitrary coset (initial index, log size) of the circle group: the flattened twiddle
 * tree upstream's slow_precompute_twiddles builds (2^coset_log_size words each: per layer the x coordinates of the coset's
 * first half in bit-reversed order, then a trailing 1) and its element-wise inverses.  The provers use the cached canonic
 * towers of cb_precompute_twiddles. */
int cb_precompute_twiddles_coset(cb_ctx* ctx, uint32_t coset_initial_index, int coset_log_size, uint32_t* twiddles_out,
                                 uint32_t* itwiddles_out);
/* ---- one large trace on several GPUs (BASELINE cfg-5) --------------------------------------------------
 * The ranks of a communicator prove ONE ChaCha20 trace together: every rank calls s2c_prove_chacha20_raw/_dev with the same
 * inputs; each rank transforms its share of the witness words (whole columns), the LDE tiles are exchanged with an NCCL
 * grouped send/recv all-to-all so that every rank hashes / evaluates constraints on its row range of all columns, and
 * rank 0 finishes the proof (ranks > 0 return an empty proof).  The proof bytes are those of the single-GPU prover.
 * cb_comm_unique_id: 128-byte NCCL id created on one rank and distributed by the caller (e.g. torch.distributed broadcast).
 * Needs libnccl.so.2 at run time (resolved with dlopen; never touched otherwise) and log_size >= 16. */
int cb_comm_unique_id(uint8_t id_out[128]);
int cb_comm_init(cb_ctx* ctx, int rank, int world, const uint8_t id[128]);
int cb_comm_destroy(cb_ctx* ctx);
/* Test hook (process-wide): route 13 <= log_size <= 20 through the generic runtime-schedule kernels that serve log_size > 20. */
int cb_debug_force_generic_fft(int on);
/* Streaming provers keep as many LDE tiles as device memory allows between the commitment pass and the constraint pass;
 * this caps that number (0 = recompute every tile, -1 = default).  Results do not depend on it. */
int cb_set_max_cached_tiles(cb_ctx* ctx, int n_tiles);
/* PolyOps::eval_at_point for n_cols polynomials at one point of the QM31 circle; point_host = {x[4], y[4]},
 * out_host = n_cols x 4 words. */
int cb_eval_at_point(cb_ctx* ctx, const uint32_t* coeffs, size_t stride, int n_cols, int log_size, const uint32_t point_host[8],
                     uint32_t* out_host);

/* ---- MerkleOps<Blake2sMerkleHasher> (lifted VCS) --------------------------------------------------------------------- */
/* build_leaves: Blake2s over the (lifted) values of all columns of a row.  Columns come as `n_groups` groups of equally
 * sized, equally strided columns.  hashes_out: 2^lifting_log x 8 words. */
int cb_merkle_build_leaves(cb_ctx* ctx, const uint32_t* const* group_base, const size_t* group_stride, const int* group_ncols,
                           const int* group_log_size, int n_groups, int lifting_log, uint32_t* hashes_out);
/* Streaming form for column tiles: state = 8 x 2^lifting_log words, `bytes_before` = bytes hashed by earlier calls;
 * every non-final call must carry a multiple of 16 columns. */
int cb_merkle_leaves_absorb(cb_ctx* ctx, const uint32_t* cols, size_t stride, int n_cols, int log_size, int lifting_log,
                            uint32_t* state, uint64_t bytes_before, int is_first, int is_final, uint32_t* hashes_out);
/* build_next_layer / legacy commit_on_layer(prev, no columns): n_parents node hashes from 2*n_parents children. */
int cb_merkle_next_layer(cb_ctx* ctx, const uint32_t* prev_hashes, uint32_t n_parents, uint32_t* out_hashes);

/* Legacy (non-lifted) MerkleOps::commit_on_layer: out[i] = Blake2s(prev[2i] || prev[2i+1] || col_0[i] || ... || col_k[i]) for the
 * 2^log_size nodes of a layer (prev_or_null = hashes of the layer below, 2^(log_size+1) x 8 words; cols_host = host array of
 * n_cols device column pointers).  The pinned reference commits through the lifted VCS only; kept for the BASELINE wording. */
int cb_commit_on_layer(cb_ctx* ctx, int log_size, const uint32_t* prev_or_null, const uint32_t* const* cols_host, int n_cols,
                       uint32_t* out);

/* ---- ComponentProver::evaluate_constraint_quotients_on_domain + AccumulationOps ------------------------------------- */
/* generate_secure_powers, reversed: out[k] = alpha^(n-1-k) (4 words each). */
int cb_generate_secure_powers_rev(cb_ctx* ctx, const uint32_t alpha_host[4], int n, uint32_t* out_dev);
/* ChaCha20 stream AIR (constraints_stream.rs:20-70) on the 2^eval_log evaluation domain; lde = 33,280 columns;
 * alpha_pows_rev from cb_generate_secure_powers_rev(n = 54,784); accum = 4 coordinate columns (accum_stride apart). */
int cb_eval_constraints_chacha_stream(cb_ctx* ctx, const uint32_t* lde, size_t stride, int eval_log, int trace_log,
                                      const uint32_t* alpha_pows_rev, uint32_t* accum, size_t accum_stride, int accumulate);

/* AccumulationOps::accumulate (dst += src over n_words words) and lift_and_accumulate (the 4 coordinate columns `small_cols` of a
 * 2^small_log accumulation, 2^small_log words apart, are lifted onto the 2^big_log domain by the lifted-VCS index map and added). */
int cb_accumulate(cb_ctx* ctx, uint32_t* dst, const uint32_t* src, size_t n_words);
int cb_lift_and_accumulate(cb_ctx* ctx, uint32_t* big, size_t big_stride, int big_log, const uint32_t* small_cols, int small_log);

/* ---- AES-128/256-CTR AIR stages (aes/lookup/{gen_ctr,ctr}.rs, aes/sbox_table.rs) ------------------------------------------ */
/* Column / constraint / lookup counts of the CTR component and the (input, output) trace columns of its S-box lookups in
 * relation order (any out pointer may be NULL; the lookup arrays need n_lookups ints). */
int cb_aes_ctr_layout(int key_len, int* n_cols, int* n_constraints, int* n_lookups, int* lookup_in_cols, int* lookup_out_cols);
/* generate_aes{128,256}_ctr_trace_with_inputs (gen_ctr.rs:386-439, 491-544): trace_out = n_cols columns of 2^log_size words
 * (`stride` apart, device), mults_out_host = the 256 S-box multiplicities, *valid = keystream xor plaintext == ciphertext. */
int cb_gen_trace_aes_ctr(cb_ctx* ctx, int key_len, const uint8_t* key, const uint8_t nonce[12], uint32_t counter, const uint8_t* pt_host,
                         const uint8_t* ct_host, uint32_t n_blocks, int log_size, uint32_t* trace_out, size_t stride,
                         uint32_t mults_out_host[256], int* valid);
/* generate_ctr_sbox_interaction_trace (gen_ctr.rs:640-683): 4 * n_lookups / 2 coordinate columns, the last QM31 column finalised
 * (LogupTraceGenerator::finalize_last); claimed_sum_out_host = the component's claimed sum. */
int cb_gen_logup_interaction_aes_ctr(cb_ctx* ctx, int key_len, const uint32_t* trace, size_t stride, int log_size,
                                     const uint32_t z_host[4], const uint32_t alpha_host[4], uint32_t* inter_out, size_t inter_stride,
                                     uint32_t claimed_sum_out_host[4]);
/* LogupTraceGenerator::finalize_last on one QM31 column (4 coordinate columns `stride` apart, device, in place): inclusive prefix
 * sum in trace-coset order of (value - claimed_sum / N); returns the claimed sum. */
int cb_logup_finalize_last(cb_ctx* ctx, uint32_t* col4, size_t stride, int log_size, uint32_t claimed_sum_out_host[4]);
/* AESCtrEvalAtRow::ctr_block + finalize_logup_in_pairs (ctr.rs:320-364) on the 2^(trace_log+1) evaluation domain, divided by the
 * trace-domain vanishing polynomial: lde = the CTR component's trace columns, inter_lde = its interaction columns,
 * alpha_pows_rev = this component's n_constraints reversed powers of the composition random coefficient (device, 4 words each),
 * accum = 4 coordinate columns (overwritten). */
int cb_eval_constraints_aes_ctr(cb_ctx* ctx, int key_len, const uint32_t* lde, size_t stride, const uint32_t* inter_lde,
                                size_t inter_stride, int trace_log, const uint32_t* alpha_pows_rev, const uint32_t z_host[4],
                                const uint32_t alpha_host[4], const uint32_t claimed_sum_host[4], uint32_t* accum, size_t accum_stride);
/* SboxTableEval::evaluate (sbox_table.rs:103-120) on its own 2^9 evaluation domain (columns of 512 words): preprocessed input /
 * output columns, multiplicity column, the table's interaction columns (4, inter_stride apart), alpha_pow = the power of the
 * composition random coefficient of its single constraint; accum = 4 x 512 words. */
int cb_eval_constraints_sbox_table(cb_ctx* ctx, const uint32_t* pre_in_lde, const uint32_t* pre_out_lde, const uint32_t* mult_lde,
                                   const uint32_t* inter_lde, size_t inter_stride, const uint32_t z_host[4], const uint32_t alpha_host[4],
                                   const uint32_t claimed_sum_host[4], const uint32_t alpha_pow_host[4], uint32_t* accum);

/* ---- QuotientOps / FriOps / GrindOps -------------------------------------------------------------------------------- */
/* accumulate_quotients for one sample point shared by all columns (the ChaCha case): sampled_host = n_cols x 4 words,
 * point_host = {x[4], y[4]}, random_coeff_host[4]; out = 4 coordinate columns on the 2^domain_log domain. */
int cb_accumulate_quotients(cb_ctx* ctx, const uint32_t* cols, size_t stride, int n_cols, int domain_log,
                            const uint32_t* sampled_host, const uint32_t point_host[8], const uint32_t random_coeff_host[4],
                            uint32_t* out, size_t out_stride);
/* accumulate_quotients in its general form: columns of different sizes (col_ptrs_host[j] = device pointer, col_logs_host[j] =
 * log size of the column's evaluation; smaller columns are read through the lifted-VCS index map), sample batches grouped by
 * point (batch b owns entries [batch_offsets[b], batch_offsets[b+1])), and for every entry the column, its sampled value and the
 * power of the quotient random coefficient assigned to that (column, sample) pair.  Batches are summed, as the pinned stwo rev
 * does.  Used by the AES-CTR prover (mask [-1, 0] samples, periodicity samples, lifted S-box table columns). */
int cb_accumulate_quotients_batches(cb_ctx* ctx, const uint32_t* const* col_ptrs_host, const int* col_logs_host, int n_cols,
                                    int domain_log, int n_batches, const uint32_t* batch_points_host, const int* batch_offsets_host,
                                    const int* entry_col_host, const uint32_t* entry_value_host, const uint32_t* entry_alpha_host,
                                    uint32_t* out, size_t out_stride);
int cb_fold_circle_into_line(cb_ctx* ctx, const uint32_t* src, size_t src_stride, int src_log, const uint32_t alpha_host[4],
                             uint32_t* dst, size_t dst_stride, int dst_is_zero);
int cb_fold_line(cb_ctx* ctx, const uint32_t* src, size_t src_stride, int src_log, const uint32_t alpha_host[4], uint32_t* dst,
                 size_t dst_stride);
/* Lowest nonce whose Blake2s(prefixed_digest || nonce) has >= pow_bits trailing zero bits. */
int cb_grind_blake2s(cb_ctx* ctx, const uint8_t prefixed_digest_host[32], uint32_t pow_bits, uint64_t* nonce_out);
int cb_gather_rows(cb_ctx* ctx, const uint32_t* cols, size_t stride, int n_cols, const uint32_t* rows_host, int n_rows,
                   uint32_t* out_host);

/* ---- trace generation ----------------------------------------------------------------------------------------------- */
/* generate_stream_trace (gen_stream.rs:226-261) in packed form: 1,040 words per row (words_out[w*stride + row]).
 * pt/ct: host byte buffers of n_blocks*64 bytes.  *valid_out = 1 iff keystream^pt == ct on every provided row. */
int cb_gen_trace_chacha_stream(cb_ctx* ctx, const uint8_t key[32], const uint8_t nonce[12], uint32_t counter,
                               const uint8_t* plaintext_host, const uint8_t* ciphertext_host, uint32_t n_blocks, int log_size,
                               uint32_t* words_out, size_t stride, int* valid_out);

/* ---- product level (wasm_api.rs exports) -------------------------------------------------------------------------- */
/* Each returns a malloc'd NUL-terminated JSON string in *json_out (release with s2c_free), identical in content to the
 * reference's return string: {"success":true,"blocks":N,"algorithm":"chacha20","proof":"<base64>","proof_size_bytes":S}
 * or {"error":"..."}.  ctx may be NULL (a process-wide context on device 0 is used). */
int s2c_generate_chacha20_proof(cb_ctx* ctx, const uint8_t* key, size_t key_len, const uint8_t* nonce, size_t nonce_len,
                                uint32_t counter, const uint8_t* plaintext, size_t plaintext_len, const uint8_t* ciphertext,
                                size_t ciphertext_len, char** json_out, size_t* json_len);
/* wasm_api.rs:652-772 / 776-896: AES-128 / AES-256 CTR proofs (16-byte blocks, counter big-endian in the counter block);
 * JSON {"success":true,"blocks":N,"algorithm":"aes128-ctr"|"aes256-ctr","proof":"<base64 bincode AESCtrProof>",
 * "proof_size_bytes":S} or {"error":"..."} with the reference's messages. */
int s2c_generate_aes128_ctr_proof(cb_ctx* ctx, const uint8_t* key, size_t key_len, const uint8_t* nonce, size_t nonce_len,
                                  uint32_t counter, const uint8_t* plaintext, size_t plaintext_len, const uint8_t* ciphertext,
                                  size_t ciphertext_len, char** json_out, size_t* json_len);
int s2c_generate_aes256_ctr_proof(cb_ctx* ctx, const uint8_t* key, size_t key_len, const uint8_t* nonce, size_t nonce_len,
                                  uint32_t counter, const uint8_t* plaintext, size_t plaintext_len, const uint8_t* ciphertext,
                                  size_t ciphertext_len, char** json_out, size_t* json_len);
/* Raw form: proof bytes (bincode AESCtrProof); key_len selects AES-128 (16) or AES-256 (32). */
int s2c_prove_aes_ctr_raw(cb_ctx* ctx, int key_len, const uint8_t* key, const uint8_t nonce[12], uint32_t counter,
                          const uint8_t* plaintext, const uint8_t* ciphertext, size_t len, uint8_t** proof_out, size_t* proof_len);
/* Raw form used by the benchmark and tests: proof bytes (bincode StreamProof) instead of base64-in-JSON. */
int s2c_prove_chacha20_raw(cb_ctx* ctx, const uint8_t key[32], const uint8_t nonce[12], uint32_t counter, const uint8_t* plaintext,
                           const uint8_t* ciphertext, size_t len, uint8_t** proof_out, size_t* proof_len);
/* Same, with plaintext/ciphertext already resident in device memory (pt_dev/ct_dev: len bytes each); used to time the prover
 * with inputs in HBM.  pt_hash/ct_hash: the two Blake2s public-input hashes (ChaChaPublicInputs::new, air_stream.rs:44-53) if the
 * caller already has them, or both NULL: the library then reads the buffers back on a side stream and hashes them on host
 * threads while the commitment pass runs (host work in the reference too). */
int s2c_prove_chacha20_dev(cb_ctx* ctx, const uint8_t key[32], const uint8_t nonce[12], uint32_t counter, const void* pt_dev,
                           const void* ct_dev, size_t len, const uint8_t pt_hash[32], const uint8_t ct_hash[32],
                           uint8_t** proof_out, size_t* proof_len);
/* prove_stream::<Blake2sMerkleChannel>(log_size, PcsConfig::default()) of the reference (air_stream.rs:237-289): its own test-data
 * generator (key 00..1f, witness nonce 00 00 00 00 4a .., block r: counter r+1, plaintext word w = 16r + w; the statement binds an
 * all-zero nonce, counter 1 and empty-string hashes).  Returns the bincode StreamProof the reference's generator returns. */
int s2c_prove_chacha20_stream_testdata(cb_ctx* ctx, int log_size, uint8_t** proof_out, size_t* proof_len);
/* Per-stage device times (ms) of the last proof on ctx when profiling is enabled: "name=ms;name=ms;..." */
int cb_set_profile(cb_ctx* ctx, int enable);
const char* cb_stage_times(cb_ctx* ctx);
/* Counters of the last streaming proof on ctx: "fft_words=..;cached_tiles=..;transient_tiles=..;" (packed witness words
 * transformed to LDE tiles over both passes, tiles kept between the passes, transient tile slots). */
/* host wall-clock milliseconds between the stage marks of the last profiled proof ("stage=ms;..."; a stage's entry runs
   from its begin to the next stage's begin, so it includes the host work and stream synchronisations in between) */
const char* cb_host_times(cb_ctx* ctx);
const char* cb_counters(cb_ctx* ctx);
/* ---- verification and prove+verify (wasm_api.rs:609-648 verify_chacha20_proof, :904-946 verify_aes_ctr_proof,
 * :61-188 prove_chacha20_encrypt, :210-330 prove_aes128_ctr_encrypt, :343-463 prove_aes256_ctr_encrypt).
 * Verification is host work (csrc/verify.cu: air_stream.rs:284-421, air_ctr.rs:619-714 + upstream verify) and needs no
 * CUDA context.  JSON as the reference: {"valid":true,"algorithm":A} | {"valid":false,"error":"<VerificationError {:?}>"} |
 * {"error":"Proof payload too large" | "Nonce must be 12 bytes, got N" | "Invalid base64: ..." | "Invalid proof format: ..."}.
 * verify_aes_ctr_proof serves both key sizes (read from the proof's statement). */
int s2c_verify_chacha20_proof(const char* proof_b64, size_t proof_b64_len, const uint8_t* nonce, size_t nonce_len, uint32_t counter,
                              const uint8_t* plaintext, size_t plaintext_len, const uint8_t* ciphertext, size_t ciphertext_len,
                              char** json_out, size_t* json_len);
int s2c_verify_aes_ctr_proof(const char* proof_b64, size_t proof_b64_len, const uint8_t* nonce, size_t nonce_len, uint32_t counter,
                             const uint8_t* plaintext, size_t plaintext_len, const uint8_t* ciphertext, size_t ciphertext_len,
                             char** json_out, size_t* json_len);
/* Raw forms: bincode proof bytes in; 0 = valid, 1 = rejected or malformed, with the reference's error rendering in *error_out
 * (s2c_free; NULL when valid). */
int s2c_verify_chacha20_raw(const uint8_t* proof, size_t proof_len, const uint8_t nonce[12], uint32_t counter, const uint8_t* plaintext,
                            size_t plaintext_len, const uint8_t* ciphertext, size_t ciphertext_len, char** error_out);
int s2c_verify_aes_ctr_raw(const uint8_t* proof, size_t proof_len, const uint8_t nonce[12], uint32_t counter, const uint8_t* plaintext,
                           size_t plaintext_len, const uint8_t* ciphertext, size_t ciphertext_len, char** error_out);
/* Prove on the GPU, verify on the host: {"success":true,"blocks":N,"algorithm":A} | {"error":"..."} (validation and prover
 * errors as the generate_* calls; "Verification failed: <VerificationError {:?}>" if the fresh proof does not verify). */
int s2c_prove_chacha20_encrypt(cb_ctx* ctx, const uint8_t* key, size_t key_len, const uint8_t* nonce, size_t nonce_len,
                               uint32_t counter, const uint8_t* plaintext, size_t plaintext_len, const uint8_t* ciphertext,
                               size_t ciphertext_len, char** json_out, size_t* json_len);
int s2c_prove_aes128_ctr_encrypt(cb_ctx* ctx, const uint8_t* key, size_t key_len, const uint8_t* nonce, size_t nonce_len,
                                 uint32_t counter, const uint8_t* plaintext, size_t plaintext_len, const uint8_t* ciphertext,
                                 size_t ciphertext_len, char** json_out, size_t* json_len);
int s2c_prove_aes256_ctr_encrypt(cb_ctx* ctx, const uint8_t* key, size_t key_len, const uint8_t* nonce, size_t nonce_len,
                                 uint32_t counter, const uint8_t* plaintext, size_t plaintext_len, const uint8_t* ciphertext,
                                 size_t ciphertext_len, char** json_out, size_t* json_len);
/* Blake2s-256 of a host buffer with the library's host implementation (the one the Fiat-Shamir channel, the public-input hashes
   and the verifier use: scalar, or SIMD with run-time dispatch, host_blake2s_simd.cpp).  Host only, no GPU needed. */
int s2c_debug_blake2s(const uint8_t* data, size_t len, uint8_t out[32]);
int s2c_debug_chacha20_keystream(const uint8_t* key, size_t key_len, const uint8_t* nonce, size_t nonce_len, uint32_t counter,
                                 char** json_out, size_t* json_len);
/* Block-AIR variant of the ChaCha20 circuit (reference: stwo/src/chacha/bitwise/air.rs:53-171 prove_bitwise / verify_bitwise,
   constraints.rs, gen.rs; `bench_bitwise` air.rs:201): 32,256 columns, 53,248 constraints, statement = log_size; the trace is the
   reference generator's (key 00..1f, nonce 00 00 00 09 00 00 00 4a 00 00 00 00, counter = row).  Proof bytes = u32 log_size ||
   bincode(StarkProof); free with s2c_free.  s2c_verify_chacha20_block returns 0 when the proof verifies, else 1 with the
   reference's VerificationError rendering in *err_out (free with s2c_free). */
int s2c_prove_chacha20_block(cb_ctx* ctx, int log_size, uint8_t** proof_out, size_t* proof_len);
/* The same for the AES-128 block AIR (reference: stwo/src/aes/lookup/air.rs:139-305 prove_aes_lookup / verify_aes_lookup,
   constraints.rs aes128_block, gen.rs): key 00..0f, input byte b of row r = (r + b) & 0xFF, log_size in [8, 19].  Proof bytes =
   u32 log_size || two QM31 claimed sums || two usize interaction column counts || bincode(StarkProof). */
int s2c_prove_aes128_block(cb_ctx* ctx, int log_size, uint8_t** proof_out, size_t* proof_len);
int s2c_verify_aes128_block(const uint8_t* proof, size_t proof_len, char** err_out, size_t* err_len);
int s2c_verify_chacha20_block(const uint8_t* proof, size_t proof_len, char** err_out, size_t* err_len);
int s2c_get_circuits_info(char** json_out, size_t* json_len);
void s2c_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
