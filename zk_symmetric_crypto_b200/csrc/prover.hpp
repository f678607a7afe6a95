// Prover-level declarations shared by the AIR drivers.
#pragma once
#include <array>
#include "ctx.hpp"

// PcsConfig::default() of the pinned stwo rev, as serialised inside every reference proof (oracle/prover.py PcsConfig):
// pow_bits 10, FriConfig{log_blowup_factor 1, log_last_layer_degree_bound 0, n_queries 3, fold_step 1}, Option::None.
struct PcsConfig {
    uint32_t pow_bits = 10;
    uint32_t log_blowup = 1;
    uint32_t log_last_layer_degree_bound = 0;
    uint64_t n_queries = 3;
    uint32_t fold_step = 1;
    void serialize(std::vector<uint8_t>& out) const {
        host::put_u32(out, pow_bits);
        host::put_u32(out, log_blowup);
        host::put_u32(out, log_last_layer_degree_bound);
        host::put_u64(out, n_queries);
        host::put_u32(out, fold_step);
        out.push_back(0);
    }
};

struct ProveOptions {
    bool empty_public_hashes = false;  // hash empty byte strings into the statement (reference test-data generator)
    const uint8_t* stmt_nonce = nullptr;  // statement nonce when it differs from the witness nonce (the same generator:
                                          // air_stream.rs:282-283 binds an all-zero nonce while the witness uses 00 00 00 00 4a ..)
    // plaintext/ciphertext already resident on the device (skips the H2D copies); the caller then supplies the two
    // Blake2s public-input hashes (ChaChaPublicInputs::new, air_stream.rs:44-53), which are host work in the reference too
    const uint32_t* pt_dev = nullptr;
    const uint32_t* ct_dev = nullptr;
    const uint8_t* pt_hash = nullptr;
    const uint8_t* ct_hash = nullptr;
    int max_cached_tiles = -1;         // cap on LDE tiles kept between the commitment and constraint passes (-1 = memory bound)
    // block AIR (chacha/bitwise/{gen,constraints,air}.rs): the 32,256-column prefix of the stream AIR (no plaintext / ciphertext),
    // statement = log_size only; the trace comes from key / nonce / counter + row alone (plaintext = ciphertext = nullptr)
    bool block_air = false;
};

struct FriProverState {
    std::vector<DBuf<uint32_t>> evals;  // evals[0] = circle layer [4][2^m]; evals[i>=1] = line layers [4][2^(m-i)]
    std::vector<DevMerkle> trees;       // one per committed layer
    std::vector<int> logs;              // log size of each committed layer
    std::vector<m31::QM31> last_poly;   // last-layer coefficients
};

FriProverState fri_commit(cb_ctx* ctx, host::Channel& ch, const PcsConfig& cfg, DBuf<uint32_t>&& quot, int m);
std::vector<uint8_t> fri_decommit(cb_ctx* ctx, FriProverState& fri, const PcsConfig& cfg, const std::vector<uint32_t>& queries);
uint64_t grind(cb_ctx* ctx, const host::Channel& ch, uint32_t pow_bits);
m31::QM31 coset_vanishing_q(int trace_log, const host::CirclePointQ& z);

size_t stark_proof_size_estimate(const uint8_t* p, size_t len, size_t stark_off);

std::string prove_chacha20(cb_ctx* ctx, const uint8_t key[32], const uint8_t nonce[12], uint32_t counter, const uint8_t* plaintext,
                           const uint8_t* ciphertext, size_t len, std::vector<uint8_t>& proof, ProveOptions opt = ProveOptions());

// AES-128/256-CTR (prove_aes.cu): key_len 16 or 32; len = multiple of 16 bytes.  Returns "" or the reference's error string.
// block_air: the AES-128 block AIR (aes/lookup/air.rs prove_aes_lookup): `plaintext` = one 16-byte input block per row
// (len = 16 << log_size), nonce / counter / ciphertext unused; proof = u32 log_size || stmt1 || bincode(StarkProof)
std::string prove_aes_ctr(cb_ctx* ctx, int key_len, const uint8_t* key, const uint8_t nonce[12], uint32_t counter,
                          const uint8_t* plaintext, const uint8_t* ciphertext, size_t len, std::vector<uint8_t>& proof,
                          bool block_air = false);

// AES-CTR AIR column bookkeeping shared by the prover and the host verifier
struct AesLayout {
    int n_cols = 0, n_constraints = 0;
    std::vector<int> lk_in, lk_out;  // S-box lookup (input, output) columns in relation order
};
AesLayout aes_make_layout(int n_rounds, bool block_air = false);
std::vector<uint8_t> aes_expand_key(const uint8_t* key, int key_len);  // aes/mod.rs:213-270
const uint8_t* aes_sbox();                                             // aes/mod.rs:10-30

// ChaCha20 stream AIR on QM31 mask values; alpha_powers_rev[k] = alpha^(K-1-k)
m31::QM31 chacha_constraints_at_mask(const std::vector<m31::QM31>& mask, const std::vector<m31::QM31>& alpha_powers_rev,
                                     bool block_air = false);

// ---- host verifier (verify.cu): air_stream.rs:343-421, air_ctr.rs:619-714 followed by upstream stwo::core::verifier::verify.
// Both return "" when the proof verifies and otherwise the reference's `{:?}` rendering of its VerificationError.
// A malformed byte string throws VerifyFormatError (the reference's "Invalid proof format: ..." branch).
struct VerifyFormatError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
std::string verify_chacha20(const uint8_t* proof, size_t len, const uint8_t nonce[12], uint32_t counter, const uint8_t* plaintext,
                            size_t pt_len, const uint8_t* ciphertext, size_t ct_len);
// verify_bitwise (chacha/bitwise/air.rs:139-171): proof = u32 log_size || bincode(StarkProof)
std::string verify_chacha20_block(const uint8_t* proof, size_t len);
// verify_aes_lookup (aes/lookup/air.rs:262-305)
std::string verify_aes128_block(const uint8_t* proof, size_t len);
// *key_size_out: 0 = AES-128, 1 = AES-256 (read from the proof's statement before any check)
std::string verify_aes_ctr(const uint8_t* proof, size_t len, const uint8_t nonce[12], uint32_t counter, const uint8_t* plaintext,
                           size_t pt_len, const uint8_t* ciphertext, size_t ct_len, int* key_size_out);
