"""CPU restatement of the reference's product-level entry points (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/stwo/src/wasm_api.rs:467-602 (generate_chacha20_proof: validation, log_size choice,
lane packing, padding-lane keystreams) and chacha/bitwise/air_stream.rs:143-234 (prove_stream_with_inputs /
prove_stream_internal: transcript order, statement mixing, tree commits, prove).
"""
import base64
import math
import struct
import numpy as np

import chacha_air as ca
from stwo_core import (U64, P, QM31, Blake2sChannel, blake2s, get_random_point)
from prover import (PcsConfig, CommitmentSchemeProver, secure_powers, coset_vanishing_on_domain,
                    finalize_composition, prove_values)


class ProofError(Exception):
    pass


def chacha_public_inputs(nonce, counter, plaintext, ciphertext):
    """ChaChaPublicInputs::new (air_stream.rs:44-53) serialised as bincode: nonce[12] counter pt_hash ct_hash."""
    return bytes(nonce) + struct.pack("<I", counter) + blake2s(bytes(plaintext)) + blake2s(bytes(ciphertext))


def _mix_chacha_statement(channel, log_size, pub):
    """StreamStatement::mix_into (air_stream.rs:120-123) + ChaChaPublicInputs::mix_into (:66-99)."""
    channel.mix_u64(log_size)
    for i in range(3):
        channel.mix_u64(struct.unpack_from("<I", pub, 4 * i)[0])
    channel.mix_u64(struct.unpack_from("<I", pub, 12)[0])
    for i in range(16):
        channel.mix_u64(struct.unpack_from("<I", pub, 16 + 4 * i)[0])


def prove_stream_internal(log_size, key, nonce, counters, pt, ct, pub, config=None, debug=None, n_input_rows=None):
    """air_stream.rs:160-234.  Inputs are per-row arrays (see chacha_air.generate_stream_trace)."""
    config = config or PcsConfig()
    if log_size < 4:
        raise ProofError("log_size (%d) must be >= LOG_N_LANES (4)" % log_size)
    if log_size > 24:
        raise ProofError("log_size (%d) must be <= MAX_LOG_SIZE (24)" % log_size)
    channel = Blake2sChannel()
    scheme = CommitmentSchemeProver(config)
    scheme.commit_polys([], channel)                                   # empty preprocessed tree
    trace, valid = ca.generate_stream_trace(log_size, key, nonce, counters, pt, ct, n_input_rows)
    if not valid:
        raise ProofError("Ciphertext does not match encryption - invalid witness")
    _mix_chacha_statement(channel, log_size, pub)
    tree1 = scheme.commit_evals(trace, channel)
    # ---- stwo::prover::prove
    random_coeff = channel.draw_secure_felt()
    eval_log = log_size + 1
    apr = secure_powers(random_coeff, ca.N_CONSTRAINTS)[::-1].copy()
    lde = np.stack(tree1.evals, axis=0)                                # eval domain == commitment domain (blowup 2)
    acc = ca.evaluate_constraints(lde, apr)
    from stwo_core import m_inv, q_mul_m31
    den_inv = m_inv(coset_vanishing_on_domain(log_size, eval_log))
    acc = q_mul_m31(acc, den_inv)
    comp_polys = finalize_composition(acc, eval_log)
    scheme.commit_polys(comp_polys, channel)
    oods = get_random_point(channel)
    sample_points = [[], [[oods]] * ca.N_COLS, [[oods]] * 8]
    proof, info = prove_values(scheme, sample_points, channel, eval_log)
    if debug is not None:
        debug.update(info, random_coeff_composition=random_coeff, oods=oods, acc=acc, scheme=scheme)
    # prove()'s closing sanity check (prover/mod.rs): composition value recombined from the sampled halves must
    # equal the constraints evaluated on the sampled mask values, else ProvingError::ConstraintsNotSatisfied.
    if not composition_oods_check(log_size, oods, info["sampled"], random_coeff):
        raise ProofError("Proof generation failed: ConstraintsNotSatisfied")
    stmt = struct.pack("<I", log_size) + pub
    return stmt + proof


def prove_bitwise(log_size, config=None):
    """chacha/bitwise/air.rs:53-137 prove_bitwise: empty preprocessed tree, trace from the fixed generator, statement =
    mix_u64(log_size), one component.  Returns u32 log_size || bincode(StarkProof) (the reference does not serialise
    BitwiseProof; the layout follows StreamProof's)."""
    config = config or PcsConfig()
    channel = Blake2sChannel()
    scheme = CommitmentSchemeProver(config)
    scheme.commit_polys([], channel)
    key_words = [0x03020100 + 0x04040404 * i for i in range(8)]
    trace = ca.generate_block_trace(log_size, key_words, [0x09000000, 0x4a000000, 0])
    channel.mix_u64(log_size)
    tree1 = scheme.commit_evals(trace, channel)
    random_coeff = channel.draw_secure_felt()
    eval_log = log_size + 1
    apr = secure_powers(random_coeff, ca.N_CONSTRAINTS_BLOCK)[::-1].copy()
    acc = ca.evaluate_constraints(np.stack(tree1.evals, axis=0), apr, block=True)
    from stwo_core import m_inv, q_mul_m31
    acc = q_mul_m31(acc, m_inv(coset_vanishing_on_domain(log_size, eval_log)))
    scheme.commit_polys(finalize_composition(acc, eval_log), channel)
    oods = get_random_point(channel)
    proof, info = prove_values(scheme, [[], [[oods]] * ca.N_COLS_BLOCK, [[oods]] * 8], channel, eval_log)
    if not composition_oods_check(log_size, oods, info["sampled"], random_coeff, block=True):
        raise ProofError("Proof generation failed: ConstraintsNotSatisfied")
    return struct.pack("<I", log_size) + proof


def composition_oods_check(log_size, oods, sampled, random_coeff, block=False):
    """Verifier-side identity (core/air/components.rs eval_composition_polynomial_at_point +
    core/verifier.rs): left(z) + pi^{n-1}(z.x) * right(z) == sum_k alpha^(K-1-k) C_k(mask(z)) / Z_H(z)."""
    from stwo_core import Coset, index_to_point
    from prover import secure_powers
    px, py = oods
    mask = np.array([[cv[0].v] for cv in sampled[1]], dtype=U64)          # [C,1,4]
    apr = secure_powers(random_coeff, ca.N_CONSTRAINTS_BLOCK if block else ca.N_CONSTRAINTS)[::-1].copy()
    num = ca.evaluate_constraints(mask, apr, block)[0]
    num = QM31(*[int(v) for v in num])
    # coset_vanishing(CanonicCoset(n).coset, z)
    coset = Coset.odds(log_size)
    t = index_to_point((-coset.initial_index + (coset.step_size >> 1)) & ((1 << 31) - 1))
    x = px * t[0] - py * t[1]
    for _ in range(1, log_size):
        x = x * x * 2 - 1
    lhs_expected = num * x.inv()
    comp = sampled[2]
    left = QM31(0); right = QM31(0)
    units = [QM31(1), QM31(0, 1), QM31(0, 0, 1), QM31(0, 0, 0, 1)]
    for c in range(4):
        left = left + comp[c][0] * units[c]
        right = right + comp[4 + c][0] * units[c]
    # pi^{n-1}(z.x): x doubled (log_size - 1) times
    pix = px
    for _ in range(log_size - 1):
        pix = pix * pix * 2 - 1
    return left + pix * right == lhs_expected


def build_chacha_inputs(key, nonce, counter, plaintext, ciphertext):
    """wasm_api.rs:495-575: log_size, per-row key/nonce/counter/pt/ct with padding lanes and zero default rows."""
    num_blocks = len(plaintext) // 64
    log_size = max(int(math.ceil(math.log2(num_blocks))), 4)
    n = 1 << log_size
    rows_needed = (num_blocks + 15) // 16
    kw = struct.unpack("<8I", key)
    nw = struct.unpack("<3I", nonce)
    K = np.zeros((n, 8), dtype=U64); NO = np.zeros((n, 3), dtype=U64); C = np.zeros(n, dtype=U64)
    PT = np.zeros((n, 16), dtype=U64); CT = np.zeros((n, 16), dtype=U64)
    m = rows_needed * 16
    K[:m] = kw; NO[:m] = nw
    C[:m] = (counter + np.arange(m, dtype=np.uint64)) & np.uint64(0xFFFFFFFF)
    PT[:num_blocks] = np.frombuffer(bytes(plaintext), dtype="<u4").reshape(num_blocks, 16)
    CT[:num_blocks] = np.frombuffer(bytes(ciphertext), dtype="<u4").reshape(num_blocks, 16)
    for row in range(num_blocks, m):
        CT[row] = ca.chacha20_block_words(kw, (counter + row) & 0xFFFFFFFF, nw)
    return log_size, K, NO, C, PT, CT, m


def generate_chacha20_proof(key, nonce, counter, plaintext, ciphertext, debug=None):
    """wasm_api.rs:467-602; returns the same JSON-shaped dict."""
    if len(key) != 32:
        return {"error": "Key must be 32 bytes, got %d" % len(key)}
    if len(nonce) != 12:
        return {"error": "Nonce must be 12 bytes, got %d" % len(nonce)}
    if len(plaintext) == 0 or len(plaintext) % 64 != 0:
        return {"error": "Plaintext must be non-empty multiple of 64 bytes, got %d" % len(plaintext)}
    if len(ciphertext) != len(plaintext):
        return {"error": "Ciphertext must be same length as plaintext, got %d vs %d" % (len(ciphertext), len(plaintext))}
    num_blocks = len(plaintext) // 64
    if num_blocks > 1 and counter + num_blocks - 1 > 0xFFFFFFFF:
        return {"error": "Counter overflow: counter %d + %d blocks would exceed u32::MAX" % (counter, num_blocks)}
    log_size, K, NO, C, PT, CT, m = build_chacha_inputs(key, nonce, counter, plaintext, ciphertext)
    pub = chacha_public_inputs(nonce, counter, plaintext, ciphertext)
    try:
        proof = prove_stream_internal(log_size, K, NO, C, PT, CT, pub, debug=debug, n_input_rows=m)
    except ProofError as e:
        return {"error": str(e)}
    return {"success": True, "blocks": num_blocks, "algorithm": "chacha20",
            "proof": base64.b64encode(proof).decode(), "proof_bytes": proof}
