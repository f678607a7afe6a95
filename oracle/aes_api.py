"""CPU restatement of the reference's AES-CTR product entry points and protocol driver (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/stwo/src/wasm_api.rs:652-896 (generate_aes{128,256}_ctr_proof: validation, log_size, lane packing,
padding-lane keystreams) and aes/lookup/air_ctr.rs:297-422 (prove_aes_ctr_with_inputs_internal: preprocessed S-box tree,
statement 0, main tree + multiplicities, lookup elements, interaction traces, statement 1, interaction tree, prove) with
the proof container of air_ctr.rs:44-184.
"""
import base64
import math
import struct

import numpy as np

import aes_air as aa
from stwo_core import (U64, P, QM31, Blake2sChannel, blake2s, get_random_point, circle_ifft, circle_fft, m_inv, q_mul_m31,
                       q_add, index_to_point)
from prover import (PcsConfig, CommitmentSchemeProver, secure_powers, coset_vanishing_on_domain, finalize_composition,
                    prove_values)


class ProofError(Exception):
    pass


def build_aes_inputs(key, nonce, counter, plaintext, ciphertext):
    """wasm_api.rs:679-746 + gen_ctr.rs:408-434 (default rows beyond the provided vec-rows)."""
    nb = len(plaintext) // 16
    log_size = max(int(math.ceil(math.log2(nb))) if nb > 1 else 0, 8)
    n = 1 << log_size
    rows_needed = (nb + 15) // 16
    m = rows_needed * 16
    nonce_rows = np.zeros((n, 12), dtype=np.uint8)
    counters = np.zeros(n, dtype=np.uint64)
    PT = np.zeros((n, 16), dtype=np.uint8)
    CT = np.zeros((n, 16), dtype=np.uint8)
    nonce_rows[:m] = np.frombuffer(bytes(nonce), dtype=np.uint8)
    counters[:m] = (counter + np.arange(m, dtype=np.uint64)) & np.uint64(0xFFFFFFFF)
    PT[:nb] = np.frombuffer(bytes(plaintext), dtype=np.uint8).reshape(nb, 16)
    CT[:nb] = np.frombuffer(bytes(ciphertext), dtype=np.uint8).reshape(nb, 16)
    for r in range(nb, m):
        CT[r] = np.frombuffer(aa.ctr_keystream_block(key, nonce, counter + r), dtype=np.uint8)
    default_ks = [np.frombuffer(aa.ctr_keystream_block(key, bytes(12), lane), dtype=np.uint8) for lane in range(16)]
    for r in range(m, n):
        counters[r] = r % 16
        CT[r] = default_ks[r % 16]
    return log_size, nonce_rows, counters, PT, CT


def _pt_add_q(p, q):
    """circle group law for (x,y) pairs of QM31 / ints."""
    return p[0] * q[0] - p[1] * q[1], p[0] * q[1] + p[1] * q[0]


def prove_aes_ctr_internal(log_size, key, nonce_rows, counters, PT, CT, pub, config=None, debug=None, block=False):
    """block=True: aes/lookup/air.rs:139-260 prove_aes_lookup (the block AIR: same flow, statement 0 = log_size alone, PT = the
    input block of every row)."""
    config = config or PcsConfig()
    if log_size < 8:
        raise ProofError("log_size (%d) must be >= 8 for S-box table" % log_size)
    if log_size > 24:
        raise ProofError("log_size (%d) must be <= MAX_LOG_SIZE (24)" % log_size)
    key_len = len(key)
    trace, lookups, mults, valid = aa.generate_ctr_trace(log_size, key, nonce_rows, counters, PT, CT, block)
    if not valid:
        raise ProofError("Ciphertext does not match encryption - invalid witness")
    channel = Blake2sChannel()
    scheme = CommitmentSchemeProver(config)
    tab = aa.sbox_table_columns()
    scheme.commit_evals([tab[0], tab[1]], channel)                                   # tree 0
    channel.mix_u64(log_size)
    if not block:
        channel.mix_u64(0 if key_len == 16 else 1)
        for i in range(3):
            channel.mix_u64(struct.unpack_from("<I", pub, 4 * i)[0])
        channel.mix_u64(struct.unpack_from("<I", pub, 12)[0])
        for i in range(16):
            channel.mix_u64(struct.unpack_from("<I", pub, 16 + 4 * i)[0])
    C = trace.shape[0]
    scheme.commit_evals([trace[j] for j in range(C)] + [mults.astype(U64)], channel)  # tree 1
    elems = aa.SboxElements.draw(channel)
    icols, csum = aa.ctr_interaction_trace(log_size, lookups, elems)
    tcols, tsum = aa.table_interaction_trace(mults, elems)
    channel.mix_felts([csum, tsum])
    scheme.commit_evals(icols + tcols, channel)                                       # tree 2
    if not (csum + tsum == QM31(0)):
        raise ProofError("LogUp sums don't balance")
    # ---- stwo::prover::prove
    random_coeff = channel.draw_secure_felt()
    K = aa.n_constraints(key_len, block)
    apr = secure_powers(random_coeff, K + 1)[::-1].copy()       # apr[k] = alpha^(K_total-1-k), ctr component first
    n_i = len(icols)
    ev1, ev2, ev0 = scheme.trees[1].evals, scheme.trees[2].evals, scheme.trees[0].evals
    main = np.stack(ev1[:C], axis=0)
    inter = np.stack(ev2[:n_i], axis=0)
    prev = aa.prev_row_index(log_size, log_size + 1)
    acc = aa.evaluate_ctr_constraints(main, inter, inter[n_i - 4:][:, prev], elems, csum, log_size, apr[:K], key_len, block)
    acc = q_mul_m31(acc, m_inv(coset_vanishing_on_domain(log_size, log_size + 1)))
    prev8 = aa.prev_row_index(8, 9)
    tinter = np.stack(ev2[n_i:], axis=0)
    acc_t = aa.evaluate_table_constraint(np.stack(ev0, axis=0), ev1[C], tinter, tinter[:, prev8], elems, tsum, apr[K])
    acc_t = q_mul_m31(acc_t, m_inv(coset_vanishing_on_domain(8, 9)))
    if log_size == 8:
        acc = q_add(acc, acc_t)
    else:
        # DomainEvaluationAccumulator::finalize with AccumulationOps::lift_and_accumulate: the smaller accumulation is
        # lifted onto the larger evaluation domain by the vcs_lifted index map (value at row i of the large domain =
        # value at row ((i >> (s+1)) << 1) | (i & 1) of the small one) and added
        from prover import _lift
        small = np.stack([_lift(acc_t[:, c], log_size + 1) for c in range(4)], axis=1)
        acc = q_add(acc, small)
    comp_polys = finalize_composition(acc, log_size + 1)
    scheme.commit_polys(comp_polys, channel)                                          # tree 3
    oods = get_random_point(channel)
    step_n = index_to_point((-(1 << (31 - log_size))) & ((1 << 31) - 1))
    step_8 = index_to_point((-(1 << (31 - 8))) & ((1 << 31) - 1))
    oods_prev_n = _pt_add_q(oods, (QM31(step_n[0]), QM31(step_n[1])))
    oods_prev_8 = _pt_add_q(oods, (QM31(step_8[0]), QM31(step_8[1])))
    sp = [[[oods]] * 2,
          [[oods]] * (C + 1),
          [[oods]] * (n_i - 4) + [[oods_prev_n, oods]] * 8,      # points on the lifting domain (prove_values doubles them
                                                                 # for the log-8 table columns)
          [[oods]] * 8]
    proof, info = prove_values(scheme, sp, channel, log_size + 1)
    if debug is not None:
        debug.update(info, scheme=scheme, oods=oods, acc=acc, random_coeff=random_coeff)
    stmt0 = struct.pack("<I", log_size) if block else struct.pack("<II", log_size, 0 if key_len == 16 else 1) + pub
    stmt1 = struct.pack("<4I", *csum.v) + struct.pack("<4I", *tsum.v) + struct.pack("<QQ", n_i, len(tcols))
    return stmt0 + stmt1 + proof


def prove_aes_lookup(log_size, config=None):
    """aes/lookup/air.rs:139-260 with its fixed generator: key 00..0f, input byte b of row r = (r + b) & 0xFF.
    Returns u32 log_size || stmt1 || bincode(StarkProof) (AESLookupProof's field order; the reference does not serialise it)."""
    n = 1 << log_size
    rows = np.arange(n, dtype=np.uint64)[:, None]
    blocks = ((rows + np.arange(16, dtype=np.uint64)[None, :]) & np.uint64(0xFF)).astype(np.uint8)
    zeros12 = np.zeros((n, 12), dtype=np.uint8)
    return prove_aes_ctr_internal(log_size, bytes(range(16)), zeros12, np.zeros(n, dtype=np.uint64), blocks,
                                  np.zeros((n, 16), dtype=np.uint8), b"", config, block=True)


def _generate(key_len, name, key, nonce, counter, plaintext, ciphertext, debug=None):
    if len(key) != key_len:
        return {"error": "Key must be %d bytes, got %d" % (key_len, len(key))}
    if len(nonce) != 12:
        return {"error": "Nonce must be 12 bytes, got %d" % len(nonce)}
    if len(plaintext) == 0 or len(plaintext) % 16 != 0:
        return {"error": "Plaintext must be non-empty multiple of 16 bytes, got %d" % len(plaintext)}
    if len(ciphertext) != len(plaintext):
        return {"error": "Ciphertext must be same length as plaintext, got %d vs %d" % (len(ciphertext), len(plaintext))}
    nb = len(plaintext) // 16
    if nb > 1 and counter + nb - 1 > 0xFFFFFFFF:
        return {"error": "Counter overflow: counter %d + %d blocks would exceed u32::MAX" % (counter, nb)}
    log_size, nonce_rows, counters, PT, CT = build_aes_inputs(key, nonce, counter, plaintext, ciphertext)
    pub = bytes(nonce) + struct.pack("<I", counter) + blake2s(bytes(plaintext)) + blake2s(bytes(ciphertext))
    try:
        proof = prove_aes_ctr_internal(log_size, bytes(key), nonce_rows, counters, PT, CT, pub, debug=debug)
    except ProofError as e:
        return {"error": str(e)}
    return {"success": True, "blocks": nb, "algorithm": name, "proof": base64.b64encode(proof).decode(), "proof_bytes": proof}


def generate_aes128_ctr_proof(key, nonce, counter, plaintext, ciphertext, debug=None):
    return _generate(16, "aes128-ctr", key, nonce, counter, plaintext, ciphertext, debug)


def generate_aes256_ctr_proof(key, nonce, counter, plaintext, ciphertext, debug=None):
    return _generate(32, "aes256-ctr", key, nonce, counter, plaintext, ciphertext, debug)
