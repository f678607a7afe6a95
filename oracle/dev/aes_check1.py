"""Dev check: AES-CTR oracle milestones (tree roots, claimed sums) against the reference binary's proof."""
import sys, os, base64, struct, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import aes_air as aa
import ref_wasm
from stwo_core import Blake2sChannel, QM31, blake2s
from prover import PcsConfig, CommitmentSchemeProver

def build_inputs(key, nonce, counter, pt, ct, log_size):
    nb = len(pt) // 16
    n = 1 << log_size
    rows_needed = (nb + 15) // 16
    nonce_rows = np.zeros((n, 12), dtype=np.uint8)
    counters = np.zeros(n, dtype=np.uint64)
    P_ = np.zeros((n, 16), dtype=np.uint8)
    C_ = np.zeros((n, 16), dtype=np.uint8)
    for r in range(n):
        if r < rows_needed * 16:
            nonce_rows[r] = list(nonce)
            counters[r] = (counter + r) & 0xFFFFFFFF
            if r < nb:
                P_[r] = list(pt[16 * r:16 * r + 16]); C_[r] = list(ct[16 * r:16 * r + 16])
            else:
                C_[r] = list(aa.ctr_keystream_block(key, nonce, counter + r))
        else:
            counters[r] = r % 16
            C_[r] = list(aa.ctr_keystream_block(key, bytes(12), r % 16))
    return nonce_rows, counters, P_, C_

key = bytes(range(16)); nonce = bytes(range(100, 112)); counter = 7
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 5
pt = bytes((i * 7 + 3) & 0xff for i in range(16 * nb)); ct = aa.ctr_encrypt(key, nonce, counter, pt)
ref = ref_wasm.generate_aes128_ctr_proof(key, nonce, counter, pt, ct)
rb = base64.b64decode(ref["proof"])
log_size = max(8, (nb - 1).bit_length())
roots = [rb[8 + 12 + 4 + 64 + 48 + 25 + 8 + 32 * i:][:32] for i in range(4)]
sums = struct.unpack_from("<8I", rb, 8 + 12 + 4 + 64)
print("log", log_size, "ref sums", sums)
t0 = time.time()
nonce_rows, counters, P_, C_ = build_inputs(key, nonce, counter, pt, ct, log_size)
trace, lookups, mults, valid = aa.generate_ctr_trace(log_size, key, nonce_rows, counters, P_, C_)
print("trace", trace.shape, lookups.shape, valid, "%.1fs" % (time.time() - t0))
cfg = PcsConfig()
ch = Blake2sChannel()
scheme = CommitmentSchemeProver(cfg)
t = aa.sbox_table_columns()
scheme.commit_evals([t[0], t[1]], ch)
print("root0", scheme.trees[0].tree.root() == roots[0])
ch.mix_u64(log_size); ch.mix_u64(0)
pub = bytes(nonce) + struct.pack("<I", counter) + blake2s(pt) + blake2s(ct)
for i in range(3): ch.mix_u64(struct.unpack_from("<I", pub, 4 * i)[0])
ch.mix_u64(counter)
for i in range(16): ch.mix_u64(struct.unpack_from("<I", pub, 16 + 4 * i)[0])
cols = [trace[j] for j in range(trace.shape[0])] + [mults.astype(np.uint64)]
scheme.commit_evals(cols, ch)
print("root1", scheme.trees[1].tree.root() == roots[1], "%.1fs" % (time.time() - t0))
el = aa.SboxElements.draw(ch)
icols, csum = aa.ctr_interaction_trace(log_size, lookups, el)
tcols, tsum = aa.table_interaction_trace(mults, el)
print("ctr sum", csum.v, "table sum", tsum.v, "match", tuple(csum.v + tsum.v) == tuple(sums))
ch.mix_felts([csum, tsum])
scheme.commit_evals(icols + tcols, ch)
print("root2", scheme.trees[2].tree.root() == roots[2], "%.1fs" % (time.time() - t0))
