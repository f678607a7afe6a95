"""Dev check: whole AES-CTR proof bytes, oracle vs reference binary; prints the first differing region."""
import sys, os, base64, struct, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import aes_air as aa, aes_api, ref_wasm
klen = int(sys.argv[2]) if len(sys.argv) > 2 else 16
key = bytes(range(klen)); nonce = bytes(range(100, 112)); counter = 7
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 5
pt = bytes((i * 7 + 3) & 0xff for i in range(16 * nb)); ct = aa.ctr_encrypt(key, nonce, counter, pt)
t0 = time.time()
ref = (ref_wasm.generate_aes128_ctr_proof if klen == 16 else ref_wasm.generate_aes256_ctr_proof)(key, nonce, counter, pt, ct)
t1 = time.time()
dbg = {}
got = (aes_api.generate_aes128_ctr_proof if klen == 16 else aes_api.generate_aes256_ctr_proof)(key, nonce, counter, pt, ct, debug=dbg)
t2 = time.time()
rb = base64.b64decode(ref["proof"]); gb = got["proof_bytes"]
print("ref %.1fs oracle %.1fs" % (t1 - t0, t2 - t1), "len", len(rb), len(gb), "equal", rb == gb)
if rb != gb:
    i = next(k for k in range(min(len(rb), len(gb))) if rb[k] != gb[k])
    hdr = 8 + 80 + 48
    print("first diff at", i, "header ends", hdr, "commitments at", hdr + 25 + 8, "..", hdr + 25 + 8 + 128)
    sys.path.insert(0, os.path.dirname(__file__))
    from aes_parse import parse
    A, B = parse(rb, 136), parse(gb, 136)
    print("commitments", [a == b for a, b in zip(A['commitments'], B['commitments'])])
    print("sampled equal", A['sampled'] == B['sampled'], [a == b for a, b in zip(A['sampled'], B['sampled'])])
    for t in range(len(A['sampled'])):
        bad = [j for j, (a, b) in enumerate(zip(A['sampled'][t], B['sampled'][t])) if a != b]
        if bad: print(" tree", t, "first bad cols", bad[:6], "count", len(bad))
    print("first layer commitment", A['first']['commitment'] == B['first']['commitment'], "pow", A['pow'], B['pow'])
    print("inner commitments", [a['commitment'] == b['commitment'] for a, b in zip(A['inner'], B['inner'])], "last", A['last'] == B['last'])
