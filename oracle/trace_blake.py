"""Debug aid (TEST INFRASTRUCTURE ONLY): log every scalar Blake2s compression the reference performs while
proving, grouped into hash invocations, to pin transcript/Merkle byte layouts of the un-vendored stwo rev."""
import ctypes, struct, sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_wasm as r

CB = ctypes.CFUNCTYPE(None, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64)
IV0 = bytes.fromhex("47e6086b")

def trace(fn, *args):
    L = r.lib()
    mem = r._mem()
    sessions = []
    def hook(f, a0, a1, a2, a3):
        st = ctypes.string_at(mem + a0, 40)
        blk = ctypes.string_at(mem + a1, 64)
        h, t = st[:32], struct.unpack("<Q", st[32:40])[0]
        if h == bytes.fromhex("47e6086b85ae67bb72f36e3c3af54fa57f520e518c68059babd9831f19cde05b"):
            sessions.append([])
        sessions[-1].append((t, blk, a2))
    cb = CB(hook)
    L.w2c_set_hook(cb)
    try:
        out = fn(*args)
    finally:
        L.w2c_set_hook(CB(0))
    return out, sessions

if __name__ == "__main__":
    key = bytes(range(32)); nonce = bytes([0,0,0,9,0,0,0,0x4a,0,0,0,0])
    ks = bytes.fromhex(r.debug_chacha20_keystream(key, nonce, 1)['keystream_hex'])
    pt = bytes((i*7) & 0xff for i in range(64)); ct = bytes(a ^ b for a, b in zip(pt, ks))
    out, ses = trace(r.generate_chacha20_proof, key, nonce, 1, pt, ct)
    print("sessions", len(ses))
    for i, s in enumerate(ses):
        nb = len(s)
        msg = b"".join(b for (_, b, _) in s)
        if nb <= 4 or i < 5:
            print(i, "blocks", nb, "t", [x[0] for x in s][:4], "flag", [x[2] for x in s][:4], msg[:160].hex())
        else:
            print(i, "blocks", nb, "...", msg[:48].hex())
        if i > 400: break
