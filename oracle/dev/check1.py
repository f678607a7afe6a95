import sys, struct, time, base64
sys.path.insert(0, '/root/repo/oracle')
import numpy as np
from stwo_core import *
import chacha_air as ca
import ref_wasm as r

key = bytes(range(32)); nonce = bytes([0,0,0,9,0,0,0,0x4a,0,0,0,0])
pt = bytes((i*7) & 0xff for i in range(64))
ks = ca.chacha20_keystream_bytes(key, nonce, 1, 1)
assert ks.hex() == r.debug_chacha20_keystream(key, nonce, 1)['keystream_hex']
ct = bytes(a ^ b for a, b in zip(pt, ks))
d = open('/tmp/cc1.bin','rb').read()

def build_inputs(key, nonce, counter, pt, ct):
    nb = len(pt)//64
    log = max((nb-1).bit_length(), 4)
    n = 1 << log
    rows_needed = (nb + 15)//16
    kw = struct.unpack('<8I', key); nw = struct.unpack('<3I', nonce)
    K = np.zeros((n,8), dtype=np.uint64); NO = np.zeros((n,3), dtype=np.uint64); C = np.zeros(n, dtype=np.uint64)
    PT = np.zeros((n,16), dtype=np.uint64); CT = np.zeros((n,16), dtype=np.uint64)
    for row in range(rows_needed*16):
        K[row] = kw; NO[row] = nw; C[row] = (counter + row) & 0xffffffff
        if row < nb:
            PT[row] = struct.unpack('<16I', pt[row*64:row*64+64]); CT[row] = struct.unpack('<16I', ct[row*64:row*64+64])
        else:
            CT[row] = ca.chacha20_block_words(kw, counter+row, nw)
    return log, K, NO, C, PT, CT

log, K, NO, C, PT, CT = build_inputs(key, nonce, 1, pt, ct)
t=time.time()
trace, valid = ca.generate_stream_trace(log, K, NO, C, PT, CT)
print('trace', trace.shape, valid, time.time()-t)
coef = circle_ifft(trace)
# sanity: FFT on same domain recovers
assert np.array_equal(circle_fft(coef), trace)
lde = circle_fft(coef, log+1)
print('lde', lde.shape, time.time()-t)
# queried values from proof: located at offset; parse
p = 799701
def u64():
    global p
    v=struct.unpack('<Q',d[p:p+8])[0]; p+=8; return v
nt=u64(); qv=[]
for tr in range(nt):
    nc=u64(); cols=[]
    for c in range(nc):
        k=u64(); cols.append(struct.unpack('<%dI'%k, d[p:p+4*k])); p+=4*k
    qv.append(np.array(cols, dtype=np.uint64))
print([q.shape for q in qv])
q1 = qv[1]
for qi in range(3):
    rows = [rr for rr in range(lde.shape[1]) if np.array_equal(lde[:,rr], q1[:,qi])]
    print('query', qi, 'matches LDE rows', rows)
mt = MerkleTree([lde[j] for j in range(lde.shape[0])])
print('root', mt.root().hex())
print('ref ', d[117+32:117+64].hex())
np.save('/tmp/cc1_lde.npy', lde); np.save('/tmp/cc1_coef.npy', coef)
